"""Drop-in installer: run the reference's own ``CascadeMVSNet`` / ``train.py`` / ``test_uni.py`` on the native hot path.

The reference builds its model from names it looks up in ``models.module`` and ``models.cas_mvsnet`` at
construction time (reference models/cas_mvsnet.py:4 ``from .module import *``; ``CostRegNet`` at :178-182,
``DepthNet`` at :188).  ``install()`` rebinds exactly the hot-path names in those two modules to this
package's classes/functions, which keep the reference's signatures and state_dict keys; everything else
(FeatureNet, GeoFeatureFusion, hypothesis sampling, losses, drivers) keeps running the reference's PyTorch
code, as BASELINE.json's north star prescribes.  Nothing of the reference is copied.

    import sys; sys.path.insert(0, "/path/to/DAMVSNet")
    import damvsnet_b200.dropin as dropin
    dropin.install()
    from models.cas_mvsnet import CascadeMVSNet      # now backed by libdamvs_b200.so
"""
from __future__ import annotations

import importlib
from typing import Dict, List

HOT_NAMES = ("homo_warping", "depth_regression", "Conv3d", "Deconv3d", "CostRegNet", "AggWeightNetVolume",
             "uncertainty_aware_samples", "cross_view_loss")

_saved: Dict[str, Dict[str, object]] = {}


def install(module_pkg: str = "models", precision: str = "fp32") -> List[str]:
    """Rebind the hot-path names of ``<module_pkg>.module`` and ``<module_pkg>.cas_mvsnet``.  Returns the patched
    qualified names.  Idempotent; ``uninstall()`` restores the reference's own definitions.

    `precision` is set explicitly: "fp32" (default) keeps the reference's arithmetic width (relative depth error
    <= 1e-4 against the reference); "bf16" opts into the reduced-precision pipeline (fp16 features, bf16 cost volume
    and tensor-core convolutions; its error bound is stated in DESIGN.md section 5)."""
    import damvsnet_b200 as dm
    dm.set_precision(precision)
    mod = importlib.import_module(module_pkg + ".module")
    cas = importlib.import_module(module_pkg + ".cas_mvsnet")
    patched = []
    for target in (mod, cas):
        saved = _saved.setdefault(target.__name__, {})
        for name in HOT_NAMES:
            if hasattr(target, name):
                saved.setdefault(name, getattr(target, name))
                setattr(target, name, getattr(dm, name))
                patched.append(f"{target.__name__}.{name}")
    saved = _saved.setdefault(cas.__name__, {})
    saved.setdefault("DepthNet", cas.DepthNet)
    cas.DepthNet = dm.DepthNet
    patched.append(f"{cas.__name__}.DepthNet")
    return patched


def uninstall() -> None:
    for modname, names in _saved.items():
        target = importlib.import_module(modname)
        for name, obj in names.items():
            setattr(target, name, obj)
    _saved.clear()
