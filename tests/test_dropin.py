"""The drop-in installer against the real reference checkout (build container only; no GPU needed:
only construction and state_dict layout are checked here -- forward parity is the GPU tests' job)."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")


def test_install_rebinds_hot_path_and_keeps_state_dict_layout():
    sys.path.insert(0, REF)
    import damvsnet_b200 as dm
    from damvsnet_b200 import dropin
    import models.cas_mvsnet as cas
    dropin.uninstall()
    ref_model = cas.CascadeMVSNet(refine=False, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], cr_base_chs=[8, 8, 8])
    ref_keys = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    try:
        patched = dropin.install()
        assert "models.cas_mvsnet.DepthNet" in patched and "models.module.homo_warping" in patched
        model = cas.CascadeMVSNet(refine=False, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], cr_base_chs=[8, 8, 8])
        assert isinstance(model.DepthNet, dm.DepthNet)
        assert all(isinstance(m, dm.CostRegNet) for m in model.cost_regularization)
        keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert keys == ref_keys                       # 580 keys, identical names and shapes
        assert len(keys) == 580
        model.load_state_dict(ref_model.state_dict(), strict=True)   # test_uni.py:224 keeps working
        # the variance variant constructs too (share_cr=True is broken in the reference itself: it passes the
        # list of stage channels as in_channels, reference models/cas_mvsnet.py:178)
        v = cas.CascadeMVSNet(agg_mode="variance")
        assert len(v.state_dict()) == 526
    finally:
        dropin.uninstall()
    assert cas.DepthNet is not dm.DepthNet
