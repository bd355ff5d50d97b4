"""damvsnet_b200 -- B200-native (sm_100a) cost-volume hot path of DA-MVSNet.

Drop-in names (same signatures as the reference's models/module.py and
models/cas_mvsnet.py): homo_warping, depth_regression, Conv3d, Deconv3d,
CostRegNet, AggWeightNetVolume, DepthNet.  The arithmetic lives in the CUDA
library behind include/damvs.h (built in-tree by ``python -m damvsnet_b200.build``).
"""
from . import ops  # noqa: F401
from .cas_mvsnet import DepthNet  # noqa: F401
from .module import (AggWeightNetVolume, Conv3d, CostRegNet, Deconv3d, depth_regression,  # noqa: F401
                     homo_warping, invalidate_packed, uncertainty_aware_samples)
from .losses import cross_view_loss  # noqa: F401
from .ops import G8Volume, precision, set_precision  # noqa: F401

__version__ = "0.2.0"
