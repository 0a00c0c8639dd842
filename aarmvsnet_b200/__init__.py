"""aarmvsnet_b200: B200-native (sm_100a) AA-RMVSNet per-depth-plane cost-volume pipeline.

Public API mirrors the reference (models/drmvsnet.py): ``EMVSNet(...)(imgs, proj_matrices, depth_values)``; in the training
phase the returned probability volume is differentiable (``loss.backward()`` reaches all 90 parameters), and
``training.TrainStep`` is train.py:288-307 with a fused flat-buffer Adam.
The arithmetic lives in lib/libaarmvs_sm100.so (C ABI: include/aarmvs.h); there is no CPU path.
"""
from . import synth                                     # noqa: F401
from .model import EMVSNet, mvsnet_cls_loss             # noqa: F401
from .ops import homo_warping_depthwise                 # noqa: F401

__version__ = "0.2.0"
__all__ = ["EMVSNet", "mvsnet_cls_loss", "homo_warping_depthwise", "synth"]
