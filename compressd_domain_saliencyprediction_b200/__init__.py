"""B200-native compressed-domain saliency hot path (sm_100a CUDA behind a C ABI).

Python here is only the host-side mirror of the reference's operator surfaces
(computecontrast5, getFrame, libsvmpredict, mv_vote, the per-clip loop of main.m); all compute
runs in libcds_b200.so.  There is no CPU fallback: calls raise CdsError without a B200.
"""
from ._lib import CdsError, lib, lib_path, device_ok, take_launch_count  # noqa: F401
from . import feeder, formats, mat73, metrics, ops, shard  # noqa: F401
from .saliency import SaliencyClip  # noqa: F401

__all__ = ["CdsError", "lib", "lib_path", "device_ok", "take_launch_count", "formats", "mat73", "metrics", "ops", "shard", "SaliencyClip"]
