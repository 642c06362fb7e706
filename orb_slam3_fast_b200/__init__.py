"""orb_slam3_fast_b200 — B200-native (sm_100a) ORB-SLAM3 tracking front-end: ORBextractor + Hamming matchers.

Python here is the host-side mirror of the reference's C++ class surface over the C ABI in include/*.h; all compute is
in liborbx.so (hand-written CUDA). There is no CPU fallback.
"""
from .lib import KP_DTYPE, OrbxError, build  # noqa: F401
from .extractor import ORBextractor, cvtColorToGray, remapLinear  # noqa: F401
from .matcher import ORBmatcher  # noqa: F401
from . import views  # noqa: F401
