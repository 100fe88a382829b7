"""lwsnet_b200 — B200-native (sm_100a) implementation of the LWSNet multi-stage stereo hot path.

Host side: a mirror of the reference's model API (``models.LWSNet`` and the ``submodules`` factories of
PrinceVictor/LWSNet models/models.py, models/submodules.py) that runs every hot-path function through the C ABI of
``include/lws.h`` (``lib/liblws_b200.so``, hand-written CUDA).  There is no CPU path: calling the model without the
CUDA library or with CPU tensors raises.
"""
from . import _lib  # noqa: F401  (fails loudly when liblws_b200.so is missing)
from .models import LWSNet, disparity_regression  # noqa: F401

__all__ = ["LWSNet", "disparity_regression"]
__version__ = "0.1.0"
