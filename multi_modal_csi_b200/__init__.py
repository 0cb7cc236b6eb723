"""B200-native THAT (WiFi-CSI two-stream transformer) train step.

Drop-in for the ``benchmark/wifi_csi`` hot path of amirhosseinmhd/multi_modal_CSI: ``THAT``, ``train``,
``run_that``, ``preset`` and the metric/result format keep the reference's names and meaning; the arithmetic
runs in hand-written sm_100a kernels behind the C ABI declared in ``include/csi_that.h``.
"""
from .that import THAT, THAT_COUNT_PRED, THAT_MULTI_HEAD, PermutationMatchingLoss  # noqa: F401
from .cnn2d import CNN_2D  # noqa: F401
from .optim import FusedAdam  # noqa: F401

__all__ = ["THAT", "THAT_COUNT_PRED", "THAT_MULTI_HEAD", "PermutationMatchingLoss", "CNN_2D", "FusedAdam"]
