"""Block partition of the radius-sorted star array over ranks (findLimits, reference src/cmc/cmc_utils.c:2299-2333)."""
import numpy as np
from . import lib


def find_limits(N, blk, procs):
    """Start[i], End[i] (1-based inclusive) exactly as findLimits / mpiFindIndicesCustom compute them."""
    S = np.zeros(procs, np.int32)
    E = np.zeros(procs, np.int32)
    lib.load().cmcb200_find_limits(int(N), int(blk), int(procs), S.ctypes.data, E.ctypes.data)
    return S, E
