"""B200-native neural min-sum LDPC decode / Monte-Carlo path (drop-in for the hot path of
ghy1228/LDPC_Error_Floor).  CUDA kernels + C-ABI live in csrc/ (libldpc_b200.so); this package
is the host-side mirror of the reference's operator surface."""
from . import formats  # noqa: F401
from ._lib import (COUNTER_NAMES, FLAG_SYND_OK, FLAG_SYND_OK_EVER, FLAG_UNCOR_ANY, FLAG_UNCOR_LAST,  # noqa: F401
                   HARVEST_NONE, HARVEST_SYND_FAIL, HARVEST_UNCOR_ANY, HARVEST_UNCOR_LAST, LdpcError)
from .formats import WeightSet  # noqa: F401
from .graph import BaseGraph, init_parameter  # noqa: F401


def __getattr__(name):
    # decoder.py imports torch; keep `import ldpc_error_floor_b200` light for format-only users
    if name in ("NMSDecoder", "DecodeResult", "check_params", "unpack_bits"):
        from . import decoder
        return getattr(decoder, name)
    if name in ("MonteCarlo", "compute_results", "create_mix_epoch", "SnrPoint"):
        from . import montecarlo
        return getattr(montecarlo, name)
    raise AttributeError(name)
