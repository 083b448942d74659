"""birda_b200 — B200-native audio front end + post-inference scoring for birda's hot path.

The product is the C-ABI library ``libbirda_b200.so`` (``include/birda_b200.h``); this
package is a thin ctypes mirror of it for tests, ``bench.py`` and Python hosts.  There is no
CPU fallback: importing works anywhere (so the host rules can be used), but creating a
``Context`` without a CUDA device raises.
"""
from ._lib import BirdaError, lib, lib_path  # noqa: F401
from .api import (  # noqa: F401
    ACT_NONE, ACT_SIGMOID, ACT_SOFTMAX, FMT_F32, FMT_S16, FMT_S24, FMT_S32,
    Context, FrontEndPlan, MelSpec, PostConfig, Segments, StandIn, Watchdog, FlacDecoder, flac_probe, mask_build, rules, wav_probe, wav_read,
)

__version__ = "0.1.0"
