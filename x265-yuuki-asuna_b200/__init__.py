"""x265-yuuki-asuna_b200 -- B200-native backend for x265's EncoderPrimitives / motion-estimation
hot path.  The product is the C-ABI library libx265b200.so (csrc/*.cu, include/x265b200.h) and the
C++ drop-in adapter (adapter/); this Python package is only the ctypes binding used by tests/ and
bench.py.  Import with importlib.import_module("x265-yuuki-asuna_b200")."""
from .capi import *  # noqa: F401,F403
from .capi import Ctx, DevBuf, X265B200Error, load, header_symbols, LIB_PATH  # noqa: F401
from .sharding import band_rows, plane_chunk, records_per_band, assemble_bands, triples_of_rank, triples_per_rank_max, assemble_triples  # noqa: F401
