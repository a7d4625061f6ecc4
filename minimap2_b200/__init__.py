"""minimap2_b200 -- B200-native seed-chain-extend mapper behind the minimap.h C API."""
from ._lib import lib, Context, LIB_PATH  # noqa: F401
