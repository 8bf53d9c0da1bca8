"""Integer helper kept for API compatibility (fss_crypto/_int.py:4-8)."""


def split_uint128(value: int) -> tuple[int, int]:
    """Split an unsigned integer into low and high 64-bit halves."""
    return value & ((1 << 64) - 1), (value >> 64) & ((1 << 64) - 1)
