"""Integer helper of the reference binding (fss_crypto/_int.py:4-8)."""


def split_uint128(value: int) -> tuple[int, int]:
    return value & ((1 << 64) - 1), (value >> 64) & ((1 << 64) - 1)
