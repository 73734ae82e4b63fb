"""B200-native (sm_100a) flash-attention forward: drop-in for the (Q,K,V)->O path of
tyler-utah/exploring_flash_attention.  See DESIGN.md."""
from ._lib import FA_DTYPE_BF16, FA_DTYPE_F16, FA_DTYPE_F32, FlashAttentionError  # noqa: F401

__all__ = ["FlashAttentionError", "FA_DTYPE_F32", "FA_DTYPE_BF16", "FA_DTYPE_F16"]
