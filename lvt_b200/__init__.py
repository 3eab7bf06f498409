"""lvt_b200 — B200-native (sm_100a) implementation of the hot path of rakhimovv/lvt.

Host side: Python mirroring the reference's `vidgen` module/registry surface.
Device side: hand-written CUDA (tcgen05 / TMEM / TMA) behind the C-ABI in include/lvt_b200.h,
built in-tree as lvt_b200/lib/liblvt_b200.so.
"""
__version__ = "0.1.0"
