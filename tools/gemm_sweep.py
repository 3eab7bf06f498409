"""GEMM timing sweep (GPU box): separates the fixed per-launch cost from the per-tile cost of lvt_gemm_bf16 by
timing M = 1, 2, 4 tiles per CTA for a few (N, K), with and without the epilogue (debug flag bit 29)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops
from lvt_b200.ops import Operand

bf = torch.bfloat16
reps = 30


def timeit(fn):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (N, K) in [(512, 512), (3072, 512), (512, 2048)]:
    for mult in (1, 2, 4, 8):
        M = 148 * 128 * mult // (N // 256) if N <= 512 else 148 * 128 * mult // 4
        M = max(256, M // 256 * 256)
        x = torch.randn(M, K, device="cuda").to(bf)
        w = torch.randn(N, K, device="cuda").to(bf)
        o = torch.empty(M, N, device="cuda", dtype=bf)
        of = torch.empty(M, N, device="cuda")
        row = []
        for flags in (0, 1 << 29):
            t = timeit(lambda: ops.gemm(M, N, K, Operand(x.data_ptr(), K), Operand(w.data_ptr(), K), Operand(o.data_ptr(), N),
                                        out_bf16=o, flags=flags))
            row.append(t)
        t32 = timeit(lambda: ops.gemm(M, N, K, Operand(x.data_ptr(), K), Operand(w.data_ptr(), K), Operand(of.data_ptr(), N),
                                      out_f32=of, res=of))
        fl = 2.0 * M * N * K
        print(f"N={N} K={K} M={M:6d} tiles128={M // 128 * (N // 256):5d}: bf16-out {row[0]:7.1f} us ({fl / row[0] / 1e6:6.0f} TF/s)  "
              f"no-epilogue {row[1]:7.1f} us  f32+res {t32:7.1f} us ({fl / t32 / 1e6:6.0f} TF/s)")
        del x, w, o, of
