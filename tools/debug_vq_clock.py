"""Timeline of the tensor-core VQ kernel (CTA 0, first tiles): clock() stamps written by lvt_dbg_vq_clock.
LVT_VQ_TC1=1 selects the round-1 kernel (see csrc/vq_tc.cu)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops, _lib
lib = _lib.load()
nfr = 1024
nhwc = os.environ.get("VQ_NHWC", "0") == "1"
g = torch.Generator().manual_seed(0)
z = (torch.randn(nfr, 256, 16, 16, generator=g) * 0.3).cuda()
cb = (torch.randn(4, 512, 64, generator=g) * 0.3).cuda()
ops.vq_argmin(z, cb)
clk = torch.zeros(24 * 4 * 16, dtype=torch.int32, device="cuda")
lib.lvt_dbg_vq_clock.argtypes = [ctypes.c_void_p]
lib.lvt_dbg_vq_clock(ctypes.c_void_p(clk.data_ptr()))
ops.vq_argmin(z, cb)
torch.cuda.synchronize()
lib.lvt_dbg_vq_clock(ctypes.c_void_p(0))
c = clk.cpu().view(24, 4, 16).numpy().astype("int64") & 0xFFFFFFFF
t0 = c[3, 0, 0]
ver = 1 if os.environ.get("LVT_VQ_TC1", "0") == "1" else 2
mma = ["a_full", "t_empty0", "t_empty1"]
if ver == 1:
    ev = {0: "start", 1: "a_full", 2: "x2done", 3: "t_full", 4: "p1", 5: "bar1", 6: "p2", 7: "bar2", 8: "exact", 9: "bar3", 10: "out"}
    whos = lambda it: ((1, "ew0/g0"), (2, "ew12/g3"), (3, "ew5/g1"))
else:
    ev = {0: "start", 1: "a_full", 2: "x2", 3: "t_full0", 4: "scan0", 5: "t_full1", 6: "scan1", 11: "R2", 14: "filtered",
          7: "listed", 15: "R3", 8: "exact", 9: "R4", 10: "out"}
    whos = lambda it: ((1, "set0 pw0"), (3, "set0 pw1")) if it % 2 == 0 else ((2, "set1 pw0"),)
for it in range(3, 15):
    print(f"tile {it}")
    print("  mma        " + " ".join(f"{n}={int(c[it, 0, i] - t0)}" for i, n in enumerate(mma)))
    for who, label in whos(it):
        print(f"  {label:10s} " + " ".join(f"{n}={int(c[it, who, i] - t0)}" for i, n in ev.items()))
