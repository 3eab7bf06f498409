"""Timeline of the tensor-core VQ kernel (CTA 0, first tiles): clock() stamps written by lvt_dbg_vq_clock."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops, _lib
lib = _lib.load()
nfr = 1024
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
t0 = c[2, 0, 0]
names = {0: ["a_full", "t_empty0", "t_empty1"],
         1: ["start", "a_full", "x2done", "t_full", "p1", "bar1", "p2", "bar2", "exact", "bar3", "out"]}
for it in range(2, 12):
    print(f"tile {it}")
    for who, label in ((0, "mma"), (1, "ew0/g0"), (2, "ew12/g3"), (3, "ew5/g1 x2")):
        nm = names[0] if who == 0 else names[1]
        print(f"  {label:10s} " + " ".join(f"{n}={int(c[it, who, i] - t0)}" for i, n in enumerate(nm)))
