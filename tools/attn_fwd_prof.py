"""Timeline of the fused attention forward kernel (CTA 0, epilogue warp 0): clock64 stamps per 128 x 256 tile as deltas.
Slots: 0 tile start | 1 S ready | 2 S in registers | 3 row max exchanged | 4 exp + row sum exchanged | 5 P in shared memory
(handed to the second MMA) | 6 P V done | 7 O staged and stored."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops
from lvt_b200.ops import Operand
H, da, L = 8, 128, 256
nb = int(os.environ.get("NB", 64))
M = nb * L
bf = torch.bfloat16
qkv = (torch.randn(M, 3 * H * da, device="cuda") * 0.5).to(bf)
o = torch.empty(M, H * da, device="cuda", dtype=bf)
lse = torch.empty(nb * H, L, device="cuda")
banks = [torch.zeros(H, 1, device="cuda"), torch.zeros(H, 31, device="cuda"), torch.zeros(H, 31, device="cuda")]
prof = torch.zeros(32 * 8, dtype=torch.int64, device="cuda")
def qkv_op(which, mn):
    return Operand(qkv.data_ptr() + 2 * which * H * da, 3 * H * da, mn_major=mn, cin=da, zdiv=H, s_zlo=da, s_zhi=L * 3 * H * da)
for _ in range(3):
    ops.gemm(L, L, da, qkv_op(0, False), qkv_op(1, False), Operand(o.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=None,
             batch=nb * H, alpha=0.088, mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL if os.environ.get("CAUSAL") == "1" else 0,
             lse=lse, banks=banks, block=(1, 16, 16), heads=H, v=qkv_op(2, True),
             o2=Operand(o.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), o2_n=da, prof=prof)
torch.cuda.synchronize()
t = prof.cpu().view(32, 8)
t0 = t[0, 0].item()
print("tile " + " ".join(f"{n:>8s}" for n in ["start", "S_ready", "S_regs", "max", "expsum", "P_smem", "PV_done", "O_out"]))
for n in range(7):
    print(f"{n:4d} " + " ".join(f"{t[n, s].item() - t0:8d}" for s in range(8)))
