"""Timeline of the fused attention backward kernel (CTA 0): clock64 stamps per 128 x 128 block, printed as
cycle deltas.  Slots: 0 producer passed mma2_done(n-1) | 1 MMA warp has its operands | 2 ... and the free accumulators |
3 MMA warp sees P, dS | 4 epilogue sees S, dP | 5 P, dS written | 6 bank sums binned | 7 second MMA group done |
8 dQ drained | 9 dV / dK drained, accumulators released."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops

H, da, L = 8, 128, 256
nb = int(os.environ.get("NB", 64))
causal = os.environ.get("CAUSAL", "0") == "1"
M = nb * L
bf = torch.bfloat16
qkv = (torch.randn(M, 3 * H * da, device="cuda") * 0.5).to(bf)
dO = torch.randn(M, H * da, device="cuda").to(bf)
dqkv = torch.empty_like(qkv)
lse = torch.randn(nb * H, L, device="cuda") + 8.0
delta = torch.randn(nb * H, L, device="cuda")
banks = [torch.zeros(H, 1, device="cuda"), torch.zeros(H, 31, device="cuda"), torch.zeros(H, 31, device="cuda")]
gb = [torch.zeros_like(b) for b in banks]
prof = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.attn_bwd(qkv, dO, dqkv, lse, delta, banks, gb, nb, H, (1, 16, 16), causal, 0.088, prof=prof)
torch.cuda.synchronize()
t = prof.cpu().view(64, 16)
t0 = t[0, 0].item()
names = ["prod", "mma_ops", "mma_free", "mma_ps", "epi_s", "epi_ps", "epi_bank", "epi_m2", "epi_dq", "epi_rel"]
print("blk " + " ".join(f"{n:>9s}" for n in names))
nblk = 3 if causal else 4
for n in range(4 * nblk):
    print(f"{n:3d} " + " ".join(f"{(t[n, s].item() - t0):9d}" for s in range(10)))
