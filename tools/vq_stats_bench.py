import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops
nfr = 512
g = torch.Generator().manual_seed(0)
z = (torch.randn(nfr * 256, 256, generator=g) * 0.3).cuda()
cb = (torch.randn(4, 512, 64, generator=g) * 0.3).cuda()
counts = torch.zeros(4, 512, device="cuda"); sums = torch.zeros(4, 512, 64, device="cuda")
def run(stats):
    if stats:
        ops.vq_argmin_nhwc(z, cb, 256, want_zq_bf16=True, counts=counts, sums=sums)
    else:
        ops.vq_argmin_nhwc(z, cb, 256, want_zq_bf16=True)
for stats in (False, True):
    for _ in range(3): run(stats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run(stats)
    e1.record(); torch.cuda.synchronize()
    print("stats" if stats else "argmin only", f"{e0.elapsed_time(e1)/10*1e3:.1f} us")
