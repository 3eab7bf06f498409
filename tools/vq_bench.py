"""VQ argmin micro-benchmark (GPU box): 2^20 latent positions, 1056 algorithmic bytes each."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import ops
nfr = int(os.environ.get("VQ_FRAMES", 4096))
init = os.environ.get("VQ_INIT", "spread")
g = torch.Generator().manual_seed(0)
z = (torch.randn(nfr, 256, 16, 16, generator=g) * 0.3).cuda()
cb = (torch.randn(4, 512, 64, generator=g) * 0.3 if init == "spread" else (torch.rand(4, 512, 64, generator=g) * 2 - 1) / 512).cuda()
nhwc = os.environ.get("VQ_NHWC", "0") == "1"   # the VQ-VAE engine's channels-last layout
if nhwc:
    zl = z.permute(0, 2, 3, 1).reshape(-1, 256).contiguous()
    run = lambda: ops.vq_argmin_nhwc(zl, cb, 256, want_zq_bf16=True)
    init += " nhwc+zq_bf16"
else:
    run = lambda: ops.vq_argmin(z, cb)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
pos = nfr * 256
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6554.9
print(f"vq_argmin[{init}] {t*1e3:.3f} ms  {pos/t/1e9:.3f} Gpos/s  {pos*1056/t/1e9:.1f} GB/s algorithmic = {pos*1056/t/1e9/peak:.3f} of HBM peak")
