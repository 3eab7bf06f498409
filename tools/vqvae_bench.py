"""PR-DVQVAE2 train-step micro-benchmark (GPU box): 512 synthetic 64x64 frames, eager launches (profile with ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200 import _lib
from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
nfr = int(os.environ.get("VQVAE_FRAMES", 512))
steps = int(os.environ.get("VQVAE_STEPS", 5))
spec = VQVAESpec(n_layers=2)
ve = VQVAEEngine(spec)
g = torch.Generator().manual_seed(7)
init = {}
for name, shp in spec.param_shapes().items():
    fan = 1
    for s_ in shp[1:]:
        fan *= s_
    init[name] = torch.randn(shp, generator=g) / (fan ** 0.5) if len(shp) > 1 else torch.zeros(shp)
ve.store.load(init)
ve.load_state_dict(codebook=torch.randn(4, 512, 64, generator=g) * 0.3, running_size=torch.full((4, 512), 5.0))
ve.init_optimizer(lr=3e-4, betas=(0.9, 0.9))
w = ve.workspace(nfr, train=True)
w.x.copy_(torch.rand((nfr, 3, 64, 64), generator=g))
for _ in range(2):
    ve.train_step(w)
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ve.train_step(w)
e1.record(); torch.cuda.synchronize()
print(f"vqvae train step: {e0.elapsed_time(e1)/steps:.3f} ms for {nfr} frames, {(_lib.launch_count()-n0)//steps} launches/step, losses {w.loss.tolist()}")
