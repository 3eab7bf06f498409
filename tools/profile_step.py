"""One eager (un-graphed) DSFVT train step between cudaProfilerStart / Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/<tag>_step_launches.csv python tools/profile_step.py
(WORKLOAD=vqvae: one PR-DVQVAE2 step instead).  Summarise with tools/launch_table.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200.data import synthetic_vt_batch
from lvt_b200.modeling.autoregressive import VTEngine, VTSpec

if os.environ.get("WORKLOAD", "dsfvt") == "vqvae":
    from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
    spec = VQVAESpec(n_layers=2)
    ve = VQVAEEngine(spec)
    g = torch.Generator().manual_seed(7)
    w = {}
    for name, shp in spec.param_shapes().items():
        fan = 1
        for s_ in shp[1:]:
            fan *= s_
        w[name] = torch.randn(shp, generator=g) / (fan ** 0.5) if len(shp) > 1 else torch.zeros(shp)
    ve.store.load(w)
    ve.load_state_dict(codebook=torch.randn(4, 512, 64, generator=g) * 0.3, running_size=torch.full((4, 512), 5.0))
    ve.init_optimizer(lr=3e-4, betas=(0.9, 0.9))
    vw = ve.workspace(512, train=True)
    vw.x.copy_(torch.rand((512, 3, 64, 64), generator=g))
    for _ in range(2):
        ve.train_step(vw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ve.train_step(vw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

B = int(os.environ.get("BATCH", 64))
spec = VTSpec()
eng = VTEngine(spec)
g = torch.Generator().manual_seed(1234)
init = {}
for name, shp in spec.param_shapes().items():
    if "layer_norm.weight" in name or name.endswith("ffn.0.weight"):
        init[name] = torch.ones(shp)
    elif len(shp) == 1:
        init[name] = torch.zeros(shp)
    else:
        fan = 1
        for s_ in shp[1:]:
            fan *= s_
        init[name] = torch.randn(shp, generator=g) / (fan ** 0.5)
eng.load_state_dict(init)
eng.init_optimizer("rmsprop", lr=2e-5, alpha=0.95, momentum=0.9, eps=1e-8)
host = synthetic_vt_batch(B, seed=1000)
ws = eng.workspace(B, (1, 16, 16), tuple(host[0].shape[2:]), train=True)
eng.set_inputs(ws, *host)
for _ in range(2):
    eng.train_step(ws)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(ws)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
