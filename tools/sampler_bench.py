"""Sampling micro-benchmark (GPU box): one 16x16 latent frame (256 positions x 4 channels) of one video through the
full 8+8-layer DSFVT, CUDA-graph replay per position.  LVT_PDL=1 turns programmatic dependent launch on."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200.config.presets import preset
from lvt_b200.modeling import build_model
B = int(os.environ.get("SAMPLER_B", 1))
cfg = preset("DSFVT", ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", "/tmp/lvt_bench_out"])
cfg.freeze()
vt = build_model(cfg)
vt.train(False)
video = torch.randint(0, 512, (B, 4, 16, 16, 16), device="cuda")
for mode, flag in (("graph", True), ("per_pixel_loop", False)):
    vt.sampler_graph = flag
    vt.sample_video(video.clone(), n_prime=15)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vt.sample_video(video.clone(), n_prime=15)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"B={B} {mode:15s} {dt*1e3/256:.3f} ms/position  {B/dt:.2f} latent frames/s  PDL={os.environ.get('LVT_PDL','0')}")
