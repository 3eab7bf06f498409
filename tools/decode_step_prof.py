"""Per-stage timeline (ns, %globaltimer of CTA 0) of the fused per-position sampler kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lvt_b200.config.presets import preset
from lvt_b200.modeling import build_model
from lvt_b200.modeling.autoregressive.incremental import IncrementalDecoder
B = int(os.environ.get("SAMPLER_B", 1))
cfg = preset("DSFVT", ["TEST.EVALUATORS", "VTSampler", "OUTPUT_DIR", "/tmp/lvt_bench_out"])
cfg.freeze()
vt = build_model(cfg)
vt.train(False)
eng = vt.model.engine
ctx = torch.randint(0, 512, (B, 4, 7, 16, 16), device="cuda")
slc = torch.randint(0, 512, (B, 4, 1, 16, 16), device="cuda")
ws = vt.model._stage(ctx, slc, torch.zeros(B, dtype=torch.int64, device="cuda"), None, train=False)
eng.encoder_forward(ws, train=False)
dec = IncrementalDecoder(eng, ws)
dec.begin_slice()
prof = torch.zeros(128, dtype=torch.int64, device="cuda")
d = dec._fused_desc(True, 1.0)
d.prof = prof.data_ptr()
for p in (0, 100, 200):
    dec.pos.fill_(p)
    dec.sample_row_fused(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dec.sample_row_fused(1.0)
e1.record(); torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per position (4 exponential_ + 1 kernel, eager)")
t = prof.cpu().tolist()
t = [x for x in t if x > 0]
d = [t[i + 1] - t[i] for i in range(len(t) - 1)]
print("total kernel ns:", t[-1] - t[0], "stamps", len(t))
print("alternating (stage work, barrier) ns:", d[:40])
work = sum(d[0::2]); bar = sum(d[1::2])
print("sum stage work ns", work, "sum barrier ns", bar)
