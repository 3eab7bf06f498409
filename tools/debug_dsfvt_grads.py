"""Debug helper (GPU box): per-parameter gradient error of the DSFVT engine vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lvt_oracle as O
from lvt_b200.modeling.autoregressive import VTEngine, VTSpec

layers, batch = int(sys.argv[1]), int(sys.argv[2])
cfg = O.VTConfig(blocks_e=tuple([(1, 16, 16)] * layers), heads_e=tuple([8] * layers),
                 blocks_d=tuple([(1, 16, 16)] * layers), heads_d=tuple([8] * layers))
weights = O.synth_weights(O.dsfvt_param_shapes(cfg), seed=1234)
context, slc, slice_idx, ignore = O.synth_vt_batch(batch, seed=77, cfg=cfg)
spec = VTSpec(blocks_e=((1, 16, 16),) * layers, heads_e=(8,) * layers, blocks_d=((1, 16, 16),) * layers, heads_d=(8,) * layers)
eng = VTEngine(spec)
eng.load_state_dict(weights)
ws = eng.workspace(batch, cfg.slice_shape, tuple(context.shape[2:]), train=True)
eng.set_inputs(ws, context, slc, slice_idx, ignore)
eng.zero_grad()
loss = eng.forward(ws, train=True)
eng.backward(ws)
torch.cuda.synchronize()
sd = {k: v.clone().requires_grad_(True) for k, v in weights.items()}
want = O.vt_supervised_loss(context, slc, slice_idx, ignore, sd, cfg)
want.backward()
print("loss", loss.item(), want.item())
for name, p in sd.items():
    gw, gg = p.grad, eng.store.g[name].cpu()
    if name == "decoder.conv.conv.weight":
        gw = gw.clone(); gw[:, :, -1, -1, 1:] = 0
    e = ((gg - gw).double().norm() / (gw.double().norm() + 1e-30)).item()
    cos = (gg.double().flatten() @ gw.double().flatten() / (gg.double().norm() * gw.double().norm() + 1e-30)).item()
    print(f"{name:60s} rel {e:9.4f} cos {cos:8.5f} |g| {gw.norm().item():.3e} |got| {gg.norm().item():.3e}")
