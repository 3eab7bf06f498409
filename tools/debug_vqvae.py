"""Debug helper (GPU box): VQ-VAE engine vs oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lvt_oracle as O
from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = O.VQVAEConfig(n_layers=L)
eshape, gshape = O.vqvae_param_shapes(cfg)
we, wg = O.synth_weights(eshape, seed=11), O.synth_weights(gshape, seed=12)
x = torch.rand((n, 3, 64, 64), generator=torch.Generator().manual_seed(1234))
with torch.no_grad():
    z_ref = O.res_encoder((x - 0.5) / 0.5, we, cfg.n_layers)
cb = torch.randn((4, 512, 64), generator=torch.Generator().manual_seed(5)) * z_ref.std()

eng = VQVAEEngine(VQVAESpec(n_layers=L))
eng.load_state_dict(we, wg, cb)
w = eng.workspace(n, train=True)
w.x.copy_(x)
recon, idx = eng.inference(w)
torch.cuda.synchronize()
z_got = w.z_e.cpu().view(n, 16, 16, 256).permute(0, 3, 1, 2)
print("z_e rel err", ((z_got - z_ref).norm() / z_ref.norm()).item(), "max", (z_got - z_ref).abs().max().item(), "scale", z_ref.abs().max().item())
with torch.no_grad():
    recon_ref, idx_ref = O.vqvae_inference(x, we, wg, cb, cfg)
agree = (idx.cpu() == idx_ref).float().mean().item()
print("index agreement", agree)
# decoder alone on the oracle's indices
xt = eng.decode_indices(w, idx_ref.cuda().contiguous()).cpu()
with torch.no_grad():
    xt_ref = O.res_decoder(O.dvq_embed(idx_ref, cb), wg, cfg.n_layers)
print("decoder rel err", ((xt - xt_ref).norm() / xt_ref.norm()).item(), "max", (xt - xt_ref).abs().max().item())
print("recon err (own indices)", (recon.cpu() - recon_ref).abs().max().item(), (recon.cpu() - recon_ref).abs().mean().item())

# training step
eng2 = VQVAEEngine(VQVAESpec(n_layers=L))
eng2.load_state_dict(we, wg, cb)
eng2.init_optimizer()
w2 = eng2.workspace(n, train=True)
w2.x.copy_(x)
eng2.store.grad.zero_()
eng2.forward_train(w2)
eng2.backward(w2)
torch.cuda.synchronize()
print("losses", w2.loss.tolist())
we_g = {k: v.clone().requires_grad_(True) for k, v in we.items()}
wg_g = {k: v.clone().requires_grad_(True) for k, v in wg.items()}
losses, aux = O.vqvae_supervised_loss(x, we_g, wg_g, cb, torch.zeros(4, 512), cb.clone(), cfg)
sum(losses.values()).backward()
print("oracle losses", {k: v.item() for k, v in losses.items()})
print("codebook after rel err", ((eng2.codebook.cpu() - aux["codebooks"]).norm() / aux["codebooks"].norm()).item())
for pre, sd in (("E.", we_g), ("G.", wg_g)):
    for k, p in sd.items():
        gw, gg = p.grad, eng2.store.g[pre + k].cpu()
        e = ((gg - gw).double().norm() / (gw.double().norm() + 1e-30)).item()
        cos = (gg.double().flatten() @ gw.double().flatten() / (gg.double().norm() * gw.double().norm() + 1e-30)).item()
        print(f"{pre + k:34s} rel {e:8.4f} cos {cos:8.5f} |g| {gw.norm().item():.3e} |got| {gg.norm().item():.3e}")
