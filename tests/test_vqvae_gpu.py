"""GPU parity of the VQ-VAE engine (ResEncoder -> DVQ/EMA -> ResDecoder, forward / backward / Adam, all
through the C-ABI) against the oracle on seeded weights and inputs.

bf16 tensor-core convolutions (fp32 accumulate) vs the fp32 oracle: z_e / reconstruction within 2e-2
relative L2; code indices from the engine's OWN z_e agree with the oracle's on >= 98 % of positions
(a flipped index needs a best-vs-second gap below the bf16 conv noise; the codebook search itself is
bit-exact, test_vq_gpu.py).  With z_e teacher-forced to the oracle's, indices are bit-exact and the
EMA / commitment / decoder path is compared tightly: losses 1e-3 relative, updated codebook 1e-4,
parameter gradients cos >= 0.99."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n, L, seed=1234):
    from oracle import lvt_oracle as O
    cfg = O.VQVAEConfig(n_layers=L)
    eshape, gshape = O.vqvae_param_shapes(cfg)
    we, wg = O.synth_weights(eshape, seed=11), O.synth_weights(gshape, seed=12)
    x = torch.rand((n, 3, 64, 64), generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        z_ref = O.res_encoder((x - 0.5) / 0.5, we, cfg.n_layers)
    cb = torch.randn((4, 512, 64), generator=torch.Generator().manual_seed(5)) * z_ref.std()
    return cfg, we, wg, x, cb, z_ref


def _rel(a, b):
    return ((a - b).double().norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("n,L", [(8, 2), (4, 4)])
def test_vqvae_inference_vs_oracle(cuda_lib, n, L):
    from oracle import lvt_oracle as O
    from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg, we, wg, x, cb, z_ref = _setup(n, L)
    eng = VQVAEEngine(VQVAESpec(n_layers=L))
    eng.load_state_dict(we, wg, cb)
    w = eng.workspace(n, train=False)
    w.x.copy_(x)
    recon, idx = eng.inference(w)
    torch.cuda.synchronize()
    z_got = w.z_e.cpu().view(n, 16, 16, 256).permute(0, 3, 1, 2)
    assert _rel(z_got, z_ref) <= 2e-2
    with torch.no_grad():
        recon_ref, idx_ref = O.vqvae_inference(x, we, wg, cb, cfg)
    assert (idx.cpu() == idx_ref).float().mean().item() >= 0.98
    # decoder alone, driven by the oracle's indices (VQVAEModel.decode, vqvae.py:103-106)
    xt = eng.decode_indices(w, idx_ref.cuda().contiguous()).cpu()
    with torch.no_grad():
        xt_ref = O.res_decoder(O.dvq_embed(idx_ref, cb), wg, cfg.n_layers)
    assert _rel(xt, xt_ref) <= 2e-2
    assert recon.min().item() >= 0.0 and recon.max().item() <= 1.0


@pytest.mark.parametrize("n,L", [(32, 2), (8, 4)])
def test_precise_encoder_latents_vs_oracle(cuda_lib, n, L):
    """BASELINE.json config 1 (PR-DVQVAE2 forward + VQ on one synthetic 64x64x3 batch of 32 frames; also the K-DVQVAE
    depth): the high-precision encoder (3-term bf16 split products, VQVAEEngine.encode_precise) gives z_e within 5e-5
    relative L2 of the fp32 oracle and code indices that agree on >= 99.9 % of the 32768 (position, codebook) pairs
    -- against ~98-99 % for the plain bf16 encoder; the rate is printed."""
    from oracle import lvt_oracle as O
    from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    cfg, we, wg, x, cb, z_ref = _setup(n, L)
    eng = VQVAEEngine(VQVAESpec(n_layers=L))
    eng.load_state_dict(we, wg, cb)
    w = eng.workspace(n, train=False)
    w.x.copy_(x)
    with torch.no_grad():
        _, idx_ref = O.vqvae_inference(x, we, wg, cb, cfg)
    rates = {}
    for precise in (False, True):
        recon, idx = eng.inference(w, precise=precise)
        torch.cuda.synchronize()
        z_got = w.z_e.cpu().view(n, 16, 16, 256).permute(0, 3, 1, 2)
        rates[precise] = ((idx.cpu() == idx_ref).float().mean().item(), _rel(z_got, z_ref))
    print(f"latent agreement with the oracle, {n} frames, N_LAYERS {L}: bf16 encoder {rates[False][0]:.5f} "
          f"(z_e rel-L2 {rates[False][1]:.2e}), split encoder {rates[True][0]:.5f} (z_e rel-L2 {rates[True][1]:.2e})")
    assert rates[True][1] <= 5e-5, rates
    assert rates[True][0] >= 0.999, rates
    assert rates[False][0] >= 0.98, rates


def test_vqvae_train_step_vs_oracle(cuda_lib):
    from oracle import lvt_oracle as O
    from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
    n, L = 8, 2
    cfg, we, wg, x, cb, z_ref = _setup(n, L)
    rs0 = torch.full((4, 512), 5.0)           # warm EMA state (a fresh one divides by ~eps for unused codes)
    rsum0 = cb * 5.0
    eng = VQVAEEngine(VQVAESpec(n_layers=L))
    eng.load_state_dict(we, wg, cb, running_size=rs0, running_sum=rsum0)
    eng.init_optimizer()
    w = eng.workspace(n, train=True)
    w.x.copy_(x)
    eng.store.grad.zero_()
    eng.encode(w)
    w.z_e.copy_(z_ref.permute(0, 2, 3, 1).reshape(-1, 256))   # teacher-force z_e -> identical indices
    eng.quantize(w, train=True)
    eng.ema_update(w)
    eng.decode(w)
    w.loss.zero_()
    k = eng.kG
    assert eng.lib.lvt_vqvae_recon_loss(w.x_tilde.data_ptr(), w.x.data_ptr(), w.dpre.data_ptr(), w.loss.data_ptr(),
                                        eng.store.gf(f"G.layers.{k + 3}.bias"), n, 0.5, 0.5, 1.0,
                                        torch.cuda.current_stream().cuda_stream) == 0
    assert eng.lib.lvt_vqvae_commit_loss(w.z_e.data_ptr(), w.zq_bar.data_ptr(), None, None, w.loss.data_ptr() + 4,
                                         w.M * 256, 1.0, torch.cuda.current_stream().cuda_stream) == 0
    eng.backward(w)
    torch.cuda.synchronize()

    we_g = {k_: v.clone().requires_grad_(True) for k_, v in we.items()}
    wg_g = {k_: v.clone().requires_grad_(True) for k_, v in wg.items()}
    losses, aux = O.vqvae_supervised_loss(x, we_g, wg_g, cb, rs0.clone(), rsum0.clone(), cfg)
    sum(losses.values()).backward()
    assert torch.equal(w.idx.cpu(), aux["idx"])
    got = w.loss.tolist()
    assert abs(got[0] - losses["loss_reconstruction"].item()) <= 1e-3 * losses["loss_reconstruction"].item()
    assert abs(got[1] - losses["loss_commitment"].item()) <= 1e-3 * losses["loss_commitment"].item()
    assert torch.allclose(eng.codebook.cpu(), aux["codebooks"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(eng.running_size.cpu(), aux["running_size"], rtol=1e-5, atol=1e-6)
    bad = []
    for pre, sd in (("E.", we_g), ("G.", wg_g)):
        for name, p in sd.items():
            gw, gg = p.grad, eng.store.g[pre + name].cpu()
            cos = (gg.double().flatten() @ gw.double().flatten() / (gg.double().norm() * gw.double().norm())).item()
            ratio = (gg.double().norm() / gw.double().norm()).item()
            if not (cos >= 0.99 and abs(ratio - 1) <= 0.03):
                bad.append((pre + name, cos, ratio))
    assert not bad, bad
    # one Adam step moves the weights like torch.optim.Adam would (sign / magnitude: lr * g/|g| at step 1)
    before = eng.store.master.clone()
    eng.optimizer_step()
    torch.cuda.synchronize()
    delta = (eng.store.master - before).abs().max().item()
    assert 0 < delta <= 3e-4 * 1.001


def test_vqvae_noema_codebook_gradient_vs_oracle(cuda_lib):
    """MODEL.CODEBOOK.EMA False (vq_embedding.py:36-38,61-66, vqvae.py:84-88): the codebook is trained by gradient.
    Losses (1e-3), the codebook gradient of lvt_vq_codebook_grad against autograd's index_select backward (1e-4: it is
    built from the same fp32 counts / sums) and the Adam update of the codebook against torch.optim.Adam."""
    from oracle import lvt_oracle as O
    from lvt_b200.modeling.vqvae_engine import VQVAEEngine, VQVAESpec
    n, L = 8, 2
    cfg, we, wg, x, cb, z_ref = _setup(n, L)
    cfg = O.VQVAEConfig(n_layers=L, ema=False)
    eng = VQVAEEngine(VQVAESpec(n_layers=L, ema=False))
    eng.load_state_dict(we, wg, cb)
    eng.init_optimizer()
    w = eng.workspace(n, train=True)
    w.x.copy_(x)
    eng.store.grad.zero_()
    eng.encode(w)
    w.z_e.copy_(z_ref.permute(0, 2, 3, 1).reshape(-1, 256))   # teacher-force z_e -> identical indices
    eng.quantize(w, train=True)
    eng._forward_train_b(w)
    eng.backward(w)
    torch.cuda.synchronize()

    we_g = {k_: v.clone().requires_grad_(True) for k_, v in we.items()}
    wg_g = {k_: v.clone().requires_grad_(True) for k_, v in wg.items()}
    cb_g = cb.clone().requires_grad_(True)
    losses, aux = O.vqvae_supervised_loss(x, we_g, wg_g, cb_g, None, None, cfg)
    sum(losses.values()).backward()
    assert torch.equal(w.idx.cpu(), aux["idx"])
    assert torch.equal(eng.codebook.cpu(), cb)                  # no EMA update happened
    got = w.loss.tolist()
    for i, k_ in enumerate(("loss_reconstruction", "loss_commitment", "loss_dict")):
        assert abs(got[i] - losses[k_].item()) <= 1e-3 * losses[k_].item(), (k_, got[i], losses[k_].item())
    g_got, g_want = eng.cb_grad.cpu(), cb_g.grad
    assert ((g_got - g_want).double().norm() / g_want.double().norm()).item() <= 1e-4
    # encoder gradient still carries straight-through + commitment (same tolerance as the EMA test)
    gw, gg = we_g["layers.4.weight"].grad, eng.store.g["E.layers.4.weight"].cpu()
    cos = (gg.double().flatten() @ gw.double().flatten() / (gg.double().norm() * gw.double().norm())).item()
    assert cos >= 0.99 and abs((gg.double().norm() / gw.double().norm()).item() - 1) <= 0.03
    # the codebook's Adam step (optimizer_c of vqvae.py:108-116) against torch.optim.Adam on the oracle's gradient
    ref = cb.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-4, betas=(0.9, 0.9))
    ref.grad = g_want.clone()
    opt.step()
    eng.optimizer_step()
    torch.cuda.synchronize()
    touched = g_want.abs() > 1e-12
    assert touched.any() and torch.allclose(eng.codebook.cpu()[touched], ref.detach()[touched], rtol=0, atol=3e-6)
    assert torch.equal(eng.codebook.cpu()[~touched], cb[~touched])


def test_vqvae_graphed_step_matches_eager(cuda_lib):
    """GraphedVQVAEStep (CUDA-graph replay of forward + EMA + backward, Adam outside) against the eager train_step
    from identical state.  Step 1 is compared tightly (same code indices, losses, EMA codebook; Adam's first update
    is lr * g/|g|, so weights agree within 2 lr); two more replayed steps must keep the loss finite and falling
    (the split-K weight gradients are accumulated atomically, so longer trajectories drift by rounding noise)."""
    from lvt_b200.modeling.vqvae_engine import GraphedVQVAEStep, VQVAEEngine, VQVAESpec
    n, L = 8, 2
    cfg, we, wg, x, cb, z_ref = _setup(n, L)
    runs = []
    for graphed in (False, True):
        eng = VQVAEEngine(VQVAESpec(n_layers=L))
        eng.load_state_dict(we, wg, cb, running_size=torch.full((4, 512), 5.0), running_sum=cb * 5.0)
        eng.init_optimizer()
        w = eng.workspace(n, train=True)
        w.x.copy_(x)
        if graphed:
            stepper = GraphedVQVAEStep(eng, w)
            state = (eng.store.master.clone(), eng.codebook.clone(), eng.running_size.clone(), eng.running_sum.clone())
            stepper.capture(warmup=1)
            # the warm-up step moved the state: restore it so that both runs start from the same point
            eng.store.master.copy_(state[0]); eng.codebook.copy_(state[1])
            eng.running_size.copy_(state[2]); eng.running_sum.copy_(state[3])
            eng.opt_m.zero_(); eng.opt_v.zero_(); eng.opt["step"] = 0
            eng.refresh_shadows()
            step = stepper.step
        else:
            step = lambda: eng.train_step(w)  # noqa: E731
        loss1 = step().clone()
        torch.cuda.synchronize()
        first = (loss1.cpu(), w.idx.cpu().clone(), eng.codebook.cpu().clone(), eng.store.master.cpu().clone())
        later = [step().clone().cpu() for _ in range(2)]
        torch.cuda.synchronize()
        runs.append((first, later))
    (l0, i0, c0, m0), later0 = runs[0]
    (l1, i1, c1, m1), later1 = runs[1]
    assert torch.allclose(l0, l1, rtol=1e-3, atol=1e-6), (l0, l1)
    assert torch.equal(i0, i1)
    assert torch.allclose(c0, c1, rtol=1e-4, atol=1e-6)
    assert (m0 - m1).abs().max().item() <= 2 * 3e-4 * 1.01
    for later in (later0, later1):
        assert all(torch.isfinite(v).all() for v in later) and later[-1][0] < l0[0]
