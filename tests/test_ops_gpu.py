"""GPU unit parity of the bandwidth-bound DSFVT operators (C-ABI) against plain PyTorch fp32
references of the same op (CPU).  Outputs stored as bf16 are compared at bf16 resolution
(rtol 8e-3 of the tensor scale); fp32 outputs and reductions at 1e-4 / 1e-3."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _close(got, want, tol):
    got, want = got.float().cpu(), want.float().cpu()
    scale = want.abs().max().item() + 1e-12
    err = (got - want).abs().max().item()
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("M,d", [(1000, 512), (256, 128), (77, 256)])
def test_layernorm_fwd_bwd(cuda_lib, M, d):
    lib = cuda_lib
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, d, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    dy = torch.randn(M, d, generator=g)
    dres = torch.randn(M, d, generator=g)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.layer_norm(xr, (d,), gr, br)
    y.backward(dy)
    xd, gd, bd = x.cuda(), gamma.cuda(), beta.cuda()
    yd = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    assert lib.lvt_layernorm_fwd(_p(xd), _p(gd), _p(bd), _p(yd), _p(mean), _p(rstd), M, d, 1e-5, _s()) == 0
    _close(yd, y.detach(), 8e-3)
    dx = torch.empty(M, d, device="cuda")
    dxb = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    assert lib.lvt_layernorm_bwd(_p(dy.cuda()), _p(xd), _p(mean), _p(rstd), _p(gd), _p(dres.cuda()), _p(dx), _p(dxb),
                                 _p(dg), _p(db), M, d, _s()) == 0
    torch.cuda.synchronize()
    _close(dx, xr.grad + dres, 1e-4)
    _close(dxb, xr.grad + dres, 8e-3)
    _close(dg, gr.grad, 1e-4)
    _close(db, br.grad, 1e-4)


def test_colsum_delta_bankgrad(cuda_lib):
    lib = cuda_lib
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3000, 512, generator=g).to(torch.bfloat16)
    out = torch.ones(512, device="cuda")
    assert lib.lvt_colsum_bf16(_p(x.cuda()), _p(out), 3000, 512, 512, _s()) == 0
    _close(out, 1 + x.float().sum(0), 1e-4)
    nb, H, L, da = 3, 8, 256, 128
    dO = torch.randn(nb * L, H * da, generator=g).to(torch.bfloat16)
    O = torch.randn(nb * L, H * da, generator=g).to(torch.bfloat16)
    delta = torch.empty(nb, H, L, device="cuda")
    assert lib.lvt_attn_delta(_p(dO.cuda()), _p(O.cuda()), _p(delta), nb, H, L, da, _s()) == 0
    want = (dO.float() * O.float()).view(nb, L, H, da).sum(-1).permute(0, 2, 1)
    _close(delta, want, 1e-4)
    for block in [(1, 16, 16), (4, 8, 8)]:
        bt, bh, bw = block
        dS = (torch.randn(nb, H, L, L, generator=g) * 0.1).to(torch.bfloat16)
        banks = [torch.zeros(H, 2 * n - 1, requires_grad=True) for n in block]
        i = torch.arange(L)
        t, h, w = i // (bh * bw), (i // bw) % bh, i % bw
        B = (banks[0][:, t[:, None] - t[None, :] + bt - 1] + banks[1][:, h[:, None] - h[None, :] + bh - 1]
             + banks[2][:, w[:, None] - w[None, :] + bw - 1])
        (B[None] * dS.float()).sum().backward()
        outs = [torch.zeros(H, 2 * n - 1, device="cuda") for n in block]
        assert lib.lvt_relpos_bank_grad(_p(dS.cuda()), _p(outs[0]), _p(outs[1]), _p(outs[2]), nb, H, bt, bh, bw, _s()) == 0
        for o, b in zip(outs, banks):
            _close(o, b.grad, 1e-3)


def test_cross_entropy(cuda_lib):
    lib = cuda_lib
    g = torch.Generator().manual_seed(2)
    B, nc, nv, thw = 3, 4, 512, 256
    M = B * thw
    logits = torch.randn(nc, M, nv, generator=g) * 3
    slc = torch.randint(0, nv, (B, nc, thw), generator=g)
    ignore = torch.rand(B, thw, generator=g) < 0.3
    lr = logits.clone().requires_grad_(True)
    target = slc.masked_fill(ignore[:, None, :], -100)
    loss = sum(F.cross_entropy(lr[k].view(B, thw, nv).permute(0, 2, 1), target[:, k], ignore_index=-100)
               for k in range(nc)) / nc
    loss.backward()
    dl = torch.empty(nc, M, nv, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(1, device="cuda")
    cnt = torch.empty(1, device="cuda", dtype=torch.int32)
    assert lib.lvt_cross_entropy(_p(logits.cuda()), _p(slc.cuda()), _p(ignore.to(torch.uint8).cuda()), _p(dl), _p(out),
                                 _p(cnt), B, nc, nv, thw, _s()) == 0
    torch.cuda.synchronize()
    assert int(cnt.item()) == int((~ignore).sum())
    assert abs(out.item() - loss.item()) <= 1e-5 * abs(loss.item())
    _close(dl, lr.grad, 8e-3)


def test_optimizers(cuda_lib):
    lib = cuda_lib
    from oracle import lvt_oracle as O
    g = torch.Generator().manual_seed(3)
    n = 4096
    p0, grads = torch.randn(n, generator=g), [torch.randn(n, generator=g) * 0.1 for _ in range(3)]
    p, sq, buf = p0.clone(), torch.zeros(n), torch.zeros(n)
    pd, sqd, bufd = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for gr in grads:
        O.rmsprop_step(p, gr, sq, buf, lr=2e-5, alpha=0.95, momentum=0.9)
        assert lib.lvt_rmsprop_step(_p(pd), _p(gr.cuda()), _p(sqd), _p(bufd), _p(pb), n, 2e-5, 0.95, 0.9, 1e-8, 1.0, _s()) == 0
    torch.cuda.synchronize()
    assert torch.allclose(pd.cpu(), p, rtol=1e-6, atol=1e-7)
    assert torch.allclose(pb.float().cpu(), p, rtol=8e-3, atol=1e-6)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    pd, md, vd = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step, gr in enumerate(grads, 1):
        O.adam_step(p, gr, m, v, step, lr=3e-4, beta1=0.9, beta2=0.9)
        assert lib.lvt_adam_step(_p(pd), _p(gr.cuda()), _p(md), _p(vd), None, n, 3e-4, 0.9, 0.9, 1e-8, step, 1.0, _s()) == 0
    torch.cuda.synchronize()
    assert torch.allclose(pd.cpu(), p, rtol=1e-5, atol=1e-6)
    # torch.optim cross-check of the oracle's restatement
    q = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([q], lr=3e-4, betas=(0.9, 0.9))
    for gr in grads:
        q.grad = gr.clone()
        opt.step()
    assert torch.allclose(q.detach(), p, rtol=1e-5, atol=1e-6)
