"""GPU parity of the tcgen05 GEMM (through the C-ABI) against a plain PyTorch fp32 reference of
the same op on the same bf16-rounded inputs.  Tolerance: fp32 accumulation of exact bf16
products differs from torch's fp32 matmul only by summation order -> rtol 2e-3 of the output
scale when the result is read back as bf16, 1e-4 when read back as fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(got, want, tol):
    got = got.float()
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item()
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (1024, 512, 512),
                                   (200, 72, 96), (384, 256, 1024), (4096, 3072, 512)])
def test_gemm_kmajor_plain(cuda_lib, M, N, K):
    from lvt_b200 import ops
    a, b = _rand((M, K), 1), _rand((N, K), 2)
    want = a.float() @ b.float().t()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(out.data_ptr(), N), out_f32=out)
    torch.cuda.synchronize()
    _close(out, want, 1e-4)
    outb = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(outb.data_ptr(), N), out_bf16=outb)
    torch.cuda.synchronize()
    _close(outb, want, 6e-3)


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (512, 384, 256), (256, 256, 2048)])
def test_gemm_mn_major(cuda_lib, a_mn, b_mn, M, N, K):
    from lvt_b200 import ops
    a, b = _rand((M, K), 3), _rand((N, K), 4)
    want = a.float() @ b.float().t()
    at = a.t().contiguous()  # [K, M]
    bt = b.t().contiguous()  # [K, N]
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(M, N, K, ops.op_mnmajor(at) if a_mn else ops.op_kmajor(a),
             ops.op_mnmajor(bt) if b_mn else ops.op_kmajor(b), ops.Operand(out.data_ptr(), N), out_f32=out)
    torch.cuda.synchronize()
    _close(out, want, 1e-4)


def test_gemm_epilogue_bias_res_relu_mask(cuda_lib):
    from lvt_b200 import ops
    M, N, K = 512, 512, 512
    a, b = _rand((M, K), 5), _rand((N, K), 6, 0.05)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    aux = _rand((M, N), 7)
    base = a.float() @ b.float().t()
    # bias + residual + relu, dual output
    o32 = torch.empty((M, N), device="cuda")
    o16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(o32.data_ptr(), N), out_f32=o32,
             out_bf16=o16, bias=bias, res=res, flags=ops.GEMM_RELU, alpha=0.5)
    torch.cuda.synchronize()
    want = torch.relu(0.5 * base + bias + res)
    _close(o32, want, 1e-4)
    _close(o16, want, 6e-3)
    # relu-backward mask
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(o32.data_ptr(), N), out_f32=o32,
             aux=aux, flags=ops.GEMM_MASK)
    torch.cuda.synchronize()
    _close(o32, base * (aux.float() > 0), 1e-4)
    # row-periodic 2-D bias table (positional encoding add)
    tab = torch.randn(128, N, device="cuda")
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(o32.data_ptr(), N), out_f32=o32,
             bias=tab, bias_mod=128)
    torch.cuda.synchronize()
    _close(o32, base + tab.repeat(M // 128, 1), 1e-4)
    # residual aliasing the output (accumulate)
    acc = res.clone()
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(acc.data_ptr(), N), out_f32=acc, res=acc)
    torch.cuda.synchronize()
    _close(acc, base + res, 1e-4)


def test_gemm_splitk_atomic(cuda_lib):
    from lvt_b200 import ops
    # weight-gradient shape: dW[N_out, K_in] = dY^T X, contraction over 4096 tokens
    T, NO, KI = 4096, 512, 384
    dy, x = _rand((T, NO), 8), _rand((T, KI), 9)
    want = dy.float().t() @ x.float()
    out = torch.zeros((NO, KI), device="cuda")
    ops.gemm(NO, KI, T, ops.op_mnmajor(dy), ops.op_mnmajor(x), ops.Operand(out.data_ptr(), KI),
             out_f32=out, splits=8, flags=ops.GEMM_ATOMIC)
    torch.cuda.synchronize()
    _close(out, want, 1e-4)


def test_gemm_blocked_qkv_weights(cuda_lib):
    """w_q/w_k/w_v are stored (head, d, da) (vt_attention.py:101-103): forward uses them as an
    MN-major B with 128-wide column blocks, dgrad as a K-major B with blocked k, wgrad writes a
    blocked output."""
    from lvt_b200 import ops
    M, d, H, da = 512, 256, 4, 128
    x = _rand((M, d), 10)
    w = _rand((3 * H, d, da), 11, 0.05)  # [q heads | k heads | v heads]
    N = 3 * H * da
    want = torch.cat([x.float() @ w[i].float() for i in range(3 * H)], dim=1)  # [M, N]
    out = torch.empty((M, N), device="cuda")
    bop = ops.Operand(w.data_ptr(), da, mn_major=True, cin=da, s_blk=d * da)
    ops.gemm(M, N, d, ops.op_kmajor(x), bop, ops.Operand(out.data_ptr(), N), out_f32=out)
    torch.cuda.synchronize()
    _close(out, want, 1e-4)
    # dgrad: dx[M, d] = dqkv[M, N] @ Wcat^T, contraction over N with blocked k
    dqkv = _rand((M, N), 12)
    wcat = torch.cat([w[i].float() for i in range(3 * H)], dim=1)  # [d, N]
    want_dx = dqkv.float() @ wcat.t()
    dx = torch.empty((M, d), device="cuda")
    bop2 = ops.Operand(w.data_ptr(), da, mn_major=False, cin=da, s_blk=d * da)
    ops.gemm(M, d, N, ops.op_kmajor(dqkv), bop2, ops.Operand(dx.data_ptr(), d), out_f32=dx)
    torch.cuda.synchronize()
    _close(dx, want_dx, 1e-4)
    # wgrad: dW[blk][k][a] = sum_m x[m,k] dqkv[m, blk*da + a]
    dw = torch.zeros((3 * H, d, da), device="cuda")
    want_dw = torch.stack([x.float().t() @ dqkv.float()[:, i * da:(i + 1) * da] for i in range(3 * H)])
    oop = ops.Operand(dw.data_ptr(), da, cin=da, s_blk=d * da)
    ops.gemm(d, N, M, ops.op_mnmajor(x), ops.op_mnmajor(dqkv), oop, out_f32=dw, splits=2,
             flags=ops.GEMM_ATOMIC)
    torch.cuda.synchronize()
    _close(dw, want_dw, 1e-4)


def _attn_ref(q, k, v, banks, block, causal):
    """vt_attention.py:61-81 + get_B (169-174) in fp32 on the bf16-rounded inputs."""
    bt, bh, bw = block
    L = bt * bh * bw
    idx = torch.arange(L)
    t, h, w = idx // (bh * bw), (idx // bw) % bh, idx % bw
    B = (banks[0][:, (t[:, None] - t[None, :] + bt - 1)] + banks[1][:, (h[:, None] - h[None, :] + bh - 1)]
         + banks[2][:, (w[:, None] - w[None, :] + bw - 1)])  # [H, L, L]
    s = torch.einsum("bhid,bhjd->bhij", q, k) / math.sqrt(q.shape[-1]) + B[None]
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(L, L), diagonal=1).bool(), -1e4)
    p = torch.softmax(s, dim=-1)
    return p, torch.logsumexp(s, dim=-1), torch.einsum("bhij,bhjd->bhid", p, v)


@pytest.mark.parametrize("block,causal", [((1, 16, 16), True), ((1, 16, 16), False), ((4, 8, 8), True)])
def test_attention_via_gemm(cuda_lib, block, causal):
    """S=QK^T (+bias, mask, softmax in the epilogue) and O=PV straight out of a packed
    [M, 3*H*da] qkv buffer, as the DSFVT layer uses them."""
    from lvt_b200 import ops
    Bsz, H, da, L = 3, 8, 128, 256
    M = Bsz * L
    qkv = _rand((M, 3 * H * da), 13, 0.5)
    g = torch.Generator().manual_seed(14)
    banks = [torch.randn(H, 2 * n - 1, generator=g) * 0.5 for n in block]
    banks_d = [b.cuda().contiguous() for b in banks]
    qf = qkv.float().cpu().view(Bsz, L, 3, H, da)
    q, k, v = [qf[:, :, i].permute(0, 2, 1, 3) for i in range(3)]  # [B,H,L,da]
    p_ref, lse_ref, o_ref = _attn_ref(q, k, v, banks, block, causal)

    ld = 3 * H * da
    P = torch.empty((Bsz, H, L, L), device="cuda", dtype=torch.bfloat16)
    lse = torch.empty((Bsz * H, L), device="cuda")

    def qkv_op(which, mn):
        return ops.Operand(qkv.data_ptr() + 2 * which * H * da, ld, mn_major=mn, cin=da, zdiv=H,
                           s_zlo=da, s_zhi=L * ld)

    ops.gemm(L, L, da, qkv_op(0, False), qkv_op(1, False),
             ops.Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=P, batch=Bsz * H,
             alpha=1.0 / math.sqrt(da), mode=ops.EPI_SOFTMAX,
             flags=ops.GEMM_CAUSAL if causal else 0, lse=lse, banks=banks_d, block=block, heads=H)
    torch.cuda.synchronize()
    assert (P.float().cpu() - p_ref).abs().max().item() < 4e-3
    assert (lse.cpu().view(Bsz, H, L) - lse_ref).abs().max().item() < 2e-3

    # O[b, i, h*da + d] = sum_j P[b,h,i,j] V[b,h,j,d]  (head-major concat, vt_attention.py:125-126)
    O = torch.empty((M, H * da), device="cuda", dtype=torch.bfloat16)
    ops.gemm(L, da, L, ops.Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), qkv_op(2, True),
             ops.Operand(O.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da),
             out_bf16=O, batch=Bsz * H)
    torch.cuda.synchronize()
    o_want = torch.einsum("bhij,bhjd->bhid", P.float().cpu(), v).permute(0, 2, 1, 3).reshape(M, H * da)
    _close(O.cpu(), o_want, 8e-3)

    # dS epilogue: dS = P * (dP - delta)
    dO = _rand((M, H * da), 15)
    delta = (dO.float() * O.float()).view(Bsz, L, H, da).sum(-1).permute(0, 2, 1).contiguous()  # [B,H,L]
    dS = torch.empty_like(P)
    ops.gemm(L, L, da, ops.Operand(dO.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da),
             qkv_op(2, False), ops.Operand(dS.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=dS,
             batch=Bsz * H, mode=ops.EPI_DS, aux=P, delta=delta)
    torch.cuda.synchronize()
    dOh = dO.float().cpu().view(Bsz, L, H, da).permute(0, 2, 1, 3)
    dP = torch.einsum("bhid,bhjd->bhij", dOh, v)
    want = P.float().cpu() * (dP - delta.cpu()[..., None])
    _close(dS.cpu(), want, 1e-2)


@pytest.mark.parametrize("M,N,K,blk,L", [(512, 1024, 512, 128, 256), (768, 384, 256, 128, 256), (256, 256, 128, 64, 128)])
def test_gemm_rowdot_epilogue(cuda_lib, M, N, K, blk, L):
    """LVT_GEMM_ROWDOT: delta[seq, head, i] = sum over the head's columns of (A B^T) * aux — the softmax
    backward row term rowsum(dO * O) (vt_attention.py:75-80) fused into the GEMM that produces dO."""
    from lvt_b200 import ops
    a, b = _rand((M, K), 21), _rand((N, K), 22, 0.05)
    aux = _rand((M, N), 23)
    want_o = a.float() @ b.float().t()
    H = N // blk
    want_d = (want_o * aux.float()).view(M // L, L, H, blk).sum(-1).permute(0, 2, 1).contiguous()  # [seq, head, L]
    out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    delta = torch.full((M // L, H, L), float("nan"), device="cuda")
    ops.gemm(M, N, K, ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(out.data_ptr(), N), out_bf16=out, aux=aux,
             rowdot=delta, rd_block=blk, rd_L=L)
    torch.cuda.synchronize()
    _close(out, want_o, 6e-3)
    _close(delta, want_d, 1e-4)


@pytest.mark.parametrize("block,causal,store_p", [((1, 16, 16), True, True), ((1, 16, 16), False, False),
                                                   ((4, 8, 8), True, True)])
@pytest.mark.parametrize("Bsz", [1, 5, 40])  # 40 x 8 heads x 2 tiles = 640 tiles: several tiles per persistent CTA
def test_fused_attention_forward(cuda_lib, block, causal, store_p, Bsz):
    """LVT_EPI_SOFTMAX with V: P = softmax(QK^T/sqrt(da) + B [mask]) and O = P V in ONE kernel (P handed to the
    second MMA through shared memory), against fp32 torch on the same bf16 inputs; P optional."""
    from lvt_b200 import ops
    H, da, L = 8, 128, 256
    M = Bsz * L
    qkv = _rand((M, 3 * H * da), 31, 0.5)
    g = torch.Generator().manual_seed(32)
    banks = [torch.randn(H, 2 * n - 1, generator=g) * 0.5 for n in block]
    banks_d = [b.cuda().contiguous() for b in banks]
    qf = qkv.float().cpu().view(Bsz, L, 3, H, da)
    q, k, v = [qf[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    p_ref, lse_ref, o_ref = _attn_ref(q, k, v, banks, block, causal)
    ld = 3 * H * da
    P = torch.full((Bsz, H, L, L), float("nan"), device="cuda", dtype=torch.bfloat16)
    O = torch.full((M, H * da), float("nan"), device="cuda", dtype=torch.bfloat16)

    def qkv_op(which, mn):
        return ops.Operand(qkv.data_ptr() + 2 * which * H * da, ld, mn_major=mn, cin=da, zdiv=H,
                           s_zlo=da, s_zhi=L * ld)

    for _ in range(2):  # twice: the persistent pipeline must leave no stale state behind
        ops.gemm(L, L, da, qkv_op(0, False), qkv_op(1, False),
                 ops.Operand(P.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=P if store_p else None, batch=Bsz * H,
                 alpha=1.0 / math.sqrt(da), mode=ops.EPI_SOFTMAX, flags=ops.GEMM_CAUSAL if causal else 0,
                 banks=banks_d, block=block, heads=H, v=qkv_op(2, True),
                 o2=ops.Operand(O.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da), o2_n=da)
    torch.cuda.synchronize()
    if store_p:
        assert (P.float().cpu() - p_ref).abs().max().item() < 4e-3
    o_want = o_ref.permute(0, 2, 1, 3).reshape(M, H * da)
    _close(O.cpu(), o_want, 1e-2)


@pytest.mark.parametrize("Bsz", [1, 3, 40])  # 40: several tiles per persistent CTA
def test_fused_attention_backward_ds_dq(cuda_lib, Bsz):
    """LVT_EPI_DS with K as second operand: dS = P * (dO V^T - delta) and dQ = scale * dS K in ONE kernel
    (dS handed to the second MMA through shared memory), against fp32 torch on the same bf16 inputs."""
    from lvt_b200 import ops
    H, da, L = 8, 128, 256
    M = Bsz * L
    scale = 1.0 / math.sqrt(da)
    qkv = _rand((M, 3 * H * da), 41, 0.5)
    dO = _rand((M, H * da), 42)
    g = torch.Generator().manual_seed(43)
    P = torch.softmax(torch.randn((Bsz, H, L, L), generator=g) * 2.0, -1).to(torch.bfloat16).cuda()
    delta = (torch.randn((Bsz, H, L), generator=g) * 0.5).cuda()
    ld = 3 * H * da

    def qkv_op(buf, which, mn):
        return ops.Operand(buf.data_ptr() + 2 * which * H * da, ld, mn_major=mn, cin=da, zdiv=H,
                           s_zlo=da, s_zhi=L * ld)

    dS = torch.full((Bsz, H, L, L), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv = torch.full((M, ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    block = (1, 16, 16)
    gbanks = [torch.full((H, 2 * n - 1), 1.0, device="cuda") for n in block]   # accumulated into (+=)
    for rep in range(2):
        ops.gemm(L, L, da, ops.Operand(dO.data_ptr(), H * da, cin=da, zdiv=H, s_zlo=da, s_zhi=L * H * da),
                 qkv_op(qkv, 2, False), ops.Operand(dS.data_ptr(), L, zdiv=1, s_zhi=L * L), out_bf16=dS,
                 batch=Bsz * H, mode=ops.EPI_DS, aux=P, delta=delta, alpha=scale,
                 v=qkv_op(qkv, 1, True), o2=qkv_op(dqkv, 0, False), o2_n=da,
                 banks=gbanks if rep == 1 else None, block=block, heads=H)   # second launch: fused bank gradient too
    torch.cuda.synchronize()
    qf = qkv.float().cpu().view(Bsz, L, 3, H, da)
    k, v = [qf[:, :, i].permute(0, 2, 1, 3) for i in (1, 2)]       # [B,H,L,da]
    dOh = dO.float().cpu().view(Bsz, L, H, da).permute(0, 2, 1, 3)
    dP = torch.einsum("bhid,bhjd->bhij", dOh, v)
    want_dS = P.float().cpu() * (dP - delta.cpu()[..., None])
    _close(dS.cpu(), want_dS, 1e-2)
    want_dQ = scale * torch.einsum("bhij,bhjd->bhid", dS.float().cpu(), k)   # from the bf16 dS the kernel used
    got_dQ = dqkv.float().cpu().view(Bsz, L, 3, H, da)[:, :, 0].permute(0, 2, 1, 3)
    _close(got_dQ, want_dQ, 1e-2)
    # bank gradients (BlockLocalAttention.get_B, vt_attention.py:169-174) from the same dS, on top of the initial 1.0
    bt, bh, bw = block
    banks = [torch.zeros(H, 2 * n - 1, requires_grad=True) for n in block]
    i = torch.arange(L)
    t, h, w = i // (bh * bw), (i // bw) % bh, i % bw
    B = (banks[0][:, t[:, None] - t[None, :] + bt - 1] + banks[1][:, h[:, None] - h[None, :] + bh - 1]
         + banks[2][:, w[:, None] - w[None, :] + bw - 1])
    (B[None] * want_dS).sum().backward()
    for got, b in zip(gbanks, banks):
        _close(got.cpu() - 1.0, b.grad, 2e-3 if b.shape[1] > 1 else 1.0)   # (H, 1) dt bank: sum of all dS, pure cancellation
