"""GPU parity of the CTA-pair (tcgen05 cta_group::2, 256 x 256 tiles, cluster of 2) form of lvt_gemm_bf16
against a plain PyTorch fp32 reference on the same bf16-rounded inputs.  The library picks the pair kernel
for linear-epilogue GEMMs with 256-wide tiles, M >= 256 and more than 37 output tiles, i.e. for every large
GEMM of the DSFVT / VQ-VAE train steps; the shapes below are chosen to land there and to cover every
operand-major combination, every store path of the epilogue, a ragged last row tile (the peer CTA of the
last pair owns no rows), split-K reduction and the implicit-GEMM convolution operands.
Tolerance: 1e-4 of the output scale for fp32 outputs, bf16 resolution for bf16 outputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(got, want, tol):
    got = got.float()
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item()
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(4096, 512, 512), (2048 + 128, 1024, 192), (16384, 256, 64)])
def test_pair_gemm_majors(cuda_lib, a_mn, b_mn, M, N, K):
    from lvt_b200 import ops
    a, b = _rand((M, K), 3), _rand((N, K), 4)
    want = a.float() @ b.float().t()
    at, bt = a.t().contiguous(), b.t().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(M, N, K, ops.op_mnmajor(at) if a_mn else ops.op_kmajor(a),
             ops.op_mnmajor(bt) if b_mn else ops.op_kmajor(b), ops.Operand(out.data_ptr(), N), out_f32=out)
    torch.cuda.synchronize()
    _close(out, want, 1e-4)
    outb = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(M, N, K, ops.op_mnmajor(at) if a_mn else ops.op_kmajor(a),
             ops.op_mnmajor(bt) if b_mn else ops.op_kmajor(b), ops.Operand(outb.data_ptr(), N), out_bf16=outb)
    torch.cuda.synchronize()
    _close(outb, want, 6e-3)


def test_pair_gemm_epilogues(cuda_lib):
    from lvt_b200 import ops
    M, N, K = 8192, 512, 512
    a, b = _rand((M, K), 5), _rand((N, K), 6, 0.05)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    aux = _rand((M, N), 7)
    base = a.float() @ b.float().t()
    o32 = torch.empty((M, N), device="cuda")
    o16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    A, B, O = ops.op_kmajor(a), ops.op_kmajor(b), ops.Operand(o32.data_ptr(), N)
    # generic staged epilogue: dual output, bias + residual + relu
    ops.gemm(M, N, K, A, B, O, out_f32=o32, out_bf16=o16, bias=bias, res=res, flags=ops.GEMM_RELU, alpha=0.5)
    torch.cuda.synchronize()
    want = torch.relu(0.5 * base + bias + res)
    _close(o32, want, 1e-4)
    _close(o16, want, 6e-3)
    # TMA store fp32 + TMA-loaded fp32 residual (+ bias): FFN2 / proj of the DSFVT layer
    ops.gemm(M, N, K, A, B, O, out_f32=o32, bias=bias, res=res)
    torch.cuda.synchronize()
    _close(o32, base + bias + res, 1e-4)
    # residual aliasing the output
    acc = res.clone()
    ops.gemm(M, N, K, A, B, ops.Operand(acc.data_ptr(), N), out_f32=acc, res=acc)
    torch.cuda.synchronize()
    _close(acc, base + res, 1e-4)
    # TMA store bf16 + bias + relu: FFN1
    ops.gemm(M, N, K, A, B, ops.Operand(o16.data_ptr(), N), out_bf16=o16, bias=bias, flags=ops.GEMM_RELU)
    torch.cuda.synchronize()
    _close(o16, torch.relu(base + bias), 6e-3)
    # relu-backward mask from a TMA-loaded bf16 tensor
    ops.gemm(M, N, K, A, B, ops.Operand(o16.data_ptr(), N), out_bf16=o16, aux=aux, flags=ops.GEMM_MASK)
    torch.cuda.synchronize()
    _close(o16, base * (aux.float() > 0), 6e-3)
    # bf16 residual add + relu (ResBlock tail)
    ops.gemm(M, N, K, A, B, ops.Operand(o16.data_ptr(), N), out_bf16=o16, bias=bias, aux=aux,
             flags=ops.GEMM_AUX_ADD | ops.GEMM_RELU)
    torch.cuda.synchronize()
    _close(o16, torch.relu(base + bias + aux.float()), 8e-3)
    # row-periodic bias table (positional encoding), generic epilogue
    tab = torch.randn(256, N, device="cuda")
    ops.gemm(M, N, K, A, B, O, out_f32=o32, bias=tab, bias_mod=256)
    torch.cuda.synchronize()
    _close(o32, base + tab.repeat(M // 256, 1), 1e-4)
    # rowdot: delta[seq, block, i] = sum over 128-column blocks of out * aux
    L, blk = 256, 128
    rd = torch.zeros((M // L, N // blk, L), device="cuda")
    ops.gemm(M, N, K, A, B, ops.Operand(o16.data_ptr(), N), out_bf16=o16, aux=aux, rowdot=rd, rd_block=blk, rd_L=L)
    torch.cuda.synchronize()
    _close(o16, base, 6e-3)
    want_rd = (base * aux.float()).view(M // L, L, N // blk, blk).sum(-1).permute(0, 2, 1)
    _close(rd, want_rd, 2e-3)


@pytest.mark.parametrize("splits", [-1, 5])
def test_pair_gemm_splitk_reduce(cuda_lib, splits):
    from lvt_b200 import ops
    # weight-gradient shapes of the DSFVT layer: dW[512, 512] and dW[512, 1024], contraction over the tokens
    for T, NO, KI in [(16384, 512, 512), (8192, 512, 1024)]:
        dy, x = _rand((T, NO), 8, 0.1), _rand((T, KI), 9, 0.1)
        want = dy.float().t() @ x.float()
        out = torch.zeros((NO, KI), device="cuda")
        ops.gemm(NO, KI, T, ops.op_mnmajor(dy), ops.op_mnmajor(x), ops.Operand(out.data_ptr(), KI),
                 out_f32=out, splits=splits, flags=ops.GEMM_ATOMIC)
        torch.cuda.synchronize()
        _close(out, want, 1e-4)


def test_pair_gemm_blocked_qkv(cuda_lib):
    """(head, d, da) weights: blocked MN-major B forward, blocked-k dgrad, blocked-output split-K wgrad."""
    from lvt_b200 import ops
    M, d, H, da = 4096, 512, 8, 128
    x = _rand((M, d), 10)
    w = _rand((3 * H, d, da), 11, 0.05)
    N = 3 * H * da
    want = torch.cat([x.float() @ w[i].float() for i in range(3 * H)], dim=1)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    bop = ops.Operand(w.data_ptr(), da, mn_major=True, cin=da, s_blk=d * da)
    ops.gemm(M, N, d, ops.op_kmajor(x), bop, ops.Operand(out.data_ptr(), N), out_bf16=out)
    torch.cuda.synchronize()
    _close(out, want, 6e-3)
    dqkv = _rand((M, N), 12)
    wcat = torch.cat([w[i].float() for i in range(3 * H)], dim=1)
    dx = torch.empty((M, d), device="cuda")
    bop2 = ops.Operand(w.data_ptr(), da, mn_major=False, cin=da, s_blk=d * da)
    ops.gemm(M, d, N, ops.op_kmajor(dqkv), bop2, ops.Operand(dx.data_ptr(), d), out_f32=dx)
    torch.cuda.synchronize()
    _close(dx, dqkv.float() @ wcat.t(), 1e-4)
    dw = torch.zeros((3 * H, d, da), device="cuda")
    want_dw = torch.stack([x.float().t() @ dqkv.float()[:, i * da:(i + 1) * da] for i in range(3 * H)])
    oop = ops.Operand(dw.data_ptr(), da, cin=da, s_blk=d * da)
    ops.gemm(d, N, M, ops.op_mnmajor(x), ops.op_mnmajor(dqkv), oop, out_f32=dw, splits=-1, flags=ops.GEMM_ATOMIC)
    torch.cuda.synchronize()
    _close(dw, want_dw, 1e-4)


def test_pair_conv3x3_fwd_and_wgrad(cuda_lib):
    from lvt_b200 import ops
    from lvt_b200.ops import ConvSpec, Operand
    n, C, CO, H, W = 48, 256, 256, 16, 16
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, C, H, W), generator=g).to(torch.bfloat16)
    w = (torch.randn((CO, C, 3, 3), generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(CO, generator=g)
    want = F.conv2d(x.float(), w.float(), bias, padding=1)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    taps = [(kh - 1, kw - 1, 0) for kh in range(3) for kw in range(3)]
    wp = w.permute(0, 2, 3, 1).reshape(CO, 9 * C).contiguous().cuda()
    M = n * H * W
    out = torch.empty((M, CO), device="cuda")
    ops.gemm(M, CO, 9 * C, Operand(x_nhwc.data_ptr(), C), Operand(wp.data_ptr(), 9 * C), Operand(out.data_ptr(), CO),
             out_f32=out, bias=bias.cuda(), conv=ConvSpec("a", C, H, W, n, taps))
    torch.cuda.synchronize()
    got = out.view(n, H, W, CO).permute(0, 3, 1, 2).cpu()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-4 * scale
    dy = (torch.randn((n, CO, H, W), generator=g) * 0.1).to(torch.bfloat16)
    wg = w.float().requires_grad_(True)
    F.conv2d(x.float(), wg, None, padding=1).backward(dy.float())
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dwp = torch.zeros((CO, 9 * C), device="cuda")
    ops.gemm(CO, 9 * C, M, Operand(dy_nhwc.data_ptr(), CO, mn_major=True), Operand(x_nhwc.data_ptr(), C, mn_major=True),
             Operand(dwp.data_ptr(), 9 * C), out_f32=dwp, splits=-1, flags=ops.GEMM_ATOMIC,
             conv=ConvSpec("b", C, H, W, n, taps))
    torch.cuda.synchronize()
    gotw = dwp.view(CO, 3, 3, C).permute(0, 3, 1, 2).cpu()
    scale = wg.grad.abs().max().item()
    assert (gotw - wg.grad).abs().max().item() <= 1e-4 * scale
