"""GPU parity of the implicit-GEMM convolution modes of lvt_gemm_bf16 (per-tap shifted TMA boxes
over NHWC bf16 activations) against torch.nn.functional convolutions in fp32 on the same
bf16-rounded inputs (ResEncoder / ResDecoder layers, encoder/resencoder.py:46-52,
generator/resdecoder.py:48-56).  Tolerance 1e-4 of the output scale (fp32 out)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# Conv2d(k=4, s=2, p=1): input row 2*o - 1 + k  ->  (parity, shift in the half-resolution grid)
K4S2 = {0: (1, -1), 1: (0, 0), 2: (1, 0), 3: (0, 1)}


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


def _close(got, want, tol):
    scale = want.abs().max().item() + 1e-9
    err = (got.float().cpu() - want).abs().max().item()
    assert err <= tol * scale, (err, scale)


def test_conv3x3_fwd_dgrad_wgrad(cuda_lib):
    from lvt_b200 import ops
    from lvt_b200.ops import ConvSpec, Operand
    n, C, CO, H, W = 3, 256, 128, 16, 16
    x = _rand((n, C, H, W), 1)            # NCHW values
    w = _rand((CO, C, 3, 3), 2, 0.05)
    bias = torch.randn(CO)
    want = F.conv2d(x.float(), w.float(), bias, padding=1)  # n, CO, H, W
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    taps = [(kh - 1, kw - 1, 0) for kh in range(3) for kw in range(3)]
    wp = w.permute(0, 2, 3, 1).reshape(CO, 9 * C).contiguous().cuda()  # [co][(kh,kw),ci]
    M = n * H * W
    out = torch.empty((M, CO), device="cuda")
    ops.gemm(M, CO, 9 * C, Operand(x_nhwc.data_ptr(), C), Operand(wp.data_ptr(), 9 * C), Operand(out.data_ptr(), CO),
             out_f32=out, bias=bias.cuda(), conv=ConvSpec("a", C, H, W, n, taps))
    torch.cuda.synchronize()
    _close(out.view(n, H, W, CO).permute(0, 3, 1, 2), want, 1e-4)

    # data gradient: dX = conv(dY, W^T flipped)
    dy = _rand((n, CO, H, W), 3)
    xg = x.float().requires_grad_(True)
    wg = w.float().requires_grad_(True)
    F.conv2d(xg, wg, None, padding=1).backward(dy.float())
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().cuda()
    taps_t = [(1 - kh, 1 - kw, 0) for kh in range(3) for kw in range(3)]
    wt = w.permute(1, 2, 3, 0).reshape(C, 9 * CO).contiguous().cuda()  # [ci][(kh,kw),co]
    dx = torch.empty((M, C), device="cuda")
    ops.gemm(M, C, 9 * CO, Operand(dy_nhwc.data_ptr(), CO), Operand(wt.data_ptr(), 9 * CO), Operand(dx.data_ptr(), C),
             out_f32=dx, conv=ConvSpec("a", CO, H, W, n, taps_t))
    torch.cuda.synchronize()
    _close(dx.view(n, H, W, C).permute(0, 3, 1, 2), xg.grad, 1e-4)

    # weight gradient: dW[co][(tap, ci)] = sum_m dY[m, co] X[m + tap, ci]   (split-K reduce-add)
    dwp = torch.zeros((CO, 9 * C), device="cuda")
    ops.gemm(CO, 9 * C, M, Operand(dy_nhwc.data_ptr(), CO, mn_major=True), Operand(x_nhwc.data_ptr(), C, mn_major=True),
             Operand(dwp.data_ptr(), 9 * C), out_f32=dwp, splits=3, flags=ops.GEMM_ATOMIC,
             conv=ConvSpec("b", C, H, W, n, taps))
    torch.cuda.synchronize()
    _close(dwp.view(CO, 3, 3, C).permute(0, 3, 1, 2), wg.grad, 1e-4)


def test_conv4x4_stride2_over_phase_major_input(cuda_lib):
    """Conv2d(128->256, k4, s2, p1) on a 32x32 input stored as 4 parity phases of 16x16."""
    from lvt_b200 import ops
    from lvt_b200.ops import ConvSpec, Operand
    n, C, CO = 2, 128, 256
    x = _rand((n, C, 32, 32), 4)
    w = _rand((CO, C, 4, 4), 5, 0.05)
    want = F.conv2d(x.float(), w.float(), None, stride=2, padding=1)  # n, CO, 16, 16
    # phase-major NHWC: [hp][wp][n][16][16][C]
    xp = x.view(n, C, 16, 2, 16, 2).permute(3, 5, 0, 2, 4, 1).contiguous().cuda()
    taps = [(K4S2[kh][1], K4S2[kw][1], K4S2[kh][0] * 2 + K4S2[kw][0]) for kh in range(4) for kw in range(4)]
    wp = w.permute(0, 2, 3, 1).reshape(CO, 16 * C).contiguous().cuda()
    M = n * 256
    out = torch.empty((M, CO), device="cuda")
    ops.gemm(M, CO, 16 * C, Operand(xp.data_ptr(), C), Operand(wp.data_ptr(), 16 * C), Operand(out.data_ptr(), CO),
             out_f32=out, conv=ConvSpec("a", C, 16, 16, n, taps, P=4, s_phase=n * 256 * C))
    torch.cuda.synchronize()
    _close(out.view(n, 16, 16, CO).permute(0, 3, 1, 2), want, 1e-4)


def test_conv_transpose4x4_stride2_as_four_phases(cuda_lib):
    """ConvTranspose2d(256->128, k4, s2, p1): each output parity phase is a 2x2-tap conv."""
    from lvt_b200 import ops
    from lvt_b200.ops import ConvSpec, Operand
    n, C, CO = 2, 256, 128
    x = _rand((n, C, 16, 16), 6)
    w = _rand((C, CO, 4, 4), 7, 0.05)  # ConvTranspose2d weight layout (in, out, kh, kw)
    want = F.conv_transpose2d(x.float(), w.float(), None, stride=2, padding=1)  # n, CO, 32, 32
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    M = n * 256
    # output row oh = 2q + ph gets taps kh with kh = ph + 1 (mod 2): ih = q + (ph + 1 - kh) / 2
    out = torch.empty((2, 2, M, CO), device="cuda")
    for ph in range(2):
        for pw in range(2):
            khs = [kh for kh in range(4) if (kh - ph - 1) % 2 == 0]
            kws = [kw for kw in range(4) if (kw - pw - 1) % 2 == 0]
            taps = [((ph + 1 - kh) // 2, (pw + 1 - kw) // 2, 0) for kh in khs for kw in kws]
            wp = torch.stack([w[:, :, kh, kw] for kh in khs for kw in kws], 0)  # [tap][ci][co]
            wp = wp.permute(2, 0, 1).reshape(CO, 4 * C).contiguous().cuda()
            ops.gemm(M, CO, 4 * C, Operand(x_nhwc.data_ptr(), C), Operand(wp.data_ptr(), 4 * C),
                     Operand(out[ph, pw].data_ptr(), CO), out_f32=out[ph, pw], conv=ConvSpec("a", C, 16, 16, n, taps))
    torch.cuda.synchronize()
    got = out.view(2, 2, n, 16, 16, CO).permute(2, 5, 3, 0, 4, 1).reshape(n, CO, 32, 32)
    _close(got, want, 1e-4)


def test_bf16_residual_epilogue(cuda_lib):
    """ResBlock tail (resencoder.py:10-21): out = relu(r + conv1x1(h) + b) with a bf16 skip tensor."""
    from lvt_b200 import ops
    from lvt_b200.ops import Operand
    M, K, N = 512, 128, 256
    h, w, r = _rand((M, K), 8), _rand((N, K), 9, 0.1), _rand((M, N), 10)
    bias = torch.randn(N)
    want = torch.relu(h.float() @ w.float().t() + bias + r.float())
    hd, wd, rd = h.cuda(), w.cuda(), r.cuda()
    ob = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(M, N, K, Operand(hd.data_ptr(), K), Operand(wd.data_ptr(), K), Operand(ob.data_ptr(), N), out_bf16=ob,
             bias=bias.cuda(), aux=rd, flags=ops.GEMM_AUX_ADD | ops.GEMM_RELU)
    of = torch.empty((M, N), device="cuda")
    ops.gemm(M, N, K, Operand(hd.data_ptr(), K), Operand(wd.data_ptr(), K), Operand(of.data_ptr(), N), out_f32=of,
             bias=bias.cuda(), aux=rd, flags=ops.GEMM_AUX_ADD)
    torch.cuda.synchronize()
    _close(ob, want, 8e-3)
    _close(of, h.float() @ w.float().t() + bias + r.float(), 1e-4)
