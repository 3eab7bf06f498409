"""GPU parity of the VQ codebook kernels (through the C-ABI) against the oracle.
Bar: code indices bit-exact (integer result); gathered z_q bit-exact (pure copy); EMA statistics
within fp32 summation-order tolerance (atomics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(n, num, K, D, h, w, seed, init="default"):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn((n, num * D, h, w), generator=g) * 0.3
    if init == "default":  # VQEmbedding.__init__: uniform(-1/K, 1/K) (vq_embedding.py:13)
        cb = (torch.rand((num, K, D), generator=g) * 2 - 1) / K
    else:                  # well-spread codes
        cb = torch.randn((num, K, D), generator=g) * z.std()
    return z, cb


@pytest.mark.parametrize("init", ["default", "spread"])
@pytest.mark.parametrize("n,h,w", [(32, 16, 16), (3, 16, 16), (1, 5, 7), (0, 16, 16)])
def test_vq_argmin_bit_exact(cuda_lib, n, h, w, init):
    from lvt_b200 import ops
    from oracle import vq as ovq
    z, cb = _inputs(n, 4, 512, 64, h, w, seed=100 + n, init=init)
    idx, zq = ops.vq_argmin(z.cuda(), cb.cuda(), want_zq=True)
    torch.cuda.synchronize()
    if n == 0:
        assert idx.shape == (0, 4, h, w)
        return
    want = ovq.vq_argmin_c(z, cb)
    assert torch.equal(idx.cpu(), want), f"{(idx.cpu() != want).sum().item()} index mismatches"
    # the C oracle itself equals the literal torch restatement of vq_utils.py on this host
    want_t = ovq.dvq_argmin_torch(z, cb)
    assert torch.equal(want, want_t)
    # straight-through value = gathered codebook rows (vq_utils.py:42-44), NCHW
    zq_want = torch.cat([cb[g][want[:, g]].permute(0, 3, 1, 2) for g in range(4)], dim=1)
    assert torch.equal(zq.cpu(), zq_want)


def test_vq_argmin_ties_first_index(cuda_lib):
    """Duplicate codebook rows force exact distance ties; torch.min returns the first index."""
    from lvt_b200 import ops
    from oracle import vq as ovq
    z, cb = _inputs(4, 4, 512, 64, 16, 16, seed=7, init="spread")
    cb[:, 256:] = cb[:, :256]
    idx = ops.vq_argmin(z.cuda(), cb.cuda())
    want = ovq.vq_argmin_c(z, cb)
    assert torch.equal(idx.cpu(), want)
    assert int(idx.max()) < 256


def test_vq_generic_path(cuda_lib):
    """CODEBOOK.NUM == 1 (Base-VQVAE.yaml): one 512 x 256 codebook."""
    from lvt_b200 import ops
    from oracle import vq as ovq
    z, cb = _inputs(2, 1, 512, 256, 16, 16, seed=9, init="spread")
    idx = ops.vq_argmin(z.cuda(), cb.cuda())
    assert torch.equal(idx.cpu(), ovq.vq_argmin_c(z, cb))


def test_vq_gather_and_ema(cuda_lib):
    from lvt_b200 import ops
    n, num, K, D = 8, 4, 512, 64
    z, cb = _inputs(n, num, K, D, 16, 16, seed=11, init="spread")
    zc, cbc = z.cuda(), cb.cuda()
    counts = torch.zeros((num, K), device="cuda")
    sums = torch.zeros((num, K, D), device="cuda")
    idx = ops.vq_argmin(zc, cbc, counts=counts, sums=sums)
    out = ops.vq_gather(idx, cbc)
    torch.cuda.synchronize()
    want = torch.cat([cb[g][idx.cpu()[:, g]].permute(0, 3, 1, 2) for g in range(num)], dim=1)
    assert torch.equal(out.cpu(), want)
    # EMA statistics (vq_embedding.py:44-55)
    rs = torch.rand(num, K) * 3
    rsum = torch.randn(num, K, D)
    cb_new, rs_new, rsum_new = [], [], []
    for g in range(num):
        ind = idx.cpu()[:, g].reshape(-1)
        x = z[:, g * D:(g + 1) * D].permute(0, 2, 3, 1).reshape(-1, D)
        size = torch.zeros(K).index_add_(0, ind, torch.ones(ind.numel()))
        s = torch.zeros(K, D).index_add_(0, ind, x)
        assert torch.equal(counts[g].cpu(), size)
        assert torch.allclose(sums[g].cpu(), s, rtol=1e-4, atol=1e-5)
        r1 = rs[g] * 0.99 + (1 - 0.99) * size
        r2 = rsum[g] * 0.99 + (1 - 0.99) * s
        nn_ = r1.sum()
        size_ = (r1 + 1e-5) / (nn_ + K * 1e-5) * nn_
        cb_new.append(r2 / size_[:, None]); rs_new.append(r1); rsum_new.append(r2)
    rs_d, rsum_d = rs.cuda(), rsum.cuda()
    ops.vq_ema_update(cbc, rs_d, rsum_d, counts, sums, 0.99, 1e-5)
    torch.cuda.synchronize()
    assert torch.allclose(rs_d.cpu(), torch.stack(rs_new), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rsum_d.cpu(), torch.stack(rsum_new), rtol=1e-4, atol=1e-5)
    assert torch.allclose(cbc.cpu(), torch.stack(cb_new), rtol=2e-4, atol=1e-5)


def test_vq_tensor_core_path_matches_exact_path_on_adversarial_codebooks(cuda_lib):
    """The tcgen05 scan + exact re-rank must give the oracle's indices even when many codes are nearly
    tied (near-duplicate rows, clustered codes, tiny default-init codebooks, large dynamic range)."""
    from lvt_b200 import ops
    from oracle import vq as ovq
    g = torch.Generator().manual_seed(21)
    n = 6
    z = torch.randn((n, 256, 16, 16), generator=g) * 0.3
    base = torch.randn((4, 512, 64), generator=g) * 0.3
    cbs = {
        "near_duplicates": base.clone(),
        "clustered": (base[:, :8].repeat_interleave(64, dim=1) + 1e-3 * torch.randn((4, 512, 64), generator=g)),
        "default_init": (torch.rand((4, 512, 64), generator=g) * 2 - 1) / 512,
        "wide_range": base * torch.logspace(-3, 1, 512)[None, :, None],
    }
    cbs["near_duplicates"][:, 1::2] = cbs["near_duplicates"][:, 0::2] * (1 + 1e-6)
    for name, cb in cbs.items():
        idx = ops.vq_argmin(z.cuda(), cb.contiguous().cuda())
        torch.cuda.synchronize()
        want = ovq.vq_argmin_c(z, cb.contiguous())
        bad = (idx.cpu() != want).sum().item()
        assert bad == 0, f"{name}: {bad} index mismatches"
    # EMA statistics / straight-through output through the same kernel
    counts = torch.zeros((4, 512), device="cuda")
    sums = torch.zeros((4, 512, 64), device="cuda")
    cb = base.contiguous()
    idx, zq = ops.vq_argmin(z.cuda(), cb.cuda(), want_zq=True, counts=counts, sums=sums)
    torch.cuda.synchronize()
    want = ovq.vq_argmin_c(z, cb)
    assert torch.equal(idx.cpu(), want)
    zq_want = torch.cat([cb[gg][want[:, gg]].permute(0, 3, 1, 2) for gg in range(4)], dim=1)
    assert torch.equal(zq.cpu(), zq_want)
    assert counts.sum().item() == n * 256 * 4
