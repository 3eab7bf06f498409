"""CPU: host-side logic of the product (slice construction, config, C-ABI surface)."""
import os
import re

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prepare_slices_matches_reference_mapper():
    from lvt_b200.data import prepare_slices, subscale_order, synthetic_latent_video
    fix = np.load(os.path.join(GOLD, "mapper.npz"))
    specs = {"DSFVT": ((7, 1, 1), (16, 1, 1), 1, 16), "DSSVT": ((1, 3, 3), (1, 2, 2), 1, 4),
             "DSTSVT": ((5, 3, 3), (4, 2, 2), 1, 16)}
    for name, (kernel, stride, n_prime, T) in specs.items():
        idx2abc, _ = subscale_order(*stride)
        for i in range(3):
            video = synthetic_latent_video(500 + i, (T, 4, 16, 16))
            abc = idx2abc[int(fix[f"{name}:{i}:slice_idx"])]
            got = prepare_slices(video, abc, kernel, stride, n_prime)
            for k in ("context", "slice", "ignore_mask"):
                assert np.array_equal(got[k].numpy(), fix[f"{name}:{i}:{k}"]), (name, i, k)


def test_synthetic_batch_agrees_with_oracle_generator():
    from lvt_b200.data import synthetic_vt_batch
    from oracle import lvt_oracle as O
    got = synthetic_vt_batch(3, seed=77)
    want = O.synth_vt_batch(3, seed=77, cfg=O.VTConfig())
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_cabi_library_exports_every_declared_symbol():
    """include/lvt_b200.h <-> liblvt_b200.so <-> ctypes table (no compute without a GPU)."""
    from lvt_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lvt_b200.h")).read()
    declared = set(re.findall(r"\b(lvt_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    assert lib.lvt_abi_version() == 1
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} missing from the ctypes table"
    assert declared == set(_lib.SYMBOLS)


def test_product_fails_loudly_without_gpu():
    from lvt_b200 import _lib, ops
    if torch.cuda.is_available():
        return
    try:
        ops.vq_argmin(torch.zeros(1, 256, 16, 16), torch.zeros(4, 512, 64))
    except _lib.LvtError:
        return
    raise AssertionError("expected LvtError on a machine without a B200")


def test_batched_slice_preparation_equals_per_sample_mapper():
    """prepare_slices_batched (one gather per tensor, any device) == stacking prepare_slices, the restatement of the
    reference mapper (dataset_mapper.py:113-149), for the three shipped VT configs incl. n_prime = 0 and 5."""
    import random
    from lvt_b200.data import prepare_slices, prepare_slices_batched, sample_abc, synthetic_latent_video
    specs = {"DSFVT": ((7, 1, 1), (16, 1, 1), 16), "DSSVT": ((1, 3, 3), (1, 2, 2), 4), "DSTSVT": ((5, 3, 3), (4, 2, 2), 16)}
    for name, (kernel, stride, T) in specs.items():
        for n_prime in (0, 1, 5):
            rng = random.Random(11)
            vids = torch.stack([synthetic_latent_video(70 + i, (T, 4, 16, 16)) for i in range(5)])
            abcs = [sample_abc(stride, T, min(n_prime, stride[0] - 1) if stride[0] > 1 else 0, rng) for _ in range(5)]
            want = [prepare_slices(vids[i], abcs[i], kernel, stride, n_prime, -1) for i in range(5)]
            got = prepare_slices_batched(vids, torch.tensor(abcs), kernel, stride, n_prime, -1)
            for key in ("context", "slice", "slice_idx", "ignore_mask"):
                w = torch.stack([x[key] for x in want])
                assert got[key].dtype == w.dtype and torch.equal(got[key], w), (name, n_prime, key)


def test_gradient_buckets_tile_the_flat_buffer_in_backward_order():
    """VTEngine.backward_plan's buckets (bucket_ranges): contiguous, cover the whole flat gradient from the top down,
    and every parameter lies in the bucket of the backward segment that completes it."""
    from lvt_b200.modeling.autoregressive.vt_engine import ParamStore, VTSpec, bucket_ranges, layer_cuts
    spec = VTSpec()
    store = ParamStore(spec.param_shapes(), "cpu")
    for parts in (1, 2, 4, 8):
        r = bucket_ranges(store.offsets, store.numel, 8, 8, parts)
        assert len(r) == 2 * parts and r[0][1] == store.numel and r[-1][0] == 0
        assert all(r[i][0] == r[i + 1][1] for i in range(len(r) - 1)) and all(lo < hi for lo, hi in r)
        cd = layer_cuts(8, parts)

        def bucket_of(name):
            o = store.offsets[name]
            return next(i for i, (lo, hi) in enumerate(r) if lo <= o < hi)
        assert bucket_of("ch_predictor.P.3.bias") == 0 and bucket_of("decoder.block_local_attention.7.ffn.3.bias") == 0
        assert bucket_of("decoder.ch_embedder.0.weight") == parts - 1 and bucket_of("decoder.conv.conv.weight") == parts - 1
        assert bucket_of("encoder.block_local_attention.7.mha.w_q") == parts
        assert bucket_of("encoder.conv.weight") == 2 * parts - 1 and bucket_of("encoder.block_local_attention.0.dt_bank") == 2 * parts - 1
        for j in range(parts):   # decoder layers cd[j]-1 .. cd[j+1] are completed by segment j
            for i in range(cd[j + 1], cd[j]):
                assert bucket_of(f"decoder.block_local_attention.{i}.ffn.1.weight") == j
                assert bucket_of(f"encoder.block_local_attention.{i}.ffn.1.weight") == parts + j


def test_reference_arm_line_contract():
    """`bench.py --impl reference` (CPU only: the reference's own implementation of the path, or the oracle port when
    baseline/_ref is absent) prints ONE JSON line with the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "latent tokens/sec DSFVT train step"
    assert line["unit"] == "latent tokens/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 1 and line["config"]["per_gpu_batch"] == 8
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
