"""The latent on-disk format (VQ-VAE -> video transformer): writer / lister / loader of lvt_b200.data.latents against
the format the reference defines (evaluation/codes_extractor.py:36-53, data/datasets/latents.py:10-40,
data/dataset_mapper.py:43-74) and, when the reference is present (authoring container), against the reference's own
lister and mapper reading OUR files."""
import os
import random

import numpy as np
import pytest
import torch

from lvt_b200.data import latents as L
from lvt_b200.data import synthetic_latent_video


def _write_tree(root, n_videos=3, T=12):
    vids = []
    for v in range(n_videos):
        video = synthetic_latent_video(100 + v, (T, 4, 16, 16))
        L.save_latent_video(video, root, "bair_train", v)
        vids.append(video)
    return vids


def test_writer_layout_dtype_and_natural_order(tmp_path):
    vids = _write_tree(str(tmp_path))
    vdir = tmp_path / "bair_train" / "video_1"
    assert sorted(os.listdir(vdir)) == sorted(f"{i}.npy" for i in range(12))
    fr = np.load(vdir / "10.npy")
    assert fr.dtype == np.int64 and fr.shape == (4, 16, 16)
    assert np.array_equal(fr, vids[1][10].numpy())
    entries = L.get_latent_video_paths(str(tmp_path / "bair_train"))
    assert [e["video_idx"] for e in entries] == [0, 1, 2]
    assert set(entries[0]) == {"video_path", "latent_paths", "video_idx"}
    by_dir = {os.path.basename(e["video_path"]): e for e in entries}
    names = [os.path.basename(p) for p in by_dir["video_2"]["latent_paths"]]
    assert names == [f"{i}.npy" for i in range(12)]          # natural order: 2.npy before 10.npy
    # the listing is cached as a pickled list of dicts and reused
    cache = tmp_path / "bair_train" / L.CACHE_NAME
    assert cache.exists()
    assert np.load(cache, allow_pickle=True).tolist() == entries
    assert L.get_latent_video_paths(str(tmp_path / "bair_train")) == entries
    # 3-dim latents (single codebook) get a channel axis, like the reference
    L.save_latent_video(torch.zeros(2, 16, 16, dtype=torch.int32), str(tmp_path), "one", 0)
    assert np.load(tmp_path / "one" / "video_0" / "1.npy").shape == (1, 16, 16)


def test_loader_window_rules_and_round_trip(tmp_path):
    vids = _write_tree(str(tmp_path), n_videos=2, T=20)
    entries = L.get_latent_video_paths(str(tmp_path / "bair_train"), use_cache=False)
    e = next(x for x in entries if x["video_path"].endswith("video_0"))
    full = L.load_latent_video(e, -1)
    assert full.dtype == torch.int64 and torch.equal(full, vids[0])
    assert torch.equal(L.load_latent_video(e, 16, is_train=False), vids[0][:16])
    rng = random.Random(3)
    start = random.Random(3).randint(0, 4)
    assert torch.equal(L.load_latent_video(e, 16, True, rng), vids[0][start:start + 16])
    assert L.load_latent_video(e, 32) is None                # shorter than the window: dropped


def test_codes_extractor_protocol(tmp_path):
    ex = L.CodesExtractor("kin", False, str(tmp_path), class_names={7: "archery"})
    ex.reset()
    lat = synthetic_latent_video(5, (3, 4, 16, 16))
    ex.process([{"video_idx": 4, "class": torch.tensor(7)}, {"video_idx": 9}], [{"latent": lat}, {"latent": lat + 0}])
    assert ex.evaluate() == {"latents": {}}
    assert np.array_equal(np.load(tmp_path / "kin" / "archery" / "video_4" / "2.npy"), lat[2].numpy())
    assert (tmp_path / "kin" / "video_9" / "0.npy").exists()


def test_slice_loader_batches(tmp_path):
    from lvt_b200.config.presets import preset
    _write_tree(str(tmp_path), n_videos=5, T=16)
    cfg = preset("DSFVT", ["SOLVER.IMS_PER_BATCH", 4, "SEED", 11])
    it = L.latent_slice_loader(cfg, str(tmp_path / "bair_train"))
    batch = next(it)
    assert len(batch) == 4
    assert batch[0]["context"].shape == (4, 7, 16, 16) and batch[0]["slice"].shape == (4, 1, 16, 16)
    assert batch[0]["context"].dtype == torch.int64 and 1 <= int(batch[0]["slice_idx"]) <= 15
    next(it)  # wraps around the 5 videos


def test_reference_reads_our_files(tmp_path):
    """The unmodified reference lists and loads the tree written by lvt_b200 (authoring container only)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not present on this machine")
    ref_shim.install()
    from vidgen.data.datasets.latents import get_latent_video_paths as ref_list
    vids = _write_tree(str(tmp_path), n_videos=3, T=12)
    root = str(tmp_path / "bair_train")
    want = ref_list(root, use_cache=False)
    got = L.get_latent_video_paths(root, use_cache=False)
    assert got == want
    # and the reference mapper's loading rule on our files: np.stack of np.load over latent_paths
    e = next(x for x in want if x["video_path"].endswith("video_2"))
    video = np.stack([np.load(p) for p in e["latent_paths"]], axis=0)
    assert video.dtype == np.int64 and np.array_equal(video, vids[2].numpy())
