"""The latent wire format between the two halves of the system: the VQ-VAE writes code indices, the video
transformer trains on them.

Format (reference: evaluation/codes_extractor.py:36-53, data/datasets/latents.py:10-40, data/dataset_mapper.py:66-74):
one ``.npy`` per frame, int64, shape (nc, 16, 16), at ``<output_dir>/<dataset_name>/[<class_name>/]video_<idx>/<frame>.npy``;
a directory counts as a video when it has no sub-directories and only ``.npy`` files; frames are ordered by natural
sort ("2.npy" < "10.npy"); the listing is cached as a pickled list of dicts in ``<root>/latent_video_paths.npy``.
"""
import os
import random
import re
from collections import OrderedDict

import numpy as np
import torch

from ..utils import comm

CACHE_NAME = "latent_video_paths.npy"


def natural_sorted(items):
    """utils/strings.py: human order ("2" before "10")."""
    return sorted(items, key=lambda t: [int(c) if c.isdigit() else c for c in re.split(r"(\d+)", t)])


def video_dir(output_dir, dataset_name, video_idx, class_name=None):
    parts = [output_dir, dataset_name] + ([class_name] if class_name is not None else []) + [f"video_{video_idx}"]
    return os.path.join(*parts)


def save_latent_video(latent, output_dir, dataset_name, video_idx, class_name=None):
    """latent: (T, nc, h, w) or (T, h, w) integer codes -> one int64 ``<frame>.npy`` per frame
    (codes_extractor.py:38-52).  Returns the video directory."""
    latent = torch.as_tensor(latent)
    if latent.dim() == 3:
        latent = latent.unsqueeze(1)
    frames = latent.detach().to("cpu", torch.int64).numpy()  # ONE device->host copy per video, not one per frame
    vdir = video_dir(output_dir, dataset_name, video_idx, class_name)
    os.makedirs(vdir, exist_ok=True)
    for frame_idx in range(frames.shape[0]):
        np.save(os.path.join(vdir, f"{frame_idx}.npy"), frames[frame_idx])
    return vdir


def get_latent_video_paths(root, use_cache=True):
    """datasets/latents.py:10-40: list of {"video_path", "latent_paths", "video_idx"}."""
    assert os.path.isdir(root) or os.path.islink(root), f"{root} is not a valid directory"
    cache_path = os.path.join(root, CACHE_NAME)
    if use_cache and os.path.exists(cache_path):
        return np.load(cache_path, allow_pickle=True).tolist()
    video_paths = []
    video_idx = 0
    for cur, dirs, files in os.walk(root):
        if len(dirs) > 0:
            continue  # a video folder contains only npy files
        files = natural_sorted(files)
        if not all(f.endswith(".npy") for f in files):
            continue
        video_paths.append({"video_path": cur, "latent_paths": [os.path.join(cur, f) for f in files],
                            "video_idx": video_idx})
        video_idx += 1
    if use_cache and not os.path.exists(cache_path):
        np.save(cache_path, video_paths)
    return video_paths


def load_latent_video(entry, n_frames=-1, is_train=True, rng=random):
    """The ``latent_paths`` branch of DatasetMapper.__call__ (dataset_mapper.py:43-49,66-69): a random window of
    n_frames frames when training, the first n_frames otherwise, everything for n_frames == -1.
    Returns int64 (T, nc, h, w), or None when the video is too short (the reference drops such samples)."""
    paths = entry["latent_paths"]
    n = len(paths)
    if n < n_frames:
        return None
    start = 0 if (n_frames == -1 or not is_train) else rng.randint(0, n - n_frames)
    end = n if n_frames == -1 else start + n_frames
    return torch.from_numpy(np.stack([np.load(p) for p in paths[start:end]], axis=0))


class CodesExtractor:
    """Evaluator with the reference's reset / process / evaluate protocol (evaluation/codes_extractor.py:13-61):
    stores the ``latent`` output of the VQ-VAE for every input video."""

    def __init__(self, dataset_name, distributed, output_dir=None, class_names=None):
        self._dataset_name = dataset_name
        self._distributed = distributed
        self._output_dir = output_dir
        self._class_names = class_names  # KINETICS_IDX_LABEL-style mapping for dicts that carry "class"

    def reset(self):
        pass

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            cls = None
            if "class" in inp:
                c = int(inp["class"])
                cls = self._class_names[c] if self._class_names is not None else str(c)
            save_latent_video(out["latent"], self._output_dir, self._dataset_name, inp["video_idx"], cls)

    def evaluate(self):
        if self._distributed:
            comm.synchronize()
            if not comm.is_main_process():
                return None
        return OrderedDict({"latents": {}})


@torch.no_grad()
def extract_codes(model, videos, dataset_name, output_dir, videos_per_batch=8):
    """TEST.EVALUATORS "CodesExtractor"  without the reference's batch-1 loop: `videos` yields dicts
    {"image_sequence": (T, 3, 64, 64) float in [0, 1], "video_idx": int[, "class": int]}; several videos go through
    one VQ-VAE inference call (all frames of a batch are one encoder launch sequence), each rank takes a contiguous
    shard of the list (samplers/distributed_sampler.py:192-195)."""
    videos = list(videos)
    world, rank = comm.get_world_size(), comm.get_rank()
    shard = (len(videos) + world - 1) // world
    mine = videos[rank * shard:(rank + 1) * shard]
    ex = CodesExtractor(dataset_name, world > 1, output_dir)
    ex.reset()
    for i in range(0, len(mine), videos_per_batch):
        batch = mine[i:i + videos_per_batch]
        ex.process(batch, model(batch, mode="inference"))
    return ex.evaluate()


def latent_slice_loader(cfg, root, seed=None):
    """Infinite iterator of list[dict] training batches for the video transformer read from a latent tree:
    per-rank batch = IMS_PER_BATCH / world_size (data/build.py:62-74), rank r takes indices[r::world] of a shared-seed
    permutation (samplers/distributed_sampler.py:45-56), samples go through the slice preparation of
    DatasetMapper (dataset_mapper.py:113-149)."""
    from .slices import prepare_slices, sample_abc
    vt = cfg.MODEL.AUTOREGRESSIVE.VT
    entries = get_latent_video_paths(root)
    assert entries, f"no latent videos under {root}"
    world, rank = comm.get_world_size(), comm.get_rank()
    per_rank = max(1, cfg.SOLVER.IMS_PER_BATCH // world)
    T = cfg.INPUT.N_FRAMES_PER_VIDEO_TRAIN
    shared = random.Random(cfg.SEED if seed is None else seed)      # same permutation on every rank
    local = random.Random((cfg.SEED if seed is None else seed) + rank)
    batch = []
    while True:
        order = list(range(len(entries)))
        shared.shuffle(order)
        for idx in order[rank::world]:
            video = load_latent_video(entries[idx], T, True, local)
            if video is None:
                continue
            abc = sample_abc(tuple(vt.STRIDE), video.shape[0], vt.N_PRIME, local)
            batch.append(prepare_slices(video, abc, tuple(vt.KERNEL), tuple(vt.STRIDE), vt.N_PRIME, vt.PAD_VALUE))
            if len(batch) == per_rank:
                yield batch
                batch = []
