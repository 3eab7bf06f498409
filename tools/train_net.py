#!/usr/bin/env python
"""Training / evaluation entry point with the reference's command line (tools/train_net.py:60-104):

    python tools/train_net.py --config-file <reference yaml | preset name> [--num-gpus N] [--eval-only] KEY VALUE ...

Batches come from the synthetic generators of lvt_b200.data, or — for the video transformer — from a tree of latent
codes in the reference's on-disk format when LVT_LATENT_ROOT=<dir> is set (written by the CodesExtractor,
lvt_b200.data.latents); a loader can also be supplied programmatically (Trainer(cfg, data_loader))."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from lvt_b200.config import get_cfg  # noqa: E402
from lvt_b200.config.presets import PRESETS, preset  # noqa: E402
from lvt_b200.data import prepare_slices, sample_abc, synthetic_latent_video  # noqa: E402
from lvt_b200.engine import Trainer, default_argument_parser, default_setup, launch  # noqa: E402
from lvt_b200.utils import comm  # noqa: E402


def setup(args):
    if args.config_file in PRESETS:
        cfg = preset(args.config_file)
    else:
        cfg = get_cfg()
        cfg.merge_from_file(args.config_file)
    cfg.merge_from_list(args.opts)
    cfg.freeze()
    default_setup(cfg, args)
    return cfg


def synthetic_loader(cfg):
    """Infinite iterator of list[dict] batches in the reference mapper's format; per-rank batch =
    IMS_PER_BATCH / world_size (data/build.py:62-74)."""
    per_rank = max(1, cfg.SOLVER.IMS_PER_BATCH // comm.get_world_size())
    rng = random.Random(cfg.SEED + comm.get_rank())
    g = torch.Generator().manual_seed(max(0, cfg.SEED) + comm.get_rank())
    step = 0
    if cfg.MODEL.META_ARCHITECTURE == "VideoTransformerModel":
        vt = cfg.MODEL.AUTOREGRESSIVE.VT
        T = cfg.INPUT.N_FRAMES_PER_VIDEO_TRAIN
        while True:
            batch = []
            for i in range(per_rank):
                video = synthetic_latent_video(rng.randrange(1 << 30), (T, vt.NC, 16, 16), vt.NV)
                batch.append(prepare_slices(video, sample_abc(tuple(vt.STRIDE), T, vt.N_PRIME, rng), tuple(vt.KERNEL),
                                            tuple(vt.STRIDE), vt.N_PRIME, vt.PAD_VALUE))
            step += 1
            yield batch
    else:
        while True:
            yield [{"image": torch.rand((3, 64, 64), generator=g)} for _ in range(per_rank)]


def main(args):
    cfg = setup(args)
    latent_root = os.environ.get("LVT_LATENT_ROOT")
    if latent_root and cfg.MODEL.META_ARCHITECTURE == "VideoTransformerModel":
        from lvt_b200.data import latent_slice_loader
        loader = latent_slice_loader(cfg, latent_root)
    else:
        loader = synthetic_loader(cfg)
    trainer = Trainer(cfg, data_loader=loader)
    trainer.resume_or_load(resume=args.resume)
    if args.eval_only:
        model = trainer.model
        model.train(False)
        out = model(next(synthetic_loader(cfg))[:2] if cfg.MODEL.META_ARCHITECTURE == "VQVAEModel" else
                    [{"image_sequence": synthetic_latent_video(0)}])
        print({k: tuple(v.shape) if hasattr(v, "shape") else type(v).__name__ for k, v in out[0].items()})
        return
    return trainer.train()


if __name__ == "__main__":
    args = default_argument_parser().parse_args()
    print("Command Line Args:", args)
    launch(main, args.num_gpus, num_machines=args.num_machines, machine_rank=args.machine_rank, dist_url=args.dist_url,
           args=(args,))
