#!/usr/bin/env python
"""Video generation with the reference's flow (scripts/generate_videos.py:53-100): priming frames -> VQ-VAE
codes -> autoregressive sampling of the remaining frames -> VQ-VAE decode -> PNGs.

    python scripts/generate_videos.py --video-dir <dir with 0.png..4.png> --config-file <VT yaml | preset>
                                      [--vqvae-config <yaml | preset>] [--out-dir out] [--n-frames 16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lvt_b200.config import get_cfg  # noqa: E402
from lvt_b200.config.presets import PRESETS, preset  # noqa: E402
from lvt_b200.engine.checkpoint import Checkpointer  # noqa: E402
from lvt_b200.modeling import build_model  # noqa: E402


def load_cfg(name, overrides=()):
    if name in PRESETS:
        cfg = preset(name, overrides)
    else:
        cfg = get_cfg()
        cfg.merge_from_file(name)
        cfg.merge_from_list(list(overrides))
    cfg.freeze()
    return cfg


def load_frames(video_dir, n):
    from PIL import Image
    frames = [np.asarray(Image.open(os.path.join(video_dir, f"{i}.png")).convert("RGB"), dtype=np.float32) / 255.
              for i in range(n)]
    return torch.from_numpy(np.stack(frames).transpose(0, 3, 1, 2).copy())


@torch.no_grad()
def sample_videos(args):
    vt_cfg = load_cfg(args.config_file, ["TEST.EVALUATORS", "VTSampler", "TEST.VT_SAMPLER.NUM_SAMPLES", 1])
    vq_name = args.vqvae_config or vt_cfg.TEST.VT_SAMPLER.VQ_VAE.CFG or "PR-DVQVAE2"
    vq_cfg = load_cfg(vq_name if (vq_name in PRESETS or os.path.exists(vq_name)) else "PR-DVQVAE2")
    vt, vqvae = build_model(vt_cfg), build_model(vq_cfg)
    s = vt_cfg.TEST.VT_SAMPLER.VQ_VAE
    for module, path in ((vt.model, vt_cfg.MODEL.GENERATOR.WEIGHTS), (vqvae.encoder, s.ENCODER_WEIGHTS),
                         (vqvae.generator, s.GENERATOR_WEIGHTS), (vqvae.codebook, s.CODEBOOK_WEIGHTS)):
        if path and os.path.exists(path):
            Checkpointer(module).load(path)
    vt.train(False)
    vqvae.train(False)
    n_prime = vt_cfg.TEST.VT_SAMPLER.N_PRIME
    frames = load_frames(args.video_dir, n_prime)                                   # (n_prime, 3, 64, 64) in [0,1]
    latent = vqvae([{"image_sequence": frames}])[0]["latent"]                       # (n_prime, nc, 16, 16)
    seq = torch.zeros((args.n_frames,) + tuple(latent.shape[1:]), dtype=torch.int64, device=latent.device)
    seq[:n_prime] = latent
    sample = vt([{"image_sequence": seq}])[0]["samples"][0]                         # (nc, T, 16, 16)
    video = vqvae.back_normalizer(vqvae.decode(sample.transpose(0, 1).contiguous())).clamp_(0, 1)
    os.makedirs(args.out_dir, exist_ok=True)
    from PIL import Image
    for i, fr in enumerate((video * 255).byte().permute(0, 2, 3, 1).cpu().numpy()):
        Image.fromarray(fr).save(os.path.join(args.out_dir, f"{i}.png"))
    return video


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--video-dir", required=True)
    ap.add_argument("--config-file", required=True)
    ap.add_argument("--vqvae-config", default="")
    ap.add_argument("--out-dir", default="generated")
    ap.add_argument("--n-frames", type=int, default=16)
    sample_videos(ap.parse_args())
