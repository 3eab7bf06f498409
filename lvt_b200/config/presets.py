"""The seven shipped experiment configurations of the reference (configs/vqvae/*.yaml, configs/vt/*.yaml) as
override lists for `get_cfg()`, so that everything runs without the reference checkout.  A YAML path from the
reference can be used instead via cfg.merge_from_file — both routes give the same tree."""
from .config import get_cfg

_BLK = lambda b: tuple([b] * 8)  # noqa: E731

_VQVAE_COMMON = [
    "MODEL.META_ARCHITECTURE", "VQVAEModel", "MODEL.INIT_TYPE", "xavier_uniform",
    "MODEL.ENCODER.NAME", "ResEncoder", "MODEL.ENCODER.NF", 256, "MODEL.ENCODER.OUT_CHANNELS", 256,
    "MODEL.ENCODER.RES_CHANNELS", 128, "MODEL.ENCODER.IN_CHANNELS", 3, "MODEL.ENCODER.NORM", "",
    "MODEL.GENERATOR.NAME", "ResDecoder", "MODEL.GENERATOR.IN_CHANNELS", 256, "MODEL.GENERATOR.RES_CHANNELS", 128,
    "MODEL.GENERATOR.NF", 256, "MODEL.GENERATOR.OUT_CHANNELS", 3, "MODEL.GENERATOR.OUT_ACTIVATION", "tanh",
    "MODEL.GENERATOR.NORM", "", "MODEL.CODEBOOK.SIZE", 512, "MODEL.CODEBOOK.DIM", 256, "MODEL.CODEBOOK.EMA", True,
    "MODEL.CODEBOOK.NUM", 4, "MODEL.PIXEL_MEAN", [0.5, 0.5, 0.5], "MODEL.PIXEL_STD", [0.5, 0.5, 0.5],
    "INPUT.FORMAT", "RGB", "INPUT.N_FRAMES_PER_VIDEO_TEST", 16, "SOLVER.IMS_PER_BATCH", 32, "SOLVER.LR_G", 0.0003,
    "SOLVER.LR_SCHEDULER_NAME", "Identity", "TEST.EVALUATORS", "MSEEvaluator,CodesExtractor", "SEED", 29871897,
]

_VT_COMMON = [
    "INPUT.SCALE_TO_ZEROONE", False, "INPUT.N_FRAMES_PER_VIDEO_TEST", 16, "INPUT.PREPARE_SLICES_TRAIN", True,
    "MODEL.META_ARCHITECTURE", "VideoTransformerModel", "MODEL.INIT_TYPE", "xavier_uniform",
    "MODEL.AUTOREGRESSIVE.NAME", "VideoTransformer", "MODEL.AUTOREGRESSIVE.VT.NC", 4, "MODEL.AUTOREGRESSIVE.VT.NV", 512,
    "MODEL.AUTOREGRESSIVE.VT.DE", 128, "MODEL.AUTOREGRESSIVE.VT.D", 512, "MODEL.AUTOREGRESSIVE.VT.DA", 128,
    "MODEL.AUTOREGRESSIVE.VT.N_HEAD_E", (8,) * 8, "MODEL.AUTOREGRESSIVE.VT.N_HEAD_D", (8,) * 8,
    "MODEL.AUTOREGRESSIVE.VT.N_PRIME", 1, "MODEL.AUTOREGRESSIVE.VT.SHARE_P", False,
    "SOLVER.IMS_PER_BATCH", 64, "SOLVER.MAX_ITER", 600000, "SOLVER.OPTIMIZER_NAME", "rmsprop", "SOLVER.LR_G", 0.00002,
    "SOLVER.RMSPROP.ALPHA_G", 0.95, "SOLVER.RMSPROP.MOMENTUM_G", 0.9, "SOLVER.LR_SCHEDULER_NAME", "Identity",
    "SOLVER.CHECKPOINT_PERIOD", 100000, "TEST.EVALUATORS", "BitsEvaluator", "TEST.VT_SAMPLER.N_PRIME", 5,
    "TEST.VT_SAMPLER.NUM_SAMPLES", 1, "SEED", 29871897, "VIS_PERIOD", 1000000000,
]


def _vt(kernel, stride, block, n_train):
    return _VT_COMMON + ["MODEL.AUTOREGRESSIVE.VT.KERNEL", kernel, "MODEL.AUTOREGRESSIVE.VT.STRIDE", stride,
                         "MODEL.AUTOREGRESSIVE.VT.BLOCKS_E", _BLK(block), "MODEL.AUTOREGRESSIVE.VT.BLOCKS_D", _BLK(block),
                         "INPUT.N_FRAMES_PER_VIDEO_TRAIN", n_train]


PRESETS = {
    "Base-VQVAE": _VQVAE_COMMON + ["MODEL.CODEBOOK.NUM", 1, "MODEL.ENCODER.N_LAYERS", 2, "MODEL.GENERATOR.N_LAYERS", 2,
                                   "SOLVER.MAX_ITER", 500000, "SOLVER.CHECKPOINT_PERIOD", 50000],
    "PR-DVQVAE2": _VQVAE_COMMON + ["MODEL.ENCODER.N_LAYERS", 2, "MODEL.GENERATOR.N_LAYERS", 2, "SOLVER.MAX_ITER", 500000,
                                   "SOLVER.CHECKPOINT_PERIOD", 50000],
    "K-DVQVAE": _VQVAE_COMMON + ["MODEL.ENCODER.N_LAYERS", 4, "MODEL.GENERATOR.N_LAYERS", 4, "SOLVER.MAX_ITER", 1000000,
                                 "SOLVER.CHECKPOINT_PERIOD", 1000000, "INPUT.N_FRAMES_PER_VIDEO_TRAIN", 1],
    "DSFVT": _vt((7, 1, 1), (16, 1, 1), (1, 16, 16), 16),
    "KDSFVT": _vt((7, 1, 1), (16, 1, 1), (1, 16, 16), 16) + ["TEST.VT_SAMPLER.N_PRIME", 5],
    "DSSVT": _vt((1, 3, 3), (1, 2, 2), (4, 8, 8), 4),
    "DSTSVT": _vt((5, 3, 3), (4, 2, 2), (4, 8, 8), 16),
}


def preset(name, overrides=()):
    cfg = get_cfg()
    cfg.merge_from_list(list(PRESETS[name]) + list(overrides))
    return cfg
