"""Default config tree: the keys and default values of the reference (vidgen/config/defaults.py:1-170),
written as one nested mapping.  Only keys matter for YAML compatibility; values that the hot path reads
are cited where they are consumed."""


def defaults():
    vt = dict(NC=0, NV=0, KERNEL=(), STRIDE=(), D=0, DA=0, DE=0, BLOCKS_E=(), N_HEAD_E=(), BLOCKS_D=(), N_HEAD_D=(),
              N_PRIME=0, PAD_VALUE=-1, SHARE_P=True, SHARE_EMBEDDINGS=False, CLASS_NUM=0)
    enc = dict(WEIGHTS="", NAME="", IN_CHANNELS=1, NF=16, RES_CHANNELS=0, OUT_CHANNELS=16, NORM="", N_LAYERS=0,
               SPECTRAL=False, OUT_ACTIVATION="")
    gen = dict(WEIGHTS="", NAME="", IN_CHANNELS=16, NF=16, RES_CHANNELS=0, OUT_CHANNELS=3, NORM="", N_LAYERS=0,
               SPECTRAL=False, OUT_ACTIVATION="")
    model = dict(DEVICE="cuda", META_ARCHITECTURE="ACAIModel", PIXEL_MEAN=[0.], PIXEL_STD=[1.], IGNORE_INDEX=-100,
                 INIT_TYPE="normal", INIT_VARIANCE=0.02,
                 AUTOREGRESSIVE=dict(NAME="", VT=vt), ENCODER=enc, GENERATOR=gen,
                 CODEBOOK=dict(NUM=1, SIZE=512, DIM=256, WEIGHTS="", EMA=False, BETA=1.0))
    solver = dict(MAX_ITER=40000, SUPERVISED_MAX_ITER=-1, LR_SCHEDULER_NAME="Identity", GAMMA=0.1, STEPS=(),
                  WARMUP_ITERS=-1, WARMUP_FACTOR=0.01, WARMUP_METHOD="linear", OPTIMIZER_NAME="adam", LR_G=0.0001,
                  LR_D=0.0004,
                  WEIGHT_DECAY=dict(BASE_G=0.0, BIAS_G=0.0, NORM_G=0.0, BASE_D=0.0, BIAS_D=0.0, NORM_D=0.0),
                  ADAM=dict(BETA1_G=0.9, BETA2_G=0.9, BETA1_D=0.9, BETA2_D=0.999),
                  RMSPROP=dict(ALPHA_G=0.99, ALPHA_D=0.99, MOMENTUM_G=0.0, MOMENTUM_D=0.0),
                  ACCUMULATION_STEPS=1, CHECKPOINT_PERIOD=50000, IMS_PER_BATCH=32, D_UPDATE_RATIO=1, D_INIT_ITERS=-1,
                  MAXUP=False)
    return dict(
        MODEL=model,
        INPUT=dict(FORMAT="L", N_FRAMES_PER_VIDEO_TRAIN=-1, N_FRAMES_PER_VIDEO_TEST=-1, SCALE_TO_ZEROONE=True,
                   PREPARE_SLICES_TRAIN=False),
        GAN_MODE_ON=False,
        DATASETS=dict(TRAIN=(), TEST=()),
        DATALOADER=dict(NUM_WORKERS=4, SAMPLER_TRAIN="TrainingSampler"),
        SOLVER=solver,
        LOSS=dict(PIXEL=dict(ONN=False, LAMBDA=1.0, MODE="l2"),
                  GAN=dict(ONN=False, LAMBDA_G=1.0, LAMBDA_D=1.0, REAL_LABEL=1.0, FAKE_LABEL=0.0, MODE="wgan")),
        TEST=dict(EXPECTED_RESULTS=[], EVAL_PERIOD=0, N_SAMPLES=0, EVALUATORS="",
                  VT_SAMPLER=dict(VQ_VAE=dict(CFG="", ENCODER_WEIGHTS="", GENERATOR_WEIGHTS="", CODEBOOK_WEIGHTS=""),
                                  N_PRIME=5, NUM_SAMPLES=10)),
        OUTPUT_DIR="./output", SEED=-1, CUDNN_BENCHMARK=True, VIS_PERIOD=100000000000, VERSION=1,
        GLOBAL=dict(HACK=1.0),
    )
