"""src/config.py:7-28 -- the same constants.  Change them BEFORE the first kernel call."""
image_resolution = (1920 * 4 // 10, 1080 * 4 // 10)

SAMPLES_PER_FRAME = 1
SAMPLES_PER_PIXEL = 1  # number of samples in one draw call
QUALITY_PER_SAMPLE = 0.8  # for russian roulette

BLACK_BACKGROUND = False
ADAPTIVE_SAMPLING = False

VISIBILITY = (1e-4, 1e4)
NOISE_THRESHOLD = 1e-4  # for self-adaptive sampling

MAX_RAYMARCH = 512
MAX_RAYTRACE = 512

ENV_IOR = 1.000277

SEED = 0          # Philox key (the reference sets no seed)
DEVICE = 0
