"""Constants of the reference's `configuration.py` that the colour path reads
(configuration.py:4, :24-32).  Names are kept so `from configuration import *` call sites port over."""
SEED = 47
BATCH_SIZE = 4
IMG_SIZE = 64
INPUT_CHANNELS = 4
OUTPUT_CHANNELS = 4
MAX_PALETTE_SIZE = 256
INVALID_INDEX_COLOR = [255, 0, 220, 255]
