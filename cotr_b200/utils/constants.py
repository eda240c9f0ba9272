"""Constants of the hot path (reference: COTR/utils/constants.py)."""
DEFAULT_PRECISION = 'float32'
MAX_SIZE = 256          # side of each network input image; the canvas is MAX_SIZE x 2*MAX_SIZE
VALID_NN_OVERLAPPING_THRESH = 0.1
