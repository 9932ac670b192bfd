"""Drop-in mirror of the reference module
wanx/train/special_attentions_local/TrainRelated/wanx_blocksparseattn.py (W) -- same module-level knobs,
same function / class names and call signatures, B200 kernels underneath (no PyTorch fallback, forward only).
The functions and classes live in `_blocksparse_common.build_api` (shared with the CogVideoX mirror); this file
holds what differs: the knobs, read at call time exactly like the reference's module globals.

ESTIMATOR DEFAULT.  `estimator = "meanpool"` selects blocks from the softmax of block-mean scores (the mask
generation kernel BASELINE.json's north_star specifies).  The reference selects from the sampled max-pooled map of
its Triton kernel (W:62-87 -> P); set `estimator = "sampled_max"` to reproduce that (same kernels otherwise, one
C-ABI call either way).  With the default the block mask -- hence the output -- is NOT the reference model's.
"""
from __future__ import annotations

from ._blocksparse_common import build_api

# ----------------------------- parameters (W:9-16) -----------------------------
use_rearrange = True
max_retain_ratio = 0.17
min_retain_ratio = 0.05
width = 52
height = 30
depth = 21
sample_gap = 30
text_length = 0
# ---------------------- literals of the reference (W:62,325,341) ----------------
block_size = 128
num_keep = 32
energy_threshold = 0.95
estimator = "meanpool"   # north-star kernel (a); the reference's sampled-max estimator is "sampled_max"
exact_merge = True
_FLAVOR = "wan"

globals().update(build_api(__name__))
