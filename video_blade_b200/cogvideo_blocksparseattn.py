"""Drop-in mirror of the reference module
cogvideox/train/special_attentions_local/TrainRelated/cogvideo_blocksparseattn.py (C) -- same knobs and names
as the reference, B200 kernels underneath (forward only).  Differences from the wan flavour, as in the reference:
text tokens (first `text_length` rows) move to the sequence tail inside the Gilbert reorder (C:141-161),
retain bounds come from an fp32 tensor multiply (C:230-231), the last two block rows/cols are forced on
(C:247-248), defaults C:9-16.  Implementation shared with the Wan mirror: `_blocksparse_common.build_api`
(function map and the note on the estimator default are in wanx_blocksparseattn.py).
"""
from __future__ import annotations

from ._blocksparse_common import build_api

# ----------------------------- parameters (C:9-16) -----------------------------
use_rearrange = True
max_retain_ratio = 0.1
min_retain_ratio = 0.05
width = 45
height = 30
depth = 13
sample_gap = 15
text_length = 226
# ---------------------- literals of the reference (C:62,341,357) ----------------
block_size = 128
num_keep = 32
energy_threshold = 0.95
estimator = "meanpool"
exact_merge = True
_FLAVOR = "cog"

globals().update(build_api(__name__))
