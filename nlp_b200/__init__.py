"""nlp_b200 -- B200-native LDA SCVB0 path of james-bowman/nlp behind the reference's own API names.

Only the hot path is here: `LatentDirichletAllocation` with `Fit` / `FitTransform` /
`Transform` / `Perplexity` / `Components`, computed by hand-written sm_100a kernels in liblda_b200.so, and -- the step
after the path (section 8f) -- `LinearScanIndex` with the `pairwise` comparers over the resulting document vectors.
"""
from .lda import (CSC, CSR, LatentDirichletAllocation, LearningSchedule,  # noqa: F401
                  NewLatentDirichletAllocation)
from . import rand  # noqa: F401
from .index import LinearScanIndex, Match, NewLinearScanIndex, SearchSharded, merge_matches  # noqa: F401
from .measures import pairwise  # noqa: F401
