"""nlp_b200 -- B200-native LDA SCVB0 path of james-bowman/nlp behind the reference's own API names.

Only the hot path is here: `LatentDirichletAllocation` with `Fit` / `FitTransform` /
`Transform` / `Perplexity` / `Components`, computed by hand-written sm_100a kernels in liblda_b200.so.
"""
from .lda import (CSC, CSR, LatentDirichletAllocation, LearningSchedule,  # noqa: F401
                  NewLatentDirichletAllocation)
from . import rand  # noqa: F401
