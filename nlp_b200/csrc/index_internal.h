// index_internal.h -- hand-over between the LDA engine and the index inside liblda_b200.so (not part of the C ABI):
// lets Transform write its normalised, transposed doc-topic vectors straight into the index's [dim][cap] matrix in HBM,
// so that `ColDo(lda.Transform(m), index.Index)` moves no vector through the host.
#pragma once
#include <stdint.h>

struct lda_b200_index;

// Reserves room for n more vectors of `dim` elements on `device` and returns where column Len() starts (*dst) and the
// row pitch in elements (*ld).  Everything the index had queued is complete on return; nothing is visible to Search
// until index_finish_append.  Returns an lda_b200_status code (text via lda_b200_index_last_error).
int index_begin_append(lda_b200_index *ix, int device, int dim, int64_t n, double **dst, int64_t *ld);
// Publishes the n vectors written behind index_begin_append's pointer under ids[0..n) (host array).  The writer must have
// completed (stream-synchronised) before the call.
int index_finish_append(lda_b200_index *ix, int64_t n, const int64_t *ids);
