/*
 * lda_b200_index.h -- C ABI of the exact nearest-neighbour scan over document vectors in liblda_b200.so:
 * the step after the LDA path, i.e. james-bowman/nlp's `LinearScanIndex`
 * (index.go:60-135) with the `pairwise.Comparer` functions of measures/pairwise/comparisons.go, on one B200.
 *
 * Like lda_b200.h: the reference has no FFI boundary, so each entry point cites the Go method it replaces; every
 * function returns 0 or an lda_b200_status code (lda_b200.h), text via lda_b200_index_last_error; all pointers are host
 * pointers owned by the caller and nothing is retained after a call returns; no CPU fallback.
 *
 * Vectors are float64 (the reference's type) and all have the dimension of the first one added.  A *batch* of vectors
 * is passed as the columns of a dim x n row-major matrix with leading dimension ld >= n -- exactly the K x D doc-topic
 * matrix Transform / FitTransform return (lda.go:452,541), so `ColDo(theta, index.Index)` is one call; a single
 * mat.Vector is the batch n = 1, ld = its increment.  ids are int64 (the Go shim maps `interface{}` ids to them).
 */
#ifndef LDA_B200_INDEX_H
#define LDA_B200_INDEX_H

#include <stdint.h>

#include "lda_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* pairwise.Comparer, measures/pairwise/comparisons.go */
typedef enum {
    LDA_B200_COSINE_SIMILARITY = 0,     /* comparisons.go:17-29  (NaN when a vector is all zeros) */
    LDA_B200_COSINE_DISTANCE = 1,       /* comparisons.go:40-42  */
    LDA_B200_ANGULAR_DISTANCE = 2,      /* comparisons.go:48-55  */
    LDA_B200_ANGULAR_SIMILARITY = 3,    /* comparisons.go:59-61  */
    LDA_B200_HAMMING_DISTANCE = 4,      /* comparisons.go:69-84  (proportion of unequal elements) */
    LDA_B200_HAMMING_SIMILARITY = 5,    /* comparisons.go:89-91  */
    LDA_B200_EUCLIDEAN_DISTANCE = 6,    /* comparisons.go:96-100 */
    LDA_B200_MANHATTEN_DISTANCE = 7,    /* comparisons.go:104-108 */
    LDA_B200_VECTOR_LEN_SIMILARITY = 8  /* comparisons.go:111-117 */
} lda_b200_comparer;

typedef struct lda_b200_index lda_b200_index;

/* NewLinearScanIndex(compareFN), index.go:74-76, on CUDA device `device` */
int lda_b200_index_create(int32_t device, int32_t comparer, lda_b200_index **out);
void lda_b200_index_destroy(lda_b200_index *ix);
/* error text of the last failed call on ix (ix == NULL: last failed create on this thread) */
const char *lda_b200_index_last_error(const lda_b200_index *ix);

/* Index(v, id), index.go:79-84, for n vectors: column j of m gets ids[j] (ids == NULL: the vector's position, i.e.
 * consecutive numbers continuing from Len()). */
int lda_b200_index_add(lda_b200_index *ix, int32_t dim, int64_t n, const double *m, int64_t ld, const int64_t *ids);

/* `ColDo(theta, func(j, v) { index.Index(v, ids[j]) })` with theta = lda.Transform(m) (lda.go:446-453), fused: the
 * Transform pipeline writes the doc-topic vectors of m's documents straight into the index's matrix in HBM, so no vector
 * crosses the host.  Arguments as lda_b200_transform_matrix; document j of m is indexed under ids[j], or under j when
 * ids is NULL.  h and ix must live on the same device.  With a communicator every rank indexes the documents of its own
 * minibatches (a sharded index: search every rank's index and merge). */
int lda_b200_transform_into_index(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                                  double rho_theta_t, lda_b200_index *ix, const int64_t *ids);
/* the same call with the matrix descriptor spread over scalar arguments (the form cgo can call, see lda_b200_fit_flat) */
int lda_b200_transform_into_index_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                                       const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                                       double rho_theta_t, lda_b200_index *ix, const int64_t *ids);

/* Remove(id), index.go:117-135: drops the first vector indexed under id, keeping the order of the others; unknown ids
 * are ignored. */
int lda_b200_index_remove(lda_b200_index *ix, int64_t id);

/* number of indexed vectors */
int64_t lda_b200_index_len(const lda_b200_index *ix);

/* Search(qv, k), index.go:85-115, for nq query vectors (the columns of q, dim x nq row-major, leading dimension ldq):
 * for query i, found_out[i] = min(k, Len()) matches are written to ids_out[i*k ...] / dist_out[i*k ...], nearest
 * (smallest comparer value) first; unused slots get id -1 and NaN.  The reference returns the same matches in heap
 * order ("unsorted", index.go:87).  Among vectors at exactly the same distance the later-indexed one ranks first, which
 * is the one the reference's `dist <= heap top` test keeps at the cut (index.go:108).  A NaN comparer value (all-zero
 * vector under the cosine family) ranks after every number, where the reference's heap would treat it inconsistently. */
int lda_b200_index_search(lda_b200_index *ix, int64_t nq, const double *q, int64_t ldq, int32_t k, int64_t *ids_out,
                          double *dist_out, int32_t *found_out);

/* device time (ms, CUDA events on the index's stream) the last lda_b200_index_search spent in the distance scan and in
 * the top-k selection, the number of kernels it launched and how many times the scan read the whole index (one pass
 * serves up to 64 queries under the dot-product comparers, else up to 16) -- tools/bench_search.py */
int lda_b200_index_last_search_time(const lda_b200_index *ix, float *scan_ms, float *select_ms, int32_t *launches,
                                    int32_t *passes);

#ifdef __cplusplus
}
#endif
#endif /* LDA_B200_INDEX_H */
