/*
 * index_oracle.c -- CPU restatement of james-bowman/nlp's exact nearest-neighbour scan over document vectors
 * (the step after the LDA path).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/ and the cpu_baseline leg of tools/bench_search.py may load it.
 *
 * Follows, in float64 and in the reference's loop order:
 *   measures/pairwise/comparisons.go:17-125   the nine pairwise.Comparer functions
 *   index.go:21-42                           resultHeap (Less = larger distance first) on Go's container/heap
 *   index.go:85-115                          LinearScanIndex.Search
 * Third-party arithmetic the reference calls and this file restates from the published algorithms (sources are not in
 * the reference tree; no go.mod pins them): sparse.Dot / mat.Dot (sum of products in index order), sparse.Norm(v, 2) /
 * mat.Norm (Euclidean norm; gonum scales to avoid overflow, which changes the result by rounding only), mat.Norm(v, 1),
 * and container/heap's Init / Push / Pop / up / down, restated exactly because they decide which of several equally
 * distant vectors survive at the cut.
 * Pinning: the reference's index_test.go holds no numeric golden vectors for this path, only properties
 * (TestIndexerIndex: every indexed vector finds itself at distance 0 +- 1e-7; TestIndexerSearch: Search(v, n) sorted
 * equals the argsort of the pairwise distances; TestIndexerRemove); tests/test_index_oracle.py checks the oracle against
 * those properties and against closed-form comparer values.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

enum {
    CMP_COSINE_SIMILARITY = 0,
    CMP_COSINE_DISTANCE = 1,
    CMP_ANGULAR_DISTANCE = 2,
    CMP_ANGULAR_SIMILARITY = 3,
    CMP_HAMMING_DISTANCE = 4,
    CMP_HAMMING_SIMILARITY = 5,
    CMP_EUCLIDEAN_DISTANCE = 6,
    CMP_MANHATTEN_DISTANCE = 7,
    CMP_VECTOR_LEN_SIMILARITY = 8
};

/* comparisons.go:17-29 */
static double cosine_similarity(int64_t n, const double *a, int64_t sa, const double *b, int64_t sb) {
    double dot = 0, na = 0, nb = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double x = a[i * sa], y = b[i * sb];
        dot += x * y;
        na += x * x;
        nb += y * y;
    }
    na = sqrt(na);
    nb = sqrt(nb);
    if (na == 0 || nb == 0) return NAN;
    return dot / (na * nb);
}

/* comparisons.go:48-55 */
static double angular_distance(int64_t n, const double *a, int64_t sa, const double *b, int64_t sb) {
    double c = cosine_similarity(n, a, sa, b, sb);
    if (c > 1) c = 1.0;
    return acos(c) / M_PI;
}

/* comparisons.go:69-84 (the generic branch; BinaryVec is a different storage of the same count) */
static double hamming_distance(int64_t n, const double *a, int64_t sa, const double *b, int64_t sb) {
    double count = 0;
    for (int64_t i = 0; i < n; ++i)
        if (a[i * sa] != b[i * sb]) count++;
    return count / (double)n;
}

double oracle_compare(int32_t comparer, int64_t n, const double *a, int64_t sa, const double *b, int64_t sb) {
    switch (comparer) {
        case CMP_COSINE_SIMILARITY:
            return cosine_similarity(n, a, sa, b, sb);
        case CMP_COSINE_DISTANCE: /* comparisons.go:40-42 */
            return 1.0 - cosine_similarity(n, a, sa, b, sb);
        case CMP_ANGULAR_DISTANCE:
            return angular_distance(n, a, sa, b, sb);
        case CMP_ANGULAR_SIMILARITY: /* comparisons.go:59-61 */
            return 1.0 - angular_distance(n, a, sa, b, sb);
        case CMP_HAMMING_DISTANCE:
            return hamming_distance(n, a, sa, b, sb);
        case CMP_HAMMING_SIMILARITY: /* comparisons.go:89-91 */
            return 1.0 - hamming_distance(n, a, sa, b, sb);
        case CMP_EUCLIDEAN_DISTANCE: { /* comparisons.go:96-100 */
            double s = 0;
            for (int64_t i = 0; i < n; ++i) {
                const double d = a[i * sa] - b[i * sb];
                s += d * d;
            }
            return sqrt(s);
        }
        case CMP_MANHATTEN_DISTANCE: { /* comparisons.go:104-108 */
            double s = 0;
            for (int64_t i = 0; i < n; ++i) s += fabs(a[i * sa] - b[i * sb]);
            return s;
        }
        case CMP_VECTOR_LEN_SIMILARITY: { /* comparisons.go:111-117 */
            double dot = 0;
            for (int64_t i = 0; i < n; ++i) dot += a[i * sa] * b[i * sb];
            if (dot == 0) return NAN;
            return sqrt(dot);
        }
    }
    return NAN;
}

/* ---- container/heap over resultHeap (index.go:21-42) ------------------------------------------------------------- */
typedef struct {
    double distance;
    int64_t id;
} match_t;

static int less(const match_t *m, int64_t i, int64_t j) { return m[i].distance > m[j].distance; } /* index.go:29 */
static void swap(match_t *m, int64_t i, int64_t j) {
    const match_t t = m[i];
    m[i] = m[j];
    m[j] = t;
}
static void up(match_t *m, int64_t j) {
    for (;;) {
        const int64_t i = (j - 1) / 2; /* parent; Go's integer division truncates toward zero, so j = 0 gives i = 0 */
        if (i == j || !less(m, j, i)) break;
        swap(m, i, j);
        j = i;
    }
}
static void down(match_t *m, int64_t i0, int64_t n) {
    int64_t i = i0;
    for (;;) {
        const int64_t j1 = 2 * i + 1;
        if (j1 >= n || j1 < 0) break;
        int64_t j = j1;
        const int64_t j2 = j1 + 1;
        if (j2 < n && less(m, j2, j1)) j = j2;
        if (!less(m, j, i)) break;
        swap(m, i, j);
        i = j;
    }
}

/* LinearScanIndex.Search (index.go:85-115).  The indexed vectors are the columns of the dim x n row-major matrix m
 * (leading dimension ld), ids[j] their ids; q has stride sq.  Writes up to k matches in the reference's (heap) order and
 * returns how many. */
int64_t oracle_linear_scan_search(int32_t comparer, int64_t dim, int64_t n, const double *m, int64_t ld, const int64_t *ids,
                                  const double *q, int64_t sq, int64_t k, int64_t *ids_out, double *dist_out) {
    if (k < 0) k = 0;
    match_t *h = (match_t *)malloc(sizeof(match_t) * (size_t)(k > 0 ? k : 1));
    int64_t len = 0, point;
    for (point = 0; point < k && point < n; point++) { /* index.go:96-100 */
        h[len].distance = oracle_compare(comparer, dim, q, sq, m + point, ld);
        h[len].id = ids[point];
        len++;
    }
    if (len >= k && k > 0) {                           /* index.go:101-103 returns early when fewer than k */
        for (int64_t i = len / 2 - 1; i >= 0; i--) down(h, i, len); /* heap.Init */
        for (int64_t i = point; i < n; i++) {                       /* index.go:106-112 */
            const double dist = oracle_compare(comparer, dim, q, sq, m + i, ld);
            if (dist <= h[0].distance) {
                swap(h, 0, len - 1); /* heap.Pop */
                down(h, 0, len - 1);
                h[len - 1].distance = dist; /* heap.Push */
                h[len - 1].id = ids[i];
                up(h, len - 1);
            }
        }
    }
    for (int64_t i = 0; i < len; ++i) {
        ids_out[i] = h[i].id;
        dist_out[i] = h[i].distance;
    }
    free(h);
    return len;
}
