// index_kernels.cuh -- sm_100a kernels of the exact nearest-neighbour scan over document vectors (reference: index.go
// LinearScanIndex, measures/pairwise/comparisons.go).  SURVEY.md section 8f rank 4: the step after the LDA path.
//
// Layout in HBM (float64, the reference's arithmetic type):
//   vec   [dim][cap]   vector j is column j -- the K x D doc-topic matrix Transform / FitTransform return is indexed as
//                      it is (ColDo(theta, index.Index)); a thread owns one vector, so every load of a warp is one
//                      contiguous 256-byte segment of a row
//   norm  [cap]        Euclidean norm of every indexed vector (sparse.Norm(v, 2), comparisons.go:21-22)
//   keys  [nq][n]      order-preserving 64-bit image of distance(query q, vector j)
// The scan is HBM-bound: dim * 8 bytes per (vector, batch of QB queries); nothing here is a dense contraction worth a
// tensor core at the batch sizes a search has.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace idxb {

enum Comparer : int {  // measures/pairwise/comparisons.go
    COSINE_SIMILARITY = 0,      // :17-29
    COSINE_DISTANCE = 1,        // :40-42
    ANGULAR_DISTANCE = 2,       // :48-55
    ANGULAR_SIMILARITY = 3,     // :59-61
    HAMMING_DISTANCE = 4,       // :69-84
    HAMMING_SIMILARITY = 5,     // :89-91
    EUCLIDEAN_DISTANCE = 6,     // :96-100
    MANHATTEN_DISTANCE = 7,     // :104-108
    VECTOR_LEN_SIMILARITY = 8,  // :111-117
    N_COMPARERS = 9
};
enum AccClass : int { ACC_DOT = 0, ACC_L2 = 1, ACC_L1 = 2, ACC_HAMMING = 3 };
__host__ __device__ inline int acc_class(int comparer) {
    return comparer == EUCLIDEAN_DISTANCE ? ACC_L2
           : comparer == MANHATTEN_DISTANCE ? ACC_L1
           : (comparer == HAMMING_DISTANCE || comparer == HAMMING_SIMILARITY) ? ACC_HAMMING
                                                                              : ACC_DOT;
}

// distances are ranked through an unsigned key: smaller distance <=> smaller key; NaN (zero vectors under the cosine
// family, comparisons.go:24-26) ranks after +Inf, -0 is folded into +0
__device__ __forceinline__ uint64_t dist_to_key(double d) {
    if (d != d) return 0xFFFFFFFFFFFFFFFFull;
    d += 0.0;
    const uint64_t b = (uint64_t)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_dist(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFull) return __longlong_as_double(0x7FF8000000000000ll);
    const uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// Euclidean norm of the vectors [j0, j0+n) (sum of squares in index order, like the oracle)
__global__ void __launch_bounds__(256) index_norms_kernel(const double *vec, int64_t cap, int dim, int64_t j0, int64_t n, double *norm) {
    for (int64_t j = j0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j0 + n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0;
        for (int k = 0; k < dim; ++k) {
            const double x = vec[(int64_t)k * cap + j];
            s = fma(x, x, s);
        }
        norm[j] = sqrt(s);
    }
}

// queries arrive as the columns of a dim x nq row-major matrix; they are repacked into chunks of QB queries,
// qpack[(chunk*dim + k)*QB + qi] (zero padded), so that a chunk's [dim][QB] block is one contiguous shared-memory
// image; qnorm[q] = Euclidean norm of query q
__global__ void index_pack_queries_kernel(const double *q, int64_t ldq, int dim, int nq, int QB, double *qpack, double *qnorm) {
    const int nchunks = (nq + QB - 1) / QB;
    const int64_t total = (int64_t)nchunks * dim * QB;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int qi = (int)(i % QB);
        const int k = (int)((i / QB) % dim);
        const int chunk = (int)(i / ((int64_t)QB * dim));
        const int qq = chunk * QB + qi;
        qpack[i] = qq < nq ? q[(int64_t)k * ldq + qq] : 0.0;
    }
    for (int qq = blockIdx.x * blockDim.x + threadIdx.x; qq < nq; qq += gridDim.x * blockDim.x) {
        double s = 0;
        for (int k = 0; k < dim; ++k) {
            const double x = q[(int64_t)k * ldq + qq];
            s = fma(x, x, s);
        }
        qnorm[qq] = sqrt(s);
    }
}

struct ScanArgs {
    const double *vec;
    const double *norm;
    const double *qpack;
    const double *qnorm;
    uint64_t *keys;  // [nq][n]
    int64_t cap, n;
    int dim, nq, comparer;
    int ktile;       // rows of the query block staged per shared-memory tile
};

template <int CLS>
__device__ __forceinline__ double acc_step(double acc, double x, double q) {
    if (CLS == ACC_DOT) return fma(x, q, acc);
    if (CLS == ACC_L2) {
        const double d = q - x;
        return fma(d, d, acc);
    }
    if (CLS == ACC_L1) return acc + fabs(q - x);
    return acc + (q != x ? 1.0 : 0.0);
}

__device__ __forceinline__ double finish_distance(int comparer, double acc, double nq_, double nx, int dim) {
    switch (comparer) {
        case COSINE_SIMILARITY:
            return (nq_ == 0 || nx == 0) ? nan("") : acc / (nq_ * nx);
        case COSINE_DISTANCE:
            return (nq_ == 0 || nx == 0) ? nan("") : 1.0 - acc / (nq_ * nx);
        case ANGULAR_DISTANCE:
        case ANGULAR_SIMILARITY: {
            if (nq_ == 0 || nx == 0) return nan("");
            double c = acc / (nq_ * nx);
            if (c > 1) c = 1.0;
            const double a = acos(c) / 3.14159265358979323846;
            return comparer == ANGULAR_DISTANCE ? a : 1.0 - a;
        }
        case HAMMING_DISTANCE:
            return acc / (double)dim;
        case HAMMING_SIMILARITY:
            return 1.0 - acc / (double)dim;
        case EUCLIDEAN_DISTANCE:
            return sqrt(acc);
        case MANHATTEN_DISTANCE:
            return acc;
        default:  // VECTOR_LEN_SIMILARITY
            return acc == 0 ? nan("") : sqrt(acc);
    }
}

// distance(query, vector) for QB queries x all vectors: grid.y = query chunk, a thread owns a vector.  The chunk's
// query values sit in shared memory (every lane reads the same address: a broadcast, two queries per 16-byte load);
// U rows of the vector are loaded before they are used so that U independent 8-byte loads per thread are in flight.
template <int QB, int CLS>
__global__ void __launch_bounds__(256) index_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) double qs[];  // [ktile][QB]
    constexpr int U = 8;
    const int chunk = blockIdx.y;
    const double *qsrc = a.qpack + (int64_t)chunk * a.dim * QB;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool one_tile = a.dim <= a.ktile;  // the whole query block fits: staged once per CTA, not once per vector batch
    if (one_tile) {
        for (int i = threadIdx.x; i < a.dim * QB; i += blockDim.x) qs[i] = qsrc[i];
        __syncthreads();
    }
    // every thread of the block takes part in every tile's staging, so the vector loop is bounded by the block's first index
    for (int64_t jb = (int64_t)blockIdx.x * blockDim.x; jb < a.n; jb += stride) {
        const int64_t j = jb + threadIdx.x;
        const bool live = j < a.n;
        double acc[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) acc[qi] = 0.0;
        for (int k0 = 0; k0 < a.dim; k0 += a.ktile) {
            const int kt = min(a.ktile, a.dim - k0);
            if (!one_tile) {
                __syncthreads();
                for (int i = threadIdx.x; i < kt * QB; i += blockDim.x) qs[i] = qsrc[(int64_t)k0 * QB + i];
                __syncthreads();
            }
            if (live) {
                const double *col = a.vec + (int64_t)k0 * a.cap + j;
                int k = 0;
                for (; k + U <= kt; k += U) {
                    double x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u] = __ldg(col + (int64_t)(k + u) * a.cap);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if constexpr (QB >= 2) {
                            const double2 *row = reinterpret_cast<const double2 *>(qs + (k + u) * QB);
#pragma unroll
                            for (int q2 = 0; q2 < QB / 2; ++q2) {
                                const double2 v = row[q2];
                                acc[2 * q2] = acc_step<CLS>(acc[2 * q2], x[u], v.x);
                                acc[2 * q2 + 1] = acc_step<CLS>(acc[2 * q2 + 1], x[u], v.y);
                            }
                        } else {
                            acc[0] = acc_step<CLS>(acc[0], x[u], qs[k + u]);
                        }
                    }
                }
                for (; k < kt; ++k) {
                    const double x = __ldg(col + (int64_t)k * a.cap);
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) acc[qi] = acc_step<CLS>(acc[qi], x, qs[k * QB + qi]);
                }
            }
        }
        if (live) {
            const double nx = a.norm[j];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const int qq = chunk * QB + qi;
                if (qq < a.nq) a.keys[(int64_t)qq * a.n + j] = dist_to_key(finish_distance(a.comparer, acc[qi], a.qnorm[qq], nx, a.dim));
            }
        }
    }
}

// ---- exact top-k by radix selection ----------------------------------------------------------------------------
// The k nearest vectors of a query are the k smallest 96-bit composites (key, ~position): equal distances rank the
// LATER-indexed vector first, which is what the reference's `dist <= heap top` replacement keeps (index.go:108).
// Composites are unique, so the k-th smallest composite T is a threshold that exactly k vectors pass.  T is found digit
// by digit, most significant first (6 digits of the key, 3 of the position), one histogram pass per digit.
constexpr int SEL_BINS = 2048;
constexpr int SEL_PASSES = 9;
__host__ __device__ inline void sel_digit(int pass, int *word, int *shift, int *bits) {
    if (pass < 6) {  // key: digits of 11,11,11,11,11,9 bits at shifts 53,42,31,20,9,0
        *word = 0;
        *shift = pass < 5 ? 53 - 11 * pass : 0;
        *bits = pass < 5 ? 11 : 9;
    } else {         // ~position: digits of 11,11,10 bits at shifts 21,10,0
        *word = 1;
        *shift = pass == 6 ? 21 : pass == 7 ? 10 : 0;
        *bits = pass == 8 ? 10 : 11;
    }
}

struct SelState {
    uint64_t t_key;   // threshold composite, built up digit by digit
    uint32_t t_inv;
    uint32_t k_rem;   // rank of the wanted element among those that share the digits fixed so far (1-based)
    int done;         // the threshold is final (all remaining candidates are taken)
    int pad;
};

__global__ void sel_init_kernel(SelState *st, int nq, uint32_t k, int take_all) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) {
        st[q].t_key = take_all ? 0xFFFFFFFFFFFFFFFFull : 0ull;
        st[q].t_inv = take_all ? 0xFFFFFFFFu : 0u;
        st[q].k_rem = k;
        st[q].done = take_all;
        st[q].pad = 0;
    }
}

// does (key, inv) agree with the threshold on every digit above this pass's digit?
__device__ __forceinline__ bool sel_matches(uint64_t key, uint32_t inv, const SelState &s, int word, int shift, int bits) {
    if (word == 0) {
        const int hi = shift + bits;
        return hi >= 64 ? true : (key >> hi) == (s.t_key >> hi);
    }
    const int hi = shift + bits;
    return key == s.t_key && (hi >= 32 ? true : (inv >> hi) == (s.t_inv >> hi));
}

__global__ void __launch_bounds__(256) sel_hist_kernel(const uint64_t *keys, int64_t n, const SelState *st, int pass, unsigned *hist) {
    __shared__ unsigned sh[SEL_BINS];
    const int q = blockIdx.y;
    const SelState s = st[q];
    if (s.done) return;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    int word, shift, bits;
    sel_digit(pass, &word, &shift, &bits);
    const uint64_t *kq = keys + (int64_t)q * n;
    const unsigned mask = (1u << bits) - 1u;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = kq[j];
        const uint32_t inv = ~(uint32_t)j;
        if (sel_matches(key, inv, s, word, shift, bits)) {
            const unsigned d = word == 0 ? (unsigned)(key >> shift) & mask : (inv >> shift) & mask;
            atomicAdd(&sh[d], 1u);
        }
    }
    __syncthreads();
    unsigned *hq = hist + (int64_t)q * SEL_BINS;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hq[i], sh[i]);
}

// one block per query: the digit whose cumulative count reaches k_rem becomes the next digit of the threshold
__global__ void __launch_bounds__(256) sel_pick_kernel(SelState *st, int pass, unsigned *hist) {
    __shared__ unsigned part[256];
    __shared__ int sel_t;
    __shared__ unsigned sel_before;
    const int q = blockIdx.x;
    SelState s = st[q];
    unsigned *hq = hist + (int64_t)q * SEL_BINS;
    if (s.done) return;
    constexpr int PER = SEL_BINS / 256;
    unsigned mine[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        mine[i] = hq[threadIdx.x * PER + i];
        sum += mine[i];
        hq[threadIdx.x * PER + i] = 0;  // ready for the next pass
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned cum = 0;
        int t = 0;
        for (; t < 255; ++t) {
            if (cum + part[t] >= s.k_rem) break;
            cum += part[t];
        }
        sel_t = t;
        sel_before = cum;
    }
    __syncthreads();
    if ((int)threadIdx.x == sel_t) {
        unsigned cum = sel_before;
        int i = 0;
        for (; i < PER - 1; ++i) {
            if (cum + mine[i] >= s.k_rem) break;
            cum += mine[i];
        }
        const unsigned d = (unsigned)(threadIdx.x * PER + i);
        int word, shift, bits;
        sel_digit(pass, &word, &shift, &bits);
        if (word == 0)
            s.t_key |= (uint64_t)d << shift;
        else
            s.t_inv |= d << shift;
        s.k_rem -= cum;
        if (mine[i] == s.k_rem || pass == SEL_PASSES - 1) {
            // every candidate that shares this digit is wanted: the remaining digits of the threshold are all ones
            if (word == 0) {
                s.t_key |= shift ? ((1ull << shift) - 1ull) : 0ull;
                s.t_inv = 0xFFFFFFFFu;
            } else {
                s.t_inv |= shift ? ((1u << shift) - 1u) : 0u;
            }
            s.done = 1;
        }
        st[q] = s;
    }
}

struct Cand {
    uint64_t key;
    uint32_t inv;  // ~position
    uint32_t pad;
};
__device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b) { return a.key < b.key || (a.key == b.key && a.inv < b.inv); }

// every vector whose composite is <= the threshold goes into the query's candidate list (exactly min(k, n) of them)
__global__ void __launch_bounds__(256) sel_collect_kernel(const uint64_t *keys, int64_t n, const SelState *st, Cand *cand, int P,
                                                          unsigned *count) {
    const int q = blockIdx.y;
    const SelState s = st[q];
    const uint64_t *kq = keys + (int64_t)q * n;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = kq[j];
        const uint32_t inv = ~(uint32_t)j;
        if (key < s.t_key || (key == s.t_key && inv <= s.t_inv)) {
            const unsigned slot = atomicAdd(&count[q], 1u);
            if (slot < (unsigned)P) {
                Cand c;
                c.key = key;
                c.inv = inv;
                c.pad = 0;
                cand[(int64_t)q * P + slot] = c;
            }
        }
    }
}

// one block per query: bitonic sort of the P (power of two) candidate slots -- unused slots hold the largest composite --
// then the first `found` become the result: nearest first, later-indexed first among equal distances
__global__ void __launch_bounds__(1024) sel_sort_emit_kernel(Cand *cand, int P, const unsigned *count, int k, const int64_t *ids,
                                                             int64_t *ids_out, double *dist_out, int32_t *found_out) {
    const int q = blockIdx.x;
    Cand *c = cand + (int64_t)q * P;
    const int have = min((int)count[q], k);
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        if (i >= have) {
            c[i].key = 0xFFFFFFFFFFFFFFFFull;
            c[i].inv = 0xFFFFFFFFu;
            c[i].pad = 1;
        }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int str = size >> 1; str > 0; str >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int p = i ^ str;
                if (p > i) {
                    const bool up = (i & size) == 0;
                    const Cand x = c[i], y = c[p];
                    if (cand_less(y, x) == up) {
                        c[i] = y;
                        c[p] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        if (i < have) {
            ids_out[(int64_t)q * k + i] = ids[(uint32_t)~c[i].inv];
            dist_out[(int64_t)q * k + i] = key_to_dist(c[i].key);
        } else {
            ids_out[(int64_t)q * k + i] = -1;
            dist_out[(int64_t)q * k + i] = __longlong_as_double(0x7FF8000000000000ll);
        }
    }
    if (threadIdx.x == 0) found_out[q] = have;
}

}  // namespace idxb
