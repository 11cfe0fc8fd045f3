// index_kernels.cuh -- sm_100a kernels of the exact nearest-neighbour scan over document vectors (reference: index.go
// LinearScanIndex, measures/pairwise/comparisons.go).  SURVEY.md section 8f rank 4: the step after the LDA path.
//
// Layout in HBM (float64, the reference's arithmetic type):
//   vec   [dim][cap]   vector j is column j -- the K x D doc-topic matrix Transform / FitTransform return is indexed as
//                      it is (ColDo(theta, index.Index)); a thread owns one vector, so every load of a warp is one
//                      contiguous 256-byte segment of a row
//   inorm [cap]        1 / Euclidean norm of every indexed vector (sparse.Norm(v, 2), comparisons.go:21-22); +Inf for an
//                      all-zero vector, which turns its cosine into 0 * Inf = NaN exactly where the reference returns NaN
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

// 1 / Euclidean norm of the vectors [j0, j0+n) (sum of squares in index order, like the oracle)
__global__ void __launch_bounds__(256) index_norms_kernel(const double *vec, int64_t cap, int dim, int64_t j0, int64_t n, double *inorm) {
    for (int64_t j = j0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j0 + n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0;
        for (int k = 0; k < dim; ++k) {
            const double x = vec[(int64_t)k * cap + j];
            s = fma(x, x, s);
        }
        inorm[j] = 1.0 / sqrt(s);
    }
}

// queries arrive as the columns of a dim x nq row-major matrix; they are repacked into chunks of QB queries,
// qpack[(chunk*dim + k)*QB + qi] (zero padded), so that a chunk's [dim][QB] block is one contiguous shared-memory
// image; qinorm[q] = 1 / Euclidean norm of query q
__global__ void index_pack_queries_kernel(const double *q, int64_t ldq, int dim, int nq, int QB, double *qpack, double *qinorm) {
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
        qinorm[qq] = 1.0 / sqrt(s);
    }
}

struct ScanArgs {
    const double *vec;
    const double *inorm;   // 1 / norm of the indexed vectors
    const double *qpack;
    const double *qinorm;  // 1 / norm of the queries
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

// comparer value from the accumulated sum.  The cosine family divides by the norms through their stored reciprocals
// (comparisons.go:28 computes dot / (norma * normb); the product of reciprocals differs by rounding only) and a zero
// norm yields NaN as in comparisons.go:24-26.  The cheap finishes are inlined -- a thread finishes up to 64 values per
// vector tile -- the rest share one out-of-line copy.
__device__ __noinline__ double finish_slow(int comparer, double acc, double iq, double ix, int dim) {
    switch (comparer) {
        case ANGULAR_DISTANCE:
        case ANGULAR_SIMILARITY: {
            double c = acc * iq * ix;  // NaN for a zero vector
            if (c > 1) c = 1.0;
            const double a = acos(c) / 3.14159265358979323846;
            return comparer == ANGULAR_DISTANCE ? a : 1.0 - a;
        }
        case HAMMING_DISTANCE:
            return acc / (double)dim;
        case HAMMING_SIMILARITY:
            return 1.0 - acc / (double)dim;
        default:  // VECTOR_LEN_SIMILARITY
            return acc == 0 ? nan("") : sqrt(acc);
    }
}
__device__ __forceinline__ double finish_distance(int comparer, double acc, double iq, double ix, int dim) {
    if (comparer == COSINE_DISTANCE) return 1.0 - acc * iq * ix;
    if (comparer == COSINE_SIMILARITY) return acc * iq * ix;
    if (comparer == EUCLIDEAN_DISTANCE) return sqrt(acc);
    if (comparer == MANHATTEN_DISTANCE) return acc;
    return finish_slow(comparer, acc, iq, ix, dim);
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
            const double nx = a.inorm[j];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const int qq = chunk * QB + qi;
                if (qq < a.nq) a.keys[(int64_t)qq * a.n + j] = dist_to_key(finish_distance(a.comparer, acc[qi], a.qinorm[qq], nx, a.dim));
            }
        }
    }
}

// ---- the batched scan for the dot-product comparers: FP64 tensor cores fed by TMA bulk copies -----------------------
// Many queries against all vectors IS a dense contraction, keys = f(Q^T [nq x dim] . X [dim x n]), and the kernel above
// shows what happens without tensor cores: at 16 queries per pass it is bound by the FP64 pipe (measured 5.5e12
// multiply-adds/s whatever the blocking -- 1, 2 or 4 vectors per thread, vectors staged through shared memory or not),
// 0.43 of the HBM roofline.  tcgen05 has no FP64 kind, so the contraction uses mma.sync.m8n8k4.f64 (DMMA):
//   A = 8 queries x 4 dims (row major, from the query block in shared memory)
//   B = 4 dims x 8 vectors (column major == the [dim][n] layout of the index, from the staged tile)
//   D = 8 x 8 comparer numerators, two per thread, accumulated over dim in registers.
// A warp owns 32 vectors x all QB = 8*MT queries of the pass; a persistent CTA of 8 warps per SM walks the
// (vector tile, row tile) pairs in order while one elected thread keeps a STAGES-deep ring of [KT rows][256 vectors]
// tiles in flight with cp.async.bulk (one 2 KB copy per row, completion on an mbarrier per stage).  Row pitches are
// padded by 4 doubles so that both fragment reads touch all 32 banks exactly once per half-warp.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int MMA_TV = 256;            // vectors per CTA tile (8 warps x 32); the index capacity is a multiple of it
constexpr int MMA_KT = 12;             // rows per stage (a multiple of 4)
constexpr int MMA_STAGES = 3;
constexpr int MMA_PITCH = MMA_TV + 4;  // doubles
__host__ __device__ constexpr int mma_q_pitch(int dim) { return (dim + 15) / 16 * 16 + 4; }
__host__ __device__ constexpr size_t scan_mma_smem_bytes(int dim, int QB) {
    return (size_t)MMA_STAGES * MMA_KT * MMA_PITCH * 8 + (size_t)QB * mma_q_pitch(dim) * 8 + MMA_STAGES * 8;
}

template <int MT>  // QB = 8*MT queries per pass; dot-product comparers only; dim % 4 == 0
__global__ void __launch_bounds__(256, 1) index_scan_mma_kernel(const ScanArgs a) {
    constexpr int QB = 8 * MT;
    extern __shared__ __align__(128) unsigned char scan_smem[];
    double *xs = reinterpret_cast<double *>(scan_smem);                    // [STAGES][KT][PITCH]
    double *qs = xs + MMA_STAGES * MMA_KT * MMA_PITCH;                     // [QB][PQ]
    const int PQ = mma_q_pitch(a.dim);
    uint64_t *full = reinterpret_cast<uint64_t *>(qs + (size_t)QB * PQ);   // [STAGES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int chunk = blockIdx.y;
    const int64_t nvt = (a.n + MMA_TV - 1) / MMA_TV;
    const int nkt = (a.dim + MMA_KT - 1) / MMA_KT;
    const int64_t my_tiles = (int64_t)blockIdx.x < nvt ? (nvt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * nkt;  // pipeline steps of this CTA: (vector tile, row tile) pairs in order

    if (tid == 0) {
        for (int s = 0; s < MMA_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const double *qsrc = a.qpack + (int64_t)chunk * a.dim * QB;  // [dim][QB]
    for (int i = tid; i < a.dim * QB; i += 256) qs[(i % QB) * PQ + i / QB] = qsrc[i];
    __syncthreads();

    // warp 0 issues a stage: lane r copies row r, lane 0 arms the barrier with the byte count of the whole stage
    auto issue = [&](int64_t s) {
        const int64_t vt = blockIdx.x + (s / nkt) * gridDim.x;
        const int k0 = (int)(s % nkt) * MMA_KT;
        const int rows = min(MMA_KT, a.dim - k0);
        const int stage = (int)(s % MMA_STAGES);
        if (lane == 0) mbar_expect_tx(&full[stage], (unsigned)rows * MMA_TV * 8u);
        __syncwarp();
        const double *src = a.vec + (int64_t)k0 * a.cap + vt * MMA_TV;
        double *dst = xs + (size_t)stage * MMA_KT * MMA_PITCH;
        if (lane < rows) bulk_g2s(dst + lane * MMA_PITCH, src + (int64_t)lane * a.cap, MMA_TV * 8u, &full[stage]);
    };
    if (warp == 0)
        for (int64_t s = 0; s < MMA_STAGES && s < total; ++s) issue(s);

    double acc[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const double *qfrag = qs + (size_t)g * PQ + t;             // + mt*8*PQ + k
    for (int64_t s = 0; s < total; ++s) {
        const int stage = (int)(s % MMA_STAGES);
        const unsigned parity = (unsigned)((s / MMA_STAGES) & 1);
        while (!mbar_try_wait(&full[stage], parity)) {
        }
        const int kt = (int)(s % nkt);
        const int k0 = kt * MMA_KT;
        const int rows = min(MMA_KT, a.dim - k0);
        const double *xfrag = xs + (size_t)stage * MMA_KT * MMA_PITCH + t * MMA_PITCH + warp * 32 + g;  // + kk*PITCH + nt*8
#pragma unroll
        for (int kk = 0; kk < MMA_KT; kk += 4) {
            if (kk < rows) {
                double bf[4], af[MT];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) bf[nt] = xfrag[kk * MMA_PITCH + nt * 8];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) af[mt] = qfrag[(size_t)mt * 8 * PQ + k0 + kk];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
        }
        if (kt == nkt - 1) {  // this vector tile is complete
            const int64_t jw = (blockIdx.x + (s / nkt) * gridDim.x) * MMA_TV + warp * 32 + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int64_t j = jw + nt * 8 + i;
                    if (j < a.n) {
                        const double nx = a.inorm[j];
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const int qq = chunk * QB + mt * 8 + g;
                            if (qq < a.nq)
                                a.keys[(int64_t)qq * a.n + j] = dist_to_key(finish_distance(a.comparer, acc[mt][nt][i], a.qinorm[qq], nx, a.dim));
                        }
                    }
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        }
        __syncthreads();  // every warp has read this stage: it may be refilled
        if (warp == 0 && s + MMA_STAGES < total) issue(s + MMA_STAGES);
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
