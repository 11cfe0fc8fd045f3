// index_b200.cu -- engine + C ABI (include/lda_b200_index.h) of the exact nearest-neighbour scan over document vectors.
// Reference behaviour: james-bowman/nlp index.go (LinearScanIndex) and measures/pairwise/comparisons.go, cited per
// function.  No CPU fallback: distances, the top-k selection and the ordering of the result are kernels from
// index_kernels.cuh on the index's stream; the host only marshals arguments and keeps the position -> id list that
// Remove needs.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/lda_b200.h"
#include "../../include/lda_b200_index.h"
#include "index_internal.h"
#include "index_kernels.cuh"

using namespace idxb;

namespace {
thread_local std::string g_index_create_error;
}

struct lda_b200_index {
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    int comparer = 0;

    // the indexed vectors (index.go:62-64: signatures, ids)
    int dim = 0;
    int64_t n = 0, cap = 0;
    double *vec = nullptr;   // [dim][cap]
    double *norm = nullptr;  // [cap] reciprocal Euclidean norms
    int64_t *ids = nullptr;  // [cap]
    std::vector<int64_t> h_ids;

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float scan_ms = 0, select_ms = 0;
    int launches = 0, passes = 0;
};

namespace {

int ifail(lda_b200_index *ix, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ix)
        ix->err = buf;
    else
        g_index_create_error = buf;
    return code;
}

#define ICU(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess)                                                                                           \
            return ifail(ix, LDA_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define ITRY(call)           \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)
#define IKCHECK()                   \
    do {                            \
        ix->launches++;             \
        ICU(cudaPeekAtLastError()); \
    } while (0)

template <typename T>
int ialloc(lda_b200_index *ix, T **p, size_t n) {
    *p = nullptr;
    ICU(cudaMallocAsync((void **)p, std::max<size_t>(n, 1) * sizeof(T), ix->stream));
    return 0;
}
template <typename T>
void ifree(lda_b200_index *ix, T **p) {
    if (*p) cudaFreeAsync(*p, ix->stream);
    *p = nullptr;
}

inline int grid_for(int64_t n, int threads, int cap) {
    const int64_t g = std::max<int64_t>(1, (n + threads - 1) / threads);
    return (int)std::min<int64_t>(g, cap);
}

// room for `extra` more vectors: capacity doubles, the rows of the [dim][cap] matrix move to their new pitch
int reserve(lda_b200_index *ix, int64_t extra) {
    if (ix->n + extra <= ix->cap) return 0;
    // a multiple of the bulk-copy tile: the copy of a whole tile of any row stays inside the row
    const int64_t ncap = (std::max<int64_t>({ix->cap * 2, ix->n + extra, 1024}) + MMA_TV - 1) / MMA_TV * MMA_TV;
    double *nvec = nullptr, *nnorm = nullptr;
    int64_t *nids = nullptr;
    auto grow = [&]() -> int {
        ITRY(ialloc(ix, &nvec, (size_t)ix->dim * ncap));
        ITRY(ialloc(ix, &nnorm, (size_t)ncap));
        ITRY(ialloc(ix, &nids, (size_t)ncap));
        if (ix->n > 0) {
            ICU(cudaMemcpy2DAsync(nvec, (size_t)ncap * 8, ix->vec, (size_t)ix->cap * 8, (size_t)ix->n * 8, (size_t)ix->dim,
                                  cudaMemcpyDeviceToDevice, ix->stream));
            ICU(cudaMemcpyAsync(nnorm, ix->norm, (size_t)ix->n * 8, cudaMemcpyDeviceToDevice, ix->stream));
            ICU(cudaMemcpyAsync(nids, ix->ids, (size_t)ix->n * 8, cudaMemcpyDeviceToDevice, ix->stream));
        }
        return 0;
    };
    const int rc = grow();
    if (rc) {  // the index keeps its old storage; nothing allocated here survives
        ifree(ix, &nvec);
        ifree(ix, &nnorm);
        ifree(ix, &nids);
        return rc;
    }
    ifree(ix, &ix->vec);
    ifree(ix, &ix->norm);
    ifree(ix, &ix->ids);
    ix->vec = nvec;
    ix->norm = nnorm;
    ix->ids = nids;
    ix->cap = ncap;
    return 0;
}

// queries per pass over the index and the kernel that runs it
struct ScanPlan {
    int QB;
    bool mma;  // FP64 tensor-core kernel (dot-product comparers, dim % 4 == 0, more than 4 queries)
};
ScanPlan plan_scan(const lda_b200_index *ix, int nb) {
    if (nb == 1) return {1, false};
    if (nb <= 4) return {4, false};
    const bool dot = acc_class(ix->comparer) == ACC_DOT && ix->dim % 4 == 0 && !getenv("LDA_B200_INDEX_NO_MMA");
    const size_t limit = (size_t)227 * 1024;
    if (dot && nb > 16 && scan_mma_smem_bytes(ix->dim, 64) <= limit) return {64, true};
    if (dot && scan_mma_smem_bytes(ix->dim, 16) <= limit) return {16, true};
    return {16, false};
}

int launch_scan(lda_b200_index *ix, const ScanArgs &a, const ScanPlan &pl, int nchunks) {
    if (pl.mma) {
        const size_t smem = scan_mma_smem_bytes(a.dim, pl.QB);
        const dim3 grid((unsigned)grid_for(a.n, MMA_TV, ix->sms), (unsigned)nchunks);
        auto go = [&](auto kern) -> int {
            ICU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 256, smem, ix->stream>>>(a);
            IKCHECK();
            return 0;
        };
        return pl.QB == 64 ? go(index_scan_mma_kernel<8>) : go(index_scan_mma_kernel<2>);
    }
    const dim3 grid((unsigned)grid_for(a.n, 256, ix->sms * 6), (unsigned)nchunks);
    const size_t smem = (size_t)a.ktile * pl.QB * sizeof(double);
    auto by_qb = [&](auto cls_tag) -> int {
        constexpr int C = decltype(cls_tag)::value;
        if (pl.QB == 1)
            index_scan_kernel<1, C><<<grid, 256, smem, ix->stream>>>(a);
        else if (pl.QB == 4)
            index_scan_kernel<4, C><<<grid, 256, smem, ix->stream>>>(a);
        else
            index_scan_kernel<16, C><<<grid, 256, smem, ix->stream>>>(a);
        IKCHECK();
        return 0;
    };
    switch (acc_class(a.comparer)) {
        case ACC_DOT:
            return by_qb(std::integral_constant<int, ACC_DOT>{});
        case ACC_L2:
            return by_qb(std::integral_constant<int, ACC_L2>{});
        case ACC_L1:
            return by_qb(std::integral_constant<int, ACC_L1>{});
        default:
            return by_qb(std::integral_constant<int, ACC_HAMMING>{});
    }
}

// Search for one batch of nb queries already on the device as a dim x nb row-major matrix
int search_batch(lda_b200_index *ix, const double *d_q, int nb, int k, int64_t *ids_out, double *dist_out, int32_t *found_out) {
    const int64_t n = ix->n;
    const int k_eff = (int)std::min<int64_t>(k, n);
    // the k survivors of a query are ordered by one thread block (bitonic network over the next power of two)
    if (k_eff > (1 << 20))
        return ifail(ix, LDA_B200_EUNSUPPORTED, "k = %d nearest neighbours per query exceeds the supported 2^20", k_eff);
    const ScanPlan pl = plan_scan(ix, nb);
    const int QB = pl.QB;
    const int nchunks = (nb + QB - 1) / QB;
    int P = 1;
    while (P < std::max(k_eff, 1)) P <<= 1;

    double *qpack = nullptr, *qnorm = nullptr, *d_dist = nullptr;
    uint64_t *keys = nullptr;
    SelState *st = nullptr;
    unsigned *hist = nullptr, *count = nullptr;
    Cand *cand = nullptr;
    int64_t *d_ids = nullptr;
    int32_t *d_found = nullptr;
    auto run = [&]() -> int {
        ITRY(ialloc(ix, &qpack, (size_t)nchunks * ix->dim * QB));
        ITRY(ialloc(ix, &qnorm, (size_t)nb));
        ITRY(ialloc(ix, &keys, (size_t)nb * n));
        ITRY(ialloc(ix, &st, (size_t)nb));
        ITRY(ialloc(ix, &hist, (size_t)nb * SEL_BINS));
        ITRY(ialloc(ix, &count, (size_t)nb));
        ITRY(ialloc(ix, &cand, (size_t)nb * P));
        ITRY(ialloc(ix, &d_ids, (size_t)nb * k));
        ITRY(ialloc(ix, &d_dist, (size_t)nb * k));
        ITRY(ialloc(ix, &d_found, (size_t)nb));

        // ---- distances: comparer(query, vector) for every indexed vector (index.go:96-98, 106-107) ----
        ICU(cudaEventRecord(ix->ev[0], ix->stream));
        index_pack_queries_kernel<<<grid_for((int64_t)nchunks * ix->dim * QB, 256, ix->sms * 4), 256, 0, ix->stream>>>(
            d_q, nb, ix->dim, nb, QB, qpack, qnorm);
        IKCHECK();
        ScanArgs a;
        a.vec = ix->vec;
        a.inorm = ix->norm;
        a.qpack = qpack;
        a.qinorm = qnorm;
        a.keys = keys;
        a.cap = ix->cap;
        a.n = n;
        a.dim = ix->dim;
        a.nq = nb;
        a.comparer = ix->comparer;
        a.ktile = std::min(ix->dim, 4096 / QB);  // <= 32 KB of shared memory
        if (n > 0) ITRY(launch_scan(ix, a, pl, nchunks));
        ix->passes += nchunks;
        ICU(cudaEventRecord(ix->ev[1], ix->stream));

        // ---- the k smallest (index.go:99-112 keeps them in a heap; here: radix selection of the k-th composite) ----
        const int take_all = k_eff >= n ? 1 : 0;
        sel_init_kernel<<<(nb + 127) / 128, 128, 0, ix->stream>>>(st, nb, (uint32_t)k_eff, take_all);
        IKCHECK();
        ICU(cudaMemsetAsync(count, 0, (size_t)nb * sizeof(unsigned), ix->stream));
        const dim3 sgrid((unsigned)grid_for(n, 256 * 8, ix->sms * 4), (unsigned)nb);
        if (!take_all) {
            ICU(cudaMemsetAsync(hist, 0, (size_t)nb * SEL_BINS * sizeof(unsigned), ix->stream));
            for (int pass = 0; pass < SEL_PASSES; ++pass) {
                sel_hist_kernel<<<sgrid, 256, 0, ix->stream>>>(keys, n, st, pass, hist);
                IKCHECK();
                sel_pick_kernel<<<nb, 256, 0, ix->stream>>>(st, pass, hist);
                IKCHECK();
            }
        }
        if (n > 0) {
            sel_collect_kernel<<<sgrid, 256, 0, ix->stream>>>(keys, n, st, cand, P, count);
            IKCHECK();
        }
        sel_sort_emit_kernel<<<nb, 1024, 0, ix->stream>>>(cand, P, count, k, ix->ids, d_ids, d_dist, d_found);
        IKCHECK();
        ICU(cudaEventRecord(ix->ev[2], ix->stream));

        ICU(cudaMemcpyAsync(ids_out, d_ids, (size_t)nb * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
        ICU(cudaMemcpyAsync(dist_out, d_dist, (size_t)nb * k * sizeof(double), cudaMemcpyDeviceToHost, ix->stream));
        if (found_out) ICU(cudaMemcpyAsync(found_out, d_found, (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ix->stream));
        ICU(cudaStreamSynchronize(ix->stream));
        float a_ms = 0, b_ms = 0;
        ICU(cudaEventElapsedTime(&a_ms, ix->ev[0], ix->ev[1]));
        ICU(cudaEventElapsedTime(&b_ms, ix->ev[1], ix->ev[2]));
        ix->scan_ms += a_ms;
        ix->select_ms += b_ms;
        return 0;
    };
    const int rc = run();
    ifree(ix, &qpack);
    ifree(ix, &qnorm);
    ifree(ix, &keys);
    ifree(ix, &st);
    ifree(ix, &hist);
    ifree(ix, &count);
    ifree(ix, &cand);
    ifree(ix, &d_ids);
    ifree(ix, &d_dist);
    ifree(ix, &d_found);
    return rc;
}

}  // namespace

int index_begin_append(lda_b200_index *ix, int device, int dim, int64_t n, double **dst, int64_t *ld) {
    if (!ix) return LDA_B200_EINVAL;
    if (device != ix->device) return ifail(ix, LDA_B200_EINVAL, "the index lives on device %d, the model on device %d", ix->device, device);
    if (dim < 1 || n < 0) return ifail(ix, LDA_B200_EINVAL, "bad vector batch");
    if (ix->dim && dim != ix->dim)
        return ifail(ix, LDA_B200_EINVAL, "vector has %d elements, the index holds vectors of %d", dim, ix->dim);
    if (ix->n + n >= (1LL << 32) - 1) return ifail(ix, LDA_B200_EUNSUPPORTED, "more than 2^32 - 2 vectors");
    ICU(cudaSetDevice(ix->device));
    {
        const int32_t old_dim = ix->dim;
        ix->dim = dim;  // reserve sizes the rows by it
        const int rc = reserve(ix, n);
        if (rc) {
            ix->dim = old_dim;  // a failed first add does not pin the dimension
            return rc;
        }
    }
    ICU(cudaStreamSynchronize(ix->stream));
    *dst = ix->vec + ix->n;
    *ld = ix->cap;
    return 0;
}

int index_finish_append(lda_b200_index *ix, int64_t n, const int64_t *ids) {
    if (!ix) return LDA_B200_EINVAL;
    if (n <= 0) return 0;
    ICU(cudaSetDevice(ix->device));
    ICU(cudaMemcpyAsync(ix->ids + ix->n, ids, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    index_norms_kernel<<<grid_for(n, 256, ix->sms * 8), 256, 0, ix->stream>>>(ix->vec, ix->cap, ix->dim, ix->n, n, ix->norm);
    IKCHECK();
    ICU(cudaStreamSynchronize(ix->stream));
    ix->h_ids.insert(ix->h_ids.end(), ids, ids + n);
    ix->n += n;
    return 0;
}

extern "C" {

const char *lda_b200_index_last_error(const lda_b200_index *ix) { return ix ? ix->err.c_str() : g_index_create_error.c_str(); }

int lda_b200_index_create(int32_t device, int32_t comparer, lda_b200_index **out) {
    lda_b200_index *ix = nullptr;
    if (!out) return ifail(ix, LDA_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (comparer < 0 || comparer >= N_COMPARERS) return ifail(ix, LDA_B200_EINVAL, "unknown comparer %d", (int)comparer);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ifail(ix, LDA_B200_ECUDA, "no CUDA device available (%s); liblda_b200 has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return ifail(ix, LDA_B200_EINVAL, "device %d out of range [0,%d)", device, ndev);
    lda_b200_index *nx = new lda_b200_index();
    nx->device = device;
    nx->comparer = comparer;
    auto init = [&](lda_b200_index *ix) -> int {
        ICU(cudaSetDevice(device));
        cudaDeviceProp prop;
        ICU(cudaGetDeviceProperties(&prop, device));
        ix->sms = prop.multiProcessorCount;
        ICU(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
        for (cudaEvent_t &ev : ix->ev) ICU(cudaEventCreate(&ev));
        cudaMemPool_t pool;
        ICU(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;  // recycle the per-search scratch instead of returning it to the driver
        ICU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        return 0;
    };
    const int rc = init(nx);
    if (rc) {
        g_index_create_error = nx->err;
        lda_b200_index_destroy(nx);
        return rc;
    }
    *out = nx;
    return 0;
}

void lda_b200_index_destroy(lda_b200_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    ifree(ix, &ix->vec);
    ifree(ix, &ix->norm);
    ifree(ix, &ix->ids);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    for (cudaEvent_t ev : ix->ev)
        if (ev) cudaEventDestroy(ev);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int64_t lda_b200_index_len(const lda_b200_index *ix) { return ix ? ix->n : 0; }

int lda_b200_index_add(lda_b200_index *ix, int32_t dim, int64_t n, const double *m, int64_t ld, const int64_t *ids) {
    if (!ix) return LDA_B200_EINVAL;
    if (dim < 1) return ifail(ix, LDA_B200_EINVAL, "vector dimension must be >= 1 (got %d)", (int)dim);
    if (ix->dim && dim != ix->dim)
        return ifail(ix, LDA_B200_EINVAL, "vector has %d elements, the index holds vectors of %d", (int)dim, ix->dim);
    if (n < 0 || (n > 0 && !m) || ld < 1 || (n > 1 && ld < n)) return ifail(ix, LDA_B200_EINVAL, "bad vector batch (n = %lld, ld = %lld)", (long long)n, (long long)ld);
    if (ix->n + n >= (1LL << 32) - 1) return ifail(ix, LDA_B200_EUNSUPPORTED, "more than 2^32 - 2 vectors");
    if (n == 0) return 0;
    ICU(cudaSetDevice(ix->device));
    {
        const int32_t old_dim = ix->dim;
        ix->dim = dim;  // reserve sizes the rows by it
        const int rc = reserve(ix, n);
        if (rc) {
            ix->dim = old_dim;  // a failed first add does not pin the dimension
            return rc;
        }
    }
    ICU(cudaMemcpy2DAsync(ix->vec + ix->n, (size_t)ix->cap * 8, m, (size_t)ld * 8, (size_t)n * 8, (size_t)dim, cudaMemcpyHostToDevice,
                          ix->stream));
    std::vector<int64_t> new_ids((size_t)n);
    for (int64_t j = 0; j < n; ++j) new_ids[(size_t)j] = ids ? ids[j] : ix->n + j;
    ICU(cudaMemcpyAsync(ix->ids + ix->n, new_ids.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    index_norms_kernel<<<grid_for(n, 256, ix->sms * 8), 256, 0, ix->stream>>>(ix->vec, ix->cap, dim, ix->n, n, ix->norm);
    IKCHECK();
    ICU(cudaStreamSynchronize(ix->stream));  // the caller's memory is not referenced after return
    ix->h_ids.insert(ix->h_ids.end(), new_ids.begin(), new_ids.end());
    ix->n += n;
    return 0;
}

int lda_b200_index_remove(lda_b200_index *ix, int64_t id) {
    if (!ix) return LDA_B200_EINVAL;
    const auto it = std::find(ix->h_ids.begin(), ix->h_ids.end(), id);  // index.go:121-122: the first match
    if (it == ix->h_ids.end()) return 0;
    ICU(cudaSetDevice(ix->device));
    const int64_t pos = it - ix->h_ids.begin();
    const int64_t tail = ix->n - pos - 1;
    if (tail > 0) {  // index.go:123-129: everything behind moves up by one, order preserved
        double *tmp = nullptr;
        ITRY(ialloc(ix, &tmp, (size_t)(ix->dim + 2) * tail));
        int64_t *tmp_ids = reinterpret_cast<int64_t *>(tmp + (size_t)(ix->dim + 1) * tail);
        ICU(cudaMemcpy2DAsync(tmp, (size_t)tail * 8, ix->vec + pos + 1, (size_t)ix->cap * 8, (size_t)tail * 8, (size_t)ix->dim,
                              cudaMemcpyDeviceToDevice, ix->stream));
        ICU(cudaMemcpyAsync(tmp + (size_t)ix->dim * tail, ix->norm + pos + 1, (size_t)tail * 8, cudaMemcpyDeviceToDevice, ix->stream));
        ICU(cudaMemcpyAsync(tmp_ids, ix->ids + pos + 1, (size_t)tail * 8, cudaMemcpyDeviceToDevice, ix->stream));
        ICU(cudaMemcpy2DAsync(ix->vec + pos, (size_t)ix->cap * 8, tmp, (size_t)tail * 8, (size_t)tail * 8, (size_t)ix->dim,
                              cudaMemcpyDeviceToDevice, ix->stream));
        ICU(cudaMemcpyAsync(ix->norm + pos, tmp + (size_t)ix->dim * tail, (size_t)tail * 8, cudaMemcpyDeviceToDevice, ix->stream));
        ICU(cudaMemcpyAsync(ix->ids + pos, tmp_ids, (size_t)tail * 8, cudaMemcpyDeviceToDevice, ix->stream));
        ifree(ix, &tmp);
    }
    ix->h_ids.erase(it);
    ix->n -= 1;
    return 0;
}

int lda_b200_index_search(lda_b200_index *ix, int64_t nq, const double *q, int64_t ldq, int32_t k, int64_t *ids_out,
                          double *dist_out, int32_t *found_out) {
    if (!ix) return LDA_B200_EINVAL;
    if (k < 1) return ifail(ix, LDA_B200_EINVAL, "k must be >= 1 (got %d)", (int)k);  // the reference panics on k = 0 (index.go:108)
    if (nq < 0 || (nq > 0 && (!q || !ids_out || !dist_out)) || ldq < 1 || (nq > 1 && ldq < nq))
        return ifail(ix, LDA_B200_EINVAL, "bad query batch (nq = %lld, ldq = %lld)", (long long)nq, (long long)ldq);
    ix->scan_ms = ix->select_ms = 0;
    ix->launches = ix->passes = 0;
    if (nq == 0) return 0;
    if (ix->n == 0) {  // index.go:101-103: fewer than k vectors -> what there is
        for (int64_t i = 0; i < nq * k; ++i) {
            ids_out[i] = -1;
            dist_out[i] = __builtin_nan("");
        }
        if (found_out)
            for (int64_t i = 0; i < nq; ++i) found_out[i] = 0;
        return 0;
    }
    ICU(cudaSetDevice(ix->device));
    // queries per round: the key matrix (8 bytes per query and vector) stays below 1 GiB
    const int64_t qmax = std::max<int64_t>(1, std::min<int64_t>(64, ((int64_t)1 << 27) / ix->n));
    double *d_q = nullptr;
    ITRY(ialloc(ix, &d_q, (size_t)ix->dim * qmax));
    int rc = 0;
    for (int64_t q0 = 0; q0 < nq && !rc; q0 += qmax) {
        const int nb = (int)std::min<int64_t>(qmax, nq - q0);
        cudaError_t e = cudaMemcpy2DAsync(d_q, (size_t)nb * 8, q + q0, (size_t)ldq * 8, (size_t)nb * 8, (size_t)ix->dim,
                                          cudaMemcpyHostToDevice, ix->stream);
        if (e != cudaSuccess) {
            rc = ifail(ix, LDA_B200_ECUDA, "query upload failed: %s", cudaGetErrorString(e));
            break;
        }
        rc = search_batch(ix, d_q, nb, k, ids_out + q0 * k, dist_out + q0 * k, found_out ? found_out + q0 : nullptr);
    }
    ifree(ix, &d_q);
    return rc;
}

int lda_b200_index_last_search_time(const lda_b200_index *ix, float *scan_ms, float *select_ms, int32_t *launches,
                                    int32_t *passes) {
    if (!ix) return LDA_B200_EINVAL;
    if (scan_ms) *scan_ms = ix->scan_ms;
    if (select_ms) *select_ms = ix->select_ms;
    if (launches) *launches = ix->launches;
    if (passes) *passes = ix->passes;
    return 0;
}

}  // extern "C"
