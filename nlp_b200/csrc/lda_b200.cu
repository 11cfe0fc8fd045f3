// lda_b200.cu -- engine + C ABI (include/lda_b200.h) of the B200 LDA SCVB0 path.
// Reference behaviour: james-bowman/nlp lda.go (cited per function).  No CPU fallback: every
// compute step is a kernel from kernels.cuh / pcg.cuh launched on the handle's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in Nsight Systems / ncu --nvtx, cost nothing without a tool attached

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/lda_b200.h"
#include "../../include/lda_b200_index.h"
#include "index_internal.h"
#include "kernels.cuh"
#include "nccl_dl.h"
#include "nvls.h"
#include "pcg.cuh"

using namespace ldab;

namespace {

thread_local std::string g_create_error;

struct Segment {
    int64_t g0, g1;  // global document range
    int64_t l0;      // first local row
};

// Word ids of a pageable caller array on their way to the device behind the first iteration's launches: the copies into
// the page-locked staging ring are host work (memcpy), so they run on their own thread while the calling thread queues the
// launches.  recorded = number of sub-ranges whose "ids in place" event has been recorded on the copy stream; a launch may
// only be made to wait on an event that has been recorded (waiting on an unrecorded event is a no-op in CUDA).
struct UploadJob {
    std::thread worker;
    std::atomic<int> recorded{0};
    int rc = 0;
    std::string err;
    ~UploadJob() {
        if (worker.joinable()) worker.join();
    }
};

// resident corpus shard + per-document statistics
struct Corpus {
    int64_t launch_range = 0;  // documents a launch range of this plan covers (a step's minibatches, a Transform range, or all)
    int64_t D = 0;       // global documents
    int64_t Dl = 0;      // local documents
    int64_t nnz = 0;     // local non-zeros
    int64_t b = 0;       // block (minibatch) size used for sharding
    int conc = 1;        // minibatches per launch on this GPU (Processes / ranks), fit only
    int64_t sub = 0;     // fit: documents per sub-range of a step's launch range (a multiple of b).  The longest-first order is
                         // per sub-range, so a step can run as one launch over its whole range or as one launch per
                         // sub-range -- the latter while word ids still stream in (first iteration) or results stream
                         // out (last iteration), so that the transfers overlap the neighbouring launches
    std::vector<int64_t> sub_begin;  // first local document of every sub-range, ascending
    std::vector<Segment> segs;
    int64_t *doc_ptr = nullptr;
    int32_t *ind = nullptr;
    float *val = nullptr;
    float *wc = nullptr;
    float *theta = nullptr;
    int *order = nullptr;  // documents of each launch range, longest first (order inside a minibatch is free:
                           // documents only interact through commutative sums)
    // word ids may still be arriving on the copy stream during the first iteration of lda_b200_fit: one event per launch
    // range, recorded when that range's ids are in place
    std::vector<cudaEvent_t> range_ev;  // one per sub-range (sub_begin)
    bool ind_deferred = false;
    std::shared_ptr<UploadJob> job;     // set while a host thread feeds the deferred ids from pageable memory
    double n_global = 0;  // sum of all values of the (global) matrix: wordsInCorpus contribution, lda.go:481
    // where the entries come from (the caller's arrays, valid for the duration of the call) and the local indptr: data
    // movement can be issued per range of local documents (upload_part)
    std::vector<int64_t> lp;
    const int64_t *src_indptr = nullptr, *src_ind = nullptr;
    const double *src_data = nullptr;
    bool src_device = false;
    int64_t W = 0;
    int64_t range_docs = 0;  // > 0: Transform streams the corpus in ranges of this many local documents
};

// a term x document matrix converted to CSC on the device (from CSR or dense input); word ids ascending inside a document
struct DeviceCSC {
    int64_t W = 0, D = 0, nnz = 0;
    int64_t *indptr = nullptr;      // device [D+1]
    int64_t *ind = nullptr;         // device [nnz]
    double *data = nullptr;         // device [nnz]
    std::vector<int64_t> h_indptr;  // host copy of indptr (document lengths drive sharding and the launch order)
};

struct EventPair {
    cudaEvent_t a, b;
    int kind;  // 0 = minibatch kernel, 1 = M-step kernels, 2 = exchange (all-reduce or peer kernel; includes the wait
               // for the slowest rank); phases of the peer-memory exchange: 3 = opening barrier (= waiting for the slowest
               // rank's minibatch kernel), 4 = reduce + blend over peer memory, 5 = column sums + closing barrier + nZ,
               // 6 = refresh of C
};

}  // namespace

struct lda_b200_handle {
    lda_b200_config cfg;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;

    // model (lda.go:137-140)
    int64_t W = 0;
    int K = 0, Kp = 0;
    bool fitted = false;
    float *nphi = nullptr;
    double *nz = nullptr;
    float *inv_z = nullptr;
    float *cmat = nullptr;  // (nPhi + eta) * invZ, [W][Kp], kept current by the M-step kernels

    // scratch
    float *nphi_hat = nullptr;
    double *nzhat = nullptr;
    double *colsum = nullptr;
    float *inv_colsum = nullptr;
    int *doc_counter = nullptr;
    int64_t pair_docs_max = 16384;  // launch ranges up to this many documents use the PAIR kernels (LDA_B200_PAIR_MAX_DOCS; 0 = never)
    int *err_flag = nullptr;
    double *dacc = nullptr;  // [2]: perplexity accumulator, N accumulator
    float *rho_tab = nullptr;
    int rho_cap = 0;
    double *partials = nullptr;  // per-block partial sums of the cross-block reductions (summed in block order)
    size_t partials_cap = 0;
    bool deterministic = false;  // sufficient statistics accumulated in 64-bit fixed point (lda_b200_set_deterministic)
    // documents per warp of the 64 < K <= 256 kernels: 1 = scvb0_docs32_kernel, 2 / 4 = scvb0_docsg_kernel
    // (LDA_B200_DOCS_PER_WARP_FIT / LDA_B200_DOCS_PER_WARP_TRANSFORM override the defaults).  Measured at K = 256 on
    // 65,536-document Transform ranges: 29.7 / 27.0 / 26.1 ms for 1 / 2 / 4; the fit launch, bound by the L2's reduction
    // throughput, does not move (53.1 / 52.4 / 53.5 ms per iteration), so it keeps one document per warp.
    int docs_per_warp_fit = 1, docs_per_warp_transform = 4;
    unsigned long long *hat_fx = nullptr;
    void *stage = nullptr;
    size_t stage_bytes = 0;
    double *h_pinned = nullptr;  // small pinned read-back area

    // page-locked staging ring for pageable caller buffers (Go slices, NumPy arrays): host threads copy between the
    // caller's memory and the ring while the copy engine moves the previous slot
    static constexpr int kStageSlots = 3;
    static constexpr size_t kStageSlotBytes = (size_t)64 << 20;
    void *pin_slot[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t pin_done[kStageSlots] = {nullptr, nullptr, nullptr};
    int host_threads = 4;

    // streaming of the result during the last iteration of lda_b200_fit (second stream)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t down_stream = nullptr;  // device -> host leg of the Transform pipeline
    cudaEvent_t copy_ev = nullptr;
    double *pending_theta_out = nullptr;  // caller's K x D buffer, set by lda_b200_fit for the duration of the call
    double *dout = nullptr;               // K x Dl normalised, transposed result on the device
    bool theta_streamed = false;

    // fit run
    Corpus fitc;
    bool fit_open = false;
    int32_t it_done = 0;
    double prev_perplexity = 0;
    bool stopped = false;

    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory path (fused reduction + M-step over NVLink); falls back to ncclAllReduce + blend when unavailable
    bool peer_ok = false;
    bool peer_wanted = true;
    PeerTable peers;
    float *hat_buf[2] = {nullptr, nullptr};  // cudaMalloc'ed (IPC-capable) nPhiHat ping-pong buffers
    float *nphi_ipc = nullptr;               // cudaMalloc'ed nPhi when peers map it
    unsigned *peer_flags = nullptr;
    double *peer_nzpart = nullptr;
    std::vector<void *> ipc_opened;
    unsigned epoch = 0;
    int hat_cur = 0;
    std::string peer_note;
    // NVLink multicast (NVLS): the two nPhiHat buffers and nPhi live in one multicast-backed region per GPU; the step's
    // exchange kernel reads sums with multimem.ld_reduce and writes every replica with multimem.st (nvls.h)
    NvlsRegion nvls;
    bool nvls_wanted = true;
    bool peer_setup_det = false;  // `deterministic` at the time of peer_setup (NVLS is not used in deterministic mode)
    float *mc_hat[2] = {nullptr, nullptr};
    float *mc_nphi = nullptr;
    std::string nvls_note;

    // stats
    std::atomic<int64_t> launches{0};
    bool profiling = false;
    std::vector<EventPair> ev_pool;
    size_t ev_used = 0;
    float mb_ms = 0, blend_ms = 0, ar_ms = 0;
    int mb_n = 0, blend_n = 0, ar_n = 0;
    float phase_ms[4] = {0, 0, 0, 0};  // kinds 3..6
};

namespace {

int fail(lda_b200_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h)
        h->err = buf;
    else
        g_create_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, LDA_B200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define NC(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess)                                                                           \
            return fail(h, LDA_B200_ENCCL, "%s failed: %s", #call, nccl().GetErrorString(r_));            \
    } while (0)
#define TRY(call)            \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)
#define KCHECK()                      \
    do {                              \
        h->launches++;                \
        CU(cudaPeekAtLastError());    \
    } while (0)

// NVTX range for the lifetime of the object: the host-facing calls, each iteration, each step's launch and exchange /
// M-step, each Transform range (the reference has no tracing hooks at all)
struct Nvtx {
    explicit Nvtx(const char *name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
    Nvtx(const Nvtx &) = delete;
    Nvtx &operator=(const Nvtx &) = delete;
};

// LDA_B200_TIMING=1: wall-clock phases of the host-facing calls on stderr (diagnostic)
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    }
    explicit PhaseTimer(cudaStream_t s) : on(getenv("LDA_B200_TIMING") != nullptr), st(s), t0(on ? now() : 0) {}
    void lap(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const double t = now();
        fprintf(stderr, "[lda_b200] %-28s %8.2f ms\n", what, 1e3 * (t - t0));
        t0 = t;
    }
};

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) with the release threshold
// lifted, so the multi-GB corpus / result buffers of one call are recycled by the next instead of paying
// cudaMalloc / cudaFree (tens to hundreds of ms at these sizes) on every Fit / Transform.
template <typename T>
int dalloc(lda_b200_handle *h, T **p, size_t n) {
    if (*p) {
        cudaFreeAsync(*p, h->stream);
        *p = nullptr;
    }
    if (n == 0) n = 1;
    CU(cudaMallocAsync((void **)p, n * sizeof(T), h->stream));
    return 0;
}
template <typename T>
void dfree(lda_b200_handle *h, T **p) {
    if (*p) cudaFreeAsync(*p, h->stream);
    *p = nullptr;
}

void free_corpus(lda_b200_handle *h, Corpus &c) {
    if (c.job && c.job->worker.joinable()) c.job->worker.join();
    c.job.reset();
    if (c.ind_deferred && h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (cudaEvent_t e : c.range_ev) cudaEventDestroy(e);
    dfree(h, &c.doc_ptr);
    dfree(h, &c.ind);
    dfree(h, &c.val);
    dfree(h, &c.wc);
    dfree(h, &c.theta);
    dfree(h, &c.order);
    c = Corpus();
}

inline int grid_for(int64_t n, int threads, int cap) {
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)std::min<int64_t>(g, cap);
}

// out[0..n) (+)= sum over the nblocks per-block partials written by the kernel just launched on st, in block order
int ensure_partials(lda_b200_handle *h, size_t doubles) {
    if (doubles <= h->partials_cap) return 0;
    TRY(dalloc(h, &h->partials, doubles));
    h->partials_cap = doubles;
    return 0;
}
int reduce_partials(lda_b200_handle *h, int nblocks, int n, double *out, bool accumulate, cudaStream_t st) {
    if (n == 1)
        sum_partials_kernel<1024><<<1, 1024, 0, st>>>(h->partials, nblocks, n, out, accumulate ? 1 : 0);
    else
        sum_partials_kernel<4><<<(n + 255) / 256, 1024, 0, st>>>(h->partials, nblocks, n, out, accumulate ? 1 : 0);
    KCHECK();
    return 0;
}

int validate_config(lda_b200_handle *h, const lda_b200_config *c) {
    if (!c) return fail(h, LDA_B200_EINVAL, "config is NULL");
    if (c->K < 1) return fail(h, LDA_B200_EINVAL, "K must be >= 1 (got %d)", c->K);
    if (c->K > 1024) return fail(h, LDA_B200_EUNSUPPORTED, "K = %d > 1024 is not supported by this build", c->K);
    if (c->batch_size < 1) return fail(h, LDA_B200_EINVAL, "BatchSize must be >= 1 (got %d)", c->batch_size);
    if (c->burn_in_passes < 0 || c->transformation_passes < 0 || c->iterations < 0)
        return fail(h, LDA_B200_EINVAL, "negative pass/iteration count");
    return 0;
}

// ---- kernel dispatch on the topic width ------------------------------------------------------------
template <bool FIT>
int launch_docs(lda_b200_handle *h, const DocsArgs &a, const int *order, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    const int vecs = a.Kp / 4;
    const int n_docs = a.doc_end - a.doc_begin;
    if (n_docs <= 0) return 0;
    auto grid_of = [&](auto kern, int gpw, int *grid) -> int {
        int occ = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
        if (occ < 1) occ = 1;
        const int64_t warps = ((int64_t)n_docs + gpw - 1) / gpw;
        *grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->sms * occ);
        return 0;
    };
    // narrow topic axes: several documents per warp, natural order
    auto narrow = [&](auto kern, int lanes) -> int {
        int grid;
        TRY(grid_of(kern, 32 / lanes, &grid));
        kern<<<grid, 256, 0, st>>>(a);
        KCHECK();
        return 0;
    };
    // K > 64: one document per warp, longest documents first, nPhi rows staged through a shared-memory ring
    auto wide = [&](auto kern, int nv) -> int {
        const int smem = docs32_smem_bytes(nv);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        if (occ < 1) occ = 1;
        const int grid = (int)std::min<int64_t>(((int64_t)n_docs + 7) / 8, (int64_t)h->sms * occ);
        kern<<<grid, 256, smem, st>>>(a, order);
        KCHECK();
        return 0;
    };
    // 64 < K <= 256: GD documents per warp (scvb0_docsg_kernel), longest first
    auto grouped = [&](auto kern, int gd, int nv) -> int {
        const int smem = docsg_smem_bytes(gd, nv);
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
        if (occ < 1) occ = 1;
        const int64_t warps = ((int64_t)n_docs + gd - 1) / gd;
        const int grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->sms * occ);
        kern<<<grid, 256, smem, st>>>(a, order);
        KCHECK();
        return 0;
    };
    // deterministic mode (fit only): the variants that accumulate the sufficient statistics in fixed point
    constexpr bool CAN_DET = FIT;
    const bool det = CAN_DET && a.hat_fx != nullptr;
    const int gd = det ? 1 : (FIT ? h->docs_per_warp_fit : h->docs_per_warp_transform);
    // a launch range too small to fill the device lasts as long as its longest document: shorten the per-token chain.
    // Fit only: a fit's result depends on its schedule anyway (the order of the fp32 reductions into nPhiHat), whereas the
    // Transform of a document must not depend on what else is in the call (test_transform_pipeline_ranges: pieces of a
    // corpus give bitwise what the whole gives), so Transform keeps one arithmetic order whatever the launch size.
    // K <= 256 only: that is where it pays (measured at K = 100 and 256); at K = 1024 it made no difference
    const bool pair = FIT && !det && a.pair && vecs > 16 && vecs <= 64;
    if (!pair && gd > 1 && vecs > 16 && vecs <= 64) {
        const bool full = vecs == 32 || vecs == 64;
        const bool wide_k = vecs > 32;  // 128 < K <= 256
#define LDA_GROUPED(GD, NV)                                                                 \
    do {                                                                                    \
        if (full) return grouped(scvb0_docsg_kernel<GD, NV, true, FIT, false>, GD, NV);     \
        return grouped(scvb0_docsg_kernel<GD, NV, false, FIT, false>, GD, NV);              \
    } while (0)
        if (gd == 2) {
            if (wide_k) LDA_GROUPED(2, 4);
            LDA_GROUPED(2, 2);
        }
        if (wide_k) LDA_GROUPED(4, 8);
        LDA_GROUPED(4, 4);
#undef LDA_GROUPED
    }
#define LDA_NARROW(L)                                                              \
    do {                                                                           \
        if (det) return narrow(scvb0_docs_kernel<L, 1, 4, FIT, CAN_DET>, L);       \
        return narrow(scvb0_docs_kernel<L, 1, 4, FIT, false>, L);                  \
    } while (0)
#define LDA_WIDE(NV, FULLK)                                                        \
    do {                                                                           \
        if (det) return wide(scvb0_docs32_kernel<NV, FULLK, FIT, CAN_DET>, NV);    \
        if constexpr (FIT) {                                                       \
            if constexpr (NV <= 2) {                                                   \
                if (pair) return wide(scvb0_docs32_kernel<NV, FULLK, true, false, true>, NV); \
            }                                                                          \
        }                                                                          \
        return wide(scvb0_docs32_kernel<NV, FULLK, FIT, false>, NV);               \
    } while (0)
    if (vecs <= 4) LDA_NARROW(4);
    if (vecs <= 8) LDA_NARROW(8);
    if (vecs <= 16) LDA_NARROW(16);
    if (vecs < 32) LDA_WIDE(1, false);
    if (vecs == 32) LDA_WIDE(1, true);
    if (vecs < 64) LDA_WIDE(2, false);
    if (vecs == 64) LDA_WIDE(2, true);
    if (vecs < 128) LDA_WIDE(4, false);
    if (vecs == 128) LDA_WIDE(4, true);
    if (vecs < 256) LDA_WIDE(8, false);
    LDA_WIDE(8, true);
#undef LDA_NARROW
#undef LDA_WIDE
}

// sum over the documents' entries of log2(dot) * v into h->dacc[0]
int launch_perplexity(lda_b200_handle *h, const PerpArgs &a) {
    const int vecs = a.Kp / 4;
    if (a.n_docs <= 0) return 0;
    auto go = [&](auto kern, int lanes) -> int {
        const int gpw = 32 / lanes;
        const int64_t warps = ((int64_t)a.n_docs + gpw - 1) / gpw;
        const int grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)h->sms * 8);
        TRY(ensure_partials(h, (size_t)grid));
        PerpArgs b = a;
        b.partials = h->partials;
        kern<<<grid, 256, 0, h->stream>>>(b);
        KCHECK();
        return reduce_partials(h, grid, 1, h->dacc, false, h->stream);
    };
    if (vecs <= 4) return go(perplexity_kernel<4, 1, 4>, 4);
    if (vecs <= 8) return go(perplexity_kernel<8, 1, 4>, 8);
    if (vecs <= 16) return go(perplexity_kernel<16, 1, 4>, 16);
    if (vecs <= 32) return go(perplexity_kernel<32, 1, 4>, 32);
    if (vecs <= 64) return go(perplexity_kernel<32, 2, 4>, 32);
    if (vecs <= 128) return go(perplexity_kernel<32, 4, 2>, 32);
    return go(perplexity_kernel<32, 8, 2>, 32);
}

inline int blend_block(int Kp) {
    const int vpr = Kp / 4;
    return vpr * std::max(1, 256 / vpr);
}

// ---- profiling events ---------------------------------------------------------------------------------
int ev_begin(lda_b200_handle *h, int kind, int *slot) {
    *slot = -1;
    if (!h->profiling || h->ev_used >= 16384) return 0;
    if (h->ev_used == h->ev_pool.size()) {
        EventPair p;
        CU(cudaEventCreate(&p.a));
        CU(cudaEventCreate(&p.b));
        h->ev_pool.push_back(p);
    }
    *slot = (int)h->ev_used++;
    h->ev_pool[*slot].kind = kind;
    CU(cudaEventRecord(h->ev_pool[*slot].a, h->stream));
    return 0;
}
int ev_end(lda_b200_handle *h, int slot) {
    if (slot >= 0) CU(cudaEventRecord(h->ev_pool[slot].b, h->stream));
    return 0;
}
int ev_collect(lda_b200_handle *h) {
    h->mb_ms = h->blend_ms = h->ar_ms = 0;
    h->mb_n = h->blend_n = h->ar_n = 0;
    for (float &x : h->phase_ms) x = 0;
    for (size_t i = 0; i < h->ev_used; ++i) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, h->ev_pool[i].a, h->ev_pool[i].b));
        if (h->ev_pool[i].kind == 0) {
            h->mb_ms += ms;
            h->mb_n++;
        } else if (h->ev_pool[i].kind == 1) {
            h->blend_ms += ms;
            h->blend_n++;
        } else if (h->ev_pool[i].kind >= 3) {
            h->phase_ms[h->ev_pool[i].kind - 3] += ms;
        } else {
            h->ar_ms += ms;
            h->ar_n++;
        }
    }
    h->ev_used = 0;
    return 0;
}

void peer_teardown(lda_b200_handle *h);

// ---- model allocation -------------------------------------------------------------------------------
int ensure_model(lda_b200_handle *h, int64_t W) {
    const int K = h->cfg.K, Kp = (K + 3) & ~3;
    if (h->W == W && h->K == K && h->nphi) return 0;
    if (W < 1) return fail(h, LDA_B200_EINVAL, "matrix has no rows (W = %lld)", (long long)W);
    if (W > 0x7fffffffLL) return fail(h, LDA_B200_EUNSUPPORTED, "W = %lld exceeds int32 word ids", (long long)W);
    if ((uint64_t)W * (uint64_t)Kp / 4u >= (1ull << 32))
        return fail(h, LDA_B200_EUNSUPPORTED, "nPhi of %lld x %d floats exceeds the 64 GB the kernels address with 32-bit float4 offsets", (long long)W, K);
    peer_teardown(h);
    h->W = W;
    h->K = K;
    h->Kp = Kp;
    h->fitted = false;
    TRY(dalloc(h, &h->nphi, (size_t)W * Kp));
    TRY(dalloc(h, &h->nphi_hat, (size_t)W * Kp));
    TRY(dalloc(h, &h->cmat, (size_t)W * Kp));
    TRY(dalloc(h, &h->nz, (size_t)Kp));
    TRY(dalloc(h, &h->nzhat, (size_t)Kp));
    TRY(dalloc(h, &h->colsum, (size_t)Kp));
    TRY(dalloc(h, &h->inv_z, (size_t)Kp));
    TRY(dalloc(h, &h->inv_colsum, (size_t)Kp));
    CU(cudaMemsetAsync(h->nphi_hat, 0, (size_t)W * Kp * sizeof(float), h->stream));
    CU(cudaMemsetAsync(h->nzhat, 0, (size_t)Kp * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->nz, 0, (size_t)Kp * sizeof(double), h->stream));
    return 0;
}

// ---- peer-memory setup ------------------------------------------------------------------------------------------
struct PeerRec {
    int pid, dev;
    cudaIpcMemHandle_t mh[5];  // hat0, hat1, nphi, flags, nzpart
    void *ptr[5];
};

void peer_teardown(lda_b200_handle *h) {
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    if (h->peer_ok || h->hat_buf[0]) {
        // the model buffers were plain cudaMalloc allocations shared with the peers, or slices of the multicast region
        const bool region = h->nvls.uc_va != 0;
        if (region && h->stream) cudaStreamSynchronize(h->stream);
        if (h->nphi_ipc) {
            if (!region) cudaFree(h->nphi_ipc);
            if (h->nphi == h->nphi_ipc) h->nphi = nullptr;
            h->nphi_ipc = nullptr;
        }
        for (int i = 0; i < 2; ++i) {
            if (h->hat_buf[i] && !region) cudaFree(h->hat_buf[i]);
            if (h->nphi_hat == h->hat_buf[i]) h->nphi_hat = nullptr;
            h->hat_buf[i] = nullptr;
        }
        nvls_release(h->nvls);
        h->mc_hat[0] = h->mc_hat[1] = h->mc_nphi = nullptr;
        if (h->peer_flags) cudaFree(h->peer_flags);
        if (h->peer_nzpart) cudaFree(h->peer_nzpart);
        h->peer_flags = nullptr;
        h->peer_nzpart = nullptr;
    }
    h->peer_ok = false;
}

// every rank contributes `bytes` bytes; all[q*bytes ..] = rank q's record (ncclAllGather through a device buffer)
int gather_records(lda_b200_handle *h, const void *mine, size_t bytes, std::vector<char> &all) {
    const int R = h->nranks;
    char *d = nullptr;
    TRY(dalloc(h, &d, bytes * (size_t)R));
    CU(cudaMemcpyAsync(d + bytes * (size_t)h->rank, mine, bytes, cudaMemcpyHostToDevice, h->stream));
    NC(nccl().AllGather(d + bytes * (size_t)h->rank, d, bytes, ncclChar, h->comm, h->stream));
    all.resize(bytes * (size_t)R);
    CU(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, &d);
    return 0;
}

// One multicast object over all ranks' GPUs with `buf_bytes` x 3 (two nPhiHat buffers, nPhi) bound on every GPU.
// Collective: every rank runs every stage, and after each stage the ranks agree whether all of them succeeded, so that a
// single failure switches everybody to the unicast peer-memory path (h->nvls.active stays false; nvls_note says why).
int nvls_setup(lda_b200_handle *h, size_t buf_bytes) {
    const int R = h->nranks;
    NvlsRegion &g = h->nvls;
    nvls_release(g);
    h->nvls_note.clear();
    DriverApi &d = drv();
    struct Rec {
        int ok, pid;
        uint64_t token;
    };
    auto note = [&](const std::string &s) {
        if (h->nvls_note.empty()) h->nvls_note = s;
    };
    auto agree = [&](bool ok, Rec *first, bool *all_ok) -> int {
        Rec mine = {ok ? 1 : 0, (int)getpid(), 0};
        if (h->rank == 0) mine.token = ((uint64_t)getpid() << 32) ^ (uint64_t)time(nullptr) ^ ((uint64_t)(uintptr_t)h >> 4);
        std::vector<char> all;
        TRY(gather_records(h, &mine, sizeof mine, all));
        *all_ok = true;
        const Rec *r = (const Rec *)all.data();
        for (int q = 0; q < R; ++q) {
            if (!r[q].ok) *all_ok = false;
            for (int p = 0; p < q; ++p)
                if (r[p].pid == r[q].pid) *all_ok = false;  // ranks of one process would need no file descriptors: not handled here
        }
        if (first) *first = r[0];
        return 0;
    };
    bool ok = h->nvls_wanted && !h->deterministic && R >= 2;
    if (!ok) note(h->deterministic ? "deterministic mode keeps the fixed-order unicast reduction" : "disabled");
    if (ok && !d.load()) {
        ok = false;
        note(d.error);
    }
    CUdevice dev = 0;
    if (ok) {
        int sup = 0;
        ok = d.DeviceGet(&dev, h->device) == CUDA_SUCCESS &&
             d.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && sup;
        if (!ok) note("the device does not support multicast objects");
    }
    Rec r0;
    bool all_ok = false;
    TRY(agree(ok, &r0, &all_ok));
    if (!all_ok) {
        note("not every rank can use multicast objects (or ranks share a process)");
        return 0;
    }
    const uint64_t token = r0.token;
    // sizes: three buffers, each on a 2 MB boundary, the whole rounded to the multicast granularity
    const size_t slot = (buf_bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    mp.numDevices = (unsigned)R;
    mp.size = 3 * slot;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    CUresult cr = d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
    ok = cr == CUDA_SUCCESS && gran > 0;
    if (!ok) note("cuMulticastGetGranularity: " + d.str(cr));
    if (ok) {
        mp.size = (mp.size + gran - 1) / gran * gran;
        g.size = mp.size;
        g.device = h->device;
    }
    // stage 1: rank 0 creates the object and exports it; the others open their sockets
    int fd = -1, lsock = -1;
    if (ok && h->rank == 0) {
        cr = d.MulticastCreate(&g.mc, &mp);
        ok = cr == CUDA_SUCCESS;
        if (!ok) note("cuMulticastCreate: " + d.str(cr));
        if (ok) {
            cr = d.MemExportToShareableHandle(&fd, g.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
            ok = cr == CUDA_SUCCESS && fd >= 0;
            if (!ok) note("cuMemExportToShareableHandle: " + d.str(cr));
        }
    } else if (ok) {
        lsock = nvls_listen(token, h->rank);
        ok = lsock >= 0;
        if (!ok) note(std::string("unix socket: ") + strerror(errno));
    }
    TRY(agree(ok, nullptr, &all_ok));
    // stage 2: the descriptor travels to every other rank
    if (all_ok) {
        if (h->rank == 0) {
            for (int q = 1; q < R; ++q)
                if (!nvls_send_fd(token, q, fd)) {
                    ok = false;
                    note("could not send the multicast descriptor to rank " + std::to_string(q));
                }
        } else {
            fd = nvls_recv_fd(lsock);
            ok = fd >= 0;
            if (!ok) note("did not receive the multicast descriptor");
            if (ok) {
                cr = d.MemImportFromShareableHandle(&g.mc, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
                ok = cr == CUDA_SUCCESS;
                if (!ok) note("cuMemImportFromShareableHandle: " + d.str(cr));
            }
        }
    }
    if (fd >= 0) close(fd);
    if (lsock >= 0) close(lsock);
    if (all_ok) TRY(agree(ok, nullptr, &all_ok));
    // stage 3: every device joins the team (all must have joined before anybody binds memory)
    if (all_ok) {
        cr = d.MulticastAddDevice(g.mc, dev);
        ok = cr == CUDA_SUCCESS;
        if (!ok) note("cuMulticastAddDevice: " + d.str(cr));
        TRY(agree(ok, nullptr, &all_ok));
    }
    // stage 4: physical memory, bound to the object, mapped twice (own memory; multicast view)
    if (all_ok) {
        CUmemAllocationProp ap;
        memset(&ap, 0, sizeof ap);
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = h->device;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CUmemAccessDesc ad;
        memset(&ad, 0, sizeof ad);
        ad.location = ap.location;
        ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        auto step = [&](CUresult r, const char *what) {
            if (ok && r != CUDA_SUCCESS) {
                ok = false;
                note(std::string(what) + ": " + d.str(r));
            }
            return ok;
        };
        if (step(d.MemCreate(&g.phys, g.size, &ap, 0), "cuMemCreate") &&
            step(d.MulticastBindMem(g.mc, 0, g.phys, 0, g.size, 0), "cuMulticastBindMem")) {
            g.bound = true;
            if (step(d.MemAddressReserve(&g.uc_va, g.size, gran, 0, 0), "cuMemAddressReserve") &&
                step(d.MemMap(g.uc_va, g.size, 0, g.phys, 0), "cuMemMap")) {
                g.uc_mapped = true;
                step(d.MemSetAccess(g.uc_va, g.size, &ad, 1), "cuMemSetAccess");
            }
            if (ok && step(d.MemAddressReserve(&g.mc_va, g.size, gran, 0, 0), "cuMemAddressReserve (multicast)") &&
                step(d.MemMap(g.mc_va, g.size, 0, g.mc, 0), "cuMemMap (multicast)")) {
                g.mc_mapped = true;
                step(d.MemSetAccess(g.mc_va, g.size, &ad, 1), "cuMemSetAccess (multicast)");
            }
        }
        TRY(agree(ok, nullptr, &all_ok));
    }
    if (!all_ok) {
        note("a rank failed to set up the multicast region");
        nvls_release(g);
        return 0;
    }
    g.active = true;
    char *uc = (char *)(uintptr_t)g.uc_va, *mc = (char *)(uintptr_t)g.mc_va;
    for (int i = 0; i < 2; ++i) {
        h->hat_buf[i] = (float *)(uc + (size_t)i * slot);
        h->mc_hat[i] = (float *)(mc + (size_t)i * slot);
    }
    h->nphi_ipc = (float *)(uc + 2 * slot);
    h->mc_nphi = (float *)(mc + 2 * slot);
    return 0;
}

// Allocates the shared model buffers with cudaMalloc, exchanges their CUDA IPC handles with every rank through an
// ncclAllGather, and maps the peers' buffers.  Any failure leaves peer_ok = false (NCCL all-reduce path).
int peer_setup(lda_b200_handle *h) {
    const int R = h->nranks;
    h->peer_ok = false;
    h->peer_note.clear();
    if (R <= 1 || !h->peer_wanted) return 0;
    if (R > MAX_PEERS) {
        h->peer_note = "more than 8 ranks";
        return 0;
    }
    const size_t n = (size_t)h->W * h->Kp;
    auto soft = [&](cudaError_t e, const char *what) -> bool {
        if (e == cudaSuccess) return true;
        h->peer_note = std::string(what) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        return false;
    };
    PeerRec mine;
    memset(&mine, 0, sizeof mine);
    mine.pid = (int)getpid();
    mine.dev = h->device;
    // The three large buffers come from a multicast-backed region when NVLink multicast is available on every rank
    // (reduction and broadcast then happen in the switch, and no rank maps another rank's copy); otherwise they are
    // cudaMalloc allocations whose CUDA IPC handles the ranks exchange.  The small flag / partial-sum arrays are always
    // shared through IPC.
    h->peer_setup_det = h->deterministic;
    TRY(nvls_setup(h, n * sizeof(float)));
    const bool nvls = h->nvls.active;
    const int first_ipc = nvls ? 3 : 0;
    bool ok = true;
    if (!nvls)
        ok = soft(cudaMalloc((void **)&h->hat_buf[0], n * sizeof(float)), "cudaMalloc") &&
             soft(cudaMalloc((void **)&h->hat_buf[1], n * sizeof(float)), "cudaMalloc") &&
             soft(cudaMalloc((void **)&h->nphi_ipc, n * sizeof(float)), "cudaMalloc");
    ok = ok && soft(cudaMalloc((void **)&h->peer_flags, MAX_PEERS * sizeof(unsigned)), "cudaMalloc") &&
         soft(cudaMalloc((void **)&h->peer_nzpart, (size_t)MAX_PEERS * h->Kp * sizeof(double)), "cudaMalloc");
    void *bufs[5] = {h->hat_buf[0], h->hat_buf[1], h->nphi_ipc, h->peer_flags, h->peer_nzpart};
    if (ok) {
        CU(cudaMemsetAsync(h->hat_buf[0], 0, n * sizeof(float), h->stream));
        CU(cudaMemsetAsync(h->hat_buf[1], 0, n * sizeof(float), h->stream));
        CU(cudaMemsetAsync(h->peer_flags, 0, MAX_PEERS * sizeof(unsigned), h->stream));
        CU(cudaMemsetAsync(h->peer_nzpart, 0, (size_t)MAX_PEERS * h->Kp * sizeof(double), h->stream));
        for (int i = first_ipc; i < 5 && ok; ++i) {
            mine.ptr[i] = bufs[i];
            ok = soft(cudaIpcGetMemHandle(&mine.mh[i], bufs[i]), "cudaIpcGetMemHandle");
        }
    }
    if (!ok) mine.pid = -1;  // tell the peers this rank cannot take part
    // exchange the records (every rank must reach this point, successful or not)
    PeerRec *drec = nullptr;
    TRY(dalloc(h, &drec, (size_t)R));
    CU(cudaMemcpyAsync(drec + h->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    NC(nccl().AllGather(drec + h->rank, drec, sizeof(PeerRec), ncclChar, h->comm, h->stream));
    std::vector<PeerRec> all((size_t)R);
    CU(cudaMemcpyAsync(all.data(), drec, sizeof(PeerRec) * (size_t)R, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, &drec);
    for (int q = 0; q < R; ++q)
        if (all[(size_t)q].pid < 0) ok = false;
    PeerTable t;
    memset(&t, 0, sizeof t);
    t.rank = h->rank;
    t.nranks = R;
    for (int q = 0; q < R && ok; ++q) {
        void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int i = first_ipc; i < 5 && ok; ++i) {
            if (q == h->rank) {
                p[i] = bufs[i];
            } else if (all[(size_t)q].pid == mine.pid) {
                // same process (e.g. one goroutine per GPU): plain peer access, no IPC
                int can = 0;
                ok = soft(cudaDeviceCanAccessPeer(&can, h->device, all[(size_t)q].dev), "cudaDeviceCanAccessPeer") && can;
                if (ok) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(all[(size_t)q].dev, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) {
                        cudaGetLastError();
                        e = cudaSuccess;
                    }
                    ok = soft(e, "cudaDeviceEnablePeerAccess");
                }
                p[i] = all[(size_t)q].ptr[i];
            } else {
                ok = soft(cudaIpcOpenMemHandle(&p[i], all[(size_t)q].mh[i], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
                if (ok) h->ipc_opened.push_back(p[i]);
            }
        }
        if (!ok) break;
        if (nvls && q == h->rank) {
            p[0] = bufs[0];
            p[1] = bufs[1];
            p[2] = bufs[2];
        }
        t.hat[0][q] = (float *)p[0];
        t.hat[1][q] = (float *)p[1];
        t.nphi[q] = (float *)p[2];
        t.flags[q] = (unsigned *)p[3];
        t.nzpart[q] = (double *)p[4];
    }
    // agreement: one rank failing to map a peer must switch everybody to the fallback
    int *dflag = nullptr;
    TRY(dalloc(h, &dflag, 1));
    const int my_ok = ok ? 0 : 1;
    CU(cudaMemcpyAsync(dflag, &my_ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NC(nccl().AllReduce(dflag, dflag, 1, ncclInt, ncclSum, h->comm, h->stream));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    dfree(h, &dflag);
    if (bad) {
        if (h->peer_note.empty()) h->peer_note = "a peer could not map the shared buffers";
        peer_teardown(h);
        return 0;
    }
    // switch the model over to the shared buffers
    dfree(h, &h->nphi);
    dfree(h, &h->nphi_hat);
    h->nphi = h->nphi_ipc;
    h->nphi_hat = h->hat_buf[0];
    h->peers = t;
    h->hat_cur = 0;
    h->epoch = 0;
    h->peer_ok = true;
    return 0;
}

int ensure_rho_table(lda_b200_handle *h, int n) {
    if (n <= h->rho_cap) return 0;
    TRY(dalloc(h, &h->rho_tab, (size_t)n));
    h->rho_cap = n;
    return 0;
}

int fill_rho_table(lda_b200_handle *h, double t0, int n) {
    TRY(ensure_rho_table(h, n));
    const lda_b200_schedule &s = h->cfg.rho_theta;
    rho_table_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(s.S, s.Tau, s.Kappa, t0, n, h->rho_tab);
    KCHECK();
    return 0;
}

// nZ[k] = sum_i nPhi[i,k] (lda.go:203) and invZ
// out[k] = sum_w src[w,k] over the W x Kp matrix src, float64, blocks added in index order
int column_sums(lda_b200_handle *h, const float *src, double *out) {
    const int bs = blend_block(h->Kp), grid = h->sms * 2;
    TRY(ensure_partials(h, (size_t)grid * h->Kp));
    colsum_kernel<<<grid, bs, (size_t)bs * 32, h->stream>>>(src, h->W, h->Kp, h->partials);
    KCHECK();
    return reduce_partials(h, grid, h->Kp, out, false, h->stream);
}
int refresh_nz_from_nphi(lda_b200_handle *h) { return column_sums(h, h->nphi, h->nz); }
// invZ from nZ and c = (nPhi + eta) invZ from nPhi (after init / set_state / a change of Eta)
int refresh_inv_z(lda_b200_handle *h) {
    inv_z_kernel<<<(h->Kp + 127) / 128, 128, 0, h->stream>>>(h->nz, h->inv_z, h->K, h->Kp, h->cfg.eta * (double)h->W);
    KCHECK();
    BlendArgs ba;
    ba.nphi = h->nphi;
    ba.nphi_hat = nullptr;
    ba.cmat = h->cmat;
    ba.inv_z = h->inv_z;
    ba.W = h->W;
    ba.Kp = h->Kp;
    ba.A = 1.f;
    ba.C = 0.f;
    ba.eta = (float)h->cfg.eta;
    phi_blend_kernel<<<h->sms * 4, blend_block(h->Kp), 0, h->stream>>>(ba);
    KCHECK();
    return 0;
}

// ---- host -> device ingestion -----------------------------------------------------------------------
std::vector<Segment> make_segments(int64_t D, int64_t b, int rank, int R) {
    std::vector<Segment> s;
    if (R <= 1) {
        s.push_back({0, D, 0});
        return s;
    }
    const int64_t M = (D + b - 1) / b;
    int64_t l0 = 0;
    for (int64_t m = rank; m < M; m += R) {
        const int64_t g0 = m * b, g1 = std::min(D, g0 + b);
        s.push_back({g0, g1, l0});
        l0 += g1 - g0;
    }
    return s;
}

// ---- host <-> device copies that are fast for pageable memory too -----------------------------------------------------
bool is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

void threaded_memcpy(void *dst, const void *src, size_t bytes, int threads) {
    if (bytes < ((size_t)4 << 20) || threads <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = (bytes / threads + 4095) & ~(size_t)4095;
    for (int t = 1; t < threads; ++t) {
        const size_t off = per * t;
        if (off >= bytes) break;
        pool.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, std::min(per, bytes - off)); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (std::thread &t : pool) t.join();
}

int ensure_pin_ring(lda_b200_handle *h) {
    if (h->pin_slot[0]) return 0;
    for (int i = 0; i < lda_b200_handle::kStageSlots; ++i) {
        CU(cudaMallocHost(&h->pin_slot[i], lda_b200_handle::kStageSlotBytes));
        CU(cudaEventCreateWithFlags(&h->pin_done[i], cudaEventDisableTiming));
    }
    // staging copies between pageable caller memory and the page-locked ring scale with the number of copying threads up to
    // the host's hardware threads (a single thread moves 1-2.5 GB/s); the ranks of a multi-GPU job share them
    h->host_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, h->nranks)));
    if (const char *e = getenv("LDA_B200_HOST_THREADS")) h->host_threads = std::max(1, atoi(e));
    return 0;
}

// host -> device on the handle's stream.  Page-locked sources go straight to the copy engine; pageable ones are
// copied by host threads into a page-locked ring whose slots are sent asynchronously, so the host copy of slot i+1
// overlaps the transfer of slot i.
int h2d(lda_b200_handle *h, void *dst, const void *src, size_t bytes, cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    if (bytes == 0) return 0;
    if (!is_pageable(src) || bytes < ((size_t)1 << 20)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    TRY(ensure_pin_ring(h));
    const size_t sb = lda_b200_handle::kStageSlotBytes;
    int i = 0;
    for (size_t off = 0; off < bytes; off += sb, ++i) {
        const int slot = i % lda_b200_handle::kStageSlots;
        const size_t n = std::min(sb, bytes - off);
        CU(cudaEventSynchronize(h->pin_done[slot]));  // the slot's previous transfer has left
        threaded_memcpy(h->pin_slot[slot], (const char *)src + off, n, h->host_threads);
        CU(cudaMemcpyAsync((char *)dst + off, h->pin_slot[slot], n, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(h->pin_done[slot], st));
    }
    return 0;
}

// device -> host of `rows` rows of `width` bytes (source pitch spitch, destination pitch dpitch); returns after the data
// is in the caller's memory
int d2h_2d(lda_b200_handle *h, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
           cudaStream_t st = nullptr) {
    if (!st) st = h->stream;
    if (width == 0 || rows == 0) return 0;
    if (!is_pageable(dst)) {
        CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return 0;
    }
    TRY(ensure_pin_ring(h));
    const size_t sb = lda_b200_handle::kStageSlotBytes;
    const size_t rows_per = std::max<size_t>(1, sb / width);
    if (width > sb) {  // very wide rows: let the driver stage them
        CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return 0;
    }
    struct Pending {
        int slot;
        size_t r0, n;
    };
    std::vector<Pending> inflight;
    auto drain_one = [&]() -> int {
        const Pending p = inflight.front();
        inflight.erase(inflight.begin());
        CU(cudaEventSynchronize(h->pin_done[p.slot]));
        if (dpitch == width) {
            threaded_memcpy((char *)dst + p.r0 * dpitch, h->pin_slot[p.slot], p.n * width, h->host_threads);
        } else {
            // strided destination (a column range of the caller's K x D matrix): rows split over the host threads -- a
            // freshly allocated result is faulted in by these writes, and the faults dominate when one thread takes them
            char *d0 = (char *)dst + p.r0 * dpitch;
            const char *s0 = (const char *)h->pin_slot[p.slot];
            const int nt = (int)std::min<size_t>((size_t)std::max(1, h->host_threads), p.n);
            auto rows_of = [=](int t) {
                for (size_t r = (size_t)t; r < p.n; r += (size_t)nt) memcpy(d0 + r * dpitch, s0 + r * width, width);
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nt; ++t) pool.emplace_back(rows_of, t);
            rows_of(0);
            for (std::thread &t : pool) t.join();
        }
        return 0;
    };
    int i = 0;
    for (size_t r0 = 0; r0 < rows; r0 += rows_per, ++i) {
        const int slot = i % lda_b200_handle::kStageSlots;
        if ((int)inflight.size() == lda_b200_handle::kStageSlots) TRY(drain_one());
        const size_t n = std::min(rows_per, rows - r0);
        CU(cudaEventSynchronize(h->pin_done[slot]));  // an upload on another stream may still be reading the slot
        CU(cudaMemcpy2DAsync(h->pin_slot[slot], width, (const char *)src + r0 * spitch, spitch, width, n, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(h->pin_done[slot], st));
        inflight.push_back({slot, r0, n});
    }
    while (!inflight.empty()) TRY(drain_one());
    return 0;
}

// copies [R][K] float64 host rows of the given segments into [Dl][Kp] fp32 device rows
int upload_rows(lda_b200_handle *h, const std::vector<Segment> &segs, const double *src, int K, int Kp, float *dst) {
    const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)(h->stage_bytes / sizeof(double)) / K);
    for (const Segment &s : segs) {
        for (int64_t g = s.g0; g < s.g1; g += rows_per_chunk) {
            const int64_t n = std::min(rows_per_chunk, s.g1 - g);
            TRY(h2d(h, h->stage, src + g * K, (size_t)n * K * sizeof(double)));
            pad_rows_kernel<<<grid_for(n * Kp, 256, h->sms * 8), 256, 0, h->stream>>>((const double *)h->stage,
                                                                                    dst + (s.l0 + (g - s.g0)) * Kp, n, K, Kp);
            KCHECK();
        }
    }
    return 0;
}

// wc of the local documents [la, lb) (lda.go:476-480) and their share of the corpus total added to h->dacc[1] (lda.go:481)
int doc_word_counts(lda_b200_handle *h, Corpus &c, int64_t la, int64_t lb, double *wc64_out, cudaStream_t st) {
    const int grid = grid_for((lb - la) * 32, 256, h->sms * 8);
    TRY(ensure_partials(h, (size_t)grid));
    doc_wc_kernel<<<grid, 256, 0, st>>>(c.doc_ptr + la, c.val, (int)(lb - la), c.wc + la, wc64_out ? wc64_out + la : nullptr, h->partials);
    KCHECK();
    return reduce_partials(h, grid, 1, h->dacc + 1, true, st);
}

// Entries of the local documents [la, lb) -> c.val / c.ind on stream st: the caller's slices of the owning segments go
// through the stage buffer (host sources) or are read in place (device sources); values float64 -> fp32, word ids
// int64 -> int32 with the range check of narrow_ind_kernel.
int upload_part(lda_b200_handle *h, Corpus &c, int64_t la, int64_t lb, cudaStream_t st, bool values, bool ids) {
    const int64_t chunk = c.src_device ? ((int64_t)1 << 40) : (int64_t)(h->stage_bytes / 8);
    for (const Segment &s : c.segs) {
        const int64_t a = std::max(la, s.l0), b = std::min(lb, s.l0 + (s.g1 - s.g0));
        if (a >= b) continue;
        const int64_t p0 = c.src_indptr[s.g0 + (a - s.l0)], p1 = c.src_indptr[s.g0 + (b - s.l0)];
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0 ? !values : !ids) continue;
            const char *src = pass == 0 ? (const char *)c.src_data : (const char *)c.src_ind;
            for (int64_t p = p0; p < p1; p += chunk) {
                const int64_t n = std::min(chunk, p1 - p);
                const void *dsrc = src + p * 8;  // 8-byte source elements where a kernel can read them
                if (!c.src_device) {
                    TRY(h2d(h, h->stage, src + p * 8, (size_t)n * 8, st));
                    dsrc = h->stage;
                }
                const int64_t off = c.lp[(size_t)a] + (p - p0);
                if (pass == 0)
                    narrow_val_kernel<<<grid_for(n, 256, h->sms * 8), 256, 0, st>>>((const double *)dsrc, c.val + off, n);
                else
                    narrow_ind_kernel<<<grid_for(n, 256, h->sms * 8), 256, 0, st>>>((const int64_t *)dsrc, c.ind + off, n, c.W, h->err_flag);
                KCHECK();
            }
        }
    }
    return 0;
}

// a step's launch range of more documents than this is tiled into sub-ranges (see Corpus::sub); LDA_B200_SUB_RANGE_DOCS
// overrides it (tests exercise the sub-range path on small corpora)
int64_t fit_sub_range_docs() {
    const char *env = getenv("LDA_B200_SUB_RANGE_DOCS");
    const long long v = env ? atoll(env) : 0;
    return v > 0 ? (int64_t)v : 32768;
}

// Uploads this rank's shard of the CSC (token order preserved, ids narrowed exactly) and computes wc / N.
// src_device: ind / data are device arrays (a matrix converted on the device, see DeviceCSC); indptr is always on the host.
// range_docs > 0 (Transform / Perplexity): only the plan is made here -- local indptr, launch order per range of
// range_docs documents -- and the entries follow range by range (upload_part) while earlier ranges compute.
int upload_corpus(lda_b200_handle *h, Corpus &c, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                  const double *data, double *wc64_out /*device, optional*/, bool per_minibatch_order,
                  bool defer_ind = false, bool src_device = false, int64_t range_docs = 0) {
    Nvtx range("lda_b200: ingest corpus");
    if (src_device) defer_ind = false;  // nothing crosses PCIe: no transfer to hide behind the first launches
    if (D < 0) return fail(h, LDA_B200_EINVAL, "negative document count");
    if (!indptr) return fail(h, LDA_B200_EINVAL, "indptr is NULL");
    PhaseTimer pt(h->stream);
    free_corpus(h, c);
    c.D = D;
    c.W = W;
    c.b = h->cfg.batch_size;
    c.src_indptr = indptr;
    c.src_ind = ind;
    c.src_data = data;
    c.src_device = src_device;
    c.range_docs = range_docs;
    // the reference runs `Processes` minibatches concurrently (lda.go:487-528); here R GPUs x conc per GPU
    c.conc = per_minibatch_order ? std::min(MAX_CONCURRENT_MB, std::max(1, h->cfg.processes / h->nranks)) : 1;
    c.segs = make_segments(D, c.b, h->rank, h->nranks);
    int64_t Dl = 0;
    for (const Segment &s : c.segs) Dl += s.g1 - s.g0;
    c.Dl = Dl;
    if (Dl > 0x7fffffffLL) return fail(h, LDA_B200_EUNSUPPORTED, "more than 2^31 documents per GPU");

    std::vector<int64_t> &lp = c.lp;
    lp.assign((size_t)Dl + 1, 0);
    for (const Segment &s : c.segs) {
        for (int64_t g = s.g0; g < s.g1; ++g) {
            const int64_t n = indptr[g + 1] - indptr[g];
            if (n < 0 || indptr[g] < 0) return fail(h, LDA_B200_EINVAL, "indptr is not non-decreasing at column %lld", (long long)g);
            if (n > 0x7fffffffLL) return fail(h, LDA_B200_EUNSUPPORTED, "document %lld has more than 2^31 entries", (long long)g);
            const int64_t l = s.l0 + (g - s.g0);
            lp[l + 1] = lp[l] + n;
        }
    }
    c.nnz = lp[Dl];
    pt.lap("ingest: local indptr");
    if (c.nnz > 0 && (!ind || !data)) return fail(h, LDA_B200_EINVAL, "ind/data is NULL");
    TRY(dalloc(h, &c.doc_ptr, (size_t)Dl + 1));
    TRY(dalloc(h, &c.ind, (size_t)c.nnz));
    TRY(dalloc(h, &c.val, (size_t)c.nnz));
    TRY(dalloc(h, &c.wc, (size_t)Dl));
    CU(cudaMemcpyAsync(c.doc_ptr, lp.data(), ((size_t)Dl + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->dacc + 1, 0, sizeof(double), h->stream));
    pt.lap("ingest: cudaMalloc corpus");

    const bool streamed = range_docs > 0;
    // values first: wc and wordsInCorpus (needed by the first M-step) depend only on them
    if (!streamed) TRY(upload_part(h, c, 0, Dl, h->stream, true, !defer_ind));
    {   // Longest-first processing order inside every launch range (`conc` minibatches for Fit, range_docs documents
        // or everything for Transform), by a stable counting sort on the document length; runs on the host while the
        // copies above are in flight.
        std::vector<int> ord((size_t)Dl);
        std::vector<int> cnt;
        const int64_t range = per_minibatch_order ? c.b * c.conc : streamed ? range_docs : std::max<int64_t>(Dl, 1);
        c.sub = range;
        c.launch_range = std::min<int64_t>(range, std::max<int64_t>(Dl, 1));
        const int64_t sub_docs = fit_sub_range_docs();
        if (per_minibatch_order && range > sub_docs) c.sub = std::max<int64_t>(1, sub_docs / c.b) * c.b;
        c.sub_begin.clear();
        for (int64_t r0 = 0; r0 < Dl; r0 += range)
            for (int64_t l0 = r0; l0 < std::min(Dl, r0 + range); l0 += c.sub) c.sub_begin.push_back(l0);
        for (size_t si = 0; si < c.sub_begin.size(); ++si) {
            const int64_t l0 = c.sub_begin[si];
            const int64_t l1 = std::min({Dl, l0 + c.sub, (l0 / range + 1) * range});
            int64_t maxlen = 0;
            for (int64_t l = l0; l < l1; ++l) maxlen = std::max(maxlen, lp[l + 1] - lp[l]);
            cnt.assign((size_t)maxlen + 2, 0);
            for (int64_t l = l0; l < l1; ++l) cnt[(size_t)(maxlen - (lp[l + 1] - lp[l])) + 1]++;   // key: descending length
            for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
            for (int64_t l = l0; l < l1; ++l) ord[(size_t)(l0 + cnt[(size_t)(maxlen - (lp[l + 1] - lp[l]))]++)] = (int)l;
        }
        TRY(dalloc(h, &c.order, (size_t)Dl));
        CU(cudaMemcpyAsync(c.order, ord.data(), (size_t)Dl * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));  // ord is a host temporary
    }
    pt.lap("ingest: H2D + narrow + length sort");
    if (streamed) return 0;
    if (Dl > 0) TRY(doc_word_counts(h, c, 0, Dl, wc64_out, h->stream));
    if (h->nranks > 1) NC(nccl().AllReduce(h->dacc + 1, h->dacc + 1, 1, ncclDouble, ncclSum, h->comm, h->stream));
    CU(cudaMemcpyAsync(h->h_pinned, h->dacc + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_pinned + 1, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    c.n_global = h->h_pinned[0];
    pt.lap("ingest: wc + N");
    int bad;
    memcpy(&bad, h->h_pinned + 1, sizeof(int));
    if (bad) return fail(h, LDA_B200_EINVAL, "a row index lies outside [0, %lld)", (long long)W);
    if (defer_ind) {
        // The stage buffer now belongs to the copy stream: the ids of sub-range i are sent, narrowed and signalled with
        // range_ev[i]; the first iteration's launches wait on their own sub-range only.
        const int64_t range = c.b * c.conc;
        CU(cudaEventRecord(h->copy_ev, h->stream));
        CU(cudaStreamWaitEvent(h->copy_stream, h->copy_ev, 0));
        for (size_t si = 0; si < c.sub_begin.size(); ++si) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c.range_ev.push_back(e);
        }
        auto feed = [h, &c, range, Dl](UploadJob *job) -> int {
            for (size_t si = 0; si < c.sub_begin.size(); ++si) {
                const int64_t l0 = c.sub_begin[si];
                TRY(upload_part(h, c, l0, std::min({Dl, l0 + c.sub, (l0 / range + 1) * range}), h->copy_stream, false, true));
                CU(cudaEventRecord(c.range_ev[si], h->copy_stream));
                if (job) job->recorded.store((int)si + 1, std::memory_order_release);
            }
            return 0;
        };
        if (!c.src_device && c.nnz > 0 && is_pageable(c.src_ind)) {
            // pageable source: every 64 MB slot is a host memcpy before its transfer; a worker thread does that while
            // the caller's thread goes on to queue the launches (c lives in the handle for the whole fit)
            c.job = std::make_shared<UploadJob>();
            UploadJob *job = c.job.get();
            const int n_sub = (int)c.sub_begin.size();
            job->worker = std::thread([h, job, feed, n_sub]() {
                cudaSetDevice(h->device);
                std::string saved = h->err;
                job->rc = feed(job);
                if (job->rc) {
                    job->err = h->err;  // fail() wrote the text into the handle; hand it over and let the caller report it
                    h->err = saved;
                }
                job->recorded.store(n_sub, std::memory_order_release);  // nobody waits for events that will never come
            });
        } else {
            TRY(feed(nullptr));
        }
        c.ind_deferred = true;
        pt.lap("ingest: word ids queued on the copy stream");
    }
    return 0;
}

// initial nTheta / theta (lda.go:472-475, 422-426): either caller-provided values or PCG draws on the device
int init_theta(lda_b200_handle *h, Corpus &c, const double *theta0, lda_b200_pcg *rnd, uint64_t first_draw) {
    TRY(dalloc(h, &c.theta, (size_t)c.Dl * h->Kp));
    if (theta0) return upload_rows(h, c.segs, theta0, h->K, h->Kp, c.theta);
    if (!rnd) return fail(h, LDA_B200_EINVAL, "neither initial values nor a PCG state were given");
    if (c.Dl > 0) {
        const int64_t work = c.Dl * ((h->K + PCG_CH - 1) / PCG_CH);
        pcg_fill_rows_kernel<<<grid_for(work, 128, h->sms * 16), 128, 0, h->stream>>>(
            rnd->high, rnd->low, first_draw, (uint64_t)c.D * (uint64_t)h->K, c.Dl, h->K, h->Kp, c.b, h->nranks, h->rank, c.theta);
        KCHECK();
    }
    return 0;
}

void advance(lda_b200_pcg *rnd, uint64_t draws) {
    Pcg p;
    p.s = mk128(rnd->high, rnd->low);
    pcg_jump(p, draws);
    rnd->high = (uint64_t)(p.s >> 64);
    rnd->low = (uint64_t)p.s;
}

// perplexity (lda.go:392-408) of the corpus under theta and the current nPhi, with denominator `sum`
int eval_perplexity(lda_b200_handle *h, const Corpus &c, double sum, double *out) {
    Nvtx range("lda_b200: perplexity");
    TRY(column_sums(h, h->nphi, h->colsum));
    reciprocal_kernel<<<(h->Kp + 127) / 128, 128, 0, h->stream>>>(h->colsum, h->inv_colsum, h->K, h->Kp);
    KCHECK();
    CU(cudaMemsetAsync(h->dacc, 0, sizeof(double), h->stream));
    PerpArgs a;
    a.doc_ptr = c.doc_ptr;
    a.ind = c.ind;
    a.val = c.val;
    a.theta = c.theta;
    a.nphi = h->nphi;
    a.inv_colsum = h->inv_colsum;
    a.partials = nullptr;  // set by launch_perplexity
    a.n_docs = (int)c.Dl;
    a.Kp = h->Kp;
    TRY(launch_perplexity(h, a));
    if (h->nranks > 1) NC(nccl().AllReduce(h->dacc, h->dacc, 1, ncclDouble, ncclSum, h->comm, h->stream));
    CU(cudaMemcpyAsync(h->h_pinned, h->dacc, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = exp2(-h->h_pinned[0] / sum);
    return 0;
}

// normaliseTheta + transposed float64 copy (lda.go:452, 541) into the caller's K x D matrix
int write_theta_out(lda_b200_handle *h, const Corpus &c, double *theta_out) {
    if (!theta_out || c.Dl == 0) return 0;
    PhaseTimer pt(h->stream);
    double *dout = nullptr;
    TRY(dalloc(h, &dout, (size_t)h->K * c.Dl));
    pt.lap("theta_out: cudaMalloc");
    transpose_out_kernel<true><<<(unsigned)((c.Dl + 31) / 32), 256, 0, h->stream>>>(c.theta, c.Dl, h->K, h->Kp, nullptr, dout, c.Dl);
    h->launches++;
    int rc = 0;
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) {
        for (const Segment &s : c.segs) {
            rc = d2h_2d(h, theta_out + s.g0, (size_t)c.D * sizeof(double), dout + s.l0, (size_t)c.Dl * sizeof(double),
                        (size_t)(s.g1 - s.g0) * sizeof(double), (size_t)h->K);
            if (rc) break;
        }
    }
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(h->stream);
    pt.lap("theta_out: normalise + D2H");
    cudaFreeAsync(dout, h->stream);
    pt.lap("theta_out: free");
    if (rc) return rc;
    if (e != cudaSuccess) return fail(h, LDA_B200_ECUDA, "theta read-back failed: %s", cudaGetErrorString(e));
    return 0;
}

// During the last iteration of a fit the nTheta rows of a launch are final as soon as that launch has finished:
// normalise + transpose them (lda.go:541) right behind the launch and copy them to the caller's K x D matrix on the
// copy stream, so the device->host transfer overlaps the remaining launches of the iteration.
int stream_theta_range(lda_b200_handle *h, const Corpus &c, int64_t l0, int64_t l1) {
    if (l1 <= l0) return 0;
    transpose_out_kernel<true><<<(unsigned)((l1 - l0 + 31) / 32), 256, 0, h->stream>>>(c.theta + l0 * h->Kp, l1 - l0, h->K, h->Kp,
                                                                                      nullptr, h->dout + l0, c.Dl);
    KCHECK();
    CU(cudaEventRecord(h->copy_ev, h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->copy_ev, 0));
    for (const Segment &s : c.segs) {
        const int64_t a = std::max(l0, s.l0), b = std::min(l1, s.l0 + (s.g1 - s.g0));
        if (a >= b) continue;
        CU(cudaMemcpy2DAsync(h->pending_theta_out + s.g0 + (a - s.l0), (size_t)c.D * sizeof(double), h->dout + a,
                             (size_t)c.Dl * sizeof(double), (size_t)(b - a) * sizeof(double), (size_t)h->K,
                             cudaMemcpyDeviceToHost, h->copy_stream));
    }
    return 0;
}

DocsArgs docs_args(lda_b200_handle *h, const Corpus &c, int passes) {
    DocsArgs a;
    a.doc_ptr = c.doc_ptr;
    a.ind = c.ind;
    a.val = c.val;
    a.wc = c.wc;
    a.theta = c.theta;
    a.cmat = h->cmat;
    a.nphi_hat = h->nphi_hat;
    a.log1m_rho = h->rho_tab;
    a.passes_out = nullptr;
    a.doc_counter = h->doc_counter;
    a.doc_begin = 0;
    a.doc_end = (int)c.Dl;
    a.K = h->K;
    a.Kp = h->Kp;
    a.passes = passes;
    a.change_freq = h->cfg.change_evaluation_frequency;
    a.change_tol = (float)h->cfg.mean_change_tolerance;
    a.alpha = (float)h->cfg.alpha;
    a.eta = (float)h->cfg.eta;
    for (float &x : a.hat_scale) x = 0.f;
    a.mb_docs = (int)std::max<int64_t>(1, c.b);
    a.mb_origin = 0;
    a.hat_fx = nullptr;
    a.fx_scale = 1.f;
    a.pair = c.launch_range > 0 && c.launch_range <= h->pair_docs_max ? 1 : 0;
    return a;
}

inline double rho_calc(const lda_b200_schedule &s, double t) { return s.S / pow(s.Tau + t, s.Kappa); }  // lda.go:30-32

// unNormalisedTransform (lda.go:420-437) and, when theta_out is given, normaliseTheta + the transposed float64 copy of
// lda.go:452, over a corpus whose plan is in place (upload_corpus).  A corpus planned with range_docs > 0 flows through
// three streams range by range -- raw entries in (copy stream: copy engine only), narrowing + wc + burnInDoc kernel +
// normalise/transpose (main stream), result out (down stream) -- so the host<->device transfers
// of neighbouring ranges hide behind the kernel of the current one; otherwise it is one launch over all documents.  Documents are independent (lda.go:429-435), so the split changes
// nothing but the longest-first order inside a launch.
// dev_out / dev_ld (optional): a device matrix that receives the K x Dl result (column l at dev_out + l, row pitch dev_ld)
// instead of a host matrix -- the index's storage, see lda_b200_transform_into_index.
int run_transform(lda_b200_handle *h, Corpus &c, double rho_theta_t, double *theta_out, int32_t *passes_out_host,
                  double *dev_out = nullptr, int64_t dev_ld = 0) {
    const int T = h->cfg.transformation_passes;
    TRY(fill_rho_table(h, rho_theta_t, T + 1));
    const bool streamed = c.range_docs > 0;
    const int64_t Dl = c.Dl, blk = streamed ? c.range_docs : std::max<int64_t>(Dl, 1);
    int32_t *dpass = nullptr;
    double *dout = nullptr;
    int64_t *raw[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev;  // per range: entries uploaded, narrowed, result ready
    auto new_event = [&](cudaEvent_t *e) -> int {
        CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        ev.push_back(*e);
        return 0;
    };
    struct Range {
        int64_t la, lb;
        cudaEvent_t done;
    };
    auto download = [&](const Range &r) -> int {  // columns of the documents [la, lb) of the K x D result
        CU(cudaStreamWaitEvent(h->down_stream, r.done, 0));
        if (getenv("LDA_B200_TIMING")) {
            const double t0 = PhaseTimer::now();
            cudaEventSynchronize(r.done);
            fprintf(stderr, "[lda_b200] transform download waited %8.2f ms for the range's kernels\n", 1e3 * (PhaseTimer::now() - t0));
        }
        for (const Segment &s : c.segs) {
            const int64_t a = std::max(r.la, s.l0), b = std::min(r.lb, s.l0 + (s.g1 - s.g0));
            if (a >= b) continue;
            TRY(d2h_2d(h, theta_out + s.g0 + (a - s.l0), (size_t)c.D * sizeof(double), dout + a, (size_t)Dl * sizeof(double),
                       (size_t)(b - a) * sizeof(double), (size_t)h->K, h->down_stream));
        }
        return 0;
    };
    auto run = [&]() -> int {
        if (passes_out_host) TRY(dalloc(h, &dpass, (size_t)Dl));
        if (theta_out && !dev_out && Dl > 0) TRY(dalloc(h, &dout, (size_t)h->K * Dl));
        double *sink = dev_out ? dev_out : dout;       // where normalise + transpose writes
        const int64_t sink_ld = dev_out ? dev_ld : Dl;
        if (streamed) {  // the copy stream starts once everything queued so far has run
            CU(cudaEventRecord(h->copy_ev, h->stream));
            CU(cudaStreamWaitEvent(h->copy_stream, h->copy_ev, 0));
        }
        DocsArgs a = docs_args(h, c, T);
        a.passes_out = dpass;
        Range prev{0, 0, nullptr};
        bool have_prev = false;
        // Host sources: the copy engine alone brings a range's raw entries (int64 ids, float64 values) into one of two
        // staging buffers; the narrowing and wc kernels run on the main stream in front of the range's launch.  (Run
        // on the copy stream they would wait for SM resources until the previous range's persistent kernel drains,
        // and the transfer they gate would not overlap anything.)
        const bool raw_staging = streamed && !c.src_device;
        int64_t max_nnz = 0;
        for (int64_t la = 0; la < Dl; la += blk) max_nnz = std::max(max_nnz, c.lp[(size_t)std::min(Dl, la + blk)] - c.lp[(size_t)la]);
        if (raw_staging)
            for (int i = 0; i < 2; ++i) TRY(dalloc(h, &raw[i], (size_t)2 * std::max<int64_t>(max_nnz, 1)));
        std::vector<cudaEvent_t> narrowed;  // per range: its staging buffer may be overwritten
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ktime;
        const bool trace = getenv("LDA_B200_TIMING") != nullptr;
        const double t_start = PhaseTimer::now();
        auto stamp = [&](const char *what, int64_t range) {
            if (trace) fprintf(stderr, "[lda_b200] transform range %2lld %-22s at %8.2f ms\n", (long long)range, what, 1e3 * (PhaseTimer::now() - t_start));
        };
        int64_t r = 0;
        for (int64_t la = 0; la < Dl; la += blk, ++r) {
            const int64_t lb = std::min(Dl, la + blk);
            const int64_t n_r = c.lp[(size_t)lb] - c.lp[(size_t)la];
            Nvtx t_range("lda_b200: transform range");
            cudaStream_t cs = h->stream;
            if (raw_staging) {
                int64_t *ids_raw = raw[r & 1];
                double *val_raw = reinterpret_cast<double *>(raw[r & 1] + max_nnz);
                if (r >= 2) CU(cudaStreamWaitEvent(h->copy_stream, narrowed[(size_t)(r - 2)], 0));
                for (const Segment &sg : c.segs) {  // the caller's slices of the owning segments, back to back
                    const int64_t sa = std::max(la, sg.l0), sb = std::min(lb, sg.l0 + (sg.g1 - sg.g0));
                    if (sa >= sb) continue;
                    const int64_t p0 = c.src_indptr[sg.g0 + (sa - sg.l0)], p1 = c.src_indptr[sg.g0 + (sb - sg.l0)];
                    const int64_t off = c.lp[(size_t)sa] - c.lp[(size_t)la];
                    TRY(h2d(h, ids_raw + off, c.src_ind + p0, (size_t)(p1 - p0) * 8, h->copy_stream));
                    TRY(h2d(h, val_raw + off, c.src_data + p0, (size_t)(p1 - p0) * 8, h->copy_stream));
                }
                cudaEvent_t up;
                TRY(new_event(&up));
                CU(cudaEventRecord(up, h->copy_stream));
                CU(cudaStreamWaitEvent(cs, up, 0));
                stamp("upload issued", r);
                if (n_r > 0) {
                    narrow_val_kernel<<<grid_for(n_r, 256, h->sms * 8), 256, 0, cs>>>(val_raw, c.val + c.lp[(size_t)la], n_r);
                    KCHECK();
                    narrow_ind_kernel<<<grid_for(n_r, 256, h->sms * 8), 256, 0, cs>>>(ids_raw, c.ind + c.lp[(size_t)la], n_r, c.W, h->err_flag);
                    KCHECK();
                }
                cudaEvent_t nd;
                TRY(new_event(&nd));
                CU(cudaEventRecord(nd, cs));
                narrowed.push_back(nd);
            } else if (streamed) {
                TRY(upload_part(h, c, la, lb, cs, true, true));  // device source: read in place
            }
            if (streamed) TRY(doc_word_counts(h, c, la, lb, nullptr, cs));
            CU(cudaMemsetAsync(a.doc_counter, 0, sizeof(int), cs));
            a.doc_begin = (int)la;
            a.doc_end = (int)lb;
            cudaEvent_t k0 = nullptr, k1 = nullptr;
            if (trace) {
                CU(cudaEventCreate(&k0));
                CU(cudaEventCreate(&k1));
                CU(cudaEventRecord(k0, cs));
            }
            TRY(launch_docs<false>(h, a, c.order, cs));
            if (trace) {
                CU(cudaEventRecord(k1, cs));
                ktime.push_back({k0, k1});
            }
            if (sink) {
                transpose_out_kernel<true><<<(unsigned)((lb - la + 31) / 32), 256, 0, cs>>>(c.theta + la * h->Kp, lb - la, h->K, h->Kp,
                                                                                                  nullptr, sink + la, sink_ld);
                KCHECK();
            }
            if (dout) {
                Range cur{la, lb, nullptr};
                TRY(new_event(&cur.done));
                CU(cudaEventRecord(cur.done, cs));
                stamp("kernels launched", r);
                if (have_prev) TRY(download(prev));  // while this range computes
                stamp("previous downloaded", r);
                prev = cur;
                have_prev = true;
            }
        }
        if (have_prev) TRY(download(prev));
        if (streamed) {  // wordCount of Perplexity (lda.go:372-385) and ids that were out of range
            if (h->nranks > 1) NC(nccl().AllReduce(h->dacc + 1, h->dacc + 1, 1, ncclDouble, ncclSum, h->comm, h->stream));
            CU(cudaMemcpyAsync(h->h_pinned, h->dacc + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(h->h_pinned + 1, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        }
        if (passes_out_host) {
            std::vector<int32_t> tmp((size_t)Dl);
            CU(cudaMemcpyAsync(tmp.data(), dpass, (size_t)Dl * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            for (const Segment &s : c.segs)
                for (int64_t g = s.g0; g < s.g1; ++g) passes_out_host[g] = tmp[(size_t)(s.l0 + g - s.g0)];
        }
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaStreamSynchronize(h->down_stream));
        for (size_t i = 0; i < ktime.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ktime[i].first, ktime[i].second);
            fprintf(stderr, "[lda_b200] transform range %2zu kernel %8.2f ms\n", i, ms);
            cudaEventDestroy(ktime[i].first);
            cudaEventDestroy(ktime[i].second);
        }
        if (streamed) {
            c.n_global = h->h_pinned[0];
            int bad;
            memcpy(&bad, h->h_pinned + 1, sizeof(int));
            if (bad) return fail(h, LDA_B200_EINVAL, "a row index lies outside [0, %lld)", (long long)c.W);
        }
        return 0;
    };
    const int rc = run();
    if (rc) {  // leave no work behind that still refers to the caller's arrays or to buffers freed below
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        cudaStreamSynchronize(h->down_stream);
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (dpass) cudaFreeAsync(dpass, h->stream);
    if (dout) cudaFreeAsync(dout, h->stream);
    for (int i = 0; i < 2; ++i)
        if (raw[i]) cudaFreeAsync(raw[i], h->stream);
    return rc;
}

// documents per range of the Transform pipeline: enough for ~14 documents per resident warp, small enough that the
// first upload and the last download (the parts no kernel hides) stay short
constexpr int64_t kTransformRangeDocs = 65536;

// ---- CSR / dense -> CSC on the device (sparse.ToCSC() lda.go:464-466; dense scan utils.go:51-57) -------------------
void free_device_csc(lda_b200_handle *h, DeviceCSC &m) {
    dfree(h, &m.indptr);
    dfree(h, &m.ind);
    dfree(h, &m.data);
    m = DeviceCSC();
}

// The radix sort on the column (document) id is stable, so inside a document the rows stay in ascending order --
// exactly what the reference's host-side conversion produces and what fixes the token order of the path.
int device_csr_to_csc(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *row_ptr, const int64_t *col_ind,
                      const double *data, DeviceCSC &out) {
    if (W < 0 || D < 0 || !row_ptr) return fail(h, LDA_B200_EINVAL, "bad CSR arguments");
    const int64_t nnz = row_ptr[W] - row_ptr[0];
    if (row_ptr[0] != 0) return fail(h, LDA_B200_EINVAL, "row_ptr[0] must be 0");
    for (int64_t r = 0; r < W; ++r)
        if (row_ptr[r + 1] < row_ptr[r]) return fail(h, LDA_B200_EINVAL, "row_ptr is not non-decreasing at row %lld", (long long)r);
    if (nnz >= (1LL << 31) || D >= (1LL << 31) || W >= (1LL << 31))
        return fail(h, LDA_B200_EUNSUPPORTED, "CSR with 2^31 or more entries / rows / columns");
    if (nnz > 0 && (!col_ind || !data)) return fail(h, LDA_B200_EINVAL, "NULL CSR array");
    int64_t *d_rowptr = nullptr;
    int32_t *d_cols = nullptr, *d_cols2 = nullptr, *d_rows = nullptr;
    uint32_t *d_perm = nullptr, *d_perm2 = nullptr;
    double *d_data = nullptr;
    void *d_tmp = nullptr;
    out.W = W;
    out.D = D;
    out.nnz = nnz;
    out.h_indptr.assign((size_t)D + 1, 0);
    auto run = [&]() -> int {
        TRY(dalloc(h, &d_rowptr, (size_t)W + 1));
        TRY(dalloc(h, &out.indptr, (size_t)D + 1));
        TRY(dalloc(h, &d_cols, (size_t)nnz));
        TRY(dalloc(h, &d_cols2, (size_t)nnz));
        TRY(dalloc(h, &d_rows, (size_t)nnz));
        TRY(dalloc(h, &d_perm, (size_t)nnz));
        TRY(dalloc(h, &d_perm2, (size_t)nnz));
        TRY(dalloc(h, &d_data, (size_t)nnz));
        TRY(dalloc(h, &out.data, (size_t)nnz));
        TRY(dalloc(h, &out.ind, (size_t)nnz));
        CU(cudaMemcpyAsync(d_rowptr, row_ptr, ((size_t)W + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemsetAsync(h->err_flag, 0, sizeof(int), h->stream));
        const int64_t chunk = (int64_t)(h->stage_bytes / 8);
        for (int64_t p = 0; p < nnz; p += chunk) {
            const int64_t n = std::min(chunk, nnz - p);
            TRY(h2d(h, h->stage, col_ind + p, (size_t)n * 8));
            narrow_ind_kernel<<<grid_for(n, 256, h->sms * 8), 256, 0, h->stream>>>((const int64_t *)h->stage, d_cols + p, n, D, h->err_flag);
            KCHECK();
        }
        TRY(h2d(h, d_data, data, (size_t)nnz * sizeof(double)));
        if (nnz > 0) {
            csr_expand_rows_kernel<<<grid_for(W * 32, 256, h->sms * 8), 256, 0, h->stream>>>(d_rowptr, W, d_rows);
            KCHECK();
            iota_kernel<<<grid_for(nnz, 256, h->sms * 8), 256, 0, h->stream>>>(d_perm, nnz);
            KCHECK();
            int bits = 1;
            while ((1LL << bits) < D) ++bits;
            size_t tmp_bytes = 0;
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_cols, d_cols2, d_perm, d_perm2, (int)nnz, 0, bits, h->stream));
            CU(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, h->stream));
            CU(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_cols, d_cols2, d_perm, d_perm2, (int)nnz, 0, bits, h->stream));
            h->launches += 4;
            csc_gather_kernel<<<grid_for(nnz, 256, h->sms * 8), 256, 0, h->stream>>>(d_perm2, d_rows, d_data, nnz, out.ind, out.data);
            KCHECK();
        }
        csc_indptr_kernel<<<grid_for(D + 1, 256, h->sms * 8), 256, 0, h->stream>>>(d_cols2, nnz, D, out.indptr);
        KCHECK();
        CU(cudaMemcpyAsync(out.h_indptr.data(), out.indptr, ((size_t)D + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (bad) return fail(h, LDA_B200_EINVAL, "a column index lies outside [0, %lld)", (long long)D);
        return 0;
    };
    const int rc = run();
    dfree(h, &d_rowptr);
    dfree(h, &d_cols);
    dfree(h, &d_cols2);
    dfree(h, &d_rows);
    dfree(h, &d_perm);
    dfree(h, &d_perm2);
    dfree(h, &d_data);
    if (d_tmp) cudaFreeAsync(d_tmp, h->stream);
    if (rc) free_device_csc(h, out);
    return rc;
}

// dense mat.Matrix (rows x cols, row-major float64): every column lists its non-zero rows in ascending order, entries
// that are exactly 0 skipped (utils.go:51-57)
int device_dense_to_csc(lda_b200_handle *h, int64_t W, int64_t D, const double *a, DeviceCSC &out) {
    if (W < 0 || D < 0) return fail(h, LDA_B200_EINVAL, "bad dense matrix shape");
    if (W > 0 && D > 0 && !a) return fail(h, LDA_B200_EINVAL, "dense data is NULL");
    if (W >= (1LL << 31) || D >= (1LL << 31)) return fail(h, LDA_B200_EUNSUPPORTED, "dense matrix with 2^31 or more rows / columns");
    double *d_a = nullptr;
    int64_t *d_cnt = nullptr;
    void *d_tmp = nullptr;
    out.W = W;
    out.D = D;
    out.h_indptr.assign((size_t)D + 1, 0);
    auto run = [&]() -> int {
        TRY(dalloc(h, &d_a, (size_t)W * D));
        TRY(dalloc(h, &d_cnt, (size_t)D + 1));
        TRY(dalloc(h, &out.indptr, (size_t)D + 1));
        TRY(h2d(h, d_a, a, (size_t)W * D * sizeof(double)));
        dense_col_count_kernel<<<grid_for(D + 1, 128, h->sms * 16), 128, 0, h->stream>>>(d_a, W, D, d_cnt);
        KCHECK();
        size_t tmp_bytes = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, out.indptr, (int)(D + 1), h->stream));
        CU(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, h->stream));
        CU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, out.indptr, (int)(D + 1), h->stream));
        h->launches += 1;
        CU(cudaMemcpyAsync(out.h_indptr.data(), out.indptr, ((size_t)D + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        out.nnz = out.h_indptr[(size_t)D];
        TRY(dalloc(h, &out.ind, (size_t)out.nnz));
        TRY(dalloc(h, &out.data, (size_t)out.nnz));
        if (out.nnz > 0) {
            dense_col_fill_kernel<<<grid_for(D, 128, h->sms * 16), 128, 0, h->stream>>>(d_a, W, D, out.indptr, out.ind, out.data);
            KCHECK();
        }
        return 0;
    };
    const int rc = run();
    dfree(h, &d_a);
    dfree(h, &d_cnt);
    if (d_tmp) cudaFreeAsync(d_tmp, h->stream);
    if (rc) free_device_csc(h, out);
    return rc;
}

// the caller's matrix as the CSC the path consumes: host arrays as they are (CSC), or converted on the device
struct MatrixView {
    int64_t W = 0, D = 0;
    const int64_t *indptr = nullptr, *ind = nullptr;
    const double *data = nullptr;
    bool on_device = false;
    DeviceCSC owned;
};

int open_matrix(lda_b200_handle *h, const lda_b200_matrix *m, MatrixView &v) {
    if (!m) return fail(h, LDA_B200_EINVAL, "matrix is NULL");
    CU(cudaSetDevice(h->device));
    v.W = m->rows;
    v.D = m->cols;
    switch (m->format) {
        case LDA_B200_CSC:
            v.indptr = m->ptr;
            v.ind = m->ind;
            v.data = m->data;
            return 0;
        case LDA_B200_CSR:
            TRY(device_csr_to_csc(h, m->rows, m->cols, m->ptr, m->ind, m->data, v.owned));
            break;
        case LDA_B200_DENSE:
            TRY(device_dense_to_csc(h, m->rows, m->cols, m->data, v.owned));
            break;
        default:
            return fail(h, LDA_B200_EINVAL, "unknown matrix format %d", (int)m->format);
    }
    v.indptr = v.owned.h_indptr.data();
    v.ind = v.owned.ind;
    v.data = v.owned.data;
    v.on_device = true;
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

void lda_b200_default_config(int32_t k, lda_b200_config *c) {
    if (!c) return;
    c->K = k;
    c->iterations = 1000;
    c->batch_size = 100;
    c->burn_in_passes = 1;
    c->transformation_passes = 500;
    c->perplexity_evaluation_frequency = 30;
    c->change_evaluation_frequency = 30;
    c->processes = (int32_t)std::max(1u, std::thread::hardware_concurrency());  // lda.go:185: runtime.GOMAXPROCS(0)
    c->perplexity_tolerance = 1e-2;
    c->mean_change_tolerance = 1e-5;
    c->alpha = 0.1;
    c->eta = 0.01;
    c->rho_phi = {10, 1000, 0.9};
    c->rho_theta = {1, 10, 0.9};
}

const char *lda_b200_last_error(const lda_b200_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int lda_b200_create(const lda_b200_config *cfg, int32_t device, lda_b200_handle **out) {
    lda_b200_handle *h = nullptr;
    if (!out) return fail(h, LDA_B200_EINVAL, "out is NULL");
    *out = nullptr;
    TRY(validate_config(h, cfg));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, LDA_B200_ECUDA, "no CUDA device available (%s); liblda_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(h, LDA_B200_EINVAL, "device %d out of range [0,%d)", device, ndev);
    lda_b200_handle *nh = new lda_b200_handle();
    nh->cfg = *cfg;
    nh->device = device;
    auto init = [&](lda_b200_handle *h) -> int {
        CU(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        h->sms = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&h->copy_ev, cudaEventDisableTiming));
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        TRY(dalloc(h, &h->doc_counter, 1));
        TRY(dalloc(h, &h->err_flag, 1));
        TRY(dalloc(h, &h->dacc, 2));
        CU(cudaMemsetAsync(h->doc_counter, 0, sizeof(int), h->stream));
        auto env_gd = [](const char *name, int dflt) {
            const char *e = getenv(name);
            const int v = e ? atoi(e) : dflt;
            return (v == 1 || v == 2 || v == 4) ? v : dflt;
        };
        if (const char *e = getenv("LDA_B200_PAIR_MAX_DOCS")) h->pair_docs_max = atoll(e);
        h->docs_per_warp_fit = env_gd("LDA_B200_DOCS_PER_WARP_FIT", h->docs_per_warp_fit);
        h->docs_per_warp_transform = env_gd("LDA_B200_DOCS_PER_WARP_TRANSFORM", h->docs_per_warp_transform);
        h->stage_bytes = (size_t)256 << 20;
        CU(cudaMalloc(&h->stage, h->stage_bytes));
        CU(cudaMallocHost((void **)&h->h_pinned, 64));
        return 0;
    };
    int rc = init(nh);
    if (rc) {
        g_create_error = nh->err;
        lda_b200_destroy(nh);
        return rc;
    }
    *out = nh;
    return 0;
}

void lda_b200_destroy(lda_b200_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    peer_teardown(h);
    if (h->comm) nccl().CommDestroy(h->comm);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->down_stream) cudaStreamSynchronize(h->down_stream);
    dfree(h, &h->dout);
    free_corpus(h, h->fitc);
    dfree(h, &h->nphi);
    dfree(h, &h->nphi_hat);
    dfree(h, &h->cmat);
    dfree(h, &h->nz);
    dfree(h, &h->nzhat);
    dfree(h, &h->colsum);
    dfree(h, &h->inv_z);
    dfree(h, &h->inv_colsum);
    dfree(h, &h->doc_counter);
    dfree(h, &h->err_flag);
    dfree(h, &h->dacc);
    dfree(h, &h->rho_tab);
    dfree(h, &h->partials);
    dfree(h, &h->hat_fx);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stage) cudaFree(h->stage);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    for (EventPair &p : h->ev_pool) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (int i = 0; i < lda_b200_handle::kStageSlots; ++i) {
        if (h->pin_slot[i]) cudaFreeHost(h->pin_slot[i]);
        if (h->pin_done[i]) cudaEventDestroy(h->pin_done[i]);
    }
    if (h->copy_ev) cudaEventDestroy(h->copy_ev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->down_stream) cudaStreamDestroy(h->down_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int lda_b200_set_config(lda_b200_handle *h, const lda_b200_config *cfg) {
    if (!h) return LDA_B200_EINVAL;
    TRY(validate_config(h, cfg));
    if (h->fit_open && cfg->K != h->cfg.K) return fail(h, LDA_B200_EINVAL, "K cannot change during a fit");
    const bool eta_changed = cfg->eta != h->cfg.eta;
    h->cfg = *cfg;
    if (h->fitted && cfg->K != h->K) {
        // nPhi / nZ are laid out for the old K: every later call sizes its buffers by the new one.  The model is dropped,
        // so Transform / Perplexity / Components / get_state fail with ESTATE until the next Fit or set_state.
        h->fitted = false;
    }
    if (eta_changed && h->fitted && !h->fit_open && h->K == cfg->K) {
        CU(cudaSetDevice(h->device));
        TRY(refresh_inv_z(h));  // invZ and c = (nPhi + eta) invZ depend on Eta
    }
    return 0;
}

int lda_b200_set_deterministic(lda_b200_handle *h, int32_t on) {
    if (!h) return LDA_B200_EINVAL;
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    h->deterministic = on != 0;
    return 0;
}

int lda_b200_set_profiling(lda_b200_handle *h, int32_t on) {
    if (!h) return LDA_B200_EINVAL;
    h->profiling = on != 0;
    return 0;
}

int lda_b200_comm_unique_id(void *id128) {
    lda_b200_handle *h = nullptr;
    if (!id128) return fail(h, LDA_B200_EINVAL, "id buffer is NULL");
    if (!nccl().load()) return fail(h, LDA_B200_ENCCL, "%s", nccl().load_error);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return 0;
}

int lda_b200_comm_init(lda_b200_handle *h, int32_t rank, int32_t nranks, const void *id128) {
    if (!h) return LDA_B200_EINVAL;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, LDA_B200_EINVAL, "bad rank %d of %d", rank, nranks);
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    CU(cudaSetDevice(h->device));
    peer_teardown(h);
    if (h->comm) {
        nccl().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    h->rank = 0;
    h->nranks = 1;
    if (nranks == 1) return 0;
    if (!id128) return fail(h, LDA_B200_EINVAL, "id is NULL");
    if (!nccl().load()) return fail(h, LDA_B200_ENCCL, "%s", nccl().load_error);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NC(nccl().CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank;
    h->nranks = nranks;
    const char *env = getenv("LDA_B200_PEER_MEMORY");  // "0" forces the ncclAllReduce + blend path
    h->peer_wanted = !(env && env[0] == '0');
    env = getenv("LDA_B200_NVLS");  // "0" keeps the peer-memory kernel on unicast loads / stores (no multicast object)
    h->nvls_wanted = !(env && env[0] == '0');
    return 0;
}

// ---- FitTransform, lda.go:463-542 ----------------------------------------------------------------------
static int fit_begin_impl(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                          const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd,
                          lda_b200_counters *ctr, bool warm, bool defer_ind = false, bool src_device = false) {
    if (!h) return LDA_B200_EINVAL;
    if (!ctr) return fail(h, LDA_B200_EINVAL, "counters are NULL");
    if (warm && (!h->fitted || h->W != W))
        return fail(h, LDA_B200_ESTATE, "PartialFit on a fitted model needs the same number of rows (have %lld, got %lld)",
                    (long long)h->W, (long long)W);
    if (((!nphi0 && !warm) || !ntheta0) && !rnd) return fail(h, LDA_B200_EINVAL, "neither initial values nor a PCG state were given");
    CU(cudaSetDevice(h->device));
    PhaseTimer pt(h->stream);
    h->fit_open = false;
    // what can be checked without touching the model is checked first: a call that fails here leaves the previous
    // model as it was
    if (D < 0) return fail(h, LDA_B200_EINVAL, "negative document count");
    if (!indptr) return fail(h, LDA_B200_EINVAL, "indptr is NULL");
    for (const Segment &sg : make_segments(D, h->cfg.batch_size, h->rank, h->nranks))
        for (int64_t g = sg.g0; g < sg.g1; ++g)
            if (indptr[g + 1] < indptr[g] || indptr[g] < 0)
                return fail(h, LDA_B200_EINVAL, "indptr is not non-decreasing at column %lld", (long long)g);
    if (!warm) h->fitted = false;  // nPhi / nZ are re-drawn below: a fit that fails from here on leaves no model behind
    // init (lda.go:193-206): nPhi/nZ are re-initialised on every fit, no warm start.  The device buffers (and the
    // peer mappings) are kept when the shape is unchanged; every value in them is rewritten below.
    const bool same_shape = h->nphi && h->W == W && h->K == h->cfg.K && (h->nranks == 1 || h->peer_ok || !h->peer_wanted) &&
                            (h->nranks == 1 || h->peer_setup_det == h->deterministic);
    if (warm && !same_shape) {
        // keep the model values while moving them into peer-shared buffers: not needed on one GPU, unsupported here
        return fail(h, LDA_B200_EUNSUPPORTED, "PartialFit across a communicator change is not supported");
    }
    if (!same_shape) {
        h->W = 0;
        TRY(ensure_model(h, W));
        TRY(peer_setup(h));
    }
    if (h->peer_ok) {
        CU(cudaMemsetAsync(h->hat_buf[0], 0, (size_t)W * h->Kp * sizeof(float), h->stream));
        CU(cudaMemsetAsync(h->hat_buf[1], 0, (size_t)W * h->Kp * sizeof(float), h->stream));
    }
    CU(cudaMemsetAsync(h->nzhat, 0, (size_t)h->Kp * sizeof(double), h->stream));
    if (h->deterministic) {
        TRY(dalloc(h, &h->hat_fx, (size_t)W * h->Kp));
        CU(cudaMemsetAsync(h->hat_fx, 0, (size_t)W * h->Kp * sizeof(unsigned long long), h->stream));
    } else {
        dfree(h, &h->hat_fx);
    }
    uint64_t draws = 0;
    if (warm) {
        // streaming step: nPhi / nZ / rhoPhiT carry over
    } else if (nphi0) {
        std::vector<Segment> all{{0, W, 0}};
        TRY(upload_rows(h, all, nphi0, h->K, h->Kp, h->nphi));
    } else {
        const int64_t work = W * ((h->K + PCG_CH - 1) / PCG_CH);
        pcg_fill_rows_kernel<<<grid_for(work, 128, h->sms * 16), 128, 0, h->stream>>>(rnd->high, rnd->low, 0, (uint64_t)W * (uint64_t)h->K,
                                                                                  W, h->K, h->Kp, W, 1, 0, h->nphi);
        KCHECK();
        draws += (uint64_t)W * (uint64_t)h->K;
    }
    if (!warm) TRY(refresh_nz_from_nphi(h));
    TRY(refresh_inv_z(h));
    pt.lap("fit_begin: model alloc + nPhi init");

    // (explicit ntheta0 rows are uploaded through the same stage buffer the deferred ids would occupy)
    TRY(upload_corpus(h, h->fitc, W, D, indptr, ind, data, nullptr, true, defer_ind && ntheta0 == nullptr, src_device));
    pt.t0 = PhaseTimer::now();
    TRY(init_theta(h, h->fitc, ntheta0, rnd, draws));
    pt.lap("fit_begin: nTheta init");
    if (!ntheta0) draws += (uint64_t)D * (uint64_t)h->K;
    if (rnd && draws) advance(rnd, draws);

    ctr->words_in_corpus += h->fitc.n_global;  // lda.go:481 (never reset)
    if (!warm) ctr->rho_phi_t = 1;             // lda.go:497
    CU(cudaMemsetAsync(h->nphi_hat, 0, (size_t)h->W * h->Kp * sizeof(float), h->stream));
    CU(cudaMemsetAsync(h->doc_counter, 0, sizeof(int), h->stream));
    h->it_done = 0;
    h->prev_perplexity = 0;
    h->stopped = false;
    h->fit_open = true;
    h->fitted = true;
    return 0;
}

int lda_b200_fit_begin(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                       const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd,
                       lda_b200_counters *ctr) {
    return fit_begin_impl(h, W, D, indptr, ind, data, nphi0, ntheta0, rnd, ctr, false);
}

// PartialFit (OnlineTransformer, vectorisers.go:36-39; a TODO for LDA at lda.go:147): `iterations` SCVB0 iterations
// over the given documents starting from the current nPhi / nZ / rhoPhiT (from a fresh random model on first use).
static int partial_fit_impl(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                            const double *data, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr,
                            int32_t iterations, double *theta_out, bool src_device) {
    const bool warm = h->fitted;
    TRY(fit_begin_impl(h, W, D, indptr, ind, data, nullptr, ntheta0, rnd, ctr, warm, iterations > 0, src_device));
    const int32_t saved_it = h->cfg.iterations, saved_f = h->cfg.perplexity_evaluation_frequency;
    h->cfg.iterations = iterations;
    h->cfg.perplexity_evaluation_frequency = 0;
    int rc = lda_b200_fit_iterate(h, iterations, ctr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    h->cfg.iterations = saved_it;
    h->cfg.perplexity_evaluation_frequency = saved_f;
    if (rc) {
        free_corpus(h, h->fitc);
        h->fit_open = false;
        return rc;
    }
    return lda_b200_fit_end(h, theta_out);
}

int lda_b200_partial_fit(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                         const double *data, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr,
                         int32_t iterations, double *theta_out) {
    if (!h) return LDA_B200_EINVAL;
    return partial_fit_impl(h, W, D, indptr, ind, data, ntheta0, rnd, ctr, iterations, theta_out, false);
}

int lda_b200_partial_fit_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *ntheta0, lda_b200_pcg *rnd,
                                lda_b200_counters *ctr, int32_t iterations, double *theta_out) {
    if (!h) return LDA_B200_EINVAL;
    MatrixView v;
    TRY(open_matrix(h, m, v));
    const int rc = partial_fit_impl(h, v.W, v.D, v.indptr, v.ind, v.data, ntheta0, rnd, ctr, iterations, theta_out, v.on_device);
    free_device_csc(h, v.owned);
    return rc;
}

int lda_b200_fit_iterate(lda_b200_handle *h, int32_t n_iterations, lda_b200_counters *ctr, double *perp_trace,
                         int32_t perp_cap, int32_t *n_evals, int32_t *iters_run, int32_t *stopped, float *elapsed_ms) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fit_open) return fail(h, LDA_B200_ESTATE, "lda_b200_fit_iterate without lda_b200_fit_begin");
    if (!ctr) return fail(h, LDA_B200_EINVAL, "counters are NULL");
    CU(cudaSetDevice(h->device));
    const Corpus &c = h->fitc;
    const lda_b200_config &cfg = h->cfg;
    const int R = h->nranks;
    const int64_t b = c.b;
    const int64_t M = (c.D + b - 1) / b;      // numMiniBatches, lda.go:487
    const int G = R * c.conc;                 // minibatches in flight per step (the reference's `processes`, lda.go:488-491)
    const int64_t nsteps = (M + G - 1) / G;
    const int B = cfg.burn_in_passes;
    int evals = 0, ran = 0;
    std::vector<double> coef;  // c_g of the current step (see below)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, h->stream));
    }
    h->ev_used = 0;

    for (int it = 0; it < n_iterations && !h->stopped && h->it_done < cfg.iterations; ++it) {
        Nvtx it_range("lda_b200: SCVB0 iteration");
        ctr->rho_theta_t += 1;  // lda.go:502
        TRY(fill_rho_table(h, ctr->rho_theta_t, B + 1));
        // last iteration of lda_b200_fit: stream the result out behind every launch
        const bool stream_out = h->pending_theta_out && h->it_done + 1 == cfg.iterations && c.Dl > 0 &&
                                !is_pageable(h->pending_theta_out);
        if (stream_out) {
            TRY(dalloc(h, &h->dout, (size_t)h->K * c.Dl));
            h->theta_streamed = true;
        }
        for (int64_t s = 0; s < nsteps; ++s) {
            // minibatches s*G .. s*G+n_s-1 are all computed from the same nPhi snapshot and then blended one after
            // the other in index order (lda.go:299-317 n_s times):  nPhi' = A nPhi + sum_g c_g nPhiHat_g
            const int n_s = (int)std::min<int64_t>(G, M - s * G);
            DocsArgs a = docs_args(h, c, B);
            int64_t l_first = -1, l_last = -1;
            // coefficients of the step in one backward pass (what lda_b200_plan_step returns for every g, same order of
            // operations): a call per minibatch would cost n_s^2 pow() evaluations -- more than a millisecond of host time
            // per step at n_s = 128, as long as an 8-GPU step lasts on the device
            double A = 1.0;
            coef.resize((size_t)n_s);
            for (int g = n_s - 1; g >= 0; --g) {
                const double rho = rho_calc(cfg.rho_phi, ctr->rho_phi_t + g);
                coef[g] = rho * A;
                A *= (1.0 - rho);
            }
            const double c_single = n_s > 0 ? coef[0] : 0.0;  // used only when n_s == 1; identical on every rank
            for (int g = 0; g < n_s; ++g) {
                const int64_t m = s * G + g;
                const double cg = coef[g];
                if (m % R != h->rank) continue;
                const int64_t l = m / R;  // local minibatch index on this rank
                if (l_first < 0) l_first = l;
                l_last = l;
                const int64_t b_actual = (m < M - 1) ? b : c.D - m * b;  // lda.go:513-518, :268
                // a lone minibatch keeps rho out of the accumulated values (blend applies it); otherwise fold c_g in
                a.hat_scale[l - l_first] = (float)(ctr->words_in_corpus / (double)b_actual * (n_s > 1 ? cg : 1.0));
            }
            if (l_first >= 0) {
                Nvtx mb_range("lda_b200: minibatch launch (burn-in + sufficient statistics)");
                const int64_t r0 = l_first * b, r1 = std::min<int64_t>(c.Dl, (l_last + 1) * b);
                a.mb_origin = (int)r0;
                int slot;
                TRY(ev_begin(h, 0, &slot));
                double inv_fx = 1.0;
                if (h->hat_fx) {
                    // a word receives at most one contribution of at most hat_scale per document of the launch, so the
                    // sums stay below hat_scale_max * documents; 2^e is the largest power of two that keeps them < 2^62
                    float smax = 0.f;
                    for (float x : a.hat_scale) smax = std::max(smax, fabsf(x));
                    const double bound = std::max(1.0, (double)smax * (double)(r1 - r0));
                    const int e = std::max(-100, std::min(100, 61 - (int)ceil(log2(bound))));
                    a.hat_fx = h->hat_fx;
                    a.fx_scale = (float)ldexp(1.0, e);
                    inv_fx = ldexp(1.0, -e);
                }
                // one launch over the whole range, or one per sub-range while ids stream in / results stream out
                const bool split = (stream_out || c.ind_deferred) && c.sub < r1 - r0;
                const int64_t piece = split ? c.sub : r1 - r0;
                for (int64_t s0 = r0; s0 < r1; s0 += piece) {
                    const int64_t s1 = std::min(r1, s0 + piece);
                    if (c.ind_deferred) {
                        const size_t i0 = (size_t)(std::lower_bound(c.sub_begin.begin(), c.sub_begin.end(), s0) - c.sub_begin.begin());
                        for (size_t i = i0; i < c.sub_begin.size() && c.sub_begin[i] < s1; ++i) {
                            if (c.job)  // the event must have been recorded by the feeding thread before it can be waited on
                                while (c.job->recorded.load(std::memory_order_acquire) <= (int)i) std::this_thread::yield();
                            CU(cudaStreamWaitEvent(h->stream, c.range_ev[i], 0));
                        }
                    }
                    if (s0 > r0) CU(cudaMemsetAsync(h->doc_counter, 0, sizeof(int), h->stream));
                    a.doc_begin = (int)s0;
                    a.doc_end = (int)s1;
                    TRY(launch_docs<true>(h, a, c.order));
                    if (stream_out) TRY(stream_theta_range(h, c, s0, s1));
                }
                a.doc_begin = (int)r0;
                a.doc_end = (int)r1;
                if (h->hat_fx) {
                    const int64_t n = h->W * h->Kp;
                    hat_fx_to_float_kernel<<<grid_for(n, 256, h->sms * 8), 256, 0, h->stream>>>(h->hat_fx, h->nphi_hat, n, inv_fx);
                    KCHECK();
                }
                TRY(ev_end(h, slot));
            }
            const double Ad = A, Cd = (n_s > 1) ? 1.0 : c_single;
            Nvtx m_range(h->peer_ok ? "lda_b200: exchange + M-step (peer memory)" : "lda_b200: exchange + M-step");
            if (h->peer_ok) {
                // fused path: [barrier] -> reduce over peer memory + M-step on this rank's row slice -> [barrier + nZ]
                BarrierArgs bar;
                bar.pt = h->peers;
                bar.finish = 0;
                bar.K = h->K;
                bar.Kp = h->Kp;
                bar.nzhat_local = h->nzhat;
                bar.nz = h->nz;
                bar.inv_z = h->inv_z;
                bar.Ad = Ad;
                bar.Cd = Cd;
                bar.etaW = cfg.eta * (double)h->W;
                bar.doc_counter = h->doc_counter;
                bar.err = h->err_flag;
                int slot;
                TRY(ev_begin(h, 2, &slot));
                bar.epoch = ++h->epoch;
                int ph;
                TRY(ev_begin(h, 3, &ph));
                xgpu_barrier_kernel<<<1, 256, 0, h->stream>>>(bar);
                KCHECK();
                TRY(ev_end(h, ph));
                FusedArgs fa;
                fa.pt = h->peers;
                fa.cur = h->hat_cur;
                const int64_t rows_per = (h->W + R - 1) / R;
                fa.row0 = std::min<int64_t>(h->W, rows_per * h->rank);
                fa.row1 = std::min<int64_t>(h->W, fa.row0 + rows_per);
                fa.W = h->W;
                fa.Kp = h->Kp;
                fa.A = (float)Ad;
                fa.C = (float)Cd;
                // blocks of the exchange kernel: the multicast variant runs fewer threads with several pipelined batches each
                const int xgrid = h->nvls.active ? h->sms : h->sms * 4;
                TRY(ensure_partials(h, (size_t)xgrid * h->Kp));
                fa.partials = h->partials;
                fa.mc_hat = h->mc_hat[h->hat_cur];
                fa.mc_nphi = h->mc_nphi;
                const int bs = blend_block(h->Kp);
                TRY(ev_begin(h, 4, &ph));
                if (h->nvls.active)
                    nvls_reduce_blend_kernel<2><<<xgrid, bs, (size_t)bs * sizeof(float4), h->stream>>>(fa);
                else if (R <= 2)
                    peer_reduce_blend_kernel<4, 2><<<h->sms * 4, bs, (size_t)bs * sizeof(float4), h->stream>>>(fa);
                else if (R <= 4)
                    peer_reduce_blend_kernel<2, 4><<<h->sms * 4, bs, (size_t)bs * sizeof(float4), h->stream>>>(fa);
                else
                    peer_reduce_blend_kernel<1, 8><<<h->sms * 4, bs, (size_t)bs * sizeof(float4), h->stream>>>(fa);
                KCHECK();
                TRY(ev_end(h, ph));
                TRY(ev_begin(h, 5, &ph));
                TRY(reduce_partials(h, xgrid, h->Kp, h->nzhat, false, h->stream));  // this rank's share of nZHat
                bar.finish = 1;
                bar.epoch = ++h->epoch;
                xgpu_barrier_kernel<<<1, 256, 0, h->stream>>>(bar);
                KCHECK();
                TRY(ev_end(h, ph));
                TRY(ev_begin(h, 6, &ph));
                BlendArgs ca;  // c = (nPhi' + eta) invZ' over the local replica
                ca.nphi = h->nphi;
                ca.nphi_hat = nullptr;
                ca.cmat = h->cmat;
                ca.inv_z = h->inv_z;
                ca.W = h->W;
                ca.Kp = h->Kp;
                ca.A = 1.f;
                ca.C = 0.f;
                ca.eta = (float)cfg.eta;
                phi_blend_kernel<<<h->sms * 4, bs, 0, h->stream>>>(ca);
                KCHECK();
                TRY(ev_end(h, ph));
                TRY(ev_end(h, slot));
                h->hat_cur ^= 1;
                h->nphi_hat = h->hat_buf[h->hat_cur];
            } else {
                if (R > 1) {
                    int slot;
                    TRY(ev_begin(h, 2, &slot));
                    NC(nccl().AllReduce(h->nphi_hat, h->nphi_hat, (size_t)h->W * h->Kp, ncclFloat, ncclSum, h->comm, h->stream));
                    TRY(ev_end(h, slot));
                }
                const int bs = blend_block(h->Kp);
                int slot;
                TRY(ev_begin(h, 1, &slot));
                TRY(ensure_partials(h, (size_t)h->sms * 4 * h->Kp));
                hat_colsum_kernel<<<h->sms * 4, bs, (size_t)bs * sizeof(float4), h->stream>>>(h->nphi_hat, h->W, h->Kp, h->partials);
                KCHECK();
                nz_update_kernel<<<1, 1024, 0, h->stream>>>(h->nz, h->partials, h->sms * 4, h->inv_z, h->doc_counter, h->K, h->Kp, Ad, Cd,
                                                           cfg.eta * (double)h->W);
                KCHECK();
                BlendArgs ba;
                ba.nphi = h->nphi;
                ba.nphi_hat = h->nphi_hat;
                ba.cmat = h->cmat;
                ba.inv_z = h->inv_z;
                ba.W = h->W;
                ba.Kp = h->Kp;
                ba.A = (float)Ad;
                ba.C = (float)Cd;
                ba.eta = (float)cfg.eta;
                phi_blend_kernel<<<h->sms * 4, bs, 0, h->stream>>>(ba);
                KCHECK();
                TRY(ev_end(h, slot));
            }
            ctr->rho_phi_t += n_s;  // lda.go:300
        }
        if (h->fitc.ind_deferred) {  // every range has been waited for; report ids that were out of range
            h->fitc.ind_deferred = false;
            if (h->fitc.job) {
                if (h->fitc.job->worker.joinable()) h->fitc.job->worker.join();
                const int jrc = h->fitc.job->rc;
                const std::string jerr = h->fitc.job->err;
                h->fitc.job.reset();
                if (jrc) {
                    h->fitted = false;
                    h->err = jerr;
                    return jrc;
                }
            }
            int bad = 0;
            CU(cudaMemcpyAsync(&bad, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (bad == 1) {
                h->fitted = false;  // the first iteration already blended statistics of the clamped ids into nPhi / nZ
                return fail(h, LDA_B200_EINVAL, "a row index lies outside [0, %lld)", (long long)h->W);
            }
        }
        h->it_done++;
        ran++;
        const int f = cfg.perplexity_evaluation_frequency;
        if (f > 0 && h->it_done % f == 0) {  // lda.go:530-539
            double p;
            TRY(eval_perplexity(h, c, ctr->words_in_corpus, &p));
            if (perp_trace && evals < perp_cap) perp_trace[evals] = p;
            evals++;
            if (h->prev_perplexity != 0 && fabs(h->prev_perplexity - p) < cfg.perplexity_tolerance) {
                h->stopped = true;
                break;
            }
            h->prev_perplexity = p;
        }
    }
    if (elapsed_ms) {
        CU(cudaEventRecord(e1, h->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(elapsed_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    } else {
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->peer_ok) {
        int bad = 0;
        CU(cudaMemcpy(&bad, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad == 2) return fail(h, LDA_B200_ENCCL, "a cross-GPU barrier timed out (a peer rank did not arrive)");
    }
    if (h->profiling) TRY(ev_collect(h));
    if (n_evals) *n_evals = evals;
    if (iters_run) *iters_run = ran;
    if (stopped) *stopped = h->stopped ? 1 : 0;
    return 0;
}

int lda_b200_fit_end(lda_b200_handle *h, double *theta_out) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fit_open) return fail(h, LDA_B200_ESTATE, "lda_b200_fit_end without lda_b200_fit_begin");
    CU(cudaSetDevice(h->device));
    PhaseTimer pt(h->stream);
    int rc = 0;
    if (h->theta_streamed && theta_out == h->pending_theta_out) {
        cudaError_t e = cudaStreamSynchronize(h->copy_stream);
        if (e != cudaSuccess) rc = fail(h, LDA_B200_ECUDA, "theta read-back failed: %s", cudaGetErrorString(e));
        pt.lap("theta_out: wait for streamed copy");
    } else {
        rc = write_theta_out(h, h->fitc, theta_out);
    }
    dfree(h, &h->dout);
    h->theta_streamed = false;
    pt.t0 = PhaseTimer::now();
    free_corpus(h, h->fitc);
    pt.lap("fit_end: free corpus");
    h->fit_open = false;
    return rc;
}

static int fit_impl(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                    const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr, double *theta_out,
                    double *perp_trace, int32_t perp_cap, int32_t *n_evals, int32_t *iters_run, bool src_device) {
    PhaseTimer pt(h->stream);
    // one call: the caller's arrays stay valid until it returns, so the word ids may stream in behind the first launches
    TRY(fit_begin_impl(h, W, D, indptr, ind, data, nphi0, ntheta0, rnd, ctr, false, h->cfg.iterations > 0, src_device));
    pt.lap("fit: begin total");
    h->pending_theta_out = theta_out;
    h->theta_streamed = false;
    int rc = lda_b200_fit_iterate(h, h->cfg.iterations, ctr, perp_trace, perp_cap, n_evals, iters_run, nullptr, nullptr);
    pt.lap("fit: iterations");
    if (rc) {
        cudaStreamSynchronize(h->copy_stream);
        h->pending_theta_out = nullptr;
        dfree(h, &h->dout);
        free_corpus(h, h->fitc);
        h->fit_open = false;
        return rc;
    }
    rc = lda_b200_fit_end(h, theta_out);
    h->pending_theta_out = nullptr;
    return rc;
}

int lda_b200_fit(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                 const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd,
                 lda_b200_counters *ctr, double *theta_out, double *perp_trace, int32_t perp_cap, int32_t *n_evals,
                 int32_t *iters_run) {
    if (!h) return LDA_B200_EINVAL;
    return fit_impl(h, W, D, indptr, ind, data, nphi0, ntheta0, rnd, ctr, theta_out, perp_trace, perp_cap, n_evals, iters_run, false);
}

int lda_b200_fit_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *nphi0, const double *ntheta0,
                        lda_b200_pcg *rnd, lda_b200_counters *ctr, double *theta_out, double *perp_trace, int32_t perp_cap,
                        int32_t *n_evals, int32_t *iters_run) {
    if (!h) return LDA_B200_EINVAL;
    MatrixView v;
    TRY(open_matrix(h, m, v));
    const int rc = fit_impl(h, v.W, v.D, v.indptr, v.ind, v.data, nphi0, ntheta0, rnd, ctr, theta_out, perp_trace, perp_cap, n_evals,
                            iters_run, v.on_device);
    free_device_csc(h, v.owned);
    return rc;
}

// ---- Transform / Perplexity / Components -----------------------------------------------------------------
static int transform_impl(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                          const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *theta_out, int32_t *passes_out,
                          bool src_device) {
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Transform called before Fit / set_state");
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    CU(cudaSetDevice(h->device));
    Corpus c;
    int rc = upload_corpus(h, c, h->W, D, indptr, ind, data, nullptr, false, false, src_device, kTransformRangeDocs);
    if (!rc) rc = init_theta(h, c, theta0, rnd, 0);
    if (!rc && !theta0) advance(rnd, (uint64_t)D * (uint64_t)h->K);
    if (!rc) rc = run_transform(h, c, rho_theta_t, theta_out, passes_out);
    free_corpus(h, c);
    return rc;
}

static int perplexity_impl(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                           const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *out, bool src_device) {
    if (!out) return fail(h, LDA_B200_EINVAL, "out is NULL");
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Perplexity called before Fit / set_state");
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    CU(cudaSetDevice(h->device));
    Corpus c;
    int rc = upload_corpus(h, c, h->W, D, indptr, ind, data, nullptr, false, false, src_device, kTransformRangeDocs);
    if (!rc) rc = init_theta(h, c, theta0, rnd, 0);
    if (!rc && !theta0) advance(rnd, (uint64_t)D * (uint64_t)h->K);
    if (!rc) rc = run_transform(h, c, rho_theta_t, nullptr, nullptr);   // also yields wordCount = c.n_global, lda.go:372-385
    if (!rc) rc = eval_perplexity(h, c, c.n_global, out);
    free_corpus(h, c);
    return rc;
}

// a matrix for Transform / Perplexity must have the model's number of rows (the reference would index out of range)
static int open_model_matrix(lda_b200_handle *h, const lda_b200_matrix *m, MatrixView &v) {
    if (m && h->fitted && m->rows != h->W)
        return fail(h, LDA_B200_EINVAL, "matrix has %lld rows, the model was fitted on %lld", (long long)m->rows, (long long)h->W);
    return open_matrix(h, m, v);
}

int lda_b200_transform(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                       const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *theta_out,
                       int32_t *passes_out) {
    if (!h) return LDA_B200_EINVAL;
    return transform_impl(h, D, indptr, ind, data, theta0, rnd, rho_theta_t, theta_out, passes_out, false);
}

int lda_b200_transform_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                              double rho_theta_t, double *theta_out, int32_t *passes_out) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Transform called before Fit / set_state");
    MatrixView v;
    TRY(open_model_matrix(h, m, v));
    const int rc = transform_impl(h, v.D, v.indptr, v.ind, v.data, theta0, rnd, rho_theta_t, theta_out, passes_out, v.on_device);
    free_device_csc(h, v.owned);
    return rc;
}

// ColDo(lda.Transform(m), func(j, v) { index.Index(v, ids[j]) }) with the vectors staying in HBM: the Transform pipeline's
// normalise + transpose kernel writes each range of documents straight into the index's [dim][cap] matrix.
int lda_b200_transform_into_index(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                                  double rho_theta_t, lda_b200_index *ix, const int64_t *ids) {
    if (!h) return LDA_B200_EINVAL;
    if (!ix) return fail(h, LDA_B200_EINVAL, "index is NULL");
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Transform called before Fit / set_state");
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    MatrixView v;
    TRY(open_model_matrix(h, m, v));
    Corpus c;
    auto run = [&]() -> int {
        TRY(upload_corpus(h, c, h->W, v.D, v.indptr, v.ind, v.data, nullptr, false, false, v.on_device, kTransformRangeDocs));
        TRY(init_theta(h, c, theta0, rnd, 0));
        if (!theta0) advance(rnd, (uint64_t)v.D * (uint64_t)h->K);
        double *dst = nullptr;
        int64_t ld = 0;
        if (index_begin_append(ix, h->device, h->K, c.Dl, &dst, &ld))
            return fail(h, LDA_B200_EINVAL, "index: %s", lda_b200_index_last_error(ix));
        TRY(run_transform(h, c, rho_theta_t, nullptr, nullptr, dst, ld));  // returns with every stream drained
        std::vector<int64_t> new_ids((size_t)c.Dl);  // this rank's documents, in local order
        for (const Segment &s : c.segs)
            for (int64_t g = s.g0; g < s.g1; ++g) new_ids[(size_t)(s.l0 + g - s.g0)] = ids ? ids[g] : g;
        if (index_finish_append(ix, c.Dl, new_ids.data())) return fail(h, LDA_B200_EINVAL, "index: %s", lda_b200_index_last_error(ix));
        return 0;
    };
    const int rc = run();
    free_corpus(h, c);
    free_device_csc(h, v.owned);
    return rc;
}

int lda_b200_perplexity(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                        const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *out) {
    if (!h) return LDA_B200_EINVAL;
    return perplexity_impl(h, D, indptr, ind, data, theta0, rnd, rho_theta_t, out, false);
}

int lda_b200_perplexity_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                               double rho_theta_t, double *out) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Perplexity called before Fit / set_state");
    MatrixView v;
    TRY(open_model_matrix(h, m, v));
    const int rc = perplexity_impl(h, v.D, v.indptr, v.ind, v.data, theta0, rnd, rho_theta_t, out, v.on_device);
    free_device_csc(h, v.owned);
    return rc;
}

int lda_b200_components(lda_b200_handle *h, double *out) {
    if (!h) return LDA_B200_EINVAL;
    if (!out) return fail(h, LDA_B200_EINVAL, "out is NULL");
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "Components called before Fit / set_state");
    CU(cudaSetDevice(h->device));
    TRY(column_sums(h, h->nphi, h->colsum));
    double *dout = nullptr;
    TRY(dalloc(h, &dout, (size_t)h->K * h->W));
    transpose_out_kernel<false><<<(unsigned)((h->W + 31) / 32), 256, 0, h->stream>>>(h->nphi, h->W, h->K, h->Kp, h->colsum, dout, h->W);
    h->launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, (size_t)h->K * h->W * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFreeAsync(dout, h->stream);
    if (e != cudaSuccess) return fail(h, LDA_B200_ECUDA, "components failed: %s", cudaGetErrorString(e));
    return 0;
}

int lda_b200_get_dims(const lda_b200_handle *h, int64_t *W, int32_t *K) {
    if (!h) return LDA_B200_EINVAL;
    if (W) *W = h->W;
    if (K) *K = h->fitted ? h->K : h->cfg.K;
    return 0;
}

int lda_b200_get_state(lda_b200_handle *h, double *nphi, double *nz) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "no model state");
    CU(cudaSetDevice(h->device));
    if (nphi) {
        const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)(h->stage_bytes / sizeof(double)) / h->K);
        for (int64_t r = 0; r < h->W; r += rows_per_chunk) {
            const int64_t n = std::min(rows_per_chunk, h->W - r);
            unpad_rows_kernel<<<grid_for(n * h->K, 256, h->sms * 8), 256, 0, h->stream>>>(h->nphi + r * h->Kp, (double *)h->stage, n, h->K, h->Kp);
            KCHECK();
            CU(cudaMemcpyAsync(nphi + r * h->K, h->stage, (size_t)n * h->K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
    }
    if (nz) {
        CU(cudaMemcpyAsync(nz, h->nz, (size_t)h->K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int lda_b200_set_state(lda_b200_handle *h, int64_t W, const double *nphi, const double *nz) {
    if (!h) return LDA_B200_EINVAL;
    if (!nphi || !nz) return fail(h, LDA_B200_EINVAL, "nphi/nz is NULL");
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    CU(cudaSetDevice(h->device));
    h->W = 0;
    TRY(ensure_model(h, W));
    std::vector<Segment> all{{0, W, 0}};
    TRY(upload_rows(h, all, nphi, h->K, h->Kp, h->nphi));
    CU(cudaMemsetAsync(h->nz, 0, (size_t)h->Kp * sizeof(double), h->stream));
    CU(cudaMemcpyAsync(h->nz, nz, (size_t)h->K * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    TRY(refresh_inv_z(h));
    CU(cudaStreamSynchronize(h->stream));
    h->fitted = true;
    return 0;
}

// ---- the *_matrix calls with the descriptor spread over scalar arguments (the form cgo can call) ----------------------
static lda_b200_matrix flat_matrix(int32_t format, int64_t rows, int64_t cols, const int64_t *ptr, const int64_t *ind, const double *data) {
    lda_b200_matrix m;
    memset(&m, 0, sizeof m);
    m.format = format;
    m.rows = rows;
    m.cols = cols;
    m.ptr = ptr;
    m.ind = ind;
    m.data = data;
    return m;
}

int lda_b200_fit_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr, const int64_t *ind,
                      const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr,
                      double *theta_out, double *perp_trace, int32_t perp_cap, int32_t *n_evals, int32_t *iters_run) {
    const lda_b200_matrix m = flat_matrix(format, rows, cols, ptr, ind, data);
    return lda_b200_fit_matrix(h, &m, nphi0, ntheta0, rnd, ctr, theta_out, perp_trace, perp_cap, n_evals, iters_run);
}

int lda_b200_partial_fit_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                              const int64_t *ind, const double *data, const double *ntheta0, lda_b200_pcg *rnd,
                              lda_b200_counters *ctr, int32_t iterations, double *theta_out) {
    const lda_b200_matrix m = flat_matrix(format, rows, cols, ptr, ind, data);
    return lda_b200_partial_fit_matrix(h, &m, ntheta0, rnd, ctr, iterations, theta_out);
}

int lda_b200_transform_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                            const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                            double rho_theta_t, double *theta_out, int32_t *passes_out) {
    const lda_b200_matrix m = flat_matrix(format, rows, cols, ptr, ind, data);
    return lda_b200_transform_matrix(h, &m, theta0, rnd, rho_theta_t, theta_out, passes_out);
}

int lda_b200_perplexity_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                             const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                             double rho_theta_t, double *out) {
    const lda_b200_matrix m = flat_matrix(format, rows, cols, ptr, ind, data);
    return lda_b200_perplexity_matrix(h, &m, theta0, rnd, rho_theta_t, out);
}

int lda_b200_transform_into_index_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                                       const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                                       double rho_theta_t, lda_b200_index *ix, const int64_t *ids) {
    const lda_b200_matrix m = flat_matrix(format, rows, cols, ptr, ind, data);
    return lda_b200_transform_into_index(h, &m, theta0, rnd, rho_theta_t, ix, ids);
}

// ---- Save / Load (layout in include/lda_b200.h) ---------------------------------------------------------------------
namespace {
constexpr uint32_t kGonumVersion = 0xFE000001u;
constexpr int64_t kDenseHeader = 40;  // uint32 version, 'G','F','A', bool unit, int64 rows, cols, ku, kl
void put_dense_header(unsigned char *p, int64_t rows, int64_t cols) {
    memset(p, 0, (size_t)kDenseHeader);
    memcpy(p, &kGonumVersion, 4);  // the host is little endian (x86-64 / aarch64)
    p[4] = 'G';
    p[5] = 'F';
    p[6] = 'A';
    p[7] = 0;
    memcpy(p + 8, &rows, 8);
    memcpy(p + 16, &cols, 8);
}
bool get_dense_header(const unsigned char *p, int64_t *rows, int64_t *cols) {
    uint32_t v;
    memcpy(&v, p, 4);
    if (v != kGonumVersion || p[4] != 'G' || p[5] != 'F' || p[6] != 'A') return false;
    memcpy(rows, p + 8, 8);
    memcpy(cols, p + 16, 8);
    return *rows >= 0 && *cols >= 0;
}
}  // namespace

int lda_b200_save_size(const lda_b200_handle *h, int64_t *bytes) {
    if (!h || !bytes) return LDA_B200_EINVAL;
    if (!h->fitted) return LDA_B200_ESTATE;
    *bytes = 32 + kDenseHeader + h->W * h->K * 8 + kDenseHeader + (int64_t)h->K * 8;
    return 0;
}

int lda_b200_save(lda_b200_handle *h, const lda_b200_counters *ctr, void *buf, int64_t cap, int64_t *written) {
    if (!h) return LDA_B200_EINVAL;
    if (!ctr || !buf) return fail(h, LDA_B200_EINVAL, "counters / buffer is NULL");
    if (!h->fitted) return fail(h, LDA_B200_ESTATE, "no model state");
    int64_t need = 0;
    TRY(lda_b200_save_size(h, &need));
    if (cap < need) return fail(h, LDA_B200_EINVAL, "buffer of %lld bytes is too small for a model of %lld bytes", (long long)cap, (long long)need);
    unsigned char *p = (unsigned char *)buf;
    const uint64_t k = (uint64_t)h->K;
    memcpy(p, &k, 8);
    memcpy(p + 8, &ctr->rho_phi_t, 8);
    memcpy(p + 16, &ctr->rho_theta_t, 8);
    memcpy(p + 24, &ctr->words_in_corpus, 8);
    p += 32;
    put_dense_header(p, h->W, h->K);
    unsigned char *nphi_at = p + kDenseHeader;
    p = nphi_at + h->W * h->K * 8;
    put_dense_header(p, 1, h->K);
    unsigned char *nz_at = p + kDenseHeader;
    // the float64 payloads may sit at any byte offset of the caller's buffer: go through aligned temporaries when needed
    if (((uintptr_t)nphi_at | (uintptr_t)nz_at) % 8 == 0) {
        TRY(lda_b200_get_state(h, (double *)nphi_at, (double *)nz_at));
    } else {
        std::vector<double> a((size_t)(h->W * h->K)), b((size_t)h->K);
        TRY(lda_b200_get_state(h, a.data(), b.data()));
        memcpy(nphi_at, a.data(), a.size() * 8);
        memcpy(nz_at, b.data(), b.size() * 8);
    }
    if (written) *written = need;
    return 0;
}

int lda_b200_load(lda_b200_handle *h, const void *buf, int64_t len, lda_b200_counters *ctr) {
    if (!h) return LDA_B200_EINVAL;
    if (!buf || !ctr) return fail(h, LDA_B200_EINVAL, "buffer / counters is NULL");
    if (h->fit_open) return fail(h, LDA_B200_ESTATE, "a fit is in progress");
    const unsigned char *p = (const unsigned char *)buf;
    if (len < 32 + kDenseHeader) return fail(h, LDA_B200_EINVAL, "unexpected EOF: %lld bytes hold no model", (long long)len);  // io.ErrUnexpectedEOF, dimreduction.go:135
    uint64_t k;
    lda_b200_counters c;
    memcpy(&k, p, 8);
    memcpy(&c.rho_phi_t, p + 8, 8);
    memcpy(&c.rho_theta_t, p + 16, 8);
    memcpy(&c.words_in_corpus, p + 24, 8);
    int64_t rows = 0, cols = 0;
    if (!get_dense_header(p + 32, &rows, &cols)) return fail(h, LDA_B200_EINVAL, "not a gonum mat.Dense record (nPhi)");
    if (k < 1 || k > 1024 || cols != (int64_t)k || rows < 1) return fail(h, LDA_B200_EINVAL, "model dimensions do not match K");
    const int64_t nphi_off = 32 + kDenseHeader, nz_hdr = nphi_off + rows * cols * 8;
    if (rows > (len - nphi_off) / (cols * 8) || len < nz_hdr + kDenseHeader + cols * 8)
        return fail(h, LDA_B200_EINVAL, "unexpected EOF: truncated mat.Dense record");
    int64_t zr = 0, zc = 0;
    if (!get_dense_header(p + nz_hdr, &zr, &zc) || zr != 1 || zc != cols) return fail(h, LDA_B200_EINVAL, "not a gonum mat.Dense record (nZ)");
    lda_b200_config cfg = h->cfg;
    cfg.K = (int32_t)k;
    TRY(lda_b200_set_config(h, &cfg));
    const unsigned char *nphi_at = p + nphi_off, *nz_at = p + nz_hdr + kDenseHeader;
    if (((uintptr_t)nphi_at | (uintptr_t)nz_at) % 8 == 0) {
        TRY(lda_b200_set_state(h, rows, (const double *)nphi_at, (const double *)nz_at));
    } else {
        std::vector<double> a((size_t)(rows * cols)), b((size_t)cols);
        memcpy(a.data(), nphi_at, a.size() * 8);
        memcpy(b.data(), nz_at, b.size() * 8);
        TRY(lda_b200_set_state(h, rows, a.data(), b.data()));
    }
    *ctr = c;
    return 0;
}

// ---- host-only planning helpers -----------------------------------------------------------------------------
int lda_b200_plan_step(const lda_b200_schedule *rho_phi, double rho_phi_t, int32_t n_s, int32_t rank, double *A,
                       double *c_rank) {
    if (!rho_phi || n_s < 0) return LDA_B200_EINVAL;
    double tail = 1.0, cg = 0.0;
    for (int g = n_s - 1; g >= 0; --g) {
        const double rho = rho_calc(*rho_phi, rho_phi_t + g);
        if (g == rank) cg = rho * tail;
        tail *= (1.0 - rho);
    }
    if (A) *A = tail;
    if (c_rank) *c_rank = cg;
    return 0;
}

int64_t lda_b200_plan_shard(int64_t D, int64_t batch_size, int32_t rank, int32_t nranks, int64_t *seg_begin,
                            int64_t *seg_end, int64_t cap) {
    if (D < 0 || batch_size < 1 || nranks < 1 || rank < 0 || rank >= nranks) return -1;
    const std::vector<Segment> s = make_segments(D, batch_size, rank, nranks);
    for (size_t i = 0; i < s.size() && (int64_t)i < cap; ++i) {
        if (seg_begin) seg_begin[i] = s[i].g0;
        if (seg_end) seg_end[i] = s[i].g1;
    }
    return (int64_t)s.size();
}

// ---- PCG host helpers ------------------------------------------------------------------------------------
void lda_b200_pcg_seed(lda_b200_pcg *rnd, uint64_t seed) {
    rnd->high = seed;
    rnd->low = seed;
}
uint64_t lda_b200_pcg_uint64(lda_b200_pcg *rnd) {
    Pcg p;
    p.s = mk128(rnd->high, rnd->low);
    const uint64_t r = pcg_next(p);
    rnd->high = (uint64_t)(p.s >> 64);
    rnd->low = (uint64_t)p.s;
    return r;
}
void lda_b200_pcg_fill(lda_b200_pcg *rnd, int64_t count, int64_t n, double *out) {
    Pcg p;
    p.s = mk128(rnd->high, rnd->low);
    for (int64_t i = 0; i < count; ++i) out[i] = pcg_draw(p, (uint64_t)n);
    rnd->high = (uint64_t)(p.s >> 64);
    rnd->low = (uint64_t)p.s;
}
void lda_b200_pcg_advance(lda_b200_pcg *rnd, uint64_t draws) { advance(rnd, draws); }

// ---- diagnostics -------------------------------------------------------------------------------------------
static int ingest_roundtrip_impl(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                                 const double *data, int64_t *indptr_out, int32_t *ind_out, float *val_out, double *wc_out,
                                 double *n_out, bool src_device) {
    CU(cudaSetDevice(h->device));
    Corpus c;
    double *wc64 = nullptr;
    int rc = dalloc(h, &wc64, (size_t)std::max<int64_t>(D, 1));
    if (!rc) rc = upload_corpus(h, c, W, D, indptr, ind, data, wc64, true, false, src_device);
    auto back = [&]() -> int {
        if (indptr_out) CU(cudaMemcpyAsync(indptr_out, c.doc_ptr, (size_t)(c.Dl + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        if (ind_out) CU(cudaMemcpyAsync(ind_out, c.ind, (size_t)c.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        if (val_out) CU(cudaMemcpyAsync(val_out, c.val, (size_t)c.nnz * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (wc_out) CU(cudaMemcpyAsync(wc_out, wc64, (size_t)c.Dl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (n_out) *n_out = c.n_global;
        return 0;
    };
    if (!rc) rc = back();
    free_corpus(h, c);
    if (wc64) cudaFreeAsync(wc64, h->stream);
    return rc;
}

int lda_b200_ingest_roundtrip(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                              const double *data, int64_t *indptr_out, int32_t *ind_out, float *val_out,
                              double *wc_out, double *n_out) {
    if (!h) return LDA_B200_EINVAL;
    return ingest_roundtrip_impl(h, W, D, indptr, ind, data, indptr_out, ind_out, val_out, wc_out, n_out, false);
}

int lda_b200_ingest_roundtrip_matrix(lda_b200_handle *h, const lda_b200_matrix *m, int64_t *nnz_out, int64_t *indptr_out,
                                     int32_t *ind_out, float *val_out, double *wc_out, double *n_out) {
    if (!h) return LDA_B200_EINVAL;
    MatrixView v;
    TRY(open_matrix(h, m, v));
    int rc = 0;
    if (nnz_out) *nnz_out = v.indptr ? v.indptr[v.D] : 0;
    if (indptr_out || ind_out || val_out || wc_out || n_out)
        rc = ingest_roundtrip_impl(h, v.W, v.D, v.indptr, v.ind, v.data, indptr_out, ind_out, val_out, wc_out, n_out, v.on_device);
    free_device_csc(h, v.owned);
    return rc;
}

// sparse.ToCSC() of a CSR term x document matrix on the device; result copied back to the host
int lda_b200_csr_to_csc(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *row_ptr, const int64_t *col_ind,
                        const double *data, int64_t *indptr_out, int64_t *ind_out, double *data_out) {
    if (!h) return LDA_B200_EINVAL;
    if (!indptr_out) return fail(h, LDA_B200_EINVAL, "bad CSR arguments");
    CU(cudaSetDevice(h->device));
    DeviceCSC m;
    int rc = device_csr_to_csc(h, W, D, row_ptr, col_ind, data, m);
    auto back = [&]() -> int {
        if (m.nnz > 0 && (!ind_out || !data_out)) return fail(h, LDA_B200_EINVAL, "NULL CSR array");
        memcpy(indptr_out, m.h_indptr.data(), ((size_t)D + 1) * sizeof(int64_t));
        if (m.nnz > 0) {
            CU(cudaMemcpyAsync(ind_out, m.ind, (size_t)m.nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(data_out, m.data, (size_t)m.nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    };
    if (!rc) rc = back();
    free_device_csc(h, m);
    return rc;
}

int64_t lda_b200_launch_count(const lda_b200_handle *h) { return h ? h->launches.load() : (int64_t)0; }

int lda_b200_last_kernel_time(const lda_b200_handle *h, float *minibatch_ms, int32_t *minibatch_launches,
                              float *blend_ms, int32_t *blend_launches) {
    if (!h) return LDA_B200_EINVAL;
    if (minibatch_ms) *minibatch_ms = h->mb_ms;
    if (minibatch_launches) *minibatch_launches = h->mb_n;
    if (blend_ms) *blend_ms = h->blend_ms;
    if (blend_launches) *blend_launches = h->blend_n;
    return 0;
}

// Memory floor of the minibatch kernel (diagnostic): replays the word ids of the first launch range of the resident
// corpus through memory_mix_probe_kernel -- row gathers only, reductions only, and the kernel's own mix of BurnInPasses+1
// gathers and one reduction per entry -- and reports the best of three timings of each.
int lda_b200_probe_memory_mix(lda_b200_handle *h, int64_t *entries, float *gather_ms, float *red_ms, float *mix_ms) {
    if (!h) return LDA_B200_EINVAL;
    if (!h->fit_open) return fail(h, LDA_B200_ESTATE, "lda_b200_probe_memory_mix needs a resident corpus (lda_b200_fit_begin)");
    CU(cudaSetDevice(h->device));
    const Corpus &c = h->fitc;
    if (c.ind_deferred) CU(cudaStreamSynchronize(h->copy_stream));
    const int64_t l1 = std::min<int64_t>(c.Dl, c.b * c.conc);
    const int64_t n = c.lp.empty() ? 0 : c.lp[(size_t)l1];
    if (entries) *entries = n;
    if (n == 0) return fail(h, LDA_B200_EINVAL, "the first launch range holds no entries");
    const int B = h->cfg.burn_in_passes;
    float *sink = nullptr;
    TRY(dalloc(h, &sink, 1));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int grid = h->sms * 8;
    auto best_of = [&](int mode, float *out) -> int {
        float best = 0.f;
        for (int rep = 0; rep < 4; ++rep) {  // the first repetition warms the caches and is not counted
            CU(cudaEventRecord(e0, h->stream));
            if (mode == 0)
                memory_mix_probe_kernel<false, 4><<<grid, 256, 0, h->stream>>>(h->cmat, h->nphi_hat, c.ind, n, h->Kp, B + 1, sink);
            else if (mode == 1)
                memory_mix_probe_kernel<true, 4><<<grid, 256, 0, h->stream>>>(h->cmat, h->nphi_hat, c.ind, n, h->Kp, 0, sink);
            else
                memory_mix_probe_kernel<true, 4><<<grid, 256, 0, h->stream>>>(h->cmat, h->nphi_hat, c.ind, n, h->Kp, B + 1, sink);
            KCHECK();
            CU(cudaEventRecord(e1, h->stream));
            CU(cudaEventSynchronize(e1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && (best == 0.f || ms < best)) best = ms;
        }
        if (out) *out = best;
        return 0;
    };
    int rc = best_of(0, gather_ms);
    if (!rc) rc = best_of(1, red_ms);
    if (!rc) rc = best_of(2, mix_ms);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dfree(h, &sink);
    return rc;
}

int lda_b200_comm_info(const lda_b200_handle *h, int32_t *rank, int32_t *nranks, int32_t *peer_memory_path) {
    if (!h) return LDA_B200_EINVAL;
    if (rank) *rank = h->rank;
    if (nranks) *nranks = h->nranks;
    if (peer_memory_path) *peer_memory_path = h->peer_ok ? (h->nvls.active ? 2 : 1) : 0;
    return 0;
}

int lda_b200_last_exchange_phases(const lda_b200_handle *h, float *phase_ms4) {
    if (!h || !phase_ms4) return LDA_B200_EINVAL;
    for (int i = 0; i < 4; ++i) phase_ms4[i] = h->phase_ms[i];
    return 0;
}

int lda_b200_last_allreduce_time(const lda_b200_handle *h, float *allreduce_ms, int32_t *allreduce_calls) {
    if (!h) return LDA_B200_EINVAL;
    if (allreduce_ms) *allreduce_ms = h->ar_ms;
    if (allreduce_calls) *allreduce_calls = h->ar_n;
    return 0;
}

}  // extern "C"
