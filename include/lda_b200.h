/*
 * lda_b200.h -- C ABI of liblda_b200.so: B200 (sm_100a) implementation of the LDA SCVB0
 * path of james-bowman/nlp (lda.go).  This is the library a Go shim binds through cgo
 * (INTEGRATION.md shows the binding); there are no C++/torch types in any signature.
 *
 * The reference has no FFI boundary of its own (it is pure Go).  Each entry point below
 * therefore cites the Go method/field of `nlp.LatentDirichletAllocation` it replaces
 * (paths relative to the reference tree).
 *
 * Conventions
 *  - every function returns 0 on success or an lda_b200_status code; text via
 *    lda_b200_last_error().  Nothing aborts or exits the host process.
 *  - all pointers are host pointers owned by the caller; inputs are read-only; nothing is
 *    retained after the call returns (cgo pointer rules); outputs are caller-allocated.
 *  - the term x document matrix is passed as CSC: indptr[D+1], ind[nnz] (row = word id),
 *    data[nnz]; int64 = Go `int`, double = Go `float64` (what sparse.CSC / ToCSC() hold,
 *    lda.go:464-466).  A dense mat.Matrix is passed as the CSC listing each column's
 *    non-zero rows in ascending order (utils.go:51-57).  Token order inside a document is
 *    preserved exactly (it matters: updates inside a document are sequential).
 *  - doc-topic results are K x D row-major float64 (lda.go:452,541), components are
 *    K x W row-major float64 (lda.go:415).
 *  - one handle = one model on one GPU, used by one caller at a time.  With a
 *    communicator of R ranks (lda_b200_comm_init) each rank owns the global minibatches
 *    m with m % R == rank (lda.go:504-528 hands consecutive minibatches to workers);
 *    every rank passes global-shaped arguments (indptr[D+1], theta_out K x D) but the library
 *    reads only the indptr entries and ind/data ranges (and writes only the theta_out columns)
 *    of its own minibatches, so a rank may leave other ranks' documents empty in its view.
 */
#ifndef LDA_B200_H
#define LDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    LDA_B200_OK = 0,
    LDA_B200_EINVAL = 1,      /* bad argument (NULL, negative size, word id out of range, ...) */
    LDA_B200_ECUDA = 2,       /* CUDA runtime error (incl. no device) */
    LDA_B200_ENCCL = 3,       /* NCCL error or libnccl.so.2 not loadable */
    LDA_B200_EUNSUPPORTED = 4,/* e.g. K > 1024 */
    LDA_B200_ESTATE = 5       /* call needs a fitted model (Transform before Fit) */
} lda_b200_status;

/* nlp.LearningSchedule, lda.go:16-32 */
typedef struct {
    double S, Tau, Kappa;
} lda_b200_schedule;

/* exported fields of nlp.LatentDirichletAllocation, lda.go:68-134; defaults lda.go:160-186 */
typedef struct {
    int32_t K;                              /* lda.go:85  */
    int32_t iterations;                     /* lda.go:70  Iterations */
    int32_t batch_size;                     /* lda.go:82  BatchSize */
    int32_t burn_in_passes;                 /* lda.go:89  BurnInPasses */
    int32_t transformation_passes;          /* lda.go:93  TransformationPasses */
    int32_t perplexity_evaluation_frequency;/* lda.go:79  */
    int32_t change_evaluation_frequency;    /* lda.go:103 */
    int32_t processes;                      /* lda.go:134 Processes: minibatches in flight per step.  With R ranks
                                               each GPU processes max(1, processes / R) (at most 256) consecutive
                                               minibatches of its own in one launch, all from the same nPhi
                                               snapshot, blended afterwards in index order (lda.go:299-317) */
    double perplexity_tolerance;            /* lda.go:75  */
    double mean_change_tolerance;           /* lda.go:98  */
    double alpha, eta;                      /* lda.go:106,109 */
    lda_b200_schedule rho_phi, rho_theta;   /* lda.go:112,115 */
} lda_b200_config;

/* private scalar state of the Go struct that survives between calls (lda.go:117-121).
 * The Go shim keeps these in its struct and passes them in/out of each call. */
typedef struct {
    double rho_phi_t;       /* lda.go:117; reset to 1 by every fit (lda.go:497) */
    double rho_theta_t;     /* lda.go:118; never reset, ++ per iteration (lda.go:502) */
    double words_in_corpus; /* lda.go:120; accumulated with += and never reset (lda.go:481) */
} lda_b200_counters;

/* state of golang.org/x/exp/rand's PCGSource (128-bit; PCGSource.MarshalBinary order is
 * high then low).  Used for `Rnd` (lda.go:126) when the caller lets the library draw the
 * initial nPhi / nTheta on the device instead of passing them. */
typedef struct {
    uint64_t high, low;
} lda_b200_pcg;

/* The `mat.Matrix` argument of Fit / FitTransform / Transform / Perplexity (term x document: rows = words W,
 * cols = documents D) in the storage the Go value already has, so the shim need not call ToCSC() (lda.go:464-466)
 * or scan a dense matrix (utils.go:51-57) on one host thread: the conversion runs on the device and its result
 * never returns to the host.
 *  LDA_B200_CSC    ptr = indptr[cols+1], ind = row (word) ids, data = values    (sparse.CSC RawMatrix)
 *  LDA_B200_CSR    ptr = indptr[rows+1], ind = column (document) ids, data      (sparse.CSR RawMatrix; what the
 *                  TfidfTransformer hands to the next pipeline step, weightings.go:79-95).  Column ids inside a row
 *                  need not be sorted; inside a document the words come out in ascending order, as ToCSC() gives.
 *  LDA_B200_DENSE  data = rows*cols values, row-major (mat.Dense RawMatrix with Stride == cols); ptr = ind = NULL.
 *                  Exact zeros are skipped, rows ascending inside a document (utils.go:51-57).
 * With a communicator every rank passes the same whole matrix for CSR / DENSE (each converts it and keeps its own
 * minibatches); CSC keeps the per-rank-shard convention described above. */
typedef enum { LDA_B200_CSC = 0, LDA_B200_CSR = 1, LDA_B200_DENSE = 2 } lda_b200_format;
typedef struct {
    int32_t format; /* lda_b200_format */
    int64_t rows, cols;
    const int64_t *ptr;
    const int64_t *ind;
    const double *data;
} lda_b200_matrix;

typedef struct lda_b200_handle lda_b200_handle;

/* fills the defaults of NewLatentDirichletAllocation(k), lda.go:145-189 (processes = the host's hardware threads, as
 * lda.go:185 takes runtime.GOMAXPROCS(0)) */
void lda_b200_default_config(int32_t k, lda_b200_config *out);

/* NewLatentDirichletAllocation, lda.go:145: allocates the model on CUDA device `device`. */
int lda_b200_create(const lda_b200_config *cfg, int32_t device, lda_b200_handle **out);
void lda_b200_destroy(lda_b200_handle *h);
/* exported fields may be changed between calls in Go.  Changing K on a fitted model drops the model (nPhi / nZ are laid
 * out for the old K): Transform / Perplexity / Components / get_state then fail with LDA_B200_ESTATE until the next fit
 * or set_state. */
int lda_b200_set_config(lda_b200_handle *h, const lda_b200_config *cfg);
/* error text of the last failed call on h (h == NULL: last failed create on this thread) */
const char *lda_b200_last_error(const lda_b200_handle *h);

/* Multi-GPU (replaces the goroutine pool, lda.go:487-528): one handle per GPU, one
 * communicator over them.  id is a 128-byte ncclUniqueId produced by
 * lda_b200_comm_unique_id on rank 0 and distributed by the host. */
int lda_b200_comm_unique_id(void *id128);
int lda_b200_comm_init(lda_b200_handle *h, int32_t rank, int32_t nranks, const void *id128);

/* Host-only planning arithmetic of the multi-GPU step (no GPU needed; exported so it can be tested on CPU).
 * lda.go:299-317 applied for the n_s minibatches of one step in index order, all computed from the same nPhi:
 *   nPhi' = A nPhi + sum_g c_g nPhiHat_g,   rho_g = RhoPhi.Calc(rho_phi_t + g),
 *   A = prod_g (1 - rho_g),   c_g = rho_g prod_{h>g} (1 - rho_h).
 * Rank g scales its nPhiHat by c_g, one allreduce(sum) follows, every rank applies nPhi' = A nPhi + sum. */
int lda_b200_plan_step(const lda_b200_schedule *rho_phi, double rho_phi_t, int32_t n_s, int32_t rank, double *A,
                       double *c_rank);
/* global document ranges [begin,end) owned by `rank` (minibatch m -> rank m % nranks), in local row order;
 * returns the number of ranges (at most cap are written) */
int64_t lda_b200_plan_shard(int64_t D, int64_t batch_size, int32_t rank, int32_t nranks, int64_t *seg_begin,
                            int64_t *seg_end, int64_t cap);

/* FitTransform, lda.go:463-542 (Fit, lda.go:211, is the same call with theta_out == NULL).
 *  nphi0   W*K  values [w*K+k] replacing the draws of lda.go:199-205, or NULL
 *  ntheta0 D*K  values [j*K+k] replacing the draws of lda.go:472-475, or NULL
 *  rnd     if nphi0/ntheta0 is NULL the corresponding draws are taken from *rnd exactly as
 *          Rnd.Int() % n would produce them, and *rnd is advanced by W*K (+ D*K) draws
 *  ctr     in/out private counters (see lda_b200_counters)
 *  theta_out   K*D doubles or NULL
 *  perp_trace  receives every perplexity evaluated inside the fit (lda.go:530-539
 *              computes them but does not expose them); perp_cap entries; may be NULL
 *  n_evals, iters_run  optional outputs
 */
int lda_b200_fit(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                 const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd,
                 lda_b200_counters *ctr, double *theta_out, double *perp_trace, int32_t perp_cap, int32_t *n_evals,
                 int32_t *iters_run);

/* PartialFit: the OnlineTransformer method (vectorisers.go:36-39) the reference lists as a TODO for LDA
 * (lda.go:147).  Runs `iterations` SCVB0 iterations (lda.go:501-528) over the given documents starting from the
 * current nPhi / nZ / rhoPhiT; on a handle without a model it first draws nPhi like lda.go:193-206 (so one call on a
 * fresh handle equals lda_b200_fit with Iterations = iterations and no perplexity evaluation).  nTheta of the given
 * documents is drawn / taken from ntheta0 as in lda_b200_fit; wordsInCorpus accumulates (lda.go:481). */
int lda_b200_partial_fit(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                         const double *data, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr,
                         int32_t iterations, double *theta_out);

/* Transform, lda.go:446-453 (unNormalisedTransform lda.go:420-437 + normaliseTheta).
 *  theta0  D*K values [j*K+k] replacing the draws of lda.go:422-426, or NULL (then *rnd)
 *  rho_theta_t  the model's current rhoThetaT (burnInDoc uses it, lda.go:231)
 *  passes_out   optional D entries: burn-in passes each document executed */
int lda_b200_transform(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                       const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *theta_out,
                       int32_t *passes_out);

/* Perplexity, lda.go:368-389 */
int lda_b200_perplexity(lda_b200_handle *h, int64_t D, const int64_t *indptr, const int64_t *ind, const double *data,
                        const double *theta0, lda_b200_pcg *rnd, double rho_theta_t, double *out);

/* The same four calls taking the matrix in its own storage (see lda_b200_matrix); other arguments as above.
 * Transform / Perplexity fail with LDA_B200_EINVAL when m->rows differs from the fitted model's W. */
int lda_b200_fit_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *nphi0, const double *ntheta0,
                        lda_b200_pcg *rnd, lda_b200_counters *ctr, double *theta_out, double *perp_trace, int32_t perp_cap,
                        int32_t *n_evals, int32_t *iters_run);
int lda_b200_partial_fit_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *ntheta0, lda_b200_pcg *rnd,
                                lda_b200_counters *ctr, int32_t iterations, double *theta_out);
int lda_b200_transform_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                              double rho_theta_t, double *theta_out, int32_t *passes_out);
int lda_b200_perplexity_matrix(lda_b200_handle *h, const lda_b200_matrix *m, const double *theta0, lda_b200_pcg *rnd,
                               double rho_theta_t, double *out);

/* The same calls once more with the matrix descriptor spread over scalar arguments: format, rows, cols, ptr, ind, data
 * are the fields of lda_b200_matrix.  This is the form a cgo caller uses -- cgo forbids passing a pointer to Go memory
 * that itself holds Go pointers (a Go-side lda_b200_matrix filled with slice addresses), while every argument here is
 * a plain pointer to a flat slice or a scalar.  go/nlp/lda_b200.go binds exactly these. */
int lda_b200_fit_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr, const int64_t *ind,
                      const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd, lda_b200_counters *ctr,
                      double *theta_out, double *perp_trace, int32_t perp_cap, int32_t *n_evals, int32_t *iters_run);
int lda_b200_partial_fit_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                              const int64_t *ind, const double *data, const double *ntheta0, lda_b200_pcg *rnd,
                              lda_b200_counters *ctr, int32_t iterations, double *theta_out);
int lda_b200_transform_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                            const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                            double rho_theta_t, double *theta_out, int32_t *passes_out);
int lda_b200_perplexity_flat(lda_b200_handle *h, int32_t format, int64_t rows, int64_t cols, const int64_t *ptr,
                             const int64_t *ind, const double *data, const double *theta0, lda_b200_pcg *rnd,
                             double rho_theta_t, double *out);

/* Components, lda.go:414-416: out = K*W doubles */
int lda_b200_components(lda_b200_handle *h, double *out);

/* model persistence (absent in the reference: TODO at lda.go:158).  nphi = W*K [w*K+k],
 * nz = K; raw un-normalised statistics (lda.go:137,140). */
int lda_b200_get_dims(const lda_b200_handle *h, int64_t *W, int32_t *K);
int lda_b200_get_state(lda_b200_handle *h, double *nphi, double *nz);
int lda_b200_set_state(lda_b200_handle *h, int64_t W, const double *nphi, const double *nz);

/* Save / Load in the library's binary style (TfidfTransformer weightings.go:97-116, TruncatedSVD dimreduction.go:111-153:
 * little-endian scalars followed by gonum `MarshalBinaryTo` records).  Layout of an LDA model, all little endian:
 *     uint64  K
 *     float64 rhoPhiT, rhoThetaT, wordsInCorpus                       (the private counters, lda.go:117-120)
 *     mat.Dense record of nPhi (W x K, [w*K+k]) , mat.Dense record of nZ (1 x K)
 * where a mat.Dense record is gonum mat/io.go's storage header -- uint32 version 0xFE000001, bytes 'G','F','A', bool unit
 * = 0, int64 rows, cols, ku = 0, kl = 0 (40 bytes) -- followed by rows*cols float64 in row-major order.
 * lda_b200_save_size gives the byte count; lda_b200_save writes exactly that many bytes into buf (cap >= size);
 * lda_b200_load replaces the handle's model and configuration K by the file's and returns the counters.
 * (The reference has no LDA Save/Load and no gonum-produced bytes are available offline: the record layout follows
 * gonum's documented storage header; see DESIGN.md "parity unpinned" note.) */
int lda_b200_save_size(const lda_b200_handle *h, int64_t *bytes);
int lda_b200_save(lda_b200_handle *h, const lda_b200_counters *ctr, void *buf, int64_t cap, int64_t *written);
int lda_b200_load(lda_b200_handle *h, const void *buf, int64_t len, lda_b200_counters *ctr);

/* host-side restatement of Rnd.Int() % n draws (lda.go:201,425,474) for callers without
 * golang.org/x/exp/rand: out[i] = float64(Int() % n) / float64(n); advances *rnd. */
void lda_b200_pcg_seed(lda_b200_pcg *rnd, uint64_t seed);
uint64_t lda_b200_pcg_uint64(lda_b200_pcg *rnd);
void lda_b200_pcg_fill(lda_b200_pcg *rnd, int64_t count, int64_t n, double *out);
void lda_b200_pcg_advance(lda_b200_pcg *rnd, uint64_t draws);

/* Test/diagnostic hook for the "CSR/token handling is bit-exact" requirement: uploads the
 * CSC exactly as lda_b200_fit does and reads the device representation back.
 *  ind_out nnz int32, val_out nnz float, wc_out D doubles (lda.go:476-482), n_out = sum */
int lda_b200_ingest_roundtrip(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                              const double *data, int64_t *indptr_out, int32_t *ind_out, float *val_out,
                              double *wc_out, double *n_out);

/* The same for a matrix in its own storage: what the *_matrix calls hold on the device after the device-side
 * conversion.  Call once with only nnz_out to size the arrays (indptr_out cols+1, ind_out / val_out nnz, wc_out cols). */
int lda_b200_ingest_roundtrip_matrix(lda_b200_handle *h, const lda_b200_matrix *m, int64_t *nnz_out, int64_t *indptr_out,
                                     int32_t *ind_out, float *val_out, double *wc_out, double *n_out);

/* sparse.ToCSC() (lda.go:464-466) of a CSR term x document matrix, on the device: row_ptr[W+1], col_ind[nnz]
 * (document ids), data[nnz] -> indptr_out[D+1], ind_out[nnz] (word ids, ascending inside a document), data_out[nnz].
 * The result is what lda_b200_fit / transform / perplexity take; fewer than 2^31 entries. */
int lda_b200_csr_to_csc(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *row_ptr, const int64_t *col_ind,
                        const double *data, int64_t *indptr_out, int64_t *ind_out, double *data_out);

/* Resident-corpus API used by bench.py to time the fit loop with the inputs already in
 * HBM (the `value` leg); lda_b200_fit is these three calls back to back. */
int lda_b200_fit_begin(lda_b200_handle *h, int64_t W, int64_t D, const int64_t *indptr, const int64_t *ind,
                       const double *data, const double *nphi0, const double *ntheta0, lda_b200_pcg *rnd,
                       lda_b200_counters *ctr);
/* runs up to n_iterations more iterations of lda.go:501-540 on the resident corpus on the
 * handle's stream; *stopped = 1 when the perplexity early stop fired.  elapsed_ms (optional)
 * receives the CUDA-event time of the iterations on the launching stream. */
int lda_b200_fit_iterate(lda_b200_handle *h, int32_t n_iterations, lda_b200_counters *ctr, double *perp_trace,
                         int32_t perp_cap, int32_t *n_evals, int32_t *iters_run, int32_t *stopped,
                         float *elapsed_ms);
int lda_b200_fit_end(lda_b200_handle *h, double *theta_out);

/* Deterministic mode (not in the reference, whose goroutine pool is itself racy for Processes > 1): with on != 0 every
 * later fit on this handle is bitwise reproducible from run to run for the same inputs, configuration and number of
 * GPUs.  The sufficient statistics nPhiHat (lda.go:293) are then accumulated as 64-bit fixed-point integers, whose
 * atomic sums do not depend on the order in which documents finish, and converted to fp32 once per launch; all other
 * cross-block sums are always taken in a fixed order.  Costs one extra pass over nPhiHat per step and 8-byte instead of
 * 4-byte reductions.  Off by default: fp32 reductions in arrival order, results equal up to rounding. */
int lda_b200_set_deterministic(lda_b200_handle *h, int32_t on);

/* per-launch CUDA-event timing of the minibatch / blend kernels on the handle's stream (off by default) */
int lda_b200_set_profiling(lda_b200_handle *h, int32_t on);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t lda_b200_launch_count(const lda_b200_handle *h);
/* CUDA-event time (ms) spent in the fused SCVB0 minibatch kernel during the last
 * lda_b200_fit_iterate call, and the number of its launches (roofline leg of bench.py) */
int lda_b200_last_kernel_time(const lda_b200_handle *h, float *minibatch_ms, int32_t *minibatch_launches,
                              float *blend_ms, int32_t *blend_launches);

/* Diagnostic for the roofline leg of bench.py: what the memory system delivers for the access mix of the minibatch
 * kernel with the arithmetic and the per-document dependency chain taken away.  Replays the word ids of the first launch
 * range of the resident corpus (between lda_b200_fit_begin and lda_b200_fit_end): (BurnInPasses+1) coalesced reads of the
 * 4*K-byte row of C per entry (gather_ms), one red.global.add.v4.f32 of a row into nPhiHat per entry (red_ms; adds zeros,
 * the model is unchanged), and both together (mix_ms).  entries = entries replayed.  Best of three launches each. */
int lda_b200_probe_memory_mix(lda_b200_handle *h, int64_t *entries, float *gather_ms, float *red_ms, float *mix_ms);

/* which multi-GPU step path the last fit used: peer_memory_path = 2 when the reduction of nPhiHat and the M-step run as
 * one kernel whose reduction and broadcast happen in the NVSwitch (NVLink multicast: multimem.ld_reduce / multimem.st),
 * 1 when that kernel uses unicast peer loads / stores over CUDA IPC mappings (no multicast support, deterministic mode,
 * or LDA_B200_NVLS=0), 0 when ncclAllReduce + blend is used (fallback, or LDA_B200_PEER_MEMORY=0) */
int lda_b200_comm_info(const lda_b200_handle *h, int32_t *rank, int32_t *nranks, int32_t *peer_memory_path);
/* same for the per-step all-reduce (multi-GPU); the time includes waiting for the slowest rank's kernel */
int lda_b200_last_allreduce_time(const lda_b200_handle *h, float *allreduce_ms, int32_t *allreduce_calls);

/* phases of the peer-memory exchange during the last profiled lda_b200_fit_iterate call, total ms each: [0] opening
 * barrier (the wait for the slowest rank's minibatch kernel), [1] reduce + blend kernel over peer memory, [2] column
 * sums + closing barrier + nZ, [3] refresh of C */
int lda_b200_last_exchange_phases(const lda_b200_handle *h, float *phase_ms4);

#ifdef __cplusplus
}
#endif
#endif /* LDA_B200_H */
