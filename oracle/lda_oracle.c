/*
 * lda_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (float64, plain C) of james-bowman/nlp's LDA SCVB0 path, used ONLY
 * as the parity checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * / --impl reference legs.  Nothing under nlp_b200/ may import, link or call it.
 *
 * The reference is Go (lda.go, utils.go) and no Go toolchain exists in this image, so
 * the reference itself cannot be compiled or run here.  Parity of this
 * restatement is pinned against the reference's own tests (lda_test.go:16-235) -- see
 * tests/test_oracle_golden.py -- including the seed-dependent topic permutation, which
 * indirectly pins the RNG stream (golang.org/x/exp/rand, unpinned HEAD dependency whose
 * source is not under /root/reference; its published PCG XSL-RR 128/64 algorithm is
 * restated in pcg_* below).
 *
 * Every function cites the reference lines it follows.  All arithmetic is float64 and
 * keeps the reference's loop structure, including its quirks:
 * per-k math.Pow calls, nPhiHat ignoring the count v, wordsInCorpus/rhoThetaT never
 * reset, short last minibatch dividing by its real size, etc.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------
 * golang.org/x/exp/rand: rand.New(rand.NewSource(seed)) -> PCGSource (PCG XSL-RR 128/64)
 * call sites in the reference: lda.go:184,201,425,474; tests seed 0 (lda_test.go:68).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint64_t hi, lo;
} oracle_pcg;

typedef unsigned __int128 u128;

static const uint64_t PCG_MUL_HI = 2549297995355413924ULL, PCG_MUL_LO = 4865540595714422341ULL;
static const uint64_t PCG_INC_HI = 6364136223846793005ULL, PCG_INC_LO = 1442695040888963407ULL;

void oracle_pcg_seed(oracle_pcg *p, uint64_t seed) {
    /* PCGSource.Seed: low = seed; high = seed */
    p->lo = seed;
    p->hi = seed;
}

uint64_t oracle_pcg_uint64(oracle_pcg *p) {
    u128 s = ((u128)p->hi << 64) | p->lo;
    u128 m = ((u128)PCG_MUL_HI << 64) | PCG_MUL_LO;
    u128 c = ((u128)PCG_INC_HI << 64) | PCG_INC_LO;
    s = s * m + c;
    p->hi = (uint64_t)(s >> 64);
    p->lo = (uint64_t)s;
    /* bits.RotateLeft64(high^low, -int(high>>58)) == rotate right by high>>58 */
    uint64_t x = p->hi ^ p->lo;
    unsigned r = (unsigned)(p->hi >> 58);
    return (x >> r) | (x << ((64 - r) & 63));
}

/* rand.Rand.Int() on a 64-bit platform: Int63() = Uint64() & (1<<63 - 1) */
int64_t oracle_pcg_int(oracle_pcg *p) { return (int64_t)(oracle_pcg_uint64(p) & 0x7fffffffffffffffULL); }

/* out[i] = float64(Rnd.Int() % n) / float64(n)   (lda.go:201, 425, 474) */
void oracle_pcg_fill(oracle_pcg *p, int64_t count, int64_t n, double *out) {
    for (int64_t i = 0; i < count; i++) out[i] = (double)(oracle_pcg_int(p) % n) / (double)n;
}

/* ------------------------------------------------------------------------------------
 * Model.  Field names follow lda.go:68-141.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    double S, Tau, Kappa;
} oracle_schedule;

typedef struct {
    /* exported fields, lda.go:69-134 */
    int64_t Iterations;
    double PerplexityTolerance;
    int64_t PerplexityEvaluationFrequency;
    int64_t BatchSize;
    int64_t K;
    int64_t BurnInPasses;
    int64_t TransformationPasses;
    double MeanChangeTolerance;
    int64_t ChangeEvaluationFrequency;
    double Alpha, Eta;
    oracle_schedule RhoPhi, RhoTheta;
    int64_t Processes;
    /* private state, lda.go:117-140 */
    double rhoPhiT, rhoThetaT;
    double wordsInCorpus;
    int64_t w, d;
    oracle_pcg Rnd;
    double *nPhi; /* [w*K + k] */
    double *nZ;   /* [k] */
    pthread_mutex_t phiMutex, zMutex;
    /* instrumentation (not in the reference): perplexity evaluated inside FitTransform
     * (lda.go:530-539 computes it but never exposes it) and iterations actually run */
    double *perpTrace;
    int64_t perpTraceCap, perpTraceLen;
    int64_t iterationsRun;
    int64_t tokenVisits; /* non-zero visits executed by the last Fit/Transform call */
    double *iterTime;    /* seconds since the start of the iteration loop at the end of each iteration */
    int64_t iterTimeCap;
    /* SyncGroup = G > 1 (not in the reference): deterministic restatement of G workers running in lock step --
     * minibatches sG..sG+G-1 are all computed from the same nPhi/nZ snapshot, then blended one after the other
     * in index order (lda.go:299-317 G times).  This is the semantics of the G-GPU path; the
     * reference's own Processes > 1 schedule is racy and not reproducible. */
    int64_t SyncGroup;
    /* HoistPow != 0 (not in the reference): pow((1 - rhoTheta), v) is evaluated once per token instead of 2K times
     * (lda.go:247-248 call it inside the loop over k).  pow is a pure function of its two arguments, so every result
     * keeps its bits (tests/test_oracle_golden.py checks that); it only makes the oracle fast enough to check whole
     * BASELINE configurations.  The CPU baseline legs of bench.py leave it off: they time the reference's loop as is. */
    int64_t HoistPow;
} oracle_lda;

/* term x document matrix in CSC: column j = document j, rows = word ids, in stored order.
 * A dense mat.Matrix is equivalent to the CSC that lists each column's non-zero rows in
 * ascending order (utils.go:51-57 scans rows ascending and skips exact zeros). */
typedef struct {
    int64_t rows, cols;
    const int64_t *indptr; /* cols+1 */
    const int64_t *ind;    /* nnz */
    const double *data;    /* nnz */
} oracle_csc;

/* lda.go:30-32 */
static double schedule_calc(oracle_schedule s, double iteration) { return s.S / pow(s.Tau + iteration, s.Kappa); }

/* lda.go:145-189 (defaults).  Rnd is time-seeded in the reference; here the caller seeds. */
oracle_lda *oracle_lda_new(int64_t k, uint64_t seed, int64_t processes) {
    oracle_lda *l = (oracle_lda *)calloc(1, sizeof(oracle_lda));
    l->Iterations = 1000;
    l->PerplexityTolerance = 1e-2;
    l->PerplexityEvaluationFrequency = 30;
    l->BatchSize = 100;
    l->K = k;
    l->BurnInPasses = 1;
    l->TransformationPasses = 500;
    l->MeanChangeTolerance = 1e-5;
    l->ChangeEvaluationFrequency = 30;
    l->Alpha = 0.1;
    l->Eta = 0.01;
    l->RhoPhi = (oracle_schedule){10, 1000, 0.9};
    l->RhoTheta = (oracle_schedule){1, 10, 0.9};
    l->rhoPhiT = 1;
    l->rhoThetaT = 1;
    oracle_pcg_seed(&l->Rnd, seed);
    l->Processes = processes;
    pthread_mutex_init(&l->phiMutex, NULL);
    pthread_mutex_init(&l->zMutex, NULL);
    return l;
}

void oracle_lda_free(oracle_lda *l) {
    if (!l) return;
    free(l->nPhi);
    free(l->nZ);
    free(l->perpTrace);
    free(l->iterTime);
    pthread_mutex_destroy(&l->phiMutex);
    pthread_mutex_destroy(&l->zMutex);
    free(l);
}

/* lda.go:193-206.  If nphi0 != NULL the values are taken from it instead of the RNG
 * (used to replay an init exported to the GPU side); the RNG is then NOT advanced. */
static void lda_init(oracle_lda *l, const oracle_csc *m, const double *nphi0) {
    int64_t r = m->rows, c = m->cols, K = l->K;
    l->w = r;
    l->d = c;
    free(l->nPhi);
    free(l->nZ);
    l->nPhi = (double *)malloc(sizeof(double) * (size_t)(K * r));
    l->nZ = (double *)calloc((size_t)K, sizeof(double));
    for (int64_t i = 0; i < r; i++) {
        for (int64_t k = 0; k < K; k++) {
            double v;
            if (nphi0)
                v = nphi0[i * K + k];
            else
                v = (double)(oracle_pcg_int(&l->Rnd) % (r * K)) / (double)(r * K);
            l->nPhi[i * K + k] = v;
            l->nZ[k] += v;
        }
    }
}

/* One token update shared by burnInDoc (lda.go:232-250) and fitMiniBatch (lda.go:275-297).
 * The k-loops and the two math.Pow calls per k are kept as in the reference. */
static inline void token_update(const oracle_lda *l, int64_t i, int64_t j, double v, double rhoTheta, double wc,
                                double *gamma, double *nTheta) {
    int64_t K = l->K;
    double gammaSum = 0;
    for (int64_t k = 0; k < K; k++) {
        /* Eqn. 5 */
        gamma[k] = ((l->nPhi[i * K + k] + l->Eta) * (nTheta[j * K + k] + l->Alpha) / (l->nZ[k] + l->Eta * (double)l->w));
        gammaSum += gamma[k];
    }
    for (int64_t k = 0; k < K; k++) gamma[k] /= gammaSum;
    if (l->HoistPow) {
        const double pw = pow((1.0 - rhoTheta), v);
        for (int64_t k = 0; k < K; k++) {
            int64_t thetaInd = j * K + k;
            nTheta[thetaInd] = ((pw * nTheta[thetaInd]) + ((1 - pw) * wc * gamma[k]));
        }
        return;
    }
    for (int64_t k = 0; k < K; k++) {
        /* Eqn. 9 */
        int64_t thetaInd = j * K + k;
        nTheta[thetaInd] = ((pow((1.0 - rhoTheta), v) * nTheta[thetaInd]) + ((1 - pow((1.0 - rhoTheta), v)) * wc * gamma[k]));
    }
}

/* lda.go:218-261 */
static int64_t burn_in_doc(oracle_lda *l, int64_t j, int64_t iterations, const oracle_csc *m, double wc, double *gamma,
                           double *nTheta) {
    int64_t K = l->K;
    double sum, prevSum = 0;
    int64_t visits = 0;
    for (int64_t counter = 1; counter <= iterations; counter++) {
        if (l->ChangeEvaluationFrequency > 0 && counter % l->ChangeEvaluationFrequency == 0 && 1 < iterations) {
            prevSum = 0;
            for (int64_t k = 0; k < K; k++) prevSum += nTheta[j * K + k];
        }
        double rhoTheta = schedule_calc(l->RhoTheta, l->rhoThetaT + (double)counter);
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) { /* ColNonZeroElemDo, utils.go:45-59 */
            token_update(l, m->ind[p], j, m->data[p], rhoTheta, wc, gamma, nTheta);
            visits++;
        }
        if (l->ChangeEvaluationFrequency > 0 && counter % l->ChangeEvaluationFrequency == 0 && counter < iterations) {
            sum = 0;
            for (int64_t k = 0; k < K; k++) sum += nTheta[j * K + k];
            if (fabs(sum - prevSum) / (double)K < l->MeanChangeTolerance) break;
        }
    }
    return visits;
}

/* lda.go:34-58 */
typedef struct {
    int64_t start, end;
    double *nPhiHat, *nZHat, *gamma;
    int64_t visits;
} mini_batch;

static void mini_batch_reset(mini_batch *b, int64_t K, int64_t w) {
    for (int64_t i = 0; i < K * w; i++) b->nPhiHat[i] = 0;
    for (int64_t i = 0; i < K; i++) b->nZHat[i] = 0;
}

/* first half of fitMiniBatch, lda.go:266-298 */
static void estep_mini_batch(oracle_lda *l, mini_batch *mb, const double *wc, double *nTheta, const oracle_csc *m) {
    int64_t K = l->K;
    int64_t batchSize = mb->end - mb->start;
    for (int64_t j = mb->start; j < mb->end; j++) {
        mb->visits += burn_in_doc(l, j, l->BurnInPasses, m, wc[j], mb->gamma, nTheta);
        double rhoTheta = schedule_calc(l->RhoTheta, l->rhoThetaT + (double)l->BurnInPasses);
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) {
            int64_t i = m->ind[p];
            double v = m->data[p];
            double gammaSum = 0;
            double *gamma = mb->gamma;
            for (int64_t k = 0; k < K; k++) {
                gamma[k] = ((l->nPhi[i * K + k] + l->Eta) * (nTheta[j * K + k] + l->Alpha) / (l->nZ[k] + l->Eta * (double)l->w));
                gammaSum += gamma[k];
            }
            for (int64_t k = 0; k < K; k++) gamma[k] /= gammaSum;
            const double pw = l->HoistPow ? pow((1.0 - rhoTheta), v) : 0.0;
            for (int64_t k = 0; k < K; k++) {
                int64_t thetaInd = j * K + k;
                if (l->HoistPow)
                    nTheta[thetaInd] = ((pw * nTheta[thetaInd]) + ((1 - pw) * wc[j] * gamma[k]));
                else
                nTheta[thetaInd] =
                    ((pow((1.0 - rhoTheta), v) * nTheta[thetaInd]) + ((1 - pow((1.0 - rhoTheta), v)) * wc[j] * gamma[k]));
                /* sufficient stats: note there is no factor v (lda.go:293) */
                double nv = l->wordsInCorpus * gamma[k] / (double)batchSize;
                mb->nPhiHat[i * K + k] += nv;
                mb->nZHat[k] += nv;
            }
            mb->visits++;
        }
    }
}

/* second half of fitMiniBatch, lda.go:299-317 */
static void blend_mini_batch(oracle_lda *l, mini_batch *mb) {
    int64_t K = l->K;
    /* lda.go:299-300 -- outside any lock in the reference (racy when Processes>1) */
    double rhoPhi = schedule_calc(l->RhoPhi, l->rhoPhiT);
    l->rhoPhiT++;

    pthread_mutex_lock(&l->phiMutex); /* Eqn. 7 */
    for (int64_t w = 0; w < l->w; w++)
        for (int64_t k = 0; k < K; k++) {
            int64_t phiInd = w * K + k;
            l->nPhi[phiInd] = ((1.0 - rhoPhi) * l->nPhi[phiInd]) + (rhoPhi * mb->nPhiHat[phiInd]);
        }
    pthread_mutex_unlock(&l->phiMutex);

    pthread_mutex_lock(&l->zMutex); /* Eqn. 8 */
    for (int64_t k = 0; k < K; k++) l->nZ[k] = ((1.0 - rhoPhi) * l->nZ[k]) + (rhoPhi * mb->nZHat[k]);
    pthread_mutex_unlock(&l->zMutex);
}

/* lda.go:323-340 (adjustment = 0.0); result may alias theta */
static void normalise_theta(const oracle_lda *l, const double *theta, int64_t c, double *result) {
    int64_t K = l->K;
    for (int64_t j = 0; j < c; j++) {
        double sum = 0;
        for (int64_t k = 0; k < K; k++) sum += theta[j * K + k] + 0.0;
        for (int64_t k = 0; k < K; k++) result[j * K + k] = (theta[j * K + k] + 0.0) / sum;
    }
}

/* lda.go:345-363 */
static void normalise_phi(const oracle_lda *l, const double *phi, double *result) {
    int64_t K = l->K;
    double *sum = (double *)calloc((size_t)K, sizeof(double));
    for (int64_t i = 0; i < l->w; i++)
        for (int64_t k = 0; k < K; k++) sum[k] += phi[i * K + k] + 0.0;
    for (int64_t i = 0; i < l->w; i++)
        for (int64_t k = 0; k < K; k++) result[i * K + k] = (phi[i * K + k] + 0.0) / sum[k];
    free(sum);
}

/* lda.go:392-408 */
static double perplexity_of(const oracle_lda *l, const oracle_csc *m, double sum, const double *nTheta, const double *nPhi) {
    int64_t K = l->K;
    double ttlLogWordProb = 0;
    for (int64_t j = 0; j < m->cols; j++) {
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) {
            int64_t i = m->ind[p];
            double dot = 0;
            for (int64_t k = 0; k < K; k++) dot += nPhi[i * K + k] * nTheta[j * K + k];
            ttlLogWordProb += log2(dot) * m->data[p];
        }
    }
    return exp2(-ttlLogWordProb / sum);
}

/* goroutine pool of lda.go:504-528: `processes` workers receive consecutive minibatch
 * indices from an unbuffered channel; restated as a shared counter handing out
 * consecutive indices to whichever worker asks next. */
typedef struct {
    oracle_lda *l;
    mini_batch *mb;
    const oracle_csc *m;
    const double *wc;
    double *nTheta;
    int64_t numMiniBatches;
    int64_t *next;
    pthread_mutex_t *chan;
} worker_arg;

static void *worker_main(void *p) {
    worker_arg *a = (worker_arg *)p;
    oracle_lda *l = a->l;
    for (;;) {
        pthread_mutex_lock(a->chan);
        int64_t j = (*a->next)++;
        pthread_mutex_unlock(a->chan);
        if (j >= a->numMiniBatches) break;
        mini_batch_reset(a->mb, l->K, l->w);
        a->mb->start = j * l->BatchSize;
        if (j < a->numMiniBatches - 1)
            a->mb->end = a->mb->start + l->BatchSize;
        else
            a->mb->end = a->m->cols;
        estep_mini_batch(l, a->mb, a->wc, a->nTheta, a->m);
        blend_mini_batch(l, a->mb);
    }
    return NULL;
}

/* SyncGroup: the E-steps of one group read the same nPhi/nZ snapshot and write disjoint documents / private
 * statistics, so they may run on separate host threads without changing any value */
static void *estep_main(void *p) {
    worker_arg *a = (worker_arg *)p;
    estep_mini_batch(a->l, a->mb, a->wc, a->nTheta, a->m);
    return NULL;
}

static void trace_push(oracle_lda *l, double v) {
    if (l->perpTraceLen == l->perpTraceCap) {
        l->perpTraceCap = l->perpTraceCap ? 2 * l->perpTraceCap : 64;
        l->perpTrace = (double *)realloc(l->perpTrace, sizeof(double) * (size_t)l->perpTraceCap);
    }
    l->perpTrace[l->perpTraceLen++] = v;
}

/* lda.go:463-542.  theta_out: K x c row-major (DenseCopyOf(NewDense(c,K,..).T())).
 * nphi0 / ntheta0 (optional) replace the RNG draws of lda.go:201 / :474. */
int oracle_lda_fit_transform(oracle_lda *l, const oracle_csc *m, const double *nphi0, const double *ntheta0,
                             double *theta_out) {
    int64_t K = l->K;
    lda_init(l, m, nphi0);
    int64_t c = m->cols;

    double *nTheta = (double *)malloc(sizeof(double) * (size_t)(K * c));
    for (int64_t i = 0; i < K * c; i++) {
        if (ntheta0)
            nTheta[i] = ntheta0[i];
        else
            nTheta[i] = (double)(oracle_pcg_int(&l->Rnd) % (c * K)) / (double)(c * K);
    }
    double *wc = (double *)calloc((size_t)c, sizeof(double));
    for (int64_t j = 0; j < c; j++) {
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) wc[j] += m->data[p];
        l->wordsInCorpus += wc[j]; /* never reset: lda.go:481 */
    }

    double *phiProb = (double *)malloc(sizeof(double) * (size_t)(K * l->w));
    double *thetaProb = (double *)malloc(sizeof(double) * (size_t)(K * c));

    int64_t numMiniBatches = (int64_t)ceil((double)c / (double)l->BatchSize);
    int64_t processes = l->Processes;
    if (l->SyncGroup > 1) processes = l->SyncGroup;
    if (numMiniBatches < processes) processes = numMiniBatches;
    if (processes < 1) processes = 1;
    mini_batch *miniBatches = (mini_batch *)calloc((size_t)processes, sizeof(mini_batch));
    for (int64_t i = 0; i < processes; i++) {
        miniBatches[i].nPhiHat = (double *)malloc(sizeof(double) * (size_t)(K * l->w));
        miniBatches[i].nZHat = (double *)malloc(sizeof(double) * (size_t)K);
        miniBatches[i].gamma = (double *)malloc(sizeof(double) * (size_t)K);
    }

    l->rhoPhiT = 1;
    double perplexity = 0, prevPerplexity = 0;
    l->perpTraceLen = 0;
    l->iterationsRun = 0;
    l->tokenVisits = 0;

    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)processes);
    worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)processes);
    pthread_mutex_t chan;
    pthread_mutex_init(&chan, NULL);

    if (l->iterTimeCap < l->Iterations) {
        l->iterTime = (double *)realloc(l->iterTime, sizeof(double) * (size_t)(l->Iterations + 1));
        l->iterTimeCap = l->Iterations;
    }
    struct timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);

    for (int64_t it = 0; it < l->Iterations; it++) {
        l->rhoThetaT++;
        int64_t next = 0;
        if (l->SyncGroup > 1) {
            for (int64_t j0 = 0; j0 < numMiniBatches; j0 += processes) {
                int64_t n = numMiniBatches - j0 < processes ? numMiniBatches - j0 : processes;
                for (int64_t g = 0; g < n; g++) {
                    mini_batch *mb = &miniBatches[g];
                    int64_t j = j0 + g;
                    mini_batch_reset(mb, K, l->w);
                    mb->start = j * l->BatchSize;
                    mb->end = j < numMiniBatches - 1 ? mb->start + l->BatchSize : c;
                    args[g] = (worker_arg){l, mb, m, wc, nTheta, numMiniBatches, &next, &chan};
                    if (n > 1) pthread_create(&threads[g], NULL, estep_main, &args[g]);
                    else estep_main(&args[g]);
                }
                if (n > 1)
                    for (int64_t g = 0; g < n; g++) pthread_join(threads[g], NULL);
                for (int64_t g = 0; g < n; g++) blend_mini_batch(l, &miniBatches[g]);
            }
        } else
        for (int64_t p = 0; p < processes; p++) {
            args[p] = (worker_arg){l, &miniBatches[p], m, wc, nTheta, numMiniBatches, &next, &chan};
            if (processes > 1) pthread_create(&threads[p], NULL, worker_main, &args[p]);
        }
        if (l->SyncGroup > 1) {
        } else if (processes > 1)
            for (int64_t p = 0; p < processes; p++) pthread_join(threads[p], NULL);
        else
            worker_main(&args[0]);
        l->iterationsRun = it + 1;
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        l->iterTime[it] = (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);

        if (l->PerplexityEvaluationFrequency > 0 && (it + 1) % l->PerplexityEvaluationFrequency == 0) {
            normalise_phi(l, l->nPhi, phiProb);
            normalise_theta(l, nTheta, c, thetaProb);
            perplexity = perplexity_of(l, m, l->wordsInCorpus, thetaProb, phiProb);
            trace_push(l, perplexity);
            if (prevPerplexity != 0 && fabs(prevPerplexity - perplexity) < l->PerplexityTolerance) break;
            prevPerplexity = perplexity;
        }
    }
    for (int64_t p = 0; p < processes; p++) l->tokenVisits += miniBatches[p].visits;

    normalise_theta(l, nTheta, c, thetaProb);
    for (int64_t j = 0; j < c; j++)
        for (int64_t k = 0; k < K; k++) theta_out[k * c + j] = thetaProb[j * K + k];

    pthread_mutex_destroy(&chan);
    free(threads);
    free(args);
    for (int64_t i = 0; i < processes; i++) {
        free(miniBatches[i].nPhiHat);
        free(miniBatches[i].nZHat);
        free(miniBatches[i].gamma);
    }
    free(miniBatches);
    free(phiProb);
    free(thetaProb);
    free(wc);
    free(nTheta);
    return 0;
}

/* lda.go:420-437; theta0 (optional) replaces the RNG draws of :425.  passes_out (optional,
 * instrumentation) receives the number of burn-in passes each document executed. */
static double *un_normalised_transform(oracle_lda *l, const oracle_csc *m, const double *theta0, int64_t *passes_out) {
    int64_t K = l->K, c = m->cols;
    double *theta = (double *)malloc(sizeof(double) * (size_t)(K * c));
    for (int64_t i = 0; i < K * c; i++) {
        if (theta0)
            theta[i] = theta0[i];
        else
            theta[i] = (double)(oracle_pcg_int(&l->Rnd) % (c * K)) / (double)(c * K);
    }
    double *gamma = (double *)malloc(sizeof(double) * (size_t)K);
    l->tokenVisits = 0;
    for (int64_t j = 0; j < c; j++) {
        double wc = 0;
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) wc += m->data[p];
        int64_t visits = burn_in_doc(l, j, l->TransformationPasses, m, wc, gamma, theta);
        l->tokenVisits += visits;
        if (passes_out) {
            int64_t n = m->indptr[j + 1] - m->indptr[j];
            passes_out[j] = n > 0 ? visits / n : -1;
        }
    }
    free(gamma);
    return theta;
}

/* lda.go:446-453 */
int oracle_lda_transform(oracle_lda *l, const oracle_csc *m, const double *theta0, double *theta_out, int64_t *passes_out) {
    int64_t K = l->K, c = m->cols;
    double *theta = un_normalised_transform(l, m, theta0, passes_out);
    normalise_theta(l, theta, c, theta);
    for (int64_t j = 0; j < c; j++)
        for (int64_t k = 0; k < K; k++) theta_out[k * c + j] = theta[j * K + k];
    free(theta);
    return 0;
}

/* lda.go:368-389 */
double oracle_lda_perplexity(oracle_lda *l, const oracle_csc *m, const double *theta0) {
    double wordCount = 0;
    for (int64_t j = 0; j < m->cols; j++)
        for (int64_t p = m->indptr[j]; p < m->indptr[j + 1]; p++) wordCount += m->data[p];
    double *theta = un_normalised_transform(l, m, theta0, NULL);
    normalise_theta(l, theta, m->cols, theta);
    double *phi = (double *)malloc(sizeof(double) * (size_t)(l->K * l->w));
    normalise_phi(l, l->nPhi, phi);
    double r = perplexity_of(l, m, wordCount, theta, phi);
    free(phi);
    free(theta);
    return r;
}

/* lda.go:414-416: out is K x W row-major */
int oracle_lda_components(oracle_lda *l, double *out) {
    int64_t K = l->K;
    double *phi = (double *)malloc(sizeof(double) * (size_t)(K * l->w));
    normalise_phi(l, l->nPhi, phi);
    for (int64_t i = 0; i < l->w; i++)
        for (int64_t k = 0; k < K; k++) out[k * l->w + i] = phi[i * K + k];
    free(phi);
    return 0;
}

/* accessors for the ctypes wrapper (avoids mirroring the struct layout in Python) */
#define FIELD_RW(type, name)                                                  \
    void oracle_lda_set_##name(oracle_lda *l, type v) { l->name = v; }        \
    type oracle_lda_get_##name(const oracle_lda *l) { return l->name; }
FIELD_RW(int64_t, Iterations)
FIELD_RW(double, PerplexityTolerance)
FIELD_RW(int64_t, PerplexityEvaluationFrequency)
FIELD_RW(int64_t, BatchSize)
FIELD_RW(int64_t, K)
FIELD_RW(int64_t, BurnInPasses)
FIELD_RW(int64_t, TransformationPasses)
FIELD_RW(double, MeanChangeTolerance)
FIELD_RW(int64_t, ChangeEvaluationFrequency)
FIELD_RW(double, Alpha)
FIELD_RW(double, Eta)
FIELD_RW(int64_t, Processes)
FIELD_RW(int64_t, SyncGroup)
FIELD_RW(int64_t, HoistPow)
FIELD_RW(double, rhoPhiT)
FIELD_RW(double, rhoThetaT)
FIELD_RW(double, wordsInCorpus)
void oracle_lda_set_RhoPhi(oracle_lda *l, double S, double Tau, double Kappa) { l->RhoPhi = (oracle_schedule){S, Tau, Kappa}; }
void oracle_lda_set_RhoTheta(oracle_lda *l, double S, double Tau, double Kappa) { l->RhoTheta = (oracle_schedule){S, Tau, Kappa}; }
void oracle_lda_seed(oracle_lda *l, uint64_t seed) { oracle_pcg_seed(&l->Rnd, seed); }
void oracle_lda_get_rnd(const oracle_lda *l, uint64_t *hi, uint64_t *lo) {
    *hi = l->Rnd.hi;
    *lo = l->Rnd.lo;
}
int64_t oracle_lda_get_w(const oracle_lda *l) { return l->w; }
int64_t oracle_lda_get_d(const oracle_lda *l) { return l->d; }
int64_t oracle_lda_get_iterationsRun(const oracle_lda *l) { return l->iterationsRun; }
int64_t oracle_lda_get_tokenVisits(const oracle_lda *l) { return l->tokenVisits; }
int64_t oracle_lda_get_perpTrace(const oracle_lda *l, double *out, int64_t cap) {
    int64_t n = l->perpTraceLen < cap ? l->perpTraceLen : cap;
    if (out) memcpy(out, l->perpTrace, sizeof(double) * (size_t)n);
    return l->perpTraceLen;
}
int64_t oracle_lda_get_iterTime(const oracle_lda *l, double *out, int64_t cap) {
    int64_t n = l->iterationsRun < cap ? l->iterationsRun : cap;
    if (out && l->iterTime) memcpy(out, l->iterTime, sizeof(double) * (size_t)n);
    return l->iterationsRun;
}
/* raw (un-normalised) state, for parity checks and for exporting a trained model */
void oracle_lda_get_state(const oracle_lda *l, double *nphi, double *nz) {
    if (nphi) memcpy(nphi, l->nPhi, sizeof(double) * (size_t)(l->K * l->w));
    if (nz) memcpy(nz, l->nZ, sizeof(double) * (size_t)l->K);
}
void oracle_lda_set_state(oracle_lda *l, int64_t w, const double *nphi, const double *nz) {
    free(l->nPhi);
    free(l->nZ);
    l->w = w;
    l->nPhi = (double *)malloc(sizeof(double) * (size_t)(l->K * w));
    l->nZ = (double *)malloc(sizeof(double) * (size_t)l->K);
    memcpy(l->nPhi, nphi, sizeof(double) * (size_t)(l->K * w));
    memcpy(l->nZ, nz, sizeof(double) * (size_t)l->K);
}
double oracle_schedule_calc(double S, double Tau, double Kappa, double iteration) {
    return schedule_calc((oracle_schedule){S, Tau, Kappa}, iteration);
}
