/* go_shim_sequence_test.c -- the call sequence go/nlp/lda_b200.go encodes, executed from plain C.
 *
 * The Go shim cannot be compiled in this image (no Go toolchain), so the sequence of C-ABI calls each of its methods
 * makes is replayed here, call for call, the way cgo would issue them: every buffer is malloc'ed (pageable, like a Go
 * slice), matrices travel as the scalar arguments of the lda_b200_*_flat entry points (no descriptor struct holding
 * pointers), and -- as in lda_test.go:68 -- Rnd has been replaced by the caller, so the initial values are drawn on the
 * host in the reference's order (draws()) and passed as nphi0 / ntheta0 / theta0.  The three tests of lda_test.go run
 * on top of that sequence with the reference's fixtures, seed and tolerances; then PartialFit and Save / Load.
 * Built by tests/test_cpp_host.py (gcc, CPU), run on the B200 box (-m gpu).  Exit code = number of failures. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lda_b200.h"

static int failures = 0;
#define Errorf(...)                   \
    do {                              \
        fprintf(stderr, __VA_ARGS__); \
        ++failures;                   \
    } while (0)

static const double dataA[81] = {3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3,
                                 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4};
static const double dataB[81] = {3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 5, 1, 0, 0, 0, 0, 0, 0, 3, 5,
                                 0, 0, 0, 0, 0, 0, 0, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4};
static const double topicsA[3][9] = {{0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33}, {0, 0, 0, 0.33, 0.33, 0.33, 0, 0, 0}};
static const double topicsB[3][9] = {{0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33}, {0, 0, 0, 0.428, 0.285, 0.285, 0, 0, 0}};
static const double docs[9][3] = {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};

/* the Go object: exported fields live in cfg, the private ones next to it (go/nlp/lda_b200.go) */
typedef struct {
    lda_b200_config cfg;
    lda_b200_counters ctr; /* rhoPhiT, rhoThetaT, wordsInCorpus */
    lda_b200_pcg own;      /* state of the shim-owned PCGSource */
    lda_b200_pcg rnd;      /* the caller's replacement generator: rand.New(rand.NewSource(0)) */
    int rnd_replaced;
    int64_t w, d;
    lda_b200_handle *h;
} shim;

static void shim_new(shim *l, int k) { /* NewLatentDirichletAllocation */
    memset(l, 0, sizeof *l);
    lda_b200_default_config(k, &l->cfg);
    l->ctr.rho_phi_t = 1;
    l->ctr.rho_theta_t = 1;
    lda_b200_pcg_seed(&l->own, 0x1234abcdu); /* time.Now().UnixNano() */
}

static lda_b200_handle *shim_handle(shim *l) { /* handle(): create or set_config, then set_deterministic */
    if (!l->h) {
        if (lda_b200_create(&l->cfg, 0, &l->h)) {
            fprintf(stderr, "nlp: %s\n", lda_b200_last_error(NULL));
            exit(100);
        }
    } else if (lda_b200_set_config(l->h, &l->cfg)) {
        Errorf("set_config: %s\n", lda_b200_last_error(l->h));
    }
    lda_b200_set_deterministic(l->h, 0);
    return l->h;
}

/* draws(): rows*K values float64(Rnd.Int() % n) / float64(n) from the caller's generator, NULL while Rnd is the shim's own */
static double *shim_draws(shim *l, int64_t rows) {
    if (!l->rnd_replaced) return NULL;
    const int64_t n = rows * l->cfg.K;
    double *out = (double *)malloc(sizeof(double) * (size_t)n);
    lda_b200_pcg_fill(&l->rnd, n, n, out);
    return out;
}

/* fit(): dense input -> LDA_B200_DENSE; a malloc'ed copy stands in for the Go slice */
static double *shim_fit(shim *l, int r, int c, const double *data, int want) {
    lda_b200_handle *h = shim_handle(l);
    double *m = (double *)malloc(sizeof(double) * (size_t)(r * c));
    memcpy(m, data, sizeof(double) * (size_t)(r * c));
    l->w = r;
    l->d = c;
    double *theta = want ? (double *)malloc(sizeof(double) * (size_t)(l->cfg.K * c)) : NULL;
    double *nphi0 = shim_draws(l, r), *ntheta0 = shim_draws(l, c);
    lda_b200_pcg rnd = l->own;
    int rc = lda_b200_fit_flat(h, LDA_B200_DENSE, r, c, NULL, NULL, m, nphi0, ntheta0, &rnd, &l->ctr, theta, NULL, 0, NULL, NULL);
    if (rc) Errorf("fit: %s\n", lda_b200_last_error(h));
    l->own = rnd;
    free(m);
    free(nphi0);
    free(ntheta0);
    return theta;
}

static double *shim_transform(shim *l, int r, int c, const double *data) {
    lda_b200_handle *h = shim_handle(l);
    double *m = (double *)malloc(sizeof(double) * (size_t)(r * c));
    memcpy(m, data, sizeof(double) * (size_t)(r * c));
    double *theta = (double *)malloc(sizeof(double) * (size_t)(l->cfg.K * c));
    double *theta0 = shim_draws(l, c);
    lda_b200_pcg rnd = l->own;
    int rc = lda_b200_transform_flat(h, LDA_B200_DENSE, r, c, NULL, NULL, m, theta0, &rnd, l->ctr.rho_theta_t, theta, NULL);
    if (rc) Errorf("transform: %s\n", lda_b200_last_error(h));
    l->own = rnd;
    free(m);
    free(theta0);
    return theta;
}

static double *shim_components(shim *l) {
    lda_b200_handle *h = shim_handle(l);
    double *out = (double *)malloc(sizeof(double) * (size_t)(l->cfg.K * l->w));
    if (lda_b200_components(h, out)) Errorf("components: %s\n", lda_b200_last_error(h));
    return out;
}

static void shim_close(shim *l) {
    lda_b200_destroy(l->h);
    l->h = NULL;
}

static void set_seed0(shim *l) { /* lda.Rnd = rand.New(rand.NewSource(uint64(0))), lda_test.go:68 */
    lda_b200_pcg_seed(&l->rnd, 0);
    l->rnd_replaced = 1;
}

static void TestLDAFit(void) { /* lda_test.go:16-89 */
    for (int ti = 0; ti < 2; ++ti) {
        shim lda;
        shim_new(&lda, 3);
        set_seed0(&lda);
        shim_fit(&lda, 9, 9, ti ? dataB : dataA, 0);
        double *components = shim_components(&lda);
        for (int i = 0; i < 3; ++i) {
            double sum = 0;
            for (int ri = 0; ri < 9; ++ri) {
                const double cv = components[i * 9 + ri], v = (ti ? topicsB : topicsA)[i][ri];
                sum += cv;
                if (fabs(cv - v) > 0.01) Errorf("Test %d: Topic (%d) over word (%d) distribution incorrect. Expected %f but received %f\n", ti, i, ri, v, cv);
            }
            if (fabs(1 - sum) > 0.00000001) Errorf("Test %d: values in topic (%d) over word distributions should sum to 1 but summed to %f\n", ti, i, sum);
        }
        free(components);
        shim_close(&lda);
    }
}

static void TestLDAFitTransform(void) { /* lda_test.go:91-177 */
    for (int ti = 0; ti < 2; ++ti) {
        shim lda;
        shim_new(&lda, 3);
        set_seed0(&lda);
        double *theta = shim_fit(&lda, 9, 9, ti ? dataB : dataA, 1);
        for (int j = 0; j < 9; ++j) {
            double sum = 0;
            for (int ri = 0; ri < 3; ++ri) {
                const double cv = theta[ri * 9 + j], v = docs[j][ri];
                sum += cv;
                if (fabs(cv - v) > 0.01) Errorf("Test %d: Document (%d) over topic (%d) distribution incorrect. Expected %f but received %f\n", ti, j, ri, v, cv);
            }
            if (fabs(1 - sum) > 0.00000001) Errorf("Test %d: values in document (%d) over topic distributions should sum to 1 but summed to %f\n", ti, j, sum);
        }
        free(theta);
        shim_close(&lda);
    }
}

static void TestLDATransform(void) { /* lda_test.go:179-235 */
    for (int ti = 0; ti < 2; ++ti) {
        shim lda;
        shim_new(&lda, 3);
        set_seed0(&lda);
        lda.cfg.perplexity_evaluation_frequency = 2;
        double *theta = shim_fit(&lda, 9, 9, ti ? dataB : dataA, 1);
        double *tTheta = shim_transform(&lda, 9, 9, ti ? dataB : dataA);
        for (int i = 0; i < 27; ++i)
            if (fabs(theta[i] - tTheta[i]) > 0.035) {
                Errorf("Test %d: Transformed matrix not equal to FitTransformed (element %d: %f vs %f)\n", ti, i, theta[i], tTheta[i]);
                break;
            }
        free(theta);
        free(tTheta);
        shim_close(&lda);
    }
}

/* the shim-owned generator (Rnd untouched): draws on the device, state handed back; PartialFit; Save / Load */
static void TestOwnSourcePartialFitSaveLoad(void) {
    shim lda;
    shim_new(&lda, 3);
    lda.cfg.iterations = 40;
    const lda_b200_pcg before = lda.own;
    double *theta = shim_fit(&lda, 9, 9, dataA, 1);
    lda_b200_pcg expect = before;
    lda_b200_pcg_advance(&expect, 27 + 27); /* W*K draws of init + D*K draws of nTheta */
    if (lda.own.high != expect.high || lda.own.low != expect.low) Errorf("PCG state after Fit is not the state after W*K + D*K draws\n");
    for (int j = 0; j < 9; ++j) {
        double s = theta[j] + theta[9 + j] + theta[18 + j];
        if (fabs(1 - s) > 1e-8) Errorf("document %d sums to %f\n", j, s);
    }
    free(theta);
    /* PartialFit(m): handle(), flatten, partial_fit_flat with iterations = 1 */
    {
        lda_b200_handle *h = shim_handle(&lda);
        double *m = (double *)malloc(sizeof dataB);
        memcpy(m, dataB, sizeof dataB);
        lda_b200_pcg rnd = lda.own;
        const double rho_before = lda.ctr.rho_phi_t;
        if (lda_b200_partial_fit_flat(h, LDA_B200_DENSE, 9, 9, NULL, NULL, m, NULL, &rnd, &lda.ctr, 1, NULL)) Errorf("partial_fit: %s\n", lda_b200_last_error(h));
        lda.own = rnd;
        if (lda.ctr.rho_phi_t != rho_before + 1) Errorf("PartialFit over one minibatch must advance rhoPhiT by one\n");
        free(m);
    }
    /* Save(w): save_size, save into a []byte; Load(r) on a new object: load, get_dims */
    lda_b200_handle *h = shim_handle(&lda);
    int64_t n = 0, written = 0;
    if (lda_b200_save_size(h, &n) || n != 32 + 40 + 27 * 8 + 40 + 3 * 8) Errorf("save_size: %lld\n", (long long)n);
    unsigned char *buf = (unsigned char *)malloc((size_t)n + 1) + 1; /* odd address: Go byte slices need not be 8-aligned */
    if (lda_b200_save(h, &lda.ctr, buf, n, &written) || written != n) Errorf("save: %s\n", lda_b200_last_error(h));
    shim other;
    shim_new(&other, 7);
    lda_b200_handle *h2 = shim_handle(&other);
    if (lda_b200_load(h2, buf, n, &other.ctr)) Errorf("load: %s\n", lda_b200_last_error(h2));
    int32_t k = 0;
    lda_b200_get_dims(h2, &other.w, &k);
    other.cfg.K = k;
    if (k != 3 || other.w != 9 || other.ctr.rho_phi_t != lda.ctr.rho_phi_t || other.ctr.words_in_corpus != lda.ctr.words_in_corpus)
        Errorf("Load did not restore K / W / counters\n");
    other.own = lda.own;
    double *a = shim_transform(&lda, 9, 9, dataA), *b = shim_transform(&other, 9, 9, dataA);
    if (memcmp(a, b, sizeof(double) * 27)) Errorf("Transform after Load differs from the saved model's\n");
    if (lda_b200_load(h2, buf, n - 9, &other.ctr) == 0) Errorf("a truncated model file must be refused\n");
    free(a);
    free(b);
    free(buf - 1);
    shim_close(&lda);
    shim_close(&other);
}

int main(void) {
    TestLDAFit();
    TestLDAFitTransform();
    TestLDATransform();
    TestOwnSourcePartialFitSaveLoad();
    if (failures == 0) printf("PASS\n");
    return failures;
}
