// lda_test.cpp -- the reference's own LDA tests (lda_test.go:16-235), restated against the C++ host mirror
// (include/nlp_lda.hpp) so that they read like the originals: same fixtures, same seed, same tolerances.
// Built by tests/test_cpp_host.py (g++, CPU) and run on the B200 box (-m gpu).  Exit code = number of failures.
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "nlp_lda.hpp"

static int failures = 0;
#define Errorf(...)                       \
    do {                                  \
        std::fprintf(stderr, __VA_ARGS__); \
        ++failures;                       \
    } while (0)

struct Fixture {
    int topics, r, c;
    std::vector<double> data;
    std::vector<std::vector<double>> expected;
};

static const std::vector<double> dataA = {
    3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3,
    3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4};
static const std::vector<double> dataB = {
    3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 5, 1, 0, 0, 0, 0, 0, 0, 3, 5,
    0, 0, 0, 0, 0, 0, 0, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4};

// TestLDAFit, lda_test.go:16-89
static void TestLDAFit() {
    std::vector<Fixture> tests = {
        {3, 9, 9, dataA, {{0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33}, {0, 0, 0, 0.33, 0.33, 0.33, 0, 0, 0}}},
        {3, 9, 9, dataB, {{0.33, 0.33, 0.33, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0.33, 0.33, 0.33}, {0, 0, 0, 0.428, 0.285, 0.285, 0, 0, 0}}},
    };
    for (size_t ti = 0; ti < tests.size(); ++ti) {
        const Fixture &test = tests[ti];
        // set Rnd to fixed constant seed for deterministic results
        std::unique_ptr<nlp::LatentDirichletAllocation> lda(nlp::NewLatentDirichletAllocation(test.topics));
        lda->Rnd = nlp::rand::New(nlp::rand::NewSource(0));
        nlp::mat::Dense in = nlp::mat::NewDense(test.r, test.c, test.data);
        lda->Fit(in);
        nlp::mat::Dense components = lda->Components();
        for (int i = 0; i < test.topics; ++i) {
            double sum = 0;
            for (size_t ri = 0; ri < test.expected[i].size(); ++ri) {
                const double cv = components.At(i, (int)ri), v = test.expected[i][ri];
                sum += cv;
                if (std::fabs(cv - v) > 0.01)
                    Errorf("Test %zu: Topic (%d) over word (%zu) distribution incorrect. Expected %f but received %f\n", ti, i, ri, v, cv);
            }
            if (std::fabs(1 - sum) > 0.00000001)
                Errorf("Test %zu: values in topic (%d) over word distributions should sum to 1 but summed to %f\n", ti, i, sum);
        }
    }
}

// TestLDAFitTransform, lda_test.go:91-177
static void TestLDAFitTransform() {
    const std::vector<std::vector<double>> expectedDocs = {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, 1},
                                                           {0, 0, 1}, {0, 1, 0}, {0, 1, 0}, {0, 1, 0}};
    const std::vector<std::vector<double>> datas = {dataA, dataB};
    for (size_t ti = 0; ti < datas.size(); ++ti) {
        std::unique_ptr<nlp::LatentDirichletAllocation> lda(nlp::NewLatentDirichletAllocation(3));
        lda->Rnd = nlp::rand::New(nlp::rand::NewSource(0));
        nlp::mat::Dense in = nlp::mat::NewDense(9, 9, datas[ti]);
        nlp::mat::Dense theta;
        std::string err;
        if (!lda->FitTransform(in, theta, &err)) Errorf("%s\n", err.c_str());
        for (int j = 0; j < 9; ++j) {
            double sum = 0;
            for (int ri = 0; ri < 3; ++ri) {
                const double cv = theta.At(ri, j), v = expectedDocs[j][ri];
                sum += cv;
                if (std::fabs(cv - v) > 0.01)
                    Errorf("Test %zu: Document (%d) over topic (%d) distribution incorrect. Expected %f but received %f\n", ti, j, ri, v, cv);
            }
            if (std::fabs(1 - sum) > 0.00000001)
                Errorf("Test %zu: values in document (%d) over topic distributions should sum to 1 but summed to %f\n", ti, j, sum);
        }
    }
}

// TestLDATransform, lda_test.go:179-235
static void TestLDATransform() {
    const std::vector<std::vector<double>> datas = {dataA, dataB};
    for (size_t ti = 0; ti < datas.size(); ++ti) {
        std::unique_ptr<nlp::LatentDirichletAllocation> lda(nlp::NewLatentDirichletAllocation(3));
        lda->Rnd = nlp::rand::New(nlp::rand::NewSource(0));
        lda->PerplexityEvaluationFrequency = 2;
        nlp::mat::Dense in = nlp::mat::NewDense(9, 9, datas[ti]);
        nlp::mat::Dense theta, tTheta;
        std::string err;
        if (!lda->FitTransform(in, theta, &err)) Errorf("%s\n", err.c_str());
        if (!lda->Transform(in, tTheta, &err)) Errorf("%s\n", err.c_str());
        if (!nlp::mat::EqualApprox(theta, tTheta, 0.035))
            Errorf("Test %zu: Transformed matrix not equal to FitTransformed\n", ti);
    }
}

// not in the reference: the sparse branch (lda.go:464-466) -- a CSR handed over as it is, and its device-side ToCSC()
// read back to the host first -- must equal the dense branch
static void TestSparseInput() {
    nlp::mat::Dense in = nlp::mat::NewDense(9, 9, dataB);
    nlp::sparse::CSR csr;
    csr.rows = 9, csr.cols = 9;
    csr.indptr.push_back(0);
    for (int i = 0; i < 9; ++i) {
        for (int j = 0; j < 9; ++j)
            if (in.At(i, j) != 0) {
                csr.ind.push_back(j);
                csr.data.push_back(in.At(i, j));
            }
        csr.indptr.push_back((int64_t)csr.ind.size());
    }
    nlp::mat::Dense a, b, c;
    for (int pass = 0; pass < 3; ++pass) {
        std::unique_ptr<nlp::LatentDirichletAllocation> lda(nlp::NewLatentDirichletAllocation(3));
        lda->Rnd = nlp::rand::New(nlp::rand::NewSource(0));
        lda->Iterations = 10;
        std::string err;
        bool ok = pass == 0   ? lda->FitTransform(in, a, &err)
                  : pass == 1 ? lda->FitTransform(csr, b, &err)
                              : lda->FitTransform(lda->ToCSC(csr), c, &err);
        if (!ok) Errorf("%s\n", err.c_str());
    }
    if (!nlp::mat::EqualApprox(a, b, 0) || !nlp::mat::EqualApprox(a, c, 0)) Errorf("sparse and dense inputs disagree\n");
}

int main() {
    TestLDAFit();
    TestLDAFitTransform();
    TestLDATransform();
    TestSparseInput();
    std::printf("%s (%d failures)\n", failures ? "FAIL" : "PASS", failures);
    return failures;
}
