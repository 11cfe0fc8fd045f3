// index_test.cpp -- the reference's index tests (index_test.go:13-135, the LinearScanIndex rows of each table), restated
// against the C++ host mirror (include/nlp_index.hpp) so that they read like the originals.  Built by
// tests/test_cpp_host.py (g++, CPU) and run on the B200 box (-m gpu).  Exit code = number of failures.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <numeric>
#include <random>
#include <vector>

#include "nlp_index.hpp"

static int failures = 0;
#define Errorf(...)                        \
    do {                                   \
        std::fprintf(stderr, __VA_ARGS__); \
        ++failures;                        \
    } while (0)

// sparse.Random(sparse.DenseFormat, r, c, 1.0): a fully dense random matrix
static nlp::mat::Dense Random(int r, int c, unsigned seed) {
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::vector<double> d((size_t)r * c);
    for (double &x : d) x = u(gen);
    return nlp::mat::NewDense(r, c, std::move(d));
}

// TestIndexerIndex, index_test.go:13-44
static void TestIndexerIndex() {
    nlp::mat::Dense m = Random(100, 10, 1);
    std::unique_ptr<nlp::LinearScanIndex> index(nlp::NewLinearScanIndex(nlp::pairwise::CosineDistance));
    for (int j = 0; j < m.c; ++j) index->Index(nlp::mat::ColView(m, j), j);
    for (int j = 0; j < m.c; ++j) {
        std::vector<nlp::Match> matches = index->Search(nlp::mat::ColView(m, j), 1);
        if (matches.size() != 1) {
            Errorf("Search expected 1 result but received %zu\n", matches.size());
            continue;
        }
        if (matches[0].ID != j) Errorf("Search expected to find %d but found %lld\n", j, (long long)matches[0].ID);
        if (matches[0].Distance < -0.0000001 || matches[0].Distance > 0.0000001)
            Errorf("Search match distance expected 0.0 but received %f\n", matches[0].Distance);
    }
}

// TestIndexerSearch, index_test.go:46-96
static void TestIndexerSearch() {
    const int numCols = 10;
    nlp::mat::Dense m = Random(100, numCols, 2);
    // build similarity matrix
    std::vector<std::vector<int>> inds((size_t)numCols);
    for (int j = 0; j < numCols; ++j) {
        std::vector<double> row((size_t)numCols);
        for (int i = 0; i < numCols; ++i) row[(size_t)i] = nlp::pairwise::CosineDistance(nlp::mat::ColView(m, j), nlp::mat::ColView(m, i));
        inds[(size_t)j].resize((size_t)numCols);
        std::iota(inds[(size_t)j].begin(), inds[(size_t)j].end(), 0);
        std::stable_sort(inds[(size_t)j].begin(), inds[(size_t)j].end(), [&](int a, int b) { return row[(size_t)a] < row[(size_t)b]; });
    }
    std::unique_ptr<nlp::LinearScanIndex> index(nlp::NewLinearScanIndex(nlp::pairwise::CosineDistance));
    for (int j = 0; j < numCols; ++j) index->Index(nlp::mat::ColView(m, j), j);
    for (int j = 0; j < numCols; ++j) {
        std::vector<nlp::Match> matches = index->Search(nlp::mat::ColView(m, j), numCols);
        if ((int)matches.size() != numCols) Errorf("Search expected %d result but received %zu\n", numCols, matches.size());
        // the reference sorts the heap's matches by distance before comparing (index_test.go:83-84); they arrive sorted here
        for (size_t i = 0; i < matches.size(); ++i)
            if (matches[i].ID != inds[(size_t)j][i]) {
                Errorf("For col #%d, Rank #%zu - expected %d but found %lld\n", j, i, inds[(size_t)j][i], (long long)matches[i].ID);
                break;
            }
    }
}

// TestIndexerRemove, index_test.go:98-135
static void TestIndexerRemove() {
    nlp::mat::Dense m = Random(100, 10, 3);
    std::unique_ptr<nlp::LinearScanIndex> index(nlp::NewLinearScanIndex(nlp::pairwise::CosineDistance));
    for (int j = 0; j < m.c; ++j) index->Index(nlp::mat::ColView(m, j), j);
    for (int j = 0; j < m.c; ++j) {
        index->Remove(j);
        std::vector<nlp::Match> matches = index->Search(nlp::mat::ColView(m, j), 1);
        if (matches.size() > 1) Errorf("Search expected less than 1 result but received %zu\n", matches.size());
        if (matches.size() == 1 && matches[0].ID == j) Errorf("Search expected not to find %d but found %lld\n", j, (long long)matches[0].ID);
    }
}

// not in the reference: ColDo(theta, index.Index) as one call, and closed-form comparer values (comparisons.go)
static void TestColumnsAndComparers() {
    nlp::mat::Dense m = Random(16, 300, 4);
    std::unique_ptr<nlp::LinearScanIndex> index(nlp::NewLinearScanIndex(nlp::pairwise::EuclideanDistance));
    index->IndexColumns(m);
    if (index->Len() != 300) Errorf("Len expected 300 but is %lld\n", (long long)index->Len());
    for (int j : {0, 150, 299}) {
        std::vector<nlp::Match> matches = index->Search(nlp::mat::ColView(m, j), 3);
        if (matches.size() != 3 || matches[0].ID != j || matches[0].Distance != 0) Errorf("column %d is not its own nearest neighbour\n", j);
        if (matches.size() == 3 && !(matches[1].Distance <= matches[2].Distance)) Errorf("matches are not sorted nearest first\n");
    }
    const nlp::mat::VecDense a({1, 2, 2}), b({2, 0, 1}), z({0, 0, 0});
    const double cos = 4.0 / (3.0 * std::sqrt(5.0));
    if (std::fabs(nlp::pairwise::CosineSimilarity(a, b) - cos) > 1e-12) Errorf("CosineSimilarity\n");
    if (std::fabs(nlp::pairwise::AngularDistance(a, b) - std::acos(cos) / M_PI) > 1e-12) Errorf("AngularDistance\n");
    if (std::fabs(nlp::pairwise::EuclideanDistance(a, b) - std::sqrt(6.0)) > 1e-12) Errorf("EuclideanDistance\n");
    if (nlp::pairwise::ManhattenDistance(a, b) != 4.0) Errorf("ManhattenDistance\n");
    if (nlp::pairwise::HammingDistance(a, b) != 1.0) Errorf("HammingDistance\n");
    if (nlp::pairwise::VectorLenSimilarity(a, b) != 2.0) Errorf("VectorLenSimilarity\n");
    if (!std::isnan(nlp::pairwise::CosineDistance(a, z))) Errorf("CosineDistance of a zero vector should be NaN\n");
}

// not in the reference: LDA's Transform feeding the index on the device (lda_b200_transform_into_index)
static void TestIndexTransform() {
    const std::vector<double> dataA = {
        3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 3, 3,
        3, 0, 0, 0, 0, 0, 0, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4};
    std::unique_ptr<nlp::LatentDirichletAllocation> lda(nlp::NewLatentDirichletAllocation(3));
    lda->Rnd = nlp::rand::New(nlp::rand::NewSource(0));
    nlp::mat::Dense in = nlp::mat::NewDense(9, 9, dataA);
    nlp::mat::Dense theta;
    std::string err;
    if (!lda->FitTransform(in, theta, &err)) Errorf("%s\n", err.c_str());
    std::unique_ptr<nlp::LinearScanIndex> index(nlp::NewLinearScanIndex(nlp::pairwise::CosineDistance));
    index->IndexTransform(*lda, in);
    if (index->Len() != 9) Errorf("Len expected 9 but is %lld\n", (long long)index->Len());
    // documents 0-2, 3-5 and 6-8 use the same words (lda_test.go:26-36): a document's nearest neighbours are its own group
    for (int j = 0; j < 9; ++j) {
        std::vector<nlp::Match> matches = index->Search(nlp::mat::ColView(theta, j), 3);
        if (matches.size() != 3) Errorf("Search expected 3 results but received %zu\n", matches.size());
        for (const nlp::Match &mt : matches)
            if (mt.ID / 3 != j / 3 || mt.Distance > 0.01) Errorf("document %d: neighbour %lld at distance %f\n", j, (long long)mt.ID, mt.Distance);
    }
}

int main() {
    try {
        TestIndexerIndex();
        TestIndexerSearch();
        TestIndexerRemove();
        TestColumnsAndComparers();
        TestIndexTransform();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 100;
    }
    std::printf("%s (%d failures)\n", failures ? "FAIL" : "PASS", failures);
    return failures;
}
