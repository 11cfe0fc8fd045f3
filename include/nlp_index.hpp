// nlp_index.hpp -- C++ host-side mirror of james-bowman/nlp's LinearScanIndex (index.go:60-135) and of the pairwise
// comparers (measures/pairwise/comparisons.go) above the C ABI of liblda_b200.so (include/lda_b200_index.h).  Same type
// and method names, argument meaning and result type (Match) as the Go code; failures throw with the library's "nlp: "
// prefix where Go would panic.  Header only; uses nlp::mat::Dense of nlp_lda.hpp for matrices.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "lda_b200_index.h"
#include "nlp_lda.hpp"

namespace nlp {

// gonum mat.Vector (only what the index touches: Len, AtVec)
namespace mat {
struct VecDense {
    std::vector<double> data;
    VecDense() = default;
    explicit VecDense(std::vector<double> d) : data(std::move(d)) {}
    int Len() const { return (int)data.size(); }
    double AtVec(int i) const { return data[(size_t)i]; }
};
// ColView: column j of a dense matrix as a vector
inline VecDense ColView(const Dense &m, int j) {
    VecDense v;
    v.data.resize((size_t)m.r);
    for (int i = 0; i < m.r; ++i) v.data[(size_t)i] = m.At(i, j);
    return v;
}
}  // namespace mat

// index.go:13-19 (ids are integers here; the Go shim maps interface{} ids to them)
struct Match {
    double Distance;
    int64_t ID;
};

class LinearScanIndex;

// measures/pairwise/comparisons.go: each comparer names the kernel specialisation and, called on two vectors, evaluates
// on the GPU (through a one-vector index)
namespace pairwise {
struct Comparer {
    int code;
    double operator()(const mat::VecDense &a, const mat::VecDense &b) const;
};
constexpr Comparer CosineSimilarity{LDA_B200_COSINE_SIMILARITY}, CosineDistance{LDA_B200_COSINE_DISTANCE},
    AngularDistance{LDA_B200_ANGULAR_DISTANCE}, AngularSimilarity{LDA_B200_ANGULAR_SIMILARITY},
    HammingDistance{LDA_B200_HAMMING_DISTANCE}, HammingSimilarity{LDA_B200_HAMMING_SIMILARITY},
    EuclideanDistance{LDA_B200_EUCLIDEAN_DISTANCE}, ManhattenDistance{LDA_B200_MANHATTEN_DISTANCE},
    VectorLenSimilarity{LDA_B200_VECTOR_LEN_SIMILARITY};
}  // namespace pairwise

// index.go:60-135
class LinearScanIndex {
   public:
    explicit LinearScanIndex(pairwise::Comparer compareFN, int device = 0) {
        if (lda_b200_index_create(device, compareFN.code, &h_)) throw std::runtime_error(std::string("nlp: ") + lda_b200_index_last_error(nullptr));
    }
    LinearScanIndex(const LinearScanIndex &) = delete;
    LinearScanIndex &operator=(const LinearScanIndex &) = delete;
    ~LinearScanIndex() { lda_b200_index_destroy(h_); }

    // index.go:79-84
    void Index(const mat::VecDense &v, int64_t id) {
        check(lda_b200_index_add(h_, v.Len(), 1, v.data.data(), 1, &id));
    }
    // ColDo(m, func(j, v) { index.Index(v, j) }) in one call
    void IndexColumns(const mat::Dense &m, const std::vector<int64_t> *ids = nullptr) {
        check(lda_b200_index_add(h_, m.r, m.c, m.data.data(), m.c, ids ? ids->data() : nullptr));
    }
    // ColDo(theta, func(j, v) { index.Index(v, ids[j]) }) with theta = lda.Transform(m), fused: the vectors go from the
    // Transform kernels into the index without leaving the GPU (lda_b200_transform_into_index); ids default to 0..D-1
    template <class M>
    void IndexTransform(LatentDirichletAllocation &lda, const M &m, const std::vector<int64_t> *ids = nullptr) {
        std::string e;
        lda_b200_handle *lh = lda.DeviceHandle(e);
        if (!lh) throw std::runtime_error(e);
        const lda_b200_matrix d = LatentDirichletAllocation::describe(m);
        if (lda_b200_transform_into_index(lh, &d, nullptr, &lda.Rnd.src.s, lda.RhoThetaT(), h_, ids ? ids->data() : nullptr))
            throw std::runtime_error(std::string("nlp: ") + lda_b200_last_error(lh));
    }
    // index.go:85-115; matches come nearest first (the reference returns them in heap order)
    std::vector<Match> Search(const mat::VecDense &qv, int k) {
        std::vector<int64_t> ids((size_t)(k > 0 ? k : 1));
        std::vector<double> dist(ids.size());
        int32_t found = 0;
        check(lda_b200_index_search(h_, 1, qv.data.data(), 1, k, ids.data(), dist.data(), &found));
        std::vector<Match> out((size_t)found);
        for (int i = 0; i < found; ++i) out[(size_t)i] = Match{dist[(size_t)i], ids[(size_t)i]};
        return out;
    }
    // index.go:117-135
    void Remove(int64_t id) { check(lda_b200_index_remove(h_, id)); }
    int64_t Len() const { return lda_b200_index_len(h_); }

   private:
    lda_b200_index *h_ = nullptr;
    void check(int rc) {
        if (rc) throw std::runtime_error(std::string("nlp: ") + lda_b200_index_last_error(h_));
    }
};

inline LinearScanIndex *NewLinearScanIndex(pairwise::Comparer compareFN) { return new LinearScanIndex(compareFN); }  // index.go:74-76

inline double pairwise::Comparer::operator()(const mat::VecDense &a, const mat::VecDense &b) const {
    if (a.Len() != b.Len()) throw std::runtime_error("nlp: vectors differ in length");
    LinearScanIndex ix(*this);
    ix.Index(b, 0);
    return ix.Search(a, 1)[0].Distance;
}

}  // namespace nlp
