// nlp_lda.hpp -- C++ host-side mirror of james-bowman/nlp's LatentDirichletAllocation (lda.go:60-542) above the C ABI of
// liblda_b200.so (include/lda_b200.h).  The reference is Go; its toolchain is absent from the build image, so this is
// the compiled-language host side: same exported field names, constructor defaults (lda.go:160-186), method names,
// argument meaning, result orientation (K x D / K x W) and error behaviour (Transform / FitTransform report errors,
// Fit / Perplexity / Components have no error slot and throw with the library's "nlp: " prefix).  Header only.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "lda_b200.h"

namespace nlp {

// ---- gonum.org/v1/gonum/mat (only what the LDA path touches: Dims, At, NewDense; lda.go:193-194,415,452,541) -------------
namespace mat {
struct Dense {
    int r = 0, c = 0;
    std::vector<double> data;  // row major
    Dense() = default;
    Dense(int rows, int cols, std::vector<double> d) : r(rows), c(cols), data(std::move(d)) {
        if ((size_t)rows * cols != data.size()) throw std::invalid_argument("mat: bad Dense dimensions");
    }
    void Dims(int &rows, int &cols) const { rows = r, cols = c; }
    double At(int i, int j) const { return data[(size_t)i * c + j]; }
};
inline Dense NewDense(int r, int c, std::vector<double> data) { return Dense(r, c, std::move(data)); }
inline bool EqualApprox(const Dense &a, const Dense &b, double eps) {
    if (a.r != b.r || a.c != b.c) return false;
    for (size_t i = 0; i < a.data.size(); ++i)
        if (std::fabs(a.data[i] - b.data[i]) > eps) return false;
    return true;
}
}  // namespace mat

// ---- github.com/james-bowman/sparse (CSC / CSR with ToCSC, lda.go:464-466) --------------------------------------------------
namespace sparse {
struct CSC {
    int64_t rows = 0, cols = 0;
    std::vector<int64_t> indptr, ind;
    std::vector<double> data;
};
struct CSR {
    int64_t rows = 0, cols = 0;
    std::vector<int64_t> indptr, ind;
    std::vector<double> data;
};
}  // namespace sparse

// ---- golang.org/x/exp/rand (lda.go:10,126,184) ------------------------------------------------------------------------------
namespace rand {
struct PCGSource {
    lda_b200_pcg s{0, 0};
    void Seed(uint64_t seed) { lda_b200_pcg_seed(&s, seed); }
    uint64_t Uint64() { return lda_b200_pcg_uint64(&s); }
};
inline PCGSource NewSource(uint64_t seed) {
    PCGSource p;
    p.Seed(seed);
    return p;
}
struct Rand {
    PCGSource src;
    uint64_t Uint64() { return src.Uint64(); }
    int64_t Int63() { return (int64_t)(src.Uint64() & 0x7fffffffffffffffULL); }
    int64_t Int() { return Int63(); }
};
inline Rand New(PCGSource src) { return Rand{src}; }
}  // namespace rand

// lda.go:16-32
struct LearningSchedule {
    double S, Tau, Kappa;
    double Calc(double iteration) const { return S / std::pow(Tau + iteration, Kappa); }
};

class LatentDirichletAllocation {
   public:
    // exported fields, lda.go:69-134
    int Iterations = 1000;
    double PerplexityTolerance = 1e-2;
    int PerplexityEvaluationFrequency = 30;
    int BatchSize = 100;
    int K;
    int BurnInPasses = 1;
    int TransformationPasses = 500;
    double MeanChangeTolerance = 1e-5;
    int ChangeEvaluationFrequency = 30;
    double Alpha = 0.1, Eta = 0.01;
    LearningSchedule RhoPhi{10, 1000, 0.9};
    LearningSchedule RhoTheta{1, 10, 0.9};
    rand::Rand Rnd;
    // minibatches in flight per step (lda.go:134; default runtime.GOMAXPROCS(0), lda.go:185); on one GPU they share a launch
    int Processes = (int)std::max(1u, std::thread::hardware_concurrency());
    bool Deterministic = false;  // not in the reference: bitwise run-to-run reproducible fits (lda_b200_set_deterministic)
    int Device = 0;     // CUDA device ordinal (not in the reference)
    std::vector<double> PerplexityTrace;  // perplexities evaluated inside the last fit (lda.go:530-539 hides them)
    int IterationsRun = 0;

    explicit LatentDirichletAllocation(int k)
        : K(k), Rnd(rand::New(rand::NewSource((uint64_t)std::chrono::steady_clock::now().time_since_epoch().count()))) {}
    LatentDirichletAllocation(const LatentDirichletAllocation &) = delete;
    LatentDirichletAllocation &operator=(const LatentDirichletAllocation &) = delete;
    ~LatentDirichletAllocation() { Close(); }
    void Close() {
        if (h_) lda_b200_destroy(h_);
        h_ = nullptr;
    }

    // lda.go:211-214
    template <class M>
    LatentDirichletAllocation &Fit(const M &m) {
        std::string err;
        fit(describe(m), false, nullptr, err);
        if (!err.empty()) throw std::runtime_error(err);
        return *this;
    }
    // lda.go:463-542: K x D; returns false and sets *err on failure (Go: (mat.Matrix, error))
    template <class M>
    bool FitTransform(const M &m, mat::Dense &theta, std::string *err = nullptr) {
        std::string e;
        fit(describe(m), true, &theta, e);
        if (err) *err = e;
        return e.empty();
    }
    // lda.go:446-453
    template <class M>
    bool Transform(const M &m, mat::Dense &theta, std::string *err = nullptr) {
        const lda_b200_matrix c = describe(m);
        std::string e;
        if (handle(e)) {
            std::vector<double> out((size_t)K * c.cols);
            if (lda_b200_transform_matrix(h_, &c, nullptr, &Rnd.src.s, rhoThetaT_, out.data(), nullptr))
                e = std::string("nlp: ") + lda_b200_last_error(h_);
            else
                theta = mat::NewDense(K, (int)c.cols, std::move(out));
        }
        if (err) *err = e;
        return e.empty();
    }
    // lda.go:368-389
    template <class M>
    double Perplexity(const M &m) {
        const lda_b200_matrix c = describe(m);
        std::string e;
        double out = 0;
        if (handle(e) && lda_b200_perplexity_matrix(h_, &c, nullptr, &Rnd.src.s, rhoThetaT_, &out))
            e = std::string("nlp: ") + lda_b200_last_error(h_);
        if (!e.empty()) throw std::runtime_error(e);
        return out;
    }
    // for LinearScanIndex::IndexTransform (nlp_index.hpp): the device-side handle (created on first use) and the private
    // counter that Transform hands to burnInDoc (lda.go:231)
    lda_b200_handle *DeviceHandle(std::string &err) { return handle(err) ? h_ : nullptr; }
    double RhoThetaT() const { return rhoThetaT_; }

    // lda.go:414-416: K x W
    mat::Dense Components() {
        std::string e;
        if (!handle(e)) throw std::runtime_error(e);
        int64_t W = 0;
        lda_b200_get_dims(h_, &W, nullptr);
        std::vector<double> out((size_t)K * W);
        if (lda_b200_components(h_, out.data())) throw std::runtime_error(std::string("nlp: ") + lda_b200_last_error(h_));
        return mat::NewDense(K, (int)W, std::move(out));
    }

    // The matrix in the storage it already has (lda_b200_matrix): sparse.ToCSC() (lda.go:464-466) and the dense scan of
    // ColNonZeroElemDo (utils.go:51-57: rows ascending, exact zeros skipped) run on the device.
    static lda_b200_matrix describe(const mat::Dense &m) {
        return lda_b200_matrix{LDA_B200_DENSE, m.r, m.c, nullptr, nullptr, ptr(m.data)};
    }
    static lda_b200_matrix describe(const sparse::CSC &m) {
        return lda_b200_matrix{LDA_B200_CSC, m.rows, m.cols, ptr(m.indptr), ptr(m.ind), ptr(m.data)};
    }
    static lda_b200_matrix describe(const sparse::CSR &m) {
        return lda_b200_matrix{LDA_B200_CSR, m.rows, m.cols, ptr(m.indptr), ptr(m.ind), ptr(m.data)};
    }
    // sparse.CSR.ToCSC() (lda.go:464-466) on the device, result back on the host
    sparse::CSC ToCSC(const sparse::CSR &m) {
        std::string e;
        if (!handle(e)) throw std::runtime_error(e);
        sparse::CSC c;
        c.rows = m.rows, c.cols = m.cols;
        c.indptr.resize((size_t)m.cols + 1);
        c.ind.resize(m.ind.size());
        c.data.resize(m.data.size());
        if (lda_b200_csr_to_csc(h_, m.rows, m.cols, m.indptr.data(), ptr(m.ind), ptr(m.data), c.indptr.data(),
                                c.ind.empty() ? nullptr : c.ind.data(), c.data.empty() ? nullptr : c.data.data()))
            throw std::runtime_error(std::string("nlp: ") + lda_b200_last_error(h_));
        return c;
    }

   private:
    lda_b200_handle *h_ = nullptr;
    double rhoPhiT_ = 1, rhoThetaT_ = 1, wordsInCorpus_ = 0;  // lda.go:117-120, 182-183

    template <class T>
    static const T *ptr(const std::vector<T> &v) {
        return v.empty() ? nullptr : v.data();
    }
    bool handle(std::string &err) {
        lda_b200_config c;
        lda_b200_default_config(K, &c);
        c.iterations = Iterations;
        c.batch_size = BatchSize;
        c.burn_in_passes = BurnInPasses;
        c.transformation_passes = TransformationPasses;
        c.perplexity_evaluation_frequency = PerplexityEvaluationFrequency;
        c.change_evaluation_frequency = ChangeEvaluationFrequency;
        c.processes = Processes;
        c.perplexity_tolerance = PerplexityTolerance;
        c.mean_change_tolerance = MeanChangeTolerance;
        c.alpha = Alpha;
        c.eta = Eta;
        c.rho_phi = {RhoPhi.S, RhoPhi.Tau, RhoPhi.Kappa};
        c.rho_theta = {RhoTheta.S, RhoTheta.Tau, RhoTheta.Kappa};
        if (!h_) {
            if (lda_b200_create(&c, Device, &h_)) {
                err = std::string("nlp: ") + lda_b200_last_error(nullptr);
                return false;
            }
        } else if (lda_b200_set_config(h_, &c)) {
            err = std::string("nlp: ") + lda_b200_last_error(h_);
            return false;
        }
        lda_b200_set_deterministic(h_, Deterministic ? 1 : 0);
        return true;
    }
    void fit(const lda_b200_matrix &c, bool want, mat::Dense *theta, std::string &err) {
        if (!handle(err)) return;
        std::vector<double> out(want ? (size_t)K * c.cols : 0);
        std::vector<double> trace(PerplexityEvaluationFrequency > 0 ? (size_t)Iterations / PerplexityEvaluationFrequency + 1 : 1);
        lda_b200_counters ctr{rhoPhiT_, rhoThetaT_, wordsInCorpus_};
        int32_t n_evals = 0, iters = 0;
        if (lda_b200_fit_matrix(h_, &c, nullptr, nullptr, &Rnd.src.s, &ctr, want ? out.data() : nullptr, trace.data(),
                                (int32_t)trace.size(), &n_evals, &iters)) {
            err = std::string("nlp: ") + lda_b200_last_error(h_);
            return;
        }
        rhoPhiT_ = ctr.rho_phi_t, rhoThetaT_ = ctr.rho_theta_t, wordsInCorpus_ = ctr.words_in_corpus;
        PerplexityTrace.assign(trace.begin(), trace.begin() + n_evals);
        IterationsRun = iters;
        if (want) *theta = mat::NewDense(K, (int)c.cols, std::move(out));
    }
};

// lda.go:145
inline LatentDirichletAllocation *NewLatentDirichletAllocation(int k) { return new LatentDirichletAllocation(k); }

}  // namespace nlp
