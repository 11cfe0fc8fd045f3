// +build cgo

// Package nlp: drop-in replacement of lda.go's LatentDirichletAllocation backed by liblda_b200.so (B200, sm_100a).
//
// NOT COMPILED OR RUN IN THIS REPOSITORY'S BUILD ENVIRONMENT: there is no Go toolchain in the image (DESIGN.md 1).
// This is the binding a maintainer of james-bowman/nlp would add next to lda.go (replacing it under a build tag); the
// same C entry points are exercised through ctypes by nlp_b200/lda.py, which mirrors this file one to one.
//
// Exported fields, defaults, method names, argument meaning, result orientation and error behaviour are those of the
// reference type (lda.go:68-189, 211, 368, 414, 446, 463): Fit returns the receiver, Transform / FitTransform return
// (mat.Matrix, error) with a K x D dense result, Perplexity / Components have no error slot and panic with an
// "nlp: " prefix like the rest of the library does on failure (e.g. dimreduction.go:41-43).
package nlp

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../nlp_b200 -llda_b200 -Wl,-rpath,${SRCDIR}/../../nlp_b200
#include <stdlib.h>
#include "lda_b200.h"
*/
import "C"

import (
	"encoding/binary"
	"errors"
	"io"
	"runtime"
	"time"
	"unsafe"

	"github.com/james-bowman/sparse"
	"golang.org/x/exp/rand"
	"gonum.org/v1/gonum/mat"
)

// LearningSchedule is lda.go:16-32 unchanged.
type LearningSchedule struct {
	S, Tau, Kappa float64
}

// LatentDirichletAllocation keeps the exported surface of lda.go:68-134.
type LatentDirichletAllocation struct {
	Iterations                    int
	PerplexityTolerance           float64
	PerplexityEvaluationFrequency int
	BatchSize                     int
	K                             int
	BurnInPasses                  int
	TransformationPasses          int
	MeanChangeTolerance           float64
	ChangeEvaluationFrequency     int
	Alpha, Eta                    float64
	RhoPhi, RhoTheta              LearningSchedule
	Rnd                           *rand.Rand
	// Processes is the number of minibatches in flight per step (lda.go:134); on one GPU they share a launch.
	Processes int
	// Device is the CUDA device ordinal (not in the reference).
	Device int
	// Deterministic makes fits bitwise reproducible from run to run (not in the reference; lda_b200_set_deterministic).
	Deterministic bool

	rhoPhiT, rhoThetaT, wordsInCorpus float64 // lda.go:117-120
	w, d                              int
	// own is the generator the constructor installed and src its PCG source.  While Rnd == own the initial nPhi / nTheta
	// are drawn on the device from src's 128-bit state (bit-identical to Rnd.Int() % n, state advanced accordingly).  A
	// caller who assigns Rnd (lda_test.go:68: lda.Rnd = rand.New(rand.NewSource(0))) gets the draws made in Go from that
	// generator, in the reference's order, and handed to the library as explicit initial values.
	own *rand.Rand
	src *rand.PCGSource
	h   *C.lda_b200_handle
}

// NewLatentDirichletAllocation mirrors lda.go:145-189.
func NewLatentDirichletAllocation(k int) *LatentDirichletAllocation {
	src := &rand.PCGSource{}
	src.Seed(uint64(time.Now().UnixNano()))
	own := rand.New(src)
	l := &LatentDirichletAllocation{
		Iterations: 1000, PerplexityTolerance: 1e-2, PerplexityEvaluationFrequency: 30, BatchSize: 100, K: k,
		BurnInPasses: 1, TransformationPasses: 500, MeanChangeTolerance: 1e-5, ChangeEvaluationFrequency: 30,
		Alpha: 0.1, Eta: 0.01,
		RhoPhi:   LearningSchedule{S: 10, Tau: 1000, Kappa: 0.9},
		RhoTheta: LearningSchedule{S: 1, Tau: 10, Kappa: 0.9},
		rhoPhiT:  1, rhoThetaT: 1,
		Rnd: own, own: own, src: src,
		Processes: runtime.GOMAXPROCS(0), // lda.go:185
	}
	runtime.SetFinalizer(l, (*LatentDirichletAllocation).Close)
	return l
}

func (l *LatentDirichletAllocation) config() C.lda_b200_config {
	return C.lda_b200_config{
		K: C.int32_t(l.K), iterations: C.int32_t(l.Iterations), batch_size: C.int32_t(l.BatchSize),
		burn_in_passes: C.int32_t(l.BurnInPasses), transformation_passes: C.int32_t(l.TransformationPasses),
		perplexity_evaluation_frequency: C.int32_t(l.PerplexityEvaluationFrequency),
		change_evaluation_frequency:     C.int32_t(l.ChangeEvaluationFrequency),
		processes:                       C.int32_t(l.Processes),
		perplexity_tolerance:            C.double(l.PerplexityTolerance),
		mean_change_tolerance:           C.double(l.MeanChangeTolerance),
		alpha:                           C.double(l.Alpha), eta: C.double(l.Eta),
		rho_phi:   C.lda_b200_schedule{S: C.double(l.RhoPhi.S), Tau: C.double(l.RhoPhi.Tau), Kappa: C.double(l.RhoPhi.Kappa)},
		rho_theta: C.lda_b200_schedule{S: C.double(l.RhoTheta.S), Tau: C.double(l.RhoTheta.Tau), Kappa: C.double(l.RhoTheta.Kappa)},
	}
}

func (l *LatentDirichletAllocation) handle() (*C.lda_b200_handle, error) {
	cfg := l.config() // a C struct of scalars: no Go pointers inside
	if l.h == nil {
		if rc := C.lda_b200_create(&cfg, C.int32_t(l.Device), &l.h); rc != 0 {
			return nil, errors.New("nlp: " + C.GoString(C.lda_b200_last_error(nil)))
		}
	} else if rc := C.lda_b200_set_config(l.h, &cfg); rc != 0 {
		return nil, l.err()
	}
	det := C.int32_t(0)
	if l.Deterministic {
		det = 1
	}
	C.lda_b200_set_deterministic(l.h, det)
	return l.h, nil
}

func (l *LatentDirichletAllocation) err() error {
	return errors.New("nlp: " + C.GoString(C.lda_b200_last_error(l.h)))
}

// Close releases the device-side model (the reference relies on the GC; device memory needs an explicit owner).
func (l *LatentDirichletAllocation) Close() {
	if l.h != nil {
		C.lda_b200_destroy(l.h)
		l.h = nil
	}
}

// flatMatrix is m in the storage it already has, as the scalar arguments of the lda_b200_*_flat calls, so that ToCSC()
// (lda.go:464-466) and the dense scan of ColNonZeroElemDo (utils.go:51-57) run on the GPU instead of on one host thread:
//   *sparse.CSR (what TfidfTransformer emits, weightings.go:79-95) -> LDA_B200_CSR, zero copy
//   *sparse.CSC                                                      -> LDA_B200_CSC, zero copy
//   *mat.Dense with Stride == cols                                  -> LDA_B200_DENSE, zero copy
//   any other sparse.TypeConverter (DOK, COO, ...)                   -> ToCSC() on the host, then LDA_B200_CSC
//   any other mat.Matrix                                             -> copied into a dense buffer, LDA_B200_DENSE
// cgo's pointer rule: Go memory handed to C must not itself contain Go pointers.  A Go-side lda_b200_matrix filled with
// slice addresses would break it, which is why the slices travel as separate arguments: each is a pointer to a flat
// []int / []float64.  The library copies to the device during the call and keeps nothing.
type flatMatrix struct {
	format     C.int32_t
	rows, cols C.int64_t
	ptr, ind   []int // Go int == int64 on the 64-bit platforms the library supports
	data       []float64
}

func flatten(m mat.Matrix) flatMatrix {
	r, c := m.Dims()
	f := flatMatrix{rows: C.int64_t(r), cols: C.int64_t(c)}
	switch t := m.(type) {
	case *sparse.CSR:
		raw := t.RawMatrix()
		f.format, f.ptr, f.ind, f.data = C.LDA_B200_CSR, raw.Indptr, raw.Ind, raw.Data
	case *sparse.CSC:
		raw := t.RawMatrix()
		f.format, f.ptr, f.ind, f.data = C.LDA_B200_CSC, raw.Indptr, raw.Ind, raw.Data
	case sparse.TypeConverter:
		raw := t.ToCSC().RawMatrix()
		f.format, f.ptr, f.ind, f.data = C.LDA_B200_CSC, raw.Indptr, raw.Ind, raw.Data
	default:
		f.format = C.LDA_B200_DENSE
		if d, ok := m.(*mat.Dense); ok && d.RawMatrix().Stride == c {
			f.data = d.RawMatrix().Data
		} else {
			f.data = make([]float64, r*c)
			for i := 0; i < r; i++ {
				for j := 0; j < c; j++ {
					f.data[i*c+j] = m.At(i, j)
				}
			}
		}
	}
	return f
}

func ptrI(s []int) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}
func ptrF(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

// pcg marshals the PCGSource state (MarshalBinary: high then low, big endian) for the device-side draws.
func (l *LatentDirichletAllocation) pcg() C.lda_b200_pcg {
	b, _ := l.src.MarshalBinary()
	return C.lda_b200_pcg{high: C.uint64_t(binary.BigEndian.Uint64(b[:8])), low: C.uint64_t(binary.BigEndian.Uint64(b[8:]))}
}
func (l *LatentDirichletAllocation) takePcg(p C.lda_b200_pcg) {
	var b [16]byte
	binary.BigEndian.PutUint64(b[:8], uint64(p.high))
	binary.BigEndian.PutUint64(b[8:], uint64(p.low))
	_ = l.src.UnmarshalBinary(b[:])
}

// draws returns rows*K initial values float64(Rnd.Int() % (rows*K)) / float64(rows*K), row-major -- the loops of
// lda.go:199-205 (nPhi), :472-475 (nTheta) and :422-426 (Transform) -- when the caller has replaced Rnd, nil otherwise
// (the library then draws on the device from the shim-owned source).
func (l *LatentDirichletAllocation) draws(rows int) []float64 {
	if l.Rnd == l.own {
		return nil
	}
	n := rows * l.K
	out := make([]float64, n)
	for i := range out {
		out[i] = float64(l.Rnd.Int()%n) / float64(n)
	}
	return out
}

func (l *LatentDirichletAllocation) counters() C.lda_b200_counters {
	return C.lda_b200_counters{rho_phi_t: C.double(l.rhoPhiT), rho_theta_t: C.double(l.rhoThetaT), words_in_corpus: C.double(l.wordsInCorpus)}
}
func (l *LatentDirichletAllocation) takeCounters(c C.lda_b200_counters) {
	l.rhoPhiT, l.rhoThetaT, l.wordsInCorpus = float64(c.rho_phi_t), float64(c.rho_theta_t), float64(c.words_in_corpus)
}

// Fit is lda.go:211-214.
func (l *LatentDirichletAllocation) Fit(m mat.Matrix) Transformer {
	if _, err := l.fit(m, false); err != nil {
		panic(err)
	}
	return l
}

// FitTransform is lda.go:463-542: K x D document-over-topic matrix.
func (l *LatentDirichletAllocation) FitTransform(m mat.Matrix) (mat.Matrix, error) { return l.fit(m, true) }

func (l *LatentDirichletAllocation) fit(m mat.Matrix, want bool) (mat.Matrix, error) {
	h, err := l.handle()
	if err != nil {
		return nil, err
	}
	f := flatten(m)
	r, c := m.Dims()
	l.w, l.d = r, c
	var theta []float64
	if want {
		theta = make([]float64, l.K*c)
	}
	nphi0 := l.draws(r)   // lda.go:199-205 (init), before ...
	ntheta0 := l.draws(c) // ... lda.go:472-475
	ctr := l.counters()
	rnd := l.pcg()
	rc := C.lda_b200_fit_flat(h, f.format, f.rows, f.cols, ptrI(f.ptr), ptrI(f.ind), ptrF(f.data), ptrF(nphi0), ptrF(ntheta0),
		&rnd, &ctr, ptrF(theta), nil, 0, nil, nil)
	runtime.KeepAlive(f)
	runtime.KeepAlive(nphi0)
	runtime.KeepAlive(ntheta0)
	if rc != 0 {
		return nil, l.err()
	}
	l.takePcg(rnd) // unchanged by the library when explicit initial values were given
	l.takeCounters(ctr)
	if !want {
		return nil, nil
	}
	return mat.NewDense(l.K, c, theta), nil
}

// PartialFit is OnlineTransformer.PartialFit (vectorisers.go:36-39; a TODO for LDA at lda.go:147): one SCVB0 iteration
// over the documents of m, continuing from the current model (a fresh random model on first use).
func (l *LatentDirichletAllocation) PartialFit(m mat.Matrix) OnlineTransformer {
	h, err := l.handle()
	if err != nil {
		panic(err)
	}
	f := flatten(m)
	r, c := m.Dims()
	if l.Rnd != l.own && l.w == 0 {
		// a first PartialFit also initialises nPhi (lda.go:199-205); with a caller-owned Rnd the draws are made here
		// and installed as the model state before the streaming step
		nphi0 := l.draws(r)
		nz := make([]float64, l.K)
		for i := 0; i < r; i++ {
			for k := 0; k < l.K; k++ {
				nz[k] += nphi0[i*l.K+k] // lda.go:203
			}
		}
		if rc := C.lda_b200_set_state(h, C.int64_t(r), ptrF(nphi0), ptrF(nz)); rc != 0 {
			panic(l.err())
		}
	}
	l.w, l.d = r, c
	ntheta0 := l.draws(c)
	ctr := l.counters()
	rnd := l.pcg()
	rc := C.lda_b200_partial_fit_flat(h, f.format, f.rows, f.cols, ptrI(f.ptr), ptrI(f.ind), ptrF(f.data), ptrF(ntheta0), &rnd, &ctr, 1, nil)
	runtime.KeepAlive(f)
	runtime.KeepAlive(ntheta0)
	if rc != 0 {
		panic(l.err())
	}
	l.takePcg(rnd)
	l.takeCounters(ctr)
	return l
}

// Transform is lda.go:446-453.
func (l *LatentDirichletAllocation) Transform(m mat.Matrix) (mat.Matrix, error) {
	h, err := l.handle()
	if err != nil {
		return nil, err
	}
	f := flatten(m)
	_, c := m.Dims()
	theta := make([]float64, l.K*c)
	theta0 := l.draws(c) // lda.go:422-426
	rnd := l.pcg()
	rc := C.lda_b200_transform_flat(h, f.format, f.rows, f.cols, ptrI(f.ptr), ptrI(f.ind), ptrF(f.data), ptrF(theta0), &rnd,
		C.double(l.rhoThetaT), ptrF(theta), nil)
	runtime.KeepAlive(f)
	runtime.KeepAlive(theta0)
	if rc != 0 {
		return nil, l.err()
	}
	l.takePcg(rnd)
	return mat.NewDense(l.K, c, theta), nil
}

// Perplexity is lda.go:368-389.
func (l *LatentDirichletAllocation) Perplexity(m mat.Matrix) float64 {
	h, err := l.handle()
	if err != nil {
		panic(err)
	}
	f := flatten(m)
	_, c := m.Dims()
	theta0 := l.draws(c)
	rnd := l.pcg()
	var out C.double
	rc := C.lda_b200_perplexity_flat(h, f.format, f.rows, f.cols, ptrI(f.ptr), ptrI(f.ind), ptrF(f.data), ptrF(theta0), &rnd,
		C.double(l.rhoThetaT), &out)
	runtime.KeepAlive(f)
	runtime.KeepAlive(theta0)
	if rc != 0 {
		panic(l.err())
	}
	l.takePcg(rnd)
	return float64(out)
}

// Components is lda.go:414-416: K x W topic-over-word matrix.
func (l *LatentDirichletAllocation) Components() mat.Matrix {
	h, err := l.handle()
	if err != nil {
		panic(err)
	}
	out := make([]float64, l.K*l.w)
	if rc := C.lda_b200_components(h, ptrF(out)); rc != 0 {
		panic(l.err())
	}
	return mat.NewDense(l.K, l.w, out)
}

// Save binary serialises the model and writes it into w, in the style of TfidfTransformer.Save (weightings.go:97-101)
// and TruncatedSVD.Save (dimreduction.go:111-121): K, the private counters, nPhi and nZ as gonum MarshalBinaryTo
// records.  The bytes are produced by the library (lda_b200_save), so every host language writes the same file.
func (l *LatentDirichletAllocation) Save(w io.Writer) error {
	h, err := l.handle()
	if err != nil {
		return err
	}
	var n C.int64_t
	if rc := C.lda_b200_save_size(h, &n); rc != 0 {
		return errors.New("nlp: no model to save")
	}
	buf := make([]byte, int(n))
	ctr := l.counters()
	if rc := C.lda_b200_save(h, &ctr, unsafe.Pointer(&buf[0]), n, &n); rc != 0 {
		return l.err()
	}
	_, err = w.Write(buf[:int(n)])
	return err
}

// Load binary deserialises a model written by Save into the receiver (dimreduction.go:127-153).  Load should only be
// performed with trusted data.
func (l *LatentDirichletAllocation) Load(r io.Reader) error {
	buf, err := io.ReadAll(r)
	if err != nil {
		return err
	}
	if len(buf) < 8 {
		return io.ErrUnexpectedEOF
	}
	h, err := l.handle()
	if err != nil {
		return err
	}
	var ctr C.lda_b200_counters
	if rc := C.lda_b200_load(h, unsafe.Pointer(&buf[0]), C.int64_t(len(buf)), &ctr); rc != 0 {
		return l.err()
	}
	var w C.int64_t
	var k C.int32_t
	C.lda_b200_get_dims(h, &w, &k)
	l.K, l.w = int(k), int(w)
	l.takeCounters(ctr)
	return nil
}
