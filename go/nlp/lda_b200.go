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
	src                               *rand.PCGSource
	h                                 *C.lda_b200_handle
}

// NewLatentDirichletAllocation mirrors lda.go:145-189.
func NewLatentDirichletAllocation(k int) *LatentDirichletAllocation {
	src := &rand.PCGSource{}
	src.Seed(uint64(time.Now().UnixNano()))
	l := &LatentDirichletAllocation{
		Iterations: 1000, PerplexityTolerance: 1e-2, PerplexityEvaluationFrequency: 30, BatchSize: 100, K: k,
		BurnInPasses: 1, TransformationPasses: 500, MeanChangeTolerance: 1e-5, ChangeEvaluationFrequency: 30,
		Alpha: 0.1, Eta: 0.01,
		RhoPhi:   LearningSchedule{S: 10, Tau: 1000, Kappa: 0.9},
		RhoTheta: LearningSchedule{S: 1, Tau: 10, Kappa: 0.9},
		rhoPhiT:  1, rhoThetaT: 1,
		Rnd: rand.New(src), src: src,
		Processes: runtime.GOMAXPROCS(0), // lda.go:185
	}
	runtime.SetFinalizer(l, (*LatentDirichletAllocation).Close)
	return l
}

// SetSource replaces Rnd (lda_test.go:68 does `lda.Rnd = rand.New(rand.NewSource(0))`; the shim needs the PCGSource
// itself to hand its 128-bit state to the device and to advance it afterwards).
func (l *LatentDirichletAllocation) SetSource(src *rand.PCGSource) { l.src, l.Rnd = src, rand.New(src) }

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
	runtime.LockOSThread() // cgo calls may hop OS threads; every entry point also sets its own device
	defer runtime.UnlockOSThread()
	cfg := l.config()
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

// matrix describes m to the library in the storage it already has, so that ToCSC() (lda.go:464-466) and the dense
// scan of ColNonZeroElemDo (utils.go:51-57) run on the GPU instead of on one host thread:
//   *sparse.CSR (what TfidfTransformer emits, weightings.go:79-95) -> LDA_B200_CSR, zero copy
//   *sparse.CSC                                                      -> LDA_B200_CSC, zero copy
//   *mat.Dense with Stride == cols                                  -> LDA_B200_DENSE, zero copy
//   any other sparse.TypeConverter (DOK, COO, ...)                   -> ToCSC() on the host, then LDA_B200_CSC
//   any other mat.Matrix                                             -> copied into a dense buffer, LDA_B200_DENSE
// keep holds the Go slices the descriptor points into; the caller passes it to runtime.KeepAlive after the C call.
func matrix(m mat.Matrix) (desc C.lda_b200_matrix, keep []interface{}) {
	r, c := m.Dims()
	desc.rows, desc.cols = C.int64_t(r), C.int64_t(c)
	sparseDesc := func(format C.int32_t, indptr, ind []int, data []float64) {
		desc.format = format
		desc.ptr, desc.ind, desc.data = ptrI(indptr), ptrI(ind), ptrF(data) // Go int == int64 on amd64
		keep = []interface{}{indptr, ind, data}
	}
	switch t := m.(type) {
	case *sparse.CSR:
		raw := t.RawMatrix()
		sparseDesc(C.LDA_B200_CSR, raw.Indptr, raw.Ind, raw.Data)
	case *sparse.CSC:
		raw := t.RawMatrix()
		sparseDesc(C.LDA_B200_CSC, raw.Indptr, raw.Ind, raw.Data)
	case sparse.TypeConverter:
		raw := t.ToCSC().RawMatrix()
		sparseDesc(C.LDA_B200_CSC, raw.Indptr, raw.Ind, raw.Data)
	default:
		var data []float64
		if d, ok := m.(*mat.Dense); ok && d.RawMatrix().Stride == c {
			data = d.RawMatrix().Data
		} else {
			data = make([]float64, r*c)
			for i := 0; i < r; i++ {
				for j := 0; j < c; j++ {
					data[i*c+j] = m.At(i, j)
				}
			}
		}
		desc.format, desc.data = C.LDA_B200_DENSE, ptrF(data)
		keep = []interface{}{data}
	}
	return
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
	desc, keep := matrix(m)
	r, c := m.Dims()
	l.w, l.d = r, c
	var theta []float64
	if want {
		theta = make([]float64, l.K*c)
	}
	ctr := C.lda_b200_counters{rho_phi_t: C.double(l.rhoPhiT), rho_theta_t: C.double(l.rhoThetaT), words_in_corpus: C.double(l.wordsInCorpus)}
	rnd := l.pcg()
	rc := C.lda_b200_fit_matrix(h, &desc, nil, nil, &rnd, &ctr, ptrF(theta), nil, 0, nil, nil)
	runtime.KeepAlive(keep)
	if rc != 0 {
		return nil, l.err()
	}
	l.takePcg(rnd)
	l.rhoPhiT, l.rhoThetaT, l.wordsInCorpus = float64(ctr.rho_phi_t), float64(ctr.rho_theta_t), float64(ctr.words_in_corpus)
	if !want {
		return nil, nil
	}
	return mat.NewDense(l.K, c, theta), nil
}

// Transform is lda.go:446-453.
func (l *LatentDirichletAllocation) Transform(m mat.Matrix) (mat.Matrix, error) {
	h, err := l.handle()
	if err != nil {
		return nil, err
	}
	desc, keep := matrix(m)
	_, c := m.Dims()
	theta := make([]float64, l.K*c)
	rnd := l.pcg()
	rc := C.lda_b200_transform_matrix(h, &desc, nil, &rnd, C.double(l.rhoThetaT), ptrF(theta), nil)
	runtime.KeepAlive(keep)
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
	desc, keep := matrix(m)
	rnd := l.pcg()
	var out C.double
	rc := C.lda_b200_perplexity_matrix(h, &desc, nil, &rnd, C.double(l.rhoThetaT), &out)
	runtime.KeepAlive(keep)
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
