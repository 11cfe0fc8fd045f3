// +build cgo

// LinearScanIndex backed by liblda_b200.so: drop-in for index.go:60-135 when the comparer is one of the library's
// pairwise functions (measures/pairwise/comparisons.go).  NOT COMPILED IN THIS REPOSITORY'S BUILD ENVIRONMENT (no Go
// toolchain, DESIGN.md 1); nlp_b200/index.py performs the identical call sequence through ctypes and is what runs in
// tests/ and tools/bench_search.py.
package nlp

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../nlp_b200 -llda_b200 -Wl,-rpath,${SRCDIR}/../../nlp_b200
#include "lda_b200_index.h"
*/
import "C"

import (
	"errors"
	"reflect"
	"runtime"
	"sync"

	"github.com/james-bowman/nlp/measures/pairwise"
	"gonum.org/v1/gonum/mat"
)

// comparerCode maps the reference's comparer functions to the kernels specialised on them.  A func value cannot be
// compared in Go, so the mapping goes through the function's code pointer.
func comparerCode(fn pairwise.Comparer) (C.int32_t, bool) {
	table := []struct {
		fn   pairwise.Comparer
		code C.int32_t
	}{
		{pairwise.CosineSimilarity, C.LDA_B200_COSINE_SIMILARITY}, {pairwise.CosineDistance, C.LDA_B200_COSINE_DISTANCE},
		{pairwise.AngularDistance, C.LDA_B200_ANGULAR_DISTANCE}, {pairwise.AngularSimilarity, C.LDA_B200_ANGULAR_SIMILARITY},
		{pairwise.HammingDistance, C.LDA_B200_HAMMING_DISTANCE}, {pairwise.HammingSimilarity, C.LDA_B200_HAMMING_SIMILARITY},
		{pairwise.EuclideanDistance, C.LDA_B200_EUCLIDEAN_DISTANCE}, {pairwise.ManhattenDistance, C.LDA_B200_MANHATTEN_DISTANCE},
		{pairwise.VectorLenSimilarity, C.LDA_B200_VECTOR_LEN_SIMILARITY},
	}
	p := reflect.ValueOf(fn).Pointer()
	for _, e := range table {
		if reflect.ValueOf(e.fn).Pointer() == p {
			return e.code, true
		}
	}
	return 0, false
}

// B200LinearScanIndex implements Indexer (index.go:46-50).
type B200LinearScanIndex struct {
	lock sync.RWMutex
	h    *C.lda_b200_index
	ids  map[int64]interface{} // handle -> caller's id (Match.ID is interface{}, index.go:16)
	rev  map[interface{}][]int64
	next int64
}

// NewB200LinearScanIndex mirrors NewLinearScanIndex (index.go:74-76); it panics, like the rest of the library does on
// construction failures, when compareFN is not one of the pairwise functions or no B200 is present.
func NewB200LinearScanIndex(compareFN pairwise.Comparer, device int) *B200LinearScanIndex {
	code, ok := comparerCode(compareFN)
	if !ok {
		panic("nlp: B200LinearScanIndex needs one of the measures/pairwise comparers")
	}
	b := &B200LinearScanIndex{ids: map[int64]interface{}{}, rev: map[interface{}][]int64{}}
	if rc := C.lda_b200_index_create(C.int32_t(device), code, &b.h); rc != 0 {
		panic("nlp: " + C.GoString(C.lda_b200_index_last_error(nil)))
	}
	runtime.SetFinalizer(b, (*B200LinearScanIndex).Close)
	return b
}

func (b *B200LinearScanIndex) Close() {
	if b.h != nil {
		C.lda_b200_index_destroy(b.h)
		b.h = nil
	}
}

func (b *B200LinearScanIndex) err() error { return errors.New("nlp: " + C.GoString(C.lda_b200_index_last_error(b.h))) }

func dense(v mat.Vector) []float64 {
	if d, ok := v.(*mat.VecDense); ok && d.RawVector().Inc == 1 {
		return d.RawVector().Data
	}
	out := make([]float64, v.Len())
	for i := range out {
		out[i] = v.AtVec(i)
	}
	return out
}

// Index is index.go:79-84.
func (b *B200LinearScanIndex) Index(v mat.Vector, id interface{}) {
	b.lock.Lock()
	defer b.lock.Unlock()
	data := dense(v)
	handle := C.int64_t(b.next)
	if rc := C.lda_b200_index_add(b.h, C.int32_t(len(data)), 1, (*C.double)(&data[0]), 1, &handle); rc != 0 {
		panic(b.err())
	}
	runtime.KeepAlive(data)
	b.ids[b.next] = id
	b.rev[id] = append(b.rev[id], b.next)
	b.next++
}

// IndexColumns indexes every column of m in one call -- ColDo(m, func(j, v) { index.Index(v, ids[j]) }) -- e.g. the
// K x D doc-topic matrix returned by LatentDirichletAllocation.Transform.
func (b *B200LinearScanIndex) IndexColumns(m *mat.Dense, ids []interface{}) {
	b.lock.Lock()
	defer b.lock.Unlock()
	raw := m.RawMatrix()
	handles := make([]C.int64_t, raw.Cols)
	for j := range handles {
		handles[j] = C.int64_t(b.next + int64(j))
	}
	if rc := C.lda_b200_index_add(b.h, C.int32_t(raw.Rows), C.int64_t(raw.Cols), (*C.double)(&raw.Data[0]), C.int64_t(raw.Stride),
		&handles[0]); rc != 0 {
		panic(b.err())
	}
	runtime.KeepAlive(raw.Data)
	for j := 0; j < raw.Cols; j++ {
		b.ids[b.next] = ids[j]
		b.rev[ids[j]] = append(b.rev[ids[j]], b.next)
		b.next++
	}
}

// IndexTransform is ColDo(theta, func(j int, v mat.Vector) { index.Index(v, ids[j]) }) with theta, _ := lda.Transform(m),
// fused: the doc-topic vectors go from the Transform kernels into the index without leaving the GPU.
func (b *B200LinearScanIndex) IndexTransform(lda *LatentDirichletAllocation, m mat.Matrix, ids []interface{}) error {
	b.lock.Lock()
	defer b.lock.Unlock()
	h, err := lda.handle()
	if err != nil {
		return err
	}
	f := flatten(m)
	_, c := m.Dims()
	handles := make([]C.int64_t, c)
	for j := range handles {
		handles[j] = C.int64_t(b.next + int64(j))
	}
	theta0 := lda.draws(c) // lda.go:422-426 when the caller assigned Rnd, nil otherwise
	rnd := lda.pcg()
	var hp *C.int64_t
	if c > 0 {
		hp = &handles[0]
	}
	rc := C.lda_b200_transform_into_index_flat(h, f.format, f.rows, f.cols, ptrI(f.ptr), ptrI(f.ind), ptrF(f.data), ptrF(theta0), &rnd,
		C.double(lda.rhoThetaT), b.h, hp)
	runtime.KeepAlive(f)
	runtime.KeepAlive(theta0)
	if rc != 0 {
		return lda.err()
	}
	lda.takePcg(rnd)
	for j := 0; j < c; j++ {
		b.ids[b.next] = ids[j]
		b.rev[ids[j]] = append(b.rev[ids[j]], b.next)
		b.next++
	}
	return nil
}

// Search is index.go:85-115; the matches come nearest first (the reference returns them in heap order).
func (b *B200LinearScanIndex) Search(qv mat.Vector, k int) []Match {
	// The engine behind the handle is not re-entrant (one stream, shared timing events and error text), so searches are
	// serialised as well: the write lock, where index.go:90 takes the read lock.
	b.lock.Lock()
	defer b.lock.Unlock()
	q := dense(qv)
	ids := make([]C.int64_t, k)
	dist := make([]C.double, k)
	var found C.int32_t
	if rc := C.lda_b200_index_search(b.h, 1, (*C.double)(&q[0]), 1, C.int32_t(k), &ids[0], &dist[0], &found); rc != 0 {
		panic(b.err())
	}
	runtime.KeepAlive(q)
	out := make([]Match, int(found))
	for i := range out {
		out[i] = Match{Distance: float64(dist[i]), ID: b.ids[int64(ids[i])]}
	}
	return out
}

// Remove is index.go:117-135.
func (b *B200LinearScanIndex) Remove(id interface{}) {
	b.lock.Lock()
	defer b.lock.Unlock()
	hs := b.rev[id]
	if len(hs) == 0 {
		return
	}
	if rc := C.lda_b200_index_remove(b.h, C.int64_t(hs[0])); rc != 0 {
		panic(b.err())
	}
	delete(b.ids, hs[0])
	if len(hs) == 1 {
		delete(b.rev, id)
	} else {
		b.rev[id] = hs[1:]
	}
}
