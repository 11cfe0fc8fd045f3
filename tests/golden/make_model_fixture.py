"""Writes tests/golden/lda_model_k3_w9.bin: an LDA model file in the layout of include/lda_b200.h (lda_b200_save).

    python tests/golden/make_model_fixture.py

The reference has no Save/Load for LDA (TODO at lda.go:158) and no Go toolchain / gonum source is available offline, so no
gonum-produced bytes exist to compare with: the mat.Dense record follows gonum's documented storage header (mat/io.go:
uint32 version, Form 'G', Packing 'F', Uplo 'A', Unit, Rows, Cols, KU, KL, little endian, then the elements row-major) --
PARITY UNPINNED against gonum itself.  What the fixture pins is that every writer in this repository (the C library and
the Python restatement used by the tests) keeps producing these exact bytes.  Values are exactly representable in
fp32, so a Load -> Save round trip through the device reproduces the file bit for bit."""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
K, W = 3, 9
COUNTERS = (5.0, 37.0, 1234.0)


def arrays():
    nphi = (np.arange(W * K, dtype=np.float64).reshape(W, K) * 0.125 + 0.5)
    nz = nphi.sum(0)
    return nphi, nz


def record(a):
    a = np.ascontiguousarray(a, dtype="<f8")
    return struct.pack("<IBBBBqqqq", 0xFE000001, ord("G"), ord("F"), ord("A"), 0, a.shape[0], a.shape[1], 0, 0) + a.tobytes()


def model_bytes():
    nphi, nz = arrays()
    return struct.pack("<Qddd", K, *COUNTERS) + record(nphi) + record(nz.reshape(1, -1))


if __name__ == "__main__":
    with open(os.path.join(HERE, "lda_model_k3_w9.bin"), "wb") as f:
        f.write(model_bytes())
    print(len(model_bytes()), "bytes")
