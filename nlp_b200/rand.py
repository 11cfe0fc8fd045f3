"""Host-side stand-in for golang.org/x/exp/rand as the reference uses it (lda.go:10,126,184,201): `Rand` over a
PCG XSL-RR 128/64 `PCGSource`.  The arithmetic lives in liblda_b200.so (csrc/pcg.cuh) so that host and device
draw the identical stream."""
import numpy as np

from . import _lib


class PCGSource:
    """rand.NewSource(seed): Seed sets low = high = seed."""

    def __init__(self, seed=0):
        self._s = _lib.Pcg()
        self.Seed(seed)

    def Seed(self, seed):
        _lib.lib().lda_b200_pcg_seed(self._s, seed & 0xFFFFFFFFFFFFFFFF)

    def Uint64(self):
        return int(_lib.lib().lda_b200_pcg_uint64(self._s))

    # PCGSource.MarshalBinary order: high, low
    @property
    def state(self):
        return int(self._s.high), int(self._s.low)

    @state.setter
    def state(self, hl):
        self._s.high, self._s.low = hl


def NewSource(seed):
    return PCGSource(seed)


class Rand:
    """rand.New(src)"""

    def __init__(self, src):
        self.src = src

    def Uint64(self):
        return self.src.Uint64()

    def Int63(self):
        return self.src.Uint64() & ((1 << 63) - 1)

    def Int(self):
        return self.Int63()

    def fill_mod(self, count, n):
        """[float64(Int() % n) / float64(n) for _ in range(count)] (lda.go:201, 425, 474)"""
        out = np.empty(count, dtype=np.float64)
        _lib.lib().lda_b200_pcg_fill(self.src._s, count, n, out.ctypes.data)
        return out


def New(src):
    return Rand(src)
