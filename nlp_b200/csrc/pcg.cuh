// pcg.cuh -- golang.org/x/exp/rand PCGSource (PCG XSL-RR 128/64) for host and device, with O(log n)
// jump-ahead so that the Rnd.Int() % n draws of lda.go:199-205, 422-426 and 472-475 can be produced on
// the GPU, bit-exact and in the reference's draw order, without shipping the initial matrices over PCIe.
// The dependency is an unpinned HEAD module whose source is not in the reference tree; this follows its
// published algorithm: state' = state*MUL + INC (mod 2^128); out = rotr64(hi ^ lo, hi >> 58); Seed(s): hi = lo = s.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldab {

typedef unsigned __int128 u128;

struct Pcg {
    u128 s;
};

__host__ __device__ __forceinline__ u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | (u128)lo; }
__host__ __device__ __forceinline__ u128 pcg_mul() { return mk128(2549297995355413924ULL, 4865540595714422341ULL); }
__host__ __device__ __forceinline__ u128 pcg_inc() { return mk128(6364136223846793005ULL, 1442695040888963407ULL); }

__host__ __device__ __forceinline__ uint64_t pcg_next(Pcg &p) {
    p.s = p.s * pcg_mul() + pcg_inc();
    const uint64_t hi = (uint64_t)(p.s >> 64), lo = (uint64_t)p.s;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    return (x >> r) | (x << ((64 - r) & 63));
}

// state after `delta` further steps (Brown, "Random number generation with arbitrary strides")
__host__ __device__ __forceinline__ void pcg_jump(Pcg &p, uint64_t delta) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mul(), cur_plus = pcg_inc();
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    p.s = acc_mult * p.s + acc_plus;
}

// rand.Rand.Int() on 64-bit: Uint64() & (1<<63 - 1);  draw = float64(Int() % n) / float64(n)
__host__ __device__ __forceinline__ double pcg_draw(Pcg &p, uint64_t n) {
    const uint64_t x = pcg_next(p) & 0x7fffffffffffffffULL;
    return (double)(x % n) / (double)n;
}

// Fills rows of a [rows][Kp] fp32 matrix from consecutive draws: local row r, topic k takes draw number
//   first_draw + global_row(r)*K + k      with global_row(r) = ((r / b)*R + rank)*b + r % b
// (b = batch size, R ranks: the rows of a rank are its minibatches back to back; R = 1 gives identity).
// One thread produces CH consecutive topics of one row.
constexpr int PCG_CH = 16;
__global__ void pcg_fill_rows_kernel(uint64_t hi, uint64_t lo, uint64_t first_draw, uint64_t modulus, int64_t rows, int K,
                                     int Kp, int64_t b, int R, int rank, float *out) {
    const int chunks = (K + PCG_CH - 1) / PCG_CH;
    const int64_t total = rows * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / chunks;
        const int k0 = (int)(i - r * chunks) * PCG_CH;
        const int64_t gr = ((r / b) * R + rank) * b + r % b;
        Pcg p;
        p.s = mk128(hi, lo);
        pcg_jump(p, first_draw + (uint64_t)gr * (uint64_t)K + (uint64_t)k0);
        const int k1 = min(K, k0 + PCG_CH);
        for (int k = k0; k < k1; ++k) out[r * Kp + k] = (float)pcg_draw(p, modulus);
        if (k1 == K)
            for (int k = K; k < Kp; ++k) out[r * Kp + k] = 0.f;
    }
}

}  // namespace ldab
