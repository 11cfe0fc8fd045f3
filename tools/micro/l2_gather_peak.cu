// Microbenchmark behind DESIGN.md section 8: what can a B200 deliver for the memory access mix of the SCVB0 minibatch
// kernel, with the arithmetic and the per-document dependency chain taken away?
//   gather : every warp reads 1 KB rows (K = 256 fp32) of a 102 MB matrix at random word ids (Zipf-ish or uniform) with
//            coalesced 128-bit loads, many rows in flight  -- the rows of C a token visit fetches
//   red    : every warp adds a 1 KB row into a second 102 MB matrix with red.global.add.v4.f32 -- the nPhiHat update
//   mix    : two gathers and one reduction per token, the ratio of a BurnInPasses = 1 iteration
// Reports bytes / second for each, which is the ceiling the kernel's algorithmic bytes can be compared with.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_gather_peak l2_gather_peak.cu && ./l2_gather_peak
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cg(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void red_add_v2(float *addr, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// the same 1 KB row reductions in other forms: RW = 1 scalar red.f32 (128 B per warp instruction), 2 = red.v2.f32,
// 3 = TMA bulk reduction (cp.reduce.async.bulk .add.f32) of a 1 KB row staged in shared memory, one instruction per row
template <int RW, int NV>
__global__ void __launch_bounds__(256) red_probe(float *hat, const int *ids, int n) {
    extern __shared__ __align__(128) float stage[];   // RW == 3: [8 warps][4 slots][row]
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const size_t row = 128 * NV;
    float *mine = stage + (threadIdx.x >> 5) * 4 * row;
    int slot = 0;
    for (int t = warp; t < n; t += nwarps) {
        float *dst = hat + (size_t)ids[t] * row;
        if (RW == 1) {
#pragma unroll
            for (int j = 0; j < 4 * NV; ++j) atomicAdd(dst + j * 32 + lane, 1.f);
        } else if (RW == 4) {   // 32-bit integer adds, 4 bytes per lane
#pragma unroll
            for (int j = 0; j < 4 * NV; ++j) atomicAdd(reinterpret_cast<unsigned *>(dst) + j * 32 + lane, 1u);
        } else if (RW == 5) {   // 64-bit integer adds, 8 bytes per lane (the deterministic mode's fixed-point image)
#pragma unroll
            for (int j = 0; j < 2 * NV; ++j) atomicAdd(reinterpret_cast<unsigned long long *>(dst) + j * 32 + lane, 1ull);
        } else if (RW == 6) {   // float64 adds, 8 bytes per lane
#pragma unroll
            for (int j = 0; j < 2 * NV; ++j) atomicAdd(reinterpret_cast<double *>(dst) + j * 32 + lane, 1.0);
        } else if (RW == 2) {
#pragma unroll
            for (int j = 0; j < 2 * NV; ++j) red_add_v2(dst + 2 * (j * 32 + lane), make_float2(1.f, 2.f));
        } else {
            asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");   // the slot's previous reduction has read it
            __syncwarp();
#pragma unroll
            for (int v = 0; v < NV; ++v) reinterpret_cast<float4 *>(mine + slot * row)[v * 32 + lane] = make_float4(1.f, 2.f, 3.f, 4.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(mine + slot * row);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src),
                             "r"((unsigned)(row * 4))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            slot = (slot + 1) & 3;
        }
    }
    if (RW == 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// MODE 0: gather, 1: red, 2: mix (2 gathers + 1 red per token).  NV float4 per lane and row (row = 128 * NV floats).
template <int MODE, int NV, int U>
__global__ void __launch_bounds__(256) probe(const float *cmat, float *hat, const int *ids, int n, float *sink) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const size_t row = 128 * NV;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t0 = warp * U; t0 + U <= n; t0 += nwarps * U) {
        int w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = ids[t0 + u];
        if (MODE == 0 || MODE == 2) {
            float4 r[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v) r[u][v] = ld_cg(reinterpret_cast<const float4 *>(cmat + (size_t)w[u] * row) + v * 32 + lane);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc.x += r[u][v].x;
                    acc.y += r[u][v].y;
                    acc.z += r[u][v].z;
                    acc.w += r[u][v].w;
                }
        }
        if (MODE == 2) {  // the second pass over the same token: another word's row (the document's rows do not stay on chip)
            float4 r[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    r[u][v] = ld_cg(reinterpret_cast<const float4 *>(cmat + (size_t)ids[(t0 + u + n / 2) % n] * row) + v * 32 + lane);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc.x += r[u][v].x;
                    acc.y += r[u][v].y;
                    acc.z += r[u][v].z;
                    acc.w += r[u][v].w;
                }
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < NV; ++v) red_add_v4(hat + (size_t)w[u] * row + 4 * (v * 32 + lane), make_float4(1.f, 2.f, 3.f, 4.f));
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const int W = 100000, NV = 2;
    const int n = argc > 1 ? atoi(argv[1]) : 13200000;   // token visits of one C3 launch (6.6M tokens x 2 passes)
    const size_t row = 128 * NV;
    float *cmat, *hat, *sink;
    int *ids;
    CK(cudaMalloc(&cmat, W * row * sizeof(float)));
    CK(cudaMalloc(&hat, W * row * sizeof(float)));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&ids, (size_t)n * sizeof(int)));
    CK(cudaMemset(cmat, 0, W * row * sizeof(float)));
    CK(cudaMemset(hat, 0, W * row * sizeof(float)));
    int sms = 148;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    sms = prop.multiProcessorCount;
    printf("device: %s, %d SMs\n", prop.name, sms);
    const char *ids_file = argc > 2 ? argv[2] : nullptr;   // int32 word ids of a real token stream (e.g. the C3 corpus)
    for (int dist = ids_file ? 2 : 0; dist < (ids_file ? 3 : 2); ++dist) {
        // word ids: uniform, or Zipf(1.07) over a shuffled vocabulary like corpus_synth's base distribution
        std::vector<int> h(n);
        std::vector<double> cdf(W);
        double s = 0;
        for (int i = 0; i < W; ++i) {
            s += dist == 0 ? 1.0 : 1.0 / pow(i + 2.7, 1.07);
            cdf[i] = s;
        }
        std::vector<int> perm(W);
        for (int i = 0; i < W; ++i) perm[i] = i;
        unsigned long long x = 88172645463325252ull;
        auto rnd = [&]() {
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            return x;
        };
        for (int i = W - 1; i > 0; --i) std::swap(perm[i], perm[rnd() % (i + 1)]);
        for (int i = 0; i < n; ++i) {
            const double u = (rnd() >> 11) * (1.0 / 9007199254740992.0) * s;
            const int j = (int)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
            h[i] = perm[j < W ? j : W - 1];
        }
        if (dist == 2) {
            FILE *f = fopen(ids_file, "rb");
            if (!f) { fprintf(stderr, "cannot open %s\n", ids_file); return 1; }
            const size_t got = fread(h.data(), sizeof(int), (size_t)n, f);
            fclose(f);
            for (size_t i = got; i < (size_t)n; ++i) h[i] = h[i - got];   // repeat the stream (the second pass)
            printf("(%zu ids read from %s)\n", got, ids_file);
        }
        CK(cudaMemcpy(ids, h.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        printf("word ids: %s\n", dist == 0 ? "uniform over 100k words" : dist == 1 ? "Zipf(1.07) over 100k words" : "token stream from file");
        for (int occ = 4; occ <= 8; occ += 4) {
            const int grid = sms * occ;
            const double gb = (double)n * row * 4 / 1e9;
            float ms;
            ms = time_ms([&] { probe<0, NV, 4><<<grid, 256>>>(cmat, hat, ids, n, sink); }, 5);
            printf("  %d CTAs/SM  gather        : %7.3f ms  %7.2f TB/s (read)\n", occ, ms, gb / ms);
            ms = time_ms([&] { probe<0, NV, 8><<<grid, 256>>>(cmat, hat, ids, n, sink); }, 5);
            printf("  %d CTAs/SM  gather (U=8)  : %7.3f ms  %7.2f TB/s (read)\n", occ, ms, gb / ms);
            ms = time_ms([&] { probe<1, NV, 4><<<grid, 256>>>(cmat, hat, ids, n / 2, sink); }, 5);
            printf("  %d CTAs/SM  red           : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { red_probe<1, NV><<<grid, 256>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  red scalar    : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { red_probe<2, NV><<<grid, 256>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  red v2        : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { red_probe<4, NV><<<grid, 256>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  red u32       : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { red_probe<5, NV><<<grid, 256>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  red u64       : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { red_probe<6, NV><<<grid, 256>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  red f64       : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            CK(cudaFuncSetAttribute(red_probe<3, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 128 * NV * 4));
            ms = time_ms([&] { red_probe<3, NV><<<grid, 256, 8 * 4 * 128 * NV * 4>>>(hat, ids, n / 2); }, 5);
            printf("  %d CTAs/SM  TMA bulk red  : %7.3f ms  %7.2f TB/s (reduction payload)\n", occ, ms, gb / 2 / ms);
            ms = time_ms([&] { probe<2, NV, 4><<<grid, 256>>>(cmat, hat, ids, n / 2, sink); }, 5);
            printf("  %d CTAs/SM  2 gather+1 red: %7.3f ms  %7.2f TB/s (read + reduction payload; the kernel's mix)\n", occ, ms,
                   gb * 1.5 / ms);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
