// Do DMMA.8x8x4 and DFMA share one FP64 datapath on B200?  Times (a) DMMA only, (b) DFMA only,
// (c) both interleaved in the same warps with equal pipe time (1 DMMA : 8 DFMA), (d) half of the
// warps doing DMMA and half DFMA.  Also the dependent-chain latencies of DMMA and DFMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: DMMA only; 1: DFMA only; 2: interleaved; 3: warp-specialised (even warps DMMA, odd DFMA)
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
    double c0[4], c1[4], f[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-9 + i;
    const bool do_mma = (MODE == 0) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_fma = (MODE == 1) || (MODE == 2) || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(c0[i], c1[i], a, b);
        }
        if (do_fma) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

__global__ void k_lat_dmma(double *out, long long *cyc, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) dmma(c0, c1, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (c0 + c1 == 123.456) out[0] = c0;
}
__global__ void k_lat_dfma(double *out, long long *cyc, int iters, double a, double b) {
    double c0 = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) c0 = fma(c0, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    if (c0 == 123.456) out[0] = c0;
}

template <typename F> static double time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out; CK(cudaMalloc(&out, 64));
    long long *cyc; CK(cudaMalloc(&cyc, 64));
    const int grid = p.multiProcessorCount * 4, iters = 8192;
    double t0 = time_ms([&] { k_mix<0><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    double t1 = time_ms([&] { k_mix<1><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    double t2 = time_ms([&] { k_mix<2><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    double t3 = time_ms([&] { k_mix<3><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    k_lat_dmma<<<1, 32>>>(out, cyc, 4096, 0.999, 1e-3);
    k_lat_dfma<<<1, 32>>>(out, cyc, 4096, 0.999, 1e-3);
    long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("{\"dmma_only_ms\": %.4f, \"dfma_only_ms\": %.4f, \"interleaved_ms\": %.4f, \"warp_specialised_ms\": %.4f, "
           "\"sum_ms\": %.4f, \"max_ms\": %.4f, \"dmma_dep_latency_cyc\": %.1f, \"dfma_dep_latency_cyc\": %.1f}\n",
           t0, t1, t2, t3, t0 + t1, t0 > t1 ? t0 : t1, h[0] / 4096.0, h[1] / 4096.0);
    return 0;
}
