// FP64 roofline denominators for B200 (sm_100a), measured because MEASURED_PEAKS.json
// carries no FP64 figure.  Prints one JSON object.
//   dfma      : dependent-free DFMA chains (CUDA-core FP64 pipe)
//   dmma_*    : mma.sync f64 shapes (legacy tensor path; tcgen05 has no f64 kind)
//   exp       : CUDA libm exp(double) throughput
//   lds_dfma  : DFMA with one broadcast shared-memory operand per RM FMAs
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs per lane. 256 FMA per instruction.
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs. 1024 FMA per instruction.
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs. 512 FMA per instruction.
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1684(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs. 2048 FMA per instruction.
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_exp(double *out, int iters, double a) {
    double x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) x[i] = -1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = exp(x[i]) - a;   // stays in (-1, 0]
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

// RM x RN register tile: per k-step, RM broadcast LDS doubles and RN lane-private LDS doubles feed RM*RN DFMA
template <int RM, int RN>
__global__ void __launch_bounds__(256) k_lds_dfma(double *out, int iters) {
    __shared__ __align__(16) double sa[64 * RM];        // broadcast operand
    __shared__ __align__(16) double sb[8 * 32 * RN];       // an 8-row window, shared by all warps
    for (int i = threadIdx.x; i < 64 * RM; i += blockDim.x) sa[i] = 1.0 + 1e-9 * i;
    for (int i = threadIdx.x; i < 8 * 32 * RN; i += blockDim.x) sb[i] = 1e-9 * i;
    __syncthreads();
    double acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0;
    const int lane_off = (threadIdx.x & 31) * RN;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            double a[RM], b[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = sa[k * RM + i];
#pragma unroll
            for (int j = 0; j < RN; ++j) b[j] = sb[(k & 7) * 32 * RN + lane_off + j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) s += acc[i][j];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out; CK(cudaMalloc(&out, 64));
    const int nsm = p.multiProcessorCount;
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, nsm);
    for (int bps = 2; bps <= 8; bps *= 2) {
        const int grid = nsm * bps; const double thr = (double)grid * 256;
        double ms;
        ms = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dfma8_tflops_bps%d\": %.3f", bps, thr * iters * 8 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dfma16_tflops_bps%d\": %.3f", bps, thr * iters * 16 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dmma884_tflops_bps%d\": %.3f", bps, (thr / 32) * iters * 8 * 512.0 / ms * 1e-9);
        ms = time_ms([&] { k_dmma1684<4><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dmma1684_tflops_bps%d\": %.3f", bps, (thr / 32) * iters * 4 * 1024.0 / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<4><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dmma1688_tflops_bps%d\": %.3f", bps, (thr / 32) * iters * 4 * 2048.0 / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<4><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
        printf(", \"dmma16816_tflops_bps%d\": %.3f", bps, (thr / 32) * iters * 4 * 4096.0 / ms * 1e-9);
        ms = time_ms([&] { k_exp<4><<<grid, 256>>>(out, iters / 4, 1.0); });
        printf(", \"exp_gexps_bps%d\": %.3f", bps, thr * (iters / 4) * 4 / ms * 1e-6);
    }
    {
        const int grid = nsm * 2; const double thr = (double)grid * 256; const int it2 = 256; double ms;
        ms = time_ms([&] { k_lds_dfma<4, 2><<<grid, 256>>>(out, it2); });
        printf(", \"lds_dfma_4x2_tflops\": %.3f", thr * it2 * 64 * 8 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_lds_dfma<8, 2><<<grid, 256>>>(out, it2); });
        printf(", \"lds_dfma_8x2_tflops\": %.3f", thr * it2 * 64 * 16 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_lds_dfma<8, 4><<<grid, 256>>>(out, it2); });
        printf(", \"lds_dfma_8x4_tflops\": %.3f", thr * it2 * 64 * 32 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_lds_dfma<16, 2><<<grid, 256>>>(out, it2); });
        printf(", \"lds_dfma_16x2_tflops\": %.3f", thr * it2 * 64 * 32 * 2 / ms * 1e-9);
        ms = time_ms([&] { k_lds_dfma<16, 4><<<grid, 256>>>(out, it2); });
        printf(", \"lds_dfma_16x4_tflops\": %.3f", thr * it2 * 64 * 64 * 2 / ms * 1e-9);
    }
    printf("}\n");
    return 0;
}
