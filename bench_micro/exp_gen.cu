// Isolates the cross-kernel fragment generation (table exp) of the scoring kernel: per warp and
// iteration, KS x NT exponentials per lane exactly as gen_fragments produces them.  Reports
// Gexp/s and the equivalent FP64-pipe cost per exp relative to the measured DFMA issue rate.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../bayesian_quadrature_b200/csrc/bq_common.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
using namespace bqb;

template <int KS, int NT, int TABN, bool TL, int MODE>
__global__ void __launch_bounds__(256, 2) k_gen(double *out, const double *tab_g, int iters, double C, double x0) {
    __shared__ double s_tab[TABN];
    __shared__ double s_xs[KS * 4], s_tol[KS * 4], s_atl[KS * 4];
    for (int i = threadIdx.x; i < TABN; i += blockDim.x) s_tab[i] = tab_g[i];
    for (int i = threadIdx.x; i < KS * 4; i += blockDim.x) { s_xs[i] = 1.25 * (i - KS * 2); s_tol[i] = 1e-4; s_atl[i] = 1e-3 * i; }
    __syncthreads();
    const int lane = threadIdx.x & 31, kq = lane & 3, pq = lane >> 2;
    double x[NT], tm[NT];
    int close[NT];
    unsigned acc = 0;
    const int dmax = exp_d2max_hi(C / ExpC<TABN>::INVN);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { x[nt] = x0 + 0.37 * (pq + 8 * nt) + 1e-3 * blockIdx.x; tm[nt] = 0; close[nt] = 0; }
    for (int it = 0; it < iters; ++it) {
        double bf[KS][NT];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k = 4 * ks + kq;
            const double xs = s_xs[k];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double d = x[nt] - xs;
                double e;
                if (MODE == 0) e = exp_kernel<TABN>(d * d, C, dmax, s_tab);
                else if (MODE == 1) e = exp((d * d) * (C / ExpC<TABN>::INVN));
                else e = exp_tab<TABN>((d * d) * (C / ExpC<TABN>::INVN), s_tab);
                bf[ks][nt] = e;
                if (TL) {
                    close[nt] |= ((__double_as_longlong(d) & 0x7fffffffffffffffLL) <= __double_as_longlong(s_tol[k]));
                    tm[nt] = fma(s_atl[k], e, tm[nt]);
                }
            }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc ^= (unsigned)__double2loint(bf[ks][nt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) x[nt] += 1e-6;
    }
    if (acc == 12345u && (!TL || (tm[0] == 1.5 && close[0] == 1))) out[0] = acc;
}
template <typename F> static double time_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}
template <int KS, int NT, int TABN, bool TL, int MODE> static void run(const char *name, int sms, double *out, const double *tab) {
    const int iters = 2000, grid = sms * 2;
    double ms = time_ms([&] { k_gen<KS, NT, TABN, TL, MODE><<<grid, 256>>>(out, tab, iters, -0.2958 * ExpC<TABN>::INVN, -3.0); });
    double exps = (double)grid * 256 * iters * KS * NT;
    // DFMA issue rate: 148 SMs x 4 SMSP x 16 lanes x 1.93 GHz
    double gexp = exps / ms * 1e-6, dfma_rate = 148.0 * 64 * 1.93e9;
    printf("\"%s\": {\"ms\": %.4f, \"gexp_per_s\": %.1f, \"dfma_slots_per_exp\": %.2f}, ", name, ms, gexp, dfma_rate / (gexp * 1e9));
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out; CK(cudaMalloc(&out, 64));
    double h[2048 + 512];
    for (int j = 0; j < 2048; ++j) h[j] = (double)exp2l((long double)j / 2048);
    for (int j = 0; j < 512; ++j) h[2048 + j] = (double)exp2l((long double)j / 512);
    double *tab; CK(cudaMalloc(&tab, sizeof(h))); CK(cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("{");
    run<16, 2, 2048, false, 0>("kernel_exp_2048_L", p.multiProcessorCount, out, tab);
    run<16, 2, 2048, true, 0>("kernel_exp_2048_TL", p.multiProcessorCount, out, tab);
    run<16, 2, 512, false, 0>("kernel_exp_512_L", p.multiProcessorCount, out, tab + 2048);
    run<16, 1, 2048, false, 0>("kernel_exp_2048_L_nt1", p.multiProcessorCount, out, tab);
    run<16, 2, 2048, false, 2>("codywaite_exp_2048_L", p.multiProcessorCount, out, tab);
    run<16, 2, 2048, false, 1>("libm_exp_L", p.multiProcessorCount, out, tab);
    printf("\"sms\": %d}\n", p.multiProcessorCount);
    return 0;
}
