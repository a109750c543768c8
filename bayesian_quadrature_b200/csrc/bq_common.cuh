// Shared definitions of the B200 expected-variance scoring path: the device-resident model
// block layout, status bits, the FP64 DMMA wrapper and the table-driven exp used to generate
// cross-kernel fragments.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bqb {

// ---- per-point status bits (mirror the reference's branches; host turns them into exceptions)
constexpr int ST_OK = 0;
constexpr int ST_SHORTCUT = 1;   // bq.py:456-459   x_a isclose to an observation -> (Zm^2, Zm)
constexpr int ST_NOTPD = 2;      // bq.py:481-490   bordered matrix not PD      -> (Zm^2, Zm)
constexpr int ST_ESM_INF = 4;    // bq.py:522-523   logger.warn
constexpr int ST_EM_INF = 8;     // bq.py:524-525   logger.warn
constexpr int ST_ESM_BAD = 16;   // bq.py:514-517   RuntimeError (NaN or negative esm)
constexpr int ST_EM_BAD = 32;    // bq.py:518-520   RuntimeError (NaN em)
constexpr int ST_XA_BAD = 64;    // bq.py:451-452   ValueError (NaN / inf x_a)

// ---- setup status (per model instance)
constexpr int SETUP_OK = 0;
constexpr int SETUP_KTL_NOTPD = 1;    // gp_log_l.Kxx not positive definite
constexpr int SETUP_KL_NOTPD = 2;     // gp_l.Kxx not positive definite
constexpr int SETUP_MEAN_TOO_LARGE = 3;  // bq.py:945-947 "GP mean is too large"
constexpr int SETUP_BAD_INPUT = 4;    // non-finite / non-positive inputs

constexpr int NC_MAX = 16;            // candidates per instance supported on device
constexpr int EXP_TAB = 512;          // entries of the 2^(j/512) table
constexpr double MAX_EXPONENT = 707.0101241711442;   // log(2^1020): gauss_c.pyx:16, bq.py:16
constexpr double EPS = 2.220446049250313e-16;        // np.finfo(float64).eps, bq_c.pyx:28

// ---- header slots of a model block (doubles)
enum Hdr {
    H_NS = 0, H_NC, H_NSP, H_STATUS, H_CL, H_NHL, H_CTL, H_NHTL, H_KAA_E, H_KAA_N, H_J1, H_KTT, H_MU,
    H_HL2, H_LB, H_LOGDETB, H_ZM, H_ZV, H_THRESH, H_LOGLH, H_BA_S, H_NDB, H_WL, H_COUNT = 32
};

// Model block layout (offsets in doubles) for an instance capacity of nsp_cap observations
// (multiple of 8) and NC_MAX candidates.  Every array is padded so that padded observations
// contribute exactly zero.
struct Layout {
    int nsp_cap;     // padded observation capacity (multiple of 8)
    int nb_cap;      // row blocks of 8
    int nks_cap;     // k-steps of 4
    int off_xs, off_tol, off_atl, off_xc, off_lc, off_s0, off_lcc0, off_wb, off_wa, off_ug0, off_ua0;
    int off_af_l_tri, off_af_l_dense, off_af_tl_tri;
    int n_small;     // doubles before the fragment arrays (header + vectors + small matrices)
    int total;       // doubles per instance (multiple of 32)
};

__host__ __device__ inline int tri_frags(int nb) { return nb * (nb + 1); }   // sum_{rb<nb} (2 rb + 2)

__host__ __device__ inline Layout make_layout(int nsp_cap) {
    Layout L;
    L.nsp_cap = nsp_cap;
    L.nb_cap = nsp_cap / 8;
    L.nks_cap = nsp_cap / 4;
    int o = H_COUNT;
    L.off_xs = o; o += nsp_cap;
    L.off_tol = o; o += nsp_cap;
    L.off_atl = o; o += nsp_cap;
    L.off_xc = o; o += NC_MAX;
    L.off_lc = o; o += NC_MAX;
    L.off_s0 = o; o += NC_MAX * NC_MAX;
    L.off_lcc0 = o; o += NC_MAX * NC_MAX;
    L.off_wb = o; o += NC_MAX;
    L.off_wa = o; o += NC_MAX;
    L.off_ug0 = o; o += NC_MAX;
    L.off_ua0 = o; o += NC_MAX;
    o = (o + 31) & ~31;
    L.n_small = o;
    L.off_af_l_tri = o; o += tri_frags(L.nb_cap) * 32;
    L.off_af_l_dense = o; o += 3 * L.nks_cap * 32;      // up to 3 dense row blocks (nc + 2 <= 24 rows)
    L.off_af_tl_tri = o; o += tri_frags(L.nb_cap) * 32;
    L.total = o;
    return L;
}

#ifdef __CUDACC__

// D(8x8) += A(8x4) * B(4x8), FP64 tensor path (SASS: DMMA.8x8x4).  Lane l holds
// A[l>>2][l&3], B[l&3][l>>2], C/D[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// exp(x) for x <= 0 (kernel exponents are -(d^2)/(2 w^2)), ~1 ulp: x = (512 e + j) ln2/512 + r,
// exp(x) = 2^e * T[j] * (1 + expm1(r)), |r| <= ln2/1024, Taylor degree 4 (remainder < 2e-18).
// 9 FP64-pipe operations instead of libm's ~21; results below 2^-1021 flush to zero.
__device__ __forceinline__ double exp_neg(double x, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52
    const double INV = 0x1.71547652b82fep+9;             // 512 / ln 2
    const double HI = 0x1.62e42fef00000p-10;             // ln2/512, top 33 bits
    const double LO = 0x1.473de6af278edp-43;
    double t = fma(x, INV, MAGIC);
    int ki = __double2loint(t);
    double kf = t - MAGIC;
    double r = fma(kf, -HI, x);
    r = fma(kf, -LO, r);
    double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    double p = fma(q, r * r, r);
    double T = tab[ki & (EXP_TAB - 1)];
    double y = fma(T, p, T);
    int hi = __double2hiint(y) + ((ki >> 9) << 20);
    y = __hiloint2double(hi, __double2loint(y));
    // x < -708 (including -inf): below the normal range -> 0.  Integer compare on the high word.
    return ((unsigned)__double2hiint(x) > 0xC0862000u) ? 0.0 : y;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace bqb
