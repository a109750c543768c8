// Setup kernel: one CTA per model instance (one BQ problem under one hyper-parameter set).
//
// Replaces, once per instance instead of once per query point, what the reference does inside
// bq.py:447-527 through the `gp` package and linalg_c (Gram build bq.py:465, dpotrf
// linalg_c.pyx:86, gp.mean/cov bq.py:493-496), and computes the two scalars every score needs:
// Z_mean (bq_c.pyx:157-213) and Z_var (bq_c.pyx:264-355, with int_int_K1_K2_K1
// gauss_c.pyx:416-531 and int_K1_K2 gauss_c.pyx:235-339 fused into the quadratic forms so the
// n x n integral matrices are never stored).  It then emits the operands of the scoring kernel
// in DMMA fragment order (see bq_score.cu).
//
// With L = chol(K_l(x_sc, x_sc)) split as [[L_ss, 0], [C, L_cc]] (observations first, candidates
// last) a jitter on candidate diagonals only changes L_cc, so everything that touches the ns x ns
// block is pattern independent:
//     v_s = L_ss^-1 k_s,   w = k_c - C v_s = k_c + W k_s  (W = -C L_ss^-1),
//     v_c = chol(S0 + j1 diag(P))^-1 w,   S0 = K_cc - C C^T.
#include "bq_common.cuh"

namespace bqb {

constexpr int SETUP_THREADS = 256;
constexpr double LOG_2PI = 1.8378770664093453;
constexpr double SQRT_2PI = 2.5066282746310002;


__device__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0;
    for (int i = 0; i < SETUP_THREADS / 32; ++i) s += red[i];   // fixed order: deterministic
    return s;
}

constexpr int SETUP_NMAX = 256 + NC_MAX;   // largest matrix order (ns <= 256, nc <= 16)

constexpr int CHB = 32;                      // Cholesky panel width
constexpr int CH_STRIDE = CHB + 1;           // padded panel stride: conflict-free when lanes walk rows
// dynamic shared memory (doubles) for matrices of order <= ncap: Cholesky block + panel, or 32 staged rows of L
__host__ __device__ inline int setup_smem_doubles(int ncap) {
    const int a = (CHB + ncap) * CH_STRIDE, b = CHB * (ncap + 1) + 2 * CHB * CH_STRIDE;     // chol_lower / tri_inverse
    return a > b ? a : b;
}

// In-place lower Cholesky of the row-major n x n matrix A (leading dimension ld, global / L2 memory).
// Returns 0, or j+1 when pivot j is not positive (same contract as LAPACK dpotrf's info,
// linalg_c.pyx:86-91).  Blocked right-looking: the 32 x 32 diagonal block is factorised in shared
// memory, the panel below it is solved one row per thread and kept in shared memory, and the trailing
// matrix receives ONE rank-32 update per panel (contiguous row accesses) instead of one global
// read-modify-write pass per column.
__device__ int chol_lower(double *A, int ld, int n, double *sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = SETUP_THREADS / 32;
    double *D = sm;                            // [CHB][CH_STRIDE]
    double *P = sm + CHB * CH_STRIDE;          // [rows below][CH_STRIDE]
    __shared__ int s_info;
    if (tid == 0) s_info = 0;
    for (int jb = 0; jb < n; jb += CHB) {
        const int w = (n - jb < CHB) ? n - jb : CHB;
        __syncthreads();
        for (int e = tid; e < w * w; e += SETUP_THREADS) {
            const int r = e / w, c = e - r * w;
            D[r * CH_STRIDE + c] = (c <= r) ? A[(size_t)(jb + r) * ld + jb + c] : 0.0;
        }
        __syncthreads();
        // diagonal block in shared memory, all threads: per column a pivot, the scaled column (32 threads) and the rank-1
        // update of the trailing block spread over the CTA; two barriers per column.  (One warp updating its rows
        // serially while seven waited cost ~15 k cycles per block.)  The pivots go to a side array so that nobody
        // overwrites D[j][j] while others still read it.
        {
            __shared__ double s_diag[CHB];
            for (int j = 0; j < w; ++j) {
                const double d = D[j * CH_STRIDE + j];
                if (!(d > 0.0)) { if (tid == 0 && s_info == 0) s_info = jb + j + 1; break; }     // uniform: every thread reads the same d
                const double dj = sqrt(d);
                const int m = w - j - 1;                                   // rows / columns still below / right of j
                if (tid == 0) s_diag[j] = dj;
                if (tid < m) D[(j + 1 + tid) * CH_STRIDE + j] /= dj;
                __syncthreads();
                for (int e = tid; e < m * m; e += SETUP_THREADS) {
                    const int rr = e / m, kk = e - rr * m;
                    if (kk <= rr) {
                        const int r = j + 1 + rr, k = j + 1 + kk;
                        D[r * CH_STRIDE + k] = fma(-D[r * CH_STRIDE + j], D[k * CH_STRIDE + j], D[r * CH_STRIDE + k]);
                    }
                }
                __syncthreads();
            }
            __syncthreads();
            if (s_info == 0 && tid < w) D[tid * CH_STRIDE + tid] = s_diag[tid];
        }
        __syncthreads();
        if (s_info) return s_info;
        for (int e = tid; e < w * w; e += SETUP_THREADS) {
            const int r = e / w, c = e - r * w;
            if (c <= r) A[(size_t)(jb + r) * ld + jb + c] = D[r * CH_STRIDE + c];
        }
        // panel: rows below the block, P[i] = A[i][jb:jb+w] D^-T (forward substitution per row, one thread per row)
        const int nbelow = n - jb - w;
        for (int i = tid; i < nbelow; i += SETUP_THREADS) {
            double *Ai = A + (size_t)(jb + w + i) * ld + jb;
            double *Pi = P + i * CH_STRIDE;
            for (int c = 0; c < w; ++c) {
                double sacc = Ai[c];
                for (int k = 0; k < c; ++k) sacc -= Pi[k] * D[c * CH_STRIDE + k];
                sacc /= D[c * CH_STRIDE + c];
                Pi[c] = sacc;
                Ai[c] = sacc;
            }
        }
        __syncthreads();
        // trailing update: A[i][k] -= P[i] . P[k] for jb+w <= k <= i  (warp per row, lanes over k)
        for (int i = warp; i < nbelow; i += nw) {
            const double *Pi = P + i * CH_STRIDE;
            double *Ai = A + (size_t)(jb + w + i) * ld + jb + w;
            for (int k = lane; k <= i; k += 32) {
                const double *Pk = P + k * CH_STRIDE;
                double sacc = 0;
                for (int c = 0; c < w; ++c) sacc = fma(Pi[c], Pk[c], sacc);
                Ai[k] -= sacc;
            }
        }
    }
    __syncthreads();
    return 0;
}

// X = L^-1 (lower, row-major, both ld; X is written as exact zeros above the diagonal), blocked by 32:
//     X_II = L_II^-1,      X_IJ = -X_II (sum_{K=J}^{I-1} L_IK X_KJ)   for J < I.
// Block rows are sequential (I needs the block rows above it); inside one, all threads work: the 32-row panel of L is
// staged in shared memory, thread (warp w, lane c) accumulates rows w, w+8, w+16, w+24 of column c of a 32 x 32 block
// product (L from shared memory as warp broadcasts, X_KJ from global / L2 memory, contiguous across the warp), the block
// passes through shared memory once, and the same thread finishes its four rows of X_IJ.  (The previous version gave
// every column to one thread, n^2 / 2 dependent global loads long: 27 % of the setup kernel's samples at ns = 128.)
__device__ void tri_inverse(const double *L, double *X, int ld, int n, double *sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = SETUP_THREADS / 32;
    const int rs = ld + 1;                               // staged row stride (ld = ncap)
    double *XD = sm + CHB * rs;                          // [32][33] X_II
    double *SJ = XD + CHB * CH_STRIDE;                   // [32][33] one block of the sum
    for (int ib = 0; ib < n; ib += CHB) {
        const int h = (n - ib < CHB) ? n - ib : CHB, wdt = ib + h;       // rows ib .. ib+h-1, columns 0 .. wdt-1
        __syncthreads();
        for (int e = tid; e < h * wdt; e += SETUP_THREADS) {
            const int r = e / wdt, k = e - r * wdt;
            sm[r * rs + k] = (k <= ib + r) ? L[(size_t)(ib + r) * ld + k] : 0.0;
        }
        __syncthreads();
        // X_II: forward substitution per column, one warp (lane = column)
        if (warp == 0) {
            const int c = lane;
            for (int r = 0; r < CHB; ++r) {
                double out = 0.0;
                if (r < h && c <= r && c < h) {
                    double sacc = (c == r) ? 1.0 : 0.0;
                    const double *row = sm + r * rs + ib;
                    for (int k = c; k < r; ++k) sacc = fma(-row[k], XD[k * CH_STRIDE + c], sacc);
                    out = sacc / row[r];
                }
                XD[r * CH_STRIDE + c] = out;             // own column only: no synchronisation needed between rows
            }
        }
        __syncthreads();
        for (int e = tid; e < h * h; e += SETUP_THREADS) {
            const int r = e / h, c = e - r * h;
            X[(size_t)(ib + r) * ld + ib + c] = XD[r * CH_STRIDE + c];
        }
        // zeros right of the diagonal block (rows of this block row, columns past it)
        for (int e = tid; e < h * (n - wdt); e += SETUP_THREADS) {
            const int r = e / (n - wdt), c = e - r * (n - wdt);
            X[(size_t)(ib + r) * ld + wdt + c] = 0.0;
        }
        // X_IJ for the block columns left of the diagonal
        for (int jb = 0; jb < ib; jb += CHB) {
            double acc[CHB / 8];
#pragma unroll
            for (int q = 0; q < CHB / 8; ++q) acc[q] = 0.0;
            for (int k = jb; k < ib; ++k) {              // X[k][jb + lane] is zero for k < jb + lane (upper part of X_JJ)
                const double xk = X[(size_t)k * ld + jb + lane];
#pragma unroll
                for (int q = 0; q < CHB / 8; ++q) acc[q] = fma(sm[(warp + nw * q) * rs + k], xk, acc[q]);
            }
#pragma unroll
            for (int q = 0; q < CHB / 8; ++q) SJ[(warp + nw * q) * CH_STRIDE + lane] = acc[q];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < CHB / 8; ++q) {
                const int r = warp + nw * q;
                if (r < h) {
                    double sacc = 0.0;
                    for (int k = 0; k <= r; ++k) sacc = fma(XD[r * CH_STRIDE + k], SJ[k * CH_STRIDE + lane], sacc);
                    X[(size_t)(ib + r) * ld + jb + lane] = -sacc;
                }
            }
            __syncthreads();                             // SJ is reused by the next block column
        }
    }
    __syncthreads();
}

// out[i] = sum_{k<=i} X[i][k] v[k]   (warp per row)
__device__ void lower_matvec(const double *X, int ld, int n, const double *v, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = SETUP_THREADS / 32;
    for (int i = warp; i < n; i += nw) {
        double s = 0;
        for (int k = lane; k <= i; k += 32) s += X[(size_t)i * ld + k] * v[k];
        s = warp_sum(s);
        if (lane == 0) out[i] = s;
    }
    __syncthreads();
}

// out[c] = sum_{i>=c} X[i][c] v[i]: warp w takes rows w, w + 8, ... with its lanes across the columns (contiguous reads),
// partial sums per warp in registers, combined through shared memory in warp order (deterministic).  (One thread per
// column walking the rows was 10 % of the kernel's samples: n dependent global loads per thread.)
__device__ void lower_matvec_t(const double *X, int ld, int n, const double *v, double *out, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = SETUP_THREADS / 32;
    constexpr int NQ = (SETUP_NMAX + 31) / 32;
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
    for (int i = warp; i < n; i += nw) {
        const double vi = v[i];
        const double *row = X + (size_t)i * ld;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            if (c <= i) acc[q] = fma(row[c], vi, acc[q]);
        }
    }
    __syncthreads();                                     // sm may still be in use by the caller's previous phase
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        if (c < n) sm[warp * ld + c] = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += SETUP_THREADS) {
        double s = 0;
        for (int w = 0; w < nw; ++w) s += sm[w * ld + c];
        out[c] = s;
    }
    __syncthreads();
}

// gp's GaussianKernel: h^2 / (sqrt(2 pi) w) * exp(-0.5 d^2 / w^2)
__device__ __forceinline__ double gauss_k(double c, double w, double d) { return c * exp(-0.5 * (d * d) / (w * w)); }

// gauss_c.pyx:20-62 for d = 1 with L = sqrt(var), logdet = 2 log L
__device__ __forceinline__ double mvn_logpdf1(double x, double m, double L, double logdet) {
    const double diff = x - m;
    const double buf = (diff / L) / L;
    return -0.5 * ((LOG_2PI + logdet) + diff * buf);
}

__global__ void __launch_bounds__(SETUP_THREADS, 3) bq_setup_kernel(SetupArgs a) {
    __shared__ double red[SETUP_THREADS / 32];
    __shared__ double sm_small[4 * NC_MAX * NC_MAX];
    extern __shared__ double s_dyn[];             // setup_smem_doubles(ncap): Cholesky block + panel / row staging of tri_inverse
    double *s_vec = s_dyn;
    __shared__ int s_fail;
    const int inst = a.inst0 + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = SETUP_THREADS / 32;
    const int ns = a.ns[inst], nc = a.nc[inst], n = ns + nc;
    const double *x_s = a.x_s + (size_t)inst * a.in_stride;
    const double *l_s = a.l_s + (size_t)inst * a.in_stride;
    const double *x_c = a.x_c + (size_t)inst * NC_MAX;
    const double h_tl = a.hyp[inst * 6 + 0], w_tl = a.hyp[inst * 6 + 1], s_tl = a.hyp[inst * 6 + 2];
    const double h_l = a.hyp[inst * 6 + 3], w_l = a.hyp[inst * 6 + 4], s_l = a.hyp[inst * 6 + 5];
    const double mu = a.prior[inst * 3 + 0], sig2 = a.prior[inst * 3 + 1], thresh = a.prior[inst * 3 + 2];
    double *M = a.models + (size_t)inst * a.lay.total;
    const Layout &lay = a.lay;

    const int ncap = a.n_cap;
    double *W0 = a.work + (size_t)blockIdx.x * a.work_stride;
    double *Ltl = W0;                                  // ns x ns   K_tl -> L_tl
    double *Xtl = Ltl + (size_t)ncap * ncap;           // ns x ns   L_tl^-1
    double *Ll = Xtl + (size_t)ncap * ncap;            // n x n     K_l -> L
    double *Xl = Ll + (size_t)ncap * ncap;             // ns x ns   L_ss^-1   (later reused for K_l + s_l^2 I)
    double *vec = Xl + (size_t)ncap * ncap;
    double *tl_s = vec, *a_tl = vec + ncap, *x_sc = vec + 2 * ncap, *l_sc = vec + 3 * ncap, *b_sc = vec + 4 * ncap,
           *u_s = vec + 5 * ncap, *ua_s = vec + 6 * ncap, *gg = vec + 7 * ncap, *ga = vec + 8 * ncap,
           *alpha = vec + 9 * ncap, *beta = vec + 10 * ncap, *tmp = vec + 11 * ncap, *tmp2 = vec + 12 * ncap,
           *Wm = vec + 13 * ncap;   // W: NC_MAX x ncap, uses the tail (sized 3 + NC_MAX vectors)

    // zero the whole model block first (padding must be exact zeros)
    for (int i = tid; i < lay.total; i += SETUP_THREADS) M[i] = 0.0;
    if (tid == 0) s_fail = SETUP_OK;
    __syncthreads();

    // ---- P0 validate
    {
        int bad = 0;
        for (int i = tid; i < ns; i += SETUP_THREADS) bad |= !(isfinite(x_s[i]) && isfinite(l_s[i]) && l_s[i] > 0.0);
        for (int i = tid; i < nc; i += SETUP_THREADS) bad |= !isfinite(x_c[i]);
        if (tid == 0)
            bad |= !(h_tl > 0 && w_tl > 0 && s_tl >= 0 && h_l > 0 && w_l > 0 && s_l >= 0 && sig2 > 0 && isfinite(mu) &&
                     ns >= 1 && ns <= lay.nsp_cap && nc >= 0 && nc <= NC_MAX);
        if (bad) s_fail = SETUP_BAD_INPUT;
    }
    __syncthreads();
    if (s_fail) { if (tid == 0) M[H_STATUS] = s_fail; return; }

    // ---- P0b: observations in ascending order of x (stable rank sort; n <= 256).  Every result of this kernel and of the
    // scoring kernel is invariant under a permutation of the observations (up to rounding), and with sorted observations
    // the k-steps that matter for a query point are few and contiguous (band skipping, bq_score.cu) whatever order the
    // caller -- or add_observation, which appends -- left them in.  Already sorted input is left exactly as it is.
    {
        double *xs_sorted = vec + 29 * (size_t)ncap, *ls_sorted = vec + 30 * (size_t)ncap;
        for (int i = tid; i < ns; i += SETUP_THREADS) {
            const double xi = x_s[i];
            int rank = 0;
            for (int j = 0; j < ns; ++j) rank += (x_s[j] < xi) || (x_s[j] == xi && j < i);
            xs_sorted[rank] = xi;
            ls_sorted[rank] = l_s[i];
        }
        __syncthreads();
        x_s = xs_sorted;
        l_s = ls_sorted;
    }

    const double c_tl = (h_tl * h_tl) / (SQRT_2PI * w_tl);
    const double c_l = (h_l * h_l) / (SQRT_2PI * w_l);

    // ---- P1/P2: tl_s = log l_s (bq.py:73); K_tl = K(x_s, x_s) + s_tl^2 I
    for (int i = tid; i < ns; i += SETUP_THREADS) tl_s[i] = log(l_s[i]);
    for (int e = tid; e < ns * ns; e += SETUP_THREADS) {
        const int i = e / ns, j = e - i * ns;
        double v = gauss_k(c_tl, w_tl, x_s[i] - x_s[j]);
        if (i == j) v += s_tl * s_tl;
        Ltl[(size_t)i * ncap + j] = v;
    }
    // ---- P3: L_tl
    if (chol_lower(Ltl, ncap, ns, s_vec)) { if (tid == 0) M[H_STATUS] = SETUP_KTL_NOTPD; return; }
    // ---- P4: L_tl^-1
    tri_inverse(Ltl, Xtl, ncap, ns, s_vec);
    // ---- P5: a_tl = K_tl^-1 tl_s
    lower_matvec(Xtl, ncap, ns, tl_s, tmp);
    lower_matvec_t(Xtl, ncap, ns, tmp, a_tl, s_vec);
    // ---- P6: l_c = exp(gp_log_l.mean(x_c))  (bq.py:985 / :942-950), one warp per candidate
    for (int j = warp; j < nc; j += nw) {
        double m = 0;
        for (int i = lane; i < ns; i += 32) m += gauss_k(c_tl, w_tl, x_c[j] - x_s[i]) * a_tl[i];
        m = warp_sum(m);
        if (a.check_max) {
            // V = diag(cov(x_c)) = ktt - |L_tl^-1 k|^2 ; LinAlgError if m + 2 sqrt(max(V, 0)) > MAX
            double q = 0;
            for (int r = lane; r < ns; r += 32) {
                double s = 0;
                for (int k = 0; k <= r; ++k) s += Xtl[(size_t)r * ncap + k] * gauss_k(c_tl, w_tl, x_c[j] - x_s[k]);
                q += s * s;
            }
            q = warp_sum(q);
            double V = c_tl - q;
            if (V < 0) V = 0;
            if (lane == 0 && m + 2 * sqrt(V) > MAX_EXPONENT) s_fail = SETUP_MEAN_TOO_LARGE;
        }
        if (lane == 0) M[lay.off_lc + j] = exp(m);
    }
    __syncthreads();
    if (s_fail) { if (tid == 0) M[H_STATUS] = s_fail; return; }
    for (int i = tid; i < n; i += SETUP_THREADS) {
        x_sc[i] = i < ns ? x_s[i] : x_c[i - ns];
        l_sc[i] = i < ns ? l_s[i] : M[lay.off_lc + (i - ns)];
    }
    __syncthreads();
    // ---- P7: K_l(x_sc, x_sc) without s_l^2 (the bordered matrix of bq.py:465 is Kxoxo) -> L
    for (int e = tid; e < n * n; e += SETUP_THREADS) {
        const int i = e / n, j = e - i * n;
        Ll[(size_t)i * ncap + j] = gauss_k(c_l, w_l, x_sc[i] - x_sc[j]);
    }
    __syncthreads();
    // keep K_cc for the Schur complement before it is overwritten
    double *Kcc = sm_small;                       // NC_MAX^2
    double *S0 = sm_small + NC_MAX * NC_MAX;      // NC_MAX^2
    for (int e = tid; e < nc * nc; e += SETUP_THREADS) {
        const int i = e / nc, j = e - i * nc;
        Kcc[i * NC_MAX + j] = Ll[(size_t)(ns + i) * ncap + ns + j];
    }
    if (chol_lower(Ll, ncap, n, s_vec)) { if (tid == 0) M[H_STATUS] = SETUP_KL_NOTPD; return; }
    // ---- P8: L_ss^-1
    tri_inverse(Ll, Xl, ncap, ns, s_vec);
    // ---- P9: b_sc = int_K (gauss_c.pyx:95-164): h^2 exp(mvn_logpdf(x; mu, w_l^2 + sigma2))
    const double var_b = sig2 + w_l * w_l;
    const double Lb = sqrt(var_b), logdet_b = 2 * log(Lb);
    for (int i = tid; i < n; i += SETUP_THREADS) b_sc[i] = (h_l * h_l) * exp(mvn_logpdf1(x_sc[i], mu, Lb, logdet_b));
    __syncthreads();
    // ---- P10: pattern-independent pieces
    lower_matvec(Xl, ncap, ns, b_sc, u_s);       // u_s  = L_ss^-1 b_s
    lower_matvec(Xl, ncap, ns, l_sc, ua_s);      // ua_s = L_ss^-1 l_s
    lower_matvec_t(Xl, ncap, ns, u_s, gg, s_vec);       // g_gamma = L_ss^-T u_s
    lower_matvec_t(Xl, ncap, ns, ua_s, ga, s_vec);      // g_alpha = L_ss^-T ua_s
    // W[j][k] = -sum_{i>=k} C[j][i] Linv[i][k],  C = L[ns + j][0:ns]
    for (int e = tid; e < nc * ns; e += SETUP_THREADS) {
        const int j = e / ns, k = e - j * ns;
        double s = 0;
        for (int i = k; i < ns; ++i) s += Ll[(size_t)(ns + j) * ncap + i] * Xl[(size_t)i * ncap + k];
        Wm[(size_t)j * ncap + k] = -s;
    }
    // wb = b_c - C u_s ; wa = l_c - C ua_s ; S0 = K_cc - C C^T  (warp per output)
    for (int j = warp; j < nc; j += nw) {
        double s1 = 0, s2 = 0;
        for (int i = lane; i < ns; i += 32) {
            const double c = Ll[(size_t)(ns + j) * ncap + i];
            s1 += c * u_s[i];
            s2 += c * ua_s[i];
        }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { M[lay.off_wb + j] = b_sc[ns + j] - s1; M[lay.off_wa + j] = l_sc[ns + j] - s2; }
    }
    for (int e = warp; e < nc * nc; e += nw) {
        const int i = e / nc, j = e - i * nc;
        double s = 0;
        for (int k = lane; k < ns; k += 32) s += Ll[(size_t)(ns + i) * ncap + k] * Ll[(size_t)(ns + j) * ncap + k];
        s = warp_sum(s);
        if (lane == 0) S0[i * NC_MAX + j] = Kcc[i * NC_MAX + j] - s;
    }
    double part = 0;
    for (int i = tid; i < ns; i += SETUP_THREADS) part += u_s[i] * ua_s[i];
    const double BA_S = block_sum(part, red);
    __syncthreads();
    // small candidate-block solves (nc <= 16): thread 0
    if (tid == 0) {
        for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j) {
                M[lay.off_s0 + i * NC_MAX + j] = S0[i * NC_MAX + j];
                M[lay.off_lcc0 + i * NC_MAX + j] = (j <= i) ? Ll[(size_t)(ns + i) * ncap + ns + j] : 0.0;
            }
        for (int i = 0; i < nc; ++i) {   // ug0 = L_cc^-1 wb ; ua0 = L_cc^-1 wa
            double s1 = M[lay.off_wb + i], s2 = M[lay.off_wa + i];
            for (int k = 0; k < i; ++k) {
                const double l = Ll[(size_t)(ns + i) * ncap + ns + k];
                s1 -= l * M[lay.off_ug0 + k];
                s2 -= l * M[lay.off_ua0 + k];
            }
            const double d = Ll[(size_t)(ns + i) * ncap + ns + i];
            M[lay.off_ug0 + i] = s1 / d;
            M[lay.off_ua0 + i] = s2 / d;
            M[lay.off_rd0 + i] = 1.0 / d;
        }
        // alpha_c = L_cc^-T ua0  (back substitution)
        for (int i = nc - 1; i >= 0; --i) {
            double s = M[lay.off_ua0 + i];
            for (int k = i + 1; k < nc; ++k) s -= Ll[(size_t)(ns + k) * ncap + ns + i] * alpha[ns + k];
            alpha[ns + i] = s / Ll[(size_t)(ns + i) * ncap + ns + i];
        }
    }
    __syncthreads();
    // ---- P11: alpha = K_l^-1 l_sc:  alpha_s = L_ss^-T (ua_s - C^T alpha_c)
    for (int i = tid; i < ns; i += SETUP_THREADS) {
        double s = ua_s[i];
        for (int j = 0; j < nc; ++j) s -= Ll[(size_t)(ns + j) * ncap + i] * alpha[ns + j];
        tmp[i] = s;
    }
    __syncthreads();
    lower_matvec_t(Xl, ncap, ns, tmp, alpha, s_vec);
    double sumlog_l = 0;
    {
        double p = 0;
        for (int i = tid; i < n; i += SETUP_THREADS) p += log(Ll[(size_t)i * ncap + i]);
        sumlog_l = block_sum(p, red);
    }
    // ---- fragments that need L / L_ss^-1 are written now, because the s_l != 0 path reuses Xl
    const int nsp = (ns + 7) & ~7, nb = nsp / 8, nks = nsp / 4;
    const int ndb = (nc + 2 + 7) / 8;
    for (int e = tid; e < tri_frags(nb) * 32; e += SETUP_THREADS) {
        const int f = e >> 5, l = e & 31;
        int rb = 0;
        while (tri_frags(rb + 1) <= f) ++rb;
        const int ks = f - tri_frags(rb);
        const int r = 8 * rb + (l >> 2), k = 4 * ks + (l & 3);
        const bool in = (r < ns) && (k <= r);
        M[lay.off_af_l_tri + e] = in ? c_l * Xl[(size_t)r * ncap + k] : 0.0;
        M[lay.off_af_tl_tri + e] = in ? c_tl * Xtl[(size_t)r * ncap + k] : 0.0;
    }
    for (int e = tid; e < ndb * nks * 32; e += SETUP_THREADS) {
        const int f = e >> 5, l = e & 31;
        const int db = f / nks, ks = f - db * nks;
        const int r = 8 * db + (l >> 2), k = 4 * ks + (l & 3);
        double v = 0.0;
        if (k < ns) {
            if (r < nc) v = c_l * Wm[(size_t)r * ncap + k];
            else if (r == nc) v = c_l * gg[k];
            else if (r == nc + 1) v = c_l * ga[k];
        }
        M[lay.off_af_l_dense + e] = v;
    }
    __syncthreads();
    // ---- s_l != 0: Z_mean / Z_var / log_lh use alpha_l = (K_l + s_l^2 I)^-1 l_sc (gp.inv_Kxx_y), while
    //      the bordered matrix above does not carry s_l^2 (SURVEY appendix A.2 asymmetry)
    if (s_l != 0.0) {
        double *Kz = Xl;
        for (int e = tid; e < n * n; e += SETUP_THREADS) {
            const int i = e / n, j = e - i * n;
            double v = gauss_k(c_l, w_l, x_sc[i] - x_sc[j]);
            if (i == j) v += s_l * s_l;
            Kz[(size_t)i * ncap + j] = v;
        }
        if (chol_lower(Kz, ncap, n, s_vec)) { if (tid == 0) M[H_STATUS] = SETUP_KL_NOTPD; return; }
        if (warp == 0) {   // forward then backward substitution, one warp
            for (int i = 0; i < n; ++i) {
                double s = 0;
                for (int k = lane; k < i; k += 32) s += Kz[(size_t)i * ncap + k] * tmp2[k];
                s = warp_sum(s);
                if (lane == 0) tmp2[i] = (l_sc[i] - s) / Kz[(size_t)i * ncap + i];
                __syncwarp();
            }
            for (int i = n - 1; i >= 0; --i) {
                double s = 0;
                for (int k = i + 1 + lane; k < n; k += 32) s += Kz[(size_t)k * ncap + i] * alpha[k];
                s = warp_sum(s);
                if (lane == 0) alpha[i] = (tmp2[i] - s) / Kz[(size_t)i * ncap + i];
                __syncwarp();
            }
        }
        __syncthreads();
        double p = 0;
        for (int i = tid; i < n; i += SETUP_THREADS) p += log(Kz[(size_t)i * ncap + i]);
        sumlog_l = block_sum(p, red);
    }
    // ---- P12: Z_mean = int_K . alpha_l  (bq_c.pyx:207-209)
    double p = 0;
    for (int i = tid; i < n; i += SETUP_THREADS) p += b_sc[i] * alpha[i];
    const double Zm = block_sum(p, red);
    // ---- P13: Z_var = alpha' M alpha - beta' K_tl^-1 beta  (bq_c.pyx:342-351)
    //   M_ij = h_l^4 h_tl^2 exp(N1_i + N1_j + N2_ij)          gauss_c.pyx:488-529 (d = 1)
    {
        const double A_ = sig2 * ((sig2 / Lb) / Lb);            // cov (W1 + cov)^-1 cov     :496-500
        const double C2 = w_tl * w_tl + 2 * sig2 - 2 * A_;     // :515
        const double L2 = sqrt(C2), logdet2 = 2 * log(L2);
        const double hh = (h_l * h_l * h_l * h_l) * (h_tl * h_tl);
        // B_i = cov (W1 + cov)^-1 x_i  -> tmp ; N1_i -> tmp2
        for (int i = tid; i < n; i += SETUP_THREADS) {
            tmp[i] = sig2 * ((x_sc[i] / Lb) / Lb);
            tmp2[i] = mvn_logpdf1(x_sc[i], mu, Lb, logdet_b);
        }
        __syncthreads();
        double acc = 0;
        for (int j = warp; j < n; j += nw) {                    // column j, lanes over i (dot12 order :343)
            double col = 0;
            for (int i = lane; i < n; i += 32)
                col += alpha[i] * (hh * exp(tmp2[i] + tmp2[j] + mvn_logpdf1(tmp[i], tmp[j], L2, logdet2)));
            col = warp_sum(col);
            if (lane == 0) acc += col * alpha[j];
        }
        const double aMa = block_sum(acc, red);
        //   int_K1_K2[i, j] = h_tl^2 h_l^2 N([x_s_i, x_sc_j] | [mu, mu], [[w_tl^2 + cov, cov], [cov, w_l^2 + cov]])
        //   gauss_c.pyx:305-337 with the 2 x 2 Cholesky done in closed form
        const double c00 = w_tl * w_tl + sig2, c11 = w_l * w_l + sig2;
        const double l00 = sqrt(c00), l10 = sig2 / l00, l11 = sqrt(c11 - l10 * l10);
        const double logdet12 = 2 * (log(l00) + log(l11));
        const double h12 = (h_tl * h_tl) * (h_l * h_l);
        for (int i = warp; i < ns; i += nw) {
            double s = 0;
            const double d0 = x_s[i] - mu;
            for (int j = lane; j < n; j += 32) {
                const double d1 = x_sc[j] - mu;
                // dpotrs: forward  y0 = d0/l00, y1 = (d1 - l10 y0)/l11 ; backward z1 = y1/l11, z0 = (y0 - l10 z1)/l00
                const double y0 = d0 / l00, y1 = (d1 - l10 * y0) / l11;
                const double z1 = y1 / l11, z0 = (y0 - l10 * z1) / l00;
                const double lp = -0.5 * ((LOG_2PI * 2 + logdet12) + (d0 * z0 + d1 * z1));
                s += (h12 * exp(lp)) * alpha[j];
            }
            s = warp_sum(s);
            if (lane == 0) beta[i] = s;
        }
        __syncthreads();
        lower_matvec(Xtl, ncap, ns, beta, tmp);                 // |L_tl^-1 beta|^2 = beta' K_tl^-1 beta
        double q = 0;
        for (int i = tid; i < ns; i += SETUP_THREADS) q += tmp[i] * tmp[i];
        const double beta2 = block_sum(q, red);
        // ---- P14: log marginal likelihood of both GPs (bq.py:546; gp.log_lh)
        double y1 = 0, y2 = 0, sl = 0;
        for (int i = tid; i < ns; i += SETUP_THREADS) { y1 += tl_s[i] * a_tl[i]; sl += log(Ltl[(size_t)i * ncap + i]); }
        for (int i = tid; i < n; i += SETUP_THREADS) y2 += l_sc[i] * alpha[i];
        const double yKy_tl = block_sum(y1, red), yKy_l = block_sum(y2, red), sumlog_tl = block_sum(sl, red);
        if (tid == 0) {
            M[H_ZM] = Zm;
            M[H_ZV] = aMa - beta2;
            M[H_LOGLH] = (-0.5 * yKy_tl - sumlog_tl - 0.5 * ns * LOG_2PI) + (-0.5 * yKy_l - sumlog_l - 0.5 * n * LOG_2PI);
        }
    }
    // ---- P15/P16: header and vectors
    if (tid == 0) {
        M[H_NS] = ns; M[H_NC] = nc; M[H_NSP] = nsp; M[H_STATUS] = SETUP_OK; M[H_NDB] = ndb;
        M[H_CL] = c_l; M[H_NHL] = -0.5 / (w_l * w_l); M[H_WL] = w_l;
        M[H_CTL] = c_tl; M[H_NHTL] = -0.5 / (w_tl * w_tl);
        // bq_c.pyx:136: jitter = max(EPS, np.max(M)) * 1e-4, np.max re-evaluated after the first pass
        const double j1 = fmax(EPS, c_l) * 1e-4;
        M[H_J1] = j1;
        M[H_KAA_E] = c_l + fmax(EPS, c_l) * 1e-4;
        M[H_KAA_N] = c_l + fmax(EPS, c_l + j1) * 1e-4;
        M[H_KTT] = c_tl;
        // isclose pre-filter of the scoring kernel: an upper bound of every tolerance squared
        double t2 = 0;
        for (int i = 0; i < ns; ++i) { const double t = 1e-4 + 1e-5 * fabs(x_s[i]); t2 = fmax(t2, t * t); }
        M[H_TOL2MAX] = t2 * 1.000001;
        M[H_MU] = mu; M[H_THRESH] = thresh; M[H_BA_S] = BA_S;
        // int_K at a new point: h^2 exp(-1/2 (log 2pi + logdet)) * exp(-1/2 diff^2 / var)  (gauss_c.pyx:110, :162)
        M[H_CB] = (h_l * h_l) * exp(-0.5 * (LOG_2PI + logdet_b)); M[H_NHB] = -0.5 / var_b;
    }
    for (int i = tid; i < lay.nsp_cap; i += SETUP_THREADS) {
        const bool in = i < ns;
        M[lay.off_xs + i] = in ? x_s[i] : 1e150;                         // padded: far away (d^2 = 1e300 stays finite, never the
                                                                         // nearest observation of a point) x zero operand columns
        M[lay.off_tol + i] = in ? 1e-4 + 1e-5 * fabs(x_s[i]) : -1.0;     // np.isclose(x_a, x_s, atol=1e-4), rtol 1e-5
        M[lay.off_atl + i] = in ? c_tl * a_tl[i] : 0.0;
    }
    for (int i = tid; i < nc; i += SETUP_THREADS) M[lay.off_xc + i] = x_c[i];
}

void launch_setup(const SetupArgs &a, int n_inst, cudaStream_t stream) {
    const int bytes = (int)sizeof(double) * setup_smem_doubles(a.n_cap);
    cudaFuncSetAttribute(bq_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    bq_setup_kernel<<<n_inst, SETUP_THREADS, bytes, stream>>>(a);
}

}  // namespace bqb
