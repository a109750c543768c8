/* TEST INFRASTRUCTURE — CPU oracle for the expected-variance active-sampling path.
 *
 * A plain-C, double-precision RESTATEMENT of the reference algorithm (not of the
 * product's bordered/DMMA algorithm): for every query point it rebuilds the full
 * bordered Gram matrix, applies the jitter passes, factorises it from scratch and
 * evaluates the closed-form integrals exactly as the reference does.  Each function
 * cites the reference file:line it follows (paths relative to /root/reference/).
 *
 * The third-party `gp` package (gaussian_processes==1.0.5, requirements.txt:2) is absent
 * from /root/reference; its published behaviour is restated in
 * the orc_gp_* functions, including the explicit-inverse route for mean/cov.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the seven
 * 12-digit goldens printed in docs/ipynb/visual-tests.ipynb and against fixtures produced
 * by the unmodified reference (oracle/_ref, tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product never does.
 *
 * 1-D only (the reference's BQ class rejects x.ndim > 1, bq.py:65).  Matrices are
 * column-major like the reference's Fortran-ordered memoryviews.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_SHORTCUT 1   /* bq.py:456-459  */
#define ORC_NOTPD 2      /* bq.py:481-490  */
#define ORC_ESM_INF 4    /* bq.py:522-523 (logger.warn) */
#define ORC_EM_INF 8     /* bq.py:524-525 (logger.warn) */
#define ORC_ESM_BAD 16   /* bq.py:514-517 (RuntimeError) */
#define ORC_EM_BAD 32    /* bq.py:518-520 (RuntimeError) */
#define ORC_XA_BAD 64    /* bq.py:451-452 (ValueError)   */

static const double ORC_EPS = 2.220446049250313e-16; /* np.finfo(float64).eps, bq_c.pyx:28 */

/* gauss_c.pyx:16  MAX = log(exp2(float64(finfo(float64).maxexp - 4))) = log(2^1020) */
static double orc_max_exponent(void) { return log(ldexp(1.0, 1020)); }

/* ---------------------------------------------------------------- linalg_c.pyx */

/* linalg_c.pyx:55-93 cho_factor -> LAPACK dpotrf('L'), column-major; returns info
 * (0 ok, j+1 if the leading minor of order j+1 is not positive definite).  Unblocked
 * left-looking variant (dpotf2 order). Upper triangle left untouched. */
int orc_cho_factor(int n, const double *C, double *L) {
    if (C != L) memcpy(L, C, sizeof(double) * (size_t)n * n);
    for (int j = 0; j < n; ++j) {
        double ajj = L[j + (size_t)j * n];
        for (int k = 0; k < j; ++k) ajj -= L[j + (size_t)k * n] * L[j + (size_t)k * n];
        if (!(ajj > 0.0) || isnan(ajj)) return j + 1;
        ajj = sqrt(ajj);
        L[j + (size_t)j * n] = ajj;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i + (size_t)j * n];
            for (int k = 0; k < j; ++k) s -= L[i + (size_t)k * n] * L[j + (size_t)k * n];
            L[i + (size_t)j * n] = s / ajj;
        }
    }
    return 0;
}

/* linalg_c.pyx:96-136 cho_solve_vec -> dpotrs: solve L y = b, then L^T x = y. */
void orc_cho_solve_vec(int n, const double *L, const double *b, double *x) {
    if (x != b) memcpy(x, b, sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i + (size_t)k * n] * x[k];
        x[i] = s / L[i + (size_t)i * n];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= L[k + (size_t)i * n] * x[k];
        x[i] = s / L[i + (size_t)i * n];
    }
}

/* linalg_c.pyx:182-210 logdet = 2 * sum(log(L_ii)) */
double orc_logdet(int n, const double *L) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += log(L[i + (size_t)i * n]);
    return 2 * s;
}

static double orc_dot(int n, const double *x, const double *y) { /* linalg_c.pyx:213 dot11 */
    double s = 0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* ---------------------------------------------------------------- gauss_c.pyx */

/* gauss_c.pyx:20-62 mvn_logpdf, general d (d = 1 or 2 on this path). */
double orc_mvn_logpdf(int d, const double *x, const double *m, const double *L, double logdet) {
    double diff[2], buf[2];
    double c = log(2 * M_PI) * d + logdet;
    for (int i = 0; i < d; ++i) diff[i] = x[i] - m[i];
    orc_cho_solve_vec(d, L, diff, buf);
    return -0.5 * (c + orc_dot(d, diff, buf));
}

/* gauss_c.pyx:65-92 int_exp_norm */
double orc_int_exp_norm(double c, double m, double S) {
    double out = (c * m) + (0.5 * (c * c) * S);
    if (out > orc_max_exponent()) return INFINITY;
    return exp(out);
}

/* gauss_c.pyx:95-164 int_K, d = 1: out_i = h^2 N(x_i | mu, w^2 + cov) */
void orc_int_K(int n, const double *x, double h, double w, double mu, double cov, double *out) {
    double W = cov + w * w, L, logdet;
    orc_cho_factor(1, &W, &L);
    logdet = orc_logdet(1, &L);
    for (int i = 0; i < n; ++i) out[i] = (h * h) * exp(orc_mvn_logpdf(1, &x[i], &mu, &L, logdet));
}

/* gauss_c.pyx:235-339 int_K1_K2, d = 1: out[i + j*n1] (n1 x n2, column-major) */
void orc_int_K1_K2(int n1, const double *x1, int n2, const double *x2, double h1, double w1,
                   double h2, double w2, double mu, double cov, double *out) {
    double C[4], m[2] = {mu, mu}, logdet;
    C[0] = w1 * w1 + cov; C[3] = w2 * w2 + cov; C[1] = cov; C[2] = cov;
    orc_cho_factor(2, C, C);
    logdet = orc_logdet(2, C);
    double hh = (h1 * h1) * (h2 * h2);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            double x[2] = {x1[i], x2[j]};
            out[i + (size_t)j * n1] = hh * exp(orc_mvn_logpdf(2, x, m, C, logdet));
        }
}

/* gauss_c.pyx:416-531 int_int_K1_K2_K1, d = 1 (the code, not the stale docstring: no |Gamma|^-1) */
void orc_int_int_K1_K2_K1(int n, const double *x, double h1, double w1, double h2, double w2,
                          double mu, double cov, double *out, double *work /* 2n */) {
    double W1 = cov + w1 * w1, L, logdet, G, A, C, L2, logdet2;
    double *B = work, *N1 = work + n;
    orc_cho_factor(1, &W1, &L);
    logdet = orc_logdet(1, &L);
    orc_cho_solve_vec(1, &L, &cov, &G);      /* cho_solve_mat(L, cov) :499 */
    A = cov * G;                             /* dot22 :500 */
    for (int i = 0; i < n; ++i) {
        double buf;
        orc_cho_solve_vec(1, &L, &x[i], &buf);
        B[i] = cov * buf;                    /* :503-505 */
        N1[i] = orc_mvn_logpdf(1, &x[i], &mu, &L, logdet);
    }
    C = w2 * w2 + 2 * cov - 2 * A;           /* :515 */
    orc_cho_factor(1, &C, &L2);
    logdet2 = orc_logdet(1, &L2);
    double hh = (h1 * h1 * h1 * h1) * (h2 * h2);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double N2 = orc_mvn_logpdf(1, &B[i], &B[j], &L2, logdet2);
            out[i + (size_t)j * n] = hh * exp(N1[i] + N1[j] + N2);
        }
}

/* gauss_c.pyx:617-713 int_int_K1_K2, d = 1 (not called by BQ; parity check only) */
void orc_int_int_K1_K2(int n, const double *x, double h1, double w1, double h2, double w2,
                       double mu, double cov, double *out) {
    double W = 2 * cov + w1 * w1, L, z = 0, N, buf, C, L2, logdet;
    orc_cho_factor(1, &W, &L);
    N = orc_mvn_logpdf(1, &z, &z, &L, orc_logdet(1, &L));
    orc_cho_solve_vec(1, &L, &cov, &buf);
    C = w2 * w2 + cov - cov * buf;
    orc_cho_factor(1, &C, &L2);
    logdet = orc_logdet(1, &L2);
    double hh = (h1 * h1) * (h2 * h2);
    for (int i = 0; i < n; ++i) out[i] = hh * exp(N + orc_mvn_logpdf(1, &x[i], &mu, &L2, logdet));
}

/* gauss_c.pyx:796-855 int_int_K, d = 1 (not called by BQ; parity check only) */
double orc_int_int_K(double h, double w, double mu, double cov) {
    double W = 2 * cov + w * w, L, z = 0;
    (void)mu;
    orc_cho_factor(1, &W, &L);
    return (h * h) * exp(orc_mvn_logpdf(1, &z, &z, &L, orc_logdet(1, &L)));
}

/* ---------------------------------------------------------------- gp stand-in (A.2) */

/* GaussianKernel.__call__: K_ij = h^2 / (sqrt(2 pi) w) * exp(-0.5 (x1_i - x2_j)^2 / w^2),
 * out is n1 x n2 column-major. */
void orc_gaussian_kernel(double h, double w, int n1, const double *x1, int n2, const double *x2,
                         double *out) {
    double c = (h * h) / (sqrt(2 * M_PI) * w);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            double d = x1[i] - x2[j];
            out[i + (size_t)j * n1] = c * exp(-0.5 * (d * d) / (w * w));
        }
}

typedef struct {
    int n;
    double h, w, s;
    double *x, *y;
    double *Kxx, *Lxx, *inv_Lxx, *inv_Kxx, *inv_Kxx_y; /* memoised properties */
    int chol_info;
} orc_gp;

static void orc_gp_free(orc_gp *g) {
    if (!g) return;
    free(g->x); free(g->y); free(g->Kxx); free(g->Lxx); free(g->inv_Lxx); free(g->inv_Kxx);
    free(g->inv_Kxx_y); free(g);
}

/* GP(K, x, y, s): Kxx = K(x,x) + s^2 I; Lxx = cholesky; inv_Lxx = inv(Lxx);
 * inv_Kxx = inv_Lxx^T inv_Lxx; inv_Kxx_y = inv_Kxx y. */
static orc_gp *orc_gp_new(double h, double w, double s, int n, const double *x, const double *y) {
    orc_gp *g = (orc_gp *)calloc(1, sizeof(orc_gp));
    size_t nn = (size_t)n * n;
    g->n = n; g->h = h; g->w = w; g->s = s;
    g->x = (double *)malloc(sizeof(double) * n); memcpy(g->x, x, sizeof(double) * n);
    g->y = (double *)malloc(sizeof(double) * n); memcpy(g->y, y, sizeof(double) * n);
    g->Kxx = (double *)malloc(sizeof(double) * nn);
    g->Lxx = (double *)calloc(nn, sizeof(double));
    g->inv_Lxx = (double *)calloc(nn, sizeof(double));
    g->inv_Kxx = (double *)calloc(nn, sizeof(double));
    g->inv_Kxx_y = (double *)calloc(n, sizeof(double));
    orc_gaussian_kernel(h, w, n, x, n, x, g->Kxx);
    for (int i = 0; i < n; ++i) g->Kxx[i + (size_t)i * n] += s * s;
    memcpy(g->Lxx, g->Kxx, sizeof(double) * nn);
    g->chol_info = orc_cho_factor(n, g->Lxx, g->Lxx);
    for (int j = 0; j < n; ++j)      /* numpy.linalg.cholesky returns a clean lower triangle */
        for (int i = 0; i < j; ++i) g->Lxx[i + (size_t)j * n] = 0.0;
    if (g->chol_info) return g;
    /* inv(L): column j solves L z = e_j */
    for (int j = 0; j < n; ++j) {
        double *z = g->inv_Lxx + (size_t)j * n;
        for (int i = j; i < n; ++i) {
            double sacc = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) sacc -= g->Lxx[i + (size_t)k * n] * z[k];
            z[i] = sacc / g->Lxx[i + (size_t)i * n];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double sacc = 0;
            int k0 = i > j ? i : j;
            for (int k = k0; k < n; ++k) sacc += g->inv_Lxx[k + (size_t)i * n] * g->inv_Lxx[k + (size_t)j * n];
            g->inv_Kxx[i + (size_t)j * n] = sacc;
        }
    for (int i = 0; i < n; ++i) {
        double sacc = 0;
        for (int j = 0; j < n; ++j) sacc += g->inv_Kxx[i + (size_t)j * n] * y[j];
        g->inv_Kxx_y[i] = sacc;
    }
    return g;
}

/* GP.mean(xo) = Kxox inv_Kxx_y ; GP.cov(xo) = Kxoxo - Kxox inv_Kxx Kxxo, single xo */
static void orc_gp_mean_cov(const orc_gp *g, double xo, double *kbuf /* n */, double *mean, double *cov) {
    int n = g->n;
    double c = (g->h * g->h) / (sqrt(2 * M_PI) * g->w);
    for (int i = 0; i < n; ++i) {
        double d = xo - g->x[i];
        kbuf[i] = c * exp(-0.5 * (d * d) / (g->w * g->w));
    }
    if (mean) *mean = orc_dot(n, kbuf, g->inv_Kxx_y);
    if (cov) {
        double q = 0;
        for (int i = 0; i < n; ++i) {
            double t = 0;
            for (int j = 0; j < n; ++j) t += g->inv_Kxx[i + (size_t)j * n] * kbuf[j];
            q += kbuf[i] * t;
        }
        *cov = c - q;   /* Kxoxo(xo) = c exp(0) */
    }
}

/* GP.log_lh = -1/2 y^T inv_Kxx_y - sum(log diag(Lxx)) - n/2 log(2 pi) */
static double orc_gp_log_lh(const orc_gp *g) {
    if (g->chol_info) return -INFINITY;
    double s = 0;
    for (int i = 0; i < g->n; ++i) s += log(g->Lxx[i + (size_t)i * g->n]);
    return -0.5 * orc_dot(g->n, g->y, g->inv_Kxx_y) - s - 0.5 * g->n * log(2 * M_PI);
}

/* ---------------------------------------------------------------- bq_c.pyx */

/* bq_c.pyx:157-213 Z_mean */
double orc_Z_mean(int n, const double *x_sc, const double *alpha_l, double h_l, double w_l,
                  double mu, double cov, double *work /* n */) {
    orc_int_K(n, x_sc, h_l, w_l, mu, cov, work);
    return orc_dot(n, work, alpha_l);
}

/* bq_c.pyx:264-355 Z_var */
double orc_Z_var(int ns, const double *x_s, int n, const double *x_sc, const double *alpha_l,
                 const double *L_tl, double h_l, double w_l, double h_tl, double w_tl, double mu,
                 double cov) {
    double *M = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *M2 = (double *)malloc(sizeof(double) * (size_t)ns * n);
    double *work = (double *)malloc(sizeof(double) * (2 * (size_t)n + 2 * (size_t)ns));
    double *alpha_int = work, *beta = work + n, *L_tl_beta = beta + ns;
    double *w2 = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    orc_int_int_K1_K2_K1(n, x_sc, h_l, w_l, h_tl, w_tl, mu, cov, M, w2);
    for (int j = 0; j < n; ++j) {          /* dot12(alpha_l, M, alpha_int) :343 */
        double s = 0;
        for (int i = 0; i < n; ++i) s += alpha_l[i] * M[i + (size_t)j * n];
        alpha_int[j] = s;
    }
    double alpha_int_alpha = orc_dot(n, alpha_int, alpha_l);
    orc_int_K1_K2(ns, x_s, n, x_sc, h_tl, w_tl, h_l, w_l, mu, cov, M2);
    for (int i = 0; i < ns; ++i) {         /* dot21(M2, alpha_l, beta) :347 */
        double s = 0;
        for (int j = 0; j < n; ++j) s += M2[i + (size_t)j * ns] * alpha_l[j];
        beta[i] = s;
    }
    orc_cho_solve_vec(ns, L_tl, beta, L_tl_beta);
    double beta2 = orc_dot(ns, beta, L_tl_beta);
    free(M); free(M2); free(work); free(w2);
    return alpha_int_alpha - beta2;
}

/* bq_c.pyx:425-490 _esm_and_em */
static void orc_esm_core(int nca, const double *int_K_l, const double *l_sc, const double *L_l,
                         double tm_a, double tC_a, double *A_sca, double *esm, double *em) {
    orc_cho_solve_vec(nca, L_l, int_K_l, A_sca);
    double A_a = A_sca[nca - 1];
    double A_sc_l = orc_dot(nca - 1, A_sca, l_sc);
    double e1 = orc_int_exp_norm(1, tm_a, tC_a);
    if (e1 == INFINITY) { *esm = INFINITY; *em = INFINITY; return; }
    double E_m = A_sc_l + A_a * e1;
    double e2 = orc_int_exp_norm(2, tm_a, tC_a);
    if (e2 == INFINITY) { *esm = INFINITY; *em = E_m; return; }
    *esm = (A_sc_l * A_sc_l) + (2 * A_sc_l * A_a * e1) + ((A_a * A_a) * e2);
    *em = E_m;
}

/* ---------------------------------------------------------------- bq.py (the BQ object) */

typedef struct {
    int ns, nc, nsc;
    double *x_s, *l_s, *tl_s, *x_c, *l_c, *x_sc, *l_sc;
    double mu, cov, thresh;
    orc_gp *gp_log_l, *gp_l;
    double Zm, Zv;
    int error; /* 1: gp_log_l not PD, 2: gp_l not PD, 3: "GP mean is too large" (bq.py:945-947) */
} orc_model;

void orc_model_free(orc_model *m) {
    if (!m) return;
    free(m->x_s); free(m->l_s); free(m->tl_s); free(m->x_c); free(m->l_c); free(m->x_sc); free(m->l_sc);
    orc_gp_free(m->gp_log_l); orc_gp_free(m->gp_l); free(m);
}

/* BQ.__init__ (bq.py:55-92) + init (bq.py:132-171) with x_c given (the draw at bq.py:978 uses
 * the host RNG) + _choose_candidates' l_c = exp(gp_log_l.mean(x_c)) (bq.py:985) which is also
 * what _set_gp_log_l_params recomputes (bq.py:942-950).  params = (h, w, s). */
orc_model *orc_model_new(int ns, const double *x_s, const double *l_s, int nc, const double *x_c,
                         const double *params_tl, const double *params_l, double mu, double cov,
                         double thresh, int check_max) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    int n = ns + nc;
    m->ns = ns; m->nc = nc; m->nsc = n; m->mu = mu; m->cov = cov; m->thresh = thresh;
    m->x_s = (double *)malloc(sizeof(double) * ns); memcpy(m->x_s, x_s, sizeof(double) * ns);
    m->l_s = (double *)malloc(sizeof(double) * ns); memcpy(m->l_s, l_s, sizeof(double) * ns);
    m->tl_s = (double *)malloc(sizeof(double) * ns);
    for (int i = 0; i < ns; ++i) m->tl_s[i] = log(l_s[i]);                 /* bq.py:73 */
    m->x_c = (double *)malloc(sizeof(double) * (nc + 1)); memcpy(m->x_c, x_c, sizeof(double) * nc);
    m->l_c = (double *)malloc(sizeof(double) * (nc + 1));
    m->x_sc = (double *)malloc(sizeof(double) * n);
    m->l_sc = (double *)malloc(sizeof(double) * n);
    m->gp_log_l = orc_gp_new(params_tl[0], params_tl[1], params_tl[2], ns, m->x_s, m->tl_s);
    if (m->gp_log_l->chol_info) { m->error = 1; return m; }
    double *kbuf = (double *)malloc(sizeof(double) * (n + 1));
    for (int j = 0; j < nc; ++j) {
        double mean, var;
        orc_gp_mean_cov(m->gp_log_l, x_c[j], kbuf, &mean, &var);
        if (check_max) {                                                    /* bq.py:942-947 */
            if (var < 0) var = 0;
            if (mean + 2 * sqrt(var) > orc_max_exponent()) m->error = 3;
        }
        m->l_c[j] = exp(mean);
    }
    memcpy(m->x_sc, x_s, sizeof(double) * ns); memcpy(m->x_sc + ns, x_c, sizeof(double) * nc);
    memcpy(m->l_sc, l_s, sizeof(double) * ns); memcpy(m->l_sc + ns, m->l_c, sizeof(double) * nc);
    free(kbuf);
    if (m->error) return m;
    m->gp_l = orc_gp_new(params_l[0], params_l[1], params_l[2], n, m->x_sc, m->l_sc);
    if (m->gp_l->chol_info) { m->error = 2; return m; }
    double *work = (double *)malloc(sizeof(double) * n);
    m->Zm = orc_Z_mean(n, m->x_sc, m->gp_l->inv_Kxx_y, m->gp_l->h, m->gp_l->w, mu, cov, work); /* bq.py:268-291 */
    m->Zv = orc_Z_var(ns, m->x_s, n, m->x_sc, m->gp_l->inv_Kxx_y, m->gp_log_l->Lxx, m->gp_l->h,
                      m->gp_l->w, m->gp_log_l->h, m->gp_log_l->w, mu, cov);                   /* bq.py:329-348 */
    free(work);
    return m;
}

int orc_model_error(const orc_model *m) { return m->error; }
double orc_model_Z_mean(const orc_model *m) { return m->Zm; }
double orc_model_Z_var(const orc_model *m) { return m->Zv; }
void orc_model_l_c(const orc_model *m, double *out) { memcpy(out, m->l_c, sizeof(double) * m->nc); }
void orc_model_alpha_l(const orc_model *m, double *out) { memcpy(out, m->gp_l->inv_Kxx_y, sizeof(double) * m->nsc); }
double orc_model_log_lh(const orc_model *m) { return orc_gp_log_lh(m->gp_log_l) + orc_gp_log_lh(m->gp_l); } /* bq.py:546 */

/* bq_c.pyx:127-140 improve_covariance_conditioning on a column-major square matrix */
static void orc_improve_conditioning(int n, double *M, const int *idx, int nidx) {
    double mx = -INFINITY;
    for (size_t i = 0; i < (size_t)n * n; ++i) if (M[i] > mx) mx = M[i];   /* np.max(M) */
    double j = fmax(ORC_EPS, mx) * 1e-4;
    for (int i = 0; i < nidx; ++i) M[idx[i] + (size_t)idx[i] * n] += j;
}

/* bq.py:447-527 _esm_and_em for ONE query point.  work: caller scratch of
 * 2*(n+1)^2 + 4*(n+1) doubles and (nc+1) ints (allocated in the batch entry below). */
static int orc_esm_point(const orc_model *m, double x_a, double *esm, double *em, double *K, double *L,
                         double *vec, int *idx) {
    const int ns = m->ns, nc = m->nc, n = m->nsc, n1 = n + 1;
    if (isnan(x_a) || isinf(x_a)) { *esm = NAN; *em = NAN; return ORC_XA_BAD; }     /* :451 */
    for (int i = 0; i < ns; ++i)                                                      /* :456 np.isclose(x_a, x_s, atol=1e-4) */
        if (fabs(x_a - m->x_s[i]) <= 1e-4 + 1e-5 * fabs(m->x_s[i])) {
            *em = m->Zm; *esm = m->Zm * m->Zm; return ORC_SHORTCUT;
        }
    double *x_sca = vec, *int_K = vec + n1, *A = vec + 2 * n1, *kbuf = vec + 3 * n1;
    memcpy(x_sca, m->x_sc, sizeof(double) * n); x_sca[n] = x_a;                       /* :462 */
    orc_gaussian_kernel(m->gp_l->h, m->gp_l->w, n1, x_sca, n1, x_sca, K);              /* :465 Kxoxo: no s^2 */
    int nidx = 0;
    for (int j = 0; j < nc; ++j) if (fabs(m->x_c[j] - x_a) < m->thresh) idx[nidx++] = ns + j; /* :470 */
    if (nidx) orc_improve_conditioning(n1, K, idx, nidx);                              /* :471-473 */
    idx[0] = n; orc_improve_conditioning(n1, K, idx, 1);                               /* :476 */
    if (orc_cho_factor(n1, K, L) > 0) {                                                /* :478-490 */
        *em = m->Zm; *esm = m->Zm * m->Zm; return ORC_NOTPD;
    }
    double tm_a, tC_a;
    orc_gp_mean_cov(m->gp_log_l, x_a, kbuf, &tm_a, &tC_a);                             /* :493-496 */
    orc_int_K(n1, x_sca, m->gp_l->h, m->gp_l->w, m->mu, m->cov, int_K);                /* bq_c.pyx:531 */
    orc_esm_core(n1, int_K, m->l_sc, L, tm_a, tC_a, A, esm, em);                       /* bq_c.pyx:534 */
    int st = ORC_OK;
    if (isnan(*esm) || *esm < 0) st |= ORC_ESM_BAD;                                    /* :514 */
    if (isnan(*em)) st |= ORC_EM_BAD;                                                  /* :518 */
    if (isinf(*esm)) st |= ORC_ESM_INF;                                                /* :522 */
    if (isinf(*em)) st |= ORC_EM_INF;                                                  /* :524 */
    return st;
}

/* bq.py:425-445 expected_squared_mean_and_mean over a vector of query points.
 * status may be NULL.  Returns the OR of all statuses. */
int orc_esm_and_em(const orc_model *m, int na, const double *x_a, double *esm, double *em, int *status) {
    const int n1 = m->nsc + 1;
    double *K = (double *)malloc(sizeof(double) * (size_t)n1 * n1);
    double *L = (double *)malloc(sizeof(double) * (size_t)n1 * n1);
    double *vec = (double *)malloc(sizeof(double) * 4 * (size_t)n1);
    int *idx = (int *)malloc(sizeof(int) * (m->nc + 1));
    int all = 0;
    for (int i = 0; i < na; ++i) {
        int st = orc_esm_point(m, x_a[i], &esm[i], &em[i], K, L, vec, idx);
        if (status) status[i] = st;
        all |= st;
    }
    free(K); free(L); free(vec); free(idx);
    return all;
}

/* bq.py:374-377 expected_Z_var = Z_mean^2 + Z_var - esm */
void orc_expected_Z_var(const orc_model *m, int na, const double *esm, double *out) {
    double msm = m->Zm * m->Zm + m->Zv;
    for (int i = 0; i < na; ++i) out[i] = msm - esm[i];
}

/* GP predictive mean / variance of the two GPs (bq.py:177-231), used by the fixture checks. */
void orc_model_gp_log_l_mean_cov(const orc_model *m, int na, const double *x, double *mean, double *cov) {
    double *kbuf = (double *)malloc(sizeof(double) * m->ns);
    for (int i = 0; i < na; ++i) orc_gp_mean_cov(m->gp_log_l, x[i], kbuf, &mean[i], &cov[i]);
    free(kbuf);
}
