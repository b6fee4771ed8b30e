/* TEST INFRASTRUCTURE -- CPU oracle, not product code.
 *
 * Plain-C restatement of the SympGPR hot path of the reference
 * (python/05_tokamak/SympGPR/sympgpr.f90 + kernels*.f90 + minpack.f90 hybrd1
 * with n = 1 + fieldlines.f90 compute_r).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product library (sympgpr_b200/libsympgpr_b200.so) never links or calls it.
 *
 * Differences from the reference that do not change results beyond rounding:
 *   - alpha = Kyinv * ztrain is hoisted out of guessP / target / calcq
 *     (sympgpr.f90:72,85,121 recompute the same matvec on every call); entry
 *     points taking Kyinv do the matvec once per call, entry points ending in
 *     _alpha take the vector directly.
 *   - ensemble loops may run under OpenMP over orbits (orbits are independent,
 *     python/functions/func.py:227-236).
 *
 * Pinned by tests/test_oracle.py against tests/golden/ (reference SymPy
 * derivation + reference pure-Python GP layer) and against SciPy's MINPACK
 * hybrd (scipy.optimize.fsolve) for the root solve.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FAM_PRODUCT 0
#define FAM_SQ 1
#define FAM_SUM 2
#define FAM_PERIOD 3

#define MAP_PENDULUM 0
#define MAP_HENON 1
#define MAP_STANDARD 2
#define MAP_TOKAMAK 3

static double sqr(double v) { return v * v; }

/* ---- scalar forms, argument order (x_a, y_a, x_b, y_b, lx, ly[, p]) ------ */
/* product: kernels.f90:1-11,58-94 ; period: implicit_period_unknown/kernels.f90
 * :1-13,64-103 (product == period with p = 0.5 up to the literal constants). */
static double kern_f(int fam, double xa, double ya, double xb, double yb, double lx, double ly, double p)
{
    switch (fam) {
    case FAM_PRODUCT: {
        double h = 0.5 * xa - 0.5 * xb;
        return exp(-0.5 * sqr(ya - yb) / sqr(ly) - 0.5 * sqr(sin(h)) / sqr(lx));
    }
    case FAM_PERIOD:
        return exp(-0.5 * sqr(ya - yb) / sqr(ly) - 0.5 * sqr(sin(p * (xa - xb))) / sqr(lx));
    case FAM_SQ: /* kernels_sq.f90:1-10 */
        return exp(-0.5 * sqr(ya - yb) / sqr(ly) - 0.5 * sqr(xa - xb) / sqr(lx));
    default: { /* kernels_expl_per_q_sq_p.f90:1-10 */
        double h = 0.5 * xa - 0.5 * xb;
        return exp((-0.5 * ya * ya + ya * yb - 0.5 * yb * yb) / sqr(ly)) + exp(-0.5 * sqr(sin(h)) / sqr(lx));
    }
    }
}

/* the three Hessian blocks at once: out[0]=d2kdxdx0, out[1]=d2kdxdy0, out[2]=d2kdydy0 */
static void hess_f(int fam, double xa, double ya, double xb, double yb, double lx, double ly, double p, double *o)
{
    double dy = ya - yb;
    switch (fam) {
    case FAM_PRODUCT: { /* kernels.f90:58-94 */
        double h = 0.5 * xa - 0.5 * xb, s = sin(h), c = cos(h);
        double E = exp(-0.5 * (sqr(lx) * sqr(dy) + sqr(ly) * sqr(s)) / (sqr(lx) * sqr(ly)));
        o[0] = 0.25 * (sqr(lx) * cos(1.0 * xa - 1.0 * xb) - sqr(s) * sqr(c)) * E / (sqr(lx) * sqr(lx));
        o[1] = -0.5 * dy * E * s * c / (sqr(lx) * sqr(ly));
        o[2] = 1.0 * (sqr(ly) - sqr(dy)) * E / (sqr(ly) * sqr(ly));
        return;
    }
    case FAM_PERIOD: { /* implicit_period_unknown/kernels.f90:64-103 */
        double u = p * (xa - xb), s = sin(u), c = cos(u);
        double E = exp(-0.5 * (sqr(lx) * sqr(dy) + sqr(ly) * sqr(s)) / (sqr(lx) * sqr(ly)));
        o[0] = 1.0 * sqr(p) * (sqr(lx) * cos(2.0 * p * (xa - xb)) - sqr(s) * sqr(c)) * E / (sqr(lx) * sqr(lx));
        o[1] = -1.0 * p * dy * E * s * c / (sqr(lx) * sqr(ly));
        o[2] = 1.0 * (sqr(ly) - sqr(dy)) * E / (sqr(ly) * sqr(ly));
        return;
    }
    case FAM_SQ: { /* kernels_sq.f90:56-87 */
        double dx = xa - xb;
        double E = exp(-0.5 * (sqr(lx) * sqr(dy) + sqr(ly) * sqr(dx)) / (sqr(lx) * sqr(ly)));
        o[0] = 1.0 * (sqr(lx) - sqr(dx)) * E / (sqr(lx) * sqr(lx));
        o[1] = -1.0 * dx * dy * E / (sqr(lx) * sqr(ly));
        o[2] = 1.0 * (sqr(ly) - sqr(dy)) * E / (sqr(ly) * sqr(ly));
        return;
    }
    default: { /* kernels_expl_per_q_sq_p.f90:56-88 */
        double h = 0.5 * xa - 0.5 * xb;
        double Ex = exp(-0.5 * sqr(sin(h)) / sqr(lx));
        double Ey = exp(0.5 * (-ya * ya + 2.0 * ya * yb - yb * yb) / sqr(ly));
        o[0] = ((1.0 / 4.0) * sqr(lx) * cos(xa - xb) - 1.0 / 16.0 * sqr(sin(xa - xb))) * Ex / (sqr(lx) * sqr(lx));
        o[1] = 0.0;
        o[2] = (sqr(ly) - sqr(dy)) * Ey / (sqr(ly) * sqr(ly));
        return;
    }
    }
}

/* ---- fills ---------------------------------------------------------------- */
/* sympgpr.f90:12-38.  K is column-major with leading dimension ldk, rows x
 * cols; N = rows/2, N0 = cols/2 (integer division); the final K = hyp(3)*K
 * scales the whole array including rows/cols the loop never wrote. */
void oracle_build_k(int fam, double p, const double *x, const double *y, const double *x0, const double *y0,
                    const double *hyp, double *K, long rows, long cols, long ldk)
{
    long N = rows / 2, N0 = cols / 2;
#pragma omp parallel for schedule(static)
    for (long j = 0; j < N0; j++) {
        for (long i = 0; i < N; i++) {
            double o[3];
            hess_f(fam, x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], p, o);
            K[i + j * ldk] = o[0];
            K[N + i + j * ldk] = o[1];
            K[i + (N0 + j) * ldk] = o[1];
            K[N + i + (N0 + j) * ldk] = o[2];
        }
    }
#pragma omp parallel for schedule(static)
    for (long j = 0; j < cols; j++)
        for (long i = 0; i < rows; i++)
            K[i + j * ldk] = hyp[2] * K[i + j * ldk];
}

/* sympgpr.f90:40-60 */
void oracle_buildkreg(int fam, double p, const double *x, const double *y, const double *x0, const double *y0,
                      const double *hyp, double *K, long rows, long cols, long ldk)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < cols; j++)
        for (long i = 0; i < rows; i++)
            K[i + j * ldk] = kern_f(fam, x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], p);
#pragma omp parallel for schedule(static)
    for (long j = 0; j < cols; j++)
        for (long i = 0; i < rows; i++)
            K[i + j * ldk] = hyp[2] * K[i + j * ldk];
}

/* ---- prediction ----------------------------------------------------------- */
typedef struct {
    int fam;
    double p;
    const double *hyp, *hypp;
    long np, nt;
    const double *xtp, *ytp, *alphap;     /* ordinary GP: Np points, alphap(Np)   */
    const double *xt, *yt, *alpha;        /* symplectic GP: Nt points, alpha(2Nt) */
    int start_delta;                      /* 1: hybrd1 starts at p + guess (guess GP trained on P - p) */
    int reverse_sum;                      /* 1: training-set sums run from the last point to the first
                                           * (rounding-sensitivity probe of the tests, not a reference mode) */
} model_t;

/* sympgpr.f90:62-73 with alphap = Kyinvp*ztrainp hoisted */
static double guessp_m(const model_t *m, double q, double pp)
{
    double acc = 0.0;
    for (long jj = 0; jj < m->np; jj++) {
        long j = m->reverse_sum ? m->np - 1 - jj : jj;
        acc += m->hypp[2] * kern_f(m->fam, m->xtp[j], m->ytp[j], q, pp, m->hypp[0], m->hypp[1], m->p) * m->alphap[j];
    }
    return acc;
}

/* rows 1 and 2 of Kstar(2,2Nt) dotted with alpha: sympgpr.f90:85,121 */
static void kstar_dot(const model_t *m, double q, double P, double *row1, double *row2)
{
    double r1a = 0.0, r1b = 0.0, r2a = 0.0, r2b = 0.0;
    long nt = m->nt;
    for (long jj = 0; jj < nt; jj++) {
        long j = m->reverse_sum ? nt - 1 - jj : jj;
        double o[3];
        hess_f(m->fam, m->xt[j], m->yt[j], q, P, m->hyp[0], m->hyp[1], m->p, o);
        r1a += m->hyp[2] * o[0] * m->alpha[j];
        r1b += m->hyp[2] * o[1] * m->alpha[nt + j];
        r2a += m->hyp[2] * o[1] * m->alpha[j];
        r2b += m->hyp[2] * o[2] * m->alpha[nt + j];
    }
    *row1 = r1a + r1b;
    *row2 = r2a + r2b;
}

typedef struct { const model_t *m; double q, p; } target_ctx;

/* sympgpr.f90:112-124 */
static double target_f(const target_ctx *c, double P)
{
    double r1, r2;
    kstar_dot(c->m, c->q, P, &r1, &r2);
    return r1 - c->p + P;
}

/* minpack.f90:1460-1594 (hybrd1) -> :911-1458 (hybrd) specialised to n = 1:
 * mode = 2, diag = 1, factor = 100, epsfcn = 0, ml = mu = 0, maxfev = 400.
 * With n = 1: qrfac gives R = -J and Q = -1 (Householder), qtf = -f;
 * dogleg (:224-432), r1updt (:5299) and r1mpyq (:5188) reduce to the scalar
 * statements below. */
static double hybrd1_n1(const target_ctx *c, double x, double tol, int *info_out, int *nfev_out)
{
    const double epsmch = DBL_EPSILON, factor = 100.0, xtol = tol, diag = 1.0;
    const int maxfev = 400;
    int info = 0, nfev, iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0, jeval;
    double fvec, fnorm, fjac, r, qtf, delta = 0.0, xnorm = 0.0;
    if (tol < 0.0) { *info_out = 0; *nfev_out = 0; return x; }
    fvec = target_f(c, x);
    nfev = 1;
    fnorm = fabs(fvec);
    for (;;) { /* outer loop (label 30) */
        jeval = 1;
        { /* fdjac1 :613-788 */
            double eps = sqrt(epsmch), h = eps * fabs(x), wa;
            if (h == 0.0) h = eps;
            wa = target_f(c, x + h);
            fjac = (wa - fvec) / h;
            nfev += 1;
        }
        { /* qrfac :4773 + forming qtf and r, qform :4672 */
            double ajnorm = fabs(fjac), hh = fjac;
            if (ajnorm != 0.0) {
                if (hh < 0.0) ajnorm = -ajnorm;
                hh = hh / ajnorm;
                hh = hh + 1.0;
            }
            r = -ajnorm;                       /* rdiag */
            if (iter == 1) {
                xnorm = fabs(diag * x);
                delta = factor * xnorm;
                if (delta == 0.0) delta = factor;
            }
            qtf = fvec;
            if (hh != 0.0) {
                double temp = -(qtf * hh) / hh;
                qtf = qtf + hh * temp;
            }
            /* qform: q = 1 - (hh*1/hh)*hh if hh != 0 */
            fjac = (hh != 0.0) ? 1.0 - ((hh * 1.0) / hh) * hh : 1.0;
        }
        for (;;) { /* inner loop (label 180) */
            double wa1, wa2, wa3, wa4, pnorm, fnorm1, actred, prered, ratio, temp;
            { /* dogleg */
                double xx, qnorm;
                temp = r;
                if (temp == 0.0) {
                    temp = fabs(r);
                    temp = (temp == 0.0) ? epsmch : epsmch * temp;
                }
                xx = qtf / temp;
                qnorm = fabs(diag * xx);
                if (qnorm > delta) {
                    double g = (r * qtf) / diag, gnorm = fabs(g), sgnorm = 0.0, alpha = delta / qnorm;
                    if (gnorm != 0.0) {
                        double t2;
                        g = (g / gnorm) / diag;
                        t2 = fabs(r * g);
                        sgnorm = (gnorm / t2) / t2;
                        alpha = 0.0;
                        if (sgnorm < delta) {
                            double bnorm = fabs(qtf);
                            double t = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                            t = t - (delta / qnorm) * sqr(sgnorm / delta)
                                + sqrt(sqr(t - (delta / qnorm))
                                       + (1.0 - sqr(delta / qnorm)) * (1.0 - sqr(sgnorm / delta)));
                            alpha = ((delta / qnorm) * (1.0 - sqr(sgnorm / delta))) / t;
                        }
                    }
                    temp = (1.0 - alpha) * fmin(sgnorm, delta);
                    xx = temp * g + alpha * xx;
                }
                wa1 = xx;
            }
            wa1 = -wa1;
            wa2 = x + wa1;
            wa3 = diag * wa1;
            pnorm = fabs(wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            wa4 = target_f(c, wa2);
            nfev += 1;
            fnorm1 = fabs(wa4);
            actred = -1.0;
            if (fnorm1 < fnorm) actred = 1.0 - sqr(fnorm1 / fnorm);
            wa3 = qtf + r * wa1;
            temp = fabs(wa3);
            prered = 0.0;
            if (temp < fnorm) prered = 1.0 - sqr(temp / fnorm);
            ratio = 0.0;
            if (0.0 < prered) ratio = actred / prered;
            if (ratio < 0.1) {
                ncsuc = 0;
                ncfail += 1;
                delta = 0.5 * delta;
            } else {
                ncfail = 0;
                ncsuc += 1;
                if (0.5 <= ratio || 1 < ncsuc) delta = fmax(delta, pnorm / 0.5);
                if (fabs(ratio - 1.0) <= 0.1) delta = pnorm / 0.5;
            }
            if (0.0001 <= ratio) {
                x = wa2;
                wa2 = diag * x;
                fvec = wa4;
                xnorm = fabs(wa2);
                fnorm = fnorm1;
                iter += 1;
            }
            nslow1 += 1;
            if (0.001 <= actred) nslow1 = 0;
            if (jeval) nslow2 += 1;
            if (0.1 <= actred) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0.0) info = 1;
            if (info != 0) goto done;
            if (maxfev <= nfev) info = 2;
            if (0.1 * fmax(0.1 * delta, pnorm) <= epsmch * xnorm) info = 3;
            if (nslow2 == 5) info = 4;
            if (nslow1 == 10) info = 5;
            if (info != 0) goto done;
            if (ncfail == 2) break; /* recompute jacobian */
            { /* rank-one (Broyden) update */
                double sum2 = wa4 * fjac;
                wa2 = (sum2 - wa3) / pnorm;
                wa1 = diag * ((diag * wa1) / pnorm);
                if (0.0001 <= ratio) qtf = sum2;
                r = r + wa2 * wa1; /* r1updt with m = n = 1 */
            }
            jeval = 0;
        }
    }
done:
    if (info == 5) info = 4;
    *info_out = info;
    *nfev_out = nfev;
    return x;
}

/* sympgpr.f90:88-125: guess, one discarded target evaluation, hybrd1(tol=1e-13) */
static double calcp_m(const model_t *m, double q, double p, int *info, int *nfev)
{
    target_ctx c = { m, q, p };
    double pg = guessp_m(m, q, p);
    /* sympgpr.f90:103-107 hands the guess GP's prediction to hybrd1 as it is.  Scripts 03/04/05 train that GP on
     * P - p (python/04_standard_map/main.py:89-90), so the prediction is the difference; start_delta adds p, the
     * consistent start for such a model (same call sequence otherwise). */
    if (m->start_delta) pg = p + pg;
    (void)target_f(&c, pg);
    return hybrd1_n1(&c, pg, 1e-13, info, nfev);
}

static double calcq_m(const model_t *m, double q, double P)
{
    double r1, r2;
    kstar_dot(m, q, P, &r1, &r2);
    return r2;
}

static void matvec_cm(const double *A, long n, const double *v, double *out)
{
    for (long i = 0; i < n; i++) out[i] = 0.0;
    for (long j = 0; j < n; j++)
        for (long i = 0; i < n; i++)
            out[i] += A[i + j * n] * v[j];
}

/* Entry points mirroring the f2py signatures (Kyinv column-major, n x n). */
double oracle_guessp(int fam, double pper, double x, double y, const double *hypp, const double *xtp,
                     const double *ytp, const double *ztp, const double *kyinvp, long np)
{
    double *a = (double *)malloc(sizeof(double) * np), res;
    model_t m = { fam, pper, NULL, hypp, np, 0, xtp, ytp, a, NULL, NULL, NULL, 0, 0 };
    matvec_cm(kyinvp, np, ztp, a);
    res = guessp_m(&m, x, y);
    free(a);
    return res;
}

double oracle_calcq(int fam, double pper, double x, double y, const double *xt, const double *yt,
                    const double *hyp, const double *kyinv, const double *zt, long nt)
{
    double *a = (double *)malloc(sizeof(double) * 2 * nt), res;
    model_t m = { fam, pper, hyp, NULL, 0, nt, NULL, NULL, NULL, xt, yt, a, 0, 0 };
    matvec_cm(kyinv, 2 * nt, zt, a);
    res = calcq_m(&m, x, y);
    free(a);
    return res;
}

double oracle_calcp_alpha(int fam, double pper, double x, double y, const double *hyp, const double *hypp,
                          const double *xtp, const double *ytp, const double *alphap, long np,
                          const double *xt, const double *yt, const double *alpha, long nt,
                          int *info, int *nfev)
{
    model_t m = { fam, pper, hyp, hypp, np, nt, xtp, ytp, alphap, xt, yt, alpha, 0, 0 };
    return calcp_m(&m, x, y, info, nfev);
}

double oracle_target_alpha(int fam, double pper, double q, double p, double P, const double *hyp,
                           const double *xt, const double *yt, const double *alpha, long nt)
{
    model_t m = { fam, pper, hyp, NULL, 0, nt, NULL, NULL, NULL, xt, yt, alpha, 0, 0 };
    target_ctx c = { &m, q, p };
    return target_f(&c, P);
}

/* fieldlines.f90:94-107 (+ f_r :82-91, Ath :34-39, dAthdr :42-47; B0 = R0 = 1) */
double oracle_compute_r(double pth, double th, double rstart)
{
    const double B0 = 1.0, R0 = 1.0;
    double r = rstart;
    for (int k = 0; k < 20; k++) {
        double yv = pth - B0 * (r * r / 2.0 - r * r * r / (3.0 * R0) * cos(th));
        double dy = -(B0 * (r - r * r / R0 * cos(th)));
        r = r - yv / dy;
    }
    return r;
}

static double np_mod(double a, double b)
{
    /* numpy.mod for b > 0: result has the sign of the divisor */
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}

/* Ensemble loop.  kind selects the post-step variant:
 *   MAP_PENDULUM python/functions/func.py:216-237
 *   MAP_HENON    python/functions/func.py:239-260
 *   MAP_STANDARD python/04_standard_map/func.py:218-254 (also fills pdiff if non-NULL)
 *   MAP_TOKAMAK  python/05_tokamak/SympGPR/func.py:182-211
 * qmap/pmap are (nm, E) C-order (row = step).  Returns total function
 * evaluations of the root solver (for the n_eval statistic).  notconv (optional, E ints) counts per
 * orbit the steps where hybrd1 returned info != 1 (the reference ignores info, sympgpr.f90:107);
 * maxres (optional, E doubles) the largest residual |f(P)| any accepted root of the orbit left. */
/* flags: bit 0 = start hybrd1 at p + guess (see calcp_m), bit 1 = reversed summation order (sensitivity probe).
 * out_every: 1 = full history (nm rows); k > 1 = rows 0, k, 2k, ... ((nm-1)/k + 1 rows); the loop is the same. */
long oracle_applymap_alpha_ex(int kind, int fam, double pper, long nm, long E, const double *q0, const double *p0,
                              const double *hyp, const double *hypp,
                              const double *xtp, const double *ytp, const double *alphap, long np,
                              const double *xt, const double *yt, const double *alpha, long nt,
                              double *qmap, double *pmap, double *pdiff, int *notconv, double *maxres,
                              int flags, long out_every)
{
    const double two_pi = 2.0 * M_PI;
    model_t m = { fam, pper, hyp, hypp, np, nt, xtp, ytp, alphap, xt, yt, alpha, flags & 1, (flags >> 1) & 1 };
    long total_fev = 0;
    if (out_every < 1) out_every = 1;
    for (long k = 0; k < E; k++) {
        pmap[k] = p0[k];
        qmap[k] = q0[k];
        if (pdiff) pdiff[k] = p0[k];
        if (notconv) notconv[k] = 0;
        if (maxres) maxres[k] = 0.0;
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_fev)
    for (long k = 0; k < E; k++) {
        double qc = q0[k], pc = p0[k], pdc = p0[k];
        for (long i = 0; i < nm - 1; i++) {
            double q = qc, p = pc, P, Pst, qn;
            const long row = ((i + 1) % out_every == 0) ? (i + 1) / out_every : -1;
            int info, nfev;
            if (kind == MAP_TOKAMAK && isnan(p)) {
                pc = NAN; qc = NAN;
                if (row >= 0) { pmap[row * E + k] = NAN; qmap[row * E + k] = NAN; }
                continue;
            }
            P = calcp_m(&m, q, p, &info, &nfev);
            total_fev += nfev + 1;
            if (notconv && info != 1) notconv[k] += 1;
            if (maxres) {
                target_ctx tc = { &m, q, p };
                double rr = fabs(target_f(&tc, P));
                if (!(rr <= maxres[k])) maxres[k] = rr;
            }
            Pst = P;
            if (kind == MAP_STANDARD) {
                pdc = pdc + (P - p);
                Pst = np_mod(P, two_pi);
            }
            if (kind == MAP_TOKAMAK) {
                double r = oracle_compute_r(P * 1e-2, q, 0.3);
                if (r > 0.5 || P < 0.0) Pst = NAN;
            }
            if (isnan(Pst)) {
                qn = NAN;
            } else {
                double dq = calcq_m(&m, q, Pst);
                qn = (kind == MAP_HENON) ? dq + q : np_mod(dq + q, two_pi);
            }
            qc = qn; pc = Pst;
            if (row >= 0) {
                pmap[row * E + k] = Pst;
                qmap[row * E + k] = qn;
                if (pdiff && kind == MAP_STANDARD) pdiff[row * E + k] = pdc;
            }
        }
    }
    return total_fev;
}

long oracle_applymap_alpha(int kind, int fam, double pper, long nm, long E, const double *q0, const double *p0,
                           const double *hyp, const double *hypp,
                           const double *xtp, const double *ytp, const double *alphap, long np,
                           const double *xt, const double *yt, const double *alpha, long nt,
                           double *qmap, double *pmap, double *pdiff, int *notconv, double *maxres)
{
    const double two_pi = 2.0 * M_PI;
    model_t m = { fam, pper, hyp, hypp, np, nt, xtp, ytp, alphap, xt, yt, alpha, 0, 0 };
    long total_fev = 0;
    for (long k = 0; k < E; k++) {
        pmap[k] = p0[k];
        qmap[k] = q0[k];
        if (pdiff) pdiff[k] = p0[k];
        if (notconv) notconv[k] = 0;
        if (maxres) maxres[k] = 0.0;
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_fev)
    for (long k = 0; k < E; k++) {
        for (long i = 0; i < nm - 1; i++) {
            double q = qmap[i * E + k], p = pmap[i * E + k], P, Pst;
            int info, nfev;
            if (kind == MAP_TOKAMAK && isnan(p)) {
                pmap[(i + 1) * E + k] = NAN;
                qmap[(i + 1) * E + k] = NAN;
                continue;
            }
            P = calcp_m(&m, q, p, &info, &nfev);
            total_fev += nfev + 1;
            if (notconv && info != 1) notconv[k] += 1;   /* steps on which hybrd1 did not report convergence */
            if (maxres) {                                 /* how good the accepted root really is */
                target_ctx tc = { &m, q, p };
                double rr = fabs(target_f(&tc, P));
                if (!(rr <= maxres[k])) maxres[k] = rr;
            }
            Pst = P;
            if (kind == MAP_STANDARD) {
                if (pdiff) pdiff[(i + 1) * E + k] = pdiff[i * E + k] + (P - p);
                Pst = np_mod(P, two_pi);
            }
            if (kind == MAP_TOKAMAK) {
                double r = oracle_compute_r(P * 1e-2, q, 0.3);
                if (r > 0.5 || P < 0.0) Pst = NAN;
            }
            pmap[(i + 1) * E + k] = Pst;
            if (isnan(Pst)) {
                qmap[(i + 1) * E + k] = NAN;
            } else {
                /* the reference passes the stored (possibly wrapped) pmap[i+1,k] to calcQ */
                double dq = calcq_m(&m, q, Pst);
                qmap[(i + 1) * E + k] = (kind == MAP_HENON) ? dq + q : np_mod(dq + q, two_pi);
            }
        }
    }
    return total_fev;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
