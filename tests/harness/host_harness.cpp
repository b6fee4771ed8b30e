// TEST HARNESS (CPU): compiles the product's device headers (forms.cuh, hybrd.cuh) with g++ so
// their scalar logic -- closed forms, the resumable Hybrd1 / Newton1 state machines and one map
// step -- can be checked against the oracle without a GPU.  Not part of the product library.
#include <math.h>
#include <stdlib.h>

#include "../../sympgpr_b200/csrc/forms.cuh"
#include "../../sympgpr_b200/csrc/hybrd.cuh"

using namespace sgp;

template <int FAM>
static void sums(const HypC& h, double q, double P, double per, const double* xt, const double* yt, const double* alpha,
                 long nt, double* F, double* dF, double* dQ)
{
    const Pt b = make_pt<FAM>(q, P, h.p);
    double f = 0, d = 0, g = 0;
    for (long j = 0; j < nt; j++) {
        const Pt a = make_pt<FAM>(xt[j], yt[j], h.p);
        const Pair<FAM> pr(a, b, h);
        f += pr.kxx(h) * alpha[j] + pr.kxy(h) * alpha[nt + j];
        d += pr.kxx_yb(h) * alpha[j] + pr.kxy_yb(h) * alpha[nt + j];
        g += pr.kxy(h) * alpha[j] + pr.kyy(h) * alpha[nt + j];
    }
    *F = h.sig * f; *dF = h.sig * d; *dQ = h.sig * g;
}

template <int FAM>
static double guess(const HypC& hp, double q, double p, const double* xtp, const double* ytp, const double* alphap, long np)
{
    const Pt b = make_pt<FAM>(q, p, hp.p);
    double s = 0;
    for (long j = 0; j < np; j++) {
        const Pair<FAM> pr(make_pt<FAM>(xtp[j], ytp[j], hp.p), b, hp);
        s += pr.k() * alphap[j];
    }
    return hp.sig * s;
}

template <int FAM>
static double calcp_t(int solver, double per, double q, double p, const double* hyp, const double* hypp, const double* xtp,
                      const double* ytp, const double* alphap, long np, const double* xt, const double* yt,
                      const double* alpha, long nt, int* info, int* nfev, double* dq_out)
{
    const HypC h = make_hypc(FAM, hyp[0], hyp[1], hyp[2], per), hp = make_hypc(FAM, hypp[0], hypp[1], hypp[2], per);
    const double pg = guess<FAM>(hp, q, p, xtp, ytp, alphap, np);
    double F, dF, dQ, P;
    int n = 0;
    if (solver == 0) {
        Hybrd1 sv; sv.start(pg);
        while (!sv.done()) { const double x = sv.query(); sums<FAM>(h, q, x, per, xt, yt, alpha, nt, &F, &dF, &dQ); sv.feed(F - p + x); n++; }
        P = sv.root(); *info = sv.info;
    } else {
        Newton1 sv; sv.start(solver == 3 ? p + pg : pg);        // 3 = SGP_SOLVER_NEWTON_DELTA: the guess GP predicts P - p
        while (!sv.done()) { const double x = sv.query(); sums<FAM>(h, q, x, per, xt, yt, alpha, nt, &F, &dF, &dQ); sv.feed(F - p + x, 1.0 + dF); n++; }
        P = sv.root(); *info = sv.info;
    }
    *nfev = n;
    sums<FAM>(h, q, P, per, xt, yt, alpha, nt, &F, &dF, &dQ);
    *dq_out = dQ;
    return P;
}

extern "C" double harness_calcp(int fam, int solver, double per, double q, double p, const double* hyp, const double* hypp,
                                const double* xtp, const double* ytp, const double* alphap, long np, const double* xt,
                                const double* yt, const double* alpha, long nt, int* info, int* nfev, double* dq_out)
{
    if (fam == FAM_SQ) return calcp_t<FAM_SQ>(solver, per, q, p, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, info, nfev, dq_out);
    if (fam == FAM_SUM) return calcp_t<FAM_SUM>(solver, per, q, p, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, info, nfev, dq_out);
    return calcp_t<FAM_PRODUCT>(solver, per, q, p, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, info, nfev, dq_out);
}

// derivative check hook: F and dF/dP at (q, P)
extern "C" void harness_f_df(int fam, double per, double q, double P, const double* hyp, const double* xt, const double* yt,
                             const double* alpha, long nt, double* F, double* dF, double* dQ)
{
    if (fam == FAM_SQ) { const HypC h = make_hypc(FAM_SQ, hyp[0], hyp[1], hyp[2], per); sums<FAM_SQ>(h, q, P, per, xt, yt, alpha, nt, F, dF, dQ); }
    else { const HypC h = make_hypc(FAM_PRODUCT, hyp[0], hyp[1], hyp[2], per); sums<FAM_PRODUCT>(h, q, P, per, xt, yt, alpha, nt, F, dF, dQ); }
}

// Ensemble loops (python/functions/func.py:216-260, 04_standard_map/func.py:218-254, 05_tokamak/SympGPR/func.py:182-211)
// on the host with the product's own arithmetic (forms.cuh closed forms, Newton1 / Hybrd1 state machines of hybrd.cuh,
// sequential sums): a CPU stand-in for map_kernel that lets the parity fixtures be checked against the device code's
// logic without a GPU.  kind: 0 pendulum, 1 henon, 2 standard, 3 tokamak.  Rows 0, every, 2 every, ... are returned.
static double np_mod_h(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}

static double compute_r_h(double pth, double th, double rstart)
{
    double r = rstart;
    const double ct = cos(th);
    for (int k = 0; k < 20; k++) {
        const double yv = pth - (r * r / 2.0 - r * r * r / 3.0 * ct);
        const double dy = -(r - r * r * ct);
        r = r - yv / dy;
    }
    return r;
}

template <int FAM>
static long applymap_t(int kind, int solver, double per, long nsteps, long E, const double* q0, const double* p0,
                       const double* hyp, const double* hypp, const double* xtp, const double* ytp, const double* alphap, long np,
                       const double* xt, const double* yt, const double* alpha, long nt, long every, double* qout, double* pout)
{
    const double two_pi = 6.283185307179586;
    long evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
    for (long k = 0; k < E; k++) {
        double q = q0[k], p = p0[k];
        qout[k] = q; pout[k] = p;
        for (long s = 1; s <= nsteps; s++) {
            double qn = NAN, Pst = NAN;
            if (q == q && p == p) {
                int info, nfev;
                double dq;
                const double P = calcp_t<FAM>(solver, per, q, p, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, &info, &nfev, &dq);
                evals += nfev + 1;
                Pst = P;
                if (kind == 2) Pst = np_mod_h(P, two_pi);
                if (kind == 3) { const double r = compute_r_h(P * 1e-2, q, 0.3); if (r > 0.5 || P < 0.0) Pst = NAN; }
                if (Pst == Pst) {
                    // the reference evaluates dq at the stored (possibly wrapped) momentum
                    const HypC h = make_hypc(FAM, hyp[0], hyp[1], hyp[2], per);
                    double F, dF, dQ;
                    sums<FAM>(h, q, Pst, per, xt, yt, alpha, nt, &F, &dF, &dQ);
                    qn = (kind == 1) ? dQ + q : np_mod_h(dQ + q, two_pi);
                }
            }
            q = qn; p = Pst;
            if (s % every == 0) { qout[(s / every) * E + k] = q; pout[(s / every) * E + k] = p; }
        }
    }
    return evals;
}

extern "C" long harness_applymap(int fam, int kind, int solver, double per, long nsteps, long E, const double* q0, const double* p0,
                                 const double* hyp, const double* hypp, const double* xtp, const double* ytp,
                                 const double* alphap, long np, const double* xt, const double* yt, const double* alpha,
                                 long nt, long every, double* qout, double* pout)
{
    if (fam == FAM_SQ) return applymap_t<FAM_SQ>(kind, solver, per, nsteps, E, q0, p0, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, every, qout, pout);
    return applymap_t<FAM_PRODUCT>(kind, solver, per, nsteps, E, q0, p0, hyp, hypp, xtp, ytp, alphap, np, xt, yt, alpha, nt, every, qout, pout);
}
