// Closed forms of the SympGPR kernel families in common-subexpression form,
// usable from device kernels and from the host scalar exports.
//
// Reference: the SymPy-generated functions of
//   python/05_tokamak/SympGPR/kernels.f90            (product:  periodic(q) * SE(P))
//   python/01_pendulum/implicit_period_unknown/kernels.f90 (product with free period p)
//   python/03_henon_heiles/kernels_sq.f90             (sq:  SE(q) * SE(P))
//   python/04_standard_map/kernels_expl_per_q_sq_p.f90 (sum: periodic(q) + SE(P))
// Argument convention of every reference function: (x_a, y_a, x_b, y_b, lx, ly[, p]).
// SURVEY.md Appendix A lists the forms; each block below cites the lines it restates.
//
// The periodic families never call sin/cos per pair: every point carries
// (u, v) = (sin(p x), cos(p x)) and the pair values follow from the angle
// addition theorem, s = sin(p (x_a - x_b)) = u_a v_b - v_a u_b, c = v_a v_b + u_a u_b.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SGP_HD __host__ __device__ __forceinline__
#else
#define SGP_HD inline
#endif

namespace sgp {

enum Family : int { FAM_PRODUCT = 0, FAM_SQ = 1, FAM_SUM = 2 };

// Per-call constants derived from (lx, ly, p); built once on the host.
struct HypC {
    double lx, ly, p;
    double lx2, ly2;          // lx^2, ly^2
    double hx, hy;            // 1/(2 lx^2), 1/(2 ly^2)
    double cxx, cxy, cyy;     // p^2/lx^4, p/(lx^2 ly^2), 1/ly^4   (sq: p = 1)
    double ily2;              // 1/ly^2
    double ilx3, ily3;        // 1/lx^3, 1/ly^3
    double cxx_lx;            // p^2/lx^7
    double cyy_lx;            // 1/(lx^3 ly^4)
    double cxy_lx;            // p/(lx^5 ly^2)
    double cxx_ly;            // p^2/(lx^4 ly^3)
    double cyy_ly;            // 1/ly^7
    double cxy_ly;            // p/(lx^2 ly^5)
    double sig;               // hyp(3)
};

inline HypC make_hypc(int fam, double lx, double ly, double sig, double p)
{
    HypC h;
    const double pe = (fam == FAM_SQ) ? 1.0 : p;
    h.lx = lx; h.ly = ly; h.p = pe;
    h.lx2 = lx * lx; h.ly2 = ly * ly;
    h.hx = 0.5 / h.lx2; h.hy = 0.5 / h.ly2;
    h.cxx = pe * pe / (h.lx2 * h.lx2);
    h.cxy = pe / (h.lx2 * h.ly2);
    h.cyy = 1.0 / (h.ly2 * h.ly2);
    h.ily2 = 1.0 / h.ly2;
    h.ilx3 = 1.0 / (h.lx2 * lx);
    h.ily3 = 1.0 / (h.ly2 * ly);
    h.cxx_lx = pe * pe / (h.lx2 * h.lx2 * h.lx2 * lx);
    h.cyy_lx = 1.0 / (h.lx2 * lx * h.ly2 * h.ly2);
    h.cxy_lx = pe / (h.lx2 * h.lx2 * lx * h.ly2);
    h.cxx_ly = pe * pe / (h.lx2 * h.lx2 * h.ly2 * ly);
    h.cyy_ly = 1.0 / (h.ly2 * h.ly2 * h.ly2 * ly);
    h.cxy_ly = pe / (h.lx2 * h.ly2 * h.ly2 * ly);
    h.sig = sig;
    return h;
}

// exp(x) for x <= 0 (every exponent of every kernel family is a negative sum of squares):
// branch-free, |relative error| < 2^-52 * 1.5 over [-708, 0]; 0 below.  k = round(x/ln2),
// r = x - k ln2 (two-term Cody-Waite), degree-13 Taylor polynomial evaluated as two interleaved
// Horner chains (even/odd) for instruction-level parallelism, scaling by 2^k through the exponent
// bits.  Replaces the libm/libdevice exp, whose range checks cost branches in the inner loops.
SGP_HD double exp_neg(double x)
{
    const double LOG2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    if (!(x > -708.0)) return (x != x) ? x : 0.0;         // underflow to 0; NaN propagates
    const double SHIFT = 6755399441055744.0;              // 1.5 * 2^52: adds round-to-nearest integer in the low bits
    const double kd = (x * LOG2E + SHIFT) - SHIFT;
    const double r = (x - kd * LN2_HI) - kd * LN2_LO;
    const double r2 = r * r;
    // even part: 1 + r^2/2! + r^4/4! + ... + r^12/12!   odd part: 1 + r^2/3! + ... + r^12/13!
    double pe = 2.08767569878680989792e-09;               // 1/12!
    double po = 1.60590438368216145994e-10;               // 1/13!
    pe = pe * r2 + 2.75573192239858906526e-07;            // 1/10!
    po = po * r2 + 2.50521083854417187751e-08;            // 1/11!
    pe = pe * r2 + 2.48015873015873015873e-05;            // 1/8!
    po = po * r2 + 2.75573192239858906526e-06;            // 1/9!
    pe = pe * r2 + 1.38888888888888888889e-03;            // 1/6!
    po = po * r2 + 1.98412698412698412698e-04;            // 1/7!
    pe = pe * r2 + 4.16666666666666666667e-02;            // 1/4!
    po = po * r2 + 8.33333333333333333333e-03;            // 1/5!
    pe = pe * r2 + 0.5;                                   // 1/2!
    po = po * r2 + 1.66666666666666666667e-01;            // 1/3!
    pe = pe * r2 + 1.0;
    po = po * r2 + 1.0;
    const double e = pe + r * po;
    // 2^k, k in [-1022, 0]
    const long long ki = (long long)kd;
    union { long long i; double d; } sc;
    sc.i = (ki + 1023LL) << 52;
    return e * sc.d;
}

// Point features: periodic families (sin(p x), cos(p x), y); sq family (x, 0, y).
struct Pt { double u, v, y; };

template <int FAM>
SGP_HD Pt make_pt(double x, double y, double p)
{
    Pt r;
    if (FAM == FAM_SQ) { r.u = x; r.v = 0.0; }
    else { double s, c; sincos(p * x, &s, &c); r.u = s; r.v = c; }
    r.y = y;
    return r;
}

// Shared intermediates of one (a, b) pair.
template <int FAM>
struct Pair {
    double dy;     // y_a - y_b
    double s, c;   // periodic: sin/cos(p dx);  sq: s = dx, c unused
    double E;      // product/sq: full exponential;  sum: Ex (periodic part)
    double Ey;     // sum only: SE part

    SGP_HD Pair(const Pt& a, const Pt& b, const HypC& h)
    {
        dy = a.y - b.y;
        if (FAM == FAM_SQ) {
            s = a.u - b.u; c = 0.0;
            E = exp_neg(-(dy * dy) * h.hy - (s * s) * h.hx);   // kernels_sq.f90:9
            Ey = 0.0;
        } else {
            s = a.u * b.v - a.v * b.u;
            c = a.v * b.v + a.u * b.u;
            if (FAM == FAM_PRODUCT) {
                E = exp_neg(-(dy * dy) * h.hy - (s * s) * h.hx);   // kernels.f90:9-10
                Ey = 0.0;
            } else {
                E = exp_neg(-(s * s) * h.hx);                  // kernels_expl_per_q_sq_p.f90:9-10
                Ey = exp_neg(-(dy * dy) * h.hy);
            }
        }
    }

    // kern_num
    SGP_HD double k() const { return FAM == FAM_SUM ? Ey + E : E; }

    // bracket of d2kdxdx0:  periodic lx^2 cos(2 p dx) - s^2 c^2 ;  sq lx^2 - dx^2
    SGP_HD double bxx(const HypC& h) const
    {
        if (FAM == FAM_SQ) return h.lx2 - s * s;               // kernels_sq.f90:63-64
        const double s2 = s * s;
        return h.lx2 * (1.0 - 2.0 * s2) - s2 * (c * c);        // kernels.f90:66-69
    }
    // s*c (periodic) or dx (sq): the odd-in-dx factor of the mixed block
    SGP_HD double odd() const { return FAM == FAM_SQ ? s : s * c; }

    // d2kdxdx0, d2kdxdy0 (= d2kdydx0), d2kdydy0       kernels.f90:58-94 / kernels_sq.f90:56-87
    SGP_HD double kxx(const HypC& h) const { return h.cxx * bxx(h) * E; }
    SGP_HD double kxy(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;                        // kernels_expl_per_q_sq_p.f90:87
        return -h.cxy * dy * odd() * E;
    }
    SGP_HD double kyy(const HypC& h) const
    {
        return h.cyy * (h.ly2 - dy * dy) * (FAM == FAM_SUM ? Ey : E);
    }

    // first derivatives dkdx (= -dkdx0) and dkdy (= -dkdy0)     kernels.f90:12-57
    SGP_HD double kx(const HypC& h) const
    {
        if (FAM == FAM_SQ) return -s * E / h.lx2;
        return -h.p * odd() * E / h.lx2;
    }
    SGP_HD double ky(const HypC& h) const { return -dy * (FAM == FAM_SUM ? Ey : E) * h.ily2; }
    // d3kdydy0dy0                                               kernels.f90:108-119
    SGP_HD double kyy_yb(const HypC& h) const
    {
        return h.cyy * h.ily2 * (3.0 * h.ly2 - dy * dy) * dy * (FAM == FAM_SUM ? Ey : E);
    }

    // d/dy_b of the two blocks of row 1 (analytic Jacobian of the implicit equation)
    // d3kdxdx0dy0, d3kdxdy0dy0                         kernels.f90:95-107,120-132
    SGP_HD double kxx_yb(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cxx * h.ily2 * dy * bxx(h) * E;
    }
    SGP_HD double kxy_yb(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cxy * h.ily2 * (h.ly2 - dy * dy) * odd() * E;
    }

    // hyper-parameter derivatives                      kernels.f90:133-231 / kernels_sq.f90:126-216
    SGP_HD double k_lx(const HypC& h) const { return (s * s) * E * h.ilx3; }
    SGP_HD double k_ly(const HypC& h) const { return (dy * dy) * (FAM == FAM_SUM ? Ey : E) * h.ily3; }

    SGP_HD double kxx_lx(const HypC& h) const
    {
        const double s2 = s * s;
        if (FAM == FAM_SQ)
            return h.cxx_lx * (-2.0 * h.lx2 * h.lx2 + 5.0 * h.lx2 * s2 - s2 * s2) * E;
        const double C2 = 1.0 - 2.0 * s2;
        return h.cxx_lx * (-2.0 * h.lx2 * h.lx2 * C2 + h.lx2 * (3.0 * C2 + 2.0) * s2 - s2 * s2 * (c * c)) * E;
    }
    SGP_HD double kyy_lx(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cyy_lx * (h.ly2 - dy * dy) * (s * s) * E;
    }
    SGP_HD double kxy_lx(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cxy_lx * (2.0 * h.lx2 - s * s) * dy * odd() * E;
    }
    SGP_HD double kxx_ly(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cxx_ly * (dy * dy) * bxx(h) * E;
    }
    SGP_HD double kyy_ly(const HypC& h) const
    {
        const double d2 = dy * dy;
        return h.cyy_ly * (-2.0 * h.ly2 * h.ly2 + 5.0 * h.ly2 * d2 - d2 * d2) * (FAM == FAM_SUM ? Ey : E);
    }
    SGP_HD double kxy_ly(const HypC& h) const
    {
        if (FAM == FAM_SUM) return 0.0;
        return h.cxy_ly * (2.0 * h.ly2 - dy * dy) * dy * odd() * E;
    }
};

}  // namespace sgp
