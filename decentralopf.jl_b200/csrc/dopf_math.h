// Per-agent math of the B200 ADMM iteration, shared by every kernel (host+device inline).
//
// What is solved (derivation in DESIGN.md section 3; reference objective:
// /root/reference/src/optimization/subproblems.jl:63-83,162-183 + penalty_terms.jl:1-53):
// after eliminating the per-agent slack copies U,K the agent at node n sees, per timestep t,
//     phi'(delta) = g0 + s1*delta + corr(delta)
// where (g0,s1) is the linearisation with the slack-clip pattern of delta = 0 ("anchor") and
// corr() collects the hinges whose clip state differs from the anchor at delta.
//   generator: root of  mc + phi'(delta) + prox*delta  on the box      (1-D, monotone)
//   storage:   min sum_t mc(D+C) + prox/2((D-Db)^2+(C-Cb)^2) + phi_t(delta_t)
//              s.t. 0<=D,C<=p, 0<=E_t=cumsum(C-D)<=emax; solved through the multiplier path
//              eta_t of the level constraints ("funnel"/planning-horizon algorithm).
#ifndef DOPF_MATH_H
#define DOPF_MATH_H

#include <math.h>
#include <stdint.h>
#if !defined(__CUDACC__) && defined(DOPF_DEBUG_WARM)
#include <cstdio>
#endif

#if defined(__CUDACC__)
#define DOPF_HD __host__ __device__ __forceinline__
#else
#define DOPF_HD inline
#endif

namespace dopf {

#if !defined(__CUDACC__) && defined(DOPF_COUNT_ITERS)
static long dopf_iters_root[202], dopf_iters_sto[202];   // test-only iteration histograms
#define DOPF_COUNT(arr, it) (++arr[(it) > 200 ? 200 : (it)])
#else
#define DOPF_COUNT(arr, it) ((void)0)
#endif

// ---------------------------------------------------------------------------------------------
// Hinges.  One hinge = one (line, side) whose per-agent slack clips at 0 when the agent moves:
//   term(delta) = s*(delta-bp)  on the active side  dir*(delta-bp) > 0,  0 otherwise.
// `anchored` = active at delta=0 (then its linear part is already inside g0,s1).
// ---------------------------------------------------------------------------------------------
struct Hinge {
    double bp;   // breakpoint in delta
    double sg;   // dir * s   (s = beta*p^2 > 0)
};

DOPF_HD void hinge_accum(const Hinge h, double delta, double &val, double &slope)
{
    const double dir = h.sg > 0.0 ? 1.0 : -1.0;
    const double s = fabs(h.sg);
    const bool anchored = dir * h.bp < 0.0;
    const double e = dir * (delta - h.bp);
    if (!anchored) {
        if (e > 0.0) { val += s * (delta - h.bp); slope += s; }
    } else {
        if (e < 0.0) { val -= s * (delta - h.bp); slope -= s; }
    }
}

// one hinge's share of: the correction value at delta, the slopes on both sides of delta and the nearest breakpoints
// strictly beyond delta on both sides.  A breakpoint within `tol` of delta counts as reached: on each side the hinge is
// in the state it has beyond the breakpoint.
DOPF_HD void hinge_accum2(const Hinge h, double delta, double tol, double &val, double &sL, double &sR, double &nL, double &nR)
{
    const double bp = h.bp;
    const double dir = h.sg > 0.0 ? 1.0 : -1.0, s = fabs(h.sg);
    const bool anchored = dir * bp < 0.0;
    const double w = delta - bp;
    const bool at = fabs(w) <= tol;
    if (!at) { if (bp > delta) { if (bp < nR) nR = bp; } else { if (bp > nL) nL = bp; } }
    const bool pos = dir * w > 0.0;                         // side of the hinge delta is on (if not at it)
    const bool actR = at ? dir > 0.0 : pos, actL = at ? dir < 0.0 : pos;
    if (!anchored) {
        if (actR) sR += s;
        if (actL) sL += s;
        if (!at && pos) val += s * w;
    } else {
        if (!actR) sR -= s;
        if (!actL) sL -= s;
        if (!at && !pos) val -= s * w;
    }
}

// view of a hinge list (possibly empty)
struct HingeList {
    const Hinge *h;
    int n;
    bool sorted;   // entries ordered by |bp| ascending: hinges beyond |delta| keep their anchor state => early exit
    DOPF_HD void eval(double delta, double &val, double &slope) const
    {
        val = 0.0; slope = 0.0;
        const double ad = fabs(delta);
        for (int i = 0; i < n; ++i) {
            if (sorted && fabs(h[i].bp) > ad) break;
            hinge_accum(h[i], delta, val, slope);
        }
    }
    // value, the slopes on both sides of delta and the nearest breakpoints strictly beyond delta on both
    // sides.  A breakpoint within rounding distance of delta counts as reached: on each side the hinge is
    // in the state it has beyond the breakpoint.  For sorted lists the first hinge beyond |delta| bounds
    // the distance to every remaining breakpoint (a conservative `next` on both sides).
    DOPF_HD void eval2(double delta, double &val, double &sL, double &sR, double &nL, double &nR, double xtol = 0.0) const
    {
        val = 0.0; sL = 0.0; sR = 0.0; nL = -1e300; nR = 1e300;
        const double ad = fabs(delta), tol = 1e-14 * (1.0 + ad) + xtol;
        for (int i = 0; i < n; ++i) {
            const double abp = fabs(h[i].bp);
            if (sorted && abp > ad + tol) {
                if (abp < nR) nR = abp;
                if (-abp > nL) nL = -abp;
                break;
            }
            hinge_accum2(h[i], delta, tol, val, sL, sR, nL, nR);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Generator: root of f(delta) = c + a*delta + corr(delta) on [lo,hi], f increasing (a>0 and
// every piece of corr keeps the total slope >= prox).  Exact for piecewise-linear f:
// safeguarded Newton that ends as soon as a Newton step stays inside one linear piece.
// ---------------------------------------------------------------------------------------------
DOPF_HD double root_monotone_pl(double c, double a, const HingeList &hl, double lo, double hi)
{
    if (hl.n == 0) {
        double d = -c / a;
        return d < lo ? lo : (d > hi ? hi : d);
    }
    double v, s;
    hl.eval(lo, v, s);
    double flo = c + a * lo + v;
    if (flo >= 0.0) return lo;
    hl.eval(hi, v, s);
    double fhi = c + a * hi + v;
    if (fhi <= 0.0) return hi;
    double x = -c / a;
    if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
    // walk piece by piece towards the root: f is linear between x and the next breakpoint on the side
    // the root lies on, so either the Newton step of that piece is the exact root or the root lies
    // beyond the breakpoint.  The direction never changes => at most n+1 evaluations.
    const int cap = 2 * hl.n + 8;
    for (int it = 0; it < cap; ++it) {
        double sL, sR, nL, nR;
        hl.eval2(x, v, sL, sR, nL, nR);
        const double f = c + a * x + v;
        if (f == 0.0) { DOPF_COUNT(dopf_iters_root, it + 1); return x; }
        if (f < 0.0) {
            const double xn = x - f / (a + sR);
            if (xn <= nR) { DOPF_COUNT(dopf_iters_root, it + 1); return xn < hi ? xn : hi; }
            x = nR;
        } else {
            const double xn = x - f / (a + sL);
            if (xn >= nL) { DOPF_COUNT(dopf_iters_root, it + 1); return xn > lo ? xn : lo; }
            x = nL;
        }
    }
    DOPF_COUNT(dopf_iters_root, 201);
    return x;
}

// ---------------------------------------------------------------------------------------------
// Storage, one timestep, for a given level multiplier eta (Lagrangian term  eta*(C-D)).
// Stationarity:  D = clip(Db - (mc+nu)/prox), C = clip(Cb - (mc-nu)/prox)   with
//                nu = g0 - eta + s1*delta + corr(delta),  delta = (D-Db) - (C-Cb).
// Psi(nu) = nu - (g0-eta) - s1*delta(nu) - corr(delta(nu)) is increasing with slope >= 1.
// ---------------------------------------------------------------------------------------------
struct StoStep {
    double Db, Cb;   // previous discharge / charge
    double g0, s1;   // anchor linearisation at this (node, t)
};

struct StoConst {
    double mc, pmax, emax, prox;
    double iprox;   // 1/prox
};

DOPF_HD double clip01(double v, double hi) { return v < 0.0 ? 0.0 : (v > hi ? hi : v); }

struct StoEval {
    double D, C;     // solution at this eta
    double dy;       // d(C-D)/d eta  (<= 0)
};

DOPF_HD void sto_dc_of_nu(const StoStep &st, const StoConst &k, double nu, double &D, double &C, int &nfree)
{
    const double du = st.Db - (k.mc + nu) * k.iprox;
    const double cu = st.Cb - (k.mc - nu) * k.iprox;
    D = clip01(du, k.pmax);
    C = clip01(cu, k.pmax);
    nfree = (du > 0.0 && du < k.pmax) + (cu > 0.0 && cu < k.pmax);
}

// ---- per-timestep clip table of the hinge-free storage step (independent of eta) ---------------------------------
// Psi(nu) = nu - (g0 - eta) - s1*delta(nu) is piecewise linear with the four clip breakpoints bb[0..3] of D(nu), C(nu);
// Psi(bb[i]) = eta - e[i] with the eta-thresholds e[i] = g0 + s1*dl[i] - bb[i] (non-increasing in i, dl[i] = delta at
// bb[i]).  The thresholds cut the eta-axis into five pieces p = #{i : eta < e[i]}; on piece p nu is affine in eta,
// nu = nu0[p] + nus[p]*eta, and dy/deta = dyv[p] is constant.  On a piece with nf free variables Psi has slope
// 1 + nf*s1/prox, whose inverse is prox*r_nf with r_nf = 1/(prox + nf*s1) (r1, r2 passed in): no division here.
//
// Closed form: with u1 = pmax - D = clip((nu - c1)/prox, 0, pmax) and u2 = C = clip((nu - c2)/prox, 0, pmax) the move is
// delta = K - (u1 + u2), K = pmax - Db + Cb, and u1 + u2 at the sorted breakpoints is 0, (bb1-bb0)/prox,
// 2 pmax - (bb3-bb2)/prox, 2 pmax; the middle piece has two free variables if the two ramps overlap and none otherwise.
DOPF_HD void sto_clip_table(const StoStep &st, const StoConst &k, double r1, double r2, double (&e)[4], double (&nu0)[5], double (&nus)[5], double (&dyv)[5])
{
    const double b0 = k.prox * (st.Db - k.pmax) - k.mc, b1 = k.prox * st.Db - k.mc;          // D leaves pmax / reaches 0
    const double b2 = k.mc - k.prox * st.Cb, b3 = k.mc + k.prox * (k.pmax - st.Cb);          // C leaves 0 / reaches pmax
    const double lo_hi = b0 > b2 ? b0 : b2, hi_lo = b1 < b3 ? b1 : b3;                        // later ramp start, earlier ramp end
    double bb[4];
    bb[0] = b0 < b2 ? b0 : b2; bb[3] = b1 > b3 ? b1 : b3;
    bb[1] = lo_hi < hi_lo ? lo_hi : hi_lo; bb[2] = lo_hi < hi_lo ? hi_lo : lo_hi;
    const bool overlap = lo_hi < hi_lo;
    const double K = k.pmax - st.Db + st.Cb;
    const double su[4] = { 0.0, (bb[1] - bb[0]) * k.iprox, 2.0 * k.pmax - (bb[3] - bb[2]) * k.iprox, 2.0 * k.pmax };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) e[i] = st.g0 + st.s1 * (K - su[i]) - bb[i];
    nu0[0] = bb[0] + e[0]; nus[0] = -1.0; dyv[0] = 0.0;           // outer pieces: both variables clipped, slope 1
    nu0[4] = bb[3] + e[3]; nus[4] = -1.0; dyv[4] = 0.0;
    const int nfp[3] = { bb[1] > bb[0] ? 1 : 0, overlap ? 2 : 0, bb[3] > bb[2] ? 1 : 0 };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int f = 0; f < 3; ++f) {     // inner piece between bb[f] and bb[f+1] = eta-piece p = f+1, anchored at breakpoint f
        const int nf = nfp[f];
        const double rn = nf == 1 ? r1 : r2;
        const double as = nf == 0 ? 1.0 : k.prox * rn;
        nu0[f + 1] = bb[f] + e[f] * as; nus[f + 1] = -as; dyv[f + 1] = nf == 0 ? 0.0 : -(double)nf * rn;
    }
}
// the same table from evaluations of D, C at the breakpoints and at the piece midpoints (the form the closed form is
// checked against in tests/test_host_math.py)
DOPF_HD void sto_clip_table_ref(const StoStep &st, const StoConst &k, double r1, double r2, double (&e)[4], double (&nu0)[5], double (&nus)[5], double (&dyv)[5])
{
    double b0 = k.prox * (st.Db - k.pmax) - k.mc, b1 = k.prox * st.Db - k.mc;
    double b2 = k.mc - k.prox * st.Cb, b3 = k.mc + k.prox * (k.pmax - st.Cb);
    double x;
    if (b0 > b2) { x = b0; b0 = b2; b2 = x; }
    if (b1 > b3) { x = b1; b1 = b3; b3 = x; }
    if (b1 > b2) { x = b1; b1 = b2; b2 = x; }
    const double bb[4] = { b0, b1, b2, b3 };
    for (int i = 0; i < 4; ++i) {
        double D, C; int nf;
        sto_dc_of_nu(st, k, bb[i], D, C, nf);
        e[i] = st.g0 + st.s1 * ((D - st.Db) - (C - st.Cb)) - bb[i];
    }
    nu0[0] = bb[0] + e[0]; nus[0] = -1.0; dyv[0] = 0.0;
    nu0[4] = bb[3] + e[3]; nus[4] = -1.0; dyv[4] = 0.0;
    for (int f = 0; f < 3; ++f) {
        double D, C; int nf;
        sto_dc_of_nu(st, k, 0.5 * (bb[f] + bb[f + 1]), D, C, nf);
        const double rn = nf == 1 ? r1 : r2;
        const double as = nf == 0 ? 1.0 : k.prox * rn;
        nu0[f + 1] = bb[f] + e[f] * as; nus[f + 1] = -as; dyv[f + 1] = nf == 0 ? 0.0 : -(double)nf * rn;
    }
}

DOPF_HD StoEval sto_eval(const StoStep &st, const StoConst &k, const HingeList &hl, double eta)
{
    const double base = st.g0 - eta;
    double D, C; int nf;
    StoEval r;
    double nu;
    {
        // closed form without hinges: Psi is piecewise linear with the 4 clip breakpoints of D(nu), C(nu)
        double b0 = k.prox * (st.Db - k.pmax) - k.mc;  // D leaves pmax
        double b1 = k.prox * st.Db - k.mc;             // D reaches 0
        double b2 = k.mc - k.prox * st.Cb;             // C leaves 0
        double b3 = k.mc + k.prox * (k.pmax - st.Cb);  // C reaches pmax
        double t;                                      // sort (b0<=b1, b2<=b3 already)
        if (b0 > b2) { t = b0; b0 = b2; b2 = t; }
        if (b1 > b3) { t = b1; b1 = b3; b3 = t; }
        if (b1 > b2) { t = b1; b1 = b2; b2 = t; }
        const double bb[4] = { b0, b1, b2, b3 };
        double psi[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i) {
            sto_dc_of_nu(st, k, bb[i], D, C, nf);
            psi[i] = bb[i] - base - st.s1 * ((D - st.Db) - (C - st.Cb));
        }
        if (psi[0] >= 0.0) nu = bb[0] - psi[0];                 // slope 1 left of all breakpoints
        else if (psi[3] <= 0.0) nu = bb[3] - psi[3];            // slope 1 right of all breakpoints
        else {
            const int f = psi[1] >= 0.0 ? 0 : (psi[2] >= 0.0 ? 1 : 2);
            const double pl = bb[f], pv = psi[f], ql = bb[f + 1], qv = psi[f + 1];
            nu = (qv == pv) ? pl : pl - pv * (ql - pl) / (qv - pv);
        }
        sto_dc_of_nu(st, k, nu, D, C, nf);
    }
    double sl = 0.0;
    if (hl.n != 0) {
        double v;
        hl.eval((D - st.Db) - (C - st.Cb), v, sl);
        if (v != 0.0 || sl != 0.0) {
            // a hinge differs from its anchor state at this delta: safeguarded Newton on Psi with hinges,
            // started at the hinge-free solution; the bracket is built lazily from the iterates (Psi is
            // increasing with slope >= 1, so |Psi| bounds the distance to the root)
            // Same walk as root_monotone_pl, in nu: Psi is linear up to the next clip breakpoint of D, C
            // or the next hinge breakpoint (delta falls when nu rises), whichever comes first.
            const double bD0 = k.prox * (st.Db - k.pmax) - k.mc, bD1 = k.prox * st.Db - k.mc;
            const double bC0 = k.mc - k.prox * st.Cb, bC1 = k.mc + k.prox * (k.pmax - st.Cb);
            const double cb[4] = { bD0, bD1, bC0, bC1 };
            const int cap = 2 * hl.n + 16;
            int it = 0;
            for (; it < cap; ++it) {
                sto_dc_of_nu(st, k, nu, D, C, nf);
                const double delta = (D - st.Db) - (C - st.Cb);
                double sLd, sRd, nLd, nRd;
                hl.eval2(delta, v, sLd, sRd, nLd, nRd, 1e-15 * (1.0 + fabs(nu)) * k.iprox);   // a step to a hinge always moves nu
                const double psi = nu - base - st.s1 * delta - v;
                if (psi == 0.0) break;
                const bool up = psi < 0.0;                           // nu has to rise
                const double tn = 1e-14 * (1.0 + fabs(nu));
                int nfd = 0;                                         // free variables on the side of travel
                double nb = up ? 1e300 : -1e300;                     // next breakpoint in nu on that side
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int q = 0; q < 2; ++q) {
                    const double b0 = cb[2 * q], b1 = cb[2 * q + 1];
                    const bool at0 = fabs(nu - b0) <= tn, at1 = fabs(nu - b1) <= tn;
                    if (at0 && at1) continue;                        // pmax = 0: never free
                    nfd += at0 ? up : (at1 ? !up : (nu > b0 && nu < b1));
                    if (!at0 && (up ? (b0 > nu && b0 < nb) : (b0 < nu && b0 > nb))) nb = b0;
                    if (!at1 && (up ? (b1 > nu && b1 < nb) : (b1 < nu && b1 > nb))) nb = b1;
                }
                const double sld = up ? sLd : sRd, nbd = up ? nLd : nRd;
                if (nfd > 0 && fabs(nbd) < 1e300) {
                    const double nh = nu + (delta - nbd) * k.prox / nfd;
                    if (up ? (nh > nu && nh < nb) : (nh < nu && nh > nb)) nb = nh;
                }
                const double slope = 1.0 + (st.s1 + sld) * nfd * k.iprox;
                const double nn = nu - psi / slope;
                if (up ? nn <= nb : nn >= nb) { nu = nn; break; }
                nu = nb;
            }
            DOPF_COUNT(dopf_iters_sto, it + 1);
            sto_dc_of_nu(st, k, nu, D, C, nf);
            hl.eval((D - st.Db) - (C - st.Cb), v, sl);
        }
    }
    r.D = D; r.C = C;
    r.dy = -(double)nf / (k.prox + (st.s1 + sl) * nf);
    return r;
}

// eta-breakpoints of one step without hinges: eta(nu_b) = g0 + s1*delta(nu_b) - nu_b at the four
// clip breakpoints nu_b of D(nu), C(nu).  Returns the nearest one strictly beyond `eta` in the
// direction `up` (or +-1e300 if none).
DOPF_HD double sto_next_break(const StoStep &st, const StoConst &k, double eta, bool up)
{
    const double nb[4] = { k.prox * (st.Db - k.pmax) - k.mc, k.prox * st.Db - k.mc,
                           k.mc - k.prox * st.Cb, k.mc + k.prox * (k.pmax - st.Cb) };
    double best = up ? 1e300 : -1e300;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
        double D, C; int nf;
        sto_dc_of_nu(st, k, nb[i], D, C, nf);
        const double e = st.g0 + st.s1 * ((D - st.Db) - (C - st.Cb)) - nb[i];
        if (up) { if (e > eta && e < best) best = e; }
        else { if (e < eta && e > best) best = e; }
    }
    return best;
}

// maximal eta-interval [ilo,ihi] containing eta on which y_t = C-D of a hinge-free step stays
// at its value y_t(eta): the union of the clip pieces (in nu) without a free variable that contain
// or touch nu(eta), mapped through the decreasing map eta(nu) = g0 + s1*delta(nu) - nu.
DOPF_HD void sto_flat_interval(const StoStep &st, const StoConst &k, double eta, double D, double C, double &ilo, double &ihi)
{
    ilo = ihi = eta;
    const double nu = st.g0 - eta + st.s1 * ((D - st.Db) - (C - st.Cb));
    double b0 = k.prox * (st.Db - k.pmax) - k.mc, b1 = k.prox * st.Db - k.mc;
    double b2 = k.mc - k.prox * st.Cb, b3 = k.mc + k.prox * (k.pmax - st.Cb);
    double t;
    if (b0 > b2) { t = b0; b0 = b2; b2 = t; }
    if (b1 > b3) { t = b1; b1 = b3; b3 = t; }
    if (b1 > b2) { t = b1; b1 = b2; b2 = t; }
    const double bb[4] = { b0, b1, b2, b3 };
    const double tol = 1e-10 * (1.0 + fabs(nu));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 5; ++j) {
        const bool linf = j == 0, rinf = j == 4;
        const double plo = linf ? 0.0 : bb[j - 1], phi = rinf ? 0.0 : bb[j];
        if (!linf && nu < plo - tol) continue;
        if (!rinf && nu > phi + tol) continue;
        if (!linf && !rinf && !(phi > plo)) continue;
        const double mid = linf ? phi - 1.0 : (rinf ? plo + 1.0 : 0.5 * (plo + phi));
        double Dm, Cm; int nf;
        sto_dc_of_nu(st, k, mid, Dm, Cm, nf);
        if (nf != 0) continue;
        const double dl = (Dm - st.Db) - (Cm - st.Cb);     // delta is constant on a flat piece
        const double eh = linf ? 1e300 : st.g0 + st.s1 * dl - plo;
        const double el = rinf ? -1e300 : st.g0 + st.s1 * dl - phi;
        ilo = el < ilo ? el : ilo; ihi = eh > ihi ? eh : ihi;
    }
}

// ---------------------------------------------------------------------------------------------
// Lane groups.  The storage solver is written once over a group of W lanes that own the
// timesteps t = lane, lane+W, ...  W=32 is a warp on the device; W=1 is the sequential
// restatement used by the host-side emulation in tests/.
// ---------------------------------------------------------------------------------------------
template <int W> struct Group;

template <> struct Group<1> {
    static DOPF_HD int lane() { return 0; }
    static DOPF_HD double sum(double v) { return v; }
    static DOPF_HD double scan_incl(double v) { return v; }
    static DOPF_HD double bcast(double v, int) { return v; }
    static DOPF_HD unsigned ballot(bool p) { return p ? 1u : 0u; }
    static DOPF_HD void sync() {}
};

#if defined(__CUDACC__)
template <> struct Group<32> {
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ double sum(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    static __device__ __forceinline__ double scan_incl(double v)
    {
        const int l = threadIdx.x & 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(0xffffffffu, v, o);
            if (l >= o) v += u;
        }
        return v;
    }
    static __device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    static __device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
};
#endif

DOPF_HD int first_bit(unsigned m)
{
#if defined(__CUDA_ARCH__)
    return __ffs(m) - 1;
#else
    for (int i = 0; i < 32; ++i) if (m & (1u << i)) return i;
    return -1;
#endif
}

// Per-storage problem view.  `step[t]` lives in shared memory (device) or a plain array (host);
// `hinges`/`hcnt` are null in the predict pass (anchor linearisation only).
struct StoProblem {
    int T;
    StoConst k;
    const StoStep *step;
    const Hinge *hinges;   // [T][hcap] or null
    const int *hcnt;       // [T] or null
    int hcap;
    DOPF_HD HingeList list(int t) const
    {
        HingeList l;
        l.h = hinges ? hinges + (size_t)t * hcap : nullptr;
        l.n = hinges ? hcnt[t] : 0;
        l.sorted = false;
        return l;
    }
};

struct StoStats { int evals; int solves; int segments; int passes; };

template <int W>
struct StoSolver {
    typedef Group<W> G;
    const StoProblem &p;
    double tolE;
    StoStats stats;

    DOPF_HD StoSolver(const StoProblem &pp) : p(pp)
    {
        tolE = 1e-9 * (p.k.emax > 1.0 ? p.k.emax : 1.0);
        stats.evals = stats.solves = stats.segments = stats.passes = 0;
    }

    // sum_{t in [t0,t1)} y_t(eta) and its derivative
    DOPF_HD void seg_eval(double eta, int t0, int t1, double &s, double &ds)
    {
        double a = 0.0, b = 0.0;
        for (int t = t0 + G::lane(); t < t1; t += W) {
            StoEval e = sto_eval(p.step[t], p.k, p.list(t), eta);
            a += e.C - e.D; b += e.dy;
        }
        s = G::sum(a); ds = G::sum(b);
        stats.evals += t1 - t0; stats.passes++;
    }

    // eta with sum_{[t0,t1)} y(eta) = target   (sum is non-increasing in eta)
    DOPF_HD double seg_solve(double target, int t0, int t1, double eta)
    {
        stats.solves++;
        double lo = -INFINITY, hi = INFINITY;   // sum(lo) > target > sum(hi)
        double rlo = 0.0, rhi = 0.0;
        double step = 1.0;
        const double tolS = 1e-13 * (1.0 + fabs(target) + p.k.pmax);
        for (int it = 0; it < 300; ++it) {
            double s, ds;
            seg_eval(eta, t0, t1, s, ds);
            const double r = s - target;
            if (fabs(r) <= tolS) return eta;
            if (r > 0.0) { lo = eta; rlo = r; } else { hi = eta; rhi = r; }
            double en;
            if (ds < -1e-300) en = eta - r / ds;
            else { en = r > 0.0 ? eta + step : eta - step; step *= 4.0; }
            if (!(en > lo && en < hi)) {
                if (isfinite(lo) && isfinite(hi)) {
                    en = lo - rlo * (hi - lo) / (rhi - rlo);
                    if (!(en > lo && en < hi)) en = 0.5 * (lo + hi);
                } else {
                    en = r > 0.0 ? eta + step : eta - step; step *= 4.0;
                }
            }
            if (isfinite(lo) && isfinite(hi) && (hi - lo) <= 1e-15 * (1.0 + fabs(lo))) return en;
            eta = en;
        }
        return eta;
    }

    // first t in [t0,t1) whose level leaves [0,emax] when eta is used from t0 on (level e0
    // before t0).  returns t (or -1) and kind = +1 (above emax) / -1 (below 0).
    DOPF_HD int scan(double eta, int t0, double e0, int t1, int &kind)
    {
        double carry = e0;
        for (int base = t0; base < t1; base += W) {
            const int t = base + G::lane();
            double y = 0.0;
            if (t < t1) {
                StoEval e = sto_eval(p.step[t], p.k, p.list(t), eta);
                y = e.C - e.D;
            }
            const double E = carry + G::scan_incl(y);
            const bool up = (t < t1) && (E > p.k.emax + tolE);
            const bool dn = (t < t1) && (E < -tolE);
            const unsigned m = G::ballot(up || dn);
            if (m) {
                const int f = first_bit(m);
                const double Ef = G::bcast(E, f);
                kind = Ef > p.k.emax ? 1 : -1;
                stats.evals += (base - t0) + W;
                return base + f;
            }
            carry = G::bcast(E, W - 1);
        }
        stats.evals += t1 - t0;
        kind = 0;
        return -1;
    }

    // Planning-horizon ("funnel") solve; writes the multiplier path eta_out[t].
    DOPF_HD void solve(double *eta_out)
    {
        const int T = p.T;
        int t0 = 0;
        double e0 = 0.0;
        while (t0 < T) {
            stats.segments++;
            double eta = 0.0;       // free end: no level constraint binds after the horizon
            int dir = 0;            // +1: eta was raised (upper level bound met at tauA), -1: lowered
            int tauA = -1;
            int end = T;
            double e_next = 0.0;
            for (int guard = 0; guard < 4 * T + 8; ++guard) {
                int kind;
                const int tv = scan(eta, t0, e0, T, kind);
                if (tv < 0) {
                    if (dir != 0) { end = tauA + 1; e_next = dir > 0 ? p.k.emax : 0.0; }
                    break;
                }
                if (dir != 0 && kind != dir) {  // opposite bound hit later: close at the anchor
                    end = tauA + 1; e_next = dir > 0 ? p.k.emax : 0.0;
                    break;
                }
                const double eta2 = seg_solve((kind > 0 ? p.k.emax : 0.0) - e0, t0, tv + 1, eta);
                int kind2;
                int tw = scan(eta2, t0, e0, tv, kind2);
                if (tw < 0) { eta = eta2; dir = kind; tauA = tv; continue; }
                // conflict between the bound at tv and the opposite bound before it
                double etac = eta2;
                int tauB = tw, kB = kind2;
                for (int g2 = 0; g2 < T + 4 && tw >= 0; ++g2) {
                    etac = seg_solve((kind2 > 0 ? p.k.emax : 0.0) - e0, t0, tw + 1, etac);
                    tauB = tw; kB = kind2;
                    tw = scan(etac, t0, e0, tv, kind2);
                }
                eta = etac; end = tauB + 1; e_next = kB > 0 ? p.k.emax : 0.0;
                break;
            }
            for (int t = t0 + G::lane(); t < end; t += W) eta_out[t] = eta;
            G::sync();
            t0 = end; e0 = e_next;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Sequential planning-horizon solve (one thread per storage).  Walks t once, keeping the
// interval [lo,hi] of level multipliers eta for which the levels of the current segment stay
// inside [0,emax]; lo/hi are tightened by exact 1-D solves when a bound is met and the segment
// is closed at the binding time when the interval becomes empty (Modigliani-Hohn).
// `emit(t, eta)` is called once per timestep in increasing t with its final multiplier.
// `Steps` provides step(t) -> StoStep and list(t) -> HingeList.
// ---------------------------------------------------------------------------------------------
template <class Steps, class Emit>
DOPF_HD void sto_funnel_seq(const Steps &sp, const StoConst &k, int T, Emit &emit, StoStats *stats = nullptr)
{
    const double tolE = 1e-9 * (k.emax > 1.0 ? k.emax : 1.0);
    const double BIG = 1e300;
    int nev = 0;
    auto yof = [&](int t, double eta, double &dy) -> double {
        if (eta >= BIG) { dy = 0.0; return -k.pmax; }
        if (eta <= -BIG) { dy = 0.0; return k.pmax; }
        StoEval e = sto_eval(sp.step(t), k, sp.list(t), eta);
        dy = e.dy; ++nev;
        return e.C - e.D;
    };
    // eta in [a,b] (a<b; sum(a) >= target >= sum(b)) with sum_{t0..t1} y(eta) = target
    auto solve = [&](int t0, int t1, double target, double a, double b) -> double {
        double ra = 0.0, rb = 0.0;
        bool fa = false, fb = false;
        double eta = (a > -BIG) ? a : ((b < BIG) ? b : 0.0);
        if (a > -BIG && b < BIG) eta = 0.5 * (a + b);
        double step = 1.0;
        const double tolS = 1e-13 * (1.0 + fabs(target) + k.pmax);
        if (stats) stats->solves++;
        for (int it = 0; it < 300; ++it) {
            double s = 0.0, ds = 0.0, dy;
            if (stats) stats->passes++;
            for (int t = t0; t <= t1; ++t) { s += yof(t, eta, dy); ds += dy; }
            const double r = s - target;
            if (fabs(r) <= tolS) return eta;
            if (r > 0.0) { a = eta; ra = r; fa = true; } else { b = eta; rb = r; fb = true; }
            double en;
            if (ds < -1e-300) en = eta - r / ds;
            else {
                // flat: every step of the segment is saturated at eta -> go to the nearest clip
                // breakpoint in the required direction (exact for hinge-free steps)
                const bool up = r > 0.0;
                double best = up ? BIG : -BIG;
                bool ok = true;
                for (int t = t0; t <= t1 && ok; ++t) {
                    if (sp.list(t).n != 0) { ok = false; break; }
                    const double e2 = sto_next_break(sp.step(t), k, eta, up);
                    if (up ? e2 < best : e2 > best) best = e2;
                }
                if (ok && fabs(best) < BIG) en = best + (up ? 1.0 : -1.0) * 1e-11 * (1.0 + fabs(best));
                else { en = up ? eta + step : eta - step; step *= 4.0; }
            }
            if (!(en > a && en < b)) {
                if (fa && fb) { en = a - ra * (b - a) / (rb - ra); if (!(en > a && en < b)) en = 0.5 * (a + b); }
                else if (a > -BIG && b < BIG) en = 0.5 * (a + b);
                else { en = r > 0.0 ? eta + step : eta - step; step *= 4.0; }
            }
            if (a > -BIG && b < BIG && (b - a) <= 1e-15 * (1.0 + fabs(a))) return en;
            eta = en;
        }
        return eta;
    };

    int t0 = 0;
    double e0 = 0.0;
    while (t0 < T) {
        double lo = -BIG, hi = BIG;     // feasible multiplier interval of the segment starting at t0
        int tlo = -1, thi = -1;         // times at which lo / hi were fixed (level = emax / 0 there)
        double Elo = e0, Ehi = e0;      // levels along the lo / hi paths
        int end = -1; double eta = 0.0, e_next = 0.0;
        for (int t = t0; t < T && end < 0; ++t) {
            double dy;
            Elo += yof(t, lo, dy);
            Ehi += yof(t, hi, dy);
            if (Elo > k.emax + tolE) {
                if (Ehi > k.emax + tolE) { end = thi; eta = hi; e_next = 0.0; break; }   // empty: lower anchor binds
                lo = solve(t0, t, k.emax - e0, lo, hi); tlo = t; Elo = k.emax;
            }
            if (Ehi < -tolE) {
                if (Elo < -tolE) { end = tlo; eta = lo; e_next = k.emax; break; }        // empty: upper anchor binds
                hi = solve(t0, t, 0.0 - e0, lo, hi); thi = t; Ehi = 0.0;
            }
        }
        if (end < 0) {                  // horizon reached: multiplier closest to 0 (free end)
            if (lo > 0.0) { end = tlo; eta = lo; e_next = k.emax; }
            else if (hi < 0.0) { end = thi; eta = hi; e_next = 0.0; }
            else { end = T - 1; eta = 0.0; }
        }
        for (int t = t0; t <= end; ++t) emit(t, eta);
        t0 = end + 1; e0 = e_next;
        if (stats) stats->segments++;
    }
    if (stats) stats->evals += nev;
}

// ---------------------------------------------------------------------------------------------
// Warm start: re-use the active set of the previous solve.  The anchors are the timesteps whose
// previous level `E_prev(t)` sat at a bound (0 or emax); between two anchors the multiplier is
// constant.  Each run is re-solved with its end level fixed (Newton from the previous multiplier
// `eta_prev`, safeguarded), then the KKT conditions are verified: levels inside [0,emax] and a
// multiplier path with the right sign at every anchor (eta may only drop after a full storage,
// only rise after an empty one, and must end at 0 / >=0 / <=0).  Because saturated steps make the
// multiplier of a run non-unique, each run carries the interval of multipliers that reproduce its
// solution and the sign conditions are checked on intervals.  Returns true and has emitted the
// exact solution if everything holds; otherwise the caller runs the cold funnel.
// ---------------------------------------------------------------------------------------------
#if !defined(__CUDACC__) && defined(DOPF_DEBUG_WARM)
static int g_warm_fail[8];
#define WFAIL(i) (g_warm_fail[i]++, false)
#else
#define WFAIL(i) false
#endif
template <class Steps, class PrevEta, class PrevE, class Emit>
DOPF_HD bool sto_warm_try(const Steps &sp, const StoConst &k, int T, const PrevEta &eta_prev, const PrevE &E_prev,
                          Emit &emit, StoStats *stats = nullptr)
{
    const double tolE = 1e-9 * (k.emax > 1.0 ? k.emax : 1.0);
    const double tolA = 1e-7 * (k.emax > 1.0 ? k.emax : 1.0);   // "was at a bound" in the previous solve
    const double BIG = 1e300;
    int nev = 0;
    int a = 0;
    double e0 = 0.0;
    double Flo = -BIG, Fhi = BIG;   // multipliers the previous run can take
    int kind_last = 0;              // bound that closed the previous run: +1 emax, -1 zero
    while (a < T) {
        int b = a, kind = 0;
        for (;; ++b) {
            const double Ep = E_prev(b);
            if (Ep >= k.emax - tolA) { kind = 1; break; }
            if (Ep <= tolA) { kind = -1; break; }
            if (b == T - 1) break;
        }
        const double target = (kind > 0 ? k.emax : 0.0) - e0;
        double eta = kind == 0 ? 0.0 : eta_prev(b);
        double lo = -BIG, hi = BIG, rlo = 0.0, rhi = 0.0, Emin = 0.0, Emax = 0.0, ds = 0.0;
        bool done = false, flo = false, fhi = false, nohinge = true;
        const double tolS = 1e-13 * (1.0 + fabs(target) + k.pmax);
        for (int it = 0; it < 40 && !done; ++it) {
            double sum = 0.0;
            ds = 0.0; Emin = BIG; Emax = -BIG;
            for (int t = a; t <= b; ++t) {
                StoEval e = sto_eval(sp.step(t), k, sp.list(t), eta);
                ++nev;
                sum += e.C - e.D; ds += e.dy;
                const double E = e0 + sum;
                Emin = E < Emin ? E : Emin; Emax = E > Emax ? E : Emax;
            }
            if (kind == 0) { done = true; break; }          // free end: eta stays 0
            const double r = sum - target;
            if (fabs(r) <= tolS) { done = true; break; }
            if (r > 0.0) { lo = eta; rlo = r; flo = true; } else { hi = eta; rhi = r; fhi = true; }
            double en;
            if (ds < -1e-300) en = eta - r / ds;
            else {
                // flat: every step saturated at eta -> nearest clip breakpoint in the needed direction
                const bool up = r > 0.0;
                double best = up ? BIG : -BIG;
                for (int t = a; t <= b; ++t) {
                    if (sp.list(t).n != 0) return WFAIL(0);
                    const double e2 = sto_next_break(sp.step(t), k, eta, up);
                    if (up ? e2 < best : e2 > best) best = e2;
                }
                if (!(fabs(best) < BIG)) return WFAIL(0);   // target unreachable: active set changed
                en = best + (up ? 1.0 : -1.0) * 1e-11 * (1.0 + fabs(best));
            }
            if (!(en > lo && en < hi)) {
                if (!(flo && fhi)) return WFAIL(1);
                en = lo - rlo * (hi - lo) / (rhi - rlo);
                if (!(en > lo && en < hi)) en = 0.5 * (lo + hi);
                if ((hi - lo) <= 1e-15 * (1.0 + fabs(lo))) { eta = en; done = true; break; }
            }
            eta = en;
        }
        if (!done) return WFAIL(2);
        if (Emin < -tolE || Emax > k.emax + tolE) return WFAIL(3);
        // multipliers that give the same run solution: a point, or the flat stretch around eta
        double Ilo = eta, Ihi = eta;
        if (!(ds < -1e-300)) {
            for (int t = a; t <= b && nohinge; ++t) nohinge = sp.list(t).n == 0;
            if (nohinge) {
                Ilo = -BIG; Ihi = BIG;
                for (int t = a; t <= b; ++t) {
                    const StoStep st = sp.step(t);
                    const StoEval e = sto_eval(st, k, sp.list(t), eta);
                    double l2, h2;
                    sto_flat_interval(st, k, eta, e.D, e.C, l2, h2);
                    Ihi = h2 < Ihi ? h2 : Ihi; Ilo = l2 > Ilo ? l2 : Ilo;
                }
            }
        }
        if (kind_last > 0) Ihi = Ihi < Fhi ? Ihi : Fhi;     // eta may not rise after a full storage
        if (kind_last < 0) Ilo = Ilo > Flo ? Ilo : Flo;     // eta may not drop after an empty storage
        if (b == T - 1) {                                   // end of horizon: eta_{T+1} = 0
            if (kind > 0) Ilo = Ilo > 0.0 ? Ilo : 0.0;
            if (kind < 0) Ihi = Ihi < 0.0 ? Ihi : 0.0;
            if (kind == 0 && (Ilo > 0.0 || Ihi < 0.0)) return WFAIL(6);
        }
        if (Ilo > Ihi) return WFAIL(4);
        const double eo = eta < Ilo ? Ilo : (eta > Ihi ? Ihi : eta);
        for (int t = a; t <= b; ++t) emit(t, eo);
        if (kind != 0) e0 = kind > 0 ? k.emax : 0.0;
        Flo = Ilo; Fhi = Ihi; kind_last = kind;
        a = b + 1;
        if (stats) stats->segments++;
    }
    if (stats) stats->evals += nev;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Shared per-(line,t) quantities (row preparation) - see DESIGN.md section 3.
// ---------------------------------------------------------------------------------------------
struct Coef {
    double gamma, w2, kk, kappa, beta, g2w;   // w2=2w, kk=2w+gamma, kappa=2w*gamma/kk, beta=4w^2/kk, g2w=gamma/(2w)
    double prox, mask_tol, eps;
    DOPF_HD static Coef make(double gamma, double w, double prox, double mask_tol, double eps)
    {
        Coef c;
        c.gamma = gamma; c.w2 = 2.0 * w; c.kk = c.w2 + gamma;
        c.kappa = c.w2 * gamma / c.kk; c.beta = c.w2 * c.w2 / c.kk; c.g2w = gamma / c.w2;
        c.prox = prox; c.mask_tol = mask_tol; c.eps = eps;
        return c;
    }
};

struct RowPrep { double bplus, bminus, M, Wt; };

DOPF_HD RowPrep row_prep(const Coef &c, double fmax, double F, double U, double K, double mu, double rho)
{
    RowPrep r;
    r.bplus = (fmax - F) + c.g2w * U;
    r.bminus = (fmax + F) + c.g2w * K;
    const bool mp = r.bplus < 0.0, mm = r.bminus < 0.0;
    r.M = mu - rho + c.kappa * (U - K + 2.0 * F) + c.beta * ((mm ? r.bminus : 0.0) - (mp ? r.bplus : 0.0));
    r.Wt = c.beta * ((mp ? 1.0 : 0.0) + (mm ? 1.0 : 0.0));
    return r;
}

// hinge of (line l, side) seen from a node with PTDF entry p; returns false if p == 0
DOPF_HD bool make_hinge(const Coef &c, double p, double b, int side /*0: upper(U), 1: lower(K)*/, Hinge &h)
{
    if (p == 0.0) return false;
    const double s = c.beta * p * p;
    if (side == 0) { h.bp = b / p; h.sg = (p > 0.0 ? s : -s); }
    else { h.bp = -b / p; h.sg = (p > 0.0 ? -s : s); }
    return true;
}

// exact positive-part sums for the average slacks: avgU = (2w/(kk*A)) * sum_i (bplus - p*delta_i)_+ ,
// avgK = (2w/(kk*A)) * sum_i (bminus + p*delta_i)_+
DOPF_HD double pospart(double v) { return v > 0.0 ? v : 0.0; }

// monotone map double -> uint64 for atomicMax on non-negative doubles
DOPF_HD unsigned long long nonneg_bits(double v)
{
    union { double d; unsigned long long u; } x;
    x.d = v < 0.0 ? 0.0 : v;
    return x.u;
}
DOPF_HD double bits_nonneg(unsigned long long u)
{
    union { double d; unsigned long long u; } x;
    x.u = u;
    return x.d;
}

}  // namespace dopf
#endif
