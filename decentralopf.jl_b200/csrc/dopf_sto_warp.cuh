// Warp-parallel storage solve (one warp per storage; lane l owns the J consecutive timesteps
// l*J .. l*J+J-1, loads stay fully coalesced because the warp covers one contiguous range).
//
// Active-set method on the level bounds 0 <= E_t <= emax (reference problem:
// /root/reference/src/optimization/subproblems.jl:107-207 after slack elimination):
//   * anchors  = timesteps whose level is fixed at a bound (kind +1: emax, -1: 0); between two anchors
//     ("run") the level multiplier eta is constant;
//   * all runs are solved at once: every lane evaluates y_t(eta) for its timesteps, run sums come from
//     segmented warp scans, the run's last timestep ("tail") does one safeguarded Newton update per pass
//     and broadcasts the new multiplier back over its run;
//   * then the KKT conditions are checked in parallel: levels inside the bounds, and a multiplier
//     path with the right sign at every anchor (intervals, because saturated runs have a non-unique
//     multiplier).  Violated levels add an anchor (first violated timestep of the run), anchors with a
//     wrong sign are dropped, and the solve repeats;
//   * the result is accepted only when the KKT conditions hold (=> exact optimum of the convex
//     problem).  Storages that do not verify within the caps go to the sequential exact solver.
// The initial active set comes from the previous levels (warm start).
//
// Cost structure: (a) a segmented scan is an in-register pass over the lane's J elements plus ONE
// 32-wide cross-lane scan of the lane aggregates; (b) everything of a timestep that does not depend
// on eta - the sorted clip breakpoints of D(nu), C(nu), the eta-thresholds at which nu(eta) passes them
// and the inverse slopes in between - is tabulated once per storage in shared memory, so an evaluation
// is four comparisons, one interpolation (no division) and one clip.
#ifndef DOPF_STO_WARP_CUH
#define DOPF_STO_WARP_CUH

#include "dopf_bodies.h"
#ifdef DOPF_STATS
#include <cstdio>
#endif

namespace dopf {

constexpr unsigned FULL = 0xffffffffu;
constexpr double WBIG = 1e300;

struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return a < b ? a : b; } };
struct OpMax { __device__ double operator()(double a, double b) const { return a > b ? a : b; } };

// lane-level reach for a set of segment heads: how many lanes back a lane may combine (one ballot: the nearest lane at or
// below mine that holds a head)
template <int J>
__device__ __forceinline__ int lane_reach_back(const bool (&head)[J])
{
    const int lane = threadIdx.x & 31;
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) any |= head[j];
    const unsigned m = __ballot_sync(FULL, any) & (0xffffffffu >> (31 - lane));      // lanes <= mine with a head
    return m ? lane - (31 - __clz(m)) : lane + 1;                                      // 0 if this lane holds a head
}
template <int J>
__device__ __forceinline__ int lane_reach_fwd(const bool (&tail)[J])
{
    const int lane = threadIdx.x & 31;
    bool any = false;
#pragma unroll
    for (int j = 0; j < J; ++j) any |= tail[j];
    const unsigned m = __ballot_sync(FULL, any) & (0xffffffffu << lane);              // lanes >= mine with a tail
    return m ? (__ffs(m) - 1) - lane : 0;                                              // 0 if this lane holds a tail (or none follows)
}

// largest reach of any lane: scan levels beyond it change nothing and are skipped (runs between level anchors are short -
// about half of the timesteps of a storage are anchored in the steady state - so most rounds need 1-3 of the 5 levels)
__device__ __forceinline__ int warp_max_reach(int r) { return (int)__reduce_max_sync(FULL, (unsigned)r); }

// forward inclusive segmented scan of two arrays at once (segments start at head[j])
template <int J, class OpA, class OpB>
__device__ __forceinline__ void seg_fwd2(double (&a)[J], double (&b)[J], const bool (&head)[J], int lreach, int mreach,
                                         OpA opa, OpB opb, double ida, double idb)
{
    const int lane = threadIdx.x & 31;
    double ra = ida, rb = idb;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (head[j]) { ra = a[j]; rb = b[j]; } else { ra = opa(ra, a[j]); rb = opb(rb, b[j]); }
        a[j] = ra; b[j] = rb;
    }
    double xa = ra, xb = rb;                      // aggregate of the segment that is open at the lane end
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        if (mreach < o) break;
        const double ya = __shfl_up_sync(FULL, xa, o), yb = __shfl_up_sync(FULL, xb, o);
        if (lreach >= o) { xa = opa(ya, xa); xb = opb(yb, xb); }
    }
    double ca = __shfl_up_sync(FULL, xa, 1), cb = __shfl_up_sync(FULL, xb, 1);
    if (lane == 0) { ca = ida; cb = idb; }
    bool open = true;                             // elements before the lane's first head continue the incoming segment
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (head[j]) open = false;
        if (open) { a[j] = opa(ca, a[j]); b[j] = opb(cb, b[j]); }
    }
}
// forward inclusive segmented scan of one array
template <int J, class OpA>
__device__ __forceinline__ void seg_fwd1(double (&a)[J], const bool (&head)[J], int lreach, int mreach, OpA opa, double ida)
{
    const int lane = threadIdx.x & 31;
    double ra = ida;
#pragma unroll
    for (int j = 0; j < J; ++j) { ra = head[j] ? a[j] : opa(ra, a[j]); a[j] = ra; }
    double xa = ra;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { if (mreach < o) break; const double ya = __shfl_up_sync(FULL, xa, o); if (lreach >= o) xa = opa(ya, xa); }
    double ca = __shfl_up_sync(FULL, xa, 1);
    if (lane == 0) ca = ida;
    bool open = true;
#pragma unroll
    for (int j = 0; j < J; ++j) { if (head[j]) open = false; if (open) a[j] = opa(ca, a[j]); }
}
// every element takes the value held by the tail of its run (one array of a 32-bit or 64-bit type)
template <int J, class TB>
__device__ __forceinline__ void seg_take_tail1(TB (&b)[J], const bool (&tail)[J], int lreach, int mreach)
{
    const int lane = threadIdx.x & 31;
    TB fb = TB();
    {
        TB cb = TB(); bool seen = false;
#pragma unroll
        for (int j = J - 1; j >= 0; --j) { if (tail[j]) { cb = b[j]; seen = true; } else if (seen) b[j] = cb; }
        fb = cb;
    }
    TB xb = fb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { if (mreach < o) break; const TB yb = __shfl_down_sync(FULL, xb, o); if (lreach >= o) xb = yb; }
    const TB cb = __shfl_down_sync(FULL, xb, 1);
    bool seen = false;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) { if (tail[j]) seen = true; if (!seen && lane < 31) b[j] = cb; }
}
template <int J>
__device__ __forceinline__ void seg_fwd_min_int(int (&a)[J], const bool (&head)[J], int lreach, int mreach)
{
    const int lane = threadIdx.x & 31;
    int r = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < J; ++j) { r = head[j] ? a[j] : min(r, a[j]); a[j] = r; }
    int x = r;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { if (mreach < o) break; const int y = __shfl_up_sync(FULL, x, o); if (lreach >= o) x = min(y, x); }
    int c = __shfl_up_sync(FULL, x, 1);
    if (lane == 0) c = 0x7fffffff;
    bool open = true;
#pragma unroll
    for (int j = 0; j < J; ++j) { if (head[j]) open = false; if (open) a[j] = min(c, a[j]); }
}
// every element takes the values held by the tail of its run
template <int J, class TB>
__device__ __forceinline__ void seg_take_tail(double (&a)[J], TB (&b)[J], const bool (&tail)[J], int lreach, int mreach)
{
    const int lane = threadIdx.x & 31;
    double fa = 0.0; TB fb = TB();               // values at the lane's FIRST tail (what lanes to the left need)
    {
        double ca = 0.0; TB cb = TB(); bool seen = false;
#pragma unroll
        for (int j = J - 1; j >= 0; --j) {
            if (tail[j]) { ca = a[j]; cb = b[j]; seen = true; } else if (seen) { a[j] = ca; b[j] = cb; }
        }
        fa = ca; fb = cb;                        // after the loop: the left-most tail of the lane (if any)
    }
    double xa = fa; TB xb = fb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        if (mreach < o) break;
        const double ya = __shfl_down_sync(FULL, xa, o); const TB yb = __shfl_down_sync(FULL, xb, o);
        if (lreach >= o) { xa = ya; xb = yb; }
    }
    const double ca = __shfl_down_sync(FULL, xa, 1); const TB cb = __shfl_down_sync(FULL, xb, 1);
    bool seen = false;                           // elements to the right of the lane's last tail take the carry
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        if (tail[j]) seen = true;
        if (!seen && lane < 31) { a[j] = ca; b[j] = cb; }
    }
}
template <int J, class T>
__device__ __forceinline__ void shift_from_prev(const T (&v)[J], T (&out)[J], T first)
{
    const int lane = threadIdx.x & 31;
    const T up = __shfl_up_sync(FULL, v[J - 1], 1);
    out[0] = lane == 0 ? first : up;
#pragma unroll
    for (int j = 1; j < J; ++j) out[j] = v[j - 1];
}
template <int J, class T>
__device__ __forceinline__ void shift_from_next(const T (&v)[J], T (&out)[J], T last)
{
    const int lane = threadIdx.x & 31;
    const T dn = __shfl_down_sync(FULL, v[0], 1);
    out[J - 1] = lane == 31 ? last : dn;
#pragma unroll
    for (int j = 0; j < J - 1; ++j) out[j] = v[j + 1];
}
template <int J>
__device__ __forceinline__ bool any_of(const bool (&p)[J])
{
    bool a = false;
#pragma unroll
    for (int j = 0; j < J; ++j) a |= p[j];
    return __any_sync(FULL, a);
}

// ---- per-timestep clip table (independent of eta) ---------------------------------------------------------
// Psi(nu) = nu - (g0 - eta) - s1*delta(nu) is piecewise linear with the four clip breakpoints bb[0..3] of
// D(nu), C(nu); Psi(bb[i]) = eta - e[i] with the eta-thresholds e[i] = g0 + s1*dl[i] - bb[i] (non-increasing in
// i, dl[i] = delta at bb[i]).  The thresholds cut the eta-axis into five pieces p = #{i : eta < e[i]}; on piece p
// nu is affine in eta, nu = NU0[p] + NUS[p]*eta, and dy/deta = DY[p] is constant.  On a piece with nf free variables
// Psi has slope 1 + nf*s1/prox, whose inverse is prox*r_nf with r_nf = 1/(prox + nf*s1) from k_node_prep: no division
// here.  An evaluation is: four comparisons, three table reads by piece index, one fma, two clips.  Components
// (component-major in shared memory, comp c of timestep t at tab[c*tstride + t], conflict-free):
//   0-3 e | 4-8 NU0 | 9-13 NUS | 14-18 DY
constexpr int TAB_E = 0, TAB_NU0 = 4, TAB_NUS = 9, TAB_DY = 14, TAB_COMPS = 19;

__device__ __forceinline__ void clip_tab_build(const StoStep &st, const StoConst &k, double r1, double r2, double *tab, int tstride, int t)
{
    double e[4], nu0[5], nus[5], dyv[5];
    sto_clip_table(st, k, r1, r2, e, nu0, nus, dyv);          // closed form, dopf_math.h
#pragma unroll
    for (int i = 0; i < 4; ++i) tab[(TAB_E + i) * tstride + t] = e[i];
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        tab[(TAB_NU0 + p) * tstride + t] = nu0[p]; tab[(TAB_NUS + p) * tstride + t] = nus[p]; tab[(TAB_DY + p) * tstride + t] = dyv[p];
    }
}
// same result as sto_eval() for a hinge-free step
__device__ __forceinline__ StoEval eval_tab(const StoStep &st, const StoConst &k, const double *tab, int tstride, int t, double eta)
{
    const double e0 = tab[(TAB_E + 0) * tstride + t], e1 = tab[(TAB_E + 1) * tstride + t];
    const double e2 = tab[(TAB_E + 2) * tstride + t], e3 = tab[(TAB_E + 3) * tstride + t];
    const int p = (eta < e0) + (eta < e1) + (eta < e2) + (eta < e3);
    const double *q = tab + p * tstride + t;
    const double nu = q[TAB_NU0 * tstride] + q[TAB_NUS * tstride] * eta;
    StoEval r;
    r.D = clip01(st.Db - (k.mc + nu) * k.iprox, k.pmax);
    r.C = clip01(st.Cb - (k.mc - nu) * k.iprox, k.pmax);
    r.dy = q[TAB_DY * tstride];
    return r;
}
// nearest eta-breakpoint strictly beyond eta (same as sto_next_break)
__device__ __forceinline__ void next_breaks_tab(const double *tab, int tstride, int t, double eta, double &up, double &dn)
{
    up = WBIG; dn = -WBIG;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double e = tab[(TAB_E + i) * tstride + t];
        if (e > eta && e < up) up = e;
        if (e < eta && e > dn) dn = e;
    }
}

// clipped step with hinges: D, C, delta and therefore the hinge force are constant while nu stays inside the range
// [plo, phi] on which both variables keep their clip state, so eta may move by (nu - plo) upwards and (phi - nu)
// downwards without changing y_t.  The range is derived from the clip STATE of D and C, not from the position of nu:
// on a saturated run the Newton update lands on a clip breakpoint, and locating nu among the breakpoints would
// pick the neighbouring (free) piece as often as the flat one.
__device__ __forceinline__ void flat_range_hinge(const StoStep &st, const StoConst &k, const HingeList &hl, double hmin,
                                                 double eta, double D, double C, double &elo, double &ehi)
{
    const double dl = (D - st.Db) - (C - st.Cb);
    double hv = 0.0, hs = 0.0;
    if (fabs(dl) >= hmin) hl.eval(dl, hv, hs);
    const double nu = st.g0 - eta + st.s1 * dl + hv;
    double plo = -WBIG, phi = WBIG;
    if (k.pmax > 0.0) {
        // D = clip(Db - (mc + nu)/prox) falls with nu, C = clip(Cb - (mc - nu)/prox) rises with nu
        if (D <= 0.0) plo = fmax(plo, k.prox * st.Db - k.mc);                     // D stays 0 for nu above
        else if (D >= k.pmax) phi = fmin(phi, k.prox * (st.Db - k.pmax) - k.mc);  // D stays pmax for nu below
        else { plo = nu; phi = nu; }                                              // free variable: no slack at all
        if (C <= 0.0) phi = fmin(phi, k.mc - k.prox * st.Cb);                     // C stays 0 for nu below
        else if (C >= k.pmax) plo = fmax(plo, k.mc + k.prox * (k.pmax - st.Cb));  // C stays pmax for nu above
        else { plo = nu; phi = nu; }
    }
    ehi = plo > -WBIG ? eta + fmax(nu - plo, 0.0) : WBIG;
    elo = phi < WBIG ? eta - fmax(phi - nu, 0.0) : -WBIG;
}

// run status bits broadcast from the tail
enum { RS_CONV = 1, RS_BAD = 2, RS_FLAT = 4, RS_EMPTY = 8, RS_FREEBAD = 16, RS_ENDBAD = 32 };

// shared memory per warp: the clip table (19 doubles per timestep)
__host__ __device__ inline size_t sto_warp_smem_per_warp(int T) { return (size_t)TAB_COMPS * (size_t)((T + 1) | 1) * sizeof(double); }

// returns true if the storage was solved and written; false => caller queues it for the exact sequential solver
// AV ("all valid"): the horizon fills the warp exactly (T = 32 J), the validity predicates fold away
template <int J, bool HINGES, bool AV = false>
__device__ bool sto_warp_solve(const View &v, int s, const Hinge *hinges, const int *hcnt, double *tab)
{
    const int lane = threadIdx.x & 31, T = v.T;
#ifdef DOPF_STATS
    unsigned long long st_t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_t0));
#endif
    const int tstride = (T + 1) | 1;
    const int cur = v.ctrl->cur, nxt = 1 - cur, n = v.sto_node[s];
    StoConst k;
    k.mc = v.sto_mc[s]; k.pmax = v.sto_pmax[s]; k.emax = v.sto_emax[s]; k.prox = v.c.prox; k.iprox = 1.0 / v.c.prox;
    const double tolE = 1e-9 * (k.emax > 1.0 ? k.emax : 1.0), tolA = 1e-7 * (k.emax > 1.0 ? k.emax : 1.0);

    StoStep st[J];
    HingeList hl[J];
    bool valid[J];
    int kind[J];          // anchor at the end of t: +1 level = emax, -1 level = 0, 0 none
    double eta[J], hmin[J], r1[J], r2[J];
    // blocked ownership (lane owns timesteps lane*J .. lane*J+J-1): the warp reads one contiguous range per array,
    // J strided passes of 8 bytes per lane that hit the same L1 lines
    __syncwarp();
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane * J + j;
        valid[j] = AV || t < T;
        const int tt = valid[j] ? t : T - 1;
        const size_t o = (size_t)s * T + tt, on = v.nt_of(n, tt);      // n: virtual node
        st[j].Db = sel(v.D, cur)[o]; st[j].Cb = sel(v.C, cur)[o];
        st[j].g0 = v.g0[on]; st[j].s1 = v.s1[on];
        r1[j] = v.rg[on]; r2[j] = v.rg2[on];          // 1/(prox + s1), 1/(prox + 2 s1) from k_node_prep
        eta[j] = v.eta[o];
        const double Ep = v.E[o];
        hl[j].h = HINGES ? hinges + (size_t)tt * v.hcap : nullptr;
        hl[j].n = HINGES ? hcnt[tt] : 0;
        hl[j].sorted = HINGES && v.hcap <= 64;   // k_sto_fix sorts lists of up to 64 entries by |bp|
        // smallest |breakpoint| of the step: moves below it leave every hinge in its anchor state (no list access)
        hmin[j] = (HINGES && hl[j].n != 0 && hl[j].sorted) ? fabs(hl[j].h[0].bp) : 0.0;
        kind[j] = !valid[j] ? 0 : (Ep >= k.emax - tolA ? 1 : (Ep <= tolA ? -1 : 0));
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < J; ++j)                    // clip tables (overwrite the staging area)
        if (valid[j]) clip_tab_build(st[j], k, r1[j], r2[j], tab, tstride, lane * J + j);
    __syncwarp();

    double D[J], C[J], pre[J];
    bool accepted = false;
#ifdef DOPF_STATS
    int st_rounds = 0, st_passes = 0;
#define DOPF_STAT(i, n) do { if (lane == 0) atomicAdd(v.counters + (HINGES ? 16 : 0) + (i), (unsigned long long)(n)); } while (0)
#else
#define DOPF_STAT(i, n) do { } while (0)
#endif
    const int as_cap = 24 + T / 2;                 // one new anchor per run and round: long horizons need more rounds from a cold start
    for (int as_it = 0; as_it < as_cap && !accepted; ++as_it) {
        // ---- run structure from the anchors (timesteps beyond T are isolated one-element runs) ------
        bool head[J], tail[J];
        int prevk[J], endk[J];
        {
            int kp[J];
            shift_from_prev<J, int>(kind, kp, 1);                        // kind of t-1 (t = 0 starts a run)
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int t = lane * J + j;
                head[j] = kp[j] != 0 || !valid[j];
                tail[j] = !valid[j] || kind[j] != 0 || t == T - 1;
            }
        }
        const int rb = lane_reach_back<J>(head), rf = lane_reach_fwd<J>(tail);
        const int mrb = warp_max_reach(rb), mrf = warp_max_reach(rf);
        {
            // kind of the anchor that closed the previous run: heads read it at t-1 and spread it forward
            int kp[J];
            shift_from_prev<J, int>(kind, kp, 0);
            int hk[J];
#pragma unroll
            for (int j = 0; j < J; ++j) hk[j] = head[j] ? -kp[j] : 2;       // non-heads hold 2 => min = minus the head's value
            seg_fwd_min_int<J>(hk, head, rb, mrb);
#pragma unroll
            for (int j = 0; j < J; ++j) { prevk[j] = -hk[j]; endk[j] = kind[j]; }
            seg_take_tail<J, int>(eta, endk, tail, rf, mrf);                  // start multiplier and end kind of my run
        }
        double e0[J], target[J];
        bool freeend[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            e0[j] = prevk[j] > 0 ? k.emax : 0.0;
            target[j] = (endk[j] > 0 ? k.emax : 0.0) - e0[j];
            freeend[j] = endk[j] == 0;
            if (freeend[j]) eta[j] = 0.0;
        }

        // ---- simultaneous safeguarded Newton on all runs (state lives at the tails) -----------------
        double lo[J], hi[J], totd[J];
        int rs[J];
#pragma unroll
        for (int j = 0; j < J; ++j) { lo[j] = -WBIG; hi[j] = WBIG; totd[j] = 0.0; rs[j] = valid[j] ? 0 : RS_CONV; }
        bool capped = true;
        for (int it = 0; it < 24; ++it) {
            double dy[J];
#ifdef DOPF_STATS
            ++st_passes;
#endif
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (valid[j]) {
                    StoEval e = eval_tab(st[j], k, tab, tstride, lane * J + j, eta[j]);
                    if (HINGES && hl[j].n != 0) {
                        // the hinge-free solution is exact unless a hinge differs from its anchor state at this delta
                        const double dl = (e.D - st[j].Db) - (e.C - st[j].Cb);
                        if (fabs(dl) >= hmin[j]) {
                            double hv, hs;
                            hl[j].eval(dl, hv, hs);
                            if (hv != 0.0 || hs != 0.0) e = sto_eval(st[j], k, hl[j], eta[j]);
                        }
                    }
                    D[j] = e.D; C[j] = e.C; pre[j] = e.C - e.D; dy[j] = e.dy;
                } else { D[j] = C[j] = pre[j] = dy[j] = 0.0; }
            }
            seg_fwd2<J>(pre, dy, head, rb, mrb, OpAdd(), OpAdd(), 0.0, 0.0);
            bool pending[J], needflat[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                totd[j] = dy[j];
                pending[j] = false; needflat[j] = false;
                if (valid[j] && tail[j] && !(rs[j] & (RS_CONV | RS_BAD))) {
                    const double r = pre[j] - target[j];
                    const double tolS = 1e-13 * (1.0 + fabs(target[j]) + k.pmax);
                    if (freeend[j] || fabs(r) <= tolS) rs[j] |= RS_CONV;
                    else { pending[j] = true; needflat[j] = !(dy[j] < -1e-300); }
                }
            }
            if (!any_of<J>(pending)) { capped = false; break; }
            double bu[J], bd[J];
            if (any_of<J>(needflat)) {
                // saturated runs: nearest clip breakpoint of any step of the run in the needed direction;
                // every element offers its own breakpoints, the tail picks (runs that are not flat ignore it)
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    bu[j] = WBIG; bd[j] = -WBIG;
                    if (valid[j]) {
                        if (HINGES && hl[j].n != 0) {
                            double elo, ehi;
                            flat_range_hinge(st[j], k, hl[j], hmin[j], eta[j], D[j], C[j], elo, ehi);
                            bu[j] = ehi; bd[j] = elo;
                        } else next_breaks_tab(tab, tstride, lane * J + j, eta[j], bu[j], bd[j]);
                    }
                }
                seg_fwd2<J>(bu, bd, head, rb, mrb, OpMin(), OpMax(), WBIG, -WBIG);
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (!pending[j]) continue;
                const double r = pre[j] - target[j];
                if (r > 0.0) lo[j] = eta[j]; else hi[j] = eta[j];
                double en;
                if (!needflat[j]) en = eta[j] - r / dy[j];
                else {
                    const bool up = r > 0.0;
                    const double b = up ? bu[j] : bd[j];
                    if (!(fabs(b) < WBIG)) { rs[j] |= RS_BAD; continue; }               // unreachable target / hinge on a flat run
                    en = b + (up ? 1.0 : -1.0) * 1e-11 * (1.0 + fabs(b));
                }
                if (!(en > lo[j] && en < hi[j])) {
                    if (lo[j] > -WBIG && hi[j] < WBIG) {
                        en = 0.5 * (lo[j] + hi[j]);
                        if ((hi[j] - lo[j]) <= 1e-15 * (1.0 + fabs(lo[j]))) rs[j] |= RS_CONV;
                    } else { rs[j] |= RS_BAD; continue; }
                }
                eta[j] = en;
            }
            seg_take_tail<J, int>(eta, rs, tail, rf, mrf);
        }
#ifdef DOPF_STATS
        ++st_rounds;
#endif
        if (capped) { DOPF_STAT(5, 1); return false; }   // Newton cap reached
        // final run state for every element: multiplier, status, "flat" (sum of derivatives at the tail)
#pragma unroll
        for (int j = 0; j < J; ++j) if (valid[j] && tail[j] && !(totd[j] < -1e-300)) rs[j] |= RS_FLAT;
        seg_take_tail<J, int>(eta, rs, tail, rf, mrf);

        // ---- KKT check ------------------------------------------------------------------------------
        bool vio_up[J], vio_dn[J];
        double Ilo[J], Ihi[J], vmag[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double E = e0[j] + pre[j];
            const bool ok = valid[j] && !(rs[j] & RS_BAD);
            vio_up[j] = ok && E > k.emax + tolE;
            vio_dn[j] = ok && E < -tolE;
            vmag[j] = vio_up[j] ? E - k.emax : (vio_dn[j] ? -E : -1.0);
            // multiplier interval of my run, contributed per element: a point unless the whole run is flat
            Ilo[j] = -WBIG; Ihi[j] = WBIG;
            if (valid[j]) {
                if (!(rs[j] & RS_FLAT)) { Ilo[j] = eta[j]; Ihi[j] = eta[j]; }
                else if (HINGES && hl[j].n != 0) flat_range_hinge(st[j], k, hl[j], hmin[j], eta[j], D[j], C[j], Ilo[j], Ihi[j]);
                else flat_range_hinge(st[j], k, hl[j], WBIG, eta[j], D[j], C[j], Ilo[j], Ihi[j]);
            }
        }
        seg_fwd2<J>(Ilo, Ihi, head, rb, mrb, OpMax(), OpMin(), -WBIG, WBIG);          // tails now hold the run interval
        // sign chain over the runs: after an upper anchor eta may not rise, after a lower anchor it may not
        // drop.  Only tails carry run values, the other elements are neutral; the chains restart at the
        // head of a run whose previous anchor has the other kind.
        double Fhi[J], Flo[J];
        {
            bool hup[J], hdn[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                Fhi[j] = (valid[j] && tail[j]) ? Ihi[j] : WBIG;
                Flo[j] = (valid[j] && tail[j]) ? Ilo[j] : -WBIG;
                hup[j] = head[j] && prevk[j] <= 0;
                hdn[j] = head[j] && prevk[j] >= 0;
            }
            const int rup = lane_reach_back<J>(hup), rdn = lane_reach_back<J>(hdn);
            seg_fwd1<J>(Fhi, hup, rup, warp_max_reach(rup), OpMin(), WBIG);
            seg_fwd1<J>(Flo, hdn, rdn, warp_max_reach(rdn), OpMax(), -WBIG);
        }
        // new anchors: the most violated timestep of every maximal stretch of consecutive violated timesteps of one
        // kind inside a run (the level path peaks where the bound finally binds).  One anchor per run and round would
        // need as many rounds as the run has separate peaks and troughs (a daily pattern over a multi-day horizon).
        bool newanchor[J];
#pragma unroll
        for (int j = 0; j < J; ++j) newanchor[j] = false;
        {
            bool anyv[J];
#pragma unroll
            for (int j = 0; j < J; ++j) anyv[j] = vio_up[j] || vio_dn[j];
            if (any_of<J>(anyv)) {
                int vk[J], vp[J], vn[J];
#pragma unroll
                for (int j = 0; j < J; ++j) vk[j] = vio_up[j] ? 1 : (vio_dn[j] ? -1 : 0);
                shift_from_prev<J, int>(vk, vp, 0);
                shift_from_next<J, int>(vk, vn, 0);
                bool shead[J], stail[J];
#pragma unroll
                for (int j = 0; j < J; ++j) { shead[j] = head[j] || vk[j] != vp[j]; stail[j] = tail[j] || vk[j] != vn[j]; }
                const int rbs = lane_reach_back<J>(shead), rfs = lane_reach_fwd<J>(stail);
                const int mrbs = warp_max_reach(rbs), mrfs = warp_max_reach(rfs);
                double m[J];
#pragma unroll
                for (int j = 0; j < J; ++j) m[j] = vmag[j];
                seg_fwd1<J>(m, shead, rbs, mrbs, OpMax(), -1.0);               // stretch tails hold the largest violation
                seg_take_tail1<J, double>(m, stail, rfs, mrfs);
                int cand[J];
#pragma unroll
                for (int j = 0; j < J; ++j) cand[j] = (anyv[j] && vmag[j] >= m[j]) ? lane * J + j : 0x7fffffff;
                seg_fwd_min_int<J>(cand, shead, rbs, mrbs);                   // first candidate of the stretch so far
#pragma unroll
                for (int j = 0; j < J; ++j) newanchor[j] = anyv[j] && vmag[j] >= m[j] && cand[j] == lane * J + j;
            }
        }
        int flag[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane * J + j;
            flag[j] = 0;
            if (valid[j] && tail[j] && !(rs[j] & RS_BAD)) {
                const double a = Flo[j], b = Fhi[j];
                // multiplier signs up to rounding: a weakly active bound (multiplier ~ 0) must not flip between
                // "wrong sign => drop" and "violated => add" for ever
                const double tolM = 1e-10 * (1.0 + fabs(a) + fabs(b));
                if (a > b + tolM) flag[j] |= RS_EMPTY;                       // wrong sign at the anchor before this run
                else if (t == T - 1 && kind[j] != 0) {                       // end of horizon: eta_{T+1} = 0
                    // only if the chain itself is consistent the end anchor is the one to go: dropping both
                    // anchors of a one-element last run re-creates them one by one and cycles
                    if (kind[j] > 0 ? (b < -tolM) : (a > tolM)) flag[j] |= RS_ENDBAD;
                }
                if (freeend[j] && (Flo[j] > tolM || Fhi[j] < -tolM)) flag[j] |= RS_FREEBAD;
            }
            if (valid[j] && tail[j] && (rs[j] & RS_BAD)) flag[j] |= RS_BAD;
        }
        seg_take_tail1<J, int>(flag, tail, rf, mrf);                     // the run's verdict, known to all its elements

#ifdef DOPF_DEBUG_STO
        if (HINGES && (v.debug & 4) && s == ((v.debug >> 8) & 0xfff) && v.ctrl->iteration == (v.debug >> 20)) {
            for (int j = 0; j < J; ++j)
                if (valid[j]) printf("r%d t%d kind %d h%d tl%d eta %.9g D %.9g C %.9g E %.9g rs %d I[%.9g,%.9g] F[%.9g,%.9g] flag %d tv %d hn %d hmin %.6g prevk %d endk %d\n", as_it, lane * J + j, kind[j], (int)head[j], (int)tail[j], eta[j], D[j], C[j], e0[j] + pre[j], rs[j], Ilo[j], Ihi[j], Flo[j], Fhi[j], flag[j], tv[j], hl[j].n, hmin[j], prevk[j], endk[j]);
        }
#endif
        // ---- repair the active set -----------------------------------------------------------------
        bool change = false, anybad = false;
        int wantdrop[J];                                            // head of a run that wants the previous anchor gone
#pragma unroll
        for (int j = 0; j < J; ++j) {
            anybad |= valid[j] && (flag[j] != 0 || vio_up[j] || vio_dn[j]);
            wantdrop[j] = (valid[j] && head[j] && ((flag[j] & (RS_EMPTY | RS_FREEBAD)) || ((flag[j] & RS_BAD) && freeend[j]))) ? 1 : 0;
            if (valid[j] && newanchor[j]) { kind[j] = vio_up[j] ? 1 : -1; change = true; }       // (1) new anchors
        }
        int nextwant[J];
        shift_from_next<J, int>(wantdrop, nextwant, 0);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane * J + j;
            if (!valid[j] || kind[j] == 0 || newanchor[j]) continue;
            bool drop = nextwant[j] != 0;                                                       // (2) wrong sign at my anchor
            if (tail[j] && (flag[j] & RS_BAD)) drop = true;                                     //     my run cannot meet its target
            if (t == T - 1 && (flag[j] & RS_ENDBAD)) drop = true;                               //     wrong sign at the horizon end
            if (drop) { kind[j] = 0; change = true; }
        }
        anybad = __any_sync(FULL, anybad);
        change = __any_sync(FULL, change);
        if (!anybad) accepted = true;
        else if (!change) { DOPF_STAT(6, 1); return false; }
        if (accepted) {
#pragma unroll
            for (int j = 0; j < J; ++j) pre[j] += e0[j];            // levels
        }
    }
    if (!accepted) { DOPF_STAT(6, 1); return false; }
#ifdef DOPF_STATS
    {
        int nruns = 0, nsingle = 0, nanch = 0, nfree = 0;
        for (int j = 0; j < J; ++j) if (valid[j]) {
            // run structure at acceptance: a singleton run is a step whose predecessor and itself are anchors
            const int t = lane * J + j;
            nanch += kind[j] != 0;
            nfree += (D[j] > 0.0 && D[j] < k.pmax) || (C[j] > 0.0 && C[j] < k.pmax);
            (void)t;
        }
        for (int o = 16; o > 0; o >>= 1) { nanch += __shfl_xor_sync(FULL, nanch, o); nfree += __shfl_xor_sync(FULL, nfree, o); }
        DOPF_STAT(0, 1); DOPF_STAT(1, st_rounds); DOPF_STAT(2, st_passes); DOPF_STAT(3, nanch); DOPF_STAT(4, nfree);
        if (nanch >= T - 1) DOPF_STAT(7, 1);            // (almost) every step anchored
        if (nfree == 0) DOPF_STAT(8, 1);                // every step clipped
        if (st_rounds == 1) DOPF_STAT(9, 1);
        if (st_rounds == 1 && st_passes <= 2) DOPF_STAT(10, 1);
        DOPF_STAT(11, st_rounds > 8 ? 1 : 0);
        if (lane == 0) atomicMax(v.counters + (HINGES ? 16 : 0) + 12, (unsigned long long)st_rounds);
        if (lane == 0) atomicMax(v.counters + (HINGES ? 16 : 0) + 13, (unsigned long long)st_passes);
        (void)nruns; (void)nsingle;
        if (lane == 0 && !HINGES && (v.debug & 128) && (st_rounds > 8 || s >= v.S - 2 || s == v.S / 2)) {
            unsigned long long st_t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_t1));
            printf("STIME it %d s %d rounds %d t0 %llu t1 %llu dur_us %.1f\n", v.ctrl->iteration, s, st_rounds, st_t0, st_t1, (st_t1 - st_t0) * 1e-3);
        }
        if (st_rounds > 8 && lane == 0 && !HINGES && (v.debug & 256)) printf("STRAG it %d s %d rounds %d passes %d anch %d free %d pmax %g emax %g mc %g node %d\n", v.ctrl->iteration, s, st_rounds, st_passes, nanch, nfree, k.pmax, k.emax, k.mc, n);
    }
#endif

    // ---- emit (blocked, same access pattern as the loads) --------------------------------------------------
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane * J + j;
        if (!valid[j]) continue;
        const size_t o = (size_t)s * T + t;
        sel(v.D, nxt)[o] = D[j]; sel(v.C, nxt)[o] = C[j];
        v.E[o] = pre[j]; v.eta[o] = eta[j];
        note_move(v, n, t, (D[j] - st[j].Db) - (C[j] - st[j].Cb));
    }
    __syncwarp();
    return true;
}

}  // namespace dopf
#endif
