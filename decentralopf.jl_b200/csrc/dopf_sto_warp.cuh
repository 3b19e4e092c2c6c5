// Warp-parallel storage solve (one warp per storage, lane l owns timesteps l, l+32, ...).
//
// Active-set method on the level bounds 0 <= E_t <= emax (reference problem:
// /root/reference/src/optimization/subproblems.jl:107-207 after slack elimination):
//   * anchors  = timesteps whose level is fixed at a bound (kind +1: emax, -1: 0); between two anchors
//     ("run") the level multiplier eta is constant;
//   * all runs are solved at once: every lane evaluates y_t(eta) for its timesteps (sto_eval), run sums
//     come from segmented warp scans, the run's last timestep ("tail") does one safeguarded Newton
//     update per pass and broadcasts the new multiplier back over its run;
//   * then the KKT conditions are checked in parallel: levels inside the bounds, and a multiplier
//     path with the right sign at every anchor (intervals, because saturated runs have a non-unique
//     multiplier).  Violated levels add an anchor (first violated timestep of the run), anchors with a
//     wrong sign are dropped, and the solve repeats;
//   * the result is accepted only when the KKT conditions hold (=> exact optimum of the convex
//     problem).  Storages that do not verify within the caps go to the sequential exact solver.
// The initial active set comes from the previous levels (warm start).
//
// Scans: the segment structure is turned once per active-set iteration into per-element "reach"
// counters (how far the element may look towards its run head / tail inside its 32-wide chunk), so a
// segmented scan step is just shuffle + compare + op, without shuffling flags.
#ifndef DOPF_STO_WARP_CUH
#define DOPF_STO_WARP_CUH

#include "dopf_bodies.h"

namespace dopf {

constexpr unsigned FULL = 0xffffffffu;
constexpr double WBIG = 1e300;

template <int J>
struct Reach {
    int back[J];    // forward scans: steps the element may look back inside its chunk (towards the run head)
    int fwd[J];     // backward scans: steps it may look ahead (towards the run tail)
    bool cin[J];    // run head lies in an earlier chunk  -> takes the forward carry
    bool cout[J];   // run tail lies in a later chunk    -> takes the backward carry
};

// position of the last head at or before t (plain inclusive max-scan of head ? t : -1)
template <int J>
__device__ __forceinline__ void last_flag_pos(const bool (&flag)[J], int (&pos)[J])
{
    const int lane = threadIdx.x & 31;
    int carry = -1;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        int x = flag[j] ? lane + 32 * j : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int xo = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x = max(x, xo);
        }
        x = max(x, carry);
        pos[j] = x;
        carry = __shfl_sync(FULL, x, 31);
    }
}
// position of the first flag at or after t (plain inclusive min-scan from the right)
template <int J>
__device__ __forceinline__ void next_flag_pos(const bool (&flag)[J], int (&pos)[J])
{
    const int lane = threadIdx.x & 31;
    int carry = 0x7fffffff;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        int x = flag[j] ? lane + 32 * j : 0x7fffffff;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int xo = __shfl_down_sync(FULL, x, o);
            if (lane + o < 32) x = min(x, xo);
        }
        x = min(x, carry);
        pos[j] = x;
        carry = __shfl_sync(FULL, x, 0);
    }
}

template <int J>
__device__ __forceinline__ void reach_back_from_heads(const bool (&head)[J], int (&back)[J], bool (&cin)[J])
{
    const int lane = threadIdx.x & 31;
    int hp[J];
    last_flag_pos<J>(head, hp);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane + 32 * j, d = t - hp[j];       // hp >= 0 because t = 0 is always a head
        back[j] = min(lane, d);
        cin[j] = hp[j] < 32 * j;
    }
}
template <int J>
__device__ __forceinline__ void reach_fwd_to_tails(const bool (&tail)[J], int (&fwd)[J], bool (&cout)[J])
{
    const int lane = threadIdx.x & 31;
    int tp[J];
    next_flag_pos<J>(tail, tp);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane + 32 * j;
        const bool none = tp[j] == 0x7fffffff;
        fwd[j] = none ? 0 : min(31 - lane, tp[j] - t);
        cout[j] = !none && tp[j] > 32 * j + 31;
    }
}

struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return a < b ? a : b; } };
struct OpMax { __device__ double operator()(double a, double b) const { return a > b ? a : b; } };

// forward inclusive segmented scan of two value arrays at once
template <int J, class OpA, class OpB>
__device__ __forceinline__ void seg_fwd2(double (&a)[J], double (&b)[J], const int (&back)[J], const bool (&cin)[J],
                                         OpA opa, OpB opb, double ida, double idb)
{
    double ca = ida, cb = idb;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        double x = a[j], y = b[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double xo = __shfl_up_sync(FULL, x, o), yo = __shfl_up_sync(FULL, y, o);
            if (back[j] >= o) { x = opa(xo, x); y = opb(yo, y); }
        }
        if (cin[j]) { x = opa(ca, x); y = opb(cb, y); }
        a[j] = x; b[j] = y;
        ca = __shfl_sync(FULL, x, 31); cb = __shfl_sync(FULL, y, 31);
    }
}
template <int J>
__device__ __forceinline__ void seg_fwd_min_int(int (&a)[J], const int (&back)[J], const bool (&cin)[J])
{
    int ca = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        int x = a[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int xo = __shfl_up_sync(FULL, x, o);
            if (back[j] >= o) x = min(xo, x);
        }
        if (cin[j]) x = min(ca, x);
        a[j] = x;
        ca = __shfl_sync(FULL, x, 31);
    }
}
// every element takes the values held by the tail of its run (a double and an int at once)
template <int J>
__device__ __forceinline__ void seg_take_tail(double (&a)[J], int (&b)[J], const int (&fwd)[J], const bool (&cout)[J])
{
    double ca = 0.0; int cb = 0;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        double x = a[j]; int y = b[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double xo = __shfl_down_sync(FULL, x, o); const int yo = __shfl_down_sync(FULL, y, o);
            if (fwd[j] >= o) { x = xo; y = yo; }
        }
        if (cout[j]) { x = ca; y = cb; }
        a[j] = x; b[j] = y;
        ca = __shfl_sync(FULL, x, 0); cb = __shfl_sync(FULL, y, 0);
    }
}
template <int J>
__device__ __forceinline__ void seg_take_tail_int2(int (&a)[J], int (&b)[J], const int (&fwd)[J], const bool (&cout)[J])
{
    int ca = 0, cb = 0;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        int x = a[j], y = b[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int xo = __shfl_down_sync(FULL, x, o), yo = __shfl_down_sync(FULL, y, o);
            if (fwd[j] >= o) { x = xo; y = yo; }
        }
        if (cout[j]) { x = ca; y = cb; }
        a[j] = x; b[j] = y;
        ca = __shfl_sync(FULL, x, 0); cb = __shfl_sync(FULL, y, 0);
    }
}
// value of the element at t-1 / t+1
template <int J, class T>
__device__ __forceinline__ void shift_from_prev(const T (&v)[J], T (&out)[J], T first)
{
    const int lane = threadIdx.x & 31;
    T carry = first;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const T up = __shfl_up_sync(FULL, v[j], 1);
        out[j] = lane == 0 ? carry : up;
        carry = __shfl_sync(FULL, v[j], 31);
    }
}
template <int J>
__device__ __forceinline__ void shift_from_next(const int (&v)[J], int (&out)[J], int last)
{
    const int lane = threadIdx.x & 31;
    int carry = last;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        const int dn = __shfl_down_sync(FULL, v[j], 1);
        out[j] = lane == 31 ? carry : dn;
        carry = __shfl_sync(FULL, v[j], 0);
    }
}
template <int J>
__device__ __forceinline__ bool any_of(const bool (&p)[J])
{
    bool a = false;
#pragma unroll
    for (int j = 0; j < J; ++j) a |= p[j];
    return __any_sync(FULL, a);
}

// run status bits broadcast from the tail
enum { RS_CONV = 1, RS_BAD = 2, RS_FLAT = 4, RS_EMPTY = 8, RS_FREEBAD = 16 };

// returns true if the storage was solved and written; false => caller queues it for the exact sequential solver
template <int J, bool HINGES>
__device__ bool sto_warp_solve(const View &v, int s, const Hinge *hinges, const int *hcnt)
{
    const int lane = threadIdx.x & 31, T = v.T;
    const int cur = v.ctrl->cur, nxt = 1 - cur, n = v.sto_node[s];
    StoConst k;
    k.mc = v.sto_mc[s]; k.pmax = v.sto_pmax[s]; k.emax = v.sto_emax[s]; k.prox = v.c.prox; k.iprox = 1.0 / v.c.prox;
    const double tolE = 1e-9 * (k.emax > 1.0 ? k.emax : 1.0), tolA = 1e-7 * (k.emax > 1.0 ? k.emax : 1.0);

    StoStep st[J];
    HingeList hl[J];
    bool valid[J];
    int kind[J];          // anchor at the end of t: +1 level = emax, -1 level = 0, 0 none
    double eta[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane + 32 * j;
        valid[j] = t < T;
        const int tt = valid[j] ? t : T - 1;
        const size_t o = (size_t)s * T + tt;
        st[j].Db = sel(v.D, cur)[o]; st[j].Cb = sel(v.C, cur)[o];
        st[j].g0 = v.g0[(size_t)n * v.ldt + tt]; st[j].s1 = v.s1[(size_t)n * v.ldt + tt];
        hl[j].h = HINGES ? hinges + (size_t)tt * v.hcap : nullptr;
        hl[j].n = HINGES ? hcnt[tt] : 0;
        eta[j] = v.eta[o];
        const double Ep = v.E[o];
        kind[j] = !valid[j] ? 0 : (Ep >= k.emax - tolA ? 1 : (Ep <= tolA ? -1 : 0));
    }

    double D[J], C[J], pre[J];
    bool accepted = false;
    for (int as_it = 0; as_it < 24 && !accepted; ++as_it) {
        // ---- run structure from the anchors (timesteps beyond T are isolated one-element runs) ------
        bool head[J], tail[J];
        int prevk[J], endk[J];
        Reach<J> R;
        {
            int kp[J];
            shift_from_prev<J, int>(kind, kp, 1);                        // kind of t-1 (t = 0 starts a run)
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int t = lane + 32 * j;
                head[j] = kp[j] != 0 || !valid[j] || t == T;
                tail[j] = !valid[j] || kind[j] != 0 || t == T - 1;
            }
            reach_back_from_heads<J>(head, R.back, R.cin);
            reach_fwd_to_tails<J>(tail, R.fwd, R.cout);
            // kind of the anchor that closed the previous run = kind at (head position - 1):
            // heads read it from t-1, then it is spread over the run with a forward "take head" scan
            double hk[J], dummy[J];
#pragma unroll
            for (int j = 0; j < J; ++j) { hk[j] = head[j] ? (double)((lane + 32 * j) == 0 ? 0 : kp[j]) : -2.0; dummy[j] = 0.0; }
            seg_fwd2<J>(hk, dummy, R.back, R.cin, OpMax(), OpAdd(), -2.0, 0.0);   // non-heads hold -2 => max = head's value
#pragma unroll
            for (int j = 0; j < J; ++j) { prevk[j] = (int)hk[j]; endk[j] = kind[j]; }
            seg_take_tail<J>(eta, endk, R.fwd, R.cout);                  // start multiplier and end kind of my run
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
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (valid[j]) {
                    const StoEval e = sto_eval(st[j], k, hl[j], eta[j]);
                    D[j] = e.D; C[j] = e.C; pre[j] = e.C - e.D; dy[j] = e.dy;
                } else { D[j] = C[j] = pre[j] = dy[j] = 0.0; }
            }
            seg_fwd2<J>(pre, dy, R.back, R.cin, OpAdd(), OpAdd(), 0.0, 0.0);
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
                        if (hl[j].n != 0) { bu[j] = -WBIG; bd[j] = WBIG; }             // hinge on a flat run: not handled here
                        else { bu[j] = sto_next_break(st[j], k, eta[j], true); bd[j] = sto_next_break(st[j], k, eta[j], false); }
                    }
                }
                seg_fwd2<J>(bu, bd, R.back, R.cin, OpMin(), OpMax(), WBIG, -WBIG);
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
            seg_take_tail<J>(eta, rs, R.fwd, R.cout);
        }
        if (capped) return false;                                             // Newton cap reached
        // final run state for every element: multiplier, status, "flat" (sum of derivatives at the tail)
#pragma unroll
        for (int j = 0; j < J; ++j) if (valid[j] && tail[j] && !(totd[j] < -1e-300)) rs[j] |= RS_FLAT;
        seg_take_tail<J>(eta, rs, R.fwd, R.cout);

        // ---- KKT check ------------------------------------------------------------------------------
        bool vio_up[J], vio_dn[J];
        double Ilo[J], Ihi[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double E = e0[j] + pre[j];
            const bool ok = valid[j] && !(rs[j] & RS_BAD);
            vio_up[j] = ok && E > k.emax + tolE;
            vio_dn[j] = ok && E < -tolE;
            // multiplier interval of my run, contributed per element: a point unless the whole run is flat
            Ilo[j] = -WBIG; Ihi[j] = WBIG;
            if (valid[j]) {
                if (!(rs[j] & RS_FLAT) || hl[j].n != 0) { Ilo[j] = eta[j]; Ihi[j] = eta[j]; }
                else sto_flat_interval(st[j], k, eta[j], D[j], C[j], Ilo[j], Ihi[j]);
            }
        }
        seg_fwd2<J>(Ilo, Ihi, R.back, R.cin, OpMax(), OpMin(), -WBIG, WBIG);      // tails now hold the run interval
        // sign chain over the runs: after an upper anchor eta may not rise, after a lower anchor it may not
        // drop.  Only tails carry run values, the other elements are neutral; the chains restart at the
        // head of a run whose previous anchor has the other kind.
        double Fhi[J], Flo[J];
        {
            bool hup[J], hdn[J];
            int bup[J], bdn[J]; bool cup[J], cdn[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                Fhi[j] = (valid[j] && tail[j]) ? Ihi[j] : WBIG;
                Flo[j] = (valid[j] && tail[j]) ? Ilo[j] : -WBIG;
                hup[j] = head[j] && prevk[j] <= 0;
                hdn[j] = head[j] && prevk[j] >= 0;
            }
            reach_back_from_heads<J>(hup, bup, cup);
            reach_back_from_heads<J>(hdn, bdn, cdn);
            double d1[J], d2[J];
#pragma unroll
            for (int j = 0; j < J; ++j) { d1[j] = 0.0; d2[j] = 0.0; }
            seg_fwd2<J>(Fhi, d1, bup, cup, OpMin(), OpAdd(), WBIG, 0.0);
            seg_fwd2<J>(Flo, d2, bdn, cdn, OpMax(), OpAdd(), -WBIG, 0.0);
        }
        int tv[J], flag[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane + 32 * j;
            tv[j] = (vio_up[j] || vio_dn[j]) ? t : 0x7fffffff;
            flag[j] = 0;
            if (valid[j] && tail[j] && !(rs[j] & RS_BAD)) {
                double a = Flo[j], b = Fhi[j];
                if (t == T - 1 && kind[j] != 0) {                  // end of horizon: eta_{T+1} = 0
                    if (kind[j] > 0) a = a > 0.0 ? a : 0.0; else b = b < 0.0 ? b : 0.0;
                }
                if (a > b) flag[j] |= RS_EMPTY;
                if (freeend[j] && (Flo[j] > 0.0 || Fhi[j] < 0.0)) flag[j] |= RS_FREEBAD;
            }
            if (valid[j] && tail[j] && (rs[j] & RS_BAD)) flag[j] |= RS_BAD;
        }
        seg_fwd_min_int<J>(tv, R.back, R.cin);                      // tails hold the first violated timestep
        seg_take_tail_int2<J>(tv, flag, R.fwd, R.cout);             // ... and tell their run

        // ---- repair the active set -----------------------------------------------------------------
        bool change = false, anybad = false;
        int wantdrop[J];                                            // head of a run that wants the previous anchor gone
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane + 32 * j;
            anybad |= valid[j] && (flag[j] != 0 || vio_up[j] || vio_dn[j]);
            wantdrop[j] = (valid[j] && head[j] && ((flag[j] & (RS_EMPTY | RS_FREEBAD)) || ((flag[j] & RS_BAD) && freeend[j]))) ? 1 : 0;
            if (valid[j] && tv[j] == t) { kind[j] = vio_up[j] ? 1 : -1; change = true; }       // (1) new anchor
        }
        int nextwant[J];
        shift_from_next<J>(wantdrop, nextwant, 0);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane + 32 * j;
            if (!valid[j] || kind[j] == 0 || tv[j] == t) continue;
            bool drop = nextwant[j] != 0;                                                       // (2) wrong sign at my anchor
            if (tail[j] && (flag[j] & RS_BAD)) drop = true;                                     //     my run cannot meet its target
            if (t == T - 1 && (flag[j] & RS_EMPTY)) drop = true;                                //     wrong sign at the horizon end
            if (drop) { kind[j] = 0; change = true; }
        }
        anybad = __any_sync(FULL, anybad);
        change = __any_sync(FULL, change);
        if (!anybad) accepted = true;
        else if (!change) return false;
        if (accepted) {
#pragma unroll
            for (int j = 0; j < J; ++j) pre[j] += e0[j];            // levels
        }
    }
    if (!accepted) return false;

    // ---- emit ------------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane + 32 * j;
        if (!valid[j]) continue;
        const size_t o = (size_t)s * T + t;
        sel(v.D, nxt)[o] = D[j]; sel(v.C, nxt)[o] = C[j]; v.E[o] = pre[j]; v.eta[o] = eta[j];
        note_move(v, n, t, (D[j] - st[j].Db) - (C[j] - st[j].Cb));
    }
    return true;
}

}  // namespace dopf
#endif
