// Warp-parallel storage solve (one warp per storage, lane l owns timesteps l, l+32, ...).
//
// Active-set method on the level bounds 0 <= E_t <= emax (reference problem:
// /root/reference/src/optimization/subproblems.jl:107-207 after slack elimination):
//   * anchors  = timesteps whose level is fixed at a bound (kind +1: emax, -1: 0); between two anchors
//     ("run") the level multiplier eta is constant;
//   * all runs are solved at once: every lane evaluates y_t(eta) for its timesteps (sto_eval), run sums
//     come from segmented warp scans, one safeguarded Newton update per run and pass;
//   * then the KKT conditions are checked in parallel: levels inside the bounds, and a multiplier
//     path with the right sign at every anchor (intervals, because saturated runs have a non-unique
//     multiplier).  Violated levels add an anchor (first violated timestep of the run), anchors with a
//     wrong sign are dropped, and the solve repeats;
//   * the result is accepted only when the KKT conditions hold (=> exact optimum of the convex
//     problem).  Storages that do not verify within the caps go to the sequential exact solver.
// The initial active set comes from the previous levels (warm start).
#ifndef DOPF_STO_WARP_CUH
#define DOPF_STO_WARP_CUH

#include "dopf_bodies.h"

namespace dopf {

constexpr unsigned FULL = 0xffffffffu;
constexpr double WBIG = 1e300;

// ---- chunked segmented scans over the strided ownership t = lane + 32*j ------------------------------
// forward inclusive scan, restarting at elements with head[j] == true
template <int J, class T, class Op>
__device__ __forceinline__ void seg_scan_fwd(T (&v)[J], const bool (&head)[J], Op op, T identity)
{
    const int lane = threadIdx.x & 31;
    T carry = identity;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        T x = v[j];
        bool f = head[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T xo = __shfl_up_sync(FULL, x, o);
            const bool fo = __shfl_up_sync(FULL, (int)f, o) != 0;
            if (lane >= o && !f) { x = op(xo, x); f = fo; }
        }
        if (!f) x = op(carry, x);          // no head between the chunk start and this element
        v[j] = x;
        carry = __shfl_sync(FULL, x, 31);
    }
}

// backward inclusive scan (from larger t to smaller), restarting at elements with tail[j] == true
template <int J, class T, class Op>
__device__ __forceinline__ void seg_scan_bwd(T (&v)[J], const bool (&tail)[J], Op op, T identity)
{
    const int lane = threadIdx.x & 31;
    T carry = identity;
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
        T x = v[j];
        bool f = tail[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T xo = __shfl_down_sync(FULL, x, o);
            const bool fo = __shfl_down_sync(FULL, (int)f, o) != 0;
            if (lane + o < 32 && !f) { x = op(xo, x); f = fo; }
        }
        if (!f) x = op(carry, x);
        v[j] = x;
        carry = __shfl_sync(FULL, x, 0);
    }
}

// value of the element at t-1 (identity for t = 0)
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
__device__ __forceinline__ bool any_of(const bool (&p)[J])
{
    bool a = false;
#pragma unroll
    for (int j = 0; j < J; ++j) a |= p[j];
    return __any_sync(FULL, a);
}

struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return a < b ? a : b; } };
struct OpMax { __device__ double operator()(double a, double b) const { return a > b ? a : b; } };
struct OpTakeFirst { template <class T> __device__ T operator()(T a, T) const { return a; } };   // op(incoming, self) = incoming
struct OpLastNonzero { __device__ int operator()(int a, int b) const { return b != 0 ? b : a; } };
struct OpMinInt { __device__ int operator()(int a, int b) const { return a < b ? a : b; } };

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

    double D[J], C[J], E[J];
    bool accepted = false;
    for (int as_it = 0; as_it < 24 && !accepted; ++as_it) {
        // ---- run structure from the anchors -----------------------------------------------------
        bool head[J], tail[J];
        int prevk[J], endk[J];
        {
            int kp[J];
            shift_from_prev<J, int>(kind, kp, 1);                     // kind of t-1 (t = 0 starts a run)
            int lastnz[J];
#pragma unroll
            for (int j = 0; j < J; ++j) { head[j] = kp[j] != 0; lastnz[j] = kind[j]; }
            bool nohead[J];
#pragma unroll
            for (int j = 0; j < J; ++j) nohead[j] = false;
            seg_scan_fwd<J, int>(lastnz, nohead, OpLastNonzero(), 0);  // last anchor kind at or before t
            shift_from_prev<J, int>(lastnz, prevk, 0);                // ... strictly before t
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int t = lane + 32 * j;
                tail[j] = valid[j] && (kind[j] != 0 || t == T - 1);
                endk[j] = kind[j];
            }
            seg_scan_bwd<J, int>(endk, tail, OpTakeFirst(), 0);       // kind at the end of my run
        }
        double e0[J], target[J];
        bool freeend[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            e0[j] = prevk[j] > 0 ? k.emax : 0.0;
            target[j] = (endk[j] > 0 ? k.emax : 0.0) - e0[j];
            freeend[j] = endk[j] == 0;
        }
        // one multiplier per run: start from the previous multiplier at the run end (0 for the free end)
        seg_scan_bwd<J, double>(eta, tail, OpTakeFirst(), 0.0);
#pragma unroll
        for (int j = 0; j < J; ++j) if (freeend[j]) eta[j] = 0.0;

        // ---- simultaneous safeguarded Newton on all runs ------------------------------------------
        double lo[J], hi[J], rlo[J], rhi[J], toty[J], totd[J], pre[J];
        bool conv[J], bad[J];          // bad: the run cannot meet its target with this active set
#pragma unroll
        for (int j = 0; j < J; ++j) { lo[j] = -WBIG; hi[j] = WBIG; rlo[j] = rhi[j] = 0.0; conv[j] = !valid[j]; bad[j] = false; }
        for (int it = 0; it < 24; ++it) {
            double dy[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (valid[j]) {
                    const StoEval e = sto_eval(st[j], k, hl[j], eta[j]);
                    D[j] = e.D; C[j] = e.C; pre[j] = e.C - e.D; dy[j] = e.dy;
                } else { D[j] = C[j] = pre[j] = dy[j] = 0.0; }
            }
            seg_scan_fwd<J, double>(pre, head, OpAdd(), 0.0);
            seg_scan_fwd<J, double>(dy, head, OpAdd(), 0.0);
#pragma unroll
            for (int j = 0; j < J; ++j) { toty[j] = pre[j]; totd[j] = dy[j]; }
            seg_scan_bwd<J, double>(toty, tail, OpTakeFirst(), 0.0);
            seg_scan_bwd<J, double>(totd, tail, OpTakeFirst(), 0.0);
            bool needflat[J], pending[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const double r = toty[j] - target[j];
                const double tolS = 1e-13 * (1.0 + fabs(target[j]) + k.pmax);
                if (valid[j] && !conv[j] && !bad[j]) conv[j] = freeend[j] || fabs(r) <= tolS;
                pending[j] = valid[j] && !conv[j] && !bad[j];
                needflat[j] = pending[j] && !(totd[j] < -1e-300);
            }
            if (!any_of<J>(pending)) break;
            double best[J];
#pragma unroll
            for (int j = 0; j < J; ++j) best[j] = 0.0;
            if (any_of<J>(needflat)) {
                // saturated runs: nearest clip breakpoint of any step of the run in the needed direction
                double bu[J], bd[J];
                bool hard[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    bu[j] = WBIG; bd[j] = -WBIG; hard[j] = false;
                    if (needflat[j]) {
                        if (hl[j].n != 0) hard[j] = true;
                        else { bu[j] = sto_next_break(st[j], k, eta[j], true); bd[j] = sto_next_break(st[j], k, eta[j], false); }
                    }
                }
                seg_scan_fwd<J, double>(bu, head, OpMin(), WBIG); seg_scan_bwd<J, double>(bu, tail, OpTakeFirst(), WBIG);
                seg_scan_fwd<J, double>(bd, head, OpMax(), -WBIG); seg_scan_bwd<J, double>(bd, tail, OpTakeFirst(), -WBIG);
                double hd[J];
#pragma unroll
                for (int j = 0; j < J; ++j) hd[j] = hard[j] ? 1.0 : 0.0;
                seg_scan_fwd<J, double>(hd, head, OpMax(), 0.0); seg_scan_bwd<J, double>(hd, tail, OpTakeFirst(), 0.0);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (!needflat[j]) continue;
                    const bool up = toty[j] - target[j] > 0.0;
                    const double b = up ? bu[j] : bd[j];
                    if (hd[j] != 0.0 || !(fabs(b) < WBIG)) bad[j] = true;          // unreachable / hinge on a flat run
                    else best[j] = b + (up ? 1.0 : -1.0) * 1e-11 * (1.0 + fabs(b));
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (!pending[j] || bad[j]) continue;
                const double r = toty[j] - target[j];
                if (r > 0.0) { lo[j] = eta[j]; rlo[j] = r; } else { hi[j] = eta[j]; rhi[j] = r; }
                double en = needflat[j] ? best[j] : eta[j] - r / totd[j];
                if (!(en > lo[j] && en < hi[j])) {
                    if (lo[j] > -WBIG && hi[j] < WBIG) {
                        en = lo[j] - rlo[j] * (hi[j] - lo[j]) / (rhi[j] - rlo[j]);
                        if (!(en > lo[j] && en < hi[j])) en = 0.5 * (lo[j] + hi[j]);
                        if ((hi[j] - lo[j]) <= 1e-15 * (1.0 + fabs(lo[j]))) conv[j] = true;
                    } else bad[j] = true;
                }
                eta[j] = en;
            }
        }
        {
            bool pend[J];
#pragma unroll
            for (int j = 0; j < J; ++j) pend[j] = valid[j] && !conv[j] && !bad[j];
            if (any_of<J>(pend)) return false;                         // Newton cap reached
        }

        // ---- KKT check ------------------------------------------------------------------------------
        bool vio_up[J], vio_dn[J], change = false;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            E[j] = e0[j] + pre[j];
            vio_up[j] = valid[j] && !bad[j] && E[j] > k.emax + tolE;
            vio_dn[j] = valid[j] && !bad[j] && E[j] < -tolE;
        }
        // multiplier interval of every run (flat runs: stretch on which every step keeps its value)
        double Ilo[J], Ihi[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            Ilo[j] = -WBIG; Ihi[j] = WBIG;                              // neutral for min/max over the run
            if (valid[j]) {
                if (totd[j] < -1e-300 || hl[j].n != 0) { Ilo[j] = eta[j]; Ihi[j] = eta[j]; }
                else sto_flat_interval(st[j], k, eta[j], D[j], C[j], Ilo[j], Ihi[j]);
            }
        }
        seg_scan_fwd<J, double>(Ilo, head, OpMax(), -WBIG); seg_scan_bwd<J, double>(Ilo, tail, OpTakeFirst(), -WBIG);
        seg_scan_fwd<J, double>(Ihi, head, OpMin(), WBIG); seg_scan_bwd<J, double>(Ihi, tail, OpTakeFirst(), WBIG);
        // sign chain: after an upper anchor eta may not rise, after a lower anchor it may not drop.
        // Fhi = running min of Ihi over consecutive runs linked by upper anchors (restart after a lower
        // anchor), Flo = running max of Ilo over runs linked by lower anchors.
        double Fhi[J], Flo[J];
        bool hup[J], hdn[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            Fhi[j] = Ihi[j]; Flo[j] = Ilo[j];
            hup[j] = head[j] && prevk[j] <= 0;      // chain of "may not rise" restarts unless the previous anchor was upper
            hdn[j] = head[j] && prevk[j] >= 0;
        }
        seg_scan_fwd<J, double>(Fhi, hup, OpMin(), WBIG);
        seg_scan_fwd<J, double>(Flo, hdn, OpMax(), -WBIG);
        bool drop_prev[J];                           // the anchor just before my run has the wrong sign
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = lane + 32 * j;
            double a = Flo[j], b = Fhi[j];
            if (valid[j] && t == T - 1 && kind[j] != 0) {          // end of horizon: eta_{T+1} = 0
                if (kind[j] > 0) a = a > 0.0 ? a : 0.0; else b = b < 0.0 ? b : 0.0;
            }
            const bool empty = valid[j] && !bad[j] && a > b;
            // report at the head of the run (the anchor to drop is the end of the previous run); for the
            // horizon end the terminal anchor itself is dropped below
            drop_prev[j] = empty;
        }
        // a run is "empty" at all of its elements or none (interval values are run-uniform) except for the
        // terminal adjustment, which only touches t = T-1; propagate that to the run head
        {
            double em[J];
#pragma unroll
            for (int j = 0; j < J; ++j) em[j] = drop_prev[j] ? 1.0 : 0.0;
            seg_scan_bwd<J, double>(em, tail, OpMax(), 0.0);
#pragma unroll
            for (int j = 0; j < J; ++j) drop_prev[j] = em[j] != 0.0;
        }
        // free end must admit eta = 0
        bool freebad[J];
#pragma unroll
        for (int j = 0; j < J; ++j) freebad[j] = valid[j] && freeend[j] && (Flo[j] > 0.0 || Fhi[j] < 0.0);

        // ---- repair the active set -----------------------------------------------------------------
        // (1) first level violation of every run becomes an anchor
        {
            int tv[J];
#pragma unroll
            for (int j = 0; j < J; ++j) tv[j] = (vio_up[j] || vio_dn[j]) ? lane + 32 * j : 0x7fffffff;
            seg_scan_fwd<J, int>(tv, head, OpMinInt(), 0x7fffffff);
            seg_scan_bwd<J, int>(tv, tail, OpTakeFirst(), 0x7fffffff);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (valid[j] && tv[j] == lane + 32 * j) { kind[j] = vio_up[j] ? 1 : -1; change = true; }
            }
        }
        // (2) anchors with a wrong multiplier sign (or closing an unsatisfiable run) are dropped:
        //     drop the anchor that ends the previous run  <=> element t-1 of a head element
        {
            bool dropme[J], nexthead_drop[J];
            // does the run starting at t+1 ask to drop me?  (shift information from t+1 to t)
            const int lane_ = lane;
            bool carry = false;
#pragma unroll
            for (int j = J - 1; j >= 0; --j) {
                const bool mine = head[j] && (drop_prev[j] || freebad[j] || (bad[j] && freeend[j])) && valid[j];
                const bool dn = __shfl_down_sync(FULL, (int)mine, 1) != 0;
                nexthead_drop[j] = lane_ == 31 ? carry : dn;
                carry = __shfl_sync(FULL, (int)mine, 0) != 0;
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int t = lane + 32 * j;
                dropme[j] = valid[j] && kind[j] != 0 && nexthead_drop[j];
                // unsatisfiable run: drop its own end anchor; terminal anchor with a wrong sign likewise
                if (valid[j] && kind[j] != 0 && tail[j] && bad[j]) dropme[j] = true;
                if (valid[j] && kind[j] != 0 && t == T - 1 && drop_prev[j]) dropme[j] = true;
                if (dropme[j]) { kind[j] = 0; change = true; }
            }
        }
        // a run that starts at t = 0 cannot drop a previous anchor: if it is empty/bad without any other
        // change we cannot repair it here
        change = __any_sync(FULL, change);
        bool anybad = false;
#pragma unroll
        for (int j = 0; j < J; ++j) anybad |= bad[j] || drop_prev[j] || freebad[j] || vio_up[j] || vio_dn[j];
        anybad = __any_sync(FULL, anybad);
        if (!anybad) accepted = true;
        else if (!change) return false;
    }
    if (!accepted) return false;

    // ---- emit ------------------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int t = lane + 32 * j;
        if (!valid[j]) continue;
        const size_t o = (size_t)s * T + t;
        sel(v.D, nxt)[o] = D[j]; sel(v.C, nxt)[o] = C[j]; v.E[o] = E[j]; v.eta[o] = eta[j];
        note_move(v, n, t, (D[j] - st[j].Db) - (C[j] - st[j].Cb));
    }
    return true;
}

}  // namespace dopf
#endif
