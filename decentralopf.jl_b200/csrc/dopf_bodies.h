// Data layout ("View") of one ADMM instance in device memory and the per-thread bodies of the
// streaming / bookkeeping kernels.  Bodies are host+device inline so that tests/host_emul can
// run the identical arithmetic sequentially on the CPU; the __global__ wrappers live in
// dopf_kernels.cu.  Reference semantics of each step are cited at the body.
#ifndef DOPF_BODIES_H
#define DOPF_BODIES_H

#include <cstddef>
#include "dopf_math.h"

namespace dopf {

#if defined(__CUDA_ARCH__)
#define DOPF_ATOMIC_MAX_U64(addr, v) atomicMax((unsigned long long *)(addr), (unsigned long long)(v))
#define DOPF_ATOMIC_MIN_U64(addr, v) atomicMin((unsigned long long *)(addr), (unsigned long long)(v))
#define DOPF_ATOMIC_ADD_I32(addr, v) atomicAdd((int *)(addr), (int)(v))
#define DOPF_ATOMIC_EXCH_I32(addr, v) atomicExch((int *)(addr), (int)(v))
#else
static inline unsigned long long dopf_host_max_u64(unsigned long long *a, unsigned long long v)
{ unsigned long long o = *a; if (v > o) *a = v; return o; }
static inline unsigned long long dopf_host_min_u64(unsigned long long *a, unsigned long long v)
{ unsigned long long o = *a; if (v < o) *a = v; return o; }
static inline int dopf_host_add_i32(int *a, int v) { int o = *a; *a = o + v; return o; }
static inline int dopf_host_exch_i32(int *a, int v) { int o = *a; *a = v; return o; }
#define DOPF_ATOMIC_MAX_U64(addr, v) dopf_host_max_u64((unsigned long long *)(addr), (unsigned long long)(v))
#define DOPF_ATOMIC_MIN_U64(addr, v) dopf_host_min_u64((unsigned long long *)(addr), (unsigned long long)(v))
#define DOPF_ATOMIC_ADD_I32(addr, v) dopf_host_add_i32((int *)(addr), (int)(v))
#define DOPF_ATOMIC_EXCH_I32(addr, v) dopf_host_exch_i32((int *)(addr), (int)(v))
#endif

// select one of a double-buffered pointer pair without dynamically indexing kernel parameters
template <class T> DOPF_HD T *sel(T *const (&a)[2], int i) { return i ? a[1] : a[0]; }

enum { DOPF_ERR_NONE = 0, DOPF_ERR_HINGE_CAP = 1, DOPF_ERR_WORK_CAP = 2 };

// device-resident control block (one per instance)
struct Ctrl {
    int cur;            // index of the buffers holding the previous iterate
    int iteration;      // admm.iteration (1-based)            structures/admm.jl:29
    int converged;      // admm.convergence.all                 structures/convergence.jl
    int conv_lambda, conv_mue, conv_rho;
    int error;          // DOPF_ERR_*; set => every later kernel is a no-op, state stays valid
    int iters_done;     // iterations executed since create
    int finish_cnt;     // blocks of k_lambda_finish that are done (scenario batches: the last one flips the buffers)
    int sto_next;       // work counter of the storage predict kernel (warps draw storages from it)
    // generator correction queue: entries (low word) and groups (high word) are advanced by ONE 64-bit atomic in k_verify
    alignas(8) int gen_work_cnt; int gen_grp_cnt;
    int sto_work_cnt, cold_work_cnt, pair_cnt, fix_node_cnt;
    int stat_sto_cold;                   // storages solved by the cold funnel in the last iteration
    int stat_fix_seq;                    // cumulated correction-pass storages that needed the sequential solver
    int stat_gen_fix, stat_sto_fix;      // cumulated corrected agents (statistics)
    int stat_tight_rows, stat_wide_rows; // of the last iteration
    int stat_tight_acc, stat_wide_acc;   // accumulators of the batched finish kernel
    unsigned long long res_bits[3];      // max |dual_{k+1}-dual_k| for lambda, mue, rho (bits)
    double res[3];
    double total_costs;
};

static_assert(offsetof(Ctrl, gen_work_cnt) % 8 == 0 && offsetof(Ctrl, gen_grp_cnt) == offsetof(Ctrl, gen_work_cnt) + 4, "packed queue counters");

struct View {
    // sizes.  A batch of C independent scenarios on one grid (BASELINE configs[3]) is laid out as extra COLUMNS of the
    // network matrices: scenario c owns the columns [c*T, (c+1)*T) (TC = C*T columns in total, one PTDF shared), and as
    // C*G generators / C*S storages living on "virtual nodes" vn = c*N + n.  C = 1 is the plain single problem.
    int N, L, T, G, S, A;   // G, S: ALL agents of the batch; A: agents of ONE scenario (the divisor of the average slacks)
    int NS, TC;             // scenarios C, total columns C*T
    int ldt;          // leading dimension of every [node|line][column] matrix (TC rounded up to 32)
    int Np, Lp;       // padded row counts (multiples of 64) of the node / line matrices
    int hcap;         // hinge capacity per (agent, t) in the correction pass
    int gen_work_cap;
    int debug;        // dopf_config.debug_flags: bit0 correction pass of the storages uses the sequential solver, bit1 predict pass too
    Coef c;
    // static problem data
    const double *ptdf;     // [Lp][Np] zero padded                               admm.ptdf
    const double *fmax;     // [Lp]
    const double *demand;   // [Np][ldt]
    const double *q;        // [Np]  sum_l ptdf[l,n]^2
    const double *prow;     // [Lp]  max_n |ptdf[l,n]|
    double *mwide;          // [Lp]  max_n |ptdf[l,n]| * box range of node n
    double *rbox;           // [Np]  largest possible |delta| of an agent at node n (max-reduced over ranks)
    const double *nagents;  // [C][Np]  number of agents at node n (of scenario c)
    const double *gen_mc, *gen_pmax; const int *gen_node; const int *gen_ptr;   // sorted by virtual node (gen_node = c*N + n); ptr [C*N+1]
    const double *sto_mc, *sto_pmax, *sto_emax; const int *sto_node; const int *sto_ptr;
    // iterate (double buffered: [cur] = previous, [1-cur] = being written)
    double *P[2];           // [G][T]
    double *D[2], *C[2];    // [S][T]
    double *E;              // [S][T] level of the newest iterate
    double *eta;            // [S][T] level-multiplier path of the last storage solve (warm start)
    double *inj[2];         // [Np][ldt]  nodal injection (all ranks' agents)
    double *injloc[2];      // [Np][ldt]  contribution of this rank's agents (== inj on one GPU)
    int demand_on;          // 1: the local injection carries -demand (single-GPU handles); 0 in the partitioned mode (subtracted after the exchange)
    double *ssum[2];        // [ldt]   sum_n inj
    double *ssum_part;      // [COLSUM_R][ldt] partial column sums (fixed row groups, folded in order by the last block)
    int *colsum_cnt;        // [ldt/32] arrival counters of k_colsum (self-resetting)
    double *flow[2];        // [Lp][ldt]
    double *avgU, *avgK;    // [Lp][ldt]
    double *lam[2];         // [ldt]
    double *mu[2], *rho[2]; // [Lp][ldt]
    // per-iteration scratch
    double *bplus, *bminus, *M, *Wt;   // [Lp][ldt]
    double *g0, *s1;                   // [Np][ldt]
    double *rg;                        // [Np][ldt] 1/(prox + s1): generator step size
    double *rg2;                       // [Np][ldt] 1/(prox + 2 s1) (storage steps with both variables free)
    unsigned long long *dn;            // [Np][ldt] max |delta| of the agents at (n,t) (bits)
    // node statistics of the moves delta_i of this rank's agents at (n,t), written by body_inject:
    //   nst[0] = min delta (<= 0), nst[1] = max delta (>= 0), nst[2] = largest negative delta (-inf if none),
    //   nst[3] = smallest positive delta (+inf if none), nst[4] = sum of negative deltas, nst[5] = sum of
    //   positive deltas, nst[6] = number of negative movers, nst[7] = number of positive movers
    double *nst[8];                    // each [ldt][Np] (timestep-major: the slack rows scan over the nodes of one t)
    unsigned long long *dmax;          // [ldt]
    unsigned char *flags;              // [ldt][Lp] bit0: U side wide candidate, bit1: K side
    int *wide, *wcnt;                  // [T][2L] entries l*2+side ; [T]
    double *wide_b;                    // [T][2L] b of the wide entries (coalesced companion)
    const double *ptdfT;               // [Np][Lp] transposed PTDF (node-major) for per-agent hinge collection
    int *cold_work;                    // [S] storages whose warm start did not verify
    int *tight, *tcnt;                 // [T][2L] ; [T]
    double *tight_b;                   // [T][2L] b of the tight entries (device path: k_verify reads it beside the list)
    int *gen_work;                     // [gen_work_cap] g*T+t
    int *gen_grp;                      // [gen_work_cap][2] groups of consecutive work entries of one (n,t): first entry, count
    int *sto_work, *sto_flag;          // [S], [S]
    int *fix_node_flag, *fix_node_list, *fix_node_slot;   // [Np] nodes with a storage on the work list (device path only)
    double *rowsumU, *rowsumK;         // [Lp][ldt] tight rows: sum_i (b -+ p delta_i)_+ minus the closed form of an uncrossed row (rowsum_is_corr) or the whole sum
    int rowsum_is_corr;
    // agent-partitioned mode: every rank multiplies PTDF with the injection of ITS agents over ITS node range only; the
    // partial flows travel in the third slab of the row-sum exchange buffer and the demand part is a constant
    double *xflow;                     // [Lp][ldt] partial flow PTDF[:, own nodes] * (injection of the own agents), summed over the ranks by the exchange
    const double *flowD;               // [Lp][ldt] PTDF * demand (null on single-GPU handles: the demand is part of the injection)
    int *pair_row, *pair_node, *pair_col; double *pair_val; int pair_cap;   // queue of (tight row, node, column) whose agents are summed one by one
    int *pbase, *pcnt;                 // [TC][2L] queue range of every tight-list entry
    unsigned long long *counters;   // [32] diagnostics (filled only by builds with -DDOPF_STATS)
    // per-scenario convergence state (convergence.jl:1-31 for every scenario of the batch separately); a converged
    // scenario is frozen: its agents, duals and average slacks are carried through unchanged while the others go on
    int *sc_iteration, *sc_converged, *sc_conv;       // [C], [C], [C][3]
    unsigned long long *sc_res_bits;                   // [C][3]
    double *sc_res;                                    // [C][3]
    Ctrl *ctrl;

    DOPF_HD int scen_of_vn(int vn) const { return NS == 1 ? 0 : vn / N; }
    DOPF_HD int scen_of_col(int col) const { return NS == 1 ? 0 : col / T; }
    // index of (agent-local timestep t) of an agent on virtual node vn in a [Np][ldt] node matrix
    DOPF_HD size_t nt_of(int vn, int t) const { const int c = scen_of_vn(vn); return (size_t)(vn - c * N) * ldt + (size_t)c * T + t; }
    DOPF_HD int col_of(int vn, int t) const { return scen_of_vn(vn) * T + t; }
};

// ---- row preparation: everything that depends on (line, t) only -------------------------------
// consumes the previous iterate's flows, average slacks and the current duals
// (penalty_terms.jl:10-52 after slack elimination)
DOPF_HD void body_row_prep(const View &v, int l, int t)
{
    const int cur = v.ctrl->cur;
    const size_t i = (size_t)l * v.ldt + t;
    RowPrep r = row_prep(v.c, v.fmax[l], sel(v.flow, cur)[i], v.avgU[i], v.avgK[i], sel(v.mu, cur)[i], sel(v.rho, cur)[i]);
    const bool real = (l < v.L) && (t < v.TC);
    v.bplus[i] = r.bplus; v.bminus[i] = r.bminus;
    v.M[i] = real ? r.M : 0.0; v.Wt[i] = real ? r.Wt : 0.0;
    unsigned char f = 0;
    if (real) {
        if (fabs(r.bplus) <= v.mwide[l]) f |= 1;
        if (fabs(r.bminus) <= v.mwide[l]) f |= 2;
    }
    v.flags[(size_t)t * v.Lp + l] = f;
}

// ---- generator predict: anchor-linearised closed form + box projection ------------------------
// (subproblems.jl:63-83 reduced; exact whenever no slack hinge lies in (0, delta])
DOPF_HD double body_gen_predict(const View &v, int g, int t, double Pprev, int n, double mc, double pmax)
{
    const size_t nt = v.nt_of(n, t);
    double Pn = Pprev - (mc + v.g0[nt]) * v.rg[nt];
    Pn = Pn < 0.0 ? 0.0 : (Pn > pmax ? pmax : Pn);
    return Pn;
}

DOPF_HD void note_move(const View &v, int vn, int t, double delta)      // agent on virtual node vn, agent-local timestep t
{
    const double ad = fabs(delta);
    if (ad == 0.0) return;
    const unsigned long long b = nonneg_bits(ad);
    unsigned long long *pn = v.dn + v.nt_of(vn, t);
    if (b > *pn) DOPF_ATOMIC_MAX_U64(pn, b);
    // dmax[t] = max_n dn[n][t] is computed by a separate column reduction (k_dmax): updating it here
    // would serialise every moving agent of a timestep on one address
}

// ---- storages: accessors over the device layout ------------------------------------------------
struct GlobalSteps {     // previous iterate and anchor linearisation of one storage, read in place
    const double *Db, *Cb, *g0, *s1;
    const Hinge *hinges; const int *hcnt; int hcap; bool sorted;
    DOPF_HD StoStep step(int t) const { StoStep st; st.Db = Db[t]; st.Cb = Cb[t]; st.g0 = g0[t]; st.s1 = s1[t]; return st; }
    DOPF_HD HingeList list(int t) const
    {
        HingeList l;
        l.h = hinges ? hinges + (size_t)t * hcap : nullptr;
        l.n = hinges ? hcnt[t] : 0;
        l.sorted = sorted;
        return l;
    }
};

struct StoEmit {         // writes D, C, E, eta of the new iterate and records the move
    const View &v; const GlobalSteps &sp; StoConst k; int s, n; double E;
    DOPF_HD StoEmit(const View &vv, const GlobalSteps &ss, const StoConst &kk, int s_, int n_) : v(vv), sp(ss), k(kk), s(s_), n(n_), E(0.0) {}
    DOPF_HD void operator()(int t, double eta)
    {
        if (t == 0) E = 0.0;
        const int nxt = 1 - v.ctrl->cur;
        const StoStep st = sp.step(t);
        const StoEval e = sto_eval(st, k, sp.list(t), eta);
        E += e.C - e.D;
        const size_t o = (size_t)s * v.T + t;
        sel(v.D, nxt)[o] = e.D; sel(v.C, nxt)[o] = e.C; v.E[o] = E; v.eta[o] = eta;
    }
};

DOPF_HD void sto_setup(const View &v, int s, StoConst &k, GlobalSteps &sp)
{
    const int cur = v.ctrl->cur, n = v.sto_node[s];
    k.mc = v.sto_mc[s]; k.pmax = v.sto_pmax[s]; k.emax = v.sto_emax[s]; k.prox = v.c.prox; k.iprox = 1.0 / v.c.prox;
    sp.Db = sel(v.D, cur) + (size_t)s * v.T; sp.Cb = sel(v.C, cur) + (size_t)s * v.T;
    sp.g0 = v.g0 + v.nt_of(n, 0); sp.s1 = v.s1 + v.nt_of(n, 0);
    sp.hinges = nullptr; sp.hcnt = nullptr; sp.hcap = 0; sp.sorted = false;
}

// record the moves of a finished storage (after its final emit pass)
DOPF_HD void sto_note_moves(const View &v, int s)
{
    const int cur = v.ctrl->cur, nxt = 1 - cur, n = v.sto_node[s];
    for (int t = 0; t < v.T; ++t) {
        const size_t o = (size_t)s * v.T + t;
        note_move(v, n, t, (sel(v.D, nxt)[o] - sel(v.D, cur)[o]) - (sel(v.C, nxt)[o] - sel(v.C, cur)[o]));
    }
}

// warm attempt for storage s; on failure the storage is queued for the cold funnel
DOPF_HD void body_sto_warm(const View &v, int s)
{
    StoConst k; GlobalSteps sp;
    sto_setup(v, s, k, sp);
    StoEmit emit(v, sp, k, s, v.sto_node[s]);
    const double *ep = v.eta + (size_t)s * v.T, *Ep = v.E + (size_t)s * v.T;
    auto prev = [ep](int t) { return ep[t]; };
    auto prevE = [Ep](int t) { return Ep[t]; };
    if (sto_warm_try(sp, k, v.T, prev, prevE, emit)) sto_note_moves(v, s);
    else v.cold_work[DOPF_ATOMIC_ADD_I32(&v.ctrl->cold_work_cnt, 1)] = s;
}

DOPF_HD void body_sto_cold(const View &v, int s, const Hinge *hinges, const int *hcnt, bool sorted = false)
{
    StoConst k; GlobalSteps sp;
    sto_setup(v, s, k, sp);
    sp.hinges = hinges; sp.hcnt = hcnt; sp.hcap = v.hcap; sp.sorted = sorted;
    StoEmit emit(v, sp, k, s, v.sto_node[s]);
    bool done = false;
    if (hinges) {
        // correction pass: the multiplier path of the predict pass (just written) is the warm start
        const double *ep = v.eta + (size_t)s * v.T, *Ep = v.E + (size_t)s * v.T;
        auto prev = [ep](int t) { return ep[t]; };
        auto prevE = [Ep](int t) { return Ep[t]; };
        done = sto_warm_try(sp, k, v.T, prev, prevE, emit);
    }
    if (!done) sto_funnel_seq(sp, k, v.T, emit);
    sto_note_moves(v, s);
}

// ---- verify: which agents at (n,t) moved across a slack hinge? --------------------------------
// nearest hinge breakpoints (lo < 0 < hi) of the tight rows around delta = 0 that an agent of (n,t) can
// have reached; false if there is none within the largest move of the node
DOPF_HD bool verify_bounds(const View &v, int n, int t, double &lo, double &hi)
{
    const size_t nt = (size_t)n * v.ldt + t;
    const double dnt = bits_nonneg(v.dn[nt]);
    if (dnt == 0.0) return false;
    lo = -INFINITY; hi = INFINITY;
    const int cnt = v.tcnt[t];
    const int *lst = v.tight + (size_t)t * 2 * v.L;
    for (int j = 0; j < cnt; ++j) {
        const int l = lst[j] >> 1, side = lst[j] & 1;
        const double p = v.ptdf[(size_t)l * v.Np + n];
        const double b = side ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
        if (fabs(b) > fabs(p) * dnt * (1.0 + 1e-12)) continue;      // |b/p| > dnt (without the division)
        Hinge h;
        if (!make_hinge(v.c, p, b, side, h)) continue;
        if (fabs(h.bp) > dnt) continue;
        if (h.bp > 0.0) { if (h.bp < hi) hi = h.bp; }
        else if (h.bp < 0.0) { if (h.bp > lo) lo = h.bp; }
        else { if (h.sg > 0.0) hi = 0.0; else lo = 0.0; }
    }
    return !(lo == -INFINITY && hi == INFINITY);
}
DOPF_HD void verify_note_gen(const View &v, int g, int t)
{
    const int slot = DOPF_ATOMIC_ADD_I32(&v.ctrl->gen_work_cnt, 1);
    if (slot < v.gen_work_cap) v.gen_work[slot] = g * v.T + t;
    else v.ctrl->error = DOPF_ERR_WORK_CAP;
}
DOPF_HD void verify_note_sto(const View &v, int s)
{
    if (DOPF_ATOMIC_EXCH_I32(&v.sto_flag[s], 1) == 0) {
        const int slot = DOPF_ATOMIC_ADD_I32(&v.ctrl->sto_work_cnt, 1);
        v.sto_work[slot] = s;
    }
}
DOPF_HD bool verify_gen_moved(const View &v, int g, int t, double lo, double hi)
{
    const int cur = v.ctrl->cur;
    const double d = sel(v.P, 1 - cur)[(size_t)g * v.T + t] - sel(v.P, cur)[(size_t)g * v.T + t];
    return d > hi || d < lo;
}
DOPF_HD bool verify_sto_moved(const View &v, int s, int t, double lo, double hi)
{
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const size_t i = (size_t)s * v.T + t;
    const double d = (sel(v.D, nxt)[i] - sel(v.D, cur)[i]) - (sel(v.C, nxt)[i] - sel(v.C, cur)[i]);
    return d > hi || d < lo;
}
DOPF_HD void body_verify(const View &v, int n, int col)
{
    double lo, hi;
    if (!verify_bounds(v, n, col, lo, hi)) return;
    const int c = v.scen_of_col(col), vn = c * v.N + n, t = col - c * v.T;
    for (int g = v.gen_ptr[vn]; g < v.gen_ptr[vn + 1]; ++g) if (verify_gen_moved(v, g, t, lo, hi)) verify_note_gen(v, g, t);
    for (int s = v.sto_ptr[vn]; s < v.sto_ptr[vn + 1]; ++s) if (verify_sto_moved(v, s, t, lo, hi)) verify_note_sto(v, s);
}

// ---- nodal injection of the new iterate (results.jl:64,88-106) --------------------------------
// injection of (n,t) and the node statistics st[0..7] of the moves (see View::nst)
DOPF_HD double inject_compute(const View &v, int n, int col, double (&st)[8])
{
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    double a = 0.0;
    double lo = 0.0, hi = 0.0, inneg = -INFINITY, inpos = INFINITY, sneg = 0.0, spos = 0.0, cneg = 0.0, cpos = 0.0;
    if (n < v.N && col < v.TC) {
        const int c = v.scen_of_col(col), vn = c * v.N + n, t = col - c * v.T;
        a = v.demand_on ? -v.demand[(size_t)n * v.ldt + col] : 0.0;
        const double *Pn = sel(v.P, nxt), *Pc = sel(v.P, cur);
        const int g1 = v.gen_ptr[vn + 1];
        int g = v.gen_ptr[vn];
#define DOPF_NOTE_MOVE(d)                                                                                   \
        if ((d) < 0.0) { lo = (d) < lo ? (d) : lo; inneg = (d) > inneg ? (d) : inneg; sneg += (d); cneg += 1.0; } \
        else if ((d) > 0.0) { hi = (d) > hi ? (d) : hi; inpos = (d) < inpos ? (d) : inpos; spos += (d); cpos += 1.0; }
        // 16 loads in flight before the dependent chain; the last (partial) batch re-reads its final row instead of falling
        // back to one load-wait per generator (the kernel is bound by load latency)
        for (; g < g1; g += 8) {
            double pn[8], pc[8];
            const int nv = g1 - g < 8 ? g1 - g : 8;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < 8; ++u) { const size_t o = (size_t)(g + (u < nv ? u : nv - 1)) * v.T + t; pn[u] = Pn[o]; pc[u] = Pc[o]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < 8; ++u) if (u < nv) { const double d = pn[u] - pc[u]; a += pn[u]; DOPF_NOTE_MOVE(d) }
        }
        const double *Dn = sel(v.D, nxt), *Dc = sel(v.D, cur), *Cn = sel(v.C, nxt), *Cc = sel(v.C, cur);
        const int s1 = v.sto_ptr[vn + 1];
        int s = v.sto_ptr[vn];
        for (; s < s1; s += 4) {
            double dn_[4], dc_[4], cn_[4], cc_[4];
            const int nv = s1 - s < 4 ? s1 - s : 4;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < 4; ++u) {
                const size_t o = (size_t)(s + (u < nv ? u : nv - 1)) * v.T + t;
                dn_[u] = Dn[o]; dc_[u] = Dc[o]; cn_[u] = Cn[o]; cc_[u] = Cc[o];
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < 4; ++u) if (u < nv) { const double d = (dn_[u] - dc_[u]) - (cn_[u] - cc_[u]); a += dn_[u] - cn_[u]; DOPF_NOTE_MOVE(d) }
        }
#undef DOPF_NOTE_MOVE
    }
    st[0] = lo; st[1] = hi; st[2] = inneg; st[3] = inpos; st[4] = sneg; st[5] = spos; st[6] = cneg; st[7] = cpos;
    return a;
}
DOPF_HD void body_inject(const View &v, int n, int t)
{
    double st[8];
    sel(v.injloc, 1 - v.ctrl->cur)[(size_t)n * v.ldt + t] = inject_compute(v, n, t, st);
    const size_t i = (size_t)t * v.Np + n;
    for (int k = 0; k < 8; ++k) v.nst[k][i] = st[k];
}

// closed form of sum_i (b + sp*delta_i)_+ over the agents of node n at time t from the node statistics;
// ok = false if the hinge threshold falls strictly between two movers on one side (then the caller sums
// over the agents).  Resting agents (delta = 0) contribute (b)_+ each.
// (the common statistics are passed in: k_slack_rows loads them for several nodes at once before the first use)
DOPF_HD double slack_node_closed_pre(const View &v, double b, double sp, int n, int t, double na, double lo, double hi, double s4, double s5, bool &ok)
{
    const size_t i = (size_t)t * v.Np + n;                 // t = column
    ok = true;
    if (na == 0.0) return 0.0;
    if (sp == 0.0) return na * pospart(b);
    const double f0 = b + sp * lo, f1 = b + sp * hi;
    const double sall = s4 + s5;
    if (f0 >= 0.0 && f1 >= 0.0) return na * b + sp * sall;            // nobody is clipped
    if (f0 <= 0.0 && f1 <= 0.0 && b <= 0.0) return 0.0;               // everybody is clipped (resting agents: b <= 0)
    // movers that push the term down are the negative ones for sp > 0, the positive ones for sp < 0
    const bool downneg = sp > 0.0;
    const double in_dn = downneg ? v.nst[2][i] : v.nst[3][i];         // the "down" mover closest to zero
    const double in_up = downneg ? v.nst[3][i] : v.nst[2][i];         // the "up" mover closest to zero
    const double s_up = downneg ? v.nst[5][i] : v.nst[4][i], c_up = downneg ? v.nst[7][i] : v.nst[6][i];
    const double s_dn = downneg ? v.nst[4][i] : v.nst[5][i], c_dn = downneg ? v.nst[6][i] : v.nst[7][i];
    if (b >= 0.0) {
        // resting and "up" movers keep a non-negative term; are ALL "down" movers clipped?
        if (c_dn == 0.0 || b + sp * in_dn <= 0.0) return (na - c_dn) * b + sp * s_up;
        // is NO "down" mover clipped?  (then f0/f1 test above would have caught it, kept for clarity)
        const double out_dn = downneg ? lo : hi;
        if (b + sp * out_dn >= 0.0) return na * b + sp * sall;
    } else {
        // resting and "down" movers are clipped; are ALL "up" movers unclipped?
        if (c_up == 0.0) return 0.0;
        if (b + sp * in_up >= 0.0) return c_up * b + sp * s_up;
        const double out_up = downneg ? hi : lo;
        if (b + sp * out_up <= 0.0) return 0.0;
    }
    (void)s_dn;
    ok = false;
    return 0.0;
}
DOPF_HD double slack_node_closed(const View &v, double b, double sp, int n, int t, bool &ok)
{
    const size_t i = (size_t)t * v.Np + n;
    return slack_node_closed_pre(v, b, sp, n, t, v.nagents[(size_t)v.scen_of_col(t) * v.Np + n], v.nst[0][i], v.nst[1][i], v.nst[4][i], v.nst[5][i], ok);
}

// contribution of node n to the closed form of an uncrossed row: every agent keeps the sign of b
DOPF_HD double slack_node_lin(const View &v, double b, double sp, int n, int t)
{
    if (!(b > 0.0)) return 0.0;
    const size_t i = (size_t)t * v.Np + n;
    return v.nagents[(size_t)v.scen_of_col(t) * v.Np + n] * b + sp * (v.nst[4][i] + v.nst[5][i]);
}

// ---- exact slack sums of one tight row at one node (results.jl:83-84,110-112) ------------------
// returns sum over the agents at node n of (b - p*delta_i)_+  (side 0)  or (b + p*delta_i)_+ (side 1)
DOPF_HD double body_slack_row_node(const View &v, int l, int side, int n, int t)
{
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const double p = v.ptdf[(size_t)l * v.Np + n];
    const double b = side ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
    const double sp = side ? p : -p;   // term is (b + sp*delta)_+
    bool ok;
    const double c = slack_node_closed(v, b, sp, n, t, ok);
    if (ok) return c;
    double a = 0.0;
    const int sc = v.scen_of_col(t), vn = sc * v.N + n, tl = t - sc * v.T;
    for (int g = v.gen_ptr[vn]; g < v.gen_ptr[vn + 1]; ++g) {
        const double d = sel(v.P, nxt)[(size_t)g * v.T + tl] - sel(v.P, cur)[(size_t)g * v.T + tl];
        a += pospart(b + sp * d);
    }
    for (int s = v.sto_ptr[vn]; s < v.sto_ptr[vn + 1]; ++s) {
        const size_t i = (size_t)s * v.T + tl;
        const double d = (sel(v.D, nxt)[i] - sel(v.D, cur)[i]) - (sel(v.C, nxt)[i] - sel(v.C, cur)[i]);
        a += pospart(b + sp * d);
    }
    return a;
}

// ---- dual update for one (line, t) (update_duals.jl:17-39) ------------------------------------
// is_tight: bit0 / bit1 = the exact row sums are available in rowsumU / rowsumK
DOPF_HD void body_dual(const View &v, int l, int t, int is_tight, double &res_mu, double &res_rho)
{
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const size_t i = (size_t)l * v.ldt + t;
    const double Fn = sel(v.flow, nxt)[i], dF = Fn - sel(v.flow, cur)[i];
    const double A = (double)v.A;
    const double bp = v.bplus[i], bm = v.bminus[i];
    // closed form of a row nobody crosses; for a tight row the device kernels store the CORRECTION to it (the nodes whose
    // agents can reach the hinge: exact minus closed form), the sequential emulation stores the whole exact sum
    const double cU = bp > 0.0 ? A * bp - dF : 0.0, cK = bm > 0.0 ? A * bm + dF : 0.0;
    const double sU = (is_tight & 1) ? (v.rowsum_is_corr ? cU + v.rowsumU[i] : v.rowsumU[i]) : cU;
    const double sK = (is_tight & 2) ? (v.rowsum_is_corr ? cK + v.rowsumK[i] : v.rowsumK[i]) : cK;
    const double scale = v.c.w2 / (v.c.kk * A);
    const double aU = scale * sU, aK = scale * sK;
    const double f = v.fmax[l];
    double mu = sel(v.mu, cur)[i] + v.c.gamma * (Fn + aU - f);
    double rho = sel(v.rho, cur)[i] + v.c.gamma * (aK - Fn - f);
    mu *= (aU <= v.c.mask_tol) ? 1.0 : 0.0;
    rho *= (aK <= v.c.mask_tol) ? 1.0 : 0.0;
    v.avgU[i] = aU; v.avgK[i] = aK;
    sel(v.mu, nxt)[i] = mu; sel(v.rho, nxt)[i] = rho;
    res_mu = fabs(mu - sel(v.mu, cur)[i]);
    res_rho = fabs(rho - sel(v.rho, cur)[i]);
}

// ---- lambda update (update_duals.jl:7-15) ------------------------------------------------------
DOPF_HD double body_lambda(const View &v, int t)
{
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const double ln = sel(v.lam, cur)[t] + v.c.gamma * sel(v.ssum, nxt)[t];
    sel(v.lam, nxt)[t] = ln;
    return fabs(ln - sel(v.lam, cur)[t]);
}

// ---- end of iteration (convergence.jl:1-31), per scenario -----------------------------------------
DOPF_HD void body_finish_scenario(const View &v, int c)
{
    if (v.sc_converged[c]) return;                     // frozen
    double r[3];
    for (int k = 0; k < 3; ++k) { r[k] = bits_nonneg(v.sc_res_bits[3 * c + k]); v.sc_res[3 * c + k] = r[k]; }
    if (v.sc_iteration[c] != 1) {
        for (int k = 0; k < 3; ++k) v.sc_conv[3 * c + k] = r[k] < v.c.eps;
        v.sc_converged[c] = v.sc_conv[3 * c] && v.sc_conv[3 * c + 1] && v.sc_conv[3 * c + 2];
    }
    if (!v.sc_converged[c]) v.sc_iteration[c] += 1;
}
// global part: the control block mirrors scenario 0 (the whole problem when C = 1); `converged` = all scenarios
DOPF_HD void body_finish_global(const View &v, int all)
{
    Ctrl *c = v.ctrl;
    for (int k = 0; k < 3; ++k) c->res[k] = v.sc_res[k];
    c->conv_lambda = v.sc_conv[0]; c->conv_mue = v.sc_conv[1]; c->conv_rho = v.sc_conv[2];
    c->iteration = v.sc_iteration[0];
    c->converged = all;
    c->cur = 1 - c->cur;
    c->iters_done += 1;
    c->finish_cnt = 0;
}
DOPF_HD void body_finish(const View &v)                // single problem
{
    for (int k = 0; k < 3; ++k) v.sc_res_bits[k] = v.ctrl->res_bits[k];
    body_finish_scenario(v, 0);
    body_finish_global(v, v.sc_converged[0]);
}

}  // namespace dopf
#endif
