// TEST-ONLY host build of the device math (decentralopf.jl_b200/csrc/dopf_math.h) with a lane
// group of width 1.  Lets the per-agent algorithms be checked against the oracle on a machine
// without a GPU.  Never loaded by the product package.
#include "../../decentralopf.jl_b200/csrc/dopf_math.h"
#include <vector>
#include <cstring>
using namespace dopf;

extern "C" {

// clip table of one hinge-free storage step: which = 0 closed form (the device build), 1 evaluation-based reference;
// out = e[4] | nu0[5] | nus[5] | dy[5]
void emul_clip_table(double Db, double Cb, double g0, double s1, double mc, double pmax, double prox, int which, double *out)
{
    StoStep st; st.Db = Db; st.Cb = Cb; st.g0 = g0; st.s1 = s1;
    StoConst k; k.mc = mc; k.pmax = pmax; k.emax = 0.0; k.prox = prox; k.iprox = 1.0 / prox;
    double e[4], nu0[5], nus[5], dy[5];
    const double r1 = 1.0 / (prox + s1), r2 = 1.0 / (prox + 2.0 * s1);
    if (which == 0) sto_clip_table(st, k, r1, r2, e, nu0, nus, dy); else sto_clip_table_ref(st, k, r1, r2, e, nu0, nus, dy);
    for (int i = 0; i < 4; ++i) out[i] = e[i];
    for (int i = 0; i < 5; ++i) { out[4 + i] = nu0[i]; out[9 + i] = nus[i]; out[14 + i] = dy[i]; }
}
// D, C, dy of the step at eta through a table (same arithmetic as eval_tab in dopf_sto_warp.cuh)
void emul_clip_eval(const double *tab, double Db, double Cb, double mc, double pmax, double prox, double eta, double *out)
{
    const int p = (eta < tab[0]) + (eta < tab[1]) + (eta < tab[2]) + (eta < tab[3]);
    const double nu = tab[4 + p] + tab[9 + p] * eta;
    out[0] = clip01(Db - (mc + nu) / prox, pmax); out[1] = clip01(Cb - (mc - nu) / prox, pmax); out[2] = tab[14 + p]; out[3] = nu;
}

// storage solve with optional per-t hinge lists (hbp/hsg [T][hcap], hcnt [T]; hcap may be 0)
void emul_storage_solve(int T, double mc, double pmax, double emax, double prox,
                        const double *Db, const double *Cb, const double *g0, const double *s1,
                        int hcap, const int *hcnt, const double *hbp, const double *hsg,
                        double *D, double *C, double *eta, int *stats)
{
    std::vector<StoStep> st(T);
    for (int t = 0; t < T; ++t) { st[t].Db = Db[t]; st[t].Cb = Cb[t]; st[t].g0 = g0[t]; st[t].s1 = s1[t]; }
    std::vector<Hinge> h((size_t)T * (hcap > 0 ? hcap : 1));
    for (int i = 0; i < T * hcap; ++i) { h[i].bp = hbp[i]; h[i].sg = hsg[i]; }
    StoProblem p;
    p.T = T; p.k.mc = mc; p.k.pmax = pmax; p.k.emax = emax; p.k.prox = prox; p.k.iprox = 1.0 / prox;
    p.step = st.data(); p.hinges = hcap > 0 ? h.data() : nullptr; p.hcnt = hcnt; p.hcap = hcap;
    StoSolver<1> s(p);
    s.solve(eta);
    for (int t = 0; t < T; ++t) {
        StoEval e = sto_eval(p.step[t], p.k, p.list(t), eta[t]);
        D[t] = e.D; C[t] = e.C;
    }
    stats[0] = s.stats.evals; stats[1] = s.stats.solves; stats[2] = s.stats.segments; stats[3] = s.stats.passes;
}

struct SeqSteps {
    const StoProblem *p;
    StoStep step(int t) const { return p->step[t]; }
    HingeList list(int t) const { return p->list(t); }
};
void emul_storage_solve_seq(int T, double mc, double pmax, double emax, double prox,
                        const double *Db, const double *Cb, const double *g0, const double *s1,
                        int hcap, const int *hcnt, const double *hbp, const double *hsg,
                        double *D, double *C, double *eta, int *stats)
{
    std::vector<StoStep> st(T);
    for (int t = 0; t < T; ++t) { st[t].Db = Db[t]; st[t].Cb = Cb[t]; st[t].g0 = g0[t]; st[t].s1 = s1[t]; }
    std::vector<Hinge> h((size_t)T * (hcap > 0 ? hcap : 1));
    for (int i = 0; i < T * hcap; ++i) { h[i].bp = hbp[i]; h[i].sg = hsg[i]; }
    StoProblem p;
    p.T = T; p.k.mc = mc; p.k.pmax = pmax; p.k.emax = emax; p.k.prox = prox; p.k.iprox = 1.0 / prox;
    p.step = st.data(); p.hinges = hcap > 0 ? h.data() : nullptr; p.hcnt = hcnt; p.hcap = hcap;
    SeqSteps sp{&p};
    StoStats ss{0, 0, 0, 0};
    auto emit = [&](int t, double e) { eta[t] = e; StoEval ev = sto_eval(p.step[t], p.k, p.list(t), e); D[t] = ev.D; C[t] = ev.C; };
    sto_funnel_seq(sp, p.k, T, emit, &ss);
    stats[0] = ss.evals; stats[1] = ss.solves; stats[2] = ss.segments; stats[3] = ss.passes;
}

// warm-started solve: eta_io holds the previous multiplier path on entry, the new one on exit.
// returns 1 if the warm path verified, 0 if the cold funnel had to be used.
int emul_storage_solve_warm(int T, double mc, double pmax, double emax, double prox,
                        const double *Db, const double *Cb, const double *g0, const double *s1,
                        double *D, double *C, double *eta_io, const double *Eprev, int *stats)
{
    std::vector<StoStep> st(T);
    for (int t = 0; t < T; ++t) { st[t].Db = Db[t]; st[t].Cb = Cb[t]; st[t].g0 = g0[t]; st[t].s1 = s1[t]; }
    StoProblem p;
    p.T = T; p.k.mc = mc; p.k.pmax = pmax; p.k.emax = emax; p.k.prox = prox; p.k.iprox = 1.0 / prox;
    p.step = st.data(); p.hinges = nullptr; p.hcnt = nullptr; p.hcap = 0;
    SeqSteps sp{&p};
    StoStats ss{0, 0, 0, 0};
    std::vector<double> prev(eta_io, eta_io + T);
    auto emit = [&](int t, double e) { eta_io[t] = e; StoEval ev = sto_eval(p.step[t], p.k, p.list(t), e); D[t] = ev.D; C[t] = ev.C; };
    auto pv = [&](int t) { return prev[t]; };
    auto pe = [&](int t) { return Eprev[t]; };
    int warm = sto_warm_try(sp, p.k, T, pv, pe, emit, &ss) ? 1 : 0;
    if (!warm) sto_funnel_seq(sp, p.k, T, emit, &ss);
    stats[0] = ss.evals; stats[1] = ss.solves; stats[2] = ss.segments; stats[3] = ss.passes;
    return warm;
}

double emul_gen_root(double c, double a, int n, const double *hbp, const double *hsg, double lo, double hi)
{
    std::vector<Hinge> h(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) { h[i].bp = hbp[i]; h[i].sg = hsg[i]; }
    HingeList l; l.h = h.data(); l.n = n; l.sorted = false;
    return root_monotone_pl(c, a, l, lo, hi);
}
}

// ------------------------------------------------------------------------------------------------
// Sequential emulation of the whole iteration pipeline (mirror of enqueue_iteration in
// dopf_kernels.cu) on host arrays with the device layout.
// ------------------------------------------------------------------------------------------------
#include "../../decentralopf.jl_b200/csrc/dopf_bodies.h"
#include <algorithm>
#include <numeric>
#include <cmath>

struct Emul {
    std::vector<unsigned long long> scbits;
    View v;
    std::vector<std::vector<double>> d;   // owned double arrays
    std::vector<std::vector<int>> iv;
    std::vector<unsigned long long> dn, dmax;
    std::vector<unsigned char> flags, tflag;
    Ctrl ctrl;
    std::vector<Hinge> scratch;
    std::vector<int> hcnt;
    // partitioned mode (mirror of dopf_set_partition / dopf_step_phase): exchange buffers and constants
    std::vector<double> rbox, dmaxd;
    double *injx = nullptr, *xrow = nullptr, *flowD = nullptr;
    bool partitioned = false;
    double *mk(size_t n) { d.emplace_back(n ? n : 1, 0.0); return d.back().data(); }
    int *mki(size_t n) { iv.emplace_back(n ? n : 1, 0); return iv.back().data(); }
};

static int rup(int a, int b) { return (a + b - 1) / b * b; }

extern "C" {

// agents must already be sorted by node
void *emul_create(int N, int L, int T, int G, int S, const double *ptdf, const double *fmax, const double *demand,
                  const double *gmc, const double *gpm, const int *gnode,
                  const double *smc, const double *spm, const double *sem, const int *snode,
                  double gamma, double w, double prox, double mask_tol, double eps, int hcap)
{
    Emul *e = new Emul();
    View &v = e->v;
    memset(&v, 0, sizeof v);
    v.N = N; v.L = L; v.T = T; v.G = G; v.S = S; v.A = G + S; v.NS = 1; v.TC = T;
    v.ldt = rup(T, 32); v.Np = rup(N, 64); v.Lp = rup(L, 64); v.hcap = hcap; v.gen_work_cap = std::max(1, G * T);
    v.c = Coef::make(gamma, w, prox, mask_tol, eps);
    v.demand_on = 1;
    const int ldt = v.ldt, Np = v.Np, Lp = v.Lp;
    double *P = e->mk((size_t)Lp * Np), *f = e->mk(Lp), *dm = e->mk((size_t)Np * ldt), *q = e->mk(Np), *prow = e->mk(Lp), *mw = e->mk(Lp), *na = e->mk(Np);
    std::vector<double> rbox(Np, 0.0);
    for (int i = 0; i < G; ++i) { na[gnode[i]] += 1; rbox[gnode[i]] = std::max(rbox[gnode[i]], gpm[i]); }
    for (int i = 0; i < S; ++i) { na[snode[i]] += 1; rbox[snode[i]] = std::max(rbox[snode[i]], 2 * spm[i]); }
    for (int l = 0; l < L; ++l) {
        f[l] = fmax[l];
        for (int n = 0; n < N; ++n) {
            double a = ptdf[(size_t)l * N + n];
            P[(size_t)l * Np + n] = a; q[n] += a * a; prow[l] = std::max(prow[l], std::fabs(a)); mw[l] = std::max(mw[l], std::fabs(a) * rbox[n]);
        }
    }
    for (int n = 0; n < N; ++n) for (int t = 0; t < T; ++t) dm[(size_t)n * ldt + t] = demand[(size_t)n * T + t];
    v.ptdf = P; v.fmax = f; v.demand = dm; v.q = q; v.prow = prow; v.mwide = mw; v.nagents = na; v.rbox = nullptr;
    e->rbox = rbox;
    double *a;
    int *ip;
    a = e->mk(G); std::copy(gmc, gmc + G, a); v.gen_mc = a;
    a = e->mk(G); std::copy(gpm, gpm + G, a); v.gen_pmax = a;
    ip = e->mki(G); std::copy(gnode, gnode + G, ip); v.gen_node = ip;
    ip = e->mki(N + 1); for (int i = 0; i < G; ++i) ip[gnode[i] + 1]++; for (int n = 0; n < N; ++n) ip[n + 1] += ip[n]; v.gen_ptr = ip;
    a = e->mk(S); std::copy(smc, smc + S, a); v.sto_mc = a;
    a = e->mk(S); std::copy(spm, spm + S, a); v.sto_pmax = a;
    a = e->mk(S); std::copy(sem, sem + S, a); v.sto_emax = a;
    ip = e->mki(S); std::copy(snode, snode + S, ip); v.sto_node = ip;
    ip = e->mki(N + 1); for (int i = 0; i < S; ++i) ip[snode[i] + 1]++; for (int n = 0; n < N; ++n) ip[n + 1] += ip[n]; v.sto_ptr = ip;
    for (int k = 0; k < 2; ++k) {
        v.P[k] = e->mk((size_t)G * T); v.D[k] = e->mk((size_t)S * T); v.C[k] = e->mk((size_t)S * T);
        v.inj[k] = e->mk((size_t)Np * ldt); v.injloc[k] = v.inj[k]; v.ssum[k] = e->mk(ldt); v.flow[k] = e->mk((size_t)Lp * ldt);
        v.lam[k] = e->mk(ldt); v.mu[k] = e->mk((size_t)Lp * ldt); v.rho[k] = e->mk((size_t)Lp * ldt);
    }
    v.E = e->mk((size_t)S * T); v.eta = e->mk((size_t)S * T); v.cold_work = e->mki(S); v.wide_b = e->mk((size_t)T * 2 * L);
    { double *pt = e->mk((size_t)Np * Lp); for (int l = 0; l < L; ++l) for (int n = 0; n < N; ++n) pt[(size_t)n * Lp + l] = P[(size_t)l * Np + n]; v.ptdfT = pt; }
    v.avgU = e->mk((size_t)Lp * ldt); v.avgK = e->mk((size_t)Lp * ldt);
    v.bplus = e->mk((size_t)Lp * ldt); v.bminus = e->mk((size_t)Lp * ldt); v.M = e->mk((size_t)Lp * ldt); v.Wt = e->mk((size_t)Lp * ldt);
    v.g0 = e->mk((size_t)Np * ldt); v.s1 = e->mk((size_t)Np * ldt); v.rg = e->mk((size_t)Np * ldt);
    e->dn.assign((size_t)Np * ldt, 0); e->dmax.assign(ldt, 0); v.dn = e->dn.data(); v.dmax = e->dmax.data();
    for (int k = 0; k < 8; ++k) v.nst[k] = e->mk((size_t)Np * ldt);
    e->flags.assign((size_t)ldt * Lp, 0); v.flags = e->flags.data(); e->tflag.assign((size_t)Lp * ldt, 0);
    v.wide = e->mki((size_t)T * 2 * L); v.wcnt = e->mki(T); v.tight = e->mki((size_t)T * 2 * L); v.tcnt = e->mki(T);
    v.gen_work = e->mki(v.gen_work_cap); v.sto_work = e->mki(S); v.sto_flag = e->mki(S);
    v.rowsumU = e->mk((size_t)Lp * ldt); v.rowsumK = e->mk((size_t)Lp * ldt);
    memset(&e->ctrl, 0, sizeof e->ctrl); e->ctrl.iteration = 1; v.ctrl = &e->ctrl;
    v.sc_iteration = e->mki(1); v.sc_iteration[0] = 1; v.sc_converged = e->mki(1); v.sc_conv = e->mki(3);
    e->scbits.assign(3, 0ull); v.sc_res_bits = e->scbits.data(); v.sc_res = e->mk(3);
    e->scratch.resize((size_t)T * hcap); e->hcnt.resize(T);
    // initial state: inj = -demand (cur buffer), flows
    const int cur = 0;
    for (int n = 0; n < Np; ++n) for (int t = 0; t < ldt; ++t) v.inj[cur][(size_t)n * ldt + t] = -dm[(size_t)n * ldt + t];
    for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.inj[cur][(size_t)n * ldt + t]; v.ssum[cur][t] = s; }
    for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += P[(size_t)l * Np + n] * v.inj[cur][(size_t)n * ldt + t]; v.flow[cur][(size_t)l * ldt + t] = s; }
    return e;
}

void emul_destroy(void *h) { delete (Emul *)h; }
#ifdef DOPF_DEBUG_WARM
void emul_warm_fail(int *out) { for (int i = 0; i < 8; ++i) { out[i] = g_warm_fail[i]; g_warm_fail[i] = 0; } }
#endif

static void compact(Emul *e, int mode, bool keep_dmax = false)
{
    View &v = e->v;
    if (mode == 1 && !keep_dmax)
        for (int t = 0; t < v.ldt; ++t) { unsigned long long m = 0; for (int n = 0; n < v.N; ++n) m = std::max(m, v.dn[(size_t)n * v.ldt + t]); v.dmax[t] = m; }
    for (int t = 0; t < v.T; ++t) {
        int cnt = 0;
        if (mode == 0) {
            for (int l = 0; l < v.L; ++l) for (int side = 0; side < 2; ++side)
                if ((v.flags[(size_t)t * v.Lp + l] >> side) & 1) {
                    v.wide_b[(size_t)t * 2 * v.L + cnt] = side ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
                    v.wide[(size_t)t * 2 * v.L + cnt++] = l * 2 + side;
                }
            v.wcnt[t] = cnt;
        } else {
            const double dm = bits_nonneg(v.dmax[t]);
            for (int j = 0; j < v.wcnt[t]; ++j) {
                int en = v.wide[(size_t)t * 2 * v.L + j]; int l = en >> 1;
                double b = (en & 1) ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
                if (std::fabs(b) <= v.prow[l] * dm) v.tight[(size_t)t * 2 * v.L + cnt++] = en;
            }
            v.tcnt[t] = cnt;
        }
    }
}

static void storage_pass(Emul *e, bool fix)
{
    View &v = e->v;
    const int T = v.T;
    if (!fix) {
        v.ctrl->cold_work_cnt = 0;
        for (int s = 0; s < v.S; ++s) body_sto_warm(v, s);
        for (int w = 0; w < v.ctrl->cold_work_cnt; ++w) body_sto_cold(v, v.cold_work[w], nullptr, nullptr);
        v.ctrl->stat_sto_cold = v.ctrl->cold_work_cnt;
        return;
    }
    for (int w = 0; w < v.ctrl->sto_work_cnt; ++w) {
        const int s = v.sto_work[w], n = v.sto_node[s];
        const double pm = v.sto_pmax[s];
        for (int t = 0; t < T; ++t) {
            const double Db_ = sel(v.D, v.ctrl->cur)[(size_t)s * T + t], Cb_ = sel(v.C, v.ctrl->cur)[(size_t)s * T + t];
            const double rlo_ = -Db_ - (pm - Cb_), rhi_ = (pm - Db_) + Cb_;
            int cnt = 0;
            for (int j = 0; j < v.wcnt[t]; ++j) {
                int en = v.wide[(size_t)t * 2 * v.L + j]; Hinge h;
                if (make_hinge(v.c, v.ptdfT[(size_t)n * v.Lp + (en >> 1)], v.wide_b[(size_t)t * 2 * v.L + j], en & 1, h) && h.bp > rlo_ && h.bp < rhi_) {
                    if (cnt < v.hcap) e->scratch[(size_t)t * v.hcap + cnt] = h;
                    cnt++;
                }
            }
            if (cnt > v.hcap) { v.ctrl->error = DOPF_ERR_HINGE_CAP; cnt = v.hcap; }
            e->hcnt[t] = cnt;
        }
        body_sto_cold(v, s, e->scratch.data(), e->hcnt.data());
        v.ctrl->stat_sto_fix++;
    }
}

static void gen_fix_pass(Emul *e)
{
    View &v = e->v;
    const int cur = v.ctrl->cur, nxt = 1 - cur, ldt = v.ldt, Np = v.Np, T = v.T;
    for (int w = 0; w < v.ctrl->gen_work_cnt; ++w) {
        const int g = v.gen_work[w] / T, t = v.gen_work[w] % T, n = v.gen_node[g];
        const double Pb = v.P[cur][(size_t)g * T + t], pmax = v.gen_pmax[g], lo = -Pb, hi = pmax - Pb;
        std::vector<Hinge> lst;
        for (int j = 0; j < v.wcnt[t]; ++j) {
            int en = v.wide[(size_t)t * 2 * v.L + j]; int l = en >> 1, side = en & 1; Hinge hh;
            if (make_hinge(v.c, v.ptdf[(size_t)l * Np + n], side ? v.bminus[(size_t)l * ldt + t] : v.bplus[(size_t)l * ldt + t], side, hh) && hh.bp > lo && hh.bp < hi) lst.push_back(hh);
        }
        HingeList hl; hl.h = lst.data(); hl.n = (int)lst.size(); hl.sorted = false;
        const size_t nt = (size_t)n * ldt + t;
        double dd = root_monotone_pl(v.gen_mc[g] + v.g0[nt], v.c.prox + v.s1[nt], hl, lo, hi);
        double Pn = Pb + dd; Pn = Pn < 0 ? 0 : (Pn > pmax ? pmax : Pn);
        v.P[nxt][(size_t)g * T + t] = Pn;
        note_move(v, n, t, Pn - Pb);
        v.ctrl->stat_gen_fix++;
    }
}

void emul_iterate(void *h)
{
    Emul *e = (Emul *)h;
    View &v = e->v;
    if (v.ctrl->converged || v.ctrl->error) return;
    const int cur = v.ctrl->cur, nxt = 1 - cur, ldt = v.ldt, Np = v.Np, Lp = v.Lp, T = v.T;
    v.ctrl->gen_work_cnt = v.ctrl->sto_work_cnt = 0;
    v.ctrl->res_bits[0] = v.ctrl->res_bits[1] = v.ctrl->res_bits[2] = 0;
    std::fill(e->dn.begin(), e->dn.end(), 0ull); std::fill(e->dmax.begin(), e->dmax.end(), 0ull);
    for (int s = 0; s < v.S; ++s) v.sto_flag[s] = 0;
    for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) body_row_prep(v, l, t);
    compact(e, 0);
    for (int n = 0; n < Np; ++n) for (int t = 0; t < ldt; ++t) {
        double a = 0, b = 0;
        for (int l = 0; l < Lp; ++l) { double p = v.ptdf[(size_t)l * Np + n]; a += p * v.M[(size_t)l * ldt + t]; b += p * p * v.Wt[(size_t)l * ldt + t]; }
        v.g0[(size_t)n * ldt + t] = v.lam[cur][t] + v.c.gamma * v.ssum[cur][t] + a;
        v.s1[(size_t)n * ldt + t] = v.c.gamma + 2.0 * v.c.kappa * v.q[n] + b;
        v.rg[(size_t)n * ldt + t] = 1.0 / (v.c.prox + v.s1[(size_t)n * ldt + t]);
    }
    for (int g = 0; g < v.G; ++g) for (int t = 0; t < T; ++t) {
        const double pp = v.P[cur][(size_t)g * T + t];
        const double pn = body_gen_predict(v, g, t, pp, v.gen_node[g], v.gen_mc[g], v.gen_pmax[g]);
        v.P[nxt][(size_t)g * T + t] = pn;
        note_move(v, v.gen_node[g], t, pn - pp);
    }
    storage_pass(e, false);
    compact(e, 1);
    for (int n = 0; n < v.N; ++n) for (int t = 0; t < T; ++t) body_verify(v, n, t);
    gen_fix_pass(e);
    storage_pass(e, true);
    compact(e, 1);
    for (int n = 0; n < Np; ++n) for (int t = 0; t < ldt; ++t) body_inject(v, n, t);
    for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.inj[nxt][(size_t)n * ldt + t]; v.ssum[nxt][t] = s; }
    for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.ptdf[(size_t)l * Np + n] * v.inj[nxt][(size_t)n * ldt + t]; v.flow[nxt][(size_t)l * ldt + t] = s; }
    std::fill(e->tflag.begin(), e->tflag.end(), 0);
    for (int t = 0; t < T; ++t) for (int j = 0; j < v.tcnt[t]; ++j) {
        int en = v.tight[(size_t)t * 2 * v.L + j]; int l = en >> 1, side = en & 1;
        double a = 0; for (int n = 0; n < v.N; ++n) a += body_slack_row_node(v, l, side, n, t);
        (side ? v.rowsumK : v.rowsumU)[(size_t)l * ldt + t] = a;
        e->tflag[(size_t)l * ldt + t] |= (1 << side);
    }
    double rm = 0, rr = 0, rl = 0;
    for (int l = 0; l < v.L; ++l) for (int t = 0; t < T; ++t) { double a, b; body_dual(v, l, t, e->tflag[(size_t)l * ldt + t], a, b); rm = std::max(rm, a); rr = std::max(rr, b); }
    for (int t = 0; t < T; ++t) rl = std::max(rl, body_lambda(v, t));
    v.ctrl->res_bits[0] = nonneg_bits(rl); v.ctrl->res_bits[1] = nonneg_bits(rm); v.ctrl->res_bits[2] = nonneg_bits(rr);
    int tr = 0, wr = 0; for (int t = 0; t < T; ++t) { tr += v.tcnt[t]; wr += v.wcnt[t]; }
    v.ctrl->stat_tight_rows = tr; v.ctrl->stat_wide_rows = wr;
    body_finish(v);
}

// ---- partitioned mode: the four phases of dopf_step_phase with the three exchange buffers of dopf_exchange_buffer -----
// (same per-element bodies as the device kernels; the row sums are stored as corrections like k_slack_rows does)
void emul_partition_init(void *h, int total_agents)
{
    Emul *e = (Emul *)h; View &v = e->v;
    const int ldt = v.ldt, Np = v.Np, Lp = v.Lp;
    v.A = total_agents; v.demand_on = 0; v.rowsum_is_corr = 1; e->partitioned = true;
    e->injx = e->mk((size_t)Np * ldt); v.injloc[0] = v.injloc[1] = e->injx;
    e->xrow = e->mk((size_t)3 * Lp * ldt); v.rowsumU = e->xrow; v.rowsumK = e->xrow + (size_t)Lp * ldt; v.xflow = e->xrow + (size_t)2 * Lp * ldt;
    e->flowD = e->mk((size_t)Lp * ldt); v.flowD = e->flowD;
    for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.ptdf[(size_t)l * Np + n] * v.demand[(size_t)n * ldt + t]; e->flowD[(size_t)l * ldt + t] = s; }
    e->dmaxd.assign(ldt, 0.0);
}
// which: 0 move maxima [ldt] (MAX), 1 local injection [Np*ldt] (SUM), 2 row-sum corrections + partial flows [3*Lp*ldt] (SUM),
// 3 per-node box ranges [Np] (MAX, set-up)
double *emul_exchange_buffer(void *h, int which, long long *count)
{
    Emul *e = (Emul *)h; View &v = e->v;
    switch (which) {
    case 0: *count = v.ldt; return e->dmaxd.data();
    case 1: *count = (long long)v.Np * v.ldt; return e->injx;
    case 2: *count = (long long)3 * v.Lp * v.ldt; return e->xrow;
    default: *count = v.Np; return e->rbox.data();
    }
}
void emul_partition_finish_setup(void *h)      // after the box ranges were max-reduced: identical candidate rows on all ranks
{
    Emul *e = (Emul *)h; View &v = e->v;
    for (int l = 0; l < v.L; ++l) { double m = 0; for (int n = 0; n < v.N; ++n) m = std::max(m, std::fabs(v.ptdf[(size_t)l * v.Np + n]) * e->rbox[n]); v.mwide[l] = m; }
}
void emul_phase(void *h, int phase)
{
    Emul *e = (Emul *)h; View &v = e->v;
    if (v.ctrl->converged || v.ctrl->error) return;
    const int cur = v.ctrl->cur, nxt = 1 - cur, ldt = v.ldt, Np = v.Np, Lp = v.Lp, T = v.T;
    if (phase == 0) {
        v.ctrl->gen_work_cnt = v.ctrl->sto_work_cnt = 0;
        v.ctrl->res_bits[0] = v.ctrl->res_bits[1] = v.ctrl->res_bits[2] = 0;
        std::fill(e->dn.begin(), e->dn.end(), 0ull); std::fill(e->dmax.begin(), e->dmax.end(), 0ull);
        for (int s = 0; s < v.S; ++s) v.sto_flag[s] = 0;
        for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) body_row_prep(v, l, t);
        std::fill(e->xrow, e->xrow + (size_t)3 * Lp * ldt, 0.0);                 // k_row_prep zeroes the exchange slabs
        compact(e, 0);
        for (int n = 0; n < Np; ++n) for (int t = 0; t < ldt; ++t) {
            double a = 0, b = 0;
            for (int l = 0; l < Lp; ++l) { double p = v.ptdf[(size_t)l * Np + n]; a += p * v.M[(size_t)l * ldt + t]; b += p * p * v.Wt[(size_t)l * ldt + t]; }
            v.g0[(size_t)n * ldt + t] = v.lam[cur][t] + v.c.gamma * v.ssum[cur][t] + a;
            v.s1[(size_t)n * ldt + t] = v.c.gamma + 2.0 * v.c.kappa * v.q[n] + b;
            v.rg[(size_t)n * ldt + t] = 1.0 / (v.c.prox + v.s1[(size_t)n * ldt + t]);
        }
        for (int g = 0; g < v.G; ++g) for (int t = 0; t < T; ++t) {
            const double pp = v.P[cur][(size_t)g * T + t];
            const double pn = body_gen_predict(v, g, t, pp, v.gen_node[g], v.gen_mc[g], v.gen_pmax[g]);
            v.P[nxt][(size_t)g * T + t] = pn;
            note_move(v, v.gen_node[g], t, pn - pp);
        }
        storage_pass(e, false);
        compact(e, 1);                                   // local maxima: enough for the rank's own agents
        for (int n = 0; n < v.N; ++n) for (int t = 0; t < T; ++t) body_verify(v, n, t);
        gen_fix_pass(e);
        storage_pass(e, true);
        for (int t = 0; t < ldt; ++t) { unsigned long long m = 0; for (int n = 0; n < v.N; ++n) m = std::max(m, v.dn[(size_t)n * ldt + t]); e->dmaxd[t] = bits_nonneg(m); }
    } else if (phase == 1) {
        for (int t = 0; t < ldt; ++t) v.dmax[t] = nonneg_bits(e->dmaxd[t]);      // exchanged maxima -> identical tight lists
        compact(e, 1, true);
        for (int n = 0; n < Np; ++n) for (int t = 0; t < ldt; ++t) body_inject(v, n, t);     // agents only (demand_on = 0)
        for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.ptdf[(size_t)l * Np + n] * e->injx[(size_t)n * ldt + t]; v.xflow[(size_t)l * ldt + t] = s; }
    } else if (phase == 2) {
        for (size_t i = 0; i < (size_t)Np * ldt; ++i) v.inj[nxt][i] = e->injx[i] - v.demand[i];                 // launch_copy_inj
        for (int t = 0; t < ldt; ++t) { double s = 0; for (int n = 0; n < Np; ++n) s += v.inj[nxt][(size_t)n * ldt + t]; v.ssum[nxt][t] = s; }
        std::fill(e->tflag.begin(), e->tflag.end(), 0);
        for (int t = 0; t < T; ++t) for (int j = 0; j < v.tcnt[t]; ++j) {
            const int en = v.tight[(size_t)t * 2 * v.L + j], l = en >> 1, side = en & 1;
            const double b = side ? v.bminus[(size_t)l * ldt + t] : v.bplus[(size_t)l * ldt + t];
            double a = 0;
            for (int n = 0; n < v.N; ++n) {              // k_slack_rows: only the nodes whose largest move reaches the hinge
                const double p = v.ptdf[(size_t)l * Np + n], sp = side ? p : -p;
                const double dnm = std::max(-v.nst[0][(size_t)t * Np + n], v.nst[1][(size_t)t * Np + n]);
                if (std::fabs(b) > std::fabs(p) * dnm) continue;
                a += body_slack_row_node(v, l, side, n, t) - slack_node_lin(v, b, sp, n, t);
            }
            (side ? v.rowsumK : v.rowsumU)[(size_t)l * ldt + t] = a;
            e->tflag[(size_t)l * ldt + t] |= (1 << side);
        }
    } else {
        for (int l = 0; l < Lp; ++l) for (int t = 0; t < ldt; ++t) v.flow[nxt][(size_t)l * ldt + t] = v.xflow[(size_t)l * ldt + t] - e->flowD[(size_t)l * ldt + t];   // k_dual epilogue
        double rm = 0, rr = 0, rl = 0;
        for (int l = 0; l < v.L; ++l) for (int t = 0; t < T; ++t) { double a, b; body_dual(v, l, t, e->tflag[(size_t)l * ldt + t], a, b); rm = std::max(rm, a); rr = std::max(rr, b); }
        for (int t = 0; t < T; ++t) rl = std::max(rl, body_lambda(v, t));
        v.ctrl->res_bits[0] = nonneg_bits(rl); v.ctrl->res_bits[1] = nonneg_bits(rm); v.ctrl->res_bits[2] = nonneg_bits(rr);
        int tr = 0, wr = 0; for (int t = 0; t < T; ++t) { tr += v.tcnt[t]; wr += v.wcnt[t]; }
        v.ctrl->stat_tight_rows = tr; v.ctrl->stat_wide_rows = wr;
        body_finish(v);
    }
}

// newest iterate; matrices dense [rows][T]
void emul_get(void *h, double *P, double *D, double *C, double *E, double *inj, double *flow, double *avgU, double *avgK,
              double *lam, double *mu, double *rho, int *status /*iteration, converged, gen_fix, sto_fix, tight, wide, error, sto_cold*/)
{
    Emul *e = (Emul *)h; View &v = e->v; const int k = v.ctrl->cur, T = v.T, ldt = v.ldt;
    std::copy(v.P[k], v.P[k] + (size_t)v.G * T, P);
    std::copy(v.D[k], v.D[k] + (size_t)v.S * T, D); std::copy(v.C[k], v.C[k] + (size_t)v.S * T, C); std::copy(v.E, v.E + (size_t)v.S * T, E);
    for (int n = 0; n < v.N; ++n) for (int t = 0; t < T; ++t) inj[(size_t)n * T + t] = v.inj[k][(size_t)n * ldt + t];
    for (int l = 0; l < v.L; ++l) for (int t = 0; t < T; ++t) {
        size_t o = (size_t)l * T + t, i = (size_t)l * ldt + t;
        flow[o] = v.flow[k][i]; avgU[o] = v.avgU[i]; avgK[o] = v.avgK[i]; mu[o] = v.mu[k][i]; rho[o] = v.rho[k][i];
    }
    for (int t = 0; t < T; ++t) lam[t] = v.lam[k][t];
    status[0] = v.ctrl->iteration; status[1] = v.ctrl->converged; status[2] = v.ctrl->stat_gen_fix; status[3] = v.ctrl->stat_sto_fix;
    status[4] = v.ctrl->stat_tight_rows; status[5] = v.ctrl->stat_wide_rows; status[6] = v.ctrl->error; status[7] = v.ctrl->stat_sto_cold;
}
}
