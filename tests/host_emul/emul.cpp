// TEST-ONLY host build of the device math (decentralopf.jl_b200/csrc/dopf_math.h) with a lane
// group of width 1.  Lets the per-agent algorithms be checked against the oracle on a machine
// without a GPU.  Never loaded by the product package.
#include "../../decentralopf.jl_b200/csrc/dopf_math.h"
#include <vector>
#include <cstring>
using namespace dopf;

extern "C" {

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
    p.T = T; p.k.mc = mc; p.k.pmax = pmax; p.k.emax = emax; p.k.prox = prox;
    p.step = st.data(); p.hinges = hcap > 0 ? h.data() : nullptr; p.hcnt = hcnt; p.hcap = hcap;
    StoSolver<1> s(p);
    s.solve(eta);
    for (int t = 0; t < T; ++t) {
        StoEval e = sto_eval(p.step[t], p.k, p.list(t), eta[t]);
        D[t] = e.D; C[t] = e.C;
    }
    stats[0] = s.stats.evals; stats[1] = s.stats.solves; stats[2] = s.stats.segments;
}

double emul_gen_root(double c, double a, int n, const double *hbp, const double *hsg, double lo, double hi)
{
    std::vector<Hinge> h(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) { h[i].bp = hbp[i]; h[i].sg = hsg[i]; }
    HingeList l; l.h = h.data(); l.n = n;
    return root_monotone_pl(c, a, l, lo, hi);
}
}
