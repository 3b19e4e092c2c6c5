/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  See dopf_oracle.h for scope and reference citations.
 *
 * Two independent restatements of the agent subproblems are provided:
 *   mode 1 "literal": variables and penalty expressions exactly as the reference builds them
 *          (P | D,C and one private copy of U[L,T], K[L,T] per agent; subproblems.jl:26-31,
 *          114-121; penalty_terms.jl:3-52), assembled as a dense QP and handed to the generic
 *          solver in qp_gi.c.  E is eliminated through E_t = sum_{tau<=t}(C-D)
 *          (subproblems.jl:150-156).  Only usable for tiny L*T.
 *   mode 0 "reduced": U,K eliminated analytically for fixed net-injection change delta
 *          (SURVEY.md Appendix A.2).  Generators: exact root of the monotone piecewise-linear
 *          optimality condition by sorted breakpoints.  Storages: Newton on the slack-clip
 *          pattern with an exact line search; each model QP (2T variables) goes to qp_gi.c.
 */
#include "dopf_oracle.h"
#include "qp_gi.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the timed baseline asks for all host cores explicitly */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

typedef struct {
    const oracle_problem *p;
    const oracle_state *s;
    double w2, k, kappa;   /* 2w, 2w+gamma, 2w*gamma/k */
    double *Sbar;          /* [T]    sum_n prev injection                      */
    double *pi;            /* [N][T] lambda_t + sum_l ptdf[l,n](mu-rho)[l,t]    */
    double *ap, *am;       /* [L][T] f - Fbar,  f + Fbar                        */
} ctx_t;

/* ---------- exact root of a monotone non-decreasing piecewise-linear function ---------- */
typedef double (*fun1d)(void *, double);

static int cmp_d(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* root of f on [lo,hi]; bp = breakpoints (any order, modified). clipped if no sign change */
static double root_pl(fun1d f, void *ud, double lo, double hi, double *bp, int nb)
{
    double flo = f(ud, lo);
    if (flo >= 0.0) return lo;
    double fhi = f(ud, hi);
    if (fhi <= 0.0) return hi;
    qsort(bp, nb, sizeof(double), cmp_d);
    int a = 0, b = nb - 1;
    while (a < nb && bp[a] <= lo) ++a;
    while (b >= 0 && bp[b] >= hi) --b;
    /* invariant: f(lo)<0<f(hi); breakpoints a..b strictly inside */
    while (a <= b) {
        int mid = (a + b) / 2;
        double fm = f(ud, bp[mid]);
        if (fm == 0.0) return bp[mid];
        if (fm < 0.0) { lo = bp[mid]; flo = fm; a = mid + 1; }
        else { hi = bp[mid]; fhi = fm; b = mid - 1; }
    }
    if (fhi == flo) return lo;
    return lo - flo * (hi - lo) / (fhi - flo);
}

/* ---------- slack elimination (Appendix A.2) ---------- */
static inline double ustar(const ctx_t *c, int l, int t, double pd)
{
    const oracle_problem *p = c->p;
    double v = (c->w2 * (c->ap[l * p->T + t] - pd) + p->gamma * c->s->avgU[l * p->T + t]) / c->k;
    return v > 0.0 ? v : 0.0;
}
static inline double kstar(const ctx_t *c, int l, int t, double pd)
{
    const oracle_problem *p = c->p;
    double v = (c->w2 * (c->am[l * p->T + t] + pd) + p->gamma * c->s->avgK[l * p->T + t]) / c->k;
    return v > 0.0 ? v : 0.0;
}
/* h_{n,t}(delta) = d/d delta of the flow+slack penalty after eliminating U,K */
static double hfun(const ctx_t *c, int n, int t, double delta)
{
    const oracle_problem *p = c->p;
    double s = 0.0;
    for (int l = 0; l < p->L; ++l) {
        double pl = p->ptdf[(size_t)l * p->N + n];
        if (pl == 0.0) continue;
        double pd = pl * delta;
        double U = ustar(c, l, t, pd), K = kstar(c, l, t, pd);
        s += c->w2 * pl * ((U - c->ap[l * p->T + t] + pd) - (K - c->am[l * p->T + t] - pd));
    }
    return s;
}
static int breakpoints(const ctx_t *c, int n, int t, double *bp)
{
    const oracle_problem *p = c->p;
    int nb = 0;
    for (int l = 0; l < p->L; ++l) {
        double pl = p->ptdf[(size_t)l * p->N + n];
        if (pl == 0.0) continue;
        int lt = l * p->T + t;
        bp[nb++] = (c->ap[lt] + p->gamma * c->s->avgU[lt] / c->w2) / pl;
        bp[nb++] = -(c->am[lt] + p->gamma * c->s->avgK[lt] / c->w2) / pl;
    }
    return nb;
}

/* ---------- generator, reduced ---------- */
typedef struct { const ctx_t *c; int n, t; double c0; } gen_ud;
static double gen_f(void *v, double delta)
{
    gen_ud *u = v;
    const oracle_problem *p = u->c->p;
    return u->c0 + (p->gamma + p->prox_weight) * delta + hfun(u->c, u->n, u->t, delta);
}
static void solve_generator_reduced(const ctx_t *c, int g, double *Pout, double *bpbuf)
{
    const oracle_problem *p = c->p;
    int n = p->gen_node[g], T = p->T;
    for (int t = 0; t < T; ++t) {
        double Pb = c->s->P[(size_t)g * T + t];
        gen_ud u = { c, n, t, p->gen_mc[g] + c->pi[n * T + t] + p->gamma * c->Sbar[t] };
        int nb = breakpoints(c, n, t, bpbuf);
        double d = root_pl(gen_f, &u, -Pb, p->gen_pmax[g] - Pb, bpbuf, nb);
        double v = Pb + d;
        if (v < 0.0) v = 0.0;
        if (v > p->gen_pmax[g]) v = p->gen_pmax[g];
        Pout[t] = v;
    }
}

/* ---------- dense QP assembly helper:  obj += c * (sum_i val_i x_idx_i + k)^2 ---------- */
static void add_square(int n, double *G, double *g, double c, const int *idx, const double *val,
                       int cnt, double k)
{
    for (int a = 0; a < cnt; ++a) {
        g[idx[a]] += 2.0 * c * k * val[a];
        for (int b = 0; b < cnt; ++b) G[(size_t)idx[a] * n + idx[b]] += 2.0 * c * val[a] * val[b];
    }
}

static void note_kkt(oracle_state *s, double r)
{
#ifdef _OPENMP
#pragma omp critical(dopf_kkt)
#endif
    if (r > s->qp_kkt_worst) s->qp_kkt_worst = r;
}

/* ---------- generator, literal (subproblems.jl:19-105) ---------- */
static int solve_generator_literal(const ctx_t *c, oracle_state *st, int g, double *Pout,
                                   double *Uout, double *Kout)
{
    const oracle_problem *p = c->p;
    const int L = p->L, T = p->T, nd = p->gen_node[g];
    const int n = 1 + 2 * L, m = 2 + 2 * L;
    double *G = malloc(sizeof(double) * n * n), *gv = malloc(sizeof(double) * n);
    double *C = malloc(sizeof(double) * m * n), *b = malloc(sizeof(double) * m);
    double *x = malloc(sizeof(double) * n), *u = malloc(sizeof(double) * m);
    int rc = 0;
    for (int t = 0; t < T; ++t) {
        memset(G, 0, sizeof(double) * n * n);
        memset(gv, 0, sizeof(double) * n);
        memset(C, 0, sizeof(double) * m * n);
        double Pb = c->s->P[(size_t)g * T + t];
        /* P*mc + P*(lambda + sum ptdf (mu-rho)) */
        gv[0] += p->gen_mc[g] + c->pi[nd * T + t];
        /* gamma/2 * (sum_n injection)^2 ; injection_node = P + (prev_inj_node - prevP) */
        { int i0 = 0; double v = 1.0; add_square(n, G, gv, 0.5 * p->gamma, &i0, &v, 1, c->Sbar[t] - Pb); }
        for (int l = 0; l < L; ++l) {
            double pl = p->ptdf[(size_t)l * p->N + nd];
            double Fb = c->s->flow[l * T + t];
            int idx[2]; double val[2];
            /* 10*(sum_n ptdf inj + U - fmax)^2 */
            idx[0] = 0; val[0] = pl; idx[1] = 1 + l; val[1] = 1.0;
            add_square(n, G, gv, p->flow_weight, idx, val, 2, Fb - pl * Pb - p->fmax[l]);
            /* 10*(K - sum_n ptdf inj - fmax)^2 */
            idx[0] = 0; val[0] = -pl; idx[1] = 1 + L + l; val[1] = 1.0;
            add_square(n, G, gv, p->flow_weight, idx, val, 2, -Fb + pl * Pb - p->fmax[l]);
            /* gamma/2 (U-avgU)^2, gamma/2 (K-avgK)^2 */
            idx[0] = 1 + l; val[0] = 1.0;
            add_square(n, G, gv, 0.5 * p->gamma, idx, val, 1, -c->s->avgU[l * T + t]);
            idx[0] = 1 + L + l;
            add_square(n, G, gv, 0.5 * p->gamma, idx, val, 1, -c->s->avgK[l * T + t]);
        }
        /* 1/2 (P-prevP)^2 */
        { int i0 = 0; double v = 1.0; add_square(n, G, gv, 0.5 * p->prox_weight, &i0, &v, 1, -Pb); }
        C[0 * n + 0] = 1.0; b[0] = 0.0;
        C[1 * n + 0] = -1.0; b[1] = -p->gen_pmax[g];
        for (int j = 0; j < 2 * L; ++j) { C[(size_t)(2 + j) * n + 1 + j] = 1.0; b[2 + j] = 0.0; }
        int it = qp_gi_solve(n, m, G, gv, C, b, x, u);
        if (it < 0) { rc = it; break; }
        note_kkt(st, qp_kkt_residual(n, m, G, gv, C, b, x, u));
        Pout[t] = x[0];
        for (int l = 0; l < L; ++l) { Uout[l * T + t] = x[1 + l]; Kout[l * T + t] = x[1 + L + l]; }
    }
    free(G); free(gv); free(C); free(b); free(x); free(u);
    return rc;
}

/* constraints shared by both storage forms: rows for D,C boxes and the level bounds.
 * variables 0..T-1 = D, T..2T-1 = C (others untouched). returns rows written (6T). */
static int storage_constraints(int T, int n, double pmax, double emax, double *C, double *b)
{
    int r = 0;
    for (int t = 0; t < T; ++t) {
        C[(size_t)r * n + t] = 1.0; b[r++] = 0.0;
        C[(size_t)r * n + t] = -1.0; b[r++] = -pmax;
        C[(size_t)r * n + T + t] = 1.0; b[r++] = 0.0;
        C[(size_t)r * n + T + t] = -1.0; b[r++] = -pmax;
    }
    for (int t = 0; t < T; ++t) { /* E_t = sum_{tau<=t} (C-D)  in [0, emax] */
        for (int tau = 0; tau <= t; ++tau) { C[(size_t)r * n + T + tau] = 1.0; C[(size_t)r * n + tau] = -1.0; }
        b[r++] = 0.0;
        for (int tau = 0; tau <= t; ++tau) { C[(size_t)r * n + T + tau] = -1.0; C[(size_t)r * n + tau] = 1.0; }
        b[r++] = -emax;
    }
    return r;
}

/* ---------- storage, literal (subproblems.jl:107-207) ---------- */
static int solve_storage_literal(const ctx_t *c, oracle_state *st, int s, double *Dout, double *Cout,
                                 double *Uout, double *Kout)
{
    const oracle_problem *p = c->p;
    const int L = p->L, T = p->T, nd = p->sto_node[s];
    const int n = (2 + 2 * L) * T, m = 6 * T + 2 * L * T;
    double *G = calloc((size_t)n * n, sizeof(double)), *gv = calloc(n, sizeof(double));
    double *C = calloc((size_t)m * n, sizeof(double)), *b = calloc(m, sizeof(double));
    double *x = malloc(sizeof(double) * n), *u = malloc(sizeof(double) * m);
#define IU(l, t) (2 * T + (l) * T + (t))
#define IK(l, t) (2 * T + L * T + (l) * T + (t))
    for (int t = 0; t < T; ++t) {
        double Db = c->s->D[(size_t)s * T + t], Cb = c->s->C[(size_t)s * T + t];
        double pr = c->pi[nd * T + t];
        gv[t] += p->sto_mc[s] + pr;       /* mc*(D+C) + (D-C)*price */
        gv[T + t] += p->sto_mc[s] - pr;
        int idx[3]; double val[3];
        idx[0] = t; val[0] = 1.0; idx[1] = T + t; val[1] = -1.0;
        add_square(n, G, gv, 0.5 * p->gamma, idx, val, 2, c->Sbar[t] - Db + Cb);
        for (int l = 0; l < L; ++l) {
            double pl = p->ptdf[(size_t)l * p->N + nd];
            double Fb = c->s->flow[l * T + t];
            double base = Fb - pl * (Db - Cb);
            idx[0] = t; val[0] = pl; idx[1] = T + t; val[1] = -pl; idx[2] = IU(l, t); val[2] = 1.0;
            add_square(n, G, gv, p->flow_weight, idx, val, 3, base - p->fmax[l]);
            idx[0] = t; val[0] = -pl; idx[1] = T + t; val[1] = pl; idx[2] = IK(l, t); val[2] = 1.0;
            add_square(n, G, gv, p->flow_weight, idx, val, 3, -base - p->fmax[l]);
            idx[0] = IU(l, t); val[0] = 1.0;
            add_square(n, G, gv, 0.5 * p->gamma, idx, val, 1, -c->s->avgU[l * T + t]);
            idx[0] = IK(l, t);
            add_square(n, G, gv, 0.5 * p->gamma, idx, val, 1, -c->s->avgK[l * T + t]);
        }
        idx[0] = t; val[0] = 1.0;
        add_square(n, G, gv, 0.5 * p->prox_weight, idx, val, 1, -Db);
        idx[0] = T + t;
        add_square(n, G, gv, 0.5 * p->prox_weight, idx, val, 1, -Cb);
    }
    int r = storage_constraints(T, n, p->sto_pmax[s], p->sto_emax[s], C, b);
    for (int j = 0; j < 2 * L * T; ++j) { C[(size_t)r * n + 2 * T + j] = 1.0; b[r++] = 0.0; }
    int it = qp_gi_solve(n, m, G, gv, C, b, x, u);
    if (it >= 0) {
        note_kkt(st, qp_kkt_residual(n, m, G, gv, C, b, x, u));
        for (int t = 0; t < T; ++t) { Dout[t] = x[t]; Cout[t] = x[T + t]; }
        for (int l = 0; l < L; ++l)
            for (int t = 0; t < T; ++t) { Uout[l * T + t] = x[IU(l, t)]; Kout[l * T + t] = x[IK(l, t)]; }
    }
#undef IU
#undef IK
    free(G); free(gv); free(C); free(b); free(x); free(u);
    return it < 0 ? it : 0;
}

/* ---------- storage, reduced ---------- */
/* linear model of h_{n,t} around the slack-clip pattern at delta:  h = c0 + s0*delta' */
static void hmodel(const ctx_t *c, int n, int t, double delta, double *c0, double *s0)
{
    const oracle_problem *p = c->p;
    double cc = 0.0, ss = 0.0;
    for (int l = 0; l < p->L; ++l) {
        double pl = p->ptdf[(size_t)l * p->N + n];
        if (pl == 0.0) continue;
        int lt = l * p->T + t;
        double pd = pl * delta;
        if (ustar(c, l, t, pd) > 0.0) { cc += c->kappa * pl * (c->s->avgU[lt] - c->ap[lt]); ss += c->kappa * pl * pl; }
        else { cc += -c->w2 * pl * c->ap[lt]; ss += c->w2 * pl * pl; }
        if (kstar(c, l, t, pd) > 0.0) { cc += -c->kappa * pl * (c->s->avgK[lt] - c->am[lt]); ss += c->kappa * pl * pl; }
        else { cc += c->w2 * pl * c->am[lt]; ss += c->w2 * pl * pl; }
    }
    *c0 = cc; *s0 = ss;
}

typedef struct {
    const ctx_t *c; int s; const double *x, *d; /* x,d: [2T] */
} ls_ud;
/* directional derivative of the true reduced storage objective at x + alpha d */
static double ls_f(void *v, double alpha)
{
    ls_ud *u = v;
    const oracle_problem *p = u->c->p;
    const int T = p->T, s = u->s, n = p->sto_node[s];
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
        double Db = u->c->s->D[(size_t)s * T + t], Cb = u->c->s->C[(size_t)s * T + t];
        double D = u->x[t] + alpha * u->d[t], C = u->x[T + t] + alpha * u->d[T + t];
        double delta = (D - Db) - (C - Cb);
        double net = u->c->pi[n * T + t] + p->gamma * (u->c->Sbar[t] + delta) + hfun(u->c, n, t, delta);
        acc += (p->sto_mc[s] + p->prox_weight * (D - Db) + net) * u->d[t];
        acc += (p->sto_mc[s] + p->prox_weight * (C - Cb) - net) * u->d[T + t];
    }
    return acc;
}

static int solve_storage_reduced(const ctx_t *c, oracle_state *st, int s, double *Dout, double *Cout)
{
    const oracle_problem *p = c->p;
    const int T = p->T, L = p->L, nd = p->sto_node[s];
    const int n = 2 * T, m = 6 * T;
    double *G = malloc(sizeof(double) * n * n), *gv = malloc(sizeof(double) * n);
    double *Cm = calloc((size_t)m * n, sizeof(double)), *b = calloc(m, sizeof(double));
    double *x = malloc(sizeof(double) * n), *xq = malloc(sizeof(double) * n), *dd = malloc(sizeof(double) * n);
    double *u = malloc(sizeof(double) * m);
    double *c0 = malloc(sizeof(double) * T), *s0 = malloc(sizeof(double) * T);
    double *c1 = malloc(sizeof(double) * T), *s1 = malloc(sizeof(double) * T);
    double *bp = malloc(sizeof(double) * (2 * (size_t)L * T + 2));
    double *bpt = malloc(sizeof(double) * (2 * (size_t)L + 2));
    storage_constraints(T, n, p->sto_pmax[s], p->sto_emax[s], Cm, b);
    for (int t = 0; t < T; ++t) { x[t] = c->s->D[(size_t)s * T + t]; x[T + t] = c->s->C[(size_t)s * T + t]; }
    int rc = 0, outer = 0;
    for (outer = 1; outer <= 100; ++outer) {
        memset(G, 0, sizeof(double) * n * n);
        memset(gv, 0, sizeof(double) * n);
        for (int t = 0; t < T; ++t) {
            double Db = c->s->D[(size_t)s * T + t], Cb = c->s->C[(size_t)s * T + t], dbar = Db - Cb;
            hmodel(c, nd, t, (x[t] - Db) - (x[T + t] - Cb), &c0[t], &s0[t]);
            double a = p->gamma + s0[t];
            G[(size_t)t * n + t] = p->prox_weight + a;
            G[(size_t)(T + t) * n + T + t] = p->prox_weight + a;
            G[(size_t)t * n + T + t] = G[(size_t)(T + t) * n + t] = -a;
            double net0 = c->pi[nd * T + t] + p->gamma * (c->Sbar[t] - dbar) + c0[t] - s0[t] * dbar;
            gv[t] = p->sto_mc[s] + net0 - p->prox_weight * Db;
            gv[T + t] = p->sto_mc[s] - net0 - p->prox_weight * Cb;
        }
        int it = qp_gi_solve(n, m, G, gv, Cm, b, xq, u);
        if (it < 0) { rc = it; break; }
        note_kkt(st, qp_kkt_residual(n, m, G, gv, Cm, b, xq, u));
        int same = 1;
        for (int t = 0; t < T; ++t) {
            double Db = c->s->D[(size_t)s * T + t], Cb = c->s->C[(size_t)s * T + t];
            hmodel(c, nd, t, (xq[t] - Db) - (xq[T + t] - Cb), &c1[t], &s1[t]);
            if (c1[t] != c0[t] || s1[t] != s0[t]) same = 0;
        }
        if (same) { memcpy(x, xq, sizeof(double) * n); break; }
        /* exact line search on the true (piecewise quadratic) objective along xq - x */
        double dmax = 0.0;
        for (int i = 0; i < n; ++i) { dd[i] = xq[i] - x[i]; if (fabs(dd[i]) > dmax) dmax = fabs(dd[i]); }
        if (dmax < 1e-13) break;
        int nb = 0;
        for (int t = 0; t < T; ++t) {
            double Db = c->s->D[(size_t)s * T + t], Cb = c->s->C[(size_t)s * T + t];
            double d0 = (x[t] - Db) - (x[T + t] - Cb), dv = dd[t] - dd[T + t];
            if (dv == 0.0) continue;
            int k = breakpoints(c, nd, t, bpt);
            for (int j = 0; j < k; ++j) {
                double al = (bpt[j] - d0) / dv;
                if (al > 0.0 && al < 1.0) bp[nb++] = al;
            }
        }
        ls_ud ud = { c, s, x, dd };
        double alpha = root_pl(ls_f, &ud, 0.0, 1.0, bp, nb);
        for (int i = 0; i < n; ++i) x[i] += alpha * dd[i];
        if (alpha * dmax < 1e-13) break;
    }
    if (outer > 100) rc = -4;
#ifdef _OPENMP
#pragma omp critical(dopf_outer)
#endif
    if (outer > st->storage_outer_max) st->storage_outer_max = outer;
    for (int t = 0; t < T; ++t) {
        double D = x[t], C = x[T + t];
        if (D < 0.0) D = 0.0;
        if (C < 0.0) C = 0.0;
        if (D > p->sto_pmax[s]) D = p->sto_pmax[s];
        if (C > p->sto_pmax[s]) C = p->sto_pmax[s];
        Dout[t] = D; Cout[t] = C;
    }
    free(G); free(gv); free(Cm); free(b); free(x); free(xq); free(dd); free(u);
    free(c0); free(s0); free(c1); free(s1); free(bp); free(bpt);
    return rc;
}

/* ---------- state ---------- */
static void matmul_ptdf(const oracle_problem *p, const double *inj, double *flow)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int l = 0; l < p->L; ++l) {
        double *out = flow + (size_t)l * p->T;
        for (int t = 0; t < p->T; ++t) out[t] = 0.0;
        for (int n = 0; n < p->N; ++n) {
            double a = p->ptdf[(size_t)l * p->N + n];
            if (a == 0.0) continue;
            const double *row = inj + (size_t)n * p->T;
            for (int t = 0; t < p->T; ++t) out[t] += a * row[t];
        }
    }
}

void oracle_init_state(const oracle_problem *p, oracle_state *s)
{
    const size_t T = p->T;
    s->iteration = 1; /* admm.jl:29 */
    s->converged = s->conv_lambda = s->conv_mue = s->conv_rho = 0;
    s->res_lambda = s->res_mue = s->res_rho = 0.0;
    s->total_costs = 0.0;
    s->qp_kkt_worst = 0.0;
    s->storage_outer_max = 0;
    memset(s->P, 0, sizeof(double) * p->G * T);
    memset(s->D, 0, sizeof(double) * p->S * T);
    memset(s->C, 0, sizeof(double) * p->S * T);
    memset(s->E, 0, sizeof(double) * p->S * T);
    /* helpers/results.jl:60-66 : previous node results are zero => injection = -demand */
    for (size_t i = 0; i < (size_t)p->N * T; ++i) s->inj[i] = -p->demand[i];
    matmul_ptdf(p, s->inj, s->flow);
    memset(s->avgU, 0, sizeof(double) * p->L * T);
    memset(s->avgK, 0, sizeof(double) * p->L * T);
    memset(s->lam, 0, sizeof(double) * T);       /* admm.jl:34-36 */
    memset(s->mu, 0, sizeof(double) * p->L * T);
    memset(s->rho, 0, sizeof(double) * p->L * T);
    memset(s->lam_prev, 0, sizeof(double) * T);
    memset(s->mu_prev, 0, sizeof(double) * p->L * T);
    memset(s->rho_prev, 0, sizeof(double) * p->L * T);
}

int oracle_iteration(const oracle_problem *p, oracle_state *s, int mode)
{
    const int N = p->N, L = p->L, T = p->T, G = p->G, S = p->S, A = G + S;
    ctx_t c;
    c.p = p; c.s = s;
    c.w2 = 2.0 * p->flow_weight;
    c.k = c.w2 + p->gamma;
    c.kappa = c.w2 * p->gamma / c.k;
    c.Sbar = malloc(sizeof(double) * T);
    c.pi = malloc(sizeof(double) * (size_t)N * T);
    c.ap = malloc(sizeof(double) * (size_t)L * T);
    c.am = malloc(sizeof(double) * (size_t)L * T);
    double *Pn = malloc(sizeof(double) * (size_t)(G > 0 ? G : 1) * T);
    double *Dn = malloc(sizeof(double) * (size_t)(S > 0 ? S : 1) * T);
    double *Cn = malloc(sizeof(double) * (size_t)(S > 0 ? S : 1) * T);
    double *sumU = calloc((size_t)L * T, sizeof(double)), *sumK = calloc((size_t)L * T, sizeof(double));
    int rc = 0;

    for (int t = 0; t < T; ++t) {
        double a = 0.0;
        for (int n = 0; n < N; ++n) a += s->inj[(size_t)n * T + t];
        c.Sbar[t] = a;
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int n = 0; n < N; ++n) {
        double *row = c.pi + (size_t)n * T;
        for (int t = 0; t < T; ++t) row[t] = s->lam[t];
        for (int l = 0; l < L; ++l) {
            const double a = p->ptdf[(size_t)l * N + n];
            if (a == 0.0) continue;
            const double *m = s->mu + (size_t)l * T, *r = s->rho + (size_t)l * T;
            for (int t = 0; t < T; ++t) row[t] += a * (m[t] - r[t]);
        }
    }
    for (int i = 0; i < L * T; ++i) {
        c.ap[i] = p->fmax[i / T] - s->flow[i];
        c.am[i] = p->fmax[i / T] + s->flow[i];
    }

    if (mode == 1) {
        /* literal: per-agent U,K come out of the QP and are summed as results.jl:83-84 does */
        double *U = malloc(sizeof(double) * (size_t)L * T), *K = malloc(sizeof(double) * (size_t)L * T);
        for (int g = 0; g < G && !rc; ++g) {
            rc = solve_generator_literal(&c, s, g, Pn + (size_t)g * T, U, K);
            for (int i = 0; i < L * T; ++i) { sumU[i] += U[i]; sumK[i] += K[i]; }
        }
        for (int q = 0; q < S && !rc; ++q) {
            rc = solve_storage_literal(&c, s, q, Dn + (size_t)q * T, Cn + (size_t)q * T, U, K);
            for (int i = 0; i < L * T; ++i) { sumU[i] += U[i]; sumK[i] += K[i]; }
        }
        free(U); free(K);
    } else {
#ifdef _OPENMP
#pragma omp parallel
#endif
        {
            double *bpbuf = malloc(sizeof(double) * (2 * (size_t)L + 2));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
            for (int g = 0; g < G; ++g) solve_generator_reduced(&c, g, Pn + (size_t)g * T, bpbuf);
            free(bpbuf);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
            for (int q = 0; q < S; ++q) {
                int r = solve_storage_reduced(&c, s, q, Dn + (size_t)q * T, Cn + (size_t)q * T);
                if (r) {
#ifdef _OPENMP
#pragma omp critical(dopf_rc)
#endif
                    rc = r;
                }
            }
        }
        /* per-agent slacks U*(delta), K*(delta) summed over all agents (results.jl:83-84) */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (int l = 0; l < L; ++l) {
            for (int g = 0; g < G; ++g) {
                double pl = p->ptdf[(size_t)l * N + p->gen_node[g]];
                for (int t = 0; t < T; ++t) {
                    double pd = pl * (Pn[(size_t)g * T + t] - s->P[(size_t)g * T + t]);
                    sumU[l * T + t] += ustar(&c, l, t, pd);
                    sumK[l * T + t] += kstar(&c, l, t, pd);
                }
            }
            for (int q = 0; q < S; ++q) {
                double pl = p->ptdf[(size_t)l * N + p->sto_node[q]];
                for (int t = 0; t < T; ++t) {
                    size_t i = (size_t)q * T + t;
                    double pd = pl * ((Dn[i] - s->D[i]) - (Cn[i] - s->C[i]));
                    sumU[l * T + t] += ustar(&c, l, t, pd);
                    sumK[l * T + t] += kstar(&c, l, t, pd);
                }
            }
        }
    }

    if (!rc) {
        /* Result(...) : results.jl:55-114 */
        memcpy(s->P, Pn, sizeof(double) * (size_t)G * T);
        memcpy(s->D, Dn, sizeof(double) * (size_t)S * T);
        memcpy(s->C, Cn, sizeof(double) * (size_t)S * T);
        for (size_t i = 0; i < (size_t)N * T; ++i) s->inj[i] = -p->demand[i];
        s->total_costs = 0.0;
        for (int g = 0; g < G; ++g) {
            double sum = 0.0;
            for (int t = 0; t < T; ++t) {
                double v = s->P[(size_t)g * T + t];
                s->inj[(size_t)p->gen_node[g] * T + t] += v;
                sum += v;
            }
            s->total_costs += sum * p->gen_mc[g];
        }
        for (int q = 0; q < S; ++q) {
            double sum = 0.0, e = 0.0;
            for (int t = 0; t < T; ++t) {
                size_t i = (size_t)q * T + t;
                s->inj[(size_t)p->sto_node[q] * T + t] += s->D[i] - s->C[i];
                sum += s->D[i] + s->C[i];
                e += s->C[i] - s->D[i];
                s->E[i] = e;
            }
            s->total_costs += sum * p->sto_mc[q];
        }
        for (int i = 0; i < L * T; ++i) { s->avgU[i] = sumU[i] / A; s->avgK[i] = sumK[i] / A; }
        matmul_ptdf(p, s->inj, s->flow);

        /* update_duals! : update_duals.jl:7-39 */
        memcpy(s->lam_prev, s->lam, sizeof(double) * T);
        memcpy(s->mu_prev, s->mu, sizeof(double) * (size_t)L * T);
        memcpy(s->rho_prev, s->rho, sizeof(double) * (size_t)L * T);
        for (int t = 0; t < T; ++t) {
            double a = 0.0;
            for (int n = 0; n < N; ++n) a += s->inj[(size_t)n * T + t];
            s->lam[t] = s->lam_prev[t] + p->gamma * a;
        }
        for (int i = 0; i < L * T; ++i) {
            double f = p->fmax[i / T];
            double mu = s->mu_prev[i] + p->gamma * (s->flow[i] + s->avgU[i] - f);
            double rho = s->rho_prev[i] + p->gamma * (s->avgK[i] - s->flow[i] - f);
            s->mu[i] = mu * (s->avgU[i] <= p->slack_mask_tol ? 1.0 : 0.0);
            s->rho[i] = rho * (s->avgK[i] <= p->slack_mask_tol ? 1.0 : 0.0);
        }

        /* check_convergence! : convergence.jl:1-31 */
        if (s->iteration != 1) {
            double rl = 0.0, rm = 0.0, rr = 0.0;
            for (int t = 0; t < T; ++t) { double d = fabs(s->lam[t] - s->lam_prev[t]); if (d > rl) rl = d; }
            for (int i = 0; i < L * T; ++i) {
                double d = fabs(s->mu[i] - s->mu_prev[i]); if (d > rm) rm = d;
                d = fabs(s->rho[i] - s->rho_prev[i]); if (d > rr) rr = d;
            }
            s->res_lambda = rl; s->res_mue = rm; s->res_rho = rr;
            s->conv_lambda = rl < p->eps; s->conv_mue = rm < p->eps; s->conv_rho = rr < p->eps;
            s->converged = s->conv_lambda && s->conv_mue && s->conv_rho;
        }
        if (!s->converged) s->iteration += 1;
    }
    free(c.Sbar); free(c.pi); free(c.ap); free(c.am);
    free(Pn); free(Dn); free(Cn); free(sumU); free(sumK);
    return rc;
}

/* ---------- PTDF (helpers/ptdf.jl:1-41) ---------- */
int oracle_ptdf(int N, int L, const int *from, const int *to, const double *b, int slack, double *out)
{
    const int M = N - 1;
    double *Bn = calloc((size_t)M * M, sizeof(double));
    double *inv = calloc((size_t)M * M, sizeof(double));
    int *map = malloc(sizeof(int) * N);
    for (int n = 0, j = 0; n < N; ++n) map[n] = (n == slack) ? -1 : j++;
    for (int l = 0; l < L; ++l) { /* Bn = A' B A restricted to non-slack */
        int f = map[from[l]], t = map[to[l]];
        if (f >= 0) Bn[(size_t)f * M + f] += b[l];
        if (t >= 0) Bn[(size_t)t * M + t] += b[l];
        if (f >= 0 && t >= 0) { Bn[(size_t)f * M + t] -= b[l]; Bn[(size_t)t * M + f] -= b[l]; }
    }
    for (int i = 0; i < M; ++i) inv[(size_t)i * M + i] = 1.0;
    int rc = 0;
    for (int col = 0; col < M && !rc; ++col) { /* Gauss-Jordan with partial pivoting */
        int piv = col;
        for (int r = col + 1; r < M; ++r)
            if (fabs(Bn[(size_t)r * M + col]) > fabs(Bn[(size_t)piv * M + col])) piv = r;
        if (fabs(Bn[(size_t)piv * M + col]) < 1e-300) { rc = -1; break; }
        if (piv != col)
            for (int k = 0; k < M; ++k) {
                double tmp = Bn[(size_t)piv * M + k]; Bn[(size_t)piv * M + k] = Bn[(size_t)col * M + k]; Bn[(size_t)col * M + k] = tmp;
                tmp = inv[(size_t)piv * M + k]; inv[(size_t)piv * M + k] = inv[(size_t)col * M + k]; inv[(size_t)col * M + k] = tmp;
            }
        double d = Bn[(size_t)col * M + col];
        for (int k = 0; k < M; ++k) { Bn[(size_t)col * M + k] /= d; inv[(size_t)col * M + k] /= d; }
        for (int r = 0; r < M; ++r) {
            if (r == col) continue;
            double f = Bn[(size_t)r * M + col];
            if (f == 0.0) continue;
            for (int k = 0; k < M; ++k) { Bn[(size_t)r * M + k] -= f * Bn[(size_t)col * M + k]; inv[(size_t)r * M + k] -= f * inv[(size_t)col * M + k]; }
        }
    }
    if (!rc)
        for (int l = 0; l < L; ++l) /* PTDF = (B A) * B_inv ; slack column stays 0 */
            for (int n = 0; n < N; ++n) {
                double v = 0.0;
                if (map[n] >= 0) {
                    int f = map[from[l]], t = map[to[l]];
                    if (f >= 0) v += b[l] * inv[(size_t)f * M + map[n]];
                    if (t >= 0) v -= b[l] * inv[(size_t)t * M + map[n]];
                }
                out[(size_t)l * N + n] = v;
            }
    free(Bn); free(inv); free(map);
    return rc;
}

void oracle_nodal_price(const oracle_problem *p, const double *lam, const double *mu,
                        const double *rho, double *out)
{
    for (int n = 0; n < p->N; ++n)
        for (int t = 0; t < p->T; ++t) {
            double a = lam[t];
            for (int l = 0; l < p->L; ++l) /* note mu + rho here (network_elements.jl:20-21) */
                a += (mu[l * p->T + t] + rho[l * p->T + t]) * p->ptdf[(size_t)l * p->N + n];
            out[(size_t)n * p->T + t] = a;
        }
}
