/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  See qp_gi.h.
 *
 * Goldfarb-Idnani dual active-set method for
 *     min 0.5 x'Gx + g'x   s.t.  C x >= b
 * written from the published algorithm (Goldfarb & Idnani 1983, Alg. in sec. 3):
 *   G = L L',  J = L^-T,  J' N = [R; 0]  for the matrix N of active normals.
 */
#include "qp_gi.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define JX(i, j) J[(size_t)(i) * n + (j)]
#define RX(i, j) R[(size_t)(i) * n + (j)]

static int cholesky(int n, const double *G, double *Lc)
{
    memcpy(Lc, G, sizeof(double) * (size_t)n * n);
    for (int j = 0; j < n; ++j) {
        double s = Lc[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) s -= Lc[(size_t)j * n + k] * Lc[(size_t)j * n + k];
        if (!(s > 0.0)) return -1;
        double d = sqrt(s);
        Lc[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double t = Lc[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) t -= Lc[(size_t)i * n + k] * Lc[(size_t)j * n + k];
            Lc[(size_t)i * n + j] = t / d;
        }
        for (int k = j + 1; k < n; ++k) Lc[(size_t)j * n + k] = 0.0;
    }
    return 0;
}

int qp_gi_solve(int n, int m, const double *G, const double *g,
                const double *C, const double *b, double *x, double *u_out)
{
    double *Lc = malloc(sizeof(double) * (size_t)n * n);
    double *J = calloc((size_t)n * n, sizeof(double));
    double *R = calloc((size_t)n * n, sizeof(double));
    double *d = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n);
    double *r = malloc(sizeof(double) * n), *u = calloc((size_t)n + 1, sizeof(double));
    int *act = malloc(sizeof(int) * (n + 1));
    char *is_act = calloc((size_t)m + 1, 1), *excl = calloc((size_t)m + 1, 1);
    int q = 0, iter = 0, ret = 0;

    if (cholesky(n, G, Lc)) { ret = -2; goto done; }

    /* J = L^-T : solve L' J = I column by column (J upper triangular) */
    for (int c = 0; c < n; ++c) {
        for (int i = n - 1; i >= 0; --i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k < n; ++k) s -= Lc[(size_t)k * n + i] * JX(k, c);
            JX(i, c) = s / Lc[(size_t)i * n + i];
        }
    }
    /* x = -G^-1 g */
    for (int i = 0; i < n; ++i) {
        double s = -g[i];
        for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * n + k] * x[k];
        x[i] = s / Lc[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= Lc[(size_t)k * n + i] * x[k];
        x[i] = s / Lc[(size_t)i * n + i];
    }

    double cscale = 1.0;
    for (int i = 0; i < m; ++i) {
        double s = fabs(b[i]);
        if (s > cscale) cscale = s;
    }
    const double ftol = 1e-13 * cscale;
    const int itmax = 50 * (n + m) + 100;

    for (;;) {
        /* step 1: most violated inactive constraint */
        int ip = -1;
        double sip = -ftol;
        for (int i = 0; i < m; ++i) {
            if (is_act[i] || excl[i]) continue;
            const double *ci = C + (size_t)i * n;
            double s = -b[i];
            for (int k = 0; k < n; ++k) s += ci[k] * x[k];
            if (s < sip) { sip = s; ip = i; }
        }
        if (ip < 0) break;
        const double *np = C + (size_t)ip * n;
        u[q] = 0.0;

        for (;;) { /* step 2 */
            if (++iter > itmax) { ret = -3; goto done; }
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += JX(k, j) * np[k];
                d[j] = s;
            }
            double znorm = 0.0;
            for (int k = 0; k < n; ++k) {
                double s = 0.0;
                for (int j = q; j < n; ++j) s += JX(k, j) * d[j];
                z[k] = s;
                znorm += s * s;
            }
            for (int i = q - 1; i >= 0; --i) {
                double s = d[i];
                for (int k = i + 1; k < q; ++k) s -= RX(i, k) * r[k];
                r[i] = s / RX(i, i);
            }
            double t1 = INFINITY, t2 = INFINITY;
            int l = -1;
            for (int k = 0; k < q; ++k)
                if (r[k] > 0.0 && u[k] / r[k] < t1) { t1 = u[k] / r[k]; l = k; }
            double znp = 0.0;
            for (int k = 0; k < n; ++k) znp += z[k] * np[k];
            if (znorm > 1e-28 && znp > 1e-300) t2 = -sip / znp;
            double t = t1 < t2 ? t1 : t2;
            if (!isfinite(t)) { ret = -1; goto done; }

            int full = (t2 <= t1);
            if (isfinite(t2)) {
                for (int k = 0; k < n; ++k) x[k] += t * z[k];
            }
            for (int k = 0; k < q; ++k) u[k] -= t * r[k];
            u[q] += t;

            if (isfinite(t2) && full) {
                /* add ip: rotate d[q+1..n-1] into d[q] */
                for (int j = n - 1; j > q; --j) {
                    double a = d[j - 1], bb = d[j];
                    if (bb == 0.0) continue;
                    double h = hypot(a, bb), cs = a / h, sn = bb / h;
                    d[j - 1] = h; d[j] = 0.0;
                    for (int k = 0; k < n; ++k) {
                        double j0 = JX(k, j - 1), j1 = JX(k, j);
                        JX(k, j - 1) = cs * j0 + sn * j1;
                        JX(k, j) = -sn * j0 + cs * j1;
                    }
                }
                if (q >= n || fabs(d[q]) < 1e-14 * sqrt(znorm + 1.0)) {
                    /* dependent on the active set although a primal step existed: numerically
                     * impossible for znorm>0, but be safe */
                    excl[ip] = 1;
                    break;
                }
                for (int i = 0; i <= q; ++i) RX(i, q) = d[i];
                act[q] = ip; is_act[ip] = 1; ++q;
                memset(excl, 0, (size_t)m);
                break; /* back to step 1 */
            }
            /* drop active constraint at position l */
            {
                is_act[act[l]] = 0;
                for (int j = l; j < q - 1; ++j) {
                    act[j] = act[j + 1];
                    u[j] = u[j + 1];
                    for (int i = 0; i <= j + 1; ++i) RX(i, j) = RX(i, j + 1);
                }
                u[q - 1] = u[q]; u[q] = 0.0;
                for (int i = 0; i < q; ++i) RX(i, q - 1) = 0.0;
                --q;
                for (int j = l; j < q; ++j) {
                    double a = RX(j, j), bb = RX(j + 1, j);
                    if (bb == 0.0) continue;
                    double h = hypot(a, bb), cs = a / h, sn = bb / h;
                    for (int k = j; k < q; ++k) {
                        double r0 = RX(j, k), r1 = RX(j + 1, k);
                        RX(j, k) = cs * r0 + sn * r1;
                        RX(j + 1, k) = -sn * r0 + cs * r1;
                    }
                    for (int k = 0; k < n; ++k) {
                        double j0 = JX(k, j), j1 = JX(k, j + 1);
                        JX(k, j) = cs * j0 + sn * j1;
                        JX(k, j + 1) = -sn * j0 + cs * j1;
                    }
                }
            }
            if (isfinite(t2)) { /* partial step: recompute slack of ip */
                sip = -b[ip];
                for (int k = 0; k < n; ++k) sip += np[k] * x[k];
                if (sip >= -ftol * 1e-3) {
                    /* became feasible through rounding; treat as satisfied */
                    u[q] = 0.0;
                    break;
                }
            }
        }
    }
    if (u_out) {
        memset(u_out, 0, sizeof(double) * (size_t)m);
        for (int k = 0; k < q; ++k) u_out[act[k]] = u[k];
    }
    ret = iter;
done:
    free(Lc); free(J); free(R); free(d); free(z); free(r); free(u); free(act); free(is_act); free(excl);
    return ret;
}

double qp_kkt_residual(int n, int m, const double *G, const double *g,
                       const double *C, const double *b, const double *x, const double *u)
{
    double worst = 0.0;
    double *grad = malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        double s = g[i];
        for (int k = 0; k < n; ++k) s += G[(size_t)i * n + k] * x[k];
        grad[i] = s;
    }
    for (int j = 0; j < m; ++j) {
        const double *cj = C + (size_t)j * n;
        double s = -b[j];
        for (int k = 0; k < n; ++k) s += cj[k] * x[k];
        if (-s > worst) worst = -s;                       /* primal feasibility */
        if (-u[j] > worst) worst = -u[j];                 /* dual feasibility */
        double comp = fabs(u[j] * s);
        if (comp > worst) worst = comp;                   /* complementarity */
        for (int k = 0; k < n; ++k) grad[k] -= u[j] * cj[k];
    }
    for (int i = 0; i < n; ++i)
        if (fabs(grad[i]) > worst) worst = fabs(grad[i]); /* stationarity */
    free(grad);
    return worst;
}
