/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the shipped product path.
 *
 * Dense strictly-convex QP solver (Goldfarb & Idnani dual active-set method,
 * Math. Programming 27 (1983)) used by the CPU oracle in place of the
 * reference's third-party solver (Gurobi through JuMP, unpinned:
 * /root/reference/src/imports.jl:1-9, src/optimization/subproblems.jl:21,86,109,186).
 * Every reference subproblem is strictly convex, so any exact solver gives the
 * same minimiser; each solve is re-checked against the KKT conditions.
 *
 *     minimise   0.5 x'Gx + g'x     subject to   C x >= b      (m rows)
 */
#ifndef DOPF_QP_GI_H
#define DOPF_QP_GI_H

#ifdef __cplusplus
extern "C" {
#endif

/* G: n*n row-major symmetric positive definite (not modified)
 * g: n,  C: m*n row-major,  b: m
 * x: n (out),  u: m (out, multipliers >= 0, may be NULL)
 * returns number of iterations (>=0), -1 = infeasible, -2 = G not PD, -3 = iteration cap */
int qp_gi_solve(int n, int m, const double *G, const double *g,
                const double *C, const double *b, double *x, double *u);

/* max KKT violation of (x,u): stationarity, primal feas, dual feas, complementarity
 * (all scaled absolute). */
double qp_kkt_residual(int n, int m, const double *G, const double *g,
                       const double *C, const double *b, const double *x, const double *u);

#ifdef __cplusplus
}
#endif
#endif
