/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the shipped product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement (plain C, fp64) of one ADMM iteration of DecentralOPF.jl:
 *   calculate_iteration!            /root/reference/src/optimization/run.jl:7-16
 *   optimize_subproblem(::Generator) src/optimization/subproblems.jl:19-105
 *   optimize_subproblem(::Storage)   src/optimization/subproblems.jl:107-207
 *   add_penalty_terms!              src/optimization/penalty_terms.jl:1-53
 *   Result(unit_to_result)          src/structures/results.jl:50-117
 *   update_duals!                   src/optimization/update_duals.jl:1-39
 *   check_convergence!              src/optimization/convergence.jl:1-31
 *   calculate_ptdf                  src/helpers/ptdf.jl:1-41
 *   get_nodal_price                 src/helpers/network_elements.jl:16-25
 *
 * Parity pinning: the reference has no tests; the oracle is pinned against the reference's
 * committed per-iteration traces results/{TNS,big_gamma,wrong_weight}_*.csv (copied as
 * tests/golden/ (npz) by tests/golden/make_golden.py) - see tests/test_oracle_golden.py.
 */
#ifndef DOPF_ORACLE_H
#define DOPF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int N, L, T, G, S;
    const double *ptdf;     /* [L][N] row-major                      admm.ptdf        */
    const double *fmax;     /* [L]                                   admm.f_max       */
    const double *demand;   /* [N][T]                                node.demand      */
    const double *gen_mc;   /* [G] marginal_costs                                     */
    const double *gen_pmax; /* [G] max_generation                                     */
    const int *gen_node;    /* [G] 0-based node index                                 */
    const double *sto_mc;   /* [S] marginal_costs                                     */
    const double *sto_pmax; /* [S] max_power                                          */
    const double *sto_emax; /* [S] max_level                                          */
    const int *sto_node;    /* [S]                                                    */
    double gamma;           /* admm.gamma                                             */
    double flow_weight;     /* the literal 10 (subproblems.jl:77-78,176-177)          */
    double prox_weight;     /* 1.0 <=> the literal 1/2*(x-prev)^2 (subproblems.jl:81) */
    double slack_mask_tol;  /* 1e-2 (update_duals.jl:24,36)                           */
    double eps;             /* 1e-3 (convergence.jl:2)                                */
} oracle_problem;

typedef struct {
    int iteration;     /* admm.iteration (1-based)                                    */
    int converged;     /* admm.convergence.all                                        */
    int conv_lambda, conv_mue, conv_rho;
    double res_lambda, res_mue, res_rho; /* max |dual_{k+1}-dual_k| of the last check  */
    double total_costs;
    double qp_kkt_worst; /* worst KKT residual of any generic QP solve so far          */
    int storage_outer_max; /* max outer (slack-pattern) iterations of a storage solve  */
    double *P;         /* [G][T]  results[k] generation                               */
    double *D, *C, *E; /* [S][T]  discharge, charge, level                            */
    double *inj;       /* [N][T]  results[k].injection                                */
    double *flow;      /* [L][T]  results[k].line_utilization                         */
    double *avgU, *avgK; /* [L][T]                                                    */
    double *lam, *mu, *rho;                /* newest duals (lambdas[end])             */
    double *lam_prev, *mu_prev, *rho_prev; /* duals used by the last iteration        */
} oracle_state;

/* state <- the reference's state before iteration 1 (admm.jl:29-36, helpers/results.jl zeros) */
void oracle_init_state(const oracle_problem *p, oracle_state *s);

/* one calculate_iteration!.  mode 0: reduced exact form,
 * mode 1: literal formulation with explicit U,K slack variables solved as one dense QP per
 * agent (tiny cases only).  returns 0 ok, <0 solver failure. */
int oracle_iteration(const oracle_problem *p, oracle_state *s, int mode);

/* PTDF from incidence/susceptance (ptdf.jl).  from/to 0-based; slack node index. out [L][N] */
int oracle_ptdf(int N, int L, const int *from, const int *to, const double *susceptance,
                int slack, double *out);

/* nodal price (network_elements.jl:16-25) out [N][T] from (lam, mu, rho) */
void oracle_nodal_price(const oracle_problem *p, const double *lam, const double *mu,
                        const double *rho, double *out);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
