/*
 * libdopf - C ABI of the B200-native ADMM iteration for DecentralOPF.jl.
 *
 * The reference has no FFI seam; the seam cut here is its Julia call surface for the hot path
 * (all citations relative to /root/reference/):
 *
 *   dopf_create        <- ADMM(gamma, nodes, generators, storages, lines)   src/structures/admm.jl:23-62
 *   dopf_step          <- run!(admm) / calculate_iteration!(admm)           src/optimization/run.jl:1-16
 *                         = optimize_all_subproblems!  src/optimization/subproblems.jl:1-17
 *                         + update_duals!              src/optimization/update_duals.jl:1-39
 *                         + check_convergence!         src/optimization/convergence.jl:1-31
 *   dopf_get_iterate   <- admm.results[end].{unit_to_result[u].generation|discharge|charge|level,
 *                         injection, line_utilization, avg_U, avg_K}        src/structures/results.jl:36-48
 *   dopf_get_duals     <- admm.lambdas[k], admm.mues[k], admm.rhos[k]       src/structures/admm.jl:4-6
 *   dopf_get_status    <- admm.iteration, admm.convergence.{lambda,mue,rho,all} src/structures/convergence.jl
 *   dopf_set_state     <- (no reference counterpart) resume / inject a mid-trace state
 *   dopf_get_nodal_price <- get_nodal_price(iteration)                      src/helpers/network_elements.jl:16-25
 *   dopf_get_total_costs <- result.total_costs                              src/structures/results.jl:95-105
 *   dopf_calculate_ptdf <- calculate_ptdf(nodes, lines)                     src/helpers/ptdf.jl:1-41
 *   dopf_get_unit_penalty <- unit_to_result[u].{penalty_term,U,K}           src/optimization/subproblems.jl:89-102
 *   dopf_get_penalty_totals <- result.penalty_term                          src/structures/results.jl:73-76
 *
 * Conventions: every matrix is row-major with the timestep index contiguous ([agent][t],
 * [node][t], [line][t]; ptdf is [line][node]).  Julia callers pass permutedims(...) of their
 * column-major matrices.  Node indices are 0-based.  The library copies
 * all inputs; callers keep ownership of their buffers.  A handle is not thread-safe.  No
 * function throws or aborts: 0 = ok, < 0 = error, text from dopf_last_error().
 * There is no CPU fallback: without a CUDA device dopf_create fails with DOPF_E_CUDA.
 */
#ifndef DOPF_H
#define DOPF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOPF_OK 0
#define DOPF_E_ARG (-1)       /* invalid argument / dimension                         */
#define DOPF_E_CUDA (-2)      /* CUDA runtime error (incl. "no device")               */
#define DOPF_E_CAPACITY (-3)  /* a device work list / hinge list capacity was exceeded */
#define DOPF_E_COMM (-4)      /* NCCL could not be loaded / a collective of dopf_comm_init or dopf_step failed */
#define DOPF_E_UNSUPPORTED (-5)

typedef struct dopf_handle dopf_handle;

typedef struct dopf_problem {
    int32_t N, L, T, G, S;      /* nodes, lines, timesteps, generators, storages          */
    const double *ptdf;         /* [L][N]  admm.ptdf (helpers/ptdf.jl)                    */
    const double *f_max;        /* [L]     line.max_capacity (admm.jl:41)                 */
    const double *demand;       /* [N][T]  node.demand                                    */
    const double *gen_mc;       /* [G]     generator.marginal_costs                       */
    const double *gen_pmax;     /* [G]     generator.max_generation                       */
    const int32_t *gen_node;    /* [G]     0-based index of generator.node                */
    const double *sto_mc;       /* [S]     storage.marginal_costs                         */
    const double *sto_pmax;     /* [S]     storage.max_power                              */
    const double *sto_emax;     /* [S]     storage.max_level                              */
    const int32_t *sto_node;    /* [S]                                                    */
} dopf_problem;

typedef struct dopf_config {
    double gamma;            /* admm.gamma (opf_admm_decentral.jl:5 uses 0.3)               */
    double flow_weight;      /* the literal 10  (subproblems.jl:77-78,176-177)             */
    double prox_weight;      /* 1.0 <=> the literal 1/2*(x-prev)^2 (subproblems.jl:81,180) */
    double slack_mask_tol;   /* 1e-2 (update_duals.jl:24,36)                               */
    double eps;              /* 10^-3 (convergence.jl:2)                                   */
    int32_t device;          /* CUDA device ordinal, -1 = current device                   */
    int32_t hinge_capacity;  /* per (agent,t) hinge list capacity of the correction pass; 0 = default 32 */
    int32_t use_graph;       /* 1 = replay one captured CUDA graph per iteration           */
    int32_t debug_flags;     /* 0; diagnostics: bit0 storage correction pass by the sequential solver, bit1 storage predict pass too */
    int32_t n_scenarios;     /* 1; > 1: a batch of independent scenarios on one grid (BASELINE configs[3]), see below */
    int32_t gemm_ksplit;     /* 0 = chosen by the launch plan; > 0 fixes the split-K (summation order) of the PTDF products */
} dopf_config;

/* fills the reference's literals: gamma 0.3, flow_weight 10, prox 1, mask 1e-2, eps 1e-3 */
void dopf_default_config(dopf_config *c);

/* Scenario batches (dopf_config.n_scenarios = C > 1): C independent problems on the same grid (ptdf, f_max and the
 * node of every agent are shared; N, L, T, G, S in dopf_problem are per scenario) run as ONE device problem with C*T
 * columns - every kernel launch, both PTDF products included, covers all scenarios.  Arrays gain a leading scenario
 * dimension: demand [C][N][T], gen_mc/gen_pmax [C][G], sto_mc/sto_pmax/sto_emax [C][S]; dopf_get_iterate returns
 * P [C][G][T], D,C,E [C][S][T], injection [C][N][T], flow/avgU/avgK [C][L][T]; dopf_get_duals lam [C][T], mu/rho
 * [C][L][T]; dopf_get_nodal_price [C][N][T]; dopf_get_total_costs [C].  Every scenario follows the reference's stop
 * rule on its own (convergence.jl:1-31): a converged scenario is frozen while the others continue, exactly as if each
 * had been run alone; dopf_step returns when all have converged.  dopf_status reports scenario 0 (iteration, flags,
 * residuals) and converged = all; dopf_get_scenario_status gives every scenario. */
typedef struct dopf_status {
    int32_t iteration;       /* admm.iteration (1-based; not advanced by the converging iteration) */
    int32_t converged;       /* admm.convergence.all                                        */
    int32_t conv_lambda, conv_mue, conv_rho;
    int32_t iterations_done; /* iterations executed since create / set_state                */
    double res_lambda, res_mue, res_rho;   /* max |dual_{k+1} - dual_k| of the last iteration */
    /* statistics of the exact-correction pass */
    int32_t gen_corrected, sto_corrected;  /* cumulated agents re-solved with explicit hinges */
    int32_t tight_rows, wide_rows;         /* candidate (line,t,side) rows of the last iteration */
    int32_t launches_per_iteration;        /* kernels enqueued per iteration                  */
    int32_t sto_cold;                      /* storages that needed the cold solve in the last iteration */
    double last_step_ms;                   /* device time of the last dopf_step (CUDA events on the library stream) */
    int32_t fix_sequential;                /* cumulated correction-pass storages that fell back to the sequential solver */
    int32_t reserved3;
} dopf_status;

int dopf_create(const dopf_problem *p, const dopf_config *c, dopf_handle **out);
void dopf_destroy(dopf_handle *h);

/* runs up to max_iters iterations, stops early when converged; synchronises before returning */
int dopf_step(dopf_handle *h, int32_t max_iters, dopf_status *out);
int dopf_get_status(dopf_handle *h, dopf_status *out);
/* per scenario: admm.iteration [C], admm.convergence.all [C], residuals of the last check [C][3]; any pointer may be NULL */
int dopf_get_scenario_status(dopf_handle *h, int32_t *iteration, int32_t *converged, double *residuals);

/* newest iterate; any pointer may be NULL (skipped).  Agent order = the order given at create. */
int dopf_get_iterate(dopf_handle *h, double *P /*[G][T]*/, double *D, double *C, double *E /*[S][T]*/,
                     double *injection /*[N][T]*/, double *flow /*[L][T]*/,
                     double *avgU /*[L][T]*/, double *avgK /*[L][T]*/);
/* which = 0: newest duals (admm.lambdas[end]); 1: the duals used by the last iteration */
int dopf_get_duals(dopf_handle *h, int32_t which, double *lam /*[T]*/, double *mu /*[L][T]*/, double *rho /*[L][T]*/);

/* overwrite the state: iteration counter, previous iterate P,D,C, average slacks and duals.
 * injection, flows and levels are re-derived on the device.  NULL = keep. */
int dopf_set_state(dopf_handle *h, int32_t iteration, const double *P, const double *D, const double *C,
                   const double *avgU, const double *avgK, const double *lam, const double *mu, const double *rho);

/* runs ONE iteration kernel by kernel with a CUDA event pair around each launch (no graph) and
 * returns the device time of every kernel in launch order.  names[i] points to static strings. */
int dopf_profile_iteration(dopf_handle *h, int32_t cap, float *ms, const char **names, int32_t *count);

int dopf_get_nodal_price(dopf_handle *h, int32_t which, double *out /*[N][T]*/);
/* get_nodal_price(k) for any dual set of the caller's history (admm.lambdas[k], admm.mues[k], admm.rhos[k]) */
int dopf_nodal_price_from(dopf_handle *h, const double *lam /*[T]*/, const double *mu /*[L][T]*/, const double *rho /*[L][T]*/,
                          double *out /*[N][T]*/);
int dopf_get_total_costs(dopf_handle *h, double *out);
/* per-unit report of the newest iterate (ResultGenerator / ResultStorage fields penalty_term, U, K;
 * src/optimization/subproblems.jl:89-102, src/structures/results.jl:1-17): kind 0 = generator, 1 = storage, index in the
 * caller's order; energy_balance/upper_flow/lower_flow are [T]; U, K are [L][T] and may be NULL. */
int dopf_get_unit_penalty(dopf_handle *h, int32_t kind, int32_t index, double *energy_balance, double *upper_flow,
                          double *lower_flow, double *U, double *K);
/* result.penalty_term = sum of the units' penalty terms (src/structures/results.jl:73-76), each [T] */
int dopf_get_penalty_totals(dopf_handle *h, double *energy_balance, double *upper_flow, double *lower_flow);

/* ---- multi-GPU (one process per GPU; SURVEY.md 8(e) "agent block") ---------------------------------
 * Every rank creates its handle with the full network data and ITS block of the agents, then calls
 * dopf_set_partition.  One iteration is run as 4 phases; between two phases the caller all-reduces the
 * named device buffer IN PLACE over the ranks (torch.distributed / NCCL on the stream given to
 * dopf_set_stream, so no host synchronisation is needed):
 *     phase 0 -> all-reduce MAX  DOPF_XBUF_DMAX   (largest agent move per timestep)
 *     phase 1 -> all-reduce SUM  DOPF_XBUF_INJ    (nodal injection of the agents; the library subtracts the demand afterwards)
 *     phase 2 -> all-reduce SUM  DOPF_XBUF_ROWSUM (exact slack row sums + partial line flows: every rank multiplies PTDF with
 *                                                  the injection of its own agents over its own node range only, so both PTDF
 *                                                  products cost each rank 1/nranks of the single-GPU work)
 *     phase 3    (dual update, convergence check, buffer flip)
 * Set-up: dopf_set_partition, all-reduce SUM DOPF_XBUF_INJ and MAX DOPF_XBUF_RBOX, dopf_step_phase(h, -1).
 * The elementwise network / dual part is replicated, so all ranks hold identical duals and convergence flags. */
#define DOPF_XBUF_DMAX 0
#define DOPF_XBUF_INJ 1
#define DOPF_XBUF_ROWSUM 2
#define DOPF_XBUF_RBOX 3     /* set-up only: per-node box range, all-reduce MAX before dopf_step_phase(h, -1) */
int dopf_set_partition(dopf_handle *h, int32_t rank, int32_t nranks, int32_t total_agents);
/* CUDA stream (cudaStream_t) on which the library enqueues all work; NULL = the library's own stream */
int dopf_set_stream(dopf_handle *h, void *cuda_stream);
/* enqueue one phase (0..3) of the current iteration; asynchronous.  After phase 3 call dopf_get_status
 * (synchronises) whenever the host needs the iteration counter / convergence flags. */
int dopf_step_phase(dopf_handle *h, int32_t phase);
/* device pointer and element count (float64, or the bit pattern of non-negative float64 for DMAX) of
 * the buffer to all-reduce after `phase` = which */
int dopf_exchange_buffer(dopf_handle *h, int32_t which, void **device_ptr, int64_t *count);

/* The same partition with the collectives INSIDE the library (SURVEY.md 8(b) "n_gpus"): the host - Julia included - only
 * moves DOPF_COMM_ID_BYTES opaque bytes from rank 0 to the other ranks (any transport: MPI, a file, a socket).
 *     rank 0:    dopf_comm_get_unique_id(id)               (ncclGetUniqueId; libnccl.so.2 is loaded at run time)
 *     all ranks: dopf_create(... the rank's block of the agents ...); dopf_comm_init(h, id, rank, nranks, total_agents)
 *     all ranks: dopf_step(h, n, &status)   - same call as on one GPU: the four phases and the three ncclAllReduce of every
 *                                             iteration run on the library's stream, captured in ONE CUDA graph (use_graph).
 * All ranks must call dopf_step with the same max_iters.  A device-side capacity error of any rank stops every rank. */
#define DOPF_COMM_ID_BYTES 128
int dopf_comm_get_unique_id(void *id_out /*[DOPF_COMM_ID_BYTES]*/);
int dopf_comm_init(dopf_handle *h, const void *id /*[DOPF_COMM_ID_BYTES]*/, int32_t rank, int32_t nranks, int32_t total_agents);

/* diagnostics: 32 device-side event counters (all zero unless the library was built with -DDOPF_STATS) */
int dopf_debug_counters(dopf_handle *h, uint64_t *out /*[32]*/, int32_t reset);

/* calculate_ptdf(nodes, lines) on the GPU (src/helpers/ptdf.jl:1-41, called by ADMM(...) src/structures/admm.jl:44):
 * line_from/line_to are 0-based node indices (incidence +1 / -1), slack = index of the first node with slack == true,
 * out is [L][N] row-major with a zero slack column.  Cholesky + triangular solves of the slack-reduced susceptance matrix
 * (cuSOLVER, loaded at run time) instead of the reference's dense inverse.  Error text: dopf_ptdf_last_error(). */
int dopf_calculate_ptdf(int32_t N, int32_t L, const int32_t *line_from, const int32_t *line_to, const double *susceptance,
                        int32_t slack, int32_t device, double *out);
const char *dopf_ptdf_last_error(void);

const char *dopf_last_error(dopf_handle *h);   /* handle may be NULL: error of the last failed dopf_create */
const char *dopf_version(void);

#ifdef __cplusplus
}
#endif
#endif
