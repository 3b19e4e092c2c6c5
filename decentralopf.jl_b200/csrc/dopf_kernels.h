// Launch plan + entry points of dopf_kernels.cu (internal to the library).
#ifndef DOPF_KERNELS_H
#define DOPF_KERNELS_H
#include <cuda_runtime.h>
#include "dopf_bodies.h"

namespace dopf {

constexpr int COLSUM_R = 8;     // row groups of the column sums of the injection (k_colsum)

struct LaunchPlan {
    View view;
    int num_sms;
    int bm_t, ksplit_t;       // tile rows / split-K of the transposed product (PTDF^T M)
    int mt_base, mt_rows;     // node rows [mt_base, mt_base + mt_rows) the transposed product is computed for (multiples of 64)
    int bm_n, ksplit_n;       // ... of the flow product (PTDF * inj)
    int bm_x, ksplit_x;       // ... of the partial flow product of the partitioned mode (K = the rank's node range)
    int sto_fix_blocks;       // grid of the storage correction pass (one 32-thread block = one affected storage)
    int sto_fix_slots;        // nodes whose hinge lists k_sto_collect gathers (one scratch slot each, shared by the node's storages)
    int gen_flat;             // 1: agent-major generator kernel (few generators per node)
    int sto_j;                // timesteps per lane of the warp-parallel storage solve (0: horizon too long)
    int slack_blocks_x;
    double *part, *part2;     // split-K partial tiles
    unsigned char *tflag;     // [Lp][ldt] bit0/bit1: exact row sums present (U/K side)
    Hinge *hinge_scratch;     // [sto_fix_slots + sto_fix_blocks][T][hcap]
    int *hcnt_scratch;        // [sto_fix_slots + sto_fix_blocks][T]
    // fork/join of independent kernel groups on a second stream (whole-iteration mode only)
    cudaStream_t side_stream; cudaEvent_t ev_fork, ev_join;
    // optional per-kernel profiling (dopf_profile_iteration): event pairs + names in launch order
    cudaEvent_t *prof_events; const char **prof_names; int prof_cap; int *prof_count;
};

// exchange points split one iteration into 4 segments (multi-GPU: the host all-reduces the buffer named
// by the exchange point between two segments).  segment = -1 enqueues the whole iteration.
enum { DOPF_X_DMAX = 0, DOPF_X_INJ = 1, DOPF_X_ROWSUM = 2, DOPF_N_SEGMENTS = 4 };
int enqueue_iteration(const LaunchPlan &lp, cudaStream_t st, int segment = -1);
void launch_mwide(const View &v, cudaStream_t st);
int set_storage_smem_attr(int T);   // opt in to > 48 KB dynamic shared memory for long horizons
int slack_rows_cap();   // returns number of kernel launches
void launch_total_costs(const View &v, double *d_out, cudaStream_t st);
void launch_profile_warm(const LaunchPlan &lp, cudaStream_t st);
void launch_nodal_price(const View &v, const double *lam, const double *mu, const double *rho, double *d_out, cudaStream_t st);
void launch_unit_penalty(const View &v, int kind, int idx, double *eb, double *up, double *lo, double *U, double *K, cudaStream_t st);
void launch_copy_inj(const View &v, const double *src, cudaStream_t st);   // inj = src - demand
void launch_flow_of_demand(const LaunchPlan &lp, double *dst, cudaStream_t st);
void launch_pack_cols(double *dev, double *host_layout, int rows, int C, int T, int ld, int to_device, cudaStream_t st);
void launch_penalty_totals(const View &v, double *eb, double *up, double *lo, cudaStream_t st);
// segment 0: local injection of the staged iterate; segment 1: column sums, flows, levels, buffer flip
void launch_rebuild_derived(const LaunchPlan &lp, cudaStream_t st, int segment = -1);  // inj/ssum/flow/E of buffer [cur] from P,D,C

}  // namespace dopf
#endif
