// Launch plan + entry points of dopf_kernels.cu (internal to the library).
#ifndef DOPF_KERNELS_H
#define DOPF_KERNELS_H
#include <cuda_runtime.h>
#include "dopf_bodies.h"

namespace dopf {

struct LaunchPlan {
    View view;
    int num_sms;
    int bm_t, ksplit_t;       // tile rows / split-K of the transposed product (PTDF^T M)
    int bm_n, ksplit_n;       // ... of the flow product (PTDF * inj)
    int sto_warps;            // warps (storages) per block of the storage kernels
    int sto_blocks;           // grid cap of the predict pass
    int sto_fix_blocks;       // grid of the correction pass (sizes hinge_scratch)
    int slack_blocks_x;
    double *part, *part2;     // split-K partial tiles
    unsigned char *tflag;     // [Lp][ldt] bit0/bit1: exact row sums present (U/K side)
    Hinge *hinge_scratch;     // [sto_fix_blocks*sto_warps][T][hcap]
};

int enqueue_iteration(const LaunchPlan &lp, cudaStream_t st);   // returns number of kernel launches
size_t storage_smem_bytes(int T, int warps);
int set_storage_smem_attr(size_t bytes);
void launch_total_costs(const View &v, double *d_out, cudaStream_t st);
void launch_nodal_price(const View &v, int which, double *d_out, cudaStream_t st);
void launch_rebuild_derived(const LaunchPlan &lp, cudaStream_t st);  // inj/ssum/flow/E of buffer [cur] from P,D,C

}  // namespace dopf
#endif
