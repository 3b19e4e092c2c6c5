// C ABI of libdopf (include/dopf.h): instance set-up, SoA packing, CUDA-graph driven stepping,
// state transfer, phase-wise stepping for the partitioned (multi-GPU) mode.  No CPU compute path exists in this library.
#include "../../include/dopf.h"
#include "dopf_kernels.h"
#include <nccl.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

using namespace dopf;

namespace {

thread_local std::string g_create_error;

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace

struct dopf_handle {
    LaunchPlan lp{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool use_graph = true;
    int launches_per_iter = 0;
    std::vector<void *> allocs;
    Ctrl *h_ctrl = nullptr;          // pinned mirror of the device control block
    double *d_scalar = nullptr;
    double *d_nodal = nullptr;
    double *d_stage = nullptr;             // scenario batches: [C][rows][T] staging of one network matrix
    std::vector<int32_t> h_sc_i; std::vector<double> h_sc_d;
    double *d_pen = nullptr;               // lazily allocated: 3*T penalty values + U,K [L][T] of one unit
    std::vector<int> gen_perm, sto_perm;   // sorted position -> caller's index
    bool gen_identity = true, sto_identity = true;
    std::vector<double> stage;             // host staging for permutation / padding
    std::string err;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_step_ms = 0.0;
    // multi-GPU (agent-block partition; exchanges are done by the caller between the phases)
    int rank = 0, nranks = 1;
    int host_cur = 0;          // host mirror of Ctrl::cur, advanced by phase 3
    bool partitioned = false;  // dopf_set_partition was called: stepped phase by phase
    // library-owned communicator (dopf_comm_init): dopf_step then runs the phases AND the all-reduces itself
    ncclComm_t comm = nullptr;
    bool cfg_use_graph = true;
    int *d_flag = nullptr;     // [1] error agreement between the ranks
    cudaStream_t own_stream = nullptr;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            h->err = buf_;                                                                         \
            return DOPF_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

template <class T> int dev_alloc(dopf_handle *h, T **p, size_t count, bool zero = true)
{
    void *q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc(&q, bytes));
    h->allocs.push_back(q);
    if (zero) CK(cudaMemsetAsync(q, 0, bytes, h->stream));
    *p = (T *)q;
    return 0;
}

template <class T> int upload(dopf_handle *h, const T **dst, const std::vector<T> &src)
{
    T *d = nullptr;
    int rc = dev_alloc(h, &d, src.size(), false);
    if (rc) return rc;
    if (!src.empty()) CK(cudaMemcpyAsync(d, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // src may be a temporary
    *dst = d;
    return 0;
}

int sync_ctrl(dopf_handle *h)
{
    CK(cudaMemcpyAsync(h->h_ctrl, h->lp.view.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

void fill_status(dopf_handle *h, dopf_status *s)
{
    const Ctrl &c = *h->h_ctrl;
    s->iteration = c.iteration; s->converged = c.converged;
    s->conv_lambda = c.conv_lambda; s->conv_mue = c.conv_mue; s->conv_rho = c.conv_rho;
    s->iterations_done = c.iters_done;
    s->res_lambda = c.res[0]; s->res_mue = c.res[1]; s->res_rho = c.res[2];
    s->gen_corrected = c.stat_gen_fix; s->sto_corrected = c.stat_sto_fix;
    s->tight_rows = c.stat_tight_rows; s->wide_rows = c.stat_wide_rows;
    s->launches_per_iteration = h->launches_per_iter;
    s->sto_cold = c.stat_sto_cold;
    s->fix_sequential = c.stat_fix_seq;
    s->last_step_ms = h->last_step_ms;
}

int check_device_error(dopf_handle *h)
{
    if (h->h_ctrl->error == DOPF_ERR_NONE) return 0;
    char buf[256];
    snprintf(buf, sizeof buf,
             h->h_ctrl->error == DOPF_ERR_HINGE_CAP
                 ? "hinge list capacity exceeded in the correction pass at iteration %d; recreate with a larger dopf_config.hinge_capacity (the iterate of iteration %d is intact)"
                 : "generator work list capacity exceeded at iteration %d (iterate of iteration %d intact)",
             h->h_ctrl->iteration, h->h_ctrl->iteration - 1);
    h->err = buf;
    return DOPF_E_CAPACITY;
}


// ---- NCCL, loaded at run time (libdopf.so itself does not depend on libnccl) --------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*get_unique_id)(ncclUniqueId *) = nullptr;
    ncclResult_t (*comm_init_rank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*all_reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*comm_destroy)(ncclComm_t) = nullptr;
    const char *(*get_error_string)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (lib) return true;
        const char *names[] = {getenv("DOPF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void *l = nullptr;
        for (const char *n : names) if (n && !l) l = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!l) { err = std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return false; }
#define SYM(field, name) field = (decltype(field))dlsym(l, name); if (!field) { err = "libnccl: missing symbol " name; return false; }
        SYM(get_unique_id, "ncclGetUniqueId"); SYM(comm_init_rank, "ncclCommInitRank"); SYM(all_reduce, "ncclAllReduce");
        SYM(comm_destroy, "ncclCommDestroy"); SYM(get_error_string, "ncclGetErrorString");
#undef SYM
        lib = l;
        return true;
    }
};
NcclApi g_nccl;

#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            h->err = std::string(#call " failed: ") + g_nccl.get_error_string(r_);                 \
            return DOPF_E_COMM;                                                                    \
        }                                                                                          \
    } while (0)

}  // namespace

namespace dopf {
int enqueue_iteration_comm(dopf_handle *h, cudaStream_t st);
}

extern "C" {

const char *dopf_version(void) { return "libdopf 0.1 (sm_100a)"; }

void dopf_default_config(dopf_config *c)
{
    c->gamma = 0.3; c->flow_weight = 10.0; c->prox_weight = 1.0; c->slack_mask_tol = 1e-2; c->eps = 1e-3;
    c->device = -1; c->hinge_capacity = 0; c->use_graph = 1; c->debug_flags = 0;
    c->n_scenarios = 1; c->gemm_ksplit = 0;
}

const char *dopf_last_error(dopf_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void dopf_destroy(dopf_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    if (h->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(h->comm);     // after the graphs that captured its collectives
    if (h->lp.ev_fork) cudaEventDestroy(h->lp.ev_fork);
    if (h->lp.ev_join) cudaEventDestroy(h->lp.ev_join);
    if (h->lp.side_stream) cudaStreamDestroy(h->lp.side_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (void *p : h->allocs) cudaFree(p);
    if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

static int create_impl(dopf_handle *h, const dopf_problem *p, const dopf_config *cfg)
{
    // a batch of C scenarios becomes ONE problem with C*T columns and C*G + C*S agents on virtual nodes c*N + n
    const int C = cfg->n_scenarios > 1 ? cfg->n_scenarios : 1;
    const int N = p->N, L = p->L, T = p->T, Gs = p->G, Ss = p->S;
    if (N < 1 || L < 1 || T < 1 || Gs < 0 || Ss < 0 || Gs + Ss < 1) { h->err = "invalid dimensions"; return DOPF_E_ARG; }
    if ((long long)C * Gs >= (1ll << 31) || (long long)C * Ss >= (1ll << 31) || (long long)C * T >= (1ll << 30) || (long long)C * N >= (1ll << 31)) { h->err = "batch too large"; return DOPF_E_UNSUPPORTED; }
    const int G = C * Gs, S = C * Ss, TC = C * T;
    if (!p->ptdf || !p->f_max || !p->demand || (G && (!p->gen_mc || !p->gen_pmax || !p->gen_node)) ||
        (S && (!p->sto_mc || !p->sto_pmax || !p->sto_emax || !p->sto_node))) { h->err = "null input array"; return DOPF_E_ARG; }
    if (!(cfg->gamma > 0.0) || !(cfg->flow_weight > 0.0) || !(cfg->prox_weight > 0.0)) { h->err = "gamma, flow_weight, prox_weight must be > 0"; return DOPF_E_ARG; }
    for (int g = 0; g < Gs; ++g) if (p->gen_node[g] < 0 || p->gen_node[g] >= N) { h->err = "gen_node out of range"; return DOPF_E_ARG; }
    for (int s = 0; s < Ss; ++s) if (p->sto_node[s] < 0 || p->sto_node[s] >= N) { h->err = "sto_node out of range"; return DOPF_E_ARG; }
    if ((long long)G * T >= (1ll << 31) || (long long)S * T >= (1ll << 31)) { h->err = "G*T or S*T exceeds 2^31"; return DOPF_E_UNSUPPORTED; }

    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0) {
        h->err = std::string("no CUDA device available (") + cudaGetErrorString(e0) + "); libdopf has no CPU path";
        return DOPF_E_CUDA;
    }
    if (cfg->device >= 0) CK(cudaSetDevice(cfg->device));
    CK(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = h->stream;
    CK(cudaMallocHost((void **)&h->h_ctrl, sizeof(Ctrl)));
    CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
    CK(cudaStreamCreateWithFlags(&h->lp.side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->lp.ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->lp.ev_join, cudaEventDisableTiming));
    memset(h->h_ctrl, 0, sizeof(Ctrl));

    LaunchPlan &lp = h->lp;
    View &v = lp.view;
    v.N = N; v.L = L; v.T = T; v.G = G; v.S = S; v.A = Gs + Ss; v.NS = C; v.TC = TC;
    v.ldt = round_up(TC, 32); v.Np = round_up(N, 64); v.Lp = round_up(L, 64);
    v.hcap = cfg->hinge_capacity > 0 ? cfg->hinge_capacity : 32;
    v.c = Coef::make(cfg->gamma, cfg->flow_weight, cfg->prox_weight, cfg->slack_mask_tol, cfg->eps);
    v.demand_on = 1;
    v.rowsum_is_corr = 1;
    v.debug = cfg->debug_flags;
    const int ldt = v.ldt, Np = v.Np, Lp = v.Lp;
    h->use_graph = cfg->use_graph != 0;
    h->cfg_use_graph = h->use_graph;

    // ---- agents sorted by node (stable), CSR offsets -------------------------------------------
    // caller's agent index of the batch: c*Gs + g (scenario-major); virtual node of it: c*N + node[g]
    auto gvn = [&](int a) { return (a / std::max(Gs, 1)) * N + p->gen_node[a % std::max(Gs, 1)]; };
    auto svn = [&](int a) { return (a / std::max(Ss, 1)) * N + p->sto_node[a % std::max(Ss, 1)]; };
    h->gen_perm.resize(G); std::iota(h->gen_perm.begin(), h->gen_perm.end(), 0);
    std::stable_sort(h->gen_perm.begin(), h->gen_perm.end(), [&](int a, int b) { return gvn(a) < gvn(b); });
    h->sto_perm.resize(S); std::iota(h->sto_perm.begin(), h->sto_perm.end(), 0);
    std::stable_sort(h->sto_perm.begin(), h->sto_perm.end(), [&](int a, int b) { return svn(a) < svn(b); });
    h->gen_identity = true; for (int i = 0; i < G; ++i) if (h->gen_perm[i] != i) h->gen_identity = false;
    h->sto_identity = true; for (int i = 0; i < S; ++i) if (h->sto_perm[i] != i) h->sto_identity = false;

    std::vector<double> gmc(G), gpm(G), smc(S), spm(S), sem(S);
    const int NV = C * N;      // virtual nodes
    std::vector<int> gnode(G), snode(S), gptr(NV + 1, 0), sptr(NV + 1, 0);
    // per-scenario agent data: [C][Gs] / [C][Ss] when C > 1 (the node of an agent is shared by all scenarios)
    for (int i = 0; i < G; ++i) { int o = h->gen_perm[i]; gmc[i] = p->gen_mc[o]; gpm[i] = p->gen_pmax[o]; gnode[i] = gvn(o); gptr[gnode[i] + 1]++; }
    for (int i = 0; i < S; ++i) { int o = h->sto_perm[i]; smc[i] = p->sto_mc[o]; spm[i] = p->sto_pmax[o]; sem[i] = p->sto_emax[o]; snode[i] = svn(o); sptr[snode[i] + 1]++; }
    for (int n = 0; n < NV; ++n) { gptr[n + 1] += gptr[n]; sptr[n + 1] += sptr[n]; }

    // ---- padded network data and derived statics -----------------------------------------------
    std::vector<double> ptdf((size_t)Lp * Np, 0.0), fmax(Lp, 0.0), demand((size_t)Np * ldt, 0.0);
    std::vector<double> q(Np, 0.0), prow(Lp, 0.0), mwide(Lp, 0.0), nag((size_t)C * Np, 0.0), rbox(Np, 0.0);
    for (int i = 0; i < G; ++i) { const int c = gnode[i] / N, n = gnode[i] % N; nag[(size_t)c * Np + n] += 1.0; rbox[n] = std::max(rbox[n], gpm[i]); }
    for (int i = 0; i < S; ++i) { const int c = snode[i] / N, n = snode[i] % N; nag[(size_t)c * Np + n] += 1.0; rbox[n] = std::max(rbox[n], 2.0 * spm[i]); }
    for (int l = 0; l < L; ++l) {
        fmax[l] = p->f_max[l];
        for (int n = 0; n < N; ++n) {
            const double a = p->ptdf[(size_t)l * N + n];
            ptdf[(size_t)l * Np + n] = a;
            q[n] += a * a;
            prow[l] = std::max(prow[l], std::fabs(a));
            mwide[l] = std::max(mwide[l], std::fabs(a) * rbox[n]);
        }
    }
    for (int c = 0; c < C; ++c) for (int n = 0; n < N; ++n) for (int t = 0; t < T; ++t)
        demand[(size_t)n * ldt + (size_t)c * T + t] = p->demand[((size_t)c * N + n) * T + t];      // [C][N][T] -> columns

    int rc;
#define UP(dst, vec) if ((rc = upload(h, &dst, vec))) return rc
    {
        std::vector<double> ptdfT((size_t)Np * Lp, 0.0);
        for (int l = 0; l < L; ++l) for (int n = 0; n < N; ++n) ptdfT[(size_t)n * Lp + l] = ptdf[(size_t)l * Np + n];
        UP(v.ptdfT, ptdfT);
    }
    UP(v.ptdf, ptdf); UP(v.fmax, fmax); UP(v.demand, demand); UP(v.q, q); UP(v.prow, prow); { const double *t_ = nullptr; UP(t_, mwide); v.mwide = const_cast<double *>(t_); UP(t_, rbox); v.rbox = const_cast<double *>(t_); } UP(v.nagents, nag);
    UP(v.gen_mc, gmc); UP(v.gen_pmax, gpm); UP(v.gen_node, gnode); UP(v.gen_ptr, gptr);
    UP(v.sto_mc, smc); UP(v.sto_pmax, spm); UP(v.sto_emax, sem); UP(v.sto_node, snode); UP(v.sto_ptr, sptr);
#undef UP
#define AL(ptr, count) if ((rc = dev_alloc(h, &ptr, (size_t)(count)))) return rc
    for (int k = 0; k < 2; ++k) {
        AL(v.P[k], (size_t)G * T); AL(v.D[k], (size_t)S * T); AL(v.C[k], (size_t)S * T);
        if (k == 0) { AL(v.ssum_part, (size_t)COLSUM_R * ldt); AL(v.colsum_cnt, ldt / 32); }
        AL(v.inj[k], (size_t)Np * ldt); AL(v.ssum[k], ldt); AL(v.flow[k], (size_t)Lp * ldt);
        AL(v.lam[k], ldt); AL(v.mu[k], (size_t)Lp * ldt); AL(v.rho[k], (size_t)Lp * ldt);
        v.injloc[k] = v.inj[k];
    }
    AL(v.E, (size_t)S * T); AL(v.eta, (size_t)S * T); AL(v.cold_work, S); AL(v.wide_b, (size_t)TC * 2 * L); AL(v.avgU, (size_t)Lp * ldt); AL(v.avgK, (size_t)Lp * ldt);
    AL(v.bplus, (size_t)Lp * ldt); AL(v.bminus, (size_t)Lp * ldt); AL(v.M, (size_t)Lp * ldt); AL(v.Wt, (size_t)Lp * ldt);
    AL(v.g0, (size_t)Np * ldt); AL(v.s1, (size_t)Np * ldt); AL(v.rg, (size_t)Np * ldt); AL(v.rg2, (size_t)Np * ldt);
    AL(v.dn, (size_t)Np * ldt); AL(v.dmax, ldt);
    for (int k = 0; k < 8; ++k) AL(v.nst[k], (size_t)Np * ldt);
    AL(v.flags, (size_t)ldt * Lp);
    AL(v.wide, (size_t)TC * 2 * L); AL(v.wcnt, TC); AL(v.tight, (size_t)TC * 2 * L); AL(v.tight_b, (size_t)TC * 2 * std::max(L, 1)); AL(v.tcnt, TC);
    v.gen_work_cap = (int)std::min<long long>((long long)G * T, 1ll << 30);
    AL(v.gen_work, (size_t)std::max(v.gen_work_cap, 1)); AL(v.gen_grp, (size_t)2 * std::max(v.gen_work_cap, 1)); AL(v.sto_work, S); AL(v.sto_flag, S);
    AL(v.fix_node_flag, NV); AL(v.fix_node_list, NV); AL(v.fix_node_slot, NV);
    v.pair_cap = 1 << 20;
    AL(v.pair_row, v.pair_cap); AL(v.pair_node, v.pair_cap); AL(v.pair_col, v.pair_cap); AL(v.pair_val, v.pair_cap);
    AL(v.pbase, (size_t)TC * 2 * L); AL(v.pcnt, (size_t)TC * 2 * L);
    AL(v.sc_iteration, C); AL(v.sc_converged, C); AL(v.sc_conv, 3 * C); AL(v.sc_res_bits, 3 * C); AL(v.sc_res, 3 * C);
    {
        std::vector<int> ones(C, 1);
        CK(cudaMemcpyAsync(v.sc_iteration, ones.data(), sizeof(int) * C, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (C > 1) AL(h->d_stage, (size_t)std::max(N, L) * TC);
    h->h_sc_i.resize((size_t)5 * C); h->h_sc_d.resize((size_t)3 * C);
    AL(v.rowsumU, (size_t)2 * Lp * ldt); v.rowsumK = v.rowsumU + (size_t)Lp * ldt;   // contiguous: one exchange
    AL(v.ctrl, 1);
    AL(v.counters, 32);
    AL(lp.tflag, (size_t)Lp * ldt);

    AL(h->d_scalar, std::max(4, C));
    AL(h->d_nodal, (size_t)C * N * T);

    // ---- launch plan -----------------------------------------------------------------------------
    lp.num_sms = prop.multiProcessorCount;
    auto plan_gemm = [&](int Mp, int Kp, int &bm, int &ks) {
        const int ct = ldt / 32;
        bm = ((Mp / 64) * ct >= lp.num_sms) ? 64 : 32;
        const int tiles = (Mp / bm) * ct;
        // split-K: the kernel runs in waves of num_sms blocks per resident slot, so pick the split with the
        // fewest waves per unit of work among those that give every SM at least two blocks
        const int ksmax = std::max(1, std::min(8, Kp / 16 / 8));
        if (cfg->gemm_ksplit > 0) { ks = std::min(cfg->gemm_ksplit, ksmax); return; }      // fixed summation order (bitwise reproducibility across batch sizes)
        double best = 1e300; ks = 1;
        for (int k = 1; k <= ksmax; ++k) {
            const int blocks = tiles * k;
            const double cost = (double)((blocks + lp.num_sms - 1) / lp.num_sms) / k + (blocks < 2 * lp.num_sms ? 1.0 : 0.0);
            if (cost < best - 1e-12) { best = cost; ks = k; }
        }
    };
    // the transposed product is only needed at the nodes that carry agents of this handle (all of them on one GPU,
    // a contiguous node range per rank in the agent-partitioned mode): restrict it to the 64-row tiles of that range
    int nmin = N, nmax = -1;
    for (int g = 0; g < Gs; ++g) { nmin = std::min(nmin, (int)p->gen_node[g]); nmax = std::max(nmax, (int)p->gen_node[g]); }
    for (int s2 = 0; s2 < Ss; ++s2) { nmin = std::min(nmin, (int)p->sto_node[s2]); nmax = std::max(nmax, (int)p->sto_node[s2]); }
    if (nmax < 0) { nmin = 0; nmax = 0; }
    lp.mt_base = (nmin / 64) * 64;
    lp.mt_rows = (nmax / 64 + 1) * 64 - lp.mt_base;
    plan_gemm(lp.mt_rows, Lp, lp.bm_t, lp.ksplit_t);
    plan_gemm(Lp, Np, lp.bm_n, lp.ksplit_n);
    AL(lp.part, (size_t)std::max(lp.ksplit_t * Np, lp.ksplit_n * Lp) * ldt);
    AL(lp.part2, (size_t)lp.ksplit_t * Np * ldt);
    {
        lp.sto_fix_blocks = std::max(1, std::min(lp.num_sms * 8, S));
        {   // every solver block owns one overflow slot of hinge lists: bound them by 1 GB for long horizons
            const size_t slot_b = (size_t)T * v.hcap * sizeof(Hinge) + (size_t)T * sizeof(int);
            lp.sto_fix_blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)lp.sto_fix_blocks, ((size_t)1 << 30) / slot_b));
        }
        {
            const int need = (T + 31) / 32;
            const int opts[6] = {1, 2, 3, 4, 6, 8};
            lp.sto_j = 0;
            for (int o : opts) if (o >= need) { lp.sto_j = o; break; }
            if (cfg->debug_flags & 32) lp.sto_j = 0;      // diagnostics: force the sequential (thread per storage) solvers
        }
        if (lp.sto_j > 0 && set_storage_smem_attr(T) != 0) { h->err = "cudaFuncSetAttribute(shared memory) failed"; return DOPF_E_CUDA; }
        // scratch: one slot per node with a storage on the work list (up to 512 MB) + one per solver block for the overflow
        const size_t slot_bytes = (size_t)T * v.hcap * sizeof(Hinge) + (size_t)T * sizeof(int);
        lp.sto_fix_slots = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(NV, 1), ((size_t)512 << 20) / slot_bytes));
        const size_t warps = (size_t)lp.sto_fix_slots + (size_t)lp.sto_fix_blocks;
        AL(lp.hinge_scratch, S ? warps * T * v.hcap : 1);
        AL(lp.hcnt_scratch, S ? warps * T : 1);
    }
    lp.gen_flat = (cfg->debug_flags & 8) ? 1 : ((cfg->debug_flags & 16) ? 0 : (G < 24ll * NV ? 1 : 0));      // bits 3/4 force one variant (A/B timing)
    if (T / (T % 4 == 0 ? 4 : (T % 2 == 0 ? 2 : 1)) > 1024) lp.gen_flat = 1;      // the node-major kernel maps the horizon to at most 1024 threads
    lp.slack_blocks_x = std::max(1, std::min(64, (8 * lp.num_sms + TC - 1) / TC));

    // ---- state before iteration 1 (admm.jl:29-36; helpers/results.jl:14-73 "zeros") ----------
    Ctrl c0;
    memset(&c0, 0, sizeof c0);
    c0.iteration = 1; c0.cur = 0;
    *h->h_ctrl = c0;
    CK(cudaMemcpyAsync(v.ctrl, h->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, h->stream));
    launch_rebuild_derived(lp, h->stream);   // P=D=C=0  =>  injection = -demand, flows, levels
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = sync_ctrl(h))) return rc;
    return 0;
#undef AL
}

int dopf_create(const dopf_problem *p, const dopf_config *c, dopf_handle **out)
{
    if (!p || !c || !out) { g_create_error = "null argument"; return DOPF_E_ARG; }
    dopf_handle *h = new dopf_handle();
    int rc = create_impl(h, p, c);
    if (rc) { g_create_error = h->err; dopf_destroy(h); *out = nullptr; return rc; }
    *out = h;
    return DOPF_OK;
}

static int build_graph(dopf_handle *h)
{
    if (h->graph_exec || !h->use_graph) return 0;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_iteration_comm(h, h->stream);
    cudaError_t e = cudaStreamEndCapture(h->stream, &h->graph);
    if (rc < 0) return rc;
    if (e != cudaSuccess) { h->err = std::string("graph capture failed: ") + cudaGetErrorString(e); return DOPF_E_CUDA; }
    h->launches_per_iter = rc;
    CK(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
    return 0;
}

int dopf_step(dopf_handle *h, int32_t max_iters, dopf_status *out)
{
    if (!h) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    int rc = 0;
    int remaining = max_iters;
    h->last_step_ms = 0.0;
    h->host_cur = h->h_ctrl->cur;
    while (remaining > 0 && !h->h_ctrl->converged && h->h_ctrl->error == 0) {
        // library-owned communicator: the very first iteration runs eagerly (NCCL sets its channels up on first use, which
        // must not happen inside a capture); the graph is captured afterwards
        const bool warm = h->comm && h->use_graph && !h->graph_exec && h->h_ctrl->iters_done == 0;
        if (!warm && (rc = build_graph(h))) {
            if (!h->comm) return rc;
            h->use_graph = false;                 // capture with collectives refused: keep going eagerly (all ranks see the same)
        }
        const int chunk = warm ? 1 : std::min(remaining, 64);
        CK(cudaEventRecord(h->ev0, h->stream));
        for (int i = 0; i < chunk; ++i) {
            if (h->graph_exec) CK(cudaGraphLaunch(h->graph_exec, h->stream));
            else {
                rc = enqueue_iteration_comm(h, h->stream);
                if (rc < 0) return rc;
                h->launches_per_iter = rc;
            }
        }
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaGetLastError());
        if ((rc = sync_ctrl(h))) return rc;
        h->host_cur = h->h_ctrl->cur;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->last_step_ms += ms;
        remaining -= chunk;
        if (h->comm) {
            // a rank that hit a device-side capacity error has turned its kernels into no-ops: every rank must stop at the
            // same chunk boundary (the others would go on all-reducing its stale buffers, and the collectives would no
            // longer pair up)
            int flag = h->h_ctrl->error != 0 ? 1 : 0;
            CK(cudaMemcpyAsync(h->d_flag, &flag, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            NK(g_nccl.all_reduce(h->d_flag, h->d_flag, 1, ncclInt32, ncclMax, h->comm, h->stream));
            CK(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (flag && h->h_ctrl->error == 0) { h->err = "another rank reported a device-side capacity error"; return DOPF_E_CAPACITY; }
        }
    }
    if ((rc = check_device_error(h))) return rc;
    if (out) fill_status(h, out);
    return DOPF_OK;
}

int dopf_profile_iteration(dopf_handle *h, int32_t cap, float *ms, const char **names, int32_t *count)
{
    if (!h || !ms || !names || !count || cap < 1) return DOPF_E_ARG;
    if (h->partitioned) { h->err = "dopf_profile_iteration: single-GPU handles only"; return DOPF_E_UNSUPPORTED; }
    CK(cudaSetDevice(h->device));
    std::vector<cudaEvent_t> ev(2 * (size_t)cap);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    std::vector<const char *> nm(cap, "");
    int n = 0;
    LaunchPlan lp = h->lp;
    lp.prof_events = ev.data(); lp.prof_names = nm.data(); lp.prof_cap = cap; lp.prof_count = &n;
    launch_profile_warm(lp, h->stream);
    enqueue_iteration(lp, h->stream);
    CK(cudaGetLastError());
    int rc = sync_ctrl(h);
    if (rc) return rc;
    n = std::min(n, (int)cap);
    for (int i = 0; i < n; ++i) { CK(cudaEventElapsedTime(&ms[i], ev[2 * i], ev[2 * i + 1])); names[i] = nm[i]; }
    for (auto &e : ev) cudaEventDestroy(e);
    *count = n;
    return check_device_error(h);
}

int dopf_get_status(dopf_handle *h, dopf_status *out)
{
    if (!h || !out) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    int rc = sync_ctrl(h);
    if (rc) return rc;
    fill_status(h, out);
    // partitioned handles are never stepped through dopf_step: this is where a device-side capacity error surfaces.
    // After an error the device stops flipping its buffers, so the host mirror of the buffer index is re-read too.
    h->host_cur = h->h_ctrl->cur;
    return check_device_error(h);
}

// copy a padded device matrix [rows][ldt] to a dense host matrix [rows][T]
// (scenario batches: host [C][rows][T] <-> device columns c*T + t, re-laid on the device through a staging buffer)
static int d2h_matrix(dopf_handle *h, double *dst, const double *src, int rows, int T, int ld)
{
    if (!dst) return 0;
    const int C = h->lp.view.NS;
    if (C > 1) {
        launch_pack_cols(const_cast<double *>(src), h->d_stage, rows, C, T, ld, 0, h->stream);
        CK(cudaMemcpyAsync(dst, h->d_stage, (size_t)C * rows * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));      // the staging buffer is reused by the next matrix
        return 0;
    }
    CK(cudaMemcpy2DAsync(dst, (size_t)T * sizeof(double), src, (size_t)ld * sizeof(double), (size_t)T * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}
static int h2d_matrix(dopf_handle *h, double *dst, const double *src, int rows, int T, int ld)
{
    if (!src) return 0;
    const int C = h->lp.view.NS;
    if (C > 1) {
        CK(cudaMemcpyAsync(h->d_stage, src, (size_t)C * rows * T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_pack_cols(dst, h->d_stage, rows, C, T, ld, 1, h->stream);
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    CK(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), src, (size_t)T * sizeof(double), (size_t)T * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
// agent matrices: rows are stored sorted by node on the device
static int d2h_agents(dopf_handle *h, double *dst, const double *src, int rows, int T, const std::vector<int> &perm, bool identity)
{
    if (!dst || rows == 0) return 0;
    if (identity) { CK(cudaMemcpyAsync(dst, src, (size_t)rows * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); return 0; }
    h->stage.resize((size_t)rows * T);
    CK(cudaMemcpyAsync(h->stage.data(), src, (size_t)rows * T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < rows; ++i) memcpy(dst + (size_t)perm[i] * T, h->stage.data() + (size_t)i * T, (size_t)T * sizeof(double));
    return 0;
}
static int h2d_agents(dopf_handle *h, double *dst, const double *src, int rows, int T, const std::vector<int> &perm, bool identity)
{
    if (!src || rows == 0) return 0;
    if (identity) { CK(cudaMemcpyAsync(dst, src, (size_t)rows * T * sizeof(double), cudaMemcpyHostToDevice, h->stream)); return 0; }
    h->stage.resize((size_t)rows * T);
    for (int i = 0; i < rows; ++i) memcpy(h->stage.data() + (size_t)i * T, src + (size_t)perm[i] * T, (size_t)T * sizeof(double));
    CK(cudaMemcpyAsync(dst, h->stage.data(), (size_t)rows * T * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int dopf_get_iterate(dopf_handle *h, double *P, double *D, double *C, double *E, double *injection, double *flow, double *avgU, double *avgK)
{
    if (!h) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    const View &v = h->lp.view;
    const int k = h->h_ctrl->cur;   // newest iterate (buffers were flipped by the last iteration)
    int rc;
    if ((rc = d2h_agents(h, P, v.P[k], v.G, v.T, h->gen_perm, h->gen_identity))) return rc;
    if ((rc = d2h_agents(h, D, v.D[k], v.S, v.T, h->sto_perm, h->sto_identity))) return rc;
    if ((rc = d2h_agents(h, C, v.C[k], v.S, v.T, h->sto_perm, h->sto_identity))) return rc;
    if ((rc = d2h_agents(h, E, v.E, v.S, v.T, h->sto_perm, h->sto_identity))) return rc;
    if ((rc = d2h_matrix(h, injection, v.inj[k], v.N, v.T, v.ldt))) return rc;
    if ((rc = d2h_matrix(h, flow, v.flow[k], v.L, v.T, v.ldt))) return rc;
    if ((rc = d2h_matrix(h, avgU, v.avgU, v.L, v.T, v.ldt))) return rc;
    if ((rc = d2h_matrix(h, avgK, v.avgK, v.L, v.T, v.ldt))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_get_duals(dopf_handle *h, int32_t which, double *lam, double *mu, double *rho)
{
    if (!h || which < 0 || which > 1) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    const View &v = h->lp.view;
    const int k = which == 0 ? h->h_ctrl->cur : 1 - h->h_ctrl->cur;
    int rc;
    if (lam) CK(cudaMemcpyAsync(lam, v.lam[k], (size_t)v.TC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if ((rc = d2h_matrix(h, mu, v.mu[k], v.L, v.T, v.ldt))) return rc;
    if ((rc = d2h_matrix(h, rho, v.rho[k], v.L, v.T, v.ldt))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_set_state(dopf_handle *h, int32_t iteration, const double *P, const double *D, const double *C,
                   const double *avgU, const double *avgK, const double *lam, const double *mu, const double *rho)
{
    if (!h || iteration < 1) return DOPF_E_ARG;
    if (h->partitioned) {     // the derived network state would need the exchanges of the partitioned mode
        h->err = "dopf_set_state is not supported after dopf_set_partition";
        return DOPF_E_UNSUPPORTED;
    }
    CK(cudaSetDevice(h->device));
    View &v = h->lp.view;
    int rc;
    if ((rc = sync_ctrl(h))) return rc;
    const int cur = h->h_ctrl->cur, nxt = 1 - cur;
    // the new "previous iterate" is staged in the inactive buffers, then the buffers are flipped
    if (!P && v.G) CK(cudaMemcpyAsync(v.P[nxt], v.P[cur], (size_t)v.G * v.T * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (!D && v.S) CK(cudaMemcpyAsync(v.D[nxt], v.D[cur], (size_t)v.S * v.T * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (!C && v.S) CK(cudaMemcpyAsync(v.C[nxt], v.C[cur], (size_t)v.S * v.T * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = h2d_agents(h, v.P[nxt], P, v.G, v.T, h->gen_perm, h->gen_identity))) return rc;
    if ((rc = h2d_agents(h, v.D[nxt], D, v.S, v.T, h->sto_perm, h->sto_identity))) return rc;
    if ((rc = h2d_agents(h, v.C[nxt], C, v.S, v.T, h->sto_perm, h->sto_identity))) return rc;
    if ((rc = h2d_matrix(h, v.avgU, avgU, v.L, v.T, v.ldt))) return rc;
    if ((rc = h2d_matrix(h, v.avgK, avgK, v.L, v.T, v.ldt))) return rc;
    const size_t lt = (size_t)v.Lp * v.ldt * sizeof(double);
    if (lam) CK(cudaMemcpyAsync(v.lam[nxt], lam, (size_t)v.TC * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    else CK(cudaMemcpyAsync(v.lam[nxt], v.lam[cur], (size_t)v.ldt * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (mu) { if ((rc = h2d_matrix(h, v.mu[nxt], mu, v.L, v.T, v.ldt))) return rc; }
    else CK(cudaMemcpyAsync(v.mu[nxt], v.mu[cur], lt, cudaMemcpyDeviceToDevice, h->stream));
    if (rho) { if ((rc = h2d_matrix(h, v.rho[nxt], rho, v.L, v.T, v.ldt))) return rc; }
    else CK(cudaMemcpyAsync(v.rho[nxt], v.rho[cur], lt, cudaMemcpyDeviceToDevice, h->stream));
    Ctrl c = *h->h_ctrl;
    c.iteration = iteration; c.converged = c.conv_lambda = c.conv_mue = c.conv_rho = 0; c.error = 0;
    c.iters_done = 0;
    {
        std::vector<int> its(v.NS, iteration);
        CK(cudaMemcpyAsync(v.sc_iteration, its.data(), sizeof(int) * v.NS, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemsetAsync(v.sc_converged, 0, sizeof(int) * v.NS, h->stream)); CK(cudaMemsetAsync(v.sc_conv, 0, sizeof(int) * 3 * v.NS, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    *h->h_ctrl = c;
    CK(cudaMemcpyAsync(v.ctrl, h->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, h->stream));
    launch_rebuild_derived(h->lp, h->stream);
    CK(cudaGetLastError());
    if ((rc = sync_ctrl(h))) return rc;
    return DOPF_OK;
}

int dopf_get_nodal_price(dopf_handle *h, int32_t which, double *out)
{
    if (!h || !out || which < 0 || which > 1) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    const View &v = h->lp.view;
    const int k = which == 0 ? h->h_ctrl->cur : 1 - h->h_ctrl->cur;
    launch_nodal_price(v, v.lam[k], v.mu[k], v.rho[k], h->d_nodal, h->stream);
    CK(cudaMemcpyAsync(out, h->d_nodal, (size_t)v.NS * v.N * v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_nodal_price_from(dopf_handle *h, const double *lam, const double *mu, const double *rho, double *out)
{
    if (!h || !lam || !mu || !rho || !out) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    const View &v = h->lp.view;
    // staging in per-iteration scratch (M, Wt, first row of g0): all three are rewritten at the start of every iteration
    int rc;
    CK(cudaMemcpyAsync(v.g0, lam, (size_t)v.TC * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = h2d_matrix(h, v.M, mu, v.L, v.T, v.ldt))) return rc;
    if ((rc = h2d_matrix(h, v.Wt, rho, v.L, v.T, v.ldt))) return rc;
    launch_nodal_price(v, v.g0, v.M, v.Wt, h->d_nodal, h->stream);
    CK(cudaMemcpyAsync(out, h->d_nodal, (size_t)v.NS * v.N * v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

static int need_iteration_done(dopf_handle *h, const char *what)
{
    if (h->h_ctrl->iters_done < 1) { h->err = std::string(what) + ": no iteration has run yet"; return DOPF_E_ARG; }
    if (h->partitioned) { h->err = std::string(what) + ": single-GPU handles only"; return DOPF_E_UNSUPPORTED; }
    return 0;
}

int dopf_get_unit_penalty(dopf_handle *h, int32_t kind, int32_t index, double *eb, double *upper, double *lower, double *U, double *K)
{
    if (!h || kind < 0 || kind > 1 || !eb || !upper || !lower) return DOPF_E_ARG;
    const View &v = h->lp.view;
    if (index < 0 || index >= (kind == 0 ? v.G : v.S)) return DOPF_E_ARG;      // batches: index = c*G_per_scenario + g
    CK(cudaSetDevice(h->device));
    int rc = need_iteration_done(h, "dopf_get_unit_penalty");
    if (rc) return rc;
    if (!h->d_pen) { if ((rc = dev_alloc(h, &h->d_pen, (size_t)3 * v.TC + (size_t)2 * v.L * v.T))) return rc; }
    // caller's index -> position in the node-sorted device order
    const std::vector<int> &perm = kind == 0 ? h->gen_perm : h->sto_perm;
    int pos = index;
    if (!(kind == 0 ? h->gen_identity : h->sto_identity)) pos = (int)(std::find(perm.begin(), perm.end(), index) - perm.begin());
    double *dU = h->d_pen + (size_t)3 * v.TC, *dK = dU + (size_t)v.L * v.T;
    launch_unit_penalty(v, kind, pos, h->d_pen, h->d_pen + v.TC, h->d_pen + 2 * v.TC, U ? dU : nullptr, K ? dK : nullptr, h->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(eb, h->d_pen, (size_t)v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(upper, h->d_pen + v.TC, (size_t)v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(lower, h->d_pen + 2 * v.TC, (size_t)v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (U) CK(cudaMemcpyAsync(U, dU, (size_t)v.L * v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (K) CK(cudaMemcpyAsync(K, dK, (size_t)v.L * v.T * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_get_penalty_totals(dopf_handle *h, double *eb, double *upper, double *lower)
{
    if (!h || !eb || !upper || !lower) return DOPF_E_ARG;
    const View &v = h->lp.view;
    CK(cudaSetDevice(h->device));
    int rc = need_iteration_done(h, "dopf_get_penalty_totals");
    if (rc) return rc;
    if (!h->d_pen) { if ((rc = dev_alloc(h, &h->d_pen, (size_t)3 * v.TC + (size_t)2 * v.L * v.T))) return rc; }
    launch_penalty_totals(v, h->d_pen, h->d_pen + v.TC, h->d_pen + 2 * v.TC, h->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(eb, h->d_pen, (size_t)v.TC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(upper, h->d_pen + v.TC, (size_t)v.TC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(lower, h->d_pen + 2 * v.TC, (size_t)v.TC * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_get_scenario_status(dopf_handle *h, int32_t *iteration, int32_t *converged, double *residuals)
{
    if (!h) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    const View &v = h->lp.view;
    const size_t C = (size_t)v.NS;
    if (iteration) CK(cudaMemcpyAsync(iteration, v.sc_iteration, C * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (converged) CK(cudaMemcpyAsync(converged, v.sc_converged, C * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (residuals) CK(cudaMemcpyAsync(residuals, v.sc_res, 3 * C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_debug_counters(dopf_handle *h, uint64_t *out, int32_t reset)
{
    if (!h || !out) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->lp.view.counters, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    if (reset) CK(cudaMemsetAsync(h->lp.view.counters, 0, 32 * sizeof(uint64_t), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_get_total_costs(dopf_handle *h, double *out)
{
    if (!h || !out) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    launch_total_costs(h->lp.view, h->d_scalar, h->stream);
    CK(cudaMemcpyAsync(out, h->d_scalar, sizeof(double) * h->lp.view.NS, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DOPF_OK;
}

int dopf_set_stream(dopf_handle *h, void *cuda_stream)
{
    if (!h) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }   // graphs are stream-agnostic, but keep it simple
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
    return DOPF_OK;
}

int dopf_set_partition(dopf_handle *h, int32_t rank, int32_t nranks, int32_t total_agents)
{
    if (!h || nranks < 1 || rank < 0 || rank >= nranks || total_agents < h->lp.view.A) return DOPF_E_ARG;
    if (h->h_ctrl->iters_done != 0) { h->err = "dopf_set_partition must precede the first iteration"; return DOPF_E_ARG; }
    if (h->lp.view.NS > 1) { h->err = "scenario batches are sharded by scenario (independent handles), not by agents"; return DOPF_E_UNSUPPORTED; }
    CK(cudaSetDevice(h->device));
    h->rank = rank; h->nranks = nranks;
    View &v = h->lp.view;
    v.A = total_agents;
    // no rank carries the demand in its local injection: the all-reduced injection of the agents gets it subtracted
    // (launch_copy_inj), and the flows are sum over the ranks of PTDF[:, own nodes] * (own injection) minus the constant
    // PTDF * demand - so the flow product costs every rank 1/nranks of the single-GPU product
    v.demand_on = 0;
    if (!h->partitioned) {
        // the local injection of BOTH iterates goes to one fixed exchange buffer: every device pointer an enqueued phase
        // uses is then independent of the ping-pong parity, so the caller may capture an iteration (phases + its own
        // collectives) in a CUDA graph and replay it
        double *q = nullptr, *q3 = nullptr, *fd = nullptr;
        int r2 = dev_alloc(h, &q, (size_t)v.Np * v.ldt);
        if (r2) return r2;
        v.injloc[0] = v.injloc[1] = q;
        // row sums and partial flows in ONE contiguous exchange buffer (one collective)
        const size_t lt = (size_t)v.Lp * v.ldt;
        if ((r2 = dev_alloc(h, &q3, 3 * lt))) return r2;
        if ((r2 = dev_alloc(h, &fd, lt))) return r2;
        v.rowsumU = q3; v.rowsumK = q3 + lt; v.xflow = q3 + 2 * lt;
        LaunchPlan &lp = h->lp;
        lp.bm_x = lp.bm_n;
        const int tiles = (v.Lp / lp.bm_x) * (v.ldt / 32);
        const int ksmax = std::max(1, std::min(8, lp.mt_rows / 16 / 8));
        lp.ksplit_x = std::max(1, std::min(ksmax, (2 * lp.num_sms + tiles - 1) / tiles));
        launch_flow_of_demand(lp, fd, h->stream);
        CK(cudaGetLastError());
        v.flowD = fd;
        h->partitioned = true;
    }
    h->use_graph = false;
    // local part of the state before iteration 1: staged injection (-demand on rank 0) in the inactive
    // buffers; the caller all-reduces DOPF_XBUF_INJ and then runs dopf_step_phase(h, -1) to finish
    int rc = sync_ctrl(h);
    if (rc) return rc;
    h->host_cur = h->h_ctrl->cur;
    launch_rebuild_derived(h->lp, h->stream, 0);
    CK(cudaGetLastError());
    return DOPF_OK;
}

int dopf_comm_get_unique_id(void *id_out)
{
    std::string err;
    if (!id_out) return DOPF_E_ARG;
    if (!g_nccl.load(err)) { g_create_error = err; return DOPF_E_COMM; }
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.get_unique_id(&id);
    if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId failed: ") + g_nccl.get_error_string(r); return DOPF_E_COMM; }
    static_assert(sizeof(ncclUniqueId) == DOPF_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof id);
    return DOPF_OK;
}

int dopf_comm_init(dopf_handle *h, const void *id_bytes, int32_t rank, int32_t nranks, int32_t total_agents)
{
    if (!h || !id_bytes) return DOPF_E_ARG;
    if (h->comm) { h->err = "dopf_comm_init: the handle already has a communicator"; return DOPF_E_ARG; }
    if (!g_nccl.load(h->err)) return DOPF_E_COMM;
    int rc = dopf_set_partition(h, rank, nranks, total_agents);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    NK(g_nccl.comm_init_rank(&h->comm, nranks, id, rank));
    if ((rc = dev_alloc(h, &h->d_flag, 1))) return rc;
    // set-up exchanges: initial injection of the agents, per-node box ranges (identical candidate rows on all ranks)
    View &v = h->lp.view;
    NK(g_nccl.all_reduce(v.injloc[0], v.injloc[0], (size_t)v.Np * v.ldt, ncclFloat64, ncclSum, h->comm, h->stream));
    NK(g_nccl.all_reduce(v.rbox, v.rbox, (size_t)v.Np, ncclFloat64, ncclMax, h->comm, h->stream));
    if ((rc = dopf_step_phase(h, -1))) return rc;
    if ((rc = sync_ctrl(h))) return rc;
    h->host_cur = h->h_ctrl->cur;
    h->use_graph = h->cfg_use_graph;          // the whole iteration - kernels and collectives - is captured by dopf_step
    return DOPF_OK;
}

int dopf_step_phase(dopf_handle *h, int32_t phase)
{
    if (!h || phase < -1 || phase >= DOPF_N_SEGMENTS) return DOPF_E_ARG;
    CK(cudaSetDevice(h->device));
    View &v = h->lp.view;
    if (phase == -1) {   // second half of the initial state after the injection / box-range exchanges
        if (!h->partitioned) { h->err = "dopf_step_phase: call dopf_set_partition first"; return DOPF_E_ARG; }
        launch_mwide(v, h->stream);
        launch_copy_inj(v, v.injloc[0], h->stream);       // the all-reduced initial injection
        launch_rebuild_derived(h->lp, h->stream, 1);
        h->host_cur = 1 - h->host_cur;
        CK(cudaGetLastError());
        return DOPF_OK;
    }
    if (!h->partitioned) { h->err = "dopf_step_phase: call dopf_set_partition first"; return DOPF_E_ARG; }
    if (phase == DOPF_X_INJ + 1) launch_copy_inj(v, v.injloc[0], h->stream);   // the injection exchanged after the previous phase
    const int n = enqueue_iteration(h->lp, h->stream, phase);
    if (phase == 0) h->launches_per_iter = 0;
    h->launches_per_iter += n + (phase == DOPF_X_INJ + 1 ? 1 : 0);
    if (phase == DOPF_N_SEGMENTS - 1) h->host_cur = 1 - h->host_cur;
    CK(cudaGetLastError());
    return DOPF_OK;
}

int dopf_exchange_buffer(dopf_handle *h, int32_t which, void **device_ptr, int64_t *count)
{
    if (!h || !device_ptr || !count) return DOPF_E_ARG;
    View &v = h->lp.view;
    if (!h->partitioned) { h->err = "dopf_exchange_buffer: call dopf_set_partition first"; return DOPF_E_ARG; }
    switch (which) {      // all four are fixed device addresses for the lifetime of the handle
    case DOPF_XBUF_DMAX: *device_ptr = v.dmax; *count = v.ldt; break;
    case DOPF_XBUF_INJ: *device_ptr = v.injloc[0]; *count = (int64_t)v.Np * v.ldt; break;
    case DOPF_XBUF_ROWSUM: *device_ptr = v.rowsumU; *count = (int64_t)3 * v.Lp * v.ldt; break;   // + the partial flows
    case DOPF_XBUF_RBOX: *device_ptr = v.rbox; *count = v.Np; break;
    default: return DOPF_E_ARG;
    }
    return DOPF_OK;
}

}  // extern "C"

namespace dopf {

// one whole iteration (single-GPU handles); returns the number of kernel launches or < 0
int enqueue_iteration_comm(dopf_handle *h, cudaStream_t st)
{
    if (h->partitioned && !h->comm) {
        h->err = "partitioned handles without a library communicator (dopf_comm_init) are stepped with dopf_step_phase (the caller does the exchanges)";
        return DOPF_E_UNSUPPORTED;
    }
    if (!h->partitioned) return enqueue_iteration(h->lp, st);
    // the four phases with the three exchanges in between, all on one stream (capturable: every buffer has a fixed address)
    const View &v = h->lp.view;
    int launches = 0;
    for (int phase = 0; phase < DOPF_N_SEGMENTS; ++phase) {
        if (phase == DOPF_X_INJ + 1) { launch_copy_inj(v, v.injloc[0], st); ++launches; }
        launches += enqueue_iteration(h->lp, st, phase);
        void *buf = nullptr; int64_t cnt = 0;
        if (phase < DOPF_N_SEGMENTS - 1) {
            dopf_exchange_buffer(h, phase, &buf, &cnt);
            NK(g_nccl.all_reduce(buf, buf, (size_t)cnt, ncclFloat64, phase == DOPF_X_DMAX ? ncclMax : ncclSum, h->comm, st));
        }
    }
    return launches;
}

}  // namespace dopf
