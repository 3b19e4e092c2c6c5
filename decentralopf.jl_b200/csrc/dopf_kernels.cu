// sm_100a kernels of one ADMM iteration (calculate_iteration!, /root/reference/src/optimization/
// run.jl:7-16).  Layout and bodies: dopf_bodies.h; per-agent math: dopf_math.h; pipeline order:
// dopf_api.cu (enqueue_iteration) and DESIGN.md section 4.
//
// Every kernel first reads the device control block: once `converged` or `error` is set all
// remaining launches of an already enqueued batch are no-ops, so run!(admm) needs no host
// round trip per iteration.
#include "dopf_kernels.h"
#include "dopf_sto_warp.cuh"

namespace dopf {

#define DOPF_ACTIVE(v) ((v).ctrl->converged == 0 && (v).ctrl->error == 0)

// ------------------------------------------------------------------------------------------------
// row preparation over the padded [Lp][ldt] grid (coalesced along t)
// ------------------------------------------------------------------------------------------------
__global__ void k_row_prep(View v, unsigned char *tflag)
{
    if (!DOPF_ACTIVE(v)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.Lp * v.ldt) return;
    body_row_prep(v, i / v.ldt, i % v.ldt);
    tflag[i] = 0; v.rowsumU[i] = 0.0; v.rowsumK[i] = 0.0;      // exact-slack-sum marks of this iteration (k_slack_rows)
}

// ordered compaction of the candidate rows of one timestep; one block (8 warps) per t: every warp
// counts its contiguous slice, the offsets come from a prefix over the 8 counts, then the slice is
// written in order (deterministic list order).
//  mode 0: wide list from the flag bytes;  mode 1: tight list = wide entries with
//  |b| <= max_n|ptdf[l,n]| * (largest move of any agent at t)
__global__ void __launch_bounds__(256) k_compact(View v, int mode, int inline_dmax)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ int wcount[8];
    __shared__ unsigned long long wmax[8];
    __shared__ int unchanged;
    const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (mode == 1 && inline_dmax) {
        // dmax[t] = column maximum of dn (largest move of any agent at t); the separate k_dmax is only needed when
        // the maxima are exchanged between ranks before the lists are built
        unsigned long long a = 0ull;
        for (int n = threadIdx.x; n < v.N; n += blockDim.x) { const unsigned long long b = v.dn[(size_t)n * v.ldt + t]; a = b > a ? b : a; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, a, o); a = y > a ? y : a; }
        if (lane == 0) wmax[warp] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < 8; ++k) a = wmax[k] > a ? wmax[k] : a;
            unchanged = inline_dmax == 2 && a == v.dmax[t];
            v.dmax[t] = a;
        }
        __syncthreads();
        if (unchanged) return;      // second pass of the iteration: the correction left this column's largest move as it was => same list
    }
    const int n_in = mode == 0 ? v.L : v.wcnt[t];
    // slice length per warp: a multiple of 128 rows in mode 0 (4 flag bytes per lane), of 32 entries in mode 1
    const int gran = mode == 0 ? 128 : 32;
    const int per = ((n_in + 8 * gran - 1) / (8 * gran)) * gran;
    const int i0 = warp * per, i1 = min(i0 + per, n_in);
    const unsigned char *fl = v.flags + (size_t)t * v.Lp;
    const int *in = v.wide + (size_t)t * 2 * v.L;
    const double dm = mode == 1 ? bits_nonneg(v.dmax[t]) : 0.0;
    int *out = (mode == 0 ? v.wide : v.tight) + (size_t)t * 2 * v.L;
    for (int pass = 0; pass < 2; ++pass) {
        int cnt = 0;
        if (pass == 1) { for (int w = 0; w < warp; ++w) cnt += wcount[w]; }
        if (mode == 0) {
            // lane owns rows base+4*lane .. +3 (one 32-bit load; Lp is a multiple of 64, rows >= L carry no flags);
            // list order = row order: exclusive prefix of the lane counts, then the lane writes its entries in order
            for (int base = i0; base < i1; base += 128) {
                const int r0 = base + 4 * lane;
                const unsigned w4 = r0 < i1 ? *reinterpret_cast<const unsigned *>(fl + r0) : 0u;
                unsigned bits = 0;                                   // bit 2*k+side of row r0+k
#pragma unroll
                for (int k = 0; k < 4; ++k) if (r0 + k < i1) bits |= ((w4 >> (8 * k)) & 3u) << (2 * k);
                const int mine = __popc(bits);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                if (pass == 1) {
                    int pos = cnt + incl - mine;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if ((bits >> q) & 1u) {
                            const int i = r0 + (q >> 1), side = q & 1;
                            out[pos] = i * 2 + side;
                            v.wide_b[(size_t)t * 2 * v.L + pos] = side ? v.bminus[(size_t)i * v.ldt + t] : v.bplus[(size_t)i * v.ldt + t];
                            ++pos;
                        }
                    }
                }
                cnt += __shfl_sync(0xffffffffu, incl, 31);
            }
        } else {
            for (int base = i0; base < i1; base += 32) {
                const int i = base + lane;
                bool p = false;
                int e = 0;
                double be = 0.0;
                if (i < i1) {
                    e = in[i];
                    be = v.wide_b[(size_t)t * 2 * v.L + i];
                    p = fabs(be) <= v.prow[e >> 1] * dm;
                }
                const unsigned m = __ballot_sync(0xffffffffu, p);
                if (pass == 1 && p) {
                    const int pos = cnt + __popc(m & ((1u << lane) - 1));
                    out[pos] = e;
                    v.tight_b[(size_t)t * 2 * v.L + pos] = be;
                }
                cnt += __popc(m);
            }
        }
        if (pass == 0) {
            if (lane == 0) wcount[warp] = cnt;
            __syncthreads();
        } else if (warp == 7 && lane == 0) {
            (mode == 0 ? v.wcnt : v.tcnt)[t] = cnt;
        }
    }
}

// the same lists for grids with few lines (L <= 1024): one WARP per column, 8 columns per block - a block per column
// would leave most of its warps without rows (scenario batches of small grids have tens of thousands of columns)
__global__ void __launch_bounds__(256) k_compact_w(View v, int mode, int inline_dmax)
{
    if (!DOPF_ACTIVE(v)) return;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= v.TC) return;
    if (mode == 1 && inline_dmax) {
        unsigned long long a = 0ull;
        for (int n = lane; n < v.N; n += 32) { const unsigned long long b = v.dn[(size_t)n * v.ldt + t]; a = b > a ? b : a; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, a, o); a = y > a ? y : a; }
        const bool same = inline_dmax == 2 && a == v.dmax[t];
        __syncwarp();
        if (lane == 0) v.dmax[t] = a;
        __syncwarp();
        if (same) return;           // (see k_compact)
    }
    const int n_in = mode == 0 ? v.L : v.wcnt[t];
    const unsigned char *fl = v.flags + (size_t)t * v.Lp;
    const int *in = v.wide + (size_t)t * 2 * v.L;
    const double dm = mode == 1 ? bits_nonneg(__shfl_sync(0xffffffffu, v.dmax[t], 0)) : 0.0;
    int *out = (mode == 0 ? v.wide : v.tight) + (size_t)t * 2 * v.L;
    int cnt = 0;
    if (mode == 0) {
        for (int base = 0; base < n_in; base += 128) {
            const int r0 = base + 4 * lane;
            const unsigned w4 = r0 < n_in ? *reinterpret_cast<const unsigned *>(fl + r0) : 0u;
            unsigned bits = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (r0 + k < n_in) bits |= ((w4 >> (8 * k)) & 3u) << (2 * k);
            const int mine = __popc(bits);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            int pos = cnt + incl - mine;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if ((bits >> q) & 1u) {
                    const int i = r0 + (q >> 1), side = q & 1;
                    out[pos] = i * 2 + side;
                    v.wide_b[(size_t)t * 2 * v.L + pos] = side ? v.bminus[(size_t)i * v.ldt + t] : v.bplus[(size_t)i * v.ldt + t];
                    ++pos;
                }
            }
            cnt += __shfl_sync(0xffffffffu, incl, 31);
        }
    } else {
        for (int base = 0; base < n_in; base += 32) {
            const int i = base + lane;
            bool p = false; int e = 0; double be = 0.0;
            if (i < n_in) { e = in[i]; be = v.wide_b[(size_t)t * 2 * v.L + i]; p = fabs(be) <= v.prow[e >> 1] * dm; }
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) { const int pos = cnt + __popc(m & ((1u << lane) - 1)); out[pos] = e; v.tight_b[(size_t)t * 2 * v.L + pos] = be; }
            cnt += __popc(m);
        }
    }
    if (lane == 0) (mode == 0 ? v.wcnt : v.tcnt)[t] = cnt;
}

// ------------------------------------------------------------------------------------------------
// fp64 tensor-pipe GEMMs (DMMA m8n8k4; tcgen05 has no f64 kind).  cp.async double buffering,
// split-K partials reduced by the epilogue kernels below (deterministic order).
//   TRANS = false:  Cpart[z] = A[Mp x Kp] * B[Kp x ldt]                      (flow = PTDF * inj)
//   TRANS = true :  Cpart[z] = A^T * B  and  C2part[z] = (A.*A)^T * B2      (PTDF^T M, (PTDF.^2)^T W)
// A is the padded PTDF [Lp][Np]; all dimensions are multiples of the tile sizes.
// ------------------------------------------------------------------------------------------------
constexpr int BN = 32, BK = 16;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int BM, bool TRANS>
__global__ void __launch_bounds__(128) k_gemm(View v, const double *__restrict__ A, int lda, int ldb,
                                              double *__restrict__ Cpart, double *__restrict__ C2part,
                                              int Mp, int Kp, int ksplit, int m_base,
                                              const double *__restrict__ Bsrc = nullptr, int k_base = 0)
{
    if (!DOPF_ACTIVE(v)) return;
    // B operand: M (and W) for the transposed product, the new injection for the flow product (Bsrc: another [K][ldt]
    // matrix; k_base: first row of the K range [k_base, k_base + Kp) of A's columns / B's rows)
    const double *__restrict__ B = Bsrc ? Bsrc : (TRANS ? v.M : sel(v.inj, 1 - v.ctrl->cur));
    const double *__restrict__ B2 = v.Wt;
    constexpr int MT = BM / 32;                       // m8 tiles per warp (4 warps along M)
    constexpr int AROW = TRANS ? (BM + 4) : (BK + 4); // padded smem row of the A tile
    constexpr int AROWS = TRANS ? BK : BM;
    // cp.async pipeline depth: 3 stages (one barrier per k-step, two tile loads in flight) where the static 48 KB allow it
    constexpr int ST = (TRANS && BM == 64) ? 2 : 3;
    __shared__ __align__(16) double As[ST][AROWS][AROW];
    __shared__ __align__(16) double Bs[ST][BK][BN + 4];
    __shared__ __align__(16) double B2s[TRANS ? ST : 1][TRANS ? BK : 1][TRANS ? BN + 4 : 2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = m_base + blockIdx.x * BM, n0 = blockIdx.y * BN, z = blockIdx.z;
    const int ksteps_total = Kp / BK;
    const int per = (ksteps_total + ksplit - 1) / ksplit;
    const int ks0 = z * per, ks1 = min(ksteps_total, ks0 + per);

    double acc[MT][4][2], acc2[TRANS ? MT : 1][4][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = 0.0; if (TRANS) { acc2[i][j][0] = acc2[i][j][1] = 0.0; } }

    auto load_tiles = [&](int buf, int ks) {
        const int k0 = k_base + ks * BK;
        if (TRANS) {   // A rows = k (lines), cols = m (nodes): BK x BM doubles
            constexpr int CH = BK * BM / 2;               // 16-byte chunks
            for (int c = tid; c < CH; c += 128) {
                const int r = c / (BM / 2), cc = (c % (BM / 2)) * 2;
                cp_async16(&As[buf][r][cc], A + (size_t)(k0 + r) * lda + m0 + cc);
            }
        } else {       // A rows = m (lines), cols = k (nodes): BM x BK doubles
            constexpr int CH = BM * BK / 2;
            for (int c = tid; c < CH; c += 128) {
                const int r = c / (BK / 2), cc = (c % (BK / 2)) * 2;
                cp_async16(&As[buf][r][cc], A + (size_t)(m0 + r) * lda + k0 + cc);
            }
        }
        constexpr int CHB = BK * BN / 2;
        for (int c = tid; c < CHB; c += 128) {
            const int r = c / (BN / 2), cc = (c % (BN / 2)) * 2;
            cp_async16(&Bs[buf][r][cc], B + (size_t)(k0 + r) * ldb + n0 + cc);
            if (TRANS) cp_async16(&B2s[buf][r][cc], B2 + (size_t)(k0 + r) * ldb + n0 + cc);
        }
        cp_async_commit();
    };

    if (ks0 < ks1) load_tiles(0, ks0);
    if (ST == 3 && ks0 + 1 < ks1) load_tiles(1, ks0 + 1);
    for (int ks = ks0; ks < ks1; ++ks) {
        const int buf = (ks - ks0) % ST;
        if (ST == 3) {
            // tile ks has landed when at most one younger group is still in flight; the barrier also tells that every
            // warp is done with tile ks-1, whose buffer the load of tile ks+2 reuses
            if (ks + 1 < ks1) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
            if (ks + 2 < ks1) load_tiles((ks + 2 - ks0) % ST, ks + 2);
        } else {
            if (ks + 1 < ks1) { load_tiles(buf ^ 1, ks + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
        }
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[MT], b[4], b2[4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = warp * (BM / 4) + i * 8 + (lane >> 2);
                a[i] = TRANS ? As[buf][kk + (lane & 3)][row] : As[buf][row][kk + (lane & 3)];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b[j] = Bs[buf][kk + (lane & 3)][j * 8 + (lane >> 2)];
                if (TRANS) b2[j] = B2s[buf][kk + (lane & 3)][j * 8 + (lane >> 2)];
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    if (TRANS) dmma(acc2[i][j][0], acc2[i][j][1], a[i] * a[i], b2[j]);
                }
        }
        if (ST == 2) __syncthreads();
    }
    // partial tile store
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = m0 + warp * (BM / 4) + i * 8 + (lane >> 2);
            const int col = n0 + j * 8 + 2 * (lane & 3);
            const size_t o = ((size_t)z * Mp + row) * ldb + col;
            *reinterpret_cast<double2 *>(Cpart + o) = make_double2(acc[i][j][0], acc[i][j][1]);
            if (TRANS) *reinterpret_cast<double2 *>(C2part + o) = make_double2(acc2[i][j][0], acc2[i][j][1]);
        }
}

// epilogue of the transposed product: g0 = lambda_t + gamma*Sbar_t + PTDF^T M,
// s1 = gamma + 2 kappa q_n + (PTDF.^2)^T W        (subproblems.jl:67-75)
__global__ void k_node_prep(View v, const double *Cpart, const double *C2part, int ksplit)
{
    if (!DOPF_ACTIVE(v)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // iteration prologue: reset the per-iteration counters, move maxima and work flags (nothing before this
    // kernel in the iteration touches them)
    if (i == 0) {
        v.ctrl->gen_work_cnt = 0; v.ctrl->sto_work_cnt = 0; v.ctrl->cold_work_cnt = 0; v.ctrl->pair_cnt = 0; v.ctrl->gen_grp_cnt = 0; v.ctrl->fix_node_cnt = 0; v.ctrl->sto_next = 0;
        v.ctrl->res_bits[0] = v.ctrl->res_bits[1] = v.ctrl->res_bits[2] = 0ull;
    }
    if (i < v.ldt) v.dmax[i] = 0ull;
    for (int k = i; k < v.S; k += gridDim.x * blockDim.x) v.sto_flag[k] = 0;
    for (int k = i; k < v.NS * v.N; k += gridDim.x * blockDim.x) v.fix_node_flag[k] = 0;
    for (int k = i; k < 3 * v.NS; k += gridDim.x * blockDim.x) v.sc_res_bits[k] = 0ull;
    if (i >= v.Np * v.ldt) return;
    v.dn[i] = 0ull;
    const int n = i / v.ldt, t = i % v.ldt, cur = v.ctrl->cur;
    double a = 0.0, b = 0.0;
    for (int z = 0; z < ksplit; ++z) { a += Cpart[(size_t)z * v.Np * v.ldt + i]; b += C2part[(size_t)z * v.Np * v.ldt + i]; }
    v.g0[i] = sel(v.lam, cur)[t] + v.c.gamma * sel(v.ssum, cur)[t] + a;
    const double s1v = v.c.gamma + 2.0 * v.c.kappa * v.q[n] + b;
    v.s1[i] = s1v;
    v.rg[i] = 1.0 / (v.c.prox + s1v);     // generator step: delta = -(mc + g0) * rg
    v.rg2[i] = 1.0 / (v.c.prox + 2.0 * s1v);
}

// epilogue of the flow product: line_utilization = ptdf * injection (results.jl:114)
__global__ void k_flow_reduce(View v, const double *Cpart, int ksplit, double *dst)
{
    if (!DOPF_ACTIVE(v)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.Lp * v.ldt) return;
    double a = 0.0;
    for (int z = 0; z < ksplit; ++z) a += Cpart[(size_t)z * v.Lp * v.ldt + i];
    (dst ? dst : sel(v.flow, 1 - v.ctrl->cur))[i] = a;
}

// ------------------------------------------------------------------------------------------------
// generator predict: one thread per (agent, VEC consecutive t); streaming, 16 B per unit
// ------------------------------------------------------------------------------------------------
// one unit of the generator update: anchor-linearised closed form + box projection
// (subproblems.jl:63-83 reduced; exact unless a slack hinge lies in (0, delta], which k_verify detects)
__device__ __forceinline__ double gen_unit(double pp, double mc, double pmax, double g0, double rg)
{
    const double x = pp - (mc + g0) * rg;
    return x < 0.0 ? 0.0 : (x > pmax ? pmax : x);
}

// Node-major generator update: one block per node, threadIdx.x = timestep slot (VEC timesteps),
// threadIdx.y = agent row.  g0 and the step size 1/(prox+s1) of (node, t) stay in registers while the
// block streams through the node's generators (contiguous, agents are node-sorted): per unit one
// 8-byte load, one 8-byte store and a few flops.  The largest move per (node, t) is accumulated in
// registers and published with one atomic per thread.
constexpr int GEN_UNR = 4;
template <int VEC>
__global__ void __launch_bounds__(256) k_gen_predict(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int n = blockIdx.x, sc = blockIdx.y, vn = sc * v.N + n;      // block = (node, scenario)
    const int ga = v.gen_ptr[vn], gb = v.gen_ptr[vn + 1];
    if (ga == gb) return;
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const double *__restrict__ Pc = sel(v.P, cur);
    double *__restrict__ Pn = sel(v.P, nxt);
    const int t = threadIdx.x * VEC, rows = blockDim.y;
    if (v.sc_converged[sc]) {          // frozen scenario: the iterate is carried through unchanged
        for (int g = ga + threadIdx.y; g < gb; g += rows)
#pragma unroll
            for (int w = 0; w < VEC; ++w) Pn[(size_t)g * v.T + t + w] = Pc[(size_t)g * v.T + t + w];
        return;
    }
    const size_t nt = (size_t)n * v.ldt + (size_t)sc * v.T + t;
    double g0[VEC], rg[VEC], mx[VEC];
#pragma unroll
    for (int w = 0; w < VEC; ++w) { g0[w] = v.g0[nt + w]; rg[w] = v.rg[nt + w]; mx[w] = 0.0; }
    for (int g = ga + threadIdx.y; g < gb; g += rows * GEN_UNR) {
        double pp[GEN_UNR][VEC], mc[GEN_UNR], pm[GEN_UNR];
#pragma unroll
        for (int u = 0; u < GEN_UNR; ++u) {                    // all loads first
            const int gg = min(g + u * rows, gb - 1);
            mc[u] = __ldg(v.gen_mc + gg); pm[u] = __ldg(v.gen_pmax + gg);
            const size_t o = (size_t)gg * v.T + t;
            if (VEC == 4) {
                const double2 a = *reinterpret_cast<const double2 *>(Pc + o), b = *reinterpret_cast<const double2 *>(Pc + o + 2);
                pp[u][0] = a.x; pp[u][1 % VEC] = a.y; pp[u][2 % VEC] = b.x; pp[u][3 % VEC] = b.y;
            } else if (VEC == 2) {
                const double2 a = *reinterpret_cast<const double2 *>(Pc + o);
                pp[u][0] = a.x; pp[u][1 % VEC] = a.y;
            } else pp[u][0] = Pc[o];
        }
#pragma unroll
        for (int u = 0; u < GEN_UNR; ++u) {
            const int gg = g + u * rows;
            if (gg >= gb) break;
            double pn[VEC];
#pragma unroll
            for (int w = 0; w < VEC; ++w) {
                pn[w] = gen_unit(pp[u][w], mc[u], pm[u], g0[w], rg[w]);
                mx[w] = fmax(mx[w], fabs(pn[w] - pp[u][w]));
            }
            const size_t o = (size_t)gg * v.T + t;
            if (VEC == 4) {
                *reinterpret_cast<double2 *>(Pn + o) = make_double2(pn[0], pn[1 % VEC]);
                *reinterpret_cast<double2 *>(Pn + o + 2) = make_double2(pn[2 % VEC], pn[3 % VEC]);
            } else if (VEC == 2) *reinterpret_cast<double2 *>(Pn + o) = make_double2(pn[0], pn[1 % VEC]);
            else Pn[o] = pn[0];
        }
    }
#pragma unroll
    for (int w = 0; w < VEC; ++w) note_move(v, vn, t + w, mx[w]);
}

// Agent-major variant for grids with few generators per node (a node-major block would be mostly idle): one thread
// per (generator, VEC consecutive timesteps), the node's g0 / step size are gathered (generators are node-sorted, so
// neighbouring threads read the same lines), the largest move per (node, t) is published per element.
template <int VEC>
__global__ void __launch_bounds__(256) k_gen_flat(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int per = v.T / VEC;
    const long long units = (long long)v.G * per;
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const double *__restrict__ Pc = sel(v.P, cur);
    double *__restrict__ Pn = sel(v.P, nxt);
    for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(u / per), t = (int)(u % per) * VEC;
        const int vn = v.gen_node[g], sc = v.scen_of_vn(vn);
        const size_t o = (size_t)g * v.T + t, nt = (size_t)(vn - sc * v.N) * v.ldt + (size_t)sc * v.T + t;
        double pp[VEC], pn[VEC];
        if (VEC == 4) {
            const double2 a = *reinterpret_cast<const double2 *>(Pc + o), b = *reinterpret_cast<const double2 *>(Pc + o + 2);
            pp[0] = a.x; pp[1 % VEC] = a.y; pp[2 % VEC] = b.x; pp[3 % VEC] = b.y;
        } else if (VEC == 2) {
            const double2 a = *reinterpret_cast<const double2 *>(Pc + o);
            pp[0] = a.x; pp[1 % VEC] = a.y;
        } else pp[0] = Pc[o];
        if (v.sc_converged[sc]) {
#pragma unroll
            for (int w = 0; w < VEC; ++w) pn[w] = pp[w];              // frozen scenario
        } else {
            const double mc = __ldg(v.gen_mc + g), pm = __ldg(v.gen_pmax + g);
#pragma unroll
            for (int w = 0; w < VEC; ++w) {
                pn[w] = gen_unit(pp[w], mc, pm, v.g0[nt + w], v.rg[nt + w]);
                note_move(v, vn, t + w, pn[w] - pp[w]);
            }
        }
        if (VEC == 4) {
            *reinterpret_cast<double2 *>(Pn + o) = make_double2(pn[0], pn[1 % VEC]);
            *reinterpret_cast<double2 *>(Pn + o + 2) = make_double2(pn[2 % VEC], pn[3 % VEC]);
        } else if (VEC == 2) *reinterpret_cast<double2 *>(Pn + o) = make_double2(pn[0], pn[1 % VEC]);
        else Pn[o] = pn[0];
    }
}

// ------------------------------------------------------------------------------------------------
// storages (subproblems.jl:107-207 reduced).  Segments between level-bound hits are only a few
// timesteps long, so the horizon offers little parallelism while the agents offer plenty:
//   k_sto_warm : one thread per storage, re-solve with last iteration's active set + KKT check
//   k_sto_cold : one thread per storage that failed the check: sequential planning-horizon solve
//   k_sto_fix  : one warp per storage that moved across a slack hinge: cooperative collection of
//                its hinges (coalesced over the wide row list), exact re-solve by lane 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sto_warm(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= v.S) return;
    if (v.sc_converged[v.scen_of_vn(v.sto_node[s])]) {     // frozen scenario
        const int cur = v.ctrl->cur;
        for (int t = 0; t < v.T; ++t) { const size_t o = (size_t)s * v.T + t; sel(v.D, 1 - cur)[o] = sel(v.D, cur)[o]; sel(v.C, 1 - cur)[o] = sel(v.C, cur)[o]; }
        return;
    }
    body_sto_warm(v, s);
}

// warp-parallel active-set solve (dopf_sto_warp.cuh); storages it cannot verify are queued for k_sto_cold
// DOPF_STO_WPB warps per block (measured on B200, target case: 1, 2, 4, 6, 8 and 12 warps per block run within 3 % of
// each other - the kernel is bound by its own instruction stream, not by block granularity; profiles/r2_sto_variants.log)
#ifndef DOPF_STO_WPB
#define DOPF_STO_WPB 4
#endif
#ifndef DOPF_STO_MINB
#define DOPF_STO_MINB (12 / DOPF_STO_WPB)
#endif
constexpr int sto_wpb(int J) { return J <= 3 ? DOPF_STO_WPB : 4; }
template <int J>
__global__ void __launch_bounds__(32 * sto_wpb(J), (J <= 3 ? DOPF_STO_MINB : (J == 4 ? 3 : 2))) k_sto_warp(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    extern __shared__ __align__(16) double sto_smem[];
    double *tab = sto_smem + (size_t)(threadIdx.x >> 5) * (sto_warp_smem_per_warp(v.T) / sizeof(double));
#ifndef DOPF_STO_STATIC
    // warps draw storages from a device counter (reset by k_node_prep): a storage that needs many active-set rounds only
    // holds its own warp, the others keep streaming
    (void)gw; (void)nw;
    for (;;) {
        int s = 0;
        if (lane == 0) s = atomicAdd(&v.ctrl->sto_next, 1);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= v.S) break;
#else
    for (int s = gw; s < v.S; s += nw) {
#endif
        if (v.sc_converged[v.scen_of_vn(v.sto_node[s])]) {     // frozen scenario: carry D, C through (E, eta stay)
            const int cur = v.ctrl->cur;
            for (int t = lane; t < v.T; t += 32) { const size_t o = (size_t)s * v.T + t; sel(v.D, 1 - cur)[o] = sel(v.D, cur)[o]; sel(v.C, 1 - cur)[o] = sel(v.C, cur)[o]; }
            continue;
        }
        const bool ok = (v.debug & 2) ? false : (v.T == 32 * J ? sto_warp_solve<J, false, true>(v, s, nullptr, nullptr, tab)
                                                               : sto_warp_solve<J, false, false>(v, s, nullptr, nullptr, tab));
        if (!ok && lane == 0) v.cold_work[atomicAdd(&v.ctrl->cold_work_cnt, 1)] = s;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(64) k_sto_cold(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int total = v.ctrl->cold_work_cnt;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x)
        body_sto_cold(v, v.cold_work[w], nullptr, nullptr);
    if (blockIdx.x == 0 && threadIdx.x == 0) v.ctrl->stat_sto_cold = total;
}

// hinges of node n at time t with breakpoint inside (lo,hi), collected by one warp from the wide
// row list; returns the (uncapped) count, stores at most cap entries
__device__ __forceinline__ int collect_hinges(const View &v, int n, int t, double lo, double hi, Hinge *out, int cap)
{
    const int lane = threadIdx.x & 31;
    const int cnt_in = v.wcnt[t];
    const int *lst = v.wide + (size_t)t * 2 * v.L;
    const double *lb = v.wide_b + (size_t)t * 2 * v.L;
    const double *prow = v.ptdfT + (size_t)n * v.Lp;
    int cnt = 0;
    constexpr int U = 8;                          // 256 list entries per iteration, all loads issued before use
    for (int b0 = 0; b0 < cnt_in; b0 += 32 * U) {
        int e[U]; double b[U], p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = min(b0 + u * 32 + lane, cnt_in - 1);
            e[u] = lst[j]; b[u] = lb[j];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) p[u] = prow[e[u] >> 1];
        Hinge h[U]; bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // breakpoint b/q inside (lo,hi), tested without the division (only the few accepted entries divide)
            const double q = (e[u] & 1) ? -p[u] : p[u];
            const double x0 = (q > 0.0 ? lo : hi) * q, x1 = (q > 0.0 ? hi : lo) * q;
            ok[u] = (b0 + u * 32 + lane < cnt_in) && q != 0.0 && b[u] > x0 && b[u] < x1;
            h[u].bp = 0.0; h[u].sg = 0.0;
            if (ok[u]) ok[u] = make_hinge(v.c, p[u], b[u], e[u] & 1, h[u]) && h[u].bp > lo && h[u].bp < hi;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, ok[u]);
            const int pos = cnt + __popc(m & ((1u << lane) - 1));
            if (ok[u] && pos < cap) out[pos] = h[u];
            cnt += __popc(m);
        }
    }
    return cnt;
}

// sort a hinge list of at most 64 entries by |bp| ascending (rank sort, one warp, 2 entries per lane)
__device__ __forceinline__ void sort_hinges(Hinge *lst, int n)
{
    const int lane = threadIdx.x & 31;
    if (n <= 1 || n > 64) return;
    Hinge h0 = lane < n ? lst[lane] : Hinge{1e308, 0.0}, h1 = lane + 32 < n ? lst[lane + 32] : Hinge{1e308, 0.0};
    const double k0 = fabs(h0.bp), k1 = fabs(h1.bp);
    int r0 = 0, r1 = 0;
    for (int i = 0; i < n; ++i) {
        const double ki = __shfl_sync(0xffffffffu, i < 32 ? k0 : k1, i & 31);
        r0 += (ki < k0) || (ki == k0 && i < lane);
        r1 += (ki < k1) || (ki == k1 && i < lane + 32);
    }
    __syncwarp();
    if (lane < n) lst[r0] = h0;
    if (lane + 32 < n) lst[r1] = h1;
    __syncwarp();
}

// hinge lists for the storage correction.  All storages of a node see the same hinge candidates (same PTDF column);
// only their boxes differ, and a hinge outside a storage's own box never changes state inside it.  So the lists are
// collected once per (node with a flagged storage, timestep) for the union of the flagged storages' boxes.
__device__ __forceinline__ void sto_collect_box(const View &v, int s, int t, double &lo, double &hi)
{
    // delta = (D-Db)-(C-Cb) with 0<=D,C<=pmax  =>  delta in [-Db-(pmax-Cb), (pmax-Db)+Cb]
    const double pm = v.sto_pmax[s];
    const double Db = sel(v.D, v.ctrl->cur)[(size_t)s * v.T + t], Cb = sel(v.C, v.ctrl->cur)[(size_t)s * v.T + t];
    lo = -Db - (pm - Cb); hi = (pm - Db) + Cb;
}
__device__ __forceinline__ void sto_collect_lists(const View &v, int n, int t, double lo, double hi, Hinge *list, int *cnt_out)
{
    const int lane = threadIdx.x & 31;
    int cnt = collect_hinges(v, n, t, lo, hi, list, v.hcap);
    if (cnt > v.hcap) { if (lane == 0) v.ctrl->error = DOPF_ERR_HINGE_CAP; cnt = v.hcap; }
    if (lane == 0) *cnt_out = cnt;
    __syncwarp();
    sort_hinges(list, cnt);                       // evaluations exit at the first hinge beyond |delta|
}

__global__ void __launch_bounds__(128) k_sto_collect(View v, Hinge *hinge_scratch, int *hcnt_scratch, int slots)
{
    if (!DOPF_ACTIVE(v)) return;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const long long tasks = (long long)min(v.ctrl->fix_node_cnt, slots) * v.T;
    for (long long k = gw; k < tasks; k += nw) {
        const int slot = (int)(k / v.T), t = (int)(k % v.T), n = v.fix_node_list[slot];      // n: virtual node, t: agent-local
        double lo = 0.0, hi = 0.0;                // union of the flagged storages' boxes (both contain 0)
        for (int s = v.sto_ptr[n] + lane; s < v.sto_ptr[n + 1]; s += 32) {
            if (!v.sto_flag[s]) continue;
            double l1, h1;
            sto_collect_box(v, s, t, l1, h1);
            lo = fmin(lo, l1); hi = fmax(hi, h1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        const int sc = v.scen_of_vn(n);
        sto_collect_lists(v, n - sc * v.N, sc * v.T + t, lo, hi, hinge_scratch + ((size_t)slot * v.T + t) * v.hcap, hcnt_scratch + (size_t)slot * v.T + t);
    }
}

// one warp (= one block, so the solver may use the whole register file) per affected storage
template <int J>
__global__ void __launch_bounds__(32) k_sto_fix(View v, Hinge *hinge_scratch, int *hcnt_scratch, int slots)
{
    if (!DOPF_ACTIVE(v)) return;
    extern __shared__ __align__(16) double sto_smem[];
    const int lane = threadIdx.x & 31;
    const int total = v.ctrl->sto_work_cnt;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int s = v.sto_work[w], n = v.sto_node[s];
        const int nslot = v.fix_node_slot[n];
        const int slot = nslot < slots ? nslot : slots + blockIdx.x;
        Hinge *mylist = hinge_scratch + (size_t)slot * v.T * v.hcap;
        int *mycnt = hcnt_scratch + (size_t)slot * v.T;
        if (nslot >= slots) {                     // more flagged nodes than scratch slots: own lists, own box
            for (int t = 0; t < v.T; ++t) {
                double lo, hi;
                sto_collect_box(v, s, t, lo, hi);
                const int sc = v.scen_of_vn(n);
                sto_collect_lists(v, n - sc * v.N, sc * v.T + t, lo, hi, mylist + (size_t)t * v.hcap, mycnt + t);
            }
            __syncwarp();
        }
        bool ok = false;
        if (J > 0 && !(v.debug & 1)) ok = sto_warp_solve<(J > 0 ? J : 1), true>(v, s, mylist, mycnt, sto_smem);
        if (lane == 0) {
            if (!ok) { body_sto_cold(v, s, mylist, mycnt, v.hcap <= 64); atomicAdd(&v.ctrl->stat_fix_seq, 1); }
            atomicAdd(&v.ctrl->stat_sto_fix, 1);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// verify: lanes along nodes (PTDF rows are node-contiguous), 8 timesteps per block
// ------------------------------------------------------------------------------------------------
// verify_bounds() with the b values read beside the list (one dependent load less per row) and two rows in flight
__device__ __forceinline__ bool verify_bounds_dev(const View &v, int n, int t, double &lo, double &hi)
{
    const double dnt = bits_nonneg(v.dn[(size_t)n * v.ldt + t]);
    if (dnt == 0.0) return false;
    lo = -INFINITY; hi = INFINITY;
    const int cnt = v.tcnt[t];
    const int *lst = v.tight + (size_t)t * 2 * v.L;
    const double *lb = v.tight_b + (size_t)t * 2 * v.L;
    auto one = [&](int e, double b, double p) {
        if (fabs(b) > fabs(p) * dnt * (1.0 + 1e-12)) return;        // |b/p| > dnt (without the division)
        Hinge h;
        if (!make_hinge(v.c, p, b, e & 1, h)) return;
        if (fabs(h.bp) > dnt) return;
        if (h.bp > 0.0) { if (h.bp < hi) hi = h.bp; }
        else if (h.bp < 0.0) { if (h.bp > lo) lo = h.bp; }
        else { if (h.sg > 0.0) hi = 0.0; else lo = 0.0; }
    };
    // eight rows in flight: the list entries (warp-uniform addresses) first, then the eight PTDF loads they address -
    // two dependent memory latencies per eight rows (the kernel is latency bound)
    for (int j = 0; j < cnt; j += 8) {
        int e[8]; double b[8], p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int jj = min(j + u, cnt - 1); e[u] = lst[jj]; b[u] = lb[jj]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = v.ptdf[(size_t)(e[u] >> 1) * v.Np + n];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (j + u < cnt) one(e[u], b[u], p[u]);
    }
    return !(lo == -INFINITY && hi == INFINITY);
}

__global__ void __launch_bounds__(256) k_verify(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    // phase 1 (lanes along the nodes): hinge bounds of (n,t); phase 2: the few (n,t) with a hinge in reach
    // are queued and their agents checked with the lanes along the agents
    __shared__ int qn[256], qcnt;
    __shared__ double qlo[256], qhi[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const int t = blockIdx.y * (blockDim.x >> 5) + warp;          // column
    if (threadIdx.x == 0) qcnt = 0;
    __syncthreads();
    if (n < v.N && t < v.TC) {
        double lo, hi;
        if (verify_bounds_dev(v, n, t, lo, hi)) {
            const int q = atomicAdd(&qcnt, 1);
            qn[q] = n * 8 + warp; qlo[q] = lo; qhi[q] = hi;
        }
    }
    __syncthreads();
    const int total = qcnt;
    for (int q = warp; q < total; q += (blockDim.x >> 5)) {
        const int np = qn[q] >> 3, col = blockIdx.y * (blockDim.x >> 5) + (qn[q] & 7);
        const int sc = v.scen_of_col(col), nn = sc * v.N + np, tt = col - sc * v.T;      // virtual node, agent-local timestep
        const double lo = qlo[q], hi = qhi[q];
        // the flagged generators of one (n,t) share their hinge candidates: they are appended as one group of
        // consecutive work entries (one atomic per group) so that k_gen_fix collects the hinges once per group
        for (int g0 = v.gen_ptr[nn]; g0 < v.gen_ptr[nn + 1]; g0 += 32) {
            const int g = g0 + lane;
            const bool hit = g < v.gen_ptr[nn + 1] && verify_gen_moved(v, g, tt, lo, hi);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            int base = 0;
            if (lane == 0) {
                // one round trip for both counters: low word += entries, high word += 1 group
                const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&v.ctrl->gen_work_cnt),
                                                         (unsigned long long)__popc(m) | (1ull << 32));
                base = (int)(unsigned)(old & 0xffffffffull);
                const int k = (int)(old >> 32);                       // k <= base: every group has at least one entry
                if (base + __popc(m) <= v.gen_work_cap) { v.gen_grp[2 * k] = base; v.gen_grp[2 * k + 1] = __popc(m); }
                else v.ctrl->error = DOPF_ERR_WORK_CAP;
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit && base + __popc(m) <= v.gen_work_cap) v.gen_work[base + __popc(m & ((1u << lane) - 1))] = g * v.T + tt;
        }
        for (int s = v.sto_ptr[nn] + lane; s < v.sto_ptr[nn + 1]; s += 32)
            if (verify_sto_moved(v, s, tt, lo, hi)) {
                verify_note_sto(v, s);
                if (atomicExch(&v.fix_node_flag[nn], 1) == 0) {      // the storages of a node share one set of hinge lists
                    const int k = atomicAdd(&v.ctrl->fix_node_cnt, 1);
                    v.fix_node_list[k] = nn; v.fix_node_slot[nn] = k;
                }
            }
    }
}

// Root of the generator's optimality condition f(x) = c + a*x + corr(x) on [lo,hi] WITHOUT a hinge list: every
// evaluation streams the wide row list of the timestep (whole warp).  Used when an agent's box holds more hinge
// breakpoints than the shared-memory list of k_gen_fix (few agents on a large grid).  f is increasing and piecewise
// linear: bracketed Newton steps with bisection as the safeguard; it ends as soon as the Newton step of the current
// piece stays inside the piece (then it is the exact root), like root_monotone_pl.
__device__ void stream_eval2(const View &v, int n, int t, double x, double &val, double &sL, double &sR, double &nL, double &nR)
{
    const int lane = threadIdx.x & 31;
    const int cnt_in = v.wcnt[t];
    const int *lst = v.wide + (size_t)t * 2 * v.L;
    const double *lb = v.wide_b + (size_t)t * 2 * v.L;
    const double *prow = v.ptdfT + (size_t)n * v.Lp;
    val = 0.0; sL = 0.0; sR = 0.0; nL = -1e300; nR = 1e300;
    const double tol = 1e-14 * (1.0 + fabs(x));
    for (int j = lane; j < cnt_in; j += 32) {
        const int e = lst[j];
        Hinge h;
        if (make_hinge(v.c, prow[e >> 1], lb[j], e & 1, h)) hinge_accum2(h, x, tol, val, sL, sR, nL, nR);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        val += __shfl_xor_sync(0xffffffffu, val, o); sL += __shfl_xor_sync(0xffffffffu, sL, o); sR += __shfl_xor_sync(0xffffffffu, sR, o);
        nL = fmax(nL, __shfl_xor_sync(0xffffffffu, nL, o)); nR = fmin(nR, __shfl_xor_sync(0xffffffffu, nR, o));
    }
}
__device__ double root_monotone_stream(const View &v, int n, int t, double c, double a, double lo, double hi)
{
    double val, sL, sR, nL, nR;
    stream_eval2(v, n, t, lo, val, sL, sR, nL, nR);
    if (c + a * lo + val >= 0.0) return lo;
    stream_eval2(v, n, t, hi, val, sL, sR, nL, nR);
    if (c + a * hi + val <= 0.0) return hi;
    double xl = lo, xr = hi;                       // f(xl) < 0 < f(xr)
    double x = -c / a;
    if (!(x > xl && x < xr)) x = 0.5 * (xl + xr);
    for (int it = 0; it < 200; ++it) {
        stream_eval2(v, n, t, x, val, sL, sR, nL, nR);
        const double f = c + a * x + val;
        if (f == 0.0) return x;
        double xn;
        if (f < 0.0) { xl = x; xn = x - f / (a + sR); if (xn <= nR) return xn < hi ? xn : hi; }
        else { xr = x; xn = x - f / (a + sL); if (xn >= nL) return xn > lo ? xn : lo; }
        x = (xn > xl && xn < xr) ? xn : 0.5 * (xl + xr);
        if (!(xr - xl > 1e-15 * (1.0 + fabs(xl)))) return x;
    }
    return x;
}

// exact re-solve of the generators on the work list: one warp per (agent, t)
__global__ void __launch_bounds__(128) k_gen_fix(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    // one warp per group (= the flagged generators of one (n,t), at most 32): the hinge candidates inside the
    // union of the agents' boxes are collected once, then every lane solves its own agent.  A hinge outside an
    // agent's own box never changes state inside it, so the shared list gives the same root.
    constexpr int CAP = 128;
    __shared__ Hinge lists[4][CAP];
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + wib, nw = gridDim.x * (blockDim.x >> 5);
    const int groups = v.ctrl->gen_grp_cnt;
    for (int k = gw; k < groups; k += nw) {
        const int base = v.gen_grp[2 * k], cnt_g = v.gen_grp[2 * k + 1];
        const bool mine = lane < cnt_g;
        const int e = v.gen_work[base + (mine ? lane : 0)];
        const int g = e / v.T, t = e % v.T, vn = v.gen_node[g];
        const int sc = v.scen_of_vn(vn), n = vn - sc * v.N, col = sc * v.T + t;            // physical node, column
        const double Pb = sel(v.P, cur)[(size_t)g * v.T + t], pmax = v.gen_pmax[g];
        const double lo = -Pb, hi = pmax - Pb;
        double ulo = mine ? lo : 0.0, uhi = mine ? hi : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ulo = fmin(ulo, __shfl_xor_sync(0xffffffffu, ulo, o)); uhi = fmax(uhi, __shfl_xor_sync(0xffffffffu, uhi, o)); }
        int cnt = collect_hinges(v, n, col, ulo, uhi, lists[wib], CAP);
        __syncwarp();
        const size_t nt = (size_t)n * v.ldt + col;
        if (cnt <= CAP) {
            if (mine) {
                HingeList hl; hl.h = lists[wib]; hl.n = cnt; hl.sorted = false;
                const double d = root_monotone_pl(v.gen_mc[g] + v.g0[nt], v.c.prox + v.s1[nt], hl, lo, hi);
                double Pn = Pb + d;
                Pn = Pn < 0.0 ? 0.0 : (Pn > pmax ? pmax : Pn);
                sel(v.P, nxt)[(size_t)g * v.T + t] = Pn;
                note_move(v, vn, t, Pn - Pb);
            }
        } else {
            // the union box holds more candidates than the list: one agent at a time with its own box
            for (int a = 0; a < cnt_g; ++a) {
                const double alo = __shfl_sync(0xffffffffu, lo, a), ahi = __shfl_sync(0xffffffffu, hi, a);
                __syncwarp();
                int c1 = collect_hinges(v, n, col, alo, ahi, lists[wib], CAP);
                __syncwarp();
                double dstream = 0.0;
                if (c1 > CAP)      // more breakpoints inside the agent's own box than the list holds: list-free solve
                    dstream = root_monotone_stream(v, n, col, __shfl_sync(0xffffffffu, v.gen_mc[g], a) + v.g0[nt], v.c.prox + v.s1[nt], alo, ahi);
                if (lane == a) {
                    HingeList hl; hl.h = lists[wib]; hl.n = c1; hl.sorted = false;
                    const double d = c1 > CAP ? dstream : root_monotone_pl(v.gen_mc[g] + v.g0[nt], v.c.prox + v.s1[nt], hl, lo, hi);
                    double Pn = Pb + d;
                    Pn = Pn < 0.0 ? 0.0 : (Pn > pmax ? pmax : Pn);
                    sel(v.P, nxt)[(size_t)g * v.T + t] = Pn;
                    note_move(v, vn, t, Pn - Pb);
                }
            }
        }
        if (lane == 0) atomicAdd(&v.ctrl->stat_gen_fix, cnt_g);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// aggregation ("coordinator gather", results.jl:55-106): injection and its column sums
// ------------------------------------------------------------------------------------------------
// block = 8 nodes x 32 timesteps: warp w sums the agents of node 8*blockIdx.x+w with the lanes along t (the
// agent arrays are t-contiguous); the node statistics are stored timestep-major ([ldt][Np], read along the
// nodes by k_slack_rows), so they go through a shared-memory transpose and leave as 64-byte runs.
__global__ void __launch_bounds__(256, 4) k_inject(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ double tile[8][32][9];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + w, t = blockIdx.y * 32 + lane;
    double st[8];
    const double a = inject_compute(v, n, t, st);
    sel(v.injloc, 1 - v.ctrl->cur)[(size_t)n * v.ldt + t] = a;
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[k][lane][w] = st[k];
    __syncthreads();
    const int tt = threadIdx.x >> 3, nn = threadIdx.x & 7;
    const size_t o = (size_t)(blockIdx.y * 32 + tt) * v.Np + blockIdx.x * 8 + nn;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.nst[k][o] = tile[k][tt][nn];
}

// dmax[t] = largest move of any agent at timestep t = column maximum of dn (bit patterns of
// non-negative doubles compare like integers); block (32,32): 32 timesteps, 32 row groups
__global__ void k_dmax(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ unsigned long long part[32][33];
    const int t = blockIdx.x * 32 + threadIdx.x;
    unsigned long long a = 0ull;
    for (int n = blockIdx.y * 32 + threadIdx.y; n < v.N; n += 32 * gridDim.y) { const unsigned long long b = v.dn[(size_t)n * v.ldt + t]; a = b > a ? b : a; }
    part[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0) {
        unsigned long long m = 0ull;
        for (int k = 0; k < 32; ++k) m = part[k][threadIdx.x] > m ? part[k][threadIdx.x] : m;
        if (m > v.dmax[t]) atomicMax(v.dmax + t, m);     // dmax is zeroed by k_node_prep and only grows within an iteration
    }
}

// grid (ldt/32, COLSUM_R), block (32,32): 32 columns x 32 rows in flight; block y sums its fixed group of node rows, the
// last block of a column group to arrive adds the COLSUM_R partial sums in order (deterministic, and 8x the blocks of a
// single pass - the kernel is latency bound)
__global__ void k_colsum(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ double part[32][33];
    __shared__ int last;
    const int t = blockIdx.x * 32 + threadIdx.x, nxt = 1 - v.ctrl->cur;
    const int rows = v.Np / COLSUM_R, n0 = blockIdx.y * rows;       // Np is a multiple of 64
    double a = 0.0;
    for (int n = n0 + threadIdx.y; n < n0 + rows; n += 32) a += sel(v.inj, nxt)[(size_t)n * v.ldt + t];
    part[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0) {
        double s = 0.0;
        for (int k = 0; k < 32; ++k) s += part[k][threadIdx.x];
        v.ssum_part[(size_t)blockIdx.y * v.ldt + t] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int arrived = atomicAdd(&v.colsum_cnt[blockIdx.x], 1);
        last = arrived == COLSUM_R - 1;
        if (last) v.colsum_cnt[blockIdx.x] = 0;
    }
    __syncthreads();
    if (last && threadIdx.y == 0) {
        __threadfence();
        double s = 0.0;
        for (int r = 0; r < COLSUM_R; ++r) s += __ldcg(&v.ssum_part[(size_t)r * v.ldt + t]);
        sel(v.ssum, nxt)[t] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// exact average-slack sums of the tight rows (results.jl:83-84,110-112):
//   sum over all agents i of (b_lt -+ p_{l,n(i)} * delta_it)_+ = closed form of an uncrossed row (k_dual) + rowsum[l,t,side],
//   the correction over the nodes whose agents can reach the hinge
// k_slack_rows: one block per (t, row).  Threads classify the nodes with the node statistics of the moves at (n,t):
//   nodes whose agents all keep the hinge on one side contribute in closed form; the few mixed nodes (the hinge
//   threshold falls between two movers of one sign) are ordered by node and appended to a global pair queue as ONE
//   contiguous range per row;
// k_slack_pairs: one warp per queued (row, node) pair sums the agents of the node (whole-GPU parallelism);
// k_slack_fold: every row adds the values of its range in order.
// Every partial sum is added in a fixed order, so the result does not depend on scheduling (a batch of scenarios
// reproduces the single runs bit for bit; only the position of a row's range in the queue varies).
constexpr int SLACK_BM_WORDS = 512;       // mixed nodes of a row are marked in a shared-memory bitmap (nodes beyond 16384: summed in place)
__global__ void __launch_bounds__(512) k_slack_rows(View v, unsigned char *tflag)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ double red[16];
    __shared__ unsigned bm[SLACK_BM_WORDS];
    __shared__ int bpre[SLACK_BM_WORDS], mtotal, mbase;
    const int t = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      // t: column
    const int cnt = v.tcnt[t];
    const int *lst = v.tight + (size_t)t * 2 * v.L;
    const int words = min(SLACK_BM_WORDS, (v.N + 31) >> 5);
    for (int j = blockIdx.x; j < cnt; j += gridDim.x) {
        const int l = lst[j] >> 1, side = lst[j] & 1;
        const double b = side ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
        for (int w = threadIdx.x; w < words; w += blockDim.x) bm[w] = 0u;
        __syncthreads();
        double a = 0.0;
        // Only the nodes whose largest move reaches the hinge (|b| <= |p| * max move) are looked at: all the others
        // contribute the closed form of an uncrossed row, which k_dual adds for the whole row from the flow change
        // (A*b -+ dF).  The row sum stored here is the correction: exact minus closed form over the reachable nodes.
        const size_t so = (size_t)t * v.Np;
        for (int n = threadIdx.x; n < v.N; n += blockDim.x) {
            const double p = v.ptdf[(size_t)l * v.Np + n];
            const double dnm = fmax(-v.nst[0][so + n], v.nst[1][so + n]);      // largest |move| of an agent at (n,t)
            if (fabs(b) > fabs(p) * dnm) continue;
            const double sp = side ? p : -p;
            bool ok;
            const double c = slack_node_closed(v, b, sp, n, t, ok);
            const double lin = slack_node_lin(v, b, sp, n, t);
            if (ok) { a += c - lin; continue; }
            a -= lin;                                             // the exact part of a mixed node follows from the pair queue
            if ((n >> 5) < SLACK_BM_WORDS) atomicOr(&bm[n >> 5], 1u << (n & 31));
            else a += body_slack_row_node(v, l, side, n, t);      // beyond the bitmap: summed in place (thread-local, still a fixed order)
        }
        a = Group<32>::sum(a);
        if (lane == 0) red[warp] = a;
        __syncthreads();
        if (warp == 0) {
            // exclusive prefix of the bitmap's popcounts: the position of a mixed node in the row's queue range = its rank by
            // node index, whatever order the threads found them in
            const int per = (words + 31) >> 5, w0 = lane * per, w1 = min(words, w0 + per);
            int sum = 0;
            for (int w = w0; w < w1; ++w) sum += __popc(bm[w]);
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            int run = incl - sum;
            for (int w = w0; w < w1; ++w) { bpre[w] = run; run += __popc(bm[w]); }
            if (lane == 31) {
                const int nm = incl;
                int base = nm ? atomicAdd(&v.ctrl->pair_cnt, nm) : 0;
                if (base + nm > v.pair_cap) base = -1;            // queue full: this row sums its mixed nodes itself
                mtotal = nm; mbase = base;
            }
        }
        __syncthreads();
        const int base = mbase, nm = mtotal;
        if (base >= 0 && nm)
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                unsigned bits = bm[w];
                int pos = base + bpre[w];
                while (bits) {
                    const int k = __ffs(bits) - 1; bits &= bits - 1;
                    v.pair_row[pos] = lst[j]; v.pair_node[pos] = w * 32 + k; v.pair_col[pos] = t; ++pos;
                }
            }
        if (threadIdx.x == 0) {
            double sum = 0.0;
            for (int k2 = 0; k2 < (int)(blockDim.x >> 5); ++k2) sum += red[k2];
            if (base < 0)       // (never seen in practice) serial fallback in node order
                for (int w = 0; w < words; ++w)
                    for (unsigned bits = bm[w]; bits; bits &= bits - 1) sum += body_slack_row_node(v, l, side, w * 32 + __ffs(bits) - 1, t);
            const size_t i = (size_t)l * v.ldt + t;
            (side ? v.rowsumK : v.rowsumU)[i] = sum;
            v.pbase[(size_t)t * 2 * v.L + j] = base; v.pcnt[(size_t)t * 2 * v.L + j] = base >= 0 ? nm : 0;
            atomicOr(reinterpret_cast<unsigned int *>(tflag) + (i >> 2), (unsigned)(1u << side) << (8 * (i & 3)));
        }
        __syncthreads();
    }
}

// mixed (row, node) pairs: one warp per pair, lanes over the agents of the node
__global__ void __launch_bounds__(256) k_slack_pairs(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int total = min(v.ctrl->pair_cnt, v.pair_cap);
    const int cur = v.ctrl->cur, nxt = 1 - cur;
    for (int q = gw; q < total; q += nw) {
        const int e = v.pair_row[q], l = e >> 1, side = e & 1;
        const int n = v.pair_node[q], t = v.pair_col[q];              // physical node, column
        const int sc = v.scen_of_col(t), vn = sc * v.N + n, tl = t - sc * v.T;
        const double b = side ? v.bminus[(size_t)l * v.ldt + t] : v.bplus[(size_t)l * v.ldt + t];
        const double p = v.ptdf[(size_t)l * v.Np + n], sp = side ? p : -p;
        double a = 0.0;
        for (int g = v.gen_ptr[vn] + lane; g < v.gen_ptr[vn + 1]; g += 32) {
            const size_t o = (size_t)g * v.T + tl;
            a += pospart(b + sp * (sel(v.P, nxt)[o] - sel(v.P, cur)[o]));
        }
        for (int s = v.sto_ptr[vn] + lane; s < v.sto_ptr[vn + 1]; s += 32) {
            const size_t o = (size_t)s * v.T + tl;
            a += pospart(b + sp * ((sel(v.D, nxt)[o] - sel(v.D, cur)[o]) - (sel(v.C, nxt)[o] - sel(v.C, cur)[o])));
        }
        a = Group<32>::sum(a);
        if (lane == 0) v.pair_val[q] = a;
    }
}

// every tight row adds the pair values of its queue range in a fixed order: FOLD_X * 8 warps per column, each looks at 32
// rows at a time (coalesced counts), and a row with pairs is summed by the whole warp (lane-strided partial sums, then
// the fixed shuffle tree)
constexpr int FOLD_X = 4;
__global__ void __launch_bounds__(256) k_slack_fold(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    const int t = blockIdx.y, lane = threadIdx.x & 31, wc = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int cnt = v.tcnt[t];
    for (int j0 = wc * 32; j0 < cnt; j0 += FOLD_X * 8 * 32) {
        const size_t k = (size_t)t * 2 * v.L + j0 + lane;
        const bool in = j0 + lane < cnt;
        const int n_l = in ? v.pcnt[k] : 0, base_l = in ? v.pbase[k] : 0, e_l = in ? v.tight[k] : 0;
        unsigned m = __ballot_sync(0xffffffffu, n_l > 0);
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const int n = __shfl_sync(0xffffffffu, n_l, src), base = __shfl_sync(0xffffffffu, base_l, src), e = __shfl_sync(0xffffffffu, e_l, src);
            double part = 0.0;
            for (int q = lane; q < n; q += 32) part += v.pair_val[base + q];
            part = Group<32>::sum(part);
            if (lane == 0) {
                double *dst = ((e & 1) ? v.rowsumK : v.rowsumU) + (size_t)(e >> 1) * v.ldt + t;
                *dst += part;
            }
        }
    }
}

// dual update + residual maxima (update_duals.jl, convergence.jl:3-12)
__global__ void __launch_bounds__(256) k_dual(View v, const unsigned char *tflag, const double *Cpart, int ksplit)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ double rm[8], rr[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0, b = 0.0;
    if (i < v.L * v.ldt) {
        const int l = i / v.ldt, t = i % v.ldt;
        {   // epilogue of the flow product: line_utilization = ptdf * injection (results.jl:114), split-K partials in order
            double f = 0.0;
            for (int z = 0; z < ksplit; ++z) f += Cpart[(size_t)z * v.Lp * v.ldt + i];
            if (v.flowD) f -= v.flowD[i];          // partitioned mode: the ranks' partial flows carry no demand
            sel(v.flow, 1 - v.ctrl->cur)[i] = f;
        }
        if (t < v.TC) {
            const int sc = v.scen_of_col(t);
            if (v.sc_converged[sc]) {          // frozen scenario: duals carried through, average slacks untouched
                const int cur = v.ctrl->cur;
                sel(v.mu, 1 - cur)[i] = sel(v.mu, cur)[i]; sel(v.rho, 1 - cur)[i] = sel(v.rho, cur)[i];
            } else {
                const int flag = tflag[i];
                body_dual(v, l, t, flag, a, b);
                if (v.NS > 1) {                 // per-scenario residual maxima (a warp may span scenarios)
                    if (a > 0.0) { const unsigned long long bits = nonneg_bits(a); if (bits > v.sc_res_bits[3 * sc + 1]) atomicMax(&v.sc_res_bits[3 * sc + 1], bits); }
                    if (b > 0.0) { const unsigned long long bits = nonneg_bits(b); if (bits > v.sc_res_bits[3 * sc + 2]) atomicMax(&v.sc_res_bits[3 * sc + 2], bits); }
                    a = 0.0; b = 0.0;
                }
            }
        }
    }
    if (v.NS > 1) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) { rm[threadIdx.x >> 5] = a; rr[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { a = fmax(a, rm[k]); b = fmax(b, rr[k]); }
        a = fmax(a, rm[0]); b = fmax(b, rr[0]);
        if (a > 0.0) atomicMax(&v.ctrl->res_bits[1], nonneg_bits(a));
        if (b > 0.0) atomicMax(&v.ctrl->res_bits[2], nonneg_bits(b));
    }
}

// lambda update (update_duals.jl:2-9) with its residual, then the convergence check and the buffer flip
__global__ void __launch_bounds__(256) k_lambda_finish(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ unsigned long long rb[8];
    __shared__ int ct[8], cw[8];
    unsigned long long r = 0ull;
    int nt = 0, nw = 0;
    for (int t = threadIdx.x; t < v.T; t += blockDim.x) {
        const double x = body_lambda(v, t);
        if (x > 0.0) { const unsigned long long b = nonneg_bits(x); r = b > r ? b : r; }
        nt += v.tcnt[t]; nw += v.wcnt[t];
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ro = __shfl_xor_sync(0xffffffffu, r, o);
        r = ro > r ? ro : r;
        nt += __shfl_xor_sync(0xffffffffu, nt, o); nw += __shfl_xor_sync(0xffffffffu, nw, o);
    }
    if ((threadIdx.x & 31) == 0) { rb[threadIdx.x >> 5] = r; ct[threadIdx.x >> 5] = nt; cw[threadIdx.x >> 5] = nw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { r = rb[k] > r ? rb[k] : r; nt += ct[k]; nw += cw[k]; }
        v.ctrl->res_bits[0] = r;
        v.ctrl->stat_tight_rows = nt; v.ctrl->stat_wide_rows = nw;
        body_finish(v);
    }
}

// the same for a batch of scenarios: one warp per scenario (lanes over its timesteps) updates lambda, takes the
// residual maximum and evaluates the scenario's own stop rule; the last block to finish does the global part
// (all-converged flag, buffer flip)
__global__ void __launch_bounds__(128) k_lambda_finish_batch(View v)
{
    if (!DOPF_ACTIVE(v)) return;
    __shared__ int last;
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c < v.NS) {
        const int cur = v.ctrl->cur, nxt = 1 - cur;
        int nt = 0, nw = 0;
        if (v.sc_converged[c]) {
            for (int t = lane; t < v.T; t += 32) sel(v.lam, nxt)[c * v.T + t] = sel(v.lam, cur)[c * v.T + t];
        } else {
            double r = 0.0;
            for (int t = lane; t < v.T; t += 32) { r = fmax(r, body_lambda(v, c * v.T + t)); nt += v.tcnt[c * v.T + t]; nw += v.wcnt[c * v.T + t]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o)); nt += __shfl_xor_sync(0xffffffffu, nt, o); nw += __shfl_xor_sync(0xffffffffu, nw, o); }
            if (lane == 0) {
                v.sc_res_bits[3 * c] = nonneg_bits(r);
                body_finish_scenario(v, c);
                atomicAdd(&v.ctrl->stat_tight_acc, nt); atomicAdd(&v.ctrl->stat_wide_acc, nw);
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&v.ctrl->finish_cnt, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence();
        int all = 1;
        for (int k = threadIdx.x; k < v.NS; k += blockDim.x) all &= *(volatile int *)&v.sc_converged[k];
        all = __syncthreads_and(all);
        if (threadIdx.x == 0) {
            v.ctrl->stat_tight_rows = v.ctrl->stat_tight_acc; v.ctrl->stat_wide_rows = v.ctrl->stat_wide_acc;
            v.ctrl->stat_tight_acc = 0; v.ctrl->stat_wide_acc = 0;
            body_finish_global(v, all);
        }
    }
}

// total_costs (results.jl:95-105) per scenario - on demand only: one warp per agent
__global__ void k_total_costs(View v, double *out /*[C]*/)
{
    const int newest = v.ctrl->cur;   // after k_finish the newest iterate sits in [cur]
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int a = gw; a < v.G + v.S; a += nw) {
        double x = 0.0; int vn;
        if (a < v.G) {
            vn = v.gen_node[a];
            for (int t = lane; t < v.T; t += 32) x += sel(v.P, newest)[(size_t)a * v.T + t];
            x *= v.gen_mc[a];
        } else {
            const int s = a - v.G; vn = v.sto_node[s];
            for (int t = lane; t < v.T; t += 32) x += sel(v.D, newest)[(size_t)s * v.T + t] + sel(v.C, newest)[(size_t)s * v.T + t];
            x *= v.sto_mc[s];
        }
        x = Group<32>::sum(x);
        if (lane == 0) atomicAdd(out + v.scen_of_vn(vn), x);
    }
}

// nodal price (network_elements.jl:16-25): lambda_t + sum_l (mu+rho)[l,t] ptdf[l,n]
__global__ void k_nodal_price(View v, const double *lam, const double *mu, const double *rho, double *out /*[N][T]*/)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;       // out is [C][N][T]
    if (i >= (long long)v.NS * v.N * v.T) return;
    const int t = (int)(i % v.T), n = (int)((i / v.T) % v.N), col = (int)(i / ((long long)v.T * v.N)) * v.T + t;
    double a = lam[col];
    for (int l = 0; l < v.L; ++l)
        a += (mu[(size_t)l * v.ldt + col] + rho[(size_t)l * v.ldt + col]) * v.ptdf[(size_t)l * v.Np + n];
    out[i] = a;
}

// per-unit report of the newest iterate (subproblems.jl:89-102): the unit's private slacks
//   U*[l,t] = (2w/k) (b+ - p delta)_+ ,  K*[l,t] = (2w/k) (b- + p delta)_+        
// and the values of its penalty expressions (penalty_terms.jl:3-37)
//   energy_balance[t] = (Sbar_t + delta_t)^2, upper_flow[t] = sum_l (Fbar + p delta + U - f)^2, lower_flow[t] = sum_l (K - Fbar - p delta - f)^2.
// Valid after an iteration has finished: [cur] holds the newest iterate, [1-cur] the previous one and bplus/bminus
// are still those of that iteration.  One block per timestep, threads over the lines.
__global__ void __launch_bounds__(128) k_unit_penalty(View v, int kind, int idx, double *eb, double *up, double *lo, double *U, double *K)
{
    __shared__ double ru[4], rl[4];
    const int t = blockIdx.x, newest = v.ctrl->cur, prev = 1 - newest;
    const int vn = kind == 0 ? v.gen_node[idx] : v.sto_node[idx];
    const int n = vn - v.scen_of_vn(vn) * v.N, col = v.col_of(vn, t);
    const size_t o = (size_t)idx * v.T + t;
    const double delta = kind == 0 ? sel(v.P, newest)[o] - sel(v.P, prev)[o]
                                   : (sel(v.D, newest)[o] - sel(v.D, prev)[o]) - (sel(v.C, newest)[o] - sel(v.C, prev)[o]);
    const double sc = v.c.w2 / v.c.kk;
    double au = 0.0, al = 0.0;
    for (int l = threadIdx.x; l < v.L; l += blockDim.x) {
        const size_t i = (size_t)l * v.ldt + col;
        const double p = v.ptdf[(size_t)l * v.Np + n], pd = p * delta, F = sel(v.flow, prev)[i], f = v.fmax[l];
        const double u = sc * pospart(v.bplus[i] - pd), k = sc * pospart(v.bminus[i] + pd);
        if (U) U[(size_t)l * v.T + t] = u;
        if (K) K[(size_t)l * v.T + t] = k;
        const double a = F + pd + u - f, b = k - F - pd - f;
        au += a * a; al += b * b;
    }
    au = Group<32>::sum(au); al = Group<32>::sum(al);
    if ((threadIdx.x & 31) == 0) { ru[threadIdx.x >> 5] = au; rl[threadIdx.x >> 5] = al; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double s = sel(v.ssum, prev)[col] + delta;
        eb[t] = s * s; up[t] = ru[0] + ru[1] + ru[2] + ru[3]; lo[t] = rl[0] + rl[1] + rl[2] + rl[3];
    }
}

// the same three penalty values summed over ALL units (result.penalty_term, results.jl:73-76): one thread per
// (line, t) walks the nodes and their agents (on demand only: A*L*T terms)
__global__ void __launch_bounds__(128) k_penalty_totals(View v, double *eb, double *up, double *lo)
{
    const int t = blockIdx.x * 32 + (threadIdx.x & 31), l = blockIdx.y * 4 + (threadIdx.x >> 5);     // t: column, outputs [C][T]
    if (t >= v.TC || l >= v.L) return;
    const int scn = v.scen_of_col(t), tl = t - scn * v.T;
    const int newest = v.ctrl->cur, prev = 1 - newest;
    const size_t i = (size_t)l * v.ldt + t;
    const double sc = v.c.w2 / v.c.kk, F = sel(v.flow, prev)[i], f = v.fmax[l], bp = v.bplus[i], bm = v.bminus[i], S = sel(v.ssum, prev)[t];
    double au = 0.0, al = 0.0, ae = 0.0;
    for (int n = 0; n < v.N; ++n) {
        const double p = v.ptdf[(size_t)l * v.Np + n];
        const int vn = scn * v.N + n;
        for (int g = v.gen_ptr[vn]; g < v.gen_ptr[vn + 1]; ++g) {
            const size_t o = (size_t)g * v.T + tl;
            const double d = sel(v.P, newest)[o] - sel(v.P, prev)[o], pd = p * d;
            const double a = F + pd + sc * pospart(bp - pd) - f, b = sc * pospart(bm + pd) - F - pd - f;
            au += a * a; al += b * b;
            if (l == 0) ae += (S + d) * (S + d);
        }
        for (int s = v.sto_ptr[vn]; s < v.sto_ptr[vn + 1]; ++s) {
            const size_t o = (size_t)s * v.T + tl;
            const double d = (sel(v.D, newest)[o] - sel(v.D, prev)[o]) - (sel(v.C, newest)[o] - sel(v.C, prev)[o]), pd = p * d;
            const double a = F + pd + sc * pospart(bp - pd) - f, b = sc * pospart(bm + pd) - F - pd - f;
            au += a * a; al += b * b;
            if (l == 0) ae += (S + d) * (S + d);
        }
    }
    atomicAdd(up + t, au); atomicAdd(lo + t, al);
    if (l == 0) eb[t] = ae;
}

// ------------------------------------------------------------------------------------------------
// host-side launcher of one iteration (captured into a CUDA graph by dopf_api.cu)
// ------------------------------------------------------------------------------------------------
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int set_storage_smem_attr(int T)
{
    cudaError_t e = cudaSuccess;
    // only the instantiation this horizon selects can be launched (32(J-1) < T <= 32J ... the plan picks the smallest fitting J)
#define SETA(K, W) do { const int bytes = (int)((W) * sto_warp_smem_per_warp(T)); \
        if (e == cudaSuccess && bytes > 48 * 1024 && bytes <= 227 * 1024) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); } while (0)
    SETA(k_sto_warp<1>, sto_wpb(1)); SETA(k_sto_warp<2>, sto_wpb(2)); SETA(k_sto_warp<3>, sto_wpb(3));
    SETA(k_sto_warp<4>, sto_wpb(4)); SETA(k_sto_warp<6>, sto_wpb(6)); SETA(k_sto_warp<8>, sto_wpb(8));
    SETA(k_sto_fix<1>, 1); SETA(k_sto_fix<2>, 1); SETA(k_sto_fix<3>, 1); SETA(k_sto_fix<4>, 1); SETA(k_sto_fix<6>, 1); SETA(k_sto_fix<8>, 1);
#undef SETA
    return (int)e;
}


int enqueue_iteration(const LaunchPlan &lp, cudaStream_t st, int segment)
{
    int cur_seg = 0;
    // the storage and the generator kernels of the predict and of the correction pass are independent: they run on
    // two streams unless per-kernel profiling is requested (works both under graph capture and for direct launches;
    // both fork/join pairs lie inside phase 0 of the partitioned mode).  Measured: anything forked beside the PTDF
    // products (wide-list compaction, slack sums) slows the iteration down, so those stay in stream order.
    const bool par = !lp.prof_events && lp.side_stream;
    cudaStream_t cs = st;
#define SEG_ON() (segment < 0 || cur_seg == segment)
#define FORK() do { if (par && SEG_ON()) { cudaEventRecord(lp.ev_fork, st); cudaStreamWaitEvent(lp.side_stream, lp.ev_fork, 0); cs = lp.side_stream; } } while (0)
#define MAIN() do { cs = st; } while (0)
#define JOIN() do { if (par && SEG_ON()) { cudaEventRecord(lp.ev_join, lp.side_stream); cudaStreamWaitEvent(st, lp.ev_join, 0); } cs = st; } while (0)
#define XCHG(what) do { ++cur_seg; } while (0)
    const View &v = lp.view;
    int launches = 0;
#define LAUNCH(...)                                                                                \
    do {                                                                                           \
        if (segment >= 0 && cur_seg != segment) break;                                             \
        const bool prof_ = lp.prof_events && launches < lp.prof_cap;                               \
        if (prof_) cudaEventRecord(lp.prof_events[2 * launches], cs);                              \
        __VA_ARGS__;                                                                               \
        if (prof_) { cudaEventRecord(lp.prof_events[2 * launches + 1], cs); lp.prof_names[launches] = #__VA_ARGS__; } \
        ++launches;                                                                                \
    } while (0)
    LAUNCH(k_row_prep<<<cdiv((long long)v.Lp * v.ldt, 256), 256, 0, cs>>>(v, lp.tflag));
    if (v.L <= 1024) LAUNCH(k_compact_w<<<cdiv(v.TC, 8), 256, 0, cs>>>(v, 0, 0));
    else LAUNCH(k_compact<<<v.TC, 256, 0, cs>>>(v, 0, 0));
    {   // PTDF^T M and (PTDF.^2)^T W
        dim3 grid(lp.mt_rows / lp.bm_t, v.ldt / BN, lp.ksplit_t);
        if (lp.bm_t == 64) LAUNCH(k_gemm<64, true><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, lp.part, lp.part2, v.Np, v.Lp, lp.ksplit_t, lp.mt_base));
        else LAUNCH(k_gemm<32, true><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, lp.part, lp.part2, v.Np, v.Lp, lp.ksplit_t, lp.mt_base));
        LAUNCH(k_node_prep<<<cdiv((long long)v.Np * v.ldt, 256), 256, 0, cs>>>(v, lp.part, lp.part2, lp.ksplit_t));
    }
    FORK();   // storages on the side stream ...
    if (v.S > 0) {
        // one storage per warp: the hardware block scheduler then balances the load - a straggler storage (many active-set
        // rounds in the start-up transient) only holds its own block while the others stream through the remaining slots
        #ifndef DOPF_STO_STATIC
        const int wpb = sto_wpb(lp.sto_j), wblocks = min(cdiv(v.S, wpb), lp.num_sms * (lp.sto_j <= 4 ? 12 : 8) / wpb);
#else
        const int wpb = sto_wpb(lp.sto_j), wblocks = (lp.view.debug & 64) ? min(cdiv(v.S, wpb), lp.num_sms * 16) : cdiv(v.S, wpb);
#endif
        switch (lp.sto_j) {
        case 1: LAUNCH(k_sto_warp<1><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        case 2: LAUNCH(k_sto_warp<2><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        case 3: LAUNCH(k_sto_warp<3><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        case 4: LAUNCH(k_sto_warp<4><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        case 6: LAUNCH(k_sto_warp<6><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        case 8: LAUNCH(k_sto_warp<8><<<wblocks, 32 * wpb, wpb * sto_warp_smem_per_warp(v.T), cs>>>(v)); break;
        default: LAUNCH(k_sto_warm<<<cdiv(v.S, 128), 128, 0, cs>>>(v)); break;   // long horizons: sequential warm start
        }
        LAUNCH(k_sto_cold<<<min(cdiv(v.S, 64), lp.num_sms * 8), 64, 0, cs>>>(v));
    }
    MAIN();   // ... generators on the main stream
    if (v.G > 0) {
        // one block per node: x = timestep slots of `vec` timesteps, y = agent rows
        const int vec = v.T % 4 == 0 ? 4 : (v.T % 2 == 0 ? 2 : 1), per = v.T / vec;   // per <= 1024 checked at create
        const dim3 blk(per, max(1, 256 / per)), grd(v.N, v.NS);
        if (lp.gen_flat) {
            const int fb = (int)min((long long)lp.num_sms * 16, ((long long)v.G * per + 255) / 256);
            if (vec == 4) LAUNCH(k_gen_flat<4><<<fb, 256, 0, cs>>>(v));
            else if (vec == 2) LAUNCH(k_gen_flat<2><<<fb, 256, 0, cs>>>(v));
            else LAUNCH(k_gen_flat<1><<<fb, 256, 0, cs>>>(v));
        }
        else if (vec == 4) LAUNCH(k_gen_predict<4><<<grd, blk, 0, cs>>>(v));
        else if (vec == 2) LAUNCH(k_gen_predict<2><<<grd, blk, 0, cs>>>(v));
        else LAUNCH(k_gen_predict<1><<<grd, blk, 0, cs>>>(v));
    }
    JOIN();
    if (v.L <= 1024) LAUNCH(k_compact_w<<<cdiv(v.TC, 8), 256, 0, cs>>>(v, 1, 1));
    else LAUNCH(k_compact<<<v.TC, 256, 0, cs>>>(v, 1, 1));   // with the column maxima of the moves computed in place
    {
        dim3 grid(cdiv(v.N, 32), cdiv(v.TC, 8));
        LAUNCH(k_verify<<<grid, 256, 0, cs>>>(v));
    }
    FORK();
    if (v.S > 0) {
        LAUNCH(k_sto_collect<<<lp.num_sms * 8, 128, 0, cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots));
        switch (lp.sto_j) {
        case 1: LAUNCH(k_sto_fix<1><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        case 2: LAUNCH(k_sto_fix<2><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        case 3: LAUNCH(k_sto_fix<3><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        case 4: LAUNCH(k_sto_fix<4><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        case 6: LAUNCH(k_sto_fix<6><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        case 8: LAUNCH(k_sto_fix<8><<<lp.sto_fix_blocks, 32, sto_warp_smem_per_warp(v.T), cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        default: LAUNCH(k_sto_fix<0><<<lp.sto_fix_blocks, 32, 0, cs>>>(v, lp.hinge_scratch, lp.hcnt_scratch, lp.sto_fix_slots)); break;
        }
    }
    MAIN();
    if (v.G > 0) LAUNCH(k_gen_fix<<<lp.num_sms * 8, 128, 0, cs>>>(v));
    JOIN();
    if (segment >= 0) LAUNCH(k_dmax<<<dim3(v.ldt / 32, 16), dim3(32, 32), 0, cs>>>(v));   // partitioned mode: maxima are exchanged
    XCHG(DOPF_X_DMAX);   // all ranks must build the same tight lists
    // moves may have grown: lists of the columns whose largest move changed are rebuilt (single-GPU mode: 2 = compare with
    // the maximum the first pass used; partitioned mode: the maxima were just exchanged)
    // (the list rebuild is a small latency-bound kernel, the aggregation a bandwidth-bound one, and neither reads what the
    // other writes: side by side)
    FORK();
    if (v.L <= 1024) LAUNCH(k_compact_w<<<cdiv(v.TC, 8), 256, 0, cs>>>(v, 1, segment < 0 ? 2 : 0));
    else LAUNCH(k_compact<<<v.TC, 256, 0, cs>>>(v, 1, segment < 0 ? 2 : 0));
    MAIN();
    LAUNCH(k_inject<<<dim3(v.Np / 8, v.ldt / 32), 256, 0, cs>>>(v));
    if (v.flowD) {   // partitioned mode: partial flow of the rank's own injection over its own node range (1/ranks of the product)
        dim3 grid(v.Lp / lp.bm_x, v.ldt / BN, lp.ksplit_x);
        double *dst = lp.ksplit_x == 1 ? v.xflow : lp.part;
        if (lp.bm_x == 64) LAUNCH(k_gemm<64, false><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, dst, nullptr, v.Lp, lp.mt_rows, lp.ksplit_x, 0, v.injloc[0], lp.mt_base));
        else LAUNCH(k_gemm<32, false><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, dst, nullptr, v.Lp, lp.mt_rows, lp.ksplit_x, 0, v.injloc[0], lp.mt_base));
        if (lp.ksplit_x > 1) LAUNCH(k_flow_reduce<<<cdiv((long long)v.Lp * v.ldt, 256), 256, 0, cs>>>(v, lp.part, lp.ksplit_x, v.xflow));
    }
    JOIN();
    XCHG(DOPF_X_INJ);    // nodal injection of all ranks' agents
    LAUNCH(k_slack_rows<<<dim3(lp.slack_blocks_x, v.TC), (v.N <= 256 ? 128 : 512), 0, cs>>>(v, lp.tflag));   // needs the local injection statistics only
    LAUNCH(k_slack_pairs<<<lp.num_sms * 2, 256, 0, cs>>>(v));
    LAUNCH(k_slack_fold<<<dim3(FOLD_X, v.TC), 256, 0, cs>>>(v));
    LAUNCH(k_colsum<<<dim3(v.ldt / 32, COLSUM_R), dim3(32, 32), 0, cs>>>(v));
    if (!v.flowD) {   // flow = PTDF * inj
        dim3 grid(v.Lp / lp.bm_n, v.ldt / BN, lp.ksplit_n);
        if (lp.bm_n == 64) LAUNCH(k_gemm<64, false><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0));
        else LAUNCH(k_gemm<32, false><<<grid, 128, 0, cs>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0));
    }
    XCHG(DOPF_X_ROWSUM); // exact slack sums (and partial flows) over all ranks' agents
    LAUNCH(k_dual<<<cdiv((long long)v.L * v.ldt, 256), 256, 0, cs>>>(v, lp.tflag, v.flowD ? v.xflow : lp.part, v.flowD ? 1 : lp.ksplit_n));
    if (v.NS == 1) LAUNCH(k_lambda_finish<<<1, 256, 0, cs>>>(v));
    else LAUNCH(k_lambda_finish_batch<<<cdiv(v.NS, 4), 128, 0, cs>>>(v));
#undef LAUNCH
#undef XCHG
#undef FORK
#undef SEG_ON
#undef MAIN
#undef JOIN
    if (lp.prof_count) *lp.prof_count = launches;
    return launches;
}

// profiling aid: keeps the stream busy while the first timed launch is queued (k_row_prep only depends on the
// previous iterate, so running it twice changes nothing); without it the first event pair also times the host's
// launch latency
void launch_profile_warm(const LaunchPlan &lp, cudaStream_t st)
{
    const View &v = lp.view;
    k_row_prep<<<cdiv((long long)v.Lp * v.ldt, 256), 256, 0, st>>>(v, lp.tflag);
}

void launch_total_costs(const View &v, double *d_out, cudaStream_t st)
{
    cudaMemsetAsync(d_out, 0, sizeof(double) * v.NS, st);
    k_total_costs<<<296, 256, 0, st>>>(v, d_out);
}

void launch_nodal_price(const View &v, const double *lam, const double *mu, const double *rho, double *d_out, cudaStream_t st)
{
    k_nodal_price<<<cdiv((long long)v.NS * v.N * v.T, 128), 128, 0, st>>>(v, lam, mu, rho, d_out);
}

void launch_unit_penalty(const View &v, int kind, int idx, double *eb, double *up, double *lo, double *U, double *K, cudaStream_t st)
{
    k_unit_penalty<<<v.T, 128, 0, st>>>(v, kind, idx, eb, up, lo, U, K);
}

void launch_penalty_totals(const View &v, double *eb, double *up, double *lo, cudaStream_t st)
{
    cudaMemsetAsync(up, 0, sizeof(double) * v.TC, st); cudaMemsetAsync(lo, 0, sizeof(double) * v.TC, st);
    k_penalty_totals<<<dim3(cdiv(v.TC, 32), cdiv(v.L, 4)), 128, 0, st>>>(v, eb, up, lo);
}

// partitioned mode: the all-reduced injection (fixed exchange buffer) becomes the injection of the new iterate
__global__ void k_copy_inj(View v, const double *src)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < (size_t)v.Np * v.ldt) sel(v.inj, 1 - v.ctrl->cur)[i] = src[i] - v.demand[i];     // no rank carries the demand in its local injection
}
void launch_copy_inj(const View &v, const double *src, cudaStream_t st)
{
    k_copy_inj<<<cdiv((long long)v.Np * v.ldt, 256), 256, 0, st>>>(v, src);
}

// scenario batches: host layout [C][rows][T] <-> device layout [rows][ld] (scenario c in the columns c*T .. c*T+T-1)
__global__ void k_pack_cols(double *dev, double *host_layout, int rows, int C, int T, int ld, int to_device)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)C * rows * T) return;
    const int t = (int)(i % T), r = (int)((i / T) % rows), c = (int)(i / ((long long)T * rows));
    double *d = dev + (size_t)r * ld + (size_t)c * T + t;
    if (to_device) *d = host_layout[i]; else host_layout[i] = *d;
}
void launch_pack_cols(double *dev, double *host_layout, int rows, int C, int T, int ld, int to_device, cudaStream_t st)
{
    k_pack_cols<<<cdiv((long long)C * rows * T, 256), 256, 0, st>>>(dev, host_layout, rows, C, T, ld, to_device);
}

// static wide-row bound mwide[l] = max_n |ptdf[l,n]| * rbox[n]; one warp per line (re-run after the
// node box ranges have been max-reduced over the ranks)
__global__ void k_mwide(View v)
{
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= v.L) return;
    double m = 0.0;
    for (int n = lane; n < v.N; n += 32) m = fmax(m, fabs(v.ptdf[(size_t)l * v.Np + n]) * v.rbox[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) v.mwide[l] = m;
}
void launch_mwide(const View &v, cudaStream_t st) { k_mwide<<<cdiv((long long)v.L * 32, 128), 128, 0, st>>>(v); }

// partitioned mode set-up: dst = PTDF * demand (the constant part of every flow)
void launch_flow_of_demand(const LaunchPlan &lp, double *dst, cudaStream_t st)
{
    const View &v = lp.view;
    dim3 grid(v.Lp / lp.bm_n, v.ldt / BN, lp.ksplit_n);
    if (lp.bm_n == 64) k_gemm<64, false><<<grid, 128, 0, st>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0, v.demand, 0);
    else k_gemm<32, false><<<grid, 128, 0, st>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0, v.demand, 0);
    k_flow_reduce<<<cdiv((long long)v.Lp * v.ldt, 256), 256, 0, st>>>(v, lp.part, lp.ksplit_n, dst);
}

// levels of the staged iterate and buffer flip, used when a state is injected from the host
__global__ void k_levels(View v)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= v.S) return;
    const int nxt = 1 - v.ctrl->cur;
    double e = 0.0;
    for (int t = 0; t < v.T; ++t) {
        const size_t o = (size_t)s * v.T + t;
        e += sel(v.C, nxt)[o] - sel(v.D, nxt)[o];
        v.E[o] = e;
    }
}
__global__ void k_flip(View v) { v.ctrl->cur = 1 - v.ctrl->cur; }

// derived quantities (injection, column sums, flows, levels) of the iterate staged in the
// inactive buffers, then flip: the staged iterate becomes the "previous iterate"
void launch_rebuild_derived(const LaunchPlan &lp, cudaStream_t st, int segment)
{
    const View &v = lp.view;
    if (segment <= 0) k_inject<<<dim3(v.Np / 8, v.ldt / 32), 256, 0, st>>>(v);
    if (segment == 0) return;
    k_colsum<<<dim3(v.ldt / 32, COLSUM_R), dim3(32, 32), 0, st>>>(v);
    dim3 grid(v.Lp / lp.bm_n, v.ldt / BN, lp.ksplit_n);
    if (lp.bm_n == 64) k_gemm<64, false><<<grid, 128, 0, st>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0);
    else k_gemm<32, false><<<grid, 128, 0, st>>>(v, v.ptdf, v.Np, v.ldt, lp.part, nullptr, v.Lp, v.Np, lp.ksplit_n, 0);
    k_flow_reduce<<<cdiv((long long)v.Lp * v.ldt, 256), 256, 0, st>>>(v, lp.part, lp.ksplit_n, nullptr);
    if (v.S > 0) k_levels<<<cdiv(v.S, 128), 128, 0, st>>>(v);
    k_flip<<<1, 1, 0, st>>>(v);
}


}  // namespace dopf
