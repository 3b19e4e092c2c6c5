#!/usr/bin/env python
"""bench.py - ADMM iteration throughput of libdopf on B200 (one "step" = one ADMM iteration).

    python bench.py --gpus N --steps K --warmup W [--workload target|cfg3|cfg2|cfg4] [--impl reference]

metric  : agent*timestep updates per second = (G+S)*T*iterations / time,
          whole-job aggregate over all ranks; `iters_per_s` is reported beside it.
workload: "target" = the north_star's synthetic 100k-agent x 96-period case (80k generators + 20k storages on the
          2000-node / 3000-line grid of BASELINE configs[2]); inputs resident in HBM; working set (~0.5 GB) exceeds the
          126 MB L2, so no explicit L2 flush is needed.  It is the case the north_star's target sentence names and the
          largest single-GPU case of the list.  "cfg2" = BASELINE configs[1] (118 nodes, 1k generators + 200 storages,
          24 periods), "cfg3" = configs[2] (2k nodes, 20k + 5k, 96 periods), "cfg4" = configs[3] (128 independent
          118-node / 24-period scenarios per GPU as ONE batched device problem; 1024 over 8 GPUs, no collective).
timing  : W warm-up iterations from the reference's cold start (all-zero iterate and duals), then exactly K timed
          iterations, CUDA events on the library's stream, max over ranks.  The timed window therefore lies in the
          start-up transient of the ADMM run; `steady_state` repeats the measurement at iterations 151.. of the same run.
params  : gamma = 0.03/A, flow_weight = 1/A (the reference's ratio 10/0.3, damped scale; see PARAMS); `alt_params` holds
          the same window for the literals divided by A and for the round-1 parameters.
N > 1   : one process per GPU.  --shard agents (default): ONE case with N x the agents of the workload on the same grid
          (weak; --scaling strong: the workload's agents split over the ranks), agents partitioned over the ranks,
          network/dual part replicated; per iteration torch.distributed (NCCL) all-reduces the per-timestep move maxima,
          the nodal injection and the exact slack row sums between the phases of libdopf; the whole
          iteration incl. the collectives is one CUDA graph.  --path partitioned runs N = 1 through the same code path.
          cfg4 / --shard scenarios: every rank runs independent scenarios, no data-path collective.
--impl reference : the CPU oracle (oracle/, OpenMP on all host cores) on a bounded sample of the same workload with
          the full case's parameters; the Julia/JuMP/Gurobi reference itself cannot be installed here (no Julia).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, L, G, S, T)
    "target": (2000, 3000, 80000, 20000, 96),
    "cfg3": (2000, 3000, 20000, 5000, 96),
    "cfg2": (118, 186, 1000, 200, 24),
    "cfg4": (118, 186, 1000, 200, 24),      # BASELINE configs[3]: SCENARIOS_PER_GPU independent scenarios of the cfg2 shape per GPU (1024 over 8)
}
SCENARIOS_PER_GPU = 128
SAMPLE_AGENTS = {"target": (200, 50), "cfg3": (200, 50), "cfg2": (1000, 200), "cfg4": (1000, 200)}   # CPU-side bounded sample
METRIC = "agent_timestep_updates_per_s"
UNIT = "agent*timestep/s"


def make_case(pkg, workload, seed, agents=None, params="ref_ratio", params_agents=None):
    """synthetic case of a workload; `agents` overrides (G, S) (bounded CPU samples, N x agents for weak scaling);
    the ADMM parameters are derived from `params_agents` (default: the agents of the case itself)"""
    N, L, G, S, T = WORKLOADS[workload]
    if agents is not None:
        G, S = agents
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed)
    return pkg.Problem.from_arrays(d), params_for(params_agents or (G + S), params)


def measure_dgemm_peak(torch, L, N, T, reps=8):
    """fp64 GEMM throughput of cuBLAS (torch.matmul) on THIS box: the two products of the iteration in their exact
    shapes and a square 4096^3 DGEMM; the best of the three is the denominator of the tensor-bound roofline
    (there is no fp64 figure in MEASURED_PEAKS.json).  CUDA events, best of `reps` after a warm-up."""
    out = {}
    shapes = {"flow_LxN_NxT": (L, N, T), "ptdfT_NxL_LxT": (N, L, T), "square_4096": (4096, 4096, 4096)}
    for name, (m, k, n) in shapes.items():
        a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda")
        c = torch.empty(m, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a, b, out=c)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = 2.0 * m * k * n / (best * 1e-3) / 1e12
        del a, b, c
    return out


# ADMM parameters of the synthetic workloads.  The reference's literals (gamma 0.3, flow weight 10, three_node with 5
# agents) do not carry over to 10^3..10^5 agents: every agent reacts to the whole imbalance / overload (Jacobi update),
# so the loop gain grows with the number of agents A.  Measured on B200 (scripts/param_scan.py,
# profiles/r2_param_scan.md): (0.3/A, 10/A) and (0.1/A, 3.33/A) end in a bang-bang limit cycle (dual residuals constant
# at 14 / 4.7, ~10 % of all generator timesteps need the hinge correction every iteration), (0.03/A, 1/A) is damped.
# Default = the reference's RATIO flow_weight/gamma = 10/0.3 at the largest damped scale; prox weight is the literal.
PARAMS = {
    "ref_ratio": (0.03, 1.0),          # gamma*A, flow_weight*A : w/gamma = 33.3 (the reference's), damped  [default]
    "ref_literal_over_A": (0.3, 10.0),  # the literals divided by A: w/gamma = 33.3, limit cycle
    "round1": (0.3, 1.0),              # round-1 bench parameters: w/gamma = 3.3
}


def params_for(A, name="ref_ratio"):
    gs, ws = PARAMS[name]
    return dict(gamma=gs / A, flow_weight=ws / A, prox_weight=1.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons; started before the warm-up, samples are time-stamped and only
    those inside the timed region are reported (nearest ones if the region is shorter than the period)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_for_samples(self, n=2, timeout=3.0):
        t0 = time.time()
        while len(self.rows) < n and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin, t_end):
        time.sleep(0.05)
        if self.proc:
            self.proc.terminate()
        good = [(ts, r) for ts, r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        inside = [r for ts, r in good if t_begin <= ts <= t_end + 0.03]
        note = "samples inside the timed region"
        if not inside and good:
            inside = [r for ts, r in sorted(good, key=lambda x: abs(x[0] - 0.5 * (t_begin + t_end)))[:3]]
            note = "timed region shorter than the sampling period: nearest samples"
        sm = [float(r[0]) for r in inside]; mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "note": note}


def oracle_rate(pkg, workload, steps, warmup, budget_s=150.0):
    """agent*timestep updates/s of the CPU oracle on a bounded sample of the workload: the sample's
    agent count is scaled (from a one-iteration probe) so that warmup+steps fit `budget_s`."""
    from oracle import oracle
    oracle.set_num_threads(len(os.sched_getaffinity(0)))      # all host cores, whatever OMP_NUM_THREADS the launcher exported
    g_s, s_s = SAMPLE_AGENTS[workload]
    full_agents = WORKLOADS[workload][2] + WORKLOADS[workload][3]       # the sample runs with the FULL case's gamma and flow weight
    prob, cfg = make_case(pkg, workload, seed=0, agents=(max(g_s // 5, 8), max(s_s // 5, 2)), params_agents=full_agents)
    ora = oracle.OracleADMM(prob, cfg["gamma"], flow_weight=cfg["flow_weight"], prox_weight=cfg["prox_weight"])
    ora.iterate(0)
    t0 = time.perf_counter(); ora.iterate(0); probe = time.perf_counter() - t0
    scale = budget_s / max(steps + warmup, 1) / max(probe, 1e-6)      # affordable multiple of the probe size
    frac = min(1.0, max(0.2, scale / 5.0))
    prob, cfg = make_case(pkg, workload, seed=0, agents=(max(int(g_s * frac), 8), max(int(s_s * frac), 2)), params_agents=full_agents)
    ora = oracle.OracleADMM(prob, cfg["gamma"], flow_weight=cfg["flow_weight"], prox_weight=cfg["prox_weight"])
    for _ in range(warmup):
        ora.iterate(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.iterate(0)
    dt = time.perf_counter() - t0
    A = prob.G + prob.S
    sample = (f"{steps} iteration(s) of the oracle on {prob.G} generators + {prob.S} storages of the '{workload}' grid "
              f"(N={prob.N}, L={prob.L}, T={prob.T}); exact per-agent solves, OpenMP")
    return A * prob.T * steps / dt, dt / steps * 1e3, oracle.num_threads(), sample


def reduced_cpu_rate(pkg, workload, params="ref_ratio"):
    """agent*timestep updates/s of a SEQUENTIAL CPU run of the same reduced algebra on the WHOLE workload: the library's
    per-element device code (csrc/dopf_math.h, dopf_bodies.h) compiled for the host and run as plain loops (tests/host_emul:
    test infrastructure that mirrors enqueue_iteration; dense products as triple loops, one thread).  One warm-up iteration
    from the cold start, then one timed iteration.  Not the reference and not the oracle - a second, stated CPU baseline on
    the same config beside `cpu_baseline`."""
    from tests.host_emul import emul
    from dopf_b200 import multi
    prob, cfg = make_case(pkg, workload, 0, params=params)
    sub, _, _ = multi.shard_problem(prob, 0, 1)             # node-sorted agents
    e = emul.EmulADMM(sub, gamma=cfg["gamma"], flow_weight=cfg["flow_weight"], prox_weight=cfg["prox_weight"], hcap=64)
    t0 = time.perf_counter(); e.iterate(); first = time.perf_counter() - t0
    which, dt = "iteration 1 from the cold start", first
    if first < 30.0:
        t0 = time.perf_counter(); e.iterate(); dt = time.perf_counter() - t0
        which = "iteration 2 from the cold start (after one warm-up iteration)"
    if int(e.status[6]) != 0:
        raise RuntimeError("host emulation reported a capacity error")
    A = prob.G + prob.S
    return {"value": A * prob.T / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"the whole '{workload}' case ({prob.G} generators + {prob.S} storages, N={prob.N}, L={prob.L}, T={prob.T}), {which}; "
                      "the library's per-element device code compiled for the host, sequential (tests/host_emul)",
            "ms_per_iteration": dt * 1e3, "same_config": True,
            "gen_corrected": int(e.status[2]), "sto_corrected": int(e.status[3]), "tight_rows": int(e.status[4])}


def run_reference(args):
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = g.load_package()
    value, ms, cores, sample = oracle_rate(pkg, args.workload, args.steps, args.warmup)
    N, L, G, S, T = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "nodes": N, "lines": L, "generators": G, "storages": S, "timesteps": T,
                       "sampled": True, "same_config": False, "sample": sample,
                       "note": "reference arm = CPU oracle port (exact per-agent solves, O(L) hinges per agent*timestep, dense QP per storage) on a bounded SAMPLE of the "
                               "workload's agents with the full case's gamma / flow weight; the rate per agent*timestep is what is reported - "
                               "Julia/JuMP/Gurobi are not installable offline"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def build_problem(pkg, args, rank, world):
    """(problem, ADMM parameters, per-rank agent*timestep units, description) of this rank"""
    N, L, G, S, T = WORKLOADS[args.workload]
    if args.workload == "cfg4":
        C = SCENARIOS_PER_GPU
        d = pkg.cases.synthetic_scenarios(N=N, L=L, G=G, S=S, T=T, n_scen=C, seed=0, first_scenario=rank * C)     # one grid, every rank its own 128 scenarios
        return pkg.Problem.from_arrays(d), params_for(G + S, args.params), C * (G + S) * T, f"{world} x {C} independent scenarios, no collective"
    partitioned = world > 1 and args.shard == "agents"
    if partitioned or args.path == "partitioned":
        mult = world if args.scaling == "weak" else 1
        prob, cfg = make_case(pkg, args.workload, seed=0, agents=(mult * G, mult * S), params=args.params)   # same case on every rank
        return prob, cfg, mult * (G + S) * T // world, (f"{world} GPUs, agents of ONE case partitioned ({mult * (G + S)} agents on one grid, {args.scaling} scaling), "
                                                         "NCCL all-reduce of move maxima / nodal injection / slack sums per iteration, iteration + collectives in one CUDA graph")
    prob, cfg = make_case(pkg, args.workload, seed=rank, params=args.params)
    return prob, cfg, (G + S) * T, f"{world} independent case(s), one per GPU, no collective"


def run_dopf(args):
    import torch
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libdopf has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = g.load_package()
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = WORKLOADS[args.workload]
    prob, cfg, units_per_rank, parallelism = build_problem(pkg, args, rank, world)
    partitioned = args.workload != "cfg4" and ((world > 1 and args.shard == "agents") or args.path == "partitioned")
    libcomm = None
    if partitioned and args.comm == "library":
        # collectives inside libdopf (dopf_comm_init): the ordinary dopf_step drives the partitioned iteration
        from dopf_b200 import multi
        libcomm = multi.LibraryCommADMM(prob, rank, world, local, hinge_capacity=64, use_graph=not args.no_graph, **cfg)
        part, dev = None, libcomm.dev
    elif partitioned:
        from dopf_b200 import multi
        part = multi.PartitionedADMM(prob, rank, world, local, hinge_capacity=64, graph=not args.no_graph, **cfg)
        dev = part.dev
    else:
        part = None
        dev = DeviceADMM(prob, device=local, hinge_capacity=64, **cfg)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(steps):
        """device time of `steps` iterations on the stream the library runs on, max over ranks"""
        barrier()
        if part is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(part.stream)
            st_ = part.step(steps, check_every=steps)
            e1.record(part.stream)
            part.stream.synchronize()
            ms_ = e0.elapsed_time(e1)
        else:
            st_ = dev.step(steps)              # CUDA events on the library's stream around the graph replays
            ms_ = st_.last_step_ms
        barrier()
        tm_ = torch.tensor([ms_], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm_, op=dist.ReduceOp.MAX)
        return float(tm_.item()), st_

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    (part or dev).step(args.warmup)
    if clocks:
        clocks.wait_for_samples()
    t_begin = time.time()
    ms_total, st = timed(args.steps)
    clk = clocks.stop(t_begin, time.time()) if clocks else None
    value = world * units_per_rank * args.steps / (ms_total * 1e-3)
    window = dict(iterations=f"{args.warmup + 1}..{args.warmup + args.steps} from the reference's cold start (all-zero iterate and duals)",
                  gen_corrected_total=st.gen_corrected, sto_corrected_total=st.sto_corrected, sto_cold_last=st.sto_cold,
                  tight_rows=st.tight_rows, wide_rows=st.wide_rows, residuals=[st.res_lambda, st.res_mue, st.res_rho])

    # ---- per-kernel device times of the NEXT iteration (CUDA event pair per launch) -> roofline ----
    prof = dev.profile_iteration() if part is None and libcomm is None else []
    kern = {}
    for name, ms in prof:
        kern[name] = kern.get(name, 0.0) + ms
    dom = max(kern, key=kern.get) if kern else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    C = SCENARIOS_PER_GPU if args.workload == "cfg4" else 1
    Gr, Sr = dev.prob.G * C, dev.prob.S * C       # agents on this rank
    # algorithmic bytes per launch
    def alg_bytes_of(name):
        if name.startswith("k_gen_predict") or name.startswith("k_gen_flat"): return 16.0 * Gr * T
        if name.startswith("k_sto_warp") or name.startswith("k_sto_warm"): return 40.0 * Sr * T
        if name.startswith("k_inject"): return 8.0 * (Gr + 2 * Sr) * T + 16.0 * N * T * C
        return None

    def gemm_flops_of(name):
        if name.startswith("k_gemm") and "true" in name: return 4.0 * L * N * T * C      # PTDF^T M and (PTDF.^2)^T W
        if name.startswith("k_gemm"): return 2.0 * L * N * T * C
        return None
    dgemm = measure_dgemm_peak(torch, L, N, T * C) if rank == 0 and part is None and libcomm is None else {}
    fp64_peak = max(dgemm.values()) if dgemm else 37.0
    if dom is None:
        roof = None
    elif gemm_flops_of(dom):
        ach = gemm_flops_of(dom) / (kern[dom] * 1e-3) / 1e12
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None,
                "peak_source": "cuBLAS DGEMM measured in this run (best of the two product shapes and a square 4096^3): %s" % json.dumps({k_: round(v_, 2) for k_, v_ in dgemm.items()})}
    else:
        b = alg_bytes_of(dom) or 40.0 * Sr * T
        ach = b / (kern[dom] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": b}
    if roof is not None:
        # DRAM bytes (read + write) per launch of that kernel from the committed ncu --set full capture of this workload
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
            roof["traffic"] = tr.get(args.workload, {}).get(dom)
        except Exception:
            pass
        roof["kernel_ms"] = kern[dom]
        roof["kernel_share_of_iteration"] = kern[dom] / sum(kern.values())
        roof["profiled_iteration"] = args.warmup + args.steps + 1
        # whole iteration against the HBM roofline: SURVEY.md 8(d) bytes per iteration / the timed ms per step
        b_iter = C * (16.0 * dev.prob.G * T + 40.0 * dev.prob.S * T + 32.0 * N * T + 80.0 * L * T) + 16.0 * L * N
        roof["whole_iteration"] = {"algorithmic_bytes": b_iter, "achieved_GBps": b_iter / (ms_total / args.steps * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": b_iter / (ms_total / args.steps * 1e-3) / 1e9 / hbm}
        gp = next((v for k_, v in kern.items() if k_.startswith("k_gen_predict") or k_.startswith("k_gen_flat")), None)
        if gp:
            roof["generator_stream"] = {"kernel_ms": gp, "achieved_GBps": 16.0 * Gr * T / (gp * 1e-3) / 1e9, "frac": 16.0 * Gr * T / (gp * 1e-3) / 1e9 / hbm}
        roof["dgemm_peaks_TFLOPs"] = {k_: round(v_, 2) for k_, v_ in dgemm.items()}
        roof["gemm_TFLOPs"] = {k_: round(gemm_flops_of(k_) / (v_ * 1e-3) / 1e12, 2) for k_, v_ in kern.items() if gemm_flops_of(k_)}

    # ---- steady state: the same measurement after the transient (iterations 151.. of this run) ----
    steady = None
    if not args.quick:
        done = args.warmup + args.steps + (1 if part is None and libcomm is None else 0)
        if done < 150:
            (part or dev).step(150 - done)
        ms_s, st_s = timed(args.steps)
        steady = {"ms_per_step": ms_s / args.steps, "value": world * units_per_rank * args.steps / (ms_s * 1e-3), "iterations": f"151..{150 + args.steps}",
                  "tight_rows": st_s.tight_rows, "residuals": [st_s.res_lambda, st_s.res_mue, st_s.res_rho]}

    # ---- end to end through the C ABI with host buffers (pinned): upload state, iterate, read back ----
    e2e_steps = max(1, min(args.steps, 10))
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    it = dev.get_iterate(("P", "D", "C", "E", "avgU", "avgK"))
    lam, mu, rho = dev.get_duals(0)
    hb = {k: pin(v.shape) for k, v in it.items()}
    for k, v in it.items():
        hb[k][...] = v
    hl, hm, hr = pin(lam.shape), pin(mu.shape), pin(rho.shape)
    hl[...] = lam; hm[...] = mu; hr[...] = rho
    h2d = sum(hb[k].nbytes for k in ("P", "D", "C", "avgU", "avgK")) + hl.nbytes + hm.nbytes + hr.nbytes
    d2h = sum(hb[k].nbytes for k in ("P", "D", "C", "E")) + hl.nbytes + hm.nbytes + hr.nbytes
    import ctypes as C_
    iteration = dev.status.iteration
    barrier()
    t0 = time.perf_counter()
    if part is not None or libcomm is not None:
        h2d = 0
    for _ in range(e2e_steps):
        if part is not None or libcomm is not None:
            (part or dev).step(1)              # state stays on the device; results are read back every step
        else:
            dev.set_state(iteration, P=hb["P"], D=hb["D"], C_=hb["C"], avgU=hb["avgU"], avgK=hb["avgK"], lam=hl, mu=hm, rho=hr)
            dev.step(1)
        dev.get_iterate(("P", "D", "C", "E"), out=hb)
        dev.lib.dopf_get_duals(dev.h, 0, hl.ctypes.data_as(C_.c_void_p), hm.ctypes.data_as(C_.c_void_p), hr.ctypes.data_as(C_.c_void_p))
        iteration = dev.status.iteration
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_rank * e2e_steps / float(tm.item())

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling if partitioned else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "nodes": N, "lines": L, "generators": G, "storages": S, "timesteps": T,
                           "scenarios_per_gpu": C, "agents_per_gpu": Gr + Sr,
                           "gamma": cfg["gamma"], "flow_weight": cfg["flow_weight"], "prox_weight": cfg["prox_weight"], "params": args.params,
                           "params_note": "gamma = %g/A, flow_weight = %g/A (A agents per case): the reference's ratio flow_weight/gamma = 10/0.3 at the largest damped scale; "
                                          "the literals divided by A end in a limit cycle (see alt_params and profiles/r2_param_scan.md)" % PARAMS[args.params],
                           "parallelism": parallelism, "l2": "working set larger than L2 (no flush)"},
                "iters_per_s": args.steps / (ms_total * 1e-3),
                "gpu_launches": dev.status.launches_per_iteration * args.steps,
                "clocks": clk, "roofline": roof,
                "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "what": ("partitioned run: iterate + dopf_get_iterate/duals(host) every step (no per-step upload)" if (part is not None or libcomm is not None)
                                                     else "dopf_set_state(host) + dopf_step(1) + dopf_get_iterate/duals(host), pinned buffers")},
                "timed_window": window, "steady_state": steady}
        if part is not None:
            line["partitioned_graph"] = {"captured": part.graph is not None, "error": part.graph_error}
        if libcomm is not None:
            line["collectives"] = "library-owned NCCL communicator (dopf_comm_init), iteration + ncclAllReduce captured in libdopf's CUDA graph"
        if world == 1 and not args.quick and part is None and libcomm is None and args.workload != "cfg4":
            # the same window (5 warm-up + 20 timed iterations from the cold start) with the other parameter sets
            alts = []
            for name in PARAMS:
                if name == args.params:
                    continue
                ad = DeviceADMM(prob, device=local, hinge_capacity=64, **params_for(G + S, name))
                try:
                    ad.step(5); g0, s0 = ad.status.gen_corrected, ad.status.sto_corrected
                    sa = ad.step(20)
                    alts.append({"params": name, "gamma_x_A": PARAMS[name][0], "flow_weight_x_A": PARAMS[name][1], "ms_per_step": sa.last_step_ms / 20,
                                 "gen_corrected_per_iter": (sa.gen_corrected - g0) / 20, "sto_corrected_per_iter": (sa.sto_corrected - s0) / 20,
                                 "tight_rows": sa.tight_rows, "residuals": [sa.res_lambda, sa.res_mue, sa.res_rho]})
                except Exception as e:
                    alts.append({"params": name, "error": str(e)})
                ad.close()
            line["alt_params"] = alts
        if world == 1 and part is None and libcomm is None:
            # time to tolerance on the reference's own case (three_node, gamma 0.3, literal weights, eps 1e-3: stops at
            # iteration 476 like src/opf_admm_decentral.jl) - SURVEY.md 8(d) "time-to-tolerance"
            try:
                tn = DeviceADMM(pkg.Problem.from_structs(*pkg.cases.three_node()), gamma=0.3, device=local)
                t0 = time.perf_counter(); st3 = tn.step(2000); wall = time.perf_counter() - t0
                line["three_node_to_tolerance"] = {"iterations": int(st3.iteration), "converged": bool(st3.converged),
                                                   "device_ms": st3.last_step_ms, "wall_ms": wall * 1e3}
                tn.close()
            except Exception as e:      # additional information only
                line["three_node_to_tolerance"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline and not args.quick and args.workload != "cfg4":
            v, ms, cores, sample = oracle_rate(pkg, args.workload, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_iteration_of_sample": ms}
            try:        # additional information: the same reduced algebra, sequential, on the whole case
                line["cpu_reduced_algebra_baseline"] = reduced_cpu_rate(pkg, args.workload, args.params)
            except Exception as e:
                line["cpu_reduced_algebra_baseline"] = {"error": str(e)[:300]}
        print(json.dumps(line), flush=True)
    # teardown order matters with a captured NCCL graph: drop the graph first, then the handle, then leave without the
    # process-group destructor (destroying the communicator under a live graph blocks at interpreter exit)
    if part is not None:
        part.close()
    else:
        dev.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="dopf", choices=["dopf", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="agents", choices=["agents", "scenarios"], help="N>1: partition the agents of one case (NCCL exchanges) or run independent scenarios")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="agent partition: N x the workload's agents (weak) or the workload's agents split over the ranks (strong)")
    ap.add_argument("--path", default="auto", choices=["auto", "partitioned"], help="partitioned: also N=1 runs through the phase-wise multi-GPU code path (like-for-like scaling baseline)")
    ap.add_argument("--params", default="ref_ratio", choices=sorted(PARAMS), help="ADMM parameter set (see PARAMS)")
    ap.add_argument("--comm", default="torch", choices=["torch", "library"], help="partitioned path: all-reduces by torch.distributed between dopf_step_phase calls, or by libdopf itself (dopf_comm_init + dopf_step)")
    ap.add_argument("--no-graph", action="store_true", help="partitioned path without the CUDA graph (eager launches from Python)")
    ap.add_argument("--quick", action="store_true", help="skip the steady-state / alt-params / cpu-baseline legs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_dopf(args)


if __name__ == "__main__":
    main()
