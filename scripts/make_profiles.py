"""Regenerate the committed round summaries under profiles/ from the scratch artefacts in gpurun_out/:
     python scripts/make_profiles.py r2 [bench.json launches.csv top.ncu-rep]
   inputs : gpurun_out/<bench.json> (plain bench line), gpurun_out/<launches.csv> (ncu launch list of the same command),
            gpurun_out/<top.ncu-rep> (ncu --set full capture of the top kernels)
   outputs: profiles/<round>_bench_target_n1.json, <round>_ncu_launches_bench.csv, <round>_kernel_shares.md,
            <round>_ncu_top_kernels.md, <round>_ncu_traffic.json"""
import csv, json, os, shutil, subprocess, sys
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
f_bench, f_launch, f_rep = (sys.argv[2:5] + ["bench.json", "launches_bench.csv", f"prof_{rnd}_top.ncu-rep"][len(sys.argv[2:5]):])
line = json.loads([l for l in open(os.path.join(go, f_bench)) if l.startswith("{")][-1])
json.dump(line, open(os.path.join(pr, f"{rnd}_bench_target_n1.json"), "w"), indent=1)
rows = [r for r in csv.reader(l for l in open(os.path.join(go, f_launch)) if l.startswith('"'))]
with open(os.path.join(pr, f"{rnd}_ncu_launches_bench.csv"), "w") as f:
    for l in open(os.path.join(go, f_launch)):
        if l.startswith('"'): f.write(l)
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = {}
for r in rows[1:]:
    k = r[ik].split("(")[0].replace("void ", "").replace("dopf::", "")
    agg.setdefault(k, []).append(float(r[iv]) / 1e3)
lib = {k: v for k, v in agg.items() if not k.startswith("k_")}       # cuBLAS / torch kernels of the DGEMM-peak measurement leg (outside the timed region)
agg = {k: v for k, v in agg.items() if k.startswith("k_")}
tot = sum(sum(v) for v in agg.values())
km = line["kernels_ms"]; ks = sum(km.values())
W = line["config"]
with open(os.path.join(pr, f"{rnd}_kernel_shares.md"), "w") as f:
    f.write(f"# Round {rnd[1:]} - kernel shares of one ADMM iteration (B200, workload '{W['workload']}': {W['nodes']} nodes / {W['lines']} lines / "
            f"{W['generators']} generators / {W['storages']} storages / {W['timesteps']} periods)\n\n")
    f.write(f"## CUDA-event times, no profiler (`bench.py --steps {line['steps']} --warmup {line['warmup']}` -> `dopf_profile_iteration`, iteration "
            f"{line['roofline'].get('profiled_iteration', '?')} = the one right after the timed window; kernels launched\n"
            "several times per iteration are summed)\n\n| kernel | ms | share |\n|---|---|---|\n")
    for k, v in sorted(km.items(), key=lambda kv: -kv[1]):
        f.write(f"| {k} | {v:.4f} | {100 * v / ks:.1f} % |\n")
    f.write(f"\nsum of kernels {ks:.3f} ms (serial, each launch bracketed by events); graph replay with the fork/join of the independent groups "
            f"(storages | generators in the predict and in the correction pass, list rebuild | aggregation): {line['ms_per_step']:.3f} ms per iteration over the {line['steps']} timed iterations.\n\n")
    f.write("## ncu launch list of the same bench command (`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline`)\n\n"
            "Cold-cache and serialised; covers the set-up, the 5 warm-up and the 20 timed iterations (the cold-start transient): compare SHARES.\n\n```\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:34s} n={len(v):3d} avg {sum(v) / len(v):9.1f} us  max {max(v):9.1f}  share {100 * sum(v) / tot:5.1f}%\n")
    f.write("```\n")
    if lib:
        f.write("\nLibrary kernels in the same capture (not part of the iteration: the cuBLAS DGEMM peak measurement `measure_dgemm_peak` and torch's random fill, after the timed region):\n\n```\n")
        for k, v in sorted(lib.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:90]:90s} n={len(v):3d} avg {sum(v) / len(v):9.1f} us\n")
        f.write("```\n")
rep = os.path.join(go, f_rep)
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines())); hh = r[0]; units = r[1]
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
            'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
    G, S, T, L, N = W["generators"], W["storages"], W["timesteps"], W["lines"], W["nodes"]
    traffic = {}
    with open(os.path.join(pr, f"{rnd}_ncu_top_kernels.md"), "w") as f:
        f.write(f"# Round {rnd[1:]} - ncu summary of the top kernels (B200, workload '{W['workload']}')\n\n"
                "Captured with `ncu --set full --clock-control none --import-source on -k regex:\"k_sto_warp|k_gen_predict|k_gemm\"` on\n"
                "`python scripts/prof_case.py 2000 3000 80000 20000 96 2 7` (iteration 8 from the cold start, i.e. inside the window the driver times; after the same command had exited 0 without ncu); read with\n"
                "`ncu -i ... --page raw --csv`.  Times under ncu are cold-cache and serialised; the bench line holds the CUDA-event times measured\n"
                "without a profiler.\n")
        for row in r[2:]:
            name = row[hh.index("Kernel Name")].split("(")[0].replace("void ", "").replace("dopf::", "")
            d = {k: (v, u) for k, v, u in zip(hh, row, units)}
            f.write(f"\n## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in keys:
                if k in d: f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            def val(k):
                v, u = d[k]; v = float(v.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            tr = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
            us = float(d['gpu__time_duration.sum'][0].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(d['gpu__time_duration.sum'][1], 1.0)
            bname = name.replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true").replace("<32, 0>", "<32, false>").replace("<32, 1>", "<32, true>")
            traffic[bname] = tr
            if name.startswith("k_gemm"):
                fl = (4.0 if ("1>" in name or "true" in name) else 2.0) * L * N * T
                pk = line["roofline"].get("dgemm_peaks_TFLOPs", {})
                f.write(f"\nalgorithmic flops per launch {fl / 1e9:.2f} GFLOP; DRAM traffic {tr / 1e6:.1f} MB (the 48 MB PTDF is read once, the rest stays in L2); "
                        f"{fl / us / 1e6:.1f} TFLOP/s fp64 under ncu; cuBLAS DGEMM measured in the bench run on the same box: {json.dumps(pk)} TFLOP/s.\n")
            else:
                b = 16.0 * G * T if name.startswith("k_gen_predict") else 40.0 * S * T
                f.write(f"\nalgorithmic bytes per launch {b / 1e6:.1f} MB ({'16 B per generator*timestep' if name.startswith('k_gen') else '40 B per storage*timestep'}); "
                        f"DRAM traffic (read+write) {tr / 1e6:.1f} MB; algorithmic throughput under ncu {b / us / 1e3:.1f} GB/s = "
                        f"{100 * b / us / 1e3 / 6541.1:.1f} % of the measured 6541 GB/s copy peak.\n")
    json.dump({W["workload"]: traffic, "source": f"gpurun_out/{f_rep} (dram__bytes_read.sum + dram__bytes_write.sum per launch; ncu --set full capture of iteration 8 of the target case)"},
              open(os.path.join(pr, f"{rnd}_ncu_traffic.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(pr)))
