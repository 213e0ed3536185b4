#!/usr/bin/env python
"""Turns the raw outputs of tools/measure_round.sh (gpurun_out/) into the tracked summaries under profiles/.
   python tools/summarize_profiles.py r2"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
for src, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("bench_ref.json", f"{tag}_bench_reference_arm.json"),
                 ("launches.csv", f"{tag}_bench_launches.csv"), ("frows.json", f"{tag}_frows.json"),
                 ("robust.jsonl", f"{tag}_robustness.jsonl"), ("sanitizer.log", f"{tag}_sanitizer.txt"),
                 ("bench_n2.json", f"{tag}_bench_n2.json"), ("bench_n4.json", f"{tag}_bench_n4.json"), ("bench_n8.json", f"{tag}_bench_n8.json")):
    if os.path.exists(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
if os.path.exists(os.path.join(G, "launches.csv")):
    rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4][:70], []).append(float(r[14]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f"{tag}_bench_launch_shares.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c2 --no-c5\n"
                "(per-launch times are cold-cache and serialised: compare shares, not absolutes)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:70s} n={len(v):4d} avg={sum(v)/len(v):8.1f} us share={sum(v)/tot:.3f}\n")
    print(open(os.path.join(P, f"{tag}_bench_launch_shares.txt")).read())
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"] + [
        f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio" for k in
        ("barrier", "short_scoreboard", "long_scoreboard", "wait", "not_selected", "mio_throttle", "math_pipe_throttle", "branch_resolving")]
out, traffic = [], {}
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
for name, key in (("sweep_plain", "C3"), ("sweep_weighted", "C3_weighted")):
    rep = os.path.join(G, f"{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines())); h, u, v = rr[0], rr[1], rr[2]
    idx = {x: i for i, x in enumerate(h)}
    have = [w for w in want if w in idx]
    if not out:
        out.append(["Kernel Name"] + have); out.append([""] + [u[idx[w]] for w in have])
    out.append([v[idx["Kernel Name"]]] + [v[idx[w]] for w in have])
    traffic[key] = sum(float(v[idx[m]]) * scale[u[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    # SASS mnemonics of the profiled kernel (the ones that prove TMA bulk copies, mbarriers and f64 REDs)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for r in list(csv.reader(src.splitlines()))[2:]:
        if len(r) > 1 and r[1].strip():
            t = r[1].strip().split()
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.rstrip(";")] += 1
    with open(os.path.join(P, f"{tag}_sass_mnemonics_{name}.txt"), "w") as f:
        f.write(f"static SASS instruction mix of {v[idx['Kernel Name']]} (ncu source page of gpurun_out/{name}.ncu-rep)\n")
        for op, n in ops.most_common():
            f.write(f"{op:28s} {n}\n")
if out:
    csv.writer(open(os.path.join(P, f"{tag}_tiled_sweep_ncu_full_summary.csv"), "w")).writerows(out)
    traffic["source"] = (f"profiles/{tag}_tiled_sweep_ncu_full_summary.csv (ncu --set full --clock-control none, one steady-state launch of "
                         "em_sweep_tiled on C3, plain and bootstrap-weighted): dram__bytes_read.sum + dram__bytes_write.sum")
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"))
    for r in zip(*out):
        print(r)
