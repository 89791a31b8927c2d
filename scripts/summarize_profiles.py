"""Turn the raw outputs of scripts/collect_evidence.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def short(name):
    n = name.replace("void ", "").replace("gpis::", "")
    return n.split("(")[0]

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_launches.csv"))) if r and r[0].isdigit()]
shutil.copy(os.path.join(G, f"{R}_launches.csv"), os.path.join(P, f"{R}_launches.csv"))
tot = collections.OrderedDict()
for r in rows:
    k = short(r[4]); v = float(r[-1]); u = r[-2]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += ms
allms = sum(v[1] for v in tot.values())
with open(os.path.join(P, f"{R}_launch_summary.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --grid 128 --steps 2 --warmup 1 --no-cpu-baseline\n")
    f.write(f"# {len(rows)} launches, {allms:.1f} ms of GPU time (cold-cache, serialised: shares, not absolutes)\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:28s} launches {n:5d}  total {ms:10.3f} ms  share {100 * ms / allms:5.1f} %  avg {ms / n:9.4f} ms\n")
print(open(os.path.join(P, f"{R}_launch_summary.txt")).read())

# ---- DRAM traffic of the evaluation launches of one 256^3 step
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_traffic.csv"))) if r and r[0].isdigit()]
shutil.copy(os.path.join(G, f"{R}_traffic.csv"), os.path.join(P, f"{R}_traffic.csv"))
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r[0]), {"kernel": short(r[4]), "grid": r[8]})
    d[r[-3]] = float(r[-1]) * (1e9 if r[-2] == "Gbyte" else 1e6 if r[-2] == "Mbyte" else 1e3 if r[-2] == "Kbyte" else 1.0) if "bytes" in r[-3] else float(r[-1])
ids = sorted(per)
half = len(ids)   # -c 16 = the 4 chunks x 2 passes x 2 batch classes of the first (device-resident) 256^3 step
step = [per[i] for i in ids[:half]]
dram = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in step)
bench = json.load(open(os.path.join(G, f"{R}_bench.json")))
tj = {"grid": bench["config"]["grid"], "frames": bench["config"]["frames"], "launches_per_step": half,
      "dram_bytes_per_step": dram, "dram_read_bytes_per_step": sum(d["dram__bytes_read.sum"] for d in step),
      "source": f"profiles/{R}_traffic.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, -k regex:k_eval_v3, python bench.py --steps 1 --warmup 0)",
      "launches": [{"kernel": d["kernel"], "grid": d["grid"], "dram_bytes": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]} for d in step]}
json.dump(tj, open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)
print("dram bytes per step: %.2f GB over %d launches (algorithmic %.2f GB)" % (dram / 1e9, half, bench["roofline"]["algorithmic_bytes_per_launch"] / 1e9))

# ---- full captures: the metrics quoted in profiles/README.md
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled", "sm__sass_inst_executed_op_shared"]
for kern in ("k_eval_v3", "k_leaf_train", "other"):
    rep = os.path.join(G, f"{R}_{kern}.ncu-rep")
    if not os.path.exists(rep): continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    with open(os.path.join(P, f"{R}_{kern}_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kern}; selected raw metrics\n")
        for r in rr[2:]:
            f.write("kernel: " + r[hdr.index("Kernel Name")][:70] + "\n")
            for i, h in enumerate(hdr):
                if any(h.startswith(w) for w in want):
                    f.write(f"  {h:86s} {r[i]} {units[i]}\n")
    print("wrote", f"{R}_{kern}_ncu_full.txt")
for fn in (f"{R}_bench.json", f"{R}_bench_reference.json", f"{R}_hbm_regime.json", f"{R}_update_profile.txt", f"{R}_train_bench.txt"):
    if os.path.exists(os.path.join(G, fn)):
        shutil.copy(os.path.join(G, fn), os.path.join(P, fn))
