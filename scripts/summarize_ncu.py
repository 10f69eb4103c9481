"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag> gpurun_out/launches_f1.csv gpurun_out/prof_f1.ncu-rep
"""
import collections, csv, json, os, subprocess, sys

tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs("profiles", exist_ok=True)
out = [f"# ncu summary `{tag}`", ""]

rows = list(csv.reader(open(launches_csv))) if launches_csv != "-" else [["ID", "Kernel Name", "Metric Value"]]
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mi:
        continue
    try:
        v = float(r[mi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
if agg:
  out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)",
        "", f"command: see scripts/profile_hjb.sh; {sum(a[0] for a in agg.values())} launches, {tot/1e6:.1f} ms total", "",
        "| share | launches | avg us | kernel |", "|---:|---:|---:|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {t/tot*100:.2f} % | {n} | {t/n/1e3:.1f} | `{k}` |")
out.append("")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active"]
stalls = [n for n in h if "issue_stalled" in n and n.endswith("per_issue_active.ratio")]
out += ["## Full capture (`ncu --set full --clock-control none --import-source on`)", ""]
traffic = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]]
    out += [f"### `{name}`", "", "| metric | value | unit |", "|---|---:|---|"]
    for w in want:
        if w in idx:
            out.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    st = sorted(((float(r[idx[s]]), s) for s in stalls), reverse=True)[:8]
    out += ["", "top warp stall reasons (warps per issue-active cycle): " +
            ", ".join(f"{s.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')} {v:.2f}" for v, s in st), ""]
    def tobytes(v, u):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tb = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
         tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    traffic.setdefault(name, []).append(tb)
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
tj = "profiles/traffic.json"
cur = json.load(open(tj)) if os.path.exists(tj) else {}
cur[tag] = {k: sum(v) / len(v) for k, v in traffic.items()}
json.dump(cur, open(tj, "w"), indent=1)
print("\n".join(out))
