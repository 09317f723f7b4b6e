"""Turn the ncu artefacts a GPU visit left in gpurun_out/ (scripts/gpu_ncu.sh) into the tracked summaries under profiles/."""
import csv, collections, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(REPO)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

for name, rep, cmd in (("gemm_tc", "gpurun_out/prof_gemm.ncu-rep", "-k regex:gemm_tc -s 60 -c 6"), ("gemm_tc_panel", "gpurun_out/prof_xf.ncu-rep", "--kernel-name-base demangled -k regex:\"gemm_tc_kernel<64, false, true>\" -s 4 -c 4"), ("attn_v2", "gpurun_out/prof_attn.ncu-rep", "-k regex:attn_v2 -s 0 -c 6")):
    if not os.path.exists(rep):
        continue
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# {tag} `ncu --set full` capture of {name}_kernel", "",
           f"`ncu --set full --clock-control none --import-source on {cmd} python scripts/ncu_target.py 1` (cfg2 shape: B=8, T=1024, S=256).",
           "Cold-cache, serialised replays: read SHARES and pipe percentages, not absolute times.  The raw .ncu-rep is scratch (gpurun_out/).", ""]
    for d in data:
        out += ["## " + d[idx["Kernel Name"]] + "  grid " + d[idx["Grid Size"]] + " block " + d[idx["Block Size"]], "", "| metric | value |", "|---|---|"]
        out += [f"| {k} | {d[idx[k]]} {units[idx[k]]} |" for k in KEYS if k in idx]
        out.append("")
    open(f"profiles/{tag}_{name}_ncu.md", "w").write("\n".join(out))

rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if r]
for i, r in enumerate(rows):
    if r[0] == "ID":
        hdr, data = r, rows[i + 1:]
        break
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    try:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ns2vc::", "").replace("ns2vc::", "").replace("<unnamed>::", "")
        met = r[idx["Metric Name"]]; val = float(r[idx["Metric Value"]].replace(",", "")); unit = r[idx["Metric Unit"]]
    except Exception:
        continue
    a = agg[name]
    if met == "gpu__time_duration.sum":
        a[0] += 1; a[1] += val / (1e3 if unit in ("ns", "nsecond") else 1.0)
    elif met == "dram__bytes_read.sum":
        a[2] += val * MUL.get(unit, 1)
    elif met == "dram__bytes_write.sum":
        a[3] += val * MUL.get(unit, 1)
tot = sum(v[1] for v in agg.values())
out = [f"# {tag} ncu launch list (scripts/ncu_target.py 2: load+pack, prepare_cond, 2 UNet forwards + 2 sampler steps, cfg2 shape)", "",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500`.",
       "Cold-cache, serialised: compare SHARES; DRAM bytes are per-launch averages under ncu's cache flush (upper bound of the warm-L2 traffic).", "",
       "| kernel | launches | total us | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
traffic = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.1%} | {v[2] / max(v[0], 1) / 1e6:.2f} | {v[3] / max(v[0], 1) / 1e6:.2f} |")
    traffic[k] = {"launches": v[0], "dram_bytes_per_launch": (v[2] + v[3]) / max(v[0], 1)}
open(f"profiles/{tag}_ncu_launches.md", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("\n".join(out[5:18]))
