"""Turn the ncu artefacts of scripts/gpu_pre_ncu.sh (gpurun_out/pre_launches.csv, prof_pre.ncu-rep) into profiles/<tag>_pre_model_ncu.md."""
import collections
import csv
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(REPO)
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def short(name):
    name = name.replace("void ns2vc::", "").replace("ns2vc::", "").replace("<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")
    for cut in ("(const", "(ns2vc", "(GemmOp", "(AttnOp", "(LinOp", "(SplitBuf"):
        name = name.split(cut)[0]
    return name.strip()


rows = [r for r in csv.reader(open("gpurun_out/pre_launches.csv")) if r]
for i, r in enumerate(rows):
    if r[0] == "ID":
        hdr, data = r, rows[i + 1:]
        break
idx = {h: i for i, h in enumerate(hdr)}
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
seq = []
for r in data:
    try:
        name = short(r[idx["Kernel Name"]]); met = r[idx["Metric Name"]]; val = float(r[idx["Metric Value"]].replace(",", "")); unit = r[idx["Metric Unit"]]
    except Exception:
        continue
    a = agg[name]
    if met == "gpu__time_duration.sum":
        us = val / (1e3 if unit in ("ns", "nsecond") else 1.0)
        a[0] += 1; a[1] += us; seq.append((name, us, r[idx["Grid Size"]]))
    elif met == "dram__bytes_read.sum":
        a[2] += val * MUL.get(unit, 1)
    elif met == "dram__bytes_write.sum":
        a[3] += val * MUL.get(unit, 1)
tot = sum(v[1] for v in agg.values())
out = [f"# {tag} ncu evidence for the condition encoders (`Pre_model.infer`, B=8, T=1024, S=256; scripts/gpu_pre_ncu.sh)", "",
       "## Launch list of ONE infer", "",
       "`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python scripts/ncu_pre_target.py`",
       "(cold-cache, serialised replays: compare SHARES; the in-stream time of the same call is `bench.py` -> `pre_model.ms_per_call`).", "",
       f"{sum(v[0] for v in agg.values())} kernels, {tot:.0f} us summed.", "",
       "| kernel | launches | total us | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.1%} | {v[2] / max(v[0], 1) / 1e6:.2f} | {v[3] / max(v[0], 1) / 1e6:.2f} |")
# one phoneme-encoder layer, in launch order
big = [i for i, (n, us, g) in enumerate(seq) if us > 100]
if big:
    i0 = max(0, big[0] - 5)
    out += ["", "One PhoneEncoder layer in launch order (T = 1024: 8192 rows):", "", "| kernel | us | grid |", "|---|---:|---|"]
    out += [f"| `{n}` | {us:.1f} | {g} |" for n, us, g in seq[i0:i0 + 6]]
rep = "gpurun_out/prof_pre.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h2, units, d2 = rr[0], rr[1], rr[2:]
    ix = {h: i for i, h in enumerate(h2)}
    KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    out += ["", "## `ncu --set full` capture of the ENC instantiation of the GEMM (`gemm_tc_kernel<64, 0, 0, 1>`: ReLU / row-mask epilogues)", "",
            "`ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_tc_kernel<(int)64, (bool)0, (bool)0, (bool)1> -s .. -c ..`", ""]
    for d in d2:
        out += ["### " + short(d[ix["Kernel Name"]]) + "  grid " + d[ix["Grid Size"]] + " block " + d[ix["Block Size"]], "", "| metric | value |", "|---|---|"]
        out += [f"| {k} | {d[ix[k]]} {units[ix[k]]} |" for k in KEYS if k in ix]
        out.append("")
    # details page of the first captured launch (the PhoneEncoder's conv-FFN GEMM)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    dr = list(csv.reader(det.splitlines()))
    if dr:
        dx = {h: i for i, h in enumerate(dr[0])}
        WANT = ("Duration", "Compute (SM) Throughput", "L2 Cache Throughput", "DRAM Throughput", "L2 Hit Rate", "Issue Slots Busy", "Executed Ipc Active",
                "SM Busy", "Mem Pipes Busy", "No Eligible", "Warp Cycles Per Issued Instruction", "Achieved Occupancy")
        seen = set()
        out += ["### details page of the first captured launch (the PhoneEncoder's conv-FFN GEMM: 8192 rows x 1024 columns, K = 8 taps x 256)", "",
                "| section | metric | value |", "|---|---|---|"]
        for r in dr[1:]:
            if r[dx["ID"]] != "0" or r[dx["Metric Name"]] not in WANT or r[dx["Metric Name"]] in seen:
                continue
            seen.add(r[dx["Metric Name"]])
            out.append(f"| {r[dx['Section Name']]} | {r[dx['Metric Name']]} | {r[dx['Metric Value']]} {r[dx['Metric Unit']]} |")
        out += ["", "Reading: 7 tiles x 32 k-blocks per CTA in 122 us = ~1070 cycles per k-block against 456 cycles of MMAs: a 128 x 64 tile moves 48 KB of",
                "operands per k-block from L2 (97.6 % hits) = ~45 B/clk per SM, the per-SM L2 -> SM rate the denoiser's deep-K launches also sit at",
                "(DESIGN.md section 5); neither the L2 as a whole (38 %) nor the tensor pipe (41-44 % active) is saturated.", ""]
open(f"profiles/{tag}_pre_model_ncu.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
