"""Summarise an .ncu-rep (ncu --set full) into a small markdown + json file under profiles/.

    python tools/ncu_summary.py gpurun_out/r1_prof4.ncu-rep profiles/r01_ncu_contract [cells]

`cells`: algorithmic cells of the profiled launch (stored so that bench.py can tell whether
its own launch has the profiled shape before quoting the executed-instruction counts).
"""
import collections
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h
         and h.endswith("_per_issue_active.ratio")]


def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


summ = []
for d in data:
    rec = {"kernel": d[idx["Kernel Name"]]}
    for k in KEYS:
        if k in idx:
            rec[k] = [d[idx[k]], units[idx[k]]]
    # executed fp64 thread-instructions (roofline section of --set full): DADD + DMUL + DFMA
    fp64 = 0.0
    for h in hdr:
        if "sass_thread_inst_executed_op_d" in h and h.endswith("_pred_on.sum") and \
                any(op in h for op in ("op_dadd", "op_dmul", "op_dfma")):
            try:
                v = float(d[idx[h]])
            except ValueError:
                continue
            rec[h] = [d[idx[h]], units[idx[h]]]
            fp64 += v
    rec["fp64_thread_insts"] = fp64 or None
    rec["dram_bytes_per_launch"] = (
        to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"]))
    st = sorted([(float(d[idx[h]]), h) for h in stall if d[idx[h]] not in ("", "n/a")],
                reverse=True)[:6]
    rec["top_stalls_cycles_per_issue"] = {
        h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):
        round(v, 2) for v, h in st}
    summ.append(rec)
# opcode mix of the first kernel from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
mix = collections.Counter()
if hi:
    h = srows[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(srows)
    ix, isrc = h.index("Instructions Executed"), h.index("Source")
    for r in srows[hi[0] + 1:end]:
        if len(r) > ix and r[ix].isdigit():
            s = r[isrc].strip().split()
            op = s[1] if s[0].startswith("@") else s[0]
            mix[op] += int(r[ix])
tot = sum(mix.values()) or 1
# executed fp64 arithmetic of the first launch from the per-instruction counts of the source
# page (warp-level counts of DFMA / DADD / DMUL x 32 lanes; an upper bound where lanes are
# predicated off) -- used when this ncu has no sass_thread_inst_executed_op_d* metrics
fp64_warp = sum(v for k, v in mix.items() if k.split(".")[0] in ("DFMA", "DADD", "DMUL"))
for rec in summ:
    if not rec.get("fp64_thread_insts") and fp64_warp:
        rec["fp64_thread_insts"] = 32.0 * fp64_warp
        rec["fp64_thread_insts_source"] = "source page of the first launch: 32 x (DFMA + DADD + DMUL)"
cells = float(sys.argv[3]) if len(sys.argv) > 3 else None  # algorithmic cells per launch
json.dump({"report": rep, "launches": summ, "cells_per_launch": cells,
           "opcode_mix_first_launch": {k: v for k, v in mix.most_common(16)}},
          open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write("# ncu --set full summary of `%s`\n\n" % rep)
    for rec in summ:
        f.write("## %s\n\n| metric | value |\n|---|---|\n" % rec["kernel"][:80])
        for k in KEYS:
            if k in rec:
                f.write("| %s | %s %s |\n" % (k, rec[k][0], rec[k][1]))
        f.write("| dram bytes per launch (read + write) | %.0f |\n" % rec["dram_bytes_per_launch"])
        f.write("| top stall reasons (cycles per issue) | %s |\n\n" % rec["top_stalls_cycles_per_issue"])
    f.write("## executed warp-instruction mix (first launch)\n\n| opcode | count | share |\n|---|---|---|\n")
    for k, v in mix.most_common(16):
        f.write("| %s | %d | %.1f%% |\n" % (k, v, 100.0 * v / tot))
print(open(out + ".md").read()[:1500])
