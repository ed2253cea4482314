"""Turns the raw outputs of tools/evidence_n1.sh (gpurun_out/ev_*) into the committed evidence
under profiles/ (round 2): bench JSON lines, launch-list summaries, one ncu summary per hot
kernel (tools/ncu_summary.py), the GPU test log.

    python tools/evidence_collect.py
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def bench_lines():
    for fn in sorted(os.listdir(G)):
        if fn.startswith(("ev_bench_", "ev8_bench_")) and fn.endswith(".json"):
            try:
                d = last_json(os.path.join(G, fn))
            except Exception as e:  # noqa: BLE001
                print("skip", fn, e)
                continue
            tag = fn.split("_bench_", 1)[1][:-5]
            n = d.get("n_gpus", 1)
            if tag.endswith("_n%d" % n):
                tag = tag[:-len("_n%d" % n)]
            out = "r02_bench_%s_n%d.json" % (tag, n)
            with open(os.path.join(P, out), "w") as f:
                json.dump(d, f, indent=1)
            print(out, d.get("ms_per_step"), d.get("value"))


def launch_summary(csv_name, out_name, title):
    path = os.path.join(G, csv_name)
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        name = r[ik].split("(")[0][:64]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(a[1] for a in agg.values()) or 1.0
    shutil.copy(path, os.path.join(P, out_name + ".csv"))
    with open(os.path.join(P, out_name + ".md"), "w") as f:
        f.write("# %s\n\nCommand: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 "
                "--csv` (raw list: %s.csv).  Per-launch times are cold-cache and serialised: compare "
                "shares, not absolutes.  The `FillFunctor<unsigned char>` launches are bench.py's L2 "
                "flush, not part of the step.\n\n| kernel | launches | total us | avg us | share |\n"
                "|---|---|---|---|---|\n" % (title, out_name))
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n"
                    % (name, n, t * scale, t * scale / n, 100 * t / tot))
    print(out_name, len(agg), "kernels")


def ncu(rep, out, cells):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        print("missing", rep)
        return
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path,
                    os.path.join(P, out), str(cells)], stdout=subprocess.DEVNULL)
    d = json.load(open(os.path.join(P, out + ".json")))
    rec = d["launches"][-1]
    print(out, rec["kernel"][:40], rec.get("gpu__time_duration.sum"), "grid",
          rec.get("launch__grid_size"), "fp64 pipe",
          rec.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    bench_lines()
    launch_summary("ev_launches_c3.csv", "r02_launches_c3",
                   "Round 2: ncu launch list of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e`")
    launch_summary("ev_launches_c4.csv", "r02_launches_c4",
                   "Round 2: ncu launch list of `python bench.py --config C4 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e`")
    # C3, one half-step of 128 walkers: cells = walkers x rows x intervals
    ncu("ev_c3_contract_kernel.ncu-rep", "ncu_c3_contract_kernel", 128 * 192 * 369)
    ncu("ev_c3_synchrotron_fused_kernel.ncu-rep", "ncu_c3_synchrotron_fused_kernel", 128 * 64 * 569)
    ncu("ev_c3_walker_prep_kernel.ncu-rep", "ncu_c3_walker_prep_kernel", 128 * 370)
    ncu("ev_c3_combine_lnprob_kernel.ncu-rep", "ncu_c3_combine_lnprob_kernel", 128 * 64)
    # C4: 128 walkers x (100 x 869) rows x 99 seed intervals
    ncu("ev_c4_ssc_inner_wt8.ncu-rep", "ncu_c4_ssc_inner_kernel", 128 * 100 * 869 * 99)
    ncu("ev_c4_ssc_rest.ncu-rep", "r02_ncu_c4_ssc_outer_seed", 128 * 100 * 868)
    if os.path.exists(os.path.join(G, "ev8_pytest_multi.log")):
        shutil.copy(os.path.join(G, "ev8_pytest_multi.log"), os.path.join(P, "r02_pytest_multi_gpu.txt"))
    if os.path.exists(os.path.join(G, "ev_pytest.log")):
        shutil.copy(os.path.join(G, "ev_pytest.log"), os.path.join(P, "r02_pytest_gpu.txt"))
