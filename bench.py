#!/usr/bin/env python
"""Benchmark of the likelihood hot path: walker-steps/s (= ensemble lnprob
evaluations/s) on BASELINE.json's RXJ1713 Syn+IC configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one ensemble step of the stretch-move sampler: every walker gets one
proposal and one likelihood evaluation (two half-ensemble batches).  At N > 1
(launched under torchrun) the ensemble is sharded over ranks, 256 walkers per
GPU (weak scaling), with one all-gather of log-probabilities per half-step.

`value`  device-resident loop (positions, random draws, tables in HBM; one CUDA
         graph replay per step); timed with CUDA events around each step, L2
         flushed between steps.
`e2e`    the public-API loop (EnsembleSampler over the traced LikelihoodPlan):
         every half-step copies the proposals host->device from pinned memory
         and reads log-probabilities + model-flux blobs back.
`--impl reference`  the reference's CPU path: oracle restatement of naima's
         NumPy lnprob mapped over all host cores with multiprocessing.Pool, as
         core.py:446-457 + emcee do (rank 0 only).
Prints ONE JSON line.
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

METRIC = "walker-steps/sec (ensemble lnprob evals/s) at 1/2/4/8 B200 vs CPU ref"  # BASELINE.json
UNIT = "walker-steps/s"
W_PER_GPU = 256
FLUSH_MIB = 160  # > the 126 MB L2 of a B200


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush (diagnostic)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------
# CPU arm (oracle): used by --impl reference and by the cpu_baseline leg only
# ------------------------------------------------------------------------------------
_ORACLE_CTX = {}


def _oracle_init(odata):
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import oracle.naima_oracle as o
    from naima_b200 import workloads as wl

    model, prior = wl.c3_oracle(o)
    _ORACLE_CTX.update(o=o, model=model, prior=prior, data=odata)


def _oracle_lnprob(p):
    c = _ORACLE_CTX
    return c["o"].lnprob(p, c["data"], c["model"], c["prior"])[0]


def oracle_workload(W):
    """Data and walkers of the C3 workload built with the oracle only (no GPU)."""
    import oracle.naima_oracle as o
    from naima_b200 import utils, workloads as wl

    model, _ = wl.c3_oracle(o)

    def flux(E):
        return model(wl.C3_PTRUE, dict(E_eV=E, unit_fac=np.ones(E.size)))

    xt, gt = wl.c3_tables(flux)
    data = utils.validate_data_table([xt, gt])
    return wl.oracle_data(data), wl.walkers(wl.C3_PTRUE, W)


def cpu_lnprob_rate(odata, P, n_eval, cores):
    """walker-steps/s of the oracle lnprob mapped over `cores` processes."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    reps = [P[i % len(P)] for i in range(n_eval)]
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(odata,)) as pool:
        pool.map(_oracle_lnprob, reps[: max(cores, 8)])  # warm the workers
        t0 = time.perf_counter()
        pool.map(_oracle_lnprob, reps)
        dt = time.perf_counter() - t0
    return n_eval / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W = W_PER_GPU * args.gpus
    odata, P = oracle_workload(W)
    _oracle_init(odata)
    t0 = time.perf_counter()
    for p in P[:4]:
        _oracle_lnprob(p)
    t1 = (time.perf_counter() - t0) / 4
    budget = 150.0  # seconds for all (warmup + steps)
    n_sample = int(max(cores, min(W, budget * cores / t1 / (args.steps + args.warmup))))
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(odata,)) as pool:
        for it in range(args.warmup + args.steps):
            q = [P[(it * n_sample + i) % W] * (1 + 1e-3 * ((it % 7) - 3)) for i in range(n_sample)]
            t0 = time.perf_counter()
            pool.map(_oracle_lnprob, q)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    value = n_sample * args.steps / total
    sample = ("%d of %d walkers per step (one lnprob each), NumPy oracle restatement of "
              "naima's lnprob, multiprocessing.Pool(%d)" % (n_sample, W, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, W),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus, W):
    return {"workload": "RXJ1713_SynIC: Synchrotron + InverseCompton(CMB+FIR+NIR) on "
                        "ExponentialCutoffPowerLaw electrons, N_E=64 (36 X-ray + 28 VHE), "
                        "IC grid 370 nodes, synchrotron grid 570 nodes, P=4",
            "walkers": W, "walkers_per_gpu": W_PER_GPU, "n_photon_energies": 64,
            "parallelism": "walkers sharded over %d GPU(s)" % n_gpus,
            "l2": "flushed between timed steps (%d MiB memset > 126 MB L2)" % FLUSH_MIB}


# ------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        """index None: no sampling (ranks > 0: one nvidia-smi loop per box is enough, and
        eight of them polling the driver slow every rank's launches down)."""
        self.rows, self.proc = [], None
        if index is None:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def time_kernel(fn, reps=50, flush=None):
    """Average device time of fn() in ms (CUDA events on the launch stream)."""
    import torch

    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def run_native(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: naima_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import naima_b200 as nb
    from naima_b200 import engine as eng
    from naima_b200 import parallel, workloads as wl
    from naima_b200.core import PlanLogProb

    W = W_PER_GPU * world
    xt, gt = wl.c3_tables(wl.c3_device_flux)
    data = nb.validate_data_table([xt, gt])
    plan = nb.LikelihoodPlan(wl.c3_model, wl.c3_prior, data, 4)
    p0 = wl.walkers(wl.C3_PTRUE, W)
    flush_buf = torch.empty(FLUSH_MIB << 20, dtype=torch.uint8, device="cuda")

    def flush():
        if not args.no_flush:
            flush_buf.zero_()

    # ---- device-resident loop ---------------------------------------------------
    if world == 1:
        ens = nb.DeviceEnsemble(plan, W, seed=wl.SEED)
    else:
        ens = parallel.ShardedDeviceEnsemble(plan, W, seed=wl.SEED)
    ens.set_state(p0)
    ens.load_draws(args.warmup + args.steps)
    ens.run_loaded(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local if rank == 0 else None)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush()
        ev[k][0].record()
        ens.run_loaded(1)
        ev[k][1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    if world > 1:
        dist.barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = W * args.steps / (total_ms * 1e-3)
    gpu_launches = ens.kernel_launches_per_step * args.steps
    lp_final = ens.lp.cpu().numpy()
    # acceptance fraction from the chain this rank holds (no collective: the sharded
    # ensemble's acceptance_counts sums per-rank counters with an all-reduce)
    ens._wait_pushes()
    torch.cuda.synchronize()
    ch = ens.chain[:args.warmup + args.steps].cpu().numpy()
    acc_frac = float(np.mean(np.any(ch[1:] != ch[:-1], axis=2))) if len(ch) > 1 else None
    assert np.all(np.isfinite(lp_final[np.isfinite(lp_final)])) and not np.any(np.isnan(lp_final))

    # ---- end-to-end loop through the public API -----------------------------------
    # naima_b200.PlanSampler is what get_sampler()/run_sampler() build for a traceable
    # model: the EnsembleSampler API over the device-resident loop.  Every step's random
    # draws go host->device from pinned memory and its chain row, log-probabilities and
    # blob records (model flux + We) come back device->host; the host consumes one State
    # per step.  L2 is flushed before every step here too (inside the wall-clock region,
    # so `value` below is conservative; `value_excl_flush` subtracts the flushes' device time).
    def e2e_flush():
        if not args.no_flush:
            flush_buf.zero_()

    # device time of one flush, measured apart (for value_excl_flush)
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush_buf.zero_()
    fe[0].record()
    for _ in range(20):
        flush_buf.zero_()
    fe[1].record()
    torch.cuda.synchronize()
    flush_one_s = 0.0 if args.no_flush else 1e-3 * fe[0].elapsed_time(fe[1]) / 20

    sampler = nb.PlanSampler(W, 4, plan, seed=wl.SEED)  # sharded over the ranks when N > 1
    sampler._device().before_step = e2e_flush
    h2d_step, d2h_step = sampler._device().io_bytes_per_step()  # rank 0 (reads the blobs too)
    if world > 1:
        d2h_step += (world - 1) * 8 * W * (4 + 1)  # the other ranks: chain + lnprob only
        h2d_step *= world
    api = ("naima_b200.PlanSampler.sample (the sampler get_sampler()/run_sampler() build "
           "for a traced model; one State per step on the host)")
    state = sampler.run_mcmc(p0, args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    gen = sampler.sample(state, iterations=args.steps, store=True)
    t0 = time.perf_counter()
    for k in range(args.steps):
        next(gen)
    torch.cuda.synchronize()
    e2e_t = time.perf_counter() - t0
    gen.close()
    torch.cuda.synchronize()
    flush_s = flush_one_s * args.steps
    if world > 1:
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    clk = clocks.stop()
    e2e_value = W * args.steps / e2e_t
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
           "d2h_bytes_per_step": int(d2h_step), "ms_per_step": 1e3 * e2e_t / args.steps,
           "value_excl_flush": W * args.steps / max(e2e_t - flush_s, 1e-9),
           "flush_ms_per_step": 1e3 * flush_s / args.steps, "api": api}
    if rank != 0:
        return
    if world > 1:  # roofline and CPU baseline are N = 1 measurements
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, W), "clocks": clk, "e2e": e2e,
            "gpu_launches": int(gpu_launches),
            "transport": getattr(ens, "transport", None),
            "collectives_per_step": 2 if getattr(ens, "transport", "") == "nccl" else 0,
            "roofline": None, "cpu_baseline": None,
            "acceptance_fraction": acc_frac,
        }
        print(json.dumps(line))
        return

    # ---- roofline of the dominant kernel (rank 0, N = 1 shapes) ------------------------
    ex = plan.executable(W_PER_GPU // 2)
    comps = {c["kind"] + str(i): (c, out) for i, (c, out) in enumerate(zip(plan.comps, ex.outs))}
    # per-kernel device times IN SEQUENCE (set-up -> components -> combine), CUDA events
    # between the launches; the L2 flush in front doubles as a blocker that keeps the GPU
    # busy while the host enqueues, so host launch latency does not leak into the numbers
    kt, reps_k = {}, 40
    stages = plan.stages(ex)
    for it in range(reps_k + 5):
        flush_buf.zero_()
        flush_buf.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        evs[0].record()
        for k, (name, fn) in enumerate(stages):
            fn()
            evs[k + 1].record()
        torch.cuda.synchronize()
        if it >= 5:
            for k, (name, fn) in enumerate(stages):
                kt[name] = kt.get(name, 0.0) + evs[k].elapsed_time(evs[k + 1]) / reps_k
    t_eval = time_kernel(lambda: plan.run(ex), flush=flush)
    # The roofline object is for the IC integration kernel (BASELINE.json names it); the
    # synchrotron kernel's figures ride along in roofline_fp64.
    peak, peak_src = measured_peaks()
    fp64_peak = eng.fp64_peak_tflops()
    Wh = ex.W
    per_kernel = {}
    for name, (c, out) in comps.items():
        g = ex.preps[c["prep"]].grid
        if c["kind"] == "syn":
            kname = "synchrotron_kernel"
            bytes_alg = 8 * (2 * Wh * g.N + 3 * g.N + Wh + plan.N_E + Wh * plan.N_E)
            cells, flops_cell = Wh * plan.N_E * (g.N - 1), 320
        else:
            kname = "contract_kernel (IC, %d seed fields)" % c["table"].n_comp
            R = c["table"].R
            bytes_alg = 8 * (2 * R * g.N + 2 * Wh * g.N + g.N + R + Wh * R)
            cells, flops_cell = Wh * R * (g.N - 1), 164
        t = kt[name] * 1e-3
        per_kernel[name] = {"kernel": kname, "launch_us": 1e6 * t, "cells_per_launch": cells,
                            "algorithmic_bytes_per_launch": bytes_alg,
                            "achieved_GBps": bytes_alg / t / 1e9,
                            "ref_order_flops_per_cell": flops_cell,
                            "achieved_ref_order_tflops": cells * flops_cell / t / 1e12,
                            "frac_of_measured_dfma_peak": cells * flops_cell / t / 1e12 / fp64_peak,
                            "cells_per_s": cells / t}
    ic = next(v for k, v in per_kernel.items() if k.startswith("table"))
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "ncu_contract.json")
    if os.path.exists(prof):
        with open(prof) as f:
            pj = json.load(f)
        traffic = pj["launches"][-1]["dram_bytes_per_launch"]
        traffic_src = "profiles/ncu_contract.json (ncu --set full, same shapes, cold L2)"
    roofline = {"bound": "hbm", "kernel": ic["kernel"], "achieved": ic["achieved_GBps"],
                "peak": peak, "unit": "GB/s", "frac": ic["achieved_GBps"] / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ic["algorithmic_bytes_per_launch"],
                "launch_us": ic["launch_us"],
                "note": "the path is fp64-pipe / latency bound by construction (arithmetic "
                        "intensity > 1e3 flop/B, everything L2 resident): the HBM fraction is "
                        "small on purpose; see roofline_fp64 and DESIGN.md section 4"}
    # what ncu measured for the binding resource (committed captures, same shapes)
    ncu = {}
    for key, fn in (("contract", "ncu_contract.json"), ("synchrotron", "ncu_synchrotron.json")):
        pth = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(pth):
            with open(pth) as f:
                rec = json.load(f)["launches"][-1]
            ncu[key] = {
                "source": "profiles/" + fn,
                "fp64_pipe_pct_of_peak": float(
                    rec["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0]),
                "issue_active_pct": float(
                    rec["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                "duration_us": float(rec["gpu__time_duration.sum"][0]),
                "dram_bytes_per_launch": rec["dram_bytes_per_launch"]}
    roofline_fp64 = {"peak_tflops_measured_dfma": fp64_peak, "kernels": per_kernel, "ncu": ncu,
                     "stage_us": {k: 1e3 * v for k, v in kt.items()},
                     "plan_eval_us": 1e3 * t_eval, "walkers_per_launch": Wh}

    # ---- CPU baseline (oracle port on the host cores; bounded sample) -----------------
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        odata = wl.oracle_data(data)
        _oracle_init(odata)
        t0 = time.perf_counter()
        for q in p0[:3]:
            _oracle_lnprob(q)
        t1 = (time.perf_counter() - t0) / 3
        n_eval = int(max(2 * cores, min(4096, 15.0 * cores / t1)))
        rate, dt = cpu_lnprob_rate(odata, p0[:W_PER_GPU], n_eval, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d lnprob evaluations of the same workload (NumPy oracle restatement "
                         "of naima's lnprob) over multiprocessing.Pool(%d), %.1f s"
                         % (n_eval, cores, dt),
               "single_core_ms_per_lnprob": 1e3 * t1}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, W), "clocks": clk, "e2e": e2e,
        "gpu_launches": int(gpu_launches), "roofline": roofline, "roofline_fp64": roofline_fp64,
        "cpu_baseline": cpu, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
        "acceptance_fraction": acc_frac,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl != "reference":
        # CUDA graphs that captured NCCL collectives are still alive: tearing the process
        # group down under them can dead-lock, so synchronise and leave without it
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
