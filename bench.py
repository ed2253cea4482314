#!/usr/bin/env python
"""Benchmark of the likelihood hot path: walker-steps/s (= ensemble lnprob
evaluations/s) on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config C1|C2|C3|C4|C5] [--walkers-per-gpu n]

Default: C3 (RXJ1713 Syn+IC, 256 walkers per GPU), the configuration BASELINE.json's
metric is quoted on.  A "step" is one ensemble step of the stretch-move sampler: every
walker gets one proposal and one likelihood evaluation (two half-ensemble batches).  At
N > 1 (launched under torchrun) the walkers are sharded over ranks; C5 (512 walkers in
total) is the strong-scaling configuration, the others keep the per-GPU count fixed
(weak scaling).  C1 is a single `Synchrotron.flux()` call (latency; a step = one call).

`value`  device-resident loop (positions, random draws, tables in HBM; one CUDA
         graph replay per step); timed with CUDA events around each step, L2
         flushed between steps.
`e2e`    the public-API loop (`PlanSampler.sample`, what get_sampler()/run_sampler()
         build): every step's random draws go
         host->device from pinned memory and its chain row and log-probabilities come back
         device->host; the host consumes one State per step.  Blob records stay in HBM
         until get_blobs() / a State's blobs are looked at (not counted in the bytes).
         Wall clock; the L2 flush before EVERY step is inside the timed region
         (`value_excl_flush` subtracts its own device time).
`--impl reference`  the reference's CPU path: oracle restatement of naima's
         NumPy lnprob mapped over all host cores with multiprocessing.Pool, as
         core.py:446-457 + emcee do (rank 0 only).
At N > 1 a 12-step chain of the sharded ensemble is first compared BITWISE with the
single-GPU chain (`sharded_chain_bitwise`).  Prints ONE JSON line.
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
FLUSH_MIB = 160  # > the 126 MB L2 of a B200
CHECK_STEPS = 12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--walkers-per-gpu", type=int, default=None,
                    help="override the configuration's walker count (C5: the total)")
    ap.add_argument("--transport", default="auto", choices=["auto", "fused", "nccl"])
    ap.add_argument("--no-multicast", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the N > 1 bitwise chain check")
    ap.add_argument("--no-clocks", action="store_true",
                    help="diagnostic: no nvidia-smi polling during the run")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush (diagnostic)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the public-API loop (diagnostic)")
    ap.add_argument("--steps-per-graph", type=int, default=1,
                    help="diagnostic: ensemble steps captured into one CUDA graph (a timed "
                         "'step' is then that many steps; the JSON line still reports per step)")
    ap.add_argument("--timeline", default=None, metavar="PREFIX",
                    help="diagnostic (implies --no-check): every rank saves its "
                         "per-half-step device time stamps to PREFIX<rank>.npy")
    ap.add_argument("--no-blobs", action="store_true",
                    help="device loop without blob records (diagnostic: cost of moving them)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------
# CPU arm (oracle): used by --impl reference and by the cpu_baseline leg only
# ------------------------------------------------------------------------------------
_ORACLE_CTX = {}


def _oracle_init(name, odata):
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import oracle.naima_oracle as o
    from oracle import bench_models as bm

    if name == "C1":
        _ORACLE_CTX.update(o=o, bm=bm, name=name, data=odata)
        return
    model, prior = bm.MODELS[name]()
    _ORACLE_CTX.update(o=o, model=model, prior=prior, data=odata, name=name)


def _oracle_lnprob(p):
    c = _ORACLE_CTX
    if c["name"] == "C1":  # one flux() call
        return float(np.sum(c["bm"].c1_flux(c["data"], p)))
    return c["o"].lnprob(p, c["data"], c["model"], c["prior"])[0]


def oracle_workload(name, W):
    """Data and walkers of a workload built with the oracle only (no GPU)."""
    import bench_workloads as wl
    from oracle import bench_models as bm

    if name == "C1":
        pars = np.array(wl.C1_PARS)
        return wl.c1_energies(), np.tile(pars, (max(W, 1), 1))
    from naima_b200 import utils

    wk = wl.WORKLOADS[name]
    model, _ = bm.MODELS[name]()

    def flux(E):
        return model(wk.p_true, dict(E_eV=E, unit_fac=np.ones(E.size)))

    data = utils.validate_data_table(wk.tables(flux))
    return bm.oracle_data(data), wk.walkers(W)


def cpu_lnprob_rate(name, odata, P, n_eval, cores):
    """evaluations/s of the oracle lnprob mapped over `cores` processes."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    reps = [P[i % len(P)] for i in range(n_eval)]
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(name, odata)) as pool:
        pool.map(_oracle_lnprob, reps[: max(cores, 8)])  # warm the workers
        t0 = time.perf_counter()
        pool.map(_oracle_lnprob, reps)
        dt = time.perf_counter() - t0
    return n_eval / dt, dt


def n_walkers(args, world):
    import bench_workloads as wl

    if args.config == "C1":
        return 1
    return wl.WORKLOADS[args.config].total_walkers(world, args.walkers_per_gpu)


def scaling_of(args):
    import bench_workloads as wl

    return "weak" if args.config == "C1" else wl.WORKLOADS[args.config].scaling


def workload_config(args, n_gpus, W):
    import bench_workloads as wl

    if args.config == "C1":
        return {"workload": "C1: Synchrotron + ExponentialCutoffPowerLaw electrons, 64 photon "
                            "energies, synchrotron grid 570 nodes, one flux() call per step "
                            "(tests/test_models.py shapes)",
                "walkers": 1, "n_photon_energies": 64,
                "parallelism": "%d independent replica(s)" % n_gpus,
                "l2": "flushed between timed calls (%d MiB memset > 126 MB L2)" % FLUSH_MIB}
    wk = wl.WORKLOADS[args.config]
    return {"workload": "%s: %s" % (wk.name, wk.title), "walkers": W,
            "walkers_per_gpu": W // n_gpus,
            "n_photon_energies": int(sum(e.size for e in wk.energies())),
            "parallelism": "walkers sharded over %d GPU(s)" % n_gpus,
            "l2": ("not flushed (--no-flush diagnostic)" if args.no_flush else
                   "flushed between timed %s (%d MiB memset > 126 MB L2)"
                   % ("steps" if args.steps_per_graph <= 1
                      else "graphs of %d steps" % args.steps_per_graph, FLUSH_MIB))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    name = args.config
    W = n_walkers(args, args.gpus)
    odata, P = oracle_workload(name, W)
    _oracle_init(name, odata)
    t0 = time.perf_counter()
    for p in P[:2]:
        _oracle_lnprob(p)
    t1 = (time.perf_counter() - t0) / 2
    budget = 150.0  # seconds for all (warmup + steps)
    n_sample = int(max(cores, min(W, budget * cores / t1 / (args.steps + args.warmup))))
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(name, odata)) as pool:
        for it in range(args.warmup + args.steps):
            q = [P[(it * n_sample + i) % len(P)] * (1 + 1e-3 * ((it % 7) - 3))
                 for i in range(n_sample)]
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
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus, W),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
                 "--format=csv,noheader,nounits", "-lms", "100"],
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


def _dist_setup():
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
    return rank, world, local


def _max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist

    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_chain_check(nb, plan, wk, W, world, transport, multicast):
    """CHECK_STEPS steps of the sharded device ensemble and of the sharded public-API
    sampler against the single-GPU ensemble on the same seed: chain, log-probabilities,
    blob records and acceptance counts must be equal BITWISE on every rank."""
    import torch
    import torch.distributed as dist
    from naima_b200 import parallel

    p0 = wk.walkers(W)
    ref = nb.DeviceEnsemble(plan, W, seed=7)
    ref.set_state(p0)
    rchain, rlp, rrows = ref.run(CHECK_STEPS)
    racc = ref.acceptance_counts.copy()
    sh = parallel.ShardedDeviceEnsemble(plan, W, seed=7, transport=transport,
                                        multicast=multicast)
    sh.set_state(p0)
    chain, lp, rows = sh.run(CHECK_STEPS)
    ok = (np.array_equal(chain, rchain) and np.array_equal(lp, rlp)
          and np.array_equal(rows, rrows) and np.array_equal(sh.acceptance_counts, racc))
    t = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    info = {"sharded_chain_bitwise": bool(t.item() == 1.0), "check_steps": CHECK_STEPS,
            "check_walkers": W, "transport": sh.transport,
            "uses_multicast": bool(getattr(sh, "uses_multicast", False))}
    del ref
    return info, sh


def run_c1(args):
    """C1: latency of one Synchrotron.flux() call on 64 photon energies."""
    import torch

    rank, world, local = _dist_setup()
    import bench_workloads as wl
    import naima_b200 as nb  # noqa: F401
    from naima_b200 import engine as eng
    from naima_b200 import units as u
    from naima_b200.models import ExponentialCutoffPowerLaw, Synchrotron

    E = wl.c1_energies()
    flush_buf = torch.empty(FLUSH_MIB << 20, dtype=torch.uint8, device="cuda")
    ECPL = ExponentialCutoffPowerLaw(wl.C1_PARS[0] / u.eV, wl.C1_PARS[1] * u.TeV, wl.C1_PARS[2],
                                     wl.C1_PARS[3] * u.TeV)
    SYN = Synchrotron(ECPL, B=wl.C1_PARS[4] * u.uG)
    Eq = u.Quantity(E, "eV")
    for _ in range(max(args.warmup, 3)):
        SYN.flux(Eq, distance=1 * u.kpc)
    torch.cuda.synchronize()
    # device-resident: the launches of one call on resident inputs
    kind, par_d, W = SYN._pd_device()
    g = SYN._grid()
    B_d, E_d = eng.to_dev([wl.C1_PARS[4] * 1e-6]), eng.photon_energies(E)
    out, fl = eng.empty(1, E.size), eng.empty(1, E.size)
    ones = eng.to_dev(np.ones(E.size))
    pr = eng.pd_prep(g, kind, par_d, 1, need_raw=False)

    def dev_call():
        eng.pd_prep(g, kind, par_d, 1, need_raw=False, out=pr)
        eng.synchrotron(g, pr, B_d, E_d, out=out)
        eng.combine([(out, 0, True, 4 * np.pi * (1e3 * 3.0856775814913673e18) ** 2, None)], 1,
                    E.size, ones, flux_out=fl)

    clocks = ClockSampler(local if rank == 0 else None)
    t_dev = time_kernel(dev_call, reps=args.steps,
                        flush=None if args.no_flush else flush_buf.zero_)
    # end to end: the user's call (host Quantity in, host Quantity out)
    tot = 0.0
    for _ in range(args.steps):
        if not args.no_flush:
            flush_buf.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = SYN.flux(Eq, distance=1 * u.kpc)
        tot += time.perf_counter() - t0
    clk = clocks.stop()
    t_dev = _max_over_ranks(t_dev, world)
    tot = _max_over_ranks(tot, world)
    if rank != 0:
        return
    import oracle.bench_models as bm

    want = bm.c1_flux(E, wl.C1_PARS)
    err = float(np.max(np.abs(f.value / want - 1)))
    t0 = time.perf_counter()
    nrep = 20
    for _ in range(nrep):
        bm.c1_flux(E, wl.C1_PARS)
    t_cpu = (time.perf_counter() - t0) / nrep
    cells = E.size * (g.N - 1)
    line = {
        "metric": METRIC, "value": world / (t_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world, 1), "clocks": clk,
        "e2e": {"value": world * args.steps / tot, "unit": UNIT,
                "h2d_bytes_per_step": int(8 * (E.size + 8 + 1)),
                "d2h_bytes_per_step": int(8 * E.size), "ms_per_step": 1e3 * tot / args.steps,
                "api": "naima_b200.models.Synchrotron.flux (one call, one walker)"},
        "gpu_launches": 3 * args.steps,
        "roofline": {"bound": "hbm", "kernel": "synchrotron_kernel (1 walker)",
                     "achieved": 8 * (6 * g.N + 2 * E.size) / (t_dev * 1e-3) / 1e9,
                     "peak": measured_peaks()[0], "unit": "GB/s",
                     "frac": 8 * (6 * g.N + 2 * E.size) / (t_dev * 1e-3) / 1e9 / measured_peaks()[0],
                     "traffic": None,
                     "note": "latency case: one walker = %d cells on one SM pair; launch "
                             "latency bound, not a throughput measurement" % cells},
        "cpu_baseline": {"value": 1.0 / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d flux() calls of the NumPy oracle restatement, one core"
                                   % nrep},
        "max_rel_err_vs_oracle": err,
    }
    print(json.dumps(line))


def run_native(args):
    import torch

    if args.config == "C1":
        return run_c1(args)
    rank, world, local = _dist_setup()
    if world > 1:
        import torch.distributed as dist
    import bench_workloads as wl
    import naima_b200 as nb
    from naima_b200 import engine as eng
    from naima_b200 import parallel

    wk = wl.WORKLOADS[args.config]
    W = n_walkers(args, world)
    if W % (2 * world):
        raise SystemExit("%d walkers do not split into two halves over %d GPUs" % (W, world))
    data = nb.validate_data_table(wk.tables())
    plan = nb.LikelihoodPlan(wk.model, wk.prior, data, wk.P)
    p0 = wk.walkers(W)
    flush_buf = torch.empty(FLUSH_MIB << 20, dtype=torch.uint8, device="cuda")
    multicast = not args.no_multicast

    def flush():
        if not args.no_flush:
            flush_buf.zero_()

    # ---- N > 1: the sharded chain must equal the single-GPU chain bit for bit ---------
    check = None
    if world == 1:
        ens = nb.DeviceEnsemble(plan, W, seed=wl.SEED, store_blobs=not args.no_blobs,
                                timeline=bool(args.timeline))
    else:
        if args.no_check or args.timeline:
            ens = parallel.ShardedDeviceEnsemble(plan, W, seed=wl.SEED, transport=args.transport,
                                                 multicast=multicast,
                                                 store_blobs=not args.no_blobs,
                                                 timeline=bool(args.timeline))
        else:
            check, ens = sharded_chain_check(nb, plan, wk, W, world, args.transport, multicast)
            ens._random = np.random.mtrand.RandomState(wl.SEED)

    # ---- device-resident loop ---------------------------------------------------
    spg = max(1, args.steps_per_graph)
    if args.steps % spg or args.warmup % spg:
        raise SystemExit("--steps and --warmup must be multiples of --steps-per-graph")
    ens.steps_per_graph = spg
    ens.set_state(p0)
    ens.load_draws(args.warmup + args.steps)
    ens.run_loaded(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local if (rank == 0 and not args.no_clocks) else None)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps // spg)]
    eng.fallback_counts(reset=True)
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps // spg):
        flush()
        ev[k][0].record()
        ens.run_loaded(spg)
        ev[k][1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    if world > 1:
        dist.barrier()
    if args.timeline:
        tl = ens.timeline()
        np.save("%s%d.npy" % (args.timeline, rank), np.concatenate(
            [tl, np.full((tl.shape[0], 1), 2 * (args.warmup + args.steps), dtype=tl.dtype)],
            axis=1))
    fb_contract, fb_ssc = eng.fallback_counts()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev]) / spg
    total_ms = _max_over_ranks(float(step_ms.sum()) * spg, world)
    value = W * args.steps / (total_ms * 1e-3)
    gpu_launches = ens.kernel_launches_per_step * args.steps
    lp_final = ens.lp.cpu().numpy()
    # acceptance fraction from the chain this rank holds
    ens._wait_pushes()
    torch.cuda.synchronize()
    ch = ens.chain[:args.warmup + args.steps].cpu().numpy()
    acc_frac = float(np.mean(np.any(ch[1:] != ch[:-1], axis=2))) if len(ch) > 1 else None
    assert not np.any(np.isnan(lp_final))

    # ---- end-to-end loop through the public API -----------------------------------
    # naima_b200.PlanSampler is what get_sampler()/run_sampler() build for a traceable
    # model: the EnsembleSampler API over the device-resident loop.  Every step's random
    # draws go host->device from pinned memory and its chain row, log-probabilities and
    # blob records come back device->host; the host consumes one State per step.  L2 is
    # flushed before every step here too (inside the wall-clock region, so `value` below
    # is conservative; `value_excl_flush` subtracts the flushes' device time).
    e2e = None
    clk = None
    if not args.no_e2e:
        fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        flush_buf.zero_()
        fe[0].record()
        for _ in range(20):
            flush_buf.zero_()
        fe[1].record()
        torch.cuda.synchronize()
        flush_one_s = 0.0 if args.no_flush else 1e-3 * fe[0].elapsed_time(fe[1]) / 20

        block = int(min(32, max(4, args.steps // 4)))
        sampler = nb.PlanSampler(W, wk.P, plan, seed=wl.SEED, block=block,
                                 transport=args.transport)  # sharded over the ranks when N > 1
        sampler._device().before_step = flush
        # per step and rank: the draws go up, the chain row and the log-probabilities come
        # down; the blob records (model flux + blobs of every walker) stay in HBM until
        # get_blobs() / a State's blobs are looked at
        h2d_step, d2h_step = sampler._device().io_bytes_per_step(rows_to_host=False)
        h2d_step, d2h_step = h2d_step * world, d2h_step * world
        api = ("naima_b200.PlanSampler.sample (the sampler get_sampler()/run_sampler() build "
               "for a traced model; one State per step on the host, blob records fetched "
               "from HBM on demand)")
        state = sampler.run_mcmc(p0, max(args.warmup, 3))
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
        e2e_t = _max_over_ranks(e2e_t, world)
        e2e_value = W * args.steps / e2e_t
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
               "d2h_bytes_per_step": int(d2h_step), "ms_per_step": 1e3 * e2e_t / args.steps,
               "value_excl_flush": W * args.steps / max(e2e_t - flush_s, 1e-9),
               "flush_ms_per_step": 1e3 * flush_s / args.steps, "block_steps": block,
               "l2": "not flushed" if args.no_flush else
                     "flushed before every step (%d MiB memset, inside the timed region)" % FLUSH_MIB,
               "api": api}
    clk = clocks.stop()
    if rank != 0:
        return
    base = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": wk.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, W), "clocks": clk, "e2e": e2e,
        "gpu_launches": int(gpu_launches), "acceptance_fraction": acc_frac,
        "step_ms_min_median_max": [float(step_ms.min()), float(np.median(step_ms)),
                                   float(step_ms.max())],
        # how often the lean integration cell met an irregular slope and the range was
        # redone with the careful cell, in the timed region of this rank
        "lean_cell_fallbacks": {"contract_walker_tiles": fb_contract, "ssc_rows": fb_ssc},
    }
    if world > 1:  # roofline and CPU baseline are N = 1 measurements
        base.update({"transport": getattr(ens, "transport", None),
                     "uses_multicast": bool(getattr(ens, "uses_multicast", False)),
                     "collectives_per_step": 2 if getattr(ens, "transport", "") == "nccl" else 0,
                     "roofline": None, "cpu_baseline": None})
        if check is not None:
            base.update(check)
        print(json.dumps(base))
        return

    # ---- roofline of the dominant kernels (rank 0, N = 1 shapes) ------------------------
    ex = plan.executable(W // 2)
    # per-kernel device times IN SEQUENCE (set-up -> components -> combine), CUDA events
    # between the launches; the L2 flush in front doubles as a blocker that keeps the GPU
    # busy while the host enqueues, so host launch latency does not leak into the numbers
    kt, reps_k = {}, 30
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
    t_eval = time_kernel(lambda: plan.run(ex), flush=flush, reps=30)
    peak, peak_src = measured_peaks()
    fp64_peak = eng.fp64_peak_tflops()
    per_kernel = plan.kernel_figures(ex)  # algorithmic cells / bytes per launch, per stage
    ncu = load_ncu_profiles(args.config)
    for name, pk in per_kernel.items():
        t = kt[name] * 1e-3
        pk.update({"launch_us": 1e6 * t, "achieved_GBps": pk["algorithmic_bytes_per_launch"] / t / 1e9,
                   "achieved_ref_order_tflops": pk["cells_per_launch"] * pk["ref_order_flops_per_cell"] / t / 1e12,
                   "cells_per_s": pk["cells_per_launch"] / t})
        # fraction of the fp64 pipe kept busy by EXECUTED fp64 instructions (ncu instruction
        # counts of the committed capture of the same launch shape / this run's time)
        prof = ncu.get(pk["kernel"].split(" ")[0])
        if prof and prof.get("fp64_thread_insts_per_launch") and \
                prof.get("cells_per_launch") == pk["cells_per_launch"]:
            pk["executed_fp64_frac"] = (prof["fp64_thread_insts_per_launch"] / t
                                        / (0.5 * fp64_peak * 1e12))
            pk["executed_fp64_source"] = prof["source"]
    # the roofline object is for the IC integration kernel where the configuration has one
    # (BASELINE.json names it), else for the longest component kernel
    names = list(per_kernel)
    ic = [n for n in names if per_kernel[n].get("is_ic")]
    dom = ic[0] if ic else max(names, key=lambda n: per_kernel[n]["launch_us"])
    d = per_kernel[dom]
    prof = ncu.get(d["kernel"].split(" ")[0]) or {}
    roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved_GBps"],
                "peak": peak, "unit": "GB/s", "frac": d["achieved_GBps"] / peak,
                "traffic": prof.get("dram_bytes_per_launch"),
                "traffic_source": prof.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                "launch_us": d["launch_us"],
                "executed_fp64_frac": d.get("executed_fp64_frac"),
                "note": "the path is fp64-pipe / latency bound by construction (arithmetic "
                        "intensity > 1e3 flop/B, everything L2 resident): the HBM fraction is "
                        "small on purpose; executed_fp64_frac = executed fp64 lane-instructions "
                        "/ time / measured DFMA issue rate is the utilisation figure; see "
                        "roofline_fp64 and DESIGN.md section 4"}
    roofline_fp64 = {"peak_tflops_measured_dfma": fp64_peak, "kernels": per_kernel,
                     "ncu": ncu, "stage_us": {k: 1e3 * v for k, v in kt.items()},
                     "plan_eval_us": 1e3 * t_eval, "walkers_per_launch": ex.W,
                     "note": "achieved_ref_order_tflops multiplies cells by the REFERENCE-order "
                             "flops per cell (SURVEY 8d) that hoisting removed: a speed-up "
                             "figure, not a utilisation"}

    # ---- CPU baseline (oracle port on the host cores; bounded sample) -----------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import bench_models as bm

        cores = os.cpu_count() or 1
        odata = bm.oracle_data(data)
        _oracle_init(args.config, odata)
        t0 = time.perf_counter()
        for q in p0[:2]:
            _oracle_lnprob(q)
        t1 = (time.perf_counter() - t0) / 2
        n_eval = int(max(2 * cores, min(4096, 15.0 * cores / t1)))
        rate, dt = cpu_lnprob_rate(args.config, odata, p0[:min(W, 256)], n_eval, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d lnprob evaluations of the same workload (NumPy oracle restatement "
                         "of naima's lnprob) over multiprocessing.Pool(%d), %.1f s"
                         % (n_eval, cores, dt),
               "single_core_ms_per_lnprob": 1e3 * t1}
    base.update({"roofline": roofline, "roofline_fp64": roofline_fp64, "cpu_baseline": cpu,
                 "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps})
    print(json.dumps(base))


def load_ncu_profiles(config):
    """Committed `ncu --set full` summaries (profiles/ncu_<config>_<kernel>.json, written by
    tools/ncu_summary.py): kernel base name -> figures of its last captured launch."""
    out = {}
    pdir = os.path.join(ROOT, "profiles")
    pre = "ncu_%s_" % config.lower()
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not (fn.startswith(pre) and fn.endswith(".json")):
            continue
        try:
            with open(os.path.join(pdir, fn)) as f:
                pj = json.load(f)
            rec = pj["launches"][-1]
            name = rec["kernel"].replace("void ", "").replace("kernel ", "").split("(")[0].split("<")[0]
            ent = {"source": "profiles/" + fn, "kernel": rec["kernel"][:80],
                   "dram_bytes_per_launch": rec.get("dram_bytes_per_launch"),
                   "duration_us": float(rec["gpu__time_duration.sum"][0]),
                   "cells_per_launch": pj.get("cells_per_launch")}
            for key, short in (
                    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct_of_active"),
                    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")):
                if key in rec:
                    ent[short] = float(rec[key][0])
            if rec.get("fp64_thread_insts"):
                ent["fp64_thread_insts_per_launch"] = rec["fp64_thread_insts"]
            out[name] = ent
        except Exception:
            continue
    return out


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
