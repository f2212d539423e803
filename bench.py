#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched-replication engine (BASELINE.json).

Workload (BASELINE.json configs[1]): M/M/1 batched -- Generator Exp(0.5) -> Processor Exp(0.333333)
capacity 14 -> Storage, 2^20 replications per GPU, every replication run with
Simulation::step_until(10 000); statistic = IID mean waiting time over replications.
Metric: simulated events/sec (one event = one events_ext or events_int call of the reference,
sim/src/simulator/mod.rs:220,241), whole job over all GPUs.

One bench "step" = one pass of the hot path over one batch: re-initialise the models of all
replications (Simulation::reset + put) and run step_until(T) on every replication.

  python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of `roofline`,
`cpu_baseline` and `e2e`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "simulated events/sec (replications x events), M/M/1 batched step_until(10000)"
UNIT = "events/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--replications", type=int, default=1 << 20, help="replications per GPU")
    p.add_argument("--until", type=float, default=10000.0)
    p.add_argument("--k", type=int, default=0, help="Simulation::step()s fused per kernel launch (0 = engine default)")
    p.add_argument("--network", default="mm1")
    p.add_argument("--cpu-replications", type=int, default=0, help="replications of the CPU baseline sample (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-k1", action="store_true", help="skip the one-step-per-launch (HBM streaming) section")
    return p.parse_args()


def workload(name):
    import networks
    models, conns = networks.ALL[name]()
    stat = {"mm1": "processor-01", "tandem": "processor-3"}.get(name)
    return models, conns, stat


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(args, threads, nreps):
    """Oracle (CPU port of the reference path) on a bounded sample of the same workload."""
    import oracle_api
    models, conns, _ = workload(args.network)
    o = oracle_api.OracleSimulation(models, conns)
    t0 = time.perf_counter()
    r = o.run_batch(42, 0, nreps, threads=threads, until=args.until, trace_hash=False)
    dt = time.perf_counter() - t0
    ev = r["totals"]["ext"] + r["totals"]["int"]
    return ev, dt, r["totals"]["steps"]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust engine cannot be
    built in this image (no cargo/rustc), so this times the C++ oracle port with every host thread,
    one replication per thread at a time (the north_star's rayon-over-replications baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge._run(["make", "-s"], os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    nreps = args.cpu_replications or cores * 512  # ~2 s of CPU work per step at ~5e6 events/s per core
    for _ in range(min(args.warmup, 1)):
        cpu_sample(args, cores, max(cores, 8))
    tot_ev, tot_dt = 0, 0.0
    for _ in range(args.steps):
        ev, dt, _ = cpu_sample(args, cores, nreps)
        tot_ev += ev
        tot_dt += dt
    v = tot_ev / tot_dt
    sample = "%d replications x step_until(%g) per step, %d threads" % (nreps, args.until, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "M/M/1 batched (BASELINE configs[1]): Generator Exp(0.5) -> Processor Exp(0.333333) cap 14 -> Storage"
                   if args.network == "mm1" else args.network,
                   "replications_per_step": nreps, "until": args.until, "sample": sample,
                   "implementation": "C++ oracle port of the reference engine (the Rust crate cannot be built in this image)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from sim_b200 import Simulation
    from sim_b200.simulator import connectors_to_json, models_to_json

    models, conns, stat = workload(args.network)
    mj, cj = models_to_json(models), connectors_to_json(conns)
    n = args.replications
    sim = Simulation.post_json(mj, cj, n, seed=42, replication_offset=rank * n, device=local_rank,
                               steps_per_launch=args.k, stat_model_id=stat)
    desc = sim.describe()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        sim.reset()                 # Simulation::reset + put: fresh models, the RNG streams continue
        sim.put(models, conns)
        sim.step_until(args.until)  # synchronous: returns when every replication has crossed `until`

    for _ in range(args.warmup):
        one_step()
    barrier()
    sim.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev_done = 0
    steps_done = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
        c = sim.get_counters()  # per-replication step/event counters restart with the models
        ev_done += c["events"]
        steps_done += c["steps"]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ctr = sim.get_counters()

    # max over ranks of the elapsed time, sum over ranks of the work
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    work = torch.tensor([float(ev_done), float(steps_done)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    dt_max = float(tmax.item())
    events_total, rsteps_total = float(work[0].item()), float(work[1].item())
    value = events_total / dt_max

    # statistics: the only collective of the path -- all-reduce of (n, sum) then of the squared deviations
    s_sum, s_n = sim.stat_partial() if stat else (0.0, 0)
    red = torch.tensor([s_sum, float(s_n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
    ci = None
    if stat and red[1].item() > 0:
        mean = float(red[0].item() / red[1].item())
        sq = torch.tensor([sim.stat_partial_sqdev(mean)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sq, op=dist.ReduceOp.SUM)
        from sim_b200.simulator import _CI, load_library
        import ctypes as C
        out = _CI()
        load_library().simb_oa_confidence_interval(mean, float(sq.item() / red[1].item()), int(red[1].item()), 0.05, C.byref(out))
        ci = {"mean": out.mean, "lower": out.lower, "upper": out.upper, "n": out.n}

    # roofline of the dominant kernel (simb_step_kernel): algorithmic bytes per launch / mean launch time
    peak, peak_src = peaks()
    ncu = {}
    ncu_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_path):
        with open(ncu_path) as f:
            ncu = json.load(f)

    def roofline_of(desc_, ctr_, rsteps_local, wall_s, tag):
        state_b = desc_["state_bytes_per_replication"]
        launches_ = max(1, ctr_["launches"])
        # ring touches per replication-step for this network: ~0.4 slots of 4 B (+8 B arrival time)
        ring_b = 0.4 * (4 + (8 if stat else 0)) if desc_["ring_u32_slots"] else 0.0
        # a pass over the batch is issued as `grids` concurrent grids on their own streams, each covering
        # 1/grids of the replications; launch_ms is the device time of the call divided by the launches
        grids = max(1, int(desc_.get("grids_per_pass", 1)))
        alg = desc_["stride"] * (2.0 * state_b) / grids + (rsteps_local / launches_) * ring_b
        launch_ms_ = ctr_["kernel_ms"] / launches_
        ach = alg / (launch_ms_ * 1e-3) / 1e9 if launch_ms_ > 0 else 0.0
        cap = ncu.get(tag, {})
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": cap.get("dram_bytes_per_launch"), "kernel": "simb_step_kernel<%s>" % desc_["kernel_shape"],
                "launches": ctr_["launches"], "grids_per_pass": grids, "launch_ms": launch_ms_,
                "algorithmic_bytes_per_launch": alg,
                "steps_per_launch": desc_["steps_per_launch"], "state_bytes_per_replication": state_b,
                "peak_source": peak_src, "kernel_share_of_step": ctr_["kernel_ms"] * 1e-3 / wall_s if wall_s > 0 else None,
                "ncu": cap or None}

    k_eff = desc["steps_per_launch"]
    launches = ctr["launches"]
    roofline = roofline_of(desc, ctr, rsteps_total / world, dt, "fused")
    roofline["note"] = ("fused steps: the state is loaded once per %d steps, so the kernel sits far below the HBM roof "
                        "and is bound by instruction issue (ncu issue-active in roofline.ncu); the HBM-streaming "
                        "formulation (one step per launch, Simulation::step granularity) is reported in hbm_streaming_k1" % k_eff)

    # the same workload with ONE step per launch (API-faithful Simulation::step granularity): every launch
    # streams the whole state through HBM; bounded to a quarter of the horizon to keep the run short
    k1 = None
    if rank == 0 and world == 1 and not args.no_k1:
        s1 = Simulation.post_json(mj, cj, n, seed=42, device=local_rank, steps_per_launch=1, stat_model_id=stat)
        s1.step_until(args.until / 16.0)
        s1.reset_counters()
        c0 = s1.get_counters()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        s1.step_until(args.until / 4.0)
        torch.cuda.synchronize()
        w1 = time.perf_counter() - t1
        c1 = s1.get_counters()
        r1 = roofline_of(s1.describe(), c1, c1["steps"] - c0["steps"], w1, "k1")
        k1 = {"value": (c1["events"] - c0["events"]) / w1, "unit": UNIT,
              "replication_steps_per_s": (c1["steps"] - c0["steps"]) / w1, "roofline": r1,
              "sample": "step_until(%g) -> step_until(%g), steps_per_launch=1" % (args.until / 16.0, args.until / 4.0)}
        s1.close()

    # end to end through the public API with host buffers: post (JSON in), run, waits read back to the host
    e2e = None
    if not args.no_e2e:
        barrier()
        e_ev = 0
        h2d = len(mj) + len(cj) + 8 * desc["f64_rows"] + 4 * desc["u32_rows"]
        d2h = n * 12
        t1 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 2))):
            s2 = Simulation.post_json(mj, cj, n, seed=4242, replication_offset=rank * n, device=local_rank,
                                      steps_per_launch=args.k, stat_model_id=stat)
            s2.step_until(args.until)
            if stat:
                s2.get_wait_stats()
            e_ev += s2.get_counters()["events"]
            s2.close()
        barrier()
        e_dt = time.perf_counter() - t1
        ev_t = torch.tensor([float(e_ev)], dtype=torch.float64, device="cuda")
        tt = torch.tensor([e_dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ev_t, op=dist.ReduceOp.SUM)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(ev_t.item() / tt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import __graft_entry__ as ge
        ge._run(["make", "-s"], os.path.join(ROOT, "oracle"))
        cores = os.cpu_count() or 1
        nreps = args.cpu_replications or cores * 2560  # ~10 s of CPU work at ~5e6 events/s per core
        ev, cdt, _ = cpu_sample(args, cores, nreps)
        cpu = {"value": ev / cdt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d replications x step_until(%g), one replication per thread, %d threads, %.1f s" % (nreps, args.until, cores, cdt)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt_max / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "M/M/1 batched (BASELINE configs[1]): Generator Exp(0.5) -> Processor Exp(0.333333) cap 14 -> Storage"
                       if args.network == "mm1" else args.network,
                       "replications_per_gpu": n, "until": args.until, "steps_per_launch": k_eff,
                       "rng": "philox4x32-10 per replication", "tile": desc["tile"],
                       "l2": "state %.0f MB + rings %.0f MB per GPU vs 126 MB L2 (inputs larger than L2; no flush)" % (
                           desc["stride"] * desc["state_bytes_per_replication"] / 1e6,
                           desc["stride"] * (desc["ring_u32_slots"] * 4 + desc["ring_f64_slots"] * 8) / 1e6),
                       "parallelism": "replications sharded, %d per GPU" % n},
            "replication_steps_per_s": rsteps_total / dt_max,
            "gpu_launches": int(launches * world + 2 * args.steps * world),
            "clocks": clocks, "roofline": roofline, "hbm_streaming_k1": k1, "cpu_baseline": cpu, "e2e": e2e,
            "statistic": ci,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
