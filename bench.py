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

Besides the headline (weak scaling: 2^20 M/M/1 replications per GPU) the same line carries
  hbm_streaming_k1  one Simulation::step per launch at 2^20 (state 101 MB: L2-assisted) and 2^22 replications
                    (403 MB: streams from HBM), algorithmic GB/s beside the ncu-measured DRAM GB/s,
  configs           BASELINE configs[2..4] at their named sizes (tandem 2^22, gateway network 2^20, the 32-model
                    network 2^23 per GPU) with a bounded horizon each,
  strong_scaling    fixed-total jobs split over the ranks: configs[4] with 2^26 replications in total and
                    M/M/1 with 2^20 in total.

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

UNIT = "events/s"
WORKLOADS = {
    "mm1": "M/M/1 batched (BASELINE configs[1]): Generator Exp(0.5) -> Processor Exp(0.333333) cap 14 -> Storage",
    "tandem": "Tandem 3-stage queue (BASELINE configs[2]): Generator -> LoadBalancer -> 3 x Processor cap 8 -> Processor "
              "-> Processor -> Storage, steady-state batch means of the last stage",
    "gateway_network": "Gateway network (BASELINE configs[3]): 16-model DAG with ExclusiveGateway, ParallelGateway split "
                       "and join, StochasticGate, Gate, Batcher, Stopwatch",
    "large_mixed": "Large mixed network (BASELINE configs[4]): 32 models, two gateway sub-networks and the tandem behind one split",
}


def metric_name(network, until):
    short = {"mm1": "M/M/1 batched", "tandem": "tandem queue", "gateway_network": "gateway network",
             "large_mixed": "32-model mixed network"}.get(network, network)
    return "simulated events/sec (replications x events), %s step_until(%g)" % (short, until)


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
    p.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    p.add_argument("--no-strong", action="store_true", help="skip the fixed-total (strong scaling) jobs")
    p.add_argument("--strong-total-log2", type=int, default=26, help="replications of the fixed-total configs[4] job (log2)")
    p.add_argument("--sub-batch-log2", type=int, default=23, help="largest shard of the 32-model network posted at once (log2)")
    return p.parse_args()


def workload(name):
    import networks
    models, conns = networks.ALL[name]()
    stat = {"mm1": "processor-01", "tandem": "processor-3"}.get(name)  # waiting-time statistic of that Processor
    return models, conns, stat


def json_safe(v):
    """NaN / infinity are not JSON: null."""
    if isinstance(v, dict):
        return {k: json_safe(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [json_safe(x) for x in v]
    if isinstance(v, float) and (v != v or v in (float("inf"), float("-inf"))):
        return None
    return v


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


def job_config(network, replications_per_gpu, until):
    """The `config` object of the JSON line: what the job IS (identical in both arms; how each arm runs it is
    described in `engine` / `cpu_baseline.sample`)."""
    return {"workload": WORKLOADS.get(network, network), "replications_per_gpu": replications_per_gpu, "until": until,
            "rng": "philox4x32-10, key = seed 42, counter = (block, global replication index)",
            "parallelism": "independent replications, contiguous index blocks per GPU, no data-path collective"}


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
        "impl": "reference", "metric": metric_name(args.network, args.until), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": job_config(args.network, args.replications, args.until),
        "implementation": "C++ oracle port of the reference engine (the Rust crate cannot be built in this image); "
                          "each step is a bounded sample of the job: " + sample,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=OUT, flush=True)


OUT = sys.stdout


def main():
    global OUT
    args = parse()
    # stdout carries ONE JSON line: whatever libraries write to file descriptor 1 on the way (NCCL's version banner)
    # goes to stderr instead
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np  # noqa: F401
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
    peak, peak_src = peaks()
    ncu = {}
    ncu_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(ncu_path):
        with open(ncu_path) as f:
            ncu = json.load(f)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(xs):
        t = torch.tensor([float(x) for x in xs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def roofline_of(desc_, ctr_, wall_s, tag):
        """Dominant kernel (simb_step_kernel) of a timed region: algorithmic bytes per launch / mean launch time.
        Algorithmic bytes = every state word of the launch's replications read once and written once (2 x state
        bytes x replications of the grid) + the queue / table bytes the launches touched in global memory, COUNTED
        on the device (simb_counters.ring_bytes)."""
        state_b = desc_["state_bytes_per_replication"]
        launches_ = max(1, ctr_["launches"])
        # a pass over the batch is issued as `grids` concurrent grids on their own streams, each covering
        # 1/grids of the replications; launch_ms is the device time of the calls divided by the launches
        grids = max(1, int(desc_.get("grids_per_pass", 1)))
        ring_per_launch = ctr_["ring_bytes"] / launches_
        alg = desc_["stride"] * (2.0 * state_b) / grids + ring_per_launch
        launch_ms_ = ctr_["kernel_ms"] / launches_
        ach = alg / (launch_ms_ * 1e-3) / 1e9 if launch_ms_ > 0 else 0.0
        cap = ncu.get(tag, {})
        r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": cap.get("dram_bytes_per_launch"), "kernel": "simb_step_kernel<%s>" % desc_["kernel_shape"],
             "launches": ctr_["launches"], "grids_per_pass": grids, "launch_ms": launch_ms_,
             "algorithmic_bytes_per_launch": alg, "ring_bytes_per_launch_counted": ring_per_launch,
             "steps_per_launch": desc_["steps_per_launch"], "state_bytes_per_replication": state_b,
             "peak_source": peak_src, "kernel_share_of_step": ctr_["kernel_ms"] * 1e-3 / wall_s if wall_s > 0 else None}
        if cap:
            # DRAM bytes of the committed ncu capture of this configuration (tools/ncu_summary.py) over the launch
            # time measured live here: the fraction of the HBM peak that really came from DRAM
            r["traffic_source"] = "profiles/ncu_summary.json[%s] (ncu --set full capture of the same configuration, %s)" % (
                tag, cap.get("source"))
            if launch_ms_ > 0 and cap.get("dram_bytes_per_launch"):
                r["dram_gbs"] = cap["dram_bytes_per_launch"] / (launch_ms_ * 1e-3) / 1e9
                r["dram_frac"] = r["dram_gbs"] / peak
            r["l2_hit_pct"] = cap.get("l2_hit_pct")
            r["ncu"] = cap
        return r

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
    ctr = sim.get_counters()

    # max over ranks of the elapsed time, sum over ranks of the work
    dt_max = reduce_max(dt)
    events_total, rsteps_total = reduce_sum([ev_done, steps_done])
    value = events_total / dt_max

    # statistics: the only collective of the path -- all-reduce of (n, sum) then of the squared deviations
    s_sum, s_n = sim.stat_partial() if stat else (0.0, 0)
    red = reduce_sum([s_sum, s_n])
    ci = None
    if stat and red[1] > 0:
        mean = red[0] / red[1]
        sq = reduce_sum([sim.stat_partial_sqdev(mean)])[0]
        from sim_b200.simulator import _CI, load_library
        import ctypes as C
        out = _CI()
        load_library().simb_oa_confidence_interval(mean, sq / red[1], int(red[1]), 0.05, C.byref(out))
        ci = {"mean": out.mean, "lower": out.lower, "upper": out.upper, "n": out.n}

    k_eff = desc["steps_per_launch"]
    launches = ctr["launches"]
    roofline = roofline_of(desc, ctr, dt, "fused")
    roofline["note"] = ("fused steps: the state is loaded once per %d steps, so the kernel sits far below the HBM roof "
                        "and is bound by instruction issue (ncu issue-active in roofline.ncu); the HBM-streaming "
                        "formulation (one step per launch, Simulation::step granularity) is reported in hbm_streaming_k1" % k_eff)

    # ---- end to end through the public API with host buffers: post (JSON in), run, waits read back to the host ----
    e2e = None
    if not args.no_e2e:
        # (the clock sampler belongs to the timed region above: an nvidia-smi query takes a driver lock that device
        # allocations wait for, and this region allocates)
        clocks = sampler.stop() if rank == 0 else None
        sampler = None

        def one_call():
            s2 = Simulation.post_json(mj, cj, n, seed=4242, replication_offset=rank * n, device=local_rank,
                                      steps_per_launch=args.k, stat_model_id=stat)
            s2.step_until(args.until)
            if stat:
                s2.get_wait_stats()
            ev = s2.get_counters()["events"]
            s2.close()
            return ev

        one_call()  # warm-up: the device memory pool grows to hold this handle next to the resident one
        barrier()
        e_ev = 0
        h2d = len(mj) + len(cj) + 8 * desc["f64_rows"] + 4 * desc["u32_rows"]
        d2h = n * 12
        t1 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            e_ev += one_call()
        barrier()
        e_dt = reduce_max(time.perf_counter() - t1)
        e2e = {"value": reduce_sum([e_ev])[0] / e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    if sampler is not None:
        clocks = sampler.stop() if rank == 0 else None
    sim.close()

    # ---- the same workload with ONE step per launch (API-faithful Simulation::step granularity): every launch
    # streams the whole state through the memory system.  2^20 replications = 101 MB of state, which the 126 MB L2
    # largely holds between passes (reverse sweeps + evict_last hints): an L2-assisted number.  2^22 replications =
    # 403 MB: the state genuinely streams from HBM.  Bounded to a slice of the horizon. ----
    k1 = None
    if rank == 0 and world == 1 and not args.no_k1:
        k1 = []
        for log2n, tag in ((20, "k1"), (22, "k1_2p22")):
            nn = 1 << log2n
            s1 = Simulation.post_json(mj, cj, nn, seed=42, device=local_rank, steps_per_launch=1, stat_model_id=stat)
            s1.step_until(args.until / 16.0)
            c0 = s1.get_counters()
            s1.reset_counters()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            s1.step_until(args.until / 8.0)
            torch.cuda.synchronize()
            w1 = time.perf_counter() - t1
            c1 = s1.get_counters()
            d1 = s1.describe()
            r1 = roofline_of(d1, c1, w1, tag)
            state_mb = d1["stride"] * d1["state_bytes_per_replication"] / 1e6
            k1.append({"replications": nn, "value": (c1["events"] - c0["events"]) / w1, "unit": UNIT,
                       "replication_steps_per_s": (c1["steps"] - c0["steps"]) / w1, "roofline": r1,
                       "state_mb": state_mb,
                       "l2": ("state %.0f MB < 126 MB L2: consecutive passes are L2-assisted, compare dram_frac" % state_mb) if state_mb < 126.0
                             else ("state %.0f MB > 126 MB L2: the state streams from HBM every pass" % state_mb),
                       "sample": "step_until(%g) -> step_until(%g), steps_per_launch=1" % (args.until / 16.0, args.until / 8.0)})
            s1.close()

    # ---- the other BASELINE configurations at their named sizes, bounded horizons ----
    def run_job(network, per_rank, until, offset0, sub=None, **post_kw):
        """Every rank runs `per_rank` replications of `network` (global indices offset0 + rank * per_rank ...) to
        `until`, in sub-batches of at most `sub` replications posted one after the other.  Timed like the
        headline: barrier + synchronize on both sides, max over ranks."""
        m_, c_, stat_ = workload(network)
        mj_, cj_ = models_to_json(m_), connectors_to_json(c_)
        sub = min(per_rank, sub or per_rank)
        tot = {"events": 0, "steps": 0, "launches": 0, "kernel_ms": 0.0, "ring_bytes": 0, "failed": 0}
        d_ = None
        # warm-up: one small batch through the same kernels (first-use module load, pool growth)
        w = Simulation.post_json(mj_, cj_, min(sub, 1 << 16), seed=42, device=local_rank, stat_model_id=stat_, **post_kw)
        w.step_until(until)
        w.close()
        if sub > (1 << 16):  # ... and one post of a full shard, so that the device memory pool has its size before the clock starts
            Simulation.post_json(mj_, cj_, sub, seed=42, device=local_rank, stat_model_id=stat_, **post_kw).close()
        barrier()
        t1 = time.perf_counter()
        done = 0
        while done < per_rank:
            cnt = min(sub, per_rank - done)
            h = Simulation.post_json(mj_, cj_, cnt, seed=42, replication_offset=offset0 + rank * per_rank + done,
                                     device=local_rank, stat_model_id=stat_, **post_kw)
            h.step_until(until)
            c_ = h.get_counters()
            for k_ in tot:
                tot[k_] += c_[k_]
            d_ = h.describe()
            h.close()
            done += cnt
        barrier()
        wall = time.perf_counter() - t1
        wall_max = reduce_max(wall)
        ev, st = reduce_sum([tot["events"], tot["steps"]])
        r_ = roofline_of(d_, tot, wall, "cfg_" + network)
        return {"metric": metric_name(network, until), "config": job_config(network, per_rank, until),
                "value": ev / wall_max, "unit": UNIT, "replication_steps_per_s": st / wall_max,
                "replications_total": per_rank * world, "sub_batch": sub, "ms": 1e3 * wall_max,
                "kernel_replication_steps_per_s_per_gpu": tot["steps"] / (tot["kernel_ms"] * 1e-3) if tot["kernel_ms"] else None,
                "failed_replications": int(reduce_sum([tot["failed"]])[0]), "roofline": r_}

    configs = None
    if not args.no_configs and args.network == "mm1":
        configs = []
        sub = 1 << args.sub_batch_log2
        for network, per_rank, until, kw in (("tandem", 1 << 22, 250.0, {"batch_warmup": 50, "batch_size": 5}),
                                             ("gateway_network", 1 << 20, 100.0, {}),
                                             ("large_mixed", 1 << 23, 4.0, {})):
            r_ = run_job(network, per_rank, until, 0, sub=sub, **kw)
            r_["scaling"] = "weak"
            configs.append(r_)

    # ---- fixed-total jobs split over the ranks (strong scaling): configs[4] with 2^26 replications in total,
    # processed by each rank in shards of at most 2^23, and M/M/1 with 2^20 in total ----
    strong = None
    if not args.no_strong and args.network == "mm1":
        strong = []
        total = 1 << args.strong_total_log2
        r_ = run_job("large_mixed", total // world, 4.0, 0, sub=1 << args.sub_batch_log2)
        r_["scaling"] = "strong"
        strong.append(r_)
        r_ = run_job("mm1", (1 << 20) // world, args.until, 0)
        r_["scaling"] = "strong"
        strong.append(r_)

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
        print(json.dumps(json_safe({
            "metric": metric_name(args.network, args.until), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": job_config(args.network, n, args.until),
            "engine": {"steps_per_launch": k_eff, "tile": desc["tile"], "kernel": roofline["kernel"],
                       "l2": "state %.0f MB + queue rings %.0f MB per GPU vs 126 MB L2; %d steps are fused per state load, so the "
                             "timed kernel is not a streaming kernel (no flush needed); see hbm_streaming_k1 for the streaming form" % (
                                 desc["stride"] * desc["state_bytes_per_replication"] / 1e6,
                                 desc["stride"] * (desc["ring_u32_slots"] * 4 + desc["ring_f64_slots"] * 8) / 1e6, k_eff)},
            "replication_steps_per_s": rsteps_total / dt_max,
            "gpu_launches": int(launches * world + 2 * args.steps * world),
            "clocks": clocks, "roofline": roofline, "hbm_streaming_k1": k1, "configs": configs, "strong_scaling": strong,
            "cpu_baseline": cpu, "e2e": e2e, "statistic": ci,
        }), allow_nan=False), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
