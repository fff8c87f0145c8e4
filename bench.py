#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched classic-control step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--env cartpole|mountain_car|pendulum]
    python bench.py --impl reference ...      # the reference-equivalent scalar CPU loop

One "step" = ONE launch of the step kernel (gymrs_step, through the C ABI) over one batch of
1,048,576 env instances with pre-generated random actions and same-launch auto-reset
(BASELINE.json configs[1]; configs[2]/[3] with --env).  Prints ONE JSON line (contract in the task
prompt / DESIGN.md section 6).

L2 policy: a 1M-env CartPole batch has a 26 MB footprint, far below the 126 MB L2, so stepping ONE
batch back to back would be served from L2 and would say nothing about HBM.  The timed region
therefore cycles through a ring of RING independent 1M-env batches (RING x 26 MB >> L2): every
launch reads and writes cold HBM lines.  The L2-resident figure (one batch, what an RL loop that
does nothing else would see) is reported separately as `l2_resident`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  NCCL prints its version banner with a plain printf when
# NCCL_DEBUG is set on the box, so file descriptor 1 is pointed at stderr for the whole run and the
# JSON line is written to a private duplicate of the original stdout (emit).
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def quiet_stdout() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

N_ENVS = 1 << 20
RING = 16
# algorithmic bytes per env-step
ALGO_BYTES = {"cartpole": 41, "mountain_car": 25, "pendulum": 37}
FOOTPRINT = {"cartpole": 25, "mountain_car": 17, "pendulum": 29}  # resident bytes per env of one batch
ACTION_SETS = 8  # pre-generated action batches per ring slot
ROLLOUT_BYTES = {"cartpole": 25, "mountain_car": 17, "pendulum": 21}  # action in + obs, reward, done out
D2H_BYTES = {"cartpole": 21, "mountain_car": 13, "pendulum": 17}  # obs + reward + done
WORKLOAD = {
    "cartpole": "CartPole-v1 1,048,576 envs/GPU, f32 SoA state, discrete int32 action, auto-reset",
    "mountain_car": "MountainCar-v0 1,048,576 envs/GPU, f32 SoA state, discrete int32 action, auto-reset",
    "pendulum": "Pendulum-v1 1,048,576 envs/GPU, f32 SoA state, continuous f32 action",
}


def config_of(args, world):
    """The workload description: the SAME dict in the B200 arm and in the reference arm (what is
    specific to one arm's method goes into its own `method` key)."""
    n, env = args.envs, args.env
    return {"workload": WORKLOAD[env], "envs_per_gpu": n,
            "l2_policy": f"inputs larger than L2: ring of {args.ring} independent {n}-env batches "
                         f"({args.ring * FOOTPRINT[env] * n / 1e6:.0f} MB footprint per GPU) stepped round-robin, so every "
                         "launch touches cold HBM lines",
            "parallelism": f"env batch sharded over {world} GPU(s) by global env id, no collective on the step path"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--env", default="cartpole", choices=list(ALGO_BYTES))
    ap.add_argument("--envs", type=int, default=N_ENVS, help="env instances per GPU")
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--vec", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=1,
                    help="launch mode of the headline measurement (1 = the library default: PDL with a full wait)")
    ap.add_argument("--wide", action="store_true",
                    help="main region with the high-occupancy build (gymrs_set_launch_occupancy); experiments / ncu captures")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--ref-repeats", type=int, default=0,
                    help="reference arm: timed regions of --steps steps (0 = as many as fit ~15 s of wall time, 5 to 200)")
    ap.add_argument("--streams", type=int, default=2,
                    help="the independent ring slots alternate over this many CUDA streams")
    ap.add_argument("--burn-in", type=int, default=300, help="untimed steps per ring slot before warm-up")
    ap.add_argument("--rollout-steps", type=int, default=64, help="0 disables the fused-rollout extra")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------
# clocks: sampled DURING the timed regions (NVML, 10 ms period)
# --------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTED = {"sw_power_cap": 0x4}

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in {**self.BAD, **self.NOTED}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this rank (and allocate its pinned host buffers) on the CPUs NVML
    reports as local to the GPU, so the e2e leg's host traffic does not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(env):
    """dram bytes per launch of the step kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(env)
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's scalar loop on the host cores
# --------------------------------------------------------------------------------------
def cpu_rollout(env, n_envs, steps, warmup, budget_s):
    import oracle
    kind = {"cartpole": oracle.CARTPOLE, "mountain_car": oracle.MOUNTAIN_CAR, "pendulum": oracle.PENDULUM}[env]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:  # a cgroup CPU quota below the affinity mask is the real number of usable cores
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            cores = max(1, min(cores, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    # calibrate on a small run, then bound the sample so the whole run fits the budget
    t_cal, _ = oracle.bench_rollout(kind, n_envs, 2, 1, cores, 0)
    per_step = max(t_cal / 2, 1e-6)
    sample_envs = n_envs
    if per_step * (steps + warmup) > budget_s:
        sample_envs = max(cores * 1024, int(n_envs * budget_s / (per_step * (steps + warmup))) // 1024 * 1024)
    t, _ = oracle.bench_rollout(kind, sample_envs, steps, warmup, cores, 0)
    value = sample_envs * steps / t
    sample = (f"{sample_envs} of {n_envs} env objects per step x {steps} steps, array-of-structs f64 scalar loop, "
              f"reset on done, {cores} pthreads statically sharded")
    return value, t, cores, sample, sample_envs


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle
    steps, warmup = args.steps, args.warmup
    kind = {"cartpole": oracle.CARTPOLE, "mountain_car": oracle.MOUNTAIN_CAR, "pendulum": oracle.PENDULUM}[args.env]
    # calibrate (and, on a slow host, bound the sample): one cold region from a fresh reset
    _, t_cal, cores, sample, sample_envs = cpu_rollout(args.env, args.envs, steps, warmup, budget_s=20.0)
    # A region of `steps` steps is short (20 steps of 1 M envs = ~40 ms on 16 cores) and, straight after the
    # synchronised initial reset, it sees an episode-end burst instead of the stationary ~4.5 % resets
    # per step.  So: ONE set of env objects and threads, an untimed burn-in that decorrelates the episode
    # phases (the B200 arm's ring is burnt in the same way), then regions of { warmup untimed, steps timed }
    # for ~15 s of wall time; the median region is reported.
    per_step = t_cal / max(steps, 1)
    burnin = int(min(300, max(50, 2.0 / max(per_step, 1e-6))))
    regions = args.ref_repeats or int(min(200, max(5, 15.0 / max(per_step * (steps + warmup), 1e-6))))
    times = oracle.bench_regions(kind, sample_envs, steps, warmup, burnin, regions, cores, 0)
    t = statistics.median(times)
    value = sample_envs * steps / t
    sample += f"; {burnin} untimed burn-in steps, then the median of {len(times)} regions of {steps} steps (first region from a fresh reset, cold: {sample_envs * steps / t_cal / 1e9:.3f} G)"
    line = {
        "impl": "reference",
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * t / steps * (args.envs / sample_envs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_of(args, args.gpus),
        "method": {"what": "C restatement (oracle/) of the reference's scalar Rust step loop on the host cores; the Rust "
                           "crate itself cannot be built in this image (no cargo, SDL2 dependency)",
                   "repeats": len(times), "region_s_min_median_max": [min(times), t, max(times)]},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def make_env(g, env, n, device, offset):
    cls = {"cartpole": g.CartPoleEnv, "mountain_car": g.MountainCarEnv, "pendulum": g.PendulumEnv}[env]
    return cls(num_envs=n, device=device, global_env_offset=offset)


def make_actions(torch, env, n, device, gen):
    if env == "pendulum":
        return torch.rand((n,), generator=gen, device=device) * 4.0 - 2.0
    hi = 2 if env == "cartpole" else 3
    return torch.randint(0, hi, (n,), generator=gen, device=device, dtype=torch.int32)


def kernel_isolated_us(env):
    """Duration of ONE isolated launch of the step kernel (ncu, cold cache, serialised) from the
    committed launch list of this bench command: profiles/kernel_isolated.json."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_isolated.json"))).get(env)
    except Exception:
        return None


def mode_floor(env):
    """What kernels that only move this env's bytes cost in each stepping mode (tools/mode_floor.cu, committed
    measurement: profiles/mode_floor.json): the denominators for the one-stream and isolated figures."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "mode_floor.json")))
        return dict(d[env], source=d["source"]) if env in d else None
    except Exception:
        return None


def pcie_ceiling(torch, dist, device, world, h2d_bytes, d2h_bytes, iters=12):
    """What plain copies of one step's bytes reach on this box with ALL ranks copying at once:
    per iteration one H2D copy of h2d_bytes and one D2H copy of d2h_bytes from / to pinned memory
    on two streams (the engines run concurrently, like the e2e pipeline).  Returns aggregate GB/s."""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device=device)
    dout = torch.zeros(d2h_bytes, dtype=torch.uint8, device=device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(k):
        for _ in range(k):
            with torch.cuda.stream(s_in):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s_out):
                hout.copy_(dout, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    run(2)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run(iters)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return world * (h2d_bytes + d2h_bytes) * iters / dt / 1e9


def run_b200(args):
    import torch
    import torch.distributed as dist

    import gym_rs_b200 as g
    from gym_rs_b200 import _capi

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    # Which GPU this rank drives: with fewer ranks than visible GPUs the ranks are spread over the box (GPU
    # r * visible / N), because neighbouring GPUs share a PCIe root complex and its host-write bandwidth -- that
    # is the e2e figure's ceiling on a multi-GPU host (gym_rs_b200/sharding.py).  GYMRS_BENCH_SPREAD=0: GPU r.
    from gym_rs_b200.sharding import device_for_rank
    visible = torch.cuda.device_count()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", min(world, visible)))
    spread = os.environ.get("GYMRS_BENCH_SPREAD", "1") != "0"
    dev_index = device_for_rank(local_rank, local_world, visible, spread=spread)
    dev_stride = visible // local_world if spread else 1
    torch.cuda.set_device(dev_index)
    device = torch.device("cuda", dev_index)
    numa = bind_to_gpu_numa_node(dev_index) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    n, K, W, env = args.envs, args.steps, args.warmup, args.env
    L = _capi.load()
    # One dedicated (non-default) stream carries everything: the handles adopt torch's current
    # stream, and the CUDA events that time the run are recorded on that same stream.
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)

    # ring of independent batches: rank r owns global env ids [r * n, (r + 1) * n) of every ring slot
    gen = torch.Generator(device=device).manual_seed(1 + rank)
    streams = [stream] + [torch.cuda.Stream(device) for _ in range(max(args.streams, 1) - 1)]
    ring = []
    for j in range(args.ring):
        e = make_env(g, env, n, dev_index, (j * world + rank) * n)
        e.set_stream(streams[j % len(streams)].cuda_stream)
        e.set_launch_config(vec=args.vec, block=args.block, pdl=args.pdl)
        e.set_launch_occupancy(args.wide)
        e.reset(seed=0)
        # ACTION_SETS pre-generated action batches per ring slot, used in rotation, so an env does
        # not see the same action at every step
        ring.append((e, [make_actions(torch, env, n, device, gen) for _ in range(ACTION_SETS)]))
    torch.cuda.synchronize(device)
    # launch schedule: consecutive steps go to consecutive ring slots; visit v of a slot uses action set v
    handles = [(e.handle, acts[v].data_ptr()) for v in range(ACTION_SETS) for e, acts in ring]
    resident_pool = [(ring[0][0].handle, a.data_ptr()) for a in ring[0][1]]
    AR = _capi.STEP_AUTORESET
    import ctypes as C
    step_many = L.gymrs_step_many
    _arrays = {}

    step_pass = L.gymrs_step_pass

    def run_steps(k, pool, begin_event=None):
        """k consecutive steps over the pool's (handle, action batch) pairs in order: gymrs_step_many, i.e.
        one gymrs_step per pair with one FFI crossing per pass over the pool (no Python between launches).
        begin_event (a raw cudaEvent_t): gymrs_step_pass records it on the first handle's stream and forks the
        other streams of the pass from it inside the same call, so that nothing but the launches follows it."""
        m = len(pool)
        if id(pool) not in _arrays:
            _arrays[id(pool)] = ((C.c_void_p * m)(*[h.value if hasattr(h, "value") else h for h, _ in pool]),
                                 (C.c_void_p * m)(*[a for _, a in pool]))
        hs, acts = _arrays[id(pool)]
        left = k
        while left > 0:
            c = min(m, left)
            if begin_event is not None:
                rc = step_pass(hs, acts, c, AR, begin_event, None, None)
                begin_event = None
            else:
                rc = step_many(hs, acts, c, AR, None)
            if rc:
                _capi.check(rc)
            left -= c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def join():    # the main stream's end event waits for the extra streams (they fork inside gymrs_step_pass)
        for s in streams[1:]:
            e_ = torch.cuda.Event()
            e_.record(s)
            stream.wait_event(e_)

    # Burn-in (untimed, not part of warm-up): all envs start from a synchronised reset, so their
    # first episodes end in a burst; a few hundred steps per slot decorrelate the episode phases
    # and bring the per-step reset rate to its stationary value.
    run_steps(args.burn_in * len(handles) // ACTION_SETS, handles)
    barrier()

    sanity = {"regions": 0, "device_ms": 0.0, "host_ms": 0.0}

    def timed(k, pool):
        """EXACTLY k steps between two CUDA events on the main stream, bracketed by a barrier + device
        synchronize on both sides.  Returns this rank's device time (ms); the max over ranks is taken
        once at the end, on the whole vector of repeats.  The host clock runs from the first launch
        to the drained device and excludes the barriers; it is only recorded (timing_sanity)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)  # creates the CUDA event; the record that counts is the one gymrs_step_pass makes below
        begin = C.c_void_p(ev0.cuda_event)
        barrier()
        w0 = time.perf_counter()
        # start event + fork of the other streams + the launches in ONE library call: the first pass of the pool
        # covers every ring slot, hence every stream (Python between the start event and the first launch would
        # be ~10 us of idle device inside a 130 us region)
        run_steps(k, pool, begin_event=begin)
        join()
        ev1.record(stream)
        torch.cuda.synchronize(device)
        w1 = time.perf_counter()
        barrier()
        ms = ev0.elapsed_time(ev1)
        sanity["regions"] += 1
        sanity["device_ms"] += ms
        sanity["host_ms"] += (w1 - w0) * 1e3
        return ms

    def over_ranks(times):
        """element-wise max over ranks of a list of per-repeat device times"""
        if world == 1:
            return list(times)
        t = torch.tensor(times, device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    run_steps(max(W, 3), handles)  # warm-up (untimed)
    barrier()
    # the timed region: EXACTLY K steps; repeated so the clock sampler sees sustained load, median reported
    est = over_ranks([timed(K, handles)])[0]
    # enough repeats for ~0.4 s under load (so that the 5 ms clock sampler sees it) whatever K is;
    # every repeat costs two barriers, so multi-rank runs are capped lower.  `est` is the max over
    # ranks, so every rank computes the same count (the barriers must pair up).
    repeats = int(min(2000 if world == 1 else 400, max(3, 400.0 / max(est, 1e-3))))
    with ClockSampler(dev_index) as cs:
        times = over_ranks([timed(K, handles) for _ in range(repeats)])
    clocks = cs.summary()
    ms = statistics.median(times)

    def one_stream(pdl, pool, reps=3, wide=args.wide):
        for e, _ in ring:
            e.set_stream(stream.cuda_stream)
            e.set_launch_config(vec=args.vec, block=args.block, pdl=pdl)
            e.set_launch_occupancy(wide)
        run_steps(len(pool), pool)
        t = statistics.median(over_ranks([timed(K, pool) for _ in range(reps)]))
        for e, _ in ring:
            e.set_launch_occupancy(args.wide)
        return t

    # Figures on ONE stream (what a caller with a single env group sees):
    #   default  -- pdl = 1, the library default: every launch waits for the whole previous grid.  Valid
    #               with a policy kernel in the loop that writes the actions right before the step.
    #   chained  -- pdl = 2: a step waits per CTA on the same CTA of the handle's previous step; valid
    #               when the actions are pre-generated (as here)
    # each on the cold ring (HBM) and on ONE batch stepped back to back (a true dependency chain, L2-resident)
    saved_streams = streams
    streams = [stream]
    default_cold = one_stream(1, handles)
    default_res = one_stream(1, resident_pool)
    chained = one_stream(2, handles)
    resident = one_stream(2, resident_pool)
    # the opt-in high-occupancy build (gymrs_set_launch_occupancy): meant for exactly this case, several
    # independent batches stepped round-robin on ONE stream
    wide_default_cold = one_stream(1, handles, wide=True)
    wide_chained = one_stream(2, handles, wide=True)
    streams = saved_streams
    for e, _ in ring:
        e.sync()  # surfaces any invalid-action / CUDA error from the timed launches

    value = world * n * K / (ms * 1e-3)
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES[env] * n * K / (ms * 1e-3) / 1e9  # per-GPU algorithmic GB/s

    def figure(t_ms, note):
        return {"value": world * n * K / (t_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": t_ms / K,
                "frac": ALGO_BYTES[env] * n * K / (t_ms * 1e-3) / 1e9 / peak, "note": note}

    # ---- e2e: same metric through the host-buffer entry points, pinned HOST memory -----------
    e2e = None
    if not args.no_e2e:
        e0 = ring[0][0]
        e0.set_launch_config(vec=args.vec, block=args.block, pdl=1)
        act_dtype = torch.float32 if env == "pendulum" else torch.int32
        S = 2  # action slots and result slots: step t's results stream out while step t + 1 is submitted
        h_acts = torch.stack([a.cpu().to(act_dtype) for a in ring[0][1][:S]]).pin_memory()
        h_obs = torch.empty((S, e0.obs_dim, n), dtype=torch.float32).pin_memory()
        h_rew = torch.empty((S, n), dtype=torch.float32).pin_memory()
        h_done = torch.empty((S, n), dtype=torch.uint8).pin_memory()
        acc = [0.0, 0]

        def consume(t, slot):
            # the device -> host read of the step's result: the consumer touches every delivered slot
            acc[0] += float(h_rew[slot, 0]) + float(h_obs[slot, 0, n - 1]) + float(h_done[slot, n // 2])
            acc[1] += 1

        def host_loop(steps, **kw):
            """gymrs_rollout_host: every step copies its actions in from pinned memory, runs the step
            kernel and copies observation / reward / done out to pinned memory; `consume` runs once
            per delivered step."""
            t0 = time.perf_counter()
            e0.rollout_host(h_acts, kw.get("obs", h_obs), h_rew, kw.get("done", h_done), None, n_steps=steps,
                            autoreset=True, on_step=consume, u8_actions=kw.get("u8", False),
                            packed_done=kw.get("packed", False))
            return time.perf_counter() - t0

        def max_over_ranks(x):
            if world == 1:
                return x
            t = torch.tensor([x], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        host_loop(4)
        barrier()
        dt = max_over_ranks(host_loop(args.e2e_steps))
        assert acc[1] == 4 + args.e2e_steps and acc[0] == acc[0]
        assert bool(torch.isfinite(h_obs).all()) and float(h_rew.abs().sum()) != 0.0
        # the synchronous call (what a scalar Env::step binding uses), for comparison
        barrier()
        t1 = time.perf_counter()
        for _ in range(10):
            e0.step_host(h_acts[0], h_obs[0], h_rew[0], h_done[0], None, autoreset=True)
        dt_sync = max_over_ranks((time.perf_counter() - t1) / 10)
        # the PCIe / host-memory ceiling of THIS box for the same bytes, all ranks copying at once
        h2d_b, d2h_b = 4 * n, D2H_BYTES[env] * n
        barrier()
        ceiling = pcie_ceiling(torch, dist, device, world, h2d_b, d2h_b)
        gbs = world * (h2d_b + d2h_b) * args.e2e_steps / dt / 1e9
        e2e = {"value": world * n * args.e2e_steps / dt, "unit": "env-steps/s",
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "steps": args.e2e_steps, "pcie_gbs": gbs,
               "pcie_ceiling_gbs": ceiling, "frac_of_pcie": gbs / ceiling,
               "ceiling_how": f"plain cudaMemcpyAsync of the same {h2d_b} B in + {d2h_b} B out per iteration from / to pinned "
                              f"memory on two streams, all {world} rank(s) at once, measured in this run",
               "host_cpus_bound_to_gpu_numa_node": numa,
               "synchronous_value": world * n / dt_sync,
               "api": "gymrs_rollout_host (the library's pipelined host loop: per step, actions in from a pinned slot; "
                      "observation, reward, done out to a pinned result slot; a consumer callback per delivered step); "
                      "synchronous_value = gymrs_step_host, one step at a time"}
        if env != "pendulum":
            # compact wire formats (lossless): uint8 actions in, done as bits out
            h_acts8 = h_acts.to(torch.uint8).pin_memory()
            h_bits = torch.empty((S, (n + 7) // 8), dtype=torch.uint8).pin_memory()
            keep = h_acts
            h_acts = h_acts8
            host_loop(4, u8=True, packed=True, done=h_bits)
            barrier()
            dtc = max_over_ranks(host_loop(args.e2e_steps, u8=True, packed=True, done=h_bits))
            h_acts = keep
            cb = n + (D2H_BYTES[env] - 1) * n + (n + 7) // 8
            e2e["compact"] = {"value": world * n * args.e2e_steps / dtc, "unit": "env-steps/s",
                              "bytes_per_step": cb, "pcie_gbs": world * cb * args.e2e_steps / dtc / 1e9,
                              "note": "GYMRS_HOST_U8_ACTIONS | GYMRS_HOST_PACKED_DONE: uint8 actions in, done as bits out "
                                      "(same information, fewer PCIe bytes); not the headline"}
        del h_acts, h_obs, h_rew, h_done

    # ---- fused rollout (labelled separately; never mixed with the single-step figure) ----------
    rollout = None
    if args.rollout_steps > 0:
        e0 = ring[0][0]
        kr = args.rollout_steps
        if env == "pendulum":
            acts = torch.rand((kr, n), generator=gen, device=device) * 4.0 - 2.0
        else:
            acts = torch.randint(0, 2 if env == "cartpole" else 3, (kr, n), generator=gen, device=device,
                                 dtype=torch.int32)
        obs_out = torch.empty((kr, e0.obs_dim, n), device=device)
        rew_out = torch.empty((kr, n), device=device)
        done_out = torch.empty((kr, n), device=device, dtype=torch.uint8)
        for _ in range(3):
            e0.rollout(acts, obs_out, rew_out, done_out, autoreset=True)
        reps = []
        for _ in range(7):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            ev0.record(stream)
            e0.rollout(acts, obs_out, rew_out, done_out, autoreset=True)
            ev1.record(stream)
            barrier()
            reps.append(ev0.elapsed_time(ev1))
        e0.sync()
        rms = statistics.median(over_ranks(reps))
        rbytes = ROLLOUT_BYTES[env] + 2 * 4 * e0.state_dim / kr
        wbytes = ROLLOUT_BYTES[env] - 4 + 4 * e0.state_dim / kr  # everything but the action row is a WRITE
        # The rollout is ~85 % writes, and write-only HBM traffic has a lower ceiling than the copy
        # the roofline peak is measured with: measure that ceiling here (1 GiB fill, best of 5).
        probe = torch.empty(1 << 30, dtype=torch.uint8, device=device)
        fill_ms = []
        for _ in range(5):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            probe.fill_(1)
            ev1.record(stream)
            torch.cuda.synchronize(device)
            fill_ms.append(ev0.elapsed_time(ev1))
        write_peak = (1 << 30) / (min(fill_ms) * 1e-3) / 1e9
        del probe
        wgbs = wbytes * n * kr / (rms * 1e-3) / 1e9
        rollout = {"value": world * n * kr / (rms * 1e-3), "unit": "env-steps/s", "steps_per_launch": kr,
                   "ms_per_launch": rms, "algorithmic_bytes_per_env_step": rbytes,
                   "write_gbs": wgbs, "hbm_write_only_gbs_measured": write_peak,
                   "frac_of_write_peak": wgbs / write_peak,
                   "achieved_gbs": rbytes * n * kr / (rms * 1e-3) / 1e9, "frac_of_copy_peak": rbytes * n * kr / (rms * 1e-3) / 1e9 / peak,
                   "api": "gymrs_rollout: one launch, state in registers, actions in / obs, reward, done out per step "
                          f"({kr * (ROLLOUT_BYTES[env]) * n / 1e6:.0f} MB streamed per launch, larger than L2); "
                          "write-dominated, so the primary figure is frac_of_write_peak (against the write-only HBM rate "
                          "measured in this run), the copy-peak fraction is given for context"}
        del acts, obs_out, rew_out, done_out

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample: 400 steps of the full 1M-env batch is ~1-2 s of wall time on 16 cores,
        # i.e. ~15-30 core-seconds of CPU work; cpu_rollout shrinks the batch if a slower host
        # would exceed the 20 s wall budget
        v, t, cores, sample, _ = cpu_rollout(env, n, 400, 5, budget_s=20.0)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": max(W, 3), "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(args, world),
            "method": {
                "launch": "one step kernel per step (gymrs_step semantics, enqueued through gymrs_step_many: one FFI "
                          "crossing per pass over the ring); the independent ring slots alternate over "
                          f"{len(streams)} CUDA stream(s) so consecutive launches overlap; pdl={args.pdl}"
                          + ("; high-occupancy build (--wide)" if args.wide else ""),
                "devices": f"rank r drives cuda:(r * {dev_stride}) "
                           f"of {visible} visible GPU(s): with fewer ranks than GPUs the ranks are spread over the box, "
                           "neighbouring GPUs share a PCIe root complex and its host-write bandwidth (sharding.device_for_rank)",
                "repeats": repeats, "timing": "CUDA events on the main stream (the other streams fork from the start "
                "event -- recorded, together with the fork and the launches, by one gymrs_step_pass call -- and join before the end event), barrier + device synchronize on both sides of every region, "
                "median of repeats of the element-wise max over ranks"},
            "timing_sanity": {"regions": sanity["regions"], "device_ms_total": sanity["device_ms"],
                              "host_ms_total": sanity["host_ms"],
                              "device_over_host": sanity["device_ms"] / max(sanity["host_ms"], 1e-9),
                              "note": "host clock from the first launch of a region to the drained device (barriers "
                                      "excluded); recorded, never asserted"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(env), "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES[env], "kernel": "step_kernel",
                         "kernel_us_isolated": kernel_isolated_us(env), "copy_only_floor_us": mode_floor(env)},
            "cpu_baseline": cpu,
            "rollout": rollout,
            "single_stream_default": {
                "cold_ring": figure(default_cold, "cold ring on ONE stream, pdl=1 (the library default): every launch "
                                    "waits for the whole previous grid -- what Env::step sees inside a policy loop"),
                "l2_resident": figure(default_res, "ONE 1M-env batch stepped back to back, pdl=1; the state stays in "
                                      "the 126 MB L2, so this is not an HBM number")},
            "single_stream_chained": figure(chained, "same cold ring on ONE stream with pipelined launches (pdl=2, per-CTA "
                                            "release/acquire flags instead of a grid-wide dependency; needs pre-generated actions)"),
            "single_stream_wide": {
                "default_cold_ring": figure(wide_default_cold, "cold ring on ONE stream, pdl=1, gymrs_set_launch_occupancy(1): the step "
                                            "kernel under a 32 / 40-register budget (more resident CTAs; same bits)"),
                "chained_cold_ring": figure(wide_chained, "cold ring on ONE stream, pdl=2, gymrs_set_launch_occupancy(1)")},
            "l2_resident": figure(resident, "ONE 1M-env batch stepped back to back on one stream, pdl=2 (a true "
                                  "dependency chain; the state stays in the 126 MB L2); not an HBM number"),
            "all_ms_per_step": [t / K for t in times][:40],
            "ms_per_step_min_max": [min(times) / K, max(times) / K],
        }
        emit(line)
    for e, _ in ring:
        e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    quiet_stdout()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    except BaseException as exc:  # noqa: BLE001 - the launcher's summary hides tracebacks: say why, last
        if isinstance(exc, SystemExit) and exc.code in (0, None):
            raise
        import traceback
        traceback.print_exc()
        sys.stderr.write(f"rank {dist_env()[0]}: bench.py failed: {type(exc).__name__}: {exc}\n")
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
