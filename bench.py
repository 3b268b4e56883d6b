#!/usr/bin/env python
"""bench.py - env-steps/sec of the batched env step (BASELINE.json metric: AntGather, 4096 envs/GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--env AntGather|AntMj|AntMaze|AntFlagrun|...]

One "step" = one batched `step()` of all envs of one GPU (4 physics sub-steps + task logic + observation).
Prints ONE JSON line (rank 0).  Timing protocol of the GPU arm:
  * untimed: W warm-up steps, then a settle phase (>= 40 steps in total) so that the ants have landed and the contact
    rows are active, exactly like the reference arm;
  * `value`: EXACTLY K steps, each bracketed by its own CUDA-event pair on the launch stream with the L2 flushed in
    between (256 MiB memset outside the pairs).  The (flush, event, step, event) quads are enqueued behind a
    device-side gate (hrl_stream_gate) in chunks and released at once, so no host-side launch gap can fall inside a
    pair; the sum of the K pairs is reduced with MAX over ranks;
  * `e2e`: the public host-buffer API (VecEnv.step(numpy) -> hrl_step_host), wall clock, median of 5 segments of K steps.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "env-steps/s"
# BASELINE.json configs: [1] AntMj 4096 x 1 GPU, [2] AntGather 4096/GPU (the metric), [3] AntMaze 4096/GPU, [4] AntFlagrun 16384/GPU
ENVS = {
    "AntGather": ("AntGatherBulletEnv-v0", 4096),
    "AntMj": ("AntMjBulletEnv-v0", 4096),
    "AntMaze": ("AntMazeBulletEnv-v0", 4096),
    "AntFlagrun": ("AntFlagrunBulletEnv-v0", 16384),
    "AntMazeMj": ("AntMazeMjEnv-v0", 4096),
    "PointGather": ("PointGatherBulletEnv-v0", 4096),
}
# untimed steps before any timed region (both arms).  The reset drops the ants 0.3 m (~15 steps airborne: no contact rows);
# the joint-limit rows take longer to reach their steady share under random actions (0.7 rows / sub-step after 40 steps,
# 2.6 after 200, measured), so both arms settle for 200 steps.
SETTLE_STEPS = 200
GATE_CHUNK = 32       # timed steps enqueued behind one gate (4 stream operations each: well inside the launch queue)
E2E_SEGMENTS = 5


def metric_name(env, n):
    return "env-steps/sec at %d %s envs/GPU" % (n, env)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


def load_ref_pybullet_mp():
    """bench/ref_pybullet_mp.py.  Imported under its own top-level name (a `bench` package would shadow this very file);
    its worker processes are spawned and must be able to import the module too, so its directory goes on sys.path."""
    d = os.path.join(ROOT, "bench")
    if d not in sys.path:
        sys.path.insert(0, d)
    import ref_pybullet_mp
    return ref_pybullet_mp


def median(xs):
    xs = sorted(xs)
    n = len(xs)
    return xs[n // 2] if n % 2 else 0.5 * (xs[n // 2 - 1] + xs[n // 2])


# ------------------------------------------------------------------------------------------
# CPU arm.  The real reference (pybullet in a multiprocessing vector env, bench/ref_pybullet_mp.py) runs when pybullet
# and gym are importable; they are not in this image (no network), so the arm is the oracle port, kind="port".
# ------------------------------------------------------------------------------------------
def cpu_port_rate(env_id, n_envs, seconds=12.0, threads=None):
    """env-steps/s of the C oracle (oracle/hrl_oracle.c, f64, the unoptimised checker) on `threads` host threads, on
    a bounded sample: n_envs envs stepped with random actions for ~`seconds` after the settle phase."""
    import numpy as np
    from oracle import oracle as O
    threads = threads or host_threads()
    env = O.OracleVecEnv.make(env_id, n_envs, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (8, n_envs, env.A)).astype(np.float32)
    for i in range(SETTLE_STEPS):
        env.step(acts[i % 8])
    t0 = time.perf_counter(); k = 0
    while True:
        env.step(acts[k % 8]); k += 1
        el = time.perf_counter() - t0
        if el >= seconds and k >= 3:
            break
    return n_envs * k / el, threads, "%d %s envs x %d steps (%.1f s), C oracle f64 (unoptimised checker), %d threads" % (
        n_envs, env_id, k, el, threads)


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    env_id, n_per_gpu = ENVS[args.env][0], args.envs_per_gpu
    threads = host_threads()
    base = {"impl": "reference", "metric": metric_name(args.env, n_per_gpu), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gpu_launches": 0}
    ref_pybullet_mp = load_ref_pybullet_mp()
    if ref_pybullet_mp.available():
        # the real thing: pybullet envs in worker processes (README.md:20-37 loop), one worker per host core
        r = ref_pybullet_mp.run(env_id, steps=args.steps, warmup=max(args.warmup, SETTLE_STEPS), workers=threads)
        base.update(value=r["value"], ms_per_step=r["ms_per_step"],
                    config={"workload": "%s, %d envs/GPU, U(-1,1) actions" % (env_id, n_per_gpu), "settle_steps": SETTLE_STEPS,
                            "note": "unmodified hrl_pybullet_envs + pybullet, %d worker processes x 1 env" % threads},
                    cpu_baseline={"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": r["sample"]},
                    e2e={"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(base))
        return
    import numpy as np
    from oracle import oracle as O
    # size the per-step sample so that warmup+steps finish in ~2 minutes
    probe_rate, _, _ = cpu_port_rate(env_id, 1024, seconds=3.0, threads=threads)
    total = max(args.steps + max(args.warmup, SETTLE_STEPS), 1)
    n = int(min(n_per_gpu * args.gpus, max(8 * threads, probe_rate * 100.0 / total)))
    n = max(threads, (n // threads) * threads)
    env = O.OracleVecEnv.make(env_id, n, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (16, n, env.A)).astype(np.float32)
    for i in range(max(args.warmup, SETTLE_STEPS)):   # same settle phase as the GPU arm
        env.step(acts[i % 16])
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(acts[i % 16])
    el = time.perf_counter() - t0
    value = n * args.steps / el
    sample = "%d of %d %s envs per step (bounded sample), %d steps" % (n, n_per_gpu * args.gpus, args.env, args.steps)
    base.update(value=value, ms_per_step=1e3 * el / max(args.steps, 1),
                config={"workload": "%s, %d envs/GPU, U(-1,1) actions" % (env_id, n_per_gpu), "settle_steps": SETTLE_STEPS,
                        "note": "reference unavailable: pybullet not installed (%s); this arm is OUR OWN CPU restatement "
                                "(oracle/hrl_oracle.c, the unoptimised f64 checker), not pybullet" % ref_pybullet_mp.why_unavailable()},
                cpu_baseline={"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under the bench's load (B200_PROFILING.md).  Started
    before the warm-up; `mark()` brackets the loaded window whose samples are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def lines(self):
        try:
            with open(self.f.name) as g:
                return sum(1 for _ in g)
        except OSError:
            return 0

    def stop(self, first_line=0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        with open(self.f.name) as g:
            for ln, line in enumerate(g):
                if ln < first_line:
                    continue
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from hrl_pybullet_envs_b200 import VecEnv, _cabi, roofline
    from hrl_pybullet_envs_b200.sharding import gather_over_ranks, max_over_ranks, shard_offset

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the ONE JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sampler = ClockSampler(local) if rank == 0 else None   # nvidia-smi needs ~0.3 s to deliver its first sample
    env_id = ENVS[args.env][0]
    N = args.envs_per_gpu
    K = args.steps
    L = _cabi.lib()
    # shard rule: GPU r owns global envs [r*N, (r+1)*N); no collective on the step path
    env = VecEnv(env_id, N, device=local, seed=0, env_index_offset=shard_offset(rank, N))
    A, D = env.A, env.D
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    ring = torch.rand(64, N, A, generator=g, device=dev) * 2 - 1          # synthetic U(-1,1) actions, pre-generated
    ring_host = ring.cpu().pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gate = torch.zeros(16, dtype=torch.int32).pin_memory()                 # host word the device-side gate spins on
    gate_np = gate.numpy()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed: warm-up + settle (ants on the ground, contacts active), and the clock sampler's first sample
    settle = max(args.warmup, SETTLE_STEPS)
    for i in range(settle):
        env.step(ring[i % 64])
    torch.cuda.synchronize()
    t_wait = time.perf_counter()
    while sampler is not None and sampler.lines() == 0 and time.perf_counter() - t_wait < 3.0:
        for i in range(50):                      # keep the GPU under the same load while nvidia-smi starts
            env.step(ring[i % 64])
        torch.cuda.synchronize()
    first_line = sampler.lines() if sampler else 0
    env.stats(reset=True)
    barrier()

    # ---- timed region 1 (`value`): device-resident inputs, one event pair per step, L2 flushed between steps, queued
    # behind a device-side gate so that the pairs hold no host launch gaps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = L.hrl_launch_count()
    seq = 0
    gate_launches = 0
    barrier()
    for c0 in range(0, K, GATE_CHUNK):
        seq += 1
        _cabi.check(L.hrl_stream_gate(C.c_void_p(gate.data_ptr()), seq, 2_000_000_000, stream))
        gate_launches += 1
        for i in range(c0, min(c0 + GATE_CHUNK, K)):
            flush.zero_()
            ev[i][0].record()
            env.step(ring[i % 64])
            ev[i][1].record()
        gate_np[0] = seq                         # open the gate: the chunk runs back to back on the device
    barrier()
    launches = L.hrl_launch_count() - l0 - gate_launches
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(per_step_ms)
    stats = env.stats(reset=True)

    # ---- timed region 2: back-to-back launches without the flush (state stays in L2), one event pair, gated too
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    b2b_ms = 0.0
    for c0 in range(0, K, 4 * GATE_CHUNK):
        seq += 1
        _cabi.check(L.hrl_stream_gate(C.c_void_p(gate.data_ptr()), seq, 2_000_000_000, stream))
        e0.record()
        for i in range(c0, min(c0 + 4 * GATE_CHUNK, K)):
            env.step(ring[i % 64])
        e1.record()
        gate_np[0] = seq
        torch.cuda.synchronize()
        b2b_ms += e0.elapsed_time(e1)
    barrier()

    # ---- timed region 2b: the same K steps as ONE CUDA-graph launch per 32-step rollout (VecEnv.capture_rollout)
    graph_ms = None
    if not args.no_graph:
        T = min(32, K)
        graph, buf = env.capture_rollout(ring[:T].contiguous())
        graph.replay(); torch.cuda.synchronize()
        reps = max(K // T, 1)
        barrier()
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        barrier()
        graph_ms = e0.elapsed_time(e1) / (reps * T)
        del graph, buf

    # ---- timed region 2c (consumer side, informational): T = 32 steps WITH a policy - the fused rollout (in-kernel
    # 46-64-64-8 tanh MLP, one launch per 32 steps) against the same MLP in torch followed by step(), per step
    policy = None
    if A == 8 and not args.no_graph:
        gp = torch.Generator(device=dev).manual_seed(7)
        def lin(o, i, sc):
            return ((torch.rand(o, i, generator=gp, device=dev) * 2 - 1) * sc, (torch.rand(o, generator=gp, device=dev) * 2 - 1) * 0.1)
        layers = (lin(64, D, 0.3), lin(64, 64, 0.2), lin(8, 64, 0.3))
        T = 32
        rbuf = env.rollout_buffer(T)
        env.rollout_mlp(layers, T, buf=rbuf); torch.cuda.synchronize()
        reps = max(min(K, 640) // T, 2)
        barrier()
        e0.record()
        for _ in range(reps):
            env.rollout_mlp(layers, T, buf=rbuf, sigma=0.1, noise_seed=1)
        e1.record()
        barrier()
        fused_us = e0.elapsed_time(e1) * 1e3 / (reps * T)
        obs_t = env.observe()
        for _ in range(3):                       # cuBLAS handle / kernel-module loading happens here, not in the timed loop
            x = obs_t
            for W, b in layers:
                x = torch.tanh(torch.addmm(b, x, W.t()))
        barrier()
        e0.record()
        for _ in range(reps * T):
            x = obs_t
            for W, b in layers:
                x = torch.tanh(torch.addmm(b, x, W.t()))
            obs_t, _, _, _ = env.step(x)
        e1.record()
        barrier()
        policy = {"fused_rollout_us_per_step": fused_us, "torch_mlp_plus_step_us_per_step": e0.elapsed_time(e1) * 1e3 / (reps * T),
                  "policy": "%d-64-64-8 tanh MLP, sigma 0.1" % D, "steps_per_launch": T,
                  "env_steps_per_s_fused": N * world / (fused_us * 1e-6)}

    # ---- timed region 3 (e2e): the public API with HOST buffers, per step: actions in pinned host memory -> device,
    # kernel, obs/rew/done/info -> pinned host memory, completion visible to the host.  Both transfer modes of
    # hrl_step_host; the default ("auto" = zero-copy for pinned buffers) is the headline.  Median of 5 segments of K steps.
    acts_np = ring_host.numpy()
    # a segment holds at least 100 steps: at the driver's K = 20 the barrier before and the synchronise after a 1.3 ms
    # segment are 4-5 % of it (59.2e6 against 62.1e6 env-steps/s in a 1000-step run of the same build)
    KE = max(K, 100)
    e2e_modes = {}
    for mode in ("copy", "zerocopy"):
        env.set_host_mode(mode)
        for i in range(10):
            env.step(acts_np[i])
        segs = []
        for sgm in range(E2E_SEGMENTS):
            barrier()
            t0 = time.perf_counter()
            for i in range(KE):
                obs, rew, done, info = env.step(acts_np[i % 64])
            torch.cuda.synchronize()
            segs.append((time.perf_counter() - t0) * K / KE)   # scaled to K steps: everything downstream is per K
        e2e_modes[mode] = segs
    env.set_host_mode("auto")
    barrier()
    clocks = sampler.stop(first_line) if sampler else None

    # max over ranks per quantity (for the e2e medians: the slowest rank's median)
    red = max_over_ranks([dev_ms, b2b_ms, median(e2e_modes["zerocopy"]) * 1e3, median(e2e_modes["copy"]) * 1e3,
                          graph_ms or 0.0], device=dev)
    dev_ms_max, b2b_ms_max, e2e_ms, e2e_copy_ms, graph_ms_max = red
    # per-rank view of the per-step event pairs (microseconds): min / median / max and the slowest rank
    mine = [min(per_step_ms) * 1e3, median(per_step_ms) * 1e3, max(per_step_ms) * 1e3, dev_ms * 1e3 / K]
    per_rank = gather_over_ranks(mine, device=dev)
    # optional episode statistics over NVLink (the only other collective; not on the step path)
    ep_stats = env.episode_stats(aggregate=True)   # in-kernel accumulators, summed over ranks

    if rank == 0:
        hbm_peak, sm_max, which = peaks()
        total_envs = N * world
        value = total_envs * K / (dev_ms_max * 1e-3)
        launch_s = dev_ms * 1e-3 / K                      # rank 0's mean launch duration (the roofline is per launch)
        kind = env.kind
        bytes_step = roofline.bytes_per_env_step(kind, D, A)
        flops = roofline.flop_per_env_step(stats["contacts_per_substep"], stats["limit_rows_per_substep"]) if A == 8 else None
        sm_clk = (clocks or {}).get("sm_mhz") or sm_max
        cpu_v = cpu_c = cpu_s = None
        if world == 1 and not args.skip_cpu:
            cpu_v, cpu_c, cpu_s = cpu_port_rate(env_id, N, seconds=10.0)
        default_n = ENVS[args.env][1]
        prof = (profile_json("r2_counts_%s.json" % args.env) or {}) if N == default_n else {}
        slow = max(range(world), key=lambda r: per_rank[r][3])
        out = {
            "metric": metric_name(args.env, N) + ("" if N == default_n else " (exploration)"),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, %d envs/GPU, obs[%d,%d], act[%d,%d], U(-1,1) actions from a 64-batch device ring, auto-reset on"
                                   % (env_id, N, N, D, N, A),
                       "settle_steps": settle,
                       "l2": "flushed between timed steps (256 MiB memset, outside the per-step event pairs)",
                       "timing": "sum of K per-step CUDA-event pairs on the launch stream, queued behind a device-side gate in "
                                 "chunks of %d steps (no host launch gap inside a pair), max over ranks" % GATE_CHUNK,
                       "per_rank_step_us": {"min": [round(r[0], 2) for r in per_rank], "median": [round(r[1], 2) for r in per_rank],
                                            "max": [round(r[2], 2) for r in per_rank], "mean": [round(r[3], 2) for r in per_rank],
                                            "slowest_rank": slow},
                       "ms_per_step_back_to_back": b2b_ms_max / K,
                       "value_back_to_back": total_envs * K / (b2b_ms_max * 1e-3),
                       "ms_per_step_cuda_graph": graph_ms_max if graph_ms is not None else None,
                       "policy_rollout": policy,
                       "parity": "task layer pinned on reference-executed fixtures; physics vs our own f64 oracle port only "
                                 "(pybullet absent: parity unpinned against real Bullet)"},
            "e2e": {"value": total_envs * K / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": N * A * 4, "d2h_bytes_per_step": N * (D * 4 + 4 + 1),
                    "api": "VecEnv.step(numpy) -> hrl_step_host, zero-copy mode: the kernel reads the actions from and writes "
                           "obs/rew/done to pinned host memory over PCIe (info stays on the device until somebody reads it); the host polls a completion word the last CTA publishes",
                    "segments": "median of %d segments of %d steps, wall clock, slowest rank" % (E2E_SEGMENTS, KE),
                    "segment_values": [total_envs * K / s for s in e2e_modes["zerocopy"]],
                    "value_copy_mode": total_envs * K / (e2e_copy_ms * 1e-3),
                    "copy_mode": "pinned H2D memcpy, kernel, ONE packed D2H memcpy, stream sync"},
            "gpu_launches": int(launches),
            "episode_stats": ep_stats,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": bytes_step * N / launch_s / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": bytes_step * N / launch_s / 1e9 / hbm_peak,
                         "traffic": prof.get("dram_bytes_per_launch"),
                         "bytes_per_env_step": bytes_step,
                         "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                         "note": "state traffic is not the binding resource; see roofline_fp32"},
        }
        if flops is not None:
            fp = {"bound": "fp32", "achieved": flops * N / launch_s / 1e12, "peak": roofline.fp32_peak_tflops(sm_max),
                  "unit": "TFLOP/s", "frac": flops * N / launch_s / 1e12 / roofline.fp32_peak_tflops(sm_max),
                  "peak_at_sampled_clock": roofline.fp32_peak_tflops(sm_clk),
                  "algorithmic_flop_per_env_step": flops, "contacts_per_substep": stats["contacts_per_substep"],
                  "limit_rows_per_substep": stats["limit_rows_per_substep"],
                  "model": "minimum of the kernel's own formulation (block-arrow elimination, hrl_pybullet_envs_b200/roofline.py); "
                           "oracle_counted_flop_per_env_step is the instrumented count of the f64 scalar restatement "
                           "(13-link ABA + one ABA impulse response per row) under the same contact statistics"}
            om = profile_json("r2_oracle_flop_model.json")
            if om:
                rows_ = 3.0 * stats["contacts_per_substep"] + stats["limit_rows_per_substep"]
                fp["oracle_counted_flop_per_env_step"] = 4 * (om["per_env_substep"] + rows_ * (om["per_row"] + 5 * om["per_row_iteration"])) + om["per_env_step_task_layer"]
            if prof.get("executed_fp32_flop_per_launch"):
                # executed thread-level fp32 operations of the committed ncu capture (same regime: settled ants, same batch)
                ex = prof["executed_fp32_flop_per_launch"] / prof.get("envs_per_launch", N)
                fp.update(executed_flop_per_env_step=ex, executed_over_algorithmic=ex / flops,
                          executed_frac=ex * N / launch_s / 1e12 / roofline.fp32_peak_tflops(sm_max),
                          executed_source=prof.get("source"))
            out["roofline_fp32"] = fp
        if prof.get("warp_instructions_per_launch"):
            wi = prof["warp_instructions_per_launch"]
            issue_peak = 148 * 4 * sm_max * 1e6
            out["roofline_issue"] = {"bound": "issue", "achieved": wi / launch_s / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
                                     "frac": wi / launch_s / issue_peak, "warp_instructions_per_launch": wi,
                                     "source": prof.get("source")}
        if cpu_v is not None:
            out["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cpu_c, "kind": "port", "sample": cpu_s}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default="AntGather", choices=sorted(ENVS),
                    help="BASELINE.json configs: AntGather 4096/GPU is the metric (default); AntMj, AntMaze 4096/GPU, AntFlagrun 16384/GPU")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="omit the CUDA-graph leg")
    ap.add_argument("--envs-per-gpu", type=int, default=None,
                    help="exploration only: the BASELINE.json configs are quoted at the defaults (4096; Flagrun 16384)")
    args = ap.parse_args()
    if args.envs_per_gpu is None:
        args.envs_per_gpu = ENVS[args.env][1]
    # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 behind Python's back (the NCCL
    # version banner does) are pointed at stderr; Python's own sys.stdout keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
