#!/usr/bin/env python
"""bench.py - env-steps/sec of the batched AntGather step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batched `step()` of 4096 AntGather envs per GPU (4 physics sub-steps + task
logic + observation).  Prints ONE JSON line (rank 0).  See the module docstring of each helper
for what is inside the timed region.
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

ENV_ID = "AntGatherBulletEnv-v0"
ENVS_PER_GPU = 4096
METRIC = "env-steps/sec at 4096 AntGather envs/GPU"
UNIT = "env-steps/s"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/dram_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None


def ncu_warp_instructions():
    """Warp-instructions per launch of the dominant kernel from the same committed capture, or None."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        return json.load(open(p))["warp_instructions_per_launch"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (pybullet is not installable in this image)
# ------------------------------------------------------------------------------------------
def cpu_port_rate(seconds=12.0, n_envs=ENVS_PER_GPU, threads=None):
    """env-steps/s of the C oracle (oracle/hrl_oracle.c, f64) on `threads` host threads, on a
    bounded sample: n_envs AntGather envs stepped with random actions for ~`seconds`."""
    import numpy as np
    from oracle import oracle as O
    threads = threads or host_threads()
    env = O.OracleVecEnv.make(ENV_ID, n_envs, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (8, n_envs, 8)).astype(np.float32)
    for i in range(40):   # land first (see run_reference)
        env.step(acts[i % 8])
    t0 = time.perf_counter(); k = 0
    while True:
        env.step(acts[k % 8]); k += 1
        el = time.perf_counter() - t0
        if el >= seconds and k >= 3:
            break
    return n_envs * k / el, threads, "%d AntGather envs x %d steps (%.1f s), C oracle f64, %d threads" % (n_envs, k, el, threads)


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores.  pybullet / gym are
    not installed (no network), so the arm is the oracle port, labelled kind="port"."""
    import importlib.util
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    threads = host_threads()
    have_pb = importlib.util.find_spec("pybullet") is not None and importlib.util.find_spec("gym") is not None
    # size the per-step sample so that warmup+steps finish in ~2 minutes
    probe_rate, _, _ = cpu_port_rate(seconds=3.0, n_envs=1024, threads=threads)
    total = max(args.steps + args.warmup, 1)
    n = int(min(ENVS_PER_GPU * args.gpus, max(8 * threads, probe_rate * 100.0 / total)))
    n = max(threads, (n // threads) * threads)
    env = O.OracleVecEnv.make(ENV_ID, n, seed=0, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, (16, n, 8)).astype(np.float32)
    # like the GPU arm's warm-up: let the ants land first (reset drops them 0.3 m, ~15 steps without any contact row),
    # so that even a short --steps/--warmup run times steps with the contacts active
    for i in range(max(args.warmup, 40)):
        env.step(acts[i % 16])
    t0 = time.perf_counter()
    for i in range(args.steps):
        env.step(acts[i % 16])
    el = time.perf_counter() - t0
    value = n * args.steps / el
    sample = "%d of %d AntGather envs per step (bounded sample), %d steps" % (n, ENVS_PER_GPU * args.gpus, args.steps)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "AntGatherBulletEnv-v0, %d envs/GPU, U(-1,1) actions" % ENVS_PER_GPU,
                   "note": "pybullet %s in this image; CPU restatement (oracle port), not pybullet" % ("present" if have_pb else "absent")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
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

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
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
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hrl_pybullet_envs_b200 import VecEnv, _cabi, roofline
    from hrl_pybullet_envs_b200.sharding import max_over_ranks, shard_offset

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the ONE JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.envs_per_gpu
    # shard rule: GPU r owns global envs [r*N, (r+1)*N); no collective on the step path
    env = VecEnv(ENV_ID, N, device=local, seed=0, env_index_offset=shard_offset(rank, N))
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    ring = torch.rand(64, N, 8, generator=g, device=dev) * 2 - 1          # synthetic U(-1,1) actions, pre-generated
    ring_host = ring.cpu().pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also settles the ants onto the ground: contacts are active in the timed region)
    for i in range(max(args.warmup, 3)):
        env.step(ring[i % 64])
    env.stats(reset=True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- timed region 1: device-resident inputs, one event pair per step, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = _cabi.lib().hrl_launch_count()
    barrier()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        env.step(ring[i % 64])
        ev[i][1].record()
    barrier()
    launches = _cabi.lib().hrl_launch_count() - l0
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(per_step_ms)
    stats = env.stats(reset=True)
    # ---- timed region 2: back-to-back launches without the flush (state stays in L2), one event pair
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        env.step(ring[i % 64])
    e1.record()
    barrier()
    b2b_ms = e0.elapsed_time(e1)
    # ---- timed region 3 (e2e): the public API with HOST buffers, per step: actions in pinned host memory ->
    # device, kernel, obs/rew/done/info -> pinned host memory, stream sync.  Measured in both transfer modes
    # of hrl_step_host; the default ("auto" = zero-copy for pinned buffers) is the headline.
    acts_np = ring_host.numpy()
    e2e_modes = {}
    for mode in ("copy", "zerocopy"):
        env.set_host_mode(mode)
        for i in range(5):
            env.step(acts_np[i])
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            obs, rew, done, info = env.step(acts_np[i % 64])
        torch.cuda.synchronize()
        e2e_modes[mode] = time.perf_counter() - t0
        barrier()
    env.set_host_mode("auto")
    e2e_s = e2e_modes["zerocopy"]
    clocks = sampler.stop() if sampler else None

    dev_ms, b2b_ms, e2e_ms, e2e_copy_ms = max_over_ranks([dev_ms, b2b_ms, e2e_s * 1e3, e2e_modes["copy"] * 1e3], device=dev)
    # optional episode statistics over NVLink (the only other collective; not on the step path)
    ep_stats = env.episode_stats(aggregate=True)   # in-kernel accumulators, summed over ranks

    if rank == 0:
        hbm_peak, sm_max, which = peaks()
        total_envs = N * world
        value = total_envs * args.steps / (dev_ms * 1e-3)
        # roofline of the dominant kernel (ant_env_kernel<0>), per launch, rank 0
        launch_s = dev_ms * 1e-3 / args.steps
        flops = roofline.flop_per_env_step(stats["contacts_per_substep"], stats["limit_rows_per_substep"])
        sm_clk = (clocks or {}).get("sm_mhz") or sm_max
        cpu_v, cpu_c, cpu_s = cpu_port_rate(seconds=10.0) if (world == 1 and not args.skip_cpu) else (None, None, None)
        out = {
            "metric": METRIC if N == ENVS_PER_GPU else "env-steps/sec at %d AntGather envs/GPU (exploration)" % N, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "AntGatherBulletEnv-v0, %d envs/GPU, obs[%d,46], act[%d,8], U(-1,1) actions from a 64-batch device ring, auto-reset on" % (N, N, N),
                       "l2": "flushed between timed steps (256 MiB memset, outside the per-step event pairs)",
                       "timing": "sum of per-step CUDA-event pairs on the launch stream, max over ranks",
                       "ms_per_step_back_to_back": b2b_ms / args.steps,
                       "value_back_to_back": total_envs * args.steps / (b2b_ms * 1e-3)},
            "e2e": {"value": total_envs * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": N * 8 * 4, "d2h_bytes_per_step": N * (46 * 4 + 4 + 1 + 16),
                    "api": "VecEnv.step(numpy) -> hrl_step_host, zero-copy mode: the kernel reads the actions from and writes "
                           "obs/rew/done/info to pinned host memory over PCIe; the host polls a completion word the last CTA publishes",
                    "value_copy_mode": total_envs * args.steps / (e2e_copy_ms * 1e-3),
                    "copy_mode": "pinned H2D memcpy, kernel, ONE packed D2H memcpy, stream sync"},
            "gpu_launches": int(launches),
            "episode_stats": ep_stats,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": roofline.BYTES_PER_ENV_STEP * N / launch_s / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": roofline.BYTES_PER_ENV_STEP * N / launch_s / 1e9 / hbm_peak,
                         "traffic": ncu_traffic() if N == ENVS_PER_GPU else None,
                         "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                         "note": "state traffic is not the binding resource; see roofline_fp32"},
            "roofline_fp32": {"bound": "fp32", "achieved": flops * N / launch_s / 1e12, "peak": roofline.fp32_peak_tflops(sm_max),
                              "unit": "TFLOP/s", "frac": flops * N / launch_s / 1e12 / roofline.fp32_peak_tflops(sm_max),
                              "peak_at_sampled_clock": roofline.fp32_peak_tflops(sm_clk),
                              "flop_per_env_step": flops, "contacts_per_substep": stats["contacts_per_substep"],
                              "limit_rows_per_substep": stats["limit_rows_per_substep"]},
        }
        wi = ncu_warp_instructions() if N == ENVS_PER_GPU else None
        if wi:
            # third view of the same launch: share of the machine's instruction-issue slots (4 schedulers x 148 SMs, one
            # warp-instruction per cycle each) - the kernel is integer / branch / shared-memory work around packed FMAs
            issue_peak = 148 * 4 * sm_max * 1e6
            out["roofline_issue"] = {"bound": "issue", "achieved": wi / launch_s / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
                                     "frac": wi / launch_s / issue_peak, "warp_instructions_per_launch": wi,
                                     "source": "ncu smsp__inst_executed.sum of the committed capture (profiles/), launch time measured live"}
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
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU,
                    help="exploration only: the BASELINE.json metric is quoted at the default 4096")
    args = ap.parse_args()
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
