#!/usr/bin/env python
"""bench.py — env-steps/sec (collect + update) of the CartPole-v1 PPO hot path on B200.

Workload (BASELINE.json configs[1]): CartPole-v1 PPO, 4096 parallel envs PER GPU (weak scaling),
MLP policy/critic, 128-step rollout, 4 PPO epochs, 1 minibatch, device-resident env.step + GAE +
ppo_update kernels.  A "step" is one iteration = 128 vec-env steps + critic pass + GAE + 4 updates
= 4096*128 env-steps per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # CPU restatement of the reference path (oracle port)

One JSON line on rank 0 (see the contract in the task statement).  Timing: CUDA events per
iteration on the launching stream, L2 flushed (256 MB write) between timed iterations, max over
ranks; clocks sampled with nvidia-smi during the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS, T, EPOCHS, MINIBATCH = 4096, 128, 4, 1
FLAGS = ["--seed", "0", "--episode_length", str(T), "--ppo_epoch", str(EPOCHS), "--num_mini_batch", str(MINIBATCH),
         "--log_interval", "1000000", "--log_each_episode", "false"]
WORKLOAD = f"CartPole-v1 PPO, {N_ENVS} envs/GPU, T={T}, {EPOCHS} epochs x {MINIBATCH} minibatch, MLP 64x64"
METRIC = "env-steps/sec (collect+update), CartPole-v1 PPO"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index
        self.window = None   # (t0, t1) wall-clock bounds of the timed region

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        lines = self.lines
        scope = "process lifetime (sampler runs from before warm-up to after the timed region)"
        if self.window is not None:
            inside = [x for x in lines if self.window[0] <= x[0] <= self.window[1] + 0.05]
            if len(inside) >= 3:
                lines, scope = inside, "timed region"
        for _, ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(mx)] if mx else sm   # samples taken under load
        return {"sm_mhz": statistics.median(busy or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


def build_agent(rank, world, parity=False):
    import torch

    from openrl_b200.configs.config import create_config_parser
    from openrl_b200.envs.common import make
    from openrl_b200.modules.common import PPONet
    from openrl_b200.runners.common import PPOAgent

    cfg = create_config_parser().parse_args(FLAGS)
    cfg.quiet = True
    dev = f"cuda:{torch.cuda.current_device()}"
    env = make("CartPole-v1", env_num=N_ENVS, device=dev, env_index_offset=rank * N_ENVS)
    net = PPONet(env, cfg=cfg, device=dev)
    return cfg, env, net, PPOAgent(net, rank=rank, world_size=world)


def make_driver(cfg, env, net, agent, rank, world):
    from openrl_b200.algorithms.ppo import PPOAlgorithm
    from openrl_b200.buffers import NormalReplayBuffer
    from openrl_b200.drivers.onpolicy_driver import OnPolicyDriver

    trainer = PPOAlgorithm(cfg, net.module, agent_num=1, device=net.device)
    buf = NormalReplayBuffer(cfg, 1, env.observation_space, env.action_space, device=net.device)
    drv = OnPolicyDriver({"cfg": cfg, "num_agents": 1, "run_dir": None, "envs": env, "device": net.device}, trainer, buf,
                         agent, rank=rank, world_size=world, logger=None, callback=None)
    drv.reset_and_buffer_init()
    return drv


def gae_roofline(flush, steps=10):
    """GAE kernel alone at the C2 config shape (L2 flushed per launch) and at a >= 1 GB shape."""
    import torch

    from openrl_b200 import lib

    L = lib.load()
    out = {}
    for tag, Tn, B in (("config", T, N_ENVS), ("1GB", 128, 1 << 21)):
        dev = torch.device("cuda")
        r = torch.randn(Tn, B, device=dev); vp = torch.randn(Tn + 1, B, device=dev)
        m = (torch.rand(Tn + 1, B, device=dev) > 0.01).float(); act = torch.ones(Tn + 1, B, device=dev)
        vn = torch.tensor([0.3, 2.0, 0.5], device=dev); ret = torch.empty(Tn + 1, B, device=dev)
        adv = torch.empty(Tn, B, device=dev); st = torch.empty(8, dtype=torch.float64, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        call = lambda: lib.check(L.orl_gae(lib.ptr(r), lib.ptr(vp), lib.ptr(m), None, lib.ptr(act), lib.ptr(vp[Tn]), lib.ptr(vn),  # noqa: E731
                                           lib.ptr(ret), lib.ptr(adv), lib.ptr(st), Tn, B, 0.99, 0.95, 5, s), "gae")
        for _ in range(3):
            call()
        ts = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = statistics.mean(ts)
        bytes_per_el = 24  # rewards, value_preds, masks, active_masks in; returns, advantages out
        out[tag] = {"T": Tn, "B": B, "bytes_per_element": bytes_per_el, "avg_s": t, "GBps": Tn * B * bytes_per_el / t / 1e9}
        del r, vp, m, act, ret, adv
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from openrl_b200.utils.logger import Logger

    cfg, env, net, agent = build_agent(rank, world)
    drv = make_driver(cfg, env, net, agent, rank, world)
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")  # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        drv.device_iteration()
    barrier()
    drv.phase_events = []
    l0 = drv.gpu_launches + drv.trainer.gpu_launches
    events = []
    barrier()
    t_wall0 = time.time()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        drv.device_iteration()
        e1.record()
        events.append((e0, e1))
    barrier()
    sampler.window = (t_wall0, time.time())
    launches = drv.gpu_launches + drv.trainer.gpu_launches - l0
    total_s = sum(a.elapsed_time(b) for a, b in events) * 1e-3
    phases = {}
    for name, a, b in drv.phase_events:
        phases.setdefault(name, []).append(a.elapsed_time(b))
    drv.phase_events = None
    t = torch.tensor([total_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s = float(t.item())
    env_steps = N_ENVS * T * args.steps * world
    value = env_steps / total_s

    # ---- e2e through the public API (PPOAgent.train): host logging + D2H metric reads inside ----
    cfg2, env2, net2, agent2 = build_agent(rank, world)
    cfg2.log_interval = 1
    agent2.train(total_time_steps=N_ENVS * T * 3, logger=Logger(quiet=True))  # warm-up call
    barrier()
    h2d0 = agent2.driver.trainer.h2d_bytes
    d2h0 = (agent2.driver.trainer.d2h_bytes, getattr(env2, "d2h_bytes", 0))
    t0 = time.perf_counter()
    agent2.train(total_time_steps=N_ENVS * T * args.steps, logger=Logger(quiet=True))
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    tr2, dr2 = agent2.driver.trainer, agent2.driver
    h2d = (tr2.h2d_bytes - h2d0 + dr2.h2d_bytes) / args.steps
    d2h = (tr2.d2h_bytes - d2h0[0] + dr2.d2h_bytes + getattr(env2, "d2h_bytes", 0) - d2h0[1]) / args.steps

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, how = peaks()
    gae = gae_roofline(flush)
    gae_in_step = statistics.mean(phases.get("gae", [0.0])) * 1e-3
    upd_s = statistics.mean(phases.get("update", [0.0])) * 1e-3
    flops_update = 53e3 * N_ENVS * T * EPOCHS  # fwd+bwd of both nets, SURVEY.md §8d
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample(n_envs=256, iters=1)
    out = {
        "metric": "env-steps/sec (collect+update), CartPole-v1 PPO", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (policy/critic/losses; fc3 GEMMs as split-fp16 tcgen05 MMAs with fp32 accumulate = fp32-class accuracy), f64 (CartPole state)",
        "data": "synthetic: device-resident CartPole-v1, random-init nets, seed 0",
        "config": {"workload": WORKLOAD, "global_envs": N_ENVS * world, "rollout_T": T, "parallelism": f"env-shard dp{world}",
                   "l2": "256 MB L2 flush between timed iterations; inside an iteration the 25 MB buffer is re-read by design",
                   "sampling": "device Philox (fast mode); parity mode is covered by tests/"},
        "gpu_launches": launches,
        "phases_ms": {k: statistics.mean(v) for k, v in phases.items()},
        "e2e": {"value": env_steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "PPOAgent.train(total_time_steps) incl. construction, per-iteration logging reads"},
        "roofline": {"bound": "hbm", "kernel": "gae_scan_kernel (orl_gae), >=1 GB shape, L2 flushed", "achieved": gae["1GB"]["GBps"],
                     "peak": hbm_peak, "peak_source": how, "unit": "GB/s", "frac": gae["1GB"]["GBps"] / hbm_peak,
                     "algorithmic_bytes_per_launch": 128 * (1 << 21) * 24,
                     "traffic": 4.303432e9 + 2.126753e9, "traffic_source": "ncu --set full, profiles/r1_ncu_summary.md (dram read+write)",
                     "config_shape": {**gae["config"], "frac": gae["config"]["GBps"] / hbm_peak,
                                      "in_step_avg_s": gae_in_step,
                                      "in_step_GBps": N_ENVS * T * 24 / max(gae_in_step, 1e-12) / 1e9}},
        "update_kernel": {"kernel": "ppo_fwdbwd_tc_kernel (tcgen05, split-fp16 operands, fp32 accumulate)" if drv.trainer.use_tensor_cores else "ppo_fwdbwd_kernel (fp32 FFMA)",
                          "bound": "issue/latency (row-wise LayerNorm + loss work between three small MMAs per 128-row tile)",
                          "avg_s_per_iteration": upd_s, "algorithmic_flop_per_row": 53e3,
                          "achieved_tflops": flops_update / max(upd_s, 1e-12) / 1e12, "fp32_peak_tflops_at_clock": fp32_peak,
                          "frac_of_fp32_pipe_peak": flops_update / max(upd_s, 1e-12) / 1e12 / fp32_peak},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _best_cpu_threads(n_envs):
    """The port is Python-dispatch-bound: torch intra-op threads mostly add overhead on these tiny
    matrices.  Time one iteration at 1 thread and at the default count and keep the faster setting
    ("all the host threads it can use" = the count that makes it fastest)."""
    import torch

    from oracle import loop as oloop

    default = torch.get_num_threads()
    best, best_t = default, None
    for nt in sorted({1, min(default, 8), default}):
        torch.set_num_threads(nt)
        cfg = oloop.cfg_from_flags(" ".join(FLAGS))
        tr = oloop.Trainer(cfg, "CartPole-v1", n_envs)
        t0 = time.perf_counter()
        tr.iteration()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline_sample(n_envs=256, iters=1):
    """cpu_baseline leg: the unmodified reference (`kind: reference`, baseline/_ref through
    oracle/run_reference.py, SyncVectorEnv, 128 of the 4096 envs, all host cores) with the oracle port
    (oracle/loop.py) timed beside it as a second, labelled number."""
    from oracle import loop as oloop

    threads = _best_cpu_threads(n_envs)
    cfg = oloop.cfg_from_flags(" ".join(FLAGS))
    tr = oloop.Trainer(cfg, "CartPole-v1", n_envs)
    tr.iteration()  # warm-up
    t0 = time.perf_counter()
    for _ in range(iters):
        tr.iteration()
    dt = time.perf_counter() - t0
    port = {"value": n_envs * T * iters / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{iters} iteration(s) of {n_envs} envs x T={T}, {EPOCHS} epochs (oracle/loop.py, torch-CPU + numpy)"}
    if not reference_available():
        return {**port, "note": "baseline/_ref missing: port only"}
    try:
        ref = reference_run(REF_ENVS, 5, 3)
    except Exception as e:  # noqa: BLE001
        return {**port, "note": f"reference run failed ({str(e)[-200:]}): port only"}
    return {"value": ref["env_steps_per_s"], "unit": "env-steps/s", "cores": ref["torch_threads"], "kind": "reference",
            "sample": f"5 PPOAgent.train iterations (after 3 warm-up) of {REF_ENVS} of the {N_ENVS} envs x T={T}, {EPOCHS} epochs; unmodified "
                      f"reference from baseline/_ref, SyncVectorEnv, {ref['host_cores']} host cores", "port": port}


def reference_run(envs, iters, warmup, asynchronous=False, timeout=900):
    """One timed run of the UNMODIFIED reference (baseline/_ref, oracle/run_reference.py) in a fresh
    process: `PPOAgent.train` on the host cores, all of them (the child resets torchrun's OMP_NUM_THREADS=1)."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), "--env", "CartPole-v1", "--envs", str(envs),
           "--T", str(T), "--epochs", str(EPOCHS), "--minibatch", str(MINIBATCH), "--iters", str(iters), "--warmup", str(warmup)]
    if asynchronous:
        cmd.append("--async")
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError(f"reference run failed (rc {r.returncode}): {r.stderr[-800:]}")
    return json.loads(lines[-1])


def reference_available():
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "openrl", "__init__.py"))


REF_ENVS = 128   # north_star's target point; the reference's per-env Python loop makes cost linear in envs


def run_reference(args):
    """Reference arm: the unmodified reference through its own public API (`make` -> `PPONet` ->
    `PPOAgent.train`, SyncVectorEnv) on the box's host cores.  Each step = one iteration of a bounded
    sample (128 of the 4096 envs; a 4096-env reference iteration takes ~30 s); the AsyncVectorEnv
    ("SubprocVecEnv", one process per env) run and the oracle port are reported beside it."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 3)
    base = {"impl": "reference", "metric": METRIC, "unit": "env-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD}}
    if reference_available():
        sync = reference_run(REF_ENVS, steps, warmup, asynchronous=False)
        value, kind = sync["env_steps_per_s"], "reference"
        extra = {"sync_128": sync}
        try:   # the north_star's "SubprocVecEnv" path: 128 worker processes + the learner
            extra["async_128"] = reference_run(REF_ENVS, max(2, min(steps, 3)), 1, asynchronous=True)
        except Exception as e:  # noqa: BLE001
            extra["async_128"] = {"error": str(e)[-300:]}
        try:
            extra["sync_8_c1"] = reference_run(8, 5, 2, asynchronous=False)
            extra["async_8_c1"] = reference_run(8, 5, 2, asynchronous=True)
        except Exception as e:  # noqa: BLE001
            extra["c1"] = {"error": str(e)[-300:]}
        extra["not_run"] = "AsyncVectorEnv at 4096 envs = 4097 processes: infeasible on this host; Sync at 4096 envs is ~30 s/iteration"
        cores, sample = sync["torch_threads"], (f"each step = one PPOAgent.train iteration of {REF_ENVS} of the {N_ENVS} envs x T={T}, {EPOCHS} epochs, "
                                                 f"unmodified reference (baseline/_ref) + SyncVectorEnv, {sync['host_cores']} host cores")
        dt = sync["seconds"]
    else:
        from oracle import loop as oloop

        n_envs = 256
        threads = _best_cpu_threads(n_envs)
        cfg = oloop.cfg_from_flags(" ".join(FLAGS))
        tr = oloop.Trainer(cfg, "CartPole-v1", n_envs)
        for _ in range(warmup):
            tr.iteration()
        t0 = time.perf_counter()
        for _ in range(steps):
            tr.iteration()
        dt = time.perf_counter() - t0
        value, kind, cores, extra = n_envs * T * steps / dt, "port", threads, {"note": "baseline/_ref missing: oracle port timed instead"}
        sample = f"each step = one iteration of {n_envs} of the {N_ENVS} envs x T={T}, {EPOCHS} epochs (oracle/loop.py)"
    cb = {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample}
    print(json.dumps({**base, "value": value, "ms_per_step": dt / steps * 1e3, "cpu_baseline": cb, "reference_runs": extra,
                      "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="development: skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
