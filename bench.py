#!/usr/bin/env python
"""bench.py -- the DTQN hot path on N B200s (one process per GPU) and the reference-CPU arm.

    python bench.py --gpus 1 --steps K --warmup W                  # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W # CPU port of the reference loop (oracle/loop.py)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                      # N ranks over NCCL

A "step" is one iteration of the reference's training loop body (run.py:290-298) over one lockstep batch of
environments: epsilon-greedy action through the DTQN (acting forward over the 50-token context of every env), env.step
+ TimeLimit + replay/context append (+ episode roll), then ONE DtqnAgent.train() (sample 32 windows per rank, 3 forwards,
TD loss, backward, gradient allreduce, clip, Adam).  Workload = BASELINE.json configs[1]: DiscreteCarFlag-v0, 4096
batched envs per GPU, ctx=50, in-embed=64, batch 32.  metric value = env-steps/s of the whole job; grad-steps/s is
reported beside it (same timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_ID = "DiscreteCarFlag-v0"
CTX, EMBED, HEADS, LAYERS, BATCH = 50, 64, 8, 2, 32
FWD_FLOP_PER_TOKEN = 231_168          # SURVEY.md section 8d, CarFlag d=64 L=50


# stdout carries exactly ONE JSON line: everything else a library prints there (NCCL banners, warnings) goes to stderr
_REAL_STDOUT = None


def _quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="lockstep envs per GPU")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--workload", default="carflag", choices=["carflag", "memory", "mixed128"],
                    help="carflag: BASELINE.json configs[1] (the headline, with a short configs[2] sub-record); memory: configs[2] "
                         "(Memory-5-v0, in-embed 128); mixed128: configs[4] (CarFlag + Memory-5 groups, ctx=128)")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the Memory-5 sub-record of the default workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.workload == "memory":
        global ENV_ID, EMBED, FWD_FLOP_PER_TOKEN
        ENV_ID, EMBED, FWD_FLOP_PER_TOKEN = "Memory-5-v0", 128, 893_440      # SURVEY.md section 8d, Memory d=128 L=50
    return args


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
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
        for line in self.f.read().strip().splitlines():
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
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def best_thread_count(lp):
    """The reference loop is made of tiny CPU ops; on a many-core host the fastest torch thread count is usually far
    below the core count.  Try a few, keep the best (this is the reference arm's best configuration, not a handicap)."""
    import torch
    cores = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for nt in sorted({1, 2, 4, 8, 16, 32, cores}):
        if nt > cores:
            continue
        torch.set_num_threads(nt)
        lp.iteration()
        t0 = time.perf_counter()
        for _ in range(6):
            lp.iteration()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = nt, dt
        if dt > 6.0:           # already hopeless at this thread count; larger counts only get worse
            break
    torch.set_num_threads(best)
    return best


REF_PREPOP = 50_000        # run.py:495
REF_MIN_ITERS = 200        # >= 200 iterations per leg
REF_EPS = 0.1              # exploration floor (run.py:420): the regime the 2M-step run spends ~90 % of its time in


def time_reference(batch, min_iters=REF_MIN_ITERS, budget_s=12.0):
    """The reference's CPU path on this host: the UNMODIFIED reference through oracle/ref_harness when /root/reference
    exists (kind "reference"), else the oracle port of the same loop (kind "port").  Legs: (i)
    run.prepopulate random-policy env-steps/s, (ii) run.step eps-greedy env-steps/s, (iii) agent.train() grad-steps/s,
    (iv) the loop body (1 env-step + 1 grad-step), all at the best torch thread count, plus (iv) at 1 thread."""
    import torch
    from oracle import ref_loop
    if ref_loop.available():
        lp, kind = ref_loop.RealReferenceLoop(ENV_ID, seed=1, inner_embed=EMBED, heads=HEADS, layers=LAYERS, context=CTX,
                                              batch=batch), "reference"
    else:
        from oracle.loop import ReferenceLoop
        lp, kind = ReferenceLoop(ENV_ID, seed=1, inner_embed=EMBED, heads=HEADS, layers=LAYERS, context=CTX, batch=batch), "port"
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    lp.prepopulate(REF_PREPOP)
    prepop_sps = REF_PREPOP / (time.perf_counter() - t0)
    lp.set_epsilon(REF_EPS)

    def leg(fn, n_min, budget):
        for _ in range(5):
            fn()
        t0 = time.perf_counter(); n = 0
        while n < n_min or time.perf_counter() - t0 < budget:
            fn(); n += 1
            if time.perf_counter() - t0 > 6 * budget + 5:
                break
        return n / (time.perf_counter() - t0), n

    one_thread, n1 = leg(lp.iteration, 50, 2.0)
    cores = best_thread_count(lp)
    loop_ps, n_loop = leg(lp.iteration, min_iters, budget_s)
    step_ps, _ = leg(lp.step_only, min_iters, 1.0)
    train_ps, _ = leg(lp.train_only, min_iters, 2.0)
    return {"value": loop_ps, "unit": "env-steps/s", "grad_steps_per_sec": loop_ps, "cores": cores,
            "host_cores": os.cpu_count(), "kind": kind,
            "sample": (f"{n_loop} iterations of the 1-env reference loop body (run.py:290-298: 1 eps-greedy env-step at eps={REF_EPS} + "
                       f"1 grad-step, batch {batch}) after {REF_PREPOP} prepopulate steps, {cores} torch threads"),
            "one_thread_value": one_thread, "one_thread_iterations": n1,
            "legs": {"prepopulate_env_steps_per_sec_1thread": prepop_sps, "run_step_env_steps_per_sec": step_ps,
                     "agent_train_grad_steps_per_sec": train_ps, "loop_iterations_per_sec": loop_ps}}


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's own loop (1 env, batch 32) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cpu = time_reference(args.batch, min_iters=max(REF_MIN_ITERS, args.steps), budget_s=15.0)
    v = cpu["value"]
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "grad_steps_per_sec": v,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": 1e3 / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ENV_ID}, DTQN ctx={CTX}, in-embed={EMBED}, 1 env, batch {args.batch}, reference CPU loop "
                               "(run.py:290-298: act + env.step + store + train per step)",
                   "threads": cpu["cores"], "host_cores": os.cpu_count(),
                   "timed_iterations": "max(200, --steps) loop iterations after 50 000 prepopulate steps (run.py:495)"},
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def timed_iterations(trainers, steps, warmup, barrier, streams=None):
    """ms per lockstep iteration of one or more independent trainers (each on its own stream when given), CUDA events."""
    import torch
    cur = torch.cuda.current_stream()

    def one():
        if streams is None:
            for t in trainers:
                t.train_iteration()
            return
        for t, s in zip(trainers, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                t.train_iteration()
        for s in streams:
            cur.wait_stream(s)
    for _ in range(warmup):
        one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def memory5_sub_record(dev, n_envs, batch, steps, barrier):
    """BASELINE.json configs[2] (Memory-5-v0, 4096 batched envs, ctx 50, in-embed 128) as a short sub-record of the default
    line: the same loop (policy in the loop, 1 grad step per lockstep env step) for `steps` graph-replayed iterations."""
    import torch
    from dtqn_b200.runner import BatchedTrainer
    tr = BatchedTrainer("Memory-5-v0", n_envs, seed=1, device=dev, inner_embed=128, heads=HEADS, layers=LAYERS, context=CTX,
                        batch=batch)
    tr.prepopulate(70)
    while not tr.agent.replay_buffer.can_sample(batch):
        tr.prepopulate(16)
    tr.enable_graphs()
    ms = timed_iterations([tr], steps, 3, barrier)
    tr.agent.check_finite()
    rec = {"workload": f"Memory-5-v0, {n_envs} batched envs per GPU, ctx={CTX}, in-embed=128, train batch {batch}", "steps": steps,
           "ms_per_step": ms, "env_steps_per_sec": n_envs * 1e3 / ms, "grad_steps_per_sec": 1e3 / ms,
           "acting_forward_algorithmic_TFLOPs_per_sec": 893_440 * n_envs * CTX / (ms * 1e-3) / 1e12}
    del tr
    torch.cuda.empty_cache()
    return rec


def run_mixed(args):
    """BASELINE.json configs[4]: CarFlag + Memory-5 at ctx = 128.  The reference cannot run this mix (one
    agent's buffer / network take envs[0]'s observation width 3 vs 10 and action count 3 vs 10, utils/agent_utils.py:87-107),
    so the supported reading is two independent agent groups per GPU -- each with its own envs, replay shard, network and
    optimiser -- stepping in the same iteration on two CUDA streams; value = env-steps/s of both groups and all ranks."""
    import torch
    import torch.distributed as dist
    from dtqn_b200.runner import BatchedTrainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    n_each, ctx = args.envs // 2, 128
    car = BatchedTrainer("DiscreteCarFlag-v0", n_each, seed=1, device=dev, inner_embed=64, heads=HEADS, layers=LAYERS,
                         context=ctx, batch=args.batch)
    mem = BatchedTrainer("Memory-5-v0", n_each, seed=1, device=dev, inner_embed=128, heads=HEADS, layers=LAYERS,
                         context=ctx, batch=args.batch, max_episode_steps=ctx)
    for t, n in ((car, 260), (mem, 160)):
        t.prepopulate(n)
        while not t.agent.replay_buffer.can_sample(args.batch):
            t.prepopulate(32)
        if not args.no_graph:
            t.enable_graphs()
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    W = max(3, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.25)
    ms = torch.tensor([timed_iterations([car, mem], args.steps, W, barrier, streams)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms = float(ms.item())
    ms_serial = timed_iterations([car, mem], max(5, args.steps // 4), 1, barrier)
    ms_car = timed_iterations([car], max(5, args.steps // 4), 1, barrier)
    ms_mem = timed_iterations([mem], max(5, args.steps // 4), 1, barrier)
    for t in (car, mem):
        t.agent.check_finite()
    if rank == 0:
        line = {
            "metric": "env_steps_per_sec", "value": 2 * n_each * world * 1e3 / ms, "unit": "env-steps/s",
            "grad_steps_per_sec": 2e3 / ms, "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mixed DiscreteCarFlag-v0 (in-embed 64) + Memory-5-v0 (in-embed 128, TimeLimit {ctx}), ctx={ctx}, "
                                   f"{n_each} batched envs per group per GPU, train batch {args.batch} per group, two independent agent "
                                   "groups per GPU on two CUDA streams (the reference cannot mix these spaces in one agent, SURVEY A-Q9)",
                       "envs_per_gpu": 2 * n_each, "parallelism": f"dp{world}", "cuda_graphs": not args.no_graph,
                       "l2": "per-step working set (acting-forward activations of 2 x %d x 128 tokens) exceeds the 126 MB L2; no flush" % n_each},
            "breakdown": {"both_groups_two_streams_ms": ms, "both_groups_one_stream_ms": ms_serial, "carflag_group_ms": ms_car,
                          "memory_group_ms": ms_mem},
            "e2e": None, "roofline": None, "cpu_baseline": None, "clocks": clocks, "gpu_launches": None,
            "note": "secondary workload line: parity for ctx = 128 is covered by tests (train kernels at ctx 128, acting path per GEMM); "
                    "the headline line with e2e / roofline / cpu_baseline is the default workload",
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "mixed128":
        return run_mixed(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ctypes as C
    from dtqn_b200 import _lib
    from dtqn_b200.runner import BatchedTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = args.envs
    tr = BatchedTrainer(ENV_ID, N, seed=1, device=dev, inner_embed=EMBED, heads=HEADS, layers=LAYERS, context=CTX,
                        batch=args.batch)
    use_graph = not args.no_graph and hasattr(tr, "enable_graphs")
    # prepopulate with the random policy until every rank can sample (>= 32 completed episodes; CarFlag episodes last
    # up to 200 steps, so ~250 lockstep steps = ~1M transitions per rank)
    tr.prepopulate(260)
    assert tr.agent.replay_buffer.can_sample(args.batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def replicas_identical(t):
        """Parameters and both Adam moments bit-identical on every rank (min == max elementwise over ranks)."""
        ok = True
        for x in (t.agent.policy_network.flat, t.agent.exp_avg, t.agent.exp_avg_sq, t.agent.target_network.flat):
            lo, hi = x.clone(), x.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = ok and bool(torch.equal(lo, hi)) and bool(torch.isfinite(x).all())
        return ok

    # ---- N > 1 self-check: the gradient collective of one real update against NCCL's allreduce of the same gradients ----
    exchange_check = None
    if world > 1:
        agent, rb = tr.agent, tr.agent.replay_buffer
        eps_i, starts_i = rb.draw_indices(agent.batch_size)
        rb.gather_windows(eps_i, starts_i, out=agent._win)
        agent.forward_backward(*agent._win[:4])
        want = agent.grads.clone()
        dist.all_reduce(want)                               # NCCL sum of the per-rank gradients
        agent.reduce_and_step()
        agent.finish_step()
        torch.cuda.synchronize()
        got = agent.exchange.reduced if agent.exchange is not None else agent.grads
        exchange_check = {"p2p_vs_nccl_max_abs": float((got - want).abs().max().item()),
                          "grad_max_abs": float(want.abs().max().item()),
                          "replicas_identical_after_first_update": replicas_identical(tr)}
    if use_graph:
        tr.enable_graphs()

    W = max(3, args.warmup)
    for _ in range(W):
        tr.train_iteration()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.train_iteration()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms = float(ms.item())
    tr.agent.check_finite()
    replicas_ok = replicas_identical(tr) if world > 1 else None
    env_sps = N * world * args.steps / (ms / 1e3)
    grad_sps = args.steps / (ms / 1e3)

    # ---- e2e: the same loop driven through the host-facing API with HOST buffers (pinned), copies inside the timing ----
    hrng = np.random.default_rng(1234 + rank)
    A = tr.env.num_actions
    h2d = N * 4
    d2h = N * A * 4 + N * tr.env.obs_dim * 4 + N * 4 + N + 8 * 4
    if use_graph:
        tr.enable_host_loop()

        def e2e_step(eps):
            q = tr.host_q()                                             # acting forward (graph) -> pinned host Q, sync
            greedy = q.numpy().argmax(1).astype(np.int32)               # host-side epsilon-greedy (user policy code)
            explore = hrng.random(N) < eps
            tr.h_act.numpy()[:] = np.where(explore, hrng.integers(0, A, N), greedy)
            tr.host_step(tr.h_act)                                      # H2D actions, env step + appends, D2H obs / reward / done
            return float(tr.host_train()[0])                            # train step (graph), D2H statistics, sync
    else:
        q_host = torch.empty((N, A), dtype=torch.float32, pin_memory=True)
        act_host = torch.empty((N,), dtype=torch.int32, pin_memory=True)
        obs_host = torch.empty((N, tr.env.obs_dim), dtype=torch.float32, pin_memory=True)
        rew_host = torch.empty((N,), dtype=torch.float32, pin_memory=True)
        done_host = torch.empty((N,), dtype=torch.uint8, pin_memory=True)
        loss_host = torch.empty((8,), dtype=torch.float32, pin_memory=True)

        def e2e_step(eps):
            q = tr.agent.q_last_batched()                              # acting forward on the device context
            q_host.copy_(q, non_blocking=True); torch.cuda.synchronize()
            greedy = q_host.numpy().argmax(1).astype(np.int32)          # host-side epsilon-greedy (user policy code)
            explore = hrng.random(N) < eps
            act_host.numpy()[:] = np.where(explore, hrng.integers(0, A, N), greedy)
            tr.env.step(actions=act_host.to(dev, non_blocking=True), mode=_lib.ACT_GIVEN)
            obs_host.copy_(tr.env.obs_out, non_blocking=True); rew_host.copy_(tr.env.reward_out, non_blocking=True)
            done_host.copy_(tr.env.done_out, non_blocking=True)
            tr.agent.train()
            loss_host.copy_(tr.agent.stats, non_blocking=True)
            torch.cuda.synchronize()
            return float(loss_host[0])

    if use_graph:
        tr.disable_graphs()
    for _ in range(3):
        e2e_step(tr.eps.val)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(10, args.steps // 2)
    for _ in range(e2e_steps):
        e2e_step(tr.eps.val)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = N * world * e2e_steps / float(e2e_s.item())

    # ---- roofline of the dominant kernel: CUDA events around every launch of the tagged kernel, in a dedicated pass ----
    roof = None
    launches = None
    lib = _lib.lib
    lib.dtqn_profile_enable.argtypes = [C.c_int32]
    lib.dtqn_profile_read.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    psteps = min(10, args.steps)
    if rank == 0:
        lib.dtqn_profile_enable(1)
    for _ in range(psteps):                 # every rank runs the pass (the loop contains the gradient collective)
        tr.train_iteration()
    barrier()
    if rank == 0:
        tags = {}
        names = ["linear_fwd", "attn_fwd", "env_step", "env_roll", "replay_gather", "dgrad", "wgrad", "attn_bwd", "ln_bwd",
                 "embed", "head", "td_loss", "clip_adam", "other", "linear_tcgen05", "seq_fused_fwd", "act_fused_tcgen05"]
        for t, nm in enumerate(names):
            ms_t, n_t, w_t = C.c_double(), C.c_int64(), C.c_double()
            lib.dtqn_profile_read(t, C.byref(ms_t), C.byref(n_t), C.byref(w_t))
            if n_t.value:
                tags[nm] = dict(ms=ms_t.value, n=n_t.value, work=w_t.value)
        lib.dtqn_profile_enable(0)
        pk = peaks()
        dom = max(tags.items(), key=lambda kv: kv[1]["ms"]) if tags else None
        if dom is not None:
            nm, v = dom
            tensor = nm in ("linear_fwd", "dgrad", "wgrad", "attn_fwd", "attn_bwd", "seq_fused_fwd", "act_fused_tcgen05")
            if tensor:
                ach = v["work"] / (v["ms"] * 1e-3) / 1e12
                roof = {"kernel": nm, "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                        "frac": ach / pk["tf_sustained"], "traffic": None, "peak_source": pk["src"] + " (sustained bf16)",
                        "launches_timed": v["n"], "avg_us_per_launch": 1e3 * v["ms"] / v["n"]}
            else:
                ach = v["work"] / (v["ms"] * 1e-3) / 1e9
                roof = {"kernel": nm, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["src"],
                        "launches_timed": v["n"], "avg_us_per_launch": 1e3 * v["ms"] / v["n"]}
            if nm == "act_fused_tcgen05":
                try:
                    src = next(n for n in ("r2_act_fused_ncu.json", "r1_act_fused_ncu.json")
                               if os.path.exists(os.path.join(ROOT, "profiles", n)))
                    with open(os.path.join(ROOT, "profiles", src)) as f:
                        aj = json.load(f)
                    roof["traffic"] = aj["dram_bytes_read_per_launch"] + aj["dram_bytes_write_per_launch"]
                    roof["traffic_note"] = (f"dram__bytes_read + write of one launch (ncu --set full, profiles/{src}): the "
                                            "context rows and the weight image; every activation stays in shared memory / TMEM")
                    roof["burst_peak_view"] = {"peak": pk["tf_burst"], "frac": ach / pk["tf_burst"],
                                               "note": "against the burst bf16 peak (a kernel timed alone); the line's frac uses the sustained "
                                                       "peak because the kernel is timed inside a long step"}
                    roof["tensor_view"] = {"sm__pipe_tensor_cycles_active_pct": aj["sm__pipe_tensor_cycles_active_pct"],
                                           "note": "achieved = ALGORITHMIC FLOPs (embed + layer 0 + final-layer in_proj + last-row attention) / "
                                                   "launch time; the bf16 hi/lo split issues 3 MMAs per algorithmic MMA, so this fraction is "
                                                   "capped at 1/3, and the attention core runs on mma.sync (TF32 x3), not on tcgen05"}
                except Exception:
                    pass
            try:
                with open(os.path.join(ROOT, "profiles", "r1_linear_traffic.json")) as f:
                    tj = json.load(f)
                if nm == "linear_tcgen05":
                    # DRAM bytes of the tcgen05 launches of one step (committed ncu --set full capture) / launches per step
                    roof["traffic"] = tj["tcgen05_bytes_per_step_MB"] * 1e6 / (v["n"] / psteps)
                    roof["traffic_note"] = ("bytes per launch averaged over the %d tcgen05 Linear launches of a step (ncu dram__bytes_read+write, "
                                            "profiles/r1_linear_traffic.json); below the algorithmic bytes because part of each output is still "
                                            "in the 126 MB L2 when the kernel ends" % round(v["n"] / psteps))
                    # the GEMMs have K = 64..256: 24-64 FLOP per byte against a machine balance of ~220 -> HBM-bound by the roofline
                    # model; the tensor-pipe view of the same launches (2*M*N*K per launch, x3 issued for the bf16 hi/lo split):
                    flops = 2.0 * N * CTX * EMBED * EMBED * (3 + 1 + 4 + 4 + 2)
                    roof["tensor_view"] = {"algorithmic_TFLOPs": flops * psteps / (v["ms"] * 1e-3) / 1e12,
                                           "of_sustained_bf16_peak": flops * psteps / (v["ms"] * 1e-3) / 1e12 / pk["tf_sustained"],
                                           "note": "3 MMAs issued per algorithmic MMA (bf16 hi/lo split): utilisation on algorithmic FLOPs is capped at 1/3"}
            except Exception:
                pass
            roof["per_kernel_share_of_timed_ms"] = {k: round(x["ms"] / sum(y["ms"] for y in tags.values()), 4) for k, x in tags.items()}
            roof["per_kernel_us_per_step"] = {k: round(1e3 * x["ms"] / psteps, 1) for k, x in tags.items()}
            roof["per_kernel_launches_per_step"] = {k: round(x["n"] / psteps, 1) for k, x in tags.items()}
            for k in ("env_step", "replay_gather"):
                if k in tags:
                    roof[k + "_GBps"] = tags[k]["work"] / (tags[k]["ms"] * 1e-3) / 1e9
        launches = int(sum(v["n"] for v in tags.values()) / max(1, psteps)) if tags else None

    # ---- breakdown (single GPU): the two halves of the iteration and the pure env kernel, each as its own CUDA graph ----
    breakdown = None
    if world == 1:
        def timed(graph_or_fn, n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(n):
                graph_or_fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        tr._sync_epsilon_to_device()
        g_act, g_train = tr.capture(tr.act_only), tr.capture(tr.train_only)
        g_env = tr.capture(lambda: tr.env.step(mode=_lib.ACT_RANDOM))
        for g in (g_act, g_train, g_env):
            for _ in range(3):
                g.replay()
        n = max(20, min(200, args.steps))
        ms_act, ms_train, ms_env = timed(g_act.replay, n), timed(g_train.replay, n), timed(g_env.replay, n)
        tr.agent.num_train_steps += n + 4
        breakdown = {"acting_half_ms": ms_act, "train_half_ms": ms_train,
                     "train_only_grad_steps_per_sec": 1e3 / ms_train,
                     "policy_in_loop_env_steps_per_sec_no_training": N * 1e3 / ms_act,
                     "random_policy_env_steps_per_sec": N * 1e3 / ms_env, "env_step_plus_roll_us": 1e3 * ms_env,
                     "env_step_GBps_algorithmic_55B": 55.0 * N / (ms_env * 1e-3) / 1e9}

    # ---- N > 1: the same loop with the NCCL allreduce as the gradient collective, timed beside the fused exchange ----
    nccl_arm = None
    if world > 1 and tr.allreduce == "p2p-fused":
        os.environ["DTQN_B200_ALLREDUCE"] = "nccl"
        tr2 = BatchedTrainer(ENV_ID, N, seed=1, device=dev, inner_embed=EMBED, heads=HEADS, layers=LAYERS, context=CTX,
                             batch=args.batch)
        os.environ.pop("DTQN_B200_ALLREDUCE", None)
        assert tr2.allreduce == "nccl"
        tr2.prepopulate(260)
        if use_graph:
            tr2.enable_graphs()
        for _ in range(W):
            tr2.train_iteration()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            tr2.train_iteration()
        a1.record()
        barrier()
        ms2 = torch.tensor([a0.elapsed_time(a1)], device=dev)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        tr2.agent.check_finite()
        nccl_arm = {"ms_per_step": float(ms2.item()) / args.steps,
                    "env_steps_per_sec": N * world * args.steps / (float(ms2.item()) / 1e3),
                    "replicas_identical": replicas_identical(tr2),
                    "note": "same loop, DTQN_B200_ALLREDUCE=nccl: torch.distributed.all_reduce of the flat gradient (outside the CUDA "
                            "graph) + sqnorm + clip/Adam kernels"}
        del tr2

    sub = None
    if world == 1 and args.workload == "carflag" and not args.no_sub_records:
        sub = {"memory5": memory5_sub_record(dev, N, args.batch, max(20, min(60, args.steps)), barrier)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_reference(args.batch)

    if rank == 0:
        line = {
            "metric": "env_steps_per_sec", "value": env_sps, "unit": "env-steps/s", "grad_steps_per_sec": grad_sps,
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ENV_ID}, {N} batched envs per GPU, ctx={CTX}, in-embed={EMBED}, heads={HEADS}, "
                                   f"layers={LAYERS}, train batch {args.batch} windows per GPU, policy-in-the-loop "
                                   "(eps-greedy through the DTQN), 1 grad step per lockstep env step",
                       "envs_per_gpu": N, "global_batch": args.batch * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set (acting-forward activations, ~600 MB) exceeds the 126 MB L2; no flush",
                       "cuda_graphs": bool(use_graph),
                       "grad_collective": {"none": "none (1 GPU)", "nccl": "NCCL allreduce + clip/Adam kernels (outside the graph)",
                                           "p2p-fused": "own kernel: flag barrier + NVLink peer reads + norm, then clip/Adam "
                                                        "(inside the CUDA graph, no NCCL call)"}[tr.allreduce]},
            "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": ("BatchedTrainer.host_q (graph, D2H Q) -> host eps-greedy -> host_step(pinned actions: H2D, env graph, D2H "
                                            "obs/reward/done) -> host_train (graph, D2H statistics)") if use_graph else
                                           ("agent.q_last_batched -> host eps-greedy -> BatchedEnv.step(host actions) -> "
                                            "host obs/reward/done -> agent.train -> host loss")},
            "gpu_launches": launches,
            "replicas_identical": replicas_ok, "exchange_check": exchange_check, "nccl_arm": nccl_arm,
            "sub_records": sub,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "breakdown": breakdown,
            "acting_forward_algorithmic_TFLOPs_per_sec": FWD_FLOP_PER_TOKEN * N * CTX * world * args.steps / (ms / 1e3) / 1e12,
        }
        _emit(line)
    if world > 1:
        assert replicas_ok, "data-parallel replicas diverged (parameters / Adam moments differ between ranks)"
        dist.destroy_process_group()


if __name__ == "__main__":
    _quiet_stdout()
    main()
