"""Experiment driver with the reference's command-line flags (run.py:16-184) driving the batched B200 loop.

    python -m dtqn_b200.run --envs DiscreteCarFlag-v0 --in-embed 64 --n-envs 4096
    torchrun --nproc-per-node 8 -m dtqn_b200.run --envs DiscreteCarFlag-v0 --in-embed 64 --n-envs 4096

One "timestep" of the reference loop (run.py:290) is one lockstep iteration here: every env takes one step and the agent
takes one gradient step, so ``--num-steps`` iterations collect ``num-steps x n-envs x world`` transitions.  New flags (no
reference analogue): ``--n-envs`` lockstep environments per GPU, ``--record-every``, ``--updates-per-step``, ``--float-context``.  Every ``--eval-frequency`` iterations rank 0 logs the
reference's keys (run.py:303-325) through ``dtqn_b200.logging_utils`` -- ``<policy_path>_results.csv`` / ``_losses.csv`` with
``--disable-wandb``, wandb otherwise -- and prints the ``--verbose`` line; ``--render`` is accepted and ignored.
Checkpoint / resume follows run.py:452-499,337-352,526-529: the same ``policies/<project>/<env>/model=..._seed=N`` path
prefix, ``_mini_checkpoint.pt`` / ``_checkpoint.pt`` / ``buffer_*.sav`` files (one set per rank, suffix ``_rank<r>`` when
world > 1), written when ``--time-limit`` expires and read back on the next launch.
"""
import argparse
import os
import time

import torch
import torch.distributed as dist


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--project-name", type=str, default="DTQN-test")
    p.add_argument("--disable-wandb", action="store_true")
    p.add_argument("--time-limit", type=float, default=None, help="hours; stop (cleanly) after this long")
    p.add_argument("--model", type=str, default="DTQN", choices=["DTQN"],
                   help="only the DTQN family is on the B200 hot path (baselines are out of scope)")
    p.add_argument("--envs", type=str, nargs="+", default=["DiscreteCarFlag-v0"])
    p.add_argument("--num-steps", type=int, default=2_000_000)
    p.add_argument("--tuf", type=int, default=10_000, help="target update frequency (gradient steps)")
    p.add_argument("--lr", type=float, default=3e-4)
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--buf-size", type=int, default=500_000)
    p.add_argument("--eval-frequency", type=int, default=5_000)
    p.add_argument("--eval-episodes", type=int, default=10)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--context", type=int, default=50)
    p.add_argument("--obs-embed", type=int, default=8)
    p.add_argument("--a-embed", type=int, default=0)
    p.add_argument("--in-embed", type=int, default=128)
    p.add_argument("--max-episode-steps", type=int, default=-1)
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--save-policy", action="store_true")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--render", action="store_true")
    p.add_argument("--history", type=int, default=50)
    p.add_argument("--heads", type=int, default=8)
    p.add_argument("--layers", type=int, default=2)
    p.add_argument("--dropout", type=float, default=0.0)
    p.add_argument("--discount", type=float, default=0.99)
    p.add_argument("--gate", type=str, default="res", choices=["res", "gru"])
    p.add_argument("--identity", action="store_true")
    p.add_argument("--pos", default="learned", choices=["learned", "sin", "none"])
    p.add_argument("--bag-size", type=int, default=0)
    p.add_argument("--slurm-job-id", default=0, type=str)
    p.add_argument("--n-envs", type=int, default=4096, help="lockstep environments per GPU (new)")
    p.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    p.add_argument("--record-every", type=int, default=1,
                   help="each env stores every K-th of its episodes in the replay (new): keeps the reference's replay horizon "
                        "(~25 %% of the run) with thousands of envs per update; 1 = store everything")
    p.add_argument("--updates-per-step", type=int, default=1, help="gradient updates per lockstep env step (new)")
    p.add_argument("--float-context", action="store_true",
                   help="keep float observations in the acting context instead of reproducing the reference's int64 context "
                        "(utils/context.py:46 truncates CarFlag observations toward zero)")
    return p.parse_args(argv)


def run_experiment(args):
    from dtqn_b200.runner import BatchedTrainer
    if len(args.envs) != 1:
        raise NotImplementedError("multi-env sampling needs identical spaces (run.py:47); run one trainer per env id")
    if args.bag_size:
        raise NotImplementedError("--bag-size > 0 (DTQN-bag) is outside the hot path (SURVEY.md section 2 #23)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local) if args.device.startswith("cuda") else torch.device(args.device)
    tr = BatchedTrainer(args.envs[0], args.n_envs, seed=args.seed, device=device, inner_embed=args.in_embed,
                        heads=args.heads, layers=args.layers, context=args.context, batch=args.batch,
                        buf_size=max(args.buf_size, 8 * args.n_envs * 200), lr=args.lr, tuf=args.tuf,
                        gamma=args.discount, history=args.history, num_steps=args.num_steps, obs_embed=args.obs_embed,
                        pos=args.pos, max_episode_steps=args.max_episode_steps, a_embed=args.a_embed, dropout=args.dropout,
                        identity=args.identity, gate=args.gate, record_every=args.record_every,
                        updates_per_step=args.updates_per_step, trunc_context_obs=not args.float_context)
    rank = tr.rank
    if rank == 0:
        n = sum(p.numel() for p in tr.agent.policy_network.parameters())
        print(f"Creating {args.model} with {n} parameters")                      # run.py:447-450
    policy_dir = os.path.join(os.getcwd(), "policies", args.project_name, *args.envs)          # run.py:452-460
    os.makedirs(policy_dir, exist_ok=True)
    policy_path = os.path.join(
        policy_dir,
        f"model={args.model}_envs={','.join(args.envs)}_obs_embed={args.obs_embed}_a_embed={args.a_embed}_in_embed={args.in_embed}"
        f"_context={args.context}_heads={args.heads}_layers={args.layers}_batch={args.batch}_gate={args.gate}"
        f"_identity={args.identity}_history={args.history}_pos={args.pos}_bag={args.bag_size}_seed={args.seed}")
    ckpt = policy_path + (f"_rank{rank}" if world > 1 else "")
    from dtqn_b200.checkpoint import RunningAverage
    have_ckpt = os.path.exists(ckpt + "_mini_checkpoint.pt")
    if world > 1:                     # every rank must take the same branch: resume only if ALL ranks hold their files
        flag = torch.tensor([int(have_ckpt)], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if have_ckpt and not bool(flag.item()):
            print(f"[rank {rank}] ignoring {ckpt}_mini_checkpoint.pt: another rank has no checkpoint; starting fresh")
        have_ckpt = bool(flag.item())
    if have_ckpt:                                                                                 # run.py:469-490
        done_steps = tr.agent.load_mini_checkpoint(ckpt)["step"]
        print(f"Found a mini checkpoint that completed {done_steps} training steps.")
        if done_steps >= args.num_steps:
            print("Removing checkpoint and exiting...")
            if os.path.exists(ckpt + "_checkpoint.pt"):
                os.remove(ckpt + "_checkpoint.pt")
            if world > 1:
                dist.destroy_process_group()
            return tr
        wandb_id, mean_success_rate, mean_reward, mean_episode_length = tr.load_checkpoint(ckpt)
        wandb_kwargs = {"resume": "must", "id": wandb_id} if wandb_id else {"resume": None}            # run.py:490
    else:
        # prepopulate 50 000 transitions (run.py:495) with the random policy, and at least until a batch can be sampled
        steps = max(1, 50_000 // args.n_envs)
        tr.prepopulate(steps)
        while not tr.agent.replay_buffer.can_sample(args.batch):
            tr.prepopulate(16)
        mean_success_rate, mean_reward, mean_episode_length = RunningAverage(10), RunningAverage(10), RunningAverage(10)
        wandb_kwargs = {"resume": None}                                                           # run.py:493
    logger = None
    if rank == 0:
        from dtqn_b200 import logging_utils
        logger = logging_utils.get_logger(policy_path, args, wandb_kwargs)                       # run.py:501
    if not args.no_graph:
        tr.enable_graphs()
    start = time.time()
    for timestep in range(tr.agent.num_train_steps, args.num_steps):
        tr.train_iteration()
        if timestep % args.eval_frequency == 0:
            # run.py:315: `eval_episodes` greedy episodes; here every lockstep eval env of every rank plays the same number
            # of episodes, the smallest count that reaches eval_episodes in total
            sr, ret, length = tr.evaluate(max(1, -(-args.eval_episodes // (args.n_envs * world))))
            mean_success_rate.add(sr); mean_reward.add(ret); mean_episode_length.add(length)     # run.py:316-318
            if rank == 0:
                hours = (time.time() - start) / 3600
                log_vals = logging_utils.loss_log_values(tr.agent, hours)                        # run.py:303-313
                log_vals.update({f"{args.envs[0]}/SuccessRate": sr, f"{args.envs[0]}/Return": ret,
                                 f"{args.envs[0]}/EpisodeLength": length})                       # run.py:318-324
                logger.log(log_vals, step=timestep)
                if args.verbose:                                                                 # run.py:332-335
                    print(f"[ {logging_utils.timestamp()} ] Training Steps: {timestep}, Env: {args.envs[0]}, Success Rate: "
                          f"{sr:.2f}, Return: {ret:.2f}, Episode Length: {length:.2f}, Hours: {hours:.2f}", flush=True)
        if args.save_policy and timestep % 50_000 == 0 and rank == 0:                           # run.py:337-338
            torch.save(tr.agent.policy_network.state_dict(), policy_path)
        stop = False
        if args.time_limit and timestep % 64 == 0:                                              # run.py:340-353
            stop = (time.time() - start) / 3600 >= args.time_limit
            if world > 1:                        # every rank must take the same branch (the update holds a collective)
                flag = torch.tensor([int(stop)], device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                stop = bool(flag.item())
        if stop:
            print(f"Reached time limit. Saving checkpoint with {tr.agent.num_train_steps} steps completed.")
            run_id = getattr(getattr(logger, "run", None), "id", None)                           # wandb.run.id (run.py:347)
            tr.save_checkpoint(ckpt, run_id, mean_success_rate, mean_reward, mean_episode_length)
            break
    else:
        tr.agent.save_mini_checkpoint(ckpt, getattr(getattr(logger, "run", None), "id", None))                                               # run.py:526-529
    if world > 1:
        dist.destroy_process_group()
    return tr


if __name__ == "__main__":
    run_experiment(get_args())
