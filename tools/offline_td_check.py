"""Long-horizon training parity on a FIXED replay buffer: the device agent and the CPU oracle start from the same
parameters, see the same (episode, start) indices for K consecutive updates (target sync every `tuf`), and are compared on
loss, on Q of a fixed probe batch, and on how well each fits the terminal transitions."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dtqn_b200.envs import BatchedEnv, ContextWindow
from dtqn_b200.utils import get_agent
from oracle import agent as oagent, network as onet

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tuf = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B, ctx = 32, 50
torch.manual_seed(0)
env = BatchedEnv("DiscreteCarFlag-v0", 64, seed=1, device="cuda")
agent = get_agent("DTQN", [env], 8, 0, 64, 64 * 8 * 200, "cuda", 3e-4, B, ctx, 200, ctx, tuf, 0.99, num_heads=8, num_layers=2, n_envs=64)
env.attach(agent.replay_buffer, agent.train_context)
env.reset_all()
for _ in range(600):
    env.step()                       # random policy
torch.cuda.synchronize()
rb = agent.replay_buffer
c = rb.counters.cpu().numpy(); used = min(int(c[1]), rb.max_size)
lens = rb.episode_lengths[:used].cpu().numpy(); open_ = rb.slot_open[:used].cpu().numpy().astype(bool)
valid = np.where(~open_ & (lens > 0))[0]
obss = rb.obss.cpu().numpy(); acts = rb.actions.cpu().numpy(); rews = rb.rewards.cpu().numpy(); dones = rb.dones.cpu().numpy()
print("completed episodes", len(valid), "terminal(+1/-1):", int((rews[valid, lens[valid] - 1, 0] > 0).sum()), int((rews[valid, lens[valid] - 1, 0] < 0).sum()))
sd = {k: v.detach().cpu().clone() for k, v in agent.policy_network.state_dict().items()}
tr = oagent.TrainerOracle(sd, 8, target_update_frequency=tuf)
rng = np.random.default_rng(0)

def gather(eps, starts):
    e = eps[:, None]; t = starts[:, None] + np.arange(ctx)[None, :]
    return (torch.from_numpy(obss[e, t]), torch.from_numpy(acts[e, t].astype(np.int64)), torch.from_numpy(rews[e, t]),
            torch.from_numpy(obss[e, t + 1]), torch.from_numpy(acts[e, t + 1].astype(np.int64)), torch.from_numpy(dones[e, t].astype(bool)))

pe = valid[:16]; ps = np.zeros(16, dtype=np.int64)
probe = torch.from_numpy(obss[pe[:, None], np.arange(ctx)[None, :]])
# windows that end at the terminal transition of episodes that reached heaven / hell
term = valid[rews[valid, lens[valid] - 1, 0] != 0][:256]
tstart = np.maximum(0, lens[term] - ctx)
t0 = time.time()
for k in range(1, K + 1):
    eps = rng.choice(valid, size=B)
    starts = np.array([rng.integers(0, max(0, lens[e] - ctx) + 1) for e in eps], dtype=np.int64)
    agent.train(indices=(torch.from_numpy(eps.astype(np.int32)).cuda(), torch.from_numpy(starts.astype(np.int32)).cuda()))
    stats, _ = tr.train_on_batch(gather(eps, starts))
    if k % 250 == 0 or k == 1:
        g_stats = agent.stats.cpu().numpy()
        with torch.no_grad():
            qo = onet.forward(tr.policy, probe, 8).numpy()
        qg = agent.policy_network(probe).cpu().numpy()
        # terminal fit: Q(s_{T-1}, a_{T-1}) vs reward, on short episodes
        xs = torch.from_numpy(obss[term[:, None], tstart[:, None] + np.arange(ctx)[None, :]])
        with torch.no_grad():
            qso = onet.forward(tr.policy, xs, 8).numpy()
        qsg = agent.policy_network(xs).cpu().numpy()
        T = lens[term] - 1
        Tw = T - tstart
        a_T = acts[term, T, 0]; r_T = rews[term, T, 0]
        fit_o = np.abs(qso[np.arange(len(term)), Tw, a_T] - r_T).mean(); fit_g = np.abs(qsg[np.arange(len(term)), Tw, a_T] - r_T).mean()
        print(json.dumps(dict(step=k, loss_gpu=float(g_stats[0]), loss_oracle=stats["loss"], gnorm_gpu=float(g_stats[7]), gnorm_oracle=stats["grad_norm"],
                              probe_q_maxdiff=float(np.abs(qo - qg).max()), probe_q_absmax=float(np.abs(qo).max()),
                              terminal_fit_err_gpu=float(fit_g), terminal_fit_err_oracle=float(fit_o), wall_s=round(time.time() - t0, 1))), flush=True)
