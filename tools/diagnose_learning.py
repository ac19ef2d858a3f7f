"""Diagnostic: train briefly, then (1) check reward / priest-signal consistency of the replay data, (2) probe Q for hand-made
contexts where the priest signal is visible."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dtqn_b200.runner import BatchedTrainer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
tr = BatchedTrainer("DiscreteCarFlag-v0", n_envs=256, seed=1, device="cuda", inner_embed=64, context=50, batch=128,
                    num_steps=200000, trunc_context_obs=False, tuf=500, buf_size=131072 * 200)
tr.prepopulate(260)
tr.enable_graphs()
for it in range(iters):
    tr.train_iteration()
torch.cuda.synchronize()
rb = tr.agent.replay_buffer
c = rb.counters.cpu().numpy()
used = min(int(c[1]), rb.max_size)
lens = rb.episode_lengths[:used].cpu().numpy(); open_ = rb.slot_open[:used].cpu().numpy().astype(bool)
obss = rb.obss[:used].cpu().numpy(); rews = rb.rewards[:used].cpu().numpy()[..., 0]; dones = rb.dones[:used].cpu().numpy()[..., 0]
acts = rb.actions[:used].cpu().numpy()[..., 0]
ok = bad = 0; term_pos = term_neg = 0
for s in np.where(~open_ & (lens > 0))[0][:20000]:
    L = lens[s]
    r = rews[s, L - 1]
    dirs = obss[s, : L + 1, 2]
    seen = dirs[dirs != 0]
    if r > 0: term_pos += 1
    if r < 0: term_neg += 1
    if r != 0 and len(seen):
        final_p = obss[s, L, 0]
        side = 1.0 if final_p >= 1 else -1.0
        heaven = seen[0]
        expect = 1.0 if side == heaven else -1.0
        if expect == r: ok += 1
        else: bad += 1
print("replay consistency: reward matches (final side == priest signal):", ok, "mismatch:", bad, "terminal +1:", term_pos, "-1:", term_neg)
net = tr.agent.policy_network
def ctx_for(heaven):
    # drive right from p=0 with full throttle: v += 0.0015 up to 0.07
    p, v, rows = 0.0, 0.0, []
    rows.append([p, v, 0.0])
    while len(rows) < 50:
        v = min(0.07, v + 0.0015); p = p + v
        d = heaven if 0.3 <= p <= 0.7 else 0.0
        rows.append([p, v, d])
        if p > 0.62: break
    return rows
for heaven in (1.0, -1.0):
    rows = ctx_for(heaven)
    x = torch.tensor([rows], dtype=torch.float32)
    q = net(x).cpu().numpy()[0]
    print(f"heaven={heaven:+.0f}  context len {len(rows)}  last obs {rows[-1]}  Q(left,stay,right) at last pos = {q[-1]}  at first priest pos = {q[[i for i,r in enumerate(rows) if r[2]!=0][0]]}")
sr = tr.evaluate(1); print("eval (success, return, length):", sr)
st = tr.env.ep_stats.cpu().numpy(); print("train-env episode stats (sum return, sum len, successes, episodes):", st)
