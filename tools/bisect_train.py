import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_net_gpu as T
from dtqn_b200 import _lib, networks
G = "/root/repo/tests/golden"
z = np.load(os.path.join(G, "train_carflag.npz"))
for tc_min, par, pipe in [(1<<30, 0, 1), (1<<30, 1, 1), (1024, 0, 1), (1024, 0, 0), (1024, 1, 1)]:
    networks.set_tc_min_tokens(tc_min); _lib.lib.dtqn_set_parallel_wgrad(par); _lib.lib.dtqn_set_tc_pipelined(pipe)
    agent = T._agent_from_golden(z, "carflag", 32)
    agent.train_on_windows(*T._windows(z, 0))
    torch.cuda.synchronize()
    st = agent.stats.cpu().numpy()
    q_all = agent._q_all.cpu().numpy()
    qe = max(T.rel_err(q_all[gi], z["step0/" + key]) for gi, key in enumerate(("q_policy_obs", "q_policy_next", "q_target_next")))
    grads = agent.policy_network.unflatten(agent.grads)
    worst = ("", 0)
    for k, g in grads.items():
        ref = z["step0/grad/" + k]; err = np.abs(g.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-12)
        if err > worst[1]: worst = (k, err)
    print(f"tc_min={tc_min} par={par} pipe={pipe}: loss {st[0]:.6f} (ref {z['stats/td_errors'][0]:.6f}) gnorm {st[7]:.5f} (ref {z['stats/grad_norms'][0]:.5f}) q_rel {qe:.2e} worst grad {worst}", "tc_err", networks.tc_error())
