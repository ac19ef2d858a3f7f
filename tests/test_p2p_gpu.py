"""Gradient exchange fused with the optimiser over peer memory (csrc/p2p.cu).

Single-GPU tests cover the kernels themselves (world of one; two "ranks" on two streams of one device); the real
two-process NVLink test spawns torchrun and needs >= 2 GPUs (skipped otherwise)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 107_780            # CarFlag flat parameter count (multiple of 4)


def _state(seed, n=N, gscale=1e-4):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return dict(p=torch.randn(n, device="cuda", generator=g), m=torch.zeros(n, device="cuda"), v=torch.zeros(n, device="cuda"),
                step=torch.zeros(1, dtype=torch.int64, device="cuda"), scratch=torch.zeros(1024, device="cuda"),
                stats=torch.zeros(8, device="cuda"), flags=torch.zeros(1, dtype=torch.int32, device="cuda"),
                ring=torch.zeros((100, 8), device="cuda"))


def _clip_adam(lib, s, grads, scale):
    rc = lib.dtqn_clip_adam(s["p"].data_ptr(), grads.data_ptr(), s["m"].data_ptr(), s["v"].data_ptr(), grads.numel(), scale,
                            1.0, 3e-4, 0.9, 0.999, 1e-8, s["step"].data_ptr(), s["scratch"].data_ptr(), s["stats"].data_ptr(),
                            s["flags"].data_ptr(), s["ring"].data_ptr(), 100, torch.cuda.current_stream().cuda_stream)
    assert rc == 0


def _fused(lib, s, bases, rank, world, reduced, stream=None):
    st = (stream or torch.cuda.current_stream()).cuda_stream
    rc = lib.dtqn_allreduce_clip_adam(s["p"].data_ptr(), bases, rank, world, reduced.numel(), reduced.data_ptr(),
                                      s["m"].data_ptr(), s["v"].data_ptr(), 1.0, 3e-4, 0.9, 0.999, 1e-8,
                                      s["step"].data_ptr(), s["scratch"].data_ptr(), s["stats"].data_ptr(),
                                      s["flags"].data_ptr(), s["ring"].data_ptr(), 100, st)
    assert rc == 0


@pytest.mark.parametrize("gscale,exact", [(1e-4, True), (1.0, False)])
def test_world_of_one_matches_clip_adam(gscale, exact):
    from dtqn_b200 import agents  # noqa: F401  (registers the ctypes signatures)
    from dtqn_b200._lib import lib
    from dtqn_b200.parallel import PeerExchange
    ex = PeerExchange.create_single(N, "cuda")
    a, b = _state(1), _state(1)
    for it in range(3):
        g = torch.randn(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(10 + it)) * gscale
        ex.grads.copy_(g)
        _clip_adam(lib, a, g.clone(), 1.0)
        _fused(lib, b, ex.bases, 0, 1, ex.reduced)
        torch.cuda.synchronize()
        assert torch.equal(ex.reduced, g)
        assert int(a["step"].item()) == int(b["step"].item()) == it + 1
        if exact:                                     # norm below max_norm: clip coefficient is exactly 1 in both
            assert torch.equal(a["p"], b["p"]) and torch.equal(a["m"], b["m"]) and torch.equal(a["v"], b["v"])
        else:                                         # clipped: the two norm reductions differ in summation order only
            assert torch.allclose(a["p"], b["p"], rtol=0, atol=1e-6)
        assert abs(float(a["stats"][7]) - float(b["stats"][7])) <= 1e-5 * float(a["stats"][7])
    assert not ex.error()


def test_graph_replay_advances_the_epoch():
    from dtqn_b200 import agents  # noqa: F401
    from dtqn_b200._lib import lib
    from dtqn_b200.parallel import PeerExchange
    ex = PeerExchange.create_single(N, "cuda")
    s = _state(2)
    ex.grads.normal_(0, 1e-4)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _fused(lib, s, ex.bases, 0, 1, ex.reduced)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        _fused(lib, s, ex.bases, 0, 1, ex.reduced)
    for _ in range(25):
        g.replay()
    torch.cuda.synchronize()
    assert int(s["step"].item()) == 26 and not ex.error()
    assert torch.equal(ex.reduced, ex.grads)


def test_two_ranks_on_two_streams_of_one_device():
    """The cross-rank protocol without a second GPU: two exchange buffers, two concurrent launches (rank 0 / rank 1)."""
    from dtqn_b200 import agents  # noqa: F401
    from dtqn_b200._lib import lib
    from dtqn_b200.parallel import PeerExchange
    ex = [PeerExchange(N, "cuda", r, 2) for r in range(2)]
    bases = (C.c_void_p * 2)(ex[0].base, ex[1].base)
    st = [_state(3), _state(3)]
    ref = _state(3)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(6):
        gs = [torch.randn(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + 2 * it + r)) * 1e-4
              for r in range(2)]
        for r in range(2):
            ex[r].grads.copy_(gs[r])
        torch.cuda.synchronize()
        for r in range(2):
            _fused(lib, st[r], bases, r, 2, ex[r].reduced, streams[r])
        torch.cuda.synchronize()
        assert not ex[0].error() and not ex[1].error(), "the two launches did not overlap (bounded wait expired)"
        total = gs[0] + gs[1]
        assert torch.equal(ex[0].reduced, total) and torch.equal(ex[1].reduced, total)
        _clip_adam(lib, ref, total.clone(), 0.5)
        torch.cuda.synchronize()
        for r in range(2):
            assert torch.equal(st[r]["p"], ref["p"]) and torch.equal(st[r]["m"], ref["m"]) and torch.equal(st[r]["v"], ref["v"])
            assert int(st[r]["step"].item()) == it + 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_processes_over_nvlink():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "p2p_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "P2P_WORKER_OK" in out.stdout


def test_two_rank_cuda_step_matches_reference_step(golden_dir):
    """SURVEY section 4(v) on the CUDA path: the reference's batch of 32 windows split 16/16 over two "ranks" (two agents,
    two exchange buffers, the fused exchange launched concurrently on two streams of one device) -- CUDA forward/backward per
    rank -> dtqn_allreduce_clip_adam -> the reduced mean gradient equals the reference's step-0 gradient on the whole batch
    (dtqn/agents/dtqn.py:243-256), the global norm its logged grad norm (:257-263), and after the fixture's three steps
    both replicas hold the reference's post-Adam parameters (:265)."""
    import numpy as np
    from test_net_gpu import _agent_from_golden, _windows
    from dtqn_b200.parallel import PeerExchange
    z = np.load(os.path.join(golden_dir, "train_carflag.npz"))
    d, layers, ctx, B, heads, n_steps = [int(v) for v in z["meta"]]
    agents = [_agent_from_golden(z, "carflag", B // 2) for _ in range(2)]
    n = agents[0].policy_network.n_flat
    ex = [PeerExchange(n, "cuda", r, 2) for r in range(2)]
    bases = (C.c_void_p * 2)(ex[0].base, ex[1].base)
    for r in range(2):
        ex[r].bases = bases
        agents[r].use_peer_exchange(ex[r])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for s in range(n_steps):
        win = _windows(z, s)
        for r in range(2):
            half = [w[r * (B // 2):(r + 1) * (B // 2)].contiguous() for w in win]
            agents[r].forward_backward(*half)
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                agents[r].reduce_and_step()
        torch.cuda.synchronize()
        for r in range(2):
            agents[r].finish_step()
            agents[r].check_finite()
        assert torch.equal(ex[0].reduced, ex[1].reduced)
        assert abs(float(agents[0].stats[7]) - z["stats/grad_norms"][s]) <= 1e-4 * max(1.0, z["stats/grad_norms"][s])
        if s == 0:
            grads = agents[0].policy_network.unflatten(ex[0].reduced * 0.5)
            gmax = max(np.abs(z["step0/grad/" + k]).max() for k in grads if ("step0/grad/" + k) in z.files)
            for k, g in grads.items():
                ref_g = z["step0/grad/" + k]
                err = np.abs(g.cpu().numpy() - ref_g).max()
                assert err <= 1e-3 * max(np.abs(ref_g).max(), 1e-3 * gmax), (k, err)
    assert torch.equal(agents[0].policy_network.flat, agents[1].policy_network.flat)          # replicas bit-identical
    assert torch.equal(agents[0].exp_avg, agents[1].exp_avg) and torch.equal(agents[0].exp_avg_sq, agents[1].exp_avg_sq)
    for k, p in agents[0].policy_network.state_dict().items():
        if not k.endswith("attn_mask"):
            assert np.abs(p.cpu().numpy() - z[f"policy{n_steps}/" + k]).max() < 2e-5, k
