"""CPU oracle restatement of the double-DQN sequence TD update (test infrastructure only).

Follows dtqn/agents/dtqn.py:162-269 (train), :76-107 (get_action), dtqn/agents/dqn.py:46-64,208-210 (ctor / target
update) and, for the third-party arithmetic at those call sites (torch, pinned 1.11.0 in requirements.txt:40):
``F.mse_loss`` mean reduction, ``clip_grad_norm_(params, 1.0, error_if_nonfinite=True)`` and ``optim.Adam`` with
defaults -- restated explicitly below.
"""
import math

import torch

from oracle import network


def td_loss(policy_sd, target_sd, batch, num_heads, gamma=0.99, history=None, identity=False):
    """dtqn.py:215-243.  ``batch`` = (obss, actions, rewards, next_obss, next_actions, dones) as returned by
    ReplayBuffer.sample.  Returns (loss, q_sel[B,H], targets[B,H]); loss carries autograd history w.r.t. policy_sd.
    Padded positions are NOT masked."""
    obss, actions, rewards, next_obss, next_actions, dones = batch
    actions = actions.long()
    fw = lambda sd, o, a: network.forward(sd, o, num_heads, actions=a, identity=identity)   # actions only feed --a-embed nets
    q = fw(policy_sd, obss, actions)                                     # :215
    q = q.gather(2, actions).squeeze(-1)                                 # :219
    with torch.no_grad():
        a_star = fw(policy_sd, next_obss, next_actions).argmax(dim=2, keepdim=True)             # :226-229
        q_next = fw(target_sd, next_obss, next_actions).gather(2, a_star).squeeze(-1)            # :230-233
        y = rewards.squeeze(-1) + (1 - dones.long().squeeze(-1)) * (q_next * gamma)            # :236-238
    h = history or q.shape[1]
    q, y = q[:, -h:], y[:, -h:]                                          # :240-241
    loss = ((q - y) ** 2).mean()                                         # :243
    return loss, q, y


def clip_coef(grads, max_norm=1.0):
    """clip_grad_norm_: total = || [||g_i||_2]_i ||_2 ; coef = min(1, max_norm / (total + 1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    if not math.isfinite(total.item()):
        raise RuntimeError("The total norm for gradients is non-finite, so it cannot be clipped.")
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_step(p, g, m, v, t, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay / amsgrad), step index t >= 1; returns new (p, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
    p = p - (lr / (1 - b1 ** t)) * (m / denom)
    return p, m, v


def _gate_base(key: str) -> str:
    import re
    return re.sub(r"transformer_layers\.\d+\.(attn_gate|mlp_gate)", r"transformer_layers.0.\1", key)


class TrainerOracle:
    """Holds policy/target state dicts + Adam moments and performs DtqnAgent.train() on a given batch."""

    def __init__(self, sd, num_heads, pos="learned", lr=3e-4, gamma=0.99, grad_norm_clip=1.0,
                 target_update_frequency=10_000, history=None, identity=False):
        self.identity = identity
        self.policy = {k: v.clone() for k, v in sd.items()}
        self.target = {k: v.clone() for k, v in sd.items()}             # dqn.py:46-50
        # GRU gates are ONE module shared by all layers (dtqn.py:107-131): the state_dict repeats them per layer; only the
        # layer-0 entry is a leaf, the others alias it so autograd sums the layers' contributions
        self.tied = {k: _gate_base(k) for k in sd if _gate_base(k) != k}
        self._tie()
        self.keys = [k for k in network.trainable_keys(sd, pos) if k not in self.tied]
        self.m = {k: torch.zeros_like(sd[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(sd[k]) for k in self.keys}
        self.num_heads, self.lr, self.gamma, self.clip = num_heads, lr, gamma, grad_norm_clip
        self.tuf, self.history = target_update_frequency, history
        self.num_train_steps = 0

    def _tie(self):
        for k, base in self.tied.items():
            self.policy[k] = self.policy[base]

    def train_on_batch(self, batch):
        for k in self.keys:
            self.policy[k].requires_grad_(True)
        loss, q, y = td_loss(self.policy, self.target, batch, self.num_heads, self.gamma, self.history, self.identity)
        grads = torch.autograd.grad(loss, [self.policy[k] for k in self.keys])
        for k in self.keys:
            self.policy[k].requires_grad_(False)
        total, coef = clip_coef(grads, self.clip)
        self.num_train_steps += 1
        out_grads = {}
        with torch.no_grad():
            for k, g in zip(self.keys, grads):
                out_grads[k] = g.clone()
                g = g * coef
                self.policy[k], self.m[k], self.v[k] = adam_step(self.policy[k], g, self.m[k], self.v[k],
                                                                 self.num_train_steps, self.lr)
        self._tie()
        if self.num_train_steps % self.tuf == 0:                        # dtqn.py:268-269
            self.target = {k: v.clone() for k, v in self.policy.items()}
        stats = dict(loss=loss.item(), grad_norm=total.item(), q_max=q.max().item(), q_mean=q.mean().item(),
                     q_min=q.min().item(), t_max=y.max().item(), t_mean=y.mean().item(), t_min=y.min().item())
        return stats, out_grads
