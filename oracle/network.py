"""CPU oracle restatement of the DTQN Q-network forward in plain torch fp32 (test infrastructure only).

Follows dtqn/networks/dtqn.py:158-218, transformer.py:63-78 (post-LN block, ReLU on the sub-layer outputs,
causal additive -inf mask :49-53; :86-101 identity-map reordering), representations.py:17-23,25-52,64-75,146-155,
position_encodings.py:19-43, gates.py:5-41 (ResGate, GRUGate) and utils/torch_utils.py:4-15 (init).  ``nn.MultiheadAttention`` is restated in closed
form: qkv = x W_in^T + b_in with rows [0:d]=Q, [d:2d]=K, [2d:3d]=V; heads are contiguous d/H
slices; S = (Q/sqrt(hd)) K^T + mask; P = softmax(S); O = concat_h(P V) W_out^T + b_out; LayerNorm eps 1e-5.
The functions take a reference-compatible ``state_dict`` (key names of SURVEY.md section 8 a11) so the same weights
drive the reference module, this oracle and the CUDA path.
"""
import math

import torch
import torch.nn.functional as F


def num_layers_of(sd) -> int:
    n = 0
    while f"transformer_layers.{n}.layernorm1.weight" in sd:
        n += 1
    return n


def init_state_dict(obs_dim, num_actions, embed_per_obs_dim, inner_embed, num_heads, num_layers, history_len,
                    discrete=False, vocab_size=None, pos="learned", generator=None, device="cpu"):
    """Fresh parameters with the reference's initialisation (utils/torch_utils.py:4-15): N(0, 0.02) for every
    Linear / Embedding / in_proj / out_proj weight, zero biases, LayerNorm (1, 0), learned position table zeros."""
    d = inner_embed
    g = generator

    def nrm(*shape):
        return torch.empty(*shape, device=device).normal_(0.0, 0.02, generator=g)

    sd = {}
    if discrete:
        sd["obs_embedding.observation_embedding.0.weight"] = nrm(vocab_size, embed_per_obs_dim)
        sd["obs_embedding.observation_embedding.2.weight"] = nrm(d, embed_per_obs_dim * obs_dim)
        sd["obs_embedding.observation_embedding.2.bias"] = torch.zeros(d, device=device)
    else:
        sd["obs_embedding.observation_embedding.weight"] = nrm(d, obs_dim)
        sd["obs_embedding.observation_embedding.bias"] = torch.zeros(d, device=device)
    if pos == "sin":
        position = torch.arange(history_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
        pe = torch.zeros(1, history_len, d)
        pe[0, :, 0::2] = torch.sin(position * div)
        pe[0, :, 1::2] = torch.cos(position * div)
        sd["position_embedding.position_encoding"] = pe.to(device)
    else:
        sd["position_embedding.position_encoding"] = torch.zeros(1, history_len, d, device=device)
    mask = torch.triu(torch.ones(history_len, history_len, device=device), diagonal=1)
    mask[mask.bool()] = -float("inf")
    for i in range(num_layers):
        p = f"transformer_layers.{i}."
        sd[p + "attn_mask"] = mask.clone()
        sd[p + "layernorm1.weight"] = torch.ones(d, device=device)
        sd[p + "layernorm1.bias"] = torch.zeros(d, device=device)
        sd[p + "layernorm2.weight"] = torch.ones(d, device=device)
        sd[p + "layernorm2.bias"] = torch.zeros(d, device=device)
        sd[p + "attention.in_proj_weight"] = nrm(3 * d, d)
        sd[p + "attention.in_proj_bias"] = torch.zeros(3 * d, device=device)
        sd[p + "attention.out_proj.weight"] = nrm(d, d)
        sd[p + "attention.out_proj.bias"] = torch.zeros(d, device=device)
        sd[p + "ffn.0.weight"] = nrm(4 * d, d)
        sd[p + "ffn.0.bias"] = torch.zeros(4 * d, device=device)
        sd[p + "ffn.2.weight"] = nrm(d, 4 * d)
        sd[p + "ffn.2.bias"] = torch.zeros(d, device=device)
    sd["ffn.0.weight"] = nrm(d, d)
    sd["ffn.0.bias"] = torch.zeros(d, device=device)
    sd["ffn.2.weight"] = nrm(num_actions, d)
    sd["ffn.2.bias"] = torch.zeros(num_actions, device=device)
    return sd


def embed(sd, obss):
    """representations.py:17-23 (+ :47-51 discrete, :74 continuous).  obss [B,L,O] float (continuous) or int."""
    if "obs_embedding.observation_embedding.0.weight" in sd:
        e = sd["obs_embedding.observation_embedding.0.weight"][obss.long()]          # [B,L,O,E]
        e = e.flatten(start_dim=-2)                                                   # nn.Flatten(-2)
        return e @ sd["obs_embedding.observation_embedding.2.weight"].T + sd["obs_embedding.observation_embedding.2.bias"]
    return obss.float() @ sd["obs_embedding.observation_embedding.weight"].T + sd["obs_embedding.observation_embedding.bias"]


def attention(sd, prefix, x, num_heads):
    """nn.MultiheadAttention(x, x, x, attn_mask=causal) in closed form (transformer.py:64-70)."""
    B, L, d = x.shape
    hd = d // num_heads
    qkv = x @ sd[prefix + "attention.in_proj_weight"].T + sd[prefix + "attention.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(B, L, num_heads, hd).transpose(1, 2) / math.sqrt(hd)
    k = k.view(B, L, num_heads, hd).transpose(1, 2)
    v = v.view(B, L, num_heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) + sd[prefix + "attn_mask"][:L, :L]
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, d)
    return o @ sd[prefix + "attention.out_proj.weight"].T + sd[prefix + "attention.out_proj.bias"]


def gate(sd, prefix, x, y):
    """gates.py: ResGate (:40-41) unless the layer carries GRUGate parameters (:5-31; the SAME two gate modules are shared by
    every layer, dtqn.py:107-131, so the state_dict repeats them under each layer's prefix)."""
    if prefix + "w_r.weight" not in sd:
        return x + y
    lin = lambda name, t: t @ sd[prefix + name + ".weight"].T
    z = torch.sigmoid(lin("w_z", y) + sd[prefix + "w_z.bias"] + lin("u_z", x))
    r = torch.sigmoid(lin("w_r", y) + lin("u_r", x))
    h = torch.tanh(lin("w_g", y) + lin("u_g", r * x))
    return (1.0 - z) * x + z * h


def layer_forward(sd, prefix, x, num_heads, identity=False):
    """transformer.py:63-78 (post-LN block) or :86-101 (TransformerIdentityLayer, the GTrXL identity-map reordering)."""
    d = x.shape[-1]
    ln1 = lambda t: F.layer_norm(t, (d,), sd[prefix + "layernorm1.weight"], sd[prefix + "layernorm1.bias"], 1e-5)
    ln2 = lambda t: F.layer_norm(t, (d,), sd[prefix + "layernorm2.weight"], sd[prefix + "layernorm2.bias"], 1e-5)
    ffn = lambda t: torch.relu(t @ sd[prefix + "ffn.0.weight"].T + sd[prefix + "ffn.0.bias"]) @ sd[prefix + "ffn.2.weight"].T + sd[prefix + "ffn.2.bias"]
    if identity:
        x = gate(sd, prefix + "attn_gate.", x, torch.relu(attention(sd, prefix, ln1(x), num_heads)))
        return gate(sd, prefix + "mlp_gate.", x, torch.relu(ffn(ln2(x))))
    x = ln1(gate(sd, prefix + "attn_gate.", x, torch.relu(attention(sd, prefix, x, num_heads))))
    return ln2(gate(sd, prefix + "mlp_gate.", x, torch.relu(ffn(x))))


def forward(sd, obss, num_heads, actions=None, identity=False):
    """dtqn.py:158-218 with bag_size = 0, dropout = 0.  Defaults (run.py :98-103,151-153,173-175): action_dim = 0 (no
    ``action_embedding.*`` keys), gate = res, identity = False.  With an action embedding (dtqn.py:184-192) the previous
    action's embedding is concatenated IN FRONT of the observation embedding; position 0 gets zeros."""
    L = obss.shape[1]
    assert L <= sd["position_embedding.position_encoding"].shape[1], "Cannot forward, history is longer than expected."
    x = embed(sd, obss)
    if "action_embedding.embedding.0.weight" in sd:
        a = sd["action_embedding.embedding.0.weight"][actions.long()].flatten(start_dim=-2)       # [B, L, action_dim]
        if L > 1:
            a = torch.roll(a, 1, 1)
            a = torch.cat([torch.zeros_like(a[:, :1]), a[:, 1:]], dim=1)
        x = torch.cat([a, x], dim=-1)
    x = x + sd["position_embedding.position_encoding"][:, :L, :]
    for i in range(num_layers_of(sd)):
        x = layer_forward(sd, f"transformer_layers.{i}.", x, num_heads, identity)
    h = torch.relu(x @ sd["ffn.0.weight"].T + sd["ffn.0.bias"])
    return h @ sd["ffn.2.weight"].T + sd["ffn.2.bias"]


def trainable_keys(sd, pos="learned"):
    """Keys that receive gradients: everything but attn_mask (requires_grad=False, transformer.py:49-53) and a
    non-learned position table (position_encodings.py:35,49-51), in state_dict order."""
    ks = []
    for k in sd:
        if k.endswith("attn_mask"):
            continue
        if k == "position_embedding.position_encoding" and pos != "learned":
            continue
        ks.append(k)
    return ks


class ModuleNet(torch.nn.Module):
    """The same network built from the torch.nn blocks the reference instantiates (nn.MultiheadAttention batch_first with a
    float causal mask, nn.LayerNorm, nn.Linear / nn.Embedding; dtqn.py:61-156, transformer.py:20-53), parameter names chosen
    so a reference-compatible state_dict loads unchanged.  Default flags only (gate res, no identity / action embedding).
    Used by the CPU baseline loop so the port executes the reference's own library calls op for op (the closed form above
    is the numerical checker; ``tests/test_oracle_vs_golden.py`` pins the two against each other)."""

    class _Block(torch.nn.Module):
        def __init__(self, d, heads, ctx):
            super().__init__()
            nn = torch.nn
            self.layernorm1, self.layernorm2 = nn.LayerNorm(d), nn.LayerNorm(d)
            self.attention = nn.MultiheadAttention(d, heads, dropout=0.0, batch_first=True)
            self.ffn = nn.Sequential(nn.Linear(d, 4 * d), nn.ReLU(), nn.Linear(4 * d, d), nn.Dropout(0.0))
            self.attn_mask = nn.Parameter(torch.zeros(ctx, ctx), requires_grad=False)

        def forward(self, x):
            L = x.size(1)
            a, _ = self.attention(x, x, x, attn_mask=self.attn_mask[:L, :L], average_attn_weights=True)
            x = self.layernorm1(x + torch.relu(a))
            return self.layernorm2(x + torch.relu(self.ffn(x)))

    class _Holder(torch.nn.Module):
        pass

    def __init__(self, sd, num_heads):
        super().__init__()
        nn = torch.nn
        ctx, d = sd["position_embedding.position_encoding"].shape[1:]
        self.obs_embedding = ModuleNet._Holder()
        if "obs_embedding.observation_embedding.0.weight" in sd:
            V, E = sd["obs_embedding.observation_embedding.0.weight"].shape
            K = sd["obs_embedding.observation_embedding.2.weight"].shape[1]
            self.obs_embedding.observation_embedding = nn.Sequential(nn.Embedding(V, E), nn.Flatten(start_dim=-2), nn.Linear(K, d))
        else:
            self.obs_embedding.observation_embedding = nn.Linear(sd["obs_embedding.observation_embedding.weight"].shape[1], d)
        self.position_embedding = ModuleNet._Holder()
        self.position_embedding.position_encoding = nn.Parameter(torch.zeros(1, ctx, d))
        self.transformer_layers = nn.Sequential(*[ModuleNet._Block(d, num_heads, ctx) for _ in range(num_layers_of(sd))])
        A = sd["ffn.2.weight"].shape[0]
        self.ffn = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, A))
        self.load_state_dict(sd)

    def forward(self, obss):
        B, L = obss.shape[:2]
        tok = self.obs_embedding.observation_embedding(obss.reshape(B * L, *obss.shape[2:])).reshape(B, L, -1)
        x = tok + self.position_embedding.position_encoding[:, :L, :]
        return self.ffn(self.transformer_layers(x))[:, -L:, :]
