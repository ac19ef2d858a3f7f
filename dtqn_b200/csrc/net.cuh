// Parameter / workspace layout of the DTQN Q-network shared by the forward, backward and optimiser kernels.
// Mirrors the module tree of dtqn/networks/dtqn.py:41-156 (state_dict names in SURVEY.md section 8 a11).
#pragma once
#include "common.cuh"

#define DTQN_MAX_LAYERS 8
#define DTQN_MAX_GROUPS 3

// offsets (in floats) into the flat parameter buffer; every tensor starts on a 16-byte boundary
struct LayerOff {
    long long ln1_w, ln1_b, ln2_w, ln2_b, in_w, in_b, out_w, out_b, f1_w, f1_b, f2_w, f2_b;
};
struct GateOff { long long w_r, u_r, w_z, b_z, u_z, w_g, u_g; };   // GRUGate (gates.py:13-18): six [d,d] matrices + w_z.bias
struct NetLayout {
    long long emb_table;   // [vocab, e]   (discrete only, else -1)   obs_embedding.observation_embedding.0.weight
    long long emb_w;       // [d - action_dim, K_in]    ...observation_embedding(.2).weight
    long long emb_b;       // [d - action_dim]
    long long act_table;   // [A, action_dim]  action_embedding.embedding.0.weight (action_dim > 0, else -1)
    long long pos;         // [ctx, d]     position_embedding.position_encoding
    GateOff gate[2];       // shared attention gate, shared mlp gate (gate_gru only)
    LayerOff layer[DTQN_MAX_LAYERS];
    long long h1_w, h1_b;  // ffn.0  [d,d], [d]
    long long h2_w, h2_b;  // ffn.2  [A,d], [A]
    long long total;       // floats, padded
    int k_in;              // O (continuous) or O*e (discrete)
};

static inline long long al4(long long x) { return (x + 3) & ~3ll; }

// any ablation flag: the general kernel-per-op path (net_var.cu) instead of the fused default-architecture kernels
static inline bool cfg_is_variant(const dtqn_net_cfg& c) {
    return c.action_dim > 0 || c.identity || c.gate_gru || c.dropout > 0.f;
}

static inline int net_layout(const dtqn_net_cfg& c, NetLayout& L) {
    if (c.n_layers < 1 || c.n_layers > DTQN_MAX_LAYERS) return DTQN_E_ARG;
    if (cfg_is_variant(c)) {
        if (c.d_model < 8 || c.d_model > 256 || c.d_model % 4) return DTQN_E_UNSUPPORTED;
        if (c.action_dim < 0 || c.action_dim >= c.d_model || c.action_dim % 4) return DTQN_E_UNSUPPORTED;
        if (c.dropout < 0.f || c.dropout >= 1.f || (c.dropout > 0.f && !c.dropout_state)) return DTQN_E_ARG;
    } else if (c.d_model != 64 && c.d_model != 128) return DTQN_E_UNSUPPORTED;
    if (c.n_heads <= 0 || c.d_model % c.n_heads || (c.d_model / c.n_heads) > 32 || (c.d_model / c.n_heads) % 4)
        return DTQN_E_UNSUPPORTED;
    if (c.context_len < 1 || c.context_len > 128 || c.obs_dim < 1 || c.obs_dim > 16 || c.num_actions < 1 ||
        c.num_actions > 32) return DTQN_E_UNSUPPORTED;
    if (c.discrete && (c.vocab < 1 || c.vocab > 64 || c.embed_per_obs < 1 || c.embed_per_obs > 16)) return DTQN_E_UNSUPPORTED;
    const long long d = c.d_model, d_obs = d - c.action_dim;
    long long o = 0;
    L.k_in = c.discrete ? c.obs_dim * c.embed_per_obs : c.obs_dim;
    if (c.discrete) { L.emb_table = o; o = al4(o + (long long)c.vocab * c.embed_per_obs); } else L.emb_table = -1;
    L.emb_w = o; o = al4(o + d_obs * L.k_in);
    L.emb_b = o; o = al4(o + d_obs);
    if (c.action_dim > 0) { L.act_table = o; o = al4(o + (long long)c.num_actions * c.action_dim); } else L.act_table = -1;
    L.pos = o;   o = al4(o + (long long)c.context_len * d);
    for (int i = 0; i < c.n_layers; ++i) {
        LayerOff& l = L.layer[i];
        l.ln1_w = o; o += d; l.ln1_b = o; o += d; l.ln2_w = o; o += d; l.ln2_b = o; o += d;
        l.in_w = o; o += 3 * d * d; l.in_b = o; o += 3 * d;
        l.out_w = o; o += d * d;    l.out_b = o; o += d;
        l.f1_w = o; o += 4 * d * d; l.f1_b = o; o += 4 * d;
        l.f2_w = o; o += 4 * d * d; l.f2_b = o; o += d;
        if (i == 0 && c.gate_gru) {
            for (int k = 0; k < 2; ++k) {
                GateOff& q = L.gate[k];
                q.w_r = o; o += d * d; q.u_r = o; o += d * d; q.w_z = o; o += d * d; q.b_z = o; o += d;
                q.u_z = o; o += d * d; q.w_g = o; o += d * d; q.u_g = o; o += d * d;
            }
        }
    }
    L.h1_w = o; o += d * d; L.h1_b = o; o += d;
    L.h2_w = o; o = al4(o + (long long)c.num_actions * d);
    L.h2_b = o; o = al4(o + c.num_actions);
    L.total = o;
    return 0;
}

// Activation workspace (floats) for G groups x n_seq sequences x L tokens.  With save = 1 every layer keeps its own
// buffers (needed by the backward pass); with save = 0 the per-layer buffers alias one set.
struct LayerAct {
    float *qkv, *o, *r1, *st1, *x1, *h, *r2, *st2, *x2;
};
struct NetAct {
    float* lut;    // discrete observations: [groups][O][vocab][d] lookup table of the token embedding (embed_lut_kernel)
    float* x0;
    LayerAct layer[DTQN_MAX_LAYERS];
    float* hh;
    float* q;      // [T, A]
    long long total;
};

static inline long long net_act_layout(const dtqn_net_cfg& c, long long T, int save, float* base, NetAct& A) {
    const long long d = c.d_model;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o = al4(o + n); return p; };
    A.lut = c.discrete ? take((long long)DTQN_MAX_GROUPS * c.obs_dim * c.vocab * d) : nullptr;
    A.x0 = take(T * d);
    for (int i = 0; i < c.n_layers; ++i) {
        if (i == 0 || save) {
            LayerAct& l = A.layer[i];
            l.qkv = take(T * 3 * d); l.o = take(T * d); l.r1 = take(T * d); l.st1 = take(T * 2); l.x1 = take(T * d);
            l.h = take(T * 4 * d);   l.r2 = take(T * d); l.st2 = take(T * 2);
            l.x2 = take(T * d);
        } else {
            A.layer[i] = A.layer[i - 1];
            // ping-pong the layer output so a layer never writes the buffer it reads as its input
            A.layer[i].x2 = (i & 1) ? A.x0 : A.layer[0].x2;
        }
    }
    A.hh = take(T * d);
    A.q = take(T * c.num_actions);
    A.total = o;
    return o;
}

// ---- shared by the CUDA-core and tcgen05 Linear kernels -----------------------------------------------------------------
struct GroupPtrs { const float* p[DTQN_MAX_GROUPS]; };
struct GroupSrc { dtqn_obs_src s[DTQN_MAX_GROUPS]; };

enum { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_RES_LN = 2 };

struct LinArgs {
    const float* X;        // [G*Tg, K]
    float* Y;              // [G*Tg, N]
    GroupPtrs P;
    long long w_off, b_off;
    int Tg, N, K;
    // EPI_RES_LN only (N == BN == d_model):
    const float* R;        // residual input [G*Tg, N]
    long long gamma_off, beta_off;
    float* r_save;         // relu(a) [G*Tg, N]   (nullable)
    float* st_save;        // (mean, rstd) [G*Tg, 2] (nullable)
};


// ablation-flag networks (net_var.cu): general fp32 kernel-per-op forward / backward
long long var_workspace_floats(const dtqn_net_cfg& c, long long T);
int var_forward(const dtqn_net_cfg& c, const NetLayout& lay, int G, const float* const* params, const dtqn_obs_src* src,
                int n_seq, int L, int q_mode, int save, float* ws, long long ws_floats, float* q_out, cudaStream_t st);
long long var_bwd_scratch_floats(const dtqn_net_cfg& c, long long T0);
int var_td_backward(const dtqn_net_cfg& c, const NetLayout& lay, const float* params, const dtqn_obs_src* obs_src,
                    const float* dq, int B, int L, float* ws, long long ws_floats, float* scratch, float* grads,
                    cudaStream_t st);

// sequence-resident fused forward (net_seq.cu)
bool seq_forward_supported(const dtqn_net_cfg& c, int L);
// wt: per-group k-major fp32 weight copies (TcPackTable::wt_off region of the packed image) or all-NULL -> weights are read
// (and transposed on the fly) from the flat parameters
int launch_seq_forward(const dtqn_net_cfg& c, const NetLayout& lay, const NetAct& act, const GroupPtrs& P, int G, int n_seq,
                       int L, int save, float* q_out, cudaStream_t st, const float* const* wt = nullptr);
