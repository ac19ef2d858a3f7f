/*
 * dtqn_b200.h -- C ABI of libdtqn_b200.so: the B200-native (sm_100a) DTQN training hot path.
 *
 * Drop-in boundary for the data-parallel hot path of kevslinger/DTQN (pure Python; it has no FFI of its own,
 * SURVEY.md section 8b), so every entry point below cites the reference *Python* interface it replaces.
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + explicit sizes inside POD structs, a cudaStream_t passed as void*;
 *   - return int: 0 = ok, < 0 = invalid argument (DTQN_E_*), > 0 = cudaError_t of a failed launch;
 *   - compute entry points never allocate, never synchronise, never throw; the caller (PyTorch) owns every buffer and
 *     keeps it alive until the stream has passed the call.  Exceptions, named where they are declared: dtqn_p2p_alloc /
 *     _open / _close / _free own the IPC-exported exchange buffer (cudaMalloc + cudaIpc*), dtqn_p2p_error /
 *     dtqn_tc_error / dtqn_profile_read copy one word back (they synchronise);
 *   - no per-call global state: one process per GPU, re-entrant across processes.  The dtqn_set_* switches and
 *     dtqn_profile_enable are process-global tuning / measurement knobs (defaults are the product path); set them before
 *     launching work, not concurrently with it.
 */
#ifndef DTQN_B200_H
#define DTQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTQN_ABI_VERSION 5

#define DTQN_E_ARG      (-1)  /* null pointer / out-of-range size */
#define DTQN_E_UNSUPPORTED (-2)

enum { DTQN_ENV_CARFLAG = 0, DTQN_ENV_MEMORY = 1 };

/* action_mode of dtqn_env_step */
enum {
    DTQN_ACT_GIVEN = 0,      /* actions[] supplied by the caller                         (env.step(a), run.py:368) */
    DTQN_ACT_RANDOM = 1,     /* a = RNG.rng.integers(A)                                  (prepopulate, run.py:394)  */
    DTQN_ACT_EPS_GREEDY = 2  /* RNG.rng.random() < eps ? integers(A) : argmax q_last     (agents/dtqn.py:78-107)    */
};

/* ---- batched POMDP environments: envs/car_flag.py:18-159, envs/memory_cards.py:43-116, gym TimeLimit ------------
 * SoA over n_envs lockstep instances.  PCG64 state is numpy's bit_generator.state verbatim
 * (state, inc as hi/lo 64-bit words; has_uint32; uinteger) and is seeded on the host BY numpy itself. */
typedef struct dtqn_env {
    int32_t kind;               /* DTQN_ENV_* */
    int32_t n_envs;
    int32_t obs_dim;            /* 3 (CarFlag) | 10 (Memory) */
    int32_t num_actions;        /* 3 | 10 */
    int32_t max_episode_steps;  /* TimeLimit: 200 | 50 (envs/__init__.py:31-48) */
    int32_t stat_episodes_per_env; /* > 0: only the first k finished episodes of each env enter ep_stats / env_acc
                                      (run.evaluate's fixed episode count, run.py:214-233); 0: every episode */
    uint64_t* rng;              /* [4][n] env np_random: state_hi, state_lo, inc_hi, inc_lo */
    uint32_t* rng_buf;          /* [2][n] has_uint32, uinteger */
    uint64_t* arng;             /* [4][n] agent-side stream = the reference's global RNG.rng (utils/random.py:31) */
    uint32_t* arng_buf;         /* [2][n] */
    double*   pos;              /* [n] CarFlag position (f64) */
    double*   vel;              /* [n] CarFlag velocity (f64) */
    int8_t*   heaven;           /* [n] CarFlag heaven side: +1 / -1 */
    uint64_t* cards;            /* [n] Memory hidden cards, 4 bits per card (values 1..5) */
    uint64_t* shown;            /* [n] Memory observation, 4 bits per card (0 hidden, 1..5 shown, 6 removed) */
    int32_t*  cur;              /* [n] Memory currently shown card */
    int32_t*  elapsed;          /* [n] TimeLimit._elapsed_steps */
    uint8_t*  done_flag;        /* [n] scratch: episode ended in this step (1: its successor is stored in the replay, 2: it is
                                   not -- see dtqn_replay.record_every); consumed by the roll kernel */
    int32_t*  block_counts;     /* [ceil(n/256)] scratch: episodes finished per CTA in this step */
    int64_t*  ep_stats;         /* [4] running sums over finished episodes: return, length, successes, episodes */
    int32_t*  ep_return;        /* [n] return of the running episode (rewards are integers in both envs) */
    int32_t*  env_acc;          /* nullable [n][4] per-env sums over its counted episodes: episodes, return, length, successes */
} dtqn_env;

/* ---- device-resident episode-major replay buffer: dtqn/buffers/replay_buffer.py:19-135 --------------------------
 * Same arrays / fill values as the reference (obss fill = obs_mask, actions 0, rewards 0, dones TRUE).
 * With n_envs lockstep envs there are n_envs open episode slots; slots are handed out in episode-START order
 * (ring of n_slots, FIFO overwrite like pos[0] % max_size, :77), deterministically by env index. */
typedef struct dtqn_replay {
    int32_t n_slots;            /* max_size = buffer_size // max_episode_steps (:27) */
    int32_t max_episode_steps;
    int32_t obs_dim;
    int32_t context_len;
    float   obs_mask;           /* -5 (Box) | 8 (MultiDiscrete), utils/env_processing.py:100-118 */
    int32_t record_every;       /* <= 1: every episode is stored (the reference); K > 1: each env stores every K-th of its
                                   episodes, so a ring of S slots spans K times more loop iterations (with thousands of
                                   lockstep envs per update almost all transitions are never sampled anyway; what the
                                   reference's 500 000-transition buffer provides is a memory of ~25 % of the RUN, i.e. of
                                   many target-network periods, and this keeps that horizon at scale) */
    float*   obss;              /* [S, E+1, O] */
    uint8_t* actions;           /* [S, E+1]    */
    float*   rewards;           /* [S, E]      */
    uint8_t* dones;             /* [S, E]      */
    int32_t* episode_lengths;   /* [S] (uint8 in the reference; widened) */
    uint8_t* slot_open;         /* [S] 1 while an env is still writing the slot (excluded from sampling, :141-145) */
    int64_t* counters;          /* [4] episodes started (current), started (next), completed, dropped */
    int32_t* env_slot;          /* [n_envs] open slot of each env, -1 = episode not recorded */
    int32_t* env_prev_len;      /* [n_envs] length of the episode previously stored in the env's open slot */
} dtqn_replay;

/* ---- acting history: utils/context.py:8-111 (obs window; the action window only feeds --a-embed > 0 networks) ------- */
typedef struct dtqn_context {
    int32_t context_len;
    int32_t obs_dim;
    int32_t trunc_obs;          /* 1 = reproduce the reference's int64 context (obs truncated toward 0, A-Q2) */
    float   obs_mask;
    float*   obs;               /* [n_envs, ctx, O] ring; token j of the window = ring[(t+1-n+j) % ctx] */
    int32_t* timestep;          /* [n_envs] Context.timestep */
    uint8_t* action;            /* nullable [n_envs, ctx] ring of Context.action (utils/context.py:50,77): row t % ctx = the
                                   action that led to the observation of timestep t; row 0 of a fresh context = the first of
                                   the ctx random padding actions Context.reset draws (the only one a window can ever show) */
} dtqn_context;

typedef struct dtqn_step_io {
    int32_t action_mode;        /* DTQN_ACT_* */
    int32_t _pad;
    double  epsilon;            /* f64 like the reference: RNG.rng.random() < epsilon compares two doubles (agents/dtqn.py:78) */
    int32_t* actions;           /* [n] in (GIVEN) / out (RANDOM, EPS_GREEDY) */
    const float* q_last;        /* [n, A] greedy Q of the last context position (EPS_GREEDY) */
    float*   obs_out;           /* [n, O] observation returned by env.step (terminal obs when done); nullable */
    float*   reward_out;        /* [n] nullable */
    uint8_t* done_out;          /* [n] env done incl. TimeLimit; nullable */
    uint8_t* truncated_out;     /* [n] info["TimeLimit.truncated"]; nullable */
    uint8_t* success_out;       /* [n] info["is_success"]; nullable */
    const double* epsilon_dev;  /* nullable: device f64 scalar read instead of `epsilon` (CUDA-graph replay) */
} dtqn_step_io;

int dtqn_version(void);

/* Exploration schedule on the device (LinearAnneal.anneal, utils/epsilon_anneal.py:33-34; run.py:298): writes the current
 * value to *eps_out (the `epsilon_dev` of the next dtqn_env_step) and advances state = {val, min, duration} (3 doubles,
 * device memory) by val <- max(min, val - (val - min) / duration) in double precision, bit-identical to the host loop. */
int dtqn_eps_anneal(double* state, double* eps_out, void* stream);

/* env.reset() of every instance + agent.context_reset(obs) (run.py:287-288): allocates slots 0..n-1, stores the first
 * observation (ReplayBuffer.store_obs :88-92) and resets the contexts.  rb / cx may be NULL. */
int dtqn_env_reset_all(const dtqn_env* env, const dtqn_replay* rb, const dtqn_context* cx, void* stream);

/* One lockstep iteration of run.step (run.py:356-377) for every env: pick the action, env.step, TimeLimit,
 * agent.observe -> Context.add_transition + ReplayBuffer.store with buffer_done = done & !truncated, and, where the
 * episode ended, replay_buffer.flush() + env.reset() + agent.context_reset() (run.py:293-296).
 * rb == NULL -> no replay writes (evaluation mode, agents/dtqn.py:159).  cx may be NULL. */
int dtqn_env_step(const dtqn_env* env, const dtqn_replay* rb, const dtqn_context* cx, const dtqn_step_io* io,
                  void* stream);

/* ReplayBuffer.sample index draw (:141-156) on the device (counter-based RNG; the reference uses CPython `random`,
 * which is not replicated -- parity tests inject indices).  Uniform over completed episodes, then uniform start in
 * [0, max(0, eplen - ctx)].  `seed`/`draw` select the stream. */
int dtqn_replay_sample_indices(const dtqn_replay* rb, int32_t batch, uint64_t seed, uint64_t draw,
                               uint64_t* draw_counter /* device, nullable: added to `draw`, then incremented */,
                               int32_t* episodes_out, int32_t* starts_out, void* stream);

/* ReplayBuffer.sample payload gather (:157-168) for given (episode, start) pairs.  obs_win is the (L+1)-row window
 * (obss == rows [0,L), next_obss == rows [1,L+1) of the same window); act_win likewise.
 *   obs_win [B, L+1, O] f32, act_win [B, L+1] u8, rew [B, L] f32, done [B, L] u8, eplen [B] i32 (clipped to L). */
int dtqn_replay_gather(const dtqn_replay* rb, int32_t batch, const int32_t* episodes, const int32_t* starts,
                       float* obs_win, uint8_t* act_win, float* rew, uint8_t* done, int32_t* eplen, void* stream);


/* ==== DTQN Q-network: dtqn/networks/dtqn.py:15-218, transformer.py:6-78, representations.py:9-75,146-155,
 *      position_encodings.py:14-51, gates.py:34-41 (ResGate) ==================================================== */
typedef struct dtqn_net_cfg {
    int32_t obs_dim;        /* O: length of the observation vector */
    int32_t num_actions;    /* A */
    int32_t d_model;        /* inner_embed_size (64 | 128) */
    int32_t n_heads;
    int32_t n_layers;
    int32_t context_len;    /* history_len: rows of the position table, max sequence length */
    int32_t discrete;       /* 0: Linear(O, d) on float obs; 1: Embedding(vocab, e) -> Flatten -> Linear(O*e, d) */
    int32_t vocab;          /* discrete only: obs_mask + 1 (utils/agent_utils.py:92-95) */
    int32_t embed_per_obs;  /* discrete only: e */
    int32_t pos_trainable;  /* 1 = learned position table (gets a gradient), 0 = sin / none (still added) */
    /* ---- ablation flags of run.py:98-103,151-167 (all 0 = the default architecture and its fused kernels); any of them
     *      selects the general fp32 kernel-per-op path of csrc/net_var.cu ---- */
    int32_t action_dim;     /* --a-embed: width of the previous-action embedding concatenated IN FRONT of the observation
                               embedding, which then has d_model - action_dim columns (dtqn.py:64-71,184-192) */
    int32_t identity;       /* --identity: TransformerIdentityLayer, LayerNorm before each sub-layer (transformer.py:81-101) */
    int32_t gate_gru;       /* --gate gru: GRUGate (gates.py:5-31); ONE attention gate and ONE mlp gate shared by all layers
                               (dtqn.py:107-131), so their gradients sum over layers */
    float   dropout;        /* --dropout p: token embedding (dtqn.py:195-199), attention probabilities (nn.MultiheadAttention,
                               transformer.py:31-36) and the FFN output (transformer.py:41); train-mode groups only */
    uint64_t* dropout_state; /* device uint64[1], required when dropout > 0: mask stream counter (advanced once per acting
                               forward and once per training step; the backward regenerates the forward's masks from it) */
} dtqn_net_cfg;

/* Where a group of sequences reads its observations.  Sequence i, token j reads row
 *   obs + i * seq_stride + j * obs_dim                                   (timestep == NULL), or, for the acting
 * context ring (utils/context.py), row ((t_i + 1 - n_i + j) mod ring_len) with n_i = min(ring_len, t_i + 1). */
typedef struct dtqn_obs_src {
    const float*   obs;
    int64_t        seq_stride;   /* floats */
    const int32_t* timestep;     /* nullable */
    int32_t        ring_len;
    float          obs_mask;     /* value of context rows not yet written (Context.reset fill, utils/context.py:46): -5 | 8 */
    const uint8_t* actions;      /* action_dim > 0 only: action of token j at actions + i * act_stride + j (same ring indexing
                                    as obs when timestep != NULL); token j is embedded with the action of token j - 1 */
    int64_t        act_stride;
    int32_t        train_mode;   /* 1: this group's network is in train() mode (dropout active); 0: eval() (target network) */
    int32_t        _pad;
} dtqn_obs_src;

/* Number of floats of the flat parameter buffer, and the offset of every tensor in it, in this fixed order:
 *   [emb_table (discrete only)], emb_w, emb_b, [act_table (action_dim > 0)], pos, then per layer: ln1_w, ln1_b, ln2_w,
 *   ln2_b, in_w, in_b, out_w, out_b, f1_w, f1_b, f2_w, f2_b -- with the two shared GRU gates right after layer 0 when
 *   gate_gru (attention gate then mlp gate, each w_r, u_r, w_z, b_z, u_z, w_g, u_g) -- then h1_w, h1_b, h2_w, h2_b.
 *   Returns the tensor count (or < 0). */
int64_t dtqn_net_param_count(const dtqn_net_cfg* cfg);
int dtqn_net_param_offsets(const dtqn_net_cfg* cfg, int64_t* offsets_out, int32_t max_entries);
/* Floats of activation workspace for n_tokens = groups * n_seq * seq_len tokens. */
int64_t dtqn_net_workspace_floats(const dtqn_net_cfg* cfg, int64_t n_tokens, int32_t save);

/* DTQN.forward (dtqn.py:158-218) for n_groups <= 3 groups of n_seq sequences of seq_len tokens, group g using the
 * flat parameter buffer params[g] and observation source src[g].
 *   q_mode 0: q_out[g, i, j, :] for every position (training, dtqn/agents/dtqn.py:215-233);
 *   q_mode 1: q_out[g, i, :] = Q at the last valid position n_i - 1 (acting, dtqn/agents/dtqn.py:101-107).
 * save = 1 keeps every layer's activations in `workspace` for dtqn_td_backward. */
int dtqn_forward(const dtqn_net_cfg* cfg, int32_t n_groups, const float* const* params,
                 const void* const* packed /* nullable: per-group dtqn_pack_weights images -> tcgen05 GEMMs */,
                 const dtqn_obs_src* src, int32_t n_seq, int32_t seq_len, int32_t q_mode, int32_t save, float* workspace,
                 int64_t workspace_floats, float* q_out, void* stream);

/* Measurement / test hook for --dropout: the keep-scale (0 or 1 / (1 - p)) the dropout kernels apply to element idx = 0..n-1
 * of mask site `site` (1 token embedding, 2 + 4 l attention probabilities of layer l, 3 + 4 l FFN output of layer l) when the
 * mask-stream counter (cfg.dropout_state[0]) holds `counter`. */
int dtqn_dropout_scales(uint64_t counter, uint32_t site, float p, int64_t n, float* out, void* stream);

/* tcgen05 operand images of the GEMM weights (in_proj / out_proj / ffn.0 / ffn.2 per layer + head ffn.0): each weight
 * split into bf16 hi + lo and tiled in the K-major shared-memory layout the tensor core reads, so the kernel fetches a
 * [N_TILE x 64] block with one TMA bulk copy.  Re-pack after every optimiser step / target update. */
int64_t dtqn_packed_bytes(const dtqn_net_cfg* cfg);
int dtqn_pack_weights(const dtqn_net_cfg* cfg, const float* params, void* packed, void* stream);
/* Groups with at least this many tokens use the tcgen05 path (default 4096); smaller ones stay on fp32 CUDA cores. */
int dtqn_set_tc_min_tokens(int32_t n_tokens);
/* 1 (default): persistent warp-specialised tcgen05 kernel where the weight image fits in shared memory; 0: simple one. */
int dtqn_set_tc_pipelined(int32_t on);
/* 1 (default): groups below the tcgen05 threshold with d_model 64, 8 heads, L <= 64 run every layer + the head in ONE
 * sequence-resident kernel (one CTA per sequence); 0: one kernel per GEMM / attention (the general path). */
int dtqn_set_seq_fused(int32_t on);
/* 1: the acting forward recomputes the token embedding inside the tcgen05 in_proj / out_proj kernels of layer 0
 * (continuous observations) instead of materialising it; 0 (default, measured faster): separate embed kernel. */
int dtqn_set_tc_fuse_embed(int32_t on);
/* 1 (default): on the tcgen05 path (d_model 64, inference) ffn.0 -> ReLU -> ffn.2 -> ReLU -> +residual -> LayerNorm run as
 * ONE kernel with the hidden activations kept in TMEM / shared memory; 0: two Linear launches. */
int dtqn_set_tc_fuse_ffn(int32_t on);
/* 1 (default): the attention core of d_model 64 / 8 heads / L <= 64 groups outside the sequence-resident kernel (the acting
 * forward) runs on the warp-level tensor cores (mma.sync TF32 hi/lo split); 0: fp32 CUDA-core kernel. */
int dtqn_set_attn_mma(int32_t on);
/* 1 (default): the acting forward (q_mode 1) of the default architecture (2 layers, d_model 64, 8 heads, continuous
 * observations, seq_len <= 52) with tcgen05 weight images runs embedding, the whole first layer, the final layer's in_proj
 * and the attention row of the last valid position as ONE persistent tcgen05 kernel per 128-token tile (no activations
 * in HBM); 0: one kernel per GEMM / attention. */
int dtqn_set_act_fused(int32_t on);
/* Discrete observations (Embedding -> Flatten -> Linear, representations.py:47-51).  1 (default): table-lookup form -- every CTA
 * builds LUT[feature][value][channel] = table[value] . W[channel, feature block] in shared memory and a token costs O row adds;
 * 2: the kernel that stages the transposed Linear weight and the table in shared memory (reference summation order);
 * 0: the generic per-channel kernel. */
int dtqn_set_embed_disc_fast(int32_t on);
/* debug: device buffer of >= 256 int64 that receives clock64 phase stamps of CTA 0 of the fused acting kernel (NULL: off). */
int dtqn_set_act_fused_timeline(void* device_buf);
/* 1 if a tcgen05 kernel ever timed out on an mbarrier (synchronises). */
int dtqn_tc_error(void);

/* Double-DQN sequence TD loss + backward (dtqn/agents/dtqn.py:215-256) for a forward made with n_groups = 3,
 * save = 1, q_mode = 0 over (policy|obs, policy|next_obs, target|next_obs).  Zeroes `grads` (flat, same layout as
 * the parameters) and accumulates dLoss/dparam of group 0 into it.
 *   act_win [B, L+1] u8 (actions = columns [0, L)), rew [B, L] f32, done [B, L] u8;
 *   stats_out[8] = loss, q_max, q_mean, q_min, target_max, target_mean, target_min, (grad_norm: dtqn_clip_adam). */
int dtqn_td_backward(const dtqn_net_cfg* cfg, const float* policy_params, const dtqn_obs_src* obs_src,
                     const float* q_all /* [3, B, L, A] */, const uint8_t* act_win, const float* rew, const uint8_t* done,
                     int32_t batch, int32_t seq_len, int32_t history, float gamma, float* workspace,
                     int64_t workspace_floats, float* scratch /* >= dtqn_td_scratch_floats */, float* grads,
                     float* stats_out, void* stream);
int64_t dtqn_td_scratch_floats(const dtqn_net_cfg* cfg, int32_t batch, int32_t seq_len);
/* 1 (default): weight-gradient GEMMs run on a library-owned side stream, forked / joined with events around each layer
 * (graph edges under capture); 0: everything on the caller's stream. */
int dtqn_set_parallel_wgrad(int32_t on);
/* Tokens per split-K chunk of the weight-gradient GEMMs (default 128, multiple of 16, >= 64): each chunk's partial tile is
 * stored and the last CTA of a tile adds the chunks in order (deterministic; no fp32 atomics). */
int dtqn_set_wgrad_chunk(int32_t tokens);
/* 1: the dependent kernels of the training step (sample, gather, forward, TD loss, backward chain, clip + Adam) are launched
 * with the programmatic-stream-serialization attribute and block in griddepcontrol.wait until their predecessor has
 * completed -- same results, the next grid is scheduled while the previous one drains.  0 (default): plain stream order.
 * No reference analogue (launch-mechanism knob). */
int dtqn_set_pdl(int32_t on);
/* Launch-shape knobs of the backward pass (no reference analogue; results stay deterministic for any setting).
 *  dgrad_rows 64 | 32 (default) | 16: token rows per CTA of the data-gradient GEMMs -- bitwise the same results, more CTAs
 *    in flight for the 1 600-token training batch;
 *  fuse_ln_bwd 1 (default) | 0: each LayerNorm backward runs as the epilogue of the data-gradient GEMM that produces its dy
 *    (4 launches fewer for 2 layers; the dgamma / dbeta partial grouping follows the GEMM's row tile, so 32 rows reproduce
 *    the stand-alone kernel bitwise);
 *  head_bwd_tokens 64 | 32 | 16 (default): tokens per CTA of the Q-head backward. */
int dtqn_set_dgrad_rows(int32_t rows);
int dtqn_set_fuse_ln_bwd(int32_t on);
int dtqn_set_head_bwd_tokens(int32_t tokens);

/* clip_grad_norm_(params, max_norm, error_if_nonfinite=True) + Adam.step (dtqn/agents/dtqn.py:257-265,
 * dtqn/agents/dqn.py:64): grads *= grad_scale (1/world after the allreduce), total = ||grads||_2,
 * coef = min(1, max_norm / (total + 1e-6)), Adam with bias correction at step *step_counter + 1 (incremented).
 * stats_out[7] = total norm; flags_out[0] = 1 if the norm is non-finite (parameters untouched in that case).
 * The 8 logged statistics of the step (the reference's RunningAverage.add(...item()) calls, agents/dtqn.py:245-263)
 * are appended to a device ring instead of being synchronised to the host every step. */
int dtqn_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                   float max_norm, float lr, float beta1, float beta2, float eps, int64_t* step_counter,
                   float* scratch /* >= 1024 floats */, float* stats_out, int32_t* flags_out,
                   float* stats_ring /* nullable [ring_len, 8]: row (step-1) % ring_len <- stats_out[0..8) */,
                   int32_t ring_len, void* stream);

/* ---- multi-GPU: gradient exchange fused with the optimiser over NVLink peer memory -----------------------------------
 * Replaces `allreduce(grads)` + dtqn_clip_adam when every rank of the node can map its peers' memory (SURVEY.md section
 * 8e: one collective per update, after loss.backward() and before clip_grad_norm_, dtqn/agents/dtqn.py:256-261).
 * Each rank allocates one exchange buffer (dtqn_p2p_alloc: header + n gradient floats, cudaMalloc'ed and IPC-exported),
 * lets its backward write the local gradient into `*grads_out`, exchanges the 64-byte handles out of band (any host
 * channel) and maps the peers with dtqn_p2p_open.  dtqn_allreduce_clip_adam then launches two kernels: (1) per-CTA
 * flag barrier across ranks -> sum of all ranks' gradients read through NVLink in rank order (bit-identical on every
 * rank) -> squared norm of the mean gradient -> end barrier; (2) the same clip + Adam kernel as dtqn_clip_adam with
 * grad_scale = 1 / world.  No host synchronisation, graph-capturable (the barrier epoch lives in device memory).
 * `bases[r]` = rank r's buffer as mapped in THIS process (own base at [rank]); n must be a multiple of 4.
 * A peer that never arrives makes the bounded waits expire (~2 s) and sets an error flag: dtqn_p2p_error(base) != 0. */
#define DTQN_P2P_MAX_RANKS 8
#define DTQN_P2P_HANDLE_BYTES 64
int dtqn_p2p_alloc(int64_t n_floats, void** base_out, float** grads_out, uint8_t* handle_out /* [64] */);
int dtqn_p2p_open(const uint8_t* handle /* [64] */, void** peer_base_out);
int dtqn_p2p_close(void* peer_base);
int dtqn_p2p_free(void* base);
int dtqn_p2p_error(const void* base);
int dtqn_allreduce_clip_adam(float* params, void* const* bases, int32_t rank, int32_t world, int64_t n,
                             float* grads_reduced /* [n] local output: sum over ranks */, float* exp_avg,
                             float* exp_avg_sq, float max_norm, float lr, float beta1, float beta2, float eps,
                             int64_t* step_counter, float* scratch, float* stats_out, int32_t* flags_out,
                             float* stats_ring, int32_t ring_len, void* stream);


/* ---- measurement hooks (bench.py roofline leg; no reference analogue) --------------------------------------------
 * When enabled, every launch of a tagged kernel is bracketed by CUDA events on its stream.  dtqn_profile_read
 * synchronises and returns the summed duration, launch count and algorithmic work (FLOPs or bytes) of one tag.
 * Tags: 0 linear-fwd GEMM, 1 attention-fwd, 2 env-step, 3 env-roll, 4 replay-gather, 5 dgrad, 6 wgrad, 7 attention-bwd,
 * 8 layernorm-bwd, 9 embed, 10 head, 11 td-loss, 12 clip+adam, 13 other, 14 tcgen05 Linear (work = algorithmic bytes),
 * 15 sequence-resident fused forward.  Process-global, not thread-safe. */
int dtqn_profile_enable(int32_t on);
int dtqn_profile_read(int32_t tag, double* total_ms, int64_t* launches, double* total_work);

#ifdef __cplusplus
}
#endif
#endif /* DTQN_B200_H */
