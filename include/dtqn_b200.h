/*
 * dtqn_b200.h -- C ABI of libdtqn_b200.so: the B200-native (sm_100a) DTQN training hot path.
 *
 * Drop-in boundary for the data-parallel hot path of kevslinger/DTQN (pure Python; it has no FFI of its own,
 * SURVEY.md section 8b), so every entry point below cites the reference *Python* interface it replaces.
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + explicit sizes inside POD structs, a cudaStream_t passed as void*;
 *   - return int: 0 = ok, < 0 = invalid argument (DTQN_E_*), > 0 = cudaError_t of a failed launch;
 *   - never allocates, never synchronises, never throws; the caller (PyTorch) owns every buffer and keeps it
 *     alive until the stream has passed the call;
 *   - no global mutable state: re-entrant across processes (one process per GPU).
 */
#ifndef DTQN_B200_H
#define DTQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTQN_ABI_VERSION 1

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
    int32_t _pad;
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
    uint8_t*  done_flag;        /* [n] scratch: episode ended in this step (consumed by the roll kernel) */
    int32_t*  block_counts;     /* [ceil(n/256)] scratch: episodes finished per CTA in this step */
    int64_t*  ep_stats;         /* [4] running sums over finished episodes: return, length, successes, episodes */
    int32_t*  ep_return;        /* [n] return of the running episode (rewards are integers in both envs) */
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
    int32_t _pad;
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

/* ---- acting history: utils/context.py:8-111 (obs window only; a_embed = 0 so the action window is unused) ------- */
typedef struct dtqn_context {
    int32_t context_len;
    int32_t obs_dim;
    int32_t trunc_obs;          /* 1 = reproduce the reference's int64 context (obs truncated toward 0, A-Q2) */
    float   obs_mask;
    float*   obs;               /* [n_envs, ctx, O] ring; token j of the window = ring[(t+1-n+j) % ctx] */
    int32_t* timestep;          /* [n_envs] Context.timestep */
} dtqn_context;

typedef struct dtqn_step_io {
    int32_t action_mode;        /* DTQN_ACT_* */
    float   epsilon;
    int32_t* actions;           /* [n] in (GIVEN) / out (RANDOM, EPS_GREEDY) */
    const float* q_last;        /* [n, A] greedy Q of the last context position (EPS_GREEDY) */
    float*   obs_out;           /* [n, O] observation returned by env.step (terminal obs when done); nullable */
    float*   reward_out;        /* [n] nullable */
    uint8_t* done_out;          /* [n] env done incl. TimeLimit; nullable */
    uint8_t* truncated_out;     /* [n] info["TimeLimit.truncated"]; nullable */
    uint8_t* success_out;       /* [n] info["is_success"]; nullable */
} dtqn_step_io;

int dtqn_version(void);

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

#ifdef __cplusplus
}
#endif
#endif /* DTQN_B200_H */
