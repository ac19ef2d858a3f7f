// Batched lockstep POMDP environments (CarFlag, Memory-5) + TimeLimit + fused replay / context append.
//
// Bit-exact restatement of envs/car_flag.py:76-133,145-159 and envs/memory_cards.py:70-116 (reference Python),
// one thread per environment instance, SoA state so every load/store is coalesced across the warp.
// CarFlag arithmetic is f64 with explicit round-to-nearest intrinsics (never contracted into FMA) in the
// reference's operation order; numpy PCG64 draws are reproduced by pcg64.cuh.
//
// Two kernels per lockstep step:
//   env_step_kernel : action pick -> env.step -> TimeLimit -> ReplayBuffer.store / Context.add_transition,
//                     close the episode where done (flush) and count finished episodes per CTA;
//   env_roll_kernel : for finished envs only: deterministic slot allocation (rank by env index = prefix over
//                     the CTA counts), env.reset(), ReplayBuffer.store_obs, Context.reset.
#include "common.cuh"
#include "pcg64.cuh"
#include "prof.cuh"

#define ENV_THREADS 256

namespace {

__device__ __forceinline__ unsigned nib_get(uint64_t v, int i) { return (unsigned)((v >> (4 * i)) & 0xFull); }
__device__ __forceinline__ uint64_t nib_set(uint64_t v, int i, unsigned x) {
    return (v & ~(0xFull << (4 * i))) | ((uint64_t)x << (4 * i));
}

template <int KIND> struct ObsDim { static constexpr int value = (KIND == DTQN_ENV_CARFLAG) ? 3 : 10; };

// ---- env.reset() -------------------------------------------------------------------------------------------------
// CarFlag (car_flag.py:145-159): heaven = +1 if integers(0,2,size=1)==0 else -1; p0 = uniform(-0.2, 0.2); v = 0.
// Memory  (memory_cards.py:70-80): cards = shuffle(repeat(1..5, 2)); obs = 0; cur = integers(10); reveal cur.
template <int KIND>
__device__ __forceinline__ void env_reset_one(const dtqn_env& e, int i, Pcg64& g, float* obs0) {
    if (KIND == DTQN_ENV_CARFLAG) {
        int8_t heaven = (g.bounded(2u) == 0u) ? (int8_t)1 : (int8_t)-1;
        double u = g.next_double();
        double p0 = __dadd_rn(-0.2, __dmul_rn(0.2 - (-0.2), u));      // low + (high - low) * u, two roundings
        e.heaven[i] = heaven;
        e.pos[i] = p0;
        e.vel[i] = 0.0;
        obs0[0] = (float)p0; obs0[1] = 0.f; obs0[2] = 0.f;
    } else {
        unsigned c[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) c[k] = (unsigned)(k / 2 + 1);
        for (int k = 9; k >= 1; --k) {                                 // Generator.shuffle: i = n-1 .. 1
            unsigned j = g.interval((unsigned)k);
            unsigned t = c[k]; c[k] = c[j]; c[j] = t;
        }
        uint64_t cards = 0;
#pragma unroll
        for (int k = 0; k < 10; ++k) cards |= (uint64_t)c[k] << (4 * k);
        int cur = (int)g.bounded(10u);
        uint64_t shown = nib_set(0ull, cur, nib_get(cards, cur));
        e.cards[i] = cards; e.shown[i] = shown; e.cur[i] = cur;
#pragma unroll
        for (int k = 0; k < 10; ++k) obs0[k] = (float)nib_get(shown, k);
    }
    e.elapsed[i] = 0;
    e.ep_return[i] = 0;
}

__device__ __forceinline__ void ctx_write(const dtqn_context& cx, int i, int row, const float* o, int O) {
    float* dst = cx.obs + ((size_t)i * cx.context_len + row) * O;
    for (int k = 0; k < O; ++k) dst[k] = cx.trunc_obs ? truncf(o[k]) : o[k];
}

// Context.reset (utils/context.py:36-54): fresh window with obs[0] = o; draws ctx bounded ints from the agent
// stream for the (unused, a_embed = 0) random action padding so the stream stays aligned with the reference.
__device__ __forceinline__ void ctx_reset(const dtqn_context& cx, int i, const float* o, int O, Pcg64& ag, unsigned A) {
    for (int k = 0; k < cx.context_len; ++k) {
        const unsigned a = ag.bounded(A);
        if (k == 0 && cx.action) cx.action[(size_t)i * cx.context_len] = (uint8_t)a;
    }
    cx.timestep[i] = 0;
    ctx_write(cx, i, 0, o, O);
}

template <int KIND>
__global__ void __launch_bounds__(ENV_THREADS)
env_reset_all_kernel(dtqn_env e, dtqn_replay rb, dtqn_context cx, int has_rb, int has_cx) {
    constexpr int O = ObsDim<KIND>::value;
    int i = blockIdx.x * ENV_THREADS + threadIdx.x;
    if (i == 0) {
        if (has_rb) { rb.counters[0] = e.n_envs; rb.counters[1] = e.n_envs; rb.counters[2] = 0; rb.counters[3] = 0; }
        e.ep_stats[0] = 0; e.ep_stats[1] = 0; e.ep_stats[2] = 0; e.ep_stats[3] = 0;
    }
    if (i >= e.n_envs) return;
    Pcg64 g; g.load(e.rng, e.rng_buf, e.n_envs, i);
    float o[O];
    env_reset_one<KIND>(e, i, g, o);
    g.store(e.rng, e.rng_buf, e.n_envs, i);
    e.done_flag[i] = 0;
    if (e.env_acc) { int* acc = e.env_acc + 4 * (size_t)i; acc[0] = acc[1] = acc[2] = acc[3] = 0; }
    if (has_rb) {                                                      // ReplayBuffer.store_obs (:88-92), slot = env index
        rb.env_slot[i] = i;
        rb.env_prev_len[i] = rb.episode_lengths[i];
        rb.episode_lengths[i] = 0;
        rb.slot_open[i] = 1;
        float* dst = rb.obss + (size_t)i * (rb.max_episode_steps + 1) * O;
#pragma unroll
        for (int k = 0; k < O; ++k) dst[k] = o[k];
    }
    if (has_cx) {
        Pcg64 ag; ag.load(e.arng, e.arng_buf, e.n_envs, i);
        ctx_reset(cx, i, o, O, ag, (unsigned)e.num_actions);
        ag.store(e.arng, e.arng_buf, e.n_envs, i);
    }
}

template <int KIND>
__global__ void __launch_bounds__(ENV_THREADS)
env_step_kernel(dtqn_env e, dtqn_replay rb, dtqn_context cx, dtqn_step_io io, int has_rb, int has_cx) {
    constexpr int O = ObsDim<KIND>::value;
    const int i = blockIdx.x * ENV_THREADS + threadIdx.x;
    if (has_rb && i == 0) rb.counters[0] = rb.counters[1];            // publish last step's allocations
    int done = 0;
    // run.evaluate plays a fixed number of episodes per env (run.py:214-233): an env that has finished its
    // stat_episodes_per_env-th episode is frozen -- no step, no RNG draw, no reset -- until the next dtqn_env_reset_all
    bool frozen = false;
    if (i < e.n_envs && e.env_acc && e.stat_episodes_per_env > 0) frozen = e.env_acc[4 * (size_t)i] >= e.stat_episodes_per_env;
    if (i < e.n_envs && frozen) e.done_flag[i] = 0;
    if (i < e.n_envs && !frozen) {
        // ---- action: run.py:394 (random), agents/dtqn.py:78-107 (eps-greedy), or supplied ----
        int a;
        const unsigned A = (unsigned)e.num_actions;
        if (io.action_mode == DTQN_ACT_GIVEN) {
            a = io.actions[i];
        } else {
            Pcg64 ag; ag.load(e.arng, e.arng_buf, e.n_envs, i);
            if (io.action_mode == DTQN_ACT_RANDOM) {
                a = (int)ag.bounded(A);
            } else {
                double u = ag.next_double();
                const double eps = io.epsilon_dev ? *io.epsilon_dev : io.epsilon;   // f64 vs f64 like agents/dtqn.py:78
                if (u < eps) {
                    a = (int)ag.bounded(A);
                } else {                                               // torch.argmax: first maximal index
                    const float* q = io.q_last + (size_t)i * A;
                    a = 0; float best = q[0];
                    for (unsigned k = 1; k < A; ++k) { float v = q[k]; if (v > best) { best = v; a = (int)k; } }
                }
            }
            ag.store(e.arng, e.arng_buf, e.n_envs, i);
            io.actions[i] = a;
        }
        // ---- env.step ----
        float o[O];
        int r = 0, success = 0;
        if (KIND == DTQN_ENV_CARFLAG) {
            double p = e.pos[i], v = e.vel[i];
            const double heaven = (double)e.heaven[i];
            v = __dadd_rn(v, __dmul_rn((double)(a - 1), 0.0015));      // car_flag.py:81,85
            if (v > 0.07) v = 0.07;                                    // :86-89
            if (v < -0.07) v = -0.07;
            p = __dadd_rn(p, v);                                       // :90
            if (p > 1.1) p = 1.1;                                      // :91-94
            if (p < -1.1) p = -1.1;
            if (p == -1.1 && v < 0.0) v = 0.0;                         // :95-96
            done = (p >= 1.0 || p <= -1.0);                            // :98-101 (heaven/hell are +-1)
            if (heaven > 0.0) { if (p >= 1.0) r = 1; if (p <= -1.0) r = -1; }   // :105-110
            else              { if (p <= -1.0) r = 1; if (p >= 1.0) r = -1; }   // :112-117
            double d = 0.0;
            if (p >= 0.5 - 0.2 && p <= 0.5 + 0.2) d = heaven;          // :119-129
            e.pos[i] = p; e.vel[i] = v;
            o[0] = (float)p; o[1] = (float)v; o[2] = (float)d;         // f64 -> f32 on the buffer write (replay_buffer.py:81)
            success = r > 0;                                           // :133
        } else {
            uint64_t cards = e.cards[i], shown = e.shown[i];
            int cur = e.cur[i];
            Pcg64 g; g.load(e.rng, e.rng_buf, e.n_envs, i);
            const unsigned ocur = nib_get(shown, cur);
            if (a == cur) { shown = nib_set(shown, cur, 0u); r = -1; }                   // memory_cards.py:89-91
            else if (nib_get(cards, a) == ocur) {                                        // :93-103
                shown = nib_set(shown, a, 6u); shown = nib_set(shown, cur, 6u); r = 0;
                if (shown == 0x6666666666ull) { done = 1; success = 1; }
            } else { shown = nib_set(shown, cur, 0u); r = -1; }                          // :104-106
            if (!done) {                                                                 // :108-114
                cur = (int)g.bounded(10u);
                while (nib_get(shown, cur) == 6u) cur = (int)g.bounded(10u);
                shown = nib_set(shown, cur, nib_get(cards, cur));
                g.store(e.rng, e.rng_buf, e.n_envs, i);
            }
            e.shown[i] = shown; e.cur[i] = cur;
#pragma unroll
            for (int k = 0; k < 10; ++k) o[k] = (float)nib_get(shown, k);
        }
        // ---- TimeLimit (gym 0.18): elapsed >= max -> truncated = !done; done = True ----
        const int t = e.elapsed[i];                                    // transitions already in this episode
        const int el = t + 1;
        int truncated = 0;
        if (el >= e.max_episode_steps) { truncated = !done; done = 1; }
        e.elapsed[i] = el;
        const int buffer_done = done && !truncated;                    // run.py:370-374
        const int ret = e.ep_return[i] + r;
        e.ep_return[i] = ret;
        if (io.obs_out) {
#pragma unroll
            for (int k = 0; k < O; ++k) io.obs_out[(size_t)i * O + k] = o[k];
        }
        if (io.reward_out) io.reward_out[i] = (float)r;
        if (io.done_out) io.done_out[i] = (uint8_t)done;
        if (io.truncated_out) io.truncated_out[i] = (uint8_t)truncated;
        if (io.success_out) io.success_out[i] = (uint8_t)success;
        // ---- agent.observe: ReplayBuffer.store (replay_buffer.py:71-86) ----
        if (has_rb) {
            const int s = rb.env_slot[i];
            if (s >= 0) {
                const int E = rb.max_episode_steps;
                float* orow = rb.obss + ((size_t)s * (E + 1) + (t + 1)) * O;
#pragma unroll
                for (int k = 0; k < O; ++k) orow[k] = o[k];
                rb.actions[(size_t)s * (E + 1) + t] = (uint8_t)a;
                rb.rewards[(size_t)s * E + t] = (float)r;
                rb.dones[(size_t)s * E + t] = (uint8_t)buffer_done;
                rb.episode_lengths[s] = el;                            // = context.timestep (agents/dtqn.py:160)
                if (done) {
                    // flush (:97-98).  The slot was not cleansed at episode start (cleanse_episode :100-135 would
                    // rewrite all E+1 rows); instead only the tail the previous occupant left beyond this
                    // episode's end is reset to the fill values, so the closed slot is byte-identical.
                    const int prev = rb.env_prev_len[i];
                    for (int tt = el; tt < prev; ++tt) {
                        float* row = rb.obss + ((size_t)s * (E + 1) + (tt + 1)) * O;
#pragma unroll
                        for (int k = 0; k < O; ++k) row[k] = rb.obs_mask;
                        rb.actions[(size_t)s * (E + 1) + tt] = 0;
                        rb.rewards[(size_t)s * E + tt] = 0.f;
                        rb.dones[(size_t)s * E + tt] = 1;
                    }
                    rb.slot_open[s] = 0;
                    atomicAdd((unsigned long long*)&rb.counters[2], 1ull);       // a stored episode completed (can_sample)
                }
            }
        }
        // ---- Context.add_transition (utils/context.py:56-80) ----
        if (has_cx && !done) {
            const int ts = cx.timestep[i] + 1;
            cx.timestep[i] = ts;
            ctx_write(cx, i, ts % cx.context_len, o, O);
            if (cx.action) cx.action[(size_t)i * cx.context_len + ts % cx.context_len] = (uint8_t)a;   // context.py:77
        }
        if (done) {
            if (e.env_acc) {
                int* acc = e.env_acc + 4 * (size_t)i;
                acc[0] += 1; acc[1] += ret; acc[2] += el; acc[3] += (success || ret > 0);
                // the last counted episode is not followed by env.reset() / context_reset (the next evaluation starts with one)
                if (e.stat_episodes_per_env > 0 && acc[0] >= e.stat_episodes_per_env) frozen = true;
            }
            atomicAdd((unsigned long long*)&e.ep_stats[0], (unsigned long long)(long long)ret);
            atomicAdd((unsigned long long*)&e.ep_stats[1], (unsigned long long)el);
            atomicAdd((unsigned long long*)&e.ep_stats[2], (unsigned long long)(success || ret > 0));  // run.py:232
            atomicAdd((unsigned long long*)&e.ep_stats[3], 1ull);
        }
        if (frozen) done = 0;                                          // nothing to roll
        // which of this env's episodes enter the replay: every record_every-th (episode index = episodes it has finished)
        int flag = done;
        if (done && has_rb && rb.record_every > 1 && e.env_acc && (e.env_acc[4 * (size_t)i] % rb.record_every) != 0) flag = 2;
        e.done_flag[i] = (uint8_t)flag;
        done = flag == 1;                                              // counted below: episodes that need a replay slot
    }
    const int cnt = __syncthreads_count(done);
    if (threadIdx.x == 0) e.block_counts[blockIdx.x] = cnt;
}

template <int KIND>
__global__ void __launch_bounds__(ENV_THREADS)
env_roll_kernel(dtqn_env e, dtqn_replay rb, dtqn_context cx, int has_rb, int has_cx) {
    constexpr int O = ObsDim<KIND>::value;
    __shared__ int s_red[ENV_THREADS / 32];
    __shared__ int s_warp_off[ENV_THREADS / 32];
    __shared__ int s_prefix, s_total;
    const int i = blockIdx.x * ENV_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rank base = number of finished envs with a smaller index in preceding CTAs (deterministic order)
    int part = 0, tot = 0;
    const int nb = gridDim.x;
    for (int b = threadIdx.x; b < nb; b += ENV_THREADS) {
        int c = e.block_counts[b];
        tot += c;
        if (b < (int)blockIdx.x) part += c;
    }
    for (int o = 16; o > 0; o >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
    __shared__ int s_tot[ENV_THREADS / 32];
    if (lane == 0) { s_red[warp] = part; s_tot[warp] = tot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int p = 0, t = 0;
        for (int w = 0; w < ENV_THREADS / 32; ++w) { p += s_red[w]; t += s_tot[w]; }
        s_prefix = p; s_total = t;
    }
    const int flag = (i < e.n_envs) ? (int)e.done_flag[i] : 0;
    const int done = flag == 1;                                        // finished AND its successor takes a replay slot
    const unsigned bal = __ballot_sync(0xffffffffu, done);
    const int any_roll = __syncthreads_or(flag != 0);
    if (s_total == 0 && !any_roll) return;                             // nothing finished in this CTA / anywhere
    if (lane == 0) s_warp_off[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < ENV_THREADS / 32; ++w) { int c = s_warp_off[w]; s_warp_off[w] = acc; acc += c; }
    }
    __syncthreads();
    if (has_rb && blockIdx.x == 0 && threadIdx.x == 0) {
        rb.counters[1] = rb.counters[0] + s_total;                     // episodes started (published by the next step)
    }
    if (!flag) return;
    const int rank = s_prefix + s_warp_off[warp] + __popc(bal & ((1u << lane) - 1u));
    Pcg64 g; g.load(e.rng, e.rng_buf, e.n_envs, i);
    float o[O];
    env_reset_one<KIND>(e, i, g, o);                                   // run.py:295-296 env.reset()
    g.store(e.rng, e.rng_buf, e.n_envs, i);
    if (has_rb && !done) rb.env_slot[i] = -1;                          // this episode is not stored (record_every)
    if (has_rb && done) {                                              // ReplayBuffer.store_obs at pos[0] % max_size
        const long long k = rb.counters[0] + rank;
        int s = (int)(k % rb.n_slots);
        if (rb.slot_open[s]) {                                         // ring wrapped onto a still-running episode
            s = -1;
            atomicAdd((unsigned long long*)&rb.counters[3], 1ull);
        } else {
            rb.env_prev_len[i] = rb.episode_lengths[s];
            rb.episode_lengths[s] = 0;
            rb.slot_open[s] = 1;
            float* dst = rb.obss + (size_t)s * (rb.max_episode_steps + 1) * O;
#pragma unroll
            for (int kk = 0; kk < O; ++kk) dst[kk] = o[kk];
        }
        rb.env_slot[i] = s;
    }
    if (has_cx) {
        Pcg64 ag; ag.load(e.arng, e.arng_buf, e.n_envs, i);
        ctx_reset(cx, i, o, O, ag, (unsigned)e.num_actions);
        ag.store(e.arng, e.arng_buf, e.n_envs, i);
    }
}

int check_env(const dtqn_env* e, const dtqn_replay* rb, const dtqn_context* cx) {
    if (!e || e->n_envs <= 0 || !e->rng || !e->rng_buf || !e->arng || !e->arng_buf || !e->elapsed || !e->done_flag ||
        !e->block_counts || !e->ep_stats || !e->ep_return) return DTQN_E_ARG;
    if (e->kind == DTQN_ENV_CARFLAG) {
        if (!e->pos || !e->vel || !e->heaven || e->obs_dim != 3 || e->num_actions != 3) return DTQN_E_ARG;
    } else if (e->kind == DTQN_ENV_MEMORY) {
        if (!e->cards || !e->shown || !e->cur || e->obs_dim != 10 || e->num_actions != 10) return DTQN_E_ARG;
    } else return DTQN_E_UNSUPPORTED;
    if (e->max_episode_steps <= 0) return DTQN_E_ARG;
    if (rb) {
        if (!rb->obss || !rb->actions || !rb->rewards || !rb->dones || !rb->episode_lengths || !rb->slot_open ||
            !rb->counters || !rb->env_slot || !rb->env_prev_len) return DTQN_E_ARG;
        if (rb->obs_dim != e->obs_dim || rb->max_episode_steps < e->max_episode_steps || rb->n_slots < e->n_envs)
            return DTQN_E_ARG;
    }
    if (cx) {
        if (!cx->obs || !cx->timestep || cx->obs_dim != e->obs_dim || cx->context_len <= 0) return DTQN_E_ARG;
    }
    return 0;
}

// LinearAnneal.anneal (utils/epsilon_anneal.py:33-34) kept on the device so a replayed CUDA graph needs no per-step host
// write: emit the current value for this step, then val <- max(min, val - (val - min) / duration), in double like the host.
__global__ void eps_anneal_kernel(double* state, double* eps_out) {
    const double val = state[0], lo = state[1], dur = state[2];
    *eps_out = val;
    state[0] = fmax(lo, __dsub_rn(val, __ddiv_rn(__dsub_rn(val, lo), dur)));
}

}  // namespace

extern "C" int dtqn_version(void) { return DTQN_ABI_VERSION; }

extern "C" int dtqn_eps_anneal(double* state, double* eps_out, void* stream) {
    if (!state || !eps_out) return DTQN_E_ARG;
    eps_anneal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, eps_out);
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int dtqn_env_reset_all(const dtqn_env* env, const dtqn_replay* rb, const dtqn_context* cx, void* stream) {
    int rc = check_env(env, rb, cx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    dtqn_replay rbv = rb ? *rb : dtqn_replay{};
    dtqn_context cxv = cx ? *cx : dtqn_context{};
    const int grid = dtqn_cdiv(env->n_envs, ENV_THREADS);
    if (env->kind == DTQN_ENV_CARFLAG)
        env_reset_all_kernel<DTQN_ENV_CARFLAG><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, rb != nullptr, cx != nullptr);
    else
        env_reset_all_kernel<DTQN_ENV_MEMORY><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, rb != nullptr, cx != nullptr);
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int dtqn_env_step(const dtqn_env* env, const dtqn_replay* rb, const dtqn_context* cx, const dtqn_step_io* io,
                             void* stream) {
    int rc = check_env(env, rb, cx);
    if (rc) return rc;
    if (!io || !io->actions) return DTQN_E_ARG;
    if (io->action_mode < DTQN_ACT_GIVEN || io->action_mode > DTQN_ACT_EPS_GREEDY) return DTQN_E_ARG;
    if (io->action_mode == DTQN_ACT_EPS_GREEDY && !io->q_last) return DTQN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    dtqn_replay rbv = rb ? *rb : dtqn_replay{};
    dtqn_context cxv = cx ? *cx : dtqn_context{};
    const int grid = dtqn_cdiv(env->n_envs, ENV_THREADS);
    // algorithmic bytes per env-step incl. the fused replay append: 55 B CarFlag, 140 B Memory
    const double step_bytes = (env->kind == DTQN_ENV_CARFLAG ? 55.0 : 140.0) * env->n_envs;
    prof_begin(PROF_ENV_STEP, st);
    if (env->kind == DTQN_ENV_CARFLAG)
        env_step_kernel<DTQN_ENV_CARFLAG><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, *io, rb != nullptr, cx != nullptr);
    else
        env_step_kernel<DTQN_ENV_MEMORY><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, *io, rb != nullptr, cx != nullptr);
    prof_end(PROF_ENV_STEP, st, step_bytes);
    DTQN_LAUNCH_CHECK();
    prof_begin(PROF_ENV_ROLL, st);
    if (env->kind == DTQN_ENV_CARFLAG)
        env_roll_kernel<DTQN_ENV_CARFLAG><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, rb != nullptr, cx != nullptr);
    else
        env_roll_kernel<DTQN_ENV_MEMORY><<<grid, ENV_THREADS, 0, st>>>(*env, rbv, cxv, rb != nullptr, cx != nullptr);
    prof_end(PROF_ENV_ROLL, st, 0.0);
    DTQN_LAUNCH_CHECK();
    return 0;
}
