// Device-resident replay buffer: index sampling and the history-window gather.
// Restates dtqn/buffers/replay_buffer.py:137-168 (ReplayBuffer.sample).  The store / flush side is fused into the
// env step kernels (env.cu).
#include "common.cuh"
#include "prof.cuh"

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// unbiased integer in [0, n) from a counter-based stream (Lemire multiply-shift with rejection)
__device__ __forceinline__ uint32_t bounded_ctr(uint64_t key, uint32_t& ctr, uint32_t n) {
    uint64_t m = (uint64_t)(uint32_t)splitmix64(key + (uint64_t)(ctr++) * 0xD1342543DE82EF95ull) * n;
    uint32_t l = (uint32_t)m;
    if (l < n) {
        uint32_t t = (0u - n) % n;
        while (l < t) {
            m = (uint64_t)(uint32_t)splitmix64(key + (uint64_t)(ctr++) * 0xD1342543DE82EF95ull) * n;
            l = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

// One CTA; thread b draws sample b (replay_buffer.py:141-156): episode ~ U(completed episodes), i.e. slots below
// min(started, S) that are not open (":141-145 exclude the current episode"; here n_envs open slots);
// start ~ U{0 .. max(0, eplen - ctx)}.
__global__ void __launch_bounds__(256)
sample_indices_kernel(dtqn_replay rb, int batch, uint64_t seed, uint64_t draw, uint64_t* draw_counter,
                      int32_t* episodes, int32_t* starts) {
    pdl_sync();
    const uint64_t d = draw + (draw_counter ? *draw_counter : 0ull);
    long long started = rb.counters[1];
    const uint32_t range = (uint32_t)(started < rb.n_slots ? started : rb.n_slots);
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const uint64_t key = splitmix64(seed ^ splitmix64(d * 0x100000001B3ull + (uint64_t)b));
        uint32_t ctr = 0;
        int e = 0;
        for (int tries = 0; tries < 4096; ++tries) {
            e = (int)bounded_ctr(key, ctr, range);
            if (!rb.slot_open[e] && rb.episode_lengths[e] > 0) break;
        }
        int len = rb.episode_lengths[e];
        int span = len - rb.context_len;
        if (span < 0) span = 0;
        episodes[b] = e;
        starts[b] = (int)bounded_ctr(key, ctr, (uint32_t)span + 1u);
    }
    __syncthreads();
    if (draw_counter && threadIdx.x == 0) *draw_counter += 1ull;
}

// One CTA per sampled window.  The (L+1) observation rows of a window are contiguous in the episode-major layout
// ((L+1)*O floats), so the copy is a coalesced stream; obss / next_obss of the reference are rows [0,L) / [1,L+1)
// of the same window and are not materialised twice.
__global__ void __launch_bounds__(128)
replay_gather_kernel(dtqn_replay rb, const int32_t* __restrict__ episodes, const int32_t* __restrict__ starts,
                     float* __restrict__ obs_win, uint8_t* __restrict__ act_win, float* __restrict__ rew,
                     uint8_t* __restrict__ done, int32_t* __restrict__ eplen) {
    pdl_sync();
    const int b = blockIdx.x;
    const int e = episodes[b], s0 = starts[b];
    const int L = rb.context_len, O = rb.obs_dim, E = rb.max_episode_steps;
    const float* src_o = rb.obss + ((size_t)e * (E + 1) + s0) * O;
    float* dst_o = obs_win + (size_t)b * (L + 1) * O;
    for (int k = threadIdx.x; k < (L + 1) * O; k += blockDim.x) dst_o[k] = __ldg(src_o + k);
    const uint8_t* src_a = rb.actions + (size_t)e * (E + 1) + s0;
    for (int k = threadIdx.x; k < L + 1; k += blockDim.x) act_win[(size_t)b * (L + 1) + k] = __ldg(src_a + k);
    const float* src_r = rb.rewards + (size_t)e * E + s0;
    const uint8_t* src_d = rb.dones + (size_t)e * E + s0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        rew[(size_t)b * L + k] = __ldg(src_r + k);
        done[(size_t)b * L + k] = __ldg(src_d + k);
    }
    if (threadIdx.x == 0 && eplen) {
        int l = rb.episode_lengths[e];
        eplen[b] = l < 0 ? 0 : (l > L ? L : l);                        // np.clip(eplen, 0, ctx) (:167)
    }
}

}  // namespace

extern "C" int dtqn_replay_sample_indices(const dtqn_replay* rb, int32_t batch, uint64_t seed, uint64_t draw,
                                          uint64_t* draw_counter, int32_t* episodes_out, int32_t* starts_out,
                                          void* stream) {
    if (!rb || batch <= 0 || !episodes_out || !starts_out || !rb->counters || !rb->slot_open || !rb->episode_lengths)
        return DTQN_E_ARG;
    prof_begin(PROF_OTHER, (cudaStream_t)stream);
    launch_k(sample_indices_kernel, 1, 256, 0, (cudaStream_t)stream, *rb, batch, seed, draw, draw_counter, episodes_out, starts_out);
    prof_end(PROF_OTHER, (cudaStream_t)stream, 0.0);
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int dtqn_replay_gather(const dtqn_replay* rb, int32_t batch, const int32_t* episodes, const int32_t* starts,
                                  float* obs_win, uint8_t* act_win, float* rew, uint8_t* done, int32_t* eplen,
                                  void* stream) {
    if (!rb || batch <= 0 || !episodes || !starts || !obs_win || !act_win || !rew || !done) return DTQN_E_ARG;
    if (rb->context_len > rb->max_episode_steps) return DTQN_E_ARG;   // the reference's fancy index would go out of range
    // algorithmic bytes per window: read (L+1)*O*4 + (L+1) + 4L + L + 1, write the same minus eplen
    const double L = rb->context_len, O = rb->obs_dim;
    const double win_bytes = 2.0 * ((L + 1) * O * 4 + (L + 1) + 4 * L + L) + 1.0;
    prof_begin(PROF_GATHER, (cudaStream_t)stream);
    launch_k(replay_gather_kernel, batch, 128, 0, (cudaStream_t)stream, *rb, episodes, starts, obs_win, act_win, rew, done, eplen);
    prof_end(PROF_GATHER, (cudaStream_t)stream, win_bytes * batch);
    DTQN_LAUNCH_CHECK();
    return 0;
}
