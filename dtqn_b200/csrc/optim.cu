// clip_grad_norm_(params, max_norm, error_if_nonfinite=True) + torch.optim.Adam.step on the flat parameter buffer
// (dtqn/agents/dtqn.py:257-265, dtqn/agents/dqn.py:64).  Two launches: sum of squares -> clip + Adam; every CTA of the
// second kernel re-reduces the partial sums in a fixed order, so the result is deterministic and identical on all CTAs
// (and, after the gradient allreduce, on all ranks).
#include "common.cuh"
#include "prof.cuh"

#define OPT_THREADS 256
#define OPT_MAX_BLOCKS 512

namespace {

__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_kernel(const float* __restrict__ g, long long n, float scale, float* __restrict__ partial, long long* step) {
    pdl_sync();
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) {
        const float v = g[i] * scale;
        s = fmaf(v, v, s);
    }
    __shared__ float red[OPT_THREADS / 32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
        if (blockIdx.x == 0) *step += 1;                       // optimizer step index t (1-based), read by adam_kernel
    }
}

__global__ void __launch_bounds__(OPT_THREADS)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float scale, float max_norm, float lr, float beta1, float beta2, float eps,
                 long long* __restrict__ step, const float* __restrict__ partial, int n_partial,
                 float* __restrict__ stats, int* __restrict__ flags, float* __restrict__ ring, int ring_len) {
    pdl_sync();
    __shared__ float s_coef;
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < n_partial; ++i) t += partial[i];
        const float total = sqrtf(t);
        const bool bad = !isfinite(total);
        s_bad = bad;
        s_coef = fminf(1.0f, max_norm / (total + 1e-6f));       // clip_coef clamped to 1
        if (blockIdx.x == 0) {
            stats[7] = total;
            // error_if_nonfinite (agents/dtqn.py:257-261 raises before optimizer.step): no update, no statistics row, and the
            // step index is handed back (nobody reads it on this path: every CTA returns before the bias correction)
            if (bad) { flags[0] = 1; *step -= 1; }
            else if (ring) {
                float* row = ring + ((*step - 1) % ring_len) * 8;
                for (int k = 0; k < 7; ++k) row[k] = stats[k];
                row[7] = total;
            }
        }
    }
    __syncthreads();
    if (s_bad) return;                                          // error_if_nonfinite: parameters untouched
    const float coef = s_coef * scale;
    const double t = (double)(*reinterpret_cast<const volatile long long*>(step));
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    const float neg_step = (float)(-(double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) {
        const float gi = g[i] * coef;
        const float mi = fmaf(w1, gi - m[i], m[i]);             // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(w2 * gi, gi, v[i] * beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        m[i] = mi; v[i] = vi;
        p[i] = fmaf(neg_step, mi / denom, p[i]);                // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
}

}  // namespace

// second half of the update on its own (the P2P gradient exchange in p2p.cu produces `partial` itself)
int launch_clip_adam_only(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float grad_scale,
                          float max_norm, float lr, float beta1, float beta2, float eps, long long* step,
                          const float* partial, int n_partial, float* stats, int* flags, float* ring, int ring_len,
                          cudaStream_t st) {
    int blocks = dtqn_cdiv(n, OPT_THREADS * 4);
    if (blocks > OPT_MAX_BLOCKS) blocks = OPT_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    launch_k(clip_adam_kernel, blocks, OPT_THREADS, 0, st, params, grads, exp_avg, exp_avg_sq, n, grad_scale, max_norm, lr, beta1, beta2,
             eps, step, partial, n_partial, stats, flags, ring, ring_len);
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int dtqn_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                              float max_norm, float lr, float beta1, float beta2, float eps, int64_t* step_counter,
                              float* scratch, float* stats_out, int32_t* flags_out, float* stats_ring, int32_t ring_len,
                              void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || !step_counter || !scratch || !stats_out || !flags_out)
        return DTQN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = dtqn_cdiv(n, OPT_THREADS * 4);
    if (blocks > OPT_MAX_BLOCKS) blocks = OPT_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    prof_begin(PROF_ADAM, st);
    launch_k(sqnorm_kernel, blocks, OPT_THREADS, 0, st, (const float*)grads, (long long)n, grad_scale, scratch, (long long*)step_counter);
    DTQN_LAUNCH_CHECK();
    launch_k(clip_adam_kernel, blocks, OPT_THREADS, 0, st, params, (const float*)grads, exp_avg, exp_avg_sq, (long long)n, grad_scale,
             max_norm, lr, beta1, beta2, eps, (long long*)step_counter, (const float*)scratch, blocks, stats_out, flags_out,
             stats_ring, stats_ring ? ring_len : 1);
    prof_end(PROF_ADAM, st, 32.0 * (double)n);     // 28 B/param Adam + 4 B/param norm pass
    DTQN_LAUNCH_CHECK();
    return 0;
}
