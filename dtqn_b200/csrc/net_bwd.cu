// Double-DQN sequence TD loss and the DTQN backward pass (dtqn/agents/dtqn.py:215-256: gather Q(s,a), a* = argmax
// policy(s'), Q_tgt(s',a*), Bellman target, mse_loss over the last `history` positions, loss.backward()).
// All gradients are derived by hand for the forward in net_fwd.cu; parity is against torch autograd on the oracle.
#include "net.cuh"
#include "gemm_simt.cuh"
#include "prof.cuh"

namespace {

// ---- deterministic cross-CTA sums -------------------------------------------------------------------------------------------
// Every parameter gradient is a sum over tokens that several CTAs ("chunks") share.  Instead of fp32 atomics (run-to-run
// order) each CTA stores its partial sums, takes a ticket, and the CTA that arrives LAST adds the partials of all chunks
// in chunk order and writes the gradient: the value never depends on scheduling, so a step is bit-reproducible (eager ==
// CUDA-graph replay == another run), like the reference's single-threaded-order autograd is.
// Call with all threads of the CTA after the partial stores; returns true in every thread of the last CTA.
__device__ __forceinline__ bool last_chunk_arrived(unsigned* ticket, unsigned n_chunks) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last = atomicAdd(ticket, 1u) == n_chunks - 1;
        if (last) *ticket = 0u;                       // self-resetting: the next launch (stream-ordered) starts from 0
        s_last = last;
    }
    __syncthreads();
    const bool last = s_last;
    if (last) __threadfence();
    return last;
}

// sum_{b < n} p[b * stride] in index order; the loads of a batch of 16 are independent (in flight together), the adds ordered
__device__ __forceinline__ float ordered_sum(const float* __restrict__ p, size_t stride, unsigned n) {
    float sum = 0.f;
    for (unsigned b0 = 0; b0 < n; b0 += 16) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (b0 + k < n) ? __ldcg(p + (size_t)(b0 + k) * stride) : 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += v[k];
    }
    return sum;
}

// ---- TD loss: emits dLoss/dQ directly + the logged statistics (agents/dtqn.py:245-253) --------------------------------
// q_all [3, B, L, A]: 0 = policy(obs), 1 = policy(next_obs), 2 = target(next_obs).
__global__ void __launch_bounds__(256)
td_loss_kernel(const float* __restrict__ q_all, const uint8_t* __restrict__ act_win, const float* __restrict__ rew,
               const uint8_t* __restrict__ done, int B, int L, int A, int history, float gamma,
               float* __restrict__ dq, float* __restrict__ partial, unsigned* __restrict__ ticket,
               float* __restrict__ stats) {
    pdl_sync();
    const long long T = (long long)B * L;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float se = 0.f, qs = 0.f, ys = 0.f, qmx = -INFINITY, qmn = INFINITY, ymx = -INFINITY, ymn = INFINITY;
    if (t < T) {
        const int b = (int)(t / L), l = (int)(t % L);
        const int a = act_win[(long long)b * (L + 1) + l];
        const float* q0 = q_all + t * A;
        const float* q1 = q_all + (T + t) * A;
        const float* q2 = q_all + (2 * T + t) * A;
        int astar = 0; float best = q1[0];
        for (int k = 1; k < A; ++k) { const float v = q1[k]; if (v > best) { best = v; astar = k; } }   // torch.argmax
        const float qsel = q0[a];
        const float y = rew[t] + (1.f - (float)done[t]) * (q2[astar] * gamma);                          // :236-238
        const bool in_hist = l >= L - history;                                                          // :240-241
        const float diff = qsel - y;
        for (int k = 0; k < A; ++k) dq[t * A + k] = 0.f;
        if (in_hist) {
            dq[t * A + a] = 2.f * diff / (float)((long long)B * history);                              // d mse / dq
            se = diff * diff; qs = qsel; ys = y; qmx = qmn = qsel; ymx = ymn = y;
        }
    }
    // block reduction (fixed order) -> per-block partials -> last block finalises in block order (deterministic)
    __shared__ float red[7][8];
    float v[7] = {se, qs, ys, qmx, qmn, ymx, ymn};
    v[0] = warp_sum(v[0]); v[1] = warp_sum(v[1]); v[2] = warp_sum(v[2]);
    v[3] = warp_max(v[3]); v[4] = warp_min(v[4]); v[5] = warp_max(v[5]); v[6] = warp_min(v[6]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int k = 0; k < 7; ++k) red[k][warp] = v[k];
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        float r[7] = {0.f, 0.f, 0.f, -INFINITY, INFINITY, -INFINITY, INFINITY};
        for (int w = 0; w < 8; ++w) {
            r[0] += red[0][w]; r[1] += red[1][w]; r[2] += red[2][w];
            r[3] = fmaxf(r[3], red[3][w]); r[4] = fminf(r[4], red[4][w]);
            r[5] = fmaxf(r[5], red[5][w]); r[6] = fminf(r[6], red[6][w]);
        }
        for (int k = 0; k < 7; ++k) partial[blockIdx.x * 8 + k] = r[k];
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float r[7] = {0.f, 0.f, 0.f, -INFINITY, INFINITY, -INFINITY, INFINITY};
        for (unsigned bb = 0; bb < gridDim.x; ++bb) {
            const volatile float* pp = partial + bb * 8;
            r[0] += pp[0]; r[1] += pp[1]; r[2] += pp[2];
            r[3] = fmaxf(r[3], pp[3]); r[4] = fminf(r[4], pp[4]); r[5] = fmaxf(r[5], pp[5]); r[6] = fminf(r[6], pp[6]);
        }
        const float n = (float)((long long)B * history);
        stats[0] = r[0] / n;              // F.mse_loss (mean)
        stats[1] = r[3]; stats[2] = r[1] / n; stats[3] = r[4];
        stats[4] = r[5]; stats[5] = r[2] / n; stats[6] = r[6];
        *ticket = 0u;
    }
}

// ---- head (ffn.2) backward: N = A is tiny -> CUDA cores --------------------------------------------------------------------
// d_hh = (dq W2) * [hh > 0];  dW2 = dq^T hh;  db2 = colsum(dq).   tok <= HB_TOK tokens per CTA (dtqn_set_head_bwd_tokens);
// per-CTA partials [A*d + A] summed in CTA order by the last CTA.
#define HB_TOK 64
#define HB_TOK_MIN 16
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ hh, const float* __restrict__ W2, int T, int d,
                int A, int tok, float* __restrict__ d_hh, float* __restrict__ gW2, float* __restrict__ gb2,
                float* __restrict__ part, unsigned* __restrict__ ticket) {
    pdl_sync();
    __shared__ float sdq[HB_TOK][33];
    const int t0 = blockIdx.x * tok;
    const int nt = min(tok, T - t0);
    for (int e = threadIdx.x; e < tok * A; e += blockDim.x) {
        const int r = e / A, a = e % A;
        sdq[r][a] = r < nt ? dq[(long long)(t0 + r) * A + a] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nt * d; e += blockDim.x) {
        const int r = e / d, c = e % d;
        float s = 0.f;
        for (int a = 0; a < A; ++a) s = fmaf(sdq[r][a], __ldg(W2 + a * d + c), s);
        const long long gi = (long long)(t0 + r) * d + c;
        d_hh[gi] = hh[gi] > 0.f ? s : 0.f;
    }
    for (int e = threadIdx.x; e < A * d; e += blockDim.x) {
        const int a = e / d, c = e % d;
        float s = 0.f;
#pragma unroll 16
        for (int r = 0; r < nt; ++r) s = fmaf(sdq[r][a], hh[(long long)(t0 + r) * d + c], s);
        part[(size_t)blockIdx.x * (A * d + A) + e] = s;
    }
    if (threadIdx.x < A) {
        float s = 0.f;
        for (int r = 0; r < nt; ++r) s += sdq[r][threadIdx.x];
        part[(size_t)blockIdx.x * (A * d + A) + A * d + threadIdx.x] = s;
    }
    if (!last_chunk_arrived(ticket, gridDim.x)) return;
    for (int e = threadIdx.x; e < A * d + A; e += blockDim.x) {
        const float s = ordered_sum(part + e, (size_t)(A * d + A), gridDim.x);
        if (e < A * d) gW2[e] = s; else gb2[e - A * d] = s;
    }
}

// ---- dgrad:  dX[T, Kf] = dY[T, Nf] W[Nf, Kf]  (+ epilogue) -----------------------------------------------------------------
enum { DG_NONE = 0, DG_MASK = 1, DG_ADD = 2 };

// BM rows per CTA (64 / 32 / 16; dtqn_set_dgrad_rows): with 1 600 tokens the 64-row tile fills 25 of 148 SMs.
// LN = true (needs Kf == BN, one CTA per row block): the tile is the dy of the LayerNorm the forward pass applied to this
// Linear's input (y = LN(x_in + relu(a))), and the epilogue is that LayerNorm's backward -- the ln_bwd_kernel maths on the
// tile staged in shared memory -- so dX is never written:  du -> ln.du (may alias aux: rows are CTA-private and the aux reads
// precede the barrier),  da = du * [r > 0] -> ln.da,  dgamma / dbeta per-CTA partials summed in CTA order by the last CTA.
struct LnBwdArgs {
    const float *xin, *r, *st, *gamma;
    float *du, *da, *ggamma, *gbeta, *part;
    unsigned* ticket;
};

template <int BM, int BN, int EPI, bool LN>
__global__ void __launch_bounds__(GEMM_THREADS)
dgrad_kernel(const float* __restrict__ dY, const float* __restrict__ W, const float* __restrict__ aux,
             float* __restrict__ dX, int T, int Nf, int Kf, LnBwdArgs ln) {
    pdl_sync();
    __shared__ __align__(16) GemmSmem<BN> sm;
    constexpr int RM = BM / 16, TN = BN / 16;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    float acc[RM][TN];
    gemm_tile<BM, BN, true>(dY, Nf, T, W, Kf, Nf, m0, n0, acc, sm);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    static_assert(!LN || BM * (BN + 1) <= (int)(sizeof(GemmSmem<BN>) / sizeof(float)), "dy tile must fit the GEMM slabs");
    float* sY = reinterpret_cast<float*>(&sm);                 // LN: [BM][BN + 1] dy tile (the k loop ended with a barrier)
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int lr = ty * RM + i, r = m0 + lr;
        if (r >= T) continue;
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
            const size_t off = (size_t)r * Kf + n0 + j4 * 64 + tx * 4;
            float v[4] = {acc[i][j4 * 4], acc[i][j4 * 4 + 1], acc[i][j4 * 4 + 2], acc[i][j4 * 4 + 3]};
            if (EPI != DG_NONE) {
                const float4 x = *reinterpret_cast<const float4*>(aux + off);
                const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = (EPI == DG_MASK) ? (xv[q] > 0.f ? v[q] : 0.f) : v[q] + xv[q];
            }
            if constexpr (LN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) sY[lr * (BN + 1) + j4 * 64 + tx * 4 + q] = v[q];
            } else {
                *reinterpret_cast<float4*>(dX + off) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if constexpr (LN) {
        constexpr int D = BN, PER = D / 32, ROWS = BM / 8;     // rows per warp
        __shared__ float sg[8][D], sb[8][D];
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float gam[PER], ag[PER], ab[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) { gam[k] = ln.gamma[lane + 32 * k]; ag[k] = 0.f; ab[k] = 0.f; }
        for (int rr = 0; rr < ROWS; ++rr) {
            const int lr = warp * ROWS + rr, t = m0 + lr;
            if (t >= T) break;
            const float mean = ln.st[2 * t], rstd = ln.st[2 * t + 1];
            float xh[PER], g[PER], rv[PER];
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const size_t o = (size_t)t * D + lane + 32 * k;
                const float dyv = sY[lr * (BN + 1) + lane + 32 * k];
                rv[k] = ln.r[o];
                xh[k] = (ln.xin[o] + rv[k] - mean) * rstd;
                g[k] = dyv * gam[k];
                m1 += g[k]; m2 = fmaf(g[k], xh[k], m2);
                ag[k] = fmaf(dyv, xh[k], ag[k]); ab[k] += dyv;
            }
            m1 = warp_sum(m1) * (1.f / D); m2 = warp_sum(m2) * (1.f / D);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const size_t o = (size_t)t * D + lane + 32 * k;
                const float v = rstd * (g[k] - m1 - xh[k] * m2);
                ln.du[o] = v;
                ln.da[o] = rv[k] > 0.f ? v : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) { sg[warp][lane + 32 * k] = ag[k]; sb[warp][lane + 32 * k] = ab[k]; }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            float s1 = 0.f, s2 = 0.f;
            for (int w = 0; w < 8; ++w) { s1 += sg[w][c]; s2 += sb[w][c]; }
            ln.part[(size_t)blockIdx.x * 2 * D + c] = s1; ln.part[(size_t)blockIdx.x * 2 * D + D + c] = s2;
        }
        if (!last_chunk_arrived(ln.ticket, gridDim.x)) return;
        for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
            const float sum = ordered_sum(ln.part + c, (size_t)(2 * D), gridDim.x);
            if (c < D) ln.ggamma[c] = sum; else ln.gbeta[c - D] = sum;
        }
    }
}

// ---- wgrad:  gW[Nf, Kf] = dY[T, Nf]^T X[T, Kf];  gb[Nf] = colsum(dY) ----------------------------------------------------------
// 64 x 64 tile of gW per CTA, tokens split over gridDim.z chunks (split-K); chunk partials go to pW / pb (same indexing as
// gW / gb, `pstride` floats between chunks); wgrad_reduce_kernel adds them in chunk order at the end of the backward pass.
#define WG_CHUNK 64
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, int T, int Nf, int Kf,
             float* __restrict__ gb, float* __restrict__ pW, float* __restrict__ pb, long long pstride, int chunk) {
    __shared__ __align__(16) float As[16][64 + 4];   // dY slab: 16 tokens x 64 n   (float4 accesses: keep 16-byte aligned
    __shared__ __align__(16) float Bs[16][64 + 4];   // X  slab: 16 tokens x 64 k    whatever else lives in static shared memory)
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int tb = blockIdx.z * chunk, te = min(T, tb + chunk);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lr = tid >> 4, lc = (tid & 15) * 4;            // slab load: row lr (token), 4 columns at lc
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    for (int t = tb; t < te; t += 16) {
        const bool ok = (t + lr) < te;
        const float4 av = ok ? *reinterpret_cast<const float4*>(dY + (size_t)(t + lr) * Nf + n0 + lc) : make_float4(0, 0, 0, 0);
        const float4 bv = ok ? *reinterpret_cast<const float4*>(X + (size_t)(t + lr) * Kf + k0 + lc) : make_float4(0, 0, 0, 0);
        *reinterpret_cast<float4*>(&As[lr][lc]) = av;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = bv;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (gb && blockIdx.y == 0 && tid < 64) {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) bsum += As[kk][tid];
        }
        __syncthreads();
    }
    const size_t poff = (size_t)blockIdx.z * pstride;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(pW + poff + (size_t)(n0 + ty * 4 + i) * Kf + k0 + tx * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (gb && blockIdx.y == 0 && tid < 64) pb[poff + n0 + tid] = bsum;
}

// Chunk-ordered sum of the split-K partials of every Linear weight / bias: grads[e] = sum_z partial[z][e] for e in the
// contiguous [in_w .. f2_b] block of each layer and the [h1_w .. h1_b] block of the head (same offsets in the partial rows as
// in the flat gradient).  One launch for the whole network, one float4 per thread, the loads of 4 chunks in flight together.
struct WgradSegs { long long begin[DTQN_MAX_LAYERS + 1], len4[DTQN_MAX_LAYERS + 1]; int n; long long total4; };
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, long long pstride, int n_chunks, WgradSegs segs, float* __restrict__ grads) {
    pdl_sync();
    long long q = (long long)blockIdx.x * 256 + threadIdx.x;
    if (q >= segs.total4) return;
    int sgi = 0;
    while (q >= segs.len4[sgi]) { q -= segs.len4[sgi]; ++sgi; }
    const long long e = segs.begin[sgi] + 4 * q;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z0 = 0; z0 < n_chunks; z0 += 4) {
        float4 v[4];
#pragma unroll
        for (int zz = 0; zz < 4; ++zz)
            v[zz] = (z0 + zz < n_chunks) ? __ldcg(reinterpret_cast<const float4*>(partial + (size_t)(z0 + zz) * pstride + e))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int zz = 0; zz < 4; ++zz) { sum.x += v[zz].x; sum.y += v[zz].y; sum.z += v[zz].z; sum.w += v[zz].w; }
    }
    *reinterpret_cast<float4*>(grads + e) = sum;
}

// ---- LayerNorm backward fused with the ReLU / residual split (transformer.py:72-73,76-77) -----------------------------------
//   y = LN(u), u = x_in + r, r = relu(a).   Given dy: du (-> residual path) and da = du * [r > 0]; dgamma, dbeta.
template <int D>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xin, const float* __restrict__ r,
              const float* __restrict__ st, const float* __restrict__ gamma, int T, float* __restrict__ du,
              float* __restrict__ da, float* __restrict__ ggamma, float* __restrict__ gbeta, float* __restrict__ part,
              unsigned* __restrict__ ticket) {
    pdl_sync();
    constexpr int PER = D / 32, ROWS = 4;                     // rows per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gam[PER], ag[PER], ab[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { gam[k] = gamma[lane + 32 * k]; ag[k] = 0.f; ab[k] = 0.f; }
    const int row0 = (blockIdx.x * 8 + warp) * ROWS;
    for (int rr = 0; rr < ROWS; ++rr) {
        const int t = row0 + rr;
        if (t >= T) break;
        const float mean = st[2 * t], rstd = st[2 * t + 1];
        float xh[PER], g[PER], rv[PER], dyv[PER];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const size_t o = (size_t)t * D + lane + 32 * k;
            rv[k] = r[o]; dyv[k] = dy[o];
            xh[k] = (xin[o] + rv[k] - mean) * rstd;
            g[k] = dyv[k] * gam[k];
            m1 += g[k]; m2 = fmaf(g[k], xh[k], m2);
            ag[k] = fmaf(dyv[k], xh[k], ag[k]); ab[k] += dyv[k];
        }
        m1 = warp_sum(m1) * (1.f / D); m2 = warp_sum(m2) * (1.f / D);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const size_t o = (size_t)t * D + lane + 32 * k;
            const float v = rstd * (g[k] - m1 - xh[k] * m2);
            du[o] = v;
            da[o] = rv[k] > 0.f ? v : 0.f;
        }
    }
    __shared__ float sg[8][D], sb[8][D];
#pragma unroll
    for (int k = 0; k < PER; ++k) { sg[warp][lane + 32 * k] = ag[k]; sb[warp][lane + 32 * k] = ab[k]; }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < 8; ++w) { s1 += sg[w][c]; s2 += sb[w][c]; }
        part[(size_t)blockIdx.x * 2 * D + c] = s1; part[(size_t)blockIdx.x * 2 * D + D + c] = s2;
    }
    if (!last_chunk_arrived(ticket, gridDim.x)) return;
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
        const float sum = ordered_sum(part + c, (size_t)(2 * D), gridDim.x);
        if (c < D) ggamma[c] = sum; else gbeta[c - D] = sum;
    }
}

// ---- attention backward, one CTA per (sequence, head); softmax probabilities are recomputed ----------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ d_o,
                float* __restrict__ d_qkv, int L, int d, float scale) {
    pdl_sync();
    __shared__ float Qs[128][HD + 1], Ks[128][HD + 1], Vs[128][HD + 1], Gs[128][HD + 1];
    __shared__ float sm_m[128], sm_il[128], sm_D[128];
    const int h = blockIdx.x;
    const size_t t0 = (size_t)blockIdx.y * L;
    const int tid = threadIdx.x;
    for (int e = tid; e < L * (HD / 4); e += blockDim.x) {
        const int r = e / (HD / 4), c4 = (e % (HD / 4)) * 4;
        const float* base = qkv + (t0 + r) * (size_t)(3 * d) + h * HD + c4;
        const float4 qv = *reinterpret_cast<const float4*>(base);
        const float4 kv = *reinterpret_cast<const float4*>(base + d);
        const float4 vv = *reinterpret_cast<const float4*>(base + 2 * d);
        const float4 gv = *reinterpret_cast<const float4*>(d_o + (t0 + r) * (size_t)d + h * HD + c4);
        Qs[r][c4] = qv.x * scale; Qs[r][c4 + 1] = qv.y * scale; Qs[r][c4 + 2] = qv.z * scale; Qs[r][c4 + 3] = qv.w * scale;
        Ks[r][c4] = kv.x; Ks[r][c4 + 1] = kv.y; Ks[r][c4 + 2] = kv.z; Ks[r][c4 + 3] = kv.w;
        Vs[r][c4] = vv.x; Vs[r][c4 + 1] = vv.y; Vs[r][c4 + 2] = vv.z; Vs[r][c4 + 3] = vv.w;
        Gs[r][c4] = gv.x; Gs[r][c4 + 1] = gv.y; Gs[r][c4 + 2] = gv.z; Gs[r][c4 + 3] = gv.w;
    }
    __syncthreads();
    const int j = tid;
    // phase A (thread = query row j): softmax statistics, D_j = dO_j . O_j, dQ_j
    if (j < L) {
        float q[HD], g[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { q[c] = Qs[j][c]; g[c] = Gs[j][c]; }
        float m = -INFINITY;
        for (int i = 0; i <= j; ++i) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) s = fmaf(q[c], Ks[i][c], s);
            m = fmaxf(m, s);
        }
        float l = 0.f;
        for (int i = 0; i <= j; ++i) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) s = fmaf(q[c], Ks[i][c], s);
            l += expf(s - m);
        }
        const float il = 1.f / l;
        float D = 0.f;
        const float* op = o + (t0 + j) * (size_t)d + h * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) D = fmaf(g[c], op[c], D);
        float dq[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = 0.f;
        for (int i = 0; i <= j; ++i) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) { s = fmaf(q[c], Ks[i][c], s); dp = fmaf(g[c], Vs[i][c], dp); }
            const float p = expf(s - m) * il;
            const float ds = p * (dp - D);
#pragma unroll
            for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, Ks[i][c], dq[c]);
        }
        sm_m[j] = m; sm_il[j] = il; sm_D[j] = D;
        float* dst = d_qkv + (t0 + j) * (size_t)(3 * d) + h * HD;
#pragma unroll
        for (int c4 = 0; c4 < HD; c4 += 4)
            *reinterpret_cast<float4*>(dst + c4) = make_float4(dq[c4] * scale, dq[c4 + 1] * scale, dq[c4 + 2] * scale, dq[c4 + 3] * scale);
    }
    __syncthreads();
    // phase B (thread = key row i): dK_i = sum_{j>=i} ds_ji (q_j * scale), dV_i = sum_{j>=i} p_ji dO_j
    const int i = tid;
    if (i < L) {
        float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { k[c] = Ks[i][c]; v[c] = Vs[i][c]; dk[c] = 0.f; dv[c] = 0.f; }
        for (int jj = i; jj < L; ++jj) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) { s = fmaf(Qs[jj][c], k[c], s); dp = fmaf(Gs[jj][c], v[c], dp); }
            const float p = expf(s - sm_m[jj]) * sm_il[jj];
            const float ds = p * (dp - sm_D[jj]);
#pragma unroll
            for (int c = 0; c < HD; ++c) { dk[c] = fmaf(ds, Qs[jj][c], dk[c]); dv[c] = fmaf(p, Gs[jj][c], dv[c]); }
        }
        float* dst = d_qkv + (t0 + i) * (size_t)(3 * d) + h * HD;
#pragma unroll
        for (int c4 = 0; c4 < HD; c4 += 4) {
            *reinterpret_cast<float4*>(dst + d + c4) = make_float4(dk[c4], dk[c4 + 1], dk[c4 + 2], dk[c4 + 3]);
            *reinterpret_cast<float4*>(dst + 2 * d + c4) = make_float4(dv[c4], dv[c4 + 1], dv[c4 + 2], dv[c4 + 3]);
        }
    }
}

// ---- embedding backward ------------------------------------------------------------------------------------------------------------
// position table: gpos[j, c] = sum_b dx0[b, j, c]   (one thread per (j, c), no atomics)
__global__ void pos_bwd_kernel(const float* __restrict__ dx0, int B, int L, int d, float* __restrict__ gpos) {
    pdl_sync();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * d) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dx0[(size_t)b * L * d + idx];
    gpos[idx] = s;
}

// obs embedding: continuous  gW[c,k] += sum_t dx0[t,c] obs[t,k];  discrete: through Embedding->Flatten->Linear.
// 32 tokens per CTA.
// Per-CTA partials [d*KI | d | vocab*E] summed in CTA order (no atomics; the table rows hit by several tokens of a CTA are
// accumulated in token order).
#define EMBED_INLINE_REDUCE_MAX 1024
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ dx0, dtqn_obs_src src, dtqn_net_cfg c, const float* __restrict__ params,
                 long long emb_table, long long emb_w, int L, int T, float* __restrict__ g_table,
                 float* __restrict__ g_w, float* __restrict__ g_b, float* __restrict__ part, unsigned* __restrict__ ticket) {
    pdl_sync();
    extern __shared__ float smem[];
    const int d = c.d_model, O = c.obs_dim, E = c.discrete ? c.embed_per_obs : 1, KI = O * E;
    const int n_tab = c.discrete ? c.vocab * E : 0, n_part = d * KI + d + n_tab;
    float* mypart = part + (size_t)blockIdx.x * n_part;
    float* sdx = smem;                 // [32][d]
    float* sin_ = smem + 32 * d;       // [32][KI]  input features of the Linear (obs or looked-up embeddings); then d in[r, f]
    int* stok = reinterpret_cast<int*>(sin_ + 32 * KI);   // [32][O] token ids (discrete)
    const int t0 = blockIdx.x * 32, nt = min(32, T - t0);
    for (int e = threadIdx.x; e < 32 * d; e += blockDim.x) {
        const int r = e / d;
        sdx[e] = r < nt ? dx0[(size_t)(t0 + r) * d + (e % d)] : 0.f;
    }
    for (int e = threadIdx.x; e < 32 * O; e += blockDim.x) {
        const int r = e / O, k = e % O;
        float ov = 0.f;
        if (r < nt) {
            const int t = t0 + r, i = t / L, j = t % L;
            ov = src.obs[(long long)i * src.seq_stride + (long long)j * O + k];
        }
        if (c.discrete) {
            int tok = (int)ov; tok = tok < 0 ? 0 : (tok >= c.vocab ? c.vocab - 1 : tok);
            stok[e] = tok;
            for (int ee = 0; ee < E; ++ee) sin_[r * KI + k * E + ee] = r < nt ? params[emb_table + tok * E + ee] : 0.f;
        } else sin_[r * KI + k] = ov;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < d * KI; e += blockDim.x) {          // gW[c, f] += sum_r dx[r, c] * in[r, f]
        const int cc = e / KI, f = e % KI;
        float s = 0.f;
        for (int r = 0; r < nt; ++r) s = fmaf(sdx[r * d + cc], sin_[r * KI + f], s);
        mypart[e] = s;
    }
    for (int cc = threadIdx.x; cc < d; cc += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < nt; ++r) s += sdx[r * d + cc];
        mypart[d * KI + cc] = s;
    }
    if (c.discrete) {
        __syncthreads();                                              // sin_ is reused for d in[r, f] = sum_c dx[r,c] W[c, f]
        for (int e = threadIdx.x; e < 32 * KI; e += blockDim.x) {
            const int r = e / KI, f = e % KI;
            float s = 0.f;
            if (r < nt)
                for (int cc = 0; cc < d; ++cc) s = fmaf(sdx[r * d + cc], __ldg(params + emb_w + (long long)cc * KI + f), s);
            sin_[e] = s;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < n_tab; e += blockDim.x) {       // d table[v, ee] = sum over (r, k) with tok(r, k) == v
            const int v = e / E, ee = e % E;
            float s = 0.f;
            for (int r = 0; r < nt; ++r)
                for (int k = 0; k < O; ++k)
                    if (stok[r * O + k] == v) s += sin_[r * KI + k * E + ee];
            mypart[d * KI + d + e] = s;
        }
    }
    // few gradient elements (continuous observations: d*O + d): the last CTA sums the partials; the ~10 k elements of a
    // discrete embedding go through embed_reduce_kernel (a single CTA walking 50 partial rows of each would serialise ~200 us)
    if (n_part > EMBED_INLINE_REDUCE_MAX) return;
    if (!last_chunk_arrived(ticket, gridDim.x)) return;
    for (int e = threadIdx.x; e < n_part; e += blockDim.x) {
        const float sum = ordered_sum(part + e, (size_t)n_part, gridDim.x);
        if (e < d * KI) g_w[e] = sum;
        else if (e < d * KI + d) g_b[e - d * KI] = sum;
        else g_table[e - d * KI - d] = sum;
    }
}
// CTA-ordered sum of the embedding-gradient partials: one thread per gradient element (the discrete embedding has ~10 k of
// them: a single last CTA walking 50 partial rows of each would serialise ~200 us)
__global__ void __launch_bounds__(256)
embed_reduce_kernel(const float* __restrict__ part, int n_blocks, int n_part, int n_w, int n_b, float* __restrict__ g_w,
                    float* __restrict__ g_b, float* __restrict__ g_table) {
    pdl_sync();
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n_part) return;
    const float sum = ordered_sum(part + e, (size_t)n_part, (unsigned)n_blocks);
    if (e < n_w) g_w[e] = sum;
    else if (e < n_w + n_b) g_b[e - n_w] = sum;
    else g_table[e - n_w - n_b] = sum;
}


// The weight-gradient GEMMs only feed the flat gradient buffer, so they run on a side stream, forked after the kernel
// that produces their dY operand and joined at the end of every layer (before the dY buffers are overwritten).  Under
// CUDA-graph capture the event record / wait pairs become graph edges.  The stream and events are created once per process.
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[8];
    int next = 0;
    bool ok = false;
    void init() {
        if (ok) return;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return;
        for (int i = 0; i < 8; ++i) if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return;
        ok = true;
    }
    void fork(cudaStream_t main_st) {           // side stream waits for everything issued on main so far
        cudaEvent_t e = ev[next]; next = (next + 1) & 7;
        cudaEventRecord(e, main_st);
        cudaStreamWaitEvent(st, e, 0);
    }
    void join(cudaStream_t main_st) {           // main waits for everything issued on the side stream so far
        cudaEvent_t e = ev[next]; next = (next + 1) & 7;
        cudaEventRecord(e, st);
        cudaStreamWaitEvent(main_st, e, 0);
    }
};
SideStream g_side;
int g_parallel_wgrad = 1;
int g_wgrad_chunk = 2 * WG_CHUNK;  // tokens per split-K chunk of the weight-gradient GEMMs (multiple of 16, >= WG_CHUNK); 128 measured best

// Defaults = the fastest setting of tools/tune_bwd.py on the bench shape (32 x 50 tokens; profiles/r2_tune_bwd.txt: training
// half 0.281 -> 0.249 ms against 64 rows / unfused / 64 tokens), and the one the network parity tests ran under.
int g_head_tok = HB_TOK_MIN;        // tokens per CTA of head_bwd_kernel (dtqn_set_head_bwd_tokens)
int g_dgrad_rows = 32;              // rows per CTA of the dgrad GEMMs (dtqn_set_dgrad_rows)
int g_fuse_ln_bwd = 1;              // LayerNorm backward as the epilogue of the dgrad that produces its dy (dtqn_set_fuse_ln_bwd)

template <int BM, int BN, int EPI, bool LN>
void launch_dgrad_tile(const float* dY, const float* W, const float* aux, float* dX, int T, int Nf, int Kf, const LnBwdArgs& ln,
                       cudaStream_t st) {
    launch_k(dgrad_kernel<BM, BN, EPI, LN>, dim3(dtqn_cdiv(T, BM), Kf / BN, 1), GEMM_THREADS, 0, st, dY, W, aux, dX, T, Nf, Kf, ln);
}
template <int BN, int EPI, bool LN>
void launch_dgrad_rows(int rows, const float* dY, const float* W, const float* aux, float* dX, int T, int Nf, int Kf,
                       const LnBwdArgs& ln, cudaStream_t st) {
    if (LN && BN == 128 && rows == 64) rows = 32;              // the 64 x 129 dy tile does not fit the GEMM slabs
    if (rows == 16) launch_dgrad_tile<16, BN, EPI, LN>(dY, W, aux, dX, T, Nf, Kf, ln, st);
    else if (rows == 32) launch_dgrad_tile<32, BN, EPI, LN>(dY, W, aux, dX, T, Nf, Kf, ln, st);
    else if constexpr (!(LN && BN == 128)) launch_dgrad_tile<64, BN, EPI, LN>(dY, W, aux, dX, T, Nf, Kf, ln, st);
}

template <int EPI>
int launch_dgrad(const float* dY, const float* W, const float* aux, float* dX, int T, int Nf, int Kf, cudaStream_t st) {
    const LnBwdArgs none{};
    prof_begin(PROF_DGRAD, st);
    if (Kf % 128 == 0) launch_dgrad_rows<128, EPI, false>(g_dgrad_rows, dY, W, aux, dX, T, Nf, Kf, none, st);
    else if (Kf % 64 == 0) launch_dgrad_rows<64, EPI, false>(g_dgrad_rows, dY, W, aux, dX, T, Nf, Kf, none, st);
    else return DTQN_E_UNSUPPORTED;
    prof_end(PROF_DGRAD, st, 2.0 * (double)T * Nf * Kf);
    DTQN_LAUNCH_CHECK();
    return 0;
}

// dgrad whose output is the dy of a LayerNorm (Kf == d_model == 64 or 128): LayerNorm backward in the epilogue
template <int EPI>
int launch_dgrad_ln(const float* dY, const float* W, const float* aux, int T, int Nf, int d, const LnBwdArgs& ln, cudaStream_t st) {
    prof_begin(PROF_DGRAD, st);
    if (d == 128) launch_dgrad_rows<128, EPI, true>(g_dgrad_rows, dY, W, aux, nullptr, T, Nf, d, ln, st);
    else if (d == 64) launch_dgrad_rows<64, EPI, true>(g_dgrad_rows, dY, W, aux, nullptr, T, Nf, d, ln, st);
    else return DTQN_E_UNSUPPORTED;
    prof_end(PROF_DGRAD, st, 2.0 * (double)T * Nf * d);
    DTQN_LAUNCH_CHECK();
    return 0;
}

// pW / pb: chunk-partial areas of this weight / bias (same offsets as in the flat gradient), pstride floats between chunks
int launch_wgrad(const float* dY, const float* X, int T, int Nf, int Kf, float* gb, float* pW, float* pb,
                 long long pstride, cudaStream_t st) {
    if (Nf % 64 || Kf % 64) return DTQN_E_UNSUPPORTED;
    dim3 grid(Nf / 64, Kf / 64, dtqn_cdiv(T, g_wgrad_chunk));
    prof_begin(PROF_WGRAD, st);
    wgrad_kernel<<<grid, 256, 0, st>>>(dY, X, T, Nf, Kf, gb, pW, pb, pstride, g_wgrad_chunk);
    prof_end(PROF_WGRAD, st, 2.0 * (double)T * Nf * Kf);
    DTQN_LAUNCH_CHECK();
    return 0;
}

struct BwdScratch {
    float *dq, *g_hh, *gx, *gu, *ga, *ga1, *gh, *gqkv, *go, *gx1, *ga_alt, *partial;
    float *pgrad;            // [n_chunks][flat parameter count] split-K partials of the weight / bias gradients
    float *psmall;           // per-CTA partials of the head / LayerNorm / embedding gradients (one launch at a time)
    unsigned* ticket;        // [0] td loss; [1] head; [2] LayerNorm; [3] embedding; [8, 72) weight-gradient tiles
    long long total;
};
long long bwd_scratch_layout(const dtqn_net_cfg& c, long long T0, float* base, BwdScratch& s) {
    const long long d = c.d_model;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o = al4(o + n); return p; };
    s.dq = take(T0 * c.num_actions); s.g_hh = take(T0 * d); s.gx = take(T0 * d); s.gu = take(T0 * d);
    s.ga = take(T0 * d); s.ga1 = take(T0 * d); s.gh = take(T0 * 4 * d); s.gqkv = take(T0 * 3 * d); s.go = take(T0 * d); s.gx1 = take(T0 * d);
    s.ga_alt = take(T0 * d);       // fused LayerNorm backward: da of the next layer is written before this layer's wgrad(ga) is joined
    s.partial = take((long long)dtqn_cdiv(T0, 256) * 8);
    NetLayout lay;
    net_layout(c, lay);
    s.pgrad = take((long long)dtqn_cdiv(T0, WG_CHUNK) * lay.total);
    const long long n_emb = d * lay.k_in + d + (c.discrete ? (long long)c.vocab * c.embed_per_obs : 0);
    long long small = (long long)dtqn_cdiv(T0, HB_TOK_MIN) * (c.num_actions * d + c.num_actions);
    if ((long long)dtqn_cdiv(T0, 16) * 2 * d > small) small = (long long)dtqn_cdiv(T0, 16) * 2 * d;   // LayerNorm: 32 (own kernel) or >= 16 rows per CTA
    if ((long long)dtqn_cdiv(T0, 32) * n_emb > small) small = (long long)dtqn_cdiv(T0, 32) * n_emb;
    s.psmall = take(small);
    s.ticket = reinterpret_cast<unsigned*>(take(72));
    s.total = o;
    return o;
}

}  // namespace

extern "C" int dtqn_set_parallel_wgrad(int32_t on) { g_parallel_wgrad = on; return 0; }
extern "C" int dtqn_set_dgrad_rows(int32_t rows) {
    if (rows != 64 && rows != 32 && rows != 16) return DTQN_E_ARG;
    g_dgrad_rows = rows;
    return 0;
}
extern "C" int dtqn_set_head_bwd_tokens(int32_t tokens) {
    if (tokens != 64 && tokens != 32 && tokens != 16) return DTQN_E_ARG;
    g_head_tok = tokens;
    return 0;
}
extern "C" int dtqn_set_fuse_ln_bwd(int32_t on) { g_fuse_ln_bwd = on ? 1 : 0; return 0; }
extern "C" int dtqn_set_wgrad_chunk(int32_t tokens) {
    if (tokens < WG_CHUNK || tokens % 16) return DTQN_E_ARG;      // the partial buffer is sized for WG_CHUNK-token chunks
    g_wgrad_chunk = tokens;
    return 0;
}

extern "C" int64_t dtqn_td_scratch_floats(const dtqn_net_cfg* cfg, int32_t batch, int32_t seq_len) {
    if (!cfg || batch < 1 || seq_len < 1) return DTQN_E_ARG;
    const long long T0 = (long long)batch * seq_len;
    if (cfg_is_variant(*cfg))       // dq | td partials | ticket | scratch of net_var.cu
        return al4(T0 * cfg->num_actions) + al4((long long)dtqn_cdiv(T0, 256) * 8) + 4 + var_bwd_scratch_floats(*cfg, T0);
    BwdScratch s;
    return bwd_scratch_layout(*cfg, T0, nullptr, s);
}

extern "C" int dtqn_td_backward(const dtqn_net_cfg* cfg, const float* params, const dtqn_obs_src* obs_src,
                                const float* q_all, const uint8_t* act_win, const float* rew, const uint8_t* done,
                                int32_t B, int32_t L, int32_t history, float gamma, float* ws, int64_t ws_floats,
                                float* scratch, float* grads, float* stats_out, void* stream) {
    if (!cfg || !params || !obs_src || !obs_src->obs || obs_src->timestep || !q_all || !act_win || !rew || !done ||
        !ws || !scratch || !grads || !stats_out || B < 1 || L < 1 || history < 1 || history > L) return DTQN_E_ARG;
    NetLayout lay;
    int rc = net_layout(*cfg, lay);
    if (rc) return rc;
    const long long T0 = (long long)B * L, T = 3 * T0;
    if (cfg_is_variant(*cfg)) {                                // ablation flags: TD loss here, backward in net_var.cu
        cudaStream_t vst = (cudaStream_t)stream;
        float* dq = scratch;
        float* partial = dq + al4(T0 * cfg->num_actions);
        unsigned* ticket = reinterpret_cast<unsigned*>(partial + al4((long long)dtqn_cdiv(T0, 256) * 8));
        float* vscratch = reinterpret_cast<float*>(ticket) + 4;
        cudaError_t ce = cudaMemsetAsync(grads, 0, sizeof(float) * lay.total, vst);
        if (ce != cudaSuccess) return (int)ce;
        prof_begin(PROF_TD, vst);
        td_loss_kernel<<<dtqn_cdiv(T0, 256), 256, 0, vst>>>(q_all, act_win, rew, done, B, L, cfg->num_actions, history, gamma, dq,
                                                             partial, ticket, stats_out);
        prof_end(PROF_TD, vst, 0.0);
        DTQN_LAUNCH_CHECK();
        return var_td_backward(*cfg, lay, params, obs_src, dq, B, L, ws, ws_floats, vscratch, grads, vst);
    }
    NetAct act;
    if (net_act_layout(*cfg, T, 1, ws, act) > ws_floats) return DTQN_E_ARG;
    BwdScratch s;
    bwd_scratch_layout(*cfg, T0, scratch, s);
    cudaStream_t st = (cudaStream_t)stream;
    const int d = cfg->d_model, H = cfg->n_heads, hd = d / H, A = cfg->num_actions, Ti = (int)T0;
    cudaError_t ce = cudaMemsetAsync(grads, 0, sizeof(float) * lay.total, st);
    if (ce != cudaSuccess) return (int)ce;

    prof_begin(PROF_TD, st);
    launch_k(td_loss_kernel, dtqn_cdiv(T0, 256), 256, 0, st, q_all, act_win, rew, done, B, L, A, history, gamma, s.dq, s.partial,
             s.ticket, stats_out);
    prof_end(PROF_TD, st, 0.0);
    DTQN_LAUNCH_CHECK();
    // head: ffn.2 then ffn.0 (group 0 rows are the first T0 rows of every activation buffer)
    prof_begin(PROF_HEAD, st);
    launch_k(head_bwd_kernel, dtqn_cdiv(T0, g_head_tok), 256, 0, st, s.dq, (const float*)act.hh, params + lay.h2_w, Ti, d, A, g_head_tok, s.g_hh,
             grads + lay.h2_w, grads + lay.h2_b, s.psmall, s.ticket + 1);
    prof_end(PROF_HEAD, st, 4.0 * (double)T0 * d * A);
    DTQN_LAUNCH_CHECK();
    g_side.init();
    const bool par = g_parallel_wgrad && g_side.ok;
    cudaStream_t ws_ = par ? g_side.st : st;                       // stream of the weight-gradient GEMMs
    auto fork = [&]() { if (par) g_side.fork(st); };
    // weight gradient of the Linear whose weight / bias sit at w_off / b_off of the flat layout (split-K, fixed-order sum)
    auto wgrad = [&](const float* dY, const float* X, int Nf, int Kf, long long w_off, long long b_off) {
        return launch_wgrad(dY, X, Ti, Nf, Kf, grads + b_off, s.pgrad + w_off, s.pgrad + b_off, lay.total, ws_);
    };
    const float* x_last = act.layer[cfg->n_layers - 1].x2;
    // LayerNorm backward either as its own kernel or (dtqn_set_fuse_ln_bwd) as the epilogue of the dgrad producing its dy.  Fused, the
    // LN2 of layer li - 1 runs inside layer li's in_proj dgrad, i.e. before layer li's weight-gradient GEMMs are joined, so its da
    // alternates between two buffers (ga / ga_alt) instead of overwriting the one wgrad(ffn.2) of layer li may still be reading.
    const bool fuse = g_fuse_ln_bwd && (d == 64 || d == 128);
    auto ga_of = [&](int li) { return (fuse && (li & 1)) ? s.ga_alt : s.ga; };
    auto ln2_args = [&](int li) {
        const LayerOff& lo = lay.layer[li];
        const LayerAct& la = act.layer[li];
        return LnBwdArgs{la.x1, la.r2, la.st2, params + lo.ln2_w, s.gu, ga_of(li), grads + lo.ln2_w, grads + lo.ln2_b, s.psmall, s.ticket + 2};
    };
    fork();
    if ((rc = wgrad(s.g_hh, x_last, d, d, lay.h1_w, lay.h1_b))) return rc;
    if (fuse) { if ((rc = launch_dgrad_ln<DG_NONE>(s.g_hh, params + lay.h1_w, nullptr, Ti, d, d, ln2_args(cfg->n_layers - 1), st))) return rc; }
    else if ((rc = launch_dgrad<DG_NONE>(s.g_hh, params + lay.h1_w, nullptr, s.gx, Ti, d, d, st))) return rc;
    for (int li = cfg->n_layers - 1; li >= 0; --li) {
        const LayerOff& lo = lay.layer[li];
        const LayerAct& la = act.layer[li];
        const float* x_in = li == 0 ? act.x0 : act.layer[li - 1].x2;
        float* ga = ga_of(li);
        if (!fuse) {
            // LN2 backward: dy = gx -> gu (du2), ga (d ffn.2 output)
            prof_begin(PROF_LN_BWD, st);
            if (d == 64) launch_k(ln_bwd_kernel<64>, dtqn_cdiv(T0, 32), 256, 0, st, s.gx, la.x1, la.r2, la.st2, params + lo.ln2_w, Ti, s.gu, s.ga, grads + lo.ln2_w, grads + lo.ln2_b, s.psmall, s.ticket + 2);
            else         launch_k(ln_bwd_kernel<128>, dtqn_cdiv(T0, 32), 256, 0, st, s.gx, la.x1, la.r2, la.st2, params + lo.ln2_w, Ti, s.gu, s.ga, grads + lo.ln2_w, grads + lo.ln2_b, s.psmall, s.ticket + 2);
            prof_end(PROF_LN_BWD, st, 0.0);
            DTQN_LAUNCH_CHECK();
        }
        // ffn.2
        fork();
        if ((rc = wgrad(ga, la.h, d, 4 * d, lo.f2_w, lo.f2_b))) return rc;
        if ((rc = launch_dgrad<DG_MASK>(ga, params + lo.f2_w, la.h, s.gh, Ti, d, 4 * d, st))) return rc;
        // ffn.0, then LN1 backward: dy = gx1 = gh W + du2 -> gu (du1), ga1 (d out_proj output)
        fork();
        if ((rc = wgrad(s.gh, la.x1, 4 * d, d, lo.f1_w, lo.f1_b))) return rc;
        if (fuse) {
            const LnBwdArgs ln1{x_in, la.r1, la.st1, params + lo.ln1_w, s.gu, s.ga1, grads + lo.ln1_w, grads + lo.ln1_b, s.psmall, s.ticket + 2};
            if ((rc = launch_dgrad_ln<DG_ADD>(s.gh, params + lo.f1_w, s.gu, Ti, 4 * d, d, ln1, st))) return rc;
        } else {
            if ((rc = launch_dgrad<DG_ADD>(s.gh, params + lo.f1_w, s.gu, s.gx1, Ti, 4 * d, d, st))) return rc;
            prof_begin(PROF_LN_BWD, st);
            if (d == 64) launch_k(ln_bwd_kernel<64>, dtqn_cdiv(T0, 32), 256, 0, st, s.gx1, x_in, la.r1, la.st1, params + lo.ln1_w, Ti, s.gu, s.ga1, grads + lo.ln1_w, grads + lo.ln1_b, s.psmall, s.ticket + 2);
            else         launch_k(ln_bwd_kernel<128>, dtqn_cdiv(T0, 32), 256, 0, st, s.gx1, x_in, la.r1, la.st1, params + lo.ln1_w, Ti, s.gu, s.ga1, grads + lo.ln1_w, grads + lo.ln1_b, s.psmall, s.ticket + 2);
            prof_end(PROF_LN_BWD, st, 0.0);
            DTQN_LAUNCH_CHECK();
        }
        // out_proj
        fork();
        if ((rc = wgrad(s.ga1, la.o, d, d, lo.out_w, lo.out_b))) return rc;
        if ((rc = launch_dgrad<DG_NONE>(s.ga1, params + lo.out_w, nullptr, s.go, Ti, d, d, st))) return rc;
        // attention core
        {
            dim3 grid(H, (unsigned)B);
            const int thr = L <= 64 ? 64 : 128;
            const float scale = 1.0f / sqrtf((float)hd);
            prof_begin(PROF_ATTN_BWD, st);
            if (hd == 8) launch_k(attn_bwd_kernel<8>, grid, thr, 0, st, la.qkv, la.o, s.go, s.gqkv, L, d, scale);
            else if (hd == 16) launch_k(attn_bwd_kernel<16>, grid, thr, 0, st, la.qkv, la.o, s.go, s.gqkv, L, d, scale);
            else if (hd == 4) launch_k(attn_bwd_kernel<4>, grid, thr, 0, st, la.qkv, la.o, s.go, s.gqkv, L, d, scale);
            else return DTQN_E_UNSUPPORTED;
            prof_end(PROF_ATTN_BWD, st, 8.0 * (double)T0 * L * d);
            DTQN_LAUNCH_CHECK();
        }
        // in_proj (+ the LN2 backward of the layer below)
        fork();
        if ((rc = wgrad(s.gqkv, x_in, 3 * d, d, lo.in_w, lo.in_b))) return rc;
        if (fuse && li > 0) { if ((rc = launch_dgrad_ln<DG_ADD>(s.gqkv, params + lo.in_w, s.gu, Ti, 3 * d, d, ln2_args(li - 1), st))) return rc; }
        else if ((rc = launch_dgrad<DG_ADD>(s.gqkv, params + lo.in_w, s.gu, s.gx, Ti, 3 * d, d, st))) return rc;
        if (par) g_side.join(st);          // the next layer overwrites gh / gqkv / ga1 (and, unfused, ga)
    }
    // every weight-gradient GEMM has been joined: chunk-ordered sum of their partials into the flat gradient
    {
        WgradSegs segs{};
        long long tot = 0;
        for (int li = 0; li <= cfg->n_layers; ++li) {
            const long long b = li < cfg->n_layers ? lay.layer[li].in_w : lay.h1_w;
            const long long e = li < cfg->n_layers ? lay.layer[li].f2_b + d : lay.h1_b + d;
            segs.begin[li] = b; segs.len4[li] = (e - b) / 4; tot += (e - b) / 4;
        }
        segs.n = cfg->n_layers + 1; segs.total4 = tot;
        prof_begin(PROF_WGRAD, st);
        launch_k(wgrad_reduce_kernel, dtqn_cdiv(tot, 256), 256, 0, st, (const float*)s.pgrad, lay.total, dtqn_cdiv(Ti, g_wgrad_chunk), segs, grads);
        prof_end(PROF_WGRAD, st, 0.0);
        DTQN_LAUNCH_CHECK();
    }
    // embedding + position table
    prof_begin(PROF_OTHER, st);
    if (cfg->pos_trainable) {
        launch_k(pos_bwd_kernel, dtqn_cdiv((long long)L * d, 256), 256, 0, st, (const float*)s.gx, B, L, d, grads + lay.pos);
        DTQN_LAUNCH_CHECK();
    }
    {
        const int KI = lay.k_in;
        const size_t smem = sizeof(float) * (32 * d + 32 * KI) + sizeof(int) * 32 * cfg->obs_dim;
        launch_k(embed_bwd_kernel, dtqn_cdiv(T0, 32), 256, smem, st, (const float*)s.gx, *obs_src, *cfg, params, lay.emb_table, lay.emb_w,
                 L, Ti, cfg->discrete ? grads + lay.emb_table : nullptr, grads + lay.emb_w, grads + lay.emb_b, s.psmall, s.ticket + 3);
        DTQN_LAUNCH_CHECK();
        const int n_part = d * KI + d + (cfg->discrete ? cfg->vocab * cfg->embed_per_obs : 0);
        if (n_part > EMBED_INLINE_REDUCE_MAX) {
            launch_k(embed_reduce_kernel, dtqn_cdiv(n_part, 256), 256, 0, st, (const float*)s.psmall, dtqn_cdiv(T0, 32), n_part, d * KI, d,
                     grads + lay.emb_w, grads + lay.emb_b, cfg->discrete ? grads + lay.emb_table : nullptr);
            DTQN_LAUNCH_CHECK();
        }
    }
    prof_end(PROF_OTHER, st, 0.0);
    return 0;
}
