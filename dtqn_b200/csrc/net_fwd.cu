// DTQN Q-network forward (dtqn/networks/dtqn.py:158-218) as sm_100a kernels.
//   embed_kernel      obs embedding (Linear(O,d) or Embedding->Flatten->Linear, representations.py:17-75) + position
//                     table add (dtqn.py:195-199); reads a dense batch, the replay window or the acting-context ring
//   linear_kernel     x W^T + b with fused epilogues: bias | bias+ReLU | bias -> ReLU -> +residual -> LayerNorm
//                     (transformer.py:64-78: x = LN(x + relu(sublayer(x))), gates.py:40-41)
//   attn_fwd_kernel   causal multi-head self-attention core (softmax((q/sqrt(hd)) k^T + mask) v), one CTA per
//                     (sequence, head), online softmax in registers (nn.MultiheadAttention closed form)
//   head_kernel       final Linear(d, A) (dtqn.py:149-153), one warp per token
#include "net.cuh"
#include "gemm_simt.cuh"
#include "prof.cuh"
#include "linear_tc.cuh"
#include "attn_mma.cuh"

namespace {


// ---- observation embedding + position ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_kernel(GroupPtrs P, GroupSrc S, dtqn_net_cfg c, long long emb_table, long long emb_w, long long emb_b,
             long long pos_off, int n_seq, int L, float obs_mask, float* __restrict__ x0) {
    const int g = blockIdx.z;
    const int d = c.d_model, dq = d / 4;
    const long long Tg = (long long)n_seq * L;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Tg * dq) return;
    const long long t = idx / dq;
    const int c0 = (int)(idx % dq) * 4;
    const int i = (int)(t / L), j = (int)(t % L);
    const dtqn_obs_src& s = S.s[g];
    int row = j;
    bool valid = true;
    if (s.timestep) {
        const int ts = s.timestep[i];
        const int n = min(s.ring_len, ts + 1);
        valid = j < n;
        row = valid ? (ts + 1 - n + j) % s.ring_len : 0;
    }
    const float* o = s.obs + (long long)i * s.seq_stride + (long long)row * c.obs_dim;
    const float* p = P.p[g];
    float acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = __ldg(p + emb_b + c0 + q);
    if (!c.discrete) {
        const int O = c.obs_dim;
        for (int k = 0; k < O; ++k) {
            const float ov = valid ? __ldg(o + k) : obs_mask;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(ov, __ldg(p + emb_w + (long long)(c0 + q) * O + k), acc[q]);
        }
    } else {
        const int O = c.obs_dim, E = c.embed_per_obs, KI = O * E;
        for (int k = 0; k < O; ++k) {
            const float ov = valid ? __ldg(o + k) : obs_mask;
            int tok = (int)ov;
            tok = tok < 0 ? 0 : (tok >= c.vocab ? c.vocab - 1 : tok);
            for (int e = 0; e < E; ++e) {
                const float tv = __ldg(p + emb_table + tok * E + e);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc[q] = fmaf(tv, __ldg(p + emb_w + (long long)(c0 + q) * KI + k * E + e), acc[q]);
            }
        }
    }
    const float4 pv = *reinterpret_cast<const float4*>(p + pos_off + (long long)j * d + c0);
    float4 out = make_float4(acc[0] + pv.x, acc[1] + pv.y, acc[2] + pv.z, acc[3] + pv.w);
    *reinterpret_cast<float4*>(x0 + ((long long)g * Tg + t) * d + c0) = out;
}


// Continuous observations, fast path: 64 tokens per CTA.  Observation rows (ring-resolved) and the tiny Linear(O, d)
// are staged in shared memory once; each thread then emits 16 contiguous channels of one token (4 threads = one 256 B row).
template <int D>
__global__ void __launch_bounds__(256)
embed_cont_kernel(GroupPtrs P, GroupSrc S, int O, long long emb_w, long long emb_b, long long pos_off, int n_seq, int L,
                  float obs_mask, float* __restrict__ x0) {
    constexpr int TOK = 1024 / (D / 16) / 4;                  // tokens per CTA: 64 (D = 64) or 32 (D = 128)
    __shared__ float sW[D * 16];
    __shared__ float sB[D];
    __shared__ float sObs[TOK][16];
    const int g = blockIdx.z, tid = threadIdx.x;
    const float* p = P.p[g];
    const long long Tg = (long long)n_seq * L;
    const long long t0 = (long long)blockIdx.x * TOK;
    for (int e = tid; e < D * O; e += 256) sW[e] = __ldg(p + emb_w + e);
    for (int e = tid; e < D; e += 256) sB[e] = __ldg(p + emb_b + e);
    if (tid < TOK) {
        const long long t = t0 + tid;
        if (t < Tg) {
            const int i = (int)(t / L), j = (int)(t % L);
            const dtqn_obs_src& s = S.s[g];
            int row = j; bool valid = true;
            if (s.timestep) {
                const int ts = s.timestep[i];
                const int n = min(s.ring_len, ts + 1);
                valid = j < n;
                row = valid ? (ts + 1 - n + j) % s.ring_len : 0;
            }
            const float* o = s.obs + (long long)i * s.seq_stride + (long long)row * O;
            for (int k = 0; k < O; ++k) sObs[tid][k] = valid ? __ldg(o + k) : obs_mask;
        }
    }
    __syncthreads();
    constexpr int TPT = D / 16;                               // threads per token
    const int tl = tid / TPT, c0 = (tid % TPT) * 16;
    const long long t = t0 + tl;
    if (tl >= TOK || t >= Tg) return;
    const int j = (int)(t % L);
    float ov[16];
    for (int k = 0; k < O; ++k) ov[k] = sObs[tl][k];
    const float* pp = p + pos_off + (long long)j * D + c0;
    float* out = x0 + ((long long)g * Tg + t) * D + c0;
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(pp + q);
        float acc[4] = {sB[c0 + q], sB[c0 + q + 1], sB[c0 + q + 2], sB[c0 + q + 3]};
        for (int k = 0; k < O; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(ov[k], sW[(c0 + q + e) * O + k], acc[e]);
        }
        *reinterpret_cast<float4*>(out + q) = make_float4(acc[0] + pv.x, acc[1] + pv.y, acc[2] + pv.z, acc[3] + pv.w);
    }
}

// Discrete observations, table-lookup form.  Embedding -> Flatten -> Linear is linear in the one-hot of every feature, so
//   x[c] = b[c] + sum_k LUT[k][tok_k][c],   LUT[k][v][c] = sum_e table[v][e] * W[c][k*E + e]
// (representations.py:47-51 regrouped: the e-sum of a feature first, then the features in order).  embed_lut_build_kernel
// writes the O x vocab x D table of every group's network once per forward (vocab is 9 for Memory-5: 46 KB at D = 128, 92 k
// MACs); embed_lut_kernel stages it in shared memory and streams tokens: one warp per token (pair), a lane adds O float4 rows
// of the table + the position row and writes 16 bytes -- O adds per channel instead of O*E multiply-adds, bound by the write
// of x0 (HBM) instead of the fp32 pipe.
__global__ void __launch_bounds__(256)
embed_lut_build_kernel(GroupPtrs P, int O, int E, int vocab, int D, long long emb_table, long long emb_w, float* __restrict__ lut) {
    const int g = blockIdx.y, n = O * vocab * D, KI = O * E;
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    const float* p = P.p[g];
    const int c = e % D, v = (e / D) % vocab, k = e / (D * vocab);
    float a = 0.f;
    for (int q = 0; q < E; ++q) a = fmaf(__ldg(p + emb_table + v * E + q), __ldg(p + emb_w + (long long)c * KI + k * E + q), a);
    lut[(size_t)g * n + e] = a;
}
template <int D>
__global__ void __launch_bounds__(256)
embed_lut_kernel(GroupPtrs P, GroupSrc S, int O, int vocab, const float* __restrict__ lut, long long emb_b,
                 long long pos_off, int n_seq, int L, float obs_mask, float* __restrict__ x0) {
    extern __shared__ __align__(16) float el_sm[];
    float* sLut = el_sm;                                       // [O][vocab][D]
    float* sB = sLut + (size_t)O * vocab * D;                  // [D]
    const int g = blockIdx.z, tid = threadIdx.x, n_lut = O * vocab * D;
    const float* p = P.p[g];
    for (int e = tid * 4; e < n_lut; e += 1024)
        *reinterpret_cast<float4*>(sLut + e) = __ldg(reinterpret_cast<const float4*>(lut + (size_t)g * n_lut + e));
    for (int e = tid; e < D; e += 256) sB[e] = __ldg(p + emb_b + e);
    __syncthreads();
    constexpr int LPT = D / 4;                                 // lanes per token (16 or 32), one float4 of channels each
    constexpr int TPW = 32 / LPT;                              // tokens per warp pass
    const int lane = tid & 31, warp = tid >> 5;
    const int c4 = (lane % LPT) * 4, sub = lane / LPT;
    const dtqn_obs_src s = S.s[g];
    const long long Tg = (long long)n_seq * L;
    for (long long t = ((long long)blockIdx.x * 8 + warp) * TPW + sub; t < Tg; t += (long long)gridDim.x * 8 * TPW) {
        const int i = (int)(t / L), j = (int)(t % L);
        int row = j; bool valid = true;
        if (s.timestep) {
            const int ts = __ldg(s.timestep + i);
            const int n = min(s.ring_len, ts + 1);
            valid = j < n;
            row = valid ? (ts + 1 - n + j) % s.ring_len : 0;
        }
        const float* ob = s.obs + (long long)i * s.seq_stride + (long long)row * O;
        float4 acc = *reinterpret_cast<const float4*>(sB + c4);
        for (int k = 0; k < O; ++k) {
            int tok = (int)(valid ? __ldg(ob + k) : obs_mask);
            tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
            const float4 r = *reinterpret_cast<const float4*>(sLut + ((size_t)k * vocab + tok) * D + c4);
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        }
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p + pos_off + (long long)j * D + c4));
        *reinterpret_cast<float4*>(x0 + ((long long)g * Tg + t) * D + c4) = make_float4(acc.x + pv.x, acc.y + pv.y, acc.z + pv.z, acc.w + pv.w);
    }
}

// Discrete observations (Embedding(vocab, E) per feature -> Flatten -> Linear(O*E, D), representations.py:47-51), fast path:
// ED_TOK tokens per CTA.  The Linear weight is staged TRANSPOSED in shared memory once per CTA ([KI][D + 4]: float4 reads
// along the channels are conflict-free), every token's flattened embedding row is gathered from the (staged) table, and each
// thread then emits 16 contiguous channels of one token with the reference's summation order (k ascending).  The generic
// embed_kernel reads W with a stride of KI floats across the lanes of a warp: 7.6 ms for 4096 x 50 Memory-5 tokens.
constexpr int ED_TOK = 64;
template <int D>
__global__ void __launch_bounds__(256)
embed_disc_kernel(GroupPtrs P, GroupSrc S, int O, int E, int vocab, long long emb_table, long long emb_w, long long emb_b,
                  long long pos_off, int n_seq, int L, float obs_mask, float* __restrict__ x0) {
    extern __shared__ __align__(16) float ed_sm[];
    constexpr int LDW = D + 4;
    const int KI = O * E;
    float* sWT = ed_sm;                                        // [KI][LDW]
    float* sB = sWT + KI * LDW;                                // [D]
    float* sTab = sB + D;                                      // [vocab * E]
    float* sE = sTab + ((vocab * E + 3) & ~3);                 // [ED_TOK][KI]
    const int g = blockIdx.z, tid = threadIdx.x;
    const float* p = P.p[g];
    const long long Tg = (long long)n_seq * L;
    for (int c = tid / KI, k = tid % KI, e = tid; e < D * KI; e += 256) {   // coalesced read of W[c][k], transposed store
        sWT[k * LDW + c] = __ldg(p + emb_w + e);
        k += 256; while (k >= KI) { k -= KI; ++c; }
    }
    for (int e = tid; e < D; e += 256) sB[e] = __ldg(p + emb_b + e);
    for (int e = tid; e < vocab * E; e += 256) sTab[e] = __ldg(p + emb_table + e);
    constexpr int TPT = D / 16;                               // threads per token
    constexpr int TPP = 256 / TPT;                            // tokens per pass
    // channel quads of this thread: c0 + q * CQ (q = 0..3) -- the TPT threads of a token read CONTIGUOUS float4s of a weight
    // row (conflict-free; a 16-channel block per thread would put every second thread on the same banks)
    constexpr int CQ = TPT * 4;
    const int c0 = (tid % TPT) * 4;
    const dtqn_obs_src s = S.s[g];
    // persistent over token chunks: the weight image is staged once per CTA
    for (long long t0 = (long long)blockIdx.x * ED_TOK; t0 < Tg; t0 += (long long)gridDim.x * ED_TOK) {
        __syncthreads();                                      // staging done / previous chunk's rows consumed
        for (int e = tid; e < ED_TOK * O; e += 256) {          // one (token, feature) pair per thread: gather E table values
            const int tl = e / O, k = e % O;
            const long long t = t0 + tl;
            if (t < Tg) {
                const int i = (int)(t / L), j = (int)(t % L);
                int row = j; bool valid = true;
                if (s.timestep) {
                    const int ts = s.timestep[i];
                    const int n = min(s.ring_len, ts + 1);
                    valid = j < n;
                    row = valid ? (ts + 1 - n + j) % s.ring_len : 0;
                }
                const float ov = valid ? __ldg(s.obs + (long long)i * s.seq_stride + (long long)row * O + k) : obs_mask;
                int tok = (int)ov;
                tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
                for (int q = 0; q < E; ++q) sE[tl * KI + k * E + q] = sTab[tok * E + q];
            }
        }
        __syncthreads();
        for (int tl = tid / TPT; tl < ED_TOK; tl += TPP) {
            const long long t = t0 + tl;
            if (t >= Tg) break;
            const int j = (int)(t % L);
            float acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = sB[c0 + (q >> 2) * CQ + (q & 3)];
            const float* er = sE + tl * KI;
#pragma unroll 4
            for (int kk = 0; kk < KI; ++kk) {
                const float tv = er[kk];
                const float* w = sWT + kk * LDW + c0;
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w + (q >> 2) * CQ);
                    acc[q] = fmaf(tv, w4.x, acc[q]); acc[q + 1] = fmaf(tv, w4.y, acc[q + 1]);
                    acc[q + 2] = fmaf(tv, w4.z, acc[q + 2]); acc[q + 3] = fmaf(tv, w4.w, acc[q + 3]);
                }
            }
            const float* pp = p + pos_off + (long long)j * D + c0;
            float* out = x0 + ((long long)g * Tg + t) * D + c0;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                const float4 pv = __ldg(reinterpret_cast<const float4*>(pp + (q >> 2) * CQ));
                *reinterpret_cast<float4*>(out + (q >> 2) * CQ) = make_float4(acc[q] + pv.x, acc[q + 1] + pv.y, acc[q + 2] + pv.z, acc[q + 3] + pv.w);
            }
        }
    }
}

// ---- Linear with fused epilogues --------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
linear_kernel(LinArgs a) {
    __shared__ GemmSmem<BN> sm;
    constexpr int TN = BN / 16;
    const int g = blockIdx.z;
    const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * BN;
    const float* p = a.P.p[g];
    const size_t grow = (size_t)g * a.Tg;
    float acc[4][TN];
    gemm_tile_64<BN, false>(a.X + grow * a.K, a.K, a.Tg, p + a.w_off, a.K, a.K, m0, n0, acc, sm);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = __ldg(p + a.b_off + n0 + gemm_col(tx, j));
    if (EPI != EPI_RES_LN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = m0 + ty * 4 + i;
            if (r >= a.Tg) continue;
            float* y = a.Y + (grow + r) * a.N + n0;
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q] = acc[i][j4 * 4 + q] + bias[j4 * 4 + q];
                    if (EPI == EPI_BIAS_RELU) v[q] = fmaxf(v[q], 0.f);
                }
                *reinterpret_cast<float4*>(y + j4 * 64 + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    } else {
        // x_out = LayerNorm(x_res + relu(acc + b)) over the full row (BN == N); the 16 lanes of a half-warp share a row
        float gam[TN], bet[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            gam[j] = __ldg(p + a.gamma_off + gemm_col(tx, j));
            bet[j] = __ldg(p + a.beta_off + gemm_col(tx, j));
        }
        const float inv_n = 1.f / (float)BN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = m0 + ty * 4 + i;
            const bool ok = r < a.Tg;                         // all 16 lanes of the row agree
            const size_t ro = (grow + (ok ? r : 0)) * (size_t)BN;
            float u[TN], rl[TN];
            float s = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 xr = *reinterpret_cast<const float4*>(a.R + ro + j4 * 64 + tx * 4);
                const float xv[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j4 * 4 + q;
                    rl[j] = fmaxf(acc[i][j] + bias[j], 0.f);
                    u[j] = xv[q] + rl[j];
                    s += u[j];
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * inv_n;
            float vs = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) { const float dlt = u[j] - mean; vs = fmaf(dlt, dlt, vs); }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
            const float rstd = 1.0f / sqrtf(vs * inv_n + 1e-5f);
            if (!ok) continue;
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const int j = j4 * 4 + q; v[q] = (u[j] - mean) * rstd * gam[j] + bet[j]; }
                *reinterpret_cast<float4*>(a.Y + ro + j4 * 64 + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
                if (a.r_save)
                    *reinterpret_cast<float4*>(a.r_save + ro + j4 * 64 + tx * 4) =
                        make_float4(rl[j4 * 4], rl[j4 * 4 + 1], rl[j4 * 4 + 2], rl[j4 * 4 + 3]);
            }
            if (a.st_save && tx == 0) { a.st_save[(grow + r) * 2] = mean; a.st_save[(grow + r) * 2 + 1] = rstd; }
        }
    }
}

// ---- causal self-attention core ------------------------------------------------------------------------------------------
// qkv [T, 3d]: q at column h*hd, k at d + h*hd, v at 2d + h*hd (packed in_proj layout).  o [T, d].
template <int HD>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o, int L, int d, float scale) {
    __shared__ float Ks[128][HD + 1];
    __shared__ float Vs[128][HD + 1];
    const int h = blockIdx.x;
    const size_t t0 = (size_t)blockIdx.y * L;
    const int tid = threadIdx.x;
    // cooperative load of K and V head slices: L rows x HD floats each
    for (int e = tid; e < L * (HD / 4); e += blockDim.x) {
        const int r = e / (HD / 4), c4 = (e % (HD / 4)) * 4;
        const float* base = qkv + (t0 + r) * (size_t)(3 * d) + h * HD + c4;
        const float4 kv = *reinterpret_cast<const float4*>(base + d);
        const float4 vv = *reinterpret_cast<const float4*>(base + 2 * d);
        Ks[r][c4] = kv.x; Ks[r][c4 + 1] = kv.y; Ks[r][c4 + 2] = kv.z; Ks[r][c4 + 3] = kv.w;
        Vs[r][c4] = vv.x; Vs[r][c4 + 1] = vv.y; Vs[r][c4 + 2] = vv.z; Vs[r][c4 + 3] = vv.w;
    }
    __syncthreads();
    const int j = tid;
    if (j >= L) return;
    float q[HD], accv[HD];
    const float* qp = qkv + (t0 + j) * (size_t)(3 * d) + h * HD;
#pragma unroll
    for (int c4 = 0; c4 < HD; c4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(qp + c4);
        q[c4] = v.x * scale; q[c4 + 1] = v.y * scale; q[c4 + 2] = v.z * scale; q[c4 + 3] = v.w * scale;
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) accv[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i <= j; ++i) {                           // causal: keys 0..j (mask is -inf above the diagonal)
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) s = fmaf(q[c], Ks[i][c], s);
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);                     // exp(-inf) = 0 on the first key
        const float pw = expf(s - mn);
        l = l * corr + pw;
#pragma unroll
        for (int c = 0; c < HD; ++c) accv[c] = fmaf(pw, Vs[i][c], accv[c] * corr);
        m = mn;
    }
    const float inv = 1.f / l;
    float* op = o + (t0 + j) * (size_t)d + h * HD;
#pragma unroll
    for (int c4 = 0; c4 < HD; c4 += 4)
        *reinterpret_cast<float4*>(op + c4) = make_float4(accv[c4] * inv, accv[c4 + 1] * inv, accv[c4 + 2] * inv, accv[c4 + 3] * inv);
}



// ---- causal self-attention, one CTA per sequence (all heads) ----------------------------------------------------------------
// K and V of the whole sequence are staged once in shared memory (coalesced loads, rows padded by 4 floats); warp h = head h.
// Lane r owns the query rows a = r and b = L-1-r (balanced causal work); keys are visited in warp-uniform blocks of 4 so
// that every K / V read is a broadcast float4 (1 shared-memory wavefront) shared by both rows, one online-softmax rescale
// per block, base-2 exponentials on the MUFU.  (A back-to-back per-lane key schedule has fewer instructions but makes the
// reads lane-divergent -- 4 wavefronts per LDS.128 -- and measured slower: shared-memory bandwidth bound.)
__device__ __forceinline__ float ex2_approx(float x) {        // MUFU.EX2: x <= 0 here, so no overflow / denormal handling needed
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_seq_kernel(const float* __restrict__ qkv, float* __restrict__ o, int L, int d, float scale) {
    extern __shared__ float sm_kv[];
    const int ld = d + 4;
    const int Lp = L + 4;                                     // 4 zero rows of padding: a key block may run past the last key
    float* Ks = sm_kv;
    float* Vs = sm_kv + (size_t)Lp * ld;
    const size_t t0 = (size_t)blockIdx.x * L;
    const int tid = threadIdx.x, dq = d / 4;
    for (int e = tid; e < Lp * dq; e += blockDim.x) {
        const int r = e / dq, c4 = (e % dq) * 4;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (r < L) {
            const float* base = qkv + (t0 + r) * (size_t)(3 * d) + c4;
            kv = *reinterpret_cast<const float4*>(base + d);
            vv = *reinterpret_cast<const float4*>(base + 2 * d);
        }
        *reinterpret_cast<float4*>(Ks + (size_t)r * ld + c4) = kv;
        *reinterpret_cast<float4*>(Vs + (size_t)r * ld + c4) = vv;
    }
    __syncthreads();
    const int h = tid >> 5, lane = tid & 31;
    const int half = (L + 1) / 2;
    scale *= 1.4426950408889634f;                             // softmax in base 2: exp(x) = exp2(x * log2 e)
    for (int r0 = 0; r0 < half; r0 += 32) {                  // L <= 64: one pass; L = 128: two
        // Lane r owns rows a = r and b = L-1-r.  Keys are visited in warp-UNIFORM blocks of 4, so every K / V read is a
        // broadcast (one shared-memory wavefront per LDS.128) shared by both rows; keys above a row's diagonal are masked.
        const int r = r0 + lane;
        const bool act = r < half;
        const int ra = act ? r : -1, rb = act ? L - 1 - r : -1;   // -1: every key masked for idle lanes
        float qa[HD], qb[HD], aa[HD], ab[HD];
        {
            const float* pa = qkv + (t0 + max(ra, 0)) * (size_t)(3 * d) + h * HD;
            const float* pb = qkv + (t0 + max(rb, 0)) * (size_t)(3 * d) + h * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 4) {
                const float4 va = *reinterpret_cast<const float4*>(pa + c), vb = *reinterpret_cast<const float4*>(pb + c);
                qa[c] = va.x * scale; qa[c + 1] = va.y * scale; qa[c + 2] = va.z * scale; qa[c + 3] = va.w * scale;
                qb[c] = vb.x * scale; qb[c + 1] = vb.y * scale; qb[c + 2] = vb.z * scale; qb[c + 3] = vb.w * scale;
            }
        }
#pragma unroll
        for (int c = 0; c < HD; ++c) { aa[c] = 0.f; ab[c] = 0.f; }
        float ma = -INFINITY, la_ = 0.f, mb = -INFINITY, lb = 0.f;
        const int last_a = min(half, r0 + 32) - 1;            // largest "a" row of this warp pass (warp-uniform)
        const float* kp = Ks + h * HD;
        const float* vp = Vs + h * HD;
        int rem_a = ra + 1, rem_b = rb + 1;                   // keys of each row not yet visited
        for (int j0 = 0; j0 < L; j0 += 4, kp += 4 * ld, vp += 4 * ld, rem_a -= 4, rem_b -= 4) {
            const bool do_a = j0 <= last_a;                   // warp-uniform: the upper half of the keys only feeds rows b
            float sa[4], sb[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float da = 0.f, db = 0.f;
#pragma unroll
                for (int c = 0; c < HD; c += 4) {
                    const float4 kv = *reinterpret_cast<const float4*>(kp + jj * ld + c);
                    db = fmaf(qb[c], kv.x, db); db = fmaf(qb[c + 1], kv.y, db); db = fmaf(qb[c + 2], kv.z, db); db = fmaf(qb[c + 3], kv.w, db);
                    if (do_a) { da = fmaf(qa[c], kv.x, da); da = fmaf(qa[c + 1], kv.y, da); da = fmaf(qa[c + 2], kv.z, da); da = fmaf(qa[c + 3], kv.w, da); }
                }
                sa[jj] = (jj < rem_a) ? da : -INFINITY;       // causal mask (additive -inf above the diagonal)
                sb[jj] = (jj < rem_b) ? db : -INFINITY;
            }
            const float nmb = fmaxf(mb, fmaxf(fmaxf(sb[0], sb[1]), fmaxf(sb[2], sb[3])));
            const float sfb = (nmb == -INFINITY) ? 0.f : nmb;
            const float cb = ex2_approx(mb - sfb);
            float pb_[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) pb_[jj] = ex2_approx(sb[jj] - sfb);
            lb = lb * cb + (pb_[0] + pb_[1]) + (pb_[2] + pb_[3]);
            mb = nmb;
            float ca = 1.f, pa_[4] = {0.f, 0.f, 0.f, 0.f};
            if (do_a) {
                const float nma = fmaxf(ma, fmaxf(fmaxf(sa[0], sa[1]), fmaxf(sa[2], sa[3])));
                const float sfa = (nma == -INFINITY) ? 0.f : nma;
                ca = ex2_approx(ma - sfa);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) pa_[jj] = ex2_approx(sa[jj] - sfa);
                la_ = la_ * ca + (pa_[0] + pa_[1]) + (pa_[2] + pa_[3]);
                ma = nma;
            }
#pragma unroll
            for (int c = 0; c < HD; ++c) { ab[c] *= cb; if (do_a) aa[c] *= ca; }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int c = 0; c < HD; c += 4) {
                    const float4 vv = *reinterpret_cast<const float4*>(vp + jj * ld + c);
                    ab[c] = fmaf(pb_[jj], vv.x, ab[c]); ab[c + 1] = fmaf(pb_[jj], vv.y, ab[c + 1]);
                    ab[c + 2] = fmaf(pb_[jj], vv.z, ab[c + 2]); ab[c + 3] = fmaf(pb_[jj], vv.w, ab[c + 3]);
                    if (do_a) {
                        aa[c] = fmaf(pa_[jj], vv.x, aa[c]); aa[c + 1] = fmaf(pa_[jj], vv.y, aa[c + 1]);
                        aa[c + 2] = fmaf(pa_[jj], vv.z, aa[c + 2]); aa[c + 3] = fmaf(pa_[jj], vv.w, aa[c + 3]);
                    }
                }
            }
        }
        if (act) {
            const float ib = 1.f / lb;
            float* ob = o + (t0 + rb) * (size_t)d + h * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 4)
                *reinterpret_cast<float4*>(ob + c) = make_float4(ab[c] * ib, ab[c + 1] * ib, ab[c + 2] * ib, ab[c + 3] * ib);
            if (ra != rb) {
                const float ia = 1.f / la_;
                float* oa = o + (t0 + ra) * (size_t)d + h * HD;
#pragma unroll
                for (int c = 0; c < HD; c += 4)
                    *reinterpret_cast<float4*>(oa + c) = make_float4(aa[c] * ia, aa[c + 1] * ia, aa[c + 2] * ia, aa[c + 3] * ia);
            }
        }
    }
}

// ---- causal self-attention on the warp-level tensor cores (mma.sync TF32 x3), one CTA per sequence, warp = head ----------------
// d_model = 64, 8 heads of 8, L <= 64.  The sequence's packed qkv rows are staged once in shared memory (coalesced float4 loads,
// q pre-scaled for the base-2 softmax); every warp runs att_head() for its head and parks the normalised outputs in the q
// columns it has just consumed, so the result leaves as full 256-byte rows.
__global__ void __launch_bounds__(256)
attn_mma_kernel(const float* __restrict__ qkv, float* __restrict__ o, int L, float scale) {
    extern __shared__ float sm_kv[];                           // [64][ATT_LD]
    const size_t t0 = (size_t)blockIdx.x * L;
    const int tid = threadIdx.x;
    const float qs = scale * 1.4426950408889634f;
    for (int e = tid; e < 64 * 48; e += 256) {
        const int r = e / 48, c4 = (e % 48) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < L) {
            v = __ldg(reinterpret_cast<const float4*>(qkv + (t0 + r) * 192 + c4));
            if (c4 < 64) { v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs; }
        }
        *reinterpret_cast<float4*>(sm_kv + r * ATT_LD + c4) = v;
    }
    __syncthreads();
    const int h = tid >> 5, lane = tid & 31;
    float* sq = sm_kv;
    att_head(sq, h, L, lane, [&](int row, int col, float v0, float v1) {
        *reinterpret_cast<float2*>(sq + row * ATT_LD + col) = make_float2(v0, v1);
    });
    __syncthreads();
    for (int e = tid; e < L * 16; e += 256) {
        const int r = e >> 4, c4 = (e & 15) * 4;
        *reinterpret_cast<float4*>(o + (t0 + r) * 64 + c4) = *reinterpret_cast<const float4*>(sm_kv + r * ATT_LD + c4);
    }
}

// The same for d_model = 128 / 8 heads (head_dim 16, the Memory-5 network): q|k|v rows of 384 floats staged with a stride of
// 388 (== 4 mod 32: conflict-free fragment loads), two k-steps per Q K^T block and two 8-column P V tiles per head.
constexpr int ATT_LD128 = 3 * 128 + 4;
__global__ void __launch_bounds__(256)
attn_mma128_kernel(const float* __restrict__ qkv, float* __restrict__ o, int L, float scale) {
    extern __shared__ float sm_kv[];                           // [64][ATT_LD128]
    const size_t t0 = (size_t)blockIdx.x * L;
    const int tid = threadIdx.x;
    const float qs = scale * 1.4426950408889634f;
    for (int e = tid; e < 64 * 96; e += 256) {
        const int r = e / 96, c4 = (e % 96) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < L) {
            v = __ldg(reinterpret_cast<const float4*>(qkv + (t0 + r) * 384 + c4));
            if (c4 < 128) { v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs; }
        }
        *reinterpret_cast<float4*>(sm_kv + r * ATT_LD128 + c4) = v;
    }
    __syncthreads();
    const int h = tid >> 5, lane = tid & 31;
    float* sq = sm_kv;
    att_head_g<2, 128, ATT_LD128>(sq, h, L, lane, [&](int row, int col, float v0, float v1) {
        *reinterpret_cast<float2*>(sq + row * ATT_LD128 + col) = make_float2(v0, v1);   // parked in the consumed q columns
    });
    __syncthreads();
    for (int e = tid; e < L * 32; e += 256) {
        const int r = e >> 5, c4 = (e & 31) * 4;
        *reinterpret_cast<float4*>(o + (t0 + r) * 128 + c4) = *reinterpret_cast<const float4*>(sm_kv + r * ATT_LD128 + c4);
    }
}

// ---- acting: attention of the LAST valid query only (the policy reads q[:, -1, :], agents/dtqn.py:107) ------------------
// ql [G*n_seq, d] = scaled-later query of the last valid token; kv [T, 2d] = (k | v) of every token.  One warp per
// (sequence, head): lanes stride over the n_i keys, warp-shuffle softmax, then HD warp reductions for P V.
template <int HD>
__global__ void __launch_bounds__(256)
attn_last_kernel(const float* __restrict__ ql, const float* __restrict__ kv, GroupSrc S, int n_seq, int L, int d,
                 float scale, float* __restrict__ ol) {
    extern __shared__ float sm_kv[];                           // [L][2d + 4]: (k | v) rows, padded -> conflict-free float4
    const int g = blockIdx.y, i = blockIdx.x;
    const int tid = threadIdx.x, h = tid >> 5, lane = tid & 31;
    const dtqn_obs_src& s = S.s[g];
    int n = L;
    if (s.timestep) n = min(min(s.ring_len, s.timestep[i] + 1), L);
    const size_t seq = (size_t)g * n_seq + i;
    const int ld = 2 * d + 4, rq = (2 * d) / 4;
    const float* kvb = kv + seq * L * (size_t)(2 * d);
    for (int e = tid; e < n * rq; e += blockDim.x) {
        const int r = e / rq, c4 = (e % rq) * 4;
        *reinterpret_cast<float4*>(sm_kv + (size_t)r * ld + c4) = *reinterpret_cast<const float4*>(kvb + (size_t)r * (2 * d) + c4);
    }
    __syncthreads();
    const float* qp = ql + seq * d + h * HD;
    float q[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) q[c] = qp[c] * scale;
    float sc[4];                                               // up to 128 keys: 4 per lane
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        sc[r] = -INFINITY;
        if (j < n) {
            const float* kp = sm_kv + (size_t)j * ld + h * HD;
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < HD; c += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kp + c);
                a = fmaf(q[c], k4.x, a); a = fmaf(q[c + 1], k4.y, a); a = fmaf(q[c + 2], k4.z, a); a = fmaf(q[c + 3], k4.w, a);
            }
            sc[r] = a;
        }
        m = fmaxf(m, sc[r]);
    }
    m = warp_max(m);
    float l = 0.f, acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        if (j < n) {
            const float pw = expf(sc[r] - m);
            l += pw;
            const float* vp = sm_kv + (size_t)j * ld + d + h * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(vp + c);
                acc[c] = fmaf(pw, v4.x, acc[c]); acc[c + 1] = fmaf(pw, v4.y, acc[c + 1]);
                acc[c + 2] = fmaf(pw, v4.z, acc[c + 2]); acc[c + 3] = fmaf(pw, v4.w, acc[c + 3]);
            }
        }
    }
    l = warp_sum(l);
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
        const float inv = 1.f / l;
        float* op = ol + seq * d + h * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) op[c] = acc[c] * inv;
    }
}

// ---- rows of the last valid position of every sequence (acting: q[:, -1, :], agents/dtqn.py:107) ---------------------
__global__ void gather_last_kernel(const float* __restrict__ x, GroupSrc S, int n_seq, int L, int d, float* __restrict__ out) {
    const int g = blockIdx.z;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_seq * d) return;
    const int i = (int)(idx / d), cch = (int)(idx % d);
    int last = L - 1;
    const dtqn_obs_src& s = S.s[g];
    if (s.timestep) { const int n = min(min(s.ring_len, s.timestep[i] + 1), L); last = n - 1; }
    out[((long long)g * n_seq + i) * d + cch] = x[(((long long)g * n_seq + i) * L + last) * d + cch];
}

// ---- final Linear(d, A): one warp per token --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ hh, GroupPtrs P, long long w_off, long long b_off, long long Tg, int d, int A,
            float* __restrict__ q) {
    const int g = blockIdx.z;
    const long long t = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (t >= Tg) return;
    const int lane = threadIdx.x & 31;
    const float* p = P.p[g];
    const float* x = hh + ((long long)g * Tg + t) * d;
    float xv[4];
    const int per = d / 32;                                   // 2 (d = 64) or 4 (d = 128)
    for (int k = 0; k < per; ++k) xv[k] = x[lane + 32 * k];
    for (int a = 0; a < A; ++a) {
        float s = 0.f;
        for (int k = 0; k < per; ++k) s = fmaf(xv[k], __ldg(p + w_off + (long long)a * d + lane + 32 * k), s);
        s = warp_sum(s);
        if (lane == 0) q[((long long)g * Tg + t) * A + a] = s + __ldg(p + b_off + a);
    }
}

template <int EPI>
int launch_linear(const LinArgs& a, int G, int d_model, cudaStream_t st) {
    dim3 grid(dtqn_cdiv(a.Tg, GEMM_BM), 1, G);
    prof_begin(PROF_LINEAR, st);
    if (EPI == EPI_RES_LN) {
        if (a.N == 64) linear_kernel<64, EPI_RES_LN><<<grid, GEMM_THREADS, 0, st>>>(a);
        else if (a.N == 128) linear_kernel<128, EPI_RES_LN><<<grid, GEMM_THREADS, 0, st>>>(a);
        else return DTQN_E_UNSUPPORTED;
    } else {
        if (a.N % 128 == 0) { grid.y = a.N / 128; linear_kernel<128, EPI><<<grid, GEMM_THREADS, 0, st>>>(a); }
        else if (a.N % 64 == 0) { grid.y = a.N / 64; linear_kernel<64, EPI><<<grid, GEMM_THREADS, 0, st>>>(a); }
        else return DTQN_E_UNSUPPORTED;
    }
    prof_end(PROF_LINEAR, st, 2.0 * (double)a.Tg * G * a.N * a.K);
    DTQN_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// ---- exported --------------------------------------------------------------------------------------------------------------
extern "C" int64_t dtqn_net_param_count(const dtqn_net_cfg* cfg) {
    if (!cfg) return DTQN_E_ARG;
    NetLayout L;
    int rc = net_layout(*cfg, L);
    return rc ? rc : L.total;
}

extern "C" int dtqn_net_param_offsets(const dtqn_net_cfg* cfg, int64_t* out, int32_t max_entries) {
    if (!cfg || !out) return DTQN_E_ARG;
    NetLayout L;
    int rc = net_layout(*cfg, L);
    if (rc) return rc;
    int n = 0;
    auto put = [&](long long v) { if (n < max_entries) out[n] = v; ++n; };
    if (cfg->discrete) put(L.emb_table);
    put(L.emb_w); put(L.emb_b);
    if (cfg->action_dim > 0) put(L.act_table);
    put(L.pos);
    for (int i = 0; i < cfg->n_layers; ++i) {
        const LayerOff& l = L.layer[i];
        put(l.ln1_w); put(l.ln1_b); put(l.ln2_w); put(l.ln2_b); put(l.in_w); put(l.in_b); put(l.out_w); put(l.out_b);
        put(l.f1_w); put(l.f1_b); put(l.f2_w); put(l.f2_b);
        if (i == 0 && cfg->gate_gru)
            for (int k = 0; k < 2; ++k) {
                const GateOff& q = L.gate[k];
                put(q.w_r); put(q.u_r); put(q.w_z); put(q.b_z); put(q.u_z); put(q.w_g); put(q.u_g);
            }
    }
    put(L.h1_w); put(L.h1_b); put(L.h2_w); put(L.h2_b);
    return n > max_entries ? DTQN_E_ARG : n;
}

extern "C" int64_t dtqn_net_workspace_floats(const dtqn_net_cfg* cfg, int64_t n_tokens, int32_t save) {
    if (!cfg || n_tokens <= 0) return DTQN_E_ARG;
    if (cfg_is_variant(*cfg)) return var_workspace_floats(*cfg, n_tokens);
    NetAct A;
    return net_act_layout(*cfg, n_tokens, save, nullptr, A);
}

// tokens per group from which the tcgen05 path (bf16x3 split, 128-row tiles) replaces the fp32 CUDA-core GEMMs.
// The 32 x 50-token training groups stay on fp32 CUDA cores: they are latency-bound either way (measured: no faster on
// tensor cores) and the exact-fp32 forward keeps ReLU masks -- hence gradients -- closest to the reference's.
static int g_tc_min_tokens = 4096;
static int g_seq_fused = 1;
static int g_attn_mma = 1;        // mma.sync TF32x3 attention core for d = 64 / 8 heads / L <= 64 (0: fp32 CUDA-core kernel)
extern "C" int dtqn_set_attn_mma(int32_t on) { g_attn_mma = on; return 0; }
static int g_embed_disc_fast = 1;  // shared-memory staged Embedding -> Flatten -> Linear (0: generic embed_kernel)
extern "C" int dtqn_set_embed_disc_fast(int32_t on) { g_embed_disc_fast = on; return 0; }
static int g_tc_fuse_embed = 0;   // measured slower (dependent timestep -> obs -> pos loads stall the producers): off by default
extern "C" int dtqn_set_tc_fuse_embed(int32_t on) { g_tc_fuse_embed = on; return 0; }
extern "C" int dtqn_set_seq_fused(int32_t on) { g_seq_fused = on; return 0; }
extern "C" int dtqn_set_tc_min_tokens(int32_t n) { g_tc_min_tokens = n; return 0; }

extern "C" int dtqn_forward(const dtqn_net_cfg* cfg, int32_t G, const float* const* params, const void* const* packed,
                            const dtqn_obs_src* src, int32_t n_seq, int32_t L, int32_t q_mode, int32_t save, float* ws,
                            int64_t ws_floats, float* q_out, void* stream) {
    if (!cfg || !params || !src || !ws || !q_out || G < 1 || G > DTQN_MAX_GROUPS || n_seq < 1 || L < 1) return DTQN_E_ARG;
    if (L > cfg->context_len) return DTQN_E_ARG;               // dtqn.py:171-173 assert
    NetLayout lay;
    int rc = net_layout(*cfg, lay);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg_is_variant(*cfg)) {                                // ablation flags: general fp32 kernel-per-op path
        for (int g = 0; g < G; ++g) {
            if (!params[g] || !src[g].obs) return DTQN_E_ARG;
            if (src[g].timestep && src[g].ring_len < L) return DTQN_E_ARG;
        }
        return var_forward(*cfg, lay, G, params, src, n_seq, L, q_mode, save, ws, ws_floats, q_out, st);
    }
    const long long Tg = (long long)n_seq * L, T = Tg * G;
    NetAct act;
    if (net_act_layout(*cfg, T, save, ws, act) > ws_floats) return DTQN_E_ARG;
    const int d = cfg->d_model, H = cfg->n_heads, hd = d / H;
    GroupPtrs P{}; GroupSrc S{};
    const uint8_t* pk[DTQN_MAX_GROUPS] = {nullptr, nullptr, nullptr};
    bool use_tc = packed != nullptr && Tg >= g_tc_min_tokens;
    for (int g = 0; g < G; ++g) {
        if (!params[g] || !src[g].obs) return DTQN_E_ARG;
        P.p[g] = params[g]; S.s[g] = src[g];
        if (src[g].timestep && src[g].ring_len < L) return DTQN_E_ARG;
        if (packed) { pk[g] = (const uint8_t*)packed[g]; if (!pk[g]) use_tc = false; }
    }
    TcPackTable tab{};
    if (use_tc) tc_pack_table(*cfg, lay, tab);
    // acting forward on the tensor-core path with continuous observations: the token embedding x0 is recomputed inside
    // the in_proj producer and the out_proj LayerNorm epilogue of layer 0 and never written to HBM
    const bool fuse_embed = use_tc && g_tc_fuse_embed && tc_pipelined_enabled() && q_mode == 1 && !cfg->discrete && cfg->obs_dim <= 4 && d == 64 &&
                            cfg->n_layers >= 2 && L >= 3 && H * 32 <= 256;
    TcEmbed emb{};
    if (fuse_embed) {
        emb.L = L; emb.O = cfg->obs_dim; emb.obs_mask = src[0].obs_mask; emb.w_off = lay.emb_w; emb.b_off = lay.emb_b; emb.pos_off = lay.pos;
        for (int g = 0; g < G; ++g) emb.src[g] = src[g];
    }
    auto linear = [&](const LinArgs& a, int epi, int tab_idx, int emb_mode = 0) -> int {
        if (use_tc) {
            if (emb_mode) { emb.mode = emb_mode; return launch_linear_tc(a, epi, G, pk, tab.e[tab_idx].pk_off, st, &emb); }
            return launch_linear_tc(a, epi, G, pk, tab.e[tab_idx].pk_off, st);
        }
        if (epi == EPI_BIAS) return launch_linear<EPI_BIAS>(a, G, d, st);
        if (epi == EPI_BIAS_RELU) return launch_linear<EPI_BIAS_RELU>(a, G, d, st);
        return launch_linear<EPI_RES_LN>(a, G, d, st);
    };
    // acting forward of the default 2-layer / d_model 64 network: ONE tcgen05 kernel from the context ring to the final
    // layer's (residual row, attention row) per sequence -- no embedding, q|k|v, attention or FFN activations in HBM
    const bool act_fused = use_tc && q_mode == 1 && !save && tab.act_img_off >= 0 && act_fused_supported(*cfg, L);
    if (!fuse_embed && !act_fused) {
        prof_begin(PROF_EMBED, st);
        if (!cfg->discrete && cfg->obs_dim <= 16) {
            if (d == 64) embed_cont_kernel<64><<<dim3(dtqn_cdiv(Tg, 64), 1, G), 256, 0, st>>>(P, S, cfg->obs_dim, lay.emb_w, lay.emb_b, lay.pos, n_seq, L, src[0].obs_mask, act.x0);
            else         embed_cont_kernel<128><<<dim3(dtqn_cdiv(Tg, 32), 1, G), 256, 0, st>>>(P, S, cfg->obs_dim, lay.emb_w, lay.emb_b, lay.pos, n_seq, L, src[0].obs_mask, act.x0);
        } else if (cfg->discrete && g_embed_disc_fast == 1 && (d == 64 || d == 128) &&
                   sizeof(float) * ((size_t)cfg->obs_dim * cfg->vocab * d + d) <= 96 * 1024) {
            const size_t smem = sizeof(float) * ((size_t)cfg->obs_dim * cfg->vocab * d + d);
            const int n_lut = cfg->obs_dim * cfg->vocab * d;
            embed_lut_build_kernel<<<dim3(dtqn_cdiv(n_lut, 256), G), 256, 0, st>>>(P, cfg->obs_dim, cfg->embed_per_obs, cfg->vocab, d,
                                                                                  lay.emb_table, lay.emb_w, act.lut);
            DTQN_LAUNCH_CHECK();
            long long passes = dtqn_cdiv(Tg, 8 * (128 / d));
            dim3 grid((unsigned)(passes < 2 * 148 ? passes : 2 * 148), 1, G);   // persistent: the table is staged once per CTA
            const float mask = (float)(cfg->vocab - 1);
            cudaError_t ae;
            if (d == 64) {
                ae = cudaFuncSetAttribute(embed_lut_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (ae == cudaSuccess)
                    embed_lut_kernel<64><<<grid, 256, smem, st>>>(P, S, cfg->obs_dim, cfg->vocab, act.lut, lay.emb_b, lay.pos, n_seq, L, mask, act.x0);
            } else {
                ae = cudaFuncSetAttribute(embed_lut_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (ae == cudaSuccess)
                    embed_lut_kernel<128><<<grid, 256, smem, st>>>(P, S, cfg->obs_dim, cfg->vocab, act.lut, lay.emb_b, lay.pos, n_seq, L, mask, act.x0);
            }
            if (ae != cudaSuccess) return (int)ae;
        } else if (cfg->discrete && g_embed_disc_fast == 2 && lay.k_in <= 128) {   // previous form (kept for A/B timing: dtqn_set_embed_disc_fast(2))
            const int KI = lay.k_in;
            const size_t smem = sizeof(float) * ((size_t)KI * (d + 4) + d + ((cfg->vocab * cfg->embed_per_obs + 3) & ~3) + (size_t)ED_TOK * KI);
            long long chunks = dtqn_cdiv(Tg, ED_TOK);
            dim3 grid((unsigned)(chunks < 3 * 148 ? chunks : 3 * 148), 1, G);   // persistent: <= 3 CTAs per SM, W staged once per CTA
            const float mask = (float)(cfg->vocab - 1);
            cudaError_t ae = cudaSuccess;
            if (d == 64) {
                ae = cudaFuncSetAttribute(embed_disc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (ae == cudaSuccess)
                    embed_disc_kernel<64><<<grid, 256, smem, st>>>(P, S, cfg->obs_dim, cfg->embed_per_obs, cfg->vocab, lay.emb_table, lay.emb_w, lay.emb_b, lay.pos, n_seq, L, mask, act.x0);
            } else {
                ae = cudaFuncSetAttribute(embed_disc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (ae == cudaSuccess)
                    embed_disc_kernel<128><<<grid, 256, smem, st>>>(P, S, cfg->obs_dim, cfg->embed_per_obs, cfg->vocab, lay.emb_table, lay.emb_w, lay.emb_b, lay.pos, n_seq, L, mask, act.x0);
            }
            if (ae != cudaSuccess) return (int)ae;
        } else {
            dim3 grid(dtqn_cdiv(Tg * (d / 4), 256), 1, G);
            embed_kernel<<<grid, 256, 0, st>>>(P, S, *cfg, lay.emb_table, lay.emb_w, lay.emb_b, lay.pos, n_seq, L,
                                                src[0].obs_mask, act.x0);
        }
        prof_end(PROF_EMBED, st, 2.0 * (double)T * lay.k_in * d);
        DTQN_LAUNCH_CHECK();
    }
    if (!use_tc && q_mode == 0 && g_seq_fused && seq_forward_supported(*cfg, L)) {     // all layers + head, one launch
        const float* wt[DTQN_MAX_GROUPS] = {nullptr, nullptr, nullptr};
        bool have_wt = packed != nullptr;
        TcPackTable tb{};
        if (have_wt) { tc_pack_table(*cfg, lay, tb); have_wt = tb.wt_off >= 0; }
        for (int g = 0; g < G && have_wt; ++g) {
            if (!pk[g]) have_wt = false;
            else wt[g] = reinterpret_cast<const float*>(pk[g] + tb.wt_off);
        }
        return launch_seq_forward(*cfg, lay, act, P, G, n_seq, L, save, q_out, st, have_wt ? wt : nullptr);
    }
    const float* x_in = act.x0;
    // acting (q_mode 1) needs the final layer's output at ONE position per sequence: that layer only projects K/V for
    // every token; its query, attention row, out_proj, FFN and LayerNorms run on n_seq rows (exactly the same values).
    const bool last_only = q_mode == 1 && L >= 3 && H * 32 <= 256;
    const int n_full = act_fused ? 0 : (last_only ? cfg->n_layers - 1 : cfg->n_layers);
    for (int li = 0; li < n_full; ++li) {
        const LayerOff& lo = lay.layer[li];
        const LayerAct& la = act.layer[li];
        LinArgs a{};
        a.P = P; a.Tg = (int)Tg;
        // in_proj
        a.X = x_in; a.Y = la.qkv; a.w_off = lo.in_w; a.b_off = lo.in_b; a.N = 3 * d; a.K = d;
        if ((rc = linear(a, EPI_BIAS, 4 * li + TC_W_IN, (fuse_embed && li == 0) ? 1 : 0))) return rc;
        // attention core
        {
            const float scale = 1.0f / sqrtf((float)hd);
            const size_t smem = sizeof(float) * 2 * (size_t)(L + 4) * (d + 4);
            const bool seq_kernel = H * 32 <= 256 && (hd == 8 || hd == 16);
            prof_begin(PROF_ATTN_FWD, st);
            if (g_attn_mma && seq_kernel && hd == 8 && d == 64 && H == 8 && L <= 64) {
                const size_t sm_mma = sizeof(float) * 64 * ATT_LD;
                static bool attr_set = false;
                if (!attr_set) {
                    cudaFuncSetAttribute(attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mma);
                    attr_set = true;
                }
                attn_mma_kernel<<<(unsigned)(n_seq * G), 256, sm_mma, st>>>(la.qkv, la.o, L, scale);
            } else if (g_attn_mma && !save && seq_kernel && hd == 16 && d == 128 && H == 8 && L <= 64) {   // inference groups only: the
                // training forward keeps the exact-fp32 kernel (ReLU masks / gradients closest to the reference's)
                const size_t sm_mma = sizeof(float) * 64 * ATT_LD128;
                static bool attr_set128 = false;
                if (!attr_set128) {
                    cudaFuncSetAttribute(attn_mma128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mma);
                    attr_set128 = true;
                }
                attn_mma128_kernel<<<(unsigned)(n_seq * G), 256, sm_mma, st>>>(la.qkv, la.o, L, scale);
            } else if (seq_kernel) {
                if (hd == 8) {
                    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_seq_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    attn_seq_kernel<8><<<(unsigned)(n_seq * G), 32 * H, smem, st>>>(la.qkv, la.o, L, d, scale);
                } else {
                    if (smem > 48 * 1024) cudaFuncSetAttribute(attn_seq_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    attn_seq_kernel<16><<<(unsigned)(n_seq * G), 32 * H, smem, st>>>(la.qkv, la.o, L, d, scale);
                }
            } else {
                dim3 grid(H, (unsigned)(n_seq * G));
                const int thr = L <= 64 ? 64 : 128;
                if (hd == 8) attn_fwd_kernel<8><<<grid, thr, 0, st>>>(la.qkv, la.o, L, d, scale);
                else if (hd == 16) attn_fwd_kernel<16><<<grid, thr, 0, st>>>(la.qkv, la.o, L, d, scale);
                else if (hd == 32) attn_fwd_kernel<32><<<grid, thr, 0, st>>>(la.qkv, la.o, L, d, scale);
                else if (hd == 4) attn_fwd_kernel<4><<<grid, thr, 0, st>>>(la.qkv, la.o, L, d, scale);
                else return DTQN_E_UNSUPPORTED;
            }
            prof_end(PROF_ATTN_FWD, st, 4.0 * (double)T * L * d);      // dense L x L count, as the reference computes it
            DTQN_LAUNCH_CHECK();
        }
        // out_proj -> relu -> +x -> LN1
        a.X = la.o; a.Y = la.x1; a.w_off = lo.out_w; a.b_off = lo.out_b; a.N = d; a.K = d;
        a.R = x_in; a.gamma_off = lo.ln1_w; a.beta_off = lo.ln1_b;
        a.r_save = save ? la.r1 : nullptr; a.st_save = save ? la.st1 : nullptr;
        if ((rc = linear(a, EPI_RES_LN, 4 * li + TC_W_OUT, (fuse_embed && li == 0) ? 2 : 0))) return rc;
        if (use_tc && !save && d == 64 && tc_ffn_fused_enabled()) {
            // ffn.0 -> relu -> ffn.2 -> relu -> +x1 -> LN2 in one tcgen05 kernel; the hidden activations stay on the SM
            if ((rc = launch_ffn_tc(la.x1, la.x2, P, G, pk, tab.e[4 * li + TC_W_F1].pk_off, tab.e[4 * li + TC_W_F2].pk_off,
                                    lo.f1_b, lo.f2_b, lo.ln2_w, lo.ln2_b, (int)Tg, st))) return rc;
            x_in = la.x2;
            continue;
        }
        // ffn.0 + relu
        a.X = la.x1; a.Y = la.h; a.w_off = lo.f1_w; a.b_off = lo.f1_b; a.N = 4 * d; a.K = d;
        if ((rc = linear(a, EPI_BIAS_RELU, 4 * li + TC_W_F1))) return rc;
        // ffn.2 -> relu -> +x1 -> LN2
        a.X = la.h; a.Y = la.x2; a.w_off = lo.f2_w; a.b_off = lo.f2_b; a.N = d; a.K = 4 * d;
        a.R = la.x1; a.gamma_off = lo.ln2_w; a.beta_off = lo.ln2_b;
        a.r_save = save ? la.r2 : nullptr; a.st_save = save ? la.st2 : nullptr;
        if ((rc = linear(a, EPI_RES_LN, 4 * li + TC_W_F2))) return rc;
        x_in = la.x2;
    }
    // Q head
    long long Th = Tg;
    const float* head_in = x_in;
    if (last_only) {
        const int li = cfg->n_layers - 1;
        const LayerOff& lo = lay.layer[li];
        const LayerAct& la = act.layer[li];
        const long long nd = (long long)G * n_seq * d;
        float* xl = la.h;                 // [G*n_seq, d] carved out of the (T x 4d) FFN buffer: 9*n_seq*d <= n_seq*L*4d
        float* qlb = xl + nd; float* olb = qlb + nd; float* x1l = olb + nd; float* hl = x1l + nd; float* x2l = hl + 4 * nd;
        LinArgs a{};
        a.P = P;
        if (act_fused) {
            if ((rc = launch_act_fused(*cfg, lay, P, S, G, pk, tab.act_img_off, n_seq, L, xl, olb, st))) return rc;
        } else {
        {
            dim3 grid(dtqn_cdiv((long long)n_seq * d, 256), 1, G);
            prof_begin(PROF_OTHER, st);
            gather_last_kernel<<<grid, 256, 0, st>>>(x_in, S, n_seq, L, d, xl);
            prof_end(PROF_OTHER, st, 0.0);
            DTQN_LAUNCH_CHECK();
        }
        // K | V of every token: rows [d, 3d) of in_proj
        a.Tg = (int)Tg; a.X = x_in; a.Y = la.qkv; a.w_off = lo.in_w + (long long)d * d; a.b_off = lo.in_b + d; a.N = 2 * d; a.K = d;
        if ((rc = linear(a, EPI_BIAS, 4 * cfg->n_layers + 1 + 2 * li))) return rc;
        // Q of the last token: rows [0, d)
        a.Tg = n_seq; a.X = xl; a.Y = qlb; a.w_off = lo.in_w; a.b_off = lo.in_b; a.N = d; a.K = d;
        if ((rc = launch_linear<EPI_BIAS>(a, G, d, st))) return rc;
        {
            dim3 grid(n_seq, G);
            const float scale = 1.0f / sqrtf((float)hd);
            prof_begin(PROF_ATTN_FWD, st);
            const size_t smem = sizeof(float) * (size_t)L * (2 * d + 4);
            if (smem > 48 * 1024) {
                if (hd == 8) cudaFuncSetAttribute(attn_last_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                else if (hd == 16) cudaFuncSetAttribute(attn_last_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                else if (hd == 32) cudaFuncSetAttribute(attn_last_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            }
            if (hd == 8) attn_last_kernel<8><<<grid, 32 * H, smem, st>>>(qlb, la.qkv, S, n_seq, L, d, scale, olb);
            else if (hd == 16) attn_last_kernel<16><<<grid, 32 * H, smem, st>>>(qlb, la.qkv, S, n_seq, L, d, scale, olb);
            else if (hd == 32) attn_last_kernel<32><<<grid, 32 * H, smem, st>>>(qlb, la.qkv, S, n_seq, L, d, scale, olb);
            else if (hd == 4) attn_last_kernel<4><<<grid, 32 * H, smem, st>>>(qlb, la.qkv, S, n_seq, L, d, scale, olb);
            else return DTQN_E_UNSUPPORTED;
            prof_end(PROF_ATTN_FWD, st, 4.0 * (double)G * n_seq * L * d);
            DTQN_LAUNCH_CHECK();
        }
        }
        a.Tg = n_seq;
        a.X = olb; a.Y = x1l; a.w_off = lo.out_w; a.b_off = lo.out_b; a.N = d; a.K = d;
        a.R = xl; a.gamma_off = lo.ln1_w; a.beta_off = lo.ln1_b; a.r_save = nullptr; a.st_save = nullptr;
        if ((rc = launch_linear<EPI_RES_LN>(a, G, d, st))) return rc;
        a.X = x1l; a.Y = hl; a.w_off = lo.f1_w; a.b_off = lo.f1_b; a.N = 4 * d; a.K = d;
        if ((rc = launch_linear<EPI_BIAS_RELU>(a, G, d, st))) return rc;
        a.X = hl; a.Y = x2l; a.w_off = lo.f2_w; a.b_off = lo.f2_b; a.N = d; a.K = 4 * d;
        a.R = x1l; a.gamma_off = lo.ln2_w; a.beta_off = lo.ln2_b;
        if ((rc = launch_linear<EPI_RES_LN>(a, G, d, st))) return rc;
        head_in = x2l; Th = n_seq;
    } else if (q_mode == 1) {
        // only the last valid position feeds the head; qkv of layer 0 is free scratch by now
        float* xl = act.layer[0].qkv;
        dim3 grid(dtqn_cdiv((long long)n_seq * d, 256), 1, G);
        prof_begin(PROF_OTHER, st);
        gather_last_kernel<<<grid, 256, 0, st>>>(x_in, S, n_seq, L, d, xl);
        prof_end(PROF_OTHER, st, 0.0);
        DTQN_LAUNCH_CHECK();
        head_in = xl; Th = n_seq;
    }
    {
        LinArgs a{};
        a.P = P; a.Tg = (int)Th; a.X = head_in; a.Y = act.hh; a.w_off = lay.h1_w; a.b_off = lay.h1_b; a.N = d; a.K = d;
        if ((rc = launch_linear<EPI_BIAS_RELU>(a, G, d, st))) return rc;
        dim3 grid(dtqn_cdiv(Th, 8), 1, G);
        prof_begin(PROF_HEAD, st);
        head_kernel<<<grid, 256, 0, st>>>(act.hh, P, lay.h2_w, lay.h2_b, Th, d, cfg->num_actions, q_out);
        prof_end(PROF_HEAD, st, 2.0 * (double)Th * G * d * cfg->num_actions);
        DTQN_LAUNCH_CHECK();
    }
    return 0;
}
