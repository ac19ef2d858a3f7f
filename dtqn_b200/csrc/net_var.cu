// General fp32 forward / backward of the DTQN network for the reference's ablation flags (run.py:98-103,151-167):
//   --a-embed N   previous-action embedding concatenated in front of the observation embedding (dtqn.py:64-71,184-192)
//   --identity    TransformerIdentityLayer: LayerNorm before each sub-layer (transformer.py:81-101)
//   --gate gru    GRUGate instead of the residual add, ONE attention gate and ONE mlp gate shared by every layer
//                 (gates.py:5-31, dtqn.py:107-131): the shared parameters receive the SUM of the layers' gradients
//   --dropout p   on the token embedding (dtqn.py:195-199), the attention probabilities (nn.MultiheadAttention) and the
//                 FFN output (transformer.py:41), train-mode networks only; masks come from a counter-based generator so
//                 the backward regenerates exactly the forward's masks (torch's own stream is not reproduced: p = 0 is
//                 exact, p > 0 is checked statistically and by finite differences)
// and any d_model (multiple of 4, <= 256).  One kernel per operation, fp32 CUDA cores, every reduction in a fixed order
// (bit-reproducible).  The default architecture never comes here: it has the fused kernels of net_seq.cu / act_fused.cu /
// linear_tc.cu; this path trades speed for generality, like the reference's own flags do.
#include "net.cuh"
#include "prof.cuh"

namespace {

// ---- activation workspace -------------------------------------------------------------------------------------------------
struct VarGate { float *z, *r, *hg, *rx; };                  // GRU gate: sigmoid / sigmoid / tanh outputs and r * x
struct VarLayer {
    float *lnA, *stA, *qkv, *o, *r1, *u1, *lnB, *stB, *h, *r2, *u2;
    VarGate g1, g2;
};
struct VarAct {
    float* x0;
    VarLayer layer[DTQN_MAX_LAYERS];
    float *hh, *tmp;                                         // tmp: [T, d] rows of the last valid position (acting)
    long long total;
};

long long var_act_layout(const dtqn_net_cfg& c, long long T, float* base, VarAct& A) {
    const long long d = c.d_model;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o = al4(o + n); return p; };
    A.x0 = take(T * d);
    for (int i = 0; i < c.n_layers; ++i) {
        VarLayer& l = A.layer[i];
        l.lnA = take(T * d); l.stA = take(T * 2); l.qkv = take(T * 3 * d); l.o = take(T * d); l.r1 = take(T * d);
        l.u1 = take(T * d); l.lnB = take(T * d); l.stB = take(T * 2); l.h = take(T * 4 * d); l.r2 = take(T * d);
        l.u2 = take(T * d);
        if (c.gate_gru) {
            for (VarGate* g : {&l.g1, &l.g2}) { g->z = take(T * d); g->r = take(T * d); g->hg = take(T * d); g->rx = take(T * d); }
        }
    }
    A.hh = take(T * d); A.tmp = take(T * d);
    A.total = o;
    return o;
}

// ---- counter-based dropout masks --------------------------------------------------------------------------------------------
// keep(seed, site, idx) with probability 1 - p; scale 1 / (1 - p).  splitmix64 finaliser over a (seed, site, index) counter.
__device__ __forceinline__ float var_uniform(unsigned long long seed, unsigned site, unsigned long long idx) {
    unsigned long long z = seed * 0xD1342543DE82EF95ull + (unsigned long long)(site + 1) * 0x9E3779B97F4A7C15ull + idx * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}
struct Drop {                                                 // p == 0 or state == NULL: identity
    const unsigned long long* state; float p; unsigned site;
    __device__ __forceinline__ float scale(unsigned long long idx) const {
        if (p <= 0.f || !state) return 1.f;
        return var_uniform(*state, site, idx) >= p ? 1.f / (1.f - p) : 0.f;
    }
};
__global__ void var_bump_kernel(unsigned long long* state) { *state += 1ull; }

// ---- generic fp32 GEMM, 64 x 64 tile, 256 threads (4 x 4 outputs each), bounds-checked ------------------------------------------
//   NT: C[m,n] = sum_k A[m,k] * B[n,k]     forward Linear (B = W [N,K])
//   NN: C[m,n] = sum_k A[m,k] * B[k,n]     input gradient (A = dY [T,N_w], B = W [N_w,K_w])
//   TN: C[m,n] = sum_t A[t,m] * B[t,n]     weight gradient (A = dY [T,N_w], B = X [T,K_w]); one CTA walks all t in order
// epilogue: + bias[n], + C (accumulate), then EPI.
enum { G_NT = 0, G_NN = 1, G_TN = 2 };
enum { E_NONE = 0, E_RELU = 1, E_MASK = 2 /* aux > 0 ? v : 0 */, E_ADD = 3 /* v + aux */ };
struct GemmArgs {
    const float *A, *B, *bias, *aux;
    float* C;
    int M, N, K, lda, ldb, ldc, accumulate;
    long long a_gs, b_gs, bias_gs, c_gs;                      // per-group (blockIdx.z) strides; B/bias come from gB[g] when set
    const float* gB[DTQN_MAX_GROUPS]; long long gB_off, gbias_off;
};
template <int MODE, int EPI>
__global__ void __launch_bounds__(256)
var_gemm_kernel(GemmArgs a) {
    __shared__ float As[16][64 + 1], Bs[16][64 + 1];
    const int g = blockIdx.z;
    const float* A = a.A + g * a.a_gs;
    const float* B = a.gB[0] ? a.gB[g] + a.gB_off : a.B + g * a.b_gs;
    const float* bias = a.gB[0] ? (a.gbias_off >= 0 ? a.gB[g] + a.gbias_off : nullptr) : (a.bias ? a.bias + g * a.bias_gs : nullptr);
    float* C = a.C + g * a.c_gs;
    const float* aux = a.aux ? a.aux + g * a.c_gs : nullptr;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.K; k0 += 16) {
        for (int e = tid; e < 16 * 64; e += 256) {
            if (MODE == G_TN) {                               // As[kk][m] = A[k0+kk, m0+m];  Bs[kk][n] = B[k0+kk, n0+n]
                const int kk = e >> 6, c = e & 63;
                const int k = k0 + kk;
                As[kk][c] = (k < a.K && m0 + c < a.M) ? A[(size_t)k * a.lda + m0 + c] : 0.f;
                Bs[kk][c] = (k < a.K && n0 + c < a.N) ? B[(size_t)k * a.ldb + n0 + c] : 0.f;
            } else {
                const int r = e >> 4, kk = e & 15;            // As[kk][m] = A[m0+m, k0+kk]
                const int k = k0 + kk;
                As[kk][r] = (k < a.K && m0 + r < a.M) ? A[(size_t)(m0 + r) * a.lda + k] : 0.f;
                if (MODE == G_NT) Bs[kk][r] = (k < a.K && n0 + r < a.N) ? B[(size_t)(n0 + r) * a.ldb + k] : 0.f;
            }
        }
        if (MODE == G_NN)
            for (int e = tid; e < 16 * 64; e += 256) {
                const int kk = e >> 6, c = e & 63;
                const int k = k0 + kk;
                Bs[kk][c] = (k < a.K && n0 + c < a.N) ? B[(size_t)k * a.ldb + n0 + c] : 0.f;
            }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            const size_t o = (size_t)m * a.ldc + n;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (a.accumulate) v += C[o];
            if (EPI == E_RELU) v = fmaxf(v, 0.f);
            else if (EPI == E_MASK) v = aux[o] > 0.f ? v : 0.f;
            else if (EPI == E_ADD) v += aux[o];
            C[o] = v;
        }
    }
}

template <int MODE, int EPI>
int var_gemm(const GemmArgs& a, int G, cudaStream_t st) {
    dim3 grid(dtqn_cdiv(a.M, 64), dtqn_cdiv(a.N, 64), G);
    prof_begin(MODE == G_NT ? PROF_LINEAR : (MODE == G_NN ? PROF_DGRAD : PROF_WGRAD), st);
    var_gemm_kernel<MODE, EPI><<<grid, 256, 0, st>>>(a);
    prof_end(MODE == G_NT ? PROF_LINEAR : (MODE == G_NN ? PROF_DGRAD : PROF_WGRAD), st, 2.0 * G * (double)a.M * a.N * a.K);
    DTQN_LAUNCH_CHECK();
    return 0;
}

// ---- token embedding: [action embedding of the PREVIOUS token's action | observation embedding] + position, dropout ------------
// dtqn.py:181-199.  With L > 1 the action embeddings are rolled right by one and position 0 gets zeros; with L == 1 the
// reference skips the roll (dtqn.py:187: `if history_len > 1`), i.e. the single token carries its own action's embedding.
__device__ __forceinline__ int var_ring_row(const dtqn_obs_src& s, int i, int j, bool& valid) {
    valid = true;
    if (!s.timestep) return j;
    const int t = s.timestep[i], n = min(s.ring_len, t + 1);
    valid = j < n;
    return valid ? (t + 1 - n + j) % s.ring_len : 0;
}
__global__ void __launch_bounds__(256)
var_embed_kernel(GroupPtrs P, GroupSrc S, dtqn_net_cfg c, NetLayout lay, int n_seq, int L, Drop drop, float* __restrict__ x0) {
    const int g = blockIdx.z, d = c.d_model, ad = c.action_dim, dob = d - ad;
    const long long Tg = (long long)n_seq * L;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Tg * d) return;
    const long long t = idx / d;
    const int col = (int)(idx % d), i = (int)(t / L), j = (int)(t % L);
    const dtqn_obs_src& s = S.s[g];
    const float* p = P.p[g];
    float v;
    if (col < ad) {
        v = 0.f;
        // the reference forwards only the n valid tokens of an acting context, so its `history_len > 1` test sees n, not L
        const int n_tok = s.timestep ? min(min(s.ring_len, s.timestep[i] + 1), L) : L;
        const int ja = n_tok > 1 ? j - 1 : j;
        if (ja >= 0) {
            bool valid; const int row = var_ring_row(s, i, ja, valid);
            const int a = valid ? (int)s.actions[(long long)i * s.act_stride + row] : 0;
            v = p[lay.act_table + (long long)min(a, c.num_actions - 1) * ad + col];
        }
    } else {
        const int cc = col - ad;
        bool valid; const int row = var_ring_row(s, i, j, valid);
        const float* ob = s.obs + (long long)i * s.seq_stride + (long long)row * c.obs_dim;
        v = p[lay.emb_b + cc];
        if (c.discrete) {
            const int E = c.embed_per_obs;
            for (int k = 0; k < c.obs_dim; ++k) {
                int tok = (int)(valid ? ob[k] : s.obs_mask);
                tok = tok < 0 ? 0 : (tok >= c.vocab ? c.vocab - 1 : tok);
                for (int e = 0; e < E; ++e) v = fmaf(p[lay.emb_table + tok * E + e], p[lay.emb_w + (long long)cc * lay.k_in + k * E + e], v);
            }
        } else {
            for (int k = 0; k < c.obs_dim; ++k) v = fmaf(valid ? ob[k] : s.obs_mask, p[lay.emb_w + (long long)cc * lay.k_in + k], v);
        }
    }
    v += p[lay.pos + (long long)j * d + col];
    const long long gi = (long long)g * Tg * d + idx;
    if (s.train_mode) v *= drop.scale((unsigned long long)gi);
    x0[gi] = v;
}

// ---- LayerNorm over rows of width d (one warp per row), saves (mean, rstd) ---------------------------------------------------------
__global__ void __launch_bounds__(256)
var_ln_kernel(const float* __restrict__ x, GroupPtrs P, long long g_off, long long b_off, long long Tg, int d,
              float* __restrict__ y, float* __restrict__ st) {
    const int g = blockIdx.z, lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= Tg) return;
    const long long row = (long long)g * Tg + t;
    const float* xr = x + row * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)d;
    float vs = 0.f;
    for (int c = lane; c < d; c += 32) { const float dl = xr[c] - mean; vs = fmaf(dl, dl, vs); }
    const float rstd = 1.0f / sqrtf(warp_sum(vs) / (float)d + 1e-5f);
    const float* p = P.p[g];
    for (int c = lane; c < d; c += 32) y[row * d + c] = (xr[c] - mean) * rstd * p[g_off + c] + p[b_off + c];
    if (lane == 0) { st[row * 2] = mean; st[row * 2 + 1] = rstd; }
}
// du = rstd * (g - mean(g) - xh * mean(g * xh)), g = dy * gamma; also t1 = dy * xh (for dgamma = colsum(t1))
__global__ void __launch_bounds__(256)
var_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ u, const float* __restrict__ st,
                  const float* __restrict__ gamma, long long T, int d, float* __restrict__ du, float* __restrict__ t1,
                  int accumulate) {
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    const float mean = st[2 * t], rstd = st[2 * t + 1];
    float m1 = 0.f, m2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float xh = (u[t * d + c] - mean) * rstd, gg = dy[t * d + c] * gamma[c];
        m1 += gg; m2 = fmaf(gg, xh, m2);
    }
    m1 = warp_sum(m1) / (float)d; m2 = warp_sum(m2) / (float)d;
    for (int c = lane; c < d; c += 32) {
        const float xh = (u[t * d + c] - mean) * rstd, dyv = dy[t * d + c], gg = dyv * gamma[c];
        const float v = rstd * (gg - m1 - xh * m2);
        t1[t * d + c] = dyv * xh;
        du[t * d + c] = accumulate ? du[t * d + c] + v : v;
    }
}
// out[c] (+)= sum_t a[t, c] in row order: block = 32 columns x 8 row phases, phases combined in order
__global__ void __launch_bounds__(256)
var_colsum_kernel(const float* __restrict__ a, long long T, int C, float* __restrict__ out, int accumulate) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ph = threadIdx.x >> 5, c = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < C) for (long long t = ph; t < T; t += 8) s += a[t * C + c];
    red[ph][cx] = s;
    __syncthreads();
    if (ph == 0 && c < C) {
        float tot = 0.f;
        for (int k = 0; k < 8; ++k) tot += red[k][cx];
        out[c] = accumulate ? out[c] + tot : tot;
    }
}

// ---- causal attention, one CTA per (head, sequence), thread = query row; optional dropout on the probabilities -------------------
template <int HD>
__global__ void __launch_bounds__(128)
var_attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o, GroupSrc S, int n_seq, int L, int d, float scale, Drop drop) {
    __shared__ float Ks[128][HD + 1], Vs[128][HD + 1];
    const int h = blockIdx.x, H = gridDim.x;
    const long long seq = blockIdx.y, t0 = seq * L;
    const bool train = S.s[seq / n_seq].train_mode != 0;
    for (int e = threadIdx.x; e < L * HD; e += blockDim.x) {
        const int r = e / HD, c = e % HD;
        const float* base = qkv + (t0 + r) * (size_t)(3 * d) + h * HD + c;
        Ks[r][c] = base[d]; Vs[r][c] = base[2 * d];
    }
    __syncthreads();
    const int j = threadIdx.x;
    if (j >= L) return;
    float q[HD], acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { q[c] = qkv[(t0 + j) * (size_t)(3 * d) + h * HD + c] * scale; acc[c] = 0.f; }
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
    for (int i = 0; i <= j; ++i) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) s = fmaf(q[c], Ks[i][c], s);
        float pw = expf(s - m) * il;
        if (train) pw *= drop.scale(((unsigned long long)(seq * H + h) * L + j) * L + i);
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = fmaf(pw, Vs[i][c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) o[(t0 + j) * (size_t)d + h * HD + c] = acc[c];
}
template <int HD>
__global__ void __launch_bounds__(128)
var_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ d_o,
                    float* __restrict__ d_qkv, int L, int d, float scale, Drop drop, int train) {
    __shared__ float Qs[128][HD + 1], Ks[128][HD + 1], Vs[128][HD + 1], Gs[128][HD + 1];
    __shared__ float sm_m[128], sm_il[128], sm_D[128];
    const int h = blockIdx.x, H = gridDim.x;
    const long long seq = blockIdx.y, t0 = seq * L;
    const int tid = threadIdx.x;
    for (int e = tid; e < L * HD; e += blockDim.x) {
        const int r = e / HD, c = e % HD;
        const float* base = qkv + (t0 + r) * (size_t)(3 * d) + h * HD + c;
        Qs[r][c] = base[0] * scale; Ks[r][c] = base[d]; Vs[r][c] = base[2 * d];
        Gs[r][c] = d_o[(t0 + r) * (size_t)d + h * HD + c];
    }
    __syncthreads();
    const int j = tid;
    if (j < L) {                                              // query row j: softmax statistics, D_j = dO_j . O_j, dQ_j
        float q[HD], gq[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { q[c] = Qs[j][c]; gq[c] = Gs[j][c]; }
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
#pragma unroll
        for (int c = 0; c < HD; ++c) D = fmaf(gq[c], o[(t0 + j) * (size_t)d + h * HD + c], D);
        float dq[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = 0.f;
        for (int i = 0; i <= j; ++i) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) { s = fmaf(q[c], Ks[i][c], s); dp = fmaf(gq[c], Vs[i][c], dp); }
            const float p = expf(s - m) * il;
            if (train) dp *= drop.scale(((unsigned long long)(seq * H + h) * L + j) * L + i);
            const float ds = p * (dp - D);
#pragma unroll
            for (int c = 0; c < HD; ++c) dq[c] = fmaf(ds, Ks[i][c], dq[c]);
        }
        sm_m[j] = m; sm_il[j] = il; sm_D[j] = D;
#pragma unroll
        for (int c = 0; c < HD; ++c) d_qkv[(t0 + j) * (size_t)(3 * d) + h * HD + c] = dq[c] * scale;
    }
    __syncthreads();
    const int i = tid;
    if (i < L) {                                              // key row i: dK_i, dV_i over the queries j >= i (in order)
        float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { k[c] = Ks[i][c]; v[c] = Vs[i][c]; dk[c] = 0.f; dv[c] = 0.f; }
        for (int jj = i; jj < L; ++jj) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) { s = fmaf(Qs[jj][c], k[c], s); dp = fmaf(Gs[jj][c], v[c], dp); }
            const float p = expf(s - sm_m[jj]) * sm_il[jj];
            const float sc = train ? drop.scale(((unsigned long long)(seq * H + h) * L + jj) * L + i) : 1.f;
            const float ds = p * (dp * sc - sm_D[jj]);
#pragma unroll
            for (int c = 0; c < HD; ++c) { dk[c] = fmaf(ds, Qs[jj][c], dk[c]); dv[c] = fmaf(p * sc, Gs[jj][c], dv[c]); }
        }
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            d_qkv[(t0 + i) * (size_t)(3 * d) + d + h * HD + c] = dk[c];
            d_qkv[(t0 + i) * (size_t)(3 * d) + 2 * d + h * HD + c] = dv[c];
        }
    }
}

// ---- elementwise pieces -------------------------------------------------------------------------------------------------------
__global__ void var_add_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float* __restrict__ u) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) u[i] = x[i] + y[i];                            // ResGate (gates.py:40-41)
}
// r2 = relu(dropout(f)) given relu(f): multiply by the mask scale of train-mode groups
__global__ void var_drop_kernel(float* __restrict__ x, long long n_per_group, GroupSrc S, Drop drop) {
    const int g = blockIdx.z;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_group || !S.s[g].train_mode) return;
    const long long gi = (long long)g * n_per_group + i;
    x[gi] *= drop.scale((unsigned long long)gi);
}
__device__ __forceinline__ float var_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }
// GRU gate, step 1: z = sigmoid(az), r = sigmoid(ar), rx = r * x        (az, ar hold the summed pre-activations)
__global__ void var_gru_zr_kernel(float* __restrict__ z, float* __restrict__ r, const float* __restrict__ x, long long n,
                                  float* __restrict__ rx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float zz = var_sigmoid(z[i]), rr = var_sigmoid(r[i]);
    z[i] = zz; r[i] = rr; rx[i] = rr * x[i];
}
// step 2: h = tanh(ag), u = (1 - z) x + z h
__global__ void var_gru_out_kernel(float* __restrict__ hg, const float* __restrict__ z, const float* __restrict__ x, long long n,
                                   float* __restrict__ u) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float hh = tanhf(hg[i]);
    hg[i] = hh;
    u[i] = (1.f - z[i]) * x[i] + z[i] * hh;
}
// backward, step 1: dag = du z (1 - h^2), daz = du (h - x) z (1 - z), dx = du (1 - z)
__global__ void var_gru_bwd1_kernel(const float* __restrict__ du, const float* __restrict__ z, const float* __restrict__ hg,
                                    const float* __restrict__ x, long long n, float* __restrict__ dag, float* __restrict__ daz,
                                    float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = du[i], zz = z[i], hh = hg[i];
    dag[i] = g * zz * (1.f - hh * hh);
    daz[i] = g * (hh - x[i]) * zz * (1.f - zz);
    dx[i] = g * (1.f - zz);
}
// step 2: dar = drx x r (1 - r), dx += drx r
__global__ void var_gru_bwd2_kernel(const float* __restrict__ drx, const float* __restrict__ r, const float* __restrict__ x,
                                    long long n, float* __restrict__ dar, float* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rr = r[i], g = drx[i];
    dar[i] = g * x[i] * rr * (1.f - rr);
    dx[i] += g * rr;
}
// da = g * [r > 0] * s   (ReLU backward, optional constant scale 1 / (1 - p) of the FFN-output dropout)
__global__ void var_relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ r, long long n, float s, float* __restrict__ da) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) da[i] = r[i] > 0.f ? g[i] * s : 0.f;
}
__global__ void var_axpy_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}
__global__ void var_gather_last_kernel(const float* __restrict__ x, GroupSrc S, int n_seq, int L, int d, float* __restrict__ out) {
    const int g = blockIdx.z;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_seq * d) return;
    const int i = (int)(idx / d), c = (int)(idx % d);
    int last = L - 1;
    const dtqn_obs_src& s = S.s[g];
    if (s.timestep) last = min(min(s.ring_len, s.timestep[i] + 1), L) - 1;
    out[((long long)g * n_seq + i) * d + c] = x[(((long long)g * n_seq + i) * L + last) * d + c];
}

// ---- embedding backward (deterministic: every output element is one thread walking the tokens in order) ------------------------
// dx0 already carries the embedding-dropout scale.  gpos[j, c] = sum_b dx0[b, j, c]
__global__ void var_pos_bwd_kernel(const float* __restrict__ dx0, int B, int L, int d, float* __restrict__ gpos) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * d) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dx0[(size_t)b * L * d + idx];
    gpos[idx] = s;
}
__global__ void var_drop_bwd_kernel(float* __restrict__ g, long long n, Drop drop) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] *= drop.scale((unsigned long long)i);
}
// one thread per gradient element of: act_table [A, ad] | emb_w [dob, KI] | emb_b [dob] | emb_table [vocab, E]
__global__ void __launch_bounds__(256)
var_embed_bwd_kernel(const float* __restrict__ dx0, dtqn_obs_src s, dtqn_net_cfg c, NetLayout lay, const float* __restrict__ p,
                     int B, int L, float* __restrict__ grads) {
    const int d = c.d_model, ad = c.action_dim, dob = d - ad, KI = lay.k_in, E = c.discrete ? c.embed_per_obs : 1;
    const int n_act = c.num_actions * ad, n_w = dob * KI, n_b = dob, n_tab = c.discrete ? c.vocab * E : 0;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_act + n_w + n_b + n_tab) return;
    const long long T = (long long)B * L;
    float sum = 0.f;
    if (e < n_act) {                                          // tokens whose PREVIOUS action is a (dtqn.py:187-190)
        const int a = e / ad, col = e % ad;
        for (long long t = 0; t < T; ++t) {
            const int i = (int)(t / L), j = (int)(t % L), ja = L > 1 ? j - 1 : j;
            if (ja < 0) continue;
            if ((int)s.actions[(long long)i * s.act_stride + ja] == a) sum += dx0[t * d + col];
        }
        grads[lay.act_table + e] = sum;
    } else if (e < n_act + n_w) {
        const int cc = (e - n_act) / KI, f = (e - n_act) % KI;
        for (long long t = 0; t < T; ++t) {
            const int i = (int)(t / L), j = (int)(t % L);
            const float* ob = s.obs + (long long)i * s.seq_stride + (long long)j * c.obs_dim;
            float in;
            if (c.discrete) {
                int tok = (int)ob[f / E]; tok = tok < 0 ? 0 : (tok >= c.vocab ? c.vocab - 1 : tok);
                in = p[lay.emb_table + tok * E + f % E];
            } else in = ob[f];
            sum = fmaf(dx0[t * d + ad + cc], in, sum);
        }
        grads[lay.emb_w + (e - n_act)] = sum;
    } else if (e < n_act + n_w + n_b) {
        const int cc = e - n_act - n_w;
        for (long long t = 0; t < T; ++t) sum += dx0[t * d + ad + cc];
        grads[lay.emb_b + cc] = sum;
    } else {
        const int v = (e - n_act - n_w - n_b) / E, ee = (e - n_act - n_w - n_b) % E;
        for (long long t = 0; t < T; ++t) {
            const int i = (int)(t / L), j = (int)(t % L);
            const float* ob = s.obs + (long long)i * s.seq_stride + (long long)j * c.obs_dim;
            for (int k = 0; k < c.obs_dim; ++k) {
                int tok = (int)ob[k]; tok = tok < 0 ? 0 : (tok >= c.vocab ? c.vocab - 1 : tok);
                if (tok != v) continue;
                float din = 0.f;                              // d in[t, k*E + ee] = sum_c dx0[t, ad + c] W[c, k*E + ee]
                for (int cc = 0; cc < dob; ++cc) din = fmaf(dx0[t * d + ad + cc], p[lay.emb_w + (long long)cc * KI + k * E + ee], din);
                sum += din;
            }
        }
        grads[lay.emb_table + v * E + ee] = sum;
    }
}
// head ffn.2 backward (N = A is tiny): d_hh = (dq W2) * [hh > 0]; dW2, db2 by one thread per element over all tokens
__global__ void __launch_bounds__(256)
var_head_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ hh, const float* __restrict__ W2, long long T, int d,
                    int A, float* __restrict__ d_hh, float* __restrict__ gW2, float* __restrict__ gb2) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < T * d) {
        const long long t = idx / d; const int c = (int)(idx % d);
        float s = 0.f;
        for (int a = 0; a < A; ++a) s = fmaf(dq[t * A + a], W2[a * d + c], s);
        d_hh[idx] = hh[idx] > 0.f ? s : 0.f;
    }
    if (idx < (long long)A * d + A) {
        float s = 0.f;
        if (idx < (long long)A * d) {
            const int a = (int)(idx / d), c = (int)(idx % d);
            for (long long t = 0; t < T; ++t) s = fmaf(dq[t * A + a], hh[t * d + c], s);
            gW2[idx] = s;
        } else {
            const int a = (int)(idx - (long long)A * d);
            for (long long t = 0; t < T; ++t) s += dq[t * A + a];
            gb2[a] = s;
        }
    }
}

inline dim3 grid1(long long n) { return dim3((unsigned)dtqn_cdiv(n, 256)); }

// Y[g] = X[g] W_g^T (+ b_g) with the weights of group g's network at w_off / b_off
template <int EPI>
int var_linear(const float* X, float* Y, const GroupPtrs& P, long long w_off, long long b_off, long long Tg, int N, int K, int G,
               int accumulate, cudaStream_t st) {
    GemmArgs a{};
    a.A = X; a.C = Y; a.M = (int)Tg; a.N = N; a.K = K; a.lda = K; a.ldb = K; a.ldc = N; a.accumulate = accumulate;
    a.a_gs = Tg * K; a.c_gs = Tg * N;
    for (int g = 0; g < G; ++g) a.gB[g] = P.p[g];
    a.gB_off = w_off; a.gbias_off = b_off;
    return var_gemm<G_NT, EPI>(a, G, st);
}

struct VarCtx {
    const dtqn_net_cfg& c; const NetLayout& lay; GroupPtrs P; GroupSrc S; int G, n_seq, L; long long Tg, T; cudaStream_t st;
    Drop drop(unsigned site) const { return Drop{c.dropout > 0.f ? (const unsigned long long*)c.dropout_state : nullptr, c.dropout, site}; }
};

// u = gate(x, y): ResGate or the shared GRUGate `k` (0 attention, 1 mlp); GRU saves z, r, h, r*x for the backward
int var_gate_fwd(const VarCtx& v, int k, const float* x, const float* y, float* u, const VarGate& gs) {
    const long long n = v.T * v.c.d_model;
    const int d = v.c.d_model;
    if (!v.c.gate_gru) {
        var_add_kernel<<<grid1(n), 256, 0, v.st>>>(x, y, n, u);
        DTQN_LAUNCH_CHECK();
        return 0;
    }
    const GateOff& q = v.lay.gate[k];
    int rc;
    if ((rc = var_linear<E_NONE>(y, gs.z, v.P, q.w_z, q.b_z, v.Tg, d, d, v.G, 0, v.st))) return rc;     // w_z(y) + b
    if ((rc = var_linear<E_NONE>(x, gs.z, v.P, q.u_z, -1, v.Tg, d, d, v.G, 1, v.st))) return rc;        // + u_z(x)
    if ((rc = var_linear<E_NONE>(y, gs.r, v.P, q.w_r, -1, v.Tg, d, d, v.G, 0, v.st))) return rc;
    if ((rc = var_linear<E_NONE>(x, gs.r, v.P, q.u_r, -1, v.Tg, d, d, v.G, 1, v.st))) return rc;
    var_gru_zr_kernel<<<grid1(n), 256, 0, v.st>>>(gs.z, gs.r, x, n, gs.rx);
    DTQN_LAUNCH_CHECK();
    if ((rc = var_linear<E_NONE>(y, gs.hg, v.P, q.w_g, -1, v.Tg, d, d, v.G, 0, v.st))) return rc;
    if ((rc = var_linear<E_NONE>(gs.rx, gs.hg, v.P, q.u_g, -1, v.Tg, d, d, v.G, 1, v.st))) return rc;
    var_gru_out_kernel<<<grid1(n), 256, 0, v.st>>>(gs.hg, gs.z, x, n, u);
    DTQN_LAUNCH_CHECK();
    return 0;
}

template <int HD> void var_attn_fwd_launch(const VarCtx& v, const float* qkv, float* o, unsigned site) {
    const int d = v.c.d_model;
    dim3 grid(v.c.n_heads, (unsigned)(v.n_seq * v.G));
    var_attn_fwd_kernel<HD><<<grid, v.L <= 64 ? 64 : 128, 0, v.st>>>(qkv, o, v.S, v.n_seq, v.L, d, 1.0f / sqrtf((float)HD), v.drop(site));
}

}  // namespace

namespace {
__global__ void var_scales_kernel(unsigned long long seed, unsigned site, float p, long long n, float* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = var_uniform(seed, site, (unsigned long long)i) >= p ? 1.f / (1.f - p) : 0.f;
}
}  // namespace
extern "C" int dtqn_dropout_scales(uint64_t counter, uint32_t site, float p, int64_t n, float* out, void* stream) {
    if (!out || n <= 0 || p < 0.f || p >= 1.f) return DTQN_E_ARG;
    var_scales_kernel<<<grid1(n), 256, 0, (cudaStream_t)stream>>>(counter, site, p, n, out);
    DTQN_LAUNCH_CHECK();
    return 0;
}

long long var_workspace_floats(const dtqn_net_cfg& c, long long T) {
    VarAct A;
    return var_act_layout(c, T, nullptr, A);
}

int var_forward(const dtqn_net_cfg& c, const NetLayout& lay, int G, const float* const* params, const dtqn_obs_src* src,
                int n_seq, int L, int q_mode, int save, float* ws, long long ws_floats, float* q_out, cudaStream_t st) {
    const int d = c.d_model, H = c.n_heads, hd = d / H;
    if (L > 128 || (hd != 4 && hd != 8 && hd != 16)) return DTQN_E_UNSUPPORTED;
    VarCtx v{c, lay, {}, {}, G, n_seq, L, (long long)n_seq * L, (long long)n_seq * L * G, st};
    for (int g = 0; g < G; ++g) {
        v.P.p[g] = params[g]; v.S.s[g] = src[g];
        if (c.action_dim > 0 && !src[g].actions) return DTQN_E_ARG;
    }
    VarAct act;
    if (var_act_layout(c, v.T, ws, act) > ws_floats) return DTQN_E_ARG;
    int rc;
    prof_begin(PROF_EMBED, st);
    var_embed_kernel<<<dim3((unsigned)dtqn_cdiv(v.Tg * d, 256), 1, G), 256, 0, st>>>(v.P, v.S, c, lay, n_seq, L, v.drop(1), act.x0);
    prof_end(PROF_EMBED, st, 0.0);
    DTQN_LAUNCH_CHECK();
    const float* x_in = act.x0;
    const dim3 ln_grid((unsigned)dtqn_cdiv(v.Tg, 8), 1, G);
    for (int li = 0; li < c.n_layers; ++li) {
        const LayerOff& lo = lay.layer[li];
        const VarLayer& la = act.layer[li];
        const float* att_in = x_in;
        if (c.identity) {                                      // transformer.py:87: x_norm1 = layernorm1(x)
            var_ln_kernel<<<ln_grid, 256, 0, st>>>(x_in, v.P, lo.ln1_w, lo.ln1_b, v.Tg, d, la.lnA, la.stA);
            DTQN_LAUNCH_CHECK();
            att_in = la.lnA;
        }
        if ((rc = var_linear<E_NONE>(att_in, la.qkv, v.P, lo.in_w, lo.in_b, v.Tg, 3 * d, d, G, 0, st))) return rc;
        prof_begin(PROF_ATTN_FWD, st);
        const unsigned site = 2 + 4 * li;
        if (hd == 4) var_attn_fwd_launch<4>(v, la.qkv, la.o, site);
        else if (hd == 8) var_attn_fwd_launch<8>(v, la.qkv, la.o, site);
        else var_attn_fwd_launch<16>(v, la.qkv, la.o, site);
        prof_end(PROF_ATTN_FWD, st, 4.0 * (double)v.T * L * d);
        DTQN_LAUNCH_CHECK();
        if ((rc = var_linear<E_RELU>(la.o, la.r1, v.P, lo.out_w, lo.out_b, v.Tg, d, d, G, 0, st))) return rc;   // relu(attention)
        if ((rc = var_gate_fwd(v, 0, x_in, la.r1, la.u1, la.g1))) return rc;                                     // :72 / :97
        const float* ffn_in;
        if (c.identity) {                                      // :98 x_norm2 = layernorm2(x)
            var_ln_kernel<<<ln_grid, 256, 0, st>>>(la.u1, v.P, lo.ln2_w, lo.ln2_b, v.Tg, d, la.lnB, la.stB);
            ffn_in = la.lnB;
        } else {                                               // :73 x = layernorm1(x)
            var_ln_kernel<<<ln_grid, 256, 0, st>>>(la.u1, v.P, lo.ln1_w, lo.ln1_b, v.Tg, d, la.lnA, la.stA);
            ffn_in = la.lnA;
        }
        DTQN_LAUNCH_CHECK();
        if ((rc = var_linear<E_RELU>(ffn_in, la.h, v.P, lo.f1_w, lo.f1_b, v.Tg, 4 * d, d, G, 0, st))) return rc;
        if ((rc = var_linear<E_RELU>(la.h, la.r2, v.P, lo.f2_w, lo.f2_b, v.Tg, d, 4 * d, G, 0, st))) return rc;  // relu(dropout(f)) = dropout(relu(f))
        if (c.dropout > 0.f) {
            var_drop_kernel<<<dim3((unsigned)dtqn_cdiv(v.Tg * d, 256), 1, G), 256, 0, st>>>(la.r2, v.Tg * d, v.S, v.drop(3 + 4 * li));
            DTQN_LAUNCH_CHECK();
        }
        if (c.identity) {                                      // :100 x = mlp_gate(x, relu(ffn)); no trailing LayerNorm
            if ((rc = var_gate_fwd(v, 1, la.u1, la.r2, la.u2, la.g2))) return rc;
            x_in = la.u2;
        } else {                                               // :76-77
            if ((rc = var_gate_fwd(v, 1, la.lnA, la.r2, la.u2, la.g2))) return rc;
            var_ln_kernel<<<ln_grid, 256, 0, st>>>(la.u2, v.P, lo.ln2_w, lo.ln2_b, v.Tg, d, la.lnB, la.stB);
            DTQN_LAUNCH_CHECK();
            x_in = la.lnB;
        }
    }
    long long Th = v.Tg;
    const float* head_in = x_in;
    if (q_mode == 1) {                                         // acting: only the last valid position feeds the head
        var_gather_last_kernel<<<dim3((unsigned)dtqn_cdiv((long long)n_seq * d, 256), 1, G), 256, 0, st>>>(x_in, v.S, n_seq, L, d, act.tmp);
        DTQN_LAUNCH_CHECK();
        head_in = act.tmp; Th = n_seq;
    }
    if ((rc = var_linear<E_RELU>(head_in, act.hh, v.P, lay.h1_w, lay.h1_b, Th, d, d, G, 0, st))) return rc;
    {
        GemmArgs a{};
        a.A = act.hh; a.C = q_out; a.M = (int)Th; a.N = c.num_actions; a.K = d; a.lda = d; a.ldb = d; a.ldc = c.num_actions;
        a.a_gs = Th * d; a.c_gs = Th * c.num_actions;
        for (int g = 0; g < G; ++g) a.gB[g] = v.P.p[g];
        a.gB_off = lay.h2_w; a.gbias_off = lay.h2_b;
        if ((rc = var_gemm<G_NT, E_NONE>(a, G, st))) return rc;
    }
    if (c.dropout > 0.f && !save) {                            // acting / inference forward: next call draws new masks
        var_bump_kernel<<<1, 1, 0, st>>>((unsigned long long*)c.dropout_state);
        DTQN_LAUNCH_CHECK();
    }
    return 0;
}

// ---- backward ----------------------------------------------------------------------------------------------------------------
namespace {
struct VarBwd { float *g_hh, *gx, *gu, *ga, *gb, *gc, *gd, *gh, *gqkv, *t1; long long total; };
long long var_bwd_layout(const dtqn_net_cfg& c, long long T0, float* base, VarBwd& s) {
    const long long d = c.d_model;
    long long o = 0;
    auto take = [&](long long n) { float* p = base ? base + o : nullptr; o = al4(o + n); return p; };
    s.g_hh = take(T0 * d); s.gx = take(T0 * d); s.gu = take(T0 * d); s.ga = take(T0 * d); s.gb = take(T0 * d);
    s.gc = take(T0 * d); s.gd = take(T0 * d); s.gh = take(T0 * 4 * d); s.gqkv = take(T0 * 3 * d); s.t1 = take(T0 * d);
    s.total = o;
    return o;
}
// dX = dY W  (W [N_w, K_w]);  optional epilogue with aux
template <int EPI>
int var_dgrad(const float* dY, const float* W, const float* aux, float* dX, long long T, int Nw, int Kw, int accumulate, cudaStream_t st) {
    GemmArgs a{};
    a.A = dY; a.B = W; a.aux = aux; a.C = dX; a.M = (int)T; a.N = Kw; a.K = Nw; a.lda = Nw; a.ldb = Kw; a.ldc = Kw; a.accumulate = accumulate;
    return var_gemm<G_NN, EPI>(a, 1, st);
}
// gW (+)= dY^T X, gb (+)= colsum(dY)
int var_wgrad(const float* dY, const float* X, long long T, int Nw, int Kw, float* gW, float* gb, int accumulate, cudaStream_t st) {
    GemmArgs a{};
    a.A = dY; a.B = X; a.C = gW; a.M = Nw; a.N = Kw; a.K = (int)T; a.lda = Nw; a.ldb = Kw; a.ldc = Kw; a.accumulate = accumulate;
    int rc = var_gemm<G_TN, E_NONE>(a, 1, st);
    if (rc) return rc;
    if (gb) {
        var_colsum_kernel<<<dtqn_cdiv(Nw, 32), 256, 0, st>>>(dY, T, Nw, gb, accumulate);
        DTQN_LAUNCH_CHECK();
    }
    return 0;
}
}  // namespace

long long var_bwd_scratch_floats(const dtqn_net_cfg& c, long long T0) {
    VarBwd s;
    return var_bwd_layout(c, T0, nullptr, s);
}

// dq [T0, A] = dLoss/dQ of group 0 (td_loss_kernel); activations of the 3-group forward (save = 1) in ws; grads zeroed by the caller.
int var_td_backward(const dtqn_net_cfg& c, const NetLayout& lay, const float* params, const dtqn_obs_src* obs_src,
                    const float* dq, int B, int L, float* ws, long long ws_floats, float* scratch, float* grads,
                    cudaStream_t st) {
    const int d = c.d_model, H = c.n_heads, hd = d / H, A = c.num_actions;
    const long long T0 = (long long)B * L, n = T0 * d;
    VarAct act;
    if (var_act_layout(c, 3 * T0, ws, act) > ws_floats) return DTQN_E_ARG;
    VarBwd s;
    var_bwd_layout(c, T0, scratch, s);
    const bool train = obs_src->train_mode != 0;
    auto drop = [&](unsigned site) { return Drop{c.dropout > 0.f && train ? (const unsigned long long*)c.dropout_state : nullptr, c.dropout, site}; };
    int rc;
    // GRU gate backward: given du, returns dx in `dx` (overwritten) and dy in `dy` (overwritten); gate gradients accumulate
    // over the layers that share the gate (`first` = first application in this backward pass)
    auto gate_bwd = [&](int k, bool first, const float* du, const float* x, const float* y, const VarGate& gs, float* dx, float* dy) -> int {
        if (!c.gate_gru) {                                     // ResGate: both inputs receive du
            cudaError_t e1 = cudaMemcpyAsync(dx, du, sizeof(float) * n, cudaMemcpyDeviceToDevice, st);
            cudaError_t e2 = cudaMemcpyAsync(dy, du, sizeof(float) * n, cudaMemcpyDeviceToDevice, st);
            return e1 != cudaSuccess ? (int)e1 : (int)e2;
        }
        const GateOff& q = lay.gate[k];
        const int acc = first ? 0 : 1;
        float *dag = s.gc, *daz = s.gd, *dar = s.t1;           // t1 doubles as drx before it becomes dar
        var_gru_bwd1_kernel<<<grid1(n), 256, 0, st>>>(du, gs.z, gs.hg, x, n, dag, daz, dx);
        DTQN_LAUNCH_CHECK();
        if ((rc = var_wgrad(dag, y, T0, d, d, grads + q.w_g, nullptr, acc, st))) return rc;
        if ((rc = var_wgrad(dag, gs.rx, T0, d, d, grads + q.u_g, nullptr, acc, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(dag, params + q.w_g, nullptr, dy, T0, d, d, 0, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(dag, params + q.u_g, nullptr, s.t1, T0, d, d, 0, st))) return rc;   // drx
        var_gru_bwd2_kernel<<<grid1(n), 256, 0, st>>>(s.t1, gs.r, x, n, dar, dx);
        DTQN_LAUNCH_CHECK();
        if ((rc = var_wgrad(dar, y, T0, d, d, grads + q.w_r, nullptr, acc, st))) return rc;
        if ((rc = var_wgrad(dar, x, T0, d, d, grads + q.u_r, nullptr, acc, st))) return rc;
        if ((rc = var_wgrad(daz, y, T0, d, d, grads + q.w_z, grads + q.b_z, acc, st))) return rc;
        if ((rc = var_wgrad(daz, x, T0, d, d, grads + q.u_z, nullptr, acc, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(dar, params + q.w_r, nullptr, dy, T0, d, d, 1, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(daz, params + q.w_z, nullptr, dy, T0, d, d, 1, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(dar, params + q.u_r, nullptr, dx, T0, d, d, 1, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(daz, params + q.u_z, nullptr, dx, T0, d, d, 1, st))) return rc;
        return 0;
    };
    // LayerNorm backward: du (= or +=) from dy; dgamma / dbeta written (each LayerNorm is applied once per layer)
    auto ln_bwd = [&](const float* dy, const float* u, const float* stt, long long g_off, long long b_off, float* du, int accumulate) -> int {
        var_ln_bwd_kernel<<<(unsigned)dtqn_cdiv(T0, 8), 256, 0, st>>>(dy, u, stt, params + g_off, T0, d, du, s.t1, accumulate);
        DTQN_LAUNCH_CHECK();
        var_colsum_kernel<<<dtqn_cdiv(d, 32), 256, 0, st>>>(s.t1, T0, d, grads + g_off, 0);
        var_colsum_kernel<<<dtqn_cdiv(d, 32), 256, 0, st>>>(dy, T0, d, grads + b_off, 0);
        DTQN_LAUNCH_CHECK();
        return 0;
    };
    // ---- head ----
    const float* x_last = c.identity ? act.layer[c.n_layers - 1].u2 : act.layer[c.n_layers - 1].lnB;
    prof_begin(PROF_HEAD, st);
    var_head_bwd_kernel<<<grid1(n), 256, 0, st>>>(dq, act.hh, params + lay.h2_w, T0, d, A, s.g_hh, grads + lay.h2_w, grads + lay.h2_b);
    prof_end(PROF_HEAD, st, 0.0);
    DTQN_LAUNCH_CHECK();
    if ((rc = var_wgrad(s.g_hh, x_last, T0, d, d, grads + lay.h1_w, grads + lay.h1_b, 0, st))) return rc;
    if ((rc = var_dgrad<E_NONE>(s.g_hh, params + lay.h1_w, nullptr, s.gx, T0, d, d, 0, st))) return rc;
    // ---- layers, last to first; s.gx = gradient w.r.t. the layer's output ----
    for (int li = c.n_layers - 1; li >= 0; --li) {
        const LayerOff& lo = lay.layer[li];
        const VarLayer& la = act.layer[li];
        const float* x_in = li == 0 ? act.x0 : (c.identity ? act.layer[li - 1].u2 : act.layer[li - 1].lnB);
        const bool first = li == c.n_layers - 1;
        const float drop_scale = (c.dropout > 0.f && train) ? 1.f / (1.f - c.dropout) : 1.f;
        // s.gu <- gradient w.r.t. u2 ; s.ga <- skip part, s.gb <- d r2
        const float* g_u2 = s.gx;
        if (!c.identity) {                                     // out = LN2(u2)
            if ((rc = ln_bwd(s.gx, la.u2, la.stB, lo.ln2_w, lo.ln2_b, s.gu, 0))) return rc;
            g_u2 = s.gu;
        }
        const float* skip_in = c.identity ? la.u1 : la.lnA;   // x of the mlp gate
        if ((rc = gate_bwd(1, first, g_u2, skip_in, la.r2, la.g2, s.ga, s.gb))) return rc;
        // r2 = relu(f) * mask / (1 - p): d f = d r2 * [r2 > 0] / (1 - p)
        var_relu_bwd_kernel<<<grid1(n), 256, 0, st>>>(s.gb, la.r2, n, drop_scale, s.gb);
        DTQN_LAUNCH_CHECK();
        if ((rc = var_wgrad(s.gb, la.h, T0, d, 4 * d, grads + lo.f2_w, grads + lo.f2_b, 0, st))) return rc;
        if ((rc = var_dgrad<E_MASK>(s.gb, params + lo.f2_w, la.h, s.gh, T0, d, 4 * d, 0, st))) return rc;
        const float* ffn_in = c.identity ? la.lnB : la.lnA;
        if ((rc = var_wgrad(s.gh, ffn_in, T0, 4 * d, d, grads + lo.f1_w, grads + lo.f1_b, 0, st))) return rc;
        // gradient w.r.t. u1 in s.gu
        if (c.identity) {                                      // ffn_in = LN2(u1); u1 also feeds the gate's skip input (s.ga)
            if ((rc = var_dgrad<E_NONE>(s.gh, params + lo.f1_w, nullptr, s.gb, T0, 4 * d, d, 0, st))) return rc;
            if (cudaError_t ce = cudaMemcpyAsync(s.gu, s.ga, sizeof(float) * n, cudaMemcpyDeviceToDevice, st)) return (int)ce;
            if ((rc = ln_bwd(s.gb, la.u1, la.stB, lo.ln2_w, lo.ln2_b, s.gu, 1))) return rc;
        } else {                                               // ffn_in = lnA = LN1(u1), which is also the gate's skip input
            if ((rc = var_dgrad<E_ADD>(s.gh, params + lo.f1_w, s.ga, s.gb, T0, 4 * d, d, 0, st))) return rc;
            if ((rc = ln_bwd(s.gb, la.u1, la.stA, lo.ln1_w, lo.ln1_b, s.gu, 0))) return rc;
        }
        // attention gate: u1 = gate(x_in, r1)
        if ((rc = gate_bwd(0, first, s.gu, x_in, la.r1, la.g1, s.ga, s.gb))) return rc;   // s.ga: skip to x_in; s.gb: d r1
        var_relu_bwd_kernel<<<grid1(n), 256, 0, st>>>(s.gb, la.r1, n, 1.f, s.gb);
        DTQN_LAUNCH_CHECK();
        if ((rc = var_wgrad(s.gb, la.o, T0, d, d, grads + lo.out_w, grads + lo.out_b, 0, st))) return rc;
        if ((rc = var_dgrad<E_NONE>(s.gb, params + lo.out_w, nullptr, s.gu, T0, d, d, 0, st))) return rc;   // d o
        {
            dim3 grid(H, (unsigned)B);
            const int thr = L <= 64 ? 64 : 128;
            const float scale = 1.0f / sqrtf((float)hd);
            const Drop dr = drop(2 + 4 * li);
            prof_begin(PROF_ATTN_BWD, st);
            if (hd == 4) var_attn_bwd_kernel<4><<<grid, thr, 0, st>>>(la.qkv, la.o, s.gu, s.gqkv, L, d, scale, dr, train);
            else if (hd == 8) var_attn_bwd_kernel<8><<<grid, thr, 0, st>>>(la.qkv, la.o, s.gu, s.gqkv, L, d, scale, dr, train);
            else var_attn_bwd_kernel<16><<<grid, thr, 0, st>>>(la.qkv, la.o, s.gu, s.gqkv, L, d, scale, dr, train);
            prof_end(PROF_ATTN_BWD, st, 8.0 * (double)T0 * L * d);
            DTQN_LAUNCH_CHECK();
        }
        const float* att_in = c.identity ? la.lnA : x_in;
        if ((rc = var_wgrad(s.gqkv, att_in, T0, 3 * d, d, grads + lo.in_w, grads + lo.in_b, 0, st))) return rc;
        if (c.identity) {                                      // att_in = LN1(x_in); x_in also gets the gate's skip gradient
            if ((rc = var_dgrad<E_NONE>(s.gqkv, params + lo.in_w, nullptr, s.gb, T0, 3 * d, d, 0, st))) return rc;
            if (cudaError_t ce = cudaMemcpyAsync(s.gx, s.ga, sizeof(float) * n, cudaMemcpyDeviceToDevice, st)) return (int)ce;
            if ((rc = ln_bwd(s.gb, x_in, la.stA, lo.ln1_w, lo.ln1_b, s.gx, 1))) return rc;
        } else {
            if ((rc = var_dgrad<E_ADD>(s.gqkv, params + lo.in_w, s.ga, s.gx, T0, 3 * d, d, 0, st))) return rc;
        }
    }
    // ---- embedding: s.gx = d (dropout(token + pos)) ----
    prof_begin(PROF_OTHER, st);
    if (c.dropout > 0.f && train) {
        var_drop_bwd_kernel<<<grid1(n), 256, 0, st>>>(s.gx, n, drop(1));
        DTQN_LAUNCH_CHECK();
    }
    if (c.pos_trainable) {
        var_pos_bwd_kernel<<<dtqn_cdiv((long long)L * d, 256), 256, 0, st>>>(s.gx, B, L, d, grads + lay.pos);
        DTQN_LAUNCH_CHECK();
    }
    {
        const int n_el = c.num_actions * c.action_dim + (d - c.action_dim) * (lay.k_in + 1) + (c.discrete ? c.vocab * c.embed_per_obs : 0);
        var_embed_bwd_kernel<<<dtqn_cdiv(n_el, 256), 256, 0, st>>>(s.gx, *obs_src, c, lay, params, B, L, grads);
        DTQN_LAUNCH_CHECK();
    }
    prof_end(PROF_OTHER, st, 0.0);
    if (c.dropout > 0.f) {                                     // the training step is over: next forward draws new masks
        var_bump_kernel<<<1, 1, 0, st>>>((unsigned long long*)c.dropout_state);
        DTQN_LAUNCH_CHECK();
    }
    return 0;
}
