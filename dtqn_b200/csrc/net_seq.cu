// Sequence-resident DTQN forward for the small, latency-bound groups (the 3 x 32 sequences of a training step): ONE kernel
// launch runs every transformer layer and the Q head for one sequence per CTA, with the activations of the sequence held
// in shared memory between the GEMMs (d_model = 64, 8 heads, L <= 64).  It replaces ~13 dependent launches of tiny GEMM /
// attention kernels whose cost was launch + tail latency, not math.  Same arithmetic as net_fwd.cu (fp32 CUDA cores, the
// same epilogues: bias | bias+ReLU | bias -> ReLU -> +residual -> LayerNorm; transformer.py:63-78, dtqn.py:216), and it
// saves the same activations for the backward pass.
#include "net.cuh"
#include "prof.cuh"

namespace {

constexpr int SQ_D = 64, SQ_H = 8, SQ_HD = 8, SQ_ROWS = 64, SQ_THREADS = 256;
constexpr int LDX = SQ_D + 1;          // sX / sO row stride (floats): +1 -> the two row groups of a warp hit different banks
constexpr int LDQ = 3 * SQ_D + 4;      // sQKV row stride: float4-aligned rows for the attention reads
constexpr int LDH = 4 * SQ_D + 1;      // sH row stride
constexpr int SW_LD = 256 + 4;         // weight slab row stride (max N = 256)

constexpr int SW_STAGES = 3;           // cp.async ring of weight slabs
struct SeqSmem {
    float x[SQ_ROWS * LDX];            // layer input (residual of LN1)
    float o[SQ_ROWS * LDX];            // attention output, then x1 (residual of LN2)
    float qkv[(SQ_ROWS + 4) * LDQ];    // packed q | k | v; 4 zero rows so a key block may run past the last key
    float h[SQ_ROWS * LDH];            // FFN hidden
    float w[SW_STAGES * 16 * SW_LD];   // weight slabs [16 k][N] (k-major), 3-stage ring
};

__device__ __forceinline__ int sq_col(int tx, int j) { return (j >> 2) * 64 + tx * 4 + (j & 3); }

// acc[i][j] (row ty*4+i, col sq_col(tx,j)) = sum_k sA[row][k] * W[col][k];  W row-major [N, K] in global memory.
template <int N>
__device__ __forceinline__ void cta_gemm64(const float* __restrict__ sA, int lda, const float* __restrict__ W, int K,
                                           float (&acc)[4][N / 16], float* __restrict__ sW) {
    constexpr int TN = N / 16, WV = N / 64;                   // float4 loads of a [N x 16] slab per thread
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float4 wr[WV];
    auto load = [&](int k0) {
#pragma unroll
        for (int v = 0; v < WV; ++v) {
            const int idx = tid + v * SQ_THREADS, nr = idx >> 2, kq = idx & 3;
            wr[v] = __ldg(reinterpret_cast<const float4*>(W + (size_t)nr * K + k0 + kq * 4));
        }
    };
    load(0);
    for (int k0 = 0; k0 < K; k0 += 16) {
        __syncthreads();                                      // previous slab fully consumed (also orders sA writes before reads)
#pragma unroll
        for (int v = 0; v < WV; ++v) {
            const int idx = tid + v * SQ_THREADS, nr = idx >> 2, kq = idx & 3;
            sW[(kq * 4 + 0) * SW_LD + nr] = wr[v].x; sW[(kq * 4 + 1) * SW_LD + nr] = wr[v].y;
            sW[(kq * 4 + 2) * SW_LD + nr] = wr[v].z; sW[(kq * 4 + 3) * SW_LD + nr] = wr[v].w;
        }
        __syncthreads();
        if (k0 + 16 < K) load(k0 + 16);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[TN];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[(ty * 4 + i) * lda + k0 + kk];
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 bv = *reinterpret_cast<const float4*>(&sW[kk * SW_LD + j4 * 64 + tx * 4]);
                b[j4 * 4] = bv.x; b[j4 * 4 + 1] = bv.y; b[j4 * 4 + 2] = bv.z; b[j4 * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
}

// Same contraction with the weights read from the K-MAJOR copy WT[K][N] (pack_weights_kernel): slabs of 16 k-rows are
// contiguous rows of N floats, so they stream through a 3-stage cp.async ring (two slabs in flight while one is consumed,
// one __syncthreads per slab) instead of load -> transpose -> store with the latency of every slab exposed.  The FMA order
// per output element is unchanged (k ascending), so the results are bit-identical to cta_gemm64.
__device__ __forceinline__ void sq_cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int N>
__device__ __forceinline__ void cta_gemm64_pipe(const float* __restrict__ sA, int lda, const float* __restrict__ WT, int K,
                                                float (&acc)[4][N / 16], float* __restrict__ sW) {
    constexpr int TN = N / 16, WV = N / 64;                   // 16-byte copies of a [16 x N] slab per thread
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const int slabs = K / 16;
    auto issue = [&](int s) {
        if (s < slabs) {
            float* dst = sW + (s % SW_STAGES) * (16 * SW_LD);
            const float* src = WT + (size_t)s * 16 * N;
#pragma unroll
            for (int v = 0; v < WV; ++v) {
                const int idx = tid + v * SQ_THREADS, kr = idx / (N / 4), nq = idx % (N / 4);
                sq_cp_async16(dst + kr * SW_LD + nq * 4, src + (size_t)kr * N + nq * 4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    issue(1);
    for (int s = 0; s < slabs; ++s) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");    // slab s has landed (this thread's copies)
        __syncthreads();                                      // ... everybody's; slab s - 1 fully consumed; sA writes ordered
        issue(s + 2);                                         // refill the slot slab s - 1 used
        const float* sWs = sW + (s % SW_STAGES) * (16 * SW_LD);
        const int k0 = s * 16;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[TN];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[(ty * 4 + i) * lda + k0 + kk];
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 bv = *reinterpret_cast<const float4*>(&sWs[kk * SW_LD + j4 * 64 + tx * 4]);
                b[j4 * 4] = bv.x; b[j4 * 4 + 1] = bv.y; b[j4 * 4 + 2] = bv.z; b[j4 * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float sq_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Causal attention of one sequence from the packed q|k|v tile in shared memory; warp = head, lane = row pair (r, L-1-r),
// warp-uniform key blocks of 4 (broadcast reads), online softmax in base 2 (see attn_seq_kernel in net_fwd.cu).
__device__ __forceinline__ void sq_attention(const float* __restrict__ sQKV, int L, float scale, float* __restrict__ sO,
                                             float* __restrict__ g_o /* nullable: global o rows of this sequence */) {
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = (L + 1) / 2;
    scale *= 1.4426950408889634f;
    const int r = lane;
    const bool act = r < half;
    const int ra = act ? r : -1, rb = act ? L - 1 - r : -1;
    float qa[SQ_HD], qb[SQ_HD], aa[SQ_HD], ab[SQ_HD];
    {
        const float* pa = sQKV + max(ra, 0) * LDQ + h * SQ_HD;
        const float* pb = sQKV + max(rb, 0) * LDQ + h * SQ_HD;
#pragma unroll
        for (int c = 0; c < SQ_HD; ++c) { qa[c] = pa[c] * scale; qb[c] = pb[c] * scale; aa[c] = 0.f; ab[c] = 0.f; }
    }
    float ma = -INFINITY, la_ = 0.f, mb = -INFINITY, lb = 0.f;
    const int last_a = half - 1;
    const float* kp = sQKV + SQ_D + h * SQ_HD;
    const float* vp = sQKV + 2 * SQ_D + h * SQ_HD;
    int rem_a = ra + 1, rem_b = rb + 1;
    for (int j0 = 0; j0 < L; j0 += 4, kp += 4 * LDQ, vp += 4 * LDQ, rem_a -= 4, rem_b -= 4) {
        const bool do_a = j0 <= last_a;
        float sa[4], sb[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float da = 0.f, db = 0.f;
#pragma unroll
            for (int c = 0; c < SQ_HD; c += 4) {
                const float4 kv = *reinterpret_cast<const float4*>(kp + jj * LDQ + c);
                db = fmaf(qb[c], kv.x, db); db = fmaf(qb[c + 1], kv.y, db); db = fmaf(qb[c + 2], kv.z, db); db = fmaf(qb[c + 3], kv.w, db);
                if (do_a) { da = fmaf(qa[c], kv.x, da); da = fmaf(qa[c + 1], kv.y, da); da = fmaf(qa[c + 2], kv.z, da); da = fmaf(qa[c + 3], kv.w, da); }
            }
            sa[jj] = (jj < rem_a) ? da : -INFINITY;
            sb[jj] = (jj < rem_b) ? db : -INFINITY;
        }
        const float nmb = fmaxf(mb, fmaxf(fmaxf(sb[0], sb[1]), fmaxf(sb[2], sb[3])));
        const float sfb = (nmb == -INFINITY) ? 0.f : nmb;
        const float cb = sq_ex2(mb - sfb);
        float pb_[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) pb_[jj] = sq_ex2(sb[jj] - sfb);
        lb = lb * cb + (pb_[0] + pb_[1]) + (pb_[2] + pb_[3]);
        mb = nmb;
        float ca = 1.f, pa_[4] = {0.f, 0.f, 0.f, 0.f};
        if (do_a) {
            const float nma = fmaxf(ma, fmaxf(fmaxf(sa[0], sa[1]), fmaxf(sa[2], sa[3])));
            const float sfa = (nma == -INFINITY) ? 0.f : nma;
            ca = sq_ex2(ma - sfa);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) pa_[jj] = sq_ex2(sa[jj] - sfa);
            la_ = la_ * ca + (pa_[0] + pa_[1]) + (pa_[2] + pa_[3]);
            ma = nma;
        }
#pragma unroll
        for (int c = 0; c < SQ_HD; ++c) { ab[c] *= cb; if (do_a) aa[c] *= ca; }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
            for (int c = 0; c < SQ_HD; c += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(vp + jj * LDQ + c);
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
#pragma unroll
        for (int c = 0; c < SQ_HD; ++c) {
            const float v = ab[c] * ib;
            sO[rb * LDX + h * SQ_HD + c] = v;
            if (g_o) g_o[(size_t)rb * SQ_D + h * SQ_HD + c] = v;
        }
        if (ra != rb) {
            const float ia = 1.f / la_;
#pragma unroll
            for (int c = 0; c < SQ_HD; ++c) {
                const float v = aa[c] * ia;
                sO[ra * LDX + h * SQ_HD + c] = v;
                if (g_o) g_o[(size_t)ra * SQ_D + h * SQ_HD + c] = v;
            }
        }
    }
}

struct SeqFwdArgs {
    GroupPtrs P;
    NetLayout lay;
    NetAct act;                        // global activation buffers ([G * n_seq * L, width]); x0 is the input
    int n_layers, n_seq, L, A, save;
    float* q_out;                      // [G, n_seq, L, A]
    const float* wt[DTQN_MAX_GROUPS];  // k-major weight copies (nullable): layer i at 12 d^2 i: in | out | ffn.0 | ffn.2; head ffn.0 last
};

// y = LayerNorm(xres + relu(acc + bias)) for the 4 rows x 4 cols this thread holds of a [64 x 64] tile; the 16 lanes that
// share a row reduce with shuffles.  Writes y to sY (smem, stride LDX) and, for rows < L, to the global buffers.
__device__ __forceinline__ void sq_res_ln(const float (&acc)[4][4], const float* __restrict__ p, long long b_off,
                                          long long g_off, long long be_off, const float* __restrict__ sRes,
                                          float* __restrict__ sY, int L, float* __restrict__ g_y, float* __restrict__ g_r,
                                          float* __restrict__ g_st) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float bias[4], gam[4], bet[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bias[j] = __ldg(p + b_off + tx * 4 + j); gam[j] = __ldg(p + g_off + tx * 4 + j); bet[j] = __ldg(p + be_off + tx * 4 + j);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = ty * 4 + i;
        float u[4], rl[4], s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rl[j] = fmaxf(acc[i][j] + bias[j], 0.f);
            u[j] = sRes[row * LDX + tx * 4 + j] + rl[j];
            s += u[j];
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / SQ_D);
        float vs = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float dl = u[j] - mean; vs = fmaf(dl, dl, vs); }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
        const float rstd = 1.0f / sqrtf(vs * (1.f / SQ_D) + 1e-5f);
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { y[j] = (u[j] - mean) * rstd * gam[j] + bet[j]; sY[row * LDX + tx * 4 + j] = y[j]; }
        if (row < L) {
            *reinterpret_cast<float4*>(g_y + (size_t)row * SQ_D + tx * 4) = make_float4(y[0], y[1], y[2], y[3]);
            if (g_r) *reinterpret_cast<float4*>(g_r + (size_t)row * SQ_D + tx * 4) = make_float4(rl[0], rl[1], rl[2], rl[3]);
            if (g_st && tx == 0) { g_st[row * 2] = mean; g_st[row * 2 + 1] = rstd; }
        }
    }
}

__global__ void __launch_bounds__(SQ_THREADS, 1)
seq_forward_kernel(SeqFwdArgs a) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t sq_raw[];
    SeqSmem& sm = *reinterpret_cast<SeqSmem*>(sq_raw);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int seq = blockIdx.x, g = seq / a.n_seq, L = a.L;
    const float* p = a.P.p[g];
    const float* wt = a.wt[g];                                // k-major weight copies, or NULL
    constexpr int D2 = SQ_D * SQ_D;
    const size_t t0 = (size_t)seq * L;                        // first token row of this sequence in every activation buffer
    const float scale = 1.0f / sqrtf((float)SQ_HD);

    // x0 rows -> sX (rows >= L zero), zero the padding rows of the qkv tile
    for (int e = tid; e < SQ_ROWS * SQ_D; e += SQ_THREADS) {
        const int r = e / SQ_D, c = e % SQ_D;
        sm.x[r * LDX + c] = r < L ? a.act.x0[(t0 + r) * SQ_D + c] : 0.f;
    }
    for (int e = tid; e < 4 * LDQ; e += SQ_THREADS) sm.qkv[SQ_ROWS * LDQ + e] = 0.f;
    for (int e = tid; e < SQ_ROWS * LDX; e += SQ_THREADS) sm.o[e] = 0.f;      // rows >= L are never written by attention

    for (int li = 0; li < a.n_layers; ++li) {
        const LayerOff& lo = a.lay.layer[li];
        const LayerAct& la = a.act.layer[li];
        // ---- in_proj: qkv = x W_in^T + b_in ----
        {
            float acc[4][12];
            if (wt) cta_gemm64_pipe<192>(sm.x, LDX, wt + (size_t)li * 12 * D2, SQ_D, acc, sm.w);
            else cta_gemm64<192>(sm.x, LDX, p + lo.in_w, SQ_D, acc, sm.w);
#pragma unroll
            for (int j4 = 0; j4 < 3; ++j4) {
                float bias[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) bias[e] = __ldg(p + lo.in_b + j4 * 64 + tx * 4 + e);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = ty * 4 + i;
                    const float4 v = make_float4(acc[i][j4 * 4] + bias[0], acc[i][j4 * 4 + 1] + bias[1],
                                                 acc[i][j4 * 4 + 2] + bias[2], acc[i][j4 * 4 + 3] + bias[3]);
                    *reinterpret_cast<float4*>(&sm.qkv[row * LDQ + j4 * 64 + tx * 4]) = v;
                    if (a.save && row < L) *reinterpret_cast<float4*>(la.qkv + (t0 + row) * (size_t)(3 * SQ_D) + j4 * 64 + tx * 4) = v;
                }
            }
        }
        __syncthreads();
        // ---- attention core ----
        sq_attention(sm.qkv, L, scale, sm.o, a.save ? la.o + t0 * SQ_D : nullptr);
        __syncthreads();
        // ---- out_proj -> relu -> +x -> LN1  (x1 overwrites the attention output tile) ----
        {
            float acc[4][4];
            if (wt) cta_gemm64_pipe<64>(sm.o, LDX, wt + (size_t)li * 12 * D2 + 3 * D2, SQ_D, acc, sm.w);
            else cta_gemm64<64>(sm.o, LDX, p + lo.out_w, SQ_D, acc, sm.w);
            __syncthreads();                                  // every thread is done reading sm.o as the A operand
            sq_res_ln(acc, p, lo.out_b, lo.ln1_w, lo.ln1_b, sm.x, sm.o, L, la.x1 + t0 * SQ_D,
                      a.save ? la.r1 + t0 * SQ_D : nullptr, a.save ? la.st1 + t0 * 2 : nullptr);
        }
        __syncthreads();
        // ---- ffn.0 + relu -> h ----
        {
            float acc[4][16];
            if (wt) cta_gemm64_pipe<256>(sm.o, LDX, wt + (size_t)li * 12 * D2 + 4 * D2, SQ_D, acc, sm.w);
            else cta_gemm64<256>(sm.o, LDX, p + lo.f1_w, SQ_D, acc, sm.w);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float bias[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) bias[e] = __ldg(p + lo.f1_b + j4 * 64 + tx * 4 + e);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = ty * 4 + i;
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { v[e] = fmaxf(acc[i][j4 * 4 + e] + bias[e], 0.f); sm.h[row * LDH + j4 * 64 + tx * 4 + e] = v[e]; }
                    if (a.save && row < L) *reinterpret_cast<float4*>(la.h + (t0 + row) * (size_t)(4 * SQ_D) + j4 * 64 + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        __syncthreads();
        // ---- ffn.2 -> relu -> +x1 -> LN2 -> next layer input (sX) ----
        {
            float acc[4][4];
            if (wt) cta_gemm64_pipe<64>(sm.h, LDH, wt + (size_t)li * 12 * D2 + 8 * D2, 4 * SQ_D, acc, sm.w);
            else cta_gemm64<64>(sm.h, LDH, p + lo.f2_w, 4 * SQ_D, acc, sm.w);
            sq_res_ln(acc, p, lo.f2_b, lo.ln2_w, lo.ln2_b, sm.o, sm.x, L, la.x2 + t0 * SQ_D,
                      a.save ? la.r2 + t0 * SQ_D : nullptr, a.save ? la.st2 + t0 * 2 : nullptr);
        }
        __syncthreads();
    }
    // ---- Q head: hh = relu(x W1^T + b1); q = hh W2^T + b2 ----
    {
        float acc[4][4];
        if (wt) cta_gemm64_pipe<64>(sm.x, LDX, wt + (size_t)a.n_layers * 12 * D2, SQ_D, acc, sm.w);
        else cta_gemm64<64>(sm.x, LDX, p + a.lay.h1_w, SQ_D, acc, sm.w);
        float bias[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) bias[e] = __ldg(p + a.lay.h1_b + tx * 4 + e);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = ty * 4 + i;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { v[e] = fmaxf(acc[i][e] + bias[e], 0.f); sm.o[row * LDX + tx * 4 + e] = v[e]; }
            if (row < L) *reinterpret_cast<float4*>(a.act.hh + (t0 + row) * SQ_D + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();
    for (int e = tid; e < L * a.A; e += SQ_THREADS) {
        const int row = e / a.A, ac = e % a.A;
        const float* w = p + a.lay.h2_w + (size_t)ac * SQ_D;
        float s = __ldg(p + a.lay.h2_b + ac);
        for (int k = 0; k < SQ_D; ++k) s = fmaf(sm.o[row * LDX + k], __ldg(w + k), s);
        a.q_out[(t0 + row) * a.A + ac] = s;
    }
}

}  // namespace

bool seq_forward_supported(const dtqn_net_cfg& c, int L) {
    return c.d_model == SQ_D && c.n_heads == SQ_H && L <= SQ_ROWS && L >= 1;
}

int launch_seq_forward(const dtqn_net_cfg& c, const NetLayout& lay, const NetAct& act, const GroupPtrs& P, int G, int n_seq,
                       int L, int save, float* q_out, cudaStream_t st, const float* const* wt) {
    SeqFwdArgs a{};
    a.P = P; a.lay = lay; a.act = act; a.n_layers = c.n_layers; a.n_seq = n_seq; a.L = L; a.A = c.num_actions; a.save = save;
    a.q_out = q_out;
    for (int g = 0; g < G; ++g) a.wt[g] = wt ? wt[g] : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(seq_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    prof_begin(PROF_SEQ_FWD, st);
    launch_k(seq_forward_kernel, G * n_seq, SQ_THREADS, sizeof(SeqSmem), st, a);
    prof_end(PROF_SEQ_FWD, st, 2.0 * (double)G * n_seq * L * (c.n_layers * 12.0 * SQ_D * SQ_D + SQ_D * SQ_D + SQ_D * c.num_actions));
    DTQN_LAUNCH_CHECK();
    return 0;
}
