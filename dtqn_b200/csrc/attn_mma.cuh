// Causal self-attention core of one (sequence, head) on the warp-level tensor-core path (mma.sync m16n8k8 TF32), for
// head_dim = 8 and L <= 64 (nn.MultiheadAttention inside transformer.py:64-70, additive -inf mask :49-53).
//
// fp32 parity: every operand is split x = hi + lo (two TF32 values, ~20 mantissa bits together) and three MMAs per
// product accumulate lo*hi + hi*lo + hi*hi in fp32 -- the same trick as the tcgen05 Linear, at TF32 granularity.
//
// Data layout: the sequence's packed in_proj output lives in shared memory as fp32 rows [position][q(64) | k(64) | v(64)]
// with a row stride of ATT_LD = 196 floats (== 4 mod 32), which makes every fragment load below conflict-free:
//   Q (A operand, row-major 16 x 8):  a0 = Q[g][t]  a1 = Q[g+8][t]  a2 = Q[g][t+4]  a3 = Q[g+8][t+4]     (g = lane / 4, t = lane % 4)
//   K (B operand, 8 x 8 "col"):       b0 = K[key g][t]  b1 = K[key g][t+4]
//   S (C fragment, 16 x 8):           c0 = S[g][2t]  c1 = S[g][2t+1]  c2 = S[g+8][2t]  c3 = S[g+8][2t+1]
// The probabilities feed the second product straight from the C fragment: the summation index of P V is relabelled
// (A column t <-> key 2t, column t+4 <-> key 2t+1), so a0 = c0, a1 = c2, a2 = c1, a3 = c3 and the V fragment is read with
// the same permutation (b0 = V[key 2t][g], b1 = V[key 2t+1][g]) -- no shuffles between the two GEMMs.
// q must already be scaled by log2(e) / sqrt(head_dim): the softmax runs in base 2 on the MUFU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ATT_LD = 196;            // floats per staged qkv row (3 * 64 + 4)
constexpr int ATT_HD = 8;

__device__ __forceinline__ float att_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x = hi + lo: hi keeps the 19 bits a TF32 operand has (the tensor core ignores the low 13 mantissa bits of a register, so
// truncation here is what the MMA would do anyway); lo = x - hi is exact in fp32 and is itself read truncated (2^-20 of x).
__device__ __forceinline__ void att_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void att_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One 16-row query tile MT of head h.  sq: row 0 of the sequence in the staged qkv image.  Keys 0..min(16 MT + 15, L - 1)
// are visited in blocks of 8; rows/keys beyond L only have to be finite (they are never read back).
// out(row, col, v0, v1): called once per thread for (row g, cols 2t, 2t+1) and once for row g + 8 with normalised outputs.
template <int MT, typename Out>
__device__ __forceinline__ void att_mtile(const float* __restrict__ sq, int h, int L, int lane, Out&& out) {
    constexpr int NB = 2 * MT + 2;                      // key blocks under the diagonal of this tile
    const int g = lane >> 2, t = lane & 3;
    const int nb_l = (L + 7) >> 3;                      // blocks that hold at least one real key (warp-uniform)
    const int r0 = 16 * MT;
    uint32_t qh[4], ql[4];
    {
        const float* q0 = sq + (r0 + g) * ATT_LD + h * ATT_HD + t;
        att_split(q0[0], qh[0], ql[0]);
        att_split(q0[8 * ATT_LD], qh[1], ql[1]);
        att_split(q0[4], qh[2], ql[2]);
        att_split(q0[8 * ATT_LD + 4], qh[3], ql[3]);
    }
    float s[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        s[b][0] = s[b][1] = s[b][2] = s[b][3] = 0.f;
        if (b < nb_l) {
            const float* kp = sq + (8 * b + g) * ATT_LD + 64 + h * ATT_HD + t;
            uint32_t kh0, kl0, kh1, kl1;
            att_split(kp[0], kh0, kl0);
            att_split(kp[4], kh1, kl1);
            att_mma(s[b], ql, kh0, kh1);
            att_mma(s[b], qh, kl0, kl1);
            att_mma(s[b], qh, kh0, kh1);
        }
    }
    // causal mask (additive -inf above the diagonal) + row maxima; only the last two blocks touch the diagonal
    float m0 = -INFINITY, m1 = -INFINITY;               // rows g and g + 8
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < nb_l) {
            if (b >= NB - 2) {
                const int key = 8 * b + 2 * t, ra = r0 + g, rb = r0 + g + 8;
                if (key > ra) s[b][0] = -INFINITY;
                if (key + 1 > ra) s[b][1] = -INFINITY;
                if (key > rb) s[b][2] = -INFINITY;
                if (key + 1 > rb) s[b][3] = -INFINITY;
            }
            m0 = fmaxf(m0, fmaxf(s[b][0], s[b][1]));
            m1 = fmaxf(m1, fmaxf(s[b][2], s[b][3]));
        }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // three independent accumulators (lo*hi, hi*lo, hi*hi): the P V products of a tile are one dependent chain otherwise
    float l0 = 0.f, l1 = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f}, o_lh[4] = {0.f, 0.f, 0.f, 0.f}, o_hl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < nb_l) {
            const float p0 = att_ex2(s[b][0] - m0), p1 = att_ex2(s[b][1] - m0);
            const float p2 = att_ex2(s[b][2] - m1), p3 = att_ex2(s[b][3] - m1);
            l0 += p0 + p1; l1 += p2 + p3;
            uint32_t ph[4], pl[4];
            att_split(p0, ph[0], pl[0]); att_split(p2, ph[1], pl[1]);
            att_split(p1, ph[2], pl[2]); att_split(p3, ph[3], pl[3]);
            const float* vp = sq + (8 * b + 2 * t) * ATT_LD + 128 + h * ATT_HD + g;
            uint32_t vh0, vl0, vh1, vl1;
            att_split(vp[0], vh0, vl0);
            att_split(vp[ATT_LD], vh1, vl1);
            att_mma(o_lh, pl, vh0, vh1);
            att_mma(o_hl, ph, vl0, vl1);
            att_mma(o, ph, vh0, vh1);
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] += o_lh[e] + o_hl[e];
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    out(r0 + g, h * ATT_HD + 2 * t, o[0] * i0, o[1] * i0);
    out(r0 + g + 8, h * ATT_HD + 2 * t, o[2] * i1, o[3] * i1);
}

// all query tiles of one (sequence, head) by one warp, L <= 64
template <typename Out>
__device__ __forceinline__ void att_head(const float* __restrict__ sq, int h, int L, int lane, Out&& out) {
    att_mtile<0>(sq, h, L, lane, out);
    if (L > 16) att_mtile<1>(sq, h, L, lane, out);
    if (L > 32) att_mtile<2>(sq, h, L, lane, out);
    if (L > 48) att_mtile<3>(sq, h, L, lane, out);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// General form for head_dim = 8 * KH (KH = 2: d_model 128 with 8 heads, the Memory-5 network): the q|k|v image has rows of
// 3 * D floats (+ pad), head h occupies columns [h * HD, (h + 1) * HD) of each third.  Q K^T runs KH k-steps of 8 per block and
// P V produces KH 8-column tiles, each with the same TF32 hi/lo split (3 MMAs per product) as above.
template <int MT, int KH, int D, int LD, typename Out>
__device__ __forceinline__ void att_mtile_g(const float* __restrict__ sq, int h, int L, int lane, Out&& out) {
    constexpr int NB = 2 * MT + 2, HD = 8 * KH;
    const int g = lane >> 2, t = lane & 3;
    const int nb_l = (L + 7) >> 3;
    const int r0 = 16 * MT;
    uint32_t qh[KH][4], ql[KH][4];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
        const float* q0 = sq + (r0 + g) * LD + h * HD + 8 * kh + t;
        att_split(q0[0], qh[kh][0], ql[kh][0]);
        att_split(q0[8 * LD], qh[kh][1], ql[kh][1]);
        att_split(q0[4], qh[kh][2], ql[kh][2]);
        att_split(q0[8 * LD + 4], qh[kh][3], ql[kh][3]);
    }
    float s[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        s[b][0] = s[b][1] = s[b][2] = s[b][3] = 0.f;
        if (b < nb_l) {
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
                const float* kp = sq + (8 * b + g) * LD + D + h * HD + 8 * kh + t;
                uint32_t kh0, kl0, kh1, kl1;
                att_split(kp[0], kh0, kl0);
                att_split(kp[4], kh1, kl1);
                att_mma(s[b], ql[kh], kh0, kh1);
                att_mma(s[b], qh[kh], kl0, kl1);
                att_mma(s[b], qh[kh], kh0, kh1);
            }
        }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < nb_l) {
            if (b >= NB - 2) {
                const int key = 8 * b + 2 * t, ra = r0 + g, rb = r0 + g + 8;
                if (key > ra) s[b][0] = -INFINITY;
                if (key + 1 > ra) s[b][1] = -INFINITY;
                if (key > rb) s[b][2] = -INFINITY;
                if (key + 1 > rb) s[b][3] = -INFINITY;
            }
            m0 = fmaxf(m0, fmaxf(s[b][0], s[b][1]));
            m1 = fmaxf(m1, fmaxf(s[b][2], s[b][3]));
        }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f, o[KH][4], o_lh[KH][4], o_hl[KH][4];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
        for (int e = 0; e < 4; ++e) { o[kh][e] = 0.f; o_lh[kh][e] = 0.f; o_hl[kh][e] = 0.f; }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b < nb_l) {
            const float p0 = att_ex2(s[b][0] - m0), p1 = att_ex2(s[b][1] - m0);
            const float p2 = att_ex2(s[b][2] - m1), p3 = att_ex2(s[b][3] - m1);
            l0 += p0 + p1; l1 += p2 + p3;
            uint32_t ph[4], pl[4];
            att_split(p0, ph[0], pl[0]); att_split(p2, ph[1], pl[1]);
            att_split(p1, ph[2], pl[2]); att_split(p3, ph[3], pl[3]);
#pragma unroll
            for (int kh = 0; kh < KH; ++kh) {
                const float* vp = sq + (8 * b + 2 * t) * LD + 2 * D + h * HD + 8 * kh + g;
                uint32_t vh0, vl0, vh1, vl1;
                att_split(vp[0], vh0, vl0);
                att_split(vp[LD], vh1, vl1);
                att_mma(o_lh[kh], pl, vh0, vh1);
                att_mma(o_hl[kh], ph, vl0, vl1);
                att_mma(o[kh], ph, vh0, vh1);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
        out(r0 + g, h * HD + 8 * kh + 2 * t, (o[kh][0] + o_lh[kh][0] + o_hl[kh][0]) * i0, (o[kh][1] + o_lh[kh][1] + o_hl[kh][1]) * i0);
        out(r0 + g + 8, h * HD + 8 * kh + 2 * t, (o[kh][2] + o_lh[kh][2] + o_hl[kh][2]) * i1, (o[kh][3] + o_lh[kh][3] + o_hl[kh][3]) * i1);
    }
}

template <int KH, int D, int LD, typename Out>
__device__ __forceinline__ void att_head_g(const float* __restrict__ sq, int h, int L, int lane, Out&& out) {
    att_mtile_g<0, KH, D, LD>(sq, h, L, lane, out);
    if (L > 16) att_mtile_g<1, KH, D, LD>(sq, h, L, lane, out);
    if (L > 32) att_mtile_g<2, KH, D, LD>(sq, h, L, lane, out);
    if (L > 48) att_mtile_g<3, KH, D, LD>(sq, h, L, lane, out);
}
