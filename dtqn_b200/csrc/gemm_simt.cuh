// fp32 CUDA-core GEMM tile used for the small / latency-bound contractions of the DTQN hot path (train batch of
// 32 x 50 tokens, K = 64..512) and for every backward GEMM.  64 x BN output tile per 256-thread CTA, BK = 32,
// register prefetch of the next k-slab, conflict-free float4 shared-memory reads.
//   NT:  C[M,N] = A[M,K] * W[N,K]^T   (forward Linear: x W^T)            B_NN = false
//   NN:  C[M,N] = A[M,K] * B[K,N]     (dgrad: dY W)                      B_NN = true
#pragma once
#include "common.cuh"

#define GEMM_BM 64
#define GEMM_BK 32
#define GEMM_THREADS 256

template <int BN>
struct GemmSmem {
    float As[GEMM_BK][GEMM_BM + 4];
    float Bs[GEMM_BK][BN + 4];
};

// acc[i][j] <-> row m0 + ty*4 + i, col n0 + gemm_col(tx, j), with ty = tid / 16, tx = tid % 16: each thread owns groups
// of 4 contiguous columns 64 apart, so the float4 reads of a B slab row are contiguous across the half-warp.
__device__ __forceinline__ int gemm_col(int tx, int j) { return (j >> 2) * 64 + tx * 4 + (j & 3); }

// BM = 64 / 32 / 16 rows per CTA (RM = BM / 16 rows per thread): the small-M contractions of the backward pass (1 600 tokens)
// trade register reuse for more CTAs; the k order of every output element is the same for all BM, so results are bitwise equal.
template <int BM, int BN, bool B_NN>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, int lda, int M,
                                          const float* __restrict__ B, int ldb, int K, int m0, int n0,
                                          float (&acc)[BM / 16][BN / 16], GemmSmem<BN>& sm) {
    static_assert(BM == 64 || BM == 32 || BM == 16, "BM");
    constexpr int RM = BM / 16;
    constexpr int TN = BN / 16;
    constexpr int AV = (BM * GEMM_BK / 4 + GEMM_THREADS - 1) / GEMM_THREADS;   // float4 loads of the A slab per thread (2, 1, 1)
    constexpr int BV = BN * GEMM_BK / 4 / GEMM_THREADS;        // float4 loads of the B slab per thread (2 or 4)
    constexpr int KQ = GEMM_BK / 4;                            // float4 per slab row along k
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 a_reg[AV], b_reg[BV];
    auto load_slab = [&](int k0) {
#pragma unroll
        for (int v = 0; v < AV; ++v) {   // A slab: BM rows x BK k, float4 along k
            const int idx = tid + v * GEMM_THREADS;
            const int row = idx / KQ, kq = idx % KQ;
            a_reg[v] = (row < BM && (m0 + row) < M) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * lda + k0 + kq * 4)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < BV; ++v) {
            const int idx = tid + v * GEMM_THREADS;
            if (B_NN) {   // B[k0 + kr, n0 + nq*4 ..]  (BK x BN slab, N contiguous)
                const int kr = idx / (BN / 4), nq = idx % (BN / 4);
                b_reg[v] = *reinterpret_cast<const float4*>(B + (size_t)(k0 + kr) * ldb + n0 + nq * 4);
            } else {      // W[n0 + nr, k0 + kq*4 ..]  (BN x BK slab, K contiguous)
                const int nr = idx / KQ, kq = idx % KQ;
                b_reg[v] = *reinterpret_cast<const float4*>(B + (size_t)(n0 + nr) * ldb + k0 + kq * 4);
            }
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int v = 0; v < AV; ++v) {
            const int idx = tid + v * GEMM_THREADS;
            const int row = idx / KQ, kq = idx % KQ;
            if (row < BM) {
                sm.As[kq * 4 + 0][row] = a_reg[v].x; sm.As[kq * 4 + 1][row] = a_reg[v].y;
                sm.As[kq * 4 + 2][row] = a_reg[v].z; sm.As[kq * 4 + 3][row] = a_reg[v].w;
            }
        }
#pragma unroll
        for (int v = 0; v < BV; ++v) {
            const int idx = tid + v * GEMM_THREADS;
            if (B_NN) {
                const int kr = idx / (BN / 4), nq = idx % (BN / 4);
                *reinterpret_cast<float4*>(&sm.Bs[kr][nq * 4]) = b_reg[v];
            } else {
                const int nr = idx / KQ, kq = idx % KQ;
                sm.Bs[kq * 4 + 0][nr] = b_reg[v].x; sm.Bs[kq * 4 + 1][nr] = b_reg[v].y;
                sm.Bs[kq * 4 + 2][nr] = b_reg[v].z; sm.Bs[kq * 4 + 3][nr] = b_reg[v].w;
            }
        }
    };
    load_slab(0);
    for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
        store_slab();
        __syncthreads();
        if (k0 + GEMM_BK < K) load_slab(k0 + GEMM_BK);
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            float a[RM];
            if constexpr (RM == 4) {
                const float4 av = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 4]);
                a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
            } else if constexpr (RM == 2) {
                const float2 av = *reinterpret_cast<const float2*>(&sm.As[kk][ty * 2]);
                a[0] = av.x; a[1] = av.y;
            } else {
                a[0] = sm.As[kk][ty];
            }
            float b[TN];
#pragma unroll
            for (int j4 = 0; j4 < TN / 4; ++j4) {
                const float4 bv = *reinterpret_cast<const float4*>(&sm.Bs[kk][j4 * 64 + tx * 4]);
                b[j4 * 4 + 0] = bv.x; b[j4 * 4 + 1] = bv.y; b[j4 * 4 + 2] = bv.z; b[j4 * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}

template <int BN, bool B_NN>
__device__ __forceinline__ void gemm_tile_64(const float* __restrict__ A, int lda, int M,
                                             const float* __restrict__ B, int ldb, int K, int m0, int n0,
                                             float (&acc)[4][BN / 16], GemmSmem<BN>& sm) {
    gemm_tile<64, BN, B_NN>(A, lda, M, B, ldb, K, m0, n0, acc, sm);
}
