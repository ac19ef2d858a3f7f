// tcgen05 (5th-gen tensor core) Linear for the large-M forward GEMMs of the DTQN hot path: the acting forward over
// n_envs x ctx tokens (QKV / out_proj / FFN projections, the only dense contractions -- transformer.py:64-77).
//
//   Y[128-token tile, N_TILE] = epi( X[128, K] W[N_TILE, K]^T + b )
//
// * fp32 parity on bf16 tensor cores: every operand is split x = hi + lo (two bf16), and three MMAs per k-step
//   accumulate hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator (error ~2^-16 per product; measured against the
//   1e-3 bar in tests).  Tensor-pipe utilisation is therefore capped at 1/3 on algorithmic FLOPs.
// * weights are pre-split and pre-tiled ("packed") by pack_weights_kernel into the exact shared-memory image of a
//   K-major SWIZZLE_NONE UMMA operand, so one cp.async.bulk (TMA bulk copy, mbarrier complete_tx) per k-chunk brings a
//   [N_TILE x 64] hi+lo block in; activations are loaded fp32 (coalesced), split in registers and written to the
//   canonical layout by all 128 threads.
// * one elected thread issues tcgen05.mma (cta_group::1, M=128); tcgen05.commit signals an mbarrier; the epilogue reads
//   the accumulator with tcgen05.ld 32x32b (thread = row), so bias / ReLU / residual / LayerNorm are thread-local.
// * single smem stage per CTA, 2-3 CTAs per SM (TMEM 64..256 of 512 columns each) overlap each other's load / MMA /
//   epilogue phases.
#include "net.cuh"
#include "prof.cuh"
#include "linear_tc.cuh"
#include "tc_common.cuh"

namespace {

// ---- weight packing ---------------------------------------------------------------------------------------------------------
// W[N, K] fp32 row-major -> blocks [n_tile][k_chunk] of { hi[8][N_TILE][8 bf16], lo[8][N_TILE][8 bf16] }.
__global__ void pack_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed, TcPackTable tab) {
    const TcPackEntry e = tab.e[blockIdx.y];
    if (e.n_tile == 0) {                                       // fp32 transpose: out[k][n] = W[n][k]
        float* out = reinterpret_cast<float*>(packed + e.pk_off);
        const long long total = (long long)e.N * e.K;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
            const int k = (int)(idx / e.N), n = (int)(idx % e.N);
            out[idx] = params[e.w_off + (long long)n * e.K + k];
        }
        return;
    }
    const int NT = e.n_tile;
    const long long chunks = (long long)e.N * (e.K / 8);          // 16-byte chunks per half
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx / (e.K / 8)), kc8 = (int)(idx % (e.K / 8));
        const float* src = params + e.w_off + (long long)n * e.K + kc8 * 8;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = src[i];
        uint4 hi, lo;
        split8(x, hi, lo);
        const int nt = n / NT, nl = n % NT, kc = (kc8 * 8) / TC_KC, c = kc8 % (TC_KC / 8);
        const long long blk = ((long long)nt * (e.K / TC_KC) + kc) * ((long long)NT * TC_KC * 4);
        uint8_t* dst = packed + e.pk_off + blk + ((long long)c * NT + nl) * 16;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + (long long)NT * TC_KC * 2) = lo;
    }
}

// ---- the GEMM -----------------------------------------------------------------------------------------------------------------
struct TcArgs {
    LinArgs a;
    const uint8_t* packed[DTQN_MAX_GROUPS];
    long long pk_off;
    TcEmbed emb;          // emb.mode: bit 0 = the A operand, bit 1 = the LayerNorm residual is the token embedding x0
};

// x0[row][c..c+3] = b_e + pos[j] + W_e obs (dtqn.py:181-199, continuous observations) recomputed on the fly, so the acting
// forward never materialises x0: token row -> (sequence i, position j) -> context-ring row (utils/context.py window).
__device__ __forceinline__ const float* emb_obs_row(const TcEmbed& e, int g, long long row, int& j) {
    const int i = (int)(row / e.L);
    j = (int)(row % e.L);
    const dtqn_obs_src& s = e.src[g];
    int rr = j;
    if (s.timestep) {
        const int ts = s.timestep[i];
        const int n = min(s.ring_len, ts + 1);
        rr = j < n ? (ts + 1 - n + j) % s.ring_len : -1;
    }
    return rr < 0 ? nullptr : s.obs + (long long)i * s.seq_stride + (long long)rr * e.O;
}
__device__ __forceinline__ float4 emb_x0_quad(const TcEmbed& e, const float* __restrict__ p, const float* __restrict__ sEW,
                                              const float* __restrict__ obs, int j, int c) {
    float ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ov[k] = (k < e.O) ? (obs ? __ldg(obs + k) : e.obs_mask) : 0.f;
    const float4 pv = __ldg(reinterpret_cast<const float4*>(p + e.pos_off + (long long)j * 64 + c));
    float o[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float acc = sEW[64 * 4 + c + q];                       // bias
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(ov[k], sEW[(c + q) * 4 + k], acc);
        o[q] += acc;
    }
    return make_float4(o[0], o[1], o[2], o[3]);
}

template <int N_TILE, int EPI>
__global__ void __launch_bounds__(128)
linear_tc_kernel(TcArgs t) {
    extern __shared__ uint8_t smem_raw[];
    const LinArgs& a = t.a;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.z, m0 = blockIdx.x * TC_M, nt = blockIdx.y;
    constexpr int TMEM_COLS = N_TILE <= 64 ? 64 : (N_TILE <= 128 ? 128 : 256);
    constexpr uint32_t B_HALF = N_TILE * TC_KC * 2;               // bytes of one of hi / lo of a weight block

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic only: keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t* sA = base;                                            // hi then lo, A_HALF_BYTES each
    uint8_t* sB = base + ((2 * A_HALF_BYTES + 1023) & ~1023);      // hi then lo, B_HALF each
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * B_HALF); // [0] weights landed, [1] MMAs retired
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
    __shared__ int s_fail;
    const uint32_t bar_b = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_mma, 1);
        s_fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    const size_t grow = (size_t)g * a.Tg;
    const float* X = a.X + grow * a.K;
    const int n_it = a.K / TC_KC;
    const uint8_t* wblk = t.packed[g] + t.pk_off + (size_t)nt * n_it * (2 * B_HALF);
    const uint32_t idesc = umma_idesc(TC_M, N_TILE);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    bool ok = true;

    for (int it = 0; it < n_it; ++it) {
        if (it > 0) ok = ok && mbar_wait(bar_mma, (uint32_t)((it - 1) & 1));   // smem stage free again
        if (tid == 0 && ok) {
            mbar_expect_tx(bar_b, 2 * B_HALF);
            bulk_g2s(sB_u, wblk + (size_t)it * (2 * B_HALF), 2 * B_HALF, bar_b);
        }
        // activations: 16 rows x 64 floats per pass; a warp reads 4 rows x 256 B (coalesced), splits, stores conflict-free
#pragma unroll
        for (int p = 0; p < TC_M / 16; ++p) {
            const int r = p * 16 + (tid >> 3), c = tid & 7;
            float x[8];
            if (m0 + r < a.Tg) {
                const float4* src = reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * a.K + it * TC_KC + c * 8);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
            }
            uint4 hi, lo;
            split8(x, hi, lo);
            *reinterpret_cast<uint4*>(sA + c * A_CHUNK_STRIDE + r * 16) = hi;
            *reinterpret_cast<uint4*>(sA + A_HALF_BYTES + c * A_CHUNK_STRIDE + r * 16) = lo;
        }
        fence_async_smem();                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0 && ok) {
            if (mbar_wait(bar_b, (uint32_t)(it & 1))) {
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
                    const uint64_t a_hi = umma_desc(sA_u + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                    const uint64_t a_lo = umma_desc(sA_u + A_HALF_BYTES + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                    const uint64_t b_hi = umma_desc(sB_u + k16 * 2 * (N_TILE * 16), N_TILE * 16, 128);
                    const uint64_t b_lo = umma_desc(sB_u + B_HALF + k16 * 2 * (N_TILE * 16), N_TILE * 16, 128);
                    umma_bf16(tmem, a_hi, b_hi, idesc, (it | k16) ? 1u : 0u);
                    umma_bf16(tmem, a_hi, b_lo, idesc, 1u);
                    umma_bf16(tmem, a_lo, b_hi, idesc, 1u);
                }
                umma_commit(bar_mma);         // arrives when every MMA issued so far has retired
            } else s_fail = 1;
        }
    }
    ok = ok && mbar_wait(bar_mma, (uint32_t)((n_it - 1) & 1));
    __syncthreads();
    ok = ok && !s_fail;
    tc_fence_after();

    if (ok) {
        const int r = m0 + warp * 32 + lane;                        // TMEM lane == tile row
        const bool row_ok = r < a.Tg;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const float* p = a.P.p[g];
        const int n0 = nt * N_TILE;
        if (EPI != EPI_RES_LN) {
#pragma unroll 1
            for (int c0 = 0; c0 < N_TILE; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                if (row_ok) {
                    float* y = a.Y + (grow + r) * (size_t)a.N + n0 + c0;
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        float o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[i] = v[q + i] + __ldg(p + a.b_off + n0 + c0 + q + i);
                            if (EPI == EPI_BIAS_RELU) o[i] = fmaxf(o[i], 0.f);
                        }
                        *reinterpret_cast<float4*>(y + q) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
        } else {
            // x_out = LayerNorm(x_res + relu(acc + b)); the whole row (N_TILE == d_model) lives in this thread
            float u[N_TILE];
            const size_t ro = (grow + (row_ok ? r : 0)) * (size_t)N_TILE;
            float s = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < N_TILE; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    const float4 xr = *reinterpret_cast<const float4*>(a.R + ro + c0 + q);
                    const float xv[4] = {xr.x, xr.y, xr.z, xr.w};
                    float rl[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        rl[i] = fmaxf(v[q + i] + __ldg(p + a.b_off + c0 + q + i), 0.f);
                        u[c0 + q + i] = xv[i] + rl[i];
                        s += u[c0 + q + i];
                    }
                    if (a.r_save && row_ok)
                        *reinterpret_cast<float4*>(a.r_save + ro + c0 + q) = make_float4(rl[0], rl[1], rl[2], rl[3]);
                }
            }
            const float mean = s * (1.f / N_TILE);
            float vs = 0.f;
#pragma unroll
            for (int j = 0; j < N_TILE; ++j) { const float dl = u[j] - mean; vs = fmaf(dl, dl, vs); }
            const float rstd = 1.0f / sqrtf(vs * (1.f / N_TILE) + 1e-5f);
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < N_TILE; q += 4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = (u[q + i] - mean) * rstd * __ldg(p + a.gamma_off + q + i) + __ldg(p + a.beta_off + q + i);
                    *reinterpret_cast<float4*>(a.Y + ro + q) = make_float4(o[0], o[1], o[2], o[3]);
                }
                if (a.st_save) { a.st_save[(grow + r) * 2] = mean; a.st_save[(grow + r) * 2 + 1] = rstd; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <int N_TILE, int EPI>
int launch_one(const TcArgs& t, int G, cudaStream_t st) {
    constexpr size_t smem = 1024 + ((2 * A_HALF_BYTES + 1023) & ~1023) + 2 * (size_t)N_TILE * TC_KC * 2 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel<N_TILE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid(dtqn_cdiv(t.a.Tg, TC_M), t.a.N / N_TILE, G);
    linear_tc_kernel<N_TILE, EPI><<<grid, 128, smem, st>>>(t);
    return 0;
}


// =================================================================================================================================
// Pipelined persistent variant (the one the acting forward runs): weights resident in shared memory for the CTA's whole
// life, warp-specialised roles connected by mbarriers, two shared-memory stages for the activation operand and two TMEM
// accumulators, so the global loads of tile i+1, the MMAs of tile i and the epilogue / stores of tile i-1 overlap:
//   warps 0-7  (256 thr)  producers: fp32 activations (coalesced) -> bf16 hi/lo -> canonical K-major smem stage
//   warp  8               one elected thread: TMA bulk load of the weight image (once), tcgen05.mma issue, tcgen05.commit
//   warps 9-16 (256 thr; 9-12 only for the LayerNorm epilogue)  epilogue: tcgen05.ld (thread = row) -> bias / ReLU /
//                          residual+LayerNorm -> padded smem staging ->
//                          full-line coalesced global stores (a thread-per-row store would touch 32 lines per instruction)
// =================================================================================================================================
constexpr int PIPE_PRODUCERS = 256;
constexpr int PIPE_THREADS_LN = PIPE_PRODUCERS + 32 + 128;     // 4 epilogue warps (thread = full row, LayerNorm in registers)
constexpr int PIPE_THREADS = PIPE_PRODUCERS + 32 + 256;        // 8 epilogue warps: two per TMEM lane quarter, half the columns each
constexpr int STG_COLS = 64;                                   // columns staged per epilogue round
constexpr int STG_ROW_BYTES = STG_COLS * 4 + 16;               // padded row -> conflict-free 16-byte stores
constexpr int STG_BYTES = TC_M * STG_ROW_BYTES;

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }


// Each epilogue warp owns 32 staged rows (padded to STG_ROW_BYTES).  Write them out row-major with full-line stores:
// lanes 0-15 cover the 256 B of one row, lanes 16-31 the next row.
__device__ __forceinline__ void warp_store_rows(const uint8_t* stg_warp, float* dst_row0, size_t row_stride_floats,
                                                int rows_valid, int lane) {
    __syncwarp();
    const int half = lane >> 4, c16 = lane & 15;
#pragma unroll 8
    for (int rr = 0; rr < 32; rr += 2) {
        const int r = rr + half;
        const float4 v = *reinterpret_cast<const float4*>(stg_warp + r * STG_ROW_BYTES + c16 * 16);
        if (r < rows_valid) *reinterpret_cast<float4*>(dst_row0 + (size_t)r * row_stride_floats + c16 * 4) = v;
    }
    __syncwarp();
}

template <int N_TILE, int K_CHUNKS, int EPI>
__global__ void __launch_bounds__(EPI == EPI_RES_LN ? PIPE_THREADS_LN : PIPE_THREADS, 1)
linear_tc_pipe_kernel(TcArgs t) {
    extern __shared__ uint8_t smem_raw[];
    const LinArgs& a = t.a;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NT_PER_G_DUMMY = 0; (void)NT_PER_G_DUMMY;
    const int n_ntiles = a.N / N_TILE;
    const int g = blockIdx.y / n_ntiles, nt = blockIdx.y % n_ntiles;
    constexpr int TMEM_COLS = (2 * N_TILE) <= 128 ? 128 : ((2 * N_TILE) <= 256 ? 256 : 512);
    constexpr uint32_t B_HALF = N_TILE * TC_KC * 2;            // one of hi / lo of one k-chunk block
    constexpr uint32_t B_BYTES = 2 * B_HALF * K_CHUNKS;

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic only: keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t* sB = base;                                        // resident weight image: K_CHUNKS x {hi, lo}
    uint8_t* sA = sB + ((B_BYTES + 1023) & ~1023u);            // 2 stages
    uint8_t* sStg = sA + 2 * A_STAGE_BYTES;                    // 2 staging buffers
    float* sBias = reinterpret_cast<float*>(sStg + 2 * STG_BYTES);       // bias | gamma | beta, N_TILE each
    float* sEW = sBias + 3 * N_TILE;                                      // obs-embedding W [64][4] (zero padded) | b [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEW + 64 * 5);           // a_full[2] a_empty[2] acc_full[2] acc_empty[2] b_full
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 9);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    enum { A_FULL = 0, A_EMPTY = 2, ACC_FULL = 4, ACC_EMPTY = 6, B_FULL = 8 };

    const float* p = a.P.p[g];
    if (t.emb.mode && tid < 64) {
        for (int k = 0; k < 4; ++k) sEW[tid * 4 + k] = k < t.emb.O ? __ldg(p + t.emb.w_off + tid * t.emb.O + k) : 0.f;
        sEW[64 * 4 + tid] = __ldg(p + t.emb.b_off + tid);
    }
    if (tid < N_TILE) {
        sBias[tid] = __ldg(p + a.b_off + nt * N_TILE + tid);
        if (EPI == EPI_RES_LN) {
            sBias[N_TILE + tid] = __ldg(p + a.gamma_off + tid);
            sBias[2 * N_TILE + tid] = __ldg(p + a.beta_off + tid);
        }
    }
    if (tid == 0) {
        mbar_init(BAR(A_FULL), PIPE_PRODUCERS); mbar_init(BAR(A_FULL + 1), PIPE_PRODUCERS);
        mbar_init(BAR(A_EMPTY), 1); mbar_init(BAR(A_EMPTY + 1), 1);
        mbar_init(BAR(ACC_FULL), 1); mbar_init(BAR(ACC_FULL + 1), 1);
        mbar_init(BAR(ACC_EMPTY), EPI == EPI_RES_LN ? 128 : 256); mbar_init(BAR(ACC_EMPTY + 1), EPI == EPI_RES_LN ? 128 : 256);
        mbar_init(BAR(B_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    const size_t grow = (size_t)g * a.Tg;
    const int m_tiles = (a.Tg + TC_M - 1) / TC_M;
    const int my_tiles = (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x, +gridDim.x, ...

    if (warp < 8) {
        // ------------------------------------------------ producers ------------------------------------------------
        // Two register sets: the global loads of item i+1 are issued before item i is converted, so ~2 tiles (64 KB) of
        // reads are in flight per SM while the MMA / epilogue of earlier tiles run.
        const float* X = a.X + grow * a.K;
        const int items = my_tiles * K_CHUNKS;
        auto issue = [&](int it, float4 (&v)[8]) {
            const int i = it / K_CHUNKS, kc = it % K_CHUNKS;
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TC_M;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                      // 32 rows x 64 floats per pass; a warp reads 4 rows x 256 B
                const int r = q * 32 + (tid >> 3), c = tid & 7;
                if (m0 + r < a.Tg) {
                    if (t.emb.mode & 1) {                      // A operand = token embedding, computed here (K == 64)
                        int j;
                        const float* ob = emb_obs_row(t.emb, g, m0 + r, j);
                        v[2 * q] = emb_x0_quad(t.emb, p, sEW, ob, j, c * 8);
                        v[2 * q + 1] = emb_x0_quad(t.emb, p, sEW, ob, j, c * 8 + 4);
                    } else {
                        const float4* src = reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * a.K + kc * TC_KC + c * 8);
                        v[2 * q] = __ldg(src); v[2 * q + 1] = __ldg(src + 1);
                    }
                } else { v[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * q + 1] = v[2 * q]; }
            }
        };
        auto finish = [&](int it, float4 (&v)[8]) -> bool {
            const int st = it & 1;
            if (it >= 2 && !mbar_wait(BAR(A_EMPTY + st), (uint32_t)(((it >> 1) - 1) & 1))) return false;
            uint8_t* dst = sA + st * A_STAGE_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + (tid >> 3), c = tid & 7;
                const float x[8] = {v[2 * q].x, v[2 * q].y, v[2 * q].z, v[2 * q].w, v[2 * q + 1].x, v[2 * q + 1].y, v[2 * q + 1].z, v[2 * q + 1].w};
                uint4 hi, lo;
                split8(x, hi, lo);
                *reinterpret_cast<uint4*>(dst + c * A_CHUNK_STRIDE + r * 16) = hi;
                *reinterpret_cast<uint4*>(dst + A_HALF_BYTES + c * A_CHUNK_STRIDE + r * 16) = lo;
            }
            fence_async_smem();
            mbar_arrive(BAR(A_FULL + st));
            return true;
        };
        float4 va[8], vb[8];
        if (items > 0) issue(0, va);
        bool okp = true;
        for (int it = 0; it < items && okp; it += 2) {
            if (it + 1 < items) issue(it + 1, vb);
            okp = finish(it, va);
            if (it + 2 < items) issue(it + 2, va);
            if (okp && it + 1 < items) okp = finish(it + 1, vb);
        }
    } else if (warp == 8) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        if (lane == 0) {
            const uint8_t* wimg = t.packed[g] + t.pk_off + (size_t)nt * B_BYTES;
            mbar_expect_tx(BAR(B_FULL), B_BYTES);
#pragma unroll 1
            for (uint32_t off = 0; off < B_BYTES; off += 32768u)
                bulk_g2s(smem_u32(sB) + off, wimg + off, (B_BYTES - off) < 32768u ? (B_BYTES - off) : 32768u, BAR(B_FULL));
            bool ok = mbar_wait(BAR(B_FULL), 0);
            const uint32_t idesc = umma_idesc(TC_M, N_TILE);
            const uint32_t sB_u = smem_u32(sB);
            int it = 0;
            for (int i = 0; i < my_tiles && ok; ++i) {
                const int as = i & 1;
                if (i >= 2) ok = mbar_wait(BAR(ACC_EMPTY + as), (uint32_t)(((i >> 1) - 1) & 1));
                const uint32_t d_tmem = tmem + (uint32_t)(as * N_TILE);
                for (int kc = 0; kc < K_CHUNKS && ok; ++kc, ++it) {
                    const int st = it & 1;
                    ok = mbar_wait(BAR(A_FULL + st), (uint32_t)((it >> 1) & 1));
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t sA_u = smem_u32(sA + st * A_STAGE_BYTES);
                    const uint32_t sBk = sB_u + (uint32_t)kc * 2u * B_HALF;
#pragma unroll
                    for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
                        const uint64_t a_hi = umma_desc(sA_u + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                        const uint64_t a_lo = umma_desc(sA_u + A_HALF_BYTES + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                        const uint64_t b_hi = umma_desc(sBk + k16 * 2 * (N_TILE * 16), N_TILE * 16, 128);
                        const uint64_t b_lo = umma_desc(sBk + B_HALF + k16 * 2 * (N_TILE * 16), N_TILE * 16, 128);
                        umma_bf16(d_tmem, a_hi, b_hi, idesc, (kc | k16) ? 1u : 0u);
                        umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
                        umma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
                    }
                    umma_commit(BAR(A_EMPTY + st));            // smem stage reusable once these MMAs retire
                }
                if (ok) umma_commit(BAR(ACC_FULL + as));       // accumulator complete
            }
        }
    } else if (EPI != EPI_RES_LN || warp < 13) {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int q4 = warp & 3;                                // TMEM lane quarter this warp may access
        const int row_in_tile = q4 * 32 + lane;
        const int n0 = nt * N_TILE;
        int round = 0; (void)round;
        for (int i = 0; i < my_tiles; ++i) {
            const int as = i & 1;
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TC_M;
            const int rows_valid = min(32, a.Tg - (m0 + q4 * 32));       // rows of this warp's quarter that exist
            if (EPI == EPI_RES_LN && N_TILE <= STG_COLS) {
                // pull this warp's residual rows into its staging rows with full-line loads while the MMAs still run
                uint8_t* stg_w = sStg + (q4 * 32) * STG_ROW_BYTES;
                const int half = lane >> 4, c16 = lane & 15;
                const float* src0 = a.R + (grow + m0 + q4 * 32) * (size_t)N_TILE;
                float4 rv[16];                                 // all 16 loads in flight before the first smem store
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) {
                    const int rw = 2 * rr + half;
                    rv[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rw < rows_valid) {
                        if (t.emb.mode & 2) {                  // residual = token embedding x0, recomputed (never stored)
                            int j;
                            const float* ob = emb_obs_row(t.emb, g, (long long)m0 + q4 * 32 + rw, j);
                            rv[rr] = emb_x0_quad(t.emb, p, sEW, ob, j, c16 * 4);
                        } else rv[rr] = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)rw * N_TILE + c16 * 4));
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 16; ++rr)
                    *reinterpret_cast<float4*>(stg_w + (2 * rr + half) * STG_ROW_BYTES + c16 * 16) = rv[rr];
                __syncwarp();
            }
            if (!mbar_wait(BAR(ACC_FULL + as), (uint32_t)((i >> 1) & 1))) break;
            tc_fence_after();
            const int r = m0 + row_in_tile;
            const bool row_ok = r < a.Tg;
            const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * N_TILE);
            if (EPI != EPI_RES_LN) {
                // warps 9-12 take the lower half of the columns, warps 13-16 the upper half, in rounds of 32 columns
                constexpr int RW = 32, RW_ROW = RW * 4 + 16;
                const int chalf = (warp - 9) >> 2;
                uint8_t* stg_w = sStg + (warp - 9) * (32 * RW_ROW);
                uint8_t* stg = stg_w + lane * RW_ROW;
#pragma unroll 1
                for (int c0 = chalf * (N_TILE / 2); c0 < (chalf + 1) * (N_TILE / 2); c0 += RW) {
                    uint32_t ra_[16], rb_[16];
                    tmem_ld16_issue(trow + c0, ra_);
                    tmem_ld16_issue(trow + c0 + 16, rb_);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0 + q);
                        const float4 b1 = *reinterpret_cast<const float4*>(sBias + c0 + 16 + q);
                        float o0[4] = {__uint_as_float(ra_[q]) + b0.x, __uint_as_float(ra_[q + 1]) + b0.y, __uint_as_float(ra_[q + 2]) + b0.z, __uint_as_float(ra_[q + 3]) + b0.w};
                        float o1[4] = {__uint_as_float(rb_[q]) + b1.x, __uint_as_float(rb_[q + 1]) + b1.y, __uint_as_float(rb_[q + 2]) + b1.z, __uint_as_float(rb_[q + 3]) + b1.w};
                        if (EPI == EPI_BIAS_RELU) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { o0[e] = fmaxf(o0[e], 0.f); o1[e] = fmaxf(o1[e], 0.f); }
                        }
                        *reinterpret_cast<float4*>(stg + q * 4) = make_float4(o0[0], o0[1], o0[2], o0[3]);
                        *reinterpret_cast<float4*>(stg + (16 + q) * 4) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                    }
                    __syncwarp();
                    if (rows_valid > 0) {
                        // 8 lanes cover the 128 B of one staged row, 4 rows per instruction: full-line stores
                        float* dst0 = a.Y + (grow + m0 + q4 * 32) * (size_t)a.N + n0 + c0;
                        const int rsub = lane >> 3, c8 = lane & 7;
#pragma unroll
                        for (int rr = 0; rr < 32; rr += 4) {
                            const int rw = rr + rsub;
                            const float4 v4 = *reinterpret_cast<const float4*>(stg_w + rw * RW_ROW + c8 * 16);
                            if (rw < rows_valid) *reinterpret_cast<float4*>(dst0 + (size_t)rw * a.N + c8 * 4) = v4;
                        }
                    }
                    __syncwarp();
                }
            } else {
                // x_out = LayerNorm(x_res + relu(acc + b)); N_TILE == d_model, the row lives in this thread.
                // The residual rows are first pulled into the warp's staging rows with full-line loads.
                float u[N_TILE];
                const size_t ro = (grow + (row_ok ? r : 0)) * (size_t)N_TILE;
                float s = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += STG_COLS) {
                    uint8_t* stg_w = sStg + (q4 * 32) * STG_ROW_BYTES;
                    if (N_TILE > STG_COLS) {
                        __syncwarp();
                        const int half = lane >> 4, c16 = lane & 15;
                        const float* src0 = a.R + (grow + m0 + q4 * 32) * (size_t)N_TILE + c0;
                        float4 rv[16];
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int rw = 2 * rr + half;
                            rv[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rw < rows_valid) rv[rr] = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)rw * N_TILE + c16 * 4));
                        }
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr)
                            *reinterpret_cast<float4*>(stg_w + (2 * rr + half) * STG_ROW_BYTES + c16 * 16) = rv[rr];
                        __syncwarp();
                    }
                    const uint8_t* myrow = stg_w + lane * STG_ROW_BYTES;
                    uint32_t tv[STG_COLS / 16][16];
#pragma unroll
                    for (int cc = 0; cc < STG_COLS; cc += 16) tmem_ld16_issue(trow + c0 + cc, tv[cc / 16]);
                    tmem_ld_wait();
#pragma unroll
                    for (int cc = 0; cc < STG_COLS; cc += 16) {
                        float v[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(tv[cc / 16][e]);
#pragma unroll
                        for (int q = 0; q < 16; q += 4) {
                            const float4 xr = *reinterpret_cast<const float4*>(myrow + (cc + q) * 4);
                            const float xv[4] = {xr.x, xr.y, xr.z, xr.w};
                            float rl[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                rl[e] = fmaxf(v[q + e] + sBias[c0 + cc + q + e], 0.f);
                                u[c0 + cc + q + e] = xv[e] + rl[e];
                                s += u[c0 + cc + q + e];
                            }
                            if (a.r_save && row_ok)
                                *reinterpret_cast<float4*>(a.r_save + ro + c0 + cc + q) = make_float4(rl[0], rl[1], rl[2], rl[3]);
                        }
                    }
                }
                const float mean = s * (1.f / N_TILE);
                float vs = 0.f;
#pragma unroll
                for (int j = 0; j < N_TILE; ++j) { const float dl = u[j] - mean; vs = fmaf(dl, dl, vs); }
                const float rstd = 1.0f / sqrtf(vs * (1.f / N_TILE) + 1e-5f);
                if (a.st_save && row_ok) { a.st_save[(grow + r) * 2] = mean; a.st_save[(grow + r) * 2 + 1] = rstd; }
#pragma unroll
                for (int c0 = 0; c0 < N_TILE; c0 += STG_COLS) {
                    uint8_t* stg_w = sStg + STG_BYTES + (q4 * 32) * STG_ROW_BYTES;
                    uint8_t* stg = stg_w + lane * STG_ROW_BYTES;
#pragma unroll
                    for (int q = 0; q < STG_COLS; q += 4) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            o[e] = (u[c0 + q + e] - mean) * rstd * sBias[N_TILE + c0 + q + e] + sBias[2 * N_TILE + c0 + q + e];
                        *reinterpret_cast<float4*>(stg + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    if (rows_valid > 0)
                        warp_store_rows(stg_w, a.Y + (grow + m0 + q4 * 32) * (size_t)N_TILE + c0, (size_t)N_TILE, rows_valid, lane);
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(ACC_EMPTY + as));                   // accumulator stage drained
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}


// =================================================================================================================================
// Fused position-wise FFN on tcgen05 (d_model = 64, hidden 256):  x2 = LayerNorm(x1 + relu(W2 relu(W1 x1 + b1) + b2))
// (transformer.py:37-42,74-77).  The [128 x 256] hidden tile never leaves the SM: MMA1 writes it to TMEM, the "hidden"
// epilogue warps read it back 64 columns at a time, apply bias + ReLU, split to bf16 hi/lo and store it as the K-major A
// operand of MMA2, whose [128 x 64] accumulator is drained by the LayerNorm epilogue warps.  Per token this moves
// 64 floats in + 64 residual + 64 out instead of 64 + 256 + 256 + 64 + 64 (the unfused ffn.0 / ffn.2 pair): 4x less HBM traffic.
//   warps 0-7   producers: x1 tile fp32 -> bf16 hi/lo -> A1 (one stage; two register sets keep the next tile's loads in flight)
//   warp  8     elected lane: TMA W1 image once (64 KB), W2 k-chunks through a 2-stage ring (16 KB each), all MMAs + commits
//   warps 9-12  hidden epilogue:  acc1[c] -> +b1 -> ReLU -> hi/lo -> A2 (K-major) ; signals A2_FULL per 64-column chunk
//   warps 13-16 LayerNorm epilogue: acc2 -> +b2 -> ReLU -> +x1 -> LN -> staged, full-line stores
// TMEM: acc1 = 256 columns (4 chunks), acc2 = 2 x 64 columns (double buffered across tiles).
// =================================================================================================================================
constexpr int FFN_THREADS = PIPE_PRODUCERS + 32 + 256 + 128;   // 8 producer + 1 MMA + 8 hidden-epilogue + 4 LayerNorm-epilogue warps

struct FfnArgs {
    const float* X;            // x1 [G*Tg, 64]  (input and LayerNorm residual)
    float* Y;                  // x2 [G*Tg, 64]
    GroupPtrs P;
    const uint8_t* packed[DTQN_MAX_GROUPS];
    long long pk_w1, pk_w2;    // byte offsets of the two weight images in the packed buffer
    long long b1_off, b2_off, gamma_off, beta_off;
    int Tg;
};

__global__ void __launch_bounds__(FFN_THREADS, 1)
ffn_tc_kernel(FfnArgs t) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int D = 64, HID = 256, NCH = HID / 64;
    constexpr uint32_t W1_BYTES = HID * TC_KC * 4;             // hi + lo of the whole [256 x 64] image
    constexpr uint32_t W1_HALF = HID * TC_KC * 2;
    constexpr uint32_t W2C_BYTES = D * TC_KC * 4;              // one k-chunk of W2 [64 x 64] hi + lo
    constexpr uint32_t W2C_HALF = D * TC_KC * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic only: keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t* sW1 = base;
    uint8_t* sW2 = sW1 + W1_BYTES;                             // 2 stages
    uint8_t* sA1 = sW2 + 2 * W2C_BYTES;
    uint8_t* sA2 = sA1 + A_STAGE_BYTES;                        // 2 stages: hidden chunk n -> stage n & 1
    float* sB1 = reinterpret_cast<float*>(sA2 + 2 * A_STAGE_BYTES);   // b1[256] | b2[64] | gamma[64] | beta[64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + HID + 3 * D);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 28);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    enum { W1_FULL = 0, A1_FULL = 1, A1_EMPTY = 2, A2_FULL = 3 /*+s*/, A2_EMPTY = 5 /*+s*/, W2_FULL = 7 /*+s*/, W2_EMPTY = 9 /*+s*/,
           ACC2_FULL = 11 /*+as*/, ACC2_EMPTY = 13 /*+as*/, ACC1_FULL = 15 /*+c*/, ACC1_EMPTY = 19 /*+c*/ };

    const float* p = t.P.p[g];
    for (int e = tid; e < HID; e += FFN_THREADS) sB1[e] = __ldg(p + t.b1_off + e);
    if (tid < D) {
        sB1[HID + tid] = __ldg(p + t.b2_off + tid);
        sB1[HID + D + tid] = __ldg(p + t.gamma_off + tid);
        sB1[HID + 2 * D + tid] = __ldg(p + t.beta_off + tid);
    }
    if (tid == 0) {
        mbar_init(BAR(W1_FULL), 1);
        mbar_init(BAR(A1_FULL), PIPE_PRODUCERS); mbar_init(BAR(A1_EMPTY), 1);
        for (int k = 0; k < 4; ++k) { mbar_init(BAR(ACC1_FULL + k), 1); mbar_init(BAR(ACC1_EMPTY + k), 256); }
        for (int k = 0; k < 2; ++k) {
            mbar_init(BAR(A2_FULL + k), 256); mbar_init(BAR(A2_EMPTY + k), 1);
            mbar_init(BAR(W2_FULL + k), 1); mbar_init(BAR(W2_EMPTY + k), 1);
            mbar_init(BAR(ACC2_FULL + k), 1); mbar_init(BAR(ACC2_EMPTY + k), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tm_acc1 = tmem, tm_acc2 = tmem + 256;       // acc1: columns [0,256); acc2[as]: [256 + 64 as, +64)

    const size_t grow = (size_t)g * t.Tg;
    const int m_tiles = (t.Tg + TC_M - 1) / TC_M;
    const int my_tiles = (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < 8) {
        // ------------------------------------------------ producers ------------------------------------------------
        const float* X = t.X + grow * D;
        auto issue = [&](int i, float4 (&v)[8]) {
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TC_M;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + (tid >> 3), c = tid & 7;
                if (m0 + r < t.Tg) {
                    const float4* src = reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * D + c * 8);
                    v[2 * q] = __ldg(src); v[2 * q + 1] = __ldg(src + 1);
                } else { v[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f); v[2 * q + 1] = v[2 * q]; }
            }
        };
        auto finish = [&](int i, float4 (&v)[8]) -> bool {
            if (i >= 1 && !mbar_wait(BAR(A1_EMPTY), (uint32_t)((i - 1) & 1))) return false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + (tid >> 3), c = tid & 7;
                const float x[8] = {v[2 * q].x, v[2 * q].y, v[2 * q].z, v[2 * q].w, v[2 * q + 1].x, v[2 * q + 1].y, v[2 * q + 1].z, v[2 * q + 1].w};
                uint4 hi, lo;
                split8(x, hi, lo);
                *reinterpret_cast<uint4*>(sA1 + c * A_CHUNK_STRIDE + r * 16) = hi;
                *reinterpret_cast<uint4*>(sA1 + A_HALF_BYTES + c * A_CHUNK_STRIDE + r * 16) = lo;
            }
            fence_async_smem();
            mbar_arrive(BAR(A1_FULL));
            return true;
        };
        float4 va[8], vb[8];
        if (my_tiles > 0) issue(0, va);
        bool okp = true;
        for (int i = 0; i < my_tiles && okp; i += 2) {
            if (i + 1 < my_tiles) issue(i + 1, vb);
            okp = finish(i, va);
            if (i + 2 < my_tiles) issue(i + 2, va);
            if (okp && i + 1 < my_tiles) okp = finish(i + 1, vb);
        }
    } else if (warp == 8) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        if (lane == 0) {
            const uint8_t* w1img = t.packed[g] + t.pk_w1;
            const uint8_t* w2img = t.packed[g] + t.pk_w2;
            mbar_expect_tx(BAR(W1_FULL), W1_BYTES);
            bulk_g2s(smem_u32(sW1), w1img, 32768u, BAR(W1_FULL));
            bulk_g2s(smem_u32(sW1) + 32768u, w1img + 32768, 32768u, BAR(W1_FULL));
            const int n_total = my_tiles * NCH;
            auto load_w2 = [&](int n) {                        // chunk n = 4 i + c -> ring stage n & 1
                const int s_ = n & 1, c = n % NCH;
                mbar_expect_tx(BAR(W2_FULL + s_), W2C_BYTES);
                bulk_g2s(smem_u32(sW2) + (uint32_t)s_ * W2C_BYTES, w2img + (size_t)c * W2C_BYTES, W2C_BYTES, BAR(W2_FULL + s_));
            };
            if (n_total > 0) load_w2(0);
            if (n_total > 1) load_w2(1);
            bool ok = mbar_wait(BAR(W1_FULL), 0);
            const uint32_t idesc = umma_idesc(TC_M, 64);
            const uint32_t sA1_u = smem_u32(sA1), sA2_u = smem_u32(sA2), sW1_u = smem_u32(sW1), sW2_u = smem_u32(sW2);
            // MMA1 of one 64-column hidden chunk: acc1[c] of tile i = x1 W1[c]^T.  acc1[c] is free as soon as the hidden epilogue
            // has copied chunk c of tile i-1 to registers, so MMA1 runs a whole tile ahead of the conversion work.
            auto mma1_chunk = [&](int i, int c) -> bool {
                bool o = true;
                if (c == 0) o = mbar_wait(BAR(A1_FULL), (uint32_t)(i & 1));
                if (o && i >= 1) o = mbar_wait(BAR(ACC1_EMPTY + c), (uint32_t)((i - 1) & 1));
                if (!o) return false;
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
                    const uint64_t a_hi = umma_desc(sA1_u + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                    const uint64_t a_lo = umma_desc(sA1_u + A_HALF_BYTES + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                    const uint32_t boff = (uint32_t)(k16 * 2 * (HID * 16) + c * 64 * 16);
                    const uint64_t b_hi = umma_desc(sW1_u + boff, HID * 16, 128);
                    const uint64_t b_lo = umma_desc(sW1_u + W1_HALF + boff, HID * 16, 128);
                    umma_bf16(tm_acc1 + (uint32_t)(c * 64), a_hi, b_hi, idesc, k16 ? 1u : 0u);
                    umma_bf16(tm_acc1 + (uint32_t)(c * 64), a_hi, b_lo, idesc, 1u);
                    umma_bf16(tm_acc1 + (uint32_t)(c * 64), a_lo, b_hi, idesc, 1u);
                }
                umma_commit(BAR(ACC1_FULL + c));
                if (c == NCH - 1) umma_commit(BAR(A1_EMPTY));   // A1 stage free once all four chunks have read it
                return true;
            };
            for (int c = 0; c < NCH && ok && my_tiles > 0; ++c) ok = mma1_chunk(0, c);
            for (int i = 0; i < my_tiles && ok; ++i) {
                const int as = i & 1;
                for (int c = 0; c < NCH && ok; ++c) {
                    const int n = i * NCH + c, s_ = n & 1, m = n >> 1;
                    if (i + 1 < my_tiles) ok = mma1_chunk(i + 1, c);               // next tile's hidden chunk c
                    if (!ok) break;
                    // MMA2: out[128 x 64] += relu(hidden chunk c)[128 x 64] W2[:, chunk c]^T
                    ok = mbar_wait(BAR(W2_FULL + s_), (uint32_t)(m & 1));
                    if (ok) ok = mbar_wait(BAR(A2_FULL + s_), (uint32_t)(m & 1));
                    if (ok && c == 0 && i >= 2) ok = mbar_wait(BAR(ACC2_EMPTY + as), (uint32_t)(((i >> 1) - 1) & 1));
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t wb = sW2_u + (uint32_t)s_ * W2C_BYTES;
                    const uint32_t ab = sA2_u + (uint32_t)s_ * A_STAGE_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
                        const uint64_t a_hi = umma_desc(ab + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                        const uint64_t a_lo = umma_desc(ab + A_HALF_BYTES + k16 * 2 * A_CHUNK_STRIDE, A_CHUNK_STRIDE, 128);
                        const uint64_t b_hi = umma_desc(wb + k16 * 2 * (D * 16), D * 16, 128);
                        const uint64_t b_lo = umma_desc(wb + W2C_HALF + k16 * 2 * (D * 16), D * 16, 128);
                        umma_bf16(tm_acc2 + (uint32_t)(as * 64), a_hi, b_hi, idesc, (c | k16) ? 1u : 0u);
                        umma_bf16(tm_acc2 + (uint32_t)(as * 64), a_hi, b_lo, idesc, 1u);
                        umma_bf16(tm_acc2 + (uint32_t)(as * 64), a_lo, b_hi, idesc, 1u);
                    }
                    umma_commit(BAR(A2_EMPTY + s_));
                    umma_commit(BAR(W2_EMPTY + s_));
                    if (n + 2 < n_total) {                     // refill this W2 ring stage once its MMAs have retired
                        ok = mbar_wait(BAR(W2_EMPTY + s_), (uint32_t)(m & 1));
                        if (ok) load_w2(n + 2);
                    }
                }
                if (ok) umma_commit(BAR(ACC2_FULL + as));
            }
        }
    } else if (warp < 17) {
        // -------------------------------- hidden epilogue: 8 warps, two per TMEM lane quarter, 32 columns each --------------------------------
        const int q4 = warp & 3, row = q4 * 32 + lane, chalf = (warp - 9) >> 2;
        for (int i = 0; i < my_tiles; ++i) {
            bool ok = true;
            for (int c = 0; c < NCH && ok; ++c) {
                const int n = i * NCH + c, s_ = n & 1, m = n >> 1;
                ok = mbar_wait(BAR(ACC1_FULL + c), (uint32_t)(i & 1));
                if (!ok) break;
                tc_fence_after();
                uint32_t tv[2][16];
                const uint32_t trow = tm_acc1 + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c * 64 + chalf * 32);
                tmem_ld16_issue(trow, tv[0]);
                tmem_ld16_issue(trow + 16, tv[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(BAR(ACC1_EMPTY + c));              // chunk copied to registers: MMA1 of the next tile may overwrite it
                if (m >= 1) ok = mbar_wait(BAR(A2_EMPTY + s_), (uint32_t)((m - 1) & 1));   // MMA2 of chunk n-2 done with this stage
                if (!ok) break;
                uint8_t* dst = sA2 + s_ * A_STAGE_BYTES;
#pragma unroll
                for (int sb = 0; sb < 4; ++sb) {
                    const int sub = chalf * 4 + sb;             // 8-column group within the 64-column chunk
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        x[e] = fmaxf(__uint_as_float(tv[sb >> 1][(sb & 1) * 8 + e]) + sB1[c * 64 + sub * 8 + e], 0.f);
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    *reinterpret_cast<uint4*>(dst + sub * A_CHUNK_STRIDE + row * 16) = hi;
                    *reinterpret_cast<uint4*>(dst + A_HALF_BYTES + sub * A_CHUNK_STRIDE + row * 16) = lo;
                }
                fence_async_smem();
                mbar_arrive(BAR(A2_FULL + s_));
            }
            if (!ok) break;
        }
    } else {
        // -------------------------------- LayerNorm epilogue: 4 warps, thread = row --------------------------------
        const int q4 = warp & 3, row_in_tile = q4 * 32 + lane;
        const float* sB2 = sB1 + HID; const float* sG = sB2 + D; const float* sBe = sG + D;
        for (int i = 0; i < my_tiles; ++i) {
            const int as = i & 1;
            const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * TC_M;
            const int r = m0 + row_in_tile;
            const bool row_ok = r < t.Tg;
            const size_t ro = (grow + (row_ok ? r : 0)) * (size_t)D;
            float u[D];
#pragma unroll
            for (int q = 0; q < D; q += 4) {                   // residual row (in flight while the MMAs run)
                const float4 xr = __ldg(reinterpret_cast<const float4*>(t.X + ro + q));
                u[q] = xr.x; u[q + 1] = xr.y; u[q + 2] = xr.z; u[q + 3] = xr.w;
            }
            if (!mbar_wait(BAR(ACC2_FULL + as), (uint32_t)((i >> 1) & 1))) break;
            tc_fence_after();
            const uint32_t trow = tm_acc2 + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * 64);
            float s = 0.f;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t tv[16];
                tmem_ld16_issue(trow + cc * 16, tv);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    u[cc * 16 + e] += fmaxf(__uint_as_float(tv[e]) + sB2[cc * 16 + e], 0.f);
                    s += u[cc * 16 + e];
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(ACC2_EMPTY + as));                 // accumulator consumed: MMA2 of tile i+2 may start
            const float mean = s * (1.f / D);
            float vs = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) { const float dl = u[j] - mean; vs = fmaf(dl, dl, vs); }
            const float rstd = 1.0f / sqrtf(vs * (1.f / D) + 1e-5f);
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < D; q += 4) {
                    float o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = (u[q + e] - mean) * rstd * sG[q + e] + sBe[q + e];
                    *reinterpret_cast<float4*>(t.Y + ro + q) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N_TILE, int K_CHUNKS, int EPI>
int launch_pipe(const TcArgs& t, int G, cudaStream_t st) {
    constexpr size_t B_BYTES = (size_t)N_TILE * TC_KC * 4 * K_CHUNKS;
    constexpr size_t smem = 1024 + ((B_BYTES + 1023) & ~(size_t)1023) + 2 * A_STAGE_BYTES + 2 * STG_BYTES + 3 * N_TILE * 4 + 64 * 5 * 4 + 128;
    static_assert(smem <= 227 * 1024, "pipelined tcgen05 Linear: shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_tc_pipe_kernel<N_TILE, K_CHUNKS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int m_tiles = dtqn_cdiv(t.a.Tg, TC_M);
    const int gy = (t.a.N / N_TILE) * G;
    int gx = 148 / gy; if (gx < 1) gx = 1; if (gx > m_tiles) gx = m_tiles;
    linear_tc_pipe_kernel<N_TILE, K_CHUNKS, EPI><<<dim3(gx, gy, 1), EPI == EPI_RES_LN ? PIPE_THREADS_LN : PIPE_THREADS, smem, st>>>(t);
    return 0;
}

// shapes whose whole [N_TILE x K] hi+lo weight image fits beside the pipeline buffers (<= 64 KB)
template <int EPI>
int launch_pipe_dispatch(const TcArgs& t, int G, int nt, cudaStream_t st, bool& handled) {
    handled = true;
    const int K = t.a.K;
    if constexpr (EPI == EPI_RES_LN) {
        if (t.a.N == 64 && K == 64) return launch_pipe<64, 1, EPI>(t, G, st);
        if (t.a.N == 64 && K == 256) return launch_pipe<64, 4, EPI>(t, G, st);
        if (t.a.N == 128 && K == 128) return launch_pipe<128, 2, EPI>(t, G, st);
    } else {
        if (nt == 64 && K == 64) return launch_pipe<64, 1, EPI>(t, G, st);
        if (nt == 128 && K == 64) return launch_pipe<128, 1, EPI>(t, G, st);
        if (nt == 192 && K == 64) return launch_pipe<192, 1, EPI>(t, G, st);
        if (nt == 256 && K == 64) return launch_pipe<256, 1, EPI>(t, G, st);
        if (nt == 128 && K == 128) return launch_pipe<128, 2, EPI>(t, G, st);
    }
    handled = false;
    return 0;
}

}  // namespace

static int g_tc_pipelined = 1;
extern "C" int dtqn_set_tc_pipelined(int32_t on) { g_tc_pipelined = on; return 0; }
bool tc_pipelined_enabled() { return g_tc_pipelined != 0; }


static int g_tc_fuse_ffn = 1;
extern "C" int dtqn_set_tc_fuse_ffn(int32_t on) { g_tc_fuse_ffn = on; return 0; }
bool tc_ffn_fused_enabled() { return g_tc_fuse_ffn != 0 && g_tc_pipelined != 0; }

int launch_ffn_tc(const float* X, float* Y, const GroupPtrs& P, int G, const uint8_t* const* packed, long long pk_w1,
                  long long pk_w2, long long b1_off, long long b2_off, long long gamma_off, long long beta_off, int Tg,
                  cudaStream_t st) {
    FfnArgs t{};
    t.X = X; t.Y = Y; t.P = P; t.pk_w1 = pk_w1; t.pk_w2 = pk_w2; t.b1_off = b1_off; t.b2_off = b2_off;
    t.gamma_off = gamma_off; t.beta_off = beta_off; t.Tg = Tg;
    for (int g = 0; g < G; ++g) t.packed[g] = packed[g];
    constexpr size_t smem = 1024 + 256 * TC_KC * 4 + 2 * 64 * TC_KC * 4 + 3 * A_STAGE_BYTES + (256 + 3 * 64) * 4 + 256;
    static_assert(smem <= 227 * 1024, "fused FFN: shared memory budget");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int m_tiles = dtqn_cdiv(Tg, TC_M);
    int gx = 148 / G; if (gx < 1) gx = 1; if (gx > m_tiles) gx = m_tiles;
    // HBM roofline of the fused pair: x1 read as operand and as residual, x2 written, both weight images once
    const double bytes = 4.0 * ((double)Tg * G * (64 + 64 + 64) + 2.0 * 256 * 64);
    prof_begin(PROF_LINEAR_TC, st);
    ffn_tc_kernel<<<dim3(gx, G, 1), FFN_THREADS, smem, st>>>(t);
    prof_end(PROF_LINEAR_TC, st, bytes);
    DTQN_LAUNCH_CHECK();
    return 0;
}

int tc_ntile(int N) {
    switch (N) {
        case 64: return 64; case 128: return 128; case 192: return 192; case 256: return 256;
        case 384: return 192; case 512: return 256; default: return 0;
    }
}

int tc_pack_table(const dtqn_net_cfg& c, const NetLayout& lay, TcPackTable& tab) {
    const int d = c.d_model;
    long long off = 0;
    int n = 0;
    auto add = [&](long long w_off, int N, int K) {
        TcPackEntry& e = tab.e[n++];
        e.w_off = w_off; e.N = N; e.K = K; e.n_tile = tc_ntile(N); e.pk_off = off;
        off += (long long)N * K * 4;
    };
    for (int i = 0; i < c.n_layers; ++i) {
        const LayerOff& l = lay.layer[i];
        add(l.in_w, 3 * d, d); add(l.out_w, d, d); add(l.f1_w, 4 * d, d); add(l.f2_w, d, 4 * d);
    }
    add(lay.h1_w, d, d);
    for (int i = 0; i < c.n_layers; ++i) {                      // acting: K|V rows and Q rows of in_proj as separate operands
        add(lay.layer[i].in_w + (long long)d * d, 2 * d, d);
        add(lay.layer[i].in_w, d, d);
    }
    // fused acting forward (act_fused.cu): one contiguous 240 KB image in consumption order -- layer-0 in_proj [192 x 64],
    // out_proj [64 x 64], ffn.0 as two [128 x 64] halves, ffn.2 as four [64 x 64] k-chunks, layer-1 in_proj [192 x 64]
    tab.act_img_off = -1;
    if (c.n_layers == 2 && d == 64) {
        auto addnt = [&](long long w_off, int N, int K, int nt) {
            TcPackEntry& e = tab.e[n++];
            e.w_off = w_off; e.N = N; e.K = K; e.n_tile = nt; e.pk_off = off;
            off += (long long)N * K * 4;
        };
        tab.act_img_off = off;
        addnt(lay.layer[0].in_w, 3 * d, d, 192); addnt(lay.layer[0].out_w, d, d, 64); addnt(lay.layer[0].f1_w, 4 * d, d, 128);
        addnt(lay.layer[0].f2_w, d, 4 * d, 64); addnt(lay.layer[1].in_w, 3 * d, d, 192);
    }
    // k-major fp32 copies for the sequence-resident training forward (net_seq.cu): per layer in_proj | out_proj | ffn.0 |
    // ffn.2 (12 d^2 floats), then the head's ffn.0
    tab.wt_off = -1;
    if (d == 64) {
        auto addt = [&](long long w_off, int N, int K) {
            TcPackEntry& e = tab.e[n++];
            e.w_off = w_off; e.N = N; e.K = K; e.n_tile = 0; e.pk_off = off;
            off += (long long)N * K * 4;
        };
        tab.wt_off = off;
        for (int i = 0; i < c.n_layers; ++i) {
            const LayerOff& l = lay.layer[i];
            addt(l.in_w, 3 * d, d); addt(l.out_w, d, d); addt(l.f1_w, 4 * d, d); addt(l.f2_w, d, 4 * d);
        }
        addt(lay.h1_w, d, d);
    }
    tab.n = n; tab.total_bytes = off;
    return 0;
}

int launch_linear_tc(const LinArgs& a, int epi, int G, const uint8_t* const* packed, long long pk_off, cudaStream_t st,
                     const TcEmbed* emb) {
    TcArgs t{};
    t.a = a; t.pk_off = pk_off;
    if (emb) t.emb = *emb;
    for (int g = 0; g < G; ++g) t.packed[g] = packed[g];
    const int nt = tc_ntile(a.N);
    if (!nt || a.K % TC_KC) return DTQN_E_UNSUPPORTED;
    int rc = DTQN_E_UNSUPPORTED;
    // HBM roofline of this launch: activations read once (K floats / token), output written once (N), the residual read
    // by the LayerNorm epilogue (N), the weight image once
    const double tc_bytes = 4.0 * ((double)a.Tg * G * (a.K + a.N + (epi == EPI_RES_LN ? a.N : 0)) + (double)a.N * a.K);
    prof_begin(PROF_LINEAR_TC, st);
    bool handled = false;
    if (emb && emb->mode && !g_tc_pipelined) return DTQN_E_UNSUPPORTED;   // only the pipelined kernel embeds on the fly
    if (g_tc_pipelined) {
        if (epi == EPI_RES_LN) rc = launch_pipe_dispatch<EPI_RES_LN>(t, G, nt, st, handled);
        else if (epi == EPI_BIAS) rc = launch_pipe_dispatch<EPI_BIAS>(t, G, nt, st, handled);
        else rc = launch_pipe_dispatch<EPI_BIAS_RELU>(t, G, nt, st, handled);
    }
    if (handled) {
        prof_end(PROF_LINEAR_TC, st, tc_bytes);
        if (rc) return rc;
        DTQN_LAUNCH_CHECK();
        return 0;
    }
    if (epi == EPI_RES_LN) {
        if (a.N == 64) rc = launch_one<64, EPI_RES_LN>(t, G, st);
        else if (a.N == 128) rc = launch_one<128, EPI_RES_LN>(t, G, st);
    } else if (epi == EPI_BIAS) {
        if (nt == 64) rc = launch_one<64, EPI_BIAS>(t, G, st);
        else if (nt == 128) rc = launch_one<128, EPI_BIAS>(t, G, st);
        else if (nt == 192) rc = launch_one<192, EPI_BIAS>(t, G, st);
        else rc = launch_one<256, EPI_BIAS>(t, G, st);
    } else {
        if (nt == 64) rc = launch_one<64, EPI_BIAS_RELU>(t, G, st);
        else if (nt == 128) rc = launch_one<128, EPI_BIAS_RELU>(t, G, st);
        else if (nt == 192) rc = launch_one<192, EPI_BIAS_RELU>(t, G, st);
        else rc = launch_one<256, EPI_BIAS_RELU>(t, G, st);
    }
    prof_end(PROF_LINEAR_TC, st, tc_bytes);
    if (rc) return rc;
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t dtqn_packed_bytes(const dtqn_net_cfg* cfg) {
    if (!cfg) return DTQN_E_ARG;
    NetLayout lay;
    int rc = net_layout(*cfg, lay);
    if (rc) return rc;
    if (cfg_is_variant(*cfg)) return 16;          // ablation-flag networks run the fp32 path of net_var.cu: no operand image
    TcPackTable tab;
    tc_pack_table(*cfg, lay, tab);
    return tab.total_bytes;
}

extern "C" int dtqn_pack_weights(const dtqn_net_cfg* cfg, const float* params, void* packed, void* stream) {
    if (!cfg || !params || !packed) return DTQN_E_ARG;
    NetLayout lay;
    int rc = net_layout(*cfg, lay);
    if (rc) return rc;
    if (cfg_is_variant(*cfg)) return 0;
    TcPackTable tab;
    tc_pack_table(*cfg, lay, tab);
    dim3 grid(16, tab.n);
    pack_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, (uint8_t*)packed, tab);
    DTQN_LAUNCH_CHECK();
    return 0;
}

extern "C" int dtqn_tc_error(void) {
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
    return v | act_fused_tc_error();
}
