// Fused acting forward of the 2-layer DTQN (DtqnAgent.get_action, dtqn/agents/dtqn.py:81-107 -> DTQN.forward,
// dtqn/networks/dtqn.py:181-216) for n_envs lockstep context windows: ONE persistent tcgen05 kernel takes the context
// ring to the two per-sequence vectors the rest of the final layer needs.  Per 128-row tile (two sequences of L <= 52
// tokens) every intermediate stays on the SM:
//
//   obs rows (context ring, utils/context.py window)  --embed (Linear(O,64) + position table, dtqn.py:181-199)-->  x0
//   x0 --in_proj (MMA)--> qkv (TMEM -> smem fp32) --causal attention (mma.sync TF32 hi/lo, attn_mma.cuh)--> o
//   o  --out_proj (MMA) -> ReLU -> +x0 -> LayerNorm1--> x1 --ffn.0 (MMA) -> ReLU -> ffn.2 (MMA) -> ReLU -> +x1 -> LayerNorm2--> x2
//   x2 --layer-1 in_proj (MMA)--> q | k | v --attention row of the LAST valid position only--> o_last
//                                                                                (transformer.py:63-78, causal mask :49-53)
// Outputs per sequence: x2[last] (residual of the final layer's LayerNorm1) and o_last; the n_seq-row remainder of the
// final layer and the Q head run on those (dtqn_forward).  HBM traffic per token: the 12-byte observation row in, nothing
// out -- against ~5.6 KB of fp32 activations per token for the kernel-per-GEMM path.
//
// Roles (576 threads, 1 CTA / SM, persistent over tiles):
//   warps 0-15  workers: embed, TMEM epilogues (thread = row, 16 columns each; LayerNorm reduced through smem),
//               attention (warp = (sequence, head)), bf16 hi/lo split into the K-major A operand
//   warp 16     lane 0 issues every tcgen05.mma (M = 128, N = 192 / 128 / 64 -- wide N amortises the A-operand fetch from
//               shared memory, which paces these K = 64 GEMMs; bf16 hi/lo split: 3 MMAs per k-step) + commits
//   warp 17     lane 0 streams the weight image (240 KB per tile, L2-resident) through a 64 KB shared-memory buffer with
//               cp.async.bulk (TMA) + mbarrier complete_tx, each block as soon as the MMAs reading its slots have retired
// TMEM (512 columns): [0,256) in_proj / ffn.0 accumulators (+ x0 stash), [256,320) out_proj / ffn.2, [320,512) final-layer in_proj.
// Tiles are serial inside a CTA except for one overlap: as soon as the final-layer in_proj has retired, the NEXT tile's
// embedding is published and its in_proj MMAs (own TMEM region) run under this tile's last-row attention.
// Measured: profiles/README.md (241 us for 4096 x 50 tokens, phase timeline, ncu --set full).
#include "net.cuh"
#include "prof.cuh"
#include "linear_tc.cuh"
#include "tc_common.cuh"
#include "attn_mma.cuh"

namespace {

constexpr int AF_WORKERS = 512;
constexpr int AF_THREADS = AF_WORKERS + 64;
constexpr uint32_t AF_WCHUNK = 64 * TC_KC * 4;    // [64 x 64] hi + lo = one 16 KB slot of the weight buffer
constexpr uint32_t AF_WBYTES = 4 * AF_WCHUNK;     // weight buffer: 4 slots
// weight image (tc_pack_table): in_proj0 [192 x 64] | out_proj0 [64 x 64] | ffn.0 as two [128 x 64] halves | ffn.2 as four
// [64 x 64] k-chunks | in_proj1 [192 x 64]; every block = { hi[8 k-chunks][N][8 bf16], lo[...] }
constexpr uint32_t IMG_IN0 = 0, IMG_OUT = 3 * AF_WCHUNK, IMG_F1 = 4 * AF_WCHUNK, IMG_F2 = 8 * AF_WCHUNK, IMG_IN1 = 12 * AF_WCHUNK;
constexpr int AF_QKV_ROWS = 116;                  // rows of the staged q|k|v tile (L + 64 <= 116)
constexpr int AF_MAX_L = 52;
// A operand tile [128 x 64] bf16, K-major SWIZZLE_NONE: 8-element k-chunks AF_CS bytes apart, UNPADDED so every 8 x 16 B core
// matrix is one aligned 128-byte line for the tensor core's operand fetch (a warp writes 32 consecutive rows of one chunk =
// 512 contiguous bytes, so the stores are conflict-free without padding)
constexpr int AF_CS = TC_M * 16, AF_AHALF = (TC_KC / 8) * AF_CS, AF_ASTAGE = 2 * AF_AHALF;
constexpr uint32_t AF_R_BYTES = AF_QKV_ROWS * ATT_LD * 4;      // q|k|v tile; the 2-stage hidden-operand ring aliases it
static_assert(AF_R_BYTES >= 2 * AF_ASTAGE, "hidden operand ring must fit in the qkv region");

// parameter vectors staged in shared memory (float offsets)
enum { P_INB0 = 0, P_OUTB0 = 192, P_LN1W = 256, P_LN1B = 320, P_F1B = 384, P_F2B = 640, P_LN2W = 704, P_LN2B = 768,
       P_INB1 = 832, P_EW = 1024, P_EB = 1280, P_POS = 1344 /* [L][68] position table, rows padded: conflict-free float4 per row */, P_TOTAL = 1344 + AF_MAX_L * 68 };

// weight buffer hand-offs (each barrier completes once per tile): F_* "landed" (TMA complete_tx), E_* "consumed" (tcgen05.commit)
enum { F_IN0 = 0, F_OUT, F_F1A, F_F1B, F_F2 /* +k */, F_IN1 = F_F2 + 4, E_IN0, E_OUT, E_F1A, E_F1B, E_F2K2, E_F2K3, E_IN1,
       B_AX0, B_AO, B_AX1, B_AX2, B_ACC_QKV, B_ACC_OUT, B_ACC_F2, B_ACC_L1, B_ACC1 /* +half */,
       B_A2_FULL = B_ACC1 + 2 /* +s */, B_A2_EMPTY = B_A2_FULL + 2 /* +s */, B_COUNT = B_A2_EMPTY + 2 };

constexpr size_t AF_SMEM = 1024 + AF_WBYTES + AF_ASTAGE + AF_R_BYTES +
                           P_TOTAL * 4 + 2 * 4 * 128 * 8 + 2 * 128 * 4 * 4 + 16 + 256 + B_COUNT * 8 + 16;
static_assert(AF_SMEM <= 227 * 1024, "fused acting forward: shared memory budget");

struct ActFusedArgs {
    GroupPtrs P;
    GroupSrc S;
    const uint8_t* img[DTQN_MAX_GROUPS];           // 240 KB weight image of each group's network (IMG_* blocks)
    long long emb_w, emb_b, pos;
    LayerOff l0, l1;
    int O, n_seq, L;
    float obs_mask;
    float* xl;                                     // [G * n_seq, 64]  x2 at the last valid position
    float* ol;                                     // [G * n_seq, 64]  final-layer attention output of that position
};

// optional phase timeline (tools/prof_act.py --timeline): clock64 stamps of worker thread 0 of CTA 0, 32 per tile
__device__ long long* g_af_dbg = nullptr;

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(AF_WORKERS) : "memory"); }

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 12 MMAs: D[128 x 64] (+)= A[128 x 64] W[64 x 64]^T with the bf16 hi/lo split (hi*hi + hi*lo + lo*hi).  The operand
// descriptors are built once per kernel (the single issuing thread is the pacing resource of every MMA phase); an smem
// address offset is added to the start-address field, which never carries out of its 14 bits.
struct ChunkDesc { uint64_t a_hi, a_lo; };
// N = rows of the weight block (its k-chunks are N * 16 bytes apart, its lo half N * 128 bytes after the hi half)
template <int N>
__device__ __forceinline__ void mma_block(uint32_t d_tmem, const ChunkDesc& a0, uint64_t b_hi0, bool acc_first) {
    constexpr uint64_t A_K16 = (2 * AF_CS) >> 4, B_K16 = (2 * N * 16) >> 4, B_LO = (N * TC_KC * 2) >> 4;
    const uint32_t idesc = umma_idesc(TC_M, N);
#pragma unroll
    for (int k16 = 0; k16 < TC_KC / 16; ++k16) {
        const uint64_t a_hi = a0.a_hi + k16 * A_K16, a_lo = a0.a_lo + k16 * A_K16;
        const uint64_t b_hi = b_hi0 + k16 * B_K16, b_lo = b_hi + B_LO;
        umma_bf16(d_tmem, a_hi, b_hi, idesc, (acc_first || k16) ? 1u : 0u);
        umma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
        umma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
    }
}

__global__ void __launch_bounds__(AF_THREADS, 1)
act_fused_kernel(ActFusedArgs t) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;
    const int L = t.L;

    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic only: keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t* sW = base;                                            // weight ring
    uint8_t* sA = sW + AF_WBYTES;                      // A operand (K = 64): x0 -> o -> x1 -> x2
    uint8_t* sR = sA + AF_ASTAGE;
    float* sQKV = reinterpret_cast<float*>(sR);                    // [AF_QKV_ROWS][ATT_LD]
    uint8_t* sA2 = sR;                                             // hidden operand ring (2 stages), aliases sQKV
    float* sPar = reinterpret_cast<float*>(sR + AF_R_BYTES);
    float* sRedA = sPar + P_TOTAL;                                 // [2][4][128] float2 LayerNorm partials (mean, M2)
    float* sObs = sRedA + 2 * 4 * 128 * 2;                                // [2][128][4] observation rows of the tile (double buffered)
    int* sMeta = reinterpret_cast<int*>(sObs + 2 * 128 * 4);       // [2][2] valid length of the tile's two sequences
    uint8_t* sFlag = reinterpret_cast<uint8_t*>(sMeta + 4);        // [2][128] 1 = row belongs to a real sequence
    uint64_t* bars = reinterpret_cast<uint64_t*>(sFlag + 256);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    volatile int* s_fail = reinterpret_cast<volatile int*>(s_tmem + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // bounded wait that never makes a role skip a hand-off: after the first time-out everybody free-runs to the end
    auto WAIT = [&](int b, uint32_t parity) {
        if (mbar_try_wait(BAR(b), parity)) return;             // fast path: already complete
        if (*s_fail) return;
        if (!mbar_wait(BAR(b), parity)) *s_fail = 1;
    };

    const float* p = t.P.p[g];
    for (int e = tid; e < P_TOTAL; e += AF_THREADS) {
        float v;
        if (e < P_OUTB0) v = __ldg(p + t.l0.in_b + e);
        else if (e < P_LN1W) v = __ldg(p + t.l0.out_b + (e - P_OUTB0));
        else if (e < P_LN1B) v = __ldg(p + t.l0.ln1_w + (e - P_LN1W));
        else if (e < P_F1B) v = __ldg(p + t.l0.ln1_b + (e - P_LN1B));
        else if (e < P_F2B) v = __ldg(p + t.l0.f1_b + (e - P_F1B));
        else if (e < P_LN2W) v = __ldg(p + t.l0.f2_b + (e - P_F2B));
        else if (e < P_LN2B) v = __ldg(p + t.l0.ln2_w + (e - P_LN2W));
        else if (e < P_INB1) v = __ldg(p + t.l0.ln2_b + (e - P_LN2B));
        else if (e < P_EW) v = __ldg(p + t.l1.in_b + (e - P_INB1));
        else if (e < P_EB) { const int c = (e - P_EW) >> 2, k = (e - P_EW) & 3; v = k < t.O ? __ldg(p + t.emb_w + c * t.O + k) : 0.f; }
        else if (e < P_POS) v = __ldg(p + t.emb_b + (e - P_EB));
        else { const int r = (e - P_POS) / 68, c = (e - P_POS) % 68; v = (r < L && c < 64) ? __ldg(p + t.pos + r * 64 + c) : 0.f; }
        sPar[e] = v;
    }
    if (tid == 0) {
        for (int b = F_IN0; b <= E_IN1; ++b) mbar_init(BAR(b), 1);
        mbar_init(BAR(B_AX0), AF_WORKERS); mbar_init(BAR(B_AO), AF_WORKERS);
        mbar_init(BAR(B_AX1), AF_WORKERS); mbar_init(BAR(B_AX2), AF_WORKERS);
        mbar_init(BAR(B_ACC_QKV), 1); mbar_init(BAR(B_ACC_OUT), 1); mbar_init(BAR(B_ACC_F2), 1); mbar_init(BAR(B_ACC_L1), 1);
        for (int c = 0; c < 2; ++c) mbar_init(BAR(B_ACC1 + c), 1);
        for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_A2_FULL + s), AF_WORKERS); mbar_init(BAR(B_A2_EMPTY + s), 1); }
        *s_fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    // TMEM columns: [0,192) layer-0 q|k|v, then [0,256) ffn.0 hidden; [192,256) doubles as the stash of the next x0 (residual
    // of LayerNorm1) while free; [256,320) out_proj, then ffn.2; [320,512) final-layer q|k|v (so the NEXT tile's in_proj can
    // run while this tile's last-row attention still reads it)
    constexpr uint32_t TM_STASH = 192, TM_OUT = 256, TM_F2 = 256, TM_L1 = 320;

    const int tiles = (t.n_seq + 1) >> 1;
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp < 16) {
        // =============================================== workers ===============================================
        const int q4 = warp & 3, cq = warp >> 2;
        const int row = q4 * 32 + lane, c0 = cq * 16;
        const int sl = row >= 2 * L ? 2 : (row >= L ? 1 : 0);      // which of the tile's sequences this row belongs to (2: padding)
        const int jrow = row - sl * L;
        const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
        const dtqn_obs_src& src = t.S.s[g];
        const float qs = 0.35355339059327373f * 1.4426950408889634f;   // log2(e) / sqrt(head_dim = 8)
        const bool loader = cq == 0;                               // warps 0-3: one thread per tile row fetches its observation

        // observation row of `row` in tile `tile` (context window position jrow of sequence 2*tile + sl)
        int pf_ts = 0;
        auto obs_ts = [&](int tile) {                              // stage 1: the sequence's timestep
            const int seq = tile * 2 + sl;
            pf_ts = (sl < 2 && seq < t.n_seq && src.timestep) ? __ldg(src.timestep + seq) : 0;
        };
        float pf_o[4]; int pf_flag = 0, pf_n = 0;
        auto obs_rows = [&](int tile) {                            // stage 2: the ring row (depends on the timestep)
            const int seq = tile * 2 + sl;
            pf_o[0] = pf_o[1] = pf_o[2] = pf_o[3] = 0.f; pf_flag = 0; pf_n = 0;
            if (sl < 2 && seq < t.n_seq) {
                int rr = jrow; bool ok = true; pf_n = L;
                if (src.timestep) {
                    const int n = min(src.ring_len, pf_ts + 1);
                    ok = jrow < n;
                    rr = ok ? (pf_ts + 1 - n + jrow) % src.ring_len : 0;
                    pf_n = min(n, L);
                }
                const float* o = src.obs + (long long)seq * src.seq_stride + (long long)rr * t.O;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < t.O) pf_o[k] = ok ? __ldg(o + k) : t.obs_mask;
                pf_flag = 1;
            }
        };
        auto obs_store = [&](int buf) {                            // stage 3: publish in shared memory
            *reinterpret_cast<float4*>(sObs + (buf * 128 + row) * 4) = make_float4(pf_o[0], pf_o[1], pf_o[2], pf_o[3]);
            sFlag[buf * 128 + row] = (uint8_t)pf_flag;
            if (sl < 2 && jrow == 0) sMeta[buf * 2 + sl] = pf_n;
        };
        // two 8-column groups of this thread's 16 columns -> bf16 hi/lo in the canonical K-major layout
        auto store_a = [&](uint8_t* dstA, const float (&y)[16]) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                float x8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) x8[e] = y[hlf * 8 + e];
                uint4 hi, lo;
                split8(x8, hi, lo);
                uint8_t* d = dstA + (cq * 2 + hlf) * AF_CS + row * 16;
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + AF_AHALF) = lo;
            }
        };
        // accumulator [128 x 192] (+ in_proj bias, q columns pre-scaled for the base-2 softmax) -> staged fp32 q|k|v tile
        auto dump_qkv = [&](int bias_off, uint32_t tm_base) {
            uint32_t r[3][16];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) tmem_ld16_issue(tlane + tm_base + (uint32_t)(cq * 48 + cc * 16), r[cc]);
            tmem_ld_wait();
            tc_fence_before();
            if (row < AF_QKV_ROWS) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const int col = cq * 48 + cc * 16;
                    const float sc = col < 64 ? qs : 1.f;
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(sPar + bias_off + col + q);
                        *reinterpret_cast<float4*>(sQKV + row * ATT_LD + col + q) =
                            make_float4((__uint_as_float(r[cc][q]) + b.x) * sc, (__uint_as_float(r[cc][q + 1]) + b.y) * sc,
                                        (__uint_as_float(r[cc][q + 2]) + b.z) * sc, (__uint_as_float(r[cc][q + 3]) + b.w) * sc);
                    }
                }
            }
        };
        // final layer: only k|v of every row and q of each sequence's last valid row are needed (TMEM reads pace the dump)
        auto dump_kv_qlast = [&](int bias_off, uint32_t tm_base, int buf) {
            uint32_t r[2][16];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) tmem_ld16_issue(tlane + tm_base + (uint32_t)(64 + cq * 32 + cc * 16), r[cc]);
            tmem_ld_wait();
            if (row < AF_QKV_ROWS) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = 64 + cq * 32 + cc * 16;
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(sPar + bias_off + col + q);
                        *reinterpret_cast<float4*>(sQKV + row * ATT_LD + col + q) =
                            make_float4(__uint_as_float(r[cc][q]) + b.x, __uint_as_float(r[cc][q + 1]) + b.y,
                                        __uint_as_float(r[cc][q + 2]) + b.z, __uint_as_float(r[cc][q + 3]) + b.w);
                    }
                }
            }
            if (cq == 0) {                                         // one warp per lane quarter: does a last row live here?
                const int n0 = sMeta[buf * 2], n1 = sMeta[buf * 2 + 1];
                const int r0 = n0 > 0 ? n0 - 1 : -1, r1 = n1 > 0 ? L + n1 - 1 : -1;
                const bool here0 = r0 >= 0 && (r0 >> 5) == q4, here1 = r1 >= 0 && (r1 >> 5) == q4;
                if (here0 || here1) {                              // warp-uniform
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        tmem_ld16_issue(tlane + tm_base + (uint32_t)(hf * 32), r[0]);
                        tmem_ld16_issue(tlane + tm_base + (uint32_t)(hf * 32 + 16), r[1]);
                        tmem_ld_wait();
                        if (row == r0 || row == r1) {
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                                for (int q = 0; q < 16; q += 4) {
                                    const int col = hf * 32 + cc * 16 + q;
                                    const float4 b = *reinterpret_cast<const float4*>(sPar + bias_off + col);
                                    *reinterpret_cast<float4*>(sQKV + row * ATT_LD + col) =
                                        make_float4((__uint_as_float(r[cc][q]) + b.x) * qs, (__uint_as_float(r[cc][q + 1]) + b.y) * qs,
                                                    (__uint_as_float(r[cc][q + 2]) + b.z) * qs, (__uint_as_float(r[cc][q + 3]) + b.w) * qs);
                                }
                        }
                    }
                }
            }
            tc_fence_before();
        };
        // y = LayerNorm(x + relu(acc + b)) over the 64 columns of `row` (4 threads x 16 columns, reduced through smem);
        // the residual x lives in this thread's registers from the phase that produced it and is replaced by y
        int ln_parity = 0;                                         // the two LayerNorms of a tile alternate reduction buffers
        auto res_ln = [&](uint32_t tm_col, int b_off, int gw_off, int gb_off, float (&y)[16]) {
            float v[16];
            tmem_ld16(tlane + tm_col + (uint32_t)c0, v);
            tc_fence_before();
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sPar + b_off + c0 + q);
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { y[q + e] += fmaxf(v[q + e] + bv[e], 0.f); s += y[q + e]; }
            }
            // mean / variance of the row from four 16-column partials (mean_i, M2_i), combined exactly (Chan et al.)
            const float mi = s * (1.f / 16.f);
            float m2 = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) { const float dl = y[e] - mi; m2 = fmaf(dl, dl, m2); }
            float2* red = reinterpret_cast<float2*>(sRedA) + (size_t)(ln_parity * 4) * 128;
            red[cq * 128 + row] = make_float2(mi, m2);
            worker_bar();
            const float2 r0 = red[row], r1 = red[128 + row], r2 = red[256 + row], r3 = red[384 + row];
            const float mean = ((r0.x + r1.x) + (r2.x + r3.x)) * 0.25f;
            const float d0 = r0.x - mean, d1 = r1.x - mean, d2 = r2.x - mean, d3 = r3.x - mean;
            const float var = (((r0.y + r1.y) + (r2.y + r3.y)) + 16.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))) * (1.f / 64.f);
            const float rstd = 1.0f / sqrtf(var + 1e-5f);
            ln_parity ^= 1;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(sPar + gw_off + c0 + q);
                const float4 e4 = *reinterpret_cast<const float4*>(sPar + gb_off + c0 + q);
                y[q] = (y[q] - mean) * rstd * g4.x + e4.x; y[q + 1] = (y[q + 1] - mean) * rstd * g4.y + e4.y;
                y[q + 2] = (y[q + 2] - mean) * rstd * g4.z + e4.z; y[q + 3] = (y[q + 3] - mean) * rstd * g4.w + e4.w;
            }
        };

        // token embedding of this thread's 16 columns of `row`: x0 = W_e obs + b_e + pos[j]  (rows outside a real sequence: 0)
        auto embed_row = [&](int buf, float (&x)[16]) {
            const float4 o4 = *reinterpret_cast<const float4*>(sObs + (buf * 128 + row) * 4);
            const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
            const bool real = sFlag[buf * 128 + row] != 0;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (real) pv = *reinterpret_cast<const float4*>(sPar + P_POS + jrow * 68 + c0 + q);
                const float pvv[4] = {pv.x, pv.y, pv.z, pv.w};
                const float4 b4 = *reinterpret_cast<const float4*>(sPar + P_EB + c0 + q);
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sPar + P_EW + (c0 + q + e) * 4);
                    float acc = bv[e];
                    acc = fmaf(ov[0], w4.x, acc); acc = fmaf(ov[1], w4.y, acc); acc = fmaf(ov[2], w4.z, acc); acc = fmaf(ov[3], w4.w, acc);
                    x[q + e] = real ? acc + pvv[e] : 0.f;
                }
            }
        };

        // token embedding of a tile -> A operand of its in_proj; the fp32 row is parked in TMEM until LayerNorm1 needs it as residual
        auto embed_publish = [&](int buf) {
            float x0[16];
            embed_row(buf, x0);
            store_a(sA, x0);
            tmem_st16(tlane + TM_STASH + (uint32_t)c0, x0);
            tmem_st_wait();
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(BAR(B_AX0));
        };

        if (my_tiles > 0 && loader) { obs_ts((int)blockIdx.x); obs_rows((int)blockIdx.x); obs_store(0); }
        worker_bar();
        if (my_tiles > 0) embed_publish(0);

        long long* dbg = (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) ? g_af_dbg : nullptr;
        auto STAMP = [&](int i, int k) { if (dbg && i < 8) dbg[i * 32 + k] = clock64(); };
        for (int i = 0; i < my_tiles; ++i) {
            STAMP(i, 0);
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const int tile_next = tile + (int)gridDim.x;
            const bool has_next = i + 1 < my_tiles;
            const uint32_t ph = (uint32_t)(i & 1);
            const int buf = i & 1;
            float xres[16];
            STAMP(i, 1);
            // ---- layer-0 q|k|v: TMEM -> shared memory ----
            WAIT(B_ACC_QKV, ph);
            STAMP(i, 2);
            tc_fence_after();
            dump_qkv(P_INB0, 0);
            worker_bar();
            STAMP(i, 3);
            // ---- causal attention, warp = (sequence, head); o -> A operand ----
            {
                const int s = warp >> 3, h = warp & 7;
                const int n = sMeta[buf * 2 + s];
                if (n > 0) {
                    const int rbase = s * L;
                    att_head(sQKV + rbase * ATT_LD, h, n, lane, [&](int r, int col, float v0, float v1) {
                        if (r < n) {
                            uint32_t hi, lo;
                            split2(v0, v1, hi, lo);
                            uint8_t* d = sA + (col >> 3) * AF_CS + (rbase + r) * 16 + (col & 7) * 2;
                            *reinterpret_cast<uint32_t*>(d) = hi;
                            *reinterpret_cast<uint32_t*>(d + AF_AHALF) = lo;
                        }
                    });
                }
                STAMP(i, 4);
                fence_async_smem();
                mbar_arrive(BAR(B_AO));
                // next tile's observation rows: the dependent timestep -> ring-row loads hide behind the out_proj wait
                if (has_next && loader) { obs_ts(tile_next); obs_rows(tile_next); obs_store(buf ^ 1); }
            }
            // ---- out_proj -> ReLU -> +x0 -> LayerNorm1 -> x1 ----
            WAIT(B_ACC_OUT, ph);
            STAMP(i, 5);
            tc_fence_after();
            tmem_ld16(tlane + TM_STASH + (uint32_t)c0, xres);    // residual x0, parked by embed_publish
            res_ln(TM_OUT, P_OUTB0, P_LN1W, P_LN1B, xres);        // xres: x0 -> x1
            store_a(sA, xres);
            fence_async_smem();
            mbar_arrive(BAR(B_AX1));
            STAMP(i, 6);
            // ---- ffn.0 accumulator chunk c -> +b1 -> ReLU -> hi/lo -> hidden operand ring ----
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int m = 2 * i + (c >> 1), s_ = c & 1;
                WAIT(B_ACC1 + (c >> 1), ph);
                STAMP(i, 7 + 3 * c);
                tc_fence_after();
                float v[16];
                tmem_ld16(tlane + (uint32_t)(c * 64 + c0), v);
                tc_fence_before();
                if (m >= 1) WAIT(B_A2_EMPTY + s_, (uint32_t)((m - 1) & 1));
                STAMP(i, 8 + 3 * c);
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sPar + P_F1B + c * 64 + c0 + q);
                    v[q] = fmaxf(v[q] + b4.x, 0.f); v[q + 1] = fmaxf(v[q + 1] + b4.y, 0.f);
                    v[q + 2] = fmaxf(v[q + 2] + b4.z, 0.f); v[q + 3] = fmaxf(v[q + 3] + b4.w, 0.f);
                }
                store_a(sA2 + s_ * AF_ASTAGE, v);
                fence_async_smem();
                mbar_arrive(BAR(B_A2_FULL + s_));
                STAMP(i, 9 + 3 * c);
            }
            // ---- ffn.2 -> ReLU -> +x1 -> LayerNorm2 -> x2 ----
            WAIT(B_ACC_F2, ph);
            STAMP(i, 19);
            tc_fence_after();
            {
                res_ln(TM_F2, P_F2B, P_LN2W, P_LN2B, xres);       // xres: x1 -> x2
                store_a(sA, xres);
                if (sl < 2) {
                    const int n = sMeta[buf * 2 + sl];
                    if (n > 0 && jrow == n - 1) {                  // residual of the final layer's LayerNorm1
                        float* dst = t.xl + ((long long)g * t.n_seq + tile * 2 + sl) * 64 + c0;
#pragma unroll
                        for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(xres[q], xres[q + 1], xres[q + 2], xres[q + 3]);
                    }
                }
                fence_async_smem();
                mbar_arrive(BAR(B_AX2));
            }
            STAMP(i, 20);
            // ---- final layer: q|k|v -> shared memory, attention row of the last valid position ----
            WAIT(B_ACC_L1, ph);
            STAMP(i, 21);
            tc_fence_after();
            // the A operand and TMEM [0,256) are free now: publish the next tile's embedding so its in_proj MMAs (and the weight
            // loads behind them) run under this tile's last-row attention
            if (has_next) embed_publish(buf ^ 1);
            dump_kv_qlast(P_INB1, TM_L1, buf);
            worker_bar();
            STAMP(i, 22);
            {
                const int s = warp >> 3, h = warp & 7;
                const int n = sMeta[buf * 2 + s];
                if (n > 0) {
                    const float* sb = sQKV + s * L * ATT_LD;
                    const float4 qa = *reinterpret_cast<const float4*>(sb + (n - 1) * ATT_LD + h * 8);
                    const float4 qb = *reinterpret_cast<const float4*>(sb + (n - 1) * ATT_LD + h * 8 + 4);
                    float sc[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int j = lane + 32 * r;
                        sc[r] = -INFINITY;
                        if (j < n) {
                            const float4 ka = *reinterpret_cast<const float4*>(sb + j * ATT_LD + 64 + h * 8);
                            const float4 kb = *reinterpret_cast<const float4*>(sb + j * ATT_LD + 64 + h * 8 + 4);
                            float a = qa.x * ka.x;
                            a = fmaf(qa.y, ka.y, a); a = fmaf(qa.z, ka.z, a); a = fmaf(qa.w, ka.w, a);
                            a = fmaf(qb.x, kb.x, a); a = fmaf(qb.y, kb.y, a); a = fmaf(qb.z, kb.z, a); a = fmaf(qb.w, kb.w, a);
                            sc[r] = a;
                        }
                    }
                    const float mx = warp_max_f(fmaxf(sc[0], sc[1]));
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, l = 0.f;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int j = lane + 32 * r;
                        if (j < n) {
                            const float pj = att_ex2(sc[r] - mx);
                            l += pj;
                            const float4 va = *reinterpret_cast<const float4*>(sb + j * ATT_LD + 128 + h * 8);
                            const float4 vb = *reinterpret_cast<const float4*>(sb + j * ATT_LD + 128 + h * 8 + 4);
                            acc[0] = fmaf(pj, va.x, acc[0]); acc[1] = fmaf(pj, va.y, acc[1]); acc[2] = fmaf(pj, va.z, acc[2]); acc[3] = fmaf(pj, va.w, acc[3]);
                            acc[4] = fmaf(pj, vb.x, acc[4]); acc[5] = fmaf(pj, vb.y, acc[5]); acc[6] = fmaf(pj, vb.z, acc[6]); acc[7] = fmaf(pj, vb.w, acc[7]);
                        }
                    }
                    l = warp_sum_f(l);
                    // 8 sums over 32 lanes by halving: after the three exchanges lane (b4 b3 b2 . .) holds column 4 b4 + 2 b3 + b2
                    float a4[4], a2[2], a1;
                    {
                        const bool up = lane & 16;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float send = up ? acc[c] : acc[c + 4];
                            a4[c] = (up ? acc[c + 4] : acc[c]) + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
                    }
                    {
                        const bool up = lane & 8;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const float send = up ? a4[c] : a4[c + 2];
                            a2[c] = (up ? a4[c + 2] : a4[c]) + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                    }
                    {
                        const bool up = lane & 4;
                        const float send = up ? a2[0] : a2[1];
                        a1 = (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                    if ((lane & 3) == 0) {
                        const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        t.ol[((long long)g * t.n_seq + tile * 2 + s) * 64 + h * 8 + c] = a1 / l;
                    }
                }
            }
            STAMP(i, 23);
            // The next tile's in_proj has usually retired by now (it ran under this phase), so nothing else orders the next
            // q|k|v dump -- which overwrites the staged tile -- after the slowest warp's last-row attention reads of it.
            worker_bar();
        }
    } else if (warp == 16) {
        // =============================================== MMA issuer ===============================================
        if (lane == 0) {
            const ChunkDesc dA{umma_desc(smem_u32(sA), AF_CS, 128), umma_desc(smem_u32(sA) + AF_AHALF, AF_CS, 128)};
            const ChunkDesc dA2[2] = {
                {umma_desc(smem_u32(sA2), AF_CS, 128), umma_desc(smem_u32(sA2) + AF_AHALF, AF_CS, 128)},
                {umma_desc(smem_u32(sA2) + AF_ASTAGE, AF_CS, 128), umma_desc(smem_u32(sA2) + AF_ASTAGE + AF_AHALF, AF_CS, 128)}};
            // B descriptors (hi half) of the blocks as they sit in the weight buffer
            const uint32_t sW_u = smem_u32(sW);
            const uint64_t dW192 = umma_desc(sW_u, 192 * 16, 128);                         // in_proj: slots 0-2
            const uint64_t dWout = umma_desc(sW_u + 3 * AF_WCHUNK, 64 * 16, 128);          // out_proj: slot 3
            const uint64_t dWf1a = umma_desc(sW_u, 128 * 16, 128), dWf1b = umma_desc(sW_u + 2 * AF_WCHUNK, 128 * 16, 128);
            const uint64_t dWf2 = umma_desc(sW_u, 64 * 16, 128);                           // ffn.2 k-chunk k: slot k
            long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0) ? g_af_dbg : nullptr;
            auto STAMP = [&](int i, int k) { if (dbg && i < 8) dbg[i * 32 + k] = clock64(); };
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t ph = (uint32_t)(i & 1);
                WAIT(B_AX0, ph);
                STAMP(i, 24);
                WAIT(F_IN0, ph);
                tc_fence_after();
                mma_block<192>(tmem, dA, dW192, false);
                umma_commit(BAR(E_IN0));
                umma_commit(BAR(B_ACC_QKV));
                STAMP(i, 25);
                WAIT(B_AO, ph);
                STAMP(i, 26);
                WAIT(F_OUT, ph);
                tc_fence_after();
                mma_block<64>(tmem + TM_OUT, dA, dWout, false);
                umma_commit(BAR(E_OUT));
                umma_commit(BAR(B_ACC_OUT));
                STAMP(i, 27);
                WAIT(B_AX1, ph);
                STAMP(i, 28);
                WAIT(F_F1A, ph);
                tc_fence_after();
                mma_block<128>(tmem, dA, dWf1a, false);
                umma_commit(BAR(E_F1A));
                umma_commit(BAR(B_ACC1));
                WAIT(F_F1B, ph);
                tc_fence_after();
                mma_block<128>(tmem + 128, dA, dWf1b, false);
                umma_commit(BAR(E_F1B));
                umma_commit(BAR(B_ACC1 + 1));
                for (int c = 0; c < 4; ++c) {
                    const int m = 2 * i + (c >> 1), s_ = c & 1;
                    WAIT(B_A2_FULL + s_, (uint32_t)(m & 1));
                    WAIT(F_F2 + c, ph);
                    tc_fence_after();
                    mma_block<64>(tmem + TM_F2, dA2[s_], dWf2 + (uint64_t)c * (AF_WCHUNK >> 4), c > 0);
                    umma_commit(BAR(B_A2_EMPTY + s_));
                    if (c == 2) umma_commit(BAR(E_F2K2));
                    if (c == 3) umma_commit(BAR(E_F2K3));
                }
                umma_commit(BAR(B_ACC_F2));
                STAMP(i, 29);
                WAIT(B_AX2, ph);
                STAMP(i, 30);
                WAIT(F_IN1, ph);
                tc_fence_after();
                mma_block<192>(tmem + TM_L1, dA, dW192, false);
                umma_commit(BAR(E_IN1));
                umma_commit(BAR(B_ACC_L1));
                STAMP(i, 31);
            }
        }
    } else {
        // =============================================== weight producer ===============================================
        // Every block is loaded as soon as the MMAs that read the slots it overwrites have retired (one E_* wait each):
        //   in_proj0 -> slots 0-2 | out_proj -> 3 | ffn.0 a -> 0-1, b -> 2-3 | ffn.2 k -> slot k | in_proj1 -> 0-2
        if (lane == 0) {
            const uint8_t* img = t.img[g];
            const uint32_t sW_u = smem_u32(sW);
            auto load = [&](int bar, uint32_t dst_off, uint32_t img_off, uint32_t bytes) {
                mbar_expect_tx(BAR(bar), bytes);
                for (uint32_t o = 0; o < bytes; o += 32768u)
                    bulk_g2s(sW_u + dst_off + o, img + img_off + o, (bytes - o) < 32768u ? (bytes - o) : 32768u, BAR(bar));
            };
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t ph = (uint32_t)(i & 1);
                if (i > 0) WAIT(E_IN1, ph ^ 1u);
                load(F_IN0, 0, IMG_IN0, 3 * AF_WCHUNK);
                if (i > 0) WAIT(E_F2K3, ph ^ 1u);
                load(F_OUT, 3 * AF_WCHUNK, IMG_OUT, AF_WCHUNK);
                WAIT(E_IN0, ph);
                load(F_F1A, 0, IMG_F1, 2 * AF_WCHUNK);
                WAIT(E_OUT, ph);
                load(F_F1B, 2 * AF_WCHUNK, IMG_F1 + 2 * AF_WCHUNK, 2 * AF_WCHUNK);
                WAIT(E_F1A, ph);
                load(F_F2 + 0, 0, IMG_F2, AF_WCHUNK);
                load(F_F2 + 1, AF_WCHUNK, IMG_F2 + AF_WCHUNK, AF_WCHUNK);
                WAIT(E_F1B, ph);
                load(F_F2 + 2, 2 * AF_WCHUNK, IMG_F2 + 2 * AF_WCHUNK, AF_WCHUNK);
                load(F_F2 + 3, 3 * AF_WCHUNK, IMG_F2 + 3 * AF_WCHUNK, AF_WCHUNK);
                WAIT(E_F2K2, ph);
                load(F_IN1, 0, IMG_IN1, 3 * AF_WCHUNK);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

static int g_act_fused = 1;
extern "C" int dtqn_set_act_fused(int32_t on) { g_act_fused = on; return 0; }

bool act_fused_supported(const dtqn_net_cfg& c, int L) {
    return g_act_fused && c.n_layers == 2 && c.d_model == 64 && c.n_heads == 8 && !c.discrete && c.obs_dim <= 4 && L >= 3 &&
           L <= AF_MAX_L;
}

int launch_act_fused(const dtqn_net_cfg& c, const NetLayout& lay, const GroupPtrs& P, const GroupSrc& S, int G,
                     const uint8_t* const* packed, long long img_off, int n_seq, int L, float* xl, float* ol, cudaStream_t st) {
    ActFusedArgs t{};
    t.P = P; t.S = S;
    for (int g = 0; g < G; ++g) t.img[g] = packed[g] + img_off;
    t.emb_w = lay.emb_w; t.emb_b = lay.emb_b; t.pos = lay.pos; t.l0 = lay.layer[0]; t.l1 = lay.layer[1];
    t.O = c.obs_dim; t.n_seq = n_seq; t.L = L; t.obs_mask = S.s[0].obs_mask; t.xl = xl; t.ol = ol;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(act_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int tiles = (n_seq + 1) / 2;
    int gx = 148 / G; if (gx < 1) gx = 1; if (gx > tiles) gx = tiles;
    // algorithmic FLOPs: embed + full layer 0 + layer-1 K|V for every token, layer-1 query + attention row per sequence
    const double T = (double)G * n_seq * L;
    const double flops = T * (2.0 * c.obs_dim * 64 + 24.0 * 64 * 64 + 4.0 * L * 64 + 2.0 * 64 * 128) +
                         (double)G * n_seq * (2.0 * 64 * 64 + 4.0 * L * 64);
    prof_begin(PROF_ACT_FUSED, st);
    act_fused_kernel<<<dim3(gx, G, 1), AF_THREADS, AF_SMEM, st>>>(t);
    prof_end(PROF_ACT_FUSED, st, flops);
    DTQN_LAUNCH_CHECK();
    return 0;
}

// debug: device buffer of >= 256 int64 receiving the phase timeline of CTA 0 (NULL: off)
extern "C" int dtqn_set_act_fused_timeline(void* buf) {
    long long* p = (long long*)buf;
    return (int)cudaMemcpyToSymbol(g_af_dbg, &p, sizeof(p));
}

int act_fused_tc_error() {
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
    return v;
}
