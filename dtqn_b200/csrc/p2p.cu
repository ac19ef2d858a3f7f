// Gradient exchange fused with the optimiser over NVLink peer memory (the ONE collective of the
// path sits between loss.backward() and clip_grad_norm_, dtqn/agents/dtqn.py:256-261).
//
// Every rank owns one cudaMalloc'ed, IPC-exported exchange buffer
//     [ header 8192 B : epoch, ticket, start flags [P2P_BLOCKS][8], end flags [P2P_BLOCKS][8] ][ gradients: n floats ]
// into which its backward kernels write the local gradient directly.  One kernel then, per CTA,
//     1. raises a flag in every peer's header ("my gradient is complete") and waits for theirs       (start barrier)
//     2. reads its slice of ALL ranks' gradients through NVLink (rank order 0..W-1, so every rank produces bit-identical
//        sums), writes the sum to a local buffer and accumulates the squared norm of the scaled sum     (allreduce + norm)
//     3. raises / waits the end flags so nobody's next backward can overwrite a buffer still being read (end barrier)
// and the existing clip + Adam kernel consumes the partial norms.  No NCCL call, no extra launch compared with the
// single-GPU update, no host involvement: the whole multi-GPU iteration stays inside one CUDA graph.
// Flags carry a per-launch epoch kept in device memory (graph replays re-use frozen kernel arguments); waits are bounded
// and raise an error flag instead of hanging the GPU if a peer never arrives.
#include "common.cuh"
#include "prof.cuh"

#define P2P_BLOCKS 64
#define P2P_THREADS 512
#define P2P_HEADER_BYTES 8192
#define P2P_SPIN_LIMIT 20000000ll          // x ~100 ns  =  ~2 s

namespace {

struct P2PHeader {
    uint32_t epoch;                        // launches completed by this rank
    uint32_t ticket;                       // CTAs of the current launch that finished
    uint32_t error;                        // set when a bounded wait expired
    uint32_t pad[61];
    uint32_t start[P2P_BLOCKS][DTQN_P2P_MAX_RANKS];
    uint32_t end[P2P_BLOCKS][DTQN_P2P_MAX_RANKS];
};
static_assert(sizeof(P2PHeader) <= P2P_HEADER_BYTES, "header overflows its page");

struct P2PPeers { char* base[DTQN_P2P_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// thread r < world signals peer r and waits for peer r's signal (flags live in the WAITER's memory: polling is local)
__device__ __forceinline__ void p2p_barrier(const P2PPeers& peers, int rank, int world, uint32_t e, bool end_phase) {
    const int r = threadIdx.x;
    if (r < world) {
        P2PHeader* theirs = reinterpret_cast<P2PHeader*>(peers.base[r]);
        P2PHeader* mine = reinterpret_cast<P2PHeader*>(peers.base[rank]);
        uint32_t* dst = end_phase ? &theirs->end[blockIdx.x][rank] : &theirs->start[blockIdx.x][rank];
        const uint32_t* src = end_phase ? &mine->end[blockIdx.x][r] : &mine->start[blockIdx.x][r];
        st_release_sys(dst, e);
        long long spins = 0;
        while (ld_acquire_sys(src) != e) {
            if (++spins > P2P_SPIN_LIMIT) { mine->error = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

template <int W>                            // W > 0: world size known at compile time (peer loads issued back to back)
__global__ void __launch_bounds__(P2P_THREADS)
p2p_reduce_sqnorm_kernel(P2PPeers peers, int rank, int world, long long n, float scale, float* __restrict__ out,
                         float* __restrict__ partial, long long* step) {
    P2PHeader* mine = reinterpret_cast<P2PHeader*>(peers.base[rank]);
    const uint32_t e = ld_volatile_u32(&mine->epoch) + 1;
    __threadfence_system();
    p2p_barrier(peers, rank, world, e, false);

    const long long n4 = n >> 2;                                  // n is a multiple of 4 (flat layout is 16-byte padded)
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * P2P_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * P2P_THREADS) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (W > 0) {
            float4 v[W > 0 ? W : 1];
#pragma unroll
            for (int r = 0; r < W; ++r) v[r] = __ldcv(reinterpret_cast<const float4*>(peers.base[r] + P2P_HEADER_BYTES) + i);
#pragma unroll
            for (int r = 0; r < W; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        } else {
            for (int r = 0; r < world; ++r) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(peers.base[r] + P2P_HEADER_BYTES) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(out)[i] = acc;
        const float a = acc.x * scale, b = acc.y * scale, c = acc.z * scale, d = acc.w * scale;
        s = fmaf(a, a, s); s = fmaf(b, b, s); s = fmaf(c, c, s); s = fmaf(d, d, s);
    }
    __shared__ float red[P2P_THREADS / 32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < P2P_THREADS / 32; ++w) t += red[w];
        // a bounded wait expired (a peer never arrived): the sum may be incomplete -> poison the norm so clip + Adam skips
        // the update and raises the non-finite flag instead of applying a partial gradient
        partial[blockIdx.x] = ld_volatile_u32(&mine->error) ? __int_as_float(0x7fc00000) : t;
    }
    p2p_barrier(peers, rank, world, e, true);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&mine->ticket, 1u) == gridDim.x - 1) {       // last CTA of this launch: every CTA has read `epoch`
            mine->ticket = 0;
            mine->epoch = e;
            *step += 1;                                            // optimiser step index, as sqnorm_kernel does
            __threadfence();
        }
    }
}

}  // namespace

int launch_clip_adam_only(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float grad_scale,
                          float max_norm, float lr, float beta1, float beta2, float eps, long long* step,
                          const float* partial, int n_partial, float* stats, int* flags, float* ring, int ring_len,
                          cudaStream_t st);

extern "C" int dtqn_p2p_alloc(int64_t n_floats, void** base_out, float** grads_out, uint8_t* handle_out) {
    if (n_floats <= 0 || (n_floats & 3) || !base_out || !grads_out || !handle_out) return DTQN_E_ARG;
    void* base = nullptr;
    const size_t bytes = P2P_HEADER_BYTES + sizeof(float) * (size_t)n_floats;
    cudaError_t e = cudaMalloc(&base, bytes);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaMemset(base, 0, bytes)) != cudaSuccess) { cudaFree(base); return (int)e; }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, base)) != cudaSuccess) { cudaFree(base); return (int)e; }
    static_assert(sizeof(cudaIpcMemHandle_t) == DTQN_P2P_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    *base_out = base;
    *grads_out = reinterpret_cast<float*>(static_cast<char*>(base) + P2P_HEADER_BYTES);
    return 0;
}

extern "C" int dtqn_p2p_open(const uint8_t* handle, void** peer_base_out) {
    if (!handle || !peer_base_out) return DTQN_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    *peer_base_out = p;
    return 0;
}

extern "C" int dtqn_p2p_close(void* peer_base) {
    if (!peer_base) return DTQN_E_ARG;
    return (int)cudaIpcCloseMemHandle(peer_base);
}

extern "C" int dtqn_p2p_free(void* base) {
    if (!base) return DTQN_E_ARG;
    return (int)cudaFree(base);
}

extern "C" int dtqn_p2p_error(const void* base) {
    if (!base) return DTQN_E_ARG;
    uint32_t v = 0;
    cudaError_t e = cudaMemcpy(&v, static_cast<const char*>(base) + offsetof(P2PHeader, error), sizeof(v), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? (int)v : -(int)e - 1000;
}

extern "C" int dtqn_allreduce_clip_adam(float* params, void* const* bases, int32_t rank, int32_t world, int64_t n,
                                        float* grads_reduced, float* exp_avg, float* exp_avg_sq, float max_norm, float lr,
                                        float beta1, float beta2, float eps, int64_t* step_counter, float* scratch,
                                        float* stats_out, int32_t* flags_out, float* stats_ring, int32_t ring_len,
                                        void* stream) {
    if (!params || !bases || world < 1 || world > DTQN_P2P_MAX_RANKS || rank < 0 || rank >= world || n <= 0 || (n & 3) ||
        !grads_reduced || !exp_avg || !exp_avg_sq || !step_counter || !scratch || !stats_out || !flags_out) return DTQN_E_ARG;
    P2PPeers peers{};
    for (int r = 0; r < world; ++r) {
        if (!bases[r]) return DTQN_E_ARG;
        peers.base[r] = static_cast<char*>(bases[r]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const float scale = 1.0f / (float)world;
    prof_begin(PROF_ADAM, st);
#define P2P_LAUNCH(W) p2p_reduce_sqnorm_kernel<W><<<P2P_BLOCKS, P2P_THREADS, 0, st>>>(peers, rank, world, n, scale, \
                                                                    grads_reduced, scratch, (long long*)step_counter)
    switch (world) {
        case 2: P2P_LAUNCH(2); break;
        case 4: P2P_LAUNCH(4); break;
        case 8: P2P_LAUNCH(8); break;
        default: P2P_LAUNCH(0); break;
    }
#undef P2P_LAUNCH
    DTQN_LAUNCH_CHECK();
    int rc = launch_clip_adam_only(params, grads_reduced, exp_avg, exp_avg_sq, n, scale, max_norm, lr, beta1, beta2, eps,
                                   (long long*)step_counter, scratch, P2P_BLOCKS, stats_out, flags_out, stats_ring,
                                   stats_ring ? ring_len : 1, st);
    prof_end(PROF_ADAM, st, (32.0 + 4.0 * world) * (double)n);
    return rc;
}
