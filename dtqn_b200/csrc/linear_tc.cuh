// Host-side interface of the tcgen05 Linear (linear_tc.cu).
#pragma once
#include "net.cuh"

struct TcPackEntry { long long w_off, pk_off; int N, K, n_tile; };
// n_tile == 0: plain fp32 TRANSPOSE of the weight ([K][N], k-major) for the sequence-resident kernel's cp.async slabs
struct TcPackTable { TcPackEntry e[6 * DTQN_MAX_LAYERS + 1 + 5 + 4 * DTQN_MAX_LAYERS + 1]; int n; long long total_bytes, act_img_off, wt_off; };

// entry order: per layer in_proj, out_proj, ffn.0, ffn.2; then the head's ffn.0; then per layer the K|V rows and the Q rows
// of in_proj (index 4*n_layers + 1 + 2*i, + 1)
enum { TC_W_IN = 0, TC_W_OUT = 1, TC_W_F1 = 2, TC_W_F2 = 3 };

// token embedding recomputed inside the tcgen05 kernel (continuous observations, d_model = 64)
struct TcEmbed {
    int mode;                         // 0 off; bit 0: A operand; bit 1: LayerNorm residual
    int L, O;
    float obs_mask;
    long long w_off, b_off, pos_off;  // offsets into the flat parameters
    dtqn_obs_src src[DTQN_MAX_GROUPS];
};

bool tc_pipelined_enabled();
bool tc_ffn_fused_enabled();
int launch_ffn_tc(const float* X, float* Y, const GroupPtrs& P, int G, const uint8_t* const* packed, long long pk_w1,
                  long long pk_w2, long long b1_off, long long b2_off, long long gamma_off, long long beta_off, int Tg,
                  cudaStream_t st);
int tc_ntile(int N);
int tc_pack_table(const dtqn_net_cfg& c, const NetLayout& lay, TcPackTable& tab);
int launch_linear_tc(const LinArgs& a, int epi, int G, const uint8_t* const* packed, long long pk_off, cudaStream_t st,
                     const TcEmbed* emb = nullptr);

// fused acting forward (act_fused.cu): context ring -> (x2[last], final-layer attention row) per sequence, one launch
bool act_fused_supported(const dtqn_net_cfg& c, int L);
int launch_act_fused(const dtqn_net_cfg& c, const NetLayout& lay, const GroupPtrs& P, const GroupSrc& S, int G,
                     const uint8_t* const* packed, long long img_off, int n_seq, int L, float* xl, float* ol, cudaStream_t st);
int act_fused_tc_error();
