// Optional per-kernel timing used by bench.py's roofline leg: CUDA events recorded on the launching stream around
// every launch of a tagged kernel, enabled only for a dedicated profiling pass (never while the headline is timed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

enum ProfTag {
    PROF_LINEAR = 0,     // forward Linear GEMMs (x W^T + b, fused epilogues)
    PROF_ATTN_FWD = 1,
    PROF_ENV_STEP = 2,   // env_step_kernel (+ fused replay / context append)
    PROF_ENV_ROLL = 3,
    PROF_GATHER = 4,     // replay history-window gather
    PROF_DGRAD = 5,
    PROF_WGRAD = 6,
    PROF_ATTN_BWD = 7,
    PROF_LN_BWD = 8,
    PROF_EMBED = 9,
    PROF_HEAD = 10,
    PROF_TD = 11,
    PROF_ADAM = 12,
    PROF_OTHER = 13,
    PROF_LINEAR_TC = 14, // tcgen05 Linear launches; work = algorithmic BYTES (fp32 activations in + out, residual, weights)
    PROF_SEQ_FWD = 15,   // sequence-resident fused forward (all layers + head); work = FLOPs
    PROF_ACT_FUSED = 16, // fused acting forward (embed + layer 0 + final-layer K|V + last-row attention); work = FLOPs
    PROF_NTAGS = 17
};

extern bool g_prof_on;
void prof_begin_impl(int tag, cudaStream_t st);
void prof_end_impl(int tag, cudaStream_t st, double work);

static inline void prof_begin(int tag, cudaStream_t st) { if (g_prof_on) prof_begin_impl(tag, st); }
static inline void prof_end(int tag, cudaStream_t st, double work) { if (g_prof_on) prof_end_impl(tag, st, work); }
