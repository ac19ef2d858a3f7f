#include "prof.cuh"
#include "../../include/dtqn_b200.h"
#include <vector>

bool g_prof_on = false;

namespace {
struct Rec { cudaEvent_t a, b; int tag; double work; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t g_open[PROF_NTAGS];

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

void prof_begin_impl(int tag, cudaStream_t st) {
    cudaEvent_t e = get_event();
    cudaEventRecord(e, st);
    g_open[tag] = e;
}

void prof_end_impl(int tag, cudaStream_t st, double work) {
    cudaEvent_t e = get_event();
    cudaEventRecord(e, st);
    g_recs.push_back(Rec{g_open[tag], e, tag, work});
}

extern "C" int dtqn_profile_enable(int32_t on) {
    for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
    g_prof_on = on != 0;
    return 0;
}

extern "C" int dtqn_profile_read(int32_t tag, double* total_ms, int64_t* launches, double* total_work) {
    if (tag < 0 || tag >= PROF_NTAGS || !total_ms || !launches || !total_work) return DTQN_E_ARG;
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return (int)ce;
    double ms = 0.0, work = 0.0; long long n = 0;
    for (auto& r : g_recs) {
        if (r.tag != tag) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; work += r.work; ++n; }
    }
    *total_ms = ms; *launches = n; *total_work = work;
    return 0;
}

int g_pdl = 0;
extern "C" int dtqn_set_pdl(int32_t on) { g_pdl = on ? 1 : 0; return 0; }
