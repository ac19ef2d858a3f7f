// numpy Generator(PCG64) draws on the device, bit-exact with numpy 2.x (the oracle of record).
// Restates numpy's published PCG64 (128-bit LCG, XSL-RR output), the buffered 32-bit half-word, Lemire's
// bounded integers (32-bit path), next_double and the masked-rejection interval used by shuffle
// (SURVEY.md Appendix B).  Reference call sites: envs/car_flag.py:147,158; envs/memory_cards.py:73,77,110-113;
// utils/context.py:50; dtqn/agents/dtqn.py:78-79; run.py:394.
#pragma once
#include <stdint.h>

struct Pcg64 {
    uint64_t s_hi, s_lo, i_hi, i_lo;
    uint32_t has32, buf32;

    __device__ __forceinline__ void load(const uint64_t* __restrict__ st, const uint32_t* __restrict__ bf, int n, int i) {
        s_hi = st[i]; s_lo = st[n + i]; i_hi = st[2 * n + i]; i_lo = st[3 * n + i];
        has32 = bf[i]; buf32 = bf[n + i];
    }
    __device__ __forceinline__ void store(uint64_t* __restrict__ st, uint32_t* __restrict__ bf, int n, int i) const {
        st[i] = s_hi; st[n + i] = s_lo;   // inc never changes
        bf[i] = has32; bf[n + i] = buf32;
    }
    __device__ __forceinline__ uint64_t next64() {
        const uint64_t M_HI = 0x2360ED051FC65DA4ull, M_LO = 0x4385DF649FCCF645ull;
        // state = state * MULT + inc  (mod 2^128)
        uint64_t lo = s_lo * M_LO;
        uint64_t hi = __umul64hi(s_lo, M_LO) + s_hi * M_LO + s_lo * M_HI;
        uint64_t nlo = lo + i_lo;
        uint64_t nhi = hi + i_hi + (nlo < lo ? 1ull : 0ull);
        s_lo = nlo; s_hi = nhi;
        uint64_t x = s_hi ^ s_lo;
        unsigned r = (unsigned)(s_hi >> 58);
        return (x >> r) | (x << ((64u - r) & 63u));
    }
    __device__ __forceinline__ uint32_t next32() {
        if (has32) { has32 = 0; return buf32; }
        uint64_t v = next64();
        has32 = 1; buf32 = (uint32_t)(v >> 32);
        return (uint32_t)v;
    }
    __device__ __forceinline__ double next_double() {
        return __dmul_rn((double)(next64() >> 11), 1.0 / 9007199254740992.0);
    }
    // Generator.integers(0, n) for n <= 2^32: Lemire with the buffered 32-bit source; n == 1 draws nothing.
    __device__ __forceinline__ uint32_t bounded(uint32_t n) {
        uint32_t rng = n - 1u;
        if (rng == 0u) return 0u;
        uint64_t m = (uint64_t)next32() * (uint64_t)n;
        uint32_t left = (uint32_t)m;
        if (left < n) {
            uint32_t thresh = (0xFFFFFFFFu - rng) % n;
            while (left < thresh) {
                m = (uint64_t)next32() * (uint64_t)n;
                left = (uint32_t)m;
            }
        }
        return (uint32_t)(m >> 32);
    }
    // random_interval(max) used by Generator.shuffle: masked rejection on 32-bit draws.
    __device__ __forceinline__ uint32_t interval(uint32_t mx) {
        if (mx == 0u) return 0u;
        uint32_t mask = mx;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next32() & mask; } while (v > mx);
        return v;
    }
};
