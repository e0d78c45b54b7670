// ss_device.cuh -- device-side data layout and arithmetic shared by every kernel.
//
// Data layout in HBM
//   lattice-major cell store, linear index c = x*N + y (x = the reference's first grid index):
//     sugar[c], cap[c], poll[c]  : f64   (Environment.grid slots 0, 2, 5; environment.py:24-34)
//     occw[c]                    : u32   occupancy word, 0 = free, else
//                                          bit 31      phase  (pending in step t  <=>  phase == (t & 1))
//                                          bit 30      Q1 flag (at-risk predecessor, lattice path)
//                                          bit 29      at-risk flag (may starve this step)
//                                          bits 28..24 vision (1..31, so an occupied word is non-zero)
//                                          bits 23..0  slot (= Agent.id - 1)
//     isugar[c]                  : u8    int(sugar[c]) mirror (what agents compare and harvest,
//                                        environment.py:98-100); lattice path only
//   agent store: one 32-byte record per slot (exactly one DRAM sector, so the random
//   gather/scatter by slot that movement implies costs one sector per agent).
//
// Arithmetic: all quantities the reference compares (sugar, pollution, utilities) are IEEE
// doubles computed with the reference's operation order; the library is compiled with
// -fmad=false so nothing is contracted into FMAs.
#pragma once
#include <stdint.h>
#include "../../include/sugarscape_b200.h"

#define SS_MAX_VISION 31
#define SS_MAXC (4 * SS_MAX_VISION + 8)   // >= 2*(2v+1) candidate positions of one cross

struct __align__(32) AgentRec {
    double   sugar;       // Agent.sugar
    uint32_t pos;         // x*N + y
    uint32_t attr;        // bits 0..7 vision, 8..23 metabolism, bit 31 alive
    uint32_t status;      // ((step + 1) << 2) | code : outcome of the agent's turn in `step`
    uint32_t pred;        // Q1: slot+1 of the at-risk predecessor in the execution order ...
    uint32_t pred_step;   // ... valid iff pred_step == step + 1
    uint32_t turn;        // ((step + 1) << 8) | harvest : deferred turn of a safe agent (ss_settle_kernel)
};
static_assert(sizeof(AgentRec) == 32, "AgentRec must be one 32-byte sector");

enum { ST_ACTED = 1, ST_DIED = 2, ST_SKIPPED = 3 };

#define ATTR_ALIVE 0x80000000u
#define ATTR_RISK  0x40000000u         // dependency rounds: the agent may starve in its next turn (Q1 edge to its successor)
__host__ __device__ inline uint32_t attr_pack(int vision, int metab, bool alive) {
    return (uint32_t)vision | ((uint32_t)metab << 8) | (alive ? ATTR_ALIVE : 0u);
}
__host__ __device__ inline int attr_vision(uint32_t a) { return (int)(a & 0xffu); }
__host__ __device__ inline int attr_metab(uint32_t a) { return (int)((a >> 8) & 0xffffu); }
__host__ __device__ inline bool attr_alive(uint32_t a) { return (a & ATTR_ALIVE) != 0; }

#define OCC_Q1   0x40000000u         // the agent has an at-risk predecessor in this step's order (Q1)
#define OCC_RISK 0x20000000u         // the agent may starve this step (its turn is settled synchronously)
#define OCC_FLAGS (OCC_Q1 | OCC_RISK)
#define OCC_SLOT_BITS 24
#define OCC_SLOT_MASK ((1u << OCC_SLOT_BITS) - 1u)
#define SS_MAX_SLOTS (1u << OCC_SLOT_BITS)
// vision >= 1, so an occupied cell's word is never 0
__host__ __device__ inline uint32_t occ_pack(uint32_t slot, int vision, uint32_t phase) {
    return (phase << 31) | ((uint32_t)vision << OCC_SLOT_BITS) | slot;
}
__host__ __device__ inline uint32_t occ_slot(uint32_t w) { return w & OCC_SLOT_MASK; }
__host__ __device__ inline int occ_vision(uint32_t w) { return (int)((w >> OCC_SLOT_BITS) & 31u); }
__host__ __device__ inline uint32_t occ_phase(uint32_t w) { return w >> 31; }

// ------------------------------------------------------------------ keyed word source ----
// Same constants as oracle/keyed_rng.py and oracle/ss_oracle.c.
#define SS_GOLDEN  0x9E3779B97F4A7C15ull
#define SS_C_AGENT 0xD1342543DE82EF95ull
#define SS_C_ROUND 0xA24BAED4963EE407ull

__host__ __device__ inline uint64_t ss_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ inline uint32_t ss_fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ inline uint64_t ss_step_key(uint64_t seed, int64_t t) {
    return ss_mix64(seed + SS_GOLDEN * (uint64_t)(t + 1));
}
__host__ __device__ inline uint64_t ss_agent_key(uint64_t skey, uint32_t id) {
    return ss_mix64(skey ^ ((uint64_t)id * SS_C_AGENT));
}
__host__ __device__ inline uint32_t ss_keyed_word(uint64_t akey, uint32_t j) {
    return (uint32_t)(ss_mix64(akey + (uint64_t)j * SS_GOLDEN) >> 32);
}
__host__ __device__ inline uint32_t ss_round_key(uint64_t skey, int r) {
    return (uint32_t)(ss_mix64(skey + (uint64_t)(r + 1) * SS_C_ROUND) >> 32);
}
__host__ __device__ inline int ss_order_half_bits(uint32_t n_slots) {
    int kb = 0;
    uint32_t v = (n_slots > 1 ? n_slots : 1) - 1;
    while (v) { kb++; v >>= 1; }
    if (kb < 1) kb = 1;
    int half = (kb + 1) / 2;
    return half < 1 ? 1 : half;
}
struct OrderKey {           // keyed bijection on [0, 2^(2*half)) : execution priority of a slot
    uint32_t rk[4];
    int half;
};
__host__ __device__ inline OrderKey ss_order_key(uint64_t skey, int half) {
    OrderKey k;
    for (int r = 0; r < 4; r++) k.rk[r] = ss_round_key(skey, r);
    k.half = half;
    return k;
}
__host__ __device__ inline uint32_t ss_priority(const OrderKey& k, uint32_t slot) {
    const uint32_t mask = (1u << k.half) - 1u;
    uint32_t L = (slot >> k.half) & mask, R = slot & mask;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint32_t t = L ^ (ss_fmix32(R ^ k.rk[r]) & mask);
        L = R; R = t;
    }
    return (L << k.half) | R;
}
__host__ __device__ inline uint32_t ss_priority_inverse(const OrderKey& k, uint32_t prio) {
    const uint32_t mask = (1u << k.half) - 1u;
    uint32_t L = (prio >> k.half) & mask, R = prio & mask;
#pragma unroll
    for (int r = 3; r >= 0; r--) {
        uint32_t t = R ^ (ss_fmix32(L ^ k.rk[r]) & mask);
        R = L; L = t;
    }
    return (L << k.half) | R;
}

// ------------------------------------------------------------------------- arithmetic ----
// agent.py:224-229 setSugar, limitedLifespan off: max(amount, .01)
__device__ inline double ss_set_sugar(double amount) { return amount > 0.01 ? amount : 0.01; }
__device__ inline double ss_max0(double v) { return v > 0.0 ? v : 0.0; }
// CPython float_floor_div (floatobject.c): `a // b`
__device__ inline double ss_py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) div -= 1.0;
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
    } else {
        fl = copysign(0.0, vx / wx);
    }
    return fl;
}
__device__ inline int ss_wrap(int v, int n) { int r = v % n; return r < 0 ? r + n : r; }
// same for v in [-n, 2n): no integer division
__device__ inline int ss_wrap1(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }
