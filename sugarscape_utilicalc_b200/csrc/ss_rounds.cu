// ss_rounds.cu -- the agent phase as dependency rounds over the WHOLE lattice (keyed RNG, rule M).
//
// View.updateGame's agent loop (sugarscape.py:517-601) is sequential in a random order.  Two agents conflict iff
// their vision crosses (agent.py:361-371) share a cell -- every read and write of a turn lies inside the agent's own
// cross -- so the sequential result is reproduced exactly when every agent acts after all conflicting agents of
// lower execution priority.  That is a DAG over the agents' start-of-step positions; this file evaluates it level by
// level without windows, lists, sorts or host round trips:
//
//   ss_dr_build_kernel   one CTA per 64 x 128-cell tile stages the occupancy words of the tile and a halo of 2*vmax
//                        cells in shared memory (TMA 2-D tile load, cp.async.bulk.tensor, for tiles that do not touch
//                        the torus edge), turns them into (priority, vision) words + an occupancy bitmap, and one
//                        thread per agent scans its conflict zone with the exact cross-intersection predicate:
//                        predecessors are COUNTED (dep[cell]), successors are recorded as a bitmap over the zone
//                        (mask[slot], 32 bytes for vision <= 7).  The list-order predecessor (the agent before it in
//                        the keyed order, found by walking the inverse priority map over the alive bitmap) adds one
//                        more edge when it may starve this step (Q1: sugarscape.py:520 + 502-503, the agent after a
//                        dying agent loses its turn).  Agents without predecessors enter the ready queue.
//   ss_dr_rounds_kernel  one persistent cooperative kernel per step.  Round k: one thread per ready agent -- cross
//                        scan on the one-byte cell words (occupied | int(sugar)), choice (agent.py:794-823 incl. the
//                        replay of the agent's Fisher-Yates draws for ties), harvest, pollution, metabolism, death
//                        (sugarscape.py:566-601), statistics -- then one atomic decrement per successor; whoever
//                        brings a counter to zero appends that agent to the queue.  A grid barrier separates the
//                        rounds; the kernel ends when a round appended nothing.
//
// Everything an agent's turn needs is settled in the turn (no deferred pass), so the only per-step kernels around
// these two are the statistics and the growback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstring>

#include "ss_world.cuh"

#define FULL 0xffffffffu
#define DR_TX 64                        // tile rows (x)
#define DR_TY 128                       // tile columns (y, contiguous in memory)
#define DR_BUILD_THREADS 512
#define DR_THREADS 512                  // rounds kernel
#define DR_PUSH_CAP 4096                // staged queue entries per CTA between flushes
#define DR_HIST_BINS 2048               // low wealth bins kept in shared memory per CTA
#define DR_SKIP 0x80000000u             // queue entry flag: the agent loses its turn (Q1)

// ------------------------------------------------------------------------------ helpers ----
__device__ static inline bool dr_crosses_meet(int dx, int dy, int a, int b) {
    const int adx = abs(dx), ady = abs(dy);
    return (adx <= a && ady <= b) || (ady <= a && adx <= b) || (dy == 0 && adx <= a + b) || (dx == 0 && ady <= a + b);
}
// bit of zone offset (dx, dy) in the successor bitmap: the square |dx|,|dy| <= V row by row, then the four arm
// extensions (V < |d| <= 2V on an axis)
__device__ static inline int dr_bit_index(int dx, int dy, int V) {
    const int S = 2 * V + 1;
    if (abs(dx) <= V && abs(dy) <= V) return (dx + V) * S + (dy + V);
    const int base = S * S;
    if (dy == 0) return dx < 0 ? base + (-dx - V - 1) : base + V + (dx - V - 1);
    return dy < 0 ? base + 2 * V + (-dy - V - 1) : base + 3 * V + (dy - V - 1);
}
__device__ static inline void dr_bit_offset(int idx, int V, uint32_t magicS, int& dx, int& dy) {
    const int S = 2 * V + 1, S2 = S * S;
    if (idx < S2) {
        const int r = (int)(((uint32_t)idx * magicS) >> 16);       // idx / S (exact for idx < 2^16 / S)
        dx = r - V; dy = idx - r * S - V;
        return;
    }
    const int t = idx - S2;
    int arm = 0, d = t;
    while (d >= V) { d -= V; arm++; }
    d += V + 1;
    dx = arm == 0 ? -d : arm == 1 ? d : 0;
    dy = arm == 2 ? -d : arm == 3 ? d : 0;
}
__host__ __device__ static inline int dr_mask_words(int V) {
    const int bits = (2 * V + 1) * (2 * V + 1) + 4 * V;
    return bits <= 256 ? 4 : 16;
}

// strip addressing: which strip holds lattice row gx (already wrapped to [0, n)) and where.  which: 0 = this strip,
// 1 = the ring neighbour below (lo), 2 = the one above (hi); the conflict zone never reaches further (rows >= 2*vmax).
// STRIPS = false (one strip = the whole torus): everything is local and the selects fold away.
struct DrLoc { int which; uint32_t idx; };
template <bool STRIPS>
__device__ static inline DrLoc dr_loc(const DevWorld& w, int gx, int gy) {
    DrLoc L;
    if (!STRIPS) { L.which = 0; L.idx = (uint32_t)gx * (uint32_t)w.n + (uint32_t)gy; return L; }
    const int r = gx - w.dr.self.x0;
    if ((unsigned)r < (unsigned)w.dr.self.rows) { L.which = 0; L.idx = (uint32_t)r * (uint32_t)w.n + (uint32_t)gy; return L; }
    int below = w.dr.self.x0 - gx;               // 1.. if gx is just below the strip (mod n)
    if (below < 0) below += w.n;
    if (below <= 4 * SS_MAX_VISION && below <= w.dr.lo.rows) {
        L.which = 1; L.idx = (uint32_t)(w.dr.lo.rows - below) * (uint32_t)w.n + (uint32_t)gy; return L;
    }
    int above = gx - (w.dr.self.x0 + w.dr.self.rows);
    if (above < 0) above += w.n;
    L.which = 2; L.idx = (uint32_t)above * (uint32_t)w.n + (uint32_t)gy;
    return L;
}
#define DR_F(field, L) ((!STRIPS || (L).which == 0) ? w.dr.self.field : ((L).which == 1 ? w.dr.lo.field : w.dr.hi.field))

template <int MW> struct MaskAcc;
template <> struct MaskAcc<4> {
    unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    __device__ inline void set(int idx) {
        const unsigned long long b = 1ull << (idx & 63);
        const int wi = idx >> 6;
        if (wi == 0) m0 |= b; else if (wi == 1) m1 |= b; else if (wi == 2) m2 |= b; else m3 |= b;
    }
    __device__ inline void store(unsigned long long* p) const {
        reinterpret_cast<ulonglong2*>(p)[0] = make_ulonglong2(m0, m1);
        reinterpret_cast<ulonglong2*>(p)[1] = make_ulonglong2(m2, m3);
    }
};
template <> struct MaskAcc<16> {
    unsigned long long m[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    __device__ inline void set(int idx) { m[idx >> 6] |= 1ull << (idx & 63); }
    __device__ inline void store(unsigned long long* p) const { for (int i = 0; i < 16; i++) p[i] = m[i]; }
};

// ------------------------------------------------------------------- mirrors and flags ----
// mw from the uploaded state; flag bit 0: a value does not fit 7 bits (the path is then refused)
__global__ void ss_dr_mirror_kernel(DevWorld w, size_t cells, int* flag) {
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const double v = w.sugar[c];
        int iv = 0;
        if (!(v >= 0.0 && v < 127.0) || !(w.cap[c] < 127.0)) *flag = 1; else iv = (int)v;
        w.dr.self.mw[c] = (uint8_t)(iv | (w.occw[c] ? 0x80 : 0));
    }
}
// alive / at-risk bitmaps and the OCC_RISK flag of the occupancy words from the agent store (after an upload).
// AT RISK = the agent starves this step unless it harvests something: sugar - metabolism <= 0.1 before its turn
// (sugarscape.py:569, 306-309; agent.py:224-229).  Everybody else survives whatever happens, so only at-risk agents
// create a Q1 edge to their successor in the list order.
__device__ static inline bool dr_at_risk(const DevWorld& w, double sugar, int metab) {
    return w.cfg.rule_can_starve && ss_set_sugar(ss_max0(sugar - (double)metab)) <= 0.1;
}
__global__ void ss_dr_flags_kernel(DevWorld w) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false, risk = false;
    if (slot < w.n_slots) {
        const AgentRec r = w.agents[slot];
        alive = attr_alive(r.attr);
        if (alive) {
            risk = dr_at_risk(w, r.sugar, attr_metab(r.attr));
            uint32_t ow = w.occw[r.pos];
            ow = risk ? (ow | OCC_RISK) : (ow & ~OCC_RISK);
            w.occw[r.pos] = ow & ~OCC_Q1;
        }
    }
    const unsigned ba = __ballot_sync(FULL, alive), br = __ballot_sync(FULL, risk);
    if ((threadIdx.x & 31) == 0 && (slot >> 5) < ((w.n_slots + 31) >> 5)) { w.dr.alive[slot >> 5] = ba; w.dr.atrisk[slot >> 5] = br; }
}

// -------------------------------------------------------------------------------- build ----
__device__ static inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct DrBuildArgs {
    OrderKey ok;
    int64_t iteration;
    int tiles_y;
    int use_tma;
};

template <int MW, bool STRIPS>
__global__ void __launch_bounds__(DR_BUILD_THREADS)
ss_dr_build_kernel(DevWorld w, DrBuildArgs ba, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ int s_n, s_ready, s_fill;
    __shared__ unsigned s_base;
    __shared__ __align__(8) unsigned long long s_bar;
    const int n = w.n, V = w.vmax, H = 2 * V, SX = DR_TX + 2 * H, SY = DR_TY + 2 * H, BW = (SY + 31) >> 5;
    uint32_t* tile = sm;                                   // [SX][SY] occupancy words -> (priority << 5 | vision)
    uint32_t* bm = tile + SX * SY;                         // [SX][BW] occupancy bitmap
    uint16_t* list = (uint16_t*)(bm + SX * BW);            // interior agents: r << 8 | q  (bit 15: ready)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = DR_BUILD_THREADS / 32;
    const DrPeer& S = w.dr.self;
    const int tx = blockIdx.x / ba.tiles_y, ty = blockIdx.x - tx * ba.tiles_y;
    const int x0 = S.x0 + tx * DR_TX, y0 = ty * DR_TY;     // lattice coordinates of the tile's first interior cell
    const int ex = min(DR_TX, S.x0 + S.rows - x0), ey = min(DR_TY, n - y0);      // interior extent
    if (tid == 0) { s_n = 0; s_ready = 0; s_fill = 0; }
    for (int i = tid; i < SX * BW; i += DR_BUILD_THREADS) bm[i] = 0u;
    // the tile and its halo lie inside this strip and inside [0, n) in y: one TMA box; otherwise plain loads with wraps
    const bool interior = ba.use_tma && x0 - H >= S.x0 && x0 + DR_TX + H <= S.x0 + S.rows && y0 - H >= 0 && y0 + DR_TY + H <= n;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(SX * SY * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(y0 - H), "r"(x0 - H - S.x0), "r"(smem_u32(&s_bar)) : "memory");
        }
        // everybody waits for the box (phase 0)
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&s_bar)) : "memory");
            }
        }
    } else {
        __syncthreads();
        for (int r = warp; r < SX; r += nwarps) {
            const int gx = ss_wrap(x0 - H + r, n);
            for (int q = lane; q < SY; q += 32) {
                const int gy = ss_wrap(y0 - H + q, n);
                const DrLoc L = dr_loc<STRIPS>(w, gx, gy);
                tile[r * SY + q] = __ldcg(&DR_F(occw, L)[L.idx]);
            }
        }
        __syncthreads();
    }
    // occupancy words -> (priority, vision), bitmap, list of the tile's own agents
    for (int r = warp; r < SX; r += nwarps) {
        const bool row_in = r >= H && r < H + ex;
        for (int q = lane; q < SY; q += 32) {
            const uint32_t ow = tile[r * SY + q];
            if (!ow) continue;
            tile[r * SY + q] = (ss_priority(ba.ok, occ_slot(ow)) << 5) | (uint32_t)occ_vision(ow);
            atomicOr(&bm[r * BW + (q >> 5)], 1u << (q & 31));
            if (row_in && q >= H && q < H + ey) list[atomicAdd(&s_n, 1)] = (uint16_t)((r << 8) | q);
        }
    }
    __syncthreads();
    const int nl = s_n;
    const uint32_t stamp = (uint32_t)(ba.iteration + 1);
    for (int k = tid; k < nl; k += DR_BUILD_THREADS) {
        const int lr = list[k] >> 8, lq = list[k] & 255;
        const uint32_t me = tile[lr * SY + lq];
        const uint32_t pa = me >> 5;
        const int a = (int)(me & 31u);
        const int gx = x0 + lr - H, gy = y0 + lq - H;
        const uint32_t c = (uint32_t)(gx - S.x0) * (uint32_t)n + (uint32_t)gy;
        const uint32_t slot = occ_slot(__ldcg(&S.occw[c]));
        MaskAcc<MW> mk;
        uint32_t indeg = 0;
        for (int dx = -H; dx <= H; dx++) {
            const int r = lr + dx, adx = abs(dx);
            if (adx > V) {                                             // beyond the square: only the cell on the axis
                const uint32_t cv = tile[r * SY + lq];
                if (cv && adx <= a + (int)(cv & 31u)) { if ((cv >> 5) < pa) indeg++; else mk.set(dr_bit_index(dx, 0, V)); }
                continue;
            }
            const int R = dx == 0 ? H : V, qlo = lq - R, qhi = lq + R;
            for (int wq = qlo >> 5; wq <= (qhi >> 5); wq++) {
                uint32_t bits = bm[r * BW + wq];
                const int base = wq << 5;
                if (base < qlo) bits &= FULL << (qlo - base);
                if (base + 31 > qhi) bits &= FULL >> (base + 31 - qhi);
                while (bits) {
                    const int q = base + __ffs(bits) - 1;
                    bits &= bits - 1u;
                    const int dy = q - lq;
                    if (dx == 0 && dy == 0) continue;
                    const uint32_t cv = tile[r * SY + q];
                    if (!dr_crosses_meet(dx, dy, a, (int)(cv & 31u))) continue;
                    if ((cv >> 5) < pa) indeg++; else mk.set(dr_bit_index(dx, dy, V));
                }
            }
        }
        // Q1: the agent before this one in the list order (the keyed order: ascending priority over the living
        // agents); if it may starve, its outcome decides whether this agent acts at all
        if (w.cfg.rule_can_starve) {
            for (uint32_t p = pa; p-- > 0u;) {
                const uint32_t s2 = ss_priority_inverse(ba.ok, p);
                if (s2 >= w.n_slots || !((__ldg(&w.dr.alive[s2 >> 5]) >> (s2 & 31)) & 1u)) continue;
                if ((__ldg(&w.dr.atrisk[s2 >> 5]) >> (s2 & 31)) & 1u) {
                    indeg++;
                    w.dr.q1wait[s2] = make_uint2(c, stamp);
                }
                break;
            }
        }
        reinterpret_cast<uint16_t*>(S.dep)[c] = (uint16_t)indeg;
        mk.store(&w.dr.mask[(size_t)slot * MW]);
        if (indeg == 0) { list[k] |= 0x8000u; atomicAdd(&s_ready, 1); }
    }
    __syncthreads();
    if (tid == 0 && s_ready) s_base = atomicAdd(&S.ctl[0], (unsigned)s_ready);
    __syncthreads();
    if (s_ready) {
        for (int k = tid; k < nl; k += DR_BUILD_THREADS) {
            if (!(list[k] & 0x8000u)) continue;
            const int lr = (list[k] >> 8) & 127, lq = list[k] & 255;
            const uint32_t c = (uint32_t)(x0 + lr - H - S.x0) * (uint32_t)n + (uint32_t)(y0 + lq - H);
            S.queue[s_base + (unsigned)atomicAdd(&s_fill, 1)] = c;
        }
    }
}

// ------------------------------------------------------------------------------- rounds ----
struct DrAccum { unsigned sum_m, sum_v, acted, deaths, small, inexact, fault; };
struct DrShared {
    unsigned hist[DR_HIST_BINS];
    uint32_t push[DR_PUSH_CAP];
    unsigned long long acc[8];
    int cnt;
    unsigned base;
};

__device__ static inline void dr_push(const DevWorld& w, DrShared& sh, uint32_t e) {
    const int pos = atomicAdd(&sh.cnt, 1);
    if (pos < DR_PUSH_CAP) sh.push[pos] = e;
    else w.dr.self.queue[atomicAdd(&w.dr.self.ctl[0], 1u)] = e;          // staging full: straight to the queue
}
// one predecessor of the agent that started the step on lattice cell (gx, gy) is resolved
template <bool STRIPS>
__device__ static inline void dr_release(const DevWorld& w, DrShared& sh, DrAccum& acc, int gx, int gy, bool lose_turn) {
    const DrLoc L = dr_loc<STRIPS>(w, gx, gy);
    uint32_t* wp = &DR_F(dep, L)[L.idx >> 1];
    const int shf = (int)(L.idx & 1u) << 4;
    const uint32_t old = lose_turn ? atomicAdd(wp, 0x7fffu << shf) : atomicSub(wp, 1u << shf);
    const uint32_t h = (old >> shf) & 0xffffu;
    if ((h & 0x7fffu) == 1u) {
        const uint32_t e = L.idx | ((lose_turn || (h >> 15)) ? DR_SKIP : 0u);
        if (!STRIPS || L.which == 0) dr_push(w, sh, e);
        else DR_F(queue, L)[atomicAdd(&DR_F(ctl, L)[0], 1u)] = e;        // a ring neighbour's agent
    } else if ((h & 0x7fffu) == 0u) acc.fault++;
}

struct DrStep {
    int64_t iteration, env_time;
    uint64_t skey;
    uint32_t next_phase, magicS;
    int single_round;               // host-driven rounds (several strips on one GPU): run one round and return
    unsigned seg_begin, seg_end;    // ... over this queue segment
    unsigned nblocks;
};

// agent.py:361-371 getFood + 794-823 move(): one THREAD scans the cross on the one-byte cell words.  Candidates are
// numbered k = offset + a on the x-arm (ascending raw x), then span + offset + a on the y-arm -- the order getFood
// appends them in; best = (key desc, raw Manhattan distance asc, position in the shuffled food list asc).  At most
// four candidates tie on (key, distance): one distance, two arms, two sides.
template <int POLT, bool STRIPS>
__device__ static inline void dr_choose(const DevWorld& w, int gx, int gy, int a, uint32_t slot, uint64_t skey, uint8_t own,
                                        DrAccum& acc, int& out_x, int& out_y, int& out_sg) {
    const int n = w.n;
    constexpr bool pol = POLT != 0;
    const int span = 2 * a + 1;
    const int sg0 = own & 0x7f;
    double pown = 0.0;
    if (pol) { const DrLoc L = dr_loc<STRIPS>(w, gx, gy); pown = __ldcg(&DR_F(poll, L)[L.idx]); }
    unsigned long long fm = 0ull;           // free candidates, bit = candidate index k
    unsigned long long tl = 0ull;           // tied best candidates, 16 bits each: harvest << 8 | k
    int nt = 0, dmin = 0x7fffffff;
    uint32_t ikmax = 0u;
    double dkmax = -1.0;
#pragma unroll 1
    for (int d = 1; d <= a; d++) {
        int cx[4], cy[4], kk[4];
        cx[0] = ss_wrap1(gx - d, n); cy[0] = gy; kk[0] = a - d;
        cx[1] = ss_wrap1(gx + d, n); cy[1] = gy; kk[1] = a + d;
        cx[2] = gx; cy[2] = ss_wrap1(gy - d, n); kk[2] = span + a - d;
        cx[3] = gx; cy[3] = ss_wrap1(gy + d, n); kk[3] = span + a + d;
        uint8_t mv[4];
        double pp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const DrLoc L = dr_loc<STRIPS>(w, cx[u], cy[u]);
            mv[u] = __ldcg(&DR_F(mw, L)[L.idx]);
            pp[u] = pol ? __ldcg(&DR_F(poll, L)[L.idx]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (mv[u] & 0x80) continue;
            const int sg = mv[u];
            fm |= 1ull << kk[u];
            const int dist = abs(gx - cx[u]) + abs(gy - cy[u]);                     // agent.py:867-868 (raw coordinates)
            const unsigned long long ent = ((unsigned long long)sg << 8) | (unsigned long long)kk[u];
            if (!pol) {
                const uint32_t ik = (((uint32_t)sg << 20) | (0xfffffu - (uint32_t)dist)) + 1u;
                if (ik > ikmax) { ikmax = ik; tl = ent; nt = 1; }
                else if (ik == ikmax) { tl = (tl << 16) | ent; nt++; }
            } else {
                const double key = (double)sg / (1.0 + (sg > 0 ? pp[u] : 0.0));
                if (key > dkmax || (key == dkmax && dist < dmin)) { dkmax = key; dmin = dist; tl = ent; nt = 1; }
                else if (key == dkmax && dist == dmin) { tl = (tl << 16) | ent; nt++; }
            }
        }
    }
    out_x = gx; out_y = gy; out_sg = sg0;
    if (nt == 0) return;                                                            // no free cell in sight
    // own cell, appended last (agent.py:806); distance 0 wins every tie on the key
    if (!pol) { if (sg0 >= (int)((ikmax - 1u) >> 20)) return; }
    else { if ((double)sg0 / (1.0 + (sg0 > 0 ? pown : 0.0)) >= dkmax) return; }
    int win = 0;
    if (nt > 1) {
        if (nt > 4) { acc.fault++; nt = 4; }
        const int n_free = __popcll(fm);
        int p[4];
#pragma unroll
        for (int t = 0; t < 4; t++) p[t] = t < nt ? __popcll(fm & ((1ull << (int)((tl >> (16 * t)) & 255ull)) - 1ull)) : -1;
        const uint64_t akey = ss_agent_key(skey, slot + 1u);
        uint32_t j = 0;
        int left = nt;
#pragma unroll 1
        for (int i = n_free - 1; i >= 1 && left > 1; i--) {                         // random.py:350 shuffle
            const uint32_t nn = (uint32_t)i + 1u;
            const int kbits = 32 - __clz(nn);
            uint32_t rj = ss_keyed_word(akey, j++) >> (32 - kbits);                 // random.py:242 _randbelow
            while (rj >= nn) {
                rj = ss_keyed_word(akey, j++) >> (32 - kbits);
                if (j > 100000u) { acc.fault++; break; }
            }
            // the element at index rj moves to index i and is final there; L[i] moves to rj
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (p[t] < 0) continue;
                if (p[t] == (int)rj) { p[t] = -1; left--; }
                else if (p[t] == i) p[t] = (int)rj;
            }
        }
        // the surviving tied candidate ends lowest in the shuffled list
#pragma unroll
        for (int t = 3; t >= 0; t--) if (p[t] >= 0) win = t;
    }
    const uint32_t ent = (uint32_t)((tl >> (16 * win)) & 0xffffull);
    const int k = (int)(ent & 255u);
    if (k < span) { out_x = ss_wrap1(gx + k - a, n); out_y = gy; }
    else { out_x = gx; out_y = ss_wrap1(gy + k - span - a, n); }
    out_sg = (int)(ent >> 8);
}

template <int POLT, int MW, bool STRIPS>
__device__ static inline void dr_turn(const DevWorld& w, const DrStep& st, DrShared& sh, DrAccum& acc, uint32_t entry) {
    const int n = w.n, V = w.vmax;
    const DrPeer& S = w.dr.self;
    const uint32_t c = entry & ~DR_SKIP;
    const bool skip = (entry & DR_SKIP) != 0u;
    const uint32_t ow = __ldcg(&S.occw[c]);
    if (!ow) { acc.fault++; return; }
    const uint32_t slot = occ_slot(ow);
    const int a = occ_vision(ow);
    const bool risky = (ow & OCC_RISK) != 0u;
    // loads the turn will want at its end, issued now: the agent's record and its successor bitmap
    AgentRec* recp = &w.agents[slot];
    double sugar = __ldcg(&recp->sugar);
    uint32_t attr = __ldcg(&recp->attr);
    unsigned long long m[MW];
    {
        const ulonglong2* mp = reinterpret_cast<const ulonglong2*>(&w.dr.mask[(size_t)slot * MW]);
#pragma unroll
        for (int i = 0; i < MW / 2; i++) { const ulonglong2 v = __ldcg(&mp[i]); m[2 * i] = v.x; m[2 * i + 1] = v.y; }
    }
    uint2 qw = make_uint2(0u, 0u);
    if (risky) qw = __ldcg(&w.dr.q1wait[slot]);
    const int lx = (int)(c / (uint32_t)n), gy = (int)(c - (uint32_t)lx * (uint32_t)n), gx = S.x0 + lx;
    bool died = false;
    if (!skip) {
        const uint8_t own = __ldcg(&S.mw[c]);
        int nx = gx, ny = gy, sg = 0;
        if (w.cfg.rule_move_eat) dr_choose<POLT, STRIPS>(w, gx, gy, a, slot, st.skey, own, acc, nx, ny, sg);
        const DrLoc D = dr_loc<STRIPS>(w, nx, ny);
        const bool moved = !(D.which == 0 && D.idx == c);
        const bool pol = POLT != 0 && w.cfg.pollution_start_time <= st.env_time;
        const int metab = attr_metab(attr);
        double p = 0.0;
        if (pol) p = __ldcg(&DR_F(poll, D)[D.idx]);
        if (w.cfg.rule_move_eat) {
            if (pol) p = p + (((double)sg * w.cfg.pollution_a) + (0.0 * w.cfg.pollution_b));     // agent.py:825-826
            sugar = ss_set_sugar(ss_max0(sugar + (double)sg));                                     // agent.py:832
            DR_F(sugar, D)[D.idx] = 0.0;                                                               // agent.py:864
        }
        if (pol && sugar > 0.0) p = p + ((0.0 * w.cfg.pollution_a) + ((double)metab * w.cfg.pollution_b));   // sugarscape.py:566-567
        if (pol) DR_F(poll, D)[D.idx] = p;
        sugar = ss_set_sugar(ss_max0(sugar - (double)metab));                                      // sugarscape.py:569
        acc.sum_m += (unsigned)metab; acc.sum_v += (unsigned)a; acc.acted++;
        if (sugar == 0.01) acc.small++;
        else if (sugar >= 0.0 && sugar < (double)SS_GINI_BINS && sugar == (double)(int)sugar) {
            const int bin = (int)sugar;
            if (bin < DR_HIST_BINS) atomicAdd(&sh.hist[bin], 1u); else atomicAdd(&w.hist[bin], 1u);
        } else acc.inexact++;
        died = w.cfg.rule_can_starve && sugar <= 0.1;                                              // sugarscape.py:306-309
        const bool risk_next = !died && dr_at_risk(w, sugar, metab);
        uint32_t new_word = 0u;
        if (!died) new_word = occ_pack(slot, a, st.next_phase) | (risk_next ? OCC_RISK : 0u);
        if (moved) { S.occw[c] = 0u; S.mw[c] = own & 0x7f; }
        DR_F(occw, D)[D.idx] = new_word;
        DR_F(mw, D)[D.idx] = died ? 0 : 0x80;
        if (died) {
            attr &= ~ATTR_ALIVE; acc.deaths++;
            atomicAnd(&w.dr.alive[slot >> 5], ~(1u << (slot & 31)));
        }
        if (risk_next != risky) atomicXor(&w.dr.atrisk[slot >> 5], 1u << (slot & 31));
        recp->sugar = sugar;
        recp->pos = D.idx;
        recp->attr = attr;
    } else {
        S.occw[c] = (ow & 0x7fffffffu) | (st.next_phase << 31);
    }
    // Q1: whoever follows this agent in the list order learns the outcome (a skipped agent did not die)
    if (risky && qw.y == (uint32_t)(st.iteration + 1)) {
        const uint32_t qx = qw.x / (uint32_t)n;
        dr_release<STRIPS>(w, sh, acc, S.x0 + (int)qx, (int)(qw.x - qx * (uint32_t)n), died);
    }
    // conflict successors: start-of-step positions relative to this agent's start cell
#pragma unroll
    for (int i = 0; i < MW; i++) {
        unsigned long long bits = m[i];
        while (bits) {
            const int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1ull;
            int dx, dy;
            dr_bit_offset(i * 64 + b, V, st.magicS, dx, dy);
            dr_release<STRIPS>(w, sh, acc, ss_wrap1(gx + dx, n), ss_wrap1(gy + dy, n), false);
        }
    }
}

// Grid barrier between rounds (cooperative launch: all CTAs are resident).  The last CTA to arrive snapshots the queue
// tail into ctl[3 + (generation & 1)] before it releases the others, so every CTA sees the same next segment even if a
// fast CTA is already appending to the queue again.  Returns the snapshot.
__device__ static inline unsigned dr_grid_barrier(unsigned int* ctl, unsigned nblocks, DrAccum& acc, unsigned* s_tail) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned gen = *(volatile unsigned*)&ctl[2];
        if (atomicAdd(&ctl[1], 1u) == nblocks - 1u) {
            ctl[1] = 0u;
            *(volatile unsigned*)&ctl[3 + (gen & 1u)] = *(volatile unsigned*)&ctl[0];
            __threadfence();
            atomicAdd(&ctl[2], 1u);
        } else {
            unsigned spins = 0;
            while (*(volatile unsigned*)&ctl[2] == gen) {
                __nanosleep(64);
                if (++spins > (1u << 26)) { acc.fault++; break; }      // a lost CTA: do not hang the GPU
            }
        }
        __threadfence();
        *s_tail = *(volatile unsigned*)&ctl[3 + (gen & 1u)];
    }
    __syncthreads();
    return *s_tail;
}

template <int POLT, int MW, bool STRIPS>
__global__ void __launch_bounds__(DR_THREADS, 2)
ss_dr_rounds_kernel(DevWorld w, DrStep st) {
    __shared__ DrShared sh;
    __shared__ unsigned s_tail;
    const DrPeer& S = w.dr.self;
    const int tid = threadIdx.x;
    for (int i = tid; i < DR_HIST_BINS; i += DR_THREADS) sh.hist[i] = 0u;
    if (tid < 8) sh.acc[tid] = 0ull;
    if (tid == 0) sh.cnt = 0;
    __syncthreads();
    DrAccum acc = {0, 0, 0, 0, 0, 0, 0};
    unsigned seg_b = st.single_round ? st.seg_begin : 0u;
    unsigned seg_e = st.single_round ? st.seg_end : __ldcg(&S.ctl[5]);     // ctl[5]: tail after the build (ss_dr_seg0_kernel)
    for (;;) {
        const unsigned cnt = seg_e - seg_b;
        if (cnt == 0u) break;
        for (unsigned base = blockIdx.x * DR_THREADS; base < cnt; base += gridDim.x * DR_THREADS) {
            const unsigned i = base + (unsigned)tid;
            if (i < cnt) dr_turn<POLT, MW, STRIPS>(w, st, sh, acc, __ldcg(&S.queue[seg_b + i]));
            __syncthreads();
            const int staged = sh.cnt;                    // the same value in every thread: nobody pushes between the barriers
            __syncthreads();
            if (staged >= DR_PUSH_CAP / 2) {
                const int k = min(staged, DR_PUSH_CAP);
                if (tid == 0) sh.base = atomicAdd(&S.ctl[0], (unsigned)k);
                __syncthreads();
                for (int j = tid; j < k; j += DR_THREADS) S.queue[sh.base + (unsigned)j] = sh.push[j];
                __syncthreads();
                if (tid == 0) sh.cnt = 0;
                __syncthreads();
            }
        }
        {
            const int k = min(sh.cnt, DR_PUSH_CAP);
            if (k > 0) {
                if (tid == 0) sh.base = atomicAdd(&S.ctl[0], (unsigned)k);
                __syncthreads();
                for (int j = tid; j < k; j += DR_THREADS) S.queue[sh.base + (unsigned)j] = sh.push[j];
                __syncthreads();
                if (tid == 0) sh.cnt = 0;
            }
        }
        if (st.single_round) break;
        seg_b = seg_e;
        seg_e = dr_grid_barrier(S.ctl, st.nblocks, acc, &s_tail);
    }
    // statistics of this CTA's turns (sugarscape.py:575-583)
    unsigned vals[7] = {acc.sum_m, acc.sum_v, acc.acted, acc.deaths, acc.small, acc.inexact, acc.fault};
#pragma unroll
    for (int q = 0; q < 7; q++) {
        unsigned v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&sh.acc[q], (unsigned long long)v);
    }
    __syncthreads();
    unsigned long long* dst[7] = {&w.acc->sum_metab, &w.acc->sum_vision, &w.acc->acted, &w.acc->deaths, &w.acc->hist_small,
                                  &w.acc->inexact, &w.acc->fault};
    if (tid < 7 && sh.acc[tid]) atomicAdd(dst[tid], sh.acc[tid]);
    for (int i = tid; i < DR_HIST_BINS; i += DR_THREADS) {
        const unsigned cc = sh.hist[i];
        if (cc) atomicAdd(&w.hist[i], cc);
    }
}

// every agent that was alive at the start of the step must have passed through the queue
__global__ void ss_dr_check_kernel(DevWorld w) {
    if (w.dr.self.ctl[0] != (unsigned)w.acc->pending) atomicAdd(&w.acc->fault, 1ull);
    w.acc->pending = 0ull;
}
__global__ void ss_dr_begin_kernel(DevWorld w) { w.dr.self.ctl[0] = 0u; }
__global__ void ss_dr_seg0_kernel(DevWorld w) { w.dr.self.ctl[5] = w.dr.self.ctl[0]; }

// ---------------------------------------------------------------------------- launchers ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int POLT, int MW, bool STRIPS>
static cudaError_t launch_rounds(const DevWorld& w, DrStep st, int sm_count, cudaStream_t s) {
    static int per_sm = 0;
    if (!per_sm) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ss_dr_rounds_kernel<POLT, MW, STRIPS>, DR_THREADS, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    st.nblocks = (unsigned)(per_sm * sm_count);
    if (st.single_round) {
        ss_dr_rounds_kernel<POLT, MW, STRIPS><<<st.nblocks, DR_THREADS, 0, s>>>(w, st);
        return cudaGetLastError();
    }
    void* args[2] = {(void*)&w, (void*)&st};
    return cudaLaunchCooperativeKernel((const void*)ss_dr_rounds_kernel<POLT, MW, STRIPS>, dim3(st.nblocks), dim3(DR_THREADS), args, 0, s);
}

extern "C" {

size_t ss_dr_build_smem(int vmax) {
    const int H = 2 * vmax, SX = DR_TX + 2 * H, SY = DR_TY + 2 * H, BW = (SY + 31) >> 5;
    return (size_t)SX * SY * 4 + (size_t)SX * BW * 4 + (size_t)DR_TX * DR_TY * 2;
}

cudaError_t ss_launch_dr_mirror(const DevWorld& w, int* flag, cudaStream_t s) {
    const size_t cells = (size_t)w.dr.self.rows * w.n;
    ss_dr_mirror_kernel<<<148 * 8, 256, 0, s>>>(w, cells, flag);
    if (w.n_slots) ss_dr_flags_kernel<<<(w.n_slots + 255) / 256, 256, 0, s>>>(w);
    return cudaGetLastError();
}

// tmap_storage: 128 bytes, 64-byte aligned, host memory owned by the engine (re-encoded when the lattice changes)
cudaError_t ss_dr_encode_tmap(const DevWorld& w, void* tmap_storage, int* ok) {
    *ok = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    const int n = w.n, H = 2 * w.vmax;
    if (!enc || n % 4 != 0) return cudaSuccess;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)w.dr.self.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
    const cuuint32_t box[2] = {(cuuint32_t)(DR_TY + 2 * H), (cuuint32_t)(DR_TX + 2 * H)};
    const cuuint32_t estr[2] = {1, 1};
    if (box[0] > 256 || box[1] > 256) return cudaSuccess;
    const CUresult r = enc((CUtensorMap*)tmap_storage, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)w.dr.self.occw, dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    *ok = r == CUDA_SUCCESS;
    return cudaSuccess;
}

cudaError_t ss_launch_dr_build(const DevWorld& w, int64_t iteration, const void* tmap_storage, int use_tma, cudaStream_t s) {
    static size_t configured = 0;
    const size_t smem = ss_dr_build_smem(w.vmax);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ss_dr_build_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    ss_dr_begin_kernel<<<1, 1, 0, s>>>(w);
    DrBuildArgs ba;
    ba.ok = ss_order_key(ss_step_key(w.cfg.keyed_seed, iteration), w.half);
    ba.iteration = iteration;
    ba.tiles_y = (w.n + DR_TY - 1) / DR_TY;
    ba.use_tma = use_tma;
    const int tiles_x = (w.dr.self.rows + DR_TX - 1) / DR_TX;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (use_tma) memcpy(&tm, tmap_storage, sizeof tm);
    const bool strips = w.dr.self.rows != w.n;
    const int grid = tiles_x * ba.tiles_y;
    if (w.dr.mwords == 4) {
        if (strips) ss_dr_build_kernel<4, true><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
        else ss_dr_build_kernel<4, false><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
    } else {
        if (strips) ss_dr_build_kernel<16, true><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
        else ss_dr_build_kernel<16, false><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
    }
    ss_dr_seg0_kernel<<<1, 1, 0, s>>>(w);
    return cudaGetLastError();
}

// single_round = 0: the whole agent phase in one cooperative launch; 1: the queue segment [seg_begin, seg_end) only
cudaError_t ss_launch_dr_rounds(const DevWorld& w, int64_t iteration, int64_t env_time, int sm_count, int single_round,
                                unsigned seg_begin, unsigned seg_end, cudaStream_t s) {
    DrStep st;
    st.iteration = iteration; st.env_time = env_time;
    st.skey = ss_step_key(w.cfg.keyed_seed, iteration);
    st.next_phase = (uint32_t)((iteration + 1) & 1);
    const uint32_t S = 2u * (uint32_t)w.vmax + 1u;
    st.magicS = (65536u + S - 1u) / S;
    st.single_round = single_round; st.seg_begin = seg_begin; st.seg_end = seg_end; st.nblocks = 0;
    const bool pol = w.cfg.rule_pollution != 0;
    const bool strips = w.dr.self.rows != w.n;
    if (strips) {
        if (w.dr.mwords == 4) return pol ? launch_rounds<1, 4, true>(w, st, sm_count, s) : launch_rounds<0, 4, true>(w, st, sm_count, s);
        return pol ? launch_rounds<1, 16, true>(w, st, sm_count, s) : launch_rounds<0, 16, true>(w, st, sm_count, s);
    }
    if (w.dr.mwords == 4) return pol ? launch_rounds<1, 4, false>(w, st, sm_count, s) : launch_rounds<0, 4, false>(w, st, sm_count, s);
    return pol ? launch_rounds<1, 16, false>(w, st, sm_count, s) : launch_rounds<0, 16, false>(w, st, sm_count, s);
}

cudaError_t ss_launch_dr_check(const DevWorld& w, cudaStream_t s) {
    ss_dr_check_kernel<<<1, 1, 0, s>>>(w);
    return cudaGetLastError();
}

int ss_dr_mask_words(int vmax) { return dr_mask_words(vmax); }

}  // extern "C"
