// ss_rounds.cu -- the agent phase as dependency rounds over the WHOLE lattice (keyed RNG, rule M).
//
// View.updateGame's agent loop (sugarscape.py:517-601) is sequential in a random order.  Two agents conflict iff
// their vision crosses (agent.py:361-371) share a cell -- every read and write of a turn lies inside the agent's own
// cross -- so the sequential result is reproduced exactly when every agent acts after all conflicting agents of
// lower execution priority.  That is a DAG over the agents' start-of-step positions; this file evaluates it level by
// level without windows, lists, sorts or host round trips:
//
//   ss_dr_build_kernel   one CTA per 64 x 64-cell tile stages the occupancy words of the tile and a halo of 2*vmax
//                        cells in shared memory (TMA 2-D tile load, cp.async.bulk.tensor, for tiles that do not touch
//                        the torus edge), turns them into (priority, vision) words + an occupancy bitmap, and one
//                        thread per agent scans its conflict zone with the exact cross-intersection predicate:
//                        predecessors are COUNTED (dep[slot], an L2-resident u16 array), successors are recorded as a
//                        list of slots (succ[slot]: one 64-byte record, the rare long lists continue in a pool).  The list-order predecessor (the agent before it in
//                        the keyed order, found by walking the inverse priority map over the alive bitmap) adds one
//                        more edge when it may starve this step (Q1: sugarscape.py:520 + 502-503, the agent after a
//                        dying agent loses its turn).  Agents without predecessors enter the ready queue.
//   ss_dr_rounds_kernel  one persistent cooperative kernel per step.  One thread per ready agent -- both arms
//                        of the cross as two 16-byte loads each on the one-byte cell words (occupied | int(sugar);
//                        the x-arm from a transposed copy), choice (agent.py:794-823 incl. the
//                        replay of the agent's Fisher-Yates draws for ties; or the utilicalc turn, agent.py:1025-1141),
//                        harvest, pollution, metabolism, death (sugarscape.py:566-601), statistics -- then one atomic
//                        decrement per successor; whoever brings a counter to zero appends that agent to the queue.
//                        The queue is run as a WORK LIST: warps claim its next positions and wait for the entries to be
//                        published, fences order a turn's stores ahead of the counters it decrements -- no barrier
//                        between dependency rounds (the version with one grid barrier per round stays: option
//                        async_queue = 0, and the host-driven strips of the tests).
//
// Everything an agent's turn needs is settled in the turn (no deferred pass), so the only per-step kernels around
// these two are the statistics and the growback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "ss_world.cuh"

#define FULL 0xffffffffu
#define DR_TX 64                        // tile rows (x)
#define DR_TY 64                        // tile columns (y, contiguous in memory)
#define DR_BUILD_THREADS 512
#ifndef DR_THREADS
#define DR_THREADS 512                  // rounds kernel
#endif
#ifndef DR_ASYNC_MUL
#define DR_ASYNC_MUL 8u                 // work-list mode: a warp claims backlog * MUL / warps positions, at most 32
#endif
#ifndef DR_ASYNC_SLEEP
#define DR_ASYNC_SLEEP 100              // ... and polls for unpublished entries every so many ns
#endif
#ifndef DR_MIN_BLOCKS
#define DR_MIN_BLOCKS 2                 // ... resident CTAs per SM the register allocation aims at
#endif
#define DR_WARP_CAP 128                 // staged queue entries per warp between flushes
#define DR_HIST_BINS 2048               // low wealth bins kept in shared memory per CTA
#define DR_SKIP 0x80000000u             // queue entry flag: the agent loses its turn (Q1)
#define DR_REMOTE 0x20000000u           // queue entry flag: released by an agent of another strip
#define DR_VALID 0x40000000u            // queue entry flag: the entry is written (strips: a neighbour may still be writing it)

// ------------------------------------------------------------------------------ helpers ----
__device__ static inline bool dr_crosses_meet(int dx, int dy, int a, int b) {
    const int adx = abs(dx), ady = abs(dy);
    return (adx <= a && ady <= b) || (ady <= a && adx <= b) || (dy == 0 && adx <= a + b) || (dx == 0 && ady <= a + b);
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
    if (below <= 2 * w.vmax) {
        L.which = 1; L.idx = (uint32_t)(w.dr.lo.rows - below) * (uint32_t)w.n + (uint32_t)gy; return L;
    }
    int above = gx - (w.dr.self.x0 + w.dr.self.rows);
    if (above < 0) above += w.n;
    L.which = 2; L.idx = (uint32_t)above * (uint32_t)w.n + (uint32_t)gy;
    return L;
}
#define DR_W(field, which) ((!STRIPS || (which) == 0) ? w.dr.self.field : ((which) == 1 ? w.dr.lo.field : w.dr.hi.field))
#define DR_F(field, L) DR_W(field, (L).which)
#define DR_A(field, rank) (STRIPS ? w.dr.all[(rank)].field : w.dr.self.field)
// the transposed cell byte of local cell idx = r * n + y of strip `which`: y * rows + r
template <bool STRIPS>
__device__ static inline size_t dr_xidx(const DevWorld& w, const DrLoc& L) {
    const uint32_t r = L.idx / (uint32_t)w.n, y = L.idx - r * (uint32_t)w.n;
    return (size_t)y * (size_t)DR_F(rows, L) + r;
}

// ------------------------------------------------------------------- mirrors and flags ----
// mw from the uploaded state; flag: a value does not fit 7 bits (the path is then refused)
__global__ void ss_dr_mirror_kernel(DevWorld w, size_t cells, int* flag) {
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const double v = w.sugar[c];
        int iv = 0;
        if (!(v >= 0.0 && v < 127.0) || !(w.cap[c] < 127.0)) *flag = 1; else iv = (int)v;
        w.dr.self.mw[c] = (uint8_t)(iv | (w.occw[c] ? 0x80 : 0));
    }
}
// mwx = mw transposed (64 x 64-byte tiles through shared memory): after an upload and after every growback
__global__ void __launch_bounds__(256)
ss_dr_transpose_kernel(const uint8_t* __restrict__ mw, uint8_t* __restrict__ mwx, int rows, int n) {
    __shared__ uint8_t t[64][68];
    const int tiles_y = (n + 63) >> 6;
    const int bx = blockIdx.x / tiles_y, by = blockIdx.x - bx * tiles_y;
    const int r0 = bx << 6, y0 = by << 6;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int k = ty; k < 64; k += 4)
        if (r0 + k < rows && y0 + tx < n) t[k][tx] = mw[(size_t)(r0 + k) * n + y0 + tx];
    __syncthreads();
    for (int k = ty; k < 64; k += 4)
        if (y0 + k < n && r0 + tx < rows) mwx[(size_t)(y0 + k) * rows + r0 + tx] = t[tx][k];
}
// The turns leave two things to be tidied lazily: a harvested cell keeps its old f64 sugar and carries the marker 127
// in its byte instead (the next growback -- or this kernel -- takes the sugar as 0), and a cell an agent left keeps
// its occupancy word (the byte's occupied bit is what counts).  This kernel makes sugar / occw canonical again: before
// a download and before any other path reads them.
#define DR_HARVESTED 127u
__global__ void ss_dr_canon_kernel(DevWorld w, size_t cells) {
    const DrPeer& S = w.dr.self;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const uint8_t b = S.mw[c];
        if (!(b & 0x80)) S.occw[c] = 0u;
        if ((b & 0x7f) == DR_HARVESTED) {
            S.sugar[c] = 0.0;
            const size_t r = c / (size_t)w.n, y = c - r * (size_t)w.n;
            S.mw[c] = b & 0x80;
            S.mwx[y * (size_t)S.rows + r] = b & 0x80;
        }
    }
}
// environment.py:133-155 growback (whole lattice or the two season rectangles) on 64 x 64-cell tiles, fused with the
// refresh of the cell bytes in both layouts: 8 + 8 + 1 bytes read, 8 + 1 + 1 written per cell.
struct DrGrowArgs {
    int seasons, tiles;
    double alpha, alpha_north, alpha_south;
};
__global__ void __launch_bounds__(256)
ss_dr_grow_kernel(DevWorld w, DrGrowArgs ga) {
    __shared__ uint8_t t[64][68];
    const DrPeer& S = w.dr.self;
    const int n = w.n, rows = S.rows;
    const int tiles_y = (n + 63) >> 6;
    // (one CTA per tile, or -- beside the conflict-graph kernel of the next step -- a few CTAs per SM that walk the tiles)
    for (int tile = blockIdx.x; tile < ga.tiles; tile += gridDim.x) {
    const int bx = tile / tiles_y, by = tile - bx * tiles_y;
    const int r0 = bx << 6, y0 = by << 6;
    const int32_t* rn = w.cfg.season_north;
    const int32_t* rs = w.cfg.season_south;
    const int n_imin = max(rn[0], 0), n_jmin = max(rn[1], 0), n_imax = min(rn[2] + 1, n), n_jmax = min(rn[3] + 1, n);
    const int s_imin = max(rs[0], 0), s_jmin = max(rs[1], 0), s_imax = min(rs[2] + 1, n), s_jmax = min(rs[3] + 1, n);
    const int tq = (threadIdx.x & 31) * 2, tr = threadIdx.x >> 5;
    const bool vec = (n & 1) == 0;
    for (int k = tr; k < 64; k += 8) {
        const int r = r0 + k;
        if (r >= rows) break;
        const int x = S.x0 + r;
        const size_t c = (size_t)r * n + y0 + tq;
        double sv[2] = {0.0, 0.0}, cv[2] = {0.0, 0.0};
        uint8_t bv[2] = {0, 0};
        const bool in0 = y0 + tq < n, in1 = y0 + tq + 1 < n;
        if (vec && in1) {
            const double2 s2 = *reinterpret_cast<const double2*>(&S.sugar[c]);
            const double2 c2 = __ldcs(reinterpret_cast<const double2*>(&w.cap[c]));
            const uchar2 b2 = *reinterpret_cast<const uchar2*>(&S.mw[c]);
            sv[0] = s2.x; sv[1] = s2.y; cv[0] = c2.x; cv[1] = c2.y; bv[0] = b2.x; bv[1] = b2.y;
        } else {
            if (in0) { sv[0] = S.sugar[c]; cv[0] = w.cap[c]; bv[0] = S.mw[c]; }
            if (in1) { sv[1] = S.sugar[c + 1]; cv[1] = w.cap[c + 1]; bv[1] = S.mw[c + 1]; }
        }
        bool grew[2] = {false, false};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int y = y0 + tq + u;
            double s = (bv[u] & 0x7f) == DR_HARVESTED ? 0.0 : sv[u];
            if (ga.seasons) {
                const bool in_n = x >= n_imin && x < n_imax && y >= n_jmin && y < n_jmax;
                const bool in_s = x >= s_imin && x < s_imax && y >= s_jmin && y < s_jmax;
                if (in_n) { const double v = s + ga.alpha_north; s = v <= cv[u] ? v : cv[u]; }
                if (in_s) { const double v = s + ga.alpha_south; s = v <= cv[u] ? v : cv[u]; }
                grew[u] = in_n || in_s;
            } else {
                const double v = s + ga.alpha; s = v <= cv[u] ? v : cv[u];
                grew[u] = true;
            }
            if (grew[u]) { sv[u] = s; bv[u] = (uint8_t)((bv[u] & 0x80) | (uint8_t)(int)s); }
        }
        if (vec && in1) {
            if (grew[0] || grew[1]) {
                // (a cell that did not grow keeps its stale sugar under the harvested marker: write back what was read)
                *reinterpret_cast<double2*>(&S.sugar[c]) = make_double2(sv[0], sv[1]);
                *reinterpret_cast<uchar2*>(&S.mw[c]) = make_uchar2(bv[0], bv[1]);
            }
        } else {
            if (in0 && grew[0]) { S.sugar[c] = sv[0]; S.mw[c] = bv[0]; }
            if (in1 && grew[1]) { S.sugar[c + 1] = sv[1]; S.mw[c + 1] = bv[1]; }
        }
        t[k][tq] = bv[0]; t[k][tq + 1] = bv[1];
    }
    __syncthreads();
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int k = ty; k < 64; k += 4)
        if (y0 + k < n && r0 + tx < rows) S.mwx[(size_t)(y0 + k) * rows + r0 + tx] = t[tx][k];
    __syncthreads();
    }
}

// alive / at-risk bitmaps and the ATTR_RISK flag from the agent store (after an upload).
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
        AgentRec* rp = &w.agents[slot];
        const uint32_t attr = rp->attr;
        alive = attr_alive(attr);
        if (alive) {
            risk = dr_at_risk(w, rp->sugar, attr_metab(attr));
            rp->attr = risk ? (attr | ATTR_RISK) : (attr & ~ATTR_RISK);
        }
    }
    const unsigned ba = __ballot_sync(FULL, alive), br = __ballot_sync(FULL, risk);
    // (strips: the bitmaps cover the agents of every strip and come from ss_dr_directory_kernel)
    if (w.dr.world == 1 && (threadIdx.x & 31) == 0 && (slot >> 5) < ((w.n_slots + 31) >> 5)) { w.dr.alive[slot >> 5] = ba; w.dr.atrisk[slot >> 5] = br; }
}
// strips: the replicated directories from the agents of the WHOLE world (every rank gets the same arrays at upload):
// alive / at-risk bitmaps and the strip that holds each agent
__global__ void ss_dr_directory_kernel(DevWorld w, int n_all, const int32_t* __restrict__ id, const int32_t* __restrict__ x,
                                       const int32_t* __restrict__ metab, const double* __restrict__ sugar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    const uint32_t slot = (uint32_t)id[i] - 1u;
    if (slot >= w.n_slots) return;
    atomicOr(&w.dr.alive[slot >> 5], 1u << (slot & 31));
    if (dr_at_risk(w, sugar[i], metab[i])) atomicOr(&w.dr.atrisk[slot >> 5], 1u << (slot & 31));
    w.dr.home[slot] = (uint8_t)(((long long)(x[i] + 1) * w.dr.world - 1) / w.n);      // rows [r n / P, (r + 1) n / P) belong to rank r
}
// strips, after the rounds: every rank applies every strip's delta list to its replicated directories
__global__ void __launch_bounds__(256)
ss_dr_apply_delta_kernel(DevWorld w) {
    const int p = blockIdx.y;
    if (p >= w.dr.world) return;
    const DrPeer& P = w.dr.all[p];
    const unsigned cnt = __ldcg(&P.ctl[DR_CTL_DELTA]);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint32_t e = __ldcg(&P.delta[i]);
        const uint32_t slot = e & OCC_SLOT_MASK, kind = e >> 24;
        if (kind & DR_DELTA_DIED) atomicAnd(&w.dr.alive[slot >> 5], ~(1u << (slot & 31)));
        if (kind & DR_DELTA_RISK) atomicXor(&w.dr.atrisk[slot >> 5], 1u << (slot & 31));
        if (kind & DR_DELTA_TO_LO) w.dr.home[slot] = (uint8_t)(p == 0 ? w.dr.world - 1 : p - 1);
        if (kind & DR_DELTA_TO_HI) w.dr.home[slot] = (uint8_t)(p == w.dr.world - 1 ? 0 : p + 1);
    }
}
// strips: the step's sums and the wealth histogram over all strips (every rank computes the same statistics)
__global__ void __launch_bounds__(256)
ss_dr_merge_stats_kernel(DevWorld w) {
    const int P = w.dr.world;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SS_GINI_BINS; i += gridDim.x * blockDim.x) {
        uint32_t h = 0;
        for (int p = 0; p < P; p++) h += __ldcg(&w.dr.all[p].hist[i]);
        w.dr.hist_g[i] = h;
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        // StepAccum: sum_metab, sum_vision, acted, deaths, pending, population, hist_small, inexact, fault, ...
        const int f = threadIdx.x == 4 ? 8 : threadIdx.x;        // (pending is per strip: its slot carries the faults here)
        if (f != 5) {
            unsigned long long v = 0;
            for (int p = 0; p < P; p++) v += __ldcg(&w.dr.all[p].acc[f]);
            w.dr.acc_g[f] = v;
        }
    }
}


// -------------------------------------------------------------------------------- build ----
__device__ static inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Q1 (sugarscape.py:520 + 502-503: the agent after a dying agent loses its turn): the list order is the keyed order --
// ascending priority over the living agents -- so an agent's neighbours in the list are found by walking the inverse
// priority map over the replicated alive bitmap.  If the agent before this one may starve this step, its outcome
// decides whether this agent acts at all: one more predecessor.
__device__ static inline bool dr_q1_pred_at_risk(const DevWorld& w, const OrderKey& ok, uint32_t prio) {
    for (uint32_t p = prio; p-- > 0u;) {
        const uint32_t s2 = ss_priority_inverse(ok, p);
        if (s2 >= w.n_slots || !((__ldg(&w.dr.alive[s2 >> 5]) >> (s2 & 31)) & 1u)) continue;
        return ((__ldg(&w.dr.atrisk[s2 >> 5]) >> (s2 & 31)) & 1u) != 0u;
    }
    return false;
}
// ... and an agent that may starve notes whom to tell: the next living agent in the list (slot + 1, 0 = nobody)
__device__ static inline uint32_t dr_q1_successor(const DevWorld& w, const OrderKey& ok, uint32_t prio) {
    const uint32_t domain = 1u << (2 * ok.half);
    for (uint32_t p = prio + 1u; p < domain; p++) {
        const uint32_t s2 = ss_priority_inverse(ok, p);
        if (s2 < w.n_slots && ((__ldg(&w.dr.alive[s2 >> 5]) >> (s2 & 31)) & 1u)) return s2 + 1u;
    }
    return 0u;
}

struct DrBuildArgs {
    OrderKey ok;
    uint32_t stamp;                 // iteration + 1
    int tiles_y;
    int use_tma;
};

// Conflict graph of one tile.  The work is split by its shape: the uniform part -- pulling the occupied cells of an
// agent's conflict zone out of the row / column bitmaps -- runs one THREAD per agent (a warp takes up to 32 agents);
// the irregular part -- one (agent, occupied cell) pair each: exact predicate, priorities, predecessor or successor --
// runs one LANE per pair on the warp's pair list in shared memory, so every lane is busy whatever the local density.
// WIDE = vision > 7 (64-bit windows, 32-bit pair entries).
template <bool WIDE> struct DrWide;
template <> struct DrWide<false> { typedef uint32_t Win; typedef uint16_t Pair; enum { SHIFT = 5 }; };
template <> struct DrWide<true> { typedef unsigned long long Win; typedef uint32_t Pair; enum { SHIFT = 7 }; };
template <bool WIDE>
__device__ static inline typename DrWide<WIDE>::Win dr_window(const uint32_t* row, int lo, int width) {
    typedef typename DrWide<WIDE>::Win T;
    const int w0 = lo >> 5, sh = lo & 31;
    if (!WIDE) {
        const uint32_t v = __funnelshift_r(row[w0], row[w0 + 1], sh);
        return (T)(v & (width >= 32 ? FULL : ((1u << width) - 1u)));
    }
    const unsigned long long a = (unsigned long long)row[w0] | ((unsigned long long)row[w0 + 1] << 32);
    unsigned long long v = a >> sh;
    if (sh) v |= (unsigned long long)row[w0 + 2] << (64 - sh);
    return (T)(v & (width >= 64 ? ~0ull : ((1ull << width) - 1ull)));
}
template <bool WIDE> __device__ static inline int dr_popc(typename DrWide<WIDE>::Win v) { return WIDE ? __popcll((unsigned long long)v) : __popc((uint32_t)v); }
template <bool WIDE> __device__ static inline int dr_ffs(typename DrWide<WIDE>::Win v) { return WIDE ? __ffsll((long long)v) - 1 : __ffs((uint32_t)v) - 1; }

#define DR_PAIR_CAP 1024                // pair list entries per warp
template <bool STRIPS, bool WIDE>
__global__ void __launch_bounds__(DR_BUILD_THREADS)
ss_dr_build_kernel(DevWorld w, DrBuildArgs ba, const __grid_constant__ CUtensorMap tmap) {
    typedef typename DrWide<WIDE>::Win WinT;
    typedef typename DrWide<WIDE>::Pair PairT;
    constexpr int PS = DrWide<WIDE>::SHIFT;                // bits per coordinate in a pair entry
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ int s_all, s_n;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_cnt[DR_BUILD_THREADS / 32][2][32];   // per warp and agent of the batch: predecessors, successors
    // HP: the halo as staged, 2V rounded up to a multiple of 4 cells (the TMA box wants rows of a multiple of 32 bytes)
    const int n = w.n, V = w.vmax, H = 2 * V, HP = (H + 3) & ~3, SX = DR_TX + 2 * HP, SY = DR_TY + 2 * HP;
    const int BW = ((SY + 31) >> 5) + (WIDE ? 2 : 1), BWT = ((SX + 31) >> 5) + (WIDE ? 2 : 1);
    uint32_t* tileS = sm;                                  // [SX][SY] occupancy words (slot | vision | flags), 0 = free
    uint32_t* tileP = tileS + SX * SY;                     // [SX][SY] priority << 5 | vision
    uint32_t* bm = tileP + SX * SY;                        // [SX][BW]  occupancy bitmap by row
    uint32_t* bmt = bm + SX * BW;                          // [SY][BWT] ... and by column
    uint16_t* ilist = (uint16_t*)(bmt + SY * BWT);         // the tile's own agents: r << 7 | q  (bit 15: ready)
    uint16_t* alist = ilist + DR_TX * DR_TY;               // every agent of tile + halo: r * SY + q; later the warps' pair lists
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = DR_BUILD_THREADS / 32;
    const DrPeer& S = w.dr.self;
    const int tx = blockIdx.x / ba.tiles_y, ty = blockIdx.x - tx * ba.tiles_y;
    const int x0 = S.x0 + tx * DR_TX, y0 = ty * DR_TY;     // lattice coordinates of the tile's first interior cell
    const int ex = min(DR_TX, S.x0 + S.rows - x0), ey = min(DR_TY, n - y0);      // interior extent
    if (tid == 0) { s_all = 0; s_n = 0; }
    for (int i = tid; i < SX * BW + SY * BWT; i += DR_BUILD_THREADS) bm[i] = 0u;
    // the tile and its halo lie inside this strip and inside [0, n) in y: one TMA box; otherwise plain loads with wraps
    const bool interior = ba.use_tma && x0 - HP >= S.x0 && x0 + DR_TX + HP <= S.x0 + S.rows && y0 - HP >= 0 && y0 + DR_TY + HP <= n;
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
                ::"r"(smem_u32(tileS)), "l"(&tmap), "r"(y0 - HP), "r"(x0 - HP - S.x0), "r"(smem_u32(&s_bar)) : "memory");
        }
        // a word counts only where the cell byte says "occupied" (the turns leave the words of vacated cells behind):
        // the bytes of the staged rows, four cells per lane, are fetched while the box is in flight
        // (rows start at multiples of 4 bytes: n % 4 == 0 with the TMA path, y0 and HP are multiples of 4; SX, SY <= 128)
        uint32_t mb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = warp + i * nwarps;
            mb[i] = 0x80808080u;
            if (r < SX && 4 * lane < SY)
                mb[i] = __ldcg(reinterpret_cast<const uint32_t*>(S.mw + (size_t)(x0 - HP - S.x0 + r) * n + (y0 - HP)) + lane);
        }
        uint32_t done = 0;                                 // everybody waits for the box (phase 0)
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&s_bar)) : "memory");
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = warp + i * nwarps;
            if (r < SX && 4 * lane < SY) {
#pragma unroll
                for (int j = 0; j < 4; j++) if (!((mb[i] >> (8 * j)) & 0x80u)) tileS[r * SY + 4 * lane + j] = 0u;
            }
        }
    } else {
        __syncthreads();
        for (int r = warp; r < SX; r += nwarps) {
            const int gx = ss_wrap(x0 - HP + r, n);
            for (int q = lane; q < SY; q += 32) {
                const int gy = ss_wrap(y0 - HP + q, n);
                const DrLoc L = dr_loc<STRIPS>(w, gx, gy);
                tileS[r * SY + q] = (__ldcg(&DR_F(mw, L)[L.idx]) & 0x80) ? __ldcg(&DR_F(occw, L)[L.idx]) : 0u;
            }
        }
    }
    __syncthreads();
    // bitmaps and the lists of occupied cells
    for (int r = warp; r < SX; r += nwarps) {
        const bool row_in = r >= HP && r < HP + ex;
        for (int q0 = 0; q0 < SY; q0 += 32) {
            const int q = q0 + lane;
            const uint32_t ow = q < SY ? tileS[r * SY + q] : 0u;
            const unsigned m = __ballot_sync(FULL, ow != 0u);
            if (!m) continue;
            if (lane == 0) bm[r * BW + (q0 >> 5)] = m;
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_all, __popc(m));
            base = __shfl_sync(FULL, base, 0);
            const bool mine = ow != 0u && row_in && q >= HP && q < HP + ey;
            const unsigned mi = __ballot_sync(FULL, mine);
            int ibase = 0;
            if (mi) { if (lane == 0) ibase = atomicAdd(&s_n, __popc(mi)); ibase = __shfl_sync(FULL, ibase, 0); }
            if (ow != 0u) {
                alist[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(r * SY + q);
                atomicOr(&bmt[q * BWT + (r >> 5)], 1u << (r & 31));
                if (mine) ilist[ibase + __popc(mi & ((1u << lane) - 1u))] = (uint16_t)((r << 7) | q);
            }
        }
    }
    __syncthreads();
    // priorities of the occupied cells (every lane busy)
    for (int k = tid; k < s_all; k += DR_BUILD_THREADS) {
        const int pos = alist[k];
        const uint32_t ow = tileS[pos];
        tileP[pos] = (ss_priority(ba.ok, occ_slot(ow)) << 5) | (uint32_t)occ_vision(ow);
    }
    __syncthreads();
    const int nl = s_n;
    PairT* pairs = reinterpret_cast<PairT*>(alist) + warp * DR_PAIR_CAP;       // (the list of all agents is dead by now)
    uint32_t* cnt_p = s_cnt[warp][0];
    uint32_t* cnt_s = s_cnt[warp][1];
    const int per_warp = min(32, (nl + nwarps - 1) / nwarps);                   // agents of a batch: spread over the warps
    const unsigned below = (1u << lane) - 1u;
    for (int b0 = 0; b0 < nl; b0 += per_warp * nwarps) {
        const int k = b0 + warp * per_warp + lane;
        const bool have = lane < per_warp && k < nl;
        int lr = 0, lq = 0, a = 0;
        uint32_t pa = 0, slot = 0;
        if (have) {
            const int e = ilist[k];
            lr = e >> 7; lq = e & 127;
            const uint32_t me = tileP[lr * SY + lq];
            pa = me >> 5; a = (int)(me & 31u);
            slot = occ_slot(tileS[lr * SY + lq]);
        }
        // the occupied cells of the zone: 2V+1 row windows and the x-arm beyond the square from the column bitmap.  The zone
        // is what the predicate can reach from THIS agent's vision a, whatever the other's (b <= V): rows |dx| <= a see
        // |dy| <= V, rows a < |dx| <= V see |dy| <= a, the axes reach a + V.
        auto radius = [&](int dx) -> int { return dx == 0 ? a + V : (abs(dx) <= a ? V : a); };
        auto row_win = [&](int dx) -> WinT {
            const int R = radius(dx);
            WinT v = dr_window<WIDE>(&bm[(lr + dx) * BW], lq - R, 2 * R + 1);
            if (dx == 0) v &= ~((WinT)1 << R);                                  // not the agent itself
            return v;
        };
        auto col_win = [&]() -> WinT {
            WinT v = dr_window<WIDE>(&bmt[lq * BWT], lr - H, 2 * H + 1);
            v &= ~((((WinT)1 << (2 * V + 1)) - (WinT)1) << (H - V));            // the square's rows are counted by the row windows
            return v & ((((WinT)2 << (2 * (a + V))) - (WinT)1) << (H - (a + V)));      // ... and nothing beyond a + V
        };
        uint32_t cnt = 0;
        if (have) {
            for (int dx = -V; dx <= V; dx++) cnt += (uint32_t)dr_popc<WIDE>(row_win(dx));
            cnt += (uint32_t)dr_popc<WIDE>(col_win());
        }
        uint32_t end = cnt;                                                     // inclusive scan over the batch
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, end, o); if (lane >= o) end += t; }
        cnt_p[lane] = 0u; cnt_s[lane] = 0u;
        uint32_t my_at = 0;
        bool pool_bad = false;
        // sub-batches whose pairs fit the list (nearly always one)
        uint32_t done_pairs = 0;
        int first = 0;
        while (first < 32) {
            // lanes first.. whose pairs end within the capacity
            const unsigned fit = __ballot_sync(FULL, lane >= first && end - done_pairs <= (uint32_t)DR_PAIR_CAP);
            int upto = first;                                                   // first lane that does not fit
            { const unsigned nf = ~fit & ~((1u << first) - 1u); upto = nf ? __ffs(nf) - 1 : 32; }
            if (upto == first) upto = first + 1;                                // (a single agent always fits: zone < capacity)
            const bool in_sub = lane >= first && lane < upto;
            const uint32_t sub_begin = __shfl_sync(FULL, end - cnt, first);
            const uint32_t sub_total = __shfl_sync(FULL, end, upto - 1) - sub_begin;
            if (in_sub && have) {
                uint32_t o = end - cnt - sub_begin;
                for (int dx = -V; dx <= V; dx++) {
                    const int R = radius(dx);
                    WinT bits = row_win(dx);
                    while (bits) {
                        const int j = dr_ffs<WIDE>(bits);
                        bits &= bits - (WinT)1;
                        pairs[o++] = (PairT)(((uint32_t)lane << (2 * PS)) | ((uint32_t)(dx + H) << PS) | (uint32_t)(j - R + H));
                    }
                }
                WinT bits = col_win();
                while (bits) {
                    const int j = dr_ffs<WIDE>(bits);
                    bits &= bits - (WinT)1;
                    pairs[o++] = (PairT)(((uint32_t)lane << (2 * PS)) | ((uint32_t)j << PS) | (uint32_t)H);
                }
            }
            __syncwarp();
            // one pair per lane; pass 1 writes the successors beyond the record into their pool runs (rare)
            const uint32_t geo = (uint32_t)lr | ((uint32_t)lq << 8) | ((uint32_t)a << 16);
            for (int pass = 0; pass < 2; pass++) {
                for (uint32_t i0 = 0; i0 < sub_total; i0 += 32) {
                    const uint32_t i = i0 + (uint32_t)lane;
                    const bool valid = i < sub_total;
                    const uint32_t e = valid ? (uint32_t)pairs[i] : 0u;
                    const int ag = (int)(e >> (2 * PS));
                    const uint32_t g = __shfl_sync(FULL, geo, ag);
                    const uint32_t pag = __shfl_sync(FULL, pa, ag), sag = __shfl_sync(FULL, slot, ag), atg = __shfl_sync(FULL, my_at, ag);
                    bool is_pred = false, is_succ = false;
                    uint32_t ent = 0;
                    if (valid) {
                        const int dx = (int)((e >> PS) & ((1u << PS) - 1u)) - H, dy = (int)(e & ((1u << PS) - 1u)) - H;
                        const int alr = (int)(g & 255u), alq = (int)((g >> 8) & 255u), aa = (int)(g >> 16);
                        const int pos = (alr + dx) * SY + alq + dy;
                        const uint32_t cv = tileP[pos];
                        if (dr_crosses_meet(dx, dy, aa, (int)(cv & 31u))) {
                            if ((cv >> 5) < pag) is_pred = true;
                            else {
                                is_succ = true;
                                ent = occ_slot(tileS[pos]);
                                if (STRIPS) { const int r2 = x0 + alr - HP - S.x0 + dx; ent |= (r2 < 0 ? 1u : (r2 >= S.rows ? 2u : 0u)) << 24; }
                            }
                        }
                    }
                    // lanes of one agent are neighbours in the list: segment masks, no atomics
                    const unsigned peers = __match_any_sync(FULL, valid ? ag : 32 + lane);
                    const unsigned bp = __ballot_sync(FULL, is_pred) & peers, bs = __ballot_sync(FULL, is_succ) & peers;
                    const uint32_t had = cnt_s[ag & 31];
                    __syncwarp();
                    if (is_succ) {
                        const uint32_t idx = had + (uint32_t)__popc(bs & below);
                        if (pass == 0) { if (idx < DR_SUCC_INLINE) reinterpret_cast<uint32_t*>(w.dr.succ)[(size_t)sag * 16 + 1 + idx] = ent; }
                        else if (idx >= DR_SUCC_INLINE) w.dr.pool[atg + idx - DR_SUCC_INLINE] = ent;
                    }
                    if (valid && (peers & below) == 0u) {                       // the segment's first lane keeps the counts
                        cnt_s[ag] = had + (uint32_t)__popc(bs);
                        if (pass == 0) cnt_p[ag] += (uint32_t)__popc(bp);
                    }
                    __syncwarp();
                }
                if (pass == 1) break;
                // successors beyond the record: a run of the pool each, then the second walk
                const uint32_t ns = cnt_s[lane];
                const bool over = in_sub && have && ns > DR_SUCC_INLINE;
                if (!__any_sync(FULL, over)) break;
                if (over) {
                    my_at = atomicAdd(&S.ctl[6], ns - DR_SUCC_INLINE);
                    if (my_at + (ns - DR_SUCC_INLINE) > w.dr.pool_cap) { atomicAdd(&S.ctl[7], 1u); my_at = 0xffffffffu; }
                }
                const bool bad = my_at == 0xffffffffu;
                if (__any_sync(FULL, bad)) { if (bad) { my_at = 0; pool_bad = true; } break; }     // pool exhausted: a fault of the step
                __syncwarp();
                if (in_sub) cnt_s[lane] = 0u;                                   // (counted again by the second walk)
                __syncwarp();
            }
            done_pairs += sub_total;
            first = upto;
            if (!__any_sync(FULL, have && lane >= first)) break;
        }
        __syncwarp();
        bool ready = false;
        if (have) {
            const uint32_t nsucc = pool_bad ? (uint32_t)DR_SUCC_INLINE : cnt_s[lane], preds = cnt_p[lane];
            uint32_t* rec = reinterpret_cast<uint32_t*>(w.dr.succ) + (size_t)slot * 16;
            rec[0] = nsucc;
            if (nsucc > DR_SUCC_INLINE) rec[15] = my_at;
            uint32_t indeg = preds;
            if (w.cfg.rule_can_starve) {
                if (dr_q1_pred_at_risk(w, ba.ok, pa)) indeg++;
                if ((__ldg(&w.dr.atrisk[slot >> 5]) >> (slot & 31)) & 1u) {
                    const uint32_t nxt = dr_q1_successor(w, ba.ok, pa);
                    w.dr.q1wait[slot] = make_uint2(nxt ? nxt - 1u : 0u, nxt ? ba.stamp : 0u);
                }
            }
            reinterpret_cast<uint16_t*>(S.dep)[slot] = (uint16_t)indeg;
            ready = indeg == 0;
        }
        // agents without predecessors enter the queue: one reservation per warp and batch (the other warps of the CTA
        // work on while it is in flight)
        const unsigned rm = __ballot_sync(FULL, ready);
        if (rm) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&S.ctl[DR_CTL_TAIL], (unsigned)__popc(rm));
            base = __shfl_sync(FULL, base, 0);
            if (ready) S.queue[base + (unsigned)__popc(rm & below)] = slot | DR_VALID;
        }
        __syncwarp();
    }
    if (tid == 0 && nl) atomicAdd(&S.ctl[DR_CTL_AGENTS], (unsigned)nl);
}

// ------------------------------------------------------------------------------- rounds ----
struct DrAccum { unsigned sum_m, sum_v, acted, deaths, small, inexact, fault; };
// work-list mode: what orders a turn's stores ahead of the counters it decrements and of the entries it publishes --
// device scope on one GPU, system scope when counters, queues and cells of other strips (GPUs) are involved
template <bool STRIPS> __device__ static inline void dr_fence() { if (STRIPS) __threadfence_system(); else __threadfence(); }
// the agent's keyed words in sequence: word j = mix64(akey + j * golden) >> 32 (ss_device.cuh), the sum kept running
struct DrKeyed { uint64_t x; uint32_t j; };
__device__ static inline void dr_keyed_open(DrKeyed& ks, uint64_t skey, uint32_t slot) { ks.x = ss_agent_key(skey, slot + 1u); ks.j = 0u; }
__device__ static inline uint32_t dr_keyed_next(DrKeyed& ks) {
    const uint64_t z = ks.x;
    ks.x += SS_GOLDEN; ks.j++;
    return (uint32_t)(ss_mix64(z) >> 32);
}
__device__ static inline uint32_t dr_randbelow(DrKeyed& ks, uint32_t nn, DrAccum& acc) {      // random.py:242 _randbelow
    const int kbits = 32 - __clz(nn);
    uint32_t r = dr_keyed_next(ks) >> (32 - kbits);
    while (r >= nn) {
        r = dr_keyed_next(ks) >> (32 - kbits);
        if (ks.j > 100000u) { acc.fault++; r = 0u; break; }
    }
    return r;
}
#define DR_WARPS (DR_THREADS / 32)
struct DrShared {
    unsigned hist[DR_HIST_BINS];
    uint32_t push[DR_WARPS][DR_WARP_CAP];
    int cnt[DR_WARPS];
    unsigned long long acc[8];
};

// appends to the warp's staging buffer (flushed by the warp between agents); a full buffer goes straight to the queue
__device__ static inline void dr_push(const DevWorld& w, DrShared& sh, int warp, uint32_t e) {
    const int pos = atomicAdd(&sh.cnt[warp], 1);
    if (pos < DR_WARP_CAP) sh.push[warp][pos] = e;
    else { __threadfence(); w.dr.self.queue[atomicAdd(&w.dr.self.ctl[0], 1u)] = e; }      // (fence: see the work-list mode)
}
__device__ static inline void dr_flush(const DevWorld& w, DrShared& sh, int warp, int lane) {
    __syncwarp();
    const int k = min(sh.cnt[warp], DR_WARP_CAP);
    if (k > 0) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&w.dr.self.ctl[0], (unsigned)k);
        base = __shfl_sync(FULL, base, 0);
        for (int j = lane; j < k; j += 32) w.dr.self.queue[base + (unsigned)j] = sh.push[warp][j];
    }
    __syncwarp();
    if (lane == 0) sh.cnt[warp] = 0;
    __syncwarp();
}
// the result of one decrement of a successor's counter: the counter's old value decides who queues the successor
template <bool STRIPS, bool ASYNC>
__device__ static inline void dr_released(const DevWorld& w, DrShared& sh, int warp, DrAccum& acc, uint32_t ent, uint32_t old,
                                          bool lose_turn) {
    const uint32_t slot2 = ent & OCC_SLOT_MASK;
    const uint32_t h = (old >> ((slot2 & 1u) << 4)) & 0xffffu;
    if ((h & 0x7fffu) == 1u) {
        const uint32_t e = slot2 | DR_VALID | ((lose_turn || (h >> 15)) ? DR_SKIP : 0u);
        const int which = STRIPS ? (int)((ent >> 24) & 3u) : 0;
        if (which == 0) dr_push(w, sh, warp, e);
        else {                                                                                  // a ring neighbour's agent
            if (ASYNC) dr_fence<STRIPS>();              // (the counter's other decrements, then the entry: see dr_flush)
            DR_W(queue, which)[atomicAdd(&DR_W(ctl, which)[0], 1u)] = e | DR_REMOTE;
        }
    } else if ((h & 0x7fffu) == 0u) acc.fault++;
}
template <bool STRIPS>
__device__ static inline uint32_t dr_dec(const DevWorld& w, uint32_t ent) {
    const uint32_t slot2 = ent & OCC_SLOT_MASK;
    const int which = STRIPS ? (int)((ent >> 24) & 3u) : 0;
    return atomicSub(&DR_W(dep, which)[slot2 >> 1], 1u << ((slot2 & 1u) << 4));
}

struct DrStep {
    int64_t iteration, env_time;
    uint64_t skey;
    uint32_t next_phase;
    int single_round;               // host-driven rounds (several strips on one GPU): run one round and return
    unsigned seg_begin, seg_end;    // ... over this queue segment
    unsigned nblocks;
};

// 17 consecutive cell bytes (offsets -a..+a of one arm, a <= 8) out of two aligned 16-byte loads
struct ArmWin { unsigned long long w0, w1, w2; };
__device__ static inline ArmWin dr_load_arm(const uint8_t* base, size_t first) {
    const size_t al = first & ~(size_t)15;
    const int off = (int)(first - al);
    const uint4 A = __ldcg(reinterpret_cast<const uint4*>(base + al)), B = __ldcg(reinterpret_cast<const uint4*>(base + al + 16));
    unsigned long long q0 = (unsigned long long)A.x | ((unsigned long long)A.y << 32), q1 = (unsigned long long)A.z | ((unsigned long long)A.w << 32);
    unsigned long long q2 = (unsigned long long)B.x | ((unsigned long long)B.y << 32), q3 = (unsigned long long)B.z | ((unsigned long long)B.w << 32);
    if (off & 8) { q0 = q1; q1 = q2; q2 = q3; q3 = 0ull; }
    const int r = (off & 7) * 8;
    ArmWin W;
    if (r) { W.w0 = (q0 >> r) | (q1 << (64 - r)); W.w1 = (q1 >> r) | (q2 << (64 - r)); W.w2 = (q2 >> r) | (q3 << (64 - r)); }
    else { W.w0 = q0; W.w1 = q1; W.w2 = q2; }
    return W;
}
__device__ static inline unsigned dr_arm_byte(const ArmWin& W, int i) {
    return (unsigned)((i < 8 ? W.w0 >> (8 * i) : (i < 16 ? W.w1 >> (8 * (i - 8)) : W.w2)) & 0xffull);
}

// agent.py:361-371 getFood + 794-823 move(): one THREAD evaluates the cross.  cellbyte(arm, o) = the cell word at offset o
// on the x-arm (arm 0) / y-arm (arm 1); cellpoll likewise the pollution.  Candidates are numbered k = offset + a on the
// x-arm (ascending raw x), then span + offset + a on the y-arm -- the order getFood appends them in; best = (key desc, raw
// Manhattan distance asc, position in the shuffled food list asc).  At most four candidates tie on (key, distance): one
// distance, two arms, two sides.  Returns the chosen offset (arm, o) and harvest; arm = -1: stay.
template <int POLT, class CellByte, class CellPoll>
__device__ static inline void dr_choose(const DevWorld& w, int gx, int gy, int a, uint32_t slot, uint64_t skey, int sg0, double pown,
                                        DrAccum& acc, CellByte cellbyte, CellPoll cellpoll, int& out_arm, int& out_o, int& out_sg) {
    const int n = w.n;
    constexpr bool pol = POLT != 0;
    const int span = 2 * a + 1;
    unsigned long long fm = 0ull;           // free candidates, bit = candidate index k
    unsigned long long tl = 0ull;           // tied best candidates, 16 bits each: harvest << 8 | k
    int nt = 0, dmin = 0x7fffffff;
    uint32_t ikmax = 0u;
    double dkmax = -1.0;
#pragma unroll 1
    for (int d = 1; d <= a; d++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int arm = u >> 1, o = (u & 1) ? d : -d;
            const unsigned mv = cellbyte(arm, o);
            if (mv & 0x80u) continue;
            const int sg = mv == DR_HARVESTED ? 0 : (int)mv;
            const int kk = arm * span + a + o;
            fm |= 1ull << kk;
            // agent.py:867-868: raw coordinates, so a candidate reached across the torus edge is far away
            const int cc = (arm ? gy : gx) + o;
            const int dist = (cc < 0 || cc >= n) ? n - d : d;
            const unsigned long long ent = ((unsigned long long)sg << 8) | (unsigned long long)kk;
            if (!pol) {
                const uint32_t ik = (((uint32_t)sg << 20) | (0xfffffu - (uint32_t)dist)) + 1u;
                if (ik > ikmax) { ikmax = ik; tl = ent; nt = 1; }
                else if (ik == ikmax) { tl = (tl << 16) | ent; nt++; }
            } else {
                const double key = (double)sg / (1.0 + (sg > 0 ? cellpoll(arm, o) : 0.0));
                if (key > dkmax || (key == dkmax && dist < dmin)) { dkmax = key; dmin = dist; tl = ent; nt = 1; }
                else if (key == dkmax && dist == dmin) { tl = (tl << 16) | ent; nt++; }
            }
        }
    }
    out_arm = -1; out_o = 0; out_sg = sg0;
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
        DrKeyed ks;
        dr_keyed_open(ks, skey, slot);
        int left = nt;
#pragma unroll 1
        for (int i = n_free - 1; i >= 1 && left > 1; i--) {                         // random.py:350 shuffle
            const uint32_t rj = dr_randbelow(ks, (uint32_t)i + 1u, acc);
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
    out_arm = k >= span ? 1 : 0;
    out_o = k - out_arm * span - a;
    out_sg = (int)(ent >> 8);
}

// ------------------------------------------------------------------------------------------
// agent.py:1025-1141 utilicalcMove(), sugar-only (utilicalcSugar agent.py:980-992): one THREAD takes the whole turn.
// The welfare evaluation is fused into the movement scan: the two arms are scanned once for free cells (the moves,
// agent.py:361-371 getFood) and occupants (agent.py:391-407 getVisionNeighbourhood), the three Fisher-Yates shuffles are
// replayed on the agent's keyed words (the sums below are compared for EQUALITY, agent.py:1105, so the neighbours are
// summed in their shuffled order, and random.choice picks among the maxima in the moves' shuffled order), and a move's
// utility is one f64 that never leaves the thread.  A (move, neighbour) term counts only when the neighbour would see
// the agent there (`cert`); the other terms are +-0.0 in the reference and leave the sum as it is, so they are skipped.
// Candidates are numbered as in dr_choose: k = offset + a on the x-arm, span + offset + a on the y-arm; 255 = own cell.
#define DR_UT_MAX (4 * 15 + 2)
#define DR_UT_OWN 255
__device__ static inline double dr_util_intensity(double sugar, int metab) {        // agent.py:870-872, 984
    const double days = (sugar >= 0.0 && sugar < 2147483647.0 && sugar == (double)(int)sugar)
                            ? (double)((int)sugar / metab) : ss_py_floordiv(sugar, (double)metab);
    return 1.0 / (1.0 + days);
}
// (site / metabolism) / maxCapacity of utilicalcSugar (agent.py:985-986) for the values that occur: cell sugar < 128,
// metabolism 1..DR_DUR_M -- two f64 divisions per (move, neighbour) term otherwise.  Filled per CTA with the same two
// IEEE divisions (ss_dr_rounds_kernel).
#define DR_DUR_M 8
__device__ static inline double dr_duration(const DevWorld& w, const double* durtab, int site, int metab) {
    if (metab >= 1 && metab <= DR_DUR_M) return durtab[(metab - 1) * 128 + site];
    return ((double)site / (double)metab) / w.cfg.max_capacity;
}
// a cell of the cross as the utilities see it: raw coordinates (agent.py:109-110 getDistance works on Agent.x / .y)
__device__ static inline void dr_cross_xy(int gx, int gy, int n, int arm, int o, int& x, int& y) {
    x = arm ? gx : ss_wrap1(gx + o, n); y = arm ? ss_wrap1(gy + o, n) : gy;
}
// NOWRAP: both arms lie inside [0, n) (raw coordinates = wrapped coordinates): which candidate cells a neighbour would
// see the agent on follows from offsets alone.
template <class CellByte, class CellWord, class CellRec>
__device__ static inline void dr_util_choose(const DevWorld& w, const double* durtab, bool nowrap, int gx, int gy, int a, uint32_t slot,
                                             uint64_t skey, unsigned own_byte, double own_sugar, int own_metab, DrAccum& acc,
                                             CellByte cellbyte, CellWord cellword, CellRec cellrec, int& out_arm, int& out_o, int& out_sg) {
    const int n = w.n, span = 2 * a + 1;
    uint16_t mv[DR_UT_MAX];             // moves: candidate number | int(sugar) << 8, own cell last
    uint8_t nk[DR_UT_MAX];              // neighbours: candidate numbers of their cells
    unsigned long long seen[DR_UT_MAX]; // ... in their shuffled order, own last: the candidates they would see the agent on (bit 63: own cell),
    uint32_t nvm[DR_UT_MAX];            //     metabolism << 8 | self << 31,
    double ni[DR_UT_MAX];               //     1 / (1 + days to starvation)
    double ut[DR_UT_MAX];
    int m = 0, nb = 0;
#pragma unroll 1
    for (int k = 0; k < 2 * span; k++) {
        const int arm = k >= span ? 1 : 0, o = k - arm * span - a;
        if (o == 0) continue;                                   // own cell: occupied by the agent, and not a neighbour
        const unsigned b = cellbyte(arm, o);
        if (b & 0x80u) nk[nb++] = (uint8_t)k;
        else mv[m++] = (uint16_t)((unsigned)k | (((b & 0x7fu) == DR_HARVESTED ? 0u : (b & 0x7fu)) << 8));
    }
    DrKeyed ks;
    dr_keyed_open(ks, skey, slot);
#pragma unroll 1
    for (int rep = 0; rep < 2; rep++)                           // getFood's shuffle (agent.py:369), then agent.py:1033
#pragma unroll 1
        for (int i = m - 1; i >= 1; i--) {                      // random.py:350
            const int jj = (int)dr_randbelow(ks, (uint32_t)i + 1u, acc);
            const uint16_t t = mv[i]; mv[i] = mv[jj]; mv[jj] = t;
        }
    mv[m++] = (uint16_t)(DR_UT_OWN | (((own_byte & 0x7fu) == DR_HARVESTED ? 0u : (own_byte & 0x7fu)) << 8));        // agent.py:1035
#pragma unroll 1
    for (int i = nb - 1; i >= 1; i--) {                         // agent.py:405
        const int jj = (int)dr_randbelow(ks, (uint32_t)i + 1u, acc);
        const uint8_t t = nk[i]; nk[i] = nk[jj]; nk[jj] = t;
    }
    // the neighbours' records in that order, four in flight: occupancy word -> slot -> record
#pragma unroll 1
    for (int q0 = 0; q0 < nb; q0 += 4) {
        uint32_t wd[4];
        uint4 rc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            wd[u] = 0u;
            if (q0 + u < nb) { const int k = nk[q0 + u], arm = k >= span ? 1 : 0; wd[u] = cellword(arm, k - arm * span - a); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            rc[u] = make_uint4(0u, 0u, 0u, 0u);
            if (q0 + u < nb) { const int k = nk[q0 + u], arm = k >= span ? 1 : 0; rc[u] = cellrec(arm, k - arm * span - a, occ_slot(wd[u])); }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (q0 + u >= nb) continue;
            if (wd[u] == 0u || !attr_alive(rc[u].w)) acc.fault++;
            const double bs = __longlong_as_double((long long)((unsigned long long)rc[u].x | ((unsigned long long)rc[u].y << 32)));
            const int bm = attr_metab(rc[u].w), bv = attr_vision(rc[u].w), k = nk[q0 + u], arm = k >= span ? 1 : 0, ob = k - arm * span - a;
            ni[q0 + u] = dr_util_intensity(bs, bm);
            nvm[q0 + u] = (uint32_t)bm << 8;
            // certainty (agent.py:982, 109-110): the neighbour sees a cell iff it shares a raw coordinate with it and the
            // other one differs by at most its vision
            unsigned long long sm = 0ull;
            if (nowrap) {
                // same arm, offsets within its vision; the own cell likewise; never a cell of the other arm
                const int lo = max(-a, ob - bv), hi = min(a, ob + bv);
                sm = ((2ull << (hi - lo)) - 1ull) << (arm * span + a + lo);
                if (abs(ob) <= bv) sm |= 1ull << 63;
            } else {
                int bx, by;
                dr_cross_xy(gx, gy, n, arm, ob, bx, by);
#pragma unroll 1
                for (int k2 = 0; k2 < 2 * span; k2++) {
                    const int arm2 = k2 >= span ? 1 : 0, o2 = k2 - arm2 * span - a;
                    int mx, my;
                    dr_cross_xy(gx, gy, n, arm2, o2, mx, my);
                    if ((bx == mx && abs(by - my) <= bv) || (by == my && abs(bx - mx) <= bv)) sm |= 1ull << k2;
                }
                if ((bx == gx && abs(by - gy) <= bv) || (by == gy && abs(bx - gx) <= bv)) sm |= 1ull << 63;
            }
            seen[q0 + u] = sm;
        }
    }
    const double extent = (double)nb;
    ni[nb] = dr_util_intensity(own_sugar, own_metab);           // agent.py:1039: the agent itself, last
    nvm[nb] = ((uint32_t)own_metab << 8) | 0x80000000u;
    {
        unsigned long long sm = ~0ull;                          // every candidate lies within the agent's own vision ...
        if (!nowrap) {                                          // ... in raw coordinates unless it was reached across the torus edge
            sm = 1ull << 63;
#pragma unroll 1
            for (int k2 = 0; k2 < 2 * span; k2++) {
                const int arm2 = k2 >= span ? 1 : 0, o2 = k2 - arm2 * span - a;
                int mx, my;
                dr_cross_xy(gx, gy, n, arm2, o2, mx, my);
                if ((gx == mx && abs(gy - my) <= a) || (gy == my && abs(gx - mx) <= a)) sm |= 1ull << k2;
            }
        }
        seen[nb] = sm;
    }
    nb++;
    double umax = 0.0;
#pragma unroll 1
    for (int p = 0; p < m; p++) {
        const unsigned e = mv[p];
        const int bit = (e & 255u) == DR_UT_OWN ? 63 : (int)(e & 255u), site = (int)(e >> 8);
        double u = 0.0;
#pragma unroll 1
        for (int q = 0; q < nb; q++) {
            // branch-free: the lanes of a warp are at different (move, neighbour) pairs of different agents.  A neighbour
            // that would not see the agent there contributes +-0.0 in the reference: the sum stays as it is.
            const uint32_t vm = nvm[q];
            const bool cert = ((seen[q] >> bit) & 1ull) != 0ull;
            const double dur = dr_duration(w, durtab, site, (int)((vm >> 8) & 0xffffu));
            const double t = ((ni[q] + dur) + 0.5 * 0.0) + extent;
            u += cert ? ((vm >> 31) ? t : -t) : 0.0;
        }
        ut[p] = u;
        if (p == 0 || u > umax) umax = u;
    }
    int nmax = 0;
    for (int p = 0; p < m; p++) nmax += ut[p] == umax ? 1 : 0;
    int pick = (int)dr_randbelow(ks, (uint32_t)nmax, acc);     // random.choice, agent.py:1108
    unsigned e = mv[m - 1];
    for (int p = 0; p < m; p++)
        if (ut[p] == umax && pick-- == 0) { e = mv[p]; break; }
    const int k = (int)(e & 255u);
    out_arm = -1; out_o = 0;
    if (k != DR_UT_OWN) { out_arm = k >= span ? 1 : 0; out_o = k - out_arm * span - a; }
    out_sg = (int)(e >> 8);                                     // agent.py:1128-1130
}

template <int POLT, bool STRIPS, bool UTIL, bool ASYNC>
__device__ static inline void dr_turn(const DevWorld& w, const DrStep& st, DrShared& sh, const double* durtab, int warp, DrAccum& acc, uint32_t entry) {
    const int n = w.n;
    const DrPeer& S = w.dr.self;
    const uint32_t slot = entry & OCC_SLOT_MASK;
    const bool skip = (entry & DR_SKIP) != 0u;
    // the agent's record and its successor list
    AgentRec* recp = &S.agents[slot];
    const uint4 r0 = __ldcg(reinterpret_cast<const uint4*>(recp));               // sugar, pos, attr
    const uint4* sp = &w.dr.succ[(size_t)slot * 4];
    uint4 sq[4];
    sq[0] = __ldcg(&sp[0]);
    double sugar = __longlong_as_double((long long)((unsigned long long)r0.x | ((unsigned long long)r0.y << 32)));
    const uint32_t c = r0.z;
    uint32_t attr = r0.w;
    if (!attr_alive(attr)) { acc.fault++; return; }
    const int a = attr_vision(attr);
    const bool risky = (attr & ATTR_RISK) != 0u;
    const uint32_t nsucc = sq[0].x;
#pragma unroll
    for (int i = 1; i < 4; i++) sq[i] = (nsucc + 1u > 4u * (uint32_t)i) ? __ldcg(&sp[i]) : make_uint4(0u, 0u, 0u, 0u);
    uint2 qw = make_uint2(0u, 0u);
    if (risky) qw = __ldcg(&w.dr.q1wait[slot]);
    const int lx = (int)(c / (uint32_t)n), gy = (int)(c - (uint32_t)lx * (uint32_t)n), gx = S.x0 + lx;
    bool died = false, wrote_far = false;
    if (!skip) {
        int arm = -1, o = 0, sg = 0;
        uint8_t own;
        // both arms inside this strip and the torus edge out of sight: two wide loads per arm
        const bool fast = a <= 8 && lx - a >= 0 && lx + a < S.rows && gy - a >= 0 && gy + a < n;
        if (UTIL || w.cfg.rule_move_eat) {
            if (fast) {
                const ArmWin wy = dr_load_arm(S.mw, (size_t)c - (size_t)a);
                const ArmWin wx = dr_load_arm(S.mwx, (size_t)gy * (size_t)S.rows + (size_t)(lx - a));
                own = (uint8_t)dr_arm_byte(wy, a);
                double pown = 0.0;
                if (POLT) pown = __ldcg(&S.poll[c]);
                if (UTIL)
                    dr_util_choose(w, durtab, true, gx, gy, a, slot, st.skey, own, sugar, attr_metab(attr), acc,
                                   [&](int am, int oo) { return dr_arm_byte(am ? wy : wx, a + oo); },
                                   [&](int am, int oo) { return __ldcg(&S.occw[am ? c + oo : (uint32_t)((int)c + oo * n)]); },
                                   [&](int, int, uint32_t s2) { return __ldcg(reinterpret_cast<const uint4*>(&S.agents[s2])); }, arm, o, sg);
                else
                dr_choose<POLT>(w, gx, gy, a, slot, st.skey, (own & 0x7f) == DR_HARVESTED ? 0 : (own & 0x7f), pown, acc,
                                [&](int am, int oo) { return dr_arm_byte(am ? wy : wx, a + oo); },
                                [&](int am, int oo) { return __ldcg(&S.poll[am ? c + oo : (uint32_t)((int)c + oo * n)]); }, arm, o, sg);
            } else {
                // near the torus edge or a strip border, or vision > 8: cell by cell through the strip addressing --
                // every load is issued before the first is looked at (a neighbour strip's rows are an NVLink round trip away)
                own = __ldcg(&S.mw[c]);
                double pown = 0.0;
                if (POLT) pown = __ldcg(&S.poll[c]);
                uint8_t cb[2][2 * 15 + 1];
#pragma unroll 1
                for (int oo = -a; oo <= a; oo++) {
                    if (oo == 0) continue;
                    const DrLoc Lx = dr_loc<STRIPS>(w, ss_wrap1(gx + oo, n), gy), Ly = dr_loc<STRIPS>(w, gx, ss_wrap1(gy + oo, n));
                    cb[0][oo + a] = __ldcg(&DR_F(mw, Lx)[Lx.idx]);
                    cb[1][oo + a] = __ldcg(&DR_F(mw, Ly)[Ly.idx]);
                }
                if (UTIL)
                    dr_util_choose(w, durtab, gx - a >= 0 && gx + a < n && gy - a >= 0 && gy + a < n, gx, gy, a, slot, st.skey, own, sugar,
                                   attr_metab(attr), acc,
                                   [&](int am, int oo) { return (unsigned)cb[am][oo + a]; },
                                   [&](int am, int oo) {
                                       const DrLoc L = dr_loc<STRIPS>(w, am ? gx : ss_wrap1(gx + oo, n), am ? ss_wrap1(gy + oo, n) : gy);
                                       return __ldcg(&DR_F(occw, L)[L.idx]);
                                   },
                                   [&](int am, int oo, uint32_t s2) {
                                       const DrLoc L = dr_loc<STRIPS>(w, am ? gx : ss_wrap1(gx + oo, n), am ? ss_wrap1(gy + oo, n) : gy);
                                       return __ldcg(reinterpret_cast<const uint4*>(&DR_F(agents, L)[s2]));
                                   }, arm, o, sg);
                else
                dr_choose<POLT>(w, gx, gy, a, slot, st.skey, (own & 0x7f) == DR_HARVESTED ? 0 : (own & 0x7f), pown, acc,
                                [&](int am, int oo) { return (unsigned)cb[am][oo + a]; },
                                [&](int am, int oo) {
                                    const DrLoc L = dr_loc<STRIPS>(w, am ? gx : ss_wrap1(gx + oo, n), am ? ss_wrap1(gy + oo, n) : gy);
                                    return __ldcg(&DR_F(poll, L)[L.idx]);
                                }, arm, o, sg);
            }
        } else own = __ldcg(&S.mw[c]);
        const int nx = arm == 0 ? ss_wrap1(gx + o, n) : gx, ny = arm == 1 ? ss_wrap1(gy + o, n) : gy;
        const DrLoc D = dr_loc<STRIPS>(w, nx, ny);
        const bool moved = arm >= 0;
        const bool pol = POLT != 0 && w.cfg.pollution_start_time <= st.env_time;
        const int metab = attr_metab(attr);
        double p = 0.0;
        if (pol) p = __ldcg(&DR_F(poll, D)[D.idx]);
        if (UTIL || w.cfg.rule_move_eat) {
            if (pol) p = p + (((double)sg * w.cfg.pollution_a) + (0.0 * w.cfg.pollution_b));     // agent.py:825-826
            sugar = ss_set_sugar(ss_max0(sugar + (double)sg));                                     // agent.py:832 / 1128-1130
            // agent.py:864 cell.sugar = 0: the cell's byte takes the marker "harvested" below; the f64 value is
            // dropped by the next growback (or ss_dr_canon_kernel), which saves a scattered 8-byte store per turn
        }
        if (pol && sugar > 0.0) p = p + ((0.0 * w.cfg.pollution_a) + ((double)metab * w.cfg.pollution_b));   // sugarscape.py:566-567
        if (pol) DR_F(poll, D)[D.idx] = p;
        sugar = ss_set_sugar(ss_max0(sugar - (double)metab));                                      // sugarscape.py:569
        acc.sum_m += (unsigned)metab; acc.sum_v += (unsigned)a; acc.acted++;
        if (sugar == 0.01) acc.small++;
        else if (sugar >= 0.0 && sugar < (double)SS_GINI_BINS && sugar == (double)(int)sugar) {
            const int bin = (int)sugar;
            if (bin < DR_HIST_BINS) atomicAdd(&sh.hist[bin], 1u); else atomicAdd(&w.hist[bin], 1u);
        } else { acc.inexact++; ss_note_outlier(w, sugar); }
        died = w.cfg.rule_can_starve && sugar <= 0.1;                                              // sugarscape.py:306-309
        const bool risk_next = !died && dr_at_risk(w, sugar, metab);
        const uint32_t new_word = died ? 0u : occ_pack(slot, a, st.next_phase);
        const uint8_t new_mw = (uint8_t)((died ? 0u : 0x80u) | ((UTIL || w.cfg.rule_move_eat) ? DR_HARVESTED : (own & 0x7fu)));
        if (moved) {
            // (the cell's occupancy word stays behind: a word counts only where the byte says "occupied")
            S.mw[c] = own & 0x7f;
            S.mwx[(size_t)gy * (size_t)S.rows + (size_t)lx] = own & 0x7f;
        }
        DR_F(occw, D)[D.idx] = new_word;
        DR_F(mw, D)[D.idx] = new_mw;
        DR_F(mwx, D)[dr_xidx<STRIPS>(w, D)] = new_mw;
        attr = (attr & ~ATTR_RISK) | (risk_next ? ATTR_RISK : 0u);
        if (died) { attr &= ~ATTR_ALIVE; acc.deaths++; }
        if (!STRIPS) {
            if (died) atomicAnd(&w.dr.alive[slot >> 5], ~(1u << (slot & 31)));
            if (risk_next != risky) atomicXor(&w.dr.atrisk[slot >> 5], 1u << (slot & 31));
        } else {
            // the directories are replicated on every strip: what changed goes to this strip's delta list
            const uint32_t kind = (died ? DR_DELTA_DIED : 0u) | (risk_next != risky ? DR_DELTA_RISK : 0u) |
                                  (D.which == 1 ? DR_DELTA_TO_LO : 0u) | (D.which == 2 ? DR_DELTA_TO_HI : 0u);
            if (kind) {
                const unsigned at = atomicAdd(&S.ctl[DR_CTL_DELTA], 1u);
                if (at < w.dr.delta_cap) S.delta[at] = slot | (kind << 24); else acc.fault++;
            }
        }
        AgentRec* dst = &DR_F(agents, D)[slot];                    // the record travels with the agent
        uint4 nr;
        const unsigned long long sb = (unsigned long long)__double_as_longlong(sugar);
        nr.x = (uint32_t)sb; nr.y = (uint32_t)(sb >> 32); nr.z = D.idx; nr.w = attr;
        *reinterpret_cast<uint4*>(dst) = nr;
        if (STRIPS && D.which != 0) { recp->attr = attr & ~ATTR_ALIVE; wrote_far = true; }      // left this strip
    }
    // (no round barrier orders this turn's stores before the turns it releases: a fence does, once, ahead of every counter
    // it decrements; whoever brings a counter to zero fences again before it publishes the queue entry, dr_flush)
    if (ASYNC) {
        // (strips: system scope only for a turn that wrote to or notifies another strip -- a move across the border, a
        // successor or the Q1 successor over there, a successor list that continues in the pool; the interior stays at
        // device scope)
        bool far = false;
        if (STRIPS) {
            const uint32_t ev[16] = {sq[0].x, sq[0].y, sq[0].z, sq[0].w, sq[1].x, sq[1].y, sq[1].z, sq[1].w,
                                     sq[2].x, sq[2].y, sq[2].z, sq[2].w, sq[3].x, sq[3].y, sq[3].z, sq[3].w};
            uint32_t strips_or = 0u;
#pragma unroll
            for (int i = 0; i < DR_SUCC_INLINE; i++) if ((uint32_t)i < nsucc) strips_or |= ev[1 + i];
            far = wrote_far || (strips_or & 0x03000000u) != 0u || nsucc > DR_SUCC_INLINE ||
                  (risky && qw.y == (uint32_t)(st.iteration + 1));
        }
        if (far) __threadfence_system(); else __threadfence();
    }
    // Q1: whoever follows this agent in the list order learns the outcome (a skipped agent did not die)
    if (risky && qw.y == (uint32_t)(st.iteration + 1)) {
        const uint32_t s2 = qw.x & OCC_SLOT_MASK;
        const int home = STRIPS ? (int)w.dr.home[s2] : 0;              // the successor may stand anywhere on the lattice
        const int shf = (int)(s2 & 1u) << 4;
        uint32_t* wp = &DR_A(dep, home)[s2 >> 1];
        const uint32_t old = died ? atomicAdd(wp, 0x7fffu << shf) : atomicSub(wp, 1u << shf);
        const uint32_t h = (old >> shf) & 0xffffu;
        if ((h & 0x7fffu) == 1u) {
            const uint32_t e = s2 | DR_VALID | ((died || (h >> 15)) ? DR_SKIP : 0u);
            if (!STRIPS || home == w.dr.rank) dr_push(w, sh, warp, e);
            else {
                if (ASYNC) dr_fence<STRIPS>();
                DR_A(queue, home)[atomicAdd(&DR_A(ctl, home)[DR_CTL_TAIL], 1u)] = e | DR_REMOTE;
            }
        } else if ((h & 0x7fffu) == 0u) acc.fault++;
    }
    // conflict successors: all decrements of the record are issued before the first result is looked at
    {
        const uint32_t e[16] = {sq[0].x, sq[0].y, sq[0].z, sq[0].w, sq[1].x, sq[1].y, sq[1].z, sq[1].w,
                                sq[2].x, sq[2].y, sq[2].z, sq[2].w, sq[3].x, sq[3].y, sq[3].z, sq[3].w};
        uint32_t old[DR_SUCC_INLINE];
#pragma unroll
        for (int i = 0; i < DR_SUCC_INLINE; i++) old[i] = (uint32_t)i < nsucc ? dr_dec<STRIPS>(w, e[1 + i]) : 0u;
#pragma unroll
        for (int i = 0; i < DR_SUCC_INLINE; i++)
            if ((uint32_t)i < nsucc) dr_released<STRIPS, ASYNC>(w, sh, warp, acc, e[1 + i], old[i], false);
        if (nsucc > DR_SUCC_INLINE) {
            const uint32_t* pp = w.dr.pool + e[15];
            const uint32_t extra = nsucc - DR_SUCC_INLINE;
            for (uint32_t b = 0; b < extra; b += 4) {
                uint32_t en[4], ol[4];
#pragma unroll
                for (int i = 0; i < 4; i++) en[i] = b + i < extra ? __ldcg(&pp[b + i]) : 0u;
#pragma unroll
                for (int i = 0; i < 4; i++) ol[i] = b + i < extra ? dr_dec<STRIPS>(w, en[i]) : 0u;
#pragma unroll
                for (int i = 0; i < 4; i++) if (b + i < extra) dr_released<STRIPS, ASYNC>(w, sh, warp, acc, en[i], ol[i], false);
            }
        }
    }
}

// Barrier over the strips (one GPU each) through peer memory: every strip publishes the next sequence number into
// every strip's flag words (stores over NVLink) and waits until its own words show that number from everybody.
// One thread; the caller has made this strip's writes visible to it (grid barrier / kernel boundary).
#define DR_CTL_SEQ 12                   // (10, 11: the two "more work" words of the round barrier)
#define DR_CTL_HEAD 13                  // work-list mode: queue positions handed out so far
__device__ static inline bool dr_strip_sync(const DevWorld& w, unsigned seq, int word, unsigned payload, unsigned* any_payload) {
    const int P = w.dr.world, me = w.dr.rank;
    __threadfence_system();
    for (int p = 0; p < P; p++) *(volatile unsigned*)&w.dr.all[p].ctl[DR_CTL_FLAGS + 2 * me + word] = (seq << 1) | (payload & 1u);
    unsigned any = 0;
    bool ok = true;
    for (int p = 0; p < P; p++) {
        unsigned v, spins = 0;
        while (((v = *(volatile unsigned*)&w.dr.self.ctl[DR_CTL_FLAGS + 2 * p + word]) >> 1) < seq) {
            __nanosleep(20);
            if (++spins > (1u << 26)) { ok = false; break; }           // a lost peer: do not hang the GPU
        }
        any |= v & 1u;
    }
    __threadfence_system();
    if (any_payload) *any_payload = any;
    return ok;
}
__global__ void ss_dr_strip_barrier_kernel(DevWorld w) {
    unsigned int* ctl = w.dr.self.ctl;
    const unsigned seq = ++ctl[DR_CTL_SEQ];
    if (!dr_strip_sync(w, seq, 0, 0u, nullptr)) atomicAdd(&w.acc->fault, 1ull);
}

// Grid barrier between rounds (cooperative launch: all CTAs are resident).  The last CTA to arrive snapshots the queue
// tail into ctl[3 + (generation & 1)] before it releases the others, so every CTA sees the same next segment even if a
// fast CTA is already appending to the queue again.  With strips on several GPUs (XSYNC) the same CTA first meets the
// other strips: one exchange per round, "this strip is through with round k, and it had agents in it".  Once every strip
// has said so, whatever the round queued anywhere is in place (the exchange fences at system scope; the CTAs' own
// fences order their stores before their arrival), and the agent phase ends after the first round in which no strip had
// any agent -- nothing can have been queued then.  (Measured alternative: strips running their rounds without a common
// barrier, waiting only for the entries they miss -- 64 instead of 37 rounds per step and slower; DESIGN.md.)
// Returns the snapshot; *more = the agent phase goes on.
template <bool XSYNC>
__device__ static inline unsigned dr_grid_barrier(const DevWorld& w, unsigned nblocks, unsigned seg_end, bool had_work, DrAccum& acc,
                                                  unsigned* s_tail, bool* more) {
    unsigned int* ctl = w.dr.self.ctl;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned gen = *(volatile unsigned*)&ctl[2];
        if (atomicAdd(&ctl[1], 1u) == nblocks - 1u) {
            ctl[1] = 0u;
            unsigned any;
            if (XSYNC) {
                const unsigned seq = ++ctl[DR_CTL_SEQ];
                if (!dr_strip_sync(w, seq, 0, had_work ? 1u : 0u, &any)) acc.fault++;
            }
            const unsigned tail = *(volatile unsigned*)&ctl[DR_CTL_TAIL];
            if (!XSYNC) any = tail != seg_end ? 1u : 0u;
            *(volatile unsigned*)&ctl[3 + (gen & 1u)] = tail;
            *(volatile unsigned*)&ctl[DR_CTL_MORE + (gen & 1u)] = any;
            __threadfence();
            atomicAdd(&ctl[2], 1u);
        } else {
            unsigned spins = 0;
            while (*(volatile unsigned*)&ctl[2] == gen) {
                __nanosleep(32);
                if (++spins > (1u << 27)) { acc.fault++; break; }      // a lost CTA: do not hang the GPU
            }
        }
        __threadfence();
        s_tail[0] = *(volatile unsigned*)&ctl[3 + (gen & 1u)];
        s_tail[1] = *(volatile unsigned*)&ctl[DR_CTL_MORE + (gen & 1u)];
    }
    __syncthreads();
    *more = s_tail[1] != 0u;
    return s_tail[0];
}

template <int POLT, bool STRIPS, bool XSYNC, bool UTIL, bool ASYNC>
__global__ void __launch_bounds__(DR_THREADS, DR_MIN_BLOCKS)
ss_dr_rounds_kernel(DevWorld w, DrStep st) {
    __shared__ DrShared sh;
    __shared__ unsigned s_tail[2];
    __shared__ double s_dur[UTIL ? DR_DUR_M * 128 : 1];
    const DrPeer& S = w.dr.self;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (UTIL)
        for (int i = tid; i < DR_DUR_M * 128; i += DR_THREADS) s_dur[i] = ((double)(i & 127) / (double)((i >> 7) + 1)) / w.cfg.max_capacity;
    for (int i = tid; i < DR_HIST_BINS; i += DR_THREADS) sh.hist[i] = 0u;
    if (tid < 8) sh.acc[tid] = 0ull;
    if (tid < DR_WARPS) sh.cnt[tid] = 0;
    __syncthreads();
    DrAccum acc = {0, 0, 0, 0, 0, 0, 0};
    unsigned seg_b = st.single_round ? st.seg_begin : 0u;
    unsigned seg_e = st.single_round ? st.seg_end : __ldcg(&S.ctl[DR_CTL_TAIL0]);       // the tail after the build (ss_dr_seg0_kernel)
    const unsigned gwarp = blockIdx.x * DR_WARPS + (unsigned)warp, nwarp = gridDim.x * DR_WARPS;
    if (ASYNC) {
        // No rounds: the queue is a work list.  A warp claims the next positions of the queue -- 32 while the backlog is
        // large, fewer (down to one: the lanes of a warp run their divergent turns one after the other) when entries
        // only trickle in -- and its lanes wait for their entries to be published (the queue was zeroed before the
        // build; an entry is written once, with DR_VALID), take their turn, and publish whom they released.  Every agent
        // the build saw passes through the queue exactly once, in an order that is a topological order of the conflict
        // graph: whoever holds the lowest unprocessed position never waits for a later one.
        const unsigned total = __ldcg(&S.ctl[DR_CTL_AGENTS]);
        unsigned int* head = &S.ctl[DR_CTL_HEAD];
        for (;;) {
            unsigned base = 0u, want = 0u;
            if (lane == 0) {
                const unsigned h = *(volatile unsigned*)head, t = *(volatile unsigned*)&S.ctl[DR_CTL_TAIL];
                const unsigned backlog = t > h ? t - h : 0u;
                want = min(32u, max(1u, (backlog * DR_ASYNC_MUL) / nwarp));
                base = h < total ? atomicAdd(head, want) : total;
            }
            base = __shfl_sync(FULL, base, 0); want = __shfl_sync(FULL, want, 0);
            if (base >= total) break;
            const unsigned pos = base + (unsigned)lane;
            bool pending = (unsigned)lane < want && pos < total;
            unsigned idle = 0;
            for (;;) {
                uint32_t e = 0u;
                if (pending) e = *(volatile uint32_t*)&S.queue[pos];
                const bool ready = pending && (e & DR_VALID) != 0u;
                if (__any_sync(FULL, ready)) {
                    if (ready) {
                        if (STRIPS && (e & DR_REMOTE)) __threadfence_system(); else __threadfence();      // the entry, then what its publisher wrote before it
                        dr_turn<POLT, STRIPS, UTIL, true>(w, st, sh, s_dur, warp, acc, e);
                        pending = false;
                    }
                    __syncwarp();
                    __threadfence();                        // (the entries published here are local; the ones for other strips fence on their own)
                    dr_flush(w, sh, warp, lane);
                    idle = 0;
                } else {
                    __nanosleep(DR_ASYNC_SLEEP);
                    if (++idle > (1u << 24)) { acc.fault++; pending = false; }      // a lost entry: do not hang the GPU
                }
                if (!__any_sync(FULL, pending)) break;
            }
        }
    } else
    for (;;) {
        const unsigned cnt = seg_e - seg_b;
        // a warp takes consecutive queue entries, up to 32 at a time, one lane each; every warp gets the same number of
        // batches (the lanes of a warp run their divergent turns one after the other: a small round is spread over all
        // warps -- fewer lanes, shorter round); no CTA-wide barrier inside a round
        const unsigned batches = max(1u, (cnt + 32u * nwarp - 1u) / (32u * nwarp));
        const unsigned per = min(32u, max(1u, (cnt + batches * nwarp - 1u) / (batches * nwarp)));
        for (unsigned base = gwarp * per; base < cnt; base += nwarp * per) {
            const unsigned i = base + (unsigned)lane;
            if ((unsigned)lane < per && i < cnt) dr_turn<POLT, STRIPS, UTIL, ASYNC>(w, st, sh, s_dur, warp, acc, __ldcg(&S.queue[seg_b + i]));
            __syncwarp();
            if (sh.cnt[warp] >= DR_WARP_CAP / 2) dr_flush(w, sh, warp, lane);
        }
        dr_flush(w, sh, warp, lane);
        if (st.single_round) break;
        seg_b = seg_e;
        bool more;
        seg_e = dr_grid_barrier<XSYNC>(w, st.nblocks, seg_e, cnt != 0u, acc, s_tail, &more);
        if (!more) break;
    }
    // statistics of this CTA's turns (sugarscape.py:575-583)
    unsigned vals[7] = {acc.sum_m, acc.sum_v, acc.acted, acc.deaths, acc.small, acc.inexact, acc.fault};
#pragma unroll
    for (int q = 0; q < 7; q++) {
        unsigned v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(&sh.acc[q], (unsigned long long)v);
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

// every agent that was alive at the start of the step must have passed through the queue; the successor pool must
// have held every list
__global__ void ss_dr_check_kernel(DevWorld w) {
    if (w.dr.self.ctl[DR_CTL_TAIL] != w.dr.self.ctl[DR_CTL_AGENTS]) atomicAdd(&w.acc->fault, 1ull);
    if (w.dr.self.ctl[7]) atomicAdd(&w.acc->fault, 1ull);
    w.acc->pending = 0ull;
}
__global__ void ss_dr_begin_kernel(DevWorld w) {
    unsigned int* c = w.dr.self.ctl;
    c[DR_CTL_TAIL] = 0u; c[DR_CTL_POOL] = 0u; c[DR_CTL_DELTA] = 0u; c[DR_CTL_AGENTS] = 0u; c[DR_CTL_HEAD] = 0u;
}
__global__ void ss_dr_seg0_kernel(DevWorld w) { w.dr.self.ctl[DR_CTL_TAIL0] = w.dr.self.ctl[DR_CTL_TAIL]; }

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

template <int POLT, bool STRIPS, bool XSYNC, bool UTIL, bool ASYNC>
static cudaError_t launch_rounds(const DevWorld& w, DrStep st, int sm_count, cudaStream_t s) {
    static int per_sm = 0;
    if (!per_sm) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ss_dr_rounds_kernel<POLT, STRIPS, XSYNC, UTIL, ASYNC>, DR_THREADS, 0);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    st.nblocks = (unsigned)(per_sm * sm_count);
    if (st.single_round) {
        ss_dr_rounds_kernel<POLT, STRIPS, XSYNC, UTIL, ASYNC><<<st.nblocks, DR_THREADS, 0, s>>>(w, st);
        return cudaGetLastError();
    }
    void* args[2] = {(void*)&w, (void*)&st};
    return cudaLaunchCooperativeKernel((const void*)ss_dr_rounds_kernel<POLT, STRIPS, XSYNC, UTIL, ASYNC>, dim3(st.nblocks), dim3(DR_THREADS), args, 0, s);
}

extern "C" {

size_t ss_dr_build_smem(int vmax) {
    const int HP = (2 * vmax + 3) & ~3, SX = DR_TX + 2 * HP, SY = DR_TY + 2 * HP, pad = vmax <= 7 ? 1 : 2;
    const int BW = ((SY + 31) >> 5) + pad, BWT = ((SX + 31) >> 5) + pad;
    const size_t pairs = (size_t)(DR_BUILD_THREADS / 32) * 1024 * (vmax <= 7 ? 2 : 4);     // the warps' pair lists
    return (size_t)SX * SY * 8 + ((size_t)SX * BW + (size_t)SY * BWT) * 4 + (size_t)DR_TX * DR_TY * 2 +
           std::max((size_t)SX * SY * 2 + 8, pairs) + 32;
}

cudaError_t ss_launch_dr_transpose(const DevWorld& w, cudaStream_t s) {
    const int rows = w.dr.self.rows, n = w.n;
    const int tiles = ((rows + 63) / 64) * ((n + 63) / 64);
    ss_dr_transpose_kernel<<<tiles, 256, 0, s>>>(w.dr.self.mw, w.dr.self.mwx, rows, n);
    return cudaGetLastError();
}

cudaError_t ss_launch_dr_canon(const DevWorld& w, cudaStream_t s) {
    ss_dr_canon_kernel<<<148 * 8, 256, 0, s>>>(w, (size_t)w.dr.self.rows * w.n);
    return cudaGetLastError();
}

// sugarscape.py:642-659: growback of the whole lattice, or of the two season regions with the season's factors
cudaError_t ss_launch_dr_grow(const DevWorld& w, int64_t iteration, int max_ctas, cudaStream_t s) {
    DrGrowArgs ga;
    ga.seasons = w.cfg.rule_seasons; ga.alpha = w.cfg.grow_factor; ga.alpha_north = ga.alpha_south = 0.0;
    if (w.cfg.rule_seasons) {
        const bool summer = (iteration % (2 * (int64_t)w.cfg.season_period)) < w.cfg.season_period;
        ga.alpha_north = summer ? w.cfg.season_on_factor : w.cfg.season_off_factor;
        ga.alpha_south = summer ? w.cfg.season_off_factor : w.cfg.season_on_factor;
    }
    const int tiles = ((w.dr.self.rows + 63) / 64) * ((w.n + 63) / 64);
    ga.tiles = tiles;
    ss_dr_grow_kernel<<<max_ctas > 0 ? std::min(tiles, max_ctas) : tiles, 256, 0, s>>>(w, ga);
    return cudaGetLastError();
}

cudaError_t ss_launch_dr_mirror(const DevWorld& w, int* flag, cudaStream_t s) {
    const size_t cells = (size_t)w.dr.self.rows * w.n;
    ss_dr_mirror_kernel<<<148 * 8, 256, 0, s>>>(w, cells, flag);
    if (w.n_slots) ss_dr_flags_kernel<<<(w.n_slots + 255) / 256, 256, 0, s>>>(w);
    return ss_launch_dr_transpose(w, s);
}

// tmap_storage: 128 bytes, 64-byte aligned, host memory owned by the engine (re-encoded when the lattice changes)
cudaError_t ss_dr_encode_tmap(const DevWorld& w, void* tmap_storage, int* ok) {
    *ok = 0;
    EncodeTiledFn enc = encode_tiled_fn();
    const int n = w.n, H = (2 * w.vmax + 3) & ~3;         // the staged halo (ss_dr_build_kernel)
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
        cudaError_t e = cudaFuncSetAttribute(ss_dr_build_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ss_dr_build_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    ss_dr_begin_kernel<<<1, 1, 0, s>>>(w);
    DrBuildArgs ba;
    ba.ok = ss_order_key(ss_step_key(w.cfg.keyed_seed, iteration), w.half);
    ba.stamp = (uint32_t)(iteration + 1);
    ba.tiles_y = (w.n + DR_TY - 1) / DR_TY;
    ba.use_tma = use_tma;
    const int tiles_x = (w.dr.self.rows + DR_TX - 1) / DR_TX;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (use_tma) memcpy(&tm, tmap_storage, sizeof tm);
    const bool strips = w.dr.world > 1;
    const int grid = tiles_x * ba.tiles_y;
    if (w.vmax <= 7) {
        if (strips) ss_dr_build_kernel<true, false><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
        else ss_dr_build_kernel<false, false><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
    } else {
        if (strips) ss_dr_build_kernel<true, true><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
        else ss_dr_build_kernel<false, true><<<grid, DR_BUILD_THREADS, smem, s>>>(w, ba, tm);
    }
    ss_dr_seg0_kernel<<<1, 1, 0, s>>>(w);
    return cudaGetLastError();
}

// single_round = 0: the whole agent phase in one cooperative launch; 1: the queue segment [seg_begin, seg_end) only
cudaError_t ss_launch_dr_rounds(const DevWorld& w, int64_t iteration, int64_t env_time, int sm_count, int single_round,
                                unsigned seg_begin, unsigned seg_end, int async_queue, cudaStream_t s) {
    DrStep st;
    st.iteration = iteration; st.env_time = env_time;
    st.skey = ss_step_key(w.cfg.keyed_seed, iteration);
    st.next_phase = (uint32_t)((iteration + 1) & 1);
    st.single_round = single_round; st.seg_begin = seg_begin; st.seg_end = seg_end; st.nblocks = 0;
    const bool pol = w.cfg.rule_pollution != 0;
    const bool strips = w.dr.world > 1;
    if (strips) {                   // (strips: no pollution rule; the in-kernel barrier over the strips unless the host drives the rounds)
        if (single_round) return launch_rounds<0, true, false, false, false>(w, st, sm_count, s);
        // (work list: every strip drains its own queue -- its neighbours publish into it over NVLink -- and the strips
        // meet again after the kernel; else one exchange per dependency round inside the kernel)
        if (async_queue) return launch_rounds<0, true, false, false, true>(w, st, sm_count, s);
        return launch_rounds<0, true, true, false, false>(w, st, sm_count, s);
    }
    if (async_queue) {              // one GPU: no rounds at all, the warps take queue entries as they are published
        if (w.cfg.rule_utilicalc) return launch_rounds<0, false, false, true, true>(w, st, sm_count, s);
        return pol ? launch_rounds<1, false, false, false, true>(w, st, sm_count, s) : launch_rounds<0, false, false, false, true>(w, st, sm_count, s);
    }
    if (w.cfg.rule_utilicalc) return launch_rounds<0, false, false, true, false>(w, st, sm_count, s);   // (never with pollution: the reference crashes)
    return pol ? launch_rounds<1, false, false, false, false>(w, st, sm_count, s) : launch_rounds<0, false, false, false, false>(w, st, sm_count, s);
}

cudaError_t ss_launch_dr_directory(const DevWorld& w, int n_all, const int32_t* id, const int32_t* x, const int32_t* metab,
                                   const double* sugar, cudaStream_t s) {
    if (n_all > 0) ss_dr_directory_kernel<<<(n_all + 255) / 256, 256, 0, s>>>(w, n_all, id, x, metab, sugar);
    return cudaGetLastError();
}
cudaError_t ss_launch_dr_after_rounds(const DevWorld& w, cudaStream_t s) {
    ss_dr_merge_stats_kernel<<<64, 256, 0, s>>>(w);
    ss_dr_apply_delta_kernel<<<dim3(64, (unsigned)w.dr.world), 256, 0, s>>>(w);
    return cudaGetLastError();
}
cudaError_t ss_launch_dr_strip_barrier(const DevWorld& w, cudaStream_t s) {
    ss_dr_strip_barrier_kernel<<<1, 1, 0, s>>>(w);
    return cudaGetLastError();
}

cudaError_t ss_launch_dr_check(const DevWorld& w, cudaStream_t s) {
    ss_dr_check_kernel<<<1, 1, 0, s>>>(w);
    return cudaGetLastError();
}

}  // extern "C"
