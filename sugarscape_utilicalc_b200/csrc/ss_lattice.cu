// ss_lattice.cu -- kernels over the lattice-major cell store (keyed RNG) that are not the dependency rounds of ss_rounds.cu:
//
//   ss_stats_kernel      population / means / Gini of a step (sugarscape.py:603-619, 269-280): every lattice path
//   ss_grow_kernel / ss_seasons_kernel   growback (environment.py:133-155) where the one-byte cell words are not in use
//   ss_diffuse_col_kernel  in-place pollution sweep as a column-chain wavefront (environment.py:317-350): every lattice path
//   upload / download helpers (mirrors, agent store + occupancy build, export in keyed order)
//   the replay kernels of the replicated multi-GPU scheme (sharded.py)
//   and the agent phase of ROUND 1, kept as tested alternatives (pass_impl 0-2) and for that scheme:
//   ss_prepare_kernel    Q1 bookkeeping: at-risk agents stamp their successor in the execution order (sugarscape.py:520 +
//                        502-503: the agent after a dying agent loses its turn)
//   ss_move_pass_kernel / ss_move_solo_kernel   window passes (agent.py:361-371 getFood, agent.py:794-865 move,
//                        sugarscape.py:566-601 pollute / metabolise / starve), repeated over shifted window tilings until no
//                        agent is pending; ss_settle_kernel finishes the turns of the agents that cannot starve
//
// Window passes.  The execution order of a step is the ascending keyed priority of the agent slot.  Two agents conflict iff
// their vision crosses share a cell; every read and write of an agent's turn lies inside its own cross, so an agent may take
// its turn as soon as every conflicting agent of lower priority has taken its own -- the result is identical to the
// sequential loop.  A launch cuts the lattice into disjoint windows (so CTAs / warps never touch the same cells); the agents
// of a window's core (>= 2*vmax from the window edge, so that every possible conflict partner is visible) are resolved in
// priority order, agents of the outer ring act as blockers and are resolved by a later launch whose tiling is shifted.
// ss_move_pass_kernel: a CTA per window staged in shared memory, a warp per region; ss_move_solo_kernel: a warp per window,
// cells read straight from L2.  The default agent phase is ss_rounds.cu (section 3.4).
#include "ss_world.cuh"

#define FULL 0xffffffffu
#define PASS_THREADS 256
// PASS_CLOCK() instrumentation of the move pass (tools/dbg_pass.py, option "debug"): compile with -DSS_PASS_TIMING
#ifdef SS_PASS_TIMING
#define PASS_CLOCK() clock64()
#else
#define PASS_CLOCK() 0ll
#endif
#ifndef PASS_MIN_CTAS
#define PASS_MIN_CTAS 2
#endif
#define PASS_WARPS (PASS_THREADS / 32)
#define PASS_RGX 2                      // regions per window: PASS_RGX x PASS_RGY == PASS_WARPS
#define PASS_RGY 4

// ------------------------------------------------------------------------ prepare (Q1) ----
// An agent is SAFE if it cannot starve this step whatever happens before its turn: it
// harvests at least int(sugar) of its own cell (nobody else can take an occupied cell and
// growback runs after the agent loop; with pollution the winner of the sugar/(1+p) ratio has
// sugar >= 1 whenever the own cell has).  Everything else is AT RISK: its turn is settled
// synchronously and its successor in the execution order waits for the outcome (Q1).
__global__ void ss_prepare_kernel(DevWorld w, int64_t iteration) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= w.n_slots) return;
    AgentRec* rec = &w.agents[slot];
    const uint32_t attr = rec->attr;
    if (!attr_alive(attr)) return;
    const double metab = (double)attr_metab(attr);
    if (ss_set_sugar(ss_max0(rec->sugar - metab)) > 0.1) return;          // sugarscape.py:569, 306-309
    const uint32_t pos = rec->pos;
    int hmin = 0;
    if (w.cfg.rule_move_eat) {
        hmin = (int)w.isugar[pos];
        if (w.cfg.rule_pollution && hmin > 1) hmin = 1;
    }
    if (ss_set_sugar(ss_max0(ss_set_sugar(ss_max0(rec->sugar + (double)hmin)) - metab)) > 0.1) return;
    atomicOr(&w.occw[pos], OCC_RISK);
    const OrderKey ok = ss_order_key(ss_step_key(w.cfg.keyed_seed, iteration), w.half);
    const uint32_t domain = 1u << (2 * w.half);
    for (uint32_t p = ss_priority(ok, slot) + 1u; p < domain; p++) {
        const uint32_t s2 = ss_priority_inverse(ok, p);
        if (s2 >= w.n_slots) continue;
        AgentRec* r2 = &w.agents[s2];
        if (!attr_alive(r2->attr)) continue;
        r2->pred = slot + 1u;
        r2->pred_step = (uint32_t)(iteration + 1);
        atomicOr(&w.occw[r2->pos], OCC_Q1);
        break;
    }
}

__global__ void ss_begin_step_kernel(DevWorld w) {
    StepAccum* a = w.acc;
    a->sum_metab = 0; a->sum_vision = 0; a->acted = 0; a->deaths = 0;
    a->pending = a->population; a->hist_small = 0; a->inexact = 0; a->outliers = 0;
}

// --------------------------------------------------------------------------- move pass ----
// Shared-memory image of one window: occupancy words and int(sugar) of every cell.
struct WinCtx {
    uint32_t* occ_s;      // [ex][emax]
    uint8_t* isug_s;      // [ex][emax]
    int ex, ey, emax;
    int gx0, gy0;         // global coordinate of local (0,0)
    bool full_x, full_y;  // window spans the whole torus in that dimension (wrap inside)
    int n;
    // cell accessors: (r,q) = window coordinates, cell = the same cell's lattice index
    __device__ inline uint32_t occ(int r, int q, int) const { return occ_s[r * emax + q]; }
    __device__ inline int isug(int r, int q, int) const { return (int)isug_s[r * emax + q]; }
    __device__ inline void put_occ(int r, int q, uint32_t v) const { occ_s[r * emax + q] = v; }
    __device__ inline void clear_isug(int r, int q) const { isug_s[r * emax + q] = 0; }
    __device__ static inline int gw(int v, int n) { return ss_wrap1(v, n); }      // lattice coordinate on the torus
    static constexpr bool kGlobal = false;
};
// The same window read straight from the lattice in HBM/L2 (ss_move_solo_kernel: one warp owns the whole window,
// so nothing else touches its cells during the launch); stores go to the lattice only.  WRAP = the window straddles
// the torus edge; all other windows (almost all) skip the coordinate wraps.
template <bool WRAP>
struct GlobCtx {
    const uint32_t* occw;
    const uint8_t* isugar;
    int ex, ey, emax;
    int gx0, gy0;
    static constexpr bool full_x = false, full_y = false;   // used with several windows per dimension only
    int n;
    __device__ inline uint32_t occ(int, int, int cell) const { return __ldcg(&occw[cell]); }
    __device__ inline int isug(int, int, int cell) const { return (int)__ldcg(&isugar[cell]); }
    __device__ inline void put_occ(int, int, uint32_t) const {}
    __device__ inline void clear_isug(int, int) const {}
    __device__ static inline int gw(int v, int n) { return WRAP ? ss_wrap1(v, n) : v; }
    static constexpr bool kGlobal = true;
};
template <class Ctx> __device__ static inline int wrap_row(const Ctx& c, int r) { return c.full_x ? ss_wrap1(r, c.ex) : r; }
template <class Ctx> __device__ static inline int wrap_col(const Ctx& c, int q) { return c.full_y ? ss_wrap1(q, c.ey) : q; }
template <class Ctx> __device__ static inline bool in_core(const Ctx& c, int r, int q, int R) {
    return (c.full_x || (r >= R && r < c.ex - R)) && (c.full_y || (q >= R && q < c.ey - R));
}
template <class Ctx> __device__ static inline int cell_of(const Ctx& c, int r, int q) {
    return Ctx::gw(c.gx0 + r, c.n) * c.n + Ctx::gw(c.gy0 + q, c.n);
}

// crosses of A (vision a) and B (vision b, displaced by dx,dy) share a cell
__device__ static inline bool crosses_meet(int dx, int dy, int a, int b) {
    const int adx = abs(dx), ady = abs(dy);
    return (adx <= a && ady <= b) || (ady <= a && adx <= b) || (dy == 0 && adx <= a + b) ||
           (dx == 0 && ady <= a + b);
}

struct ThreadStats {
    unsigned sum_m, sum_v, acted, deaths, resolved, small, inexact;
    unsigned long long log_pos, log_end;      // multi-GPU: this warp's reserved slice of the turn log
    unsigned long long clog_pos, clog_end;    // ... and of the compact log of safe turns
};
#define LOG_CHUNK 8
// next free entry of the rank's turn log: one global atomic per LOG_CHUNK entries of a warp
__device__ static inline ShardLogEntry* log_slot(const DevWorld& w, ThreadStats& ts) {
    if (ts.log_pos == ts.log_end) { ts.log_pos = atomicAdd(w.log_count, (unsigned long long)LOG_CHUNK); ts.log_end = ts.log_pos + LOG_CHUNK; }
    return &w.log[ts.log_pos++];
}

#define CLOG_UNUSED 0xffffffffu
__device__ static inline uint2* clog_slot(const DevWorld& w, ThreadStats& ts) {
    if (ts.clog_pos == ts.clog_end) { ts.clog_pos = atomicAdd(w.clog_count, (unsigned long long)LOG_CHUNK); ts.clog_end = ts.clog_pos + LOG_CHUNK; }
    return &w.clog[ts.clog_pos++];
}
// the reserved but unused entries of a warp's log slices are marked before the kernel ends
__device__ static inline void close_log_slices(const DevWorld& w, ThreadStats& ts) {
    while (ts.log_pos < ts.log_end) w.log[ts.log_pos++].kind_harvest = 255u;
    while (ts.clog_pos < ts.clog_end) w.clog[ts.clog_pos++] = make_uint2(CLOG_UNUSED, 0u);
}

struct Choice { int lr, lq, sg; };

// agent.py:361-371 getFood + agent.py:794-823 move(), one warp per agent: lane k evaluates
// candidate k of the cross (x-arm by ascending raw x, then y-arm; K = 2(2a+1) <= 62, so at
// most two candidates per lane).  Occupancy and int(sugar) come from the window's
// shared-memory copy; with the pollution rule each lane fetches its candidate's pollution
// from HBM (32 loads in flight).  Best = (key desc, distance asc, shuffled position asc).
// Without pollution key and distance pack into one integer and a single warp REDUX finds the
// best; ties are broken by replaying the agent's keyed Fisher-Yates draws (agent.py:369,
// random.py:350/242), each tied lane tracking the position of its own candidate; positions
// are finalised from the top, so the replay stops once a single tied candidate is left.
// The loads of one agent's cross, issued ahead of their use (the warp-per-window kernel overlaps them with its
// conflict test): per lane the occupancy word, int(sugar) and -- lattice-direct windows with the pollution rule --
// the pollution of its two candidates, plus the agent's own cell.
struct CrossLoads {         // raw loaded values: nothing may depend on them before they are consumed
    uint32_t occ[2];        // occupancy word of the candidate (1 = no candidate for this lane)
    int sg[2];
    double poll[2];
    int sg0;
    double pown;
};
// POLT: pollution rule known at compile time (0/1) or read from the configuration (-1)
template <int POLT, class Ctx>
__device__ static inline void warp_load_cross(const DevWorld& w, const Ctx& c, int lx, int ly, int gx, int gy, int a,
                                              int lane, CrossLoads& L) {
    const int n = c.n;
    const bool pol = POLT < 0 ? (w.cfg.rule_pollution != 0) : (POLT != 0);
    const int span = 2 * a + 1, K = 2 * span;
    L.sg0 = c.isug(lx, ly, gx * n + gy);
    L.pown = 0.0;
    if (Ctx::kGlobal && pol) L.pown = __ldcg(&w.poll[gx * n + gy]);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = h * 32 + lane;
        L.occ[h] = 1u; L.sg[h] = 0; L.poll[h] = 0.0;
        if (k < K) {
            int r, q, cx, cy;
            if (k < span) { const int o = k - a; r = wrap_row(c, lx + o); q = ly; cx = Ctx::gw(gx + o, n); cy = gy; }
            else { const int o = k - span - a; r = lx; q = wrap_col(c, ly + o); cx = gx; cy = Ctx::gw(gy + o, n); }
            const int ci = cx * n + cy;
            L.occ[h] = c.occ(r, q, ci);
            L.sg[h] = c.isug(r, q, ci);
            if (Ctx::kGlobal && pol) L.poll[h] = __ldcg(&w.poll[ci]);
        }
    }
}

template <int POLT, class Ctx>
__device__ static Choice warp_choose_cell(const DevWorld& w, const Ctx& c, int lx, int ly, int gx, int gy, int a,
                                          uint32_t slot, uint64_t skey, int lane, const CrossLoads& L) {
    const int n = c.n;
    const bool pol = POLT < 0 ? (w.cfg.rule_pollution != 0) : (POLT != 0);
    const int span = 2 * a + 1, K = 2 * span;
    int dist[2] = {0, 0}, cell[2] = {0, 0}, sg[2] = {0, 0}, qidx[2] = {-1, -1};
    uint32_t loc[2] = {0, 0};
    int n_free = 0;
    // own cell, appended last (agent.py:806); distance 0 wins every tie on the key
    const int sg0 = L.sg0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = h * 32 + lane;
        bool fr = false;
        if (h * 32 < K) {
            if (k < K && L.occ[h] == 0u) {
                int r, q, cx, cy;
                if (k < span) { const int o = k - a; r = wrap_row(c, lx + o); q = ly; cx = Ctx::gw(gx + o, n); cy = gy; }
                else { const int o = k - span - a; r = lx; q = wrap_col(c, ly + o); cx = gx; cy = Ctx::gw(gy + o, n); }
                fr = true;
                loc[h] = (uint32_t)(r << 8 | q); cell[h] = cx * n + cy; sg[h] = L.sg[h];
                dist[h] = abs(gx - cx) + abs(gy - cy);                       // agent.py:867-868
            }
            const unsigned fb = __ballot_sync(FULL, fr);
            if (fr) qidx[h] = n_free + __popc(fb & ((1u << lane) - 1u));
            n_free += __popc(fb);
        }
    }
    Choice stay; stay.lr = lx; stay.lq = ly; stay.sg = sg0;
    if (n_free == 0) return stay;
    bool tie0, tie1;
    if (!pol) {
        const uint32_t k0 = qidx[0] >= 0 ? (((uint32_t)sg[0] << 20) | (0xfffffu - (uint32_t)dist[0])) + 1u : 0u;
        const uint32_t k1 = qidx[1] >= 0 ? (((uint32_t)sg[1] << 20) | (0xfffffu - (uint32_t)dist[1])) + 1u : 0u;
        const uint32_t kmax = __reduce_max_sync(FULL, k0 > k1 ? k0 : k1);
        if (sg0 >= (int)((kmax - 1u) >> 20)) return stay;
        tie0 = k0 == kmax; tie1 = k1 == kmax;
    } else {
        double key[2] = {-1.0, -1.0};
        double p0 = 0.0, p1 = 0.0, pown = 0.0;
        if (qidx[0] >= 0 && sg[0] > 0) p0 = Ctx::kGlobal ? L.poll[0] : __ldcg(&w.poll[cell[0]]);
        if (qidx[1] >= 0 && sg[1] > 0) p1 = Ctx::kGlobal ? L.poll[1] : __ldcg(&w.poll[cell[1]]);
        if (sg0 > 0) pown = Ctx::kGlobal ? L.pown : __ldcg(&w.poll[gx * n + gy]);
        if (qidx[0] >= 0) key[0] = (double)sg[0] / (1.0 + p0);
        if (qidx[1] >= 0) key[1] = (double)sg[1] / (1.0 + p1);
        const double key0 = (double)sg0 / (1.0 + pown);
        double kmax = key[0] > key[1] ? key[0] : key[1];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(FULL, kmax, o); kmax = t > kmax ? t : kmax; }
        if (key0 >= kmax) return stay;
        int dmin = 0x7fffffff;
        if (qidx[0] >= 0 && key[0] == kmax) dmin = dist[0];
        if (qidx[1] >= 0 && key[1] == kmax && dist[1] < dmin) dmin = dist[1];
        dmin = __reduce_min_sync(FULL, dmin);
        tie0 = qidx[0] >= 0 && key[0] == kmax && dist[0] == dmin;
        tie1 = qidx[1] >= 0 && key[1] == kmax && dist[1] == dmin;
    }
    int left = __popc(__ballot_sync(FULL, tie0)) + __popc(__ballot_sync(FULL, tie1));
    if (left > 1) {
        // the agent's keyed words 32 at a time (lane l holds word base + l), handed round by shuffle
        const uint64_t akey = ss_agent_key(skey, slot + 1u);
        uint32_t j = 0, base = 0;
        uint32_t batch = ss_keyed_word(akey, (uint32_t)lane);
        auto next_word = [&]() {
            if (j - base >= 32u) { base += 32u; batch = ss_keyed_word(akey, base + (uint32_t)lane); }
            return __shfl_sync(FULL, batch, (int)(j++ - base));
        };
        int p0 = qidx[0], p1 = qidx[1];
        for (int i = n_free - 1; i >= 1 && left > 1; i--) {           // random.py:350 shuffle
            const uint32_t nn = (uint32_t)i + 1u;
            const int kbits = 32 - __clz(nn);
            uint32_t rj = next_word() >> (32 - kbits);               // random.py:242 _randbelow
            while (rj >= nn) {
                rj = next_word() >> (32 - kbits);
                if (j > 100000u) { atomicAdd(&w.acc->fault, 1ull); w.acc->wd_info[2] = 0x3333; break; }
            }
            // the element at index rj moves to index i and is final there; L[i] moves to rj
            bool fin = false;
            if (tie0) { if (p0 == (int)rj) { tie0 = false; fin = true; } else if (p0 == i) p0 = (int)rj; }
            if (tie1) { if (p1 == (int)rj) { tie1 = false; fin = true; } else if (p1 == i) p1 = (int)rj; }
            left -= __popc(__ballot_sync(FULL, fin));
        }
    }
    // the surviving tied candidate ends lowest in the shuffled list
    const unsigned who = __ballot_sync(FULL, tie0 || tie1);
    const int src = __ffs(who) - 1;
    const uint32_t payload = tie0 ? (((uint32_t)sg[0] << 16) | loc[0]) : (((uint32_t)sg[1] << 16) | loc[1]);
    const uint32_t got = __shfl_sync(FULL, payload, src);
    Choice ch;
    ch.lr = (int)((got >> 8) & 255u); ch.lq = (int)(got & 255u); ch.sg = (int)(got >> 16);
    return ch;
}

// One agent's turn (one warp per agent; lane 0 commits).  SAFE agents (no OCC_RISK flag)
// cannot starve, so only their move is done here -- shared-memory reads, fire-and-forget
// HBM stores; harvest, pollution, metabolism and statistics are deferred to ss_settle_kernel
// (nobody can observe them before the agent loop ends: the destination stays occupied).
// AT-RISK agents are settled synchronously because a death frees the cell and decides the
// successor's turn (Q1).
// Commit of one agent's turn by ONE thread, once the destination is chosen: (lx,ly) -> (lr,lq) in window coordinates,
// pos -> cell on the lattice, harvest sg.  sugar / attr are the agent's record fields (read only for at-risk agents).
// risky: settle the whole turn now (at-risk agents; every agent under the utilicalc rule, whose neighbours read its
// sugar); publish: somebody may be waiting for the outcome (Q1) -- fence and status word.
template <class Ctx>
__device__ static inline void commit_turn(const DevWorld& w, const Ctx& c, int lx, int ly, int lr, int lq, int pos, int cell,
                                          int sg, int a, bool risky, uint32_t slot, double sugar, uint32_t attr,
                                          int64_t iteration, int64_t env_time, uint32_t next_phase, ThreadStats& ts,
                                          bool publish = true) {
    AgentRec* recp = &w.agents[slot];
    uint32_t new_word = occ_pack(slot, a, next_phase);
    ts.resolved++;
    if (!risky) {
        c.put_occ(lx, ly, 0u);
        c.put_occ(lr, lq, new_word);
        if (cell != pos) w.occw[pos] = 0u;
        w.occw[cell] = new_word;
        if (w.cfg.rule_move_eat) {
            c.clear_isug(lr, lq);
            w.sugar[cell] = 0.0;                                      // agent.py:864
            w.isugar[cell] = 0;
        }
        recp->pos = (uint32_t)cell;
        recp->turn = ((uint32_t)(iteration + 1) << 8) | (uint32_t)sg;
        if (w.log) *clog_slot(w, ts) = make_uint2(slot | ((uint32_t)sg << 24), (uint32_t)cell);
        return;
    }
    const bool pol = w.cfg.rule_pollution != 0 && w.cfg.pollution_start_time <= env_time;
    const int metab = attr_metab(attr);
    if (w.cfg.rule_move_eat || w.cfg.rule_utilicalc) {
        if (pol) w.poll[cell] = __ldcg(&w.poll[cell]) + (((double)sg * w.cfg.pollution_a) + (0.0 * w.cfg.pollution_b));
        sugar = ss_set_sugar(ss_max0(sugar + (double)sg));            // agent.py:832 / 1128-1130
        c.clear_isug(lr, lq);
        w.sugar[cell] = 0.0;
        w.isugar[cell] = 0;
    }
    if (pol && sugar > 0.0)                                           // sugarscape.py:566-567
        w.poll[cell] = __ldcg(&w.poll[cell]) + ((0.0 * w.cfg.pollution_a) + ((double)metab * w.cfg.pollution_b));
    sugar = ss_set_sugar(ss_max0(sugar - (double)metab));             // sugarscape.py:569
    ts.sum_m += (unsigned)metab;
    ts.sum_v += (unsigned)a;
    ts.acted++;
    if (sugar == 0.01) ts.small++;
    else if (sugar >= 0.0 && sugar < (double)SS_GINI_BINS && sugar == (double)(int)sugar) atomicAdd(&w.hist[(int)sugar], 1u);
    else { ts.inexact++; ss_note_outlier(w, sugar); }
    const bool died = w.cfg.rule_can_starve && sugar <= 0.1;          // sugarscape.py:306-309
    if (died) new_word = 0u;
    c.put_occ(lx, ly, 0u);
    c.put_occ(lr, lq, new_word);
    if (cell != pos) w.occw[pos] = 0u;
    w.occw[cell] = new_word;
    if (died) { attr &= ~ATTR_ALIVE; ts.deaths++; }
    recp->sugar = sugar;
    recp->pos = (uint32_t)cell;
    recp->attr = attr;
    const uint32_t status = ((uint32_t)(iteration + 1) << 2) | (died ? ST_DIED : ST_ACTED);
    if (publish) {
        __threadfence();                   // the turn's effects before its status (Q1 readers)
        *((volatile uint32_t*)&recp->status) = status;
    }
    if (w.log) {
        ShardLogEntry en;
        en.slot = slot; en.old_cell = (uint32_t)pos; en.new_cell = (uint32_t)cell; en.new_word = new_word;
        en.sugar = sugar; en.poll = w.cfg.rule_pollution ? w.poll[cell] : 0.0; en.attr = attr; en.status = status; en.turn = 0u;
        en.kind_harvest = 1u | ((uint32_t)sg << 8);
        *log_slot(w, ts) = en;
    }
}

// One agent's turn (one warp per agent; lane 0 commits).  SAFE agents (no OCC_RISK flag)
// cannot starve, so only their move is done here -- shared-memory reads, fire-and-forget
// HBM stores; harvest, pollution, metabolism and statistics are deferred to ss_settle_kernel
// (nobody can observe them before the agent loop ends: the destination stays occupied).
// AT-RISK agents are settled synchronously because a death frees the cell and decides the
// successor's turn (Q1).
template <int POLT, class Ctx>
__device__ static void warp_agent_turn(const DevWorld& w, const Ctx& c, int lx, int ly, uint32_t ow, uint32_t slot,
                                       int64_t iteration, int64_t env_time, uint64_t skey, uint32_t next_phase,
                                       ThreadStats& ts, int lane, const CrossLoads& L) {
    const int n = c.n;
    const int a = occ_vision(ow);
    const int gx = Ctx::gw(c.gx0 + lx, n), gy = Ctx::gw(c.gy0 + ly, n);
    const int pos = gx * n + gy;
    const bool risky = (ow & OCC_RISK) != 0;
    const AgentRec* recp = &w.agents[slot];
    double sugar = 0.0;
    uint32_t attr = 0;
    if (risky && lane == 0) { sugar = recp->sugar; attr = recp->attr; }   // issued early, used after the scan
    Choice ch;
    if (w.cfg.rule_move_eat) ch = warp_choose_cell<POLT>(w, c, lx, ly, gx, gy, a, slot, skey, lane, L);
    else { ch.lr = lx; ch.lq = ly; ch.sg = 0; }
    if (lane != 0) return;
    const int cell = (ch.lr == lx && ch.lq == ly) ? pos
                   : Ctx::gw(c.gx0 + ch.lr, n) * n + Ctx::gw(c.gy0 + ch.lq, n);
    commit_turn(w, c, lx, ly, ch.lr, ch.lq, pos, cell, ch.sg, a, risky, slot, sugar, attr, iteration, env_time, next_phase, ts);
}

// ------------------------------------------------------------------------------------------
// agent.py:1025-1141 utilicalcMove(), sugar-only (utilicalcSugar agent.py:980-992), one warp per agent in the
// warp-per-window kernel -- the welfare evaluation is fused into the movement scan: the lanes scan the cross once for
// free cells (moves) and occupants (the vision neighbourhood), lane 0 replays the three Fisher-Yates shuffles on the
// agent's keyed words, the utility of every move is summed by one lane over the neighbours in list order (the sums are
// compared for equality, agent.py:1105) and never leaves shared memory.  Every turn is settled on the spot: the
// neighbours' utilities read Agent.sugar.
#define UT_MAXC 64                      // >= 2(2v+1)+1 for v <= 15
struct UtilShared {
    double util[UT_MAXC], nb_inten[UT_MAXC], nb_metab[UT_MAXC];
    int32_t food[UT_MAXC], neigh[UT_MAXC], nb_x[UT_MAXC], nb_y[UT_MAXC], nb_v[UT_MAXC], fsug[UT_MAXC];
};
// The agent's keyed words, 32 at a time (lane l holds word base + l) and handed round by shuffle: every lane sees the
// same draw, so the whole warp replays CPython's algorithms in step.
struct KeyedStream { uint64_t akey; uint32_t j, base, batch; };
__device__ static inline void ks_open(KeyedStream& ks, uint64_t akey, int lane) {
    ks.akey = akey; ks.j = 0u; ks.base = 0u; ks.batch = ss_keyed_word(akey, (uint32_t)lane);
}
__device__ static inline uint32_t ks_word(KeyedStream& ks, int lane) {
    if (ks.j - ks.base >= 32u) { ks.base += 32u; ks.batch = ss_keyed_word(ks.akey, ks.base + (uint32_t)lane); }
    return __shfl_sync(FULL, ks.batch, (int)(ks.j++ - ks.base));
}
__device__ static inline uint32_t ks_randbelow(KeyedStream& ks, uint32_t nn, int lane) {         // random.py:242
    const int k = 32 - __clz(nn);
    uint32_t r = ks_word(ks, lane) >> (32 - k);
    while (r >= nn) r = ks_word(ks, lane) >> (32 - k);
    return r;
}
// random.py:350 shuffle of a[0..n), n <= 64: the draws are sequential, the swaps are not -- every lane follows the
// positions of its (up to) two elements through the swaps and stores them where they end
// (p0, p1: where the elements that started at positions lane and lane + 32 are now; several shuffles compose)
__device__ static inline void ks_shuffle_track(KeyedStream& ks, int n, int lane, int& p0, int& p1) {
    for (int i = n - 1; i >= 1; i--) {
        const int j = (int)ks_randbelow(ks, (uint32_t)i + 1u, lane);
        p0 = p0 == i ? j : (p0 == j ? i : p0);
        p1 = p1 == i ? j : (p1 == j ? i : p1);
    }
}
// the neighbour's share of utilicalcSugar that does not depend on the move (agent.py:870-872, 984): the days its sugar
// lasts are sugar // metabolism; an integer-valued sugar (everything but the 0.01 floor) takes the integer quotient
__device__ static inline double util_intensity(double sugar, int metab) {
    const double days = (sugar >= 0.0 && sugar < 2147483647.0 && sugar == (double)(int)sugar)
                            ? (double)((int)sugar / metab) : ss_py_floordiv(sugar, (double)metab);
    return 1.0 / (1.0 + days);
}
template <bool WRAP, class Ctx>
__device__ static void warp_utilicalc_turn(const DevWorld& w, const Ctx& c, UtilShared& us, int lx, int ly, int gx, int gy,
                                           int a, uint32_t ow, uint32_t slot, int64_t iteration, int64_t env_time,
                                           uint64_t skey, uint32_t next_phase, ThreadStats& ts, int lane) {
    const int n = w.n, pos = gx * n + gy, span = 2 * a + 1, K = 2 * span;
    int m = 0, nb = 0;
    for (int base = 0; base < K; base += 32) {          // agent.py:361-371 getFood / 391-407 getVisionNeighbourhood
        const int k = base + lane;
        bool fr = false, oc = false;
        int cell = 0;
        uint32_t word = 0u;
        if (k < K) {
            int cx, cy;
            bool own;
            if (k < span) { cx = WRAP ? ss_wrap1(gx - a + k, n) : gx - a + k; cy = gy; own = cx == gx; }
            else { cx = gx; cy = WRAP ? ss_wrap1(gy - a + (k - span), n) : gy - a + (k - span); own = cy == gy; }
            cell = cx * n + cy;
            word = __ldcg(&w.occw[cell]);
            fr = word == 0u;
            oc = word != 0u && !own;
        }
        const unsigned bf = __ballot_sync(FULL, fr), bo = __ballot_sync(FULL, oc), lt = (1u << lane) - 1u;
        if (fr) us.food[m + __popc(bf & lt)] = cell;
        if (oc) us.neigh[nb + __popc(bo & lt)] = (int32_t)occ_slot(word);
        m += __popc(bf); nb += __popc(bo);
    }
    __syncwarp();
    // the loads the evaluation will want are issued now and land while the shuffles are replayed: int(sugar) of the
    // moves, the records of the neighbours, the agent's own record and cell
    const AgentRec* recp = &w.agents[slot];
    double own_sugar = 0.0;
    uint32_t own_attr = 0u;
    int own_site = 0;
    if (lane == 0) { own_sugar = __ldcg(&recp->sugar); own_attr = __ldcg(&recp->attr); own_site = (int)__ldcg(&w.isugar[pos]); }
    const int32_t f0 = lane < m ? us.food[lane] : 0, f1 = lane + 32 < m ? us.food[lane + 32] : 0;
    const int fs0 = lane < m ? (int)__ldcg(&w.isugar[f0]) : 0, fs1 = lane + 32 < m ? (int)__ldcg(&w.isugar[f1]) : 0;
    const int32_t s0 = lane < nb ? us.neigh[lane] : 0, s1 = lane + 32 < nb ? us.neigh[lane + 32] : 0;
    double bs0 = 0.0, bs1 = 0.0;
    uint32_t ba0 = 0u, ba1 = 0u, bp0 = 0u, bp1 = 0u;
    if (lane < nb) { const AgentRec* b = &w.agents[s0]; bs0 = __ldcg(&b->sugar); ba0 = __ldcg(&b->attr); bp0 = __ldcg(&b->pos); }
    if (lane + 32 < nb) { const AgentRec* b = &w.agents[s1]; bs1 = __ldcg(&b->sugar); ba1 = __ldcg(&b->attr); bp1 = __ldcg(&b->pos); }
    KeyedStream ks;
    ks_open(ks, ss_agent_key(skey, slot + 1u), lane);
    int p0 = lane, p1 = lane + 32, q0 = lane, q1 = lane + 32;
    ks_shuffle_track(ks, m, lane, p0, p1);      // getFood's shuffle, agent.py:369
    ks_shuffle_track(ks, m, lane, p0, p1);      // agent.py:1033
    ks_shuffle_track(ks, nb, lane, q0, q1);     // agent.py:405
    __syncwarp();
    if (lane < m) { us.food[p0] = f0; us.fsug[p0] = fs0; }
    if (lane + 32 < m) { us.food[p1] = f1; us.fsug[p1] = fs1; }
    if (lane < nb) {
        us.neigh[q0] = s0; us.nb_inten[q0] = util_intensity(bs0, attr_metab(ba0)); us.nb_metab[q0] = (double)attr_metab(ba0);
        us.nb_x[q0] = (int)bp0 / n; us.nb_y[q0] = (int)bp0 % n; us.nb_v[q0] = attr_vision(ba0);
    }
    if (lane + 32 < nb) {
        us.neigh[q1] = s1; us.nb_inten[q1] = util_intensity(bs1, attr_metab(ba1)); us.nb_metab[q1] = (double)attr_metab(ba1);
        us.nb_x[q1] = (int)bp1 / n; us.nb_y[q1] = (int)bp1 % n; us.nb_v[q1] = attr_vision(ba1);
    }
    if (lane == 0) {
        us.food[m] = pos; us.fsug[m] = own_site;                                    // agent.py:1035
        us.neigh[nb] = (int32_t)slot;                                               // agent.py:1039
        us.nb_inten[nb] = util_intensity(own_sugar, attr_metab(own_attr)); us.nb_metab[nb] = (double)attr_metab(own_attr);
        us.nb_x[nb] = gx; us.nb_y[nb] = gy; us.nb_v[nb] = a;
    }
    m += 1; nb += 1;
    __syncwarp();
    const double extent = (double)(nb - 1);
    double lmax = 0.0;
    bool have = false;
    for (int p = lane; p < m; p += 32) {
        const int cell = us.food[p], mx = cell / n, my = cell % n;
        const double site = (double)us.fsug[p];
        double u = 0.0;
        for (int q = 0; q < nb; q++) {
            const double dur = (site / us.nb_metab[q]) / w.cfg.max_capacity;
            const int bx = us.nb_x[q], by = us.nb_y[q], bv = us.nb_v[q];
            const int cert = (bx == mx && abs(by - my) <= bv) || (by == my && abs(bx - mx) <= bv);
            double t = 0.0;
            t += (double)cert * (((us.nb_inten[q] + dur) + 0.5 * 0.0) + extent);
            if (us.neigh[q] != (int32_t)slot) t *= -1.0;
            u += t;
        }
        us.util[p] = u;
        if (!have || u > lmax) { lmax = u; have = true; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(FULL, lmax, o);
        const int oh = __shfl_xor_sync(FULL, (int)have, o);
        if (oh && (!have || om > lmax)) { lmax = om; have = true; }
    }
    __syncwarp();
    // max_locations in move-list order (agent.py:1105-1107): the maxima among moves 0..31 and 32..63 as two ballots
    const unsigned b0 = __ballot_sync(FULL, lane < m && us.util[lane] == lmax);
    const unsigned b1 = __ballot_sync(FULL, lane + 32 < m && us.util[min(lane + 32, UT_MAXC - 1)] == lmax);
    const int n0 = __popc(b0), nmax = n0 + __popc(b1);
    const int pick = (int)ks_randbelow(ks, (uint32_t)nmax, lane);                   // random.choice, agent.py:1108
    const int pmove = pick < n0 ? (int)__fns(b0, 0u, pick + 1) : 32 + (int)__fns(b1, 0u, pick - n0 + 1);
    if (lane == 0) {
        const int dest = us.food[pmove];
        const int sg = us.fsug[pmove];                                              // agent.py:1128-1130
        commit_turn(w, c, lx, ly, lx, ly, pos, dest, sg, a, true, slot, own_sugar, own_attr, iteration, env_time, next_phase,
                    ts, (ow & OCC_RISK) != 0u);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Lane-per-agent turn (ss_move_solo_kernel<POLT, true>): one THREAD scans its agent's cross and chooses the cell --
// agent.py:361-371 getFood + agent.py:794-823 move().  Candidates are numbered as in warp_choose_cell (x-arm by
// ascending offset, then y-arm); best = (key desc, raw Manhattan distance asc, position in the shuffled food list
// asc).  At most four candidates can tie on (key, distance) -- one distance, two arms, two sides -- so the tied set
// is four (harvest, index) pairs in a register; the Fisher-Yates replay tracks their positions as warp_choose_cell does.
struct LaneChoice { int cell, sg; };
template <int POLT, bool WRAP>
__device__ static inline LaneChoice lane_choose_cell(const DevWorld& w, int gx, int gy, int a, uint32_t slot, uint64_t skey) {
    const int n = w.n;
    constexpr bool pol = POLT != 0;
    const int span = 2 * a + 1;
    const int pos = gx * n + gy;
    const int sg0 = (int)__ldcg(&w.isugar[pos]);
    double pown = 0.0;
    if (pol) pown = __ldcg(&w.poll[pos]);
    unsigned long long fm = 0ull;           // free candidates, bit = candidate index k
    unsigned long long tl = 0ull;           // tied best candidates, 16 bits each: harvest << 8 | k
    int nt = 0, dmin = 0x7fffffff;
    uint32_t ikmax = 0u;
    double dkmax = -1.0;
#pragma unroll 1
    for (int d = 1; d <= a; d++) {
        int cx[4], cy[4], kk[4];
        cx[0] = WRAP ? ss_wrap1(gx - d, n) : gx - d; cy[0] = gy; kk[0] = a - d;
        cx[1] = WRAP ? ss_wrap1(gx + d, n) : gx + d; cy[1] = gy; kk[1] = a + d;
        cx[2] = gx; cy[2] = WRAP ? ss_wrap1(gy - d, n) : gy - d; kk[2] = span + a - d;
        cx[3] = gx; cy[3] = WRAP ? ss_wrap1(gy + d, n) : gy + d; kk[3] = span + a + d;
        uint32_t oc[4];
        int sg[4];
        double pp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int ci = cx[u] * n + cy[u];
            oc[u] = __ldcg(&w.occw[ci]);
            sg[u] = (int)__ldcg(&w.isugar[ci]);
            pp[u] = pol ? __ldcg(&w.poll[ci]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (oc[u] != 0u) continue;
            fm |= 1ull << kk[u];
            const int dist = WRAP ? abs(gx - cx[u]) + abs(gy - cy[u]) : d;          // agent.py:867-868 (raw coordinates)
            const unsigned long long ent = ((unsigned long long)sg[u] << 8) | (unsigned long long)kk[u];
            if (!pol) {
                const uint32_t ik = (((uint32_t)sg[u] << 20) | (0xfffffu - (uint32_t)dist)) + 1u;
                if (ik > ikmax) { ikmax = ik; tl = ent; nt = 1; }
                else if (ik == ikmax) { tl = (tl << 16) | ent; nt++; }
            } else {
                const double key = (double)sg[u] / (1.0 + (sg[u] > 0 ? pp[u] : 0.0));
                if (key > dkmax || (key == dkmax && dist < dmin)) { dkmax = key; dmin = dist; tl = ent; nt = 1; }
                else if (key == dkmax && dist == dmin) { tl = (tl << 16) | ent; nt++; }
            }
        }
    }
    LaneChoice ch;
    ch.cell = pos; ch.sg = sg0;
    if (nt == 0) return ch;                                                         // no free cell in sight
    // own cell, appended last (agent.py:806); distance 0 wins every tie on the key
    if (!pol) { if (sg0 >= (int)((ikmax - 1u) >> 20)) return ch; }
    else { if ((double)sg0 / (1.0 + (sg0 > 0 ? pown : 0.0)) >= dkmax) return ch; }
    int win = 0;
    if (nt > 1) {
        if (nt > 4) { atomicAdd(&w.acc->fault, 1ull); w.acc->wd_info[2] = 0x4444; nt = 4; }
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
                if (j > 100000u) { atomicAdd(&w.acc->fault, 1ull); w.acc->wd_info[2] = 0x3333; break; }
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
    int cx, cy;
    if (k < span) { const int o = k - a; cx = WRAP ? ss_wrap1(gx + o, n) : gx + o; cy = gy; }
    else { const int o = k - span - a; cx = gx; cy = WRAP ? ss_wrap1(gy + o, n) : gy + o; }
    ch.cell = cx * n + cy; ch.sg = (int)(ent >> 8);
    return ch;
}

// ------------------------------------------------------------------------------------------
// Move pass = conservative parallel discrete-event simulation with the execution priority as
// the time stamp.  A CTA stages one window in shared memory and splits it into PASS_WARPS
// regions, one warp each.  A warp takes the pending agents of its region strictly in
// ascending priority (repeated warp-min selection over keys held in registers), so agents of
// one region never need a conflict test against each other.  Before an agent A acts, the warp
// publishes its frontier f = priority(A) ("everything below f in my region is finished") and
// waits until every region its conflict zone reaches has a frontier above priority(A): then
// all conflicting agents of lower priority have finished, and all of higher priority are still
// waiting for A.  The region holding the globally smallest unfinished priority never waits,
// so the scheme cannot deadlock.  Agents that cannot be resolved in this launch (window ring,
// unresolved Q1 predecessor, conflict with such an agent) are appended to the region's
// blocked set, which later agents within reach check with the exact cross-intersection test.
#define PASS_LIST 256                    // listed agents per region and launch
#define PASS_KPL (PASS_LIST / 32)
#define PRIO_INF 0xffffffffu
#define PASS_BROWS 6                     // 16x16-cell blocks per region: up to 6 x 5 (regions up to 96 x 80 cells)
#define PASS_BCOLS 5
struct RegionShared {
    volatile uint32_t bfront[32];        // per block: lowest priority among its unfinished listed agents
    volatile int n_list;
    volatile int done;                   // the region's launch is over (fronts may stay finite if the list was trimmed)
    uint64_t list[PASS_LIST];            // staging for the per-lane keys: priority << 16 | pos
};
struct PassShared {
    unsigned red[PASS_WARPS];
    int go;
    int dirty;
    int count[PASS_WARPS];
};

__global__ void __launch_bounds__(PASS_THREADS, PASS_MIN_CTAS)
ss_move_pass_kernel(DevWorld w, PassGeom g, int64_t iteration, int64_t env_time) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ PassShared sh;
    const int n = w.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sh.go = (w.acc->pending != 0ull);
    __syncthreads();
    if (!sh.go) return;
    // ---- window geometry: balanced boundaries in units of `align` cells, shifted tiling ----
    const int bx = g.bx_lo + (int)blockIdx.x / g.nwin, by = (int)blockIdx.x % g.nwin;
    const int nu = n / g.align;
    const int x_lo = g.align * (int)(((long long)bx * nu) / g.nwin), x_hi = g.align * (int)(((long long)(bx + 1) * nu) / g.nwin);
    const int y_lo = g.align * (int)(((long long)by * nu) / g.nwin), y_hi = g.align * (int)(((long long)(by + 1) * nu) / g.nwin);
    const int V = w.vmax, R = g.ring;
    WinCtx c;
    c.n = n; c.ex = x_hi - x_lo; c.ey = y_hi - y_lo; c.emax = g.emax;
    c.gx0 = x_lo + g.shift_x; c.gy0 = y_lo + g.shift_y;
    c.full_x = (c.ex == n); c.full_y = (c.ey == n);
    c.occ_s = smem;
    c.isug_s = (uint8_t*)(smem + g.emax * g.emax);
    RegionShared* regs = (RegionShared*)((uint8_t*)smem + (((size_t)g.emax * g.emax * 5 + 15) & ~(size_t)15));
    uint32_t* blk_bits = (uint32_t*)(regs + PASS_WARPS);     // [emax][ew]: pending agents this launch cannot resolve
    const int ew = g.ew;
    const uint32_t cur_phase = (uint32_t)(iteration & 1), next_phase = cur_phase ^ 1u;
    // ---- coarse pending map: a window whose core holds no agent left over by the previous launch
    //      hands the map entries of its blocks on and leaves without staging anything ----
    const int nb = w.pend_nb;
    if (g.map_read >= 0) {
        const uint32_t* mr = w.pendmap + (size_t)g.map_read * nb * nb;
        const int cx0 = (c.gx0 + R) >> 5, cx1 = (c.gx0 + c.ex - R - 1) >> 5;      // blocks overlapping the core
        const int cy0 = (c.gy0 + R) >> 5, cy1 = (c.gy0 + c.ey - R - 1) >> 5;
        const int ncx = cx1 - cx0 + 1, ncy = cy1 - cy0 + 1;
        int any = 0;
        for (int i = tid; i < ncx * ncy; i += PASS_THREADS)
            any |= mr[((cx0 + i / ncy) % nb) * nb + ((cy0 + i % ncy) % nb)] != 0u;
        if (!__syncthreads_or(any)) {
            if (g.map_write >= 0) {
                uint32_t* mw = w.pendmap + (size_t)g.map_write * nb * nb;
                const int wx0 = c.gx0 >> 5, wx1 = (c.gx0 + c.ex - 1) >> 5, wy0 = c.gy0 >> 5, wy1 = (c.gy0 + c.ey - 1) >> 5;
                const int nwx = wx1 - wx0 + 1, nwy = wy1 - wy0 + 1;
                for (int i = tid; i < nwx * nwy; i += PASS_THREADS) {
                    const int bi = ((wx0 + i / nwy) % nb) * nb + ((wy0 + i % nwy) % nb);
                    const uint32_t v = mr[bi];
                    if (v) atomicAdd(&mw[bi], v);
                }
            }
            return;
        }
    }
    if (tid == 0) sh.dirty = 0;
    // what this launch leaves pending (ring + blocked agents; everything if a region's list was trimmed)
    auto write_map = [&]() {
        if (g.map_write < 0) return;
        uint32_t* mw = w.pendmap + (size_t)g.map_write * nb * nb;
        if (sh.dirty) {
            const int wx0 = c.gx0 >> 5, wx1 = (c.gx0 + c.ex - 1) >> 5, wy0 = c.gy0 >> 5, wy1 = (c.gy0 + c.ey - 1) >> 5;
            const int nwx = wx1 - wx0 + 1, nwy = wy1 - wy0 + 1;
            for (int i = tid; i < nwx * nwy; i += PASS_THREADS)
                atomicAdd(&mw[((wx0 + i / nwy) % nb) * nb + ((wy0 + i % nwy) % nb)], 1u);
            return;
        }
        for (int wi = tid; wi < c.ex * ew; wi += PASS_THREADS) {
            const uint32_t bits = blk_bits[wi];
            if (bits == 0u) continue;
            const int lx = wi / ew, ly0 = (wi - lx * ew) << 5;
            const int bxi = (ss_wrap1(c.gx0 + lx, n) >> 5) * nb;
            const int ya = ss_wrap1(c.gy0 + ly0, n) >> 5, yb = ss_wrap1(c.gy0 + min(ly0 + 31, c.ey - 1), n) >> 5;
            atomicAdd(&mw[bxi + ya], 1u);
            if (yb != ya) atomicAdd(&mw[bxi + yb], 1u);
        }
    };
    const long long t_start = PASS_CLOCK();
    long long t_wait = 0, t_turn = 0, t_m;
    int n_turns = 0, n_blocked_me = 0;
    // ---- stage 1: copy occupancy words + int(sugar) into shared memory ----
    if (g.align == 4) {           // 128-bit / 32-bit vector loads, 4 cells per request, 4 requests in flight
        const int gpr = c.ey >> 2, total = c.ex * gpr;
        for (int base = 0; base < total; base += 4 * PASS_THREADS) {
            uint4 ov[4];
            uchar4 sv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = base + u * PASS_THREADS + tid;
                if (idx < total) {
                    const int lx = idx / gpr, ly = (idx - lx * gpr) << 2;
                    const size_t gi = (size_t)ss_wrap1(c.gx0 + lx, n) * n + ss_wrap1(c.gy0 + ly, n);
                    ov[u] = __ldcg(reinterpret_cast<const uint4*>(w.occw + gi));
                    sv[u] = __ldcg(reinterpret_cast<const uchar4*>(w.isugar + gi));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = base + u * PASS_THREADS + tid;
                if (idx < total) {
                    const int lx = idx / gpr, ly = (idx - lx * gpr) << 2;
                    uint32_t* o = c.occ_s + lx * c.emax + ly;
                    o[0] = ov[u].x; o[1] = ov[u].y; o[2] = ov[u].z; o[3] = ov[u].w;
                    uint8_t* q = c.isug_s + lx * c.emax + ly;
                    q[0] = sv[u].x; q[1] = sv[u].y; q[2] = sv[u].z; q[3] = sv[u].w;
                }
            }
        }
    } else {
        for (int idx = tid; idx < c.ex * c.ey; idx += PASS_THREADS) {
            const int lx = idx / c.ey, ly = idx - lx * c.ey;
            const size_t gi = (size_t)ss_wrap1(c.gx0 + lx, n) * n + ss_wrap1(c.gy0 + ly, n);
            c.occ_s[lx * c.emax + ly] = __ldcg(&w.occw[gi]);
            c.isug_s[lx * c.emax + ly] = __ldcg(&w.isugar[gi]);
        }
    }
    RegionShared& me = regs[warp];
    if (lane == 0) { me.n_list = 0; me.done = 0; sh.count[warp] = 0; }
    me.bfront[lane] = PRIO_INF;
    for (int i = tid; i < g.emax * ew; i += PASS_THREADS) blk_bits[i] = 0u;
    __syncthreads();
    const long long t_loaded = PASS_CLOCK();
    // ---- stage 2 (all threads, one pass over the window): pending agents -> their region's list
    //      (core) or blocked set (ring); per-block frontiers start at the block's lowest priority ----
    const uint64_t skey = ss_step_key(w.cfg.keyed_seed, iteration);
    const OrderKey ok = ss_order_key(skey, w.half);
    auto rg_of_x = [&](int x) { int r = (x * PASS_RGX) / c.ex; while (r > 0 && (r * c.ex) / PASS_RGX > x) r--; while (r + 1 < PASS_RGX && ((r + 1) * c.ex) / PASS_RGX <= x) r++; return r; };
    auto rg_of_y = [&](int y) { int r = (y * PASS_RGY) / c.ey; while (r > 0 && (r * c.ey) / PASS_RGY > y) r--; while (r + 1 < PASS_RGY && ((r + 1) * c.ey) / PASS_RGY <= y) r++; return r; };
    auto rg_x0 = [&](int ix) { return (ix * c.ex) / PASS_RGX; };
    auto rg_y0 = [&](int iy) { return (iy * c.ey) / PASS_RGY; };
    auto enlist = [&](int lx, int ly, uint32_t ow) {
        const uint32_t p = ss_priority(ok, occ_slot(ow));
        const int ix = rg_of_x(lx), iy = rg_of_y(ly);
        RegionShared& rg = regs[iy * PASS_RGX + ix];
        if (in_core(c, lx, ly, R)) {
            const int at = atomicAdd((int*)&rg.n_list, 1);
            if (at < PASS_LIST) rg.list[at] = ((uint64_t)p << 16) | (uint64_t)((lx << 8) | ly);
            const int b = (((lx - rg_x0(ix)) >> 4) * PASS_BCOLS) + ((ly - rg_y0(iy)) >> 4);
            atomicMin((uint32_t*)&rg.bfront[b], p);
        } else {                                            // ring agents belong to a later launch
            atomicOr(&blk_bits[lx * ew + (ly >> 5)], 1u << (ly & 31));
        }
    };
    if (g.align == 4 && (c.emax & 3) == 0) {                // four cells per 128-bit shared-memory load
        const int gpr = c.ey >> 2, total = c.ex * gpr;
        for (int idx = tid; idx < total; idx += PASS_THREADS) {
            const int lx = idx / gpr, ly = (idx - lx * gpr) << 2;
            const uint4 o4 = *reinterpret_cast<const uint4*>(c.occ_s + lx * c.emax + ly);
            if (o4.x != 0u && (o4.x >> 31) == cur_phase) enlist(lx, ly, o4.x);
            if (o4.y != 0u && (o4.y >> 31) == cur_phase) enlist(lx, ly + 1, o4.y);
            if (o4.z != 0u && (o4.z >> 31) == cur_phase) enlist(lx, ly + 2, o4.z);
            if (o4.w != 0u && (o4.w >> 31) == cur_phase) enlist(lx, ly + 3, o4.w);
        }
    } else {
        for (int idx = tid; idx < c.ex * c.ey; idx += PASS_THREADS) {
            const int lx = idx / c.ey, ly = idx - lx * c.ey;
            const uint32_t ow = c.occ_s[lx * c.emax + ly];
            if (ow != 0u && (ow >> 31) == cur_phase) enlist(lx, ly, ow);
        }
    }
    __syncthreads();
    // ---- my region ----
    const int rgx = warp % PASS_RGX, rgy = warp / PASS_RGX;
    const int rx0 = rg_x0(rgx), rx1 = rg_x0(rgx + 1), ry0 = rg_y0(rgy), ry1 = rg_y0(rgy + 1);
    uint32_t thr = PRIO_INF;              // listed agents have priority < thr
    int n_mine = me.n_list;
    if (n_mine > PASS_LIST) {
        // too many for one launch: keep the lower priorities (their possible blockers are lower still);
        // the unlisted agents keep their blocks' frontiers finite, so neighbours above them give up
        // bisection on the priority threshold: largest tested thr whose count fits the list
        const int rw = ry1 - ry0, cells = (rx1 - rx0) * rw;
        auto fill = [&](uint32_t t) {
            int cnt = 0;
            for (int base = 0; base < cells; base += 32) {
                const int i = base + lane;
                bool want = false;
                uint64_t key = 0;
                if (i < cells) {
                    const int lx = rx0 + i / rw, ly = ry0 + i % rw;
                    const uint32_t ow = c.occ_s[lx * c.emax + ly];
                    if (ow != 0u && (ow >> 31) == cur_phase && in_core(c, lx, ly, R)) {
                        const uint32_t p = ss_priority(ok, occ_slot(ow));
                        if (p < t) { want = true; key = ((uint64_t)p << 16) | (uint64_t)((lx << 8) | ly); }
                    }
                }
                const unsigned m = __ballot_sync(FULL, want);
                if (want) {
                    const int at = cnt + __popc(m & ((1u << lane) - 1u));
                    if (at < PASS_LIST) me.list[at] = key;
                }
                cnt += __popc(m);
            }
            return cnt;
        };
        uint32_t lo = 0u, hi = 1u << (2 * w.half);
        while (hi - lo > 1u) {
            const uint32_t mid = lo + (hi - lo) / 2u;
            if (fill(mid) <= PASS_LIST) lo = mid; else hi = mid;
        }
        thr = lo;
        n_mine = fill(thr);
        if (lane == 0) sh.dirty = 1;
        __syncwarp();
    }
    uint64_t keys[PASS_KPL];
#pragma unroll
    for (int k = 0; k < PASS_KPL; k++) {
        const int at = k * 32 + lane;
        keys[k] = at < n_mine ? me.list[at] : ~0ull;
    }
    // a window without a single resolvable agent leaves right away
    if (lane == 0 && n_mine > 0) atomicAdd(&sh.count[0], 1);
    __syncthreads();
    if (sh.count[0] == 0) { write_map(); return; }
    const long long t_listed = PASS_CLOCK();
    ThreadStats ts = {0, 0, 0, 0, 0, 0, 0, 0ull, 0ull, 0ull, 0ull};
    // ---- the region's agents in ascending priority ----
    for (int iter = 0;; iter++) {
        if (iter > PASS_LIST + 8) { if (lane == 0) { atomicAdd(&w.acc->fault, 1ull); w.acc->wd_info[2] = 0x2222; } break; }
        uint64_t km = keys[0];
#pragma unroll
        for (int k = 1; k < PASS_KPL; k++) km = keys[k] < km ? keys[k] : km;
        uint64_t kmin = km;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const uint64_t t = __shfl_xor_sync(FULL, kmin, o); kmin = t < kmin ? t : kmin; }
        if (kmin == ~0ull) break;
        const uint32_t pa = (uint32_t)(kmin >> 16);
        const int lx = (int)((kmin >> 8) & 255u), ly = (int)(kmin & 255u);
        const uint32_t ow = c.occ_s[lx * c.emax + ly];
        const int a = occ_vision(ow);
        bool blocked = false;
        // conflict zone of A: box |d| <= V plus the arm extensions a + V.  Wait until every block of
        // another region that the zone touches has its frontier above priority(A).
        unsigned reach = 0u;
        {
            t_m = PASS_CLOCK();
            const int zx0 = lx - V, zx1 = lx + V, zy0 = ly - V, zy1 = ly + V;
            const int ax0 = lx - a - V, ax1 = lx + a + V, ay0 = ly - a - V, ay1 = ly + a + V;
            const bool torus = c.full_x || c.full_y;
            int stuck = 0;
            // lane y < PASS_WARPS tests region y; then only the touched regions are visited
            unsigned touched;
            {
                bool touch = false;
                if (lane < PASS_WARPS && lane != warp) {
                    const int yx0 = rg_x0(lane % PASS_RGX), yx1 = rg_x0(lane % PASS_RGX + 1) - 1;
                    const int yy0 = rg_y0(lane / PASS_RGX), yy1 = rg_y0(lane / PASS_RGX + 1) - 1;
                    touch = torus || (zx0 <= yx1 && zx1 >= yx0 && zy0 <= yy1 && zy1 >= yy0) ||
                            (ax0 <= yx1 && ax1 >= yx0 && ly >= yy0 && ly <= yy1) ||
                            (lx >= yx0 && lx <= yx1 && ay0 <= yy1 && ay1 >= yy0);
                }
                touched = __ballot_sync(FULL, touch);
            }
            while (touched) {
                const int y = __ffs(touched) - 1;
                touched &= touched - 1;
                const int yx0 = rg_x0(y % PASS_RGX), yx1 = rg_x0(y % PASS_RGX + 1) - 1;
                const int yy0 = rg_y0(y / PASS_RGX), yy1 = rg_y0(y / PASS_RGX + 1) - 1;
                reach |= 1u << y;
                // lane = block of region y
                const int bxl = yx0 + ((lane / PASS_BCOLS) << 4), byl = yy0 + ((lane % PASS_BCOLS) << 4);
                const int bxh = min(bxl + 15, yx1), byh = min(byl + 15, yy1);
                bool mine = bxl <= yx1 && byl <= yy1 && lane < PASS_BCOLS * PASS_BROWS;
                if (mine && !torus)
                    mine = (zx0 <= bxh && zx1 >= bxl && zy0 <= byh && zy1 >= byl) ||
                           (ax0 <= bxh && ax1 >= bxl && ly >= byl && ly <= byh) ||
                           (lx >= bxl && lx <= bxh && ay0 <= byh && ay1 >= byl);
                unsigned bm = __ballot_sync(FULL, mine);
                if (lane == 0) {                      // a single lane spins; the others wait at the shuffle below
                    while (bm) {
                        const int b = __ffs(bm) - 1;
                        bm &= bm - 1;
                        const volatile uint32_t* f = &regs[y].bfront[b];
                        int spins = 0;
                        while (*f <= pa) {
                            if (regs[y].done) { if (*f <= pa) stuck = 1; break; }    // trimmed list: stopped below pa
                            __nanosleep(40);
                            if (++spins > 20000) {     // watchdog (~10 ms): never hang the GPU; retry in a later launch
                                stuck = 1;
                                if (atomicAdd(&w.acc->watchdog, 1ull) == 0ull) {
                                    w.acc->wd_info[0] = ((unsigned long long)warp << 48) | ((unsigned long long)y << 40) | ((unsigned long long)b << 32) | pa;
                                    w.acc->wd_info[1] = ((unsigned long long)*f << 32) | (unsigned)regs[y].done;
                                    w.acc->wd_info[2] = ((unsigned long long)blockIdx.x << 32) | (unsigned)((lx << 8) | ly);
                                }
                                break;
                            }
                        }
                        if (stuck) break;
                    }
                }
                stuck = __shfl_sync(FULL, stuck, 0);
                if (stuck) break;
            }
            blocked = stuck != 0;
            __threadfence_block();
            t_wait += PASS_CLOCK() - t_m;
        }
        // exact conflict test against the pending agents this launch cannot resolve (ring agents,
        // unresolved Q1 successors, and everything already blocked by those): lane l < 2V+1 owns zone
        // row dx = l - V of the blocked bitmap, lane 31 the same-row arm extensions, lanes < 2a the
        // same-column arm extension cells
        if (!blocked) {
            bool hit = false;
            auto consider = [&](int r, int q, int dx, int dy) {
                const uint32_t bw = c.occ_s[r * c.emax + q];
                if (bw == 0u || (r == lx && q == ly)) return;
                if (!crosses_meet(dx, dy, a, occ_vision(bw))) return;
                if (ss_priority(ok, occ_slot(bw)) < pa) hit = true;
            };
            auto bits_of = [&](int r, int c0, int len) {
                uint32_t out = 0u;
                if (c0 >= 0 && c0 + len <= c.ey) {
                    const volatile uint32_t* row = blk_bits + r * ew;
                    const int w0 = c0 >> 5, shf = c0 & 31;
                    const uint32_t lo = row[w0], hi = (w0 + 1 < ew) ? row[w0 + 1] : 0u;
                    out = __funnelshift_r(lo, hi, shf) & (len >= 32 ? 0xffffffffu : ((1u << len) - 1u));
                } else {
                    for (int i = 0; i < len; i++) {
                        int cc = c0 + i;
                        if (c.full_y) cc = ss_wrap1(cc, c.ey); else if (cc < 0 || cc >= c.ey) continue;
                        out |= ((((const volatile uint32_t*)blk_bits)[r * ew + (cc >> 5)] >> (cc & 31)) & 1u) << i;
                    }
                }
                return out;
            };
            if (lane < 2 * V + 1) {
                const int dx = lane - V, r = wrap_row(c, lx + dx);
                if (r >= 0 && r < c.ex) {
                    uint32_t bits = bits_of(r, ly - V, 2 * V + 1);
                    while (bits) { const int i = __ffs(bits) - 1; bits &= bits - 1; consider(r, wrap_col(c, ly + i - V), dx, i - V); }
                }
            } else if (lane == 31) {
                uint32_t lo = bits_of(lx, ly - V - a, a);
                while (lo) { const int i = __ffs(lo) - 1; lo &= lo - 1; const int dy = i - V - a; consider(lx, wrap_col(c, ly + dy), 0, dy); }
                uint32_t hi = bits_of(lx, ly + V + 1, a);
                while (hi) { const int i = __ffs(hi) - 1; hi &= hi - 1; const int dy = V + 1 + i; consider(lx, wrap_col(c, ly + dy), 0, dy); }
            }
            if (lane < 2 * a) {
                const int dx = (lane & 1) ? (V + 1 + (lane >> 1)) : -(V + 1 + (lane >> 1));
                const int r = wrap_row(c, lx + dx);
                if (r >= 0 && r < c.ex && ((((const volatile uint32_t*)blk_bits)[r * ew + (ly >> 5)] >> (ly & 31)) & 1u)) consider(r, ly, dx, 0);
            }
            blocked = __any_sync(FULL, hit);
        }
        uint32_t slot = 0;
        bool skipped = false;
        if (!blocked) {
            slot = occ_slot(ow);
            if (ow & OCC_Q1) {      // Q1: the at-risk predecessor decides whether A gets its turn
                const AgentRec* rec = &w.agents[slot];
                if (rec->pred_step == (uint32_t)(iteration + 1)) {
                    const volatile uint32_t* stp = &w.agents[rec->pred - 1u].status;
                    uint32_t st = *stp;
                    for (int tries = 0; tries < 4 && (st >> 2) != (uint32_t)(iteration + 1); tries++) { __nanosleep(250); st = *stp; }
                    if ((st >> 2) != (uint32_t)(iteration + 1)) blocked = true;     // retry in a later launch
                    else skipped = ((st & 3u) == ST_DIED);
                }
            }
        }
        if (blocked) {
            n_blocked_me++;
            if (lane == 0) atomicOr(&blk_bits[lx * ew + (ly >> 5)], 1u << (ly & 31));
        } else {
            t_m = PASS_CLOCK();
            if (skipped) {
                if (lane == 0) {      // loses its turn: stays put, no metabolism (Q1)
                    const uint32_t nw = occ_pack(slot, a, next_phase);
                    c.occ_s[lx * c.emax + ly] = nw;
                    const int gx = ss_wrap1(c.gx0 + lx, n), gy = ss_wrap1(c.gy0 + ly, n);
                    w.occw[(size_t)gx * n + gy] = nw;
                    if (ow & OCC_RISK)
                        *((volatile uint32_t*)&w.agents[slot].status) = ((uint32_t)(iteration + 1) << 2) | ST_SKIPPED;
                    ts.resolved++;
                    if (w.log) {
                        ShardLogEntry en;
                        en.slot = slot; en.old_cell = en.new_cell = (uint32_t)(gx * n + gy); en.new_word = nw;
                        en.sugar = 0.0; en.poll = 0.0; en.attr = 0u; en.turn = 0u; en.kind_harvest = 2u;
                        en.status = (ow & OCC_RISK) ? (((uint32_t)(iteration + 1) << 2) | ST_SKIPPED) : 0u;
                        *log_slot(w, ts) = en;
                    }
                }
            } else {
                CrossLoads cl;
                if (w.cfg.rule_move_eat) warp_load_cross<-1>(w, c, lx, ly, ss_wrap1(c.gx0 + lx, n), ss_wrap1(c.gy0 + ly, n), a, lane, cl);
                warp_agent_turn<-1>(w, c, lx, ly, ow & ~OCC_Q1, slot, iteration, env_time, skey, next_phase, ts, lane, cl);
                n_turns++;
            }
            t_turn += PASS_CLOCK() - t_m;
        }
        // A is finished (or parked in the blocked set): advance the frontier of its block
        __syncwarp();
        __threadfence_block();
        {
            const int bx_a = (lx - rx0) >> 4, by_a = (ly - ry0) >> 4;
            uint32_t mn = PRIO_INF;
#pragma unroll
            for (int k = 0; k < PASS_KPL; k++) {
                if (keys[k] == kmin) keys[k] = ~0ull;
                if (keys[k] != ~0ull) {
                    const int kx = (int)((keys[k] >> 8) & 255u), ky = (int)(keys[k] & 255u);
                    if (((kx - rx0) >> 4) == bx_a && ((ky - ry0) >> 4) == by_a) {
                        const uint32_t kp = (uint32_t)(keys[k] >> 16);
                        mn = kp < mn ? kp : mn;
                    }
                }
            }
            mn = __reduce_min_sync(FULL, mn);
            if (thr != PRIO_INF && mn == PRIO_INF) mn = thr;      // unlisted agents may sit in this block
            if (lane == 0) me.bfront[bx_a * PASS_BCOLS + by_a] = mn;
        }
    }
    if (lane == 0) {
        __threadfence_block(); me.done = 1;
        close_log_slices(w, ts);
    }
    __syncthreads();
    write_map();
    if (w.dbg) {
        if (lane == 0) {
            atomicAdd(&w.dbg[9], (unsigned long long)n_turns);
            atomicAdd(&w.dbg[12], (unsigned long long)n_blocked_me);
            atomicAdd(&w.dbg[15], (unsigned long long)t_turn);
            atomicAdd(&w.dbg[16], (unsigned long long)t_wait);
            atomicAdd(&w.dbg[13], (unsigned long long)n_mine);
        }
        if (tid == 0) {
            atomicAdd(&w.dbg[0], 1ull);                                   // windows with work
            atomicAdd(&w.dbg[3], (unsigned long long)(t_listed - t_start));
            atomicAdd(&w.dbg[11], (unsigned long long)(t_loaded - t_start));
            atomicAdd(&w.dbg[6], (unsigned long long)(PASS_CLOCK() - t_start));
        }
    }
    // ---- block reduction of the step accumulators ----
    unsigned vals[7] = {ts.sum_m, ts.sum_v, ts.acted, ts.deaths, ts.resolved, ts.small, ts.inexact};
    unsigned long long* dst[7] = {&w.acc->sum_metab, &w.acc->sum_vision, &w.acc->acted, &w.acc->deaths,
                                  nullptr, &w.acc->hist_small, &w.acc->inexact};
#pragma unroll
    for (int q = 0; q < 7; q++) {
        unsigned v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) sh.red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            unsigned long long s = 0;
            for (int i = 0; i < PASS_WARPS; i++) s += sh.red[i];
            if (s) {
                if (q == 4) atomicAdd(&w.acc->pending, (unsigned long long)(-(long long)s));
                else atomicAdd(dst[q], s);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Solo move pass: ONE WARP owns one whole window, so a launch has no inter-warp dependency at all -- the warp
// takes the pending agents of the window's core in ascending priority, one after the other, and every conflict
// partner of lower priority is either earlier in its own list or in the blocked set (ring agents, unresolved Q1
// successors and whatever those block).  Cells are read straight from the lattice (L2/HBM) instead of a staged
// shared-memory image, which leaves room for 32 resident warps per SM, each with an independent window: the
// latency of one agent's turn is hidden by the other windows instead of stalling a CTA.  Shared memory per warp:
// the sorted key list and the blocked bitmap.
#ifndef SOLO_LIST
#define SOLO_LIST 1024
#endif
#ifndef SOLO_SCAN_ROWS
#define SOLO_SCAN_ROWS 8
#endif
__device__ static inline void ss_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifdef SS_PASS_TIMING       // phase clocks of the warp-per-window kernels -> w.dbg[17..31] (tools/dbg_solo.py)
#define SOLO_TICK(acc) do { const long long t_now_ = clock64(); (acc) += t_now_ - t_mark; t_mark = t_now_; } while (0)
#else
#define SOLO_TICK(acc) do { } while (0)
#endif
#ifndef SOLO_MIN_CTAS
#define SOLO_MIN_CTAS 25
#endif
template <int POLT, bool WRAP, bool LANES, bool UTIL>
__device__ __forceinline__ void solo_window(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time,
                                            uint32_t* smem, int& cnt_s, int gx0, int gy0, int ex, int ey) {
    typedef GlobCtx<WRAP> GC;
    const int n = w.n, lane = threadIdx.x;
    const int V = w.vmax, R = g.ring;
    GC c;
    c.occw = w.occw; c.isugar = w.isugar;
    c.n = n; c.ex = ex; c.ey = ey; c.emax = g.emax;
    c.gx0 = gx0; c.gy0 = gy0;
    uint32_t* keys = smem;                                  // [SOLO_LIST] priority << 7 | vision + flags (sort key)
    uint32_t* blk_bits = smem + SOLO_LIST;                  // [emax][ew]
    uint16_t* poss = (uint16_t*)(blk_bits + g.emax * g.ew); // [SOLO_LIST] lx << 8 | ly, permuted with the keys
    uint32_t* shd_bits = (uint32_t*)(poss + SOLO_LIST);     // [emax][ew] lane-per-agent loop: crosses of the core agents blocked in this launch
    (void)shd_bits;
    UtilShared* us = (UtilShared*)(((uintptr_t)(poss + SOLO_LIST) + 7u) & ~(uintptr_t)7u);      // utilicalc rule: moves / neighbours
    (void)us;
    const int ew = g.ew;
    const uint32_t cur_phase = (uint32_t)(iteration & 1), next_phase = cur_phase ^ 1u;
    const int nb = w.pend_nb;
    if (g.map_read >= 0) {          // coarse pending map: nothing left in this window's core -> hand the entries on
        const uint32_t* mr = w.pendmap + (size_t)g.map_read * nb * nb;
        const int Rm = g.ring_per_agent ? V + 1 : R;    // core of the smallest vision
        const int cx0 = (c.gx0 + Rm) >> 5, cx1 = (c.gx0 + c.ex - Rm - 1) >> 5;
        const int cy0 = (c.gy0 + Rm) >> 5, cy1 = (c.gy0 + c.ey - Rm - 1) >> 5;
        const int ncx = cx1 - cx0 + 1, ncy = cy1 - cy0 + 1;
        int any = 0;
        for (int i = lane; i < ncx * ncy; i += 32)
            any |= mr[((cx0 + i / ncy) % nb) * nb + ((cy0 + i % ncy) % nb)] != 0u;
        if (!__any_sync(FULL, any)) {
            if (g.map_write >= 0) {
                uint32_t* mw = w.pendmap + (size_t)g.map_write * nb * nb;
                const int wx0 = c.gx0 >> 5, wx1 = (c.gx0 + c.ex - 1) >> 5, wy0 = c.gy0 >> 5, wy1 = (c.gy0 + c.ey - 1) >> 5;
                const int nwx = wx1 - wx0 + 1, nwy = wy1 - wy0 + 1;
                for (int i = lane; i < nwx * nwy; i += 32) {
                    const int bi = ((wx0 + i / nwy) % nb) * nb + ((wy0 + i % nwy) % nb);
                    const uint32_t v = mr[bi];
                    if (v) atomicAdd(&mw[bi], v);
                }
            }
            return;
        }
    }
    long long t_mark = PASS_CLOCK(), t_scan = 0, t_sort = 0, t_blk = 0, t_dep = 0, t_q1 = 0, t_act = 0, t_fill = 0;
    unsigned n_rounds = 0, n_blocked = 0, n_lane_slots = 0;
    (void)t_mark; (void)t_scan; (void)t_sort; (void)t_blk; (void)t_dep; (void)t_q1; (void)t_act; (void)t_fill;
    (void)n_rounds; (void)n_blocked; (void)n_lane_slots;
    bool dirty = false;
    auto write_map = [&]() {
        if (g.map_write < 0) return;
        uint32_t* mw = w.pendmap + (size_t)g.map_write * nb * nb;
        if (dirty) {
            const int wx0 = c.gx0 >> 5, wx1 = (c.gx0 + c.ex - 1) >> 5, wy0 = c.gy0 >> 5, wy1 = (c.gy0 + c.ey - 1) >> 5;
            const int nwx = wx1 - wx0 + 1, nwy = wy1 - wy0 + 1;
            for (int i = lane; i < nwx * nwy; i += 32)
                atomicAdd(&mw[((wx0 + i / nwy) % nb) * nb + ((wy0 + i % nwy) % nb)], 1u);
            return;
        }
        for (int wi = lane; wi < c.ex * ew; wi += 32) {
            const uint32_t bits = blk_bits[wi];
            if (bits == 0u) continue;
            const int lx = wi / ew, ly0 = (wi - lx * ew) << 5;
            const int bxi = (ss_wrap1(c.gx0 + lx, n) >> 5) * nb;
            const int ya = ss_wrap1(c.gy0 + ly0, n) >> 5, yb = ss_wrap1(c.gy0 + min(ly0 + 31, c.ey - 1), n) >> 5;
            atomicAdd(&mw[bxi + ya], 1u);
            if (yb != ya) atomicAdd(&mw[bxi + yb], 1u);
        }
    };
    // ---- scan the window: pending core agents -> key list, pending ring agents -> blocked set ----
    // (one copy of the per-cell code, rolled loops: the kernel's instruction footprint is what limits it)
    const uint64_t skey = ss_step_key(w.cfg.keyed_seed, iteration);
    const OrderKey ok = ss_order_key(skey, w.half);
    for (int i = lane; i < g.emax * ew; i += 32) blk_bits[i] = 0u;
    if (LANES) for (int i = lane; i < g.emax * ew; i += 32) shd_bits[i] = 0u;
    if (lane == 0) cnt_s = 0;
    __syncwarp();
    uint32_t pmin = PRIO_INF;       // lowest pending priority of the core (per lane)
    uint32_t thr = PRIO_INF, p0 = 0u, span = 0u;
    bool first = true;
    int n_mine;
    const int per = g.align == 4 ? 4 : 1;                   // cells per load
    const int gpr = c.ey / per;                             // loads per window row
    // one group of `per` cells of window row lx, starting at column ly0
    auto visit = [&](int lx, int ly0, const uint4& o) {
        unsigned m = (o.x != 0u && (o.x >> 31) == cur_phase ? 1u : 0u) | (o.y != 0u && (o.y >> 31) == cur_phase ? 2u : 0u) |
                     (o.z != 0u && (o.z >> 31) == cur_phase ? 4u : 0u) | (o.w != 0u && (o.w >> 31) == cur_phase ? 8u : 0u);
        if (m == 0u) return;
        // an agent of vision a is core iff it stands >= a + vmax from every window edge: then no agent outside the
        // window can share a cell with its cross (R = 2 vmax is that bound for the largest vision)
        uint32_t ringbits = 0u;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t ow = j == 0 ? o.x : j == 1 ? o.y : j == 2 ? o.z : o.w;
            const int ly = ly0 + j;
            const int Ra = g.ring_per_agent ? (int)((ow >> 24) & 31u) + V : R;
            if (!(lx >= Ra && lx < c.ex - Ra && ly >= Ra && ly < c.ey - Ra)) { ringbits |= 1u << (ly & 31); continue; }
            uint32_t key = ow;                                      // first scan: hashed below
            if (!first) {
                const uint32_t p = ss_priority(ok, occ_slot(ow));
                if (p >= thr) continue;
                key = (p << 7) | ((ow >> 24) & 0x7fu);
            }
            const int at = atomicAdd(&cnt_s, 1);
            if (at < SOLO_LIST) { keys[at] = key; poss[at] = (uint16_t)((lx << 8) | ly); }
        }
        if (first && ringbits) atomicOr(&blk_bits[lx * ew + (ly0 >> 5)], ringbits);
    };
    for (;;) {
        // SOLO_SCAN_ROWS window rows per iteration (as many loads in flight per lane), lanes across the row
#pragma unroll 1
        for (int lx0 = 0; lx0 < c.ex; lx0 += SOLO_SCAN_ROWS) {
#pragma unroll 1
            for (int g0 = 0; g0 < gpr; g0 += 32) {
                const int grp = g0 + lane, ly0 = grp * per;
                const int gyw = GC::gw(c.gy0 + ly0, n);
                uint4 ov[SOLO_SCAN_ROWS];
#pragma unroll
                for (int u = 0; u < SOLO_SCAN_ROWS; u++) {
                    ov[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (grp < gpr && lx0 + u < c.ex) {
                        const size_t gi = (size_t)GC::gw(c.gx0 + lx0 + u, n) * n + gyw;
                        if (per == 4) ov[u] = __ldcg(reinterpret_cast<const uint4*>(w.occw + gi));
                        else ov[u].x = __ldcg(&w.occw[gi]);
                    }
                }
#pragma unroll
                for (int u = 0; u < SOLO_SCAN_ROWS; u++) visit(lx0 + u, ly0, ov[u]);
            }
        }
        __syncwarp();
        n_mine = *((volatile int*)&cnt_s);
        if (n_mine <= SOLO_LIST) break;
        // too many for one launch: keep the lowest priorities (nothing they depend on has a higher one); the
        // unlisted agents stay pending, and the window reports itself dirty to the pending map.  Pending
        // priorities are spread over [lowest pending, domain): threshold for ~13/16 of the list, halved until it fits
        if (first) {
            dirty = true;
            first = false;
            // lowest priority among the SOLO_LIST agents that made it into the list (>= the lowest pending one)
            for (int i = lane; i < SOLO_LIST; i += 32) {
                const uint32_t p = ss_priority(ok, occ_slot(keys[i]));
                pmin = p < pmin ? p : pmin;
            }
            p0 = __reduce_min_sync(FULL, pmin);
            span = (uint32_t)((((unsigned long long)((1u << (2 * w.half)) - p0) * SOLO_LIST) / (unsigned long long)n_mine) * 13ull / 16ull);
        } else {
            span >>= 1;
        }
        if (span < 1u) span = 1u;
        thr = p0 + span;
        __syncwarp();
        if (lane == 0) cnt_s = 0;
        __syncwarp();
    }
    SOLO_TICK(t_scan);
    if (n_mine == 0) { write_map(); return; }
    if (first) {        // every pending core agent is listed: hash them now, 32 at a time
        for (int i = lane; i < n_mine; i += 32) {
            const uint32_t ow = keys[i];
            keys[i] = (ss_priority(ok, occ_slot(ow)) << 7) | ((ow >> 24) & 0x7fu);
        }
        __syncwarp();
    }
    {   // bitonic sort of the keys (unique), padded to a power of two
        int P = 32;
        while (P < n_mine) P <<= 1;
        for (int i = n_mine + lane; i < P; i += 32) keys[i] = 0xffffffffu;
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < (P >> 1); i += 32) {
                    const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                    const uint32_t x = keys[lo], y = keys[hi];
                    const bool up = (lo & k) == 0;
                    if ((x > y) == up) {
                        keys[lo] = y; keys[hi] = x;
                        const uint16_t t = poss[lo]; poss[lo] = poss[hi]; poss[hi] = t;
                    }
                }
                __syncwarp();
            }
        }
    }
    SOLO_TICK(t_sort);
    ThreadStats ts = {0, 0, 0, 0, 0, 0, 0, 0ull, 0ull, 0ull, 0ull};
    uint32_t slot_batch = 0u;
    if (LANES) {
        // ---- one LANE per agent.  The warp holds the (up to) 32 unresolved list entries of lowest priority, in list
        // order.  An entry acts in this round iff no blocked agent of lower priority conflicts with it (exact test
        // against the blocked set, as below) and its cross meets the cross of no earlier entry of the batch; entries
        // that conflict inside the batch stay for the next round, blocked ones join the blocked set.  Lanes that act
        // in the same round have disjoint crosses, so their turns commute; the first entry of the batch never waits,
        // so every round resolves at least one entry.  Everything of lower priority outside the batch is finished
        // or blocked -- the same invariant as the one-agent-at-a-time loop.
        int my = -1, next = 0;
#pragma unroll 1
        for (;;) {
            {   // order-preserving refill: waiting entries move to the low lanes, new entries follow
                const unsigned wm = __ballot_sync(FULL, my >= 0);
                const int nwait = __popc(wm);
                const int src = (int)__fns(wm, 0u, lane + 1) & 31;
                const int moved = __shfl_sync(FULL, my, src);
                const int fresh = next + lane - nwait;
                my = lane < nwait ? moved : (fresh < n_mine ? fresh : -1);
                next = min(n_mine, next + 32 - nwait);
            }
            if (!__any_sync(FULL, my >= 0)) break;
            const bool have = my >= 0;
            SOLO_TICK(t_fill);
#ifdef SS_PASS_TIMING
            n_rounds++; n_lane_slots += __popc(__ballot_sync(FULL, have));
#endif
            const uint32_t key = have ? keys[my] : 0u;
            const uint32_t pa = key >> 7;
            const int lx = have ? (int)(poss[my] >> 8) : 0, ly = have ? (int)(poss[my] & 255u) : 0;
            const int a = (int)((key >> 0) & 31u);
            const int gx = GC::gw(c.gx0 + lx, n), gy = GC::gw(c.gy0 + ly, n);
            const uint32_t info = have ? ((uint32_t)lx | ((uint32_t)ly << 8) | ((uint32_t)a << 16)) : 0xffffffffu;
            bool blocked = false, dep = false;
            {   // conflicts inside the batch: lane m < lane with meeting crosses (four shuffles in flight)
                const int nb_lanes = __popc(__ballot_sync(FULL, have));        // entries sit in lanes 0 .. nb_lanes-1
#pragma unroll 1
                for (int m0 = 0; m0 + 1 < nb_lanes; m0 += 4) {
                    uint32_t im[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) im[u] = __shfl_sync(FULL, info, (m0 + u) & 31);
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (m0 + u < lane && have)
                            dep |= crosses_meet((int)(im[u] & 255u) - lx, (int)((im[u] >> 8) & 255u) - ly, a, (int)(im[u] >> 16));
                }
            }
            SOLO_TICK(t_dep);
            const bool cand = have && !dep;
            if (cand && w.cfg.rule_move_eat) {      // the cross is wanted soon: start it on its way into L2
                const bool polp = POLT != 0;
#pragma unroll 1
                for (int d = 1; d <= a; d++) {
                    const int ci0 = GC::gw(gx - d, n) * n + gy, ci1 = GC::gw(gx + d, n) * n + gy;
                    const int ci2 = gx * n + GC::gw(gy - d, n), ci3 = gx * n + GC::gw(gy + d, n);
                    ss_prefetch_l2(&w.occw[ci0]); ss_prefetch_l2(&w.occw[ci1]);
                    ss_prefetch_l2(&w.isugar[ci0]); ss_prefetch_l2(&w.isugar[ci1]);
                    if (d == 1 || ((gy - d) & 7) == 7) ss_prefetch_l2(&w.occw[ci2]);        // 8 words per sector along y
                    if (d == 1 || ((gy + d) & 7) == 0) ss_prefetch_l2(&w.occw[ci3]);
                    if (d == 1 || ((gy - d) & 31) == 31) ss_prefetch_l2(&w.isugar[ci2]);
                    if (d == 1 || ((gy + d) & 31) == 0) ss_prefetch_l2(&w.isugar[ci3]);
                    if (polp) {
                        ss_prefetch_l2(&w.poll[ci0]); ss_prefetch_l2(&w.poll[ci1]);
                        if (d == 1 || ((gy - d) & 3) == 3) ss_prefetch_l2(&w.poll[ci2]);
                        if (d == 1 || ((gy + d) & 3) == 0) ss_prefetch_l2(&w.poll[ci3]);
                    }
                }
            }
            // Conflict test against the blocked set.  Filter, per lane: A can only be blocked if it stands within reach
            // of the ring (ring agents are blocked from the start) or its cross touches the shadow -- the union of the
            // crosses of the core agents blocked during this launch.  Everything else is free without looking further.
            bool flagged = false;
            if (cand) {
                flagged = lx - a - V < R || lx + a + V >= c.ex - R || ly - a - V < R || ly + a + V >= c.ey - R;
                if (!flagged) {
                    const volatile uint32_t* row = shd_bits + lx * ew;
                    const int c0 = ly - a, w0 = c0 >> 5;
                    const uint32_t lo = row[w0], hi = (w0 + 1 < ew) ? row[w0 + 1] : 0u;
                    uint32_t hitbits = __funnelshift_r(lo, hi, c0 & 31) & ((2u << (2 * a)) - 1u);
                    const volatile uint32_t* col = shd_bits + (lx - a) * ew + (ly >> 5);
#pragma unroll 1
                    for (int o = 0; o <= 2 * a; o++) hitbits |= (col[o * ew] >> (ly & 31)) & 1u;
                    flagged = hitbits != 0u;
                }
            }
            // Exact test for the flagged lanes, one after the other with the whole warp (as in the one-at-a-time loop):
            // the zone of A cut into segments of one bitmap row each, a segment per lane
            for (unsigned fl = __ballot_sync(FULL, flagged); fl != 0u; fl &= fl - 1u) {
                const int f = __ffs(fl) - 1;
                const uint32_t iF = __shfl_sync(FULL, info, f), paF = __shfl_sync(FULL, pa, f);
                const int lxF = (int)(iF & 255u), lyF = (int)((iF >> 8) & 255u), aF = (int)(iF >> 16);
                bool hit = false;
                const int nseg = 2 * V + 3 + 2 * aF;
#pragma unroll 1
                for (int seg = lane; seg < nseg; seg += 32) {
                    int dx, c0, len, dy0;
                    if (seg <= 2 * V) { dx = seg - V; c0 = lyF - V; len = 2 * V + 1; dy0 = -V; }
                    else if (seg == 2 * V + 1) { dx = 0; c0 = lyF - V - aF; len = aF; dy0 = -V - aF; }
                    else if (seg == 2 * V + 2) { dx = 0; c0 = lyF + V + 1; len = aF; dy0 = V + 1; }
                    else { const int k = seg - (2 * V + 3); dx = (k & 1) ? (V + 1 + (k >> 1)) : -(V + 1 + (k >> 1)); c0 = lyF; len = 1; dy0 = 0; }
                    const int r = lxF + dx;
                    const volatile uint32_t* row = blk_bits + r * ew;
                    const int w0 = c0 >> 5, shf = c0 & 31;
                    const uint32_t lo = row[w0], hi = (w0 + 1 < ew) ? row[w0 + 1] : 0u;
                    uint32_t bits = __funnelshift_r(lo, hi, shf) & ((1u << len) - 1u);
                    while (bits) {
                        const int i = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int q = c0 + i, dy = dy0 + i;
                        if ((dx == 0 && dy == 0) || !crosses_meet(dx, dy, aF, V)) continue;
                        const uint32_t bw = __ldcg(&w.occw[cell_of(c, r, q)]);
                        if (bw != 0u && crosses_meet(dx, dy, aF, occ_vision(bw)) && ss_priority(ok, occ_slot(bw)) < paF) hit = true;
                    }
                }
                const bool any_hit = __any_sync(FULL, hit);
                if (lane == f) blocked = any_hit;
            }
            __syncwarp();
            SOLO_TICK(t_blk);
            const uint32_t slot = ss_priority_inverse(ok, pa);
            const uint32_t ow = (cur_phase << 31) | ((key & 0x7fu) << 24) | slot;
            bool skipped = false;
            const bool go = have && !blocked && !dep;
            if (go && (ow & OCC_Q1)) {      // Q1: the at-risk predecessor decides whether A gets its turn
                const AgentRec* rec = &w.agents[slot];
                if (rec->pred_step == (uint32_t)(iteration + 1)) {
                    const volatile uint32_t* stp = &w.agents[rec->pred - 1u].status;
                    uint32_t st = *stp;
                    for (int tries = 0; tries < 4 && (st >> 2) != (uint32_t)(iteration + 1); tries++) { __nanosleep(250); st = *stp; }
                    if ((st >> 2) != (uint32_t)(iteration + 1)) blocked = true;     // retry in a later launch
                    else skipped = ((st & 3u) == ST_DIED);
                }
            }
            __syncwarp();
            SOLO_TICK(t_q1);
#ifdef SS_PASS_TIMING
            n_blocked += __popc(__ballot_sync(FULL, have && blocked));
#endif
            if (have && blocked) {          // joins the blocked set; its cross joins the shadow
                atomicOr(&blk_bits[lx * ew + (ly >> 5)], 1u << (ly & 31));
                const int c0 = ly - a, w0 = c0 >> 5, shf = c0 & 31;
                const unsigned long long m = (unsigned long long)((2u << (2 * a)) - 1u) << shf;
                atomicOr(&shd_bits[lx * ew + w0], (uint32_t)m);
                if ((uint32_t)(m >> 32) != 0u) atomicOr(&shd_bits[lx * ew + w0 + 1], (uint32_t)(m >> 32));
#pragma unroll 1
                for (int o = -a; o <= a; o++) atomicOr(&shd_bits[(lx + o) * ew + (ly >> 5)], 1u << (ly & 31));
                my = -1;
            } else if (go && skipped) {     // loses its turn: stays put, no metabolism (Q1)
                const uint32_t nw = occ_pack(slot, a, next_phase);
                w.occw[(size_t)gx * n + gy] = nw;
                if (ow & OCC_RISK)
                    *((volatile uint32_t*)&w.agents[slot].status) = ((uint32_t)(iteration + 1) << 2) | ST_SKIPPED;
                ts.resolved++;
                if (w.log) {
                    ShardLogEntry en;
                    en.slot = slot; en.old_cell = en.new_cell = (uint32_t)(gx * n + gy); en.new_word = nw;
                    en.sugar = 0.0; en.poll = 0.0; en.attr = 0u; en.turn = 0u; en.kind_harvest = 2u;
                    en.status = (ow & OCC_RISK) ? (((uint32_t)(iteration + 1) << 2) | ST_SKIPPED) : 0u;
                    *log_slot(w, ts) = en;
                }
                my = -1;
            } else if (go) {
                const bool risky = (ow & OCC_RISK) != 0;
                double sugar = 0.0;
                uint32_t attr = 0u;
                if (risky) { sugar = w.agents[slot].sugar; attr = w.agents[slot].attr; }
                LaneChoice ch;
                if (w.cfg.rule_move_eat) ch = lane_choose_cell<POLT, WRAP>(w, gx, gy, a, slot, skey);
                else { ch.cell = gx * n + gy; ch.sg = 0; }
                commit_turn(w, c, lx, ly, lx, ly, gx * n + gy, ch.cell, ch.sg, a, risky, slot, sugar, attr, iteration, env_time,
                            next_phase, ts);
                my = -1;
            }
            __syncwarp();
            SOLO_TICK(t_act);
        }
    } else
    for (int it = 0; it < n_mine; it++) {
        // the key carries everything of the agent's occupancy word: the slot is the inverse of the priority (computed
        // for 32 list entries at a time, one per lane), and nobody else changes the word before the agent's own turn
        if ((it & 31) == 0) slot_batch = ss_priority_inverse(ok, keys[min(it + lane, n_mine - 1)] >> 7);
        const uint32_t key = keys[it];
        const uint32_t pa = key >> 7;
        const int lx = (int)(poss[it] >> 8), ly = (int)(poss[it] & 255u);
        const int gx = GC::gw(c.gx0 + lx, n), gy = GC::gw(c.gy0 + ly, n);
        const uint32_t ow = (cur_phase << 31) | (((uint32_t)key & 0x7fu) << 24) | __shfl_sync(FULL, slot_batch, it & 31);
        const int a = occ_vision(ow);
        bool blocked = false;
        // A's cross: loads issued now, consumed after the conflict test (wasted if A turns out blocked).  Issuing them
        // a whole turn earlier (when the previous agent's cross cannot meet A's) was measured and is slower.
        CrossLoads cl;
        if (w.cfg.rule_move_eat) warp_load_cross<POLT>(w, c, lx, ly, gx, gy, a, lane, cl);
        {   // exact conflict test against the blocked set.  The zone of A is cut into segments of one bitmap row
            // each -- 2V+1 rows of the box |d| <= V, the two same-row arm extensions, and the 2a cells of the
            // same-column arm extensions -- and every lane takes segments lane, lane + 32 (windows are never
            // full here, and the zone of a core agent lies inside the window)
            bool hit = false;
            const int nseg = 2 * V + 3 + 2 * a;
#pragma unroll 1
            for (int seg = lane; seg < nseg; seg += 32) {
                int dx, c0, len, dy0;
                if (seg <= 2 * V) { dx = seg - V; c0 = ly - V; len = 2 * V + 1; dy0 = -V; }
                else if (seg == 2 * V + 1) { dx = 0; c0 = ly - V - a; len = a; dy0 = -V - a; }
                else if (seg == 2 * V + 2) { dx = 0; c0 = ly + V + 1; len = a; dy0 = V + 1; }
                else { const int k = seg - (2 * V + 3); dx = (k & 1) ? (V + 1 + (k >> 1)) : -(V + 1 + (k >> 1)); c0 = ly; len = 1; dy0 = 0; }
                const int r = lx + dx;
                const volatile uint32_t* row = blk_bits + r * ew;
                const int w0 = c0 >> 5, shf = c0 & 31;
                const uint32_t lo = row[w0], hi = (w0 + 1 < ew) ? row[w0 + 1] : 0u;
                uint32_t bits = __funnelshift_r(lo, hi, shf) & ((1u << len) - 1u);
                while (bits) {
                    const int i = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int q = c0 + i, dy = dy0 + i;
                    if ((dx == 0 && dy == 0) || !crosses_meet(dx, dy, a, V)) continue;     // V >= the blocker's vision: no load needed
                    const uint32_t bw = __ldcg(&w.occw[cell_of(c, r, q)]);
                    if (bw != 0u && crosses_meet(dx, dy, a, occ_vision(bw)) && ss_priority(ok, occ_slot(bw)) < pa) hit = true;
                }
            }
            blocked = __any_sync(FULL, hit);
        }
        uint32_t slot = 0;
        bool skipped = false;
        if (!blocked) {
            slot = occ_slot(ow);
            if (ow & OCC_Q1) {      // Q1: the at-risk predecessor decides whether A gets its turn
                const AgentRec* rec = &w.agents[slot];
                if (rec->pred_step == (uint32_t)(iteration + 1)) {
                    const volatile uint32_t* stp = &w.agents[rec->pred - 1u].status;
                    uint32_t st = *stp;
                    for (int tries = 0; tries < 4 && (st >> 2) != (uint32_t)(iteration + 1); tries++) { __nanosleep(250); st = *stp; }
                    if ((st >> 2) != (uint32_t)(iteration + 1)) blocked = true;     // retry in a later launch
                    else skipped = ((st & 3u) == ST_DIED);
                }
            }
        }
        if (blocked) {
            if (lane == 0) atomicOr(&blk_bits[lx * ew + (ly >> 5)], 1u << (ly & 31));
        } else if (skipped) {
            if (lane == 0) {      // loses its turn: stays put, no metabolism (Q1)
                const uint32_t nw = occ_pack(slot, a, next_phase);
                w.occw[(size_t)gx * n + gy] = nw;
                if (ow & OCC_RISK)
                    *((volatile uint32_t*)&w.agents[slot].status) = ((uint32_t)(iteration + 1) << 2) | ST_SKIPPED;
                ts.resolved++;
                if (w.log) {
                    ShardLogEntry en;
                    en.slot = slot; en.old_cell = en.new_cell = (uint32_t)(gx * n + gy); en.new_word = nw;
                    en.sugar = 0.0; en.poll = 0.0; en.attr = 0u; en.turn = 0u; en.kind_harvest = 2u;
                    en.status = (ow & OCC_RISK) ? (((uint32_t)(iteration + 1) << 2) | ST_SKIPPED) : 0u;
                    *log_slot(w, ts) = en;
                }
            }
        } else if (UTIL) {
            warp_utilicalc_turn<WRAP>(w, c, *us, lx, ly, gx, gy, a, ow & ~OCC_Q1, slot, iteration, env_time, skey, next_phase, ts, lane);
        } else {
            warp_agent_turn<POLT>(w, c, lx, ly, ow & ~OCC_Q1, slot, iteration, env_time, skey, next_phase, ts, lane, cl);
        }
        __syncwarp();
    }
    close_log_slices(w, ts);
    __syncwarp();
    if (!LANES) SOLO_TICK(t_act);
    write_map();
#ifdef SS_PASS_TIMING
    if (w.dbg && lane == 0) {
        atomicAdd(&w.dbg[17], 1ull); atomicAdd(&w.dbg[18], (unsigned long long)t_scan); atomicAdd(&w.dbg[19], (unsigned long long)t_sort);
        atomicAdd(&w.dbg[20], (unsigned long long)t_fill); atomicAdd(&w.dbg[21], (unsigned long long)t_blk);
        atomicAdd(&w.dbg[22], (unsigned long long)t_dep); atomicAdd(&w.dbg[23], (unsigned long long)t_q1);
        atomicAdd(&w.dbg[24], (unsigned long long)t_act); atomicAdd(&w.dbg[25], (unsigned long long)n_rounds);
        atomicAdd(&w.dbg[26], (unsigned long long)n_lane_slots); atomicAdd(&w.dbg[27], (unsigned long long)n_blocked);
        atomicAdd(&w.dbg[28], (unsigned long long)n_mine);
    }
#endif
    if (LANES) {            // every lane commits turns: add the counters up
        ts.sum_m = __reduce_add_sync(FULL, ts.sum_m); ts.sum_v = __reduce_add_sync(FULL, ts.sum_v);
        ts.acted = __reduce_add_sync(FULL, ts.acted); ts.deaths = __reduce_add_sync(FULL, ts.deaths);
        ts.small = __reduce_add_sync(FULL, ts.small); ts.inexact = __reduce_add_sync(FULL, ts.inexact);
        ts.resolved = __reduce_add_sync(FULL, ts.resolved);
    }
    // (otherwise only lane 0 commits turns, so its counters are the warp's)
    if (lane == 0) {
        if (ts.sum_m) atomicAdd(&w.acc->sum_metab, (unsigned long long)ts.sum_m);
        if (ts.sum_v) atomicAdd(&w.acc->sum_vision, (unsigned long long)ts.sum_v);
        if (ts.acted) atomicAdd(&w.acc->acted, (unsigned long long)ts.acted);
        if (ts.deaths) atomicAdd(&w.acc->deaths, (unsigned long long)ts.deaths);
        if (ts.small) atomicAdd(&w.acc->hist_small, (unsigned long long)ts.small);
        if (ts.inexact) atomicAdd(&w.acc->inexact, (unsigned long long)ts.inexact);
        if (ts.resolved) atomicAdd(&w.acc->pending, (unsigned long long)(-(long long)ts.resolved));
    }
}

#ifndef LANES_MIN_CTAS
#define LANES_MIN_CTAS 20
#endif
#ifndef UTIL_MIN_CTAS
#define UTIL_MIN_CTAS 16
#endif
template <int POLT, bool LANES, bool UTIL>
__global__ void __launch_bounds__(32, UTIL ? UTIL_MIN_CTAS : LANES ? LANES_MIN_CTAS : SOLO_MIN_CTAS)
ss_move_solo_kernel(DevWorld w, PassGeom g, int64_t iteration, int64_t env_time) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int cnt_s;
    const int n = w.n;
    if (*((volatile unsigned long long*)&w.acc->pending) == 0ull) return;
    const int bx = g.bx_lo + (int)blockIdx.x / g.nwin, by = (int)blockIdx.x % g.nwin;
    const int nu = n / g.align;
    const int x_lo = g.align * (int)(((long long)bx * nu) / g.nwin), x_hi = g.align * (int)(((long long)(bx + 1) * nu) / g.nwin);
    const int y_lo = g.align * (int)(((long long)by * nu) / g.nwin), y_hi = g.align * (int)(((long long)(by + 1) * nu) / g.nwin);
    const int ex = x_hi - x_lo, ey = y_hi - y_lo;
    const int gx0 = ss_wrap1(x_lo + g.shift_x, n), gy0 = ss_wrap1(y_lo + g.shift_y, n);
    if (gx0 + ex > n || gy0 + ey > n) solo_window<POLT, true, LANES, UTIL>(w, g, iteration, env_time, smem, cnt_s, gx0, gy0, ex, ey);
    else solo_window<POLT, false, LANES, UTIL>(w, g, iteration, env_time, smem, cnt_s, gx0, gy0, ex, ey);
}

// -------------------------------------------------------------------- multi-GPU replicas ----
// Every rank holds the whole lattice and agent store and runs the move pass on its share of the
// window rows; the turns it resolved travel as ShardLogEntry records (NCCL all-gather, sharded.py)
// and are replayed here on the other ranks' replicas, so the replicas stay bit-identical.
// A cell can be touched by several entries of one launch (A leaves it, B enters; or an at-risk
// agent dies on it and another moves in), always in the order leave/die ... enter, and at most one
// entry leaves a living agent on it.  Replay therefore runs in two phases: 0 = clears (cells left,
// cells agents died on), 1 = sets (living agents' words) and everything that commutes.
// Gathered form (all ranks' slices in one buffer, NCCL all-gather layout): blockIdx.y = source rank, its slice starts
// `stride` bytes after the previous one, its entry count is counts[2 * rank + which]; the own rank's slice is skipped.
struct GatheredLogs {
    const unsigned char* base;      // null: plain form (entries, n_entries)
    unsigned long long stride, offset;
    const long long* counts;
    int which, skip_rank;
};
template <class Entry>
__device__ static inline bool gathered_slice(const GatheredLogs& gl, const Entry*& entries, unsigned long long& n_entries) {
    if (!gl.base) return true;
    const int j = (int)blockIdx.y;
    if (j == gl.skip_rank) return false;
    entries = (const Entry*)(gl.base + (unsigned long long)j * gl.stride + gl.offset);
    n_entries = (unsigned long long)gl.counts[2 * j + gl.which];
    return true;
}
__global__ void __launch_bounds__(256)
ss_apply_log_kernel(DevWorld w, const ShardLogEntry* entries, unsigned long long n_entries, int phase, GatheredLogs gl) {
    if (!gathered_slice(gl, entries, n_entries)) return;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < n_entries;
    ShardLogEntry en;
    uint32_t kind = 255u;
    if (in_range) { en = entries[i]; kind = en.kind_harvest & 255u; }
    if (phase == 1) {          // the replicated pending counter drops by the number of valid entries
        const unsigned m = __ballot_sync(FULL, kind != 255u);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.acc->pending, (unsigned long long)(-(long long)__popc(m)));
    }
    if (kind == 255u) return;
    if (phase == 0) {
        if (kind == 2u) return;
        if (en.old_cell != en.new_cell) w.occw[en.old_cell] = 0u;
        if (en.new_word == 0u) w.occw[en.new_cell] = 0u;
        return;
    }
    AgentRec* rec = &w.agents[en.slot];
    if (kind == 2u) {                              // skipped (Q1)
        w.occw[en.new_cell] = en.new_word;
        if (en.status) rec->status = en.status;
        return;
    }
    if (en.new_word != 0u) w.occw[en.new_cell] = en.new_word;
    if (w.cfg.rule_move_eat || w.cfg.rule_utilicalc) { w.sugar[en.new_cell] = 0.0; w.isugar[en.new_cell] = 0; }
    rec->pos = en.new_cell;
    if (kind == 0u) { rec->turn = en.turn; return; }
    // at-risk agent, settled synchronously by its rank: take over the results.  Pollution only grows
    // during the agent loop (A, B >= 0), so the last writer of a cell holds the largest value and the
    // bit pattern of a non-negative double orders like an integer.
    if (w.cfg.rule_pollution)
        atomicMax((unsigned long long*)&w.poll[en.new_cell], (unsigned long long)__double_as_longlong(en.poll));
    rec->sugar = en.sugar;
    rec->attr = en.attr;
    rec->status = en.status;
    atomicAdd(&w.acc->sum_metab, (unsigned long long)attr_metab(en.attr));
    atomicAdd(&w.acc->sum_vision, (unsigned long long)attr_vision(en.attr));
    atomicAdd(&w.acc->acted, 1ull);
    if (!attr_alive(en.attr)) atomicAdd(&w.acc->deaths, 1ull);
    const double sugar = en.sugar;
    if (sugar == 0.01) atomicAdd(&w.acc->hist_small, 1ull);
    else if (sugar >= 0.0 && sugar < (double)SS_GINI_BINS && sugar == (double)(int)sugar) atomicAdd(&w.hist[(int)sugar], 1u);
    else { atomicAdd(&w.acc->inexact, 1ull); ss_note_outlier(w, sugar); }
}

// The compact entries (safe agents: move now, settle later).  Phase 0 clears the cell the agent left -- its position in
// the replica's own record --, phase 1 seats it on the new cell and stamps the deferred turn, exactly as commit_turn does.
__global__ void __launch_bounds__(256)
ss_apply_clog_kernel(DevWorld w, const uint2* entries, unsigned long long n_entries, int phase, int64_t iteration,
                     GatheredLogs gl) {
    if (!gathered_slice(gl, entries, n_entries)) return;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint2 en = make_uint2(CLOG_UNUSED, 0u);
    if (i < n_entries) en = entries[i];
    const bool valid = en.x != CLOG_UNUSED;
    if (phase == 1) {
        const unsigned m = __ballot_sync(FULL, valid);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.acc->pending, (unsigned long long)(-(long long)__popc(m)));
    }
    if (!valid) return;
    const uint32_t slot = en.x & OCC_SLOT_MASK, sg = en.x >> 24, cell = en.y;
    AgentRec* rec = &w.agents[slot];
    if (phase == 0) {
        const uint32_t old = rec->pos;
        if (old != cell) w.occw[old] = 0u;
        return;
    }
    w.occw[cell] = occ_pack(slot, attr_vision(rec->attr), (uint32_t)((iteration + 1) & 1));
    if (w.cfg.rule_move_eat) { w.sugar[cell] = 0.0; w.isugar[cell] = 0; }
    rec->pos = cell;
    rec->turn = ((uint32_t)(iteration + 1) << 8) | sg;
}

// ------------------------------------------------------------------------------ settle ----
// Deferred part of every SAFE agent's turn, one coalesced streaming pass over the agent store:
// harvest (agent.py:832), production + consumption pollution at the destination
// (agent.py:825-826, sugarscape.py:566-567; environment.py:312-315), metabolism
// (sugarscape.py:569), statistics (sugarscape.py:575-583).
#define SETTLE_BINS 2048            // low wealth bins kept in shared memory per CTA
#define SETTLE_PER_THREAD 8
__global__ void __launch_bounds__(256)
ss_settle_kernel(DevWorld w, int64_t iteration, int64_t env_time) {
    __shared__ unsigned hist_s[SETTLE_BINS];
    __shared__ unsigned long long acc_s[6];
    for (int i = threadIdx.x; i < SETTLE_BINS; i += blockDim.x) hist_s[i] = 0u;
    if (threadIdx.x < 6) acc_s[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned sum_m = 0, sum_v = 0, acted = 0, small = 0, inexact = 0, fault = 0;
    const uint32_t stamp = (uint32_t)(iteration + 1) & 0xffffffu;
    const bool pol = w.cfg.rule_pollution && w.cfg.pollution_start_time <= env_time;
    // a CTA streams SETTLE_PER_THREAD * 256 consecutive records (coalesced 32-byte loads)
    const uint32_t base = blockIdx.x * (256u * SETTLE_PER_THREAD);
#pragma unroll 2
    for (int u = 0; u < SETTLE_PER_THREAD; u++) {
        const uint32_t slot = base + u * 256u + threadIdx.x;
        if (slot >= w.n_slots) break;
        AgentRec* rec = &w.agents[slot];
        const uint32_t attr = rec->attr, turn = rec->turn;
        if (!attr_alive(attr) || (turn >> 8) != stamp) continue;
        const int h = (int)(turn & 255u), metab = attr_metab(attr);
        double sugar = rec->sugar;
        if (w.cfg.rule_move_eat) sugar = ss_set_sugar(ss_max0(sugar + (double)h));
        if (pol) {
            double p = w.poll[rec->pos];
            if (w.cfg.rule_move_eat) p += ((double)h * w.cfg.pollution_a) + (0.0 * w.cfg.pollution_b);
            if (sugar > 0.0) p += (0.0 * w.cfg.pollution_a) + ((double)metab * w.cfg.pollution_b);
            w.poll[rec->pos] = p;
        }
        sugar = ss_set_sugar(ss_max0(sugar - (double)metab));
        rec->sugar = sugar;
        sum_m += (unsigned)metab; sum_v += (unsigned)attr_vision(attr); acted++;
        if (sugar == 0.01) small++;
        else if (sugar >= 0.0 && sugar < (double)SS_GINI_BINS && sugar == (double)(int)sugar) {
            const int bin = (int)sugar;
            if (bin < SETTLE_BINS) atomicAdd(&hist_s[bin], 1u); else atomicAdd(&w.hist[bin], 1u);
        } else { inexact++; ss_note_outlier(w, sugar); }
        if (w.cfg.rule_can_starve && sugar <= 0.1) fault++;           // a "safe" agent starved: must not happen
    }
    unsigned vals[6] = {sum_m, sum_v, acted, small, inexact, fault};
#pragma unroll
    for (int q = 0; q < 6; q++) {
        unsigned v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&acc_s[q], (unsigned long long)v);
    }
    __syncthreads();
    unsigned long long* dst[6] = {&w.acc->sum_metab, &w.acc->sum_vision, &w.acc->acted, &w.acc->hist_small,
                                  &w.acc->inexact, &w.acc->fault};
    if (threadIdx.x < 6 && acc_s[threadIdx.x]) atomicAdd(dst[threadIdx.x], acc_s[threadIdx.x]);
    for (int i = threadIdx.x; i < SETTLE_BINS; i += blockDim.x) {
        const unsigned c = hist_s[i];
        if (c) atomicAdd(&w.hist[i], c);
    }
}

// ------------------------------------------------------------------------------- stats ----
#define STATS_THREADS 1024
__global__ void __launch_bounds__(STATS_THREADS, 1)
ss_stats_kernel(DevWorld w, int stats_index) {
    __shared__ double s_cnt[STATS_THREADS], s_sum[STATS_THREADS];
    __shared__ double s_area[STATS_THREADS];
    const int tid = threadIdx.x;
    const int per = SS_GINI_BINS / STATS_THREADS;
    StepAccum* a = w.acc;
    const double c0 = (double)a->hist_small;
    // per-thread totals over its `per` consecutive bins
    double cnt = 0.0, sum = 0.0;
    for (int i = 0; i < per; i++) {
        const int v = tid * per + i;
        const double c = (double)w.hist[v];
        cnt += c; sum += c * (double)v;
    }
    s_cnt[tid] = cnt; s_sum[tid] = sum;
    __syncthreads();
    // exclusive prefix of the wealth mass below my bins (Hillis-Steele on 1024 entries)
    for (int o = 1; o < STATS_THREADS; o <<= 1) {
        double add = tid >= o ? s_sum[tid - o] : 0.0;
        __syncthreads();
        s_sum[tid] += add;
        __syncthreads();
    }
    double H = 0.01 * c0 + (s_sum[tid] - sum);      // height before my first bin
    double area = 0.0;
    for (int i = 0; i < per; i++) {
        const int v = tid * per + i;
        const double c = (double)w.hist[v];
        // c entries of value v after height H:  sum_{k=1..c} (H + k v - v/2) = c H + v c^2 / 2
        area += c * H + (double)v * c * c / 2.0;
        H += c * (double)v;
    }
    s_area[tid] = area;
    __syncthreads();
    for (int o = STATS_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) { s_area[tid] += s_area[tid + o]; s_cnt[tid] += s_cnt[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const double nw = s_cnt[0] + c0;
        const double height = 0.01 * c0 + s_sum[STATS_THREADS - 1];
        const double total_area = s_area[0] + 0.01 * c0 * c0 / 2.0;
        const double fair = height * nw / 2.0;
        double gini = fair == 0.0 ? 1.0 : (fair - total_area) / fair;
        if (a->inexact && !a->fault && w.outliers && a->outliers == a->inexact && a->outliers <= SS_GINI_OUTLIERS) {
            // wealth values outside the histogram (rare): the same sums, one thread, bins and sorted outliers merged in
            // value order -- c entries of value v after height H add c H + v c^2 / 2 (sugarscape.py:269-280)
            const int no = (int)a->outliers;
            double* o = w.outliers;
            for (int i = 1; i < no; i++) {                            // insertion sort (at most SS_GINI_OUTLIERS values)
                const double v = o[i];
                int j = i - 1;
                while (j >= 0 && o[j] > v) { o[j + 1] = o[j]; j--; }
                o[j + 1] = v;
            }
            double H = 0.0, area = 0.0, cnt_all = 0.0;
            int k = 0;
            auto add = [&](double c, double v) { area += c * H + v * c * c / 2.0; H += c * v; cnt_all += c; };
            while (k < no && o[k] < 0.01) add(1.0, o[k++]);
            add(c0, 0.01);
            for (int v = 0; v < SS_GINI_BINS; v++) {
                while (k < no && o[k] < (double)v) add(1.0, o[k++]);
                const double c = (double)w.hist[v];
                if (c > 0.0) add(c, (double)v);
            }
            while (k < no) add(1.0, o[k++]);
            const double fair2 = H * cnt_all / 2.0;
            gini = fair2 == 0.0 ? 1.0 : (fair2 - area) / fair2;
        } else if (a->inexact || a->fault) gini = __longlong_as_double(0x7ff8000000000000ll);
        a->population -= a->deaths;
        const double pop = (double)a->population;
        ss_step_stats st;
        st.population = pop;
        st.metabolism_mean = pop > 0 ? (double)a->sum_metab / pop : 0.0;
        st.vision_mean = pop > 0 ? (double)a->sum_vision / pop : 0.0;
        st.gini = pop > 0 ? gini : 0.0;
        st.acted = (double)a->acted;
        st.deaths = (double)a->deaths;
        w.stats[stats_index] = st;
    }
}

// ---------------------------------------------------------------------------- growback ----
// environment.py:133-144: sugar = min(sugar + alpha, cap), pointwise -> 24 B/cell, HBM-bound.
// Each thread handles 4 consecutive cells (two 128-bit loads per array) and also refreshes
// the int(sugar) byte mirror (+1 B/cell) the move passes stage into shared memory.
__device__ static inline double grow1(double s, double alpha, double c) { const double v = s + alpha; return v <= c ? v : c; }

__global__ void __launch_bounds__(256)
ss_grow_kernel(double* __restrict__ sugar, const double* __restrict__ cap, uint8_t* __restrict__ isugar, double alpha,
               size_t cells, int keep_bit7) {
    const size_t nquad = cells / 4;
    double2* s2 = reinterpret_cast<double2*>(sugar);
    const double2* c2 = reinterpret_cast<const double2*>(cap);
    uchar4* i4 = reinterpret_cast<uchar4*>(isugar);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquad; i += stride) {
        double2 a = s2[2 * i], b = s2[2 * i + 1];
        const double2 ca = __ldcs(&c2[2 * i]), cb = __ldcs(&c2[2 * i + 1]);
        a.x = grow1(a.x, alpha, ca.x); a.y = grow1(a.y, alpha, ca.y);
        b.x = grow1(b.x, alpha, cb.x); b.y = grow1(b.y, alpha, cb.y);
        s2[2 * i] = a; s2[2 * i + 1] = b;
        if (isugar) {
            uchar4 o = make_uchar4(0, 0, 0, 0);
            if (keep_bit7) { o = i4[i]; o.x &= 0x80; o.y &= 0x80; o.z &= 0x80; o.w &= 0x80; }      // occupied bit of the cell words
            i4[i] = make_uchar4(o.x | (unsigned char)(int)a.x, o.y | (unsigned char)(int)a.y, o.z | (unsigned char)(int)b.x,
                                o.w | (unsigned char)(int)b.y);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (cells & 3)) {
        const size_t c = nquad * 4 + threadIdx.x;
        const double v = grow1(sugar[c], alpha, cap[c]);
        sugar[c] = v;
        if (isugar) isugar[c] = (unsigned char)((keep_bit7 ? (isugar[c] & 0x80) : 0) | (unsigned char)(int)v);
    }
}

// sugarscape.py:642-655 + environment.py:146-155: two inclusive rectangles, north first
__global__ void __launch_bounds__(256)
ss_seasons_kernel(DevWorld w, double alpha_north, double alpha_south, uint8_t* mirror, int keep_bit7) {
    const int n = w.n;
    const size_t cells = (size_t)n * n;
    const int32_t* rn = w.cfg.season_north;
    const int32_t* rs = w.cfg.season_south;
    const int n_imin = max(rn[0], 0), n_jmin = max(rn[1], 0), n_imax = min(rn[2] + 1, n), n_jmax = min(rn[3] + 1, n);
    const int s_imin = max(rs[0], 0), s_jmin = max(rs[1], 0), s_imax = min(rs[2] + 1, n), s_jmax = min(rs[3] + 1, n);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(c / n), y = (int)(c % n);
        const bool in_n = x >= n_imin && x < n_imax && y >= n_jmin && y < n_jmax;
        const bool in_s = x >= s_imin && x < s_imax && y >= s_jmin && y < s_jmax;
        if (!in_n && !in_s) continue;
        double s = w.sugar[c];
        const double cp = w.cap[c];
        if (in_n) s = grow1(s, alpha_north, cp);
        if (in_s) s = grow1(s, alpha_south, cp);
        w.sugar[c] = s;
        if (mirror) mirror[c] = (unsigned char)((keep_bit7 ? (mirror[c] & 0x80) : 0) | (unsigned char)(int)s);
    }
}

// --------------------------------------------------------------------------- diffusion ----
// environment.py:317-350: `for i: for j:` in-place 4-neighbour average.  Cell (x,y) sees the
// NEW values of (x,y-1) and (x-1,y) and the OLD values of (x,y+1) and (x+1,y); the sum is
// accumulated in the order n, s, e, w = (x,y+1), (x,y-1), (x+1,y), (x-1,y).
// Skewed wavefront: one warp owns a band of 32 consecutive rows (lane = row); in iteration k
// lane l updates column k - l, so its west neighbour (lane l-1, same column) was produced in
// the previous iteration (warp shuffle) and its south neighbour is its own previous result.
// Old values are staged band-chunk by band-chunk into a shared ring with coalesced loads
// and written back coalesced; bands are chained through global progress flags (band b+1
// needs the new last row of band b).  Bands take tickets in launch order, so a waiting
// band only ever waits on a band that is already running.
#define DIFF_RING 128                    // columns held per row (power of two): chunks j-1, j, j+1 (+ spare)
// x / 3.0, correctly rounded, without the long DDIV sequence on the wavefront's critical
// path: q = RN(x*r), one FMA residual correction (Markstein); exact IEEE division outside the
// comfortably-normal range.  Checked against x / 3.0 on the CPU (tests/ctests/div3_check.c).
// (the IEEE fallback is a call the compiler cannot hoist or if-convert: speculated, its zero/denormal slow path
// would run for every cell of an unpolluted region)
__device__ static __noinline__ double div3_ieee(double x) { return x / 3.0; }
__device__ static inline double div3(double x) {
    const double ax = fabs(x);
    if (!(ax > 1e-280 && ax < 1e280)) return div3_ieee(x);
    const double r = 1.0 / 3.0;
    const double q = x * r;
    const double rem = __fma_rn(-3.0, q, x);
    return __fma_rn(rem, r, q);
}
__device__ static inline int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ static inline void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ static inline void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ static inline void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ static inline void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__global__ void __launch_bounds__(32)
ss_diffuse_kernel(double* __restrict__ p, int n, int* ticket, int* progress) {
    __shared__ double tile[34][DIFF_RING];     // row 0 = above (new), 1..32 = band, 33 = below (old)
    const int lane = threadIdx.x;
    int b = 0;
    if (lane == 0) b = atomicAdd(ticket, 1);
    b = __shfl_sync(FULL, b, 0);
    const int x0 = b * 32;
    const int rows = min(32, n - x0);
    const int r = x0 + lane;
    const bool row_ok = lane < rows;
    const bool has_above = x0 > 0, has_below = x0 + 32 < n;
    const bool has_e = r + 1 < n, has_w = r >= 1;
    const int nchunks = (n + 31) / 32;
    const int nblocks = (n + 31 + 31) / 32;          // iterations 0 .. n+30
    // asynchronous chunk load (cp.async, no register staging): columns [32ch, 32ch+32) of my rows + the row below
    auto load_chunk = [&](int ch) {
        const int col = 32 * ch + lane;
        if (ch < nchunks && col < n) {
            const int slot = col & (DIFF_RING - 1);
            for (int rr = 1; rr <= rows; rr++) cp_async8(&tile[rr][slot], &p[(size_t)(x0 - 1 + rr) * n + col]);
            if (has_below) cp_async8(&tile[33][slot], &p[(size_t)(x0 + 32) * n + col]);
        }
        cp_async_commit();
    };
    auto load_above = [&](int ch) {
        const int col = 32 * ch + lane;
        if (!has_above || ch >= nchunks) return;
        const int need = min(n, 32 * (ch + 1));
        while (ld_acquire(&progress[b - 1]) < need) { }       // all lanes (same address): the warp stays converged
        __syncwarp();
        if (col < n) tile[0][col & (DIFF_RING - 1)] = __ldcg(&p[(size_t)(x0 - 1) * n + col]);
    };
    auto flush_chunk = [&](int ch) {
        if (ch < 0 || ch >= nchunks) return;
        const int col = 32 * ch + lane;
        if (col < n) {
            const int slot = col & (DIFF_RING - 1);
            for (int rr = 1; rr <= rows; rr++) p[(size_t)(x0 - 1 + rr) * n + col] = tile[rr][slot];
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(&progress[b], min(n, 32 * (ch + 1)));
    };
    load_chunk(0);
    load_chunk(1);
    double res_prev = 0.0;
    for (int j = 0; j < nblocks; j++) {
        load_chunk(j + 2);                      // chunk j+2 lands while blocks j, j+1 compute
        cp_async_wait<1>();                     // chunks <= j+1 have landed
        load_above(j);
        __syncwarp();
        const int k0 = 32 * j;
        const bool interior = row_ok && rows == 32 && has_above && has_below && (k0 - 31 >= 1) && (k0 + 32 < n);
        if (__all_sync(FULL, interior)) {
            // every lane has all four neighbours: ((n + s) + e) + w, times 0.25
#pragma unroll 4
            for (int k = k0; k < k0 + 32; k++) {
                const int y = k - lane;
                const double n_val = tile[lane + 1][(y + 1) & (DIFF_RING - 1)];
                double e_val = __shfl_down_sync(FULL, n_val, 1);
                double w_val = __shfl_up_sync(FULL, res_prev, 1);
                if (lane == 31) e_val = tile[33][y & (DIFF_RING - 1)];
                if (lane == 0) w_val = tile[0][y & (DIFF_RING - 1)];
                const double res = (((n_val + res_prev) + e_val) + w_val) * 0.25;
                tile[lane + 1][y & (DIFF_RING - 1)] = res;
                res_prev = res;
            }
        } else {
            for (int k = k0; k < k0 + 32; k++) {
                const int y = k - lane;
                const bool act = row_ok && y >= 0 && y < n;
                const bool has_n = y + 1 < n, has_s = y >= 1;
                const double n_val = (row_ok && y + 1 >= 0 && has_n) ? tile[lane + 1][(y + 1) & (DIFF_RING - 1)] : 0.0;
                double e_val = __shfl_down_sync(FULL, n_val, 1);
                double w_val = __shfl_up_sync(FULL, res_prev, 1);
                if (lane == 31 && act && has_below) e_val = tile[33][y & (DIFF_RING - 1)];
                if (lane == 0 && act && has_above) w_val = tile[0][y & (DIFF_RING - 1)];
                if (act) {
                    double t = 0.0;
                    int cnt = 0;
                    if (has_n) { t += n_val; cnt++; }
                    if (has_s) { t += res_prev; cnt++; }
                    if (has_e) { t += e_val; cnt++; }
                    if (has_w) { t += w_val; cnt++; }
                    const double res = cnt == 4 ? t * 0.25 : (cnt == 2 ? t * 0.5 : div3(t));
                    tile[lane + 1][y & (DIFF_RING - 1)] = res;
                    res_prev = res;
                }
            }
        }
        __syncwarp();
        flush_chunk(j - 1);
    }
    cp_async_wait<0>();
    flush_chunk(nblocks - 1);
    __syncwarp();
    if (lane == 0) st_release(&progress[b], n);
}

// Column-chain wavefront.  The reference adds the four neighbours in the order n, s, e, w = (x,y+1), (x,y-1),
// (x+1,y), (x-1,y): the NEW value of the row above enters LAST.  A thread that walks DOWN a column segment therefore
// carries only one add and one multiply per cell on its dependent chain (res[r] = (t[r] + res[r-1]) * 0.25 with
// t[r] = (n + s) + e known a step earlier), against three adds, a multiply and a shuffle per cell when it walks along
// a row (ss_diffuse_kernel).  A CTA owns a band of 32 * DCOL_R rows.  Warp 0 computes: lane l owns DCOL_R consecutive
// rows, works on column k - S*l in step k and receives the new value above its first row from lane l-1 by shuffle
// (S = 2: the value has a whole step to travel).  DCOL_H helper warps (one per scheduler, so they never take issue
// slots from the compute warp) keep a shared ring of old values filled with cp.async, write finished chunks back, load
// the row above the band once the previous band (previous ticket) has published it through its global progress flag,
// and publish this band's progress.  Producer/consumer counters live in shared memory.
#ifndef DCOL_R
#define DCOL_R 4
#endif
#define DCOL_ROWS (32 * DCOL_R)
#define DCOL_H 3
__device__ __forceinline__ double dcol_lds(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void dcol_sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// x / 3.0 for the fast path: Markstein step in the normal range (also exact for zeros), IEEE division otherwise
__device__ __forceinline__ double dcol_div3(double x) {
    const double r3 = 1.0 / 3.0;
    const double q = x * r3;
    double d = __fma_rn(__fma_rn(-3.0, q, x), r3, q);
    const double ax = fabs(x);
    if (!(ax > 1e-280 && ax < 1e280) && x != 0.0) d = div3_ieee(x);
    return d;
}
struct DcolState {
    double s_prev[DCOL_R], old_cur[DCOL_R], hand[2];
    uint32_t a_row;         // shared address of (my first row, slot 0)
    uint32_t a_above;       // shared address of (row above the band, slot 0)
    uint32_t off_s;         // byte offset of my current column's ring slot
    double* out;            // lane 31: where the band below picks up my last row (column of this step)
};
// 32 steps of the column-chain wavefront for a full band.  TOP / BOT: the band holds the lattice's first / last
// row (lane 0's first row has no w, lane 31's last row no e: three neighbours, x/3).  START: the first and the last
// blocks of the sweep, where lanes are waiting to enter or have left (column < 0 or >= n: idle) or stand on the first
// / last column (no s / no n neighbour: the missing operand is 0.0 -- the initial s_prev, a masked load -- which
// adds exactly; one neighbour less in the divisor); otherwise every lane is on columns 1 .. n-2.
template <int S, int RING, bool TOP, bool BOT, bool START>
__device__ __forceinline__ void dcol_fast_block(DcolState& st, int lane, int k0, int n) {
    constexpr uint32_t ROWB = RING * 8u;
    int y = k0 - S * lane;
#pragma unroll 2
    for (int k = 0; k < 32; k++, y++) {
        const uint32_t off_n = st.off_s + 8u == ROWB ? 0u : st.off_s + 8u;
        const bool act = !START || (y >= 0 && y < n), col0 = START && (y == 0 || y == n - 1);
        double w_in = __shfl_up_sync(FULL, st.hand[0], 1);
        double nv[DCOL_R];
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) nv[r] = dcol_lds(st.a_row + r * ROWB + off_n);
        double e_last = 0.0;
        if (!BOT || lane != 31) e_last = dcol_lds(st.a_row + DCOL_R * ROWB + st.off_s);
        if (!TOP) { const double ab = dcol_lds(st.a_above + st.off_s); if (lane == 0) w_in = ab; }
        double t[DCOL_R];
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) {
            t[r] = ((START && y == n - 1) ? 0.0 : nv[r]) + st.s_prev[r];
            if (r + 1 < DCOL_R) t[r] = t[r] + st.old_cur[r + 1];
        }
        const double t_last_e = t[DCOL_R - 1] + e_last;     // the usual last row: n, s, e (, w)
        double up = w_in;
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) {
            // sum of the present neighbours in the reference's order; `edge` = this cell is on the first / last row
            const bool edge = (TOP && r == 0 && lane == 0) || (BOT && r == DCOL_R - 1 && lane == 31);
            double sum;
            if (r + 1 < DCOL_R) sum = (TOP && r == 0 && lane == 0) ? t[0] : t[r] + up;
            else sum = (BOT && lane == 31) ? t[r] + up : t_last_e + up;
            double res = sum * 0.25;
            if (TOP && r == 0 || BOT && r == DCOL_R - 1 || START) {
                const double d3 = dcol_div3(sum);
                if (START) res = col0 ? (edge ? sum * 0.5 : d3) : (edge ? d3 : res);
                else if (edge) res = d3;
            }
            if (act) dcol_sts(st.a_row + r * ROWB + st.off_s, res);
            if (START) { st.s_prev[r] = act ? res : 0.0; up = act ? res : up; }
            else { st.s_prev[r] = res; up = res; }
            st.old_cur[r] = nv[r];
        }
#pragma unroll
        for (int i = 0; i + 1 < S; i++) st.hand[i] = st.hand[i + 1];
        st.hand[S - 1] = up;
        if (!BOT && lane == 31 && act) *st.out = up;
        st.out++;
        st.off_s = off_n;
    }
}
// The same 32 steps for blocks that touch the lattice border or start / end the skewed sweep: per cell the present
// neighbours are added in the reference's order (n, s, e, w) behind select masks, the divisor follows their count,
// lanes outside the lattice idle -- branch-free, so the first blocks of a band (which every later band waits for)
// cost little more than the interior ones.
template <int S, int RING>
__device__ __forceinline__ void dcol_edge_block(DcolState& st, int lane, int k0, int n, int x_first, bool has_above,
                                                bool has_below) {
    constexpr uint32_t ROWB = RING * 8u;
#pragma unroll 1
    for (int k = 0; k < 32; k++) {
        const int y = k0 + k - S * lane;
        const bool col_ok = y >= 0 && y < n, has_n = y + 1 < n, has_s = y >= 1;
        const uint32_t off_n = st.off_s + 8u == ROWB ? 0u : st.off_s + 8u;
        double w_in = __shfl_up_sync(FULL, st.hand[0], 1);
        double nv[DCOL_R];
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) nv[r] = dcol_lds(st.a_row + r * ROWB + off_n);
        const double e_last = dcol_lds(st.a_row + DCOL_R * ROWB + st.off_s);
        const double ab = dcol_lds(st.a_above + st.off_s);
        if (lane == 0) w_in = ab;
        double up = w_in;
#pragma unroll
        for (int r = 0; r < DCOL_R; r++) {
            const int x = x_first + r;
            const bool has_e = x + 1 < n, has_w = x >= 1 && (r > 0 || lane > 0 || has_above);
            const double e = r + 1 < DCOL_R ? st.old_cur[r + 1] : e_last;
            double t = 0.0 + (has_n ? nv[r] : 0.0);
            t = t + (has_s ? st.s_prev[r] : 0.0);
            t = t + (has_e ? e : 0.0);
            t = t + (has_w ? up : 0.0);
            const int cnt = (int)has_n + (int)has_s + (int)has_e + (int)has_w;
            const double d3 = dcol_div3(t);
            const double res = cnt == 4 ? t * 0.25 : (cnt == 3 ? d3 : (cnt == 2 ? t * 0.5 : t));
            const bool active = col_ok && x < n;
            if (active) dcol_sts(st.a_row + r * ROWB + st.off_s, res);
            st.s_prev[r] = active ? res : st.s_prev[r];
            up = active ? res : up;
            st.old_cur[r] = nv[r];
        }
#pragma unroll
        for (int i = 0; i + 1 < S; i++) st.hand[i] = st.hand[i + 1];
        st.hand[S - 1] = up;
        if (lane == 31 && has_below && col_ok) st.out[k] = up;
        st.off_s = off_n;
    }
    st.out += 32;
}
struct DcolShared {
    volatile int loaded[DCOL_H];    // chunks whose rows of helper h have landed in the ring
    volatile int above;             // chunks of the row above the band that are in the ring
    volatile int done;              // blocks of 32 steps the compute warp has finished
};
template <int S>
__global__ void __launch_bounds__(32 * (1 + DCOL_H), 1)
ss_diffuse_col_kernel(double* __restrict__ p, int n, int* ticket, double* handoff) {
    constexpr int NR = S + 3;                           // ring chunks: j-S .. j+2 are live in block j
    constexpr int RING = 32 * NR;
    extern __shared__ __align__(16) double dcol_tile[]; // [DCOL_ROWS + 2][RING]: row 0 = above (new), last = below (old)
    __shared__ DcolShared sh;
    __shared__ int band_s;
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        band_s = atomicAdd(ticket, 1);
        for (int h = 0; h < DCOL_H; h++) sh.loaded[h] = 0;
        sh.above = 0; sh.done = 0;
    }
    __syncthreads();
    const int b = band_s;
    const int x0 = b * DCOL_ROWS;
    const int rows = min(DCOL_ROWS, n - x0);
    const bool has_above = x0 > 0, has_below = x0 + DCOL_ROWS < n;
    const int nchunks = (n + 31) / 32;
    const int nsteps = n + 31 * S;                      // steps 0 .. n-1+31*S
    const int nblocks = (nsteps + 31) / 32;
    // every lane polls (one broadcast read): a single spinning lane would leave the warp split in two
    // independently scheduled halves afterwards, and every shuffle of the compute loop would take the divergent path
    auto wait_ge = [&](const volatile int* f, int need) {
        while (*f < need) { if (wq > 0) __nanosleep(100); }
        __syncwarp();
        __threadfence_block();
    };
    if (wq > 0) {
        // ------------------------------------------------------------------ helper warps ----
        const int h = wq - 1;
        // tile rows of this helper: an even share of the band's rows; the last helper also owns the row below
        const int per = (DCOL_ROWS + DCOL_H - 1) / DCOL_H;
        const int r_lo = 1 + h * per, r_hi = min(rows, (h + 1) * per);         // tile rows r_lo .. r_hi
        const bool last = h == DCOL_H - 1, first = h == 0;
        for (int i = 0; i < nchunks + 2 + S; i++) {
            if (i < nchunks) {                          // (a) load chunk i (its ring slot was written back in (c) of i-1)
                const int col = 32 * i + lane;
                if (col < n) {
                    double* dst = dcol_tile + (size_t)(32 * (i % NR) + lane);
                    for (int rr = r_lo; rr <= r_hi; rr++) cp_async8(dst + (size_t)rr * RING, &p[(size_t)(x0 - 1 + rr) * n + col]);
                    if (last && has_below) cp_async8(dst + (size_t)(DCOL_ROWS + 1) * RING, &p[(size_t)(x0 + DCOL_ROWS) * n + col]);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
                if (lane == 0) { __threadfence_block(); sh.loaded[h] = i + 1; }
            }
            if (first && i >= 1 && i - 1 < nchunks) {   // (b) row above the band, chunk i-1 (needed by block i-1)
                const int ch = i - 1;
                if (has_above) {
                    // the band above (previous ticket) stores its last row's new values into handoff[] as it computes
                    // them; the buffer starts as all-ones NaNs, so a value is its own "ready" flag -- no fences
                    const int col = 32 * ch + lane;
                    const volatile double* src = handoff + (size_t)(b - 1) * n + col;
                    double v = 0.0;
                    bool got = col >= n;
                    for (;;) {
                        if (!got) { v = *src; got = __double_as_longlong(v) != -1ll; }
                        if (__all_sync(FULL, got)) break;
                    }
                    if (col < n) dcol_tile[32 * (ch % NR) + lane] = v;
                    __syncwarp();
                }
                if (lane == 0) { __threadfence_block(); sh.above = ch + 1; }
            }
            const int f = i - 2 - S;                    // (c) write chunk f back once every lane has passed it
            if (f >= 0 && f < nchunks) {
                wait_ge(&sh.done, min(f + S + 1, nblocks));
                const int col = 32 * f + lane;
                if (col < n) {
                    const double* src = dcol_tile + (size_t)(32 * (f % NR) + lane);
                    for (int rr = r_lo; rr <= r_hi; rr++) p[(size_t)(x0 - 1 + rr) * n + col] = src[(size_t)rr * RING];
                }
                __syncwarp();
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warp ----
    const int row0 = 1 + DCOL_R * lane;                 // tile row of my first lattice row
    DcolState st;                                       // new values of column y-1, old values of column y (my rows), ...
#pragma unroll
    for (int r = 0; r < DCOL_R; r++) { st.s_prev[r] = 0.0; st.old_cur[r] = 0.0; }
    st.hand[0] = st.hand[1] = 0.0;                      // my last row's results of the last S steps (oldest first)
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(dcol_tile);
    st.a_row = tile_s + (uint32_t)(row0 * RING) * 8u;
    st.a_above = tile_s;
    st.off_s = (uint32_t)((RING - ((S * lane) % RING)) % RING) * 8u;    // ring slot of my column y = k - S*lane at k = 0
    st.out = handoff + (size_t)b * n - S * 31;          // lane 31's column of step 0
    for (int j = 0; j < nblocks; j++) {
        for (int h = 0; h < DCOL_H; h++) wait_ge(&sh.loaded[h], min(j + 2, nchunks));
        wait_ge(&sh.above, min(j + 1, nchunks));
        if (j == 0) {                                   // old values of my first column (lane 0 starts right away)
#pragma unroll
            for (int r = 0; r < DCOL_R; r++) st.old_cur[r] = dcol_lds(st.a_row + r * (RING * 8u) + st.off_s);
        }
        const int k0 = 32 * j;
        // interior blocks: a full band, every lane on columns 1 .. n-2.  The lattice's first and last row (three
        // neighbours: no w / no e) are handled by the two lanes that own them, so the first and the last band --
        // the first paces every other band -- run them too.
        const bool full = rows == DCOL_ROWS && (has_above || has_below) && n >= 64 + 31 * S;
        const bool interior = full && (k0 - 31 * S >= 1) && (k0 + 32 < n);
        if (interior) {
            if (!has_above) dcol_fast_block<S, RING, true, false, false>(st, lane, k0, n);
            else if (!has_below) dcol_fast_block<S, RING, false, true, false>(st, lane, k0, n);
            else dcol_fast_block<S, RING, false, false, false>(st, lane, k0, n);
        } else if (full) {          // the sweep is starting or ending: lanes enter / leave one after the other
            if (!has_above) dcol_fast_block<S, RING, true, false, true>(st, lane, k0, n);
            else if (!has_below) dcol_fast_block<S, RING, false, true, true>(st, lane, k0, n);
            else dcol_fast_block<S, RING, false, false, true>(st, lane, k0, n);
        } else {
            dcol_edge_block<S, RING>(st, lane, k0, n, x0 + DCOL_R * lane, has_above, has_below);
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sh.done = j + 1; }
    }
}

// simple reference implementation on the device (anti-diagonal sweep, one CTA): A/B check only
__global__ void __launch_bounds__(1024, 1)
ss_diffuse_diag_kernel(double* p, int n) {
    for (int d = 0; d <= 2 * n - 2; d++) {
        const int xa = max(0, d - n + 1), xb = min(d, n - 1);
        for (int x = xa + (int)threadIdx.x; x <= xb; x += blockDim.x) {
            const int y = d - x;
            double t = 0.0;
            int c = 0;
            if (y + 1 < n) { t += p[(size_t)x * n + y + 1]; c++; }
            if (y - 1 >= 0) { t += p[(size_t)x * n + y - 1]; c++; }
            if (x + 1 < n) { t += p[(size_t)(x + 1) * n + y]; c++; }
            if (x - 1 >= 0) { t += p[(size_t)(x - 1) * n + y]; c++; }
            p[(size_t)x * n + y] = t / (double)c;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------- upload / download helpers ----
// int(sugar) mirror from the uploaded sugar (and a flag if a value or capacity does not fit a byte)
__global__ void ss_isugar_kernel(const double* __restrict__ sugar, const double* __restrict__ cap, uint8_t* __restrict__ isugar,
                                 size_t cells, int* flag) {
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const double v = sugar[c];
        if (!(v >= 0.0 && v < 256.0) || !(cap[c] < 256.0)) { atomicOr(flag, 1); isugar[c] = 0; }
        else isugar[c] = (uint8_t)(int)v;
        if (!(v < 127.0) || !(cap[c] < 127.0)) atomicOr(flag, 2);      // the one-byte cell words of ss_rounds.cu hold 7 bits
    }
}
// agent store + occupancy words from the uploaded View.agents arrays; err: 1 duplicate id, 2 two agents on a cell
__global__ void ss_validate_rows_kernel(int n, const int32_t* __restrict__ x, int x0, int rows, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (x[i] < x0 || x[i] >= x0 + rows)) *out = 1;
}
__global__ void ss_build_agents_kernel(int n, int N, int x0, uint32_t phase, const int32_t* __restrict__ id, const int32_t* __restrict__ x,
                                       const int32_t* __restrict__ y, const int32_t* __restrict__ vision,
                                       const int32_t* __restrict__ metab, const double* __restrict__ sugar, AgentRec* recs,
                                       uint32_t* occw, int32_t* order, int* err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = (uint32_t)id[i] - 1u;
    const uint32_t c = (uint32_t)(x[i] - x0) * (uint32_t)N + (uint32_t)y[i];
    const uint32_t attr = attr_pack(vision[i], metab[i], true);
    if (atomicCAS(&recs[s].attr, 0u, attr) != 0u) { atomicMax(err, 1); return; }
    if (atomicCAS(&occw[c], 0u, occ_pack(s, vision[i], phase)) != 0u) { atomicMax(err, 2); return; }
    recs[s].sugar = sugar[i];
    recs[s].pos = c;
    order[i] = (int32_t)s;
}
// argument checks of ss_upload_agents on the staged columns: out[0] = error code (0 ok), out[1] = max id, out[2] = max vision
__global__ void ss_validate_agents_kernel(int n, int N, const int32_t* __restrict__ id, const int32_t* __restrict__ x,
                                          const int32_t* __restrict__ y, const int32_t* __restrict__ vision,
                                          const int32_t* __restrict__ metab, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int err = 0, mid = 0, mv = 0;
    if (i < n) {
        if (id[i] < 1) err = 3;
        else if (vision[i] < 1 || vision[i] > SS_MAX_VISION) err = 4;
        else if (2 * vision[i] + 1 > N) err = 5;
        else if (metab[i] < 1 || metab[i] > 65535) err = 6;
        else if (x[i] < 0 || x[i] >= N || y[i] < 0 || y[i] >= N) err = 7;
        mid = id[i]; mv = vision[i];
    }
    err = __reduce_max_sync(0xffffffffu, err);
    mid = __reduce_max_sync(0xffffffffu, mid);
    mv = __reduce_max_sync(0xffffffffu, mv);
    if ((threadIdx.x & 31) == 0) {
        if (err) atomicMax(&out[0], err);
        atomicMax(&out[1], mid);
        atomicMax(&out[2], mv);
    }
}
__global__ void ss_count_alive_kernel(const AgentRec* __restrict__ agents, uint32_t n_slots, unsigned long long* out) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, slot < n_slots && attr_alive(agents[slot].attr));
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}
__global__ void ss_occ_ids_kernel(const uint32_t* __restrict__ occw, int32_t* __restrict__ ids, size_t cells) {
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w = occw[c];
        ids[c] = w ? (int32_t)(occ_slot(w) + 1u) : 0;
    }
}


// ss_download_agents for the keyed order: the reference's list after its shuffle is the living agents in ascending
// execution priority (sugarscape.py:511-512 as restated by oracle/keyed_rng.py).  The priority is a bijection of the
// slot, so the list is the stream compaction of priority_inverse(0..domain-1): count, scan, emit -- no sort.
#define EXPORT_PER_THREAD 8
#define EXPORT_THREADS 256
#define EXPORT_PER_BLOCK (EXPORT_PER_THREAD * EXPORT_THREADS)
__device__ static inline bool export_slot(const AgentRec* __restrict__ agents, const OrderKey& ok, uint32_t n_slots,
                                          uint32_t p, uint32_t domain, uint32_t* slot) {
    if (p >= domain) return false;
    const uint32_t s = ss_priority_inverse(ok, p);
    *slot = s;
    return s < n_slots && attr_alive(agents[s].attr);
}
__global__ void __launch_bounds__(EXPORT_THREADS)
ss_export_count_kernel(const AgentRec* __restrict__ agents, OrderKey ok, uint32_t n_slots, uint32_t domain,
                       uint32_t* __restrict__ blockcnt) {
    __shared__ uint32_t wsum[EXPORT_THREADS / 32];
    const uint32_t p0 = blockIdx.x * EXPORT_PER_BLOCK + threadIdx.x * EXPORT_PER_THREAD;
    uint32_t c = 0, s;
#pragma unroll
    for (int k = 0; k < EXPORT_PER_THREAD; k++) c += export_slot(agents, ok, n_slots, p0 + k, domain, &s) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < EXPORT_THREADS / 32; i++) t += wsum[i];
        blockcnt[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024)
ss_export_scan_kernel(uint32_t* __restrict__ blockcnt, uint32_t nblocks, uint32_t* __restrict__ total) {
    // exclusive scan of the block counts, in place, by one CTA (nblocks <= 2^24 / 2048)
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? blockcnt[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = wsum[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0u) + inc - v;
        if (i < nblocks) blockcnt[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(EXPORT_THREADS)
ss_export_emit_kernel(const AgentRec* __restrict__ agents, OrderKey ok, uint32_t n_slots, uint32_t domain, int n, int x0,
                      const uint32_t* __restrict__ blockoff, int32_t* __restrict__ id, int32_t* __restrict__ x,
                      int32_t* __restrict__ y, int32_t* __restrict__ vision, int32_t* __restrict__ metab,
                      double* __restrict__ sugar) {
    __shared__ uint32_t wsum[EXPORT_THREADS / 32];
    const uint32_t p0 = blockIdx.x * EXPORT_PER_BLOCK + threadIdx.x * EXPORT_PER_THREAD;
    uint32_t slot[EXPORT_PER_THREAD];
    uint32_t live = 0, c = 0;
#pragma unroll
    for (int k = 0; k < EXPORT_PER_THREAD; k++)
        if (export_slot(agents, ok, n_slots, p0 + k, domain, &slot[k])) { live |= 1u << k; c++; }
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t at = blockoff[blockIdx.x] + inc - c;
    for (int i = 0; i < (int)(threadIdx.x >> 5); i++) at += wsum[i];
#pragma unroll
    for (int k = 0; k < EXPORT_PER_THREAD; k++) {
        if (!((live >> k) & 1u)) continue;
        const AgentRec r = agents[slot[k]];
        id[at] = (int32_t)(slot[k] + 1u);
        x[at] = x0 + (int32_t)(r.pos / (uint32_t)n);
        y[at] = (int32_t)(r.pos % (uint32_t)n);
        vision[at] = attr_vision(r.attr);
        metab[at] = attr_metab(r.attr);
        sugar[at] = r.sugar;
        at++;
    }
}

// ---------------------------------------------------------------------------- launchers ----
extern "C" {

cudaError_t ss_launch_isugar(const DevWorld& w, int* flag, cudaStream_t s) {
    ss_isugar_kernel<<<148 * 8, 256, 0, s>>>(w.sugar, w.cap, w.isugar, (size_t)w.rows * w.n, flag);
    return cudaGetLastError();
}
cudaError_t ss_launch_build_agents(const DevWorld& w, int n, uint32_t phase, const int32_t* id, const int32_t* x, const int32_t* y,
                                   const int32_t* vision, const int32_t* metab, const double* sugar, int* err, cudaStream_t s) {
    if (n > 0)
        ss_build_agents_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, w.n, w.x0, phase, id, x, y, vision, metab, sugar, w.agents, w.occw,
                                                               w.order, err);
    return cudaGetLastError();
}
uint32_t ss_export_blocks(uint32_t domain) { return (domain + EXPORT_PER_BLOCK - 1) / EXPORT_PER_BLOCK; }
cudaError_t ss_launch_export_agents(const DevWorld& w, uint64_t step_key, uint32_t* blockcnt, uint32_t* total, int32_t* cols,
                                    size_t stride, double* sugar, cudaStream_t s) {
    // cols = five int32 columns (id, x, y, vision, metab) of `stride` entries each; *total receives the population
    const OrderKey ok = ss_order_key(step_key, w.half);
    const uint32_t domain = 1u << (2 * w.half);
    const uint32_t nb = ss_export_blocks(domain);
    ss_export_count_kernel<<<nb, EXPORT_THREADS, 0, s>>>(w.agents, ok, w.n_slots, domain, blockcnt);
    ss_export_scan_kernel<<<1, 1024, 0, s>>>(blockcnt, nb, total);
    ss_export_emit_kernel<<<nb, EXPORT_THREADS, 0, s>>>(w.agents, ok, w.n_slots, domain, w.n, w.x0, blockcnt, cols, cols + stride,
                                                       cols + 2 * stride, cols + 3 * stride, cols + 4 * stride, sugar);
    return cudaGetLastError();
}
cudaError_t ss_launch_validate_rows(int n, const int32_t* x, int x0, int rows, int* out, cudaStream_t s) {
    if (n > 0) ss_validate_rows_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, x, x0, rows, out);
    return cudaGetLastError();
}
cudaError_t ss_launch_validate_agents(int n, int N, const int32_t* id, const int32_t* x, const int32_t* y, const int32_t* vision,
                                      const int32_t* metab, int* out3, cudaStream_t s) {
    if (n > 0) ss_validate_agents_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, N, id, x, y, vision, metab, out3);
    return cudaGetLastError();
}
cudaError_t ss_launch_count_alive(const DevWorld& w, unsigned long long* out, cudaStream_t s) {
    if (w.n_slots) ss_count_alive_kernel<<<(w.n_slots + 255) / 256, 256, 0, s>>>(w.agents, w.n_slots, out);
    return cudaGetLastError();
}
cudaError_t ss_launch_occ_ids(const DevWorld& w, int32_t* ids, cudaStream_t s) {
    ss_occ_ids_kernel<<<148 * 8, 256, 0, s>>>(w.occw, ids, (size_t)w.rows * w.n);
    return cudaGetLastError();
}

cudaError_t ss_launch_prepare(const DevWorld& w, int64_t iteration, int stamp_q1, cudaStream_t s) {
    ss_begin_step_kernel<<<1, 1, 0, s>>>(w);
    if (stamp_q1 && w.cfg.rule_can_starve && w.n_slots > 0)
        ss_prepare_kernel<<<(w.n_slots + 255) / 256, 256, 0, s>>>(w, iteration);
    return cudaGetLastError();
}

size_t ss_pass_smem_bytes(int emax, int ew, int vmax) {
    (void)vmax;
    return (((size_t)emax * emax * 5 + 15) & ~(size_t)15) + sizeof(RegionShared) * PASS_WARPS + (size_t)emax * ew * 4 + 16;
}

// one phase (0 clears, 1 sets) of the replay of another rank's entries; the caller runs phase 0 of both logs before phase 1
cudaError_t ss_launch_apply_log(const DevWorld& w, const void* entries, unsigned long long n, int phase, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    GatheredLogs gl{};
    ss_apply_log_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, (const ShardLogEntry*)entries, n, phase, gl);
    return cudaGetLastError();
}
// one phase of the replay of every other rank's slices of an all-gathered buffer: one launch per log
cudaError_t ss_launch_apply_gathered(const DevWorld& w, const void* base, unsigned long long stride, unsigned long long off_compact,
                                     const long long* counts, int world, int skip_rank, unsigned long long max_full,
                                     unsigned long long max_compact, int phase, int64_t iteration, cudaStream_t s) {
    GatheredLogs gl;
    gl.base = (const unsigned char*)base; gl.stride = stride; gl.counts = counts; gl.skip_rank = skip_rank;
    if (max_full) {
        gl.offset = 0; gl.which = 0;
        ss_apply_log_kernel<<<dim3((unsigned)((max_full + 255) / 256), (unsigned)world), 256, 0, s>>>(w, nullptr, 0, phase, gl);
    }
    if (max_compact) {
        gl.offset = off_compact; gl.which = 1;
        ss_apply_clog_kernel<<<dim3((unsigned)((max_compact + 255) / 256), (unsigned)world), 256, 0, s>>>(w, nullptr, 0, phase, iteration, gl);
    }
    return cudaGetLastError();
}

cudaError_t ss_launch_apply_clog(const DevWorld& w, const void* entries, unsigned long long n, int phase, int64_t iteration,
                                 cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    GatheredLogs gl{};
    ss_apply_clog_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, (const uint2*)entries, n, phase, iteration, gl);
    return cudaGetLastError();
}

cudaError_t ss_launch_settle(const DevWorld& w, int64_t iteration, int64_t env_time, cudaStream_t s) {
    if (w.n_slots > 0)
        ss_settle_kernel<<<(w.n_slots + 256 * SETTLE_PER_THREAD - 1) / (256 * SETTLE_PER_THREAD), 256, 0, s>>>(w, iteration, env_time);
    return cudaGetLastError();
}

cudaError_t ss_launch_move_pass(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time,
                                cudaStream_t s) {
    static size_t configured = 0;
    const size_t smem = ss_pass_smem_bytes(g.emax, g.ew, w.vmax);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ss_move_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    if (g.bx_cnt <= 0) return cudaSuccess;
    ss_move_pass_kernel<<<g.nwin * g.bx_cnt, PASS_THREADS, smem, s>>>(w, g, iteration, env_time);
    return cudaGetLastError();
}

// key list + positions + blocked bitmap (+ the shadow bitmap of the lane-per-agent loop)
// mode: 0 one agent at a time, 1 lane per agent (+ shadow bitmap), 2 utilicalc rule (+ the moves / neighbours of one agent)
size_t ss_solo_smem_bytes(int emax, int ew, int mode) {
    return (size_t)SOLO_LIST * 6 + (size_t)emax * ew * (mode == 1 ? 8 : 4) + (mode == 2 ? sizeof(UtilShared) + 8 : 0);
}
cudaError_t ss_launch_move_solo(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time, int lanes,
                                cudaStream_t s) {
    if (g.bx_cnt <= 0) return cudaSuccess;
    const size_t smem = ss_solo_smem_bytes(g.emax, g.ew, w.cfg.rule_utilicalc ? 2 : lanes ? 1 : 0);
    const int grid = g.nwin * g.bx_cnt;
    if (w.cfg.rule_utilicalc) {
        ss_move_solo_kernel<0, false, true><<<grid, 32, smem, s>>>(w, g, iteration, env_time);
    } else if (lanes) {
        if (w.cfg.rule_pollution) ss_move_solo_kernel<1, true, false><<<grid, 32, smem, s>>>(w, g, iteration, env_time);
        else ss_move_solo_kernel<0, true, false><<<grid, 32, smem, s>>>(w, g, iteration, env_time);
    } else {
        if (w.cfg.rule_pollution) ss_move_solo_kernel<1, false, false><<<grid, 32, smem, s>>>(w, g, iteration, env_time);
        else ss_move_solo_kernel<0, false, false><<<grid, 32, smem, s>>>(w, g, iteration, env_time);
    }
    return cudaGetLastError();
}

cudaError_t ss_launch_stats(const DevWorld& w, int stats_index, cudaStream_t s) {
    ss_stats_kernel<<<1, STATS_THREADS, 0, s>>>(w, stats_index);
    return cudaGetLastError();
}

cudaError_t ss_launch_grow(const DevWorld& w, int64_t iteration, int sm_count, int dr_mirror, cudaStream_t s) {
    const size_t cells = (size_t)w.n * w.n;
    uint8_t* mirror = dr_mirror ? w.dr.self.mw : w.isugar;
    if (w.cfg.rule_seasons) {
        if (!w.cfg.rule_grow) return cudaSuccess;
        const bool summer = (iteration % (2 * (int64_t)w.cfg.season_period)) < w.cfg.season_period;
        const double an = summer ? w.cfg.season_on_factor : w.cfg.season_off_factor;
        const double as = summer ? w.cfg.season_off_factor : w.cfg.season_on_factor;
        size_t blocks = (cells + 255) / 256;
        if (blocks > (size_t)sm_count * 16) blocks = (size_t)sm_count * 16;
        ss_seasons_kernel<<<(unsigned)blocks, 256, 0, s>>>(w, an, as, mirror, dr_mirror);
    } else if (w.cfg.rule_grow) {
        size_t blocks = (cells / 4 + 255) / 256;
        if (blocks > (size_t)sm_count * 16) blocks = (size_t)sm_count * 16;
        if (blocks < 1) blocks = 1;
        ss_grow_kernel<<<(unsigned)blocks, 256, 0, s>>>(w.sugar, w.cap, mirror, w.cfg.grow_factor, cells, dr_mirror);
    }
    return cudaGetLastError();
}

size_t ss_diffuse_handoff_bytes(int n) { return sizeof(double) * (size_t)((n + DCOL_ROWS - 1) / DCOL_ROWS) * n; }
cudaError_t ss_launch_diffuse(const DevWorld& w, int* ticket_progress, double* handoff, int impl, cudaStream_t s) {
    const int n = w.n;
    if (impl == 1) {
        ss_diffuse_diag_kernel<<<1, 1024, 0, s>>>(w.poll, n);
        return cudaGetLastError();
    }
    const int nbands = (n + 31) / 32;
    cudaError_t e = cudaMemsetAsync(ticket_progress, 0, sizeof(int) * (size_t)(nbands + 1), s);
    if (e != cudaSuccess) return e;
    if (impl == 2) {        // one warp per band, one band per CTA, every hand-off through global flags
        ss_diffuse_kernel<<<nbands, 32, 0, s>>>(w.poll, n, ticket_progress, ticket_progress + 1);
        return cudaGetLastError();
    }
    static bool configured = false;
    constexpr int S = 2;
    const size_t smem = sizeof(double) * (DCOL_ROWS + 2) * 32 * (S + 3);
    if (!configured) {
        e = cudaFuncSetAttribute(ss_diffuse_col_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int ncol = (n + DCOL_ROWS - 1) / DCOL_ROWS;
    e = cudaMemsetAsync(handoff, 0xff, sizeof(double) * (size_t)ncol * n, s);       // all-ones NaN = "not there yet"
    if (e != cudaSuccess) return e;
    ss_diffuse_col_kernel<S><<<ncol, 32 * (1 + DCOL_H), smem, s>>>(w.poll, n, ticket_progress, handoff);
    return cudaGetLastError();
}

}  // extern "C"
