// ss_lattice.cu -- parallel timestep kernels over the lattice-major cell store (keyed RNG).
//
// One timestep (View.updateGame, sugarscape.py:506-686) =
//   ss_prepare_kernel   Q1 bookkeeping: at-risk agents stamp their successor in the execution
//                       order (sugarscape.py:520 + 502-503: the agent after a dying agent
//                       loses its turn)
//   ss_move_pass_kernel agent phase in dependency rounds (agent.py:361-371 getFood,
//                       agent.py:794-865 move, sugarscape.py:566-601 pollute / metabolise /
//                       starve), repeated over shifted window tilings until no agent is pending
//   ss_stats_kernel     population / means / Gini (sugarscape.py:603-619, 269-280)
//   ss_grow_kernel / ss_seasons_kernel   growback (environment.py:133-155)
//   ss_diffuse_kernel   in-place pollution sweep as a skewed wavefront (environment.py:317-350)
//
// Dependency rounds.  The execution order of a step is the ascending keyed priority of the
// agent slot.  Two agents conflict iff their vision crosses share a cell; every read and
// write of an agent's turn lies inside its own cross, so an agent may take its turn as
// soon as every conflicting agent of lower priority has taken its own -- the result is
// identical to the sequential loop.  A CTA owns one window of the lattice (disjoint
// windows per launch, so CTAs never touch the same cells), stages the window's occupancy
// words in shared memory, derives a pending bitmap, and resolves the agents of the
// window's core (>= 2*vmax from the window edge, so that every possible conflict partner
// is visible) round by round; agents of the outer ring act as blockers and are resolved
// by a later launch whose tiling is shifted by half a window.
#include "ss_world.cuh"

#define FULL 0xffffffffu
#define PASS_THREADS 256
#define PASS_WPT 5                      // bitmap words per thread (covers extents up to 192)

// ------------------------------------------------------------------------ prepare (Q1) ----
__global__ void ss_prepare_kernel(DevWorld w, int64_t iteration) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= w.n_slots) return;
    AgentRec* rec = &w.agents[slot];
    const uint32_t attr = rec->attr;
    if (!attr_alive(attr)) return;
    // safe  <=>  cannot starve this step even with an empty harvest (sugarscape.py:569, 306-309)
    const double after = ss_set_sugar(ss_max0(rec->sugar - (double)attr_metab(attr)));
    if (after > 0.1) return;
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
    a->pending = a->population; a->hist_small = 0; a->inexact = 0;
}

// --------------------------------------------------------------------------- move pass ----
__device__ static inline uint32_t bits_mask(int len) { return len >= 32 ? 0xffffffffu : ((1u << len) - 1u); }

struct WinCtx {
    uint32_t* occ_s;      // [ex][emax]
    uint32_t* pend_s;     // [ex][ew]
    int ex, ey, emax, ew;
    int gx0, gy0;         // global coordinate of local (0,0)
    bool full_x, full_y;  // window spans the whole torus in that dimension (wrap inside)
    int n;
};

// pending bits of local row r, columns [c0, c0+len), len <= 32
__device__ static inline uint32_t pend_bits(const WinCtx& c, int r, int c0, int len) {
    const uint32_t* row = c.pend_s + r * c.ew;
    if (c0 >= 0 && c0 + len <= c.ey) {
        const int w0 = c0 >> 5, sh = c0 & 31;
        const uint32_t lo = row[w0];
        const uint32_t hi = (w0 + 1 < c.ew) ? row[w0 + 1] : 0u;
        return __funnelshift_r(lo, hi, sh) & bits_mask(len);
    }
    uint32_t out = 0;                                   // torus seam inside a full-span window
    for (int i = 0; i < len; i++) {
        int cc = c0 + i;
        if (c.full_y) cc = ss_wrap(cc, c.ey); else if (cc < 0 || cc >= c.ey) continue;
        out |= ((row[cc >> 5] >> (cc & 31)) & 1u) << i;
    }
    return out;
}
__device__ static inline int wrap_row(const WinCtx& c, int r) { return c.full_x ? ss_wrap(r, c.ex) : r; }
__device__ static inline int wrap_col(const WinCtx& c, int q) { return c.full_y ? ss_wrap(q, c.ey) : q; }

// crosses of A (vision a) and B (vision b, displaced by dx,dy) share a cell
__device__ static inline bool crosses_meet(int dx, int dy, int a, int b) {
    const int adx = abs(dx), ady = abs(dy);
    return (adx <= a && ady <= b) || (ady <= a && adx <= b) || (dy == 0 && adx <= a + b) ||
           (dx == 0 && ady <= a + b);
}

// Scan the conflict zone of the agent at local (lx,ly); returns the local position
// (r << 8 | q) + 1 of the pending conflicting agent with the LARGEST priority below `pa`
// (the blocker most likely to resolve last), or 0 if nothing blocks.
__device__ static uint32_t find_blocker(const WinCtx& c, const OrderKey& ok, int lx, int ly, int a, uint32_t pa,
                                        uint32_t self_slot, int V) {
    uint32_t best_p = 0, best_pos = 0;
    bool found = false;
    auto consider = [&](int r, int q, int dx, int dy) {
        const uint32_t ow = c.occ_s[r * c.emax + q];
        const uint32_t sb = occ_slot(ow);
        if (sb == self_slot) return;
        if (!crosses_meet(dx, dy, a, occ_vision(ow))) return;
        const uint32_t pb = ss_priority(ok, sb);
        if (pb < pa && (!found || pb > best_p)) { found = true; best_p = pb; best_pos = ((uint32_t)r << 8 | (uint32_t)q) + 1u; }
    };
    // box |dx| <= V, |dy| <= V
    for (int dx = -V; dx <= V; dx++) {
        const int r = wrap_row(c, lx + dx);
        uint32_t bits = pend_bits(c, r, ly - V, 2 * V + 1);
        while (bits) {
            const int i = __ffs(bits) - 1;
            bits &= bits - 1;
            const int dy = i - V;
            consider(r, wrap_col(c, ly + dy), dx, dy);
        }
    }
    // same x (dx == 0), V < |dy| <= a + V
    {
        const int r = lx;
        uint32_t lo = pend_bits(c, r, ly - V - a, a);
        while (lo) { const int i = __ffs(lo) - 1; lo &= lo - 1; const int dy = i - V - a; consider(r, wrap_col(c, ly + dy), 0, dy); }
        uint32_t hi = pend_bits(c, r, ly + V + 1, a);
        while (hi) { const int i = __ffs(hi) - 1; hi &= hi - 1; const int dy = V + 1 + i; consider(r, wrap_col(c, ly + dy), 0, dy); }
    }
    // same y (dy == 0), V < |dx| <= a + V
    for (int k = V + 1; k <= V + a; k++) {
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const int dx = sgn * k;
            const int r = wrap_row(c, lx + dx);
            if ((c.pend_s[r * c.ew + (ly >> 5)] >> (ly & 31)) & 1u) consider(r, ly, dx, 0);
        }
    }
    return found ? best_pos : 0u;
}

struct ThreadStats { unsigned sum_m, sum_v, acted, deaths, resolved, small, inexact; };

// One agent's whole turn (thread-per-agent): getFood + move + pollute + metabolise + starve.
__device__ static void agent_turn(const DevWorld& w, const WinCtx& c, int lx, int ly, uint32_t ow, int64_t iteration,
                                  int64_t env_time, uint64_t skey, uint32_t next_phase, ThreadStats& ts) {
    const int n = c.n;
    const uint32_t slot = occ_slot(ow);
    const int a = occ_vision(ow);
    AgentRec* recp = &w.agents[slot];
    AgentRec rec = *recp;
    const int gx = ss_wrap(c.gx0 + lx, n), gy = ss_wrap(c.gy0 + ly, n);
    const int pos = gx * n + gy;
    const bool pol = w.cfg.rule_pollution != 0;
    // ---- scan the cross: x-arm ascending raw x, then y-arm (agent.py:361-368) ----
    double best_key = -1.0;
    int best_d = 0, best_sg = 0, best_cell = -1, best_lr = 0, best_lq = 0;
    int ties[8];                 // pre-shuffle indices of the candidates tied with the best
    int tie_cell[8], tie_lr[8], tie_lq[8], tie_sg[8];
    int n_ties = 0, n_free = 0;
    if (w.cfg.rule_move_eat) {
        const int span = 2 * a + 1;
        for (int k = 0; k < 2 * span; k++) {
            int r, q, cx, cy;
            if (k < span) { const int o = k - a; r = wrap_row(c, lx + o); q = ly; cx = ss_wrap(gx + o, n); cy = gy; }
            else { const int o = k - span - a; r = lx; q = wrap_col(c, ly + o); cx = gx; cy = ss_wrap(gy + o, n); }
            if (c.occ_s[r * c.emax + q] != 0u) continue;
            const int cell = cx * n + cy;
            const int sg = (int)__ldcg(&w.sugar[cell]);                              // environment.py:98-100
            const int d = abs(gx - cx) + abs(gy - cy);                               // agent.py:867-868
            const double key = pol ? (double)sg / (1.0 + __ldcg(&w.poll[cell])) : (double)sg;
            const int q_idx = n_free++;
            if (key > best_key || (key == best_key && d < best_d)) {
                best_key = key; best_d = d; n_ties = 0;
            } else if (!(key == best_key && d == best_d)) {
                continue;
            }
            if (n_ties < 8) { ties[n_ties] = q_idx; tie_cell[n_ties] = cell; tie_lr[n_ties] = r; tie_lq[n_ties] = q; tie_sg[n_ties] = sg; }
            n_ties++;
        }
        // own cell, appended last (agent.py:806); distance 0 wins every tie on the key
        const int sg0 = (int)__ldcg(&w.sugar[pos]);
        const double key0 = pol ? (double)sg0 / (1.0 + __ldcg(&w.poll[pos])) : (double)sg0;
        if (n_free == 0 || key0 >= best_key) {
            best_cell = pos; best_sg = sg0; best_lr = lx; best_lq = ly;
        } else {
            int pick = 0;
            if (n_ties > 1) {
                // Tie among free cells: the winner is the first in the Fisher-Yates-shuffled
                // list (agent.py:369, random.py:350).  Replay the agent's keyed draws and track
                // only the tied candidates' positions.
                if (n_ties > 8) n_ties = 8;
                const uint64_t akey = ss_agent_key(skey, slot + 1u);
                uint32_t j = 0;
                int posn[8];
                for (int t = 0; t < n_ties; t++) posn[t] = ties[t];
                for (int i = n_free - 1; i >= 1; i--) {
                    const uint32_t nn = (uint32_t)i + 1u;
                    const int kbits = 32 - __clz(nn);
                    uint32_t rj = ss_keyed_word(akey, j++) >> (32 - kbits);
                    while (rj >= nn) rj = ss_keyed_word(akey, j++) >> (32 - kbits);
                    for (int t = 0; t < n_ties; t++) {
                        if (posn[t] == i) posn[t] = (int)rj;
                        else if (posn[t] == (int)rj) posn[t] = i;
                    }
                }
                for (int t = 1; t < n_ties; t++) if (posn[t] < posn[pick]) pick = t;
            }
            best_cell = tie_cell[pick]; best_lr = tie_lr[pick]; best_lq = tie_lq[pick]; best_sg = tie_sg[pick];
        }
        // ---- commit (agent.py:825-865) ----
        if (pol && w.cfg.pollution_start_time <= env_time)
            w.poll[best_cell] = __ldcg(&w.poll[best_cell]) + (((double)best_sg * w.cfg.pollution_a) + (0.0 * w.cfg.pollution_b));
        rec.sugar = ss_set_sugar(ss_max0(rec.sugar + (double)best_sg));
        w.sugar[best_cell] = 0.0;
    } else {
        best_cell = pos; best_lr = lx; best_lq = ly;
    }
    // ---- sugarscape.py:566-601 ----
    const int metab = attr_metab(rec.attr);
    if (pol && rec.sugar > 0.0 && w.cfg.pollution_start_time <= env_time)
        w.poll[best_cell] = __ldcg(&w.poll[best_cell]) + ((0.0 * w.cfg.pollution_a) + ((double)metab * w.cfg.pollution_b));
    rec.sugar = ss_set_sugar(ss_max0(rec.sugar - (double)metab));
    ts.sum_m += (unsigned)metab;
    ts.sum_v += (unsigned)a;
    ts.acted++;
    ts.resolved++;
    {   // wealth histogram for the Gini coefficient
        const double ws = rec.sugar;
        if (ws == 0.01) ts.small++;
        else if (ws >= 0.0 && ws < (double)SS_GINI_BINS && ws == (double)(int)ws) atomicAdd(&w.hist[(int)ws], 1u);
        else ts.inexact++;
    }
    const bool died = w.cfg.rule_can_starve && rec.sugar <= 0.1;
    const uint32_t new_word = died ? 0u : occ_pack(slot, a, next_phase);
    c.occ_s[lx * c.emax + ly] = 0u;
    c.occ_s[best_lr * c.emax + best_lq] = new_word;
    if (best_cell != pos) w.occw[pos] = 0u;
    w.occw[best_cell] = new_word;
    rec.pos = (uint32_t)best_cell;
    rec.status = ((uint32_t)(iteration + 1) << 2) | (died ? ST_DIED : ST_ACTED);
    if (died) { rec.attr &= ~ATTR_ALIVE; ts.deaths++; }
    recp->sugar = rec.sugar;
    recp->pos = rec.pos;
    recp->attr = rec.attr;
    __threadfence();                       // the turn's effects before its status (Q1 readers)
    *((volatile uint32_t*)&recp->status) = rec.status;
}

__global__ void __launch_bounds__(PASS_THREADS, 2)
ss_move_pass_kernel(DevWorld w, PassGeom g, int64_t iteration, int64_t env_time) {
    extern __shared__ uint32_t smem[];
    __shared__ unsigned red[8];
    const int n = w.n, tid = threadIdx.x, lane = tid & 31;
    __shared__ int go;
    if (tid == 0) go = (w.acc->pending != 0ull);
    __syncthreads();
    if (!go) return;
    // ---- window geometry (balanced boundaries b_k = floor(k*N/nwin), shifted tiling) ----
    const int bx = blockIdx.x % g.nwin, by = blockIdx.x / g.nwin;
    const int x_lo = (int)(((long long)bx * n) / g.nwin), x_hi = (int)(((long long)(bx + 1) * n) / g.nwin);
    const int y_lo = (int)(((long long)by * n) / g.nwin), y_hi = (int)(((long long)(by + 1) * n) / g.nwin);
    WinCtx c;
    c.n = n; c.ex = x_hi - x_lo; c.ey = y_hi - y_lo; c.emax = g.emax; c.ew = g.ew;
    c.gx0 = x_lo + g.shift_x; c.gy0 = y_lo + g.shift_y;
    c.full_x = (c.ex == n); c.full_y = (c.ey == n);
    c.occ_s = smem;
    c.pend_s = smem + g.emax * g.emax;
    const uint32_t cur_phase = (uint32_t)(iteration & 1), next_phase = cur_phase ^ 1u;
    const int V = w.vmax, R = g.ring;
    // ---- stage occupancy words, derive the pending bitmap ----
    const int eyp = c.ew * 32;
    for (int idx = tid; idx < c.ex * eyp; idx += blockDim.x) {
        const int lx = idx / eyp, ly = idx - lx * eyp;
        uint32_t ow = 0u;
        if (ly < c.ey) {
            const int gx = ss_wrap(c.gx0 + lx, n), gy = ss_wrap(c.gy0 + ly, n);
            ow = __ldcg(&w.occw[(size_t)gx * n + gy]);
            c.occ_s[lx * c.emax + ly] = ow;
        }
        const unsigned b = __ballot_sync(FULL, ow != 0u && (ow >> 31) == cur_phase);
        if (lane == 0) c.pend_s[lx * c.ew + (ly >> 5)] = b;
    }
    __syncthreads();
    // ---- my bitmap words; active = pending agents of the core ----
    const uint64_t skey = ss_step_key(w.cfg.keyed_seed, iteration);
    const OrderKey ok = ss_order_key(skey, w.half);
    const int nwords = c.ex * c.ew;
    uint32_t active[PASS_WPT];
    uint16_t* blk_s = (uint16_t*)(c.pend_s + g.emax * g.ew);     // cached blocker per cell
#pragma unroll
    for (int k = 0; k < PASS_WPT; k++) {
        const int wi = tid + k * PASS_THREADS;
        uint32_t m = 0u;
        if (wi < nwords) {
            const int lx = wi / c.ew, wc = wi - lx * c.ew;
            if (c.full_x || (lx >= R && lx < c.ex - R)) {
                m = c.pend_s[wi];
                if (!c.full_y) {
                    const int lo = R - wc * 32, hi = (c.ey - R) - wc * 32;     // core columns [lo, hi)
                    uint32_t cm = 0u;
                    if (hi > 0 && lo < 32) cm = bits_mask(hi < 32 ? hi : 32) & ~bits_mask(lo > 0 ? lo : 0);
                    m &= cm;
                }
            }
            uint32_t t = m;
            while (t) { const int i = __ffs(t) - 1; t &= t - 1; blk_s[lx * c.emax + wc * 32 + i] = 0; }
        }
        active[k] = m;
    }
    ThreadStats ts = {0, 0, 0, 0, 0, 0, 0};
    // ---- dependency rounds ----
    for (;;) {
        uint32_t ready[PASS_WPT];
        int any = 0;
#pragma unroll
        for (int k = 0; k < PASS_WPT; k++) {
            ready[k] = 0u;
            const int wi = tid + k * PASS_THREADS;
            uint32_t t = active[k];
            if (t == 0u) continue;
            const int lx = wi / c.ew, wc = wi - lx * c.ew;
            while (t) {
                const int i = __ffs(t) - 1;
                t &= t - 1;
                const int ly = wc * 32 + i;
                const uint32_t bl = blk_s[lx * c.emax + ly];
                if (bl) {   // cached blocker still pending?
                    const int br = (bl - 1) >> 8, bq = (bl - 1) & 255;
                    if ((c.pend_s[br * c.ew + (bq >> 5)] >> (bq & 31)) & 1u) continue;
                }
                const uint32_t ow = c.occ_s[lx * c.emax + ly];
                const uint32_t slot = occ_slot(ow);
                const uint32_t pa = ss_priority(ok, slot);
                const uint32_t nb = find_blocker(c, ok, lx, ly, occ_vision(ow), pa, slot, V);
                if (nb) {
                    blk_s[lx * c.emax + ly] = (uint16_t)nb;
                    // a blocker outside the core is never resolved by this launch
                    const int br = (nb - 1) >> 8, bq = (nb - 1) & 255;
                    const bool bcore = (c.full_x || (br >= R && br < c.ex - R)) && (c.full_y || (bq >= R && bq < c.ey - R));
                    if (!bcore) active[k] &= ~(1u << i);
                    continue;
                }
                blk_s[lx * c.emax + ly] = 0;
                if (ow & OCC_Q1) {      // Q1: wait for the at-risk predecessor's outcome
                    const AgentRec* rec = &w.agents[slot];
                    if (rec->pred_step == (uint32_t)(iteration + 1)) {
                        const uint32_t st = *((volatile const uint32_t*)&w.agents[rec->pred - 1u].status);
                        if ((st >> 2) != (uint32_t)(iteration + 1)) continue;       // not resolved yet
                    }
                }
                ready[k] |= 1u << i;
                any = 1;
            }
        }
        if (!__syncthreads_or(any)) break;
#pragma unroll
        for (int k = 0; k < PASS_WPT; k++) {
            uint32_t t = ready[k];
            if (t == 0u) continue;
            const int wi = tid + k * PASS_THREADS;
            const int lx = wi / c.ew, wc = wi - lx * c.ew;
            active[k] &= ~t;
            c.pend_s[wi] &= ~t;
            while (t) {
                const int i = __ffs(t) - 1;
                t &= t - 1;
                const int ly = wc * 32 + i;
                const uint32_t ow = c.occ_s[lx * c.emax + ly];
                bool skipped = false;
                if (ow & OCC_Q1) {
                    const uint32_t slot = occ_slot(ow);
                    const AgentRec* rec = &w.agents[slot];
                    if (rec->pred_step == (uint32_t)(iteration + 1)) {
                        const uint32_t st = *((volatile const uint32_t*)&w.agents[rec->pred - 1u].status);
                        skipped = ((st & 3u) == ST_DIED);
                    }
                    if (skipped) {      // loses its turn: stays put, no metabolism (Q1)
                        const uint32_t nw = occ_pack(slot, occ_vision(ow), next_phase);
                        c.occ_s[lx * c.emax + ly] = nw;
                        const int gx = ss_wrap(c.gx0 + lx, n), gy = ss_wrap(c.gy0 + ly, n);
                        w.occw[(size_t)gx * n + gy] = nw;
                        *((volatile uint32_t*)&w.agents[slot].status) = ((uint32_t)(iteration + 1) << 2) | ST_SKIPPED;
                        ts.resolved++;
                    }
                }
                if (!skipped) agent_turn(w, c, lx, ly, ow & ~OCC_Q1, iteration, env_time, skey, next_phase, ts);
            }
        }
        __syncthreads();
    }
    // ---- block reduction of the step accumulators ----
    unsigned vals[7] = {ts.sum_m, ts.sum_v, ts.acted, ts.deaths, ts.resolved, ts.small, ts.inexact};
    unsigned long long* dst[7] = {&w.acc->sum_metab, &w.acc->sum_vision, &w.acc->acted, &w.acc->deaths,
                                  nullptr, &w.acc->hist_small, &w.acc->inexact};
#pragma unroll
    for (int q = 0; q < 7; q++) {
        unsigned v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            unsigned long long s = 0;
            for (int i = 0; i < PASS_THREADS / 32; i++) s += red[i];
            if (s) {
                if (q == 4) atomicAdd(&w.acc->pending, (unsigned long long)(-(long long)s));
                else atomicAdd(dst[q], s);
            }
        }
        __syncthreads();
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
        if (a->inexact) gini = __longlong_as_double(0x7ff8000000000000ll);
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
__global__ void __launch_bounds__(256)
ss_grow_kernel(double* __restrict__ sugar, const double* __restrict__ cap, double alpha, size_t cells) {
    const size_t nvec = cells / 2;
    double2* s2 = reinterpret_cast<double2*>(sugar);
    const double2* c2 = reinterpret_cast<const double2*>(cap);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        double2 s[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { s[u] = s2[i + u * stride]; c[u] = __ldcs(&c2[i + u * stride]); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            double a = s[u].x + alpha, b = s[u].y + alpha;
            s[u].x = a <= c[u].x ? a : c[u].x;
            s[u].y = b <= c[u].y ? b : c[u].y;
            s2[i + u * stride] = s[u];
        }
    }
    for (; i < nvec; i += stride) {
        double2 s = s2[i], c = __ldcs(&c2[i]);
        double a = s.x + alpha, b = s.y + alpha;
        s.x = a <= c.x ? a : c.x;
        s.y = b <= c.y ? b : c.y;
        s2[i] = s;
    }
    if ((cells & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double a = sugar[cells - 1] + alpha, c = cap[cells - 1];
        sugar[cells - 1] = a <= c ? a : c;
    }
}

// sugarscape.py:642-655 + environment.py:146-155: two inclusive rectangles, north first
__global__ void __launch_bounds__(256)
ss_seasons_kernel(DevWorld w, double alpha_north, double alpha_south) {
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
        if (in_n) { const double v = s + alpha_north; s = v <= cp ? v : cp; }
        if (in_s) { const double v = s + alpha_south; s = v <= cp ? v : cp; }
        w.sugar[c] = s;
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
#define DIFF_RING 96
// x / 3.0, correctly rounded, without the long DDIV sequence on the wavefront's critical
// path: q = RN(x*r), one FMA residual correction (Markstein); exact IEEE division outside the
// comfortably-normal range.  Checked against x / 3.0 on the CPU (tests/test_div3.py).
__device__ static inline double div3(double x) {
    const double ax = fabs(x);
    if (!(ax > 1e-280 && ax < 1e280)) return x / 3.0;
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
    const int nchunks = (n + 31) / 32;
    const int nblocks = (n + 31 + 31) / 32;          // iterations 0 .. n+30
    auto load_chunk = [&](int ch, bool with_above) { // columns [32ch, 32ch+32)
        const int col = 32 * ch + lane;
        if (ch >= nchunks || col >= n) return;
        const int slot = col % DIFF_RING;
#pragma unroll 8
        for (int rr = 1; rr <= 33; rr++) {
            const int gr = x0 - 1 + rr;
            if (rr <= rows || (rr == 33 && has_below)) tile[rr][slot] = __ldcg(&p[(size_t)gr * n + col]);
        }
        (void)with_above;
    };
    auto load_above = [&](int ch) {
        const int col = 32 * ch + lane;
        if (!has_above || ch >= nchunks) return;
        const int need = min(n, 32 * (ch + 1));
        if (lane == 0) while (ld_acquire(&progress[b - 1]) < need) { }
        __syncwarp();
        if (col < n) tile[0][col % DIFF_RING] = __ldcg(&p[(size_t)(x0 - 1) * n + col]);
    };
    auto flush_chunk = [&](int ch) {
        if (ch < 0 || ch >= nchunks) return;
        const int col = 32 * ch + lane;
        if (col < n) {
            const int slot = col % DIFF_RING;
            for (int rr = 1; rr <= rows; rr++) p[(size_t)(x0 - 1 + rr) * n + col] = tile[rr][slot];
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(&progress[b], min(n, 32 * (ch + 1)));
    };
    load_chunk(0, false);
    load_chunk(1, false);
    double res_prev = 0.0;
    for (int j = 0; j < nblocks; j++) {
        if (j >= 1) load_chunk(j + 1, false);
        load_above(j);
        __syncwarp();
        for (int k = 32 * j; k < 32 * j + 32; k++) {
            const int y = k - lane;
            const bool act = row_ok && y >= 0 && y < n;
            const bool has_n = row_ok && (y + 1 >= 0) && (y + 1 < n);
            const double n_val = has_n ? tile[lane + 1][(y + 1) % DIFF_RING] : 0.0;
            double e_val = __shfl_down_sync(FULL, n_val, 1);
            if (lane == 31 && act && has_below) e_val = tile[33][y % DIFF_RING];
            double w_val = __shfl_up_sync(FULL, res_prev, 1);
            if (lane == 0 && act && has_above) w_val = tile[0][y % DIFF_RING];
            if (act) {
                double t = 0.0;
                int cnt = 0;
                if (y + 1 < n) { t += n_val; cnt++; }
                if (y - 1 >= 0) { t += res_prev; cnt++; }
                if (r + 1 < n) { t += e_val; cnt++; }
                if (r - 1 >= 0) { t += w_val; cnt++; }
                const double res = cnt == 4 ? t * 0.25 : (cnt == 2 ? t * 0.5 : div3(t));
                tile[lane + 1][y % DIFF_RING] = res;
                res_prev = res;
            }
        }
        __syncwarp();
        flush_chunk(j - 1);
    }
    flush_chunk(nblocks - 1);
    __syncwarp();
    if (lane == 0) st_release(&progress[b], n);
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

// ---------------------------------------------------------------------------- launchers ----
extern "C" {

cudaError_t ss_launch_prepare(const DevWorld& w, int64_t iteration, cudaStream_t s) {
    ss_begin_step_kernel<<<1, 1, 0, s>>>(w);
    if (w.cfg.rule_can_starve && w.n_slots > 0)
        ss_prepare_kernel<<<(w.n_slots + 255) / 256, 256, 0, s>>>(w, iteration);
    return cudaGetLastError();
}

size_t ss_pass_smem_bytes(int emax, int ew) {
    return (size_t)emax * emax * 4 + (size_t)emax * ew * 4 + (size_t)emax * emax * 2;
}

cudaError_t ss_launch_move_pass(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time,
                                cudaStream_t s) {
    static size_t configured = 0;
    const size_t smem = ss_pass_smem_bytes(g.emax, g.ew);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ss_move_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    ss_move_pass_kernel<<<g.nwin * g.nwin, PASS_THREADS, smem, s>>>(w, g, iteration, env_time);
    return cudaGetLastError();
}

cudaError_t ss_launch_stats(const DevWorld& w, int stats_index, cudaStream_t s) {
    ss_stats_kernel<<<1, STATS_THREADS, 0, s>>>(w, stats_index);
    return cudaGetLastError();
}

cudaError_t ss_launch_grow(const DevWorld& w, int64_t iteration, int sm_count, cudaStream_t s) {
    const size_t cells = (size_t)w.n * w.n;
    if (w.cfg.rule_seasons) {
        if (!w.cfg.rule_grow) return cudaSuccess;
        const bool summer = (iteration % (2 * (int64_t)w.cfg.season_period)) < w.cfg.season_period;
        const double an = summer ? w.cfg.season_on_factor : w.cfg.season_off_factor;
        const double as = summer ? w.cfg.season_off_factor : w.cfg.season_on_factor;
        size_t blocks = (cells + 255) / 256;
        if (blocks > (size_t)sm_count * 16) blocks = (size_t)sm_count * 16;
        ss_seasons_kernel<<<(unsigned)blocks, 256, 0, s>>>(w, an, as);
    } else if (w.cfg.rule_grow) {
        size_t blocks = (cells / 2 + 255) / 256;
        if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
        if (blocks < 1) blocks = 1;
        ss_grow_kernel<<<(unsigned)blocks, 256, 0, s>>>(w.sugar, w.cap, w.cfg.grow_factor, cells);
    }
    return cudaGetLastError();
}

cudaError_t ss_launch_diffuse(const DevWorld& w, int* ticket_progress, int impl, cudaStream_t s) {
    const int n = w.n;
    if (impl == 1) {
        ss_diffuse_diag_kernel<<<1, 1024, 0, s>>>(w.poll, n);
        return cudaGetLastError();
    }
    const int nbands = (n + 31) / 32;
    cudaError_t e = cudaMemsetAsync(ticket_progress, 0, sizeof(int) * (size_t)(nbands + 1), s);
    if (e != cudaSuccess) return e;
    ss_diffuse_kernel<<<nbands, 32, 0, s>>>(w.poll, n, ticket_progress, ticket_progress + 1);
    return cudaGetLastError();
}

}  // extern "C"
