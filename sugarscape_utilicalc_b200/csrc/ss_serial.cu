// ss_serial.cu -- serial-order timestep kernel (one CTA).
//
// Reproduces the reference's sequential agent loop exactly, including CPython's global
// MT19937 stream (SS_RNG_MT19937) -- the mode that is bit-exact against the unmodified
// reference on its shipped 50x50 presets -- and also serves as the universal fallback for
// keyed runs the lattice path refuses.  One launch runs `nsteps` whole timesteps
// (View.updateGame, sugarscape.py:506-686) back to back: order shuffle, agent loop,
// statistics, growback / seasons, pollution diffusion.  With a single CTA on one SM the
// 50x50 working set (70 KB of cells + 10 KB of agent records) stays L1-resident across
// steps, so the kernel addresses the canonical HBM arrays directly.
//
// Warp 0 walks the execution order; its 32 lanes scan an agent's cross in parallel
// (ballot-compacted candidate list, warp-shuffle argmax); lane 0 performs the inherently
// sequential parts (Fisher-Yates shuffles on the RNG stream, the commit).  All warps join
// for the order sort (keyed mode), the Gini sort and the environment phase.
#include "ss_world.cuh"
#include "ss_life.cuh"
#define SS_POW_TABLE_QUAL __device__ const
#include "ss_pow.h"          // x ** y of the reference, bit for bit

#define SERIAL_THREADS 256
#define FULL 0xffffffffu

struct SerialShared {
    uint32_t mt[624];
    int mti;
    int rng_mode;
    unsigned long long skey, akey;
    uint32_t j;
    int32_t food[SS_MAXC];
    int32_t neigh[SS_MAXC];
    double util[SS_MAXC];
    double nb_inten[SS_MAXC];
    double nb_metab[SS_MAXC];
    int32_t nb_x[SS_MAXC], nb_y[SS_MAXC], nb_v[SS_MAXC];
    double sum_m, sum_m2, sum_v;
    int nw, kept, acted, deaths;
};

// ---------------------------------------------------------------- RNG (lane 0 of warp 0) ----
__device__ static uint32_t mt_next(SerialShared& sh) {
    uint32_t* mt = sh.mt;
    if (sh.mti >= 624) {
        int kk;
        uint32_t y;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        sh.mti = 0;
    }
    uint32_t y = mt[sh.mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ static inline uint32_t rng_word(SerialShared& sh) {
    if (sh.rng_mode == SS_RNG_MT19937) return mt_next(sh);
    return ss_keyed_word(sh.akey, sh.j++);
}
// random.py:242 _randbelow_with_getrandbits
__device__ static uint32_t randbelow(SerialShared& sh, uint32_t n) {
    int k = 32 - __clz(n);
    uint32_t r = rng_word(sh) >> (32 - k);
    while (r >= n) r = rng_word(sh) >> (32 - k);
    return r;
}
// agent.py:224-229 setSugar: max(amount, 0 if limitedLifespan else .01)
__device__ static inline double set_sugar_floor(double amount, double fl) { return amount > fl ? amount : fl; }
// random.py:350 shuffle
__device__ static void shuffle_i32(SerialShared& sh, int32_t* a, int n) {
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = randbelow(sh, (uint32_t)i + 1u);
        int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}

// ------------------------------------------------------------------ CTA-wide bitonic sort ----
template <class T>
__device__ static void cta_bitonic_sort(T* a, int P) {
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool asc = ((i & k) == 0);
                    T u = a[i], v = a[ixj];
                    if ((u > v) == asc) { a[i] = v; a[ixj] = u; }
                }
            }
            __syncthreads();
        }
}
__device__ static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// --------------------------------------------------------------------- warp-level scans ----
// agent.py:361-371 getFood (free cells, x-arm then y-arm, torus) / agent.py:391-407
// getVisionNeighbourhood (occupants, own wrapped coordinate excluded).  Unshuffled.
template <bool NEIGH>
__device__ static int warp_scan_cross(const DevWorld& w, int x, int y, int v, int32_t* out, int lane) {
    const int n = w.n, span = 2 * v + 1, K = 2 * span;
    int m = 0;
    for (int base = 0; base < K; base += 32) {
        int k = base + lane;
        bool take = false;
        int val = 0;
        if (k < K) {
            int c, own;
            if (k < span) { int wx = ss_wrap(x - v + k, n); c = wx * n + y; own = (wx == x); }
            else { int wy = ss_wrap(y - v + (k - span), n); c = x * n + wy; own = (wy == y); }
            uint32_t ow = w.occw[c];
            if (NEIGH) { take = (ow != 0u) && !own; val = (int)occ_slot(ow); }
            else { take = (ow == 0u); val = c; }
        }
        unsigned b = __ballot_sync(FULL, take);
        if (take) out[m + __popc(b & ((1u << lane) - 1u))] = val;
        m += __popc(b);
    }
    __syncwarp();
    return m;
}

struct Best { double key; int d, p, c, sg; };
__device__ static inline bool better(const Best& a, const Best& b) {
    if (a.p < 0) return false;
    if (b.p < 0) return true;
    if (a.key != b.key) return a.key > b.key;
    if (a.d != b.d) return a.d < b.d;
    return a.p < b.p;
}

// environment.py:312-315
__device__ static inline void pollute_site(const DevWorld& w, int64_t env_time, int c, double production,
                                           double consumption) {
    if (w.cfg.pollution_start_time <= env_time)
        w.poll[c] += (production * w.cfg.pollution_a) + (consumption * w.cfg.pollution_b);
}

// agent.py:794-865 move(), no-spice branch
__device__ static void warp_move(const DevWorld& w, SerialShared& sh, int s, int64_t env_time, int lane, double fl) {
    const int n = w.n;
    AgentRec* rec = &w.agents[s];
    const int pos = (int)rec->pos, x = pos / n, y = pos % n, v = attr_vision(rec->attr);
    int m = warp_scan_cross<false>(w, x, y, v, sh.food, lane);
    if (lane == 0) { shuffle_i32(sh, sh.food, m); sh.food[m] = pos; }    // agent.py:369, 806
    m += 1;
    __syncwarp();
    Best best; best.p = -1; best.key = 0; best.d = 0; best.c = 0; best.sg = 0;
    for (int p = lane; p < m; p += 32) {
        Best b;
        b.c = sh.food[p]; b.p = p;
        b.sg = (int)w.sugar[b.c];                                       // environment.py:98-100
        b.d = abs(x - b.c / n) + abs(y - b.c % n);                      // agent.py:867-868
        b.key = w.cfg.rule_pollution ? (double)b.sg / (1.0 + w.poll[b.c]) : (double)b.sg;
        if (better(b, best)) best = b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best b;
        b.key = __shfl_xor_sync(FULL, best.key, o);
        b.d = __shfl_xor_sync(FULL, best.d, o);
        b.p = __shfl_xor_sync(FULL, best.p, o);
        b.c = __shfl_xor_sync(FULL, best.c, o);
        b.sg = __shfl_xor_sync(FULL, best.sg, o);
        if (better(b, best)) best = b;
    }
    if (lane == 0) {
        if (w.cfg.rule_pollution) pollute_site(w, env_time, best.c, (double)best.sg, 0.0);  // agent.py:825
        rec->sugar = set_sugar_floor(ss_max0(rec->sugar + (double)best.sg), fl);           // agent.py:832
        if (best.c != pos) {                                                                // agent.py:856-862
            uint32_t ow = w.occw[pos];
            w.occw[pos] = 0u;
            w.occw[best.c] = ow;
            rec->pos = (uint32_t)best.c;
        }
        w.sugar[best.c] = 0.0;                                                              // agent.py:864
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// agent.py:794-865 move(), spice branch (833-853) with the Cobb-Douglas welfare of agent.py:1220-1251 (no tags, no
// foresight): w1 ** (m1 / mt) * w2 ** (m2 / mt) in the wealth the agent would have on the cell -- `if x and y:` (Q9): a cell on
// row 0 or column 0 does not count its resources.  The arg-max compares these floats, so x ** y is the reference's libm pow
// bit for bit (ss_pow.h); float_pow answers y == 0 and x == 1 itself.  One lane per candidate.
__device__ static inline double spice_pow(const SpiceWorld& sw, double x, double y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    ss_pow_tables T;
    T.log_hdr = SS_POW_LOG_HDR; T.log_tab = SS_POW_LOG_TAB; T.exp_hdr = SS_POW_EXP_HDR; T.exp_tab = SS_POW_EXP_TAB;
    int bad = 0;
    const double r = ss_pow_pos(&T, x, y, &bad);
    if (bad) atomicAdd(sw.pow_faults, 1);
    return r;
}
__device__ static void warp_move_spice(const DevWorld& w, const SpiceWorld& sw, SerialShared& sh, int s, int lane, double fl) {
    const int n = w.n;
    AgentRec* rec = &w.agents[s];
    const int pos = (int)rec->pos, x = pos / n, y = pos % n, v = attr_vision(rec->attr);
    int m = warp_scan_cross<false>(w, x, y, v, sh.food, lane);
    if (lane == 0) { shuffle_i32(sh, sh.food, m); sh.food[m] = pos; }    // agent.py:369, 806
    m += 1;
    __syncwarp();
    const double own1 = rec->sugar, own2 = sw.aspice[s];
    const int im1 = attr_metab(rec->attr), im2 = sw.asp_metab[s];
    const double e1 = (double)im1 / (double)(im1 + im2), e2 = (double)im2 / (double)(im1 + im2);
    Best best; best.p = -1; best.key = 0; best.d = 0; best.c = 0; best.sg = 0;
    for (int p = lane; p < m; p += 32) {
        Best b;
        b.c = sh.food[p]; b.p = p; b.sg = 0;
        const int cx = b.c / n, cy = b.c % n;
        double w1 = own1, w2 = own2;
        if (cx != 0 && cy != 0) { w1 = w1 + (double)(int)w.sugar[b.c]; w2 = w2 + (double)(int)sw.spice[b.c]; }
        b.key = (w1 > 0.0 && w2 > 0.0) ? spice_pow(sw, w1, e1) * spice_pow(sw, w2, e2) : 0.0;
        b.d = abs(x - cx) + abs(y - cy);                               // agent.py:867-868
        if (better(b, best)) best = b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best b;
        b.key = __shfl_xor_sync(FULL, best.key, o);
        b.d = __shfl_xor_sync(FULL, best.d, o);
        b.p = __shfl_xor_sync(FULL, best.p, o);
        b.c = __shfl_xor_sync(FULL, best.c, o);
        b.sg = 0;
        if (better(b, best)) best = b;
    }
    if (lane == 0) {
        const int sg = (int)w.sugar[best.c], sp = (int)sw.spice[best.c];
        sw.aspice[s] = set_sugar_floor(ss_max0(sw.aspice[s] + (double)sp), fl);            // agent.py:851
        rec->sugar = set_sugar_floor(ss_max0(rec->sugar + (double)sg), fl);                // agent.py:852
        if (best.c != pos) {
            const uint32_t ow = w.occw[pos];
            w.occw[pos] = 0u;
            w.occw[best.c] = ow;
            rec->pos = (uint32_t)best.c;
        }
        w.sugar[best.c] = 0.0;                                                              // agent.py:864-865
        sw.spice[best.c] = 0.0;
    }
    __syncwarp();
}

// agent.py:1025-1141 utilicalcMove(), sugar-only (utilicalcSugar agent.py:980-992)
__device__ static void warp_utilicalc(const DevWorld& w, SerialShared& sh, int s, int lane) {
    const int n = w.n;
    AgentRec* rec = &w.agents[s];
    const int pos = (int)rec->pos, x = pos / n, y = pos % n, v = attr_vision(rec->attr);
    int m = warp_scan_cross<false>(w, x, y, v, sh.food, lane);
    if (lane == 0) {
        shuffle_i32(sh, sh.food, m);          // getFood's shuffle, agent.py:369
        shuffle_i32(sh, sh.food, m);          // agent.py:1033
        sh.food[m] = pos;                     // agent.py:1035
    }
    m += 1;
    int nb = warp_scan_cross<true>(w, x, y, v, sh.neigh, lane);
    if (lane == 0) { shuffle_i32(sh, sh.neigh, nb); sh.neigh[nb] = s; }   // agent.py:405, 1039
    nb += 1;
    __syncwarp();
    for (int q = lane; q < nb; q += 32) {
        const AgentRec* b = &w.agents[sh.neigh[q]];
        double bm = (double)attr_metab(b->attr);
        double days = ss_py_floordiv(b->sugar, bm);                      // agent.py:870-872
        sh.nb_inten[q] = 1.0 / (1.0 + days);
        sh.nb_metab[q] = bm;
        sh.nb_x[q] = (int)b->pos / n;
        sh.nb_y[q] = (int)b->pos % n;
        sh.nb_v[q] = attr_vision(b->attr);
    }
    __syncwarp();
    const double extent = (double)(nb - 1);
    double lmax = 0.0;
    bool have = false;
    for (int p = lane; p < m; p += 32) {
        int c = sh.food[p];
        int mx = c / n, my = c % n;
        double site = (double)(int)w.sugar[c];
        double u = 0.0;
        for (int q = 0; q < nb; q++) {
            double dur = (site / sh.nb_metab[q]) / w.cfg.max_capacity;
            int bx = sh.nb_x[q], by = sh.nb_y[q], bv = sh.nb_v[q];
            int cert = (bx == mx && abs(by - my) <= bv) || (by == my && abs(bx - mx) <= bv);
            double t = 0.0;
            t += (double)cert * (((sh.nb_inten[q] + dur) + 0.5 * 0.0) + extent);
            if (sh.neigh[q] != s) t *= -1.0;
            u += t;
        }
        sh.util[p] = u;
        if (!have || u > lmax) { lmax = u; have = true; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double om = __shfl_xor_sync(FULL, lmax, o);
        int oh = __shfl_xor_sync(FULL, (int)have, o);
        if (oh && (!have || om > lmax)) { lmax = om; have = true; }
    }
    __syncwarp();
    if (lane == 0) {
        int nmax = 0;
        for (int p = 0; p < m; p++)
            if (sh.util[p] == lmax) sh.neigh[nmax++] = sh.food[p];      // neigh[] reused as max_locations
        int dest = sh.neigh[randbelow(sh, (uint32_t)nmax)];              // random.choice, agent.py:1108
        uint32_t ow = w.occw[pos];                                       // agent.py:1121-1125
        w.occw[pos] = 0u;
        w.occw[dest] = ow;
        rec->pos = (uint32_t)dest;
        int sg = (int)w.sugar[dest];                                     // agent.py:1128-1130
        rec->sugar = ss_set_sugar(ss_max0(rec->sugar + (double)sg));
        w.sugar[dest] = 0.0;                                             // agent.py:1134
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// procreation (sugarscape.py:543-554; agent.py:653-679 mate, 681-696 findFreeLocationAround, 698-750 createChild,
// 764-773 isFertile, 373-389 getNeighbourhood).  At most four neighbours: lane 0 does all of it, every draw on the MT19937 stream
// in the reference's order.
__device__ static inline bool life_fertile(const DevWorld& w, const LifeWorld& lw, const SpiceWorld& sw, int s) {
    const LifeRec& r = lw.rec[s];
    return r.fert_lo <= r.age && r.age <= r.fert_hi && w.agents[s].sugar >= r.endow &&
           (!w.cfg.rule_spice || sw.aspice[s] >= sw.sp_endow[s]);
}
__device__ static int life_free_around(const DevWorld& w, SerialShared& sh, int x, int y) {
    const int n = w.n;
    int loc[6], m = 0;
    for (int i = x - 1; i <= x + 1; i++)                                  // no torus: isLocationValid (environment.py:165-167)
        if (i >= 0 && i < n && w.occw[i * n + y] == 0u) loc[m++] = i * n + y;
    for (int j = y - 1; j <= y + 1; j++)
        if (j >= 0 && j < n && w.occw[x * n + j] == 0u) loc[m++] = x * n + j;
    if (!m) return -1;
    return loc[randbelow(sh, (uint32_t)m)];                               // random.randint(0, length - 1)
}
__device__ static inline double life_random(SerialShared& sh) {           // random.random(): 53 bits of two words
    const uint32_t a = rng_word(sh) >> 5, b = rng_word(sh) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}
__device__ static inline void life_event(const LifeWorld& lw, int step, int kind, int a, int b, int c) {
    const int at = *lw.n_events;
    if (at >= lw.ev_cap) { lw.status[1]++; return; }
    int32_t* e = lw.events + 4 * (size_t)at;
    e[0] = kind | (step << 8); e[1] = a; e[2] = b; e[3] = c;
    *lw.n_events = at + 1;
}
// returns the new length of the list (the child joins its end and is met by the running loop in this very step)
__device__ static int life_procreate(const DevWorld& w, const LifeWorld& lw, const SpiceWorld& sw, SerialShared& sh, int s, int n_order,
                                     int step, double fl) {
    const int n = w.n;
    const int pos = (int)w.agents[s].pos, x = pos / n, y = pos % n;
    int32_t nb[4];
    int m = 0;
    for (int xx = x - 1; xx <= x + 1; xx++) {                             // getNeighbourhood: torus, x-line then y-line
        if (xx == x) continue;
        const uint32_t ow = w.occw[ss_wrap(xx, n) * n + y];
        if (ow != 0u) nb[m++] = (int32_t)occ_slot(ow);
    }
    for (int yy = y - 1; yy <= y + 1; yy++) {
        if (yy == y) continue;
        const uint32_t ow = w.occw[x * n + ss_wrap(yy, n)];
        if (ow != 0u) nb[m++] = (int32_t)occ_slot(ow);
    }
    shuffle_i32(sh, nb, m);
    for (int k = 0; k < m; k++) {
        const int b = nb[k];
        if (lw.rec[b].sex == lw.rec[s].sex || !life_fertile(w, lw, sw, b)) continue;
        int loc = life_free_around(w, sh, x, y);
        if (loc < 0) { const int pb = (int)w.agents[b].pos; loc = life_free_around(w, sh, pb / n, pb % n); }
        if (loc < 0 || !life_fertile(w, lw, sw, s)) continue;
        // createChild: genitors = [self, partner]
        const int g[2] = {s, b};
        const int metab = attr_metab(w.agents[g[randbelow(sh, 2u)]].attr);
        const int gsp = g[randbelow(sh, 2u)];                              // spiceMetabolism (drawn with or without the spice rule)
        const int vision = attr_vision(w.agents[g[randbelow(sh, 2u)]].attr);
        const double endow = 0.5 * (lw.rec[s].endow + lw.rec[b].endow);
        w.agents[s].sugar = set_sugar_floor(ss_max0(w.agents[s].sugar - 0.5 * lw.rec[s].endow), fl);
        w.agents[b].sugar = set_sugar_floor(ss_max0(w.agents[b].sugar - 0.5 * lw.rec[b].endow), fl);
        int sp_metab = 0;
        double sp_endow = 0.0;
        if (w.cfg.rule_spice) {                                           // agent.py:712-714
            sp_metab = sw.asp_metab[gsp];
            sp_endow = 0.5 * (sw.sp_endow[s] + sw.sp_endow[b]);
            sw.aspice[s] = set_sugar_floor(ss_max0(sw.aspice[s] - 0.5 * sw.sp_endow[s]), fl);
            sw.aspice[b] = set_sugar_floor(ss_max0(sw.aspice[b] - 0.5 * sw.sp_endow[b]), fl);
        }
        const int max_age = lw.rec[g[randbelow(sh, 2u)]].max_age;
        const int sex = (int)randbelow(sh, 2u);
        const int fg = g[randbelow(sh, 2u)];
        int32_t mask = 1, tags = 0;
        for (int t = 0; t < w.cfg.tags_length; t++) {                     // agent.py:727-737
            const int32_t t1 = lw.rec[s].tags & mask, t2 = lw.rec[b].tags & mask;
            if (t1 == t2 || life_random(sh) < 0.5) tags |= t1; else tags |= t2;
            mask <<= 1;
        }
        const int c = *lw.id_increment;                                   // slot = id - 1 (Environment.getNewId)
        (void)randbelow(sh, (uint32_t)(w.cfg.foresight_hi - w.cfg.foresight_lo + 1));      // Agent.__init__: generateForesight
        if (c >= lw.capacity) { lw.status[1]++; continue; }               // (the launch stops before a step that could get here)
        *lw.id_increment = c + 1;
        AgentRec a;
        a.sugar = set_sugar_floor(endow, fl);                             // setInitialSugarEndowment
        a.pos = (uint32_t)loc; a.attr = attr_pack(vision, metab, true);
        a.status = 0u; a.pred = 0u; a.pred_step = 0u; a.turn = 0u;
        w.agents[c] = a;
        LifeRec r;
        r.age = 0; r.max_age = max_age; r.sex = sex; r.fert_lo = lw.rec[fg].fert_lo; r.fert_hi = lw.rec[fg].fert_hi;
        r.tags = tags; r.living = 1; r.pad = 0; r.endow = endow;
        lw.rec[c] = r;
        if (w.cfg.rule_spice) { sw.asp_metab[c] = sp_metab; sw.sp_endow[c] = sp_endow; sw.aspice[c] = set_sugar_floor(sp_endow, fl); }
        w.occw[loc] = occ_pack((uint32_t)c, vision, 0u);
        w.order[n_order++] = c;                                           // sugarscape.py:550
        life_event(lw, step, 1, s + 1, b + 1, c + 1);
    }
    return n_order;
}

// ---------------------------------------------------------------- environment phase (CTA) ----
__device__ static inline void grow_cell(const DevWorld& w, const SpiceWorld& sw, int c, double alpha) {      // environment.py:133-142
    double v = w.sugar[c] + alpha;
    double cp = w.cap[c];
    w.sugar[c] = v <= cp ? v : cp;
    if (sw.spice) {
        const double u = sw.spice[c] + alpha, cs = sw.spice_cap[c];
        sw.spice[c] = u <= cs ? u : cs;
    }
}
__device__ static inline bool in_region(const int32_t r[4], int n, int x, int y) {     // environment.py:146-155
    int imin = max(r[0], 0), jmin = max(r[1], 0), imax = min(r[2] + 1, n), jmax = min(r[3] + 1, n);
    return x >= imin && x < imax && y >= jmin && y < jmax;
}
__device__ static void cta_environment(const DevWorld& w, const SpiceWorld& sw, int64_t iteration) {
    const int n = w.n, cells = n * n;
    if (w.cfg.rule_seasons) {                                             // sugarscape.py:642-655
        if (w.cfg.rule_grow) {
            bool summer = (iteration % (2 * (int64_t)w.cfg.season_period)) < w.cfg.season_period;
            double an = summer ? w.cfg.season_on_factor : w.cfg.season_off_factor;
            double as = summer ? w.cfg.season_off_factor : w.cfg.season_on_factor;
            for (int c = threadIdx.x; c < cells; c += blockDim.x) {
                int x = c / n, y = c % n;
                if (in_region(w.cfg.season_north, n, x, y)) grow_cell(w, sw, c, an);
                if (in_region(w.cfg.season_south, n, x, y)) grow_cell(w, sw, c, as);
            }
        }
    } else if (w.cfg.rule_grow) {                                         // sugarscape.py:657-659
        for (int c = threadIdx.x; c < cells; c += blockDim.x) grow_cell(w, sw, c, w.cfg.grow_factor);
    }
    __syncthreads();
    // sugarscape.py:661-663 + environment.py:317-350: in-place sweep == anti-diagonal wavefront
    if (w.cfg.rule_pollution && iteration >= w.cfg.diffusion_start_time &&
        iteration % w.cfg.diffusion_rate == 0) {
        double* p = w.poll;
        for (int d = 0; d <= 2 * n - 2; d++) {
            int x0 = max(0, d - n + 1), x1 = min(d, n - 1);
            for (int x = x0 + (int)threadIdx.x; x <= x1; x += blockDim.x) {
                int y = d - x;
                double t = 0.0;
                int c = 0;
                if (y + 1 < n) { t += p[x * n + y + 1]; c++; }
                if (y - 1 >= 0) { t += p[x * n + y - 1]; c++; }
                if (x + 1 < n) { t += p[(x + 1) * n + y]; c++; }
                if (x - 1 >= 0) { t += p[(x - 1) * n + y]; c++; }
                p[x * n + y] = t / (double)c;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------ kernel ----
__global__ void __launch_bounds__(SERIAL_THREADS, 1)
ss_serial_kernel(DevWorld w, LifeWorld lw, SpiceWorld sw, int64_t iteration0, int64_t env_time0, int nsteps, int fused_env, int stats_base) {
    __shared__ SerialShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 624; i += blockDim.x) sh.mt[i] = w.mt[i];
    if (tid == 0) { sh.mti = *w.mti; sh.rng_mode = w.cfg.rng_mode; }
    __syncthreads();
    const OrderKey dummy = ss_order_key(0, w.half);
    const bool life = lw.rec != nullptr;
    const double fl = w.cfg.rule_limited_lifespan ? 0.0 : 0.01;           // agent.py:226
    int step = 0;
    for (; step < nsteps; step++) {
        const int64_t iteration = iteration0 + step, env_time = env_time0 + step;
        int n_order = *w.n_order;
        // births: a turn creates at most four children and logs at most five events; the launch stops before a step the
        // agent store or the event buffer may not hold (the host enlarges them and continues)
        if (life && ((long long)*lw.id_increment + 8ll * n_order + 64 > lw.capacity ||
                     (long long)*lw.n_events + 10ll * n_order + 64 > lw.ev_cap)) break;
        // ---- execution order, sugarscape.py:517 ----
        if (w.cfg.rng_mode == SS_RNG_MT19937) {
            if (tid == 0) shuffle_i32(sh, w.order, n_order);
        } else {
            unsigned long long skey = ss_step_key(w.cfg.keyed_seed, iteration);
            if (tid == 0) sh.skey = skey;
            OrderKey ok = ss_order_key(skey, w.half);
            int P = next_pow2(n_order);
            for (int i = tid; i < P; i += blockDim.x)
                w.sortkeys[i] = i < n_order
                    ? (((unsigned long long)ss_priority(ok, (uint32_t)w.order[i]) << 32) | (uint32_t)w.order[i])
                    : ~0ull;
            __syncthreads();
            cta_bitonic_sort(w.sortkeys, P);
            for (int i = tid; i < n_order; i += blockDim.x) w.order[i] = (int32_t)(w.sortkeys[i] & 0xffffffffu);
        }
        (void)dummy;
        if (tid == 0) { sh.sum_m = 0.0; sh.sum_m2 = 0.0; sh.sum_v = 0.0; sh.nw = 0; sh.kept = 0; sh.acted = 0; sh.deaths = 0; }
        __syncthreads();
        // ---- agent loop, sugarscape.py:520-601 (warp 0) ----
        if (warp == 0) {
            int skip_next = 0, kept = 0;
            for (int i = 0; i < n_order; i++) {
                const int s = w.order[i];
                if (skip_next) {                       // Q1: the agent after a dying agent loses its turn
                    skip_next = 0;
                    if (lane == 0) w.order[kept] = s;
                    kept++;
                    continue;
                }
                if (lane == 0 && w.cfg.rng_mode == SS_RNG_KEYED) {
                    sh.akey = ss_agent_key(sh.skey, (uint32_t)s + 1u);
                    sh.j = 0;
                }
                __syncwarp();
                if (w.cfg.rule_move_eat) {
                    if (w.cfg.rule_spice) warp_move_spice(w, sw, sh, s, lane, fl);
                    else warp_move(w, sh, s, env_time, lane, fl);
                }
                if (w.cfg.rule_utilicalc) warp_utilicalc(w, sh, s, lane);
                if (life && w.cfg.rule_procreate) {                                  // sugarscape.py:543-554
                    if (lane == 0 && life_fertile(w, lw, sw, s)) n_order = life_procreate(w, lw, sw, sh, s, n_order, stats_base + step, fl);
                    n_order = __shfl_sync(FULL, n_order, 0);
                }
                int died = 0;
                if (lane == 0) {
                    AgentRec* rec = &w.agents[s];
                    const int metab = attr_metab(rec->attr);
                    if (w.cfg.rule_pollution && rec->sugar > 0.0)                     // sugarscape.py:566-567
                        pollute_site(w, env_time, (int)rec->pos, 0.0, (double)metab);
                    rec->sugar = set_sugar_floor(ss_max0(rec->sugar - (double)metab), fl);   // sugarscape.py:569
                    sh.sum_m += (double)metab;
                    sh.sum_v += (double)attr_vision(rec->attr);
                    double wealth = rec->sugar;
                    if (w.cfg.rule_spice) {                                           // sugarscape.py:571-572, 576-583
                        const int m2 = sw.asp_metab[s];
                        sw.aspice[s] = set_sugar_floor(ss_max0(sw.aspice[s] - (double)m2), fl);
                        sh.sum_m2 += (double)m2;
                        wealth = rec->sugar + sw.aspice[s];
                    }
                    w.wealth[sh.nw++] = wealth;
                    sh.acted++;
                    bool living = life ? lw.rec[s].living != 0 : true;              // Agent.alive (0: reached maxAge last step)
                    if (w.cfg.rule_can_starve && rec->sugar <= 0.1) living = false;   // sugarscape.py:306-309
                    if (w.cfg.rule_can_starve && w.cfg.rule_spice && sw.aspice[s] <= 0.1) living = false;    // sugarscape.py:311-314
                    if (living && w.cfg.rule_limited_lifespan) {                      // sugarscape.py:588-591, agent.py:636-639
                        LifeRec& lr = lw.rec[s];
                        lr.age++;
                        lr.living = lr.max_age > lr.age ? 1 : 0;                      // (it stays for one more turn)
                    }
                    if (!living) {                                                    // sugarscape.py:593-601
                        rec->attr &= ~ATTR_ALIVE;
                        w.occw[rec->pos] = 0u;                                        // sugarscape.py:595
                        if (life) {
                            lw.rec[s].living = 0;
                            life_event(lw, stats_base + step, w.cfg.rule_limited_lifespan && lw.rec[s].age > lw.rec[s].max_age ? 3 : 2, s + 1, 0, 0);
                        }
                        sh.deaths++;
                        died = 1;
                    } else {
                        w.order[kept] = s;
                    }
                }
                died = __shfl_sync(FULL, died, 0);
                if (died) skip_next = 1; else kept++;
            }
            if (lane == 0) { sh.kept = kept; *w.n_order = kept; }
        }
        __syncthreads();
        // ---- statistics, sugarscape.py:603-619 + calculateGini 269-280 ----
        {
            const int nw = sh.nw, kept = sh.kept;
            int P = next_pow2(nw > 0 ? nw : 1);
            for (int i = nw + tid; i < P; i += blockDim.x) w.wealth[i] = __longlong_as_double(0x7ff0000000000000ll);
            __syncthreads();
            cta_bitonic_sort(w.wealth, P);
            if (tid == 0) {
                ss_step_stats st;
                st.population = (double)kept;
                st.metabolism_mean = kept > 0 ? (w.cfg.rule_spice ? ((sh.sum_m + sh.sum_m2) / 2.0) / (double)kept    // sugarscape.py:608-612
                                                                   : sh.sum_m / (double)kept) : 0.0;
                st.vision_mean = kept > 0 ? sh.sum_v / (double)kept : 0.0;
                double height = 0.0, area = 0.0;
                for (int i = 0; i < nw; i++) {
                    double v = w.wealth[i];
                    height += v;
                    area += height - v / 2.0;
                }
                double fair = height * (double)nw / 2.0;
                double g = fair == 0.0 ? 1.0 : (fair - area) / fair;
                st.gini = kept > 0 ? g : 0.0;
                st.acted = (double)sh.acted;
                st.deaths = (double)sh.deaths;
                w.stats[stats_base + step] = st;
            }
        }
        __syncthreads();
        if (fused_env) cta_environment(w, sw, iteration);
        __syncthreads();
    }
    for (int i = tid; i < 624; i += blockDim.x) w.mt[i] = sh.mt[i];
    if (tid == 0) *w.mti = sh.mti;
    if (life && tid == 0) lw.status[0] = step;
}

extern "C" cudaError_t ss_launch_serial(const DevWorld& w, const LifeWorld& lw, const SpiceWorld& sw, int64_t iteration, int64_t env_time,
                                        int nsteps, int fused_env, int stats_base, cudaStream_t stream) {
    ss_serial_kernel<<<1, SERIAL_THREADS, 0, stream>>>(w, lw, sw, iteration, env_time, nsteps, fused_env, stats_base);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// World initialisation of the reference's `__main__` (sugarscape.py:1485-1555), SURVEY 8 f4.
//
// ss_init_sites_kernel: Environment.addSiteHelper (environment.py:114-122) for every sugar site, then
// env.grow(maxCapacity) (environment.py:133-139) -- a cell's capacity is the largest radial term; pointwise.
struct InitSites {
    int n_sites;
    int sx[4], sy[4];
    double D[4];            // distance(max(si, W-si), max(sj, H-sj)) * (r / float(W)), computed on the host
    double max_capacity;
    int grow;
};
__global__ void ss_init_sites_kernel(DevWorld w, InitSites sp, double* cap_out, double* val_out, int clear_rest) {
    const int n = w.n;
    const size_t cells = (size_t)n * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < cells; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / n), j = (int)(idx % n);
        double cap = 0.0;
        for (int s = 0; s < sp.n_sites; s++) {
            const int di = sp.sx[s] - i, dj = sp.sy[s] - j;
            const double d = sqrt((double)((long long)di * di + (long long)dj * dj));
            double c = 1.0 + sp.max_capacity * (1.0 - d / sp.D[s]);
            c = c <= sp.max_capacity ? c : sp.max_capacity;                   // min(c, maxCapacity)
            if (c > cap) cap = c;
        }
        cap_out[idx] = cap;
        const double grown = 0.0 + sp.max_capacity;                           // environment.py:137: min(cur + alpha, cap)
        val_out[idx] = sp.grow ? (grown <= cap ? grown : cap) : 0.0;
        if (clear_rest) { w.poll[idx] = 0.0; w.occw[idx] = 0u; }
    }
}

// ss_init_agents_kernel: `Agent(env)` + `initAgent` per agent, one warp.  Lane 0 owns the MT19937 stream (every draw
// in the reference's order); the warp finds the k-th free cell of range(0, N-1) x range(0, N-1) in row-major order
// (environment.py:173-182 builds that list and indexes it) through per-row and per-32-row free counts.
struct InitAgents {
    int n_groups;
    int group_count[4];
    int max_metabolism, max_vision, endow_min, endow_max, age_min, age_max;
    int childbearing[2][4];
    int foresight_lo, foresight_hi;
    int group_tags[4];
};
// position of the k-th unit (0-based) in the prefix sums of 32 lane values; returns the lane and leaves k relative to it
__device__ static inline int warp_select(int v, int lane, long long& k, long long& total_out) {
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    total_out = __shfl_sync(FULL, incl, 31);
    if (k >= total_out) return -1;
    const unsigned b = __ballot_sync(FULL, incl > k);
    const int L = __ffs(b) - 1;
    const long long excl = __shfl_sync(FULL, incl - v, L);
    k -= excl;
    return L;
}
__global__ void __launch_bounds__(32)
ss_init_agents_kernel(DevWorld w, InitAgents ip, int32_t* rowfree, int32_t* superfree, int32_t* id_out, int32_t* x_out,
                      int32_t* y_out, int32_t* vision_out, int32_t* metab_out, double* sugar_out, int32_t* life_out, size_t life_stride,
                      int32_t* spice_out, int32_t* result) {
    __shared__ SerialShared sh;
    const int lane = threadIdx.x, n = w.n, n1 = n - 1, nsuper = (n1 + 31) / 32;
    for (int i = lane; i < 624; i += 32) sh.mt[i] = w.mt[i];
    if (lane == 0) { sh.mti = *w.mti; sh.rng_mode = SS_RNG_MT19937; }
    for (int i = lane; i < n1; i += 32) rowfree[i] = n1;
    for (int s = lane; s < nsuper; s += 32) superfree[s] = n1 * min(32, n1 - s * 32);
    __syncwarp();
    long long F = (long long)n1 * n1;
    int nid = 0, count = 0;
    for (int g = 0; g < ip.n_groups; g++) {
        for (int a = 0; a < ip.group_count[g]; a++) {
            nid++;
            long long k = 0;
            if (lane == 0) {
                (void)randbelow(sh, (uint32_t)(ip.foresight_hi - ip.foresight_lo + 1));        // agent.py:165 generateForesight
                if (F > 0) k = (long long)randbelow(sh, (uint32_t)F);                         // environment.py:181
            }
            if (F == 0) continue;                                                             // initAgent returns False
            k = __shfl_sync(FULL, k, 0);
            int sr = -1;
            for (int base = 0; base < nsuper && sr < 0; base += 32) {
                long long tot;
                const int L = warp_select(base + lane < nsuper ? superfree[base + lane] : 0, lane, k, tot);
                if (L >= 0) sr = base + L; else k -= tot;
            }
            long long tot;
            const int rl = warp_select(sr * 32 + lane < n1 ? rowfree[sr * 32 + lane] : 0, lane, k, tot);
            const int row = sr * 32 + rl;
            int col = -1;
            for (int base = 0; base < n1 && col < 0; base += 32) {
                const int c = base + lane;
                const unsigned b = __ballot_sync(FULL, c < n1 && w.occw[(size_t)row * n + c] == 0u);
                const int cnt = __popc(b);
                if (k < cnt) col = base + (int)__fns(b, 0u, (int)k + 1); else k -= cnt;
            }
            if (lane == 0) {
                w.occw[(size_t)row * n + col] = 1u;
                rowfree[row]--; superfree[sr]--;
                const int metab = 1 + (int)randbelow(sh, (uint32_t)ip.max_metabolism);            // sugarscape.py:250
                const int endow = ip.endow_min + (int)randbelow(sh, (uint32_t)(ip.endow_max - ip.endow_min + 1));
                if (spice_out) {                                                                  // :253-255 (rules.spice)
                    spice_out[count] = 1 + (int)randbelow(sh, (uint32_t)ip.max_metabolism);
                    spice_out[life_stride + count] = ip.endow_min + (int)randbelow(sh, (uint32_t)(ip.endow_max - ip.endow_min + 1));
                }
                const int vision = 1 + (int)randbelow(sh, (uint32_t)ip.max_vision);               // :256
                const int max_age = ip.age_min + (int)randbelow(sh, (uint32_t)(ip.age_max - ip.age_min + 1));     // :257 setAge
                const int sex = (int)randbelow(sh, 2u);                                           // :258
                const int f_lo = ip.childbearing[sex][0] + (int)randbelow(sh, (uint32_t)(ip.childbearing[sex][1] - ip.childbearing[sex][0] + 1));   // :260
                const int f_hi = ip.childbearing[sex][2] + (int)randbelow(sh, (uint32_t)(ip.childbearing[sex][3] - ip.childbearing[sex][2] + 1));   // :261
                id_out[count] = nid; x_out[count] = row; y_out[count] = col;
                vision_out[count] = vision; metab_out[count] = metab; sugar_out[count] = (double)endow;
                if (life_out) {
                    life_out[count] = max_age; life_out[life_stride + count] = sex; life_out[2 * life_stride + count] = f_lo;
                    life_out[3 * life_stride + count] = f_hi; life_out[4 * life_stride + count] = ip.group_tags[g];
                }
            }
            F--; count++;
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = lane; i < 624; i += 32) w.mt[i] = sh.mt[i];
    if (lane == 0) { *w.mti = sh.mti; result[0] = count; result[1] = nid; }
}

extern "C" cudaError_t ss_launch_init_sites(const DevWorld& w, int n_sites, const int* sx, const int* sy, const double* D,
                                            double max_capacity, int grow, int sm_count, double* cap_out, double* val_out,
                                            cudaStream_t stream) {
    InitSites sp;
    sp.n_sites = n_sites;
    for (int s = 0; s < 4; s++) { sp.sx[s] = s < n_sites ? sx[s] : 0; sp.sy[s] = s < n_sites ? sy[s] : 0; sp.D[s] = s < n_sites ? D[s] : 1.0; }
    sp.max_capacity = max_capacity; sp.grow = grow;
    // (cap_out / val_out: the sugar arrays, or the spice arrays of SURVEY 8 f2)
    ss_init_sites_kernel<<<sm_count * 8, 256, 0, stream>>>(w, sp, cap_out ? cap_out : w.cap, val_out ? val_out : w.sugar, cap_out == nullptr);
    return cudaGetLastError();
}

extern "C" cudaError_t ss_launch_init_agents(const DevWorld& w, const ss_init_params* p, int32_t* rowfree, int32_t* superfree,
                                             int32_t* cols, size_t stride, double* sugar, int32_t* life_cols, int32_t* spice_cols,
                                             int32_t* result, cudaStream_t stream) {
    InitAgents ip;
    ip.n_groups = p->n_groups;
    for (int g = 0; g < 4; g++) ip.group_count[g] = g < p->n_groups ? p->group_count[g] : 0;
    ip.max_metabolism = p->max_metabolism; ip.max_vision = p->max_vision;
    ip.endow_min = p->endowment_min; ip.endow_max = p->endowment_max; ip.age_min = p->age_min; ip.age_max = p->age_max;
    for (int s = 0; s < 2; s++) for (int k = 0; k < 4; k++) ip.childbearing[s][k] = p->childbearing[s][k];
    ip.foresight_lo = p->foresight_lo; ip.foresight_hi = p->foresight_hi;
    for (int g = 0; g < 4; g++) ip.group_tags[g] = p->group_tags[g];
    ss_init_agents_kernel<<<1, 32, 0, stream>>>(w, ip, rowfree, superfree, cols, cols + stride, cols + 2 * stride, cols + 3 * stride,
                                                 cols + 4 * stride, sugar, life_cols, stride, spice_cols, result);
    return cudaGetLastError();
}
