/* ss_oracle.c -- plain-C restatement of the reference's per-timestep hot path.
 *
 * TEST INFRASTRUCTURE: see ss_oracle.h.  Single-threaded and strictly sequential,
 * exactly like the reference's Python loop.  Compile with -ffp-contract=off so no
 * FMA is formed: every float operation below is one IEEE-754 double operation in
 * the order the reference performs it.
 *
 * Reference citations are into /root/reference:
 *   View.updateGame            sugarscape.py:506-686
 *   hasStarved                 sugarscape.py:306-314
 *   calculateGini              sugarscape.py:269-280
 *   Agent.getFood              agent.py:361-371
 *   Agent.getVisionNeighbourhood agent.py:391-407
 *   Agent.move                 agent.py:794-865
 *   Agent.utilicalcSugar       agent.py:980-992
 *   Agent.utilicalcMove        agent.py:1025-1141
 *   Agent.setSugar             agent.py:224-229
 *   Agent.incAge / mate / findFreeLocationAround / createChild / isFertile / getNeighbourhood
 *                              agent.py:636-639, 653-679, 681-696, 698-750, 764-773, 373-389   
 *   Environment.grow/growRegion environment.py:133-155
 *   Environment.polluteSite    environment.py:312-315
 *   Environment.spreadPollution environment.py:317-350
 * RNG: CPython 3.12 random.py:242 (_randbelow_with_getrandbits), :350 (shuffle),
 *   :341 (choice) over _randommodule.c's MT19937 (getrandbits(k<=32) =
 *   genrand_uint32() >> (32-k)); or the keyed word source of oracle/keyed_rng.py.
 */
#include "ss_oracle.h"
#include "../sugarscape_utilicalc_b200/csrc/ss_pow.h"   /* x ** y of the reference, bit for bit */

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXC 132 /* 4*vision+1 with vision <= 32 */

static char g_err[256];
const char* sso_last_error(void) { return g_err; }
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return -1; }

struct sso_world {
    sso_config cfg;
    int n;
    double *sugar, *cap, *poll;
    double *spice, *spice_cap;    /* Environment.grid slots [1], [3] (environment.py:24-34) */
    int32_t* occ;             /* slot or -1 */
    /* agent store, indexed by slot = id - 1 */
    int32_t n_slots;
    int32_t min_slots;
    int32_t *ax, *ay, *avision, *ametab;
    double* asugar;
    uint8_t* alive;           /* in View.agents (and on the lattice) */
    /* SURVEY 8 f3 (limitedLifespan / procreate): Agent.age, maxAge, sex, fertility, sugarEndowment (= sugarCostForMating,
     * agent.py:204-206 with the (0, 0) cost range), tags, Agent.alive (0 for an agent that reached maxAge: it stays in the
     * list and on the lattice for one more turn, sugarscape.py:588-601) */
    int32_t cap_slots;        /* allocation of the per-slot arrays and of order / wealth */
    int32_t id_increment;     /* Environment.idIncrement: ids handed out so far */
    int have_life;
    int32_t *age, *max_age, *sex, *fert_lo, *fert_hi, *tags;
    double* endow;
    uint8_t* living;
    /* Agent.spice, spiceMetabolism, spiceEndowment (= spiceCostForMating) */
    int have_spice;
    double *aspice, *sp_endow;
    int32_t* asp_metab;
    int32_t* events;          /* 4 per event: kind | step << 8, a, b, c -- kind 1: a mated with b and created c; 2: a died of
                                 starvation; 3: a died of old age (sugarscape.py:551, 597-600) */
    int32_t n_events, cap_events, step_in_call;
    /* persistent list order (slots) */
    int32_t* order;
    int32_t n_order;
    int64_t iteration, env_time;
    /* MT19937 */
    uint32_t mt[624];
    int mti;
    /* keyed stream of the acting agent */
    uint64_t skey, akey;
    uint32_t j;
    uint32_t rk[4];
    int half;
    /* scratch */
    double* wealth;
    uint64_t *sortk, *sortk2;
};

/* ---------------------------------------------------------------- MT19937 ---- */
static uint32_t mt_next(sso_world* w) {
    uint32_t* mt = w->mt;
    if (w->mti >= 624) {
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
        w->mti = 0;
    }
    uint32_t y = mt[w->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* ------------------------------------------------------------ keyed source ---- */
#define GOLDEN 0x9E3779B97F4A7C15ull
#define C_AGENT 0xD1342543DE82EF95ull
#define C_ROUND 0xA24BAED4963EE407ull

static uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
static int order_half_bits(int32_t n_slots) {
    int kb = 0;
    uint32_t v = (uint32_t)((n_slots > 1 ? n_slots : 1) - 1);
    while (v) { kb++; v >>= 1; }
    if (kb < 1) kb = 1;
    int half = (kb + 1) / 2;
    return half < 1 ? 1 : half;
}
static uint32_t priority_of(const sso_world* w, uint32_t slot) {
    uint32_t mask = (1u << w->half) - 1u;
    uint32_t L = (slot >> w->half) & mask, R = slot & mask;
    for (int r = 0; r < 4; r++) {
        uint32_t t = L ^ (fmix32(R ^ w->rk[r]) & mask);
        L = R; R = t;
    }
    return (L << w->half) | R;
}
static void keyed_begin_step(sso_world* w, int64_t t) {
    w->skey = mix64(w->cfg.keyed_seed + GOLDEN * (uint64_t)(t + 1));
    for (int r = 0; r < 4; r++) w->rk[r] = (uint32_t)(mix64(w->skey + (uint64_t)(r + 1) * C_ROUND) >> 32);
}
static void keyed_begin_agent(sso_world* w, int32_t id) {
    w->akey = mix64(w->skey ^ ((uint64_t)id * C_AGENT));
    w->j = 0;
}

/* ------------------------------------------------- CPython random algorithms ---- */
static uint32_t next_word(sso_world* w) {
    if (w->cfg.rng_mode == 0) return mt_next(w);
    uint32_t r = (uint32_t)(mix64(w->akey + (uint64_t)w->j * GOLDEN) >> 32);
    w->j++;
    return r;
}
/* random.py:242 _randbelow_with_getrandbits, n >= 1, n < 2^31 */
static uint32_t randbelow(sso_world* w, uint32_t n) {
    int k = 0;
    for (uint32_t v = n; v; v >>= 1) k++;
    uint32_t r = next_word(w) >> (32 - k);
    while (r >= n) r = next_word(w) >> (32 - k);
    return r;
}
/* random.py:350 shuffle */
static void shuffle_i32(sso_world* w, int32_t* a, int n) {
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = randbelow(w, (uint32_t)i + 1u);
        int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}

/* ------------------------------------------------------------------ helpers ---- */
static inline int wrap(int v, int n) { int r = v % n; return r < 0 ? r + n : r; }
/* Python float // int (floatobject.c float_floor_div) */
static double py_floordiv(double vx, double wx) {
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
/* agent.py:224-229 setSugar: max(amount, 0 if limitedLifespan else .01) */
static inline double set_sugar(const sso_world* w, double amount) {
    const double fl = w->cfg.rule_limited_lifespan ? 0.0 : 0.01;
    return amount > fl ? amount : fl;
}
static inline double max0(double v) { return v > 0.0 ? v : 0.0; }

/* environment.py:312-315 */
static void pollute_site(sso_world* w, int cell, double production, double consumption) {
    if (w->cfg.pollution_start_time <= w->env_time)
        w->poll[cell] += (production * w->cfg.pollution_a) + (consumption * w->cfg.pollution_b);
}

/* agent.py:361-371 getFood: free cells on the x-arm then the y-arm (torus), shuffled.
 * Cells are returned as linear indices x*N + y. */
static int get_food(sso_world* w, int s, int32_t* out) {
    int n = w->n, x = w->ax[s], y = w->ay[s], v = w->avision[s], m = 0;
    for (int xx = x - v; xx <= x + v; xx++) {
        int c = wrap(xx, n) * n + y;
        if (w->occ[c] < 0) out[m++] = c;
    }
    for (int yy = y - v; yy <= y + v; yy++) {
        int c = x * n + wrap(yy, n);
        if (w->occ[c] < 0) out[m++] = c;
    }
    shuffle_i32(w, out, m);
    return m;
}

/* agent.py:391-407 getVisionNeighbourhood: occupants on the arms, shuffled (slots) */
static int get_vision_neighbourhood(sso_world* w, int s, int32_t* out) {
    int n = w->n, x = w->ax[s], y = w->ay[s], v = w->avision[s], m = 0;
    for (int xx = x - v; xx <= x + v; xx++) {
        int wx = wrap(xx, n);
        int c = wx * n + y;
        if (w->occ[c] >= 0 && wx != x) out[m++] = w->occ[c];
    }
    for (int yy = y - v; yy <= y + v; yy++) {
        int wy = wrap(yy, n);
        int c = x * n + wy;
        if (w->occ[c] >= 0 && wy != y) out[m++] = w->occ[c];
    }
    shuffle_i32(w, out, m);
    return m;
}

/* agent.py:794-865 move(), no-spice branch */
static void agent_move(sso_world* w, int s) {
    int n = w->n, x = w->ax[s], y = w->ay[s];
    int32_t food[MAXC];
    int m = get_food(w, s, food);
    food[m++] = x * n + y;                       /* agent.py:806 */
    int best = -1, best_d = 0, best_sg = 0;
    double best_key = 0.0;
    for (int p = 0; p < m; p++) {
        int c = food[p];
        int cx = c / n, cy = c % n;
        int sg = (int)w->sugar[c];                /* environment.py:98-100 int() */
        int d = abs(x - cx) + abs(y - cy);        /* agent.py:867-868, raw coordinates */
        double key = w->cfg.rule_pollution ? (double)sg / (1.0 + w->poll[c]) : (double)sg;
        /* stable sort by distance then first max  ==  (key desc, d asc, p asc) */
        if (best < 0 || key > best_key || (key == best_key && d < best_d)) {
            best = c; best_key = key; best_d = d; best_sg = sg;
        }
    }
    if (w->cfg.rule_pollution) pollute_site(w, best, (double)best_sg, 0.0);   /* agent.py:825-826 */
    w->asugar[s] = set_sugar(w, max0(w->asugar[s] + (double)best_sg));            /* agent.py:832 */
    int old = x * n + y;
    if (best != old) {                                                          /* agent.py:856-862 */
        w->occ[old] = -1;
        w->occ[best] = s;
        w->ax[s] = best / n;
        w->ay[s] = best % n;
    }
    w->sugar[best] = 0.0;                                                       /* agent.py:864 */
}

/* agent.py:231-236 setSpice: same floor as setSugar */
#define set_spice set_sugar
/* float_pow (CPython floatobject.c) for positive finite bases: iw == 0 and iv == 1 are answered before libm's pow() */
static const ss_pow_tables POW_T = {SS_POW_LOG_HDR, SS_POW_LOG_TAB, SS_POW_EXP_HDR, SS_POW_EXP_TAB};
static int g_pow_bad;
static double py_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    return ss_pow_pos(&POW_T, x, y, &g_pow_bad);
}
/* agent.py:1220-1251 getWelfare(x=, y=), no tags, no foresight: Cobb-Douglas in the wealth the agent would have on the cell --
 * `if x and y:` (Q9): a cell on row 0 or column 0 does not count its sugar and spice */
static double welfare_at(sso_world* w, int s, int cx, int cy) {
    double w1 = w->asugar[s], w2 = w->aspice[s];
    if (cx != 0 && cy != 0) {
        int c = cx * w->n + cy;
        w1 = w1 + (double)(int)w->sugar[c];
        w2 = w2 + (double)(int)w->spice[c];
    }
    double m1 = (double)w->ametab[s], m2 = (double)w->asp_metab[s], mt = (double)(w->ametab[s] + w->asp_metab[s]);
    if (!(w1 > 0.0 && w2 > 0.0)) return 0.0;
    return py_pow(w1, m1 / mt) * py_pow(w2, m2 / mt);
}
/* agent.py:794-865 move(), spice branch (833-853): closest cell of highest welfare, both resources harvested */
static void agent_move_spice(sso_world* w, int s) {
    int n = w->n, x = w->ax[s], y = w->ay[s];
    int32_t food[MAXC];
    int m = get_food(w, s, food);
    food[m++] = x * n + y;
    int best = -1, best_d = 0;
    double best_w = 0.0;
    for (int p = 0; p < m; p++) {
        int c = food[p], cx = c / n, cy = c % n;
        double wf = welfare_at(w, s, cx, cy);
        int d = abs(x - cx) + abs(y - cy);
        if (best < 0 || wf > best_w || (wf == best_w && d < best_d)) { best = c; best_w = wf; best_d = d; }
    }
    int sg = (int)w->sugar[best], sp = (int)w->spice[best];
    w->aspice[s] = set_spice(w, max0(w->aspice[s] + (double)sp));                /* agent.py:851 */
    w->asugar[s] = set_sugar(w, max0(w->asugar[s] + (double)sg));                /* agent.py:852 */
    int old = x * n + y;
    if (best != old) {
        w->occ[old] = -1; w->occ[best] = s;
        w->ax[s] = best / n; w->ay[s] = best % n;
    }
    w->sugar[best] = 0.0;                                                        /* agent.py:864-865 */
    w->spice[best] = 0.0;
}

/* agent.py:1025-1141 utilicalcMove(), sugar-only variant (utilicalcSugar 980-992) */
static void agent_utilicalc(sso_world* w, int s) {
    int n = w->n, x = w->ax[s], y = w->ay[s];
    int32_t moves[MAXC], neigh[MAXC];
    double util[MAXC];
    int m = get_food(w, s, moves);               /* shuffle #1 */
    shuffle_i32(w, moves, m);                    /* shuffle #2, agent.py:1033 */
    moves[m++] = x * n + y;
    int nb = get_vision_neighbourhood(w, s, neigh); /* shuffle #3 */
    neigh[nb++] = s;
    double extent = (double)(nb - 1);
    double umax = 0.0;
    for (int p = 0; p < m; p++) {
        int c = moves[p];
        int mx = c / n, my = c % n;
        double site = (double)(int)w->sugar[c];
        double u = 0.0;
        for (int q = 0; q < nb; q++) {
            int b = neigh[q];
            double bs = w->asugar[b], bm = (double)w->ametab[b];
            double days = py_floordiv(bs, bm);                     /* agent.py:870-872 */
            double inten = 1.0 / (1.0 + days);
            double dur = (site / bm) / w->cfg.max_capacity;
            int bx = w->ax[b], by = w->ay[b], bv = w->avision[b];
            int cert = (bx == mx && abs(by - my) <= bv) || (by == my && abs(bx - mx) <= bv);
            double t = 0.0;
            t += (double)cert * (((inten + dur) + 0.5 * 0.0) + extent);
            if (b != s) t *= -1.0;
            u += t;
        }
        util[p] = u;
        if (p == 0 || u > umax) umax = u;
    }
    int32_t maxloc[MAXC];
    int nmax = 0;
    for (int p = 0; p < m; p++)
        if (util[p] == umax) maxloc[nmax++] = moves[p];
    int dest = maxloc[randbelow(w, (uint32_t)nmax)];                /* random.choice */
    w->occ[x * n + y] = -1;
    w->occ[dest] = s;
    w->ax[s] = dest / n;
    w->ay[s] = dest % n;
    int sg = (int)w->sugar[dest];
    w->asugar[s] = set_sugar(w, max0(w->asugar[s] + (double)sg));
    w->sugar[dest] = 0.0;
}

/* ------------------------------------------------------------ lifespan + procreation ---- */
static int ensure_capacity(sso_world* w, int32_t need) {
    if (need <= w->cap_slots) return 0;
    int32_t nc = w->cap_slots > 0 ? w->cap_slots : 1;
    while (nc < need) nc *= 2;
#define GROW(p, T) do { T* q = (T*)realloc(p, (size_t)nc * sizeof(T)); if (!q) return fail("out of memory"); \
                        memset(q + w->cap_slots, 0, (size_t)(nc - w->cap_slots) * sizeof(T)); p = q; } while (0)
    GROW(w->ax, int32_t); GROW(w->ay, int32_t); GROW(w->avision, int32_t); GROW(w->ametab, int32_t);
    GROW(w->asugar, double); GROW(w->alive, uint8_t); GROW(w->order, int32_t); GROW(w->wealth, double);
    GROW(w->sortk, uint64_t); GROW(w->sortk2, uint64_t);
    GROW(w->age, int32_t); GROW(w->max_age, int32_t); GROW(w->sex, int32_t); GROW(w->fert_lo, int32_t);
    GROW(w->fert_hi, int32_t); GROW(w->tags, int32_t); GROW(w->endow, double); GROW(w->living, uint8_t);
    GROW(w->aspice, double); GROW(w->sp_endow, double); GROW(w->asp_metab, int32_t);
#undef GROW
    w->cap_slots = nc;
    return 0;
}
static void add_event(sso_world* w, int kind, int a, int b, int c) {
    if (w->n_events == w->cap_events) {
        int32_t nc = w->cap_events ? 2 * w->cap_events : 1024;
        int32_t* q = (int32_t*)realloc(w->events, (size_t)nc * 4 * sizeof(int32_t));
        if (!q) return;
        w->events = q; w->cap_events = nc;
    }
    int32_t* e = w->events + 4 * (size_t)w->n_events++;
    e[0] = kind | (w->step_in_call << 8); e[1] = a; e[2] = b; e[3] = c;
}
/* agent.py:764-773 isFertile, no-spice branch (sugarCostForMating = the endowment, agent.py:204-206) */
static int is_fertile(const sso_world* w, int s) {
    return w->fert_lo[s] <= w->age[s] && w->age[s] <= w->fert_hi[s] && w->asugar[s] >= w->endow[s] &&
           (!w->cfg.rule_spice || w->aspice[s] >= w->sp_endow[s]);
}
/* agent.py:681-696 findFreeLocationAround: no torus here (isLocationValid), x-line then y-line, randint over the list */
static int find_free_around(sso_world* w, int x, int y) {
    int n = w->n, loc[6], m = 0;
    for (int i = x - 1; i <= x + 1; i++)
        if (i >= 0 && i < n && w->occ[i * n + y] < 0) loc[m++] = i * n + y;
    for (int j = y - 1; j <= y + 1; j++)
        if (j >= 0 && j < n && w->occ[x * n + j] < 0) loc[m++] = x * n + j;
    if (!m) return -1;
    return loc[randbelow(w, (uint32_t)m)];
}
/* random.py random(): (a * 2^26 + b) / 2^53 with a = genrand >> 5, b = genrand >> 6 (_randommodule.c) */
static double random_random(sso_world* w) {
    uint32_t a = next_word(w) >> 5, b = next_word(w) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}
/* agent.py:698-750 createChild(parent = b) of agent s at cell loc; the child joins the end of View.agents
 * (sugarscape.py:550) and is met by the running loop in this very step */
static int create_child(sso_world* w, int s, int b, int loc) {
    int g[2] = {s, b};
    int metab = w->ametab[g[randbelow(w, 2u)]];
    int sp_metab = w->asp_metab[g[randbelow(w, 2u)]];         /* spiceMetabolism (drawn with or without the spice rule) */
    int vision = w->avision[g[randbelow(w, 2u)]];
    double endow = 0.5 * (w->endow[s] + w->endow[b]);
    double sp_endow = 0.5 * (w->sp_endow[s] + w->sp_endow[b]);
    w->asugar[s] = set_sugar(w, max0(w->asugar[s] - 0.5 * w->endow[s]));
    w->asugar[b] = set_sugar(w, max0(w->asugar[b] - 0.5 * w->endow[b]));
    if (w->cfg.rule_spice) {                                  /* agent.py:712-714 */
        w->aspice[s] = set_spice(w, max0(w->aspice[s] - 0.5 * w->sp_endow[s]));
        w->aspice[b] = set_spice(w, max0(w->aspice[b] - 0.5 * w->sp_endow[b]));
    }
    int max_age = w->max_age[g[randbelow(w, 2u)]];
    int sex = (int)randbelow(w, 2u);
    int fg = g[randbelow(w, 2u)];
    int32_t mask = 1, tags = 0;
    for (int t = 0; t < w->cfg.tags_length; t++) {            /* agent.py:727-737 */
        int32_t t1 = w->tags[s] & mask, t2 = w->tags[b] & mask;
        if (t1 == t2 || random_random(w) < 0.5) tags |= t1; else tags |= t2;
        mask <<= 1;
    }
    int c = w->id_increment;                                  /* slot = id - 1; Environment.getNewId */
    if (ensure_capacity(w, c + 1)) return -1;
    w->id_increment = c + 1;
    if (c + 1 > w->n_slots) w->n_slots = c + 1;
    w->ax[c] = loc / w->n; w->ay[c] = loc % w->n;
    w->ametab[c] = metab; w->avision[c] = vision;
    w->endow[c] = endow;
    w->asugar[c] = set_sugar(w, endow);                       /* setInitialSugarEndowment */
    w->asp_metab[c] = sp_metab; w->sp_endow[c] = sp_endow;
    w->aspice[c] = w->cfg.rule_spice ? set_spice(w, sp_endow) : 0.0;
    w->age[c] = 0; w->max_age[c] = max_age; w->sex[c] = sex;
    w->fert_lo[c] = w->fert_lo[fg]; w->fert_hi[c] = w->fert_hi[fg];
    w->tags[c] = tags;
    w->alive[c] = 1; w->living[c] = 1;
    (void)randbelow(w, (uint32_t)(w->cfg.foresight_hi - w->cfg.foresight_lo + 1));   /* generateForesight, agent.py:165 */
    w->occ[loc] = c;
    w->order[w->n_order++] = c;
    add_event(w, 1, s + 1, b + 1, c + 1);
    return 0;
}
/* sugarscape.py:543-554 + agent.py:653-679 mate(): every fertile neighbour of the other sex, in shuffled order */
static void procreate(sso_world* w, int s) {
    int n = w->n, x = w->ax[s], y = w->ay[s];
    int32_t nb[4];
    int m = 0;
    for (int xx = x - 1; xx <= x + 1; xx++) {                 /* agent.py:373-389 getNeighbourhood */
        if (xx == x) continue;
        int c = wrap(xx, n) * n + y;
        if (w->occ[c] >= 0) nb[m++] = w->occ[c];
    }
    for (int yy = y - 1; yy <= y + 1; yy++) {
        if (yy == y) continue;
        int c = x * n + wrap(yy, n);
        if (w->occ[c] >= 0) nb[m++] = w->occ[c];
    }
    shuffle_i32(w, nb, m);
    for (int k = 0; k < m; k++) {
        int b = nb[k];
        if (w->sex[b] == w->sex[s] || !is_fertile(w, b)) continue;
        int loc = find_free_around(w, w->ax[s], w->ay[s]);
        if (loc < 0) loc = find_free_around(w, w->ax[b], w->ay[b]);
        if (loc >= 0 && is_fertile(w, s)) create_child(w, s, b, loc);
    }
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}
/* sugarscape.py:269-280 */
static double calculate_gini(double* wealth, int n) {
    qsort(wealth, (size_t)n, sizeof(double), cmp_double);
    double height = 0.0, area = 0.0;
    for (int i = 0; i < n; i++) {
        height += wealth[i];
        area += height - wealth[i] / 2.0;
    }
    double fair_area = height * (double)n / 2.0;
    if (fair_area == 0.0) return 1.0;
    return (fair_area - area) / fair_area;
}

/* environment.py:133-139 */
static inline void grow_cell(sso_world* w, int c, double alpha) {
    double v = w->sugar[c] + alpha;
    w->sugar[c] = v <= w->cap[c] ? v : w->cap[c];
    if (w->cfg.rule_spice) {                       /* environment.py:141-142 */
        double u = w->spice[c] + alpha;
        w->spice[c] = u <= w->spice_cap[c] ? u : w->spice_cap[c];
    }
}
/* environment.py:146-155 */
static void grow_region(sso_world* w, const int32_t r[4], double alpha) {
    int n = w->n;
    int imin = r[0] > 0 ? r[0] : 0, jmin = r[1] > 0 ? r[1] : 0;
    int imax = r[2] + 1 < n ? r[2] + 1 : n, jmax = r[3] + 1 < n ? r[3] + 1 : n;
    for (int j = jmin; j < jmax; j++)
        for (int i = imin; i < imax; i++) grow_cell(w, i * n + j, alpha);
}
/* environment.py:317-350: in-place sweep, neighbours appended n,s,e,w */
static void spread_pollution(sso_world* w) {
    int n = w->n;
    double* p = w->poll;
    for (int x = 0; x < n; x++)
        for (int y = 0; y < n; y++) {
            double t = 0.0;
            int c = 0;
            if (y + 1 < n) { t += p[x * n + y + 1]; c++; }
            if (y - 1 >= 0) { t += p[x * n + y - 1]; c++; }
            if (x + 1 < n) { t += p[(x + 1) * n + y]; c++; }
            if (x - 1 >= 0) { t += p[(x - 1) * n + y]; c++; }
            p[x * n + y] = t / (double)c;
        }
}

static void radix_sort_u64(uint64_t* a, uint64_t* tmp, int n) {
    /* sort by the high 32 bits (priority), 4 passes of 8 bits */
    for (int pass = 0; pass < 4; pass++) {
        int shift = 32 + pass * 8;
        size_t cnt[257] = {0};
        for (int i = 0; i < n; i++) cnt[((a[i] >> shift) & 0xff) + 1]++;
        for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
        for (int i = 0; i < n; i++) tmp[cnt[(a[i] >> shift) & 0xff]++] = a[i];
        uint64_t* t = a; a = tmp; tmp = t;
    }
    /* 4 passes -> result is back in the original buffer */
}

/* sugarscape.py:506-686 */
static void one_step(sso_world* w, sso_step_stats* st) {
    const sso_config* cfg = &w->cfg;
    int n = w->n;
    /* sugarscape.py:517 execution order */
    if (cfg->rng_mode == 0) {
        shuffle_i32(w, w->order, w->n_order);
    } else {
        keyed_begin_step(w, w->iteration);
        for (int i = 0; i < w->n_order; i++)
            w->sortk[i] = ((uint64_t)priority_of(w, (uint32_t)w->order[i]) << 32) | (uint32_t)w->order[i];
        radix_sort_u64(w->sortk, w->sortk2, w->n_order);
        for (int i = 0; i < w->n_order; i++) w->order[i] = (int32_t)(w->sortk[i] & 0xffffffffu);
    }
    double sum_m = 0.0, sum_m2 = 0.0, sum_v = 0.0;
    int nw = 0, kept = 0, acted = 0, deaths = 0, skip_next = 0;
    /* sugarscape.py:520: `for agent in self.agents` while removeAgent() does list.remove:
       the agent following each dying agent loses its turn (Q1). */
    for (int i = 0; i < w->n_order; i++) {
        int s = w->order[i];
        if (skip_next) { skip_next = 0; w->order[kept++] = s; continue; }
        if (cfg->rng_mode == 1) keyed_begin_agent(w, s + 1);
        if (cfg->rule_move_eat) { if (cfg->rule_spice) agent_move_spice(w, s); else agent_move(w, s); }
        if (cfg->rule_utilicalc) agent_utilicalc(w, s);
        if (cfg->rule_procreate && is_fertile(w, s)) procreate(w, s);        /* sugarscape.py:543-554 */
        /* sugarscape.py:566-567 consumption pollution */
        if (cfg->rule_pollution && w->asugar[s] > 0.0)
            pollute_site(w, w->ax[s] * n + w->ay[s], 0.0, (double)w->ametab[s]);
        /* sugarscape.py:569 */
        w->asugar[s] = set_sugar(w, max0(w->asugar[s] - (double)w->ametab[s]));
        if (cfg->rule_spice)                                    /* sugarscape.py:571-572 */
            w->aspice[s] = set_spice(w, max0(w->aspice[s] - (double)w->asp_metab[s]));
        sum_m += (double)w->ametab[s];
        if (cfg->rule_spice) sum_m2 += (double)w->asp_metab[s];
        sum_v += (double)w->avision[s];
        w->wealth[nw++] = cfg->rule_spice ? w->asugar[s] + w->aspice[s] : w->asugar[s];      /* sugarscape.py:580-583 */
        acted++;
        int living = w->have_life ? w->living[s] : 1;          /* Agent.alive: 0 for an agent that reached maxAge last step */
        if (cfg->rule_can_starve && w->asugar[s] <= 0.1) living = 0;    /* sugarscape.py:306-309 */
        if (cfg->rule_can_starve && cfg->rule_spice && w->aspice[s] <= 0.1) living = 0;      /* sugarscape.py:311-314 */
        if (living && cfg->rule_limited_lifespan) {             /* sugarscape.py:588-591 + agent.py:636-639 incAge */
            w->age[s]++;
            w->living[s] = (uint8_t)(w->max_age[s] > w->age[s]);        /* (the agent stays for one more turn) */
        }
        if (!living) {                                          /* sugarscape.py:593-601 */
            w->alive[s] = 0;
            if (w->have_life) {
                w->living[s] = 0;
                add_event(w, cfg->rule_limited_lifespan && w->age[s] > w->max_age[s] ? 3 : 2, s + 1, 0, 0);
            }
            w->occ[w->ax[s] * n + w->ay[s]] = -1;
            deaths++;
            skip_next = 1;
        } else {
            w->order[kept++] = s;
        }
    }
    w->n_order = kept;
    st->population = (double)kept;
    st->metabolism_mean = kept > 0 ? (cfg->rule_spice ? ((sum_m + sum_m2) / 2.0) / (double)kept : sum_m / (double)kept) : 0.0;   /* sugarscape.py:608-612 */
    st->vision_mean = kept > 0 ? sum_v / (double)kept : 0.0;
    st->gini = kept > 0 ? calculate_gini(w->wealth, nw) : 0.0;
    st->acted = (double)acted;
    st->deaths = (double)deaths;
    /* sugarscape.py:642-659 */
    if (cfg->rule_seasons) {
        double S = (double)(w->iteration % (2 * (int64_t)cfg->season_period)) / (double)cfg->season_period;
        if (cfg->rule_grow) {
            if (S < 1.0) {
                grow_region(w, cfg->season_north, cfg->season_on_factor);
                grow_region(w, cfg->season_south, cfg->season_off_factor);
            } else {
                grow_region(w, cfg->season_north, cfg->season_off_factor);
                grow_region(w, cfg->season_south, cfg->season_on_factor);
            }
        }
    } else if (cfg->rule_grow) {
        for (int c = 0; c < n * n; c++) grow_cell(w, c, cfg->grow_factor);
    }
    /* sugarscape.py:661-663 */
    if (cfg->rule_pollution && w->iteration >= cfg->diffusion_start_time &&
        w->iteration % cfg->diffusion_rate == 0)
        spread_pollution(w);
    w->iteration++;
    w->env_time++;
}

/* ---------------------------------------------------------------- public API ---- */
int sso_create(const sso_config* cfg, sso_world** out) {
    if (!cfg || !out) return fail("null argument");
    if (cfg->grid_size < 1) return fail("grid_size must be >= 1");
    if (cfg->rule_move_eat && cfg->rule_utilicalc) return fail("moveEat and utilicalc cannot be used together");
    if (cfg->rule_utilicalc && cfg->rule_pollution) return fail("utilicalc+pollution crashes in the reference (agent.py:880)");
    if (cfg->rule_pollution && cfg->diffusion_rate < 1) return fail("diffusion_rate must be >= 1");
    if (cfg->rule_seasons && cfg->season_period < 1) return fail("season period must be >= 1");
    if (cfg->rule_spice && (!cfg->rule_move_eat || cfg->rule_pollution)) return fail("spice: with rule M, without pollution (sugarscape.py:195-197)");
    sso_world* w = (sso_world*)calloc(1, sizeof *w);
    if (!w) return fail("out of memory");
    w->cfg = *cfg;
    w->n = cfg->grid_size;
    size_t cells = (size_t)w->n * (size_t)w->n;
    w->sugar = (double*)calloc(cells, sizeof(double));
    w->cap = (double*)calloc(cells, sizeof(double));
    w->poll = (double*)calloc(cells, sizeof(double));
    w->occ = (int32_t*)malloc(cells * sizeof(int32_t));
    w->spice = (double*)calloc(cells, sizeof(double));
    w->spice_cap = (double*)calloc(cells, sizeof(double));
    if (!w->sugar || !w->cap || !w->poll || !w->occ || !w->spice || !w->spice_cap) { sso_destroy(w); return fail("out of memory"); }
    for (size_t c = 0; c < cells; c++) w->occ[c] = -1;
    w->iteration = cfg->iteration;
    w->env_time = cfg->env_time;
    w->mti = 624;
    w->half = 1;
    *out = w;
    return 0;
}

static void free_agents(sso_world* w) {
    free(w->ax); free(w->ay); free(w->avision); free(w->ametab); free(w->asugar); free(w->alive);
    free(w->order); free(w->wealth); free(w->sortk); free(w->sortk2);
    free(w->age); free(w->max_age); free(w->sex); free(w->fert_lo); free(w->fert_hi); free(w->tags); free(w->endow); free(w->living);
    free(w->aspice); free(w->sp_endow); free(w->asp_metab);
    w->aspice = w->sp_endow = NULL; w->asp_metab = NULL;
    w->ax = w->ay = w->avision = w->ametab = w->order = w->age = w->max_age = w->sex = w->fert_lo = w->fert_hi = w->tags = NULL;
    w->asugar = w->wealth = w->endow = NULL; w->alive = w->living = NULL; w->sortk = w->sortk2 = NULL;
    w->cap_slots = 0;
}
void sso_destroy(sso_world* w) {
    if (!w) return;
    free(w->sugar); free(w->cap); free(w->poll); free(w->occ); free(w->spice); free(w->spice_cap);
    free_agents(w);
    free(w->events);
    free(w);
}

int sso_upload_cells(sso_world* w, const double* sugar, const double* cap, const double* pollution) {
    size_t cells = (size_t)w->n * (size_t)w->n;
    memcpy(w->sugar, sugar, cells * sizeof(double));
    memcpy(w->cap, cap, cells * sizeof(double));
    if (pollution) memcpy(w->poll, pollution, cells * sizeof(double));
    else memset(w->poll, 0, cells * sizeof(double));
    return 0;
}

int sso_upload_agents(sso_world* w, int32_t n, const int32_t* id, const int32_t* x, const int32_t* y,
                      const int32_t* vision, const int32_t* metab, const double* sugar) {
    int32_t max_id = 0;
    for (int i = 0; i < n; i++) {
        if (id[i] < 1) return fail("agent ids must be >= 1");
        if (id[i] > max_id) max_id = id[i];
        if (vision[i] < 1 || vision[i] > 32) return fail("vision out of range [1,32]");
        if (2 * vision[i] + 1 > w->n) return fail("2*vision+1 exceeds the grid size");
        if (metab[i] < 1) return fail("metabolism must be >= 1");
        if (x[i] < 0 || x[i] >= w->n || y[i] < 0 || y[i] >= w->n) return fail("agent off the grid");
    }
    free_agents(w);
    int32_t ns = max_id > 0 ? max_id : 1;
    if (w->min_slots > ns) ns = w->min_slots;      /* a resumed run keeps the order width of the original upload */
    w->n_slots = ns;
    w->id_increment = max_id;
    w->have_life = 0;
    w->have_spice = 0;
    if (ensure_capacity(w, ns)) return -1;         /* every per-slot array, the list and the scratch arrays (n <= ns) */
    size_t cells = (size_t)w->n * (size_t)w->n;
    for (size_t c = 0; c < cells; c++) w->occ[c] = -1;
    for (int i = 0; i < n; i++) {
        int s = id[i] - 1;
        if (w->alive[s]) return fail("duplicate agent id");
        int c = x[i] * w->n + y[i];
        if (w->occ[c] >= 0) return fail("two agents on one cell");
        w->alive[s] = 1;
        w->ax[s] = x[i]; w->ay[s] = y[i];
        w->avision[s] = vision[i]; w->ametab[s] = metab[i];
        w->asugar[s] = sugar[i];
        w->occ[c] = s;
        w->order[i] = s;
    }
    w->n_order = n;
    w->half = order_half_bits(ns);
    return 0;
}

/* The keyed execution order is a bijection on a power-of-two domain sized by the largest Agent.id at upload
 * (sugarscape.py:517 as restated by oracle/keyed_rng.py); a run resumed from a checkpoint must keep that size even if the
 * agent with the largest id has died since -- before sso_upload_agents. */
int sso_set_min_slots(sso_world* w, int32_t n_slots) { w->min_slots = n_slots; return 0; }
int32_t sso_get_slots(sso_world* w) { return w->n_slots; }

int sso_set_rng_mt19937(sso_world* w, const uint32_t state[624], int32_t index) {
    if (index < 0 || index > 624) return fail("MT19937 index out of range");
    memcpy(w->mt, state, 624 * sizeof(uint32_t));
    w->mti = index;
    return 0;
}
int sso_get_rng_mt19937(sso_world* w, uint32_t state[624], int32_t* index) {
    memcpy(state, w->mt, 624 * sizeof(uint32_t));
    *index = w->mti;
    return 0;
}

/* Environment.grid slots [1] spice and [3] spice capacity; Agent.spice / spiceMetabolism / spiceEndowment of the n
 * agents of the last sso_upload_agents, same order (endowment may be NULL: = spice) */
int sso_upload_cells_spice(sso_world* w, const double* spice, const double* spice_cap) {
    size_t cells = (size_t)w->n * (size_t)w->n;
    memcpy(w->spice, spice, cells * sizeof(double));
    memcpy(w->spice_cap, spice_cap, cells * sizeof(double));
    return 0;
}
int sso_download_cells_spice(sso_world* w, double* spice, double* spice_cap) {
    size_t cells = (size_t)w->n * (size_t)w->n;
    if (spice) memcpy(spice, w->spice, cells * sizeof(double));
    if (spice_cap) memcpy(spice_cap, w->spice_cap, cells * sizeof(double));
    return 0;
}
int sso_upload_agents_spice(sso_world* w, int32_t n, const int32_t* spice_metab, const double* spice, const double* endowment) {
    if (n != w->n_order) return fail("sso_upload_agents_spice: n differs from the uploaded agents");
    for (int i = 0; i < n; i++) {
        int s = w->order[i];
        if (spice_metab[i] < 1) return fail("spice metabolism must be >= 1");
        w->asp_metab[s] = spice_metab[i]; w->aspice[s] = spice[i]; w->sp_endow[s] = endowment ? endowment[i] : spice[i];
    }
    w->have_spice = 1;
    return 0;
}
int sso_download_agents_spice(sso_world* w, int32_t* spice_metab, double* spice, double* endowment) {
    for (int i = 0; i < w->n_order; i++) {
        int s = w->order[i];
        if (spice_metab) spice_metab[i] = w->asp_metab[s];
        if (spice) spice[i] = w->aspice[s];
        if (endowment) endowment[i] = w->sp_endow[s];
    }
    return 0;
}
int sso_pow_faults(void) { return g_pow_bad; }

int sso_step(sso_world* w, int32_t nsteps, sso_step_stats* out) {
    sso_step_stats tmp;
    if (w->cfg.rule_spice && !w->have_spice) return fail("spice: sso_upload_agents_spice first");
    if ((w->cfg.rule_limited_lifespan || w->cfg.rule_procreate) && !w->have_life)
        return fail("limitedLifespan / procreate: sso_upload_agents_life first");
    w->n_events = 0;
    for (int t = 0; t < nsteps; t++) { w->step_in_call = t; one_step(w, out ? &out[t] : &tmp); }
    return 0;
}

/* the Agent fields of the lifespan / procreation rules for the n agents of the last sso_upload_agents, same
 * order; id_increment = Environment.idIncrement (ids handed out so far, >= the largest id) */
int sso_upload_agents_life(sso_world* w, int32_t n, int32_t id_increment, const int32_t* age, const int32_t* max_age,
                           const int32_t* sex, const int32_t* fert_lo, const int32_t* fert_hi, const double* endowment,
                           const int32_t* tags, const int32_t* alive) {
    if (n != w->n_order) return fail("sso_upload_agents_life: n differs from the uploaded agents");
    if (w->cfg.rng_mode != 0) return fail("limitedLifespan / procreate: MT19937 mode only");
    if (w->cfg.rule_utilicalc) return fail("limitedLifespan / procreate: rule M only");
    if (w->cfg.tags_length < 0 || w->cfg.tags_length > 30) return fail("tags_length out of range [0,30]");
    if (id_increment < w->n_slots && id_increment < w->id_increment) return fail("id_increment below the largest id");
    for (int i = 0; i < n; i++) {
        int s = w->order[i];
        w->age[s] = age[i]; w->max_age[s] = max_age[i]; w->sex[s] = sex[i];
        w->fert_lo[s] = fert_lo[i]; w->fert_hi[s] = fert_hi[i];
        w->endow[s] = endowment[i]; w->tags[s] = tags[i];
        w->living[s] = (uint8_t)(alive ? (alive[i] != 0) : 1);
    }
    if (id_increment > w->id_increment) w->id_increment = id_increment;
    w->have_life = 1;
    return 0;
}
int sso_download_agents_life(sso_world* w, int32_t* age, int32_t* max_age, int32_t* sex, int32_t* fert_lo, int32_t* fert_hi,
                             double* endowment, int32_t* tags, int32_t* alive) {
    if (!w->have_life) return fail("no life columns uploaded");
    for (int i = 0; i < w->n_order; i++) {
        int s = w->order[i];
        if (age) age[i] = w->age[s];
        if (max_age) max_age[i] = w->max_age[s];
        if (sex) sex[i] = w->sex[s];
        if (fert_lo) fert_lo[i] = w->fert_lo[s];
        if (fert_hi) fert_hi[i] = w->fert_hi[s];
        if (endowment) endowment[i] = w->endow[s];
        if (tags) tags[i] = w->tags[s];
        if (alive) alive[i] = w->living[s];
    }
    return 0;
}
int32_t sso_id_increment(sso_world* w) { return w->id_increment; }
/* events of the last sso_step call, in execution order: 4 ints each (see struct sso_world) */
int32_t sso_event_count(sso_world* w) { return w->n_events; }
int sso_get_events(sso_world* w, int32_t max_events, int32_t* out) {
    int32_t k = w->n_events < max_events ? w->n_events : max_events;
    if (k > 0) memcpy(out, w->events, (size_t)k * 4 * sizeof(int32_t));
    return 0;
}

int sso_download_cells(sso_world* w, double* sugar, double* cap, double* pollution, int32_t* occ_id) {
    size_t cells = (size_t)w->n * (size_t)w->n;
    if (sugar) memcpy(sugar, w->sugar, cells * sizeof(double));
    if (cap) memcpy(cap, w->cap, cells * sizeof(double));
    if (pollution) memcpy(pollution, w->poll, cells * sizeof(double));
    if (occ_id)
        for (size_t c = 0; c < cells; c++) occ_id[c] = w->occ[c] >= 0 ? w->occ[c] + 1 : 0;
    return 0;
}

int sso_agent_count(sso_world* w) { return w->n_order; }

int sso_download_agents(sso_world* w, int32_t* id, int32_t* x, int32_t* y, int32_t* vision,
                        int32_t* metab, double* sugar) {
    for (int i = 0; i < w->n_order; i++) {
        int s = w->order[i];
        if (id) id[i] = s + 1;
        if (x) x[i] = w->ax[s];
        if (y) y[i] = w->ay[s];
        if (vision) vision[i] = w->avision[s];
        if (metab) metab[i] = w->ametab[s];
        if (sugar) sugar[i] = w->asugar[s];
    }
    return 0;
}
