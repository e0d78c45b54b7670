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
 *   Environment.grow/growRegion environment.py:133-155
 *   Environment.polluteSite    environment.py:312-315
 *   Environment.spreadPollution environment.py:317-350
 * RNG: CPython 3.12 random.py:242 (_randbelow_with_getrandbits), :350 (shuffle),
 *   :341 (choice) over _randommodule.c's MT19937 (getrandbits(k<=32) =
 *   genrand_uint32() >> (32-k)); or the keyed word source of oracle/keyed_rng.py.
 */
#include "ss_oracle.h"

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
    int32_t* occ;             /* slot or -1 */
    /* agent store, indexed by slot = id - 1 */
    int32_t n_slots;
    int32_t min_slots;
    int32_t *ax, *ay, *avision, *ametab;
    double* asugar;
    uint8_t* alive;
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
/* agent.py:224-229 setSugar (limitedLifespan off -> floor .01) */
static inline double set_sugar(double amount) { return amount > 0.01 ? amount : 0.01; }
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
    w->asugar[s] = set_sugar(max0(w->asugar[s] + (double)best_sg));            /* agent.py:832 */
    int old = x * n + y;
    if (best != old) {                                                          /* agent.py:856-862 */
        w->occ[old] = -1;
        w->occ[best] = s;
        w->ax[s] = best / n;
        w->ay[s] = best % n;
    }
    w->sugar[best] = 0.0;                                                       /* agent.py:864 */
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
    w->asugar[s] = set_sugar(max0(w->asugar[s] + (double)sg));
    w->sugar[dest] = 0.0;
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
    double sum_m = 0.0, sum_v = 0.0;
    int nw = 0, kept = 0, acted = 0, deaths = 0, skip_next = 0;
    /* sugarscape.py:520: `for agent in self.agents` while removeAgent() does list.remove:
       the agent following each dying agent loses its turn (Q1). */
    for (int i = 0; i < w->n_order; i++) {
        int s = w->order[i];
        if (skip_next) { skip_next = 0; w->order[kept++] = s; continue; }
        if (cfg->rng_mode == 1) keyed_begin_agent(w, s + 1);
        if (cfg->rule_move_eat) agent_move(w, s);
        if (cfg->rule_utilicalc) agent_utilicalc(w, s);
        /* sugarscape.py:566-567 consumption pollution */
        if (cfg->rule_pollution && w->asugar[s] > 0.0)
            pollute_site(w, w->ax[s] * n + w->ay[s], 0.0, (double)w->ametab[s]);
        /* sugarscape.py:569 */
        w->asugar[s] = set_sugar(max0(w->asugar[s] - (double)w->ametab[s]));
        sum_m += (double)w->ametab[s];
        sum_v += (double)w->avision[s];
        w->wealth[nw++] = w->asugar[s];
        acted++;
        if (cfg->rule_can_starve && w->asugar[s] <= 0.1) {      /* sugarscape.py:306-309, 593-601 */
            w->alive[s] = 0;
            w->occ[w->ax[s] * n + w->ay[s]] = -1;
            deaths++;
            skip_next = 1;
        } else {
            w->order[kept++] = s;
        }
    }
    w->n_order = kept;
    st->population = (double)kept;
    st->metabolism_mean = kept > 0 ? sum_m / (double)kept : 0.0;
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
    sso_world* w = (sso_world*)calloc(1, sizeof *w);
    if (!w) return fail("out of memory");
    w->cfg = *cfg;
    w->n = cfg->grid_size;
    size_t cells = (size_t)w->n * (size_t)w->n;
    w->sugar = (double*)calloc(cells, sizeof(double));
    w->cap = (double*)calloc(cells, sizeof(double));
    w->poll = (double*)calloc(cells, sizeof(double));
    w->occ = (int32_t*)malloc(cells * sizeof(int32_t));
    if (!w->sugar || !w->cap || !w->poll || !w->occ) { sso_destroy(w); return fail("out of memory"); }
    for (size_t c = 0; c < cells; c++) w->occ[c] = -1;
    w->iteration = cfg->iteration;
    w->env_time = cfg->env_time;
    w->mti = 624;
    w->half = 1;
    *out = w;
    return 0;
}

void sso_destroy(sso_world* w) {
    if (!w) return;
    free(w->sugar); free(w->cap); free(w->poll); free(w->occ);
    free(w->ax); free(w->ay); free(w->avision); free(w->ametab); free(w->asugar); free(w->alive);
    free(w->order); free(w->wealth); free(w->sortk); free(w->sortk2);
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
    free(w->ax); free(w->ay); free(w->avision); free(w->ametab); free(w->asugar); free(w->alive);
    free(w->order); free(w->wealth); free(w->sortk); free(w->sortk2);
    int32_t ns = max_id > 0 ? max_id : 1;
    if (w->min_slots > ns) ns = w->min_slots;      /* a resumed run keeps the order width of the original upload */
    w->n_slots = ns;
    w->ax = (int32_t*)calloc((size_t)ns, sizeof(int32_t));
    w->ay = (int32_t*)calloc((size_t)ns, sizeof(int32_t));
    w->avision = (int32_t*)calloc((size_t)ns, sizeof(int32_t));
    w->ametab = (int32_t*)calloc((size_t)ns, sizeof(int32_t));
    w->asugar = (double*)calloc((size_t)ns, sizeof(double));
    w->alive = (uint8_t*)calloc((size_t)ns, 1);
    w->order = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    w->wealth = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    w->sortk = (uint64_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(uint64_t));
    w->sortk2 = (uint64_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(uint64_t));
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

int sso_step(sso_world* w, int32_t nsteps, sso_step_stats* out) {
    sso_step_stats tmp;
    for (int t = 0; t < nsteps; t++) one_step(w, out ? &out[t] : &tmp);
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
