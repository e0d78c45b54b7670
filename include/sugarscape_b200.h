/* sugarscape_b200.h -- C-ABI of the B200-native Sugarscape timestep engine.
 *
 * Drop-in boundary for the per-timestep hot path of joshuapalicka/sugarscape-utilicalc.
 * The reference has no FFI of its own; the seam is its L3 entry point
 *     View.updateGame(self) -> None                     sugarscape.py:506-686
 * called once per frame from View.step (sugarscape.py:1368-1372).  Everything that call
 * reads and writes -- the lattice (Environment.grid, environment.py:24-34), the agents
 * (Agent fields, agent.py:120-165), the rule switches and parameters flattened from
 * settings.json (sugarscape.py:26-162), the global `random` state (sugarscape.py:160-162)
 * and the four stat series it appends to (sugarscape.py:603-619) -- crosses this boundary
 * as plain pointers and sizes.  INTEGRATION.md shows the ctypes stub a maintainer would
 * add to sugarscape.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; the message is
 *     available from ss_last_error(engine) (ss_last_error(NULL) for ss_create failures).
 *   - the caller owns every host buffer; the engine owns device memory and keeps no host
 *     pointer past a call.
 *   - one host thread per engine, non-reentrant; the engine owns one CUDA stream.
 *   - lattice arrays are (N,N) row-major with linear index x*N + y, where x is the
 *     reference's FIRST grid index (grid[i][j], i = agent.x) -- "lattice-major".
 *   - unsupported rule combinations fail in ss_create: there is no CPU fallback.
 *
 * Limits (a run outside them fails with SS_ERR_UNSUPPORTED in ss_create / ss_upload_agents, never silently)
 *   - Agent.id <= 2^24 = 16,777,216 on the lattice paths (a cell's occupancy word keeps the slot in 24 bits; BASELINE
 *     config 5 sits exactly on it), <= 2^22 on the serial path (SS_RNG_MT19937: one CTA walks the list).
 *   - vision <= 15 and N >= 4 * max vision + 1 on the lattice paths (larger: serial path).
 *   - cell sugar and capacity < 127 for the dependency rounds (one byte per cell: occupied | int(sugar)), < 256 for the
 *     window passes; larger capacities run on the serial path.
 *   - wealth statistics: exact for any wealth (integer wealth < 65536 through a histogram, everything else through an
 *     outlier list of 4096 values per step; beyond that the step reports gini = NaN and raises the `inexact` counter).
 *   - row strips (ss_create_strip): rule M without pollution, keyed order, >= 4 * max vision + 1 rows per strip.
 */
#ifndef SUGARSCAPE_B200_H
#define SUGARSCAPE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_RNG_MT19937 0   /* CPython-exact global stream (serial order; shipped presets) */
#define SS_RNG_KEYED   1   /* keyed word source (parallel dependency rounds; scale configs) */

#define SS_OK            0
#define SS_ERR_ARG      -1
#define SS_ERR_UNSUPPORTED -2
#define SS_ERR_CUDA     -3
#define SS_ERR_STATE    -4

/* Replaces the module globals sugarscape.py:32-160 builds from settings.json
 * (schema kept verbatim; see sugarscape_utilicalc_b200/settings.py). */
typedef struct {
    int32_t grid_size;            /* view.grid_size.x (== y; the reference is square-only) */
    int32_t rule_grow;            /* rules.grow           sugarscape.py:657-659            */
    int32_t rule_seasons;         /* rules.seasons        sugarscape.py:642-655            */
    int32_t rule_move_eat;        /* rules.moveEat        sugarscape.py:522-523            */
    int32_t rule_utilicalc;       /* rules.utilicalc      sugarscape.py:525-529 (sugar-only) */
    int32_t rule_can_starve;      /* rules.canStarve      sugarscape.py:585-586            */
    int32_t rule_pollution;       /* rules.pollution      sugarscape.py:566-567, 661-663   */
    int32_t season_period;        /* rule_params.seasons.period                            */
    int32_t season_north[4];      /* start_x, start_y, end_x, end_y (inclusive rectangle)  */
    int32_t season_south[4];
    int32_t diffusion_rate;       /* rule_params.pollution.diffusion_rate                  */
    int32_t pollution_start_time; /* rule_params.pollution.pollution_start_time            */
    int32_t diffusion_start_time; /* rule_params.pollution.diffusion_start_time            */
    int32_t rng_mode;             /* SS_RNG_*                                              */
    double  grow_factor;          /* env.growth_factor.grow_factor                         */
    double  season_on_factor;     /* grow_factor_on_season                                 */
    double  season_off_factor;    /* on / grow_factor_off_season_divisor (sugarscape.py:85) */
    double  pollution_a;          /* rule_params.pollution.A (production)                  */
    double  pollution_b;          /* rule_params.pollution.B (consumption)                 */
    double  max_capacity;         /* env.max_capacity                                      */
    uint64_t keyed_seed;          /* seed of the keyed word source (SS_RNG_KEYED)          */
    int64_t iteration;            /* View.iteration at upload                              */
    int64_t env_time;             /* Environment.time at upload                            */
    /* limited lifespan + procreation (serial path, SS_RNG_MT19937; ss_upload_agents_life) */
    int32_t rule_limited_lifespan;/* rules.limitedLifespan sugarscape.py:588-597; setSugar floors at 0 (agent.py:226) */
    int32_t rule_procreate;       /* rules.procreate       sugarscape.py:543-554, agent.py:653-773 */
    int32_t tags_length;          /* rule_params.tags.tags_length: createChild draws random() per differing tag bit
                                     (agent.py:727-737), with or without the tags rule                               */
    int32_t foresight_lo;         /* Environment.foresightRange, (0, 0) unless rules.foresight: Agent.__init__ of a  */
    int32_t foresight_hi;         /* child spends randint(lo, hi) on generateForesight (agent.py:165, 329-333)       */
    int32_t rule_spice;           /* rules.spice: second resource, Cobb-Douglas welfare in move()
                                     (agent.py:833-853, 1220-1251), spice metabolism / starvation (sugarscape.py:571-572,
                                     311-314); serial path; ss_upload_cells_spice + ss_upload_agents_spice               */
} ss_config;

/* What one updateGame() appends to View.population / metabolismMean / visionMean / gini
 * (sugarscape.py:603-619), plus two counters. */
typedef struct {
    double population;
    double metabolism_mean;
    double vision_mean;
    double gini;
    double acted;                 /* agents that took their turn this step                 */
    double deaths;
} ss_step_stats;

typedef struct ss_engine ss_engine;

const char* ss_version(void);

/* Replaces: module import + Environment(gridSize) + the rule setters of sugarscape.py:1485-1519.
 * device = CUDA ordinal. */
int  ss_create(const ss_config* cfg, int32_t device, ss_engine** out);
void ss_destroy(ss_engine* e);
const char* ss_last_error(ss_engine* e);

/* Replaces: `random.seed(seed)` at import (sugarscape.py:160-162) for a non-negative integer seed: CPython seeds
 * MT19937 with init_by_array over the seed's 32-bit words (Modules/_randommodule.c). */
int  ss_seed_mt19937(ss_engine* e, uint64_t seed);

/* The world the reference's `__main__` builds before the first step (sugarscape.py:1485-1555), SURVEY 8 f4. */
typedef struct {
    int32_t n_sites;              /* sugar sites added with addSugarSite (sugarscape.py:1494-1497: ne, sw) */
    int32_t site_x[4], site_y[4], site_radius[4];     /* env.sites.<name>.{x,y,radius}                     */
    int32_t grow_to_capacity;     /* rules.grow: env.grow(maxCapacity) before the agents (sugarscape.py:1525) */
    int32_t n_groups;             /* entries of `distributions` (sugarscape.py:141-146): blue, red         */
    int32_t group_count[4];       /* env.total_agents.{blue,red}                                           */
    int32_t max_metabolism;       /* agents.max_agent_metabolism                                           */
    int32_t max_vision;           /* agents.max_agent_vision                                               */
    int32_t endowment_min, endowment_max;             /* agents.initial_endowments                         */
    int32_t age_min, age_max;     /* agents.death_age_range (drawn, unused without limitedLifespan)        */
    int32_t childbearing[2][4];   /* by the drawn sex 0/1: (min_start, max_start, min_end, max_end) of
                                     `childbearing = fertility[0], fertility[1]` (sugarscape.py:103-108): the
                                     reference indexes the "male" tuple with sex 0                          */
    int32_t foresight_lo, foresight_hi;               /* Environment.foresightRange, (0,0) without the rule:
                                     Agent.__init__ draws from it regardless (agent.py:165, 329-333)        */
    int32_t group_tags[4];        /* `distributions`: the tags of each group (tags0, 2^tags_length - 1; sugarscape.py:104-105,
                                     142-148) -- kept with limitedLifespan / procreate, where ss_init_world also
                                     leaves the engine as after ss_upload_agents_life (age 0, maxAge, sex, fertility drawn) */
    int32_t n_spice_sites;        /* rules.spice: addSpiceSite(southeast), (northwest) (sugarscape.py:1508-1513);   */
    int32_t spice_site_x[4], spice_site_y[4], spice_site_radius[4];   /* each agent then also draws spiceMetabolism and its
                                     spice endowment (sugarscape.py:253-255); the engine is left as after the spice uploads   */
} ss_init_params;

/* Replaces: Environment.addSugarSite x n_sites (environment.py:114-126), setMaxCapacity, env.grow(maxCapacity),
 * and per agent `Agent(env)` + `initAgent` (sugarscape.py:242-265: getRandomFreeLocation environment.py:173-182 over
 * range(0, N-1) x range(0, N-1), metabolism, endowment, vision, age, sex, fertility) with every draw taken from the
 * engine's MT19937 stream in the reference's order; agents get ids 1.. in creation order (an agent that finds no free
 * cell keeps its id and is dropped).  Leaves the engine as after ss_upload_cells + ss_upload_agents.  Needs
 * SS_RNG_MT19937 state (ss_seed_mt19937 / ss_set_rng_mt19937); keyed engines may call it too -- the initial world is
 * then the same one, the stepping differs. */
int  ss_init_world(ss_engine* e, const ss_init_params* p);

/* Replaces: Environment.grid slots [0] sugar, [2] capacity, [5] pollution (environment.py:24-34).
 * pollution may be NULL (zeros). */
int  ss_upload_cells(ss_engine* e, const double* sugar, const double* cap, const double* pollution);

/* Replaces: View.agents (list order = array order) and Environment.grid slot [4].
 * id[] are the reference's Agent.id (>= 1, unique); x,y = Agent.x,y; vision, metab =
 * Agent.vision, sugarMetabolism; sugar = Agent.sugar (agent.py:120-165). */
int  ss_upload_agents(ss_engine* e, int32_t n, const int32_t* id, const int32_t* x, const int32_t* y,
                      const int32_t* vision, const int32_t* metab, const double* sugar);

/* SURVEY 8 f3 -- limitedLifespan + procreate (rule M, SS_RNG_MT19937, serial path).
 * Replaces: Agent.age / maxAge / sex / fertility / sugarEndowment (= sugarCostForMating) / tags / alive (agent.py:134-151,
 * 204-206) as read by incAge (agent.py:636-639), isFertile (764-773), mate (653-679), createChild (698-750) and
 * updateGame's aging / removal branch (sugarscape.py:543-554, 588-601); Environment.idIncrement (environment.py:202-204).
 * ss_upload_agents_life: the columns of the n agents of the last ss_upload_agents, same order (alive may be NULL = all 1);
 * required before ss_step when either rule is on.  ss_download_agents_life: the same columns in View.agents order
 * (ss_agent_count entries -- an agent that reached maxAge stays in the list with alive = 0 for one more step, exactly
 * as in the reference).  Children get ids id_increment + 1, ...; ss_download_agents / ss_download_cells report them.
 * ss_event_count / ss_get_events: the appendToLog events of the last ss_step call in execution order, 4 ints each:
 * kind | (step << 8), a, b, c -- kind 1 "Agent a mated with agent b and created agent c" (sugarscape.py:551), 2 "Agent a
 * died of starvation", 3 "Agent a died of old age" (sugarscape.py:597-600). */
int  ss_upload_agents_life(ss_engine* e, int32_t n, int32_t id_increment, const int32_t* age, const int32_t* max_age,
                           const int32_t* sex, const int32_t* fert_lo, const int32_t* fert_hi, const double* endowment,
                           const int32_t* tags, const int32_t* alive);
int  ss_download_agents_life(ss_engine* e, int32_t* age, int32_t* max_age, int32_t* sex, int32_t* fert_lo, int32_t* fert_hi,
                             double* endowment, int32_t* tags, int32_t* alive, int32_t* id_increment);
int  ss_event_count(ss_engine* e, int32_t* count);
int  ss_get_events(ss_engine* e, int32_t max_events, int32_t* out);

/* SURVEY 8 f2 -- spice (rule M, serial path).
 * Replaces: Environment.grid slots [1] spice and [3] spice capacity (environment.py:24-34, 102-104, 141-142) and
 * Agent.spice / spiceMetabolism / spiceEndowment (= spiceCostForMating; agent.py:127-131, 210-216) as read by the spice branch
 * of move() (agent.py:833-853: arg-max of the Cobb-Douglas welfare getWelfare, 1220-1251 -- w1 ** (m1/mt) * w2 ** (m2/mt),
 * with `x ** y` computed as the reference's libm does, bit for bit: csrc/ss_pow.h), by the spice metabolism and starvation
 * of updateGame (sugarscape.py:571-572, 311-314) and by createChild / isFertile (agent.py:712-714, 768-771).
 * Both uploads are required before ss_step when rules.spice is on; endowment may be NULL (= spice). */
int  ss_upload_cells_spice(ss_engine* e, const double* spice, const double* spice_cap);
int  ss_download_cells_spice(ss_engine* e, double* spice, double* spice_cap);
int  ss_upload_agents_spice(ss_engine* e, int32_t n, const int32_t* spice_metab, const double* spice, const double* endowment);
int  ss_download_agents_spice(ss_engine* e, int32_t* spice_metab, double* spice, double* endowment);

/* Replaces: the global `random` state (random.getstate()[1]; sugarscape.py:160-162). */
int  ss_set_rng_mt19937(ss_engine* e, const uint32_t state[624], int32_t index);
int  ss_get_rng_mt19937(ss_engine* e, uint32_t state[624], int32_t* index);
int  ss_set_rng_keyed(ss_engine* e, uint64_t seed);

/* Replaces: nsteps calls of View.updateGame() (sugarscape.py:506-686).  out may be NULL,
 * else it receives nsteps entries. */
int  ss_step(ss_engine* e, int32_t nsteps, ss_step_stats* out);

/* Replaces: len(View.agents). */
int  ss_agent_count(ss_engine* e);

/* Replaces: reads of Environment.grid (GUI draw, sugarscape.py:713-750).  Any pointer may be
 * NULL.  occ_id receives the occupant's Agent.id or 0. */
int  ss_download_cells(ss_engine* e, double* sugar, double* cap, double* pollution, int32_t* occ_id);

/* Replaces: reads of View.agents.  Living agents, in list order for SS_RNG_MT19937 and in
 * ascending keyed priority of the last step for SS_RNG_KEYED (what the reference's list
 * holds after its shuffle).  Any pointer may be NULL. */
int  ss_download_agents(ss_engine* e, int32_t* id, int32_t* x, int32_t* y, int32_t* vision,
                        int32_t* metab, double* sugar);

/* Page-locked host buffers for the arrays that cross the boundary (the copies then run at PCIe speed without a
 * staging pass).  Optional: every entry point accepts ordinary pageable memory too.  NULL on failure. */
void* ss_host_alloc(size_t bytes);
void  ss_host_free(void* p);

/* Instrumentation.  counters: [0] kernel launches since create, [1] move passes,
 * [2] steps, [3] path (0 serial fused, 1 serial + lattice env kernels, 2 lattice),
 * [4] host syncs, [5] window extent | move-pass implementation << 16 (0 = CTA per window, 1 = warp per window, 2 = warp per window with a lane
 * per agent); [6] watchdog firings and [10] internal faults of the lattice kernels (both must stay 0). */
int  ss_get_counters(ss_engine* e, int64_t* out, int32_t n);
/* accumulated device milliseconds per kernel class (needs option "profile"=1):
 * [0] prepare, [1] move passes, [2] stats, [3] grow/seasons, [4] diffusion, [5] serial;
 * [6] device milliseconds of the last ss_step call (CUDA events on the engine's stream, always on);
 * [7] settle. */
int  ss_get_kernel_times(ss_engine* e, double* ms_out, int32_t n);
/* options: "path" (-1 auto, 0 serial, 2 lattice), "profile" (0/1), "reset_times",
 * "reset_counters", "diffuse_impl" (0 wavefront with several bands per CTA, 1 diagonal sweep and 2 one band per CTA, for A/B checks),
 * "window" (max window extent of the CTA-per-window move pass, 32..176), "pass_impl" (-1 auto, 0 CTA per window,
 * 1 warp per window taking its agents one at a time, 2 warp per window with one lane per agent -- bit-identical results,
 * measured slower, kept for A/B runs), "pendmap" (0/1 coarse pending map), "solo_window" (max window extent of the warp-per-window move pass; 0 = chosen from the lattice size) */
int  ss_set_option(ss_engine* e, const char* name, int64_t value);

/* Multi-GPU (one process per GPU; sugarscape_utilicalc_b200/sharded.py).  Every rank holds a replica of the
 * state and resolves the agents of its share of the window rows; the turns it resolved are exchanged as log entries
 * (NCCL all-gather) and replayed on the other replicas: 8 bytes for the turn of an agent that cannot starve (slot |
 * harvest << 24, new cell -- the rest follows from the replica's own record), 48 bytes (SS_SHARD_ENTRY_BYTES) for a
 * synchronously settled turn or a Q1 skip.  One step =
 *   ss_shard_begin_step; { ss_shard_pass; <all-gather logs>; ss_shard_apply(each other rank's entries) }
 *   until ss_shard_pending == 0; ss_shard_end_step.
 * log_base/log_begin/log_end (48-byte entries) and clog_* (8-byte entries) describe this rank's entries of the pass in
 * device memory. */
int  ss_shard_enable(ss_engine* e, int32_t rank, int32_t world);
int  ss_shard_begin_step(ss_engine* e);
int  ss_shard_pass(ss_engine* e, uint64_t* log_begin, uint64_t* log_end, const void** log_base, uint64_t* clog_begin,
                   uint64_t* clog_end, const void** clog_base);
int  ss_shard_apply(ss_engine* e, const void* dev_entries, uint64_t n, const void* dev_compact, uint64_t n_compact);
/* ss_shard_apply for every other rank at once, straight from the all-gathered buffer: rank j's slice starts at
 * base + j * stride (48-byte entries first, 8-byte entries at off_compact); counts = device int64 [world][2]. */
int  ss_shard_apply_gathered(ss_engine* e, const void* base, uint64_t stride, uint64_t off_compact, const int64_t* counts,
                             int32_t world, uint64_t max_full, uint64_t max_compact);
int  ss_shard_pending(ss_engine* e, uint64_t* pending);
int  ss_shard_end_step(ss_engine* e, ss_step_stats* out);
#define SS_SHARD_ENTRY_BYTES 48

/* Row strips (one engine per GPU; sugarscape_utilicalc_b200/strips.py) -- BASELINE config 5: "row-strip sharded with
 * NVLink halo exchange and agent migration".  Rank r of `world` holds the lattice rows [r N / world, (r+1) N / world) --
 * cells, occupancy and the records of the agents standing on them; the toroidal neighbour relation of the reference
 * (agent.py:362-368) makes the strips a ring.  Nothing is exchanged by the host: the kernels reach the two ring
 * neighbours' border rows (halo of the conflict graph, crosses and moves across the border, migrating agent records) and
 * every strip's dependency counters (the list-order quirk Q1 couples arbitrary agents) through peer-mapped device
 * pointers over NVLink; small replicated directories (alive / at-risk bitmaps, the strip that holds each agent) are
 * kept in step by per-strip delta lists; the per-step sums and the wealth histogram are merged over the strips, so
 * every rank reports the same statistics.
 *   ss_create_strip; ss_set_option("global_slots" = max Agent.id over all strips, "global_vmax" = max vision);
 *   ss_upload_cells (rows x N arrays of the strip); ss_upload_agents (the agents standing on the strip, lattice
 *   coordinates); ss_strip_init_directory (id, x, metabolism, sugar of the agents of ALL strips);
 *   ss_strip_ipc_export -> exchange the handles -> ss_strip_ipc_connect for every other rank (ss_strip_export /
 *   ss_strip_connect when the peers live in the same process); a barrier over the ranks; then ss_strip_step.
 * ss_download_cells / ss_download_agents / ss_agent_count return the strip's share.  SS_STRIP_NPTR pointers / handles
 * per strip. */
#define SS_STRIP_NPTR 11
#define SS_STRIP_HANDLE_BYTES 72   /* cudaIpcMemHandle_t of the allocation + offset of the array inside it */
int  ss_create_strip(const ss_config* cfg, int32_t device, int32_t rank, int32_t world, ss_engine** out);
int  ss_strip_init_directory(ss_engine* e, int32_t n_all, const int32_t* id, const int32_t* x, const int32_t* metab,
                             const double* sugar);
int  ss_strip_export(ss_engine* e, uint64_t* ptrs);
int  ss_strip_connect(ss_engine* e, int32_t peer_rank, const uint64_t* ptrs);
int  ss_strip_ipc_export(ss_engine* e, void* handles);
int  ss_strip_ipc_connect(ss_engine* e, int32_t peer_rank, const void* handles);
/* Replaces: nsteps calls of View.updateGame() on the whole lattice, this rank's share; the strips synchronise on the
 * device (flag words in peer memory between the phases and between the dependency rounds). */
int  ss_strip_step(ss_engine* e, int32_t nsteps, ss_step_stats* out);
/* The same step phase by phase with the host as the barrier (several strips on one GPU, one process: the parity
 * tests).  phase 0 begin + conflict graph, 1 one dependency round over the queue segment [seg_begin, seg_end), 2 after
 * the rounds, 3 statistics + growback.  Every strip finishes a phase before any strip starts the next. */
int  ss_strip_phase(ss_engine* e, int32_t phase, uint32_t seg_begin, uint32_t seg_end, uint32_t* tail_out,
                    ss_step_stats* stats_out);

#ifdef __cplusplus
}
#endif
#endif
