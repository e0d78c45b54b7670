/* ss_oracle.h -- CPU restatement of the reference timestep (TEST INFRASTRUCTURE).
 *
 * This library is the checker for the CUDA engine.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it;
 * the product package never does.  Every function cites the reference lines it
 * restates (paths relative to /root/reference).  Parity is pinned: the restatement
 * is compared step by step with the unmodified reference run headless
 * (oracle/ref_harness.py, tests/test_oracle_vs_reference.py) and with the golden
 * vectors that run produced (tests/golden/).
 */
#ifndef SS_ORACLE_H
#define SS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field layout as ss_config in include/sugarscape_b200.h (kept in sync by
 * tests/test_abi.py) so one ctypes structure fills both. */
typedef struct {
    int32_t grid_size;            /* settings view.grid_size.x (== y; square only, Q12)      */
    int32_t rule_grow;            /* rules.grow                                             */
    int32_t rule_seasons;         /* rules.seasons                                          */
    int32_t rule_move_eat;        /* rules.moveEat                                          */
    int32_t rule_utilicalc;       /* rules.utilicalc (sugar-only variant)                   */
    int32_t rule_can_starve;      /* rules.canStarve                                        */
    int32_t rule_pollution;       /* rules.pollution                                        */
    int32_t season_period;        /* rule_params.seasons.period                             */
    int32_t season_north[4];      /* start_x, start_y, end_x, end_y (inclusive)             */
    int32_t season_south[4];
    int32_t diffusion_rate;       /* rule_params.pollution.diffusion_rate                   */
    int32_t pollution_start_time;
    int32_t diffusion_start_time;
    int32_t rng_mode;             /* 0 = MT19937 (CPython-exact), 1 = keyed                 */
    double  grow_factor;          /* env.growth_factor.grow_factor                          */
    double  season_on_factor;     /* grow_factor_on_season                                  */
    double  season_off_factor;    /* on / grow_factor_off_season_divisor                    */
    double  pollution_a;          /* rule_params.pollution.A                                */
    double  pollution_b;          /* rule_params.pollution.B                                */
    double  max_capacity;         /* env.max_capacity                                       */
    uint64_t keyed_seed;          /* seed of the keyed word source                          */
    int64_t iteration;            /* View.iteration at upload                               */
    int64_t env_time;             /* Environment.time at upload                             */
    int32_t rule_limited_lifespan;/* rules.limitedLifespan                                  */
    int32_t rule_procreate;       /* rules.procreate                                        */
    int32_t tags_length;          /* rule_params.tags.tags_length                           */
    int32_t foresight_lo, foresight_hi;   /* Environment.foresightRange                     */
    int32_t rule_spice;           /* rules.spice                                            */
} sso_config;

typedef struct {
    double population;            /* len(self.agents) after the loop   sugarscape.py:604     */
    double metabolism_mean;       /* sugarscape.py:608-609                                  */
    double vision_mean;           /* sugarscape.py:614                                      */
    double gini;                  /* sugarscape.py:619, calculateGini 269-280               */
    double acted;                 /* agents that took their turn (Q1 skips excluded)        */
    double deaths;
} sso_step_stats;

typedef struct sso_world sso_world;

int  sso_create(const sso_config* cfg, sso_world** out);
void sso_destroy(sso_world* w);
/* cells are (N,N) row-major with index x*N + y (x = the reference's first index) */
int  sso_upload_cells(sso_world* w, const double* sugar, const double* cap, const double* pollution);
/* n agents in LIST ORDER (View.agents); id[] are the reference ids (>=1, unique) */
int  sso_upload_agents(sso_world* w, int32_t n, const int32_t* id, const int32_t* x, const int32_t* y,
                       const int32_t* vision, const int32_t* metab, const double* sugar);
int  sso_set_min_slots(sso_world* w, int32_t n_slots);   /* before sso_upload_agents: order width of a resumed run */
int32_t sso_get_slots(sso_world* w);
int  sso_set_rng_mt19937(sso_world* w, const uint32_t state[624], int32_t index);
int  sso_get_rng_mt19937(sso_world* w, uint32_t state[624], int32_t* index);
int  sso_step(sso_world* w, int32_t nsteps, sso_step_stats* out /* nsteps entries */);
int  sso_download_cells(sso_world* w, double* sugar, double* cap, double* pollution, int32_t* occ_id);
int  sso_agent_count(sso_world* w);
/* living agents in list order */
int  sso_download_agents(sso_world* w, int32_t* id, int32_t* x, int32_t* y, int32_t* vision,
                         int32_t* metab, double* sugar);
/* SURVEY 8 f3 (limitedLifespan / procreate; MT19937 mode, rule M): the Agent fields the rules read, for the n agents of
 * the last sso_upload_agents in the same order; id_increment = Environment.idIncrement */
int  sso_upload_agents_life(sso_world* w, int32_t n, int32_t id_increment, const int32_t* age, const int32_t* max_age,
                            const int32_t* sex, const int32_t* fert_lo, const int32_t* fert_hi, const double* endowment,
                            const int32_t* tags, const int32_t* alive);
int  sso_download_agents_life(sso_world* w, int32_t* age, int32_t* max_age, int32_t* sex, int32_t* fert_lo, int32_t* fert_hi,
                              double* endowment, int32_t* tags, int32_t* alive);
int32_t sso_id_increment(sso_world* w);
/* events of the last sso_step call in execution order, 4 ints each: kind | step << 8, a, b, c -- kind 1: a mated with b
 * and created c (sugarscape.py:551); 2: a died of starvation; 3: a died of old age (sugarscape.py:597-600) */
int32_t sso_event_count(sso_world* w);
int  sso_get_events(sso_world* w, int32_t max_events, int32_t* out);
/* Environment.grid slots [1], [3]; Agent.spice / spiceMetabolism / spiceEndowment (NULL: = spice) of the
 * agents of the last sso_upload_agents, same order */
int  sso_upload_cells_spice(sso_world* w, const double* spice, const double* spice_cap);
int  sso_download_cells_spice(sso_world* w, double* spice, double* spice_cap);
int  sso_upload_agents_spice(sso_world* w, int32_t n, const int32_t* spice_metab, const double* spice, const double* endowment);
int  sso_download_agents_spice(sso_world* w, int32_t* spice_metab, double* spice, double* endowment);
int  sso_pow_faults(void);   /* operands ss_pow_pos had no answer for (must stay 0) */
const char* sso_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
