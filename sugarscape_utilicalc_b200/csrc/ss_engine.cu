// ss_engine.cu -- host side of the C-ABI (include/sugarscape_b200.h): device memory, uploads /
// downloads and the per-timestep launch sequence.  No compute happens on the host; if a rule
// combination or geometry is not supported the call fails -- there is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ss_world.cuh"
#include "ss_life.cuh"

extern "C" {
cudaError_t ss_launch_serial(const DevWorld& w, const LifeWorld& lw, const SpiceWorld& sw, int64_t iteration, int64_t env_time, int nsteps, int fused_env,
                             int stats_base, cudaStream_t stream);
cudaError_t ss_launch_prepare(const DevWorld& w, int64_t iteration, int stamp_q1, cudaStream_t s);
size_t ss_dr_build_smem(int vmax);
cudaError_t ss_launch_dr_transpose(const DevWorld& w, cudaStream_t s);
cudaError_t ss_launch_dr_canon(const DevWorld& w, cudaStream_t s);
cudaError_t ss_launch_dr_grow(const DevWorld& w, int64_t iteration, int max_ctas, cudaStream_t s);
cudaError_t ss_launch_dr_mirror(const DevWorld& w, int* flag, cudaStream_t s);
cudaError_t ss_dr_encode_tmap(const DevWorld& w, void* tmap_storage, int* ok);
cudaError_t ss_launch_dr_build(const DevWorld& w, int64_t iteration, const void* tmap_storage, int use_tma, cudaStream_t s);
cudaError_t ss_launch_dr_rounds(const DevWorld& w, int64_t iteration, int64_t env_time, int sm_count, int single_round,
                                unsigned seg_begin, unsigned seg_end, int async_queue, cudaStream_t s);
cudaError_t ss_launch_dr_check(const DevWorld& w, cudaStream_t s);
cudaError_t ss_launch_dr_directory(const DevWorld& w, int n_all, const int32_t* id, const int32_t* x, const int32_t* metab,
                                   const double* sugar, cudaStream_t s);
cudaError_t ss_launch_dr_after_rounds(const DevWorld& w, cudaStream_t s);
cudaError_t ss_launch_dr_strip_barrier(const DevWorld& w, cudaStream_t s);
cudaError_t ss_launch_isugar(const DevWorld& w, int* flag, cudaStream_t s);
cudaError_t ss_launch_init_sites(const DevWorld& w, int n_sites, const int* sx, const int* sy, const double* D, double max_capacity,
                                 int grow, int sm_count, double* cap_out, double* val_out, cudaStream_t stream);
cudaError_t ss_launch_init_agents(const DevWorld& w, const ss_init_params* p, int32_t* rowfree, int32_t* superfree, int32_t* cols,
                                  size_t stride, double* sugar, int32_t* life_cols, int32_t* spice_cols, int32_t* result, cudaStream_t stream);
cudaError_t ss_launch_build_agents(const DevWorld& w, int n, uint32_t phase, const int32_t* id, const int32_t* x, const int32_t* y,
                                   const int32_t* vision, const int32_t* metab, const double* sugar, int* err, cudaStream_t s);
cudaError_t ss_launch_validate_rows(int n, const int32_t* x, int x0, int rows, int* out, cudaStream_t s);
cudaError_t ss_launch_occ_ids(const DevWorld& w, int32_t* ids, cudaStream_t s);
cudaError_t ss_launch_count_alive(const DevWorld& w, unsigned long long* out, cudaStream_t s);
cudaError_t ss_launch_validate_agents(int n, int N, const int32_t* id, const int32_t* x, const int32_t* y, const int32_t* vision,
                                      const int32_t* metab, int* out3, cudaStream_t s);
uint32_t ss_export_blocks(uint32_t domain);
cudaError_t ss_launch_export_agents(const DevWorld& w, uint64_t step_key, uint32_t* blockcnt, uint32_t* total, int32_t* cols,
                                    size_t stride, double* sugar, cudaStream_t s);
size_t ss_pass_smem_bytes(int emax, int ew, int vmax);
size_t ss_solo_smem_bytes(int emax, int ew, int mode);
cudaError_t ss_launch_move_solo(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time, int lanes, cudaStream_t s);
cudaError_t ss_launch_move_pass(const DevWorld& w, const PassGeom& g, int64_t iteration, int64_t env_time,
                                cudaStream_t s);
cudaError_t ss_launch_stats(const DevWorld& w, int stats_index, cudaStream_t s);
cudaError_t ss_launch_settle(const DevWorld& w, int64_t iteration, int64_t env_time, cudaStream_t s);
cudaError_t ss_launch_apply_log(const DevWorld& w, const void* entries, unsigned long long n, int phase, cudaStream_t s);
cudaError_t ss_launch_apply_clog(const DevWorld& w, const void* entries, unsigned long long n, int phase, int64_t iteration,
                                 cudaStream_t s);
cudaError_t ss_launch_apply_gathered(const DevWorld& w, const void* base, unsigned long long stride, unsigned long long off_compact,
                                     const long long* counts, int world, int skip_rank, unsigned long long max_full,
                                     unsigned long long max_compact, int phase, int64_t iteration, cudaStream_t s);
cudaError_t ss_launch_grow(const DevWorld& w, int64_t iteration, int sm_count, int dr_mirror, cudaStream_t s);
size_t ss_diffuse_handoff_bytes(int n);
cudaError_t ss_launch_diffuse(const DevWorld& w, int* ticket_progress, double* handoff, int impl, cudaStream_t s);
}

enum { PATH_SERIAL_FUSED = 0, PATH_SERIAL_ENV = 1, PATH_LATTICE = 2 };
enum { KT_PREPARE = 0, KT_MOVE, KT_STATS, KT_GROW, KT_DIFFUSE, KT_SERIAL, KT_LAST_CALL, KT_SETTLE, KT_COUNT };

struct ss_engine {
    ss_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    DevWorld w;
    size_t cells = 0;
    int64_t iteration = 0, env_time = 0;
    bool have_cells = false, have_agents = false, stepped = false;
    int path = -1, path_option = -1;
    PassGeom geom{};
    int ntile = 1, emin = 0;
    // move-pass implementation: 0 = one CTA per window, a warp per region (ss_move_pass_kernel);
    // 1 = one warp per window (ss_move_solo_kernel); -1 = by window count
    // 3 = dependency rounds over the whole lattice (ss_rounds.cu: ss_dr_build_kernel + ss_dr_rounds_kernel)
    int pass_impl_option = -1, pass_impl = 0;
    uint32_t dr_slots = 0;                     // allocation size of the dependency-round state
    int mirror = 0;                            // which int(sugar) mirror is current: 1 = isugar (window passes), 2 = mw (rounds)
    bool cap_fits_u7 = true;
    bool dr_dirty = false;                      // sugar / occw hold what the rounds tidy lazily (dr_tidy)
    // dependency rounds: statistics, growback and diffusion of step t run on a second stream beside the conflict-graph
    // kernel of step t + 1 (which reads occupancy and priorities only); the rounds of t + 1 wait for them
    cudaStream_t stream_env = nullptr;
    cudaEvent_t ev_rounds = nullptr, ev_env = nullptr;
    bool env_in_flight = false;
    int overlap_option = 1;
    int async_queue_strips = 1;                 // option "async_queue_strips": the same over the strips' peer-mapped queues
    int async_queue = 1;                        // option "async_queue": the rounds kernel as a work list without round barriers
    // SURVEY 8 f3 (limitedLifespan / procreate; serial path, MT19937 mode): see ss_life.cuh
    LifeWorld lw{};
    bool have_life = false;
    uint32_t slot_capacity = 0;                 // allocation of agents / order / wealth / sortkeys (>= n_slots; births need room)
    int32_t n_events = 0;                       // events of the last ss_step call
    // SURVEY 8 f2 (spice; serial path with the fused environment phase)
    SpiceWorld sw{};
    bool have_spice = false;
    int life_slack = -1;                        // option "life_capacity_slack" (tests): room for births instead of the default
    int grow_ctas_per_sm = 2;                   // ... as a few CTAs per SM that fit beside the conflict-graph CTAs
    // row strips (one engine per GPU): this engine holds the lattice rows [x0, x0 + rows)
    int strip_rank = 0, strip_world = 1, x0 = 0, rows = 0;
    uint32_t global_slots = 0;                  // max Agent.id over all strips (options "global_slots", "global_vmax")
    int global_vmax = 0;
    int peers_connected = 0;
    bool peers_ipc = false;
    void* ipc_ptrs[DR_MAX_STRIPS][16];
    int use_tma = 1, tma_ok = 0;
    int l2_persist = 1;
    alignas(64) unsigned char tmap[128];
    int solo_window_max = 0;                   // 0 = by lattice size
    int shard_rank = 0, shard_world = 1;       // multi-GPU replicas (sharded.py)
    int shard_pass = 0;
    int last_passes = 4;
    int stats_cap = 0;
    int* d_flags = nullptr;
    double* d_handoff = nullptr;                // diffusion: last rows handed from band to band
    int sm_count = 148;
    int diffuse_impl = 0;
    int window_max = 128;
    int use_pendmap = 1;
    int trace = 0;
    bool cap_fits_u8 = true;
    unsigned long long* h_pending = nullptr;   // pinned
    unsigned long long* h_dev_count = nullptr; // device scalar
    int n_uploaded = 0;                        // agents of the last upload (w.order holds their slots in list order)
    int64_t launches = 0, passes = 0, steps = 0, syncs = 0;
    bool profile = false;
    double kt[KT_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs = nullptr, eve = nullptr;
    std::string err;
};

static std::string g_err;
static int set_err(ss_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_err = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return set_err(e, SS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));   \
    } while (0)

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

extern "C" const char* ss_version(void) { return "sugarscape_b200 0.1 (sm_100a)"; }
extern "C" const char* ss_last_error(ss_engine* e) { return e ? e->err.c_str() : g_err.c_str(); }

static int create_impl(const ss_config* cfg, int32_t device, int32_t rank, int32_t world, ss_engine** out);
extern "C" int ss_create(const ss_config* cfg, int32_t device, ss_engine** out) { return create_impl(cfg, device, 0, 1, out); }
// one strip of the lattice rows per engine / GPU: rank r of `world` holds the rows [r N / world, (r + 1) N / world)
extern "C" int ss_create_strip(const ss_config* cfg, int32_t device, int32_t rank, int32_t world, ss_engine** out) {
    ss_engine* e = nullptr;
    if (!cfg || !out) return set_err(e, SS_ERR_ARG, "null argument");
    if (world < 1 || world > DR_MAX_STRIPS || rank < 0 || rank >= world) return set_err(e, SS_ERR_ARG, "bad rank / world (at most 8 strips)");
    if (world > 1) {
        if (cfg->rng_mode != SS_RNG_KEYED || !cfg->rule_move_eat || cfg->rule_pollution)
            return set_err(e, SS_ERR_UNSUPPORTED, "row strips run rule M with the keyed order, without the pollution rule");
        if (cfg->grid_size / world < 8) return set_err(e, SS_ERR_UNSUPPORTED, "strips need at least 8 rows each");
    }
    return create_impl(cfg, device, rank, world, out);
}
static int create_impl(const ss_config* cfg, int32_t device, int32_t rank, int32_t world, ss_engine** out) {
    ss_engine* e = nullptr;
    if (!cfg || !out) return set_err(e, SS_ERR_ARG, "null argument");
    if (cfg->grid_size < 3) return set_err(e, SS_ERR_ARG, "grid_size must be >= 3");
    if (cfg->grid_size > 46340) return set_err(e, SS_ERR_UNSUPPORTED, "grid_size too large (N*N must fit int32)");
    if (cfg->rule_move_eat && cfg->rule_utilicalc)
        return set_err(e, SS_ERR_ARG, "moveEat and utilicalc cannot be used together");          // sugarscape.py:186
    if (cfg->rule_utilicalc && cfg->rule_pollution)
        return set_err(e, SS_ERR_UNSUPPORTED,
                       "utilicalc+pollution raises KeyError 'duration' in the reference (agent.py:880); parity unpinned");
    if (cfg->rule_pollution && cfg->diffusion_rate < 1) return set_err(e, SS_ERR_ARG, "diffusion_rate must be >= 1");
    if (cfg->rule_seasons && cfg->season_period < 1) return set_err(e, SS_ERR_ARG, "season period must be >= 1");
    if (cfg->rng_mode != SS_RNG_MT19937 && cfg->rng_mode != SS_RNG_KEYED) return set_err(e, SS_ERR_ARG, "bad rng_mode");
    if (cfg->rule_limited_lifespan || cfg->rule_procreate) {
        if (cfg->rng_mode != SS_RNG_MT19937)
            return set_err(e, SS_ERR_UNSUPPORTED, "limitedLifespan / procreate run in the MT19937-exact mode (births change the id range the keyed order is sized by)");
        if (!cfg->rule_move_eat || cfg->rule_utilicalc)
            return set_err(e, SS_ERR_UNSUPPORTED, "limitedLifespan / procreate are built with rule M (moveEat); with utilicalc the reference needs spice");
        if (world > 1) return set_err(e, SS_ERR_UNSUPPORTED, "limitedLifespan / procreate: one GPU (serial order)");
        if (cfg->tags_length < 0 || cfg->tags_length > 30) return set_err(e, SS_ERR_ARG, "tags_length out of range [0,30]");
        if (cfg->foresight_hi < cfg->foresight_lo) return set_err(e, SS_ERR_ARG, "foresight range");
    }
    if (cfg->rule_spice) {
        if (!cfg->rule_move_eat || cfg->rule_utilicalc)
            return set_err(e, SS_ERR_UNSUPPORTED, "spice is built with rule M (moveEat); utilicalc + spice raises KeyError 'proximity' in the reference");
        if (cfg->rule_pollution) return set_err(e, SS_ERR_ARG, "pollution and spice cannot be used together");   // sugarscape.py:195-197
        if (world > 1) return set_err(e, SS_ERR_UNSUPPORTED, "spice: one GPU (serial order)");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_err(e, SS_ERR_CUDA, "no CUDA device: the engine has no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(e, SS_ERR_ARG, "bad device ordinal");
    e = new ss_engine();
    e->cfg = *cfg;
    e->device = device;
    memset(&e->w, 0, sizeof e->w);
    e->w.cfg = *cfg;
    e->w.n = cfg->grid_size;
    e->strip_rank = rank; e->strip_world = world;
    e->x0 = (int)(((long long)rank * cfg->grid_size) / world);
    e->rows = (int)(((long long)(rank + 1) * cfg->grid_size) / world) - e->x0;
    e->w.x0 = e->x0; e->w.rows = e->rows;
    e->w.dr.world = world; e->w.dr.rank = rank;
    memset(e->ipc_ptrs, 0, sizeof e->ipc_ptrs);
    e->cells = (size_t)e->rows * cfg->grid_size;
    e->iteration = cfg->iteration;
    e->env_time = cfg->env_time;
    auto bail = [&](cudaError_t ce, const char* what) {
        g_err = std::string(what) + ": " + cudaGetErrorString(ce);
        ss_destroy(e);
        return SS_ERR_CUDA;
    };
    cudaError_t ce;
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail(ce, "cudaSetDevice");
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "stream");
    {
        int lo = 0, hi = 0;             // (greatest priority = lowest number: the second stream's CTAs go first when an SM has room)
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((ce = cudaStreamCreateWithPriority(&e->stream_env, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(ce, "stream");
    }
    if ((ce = cudaEventCreateWithFlags(&e->ev_rounds, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_env, cudaEventDisableTiming)) != cudaSuccess) return bail(ce, "cudaEventCreate");
    const size_t cb = e->cells * sizeof(double);
    if ((ce = cudaMalloc(&e->w.sugar, cb)) != cudaSuccess) return bail(ce, "cudaMalloc sugar");
    if ((ce = cudaMalloc(&e->w.cap, cb)) != cudaSuccess) return bail(ce, "cudaMalloc cap");
    if ((ce = cudaMalloc(&e->w.poll, cb)) != cudaSuccess) return bail(ce, "cudaMalloc pollution");
    if ((ce = cudaMalloc(&e->w.occw, e->cells * sizeof(uint32_t))) != cudaSuccess) return bail(ce, "cudaMalloc occ");
    if ((ce = cudaMalloc(&e->w.isugar, e->cells + 16)) != cudaSuccess) return bail(ce, "cudaMalloc isugar");
    if ((ce = cudaMalloc(&e->w.mt, 624 * sizeof(uint32_t))) != cudaSuccess) return bail(ce, "cudaMalloc mt");
    if ((ce = cudaMalloc(&e->w.mti, sizeof(int32_t))) != cudaSuccess) return bail(ce, "cudaMalloc mti");
    if ((ce = cudaMalloc(&e->w.n_order, sizeof(int32_t))) != cudaSuccess) return bail(ce, "cudaMalloc n_order");
    if ((ce = cudaMalloc(&e->w.acc, sizeof(StepAccum))) != cudaSuccess) return bail(ce, "cudaMalloc acc");
    if ((ce = cudaMalloc(&e->w.hist, SS_GINI_BINS * sizeof(uint32_t))) != cudaSuccess) return bail(ce, "cudaMalloc hist");
    if ((ce = cudaMalloc(&e->w.outliers, SS_GINI_OUTLIERS * sizeof(double))) != cudaSuccess) return bail(ce, "cudaMalloc outliers");
    if (cfg->rule_spice) {
        if ((ce = cudaMalloc(&e->sw.spice, cb)) != cudaSuccess || (ce = cudaMalloc(&e->sw.spice_cap, cb)) != cudaSuccess ||
            (ce = cudaMalloc(&e->sw.pow_faults, sizeof(int32_t))) != cudaSuccess) return bail(ce, "cudaMalloc spice");
        if ((ce = cudaMemset(e->sw.spice, 0, cb)) != cudaSuccess || (ce = cudaMemset(e->sw.spice_cap, 0, cb)) != cudaSuccess ||
            (ce = cudaMemset(e->sw.pow_faults, 0, sizeof(int32_t))) != cudaSuccess) return bail(ce, "initialising device memory");
    }
    const int nbands = (cfg->grid_size + 31) / 32;
    if ((ce = cudaMalloc(&e->d_flags, sizeof(int) * (size_t)(nbands + 1))) != cudaSuccess) return bail(ce, "cudaMalloc flags");
    if ((ce = cudaMallocHost(&e->h_pending, sizeof(unsigned long long))) != cudaSuccess) return bail(ce, "cudaMallocHost");
    if ((ce = cudaMalloc(&e->h_dev_count, sizeof(unsigned long long))) != cudaSuccess) return bail(ce, "cudaMalloc count");
    int32_t mti0 = 624;
    if ((ce = cudaMemset(e->w.sugar, 0, cb)) != cudaSuccess || (ce = cudaMemset(e->w.cap, 0, cb)) != cudaSuccess ||
        (ce = cudaMemset(e->w.poll, 0, cb)) != cudaSuccess || (ce = cudaMemset(e->w.occw, 0, e->cells * sizeof(uint32_t))) != cudaSuccess ||
        (ce = cudaMemset(e->w.mt, 0, 624 * sizeof(uint32_t))) != cudaSuccess ||
        (ce = cudaMemcpy(e->w.mti, &mti0, sizeof mti0, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (ce = cudaMemset(e->w.n_order, 0, sizeof(int32_t))) != cudaSuccess || (ce = cudaMemset(e->w.acc, 0, sizeof(StepAccum))) != cudaSuccess)
        return bail(ce, "initialising device memory");
    if ((ce = cudaEventCreate(&e->ev0)) != cudaSuccess || (ce = cudaEventCreate(&e->ev1)) != cudaSuccess ||
        (ce = cudaEventCreate(&e->evs)) != cudaSuccess || (ce = cudaEventCreate(&e->eve)) != cudaSuccess)
        return bail(ce, "cudaEventCreate");
    *out = e;
    return SS_OK;
}

static void dr_free(ss_engine* e);
extern "C" void ss_destroy(ss_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->w.sugar); cudaFree(e->w.cap); cudaFree(e->w.poll); cudaFree(e->w.occw); cudaFree(e->w.isugar);
    cudaFree(e->w.agents); cudaFree(e->w.order); cudaFree(e->w.n_order); cudaFree(e->w.wealth);
    cudaFree(e->w.sortkeys); cudaFree(e->w.mt); cudaFree(e->w.mti); cudaFree(e->w.stats);
    cudaFree(e->w.pendmap);
    cudaFree(e->w.acc); cudaFree(e->w.hist); cudaFree(e->w.outliers); cudaFree(e->d_flags); cudaFree(e->d_handoff); cudaFree(e->w.log); cudaFree(e->w.log_count); cudaFree(e->w.clog); cudaFree(e->w.clog_count);
    dr_free(e);
    cudaFree(e->lw.rec); cudaFree(e->lw.id_increment); cudaFree(e->lw.events); cudaFree(e->lw.n_events); cudaFree(e->lw.status);
    cudaFree(e->sw.spice); cudaFree(e->sw.spice_cap); cudaFree(e->sw.aspice); cudaFree(e->sw.sp_endow); cudaFree(e->sw.asp_metab); cudaFree(e->sw.pow_faults);
    if (e->h_pending) cudaFreeHost(e->h_pending);
    cudaFree(e->h_dev_count);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->evs) cudaEventDestroy(e->evs);
    if (e->eve) cudaEventDestroy(e->eve);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream_env) cudaStreamDestroy(e->stream_env);
    if (e->ev_rounds) cudaEventDestroy(e->ev_rounds);
    if (e->ev_env) cudaEventDestroy(e->ev_env);
    delete e;
}

static int choose_path(ss_engine* e);
static int dr_alloc(ss_engine* e);
static int dr_tidy(ss_engine* e);
static int finish_agents(ss_engine* e, int32_t n, int32_t* d_i32, size_t nn, double* d_f64, int* d_chk);
static int life_grow(ss_engine* e, uint32_t cap);
static int life_after_launch(ss_engine* e, int* steps);

// the cell arrays are in place on the device: derive what the kernels keep beside them
static int finish_cells(ss_engine* e) {
    {   // int(sugar) byte mirror for the lattice path (built on the device); needs capacities < 256
        int flag = 0;
        CK(cudaMemsetAsync(e->d_flags, 0, sizeof(int), e->stream));
        CK(ss_launch_isugar(e->w, e->d_flags, e->stream));
        CK(cudaMemcpyAsync(&flag, e->d_flags, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->cap_fits_u8 = (flag & 1) == 0;
        e->cap_fits_u7 = flag == 0;
        e->mirror = 1;
    }
    e->have_cells = true;
    if (e->have_agents) return choose_path(e);
    return SS_OK;
}

extern "C" int ss_upload_cells(ss_engine* e, const double* sugar, const double* cap, const double* pollution) {
    if (!e || !sugar || !cap) return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    { const int rc = dr_tidy(e); if (rc) return rc; }
    const size_t cb = e->cells * sizeof(double);
    CK(cudaMemcpyAsync(e->w.sugar, sugar, cb, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->w.cap, cap, cb, cudaMemcpyHostToDevice, e->stream));
    if (pollution) CK(cudaMemcpyAsync(e->w.poll, pollution, cb, cudaMemcpyHostToDevice, e->stream));
    else CK(cudaMemsetAsync(e->w.poll, 0, cb, e->stream));
    return finish_cells(e);
}

#ifndef SS_DEFAULT_SOLO_IMPL
#define SS_DEFAULT_SOLO_IMPL 1
#endif
static int choose_path(ss_engine* e) {
    const int n = e->w.n, V = e->w.vmax;
    // (the utilicalc rule runs on the lattice through the warp-per-window kernel only: several windows needed)
    bool lattice_ok = e->cfg.rng_mode == SS_RNG_KEYED && V <= 15 && 4 * V + 1 <= n &&
                      e->cap_fits_u8 && e->w.n_slots <= SS_MAX_SLOTS;
    PassGeom g{};
    int ntile = 1, emin = 0, impl = 0;
    if (lattice_ok) {
        // window tiling for a maximum window extent; false if the lattice cannot be tiled with it
        auto tiling = [&](int wmax, PassGeom* out, int* ntile_out, int* emin_out) {
            PassGeom t{};
            t.ring = 2 * V;
            t.align = (n % 4 == 0) ? 4 : 1;
            const int nu = n / t.align;
            int em;
            if (n <= wmax) {
                t.nwin = 1; t.emax = n; em = n;
                *ntile_out = 1;
            } else {
                t.nwin = (n + wmax - 1) / wmax;
                t.emax = t.align * ((nu + t.nwin - 1) / t.nwin);
                em = t.align * (nu / t.nwin);
                // three diagonal tilings shifted by thirds of a window resolve every cell (a cell can be in
                // the ring of at most two of them: once through x, once through y) if the ring strips of
                // the three tilings are disjoint; otherwise four half-shifted tilings
                if (em >= 6 * t.ring + 3 * t.align) *ntile_out = 3;
                else if (em >= 4 * t.ring + 2 + 2 * t.align) *ntile_out = 4;
                else return false;
            }
            t.ew = (t.emax + 31) / 32;
            *emin_out = em;
            *out = t;
            return true;
        };
        const bool cta_ok = tiling(std::max(32, std::min(e->window_max, 176)), &g, &ntile, &emin) &&
                            ss_pass_smem_bytes(g.emax, g.ew, V) <= 216 * 1024;
        PassGeom gs{};
        int ntile_s = 1, emin_s = 0;
        // the solo kernel needs many windows to fill the GPU (one warp each, 32 per SM) and window coordinates < 256
        auto solo_tiling = [&](int wmax) {
            return tiling(wmax, &gs, &ntile_s, &emin_s) && gs.nwin > 1 && ss_solo_smem_bytes(gs.emax, gs.ew, e->cfg.rule_utilicalc ? 2 : 1) <= 48 * 1024;
        };
        bool solo_ok = false;
        if (e->solo_window_max > 0) {
            solo_ok = solo_tiling(std::max(32, std::min(e->solo_window_max, 176)));
        } else {
            // largest window that still leaves ~17 windows per SM; small lattices take the smallest window that tiles
            const int sizes[5] = {128, 112, 96, 80, 64};
            for (int k = 0; k < 5 && !solo_ok; k++)
                solo_ok = solo_tiling(sizes[k]) && ((long long)gs.nwin * gs.nwin >= 17ll * e->sm_count * e->shard_world || k == 4);
        }
        const bool want_solo = e->pass_impl_option >= 1 ||
                               (e->pass_impl_option < 0 && solo_ok && (long long)gs.nwin * gs.nwin >= 8ll * e->sm_count * e->shard_world);
        // 1 = the warp takes its agents one at a time, 2 = one lane per agent in batches of 32 (SS_DEFAULT_SOLO_IMPL)
        if (e->cfg.rule_utilicalc) {
            if (!solo_ok && e->solo_window_max <= 0) {      // any tiling with several windows beats the serial kernel
                const int sizes[5] = {64, 80, 96, 112, 128};
                for (int k = 0; k < 5 && !solo_ok; k++) solo_ok = solo_tiling(sizes[k]);
            }
            if (solo_ok && e->pass_impl_option != 0) { g = gs; ntile = ntile_s; emin = emin_s; impl = 1; }
            else lattice_ok = false;
        } else if (want_solo && solo_ok) { g = gs; ntile = ntile_s; emin = emin_s; impl = e->pass_impl_option >= 1 ? e->pass_impl_option : SS_DEFAULT_SOLO_IMPL; }
        else if (!cta_ok) lattice_ok = false;
    }
    // dependency rounds over the whole lattice: rule M with the keyed order, any lattice the conflict zone fits on
    // (and the utilicalc rule, whose turn reads the records of the agents on its cross: same conflict predicate; one GPU)
    const bool dr_ok = e->cfg.rng_mode == SS_RNG_KEYED && (e->cfg.rule_move_eat || (e->cfg.rule_utilicalc && e->strip_world <= 1)) &&
                       V <= 15 && 4 * V + 1 <= n && e->cap_fits_u7 && e->w.n_slots <= SS_MAX_SLOTS && e->shard_world == 1;
    if (e->strip_world > 1) {
        if (!dr_ok) return set_err(e, SS_ERR_UNSUPPORTED, "row strips need the dependency rounds (keyed order, rule M, vision <= 15, capacities < 127)");
        // a conflict zone (2 vmax rows either side) must not reach beyond the two ring neighbours
        if (n / e->strip_world < 4 * V + 1) return set_err(e, SS_ERR_UNSUPPORTED, "strips need at least 4 * max vision + 1 rows each");
        e->geom = g; e->ntile = 1; e->emin = 0; e->pass_impl = 3; e->path = PATH_LATTICE;
        return dr_alloc(e);
    }
    if (dr_ok && (e->pass_impl_option == 3 || e->pass_impl_option < 0)) { lattice_ok = true; impl = 3; }
    else if (e->pass_impl_option == 3) lattice_ok = false;
    int path;
    if (e->path_option == PATH_LATTICE) {
        if (!lattice_ok)
            return set_err(e, SS_ERR_UNSUPPORTED,
                           "lattice path needs keyed RNG, vision <= 15, capacities < 256, <= 2^24 agents and windows >= 8*vmax+2 (utilicalc: several windows)");
        path = PATH_LATTICE;
    } else if (e->path_option == PATH_SERIAL_FUSED || e->path_option == PATH_SERIAL_ENV) {
        path = n <= 256 ? PATH_SERIAL_FUSED : PATH_SERIAL_ENV;
    } else {
        path = lattice_ok ? PATH_LATTICE : (n <= 256 ? PATH_SERIAL_FUSED : PATH_SERIAL_ENV);
    }
    if (e->cfg.rule_spice) {            // the serial kernel's own environment phase grows the spice too
        if (e->path_option == PATH_LATTICE) return set_err(e, SS_ERR_UNSUPPORTED, "spice runs on the serial path");
        path = PATH_SERIAL_FUSED;
    }
    if (path != PATH_LATTICE && e->w.n_slots > (1u << 22))
        return set_err(e, SS_ERR_UNSUPPORTED, "serial path limited to 4M agents");
    e->geom = g;
    e->ntile = ntile; e->emin = emin; e->pass_impl = impl;
    e->path = path;
    if (path == PATH_LATTICE && impl == 3) { const int rc = dr_alloc(e); if (rc) return rc; }
    if (path == PATH_LATTICE && g.nwin > 1 && n % 32 == 0 && !e->w.pendmap) {
        e->w.pend_nb = n / 32;
        if (cudaMalloc(&e->w.pendmap, sizeof(uint32_t) * 2 * (size_t)e->w.pend_nb * e->w.pend_nb) != cudaSuccess) e->w.pendmap = nullptr;
    }
    return SS_OK;
}

extern "C" int ss_upload_agents(ss_engine* e, int32_t n, const int32_t* id, const int32_t* x, const int32_t* y,
                                const int32_t* vision, const int32_t* metab, const double* sugar) {
    if (!e || n < 0 || (n > 0 && (!id || !x || !y || !vision || !metab || !sugar)))
        return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    { const int rc = dr_tidy(e); if (rc) return rc; }
    // stage the six View.agents columns; argument checks, the agent store and the occupancy words are built on the device
    const size_t nn = (size_t)std::max(n, 1);
    int32_t* d_i32 = nullptr;
    double* d_f64 = nullptr;
    int* d_chk = nullptr;
    struct Staging {        // freed on every exit
        int32_t*& a; double*& b; int*& c;
        ~Staging() { cudaFree(a); cudaFree(b); cudaFree(c); }
    } staging{d_i32, d_f64, d_chk};
    CK(cudaMalloc(&d_i32, sizeof(int32_t) * 5 * nn));
    CK(cudaMalloc(&d_f64, sizeof(double) * nn));
    CK(cudaMalloc(&d_chk, sizeof(int) * 4));
    const int32_t* cols[5] = {id, x, y, vision, metab};
    for (int k = 0; k < 5 && n > 0; k++)
        CK(cudaMemcpyAsync(d_i32 + k * nn, cols[k], sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    if (n > 0) CK(cudaMemcpyAsync(d_f64, sugar, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    return finish_agents(e, n, d_i32, nn, d_f64, d_chk);
}

// the six View.agents columns are staged on the device (five int32 columns `nn` apart: id, x, y, vision, metabolism;
// sugar): argument checks, agent store, occupancy words, execution path
static int finish_agents(ss_engine* e, int32_t n, int32_t* d_i32, size_t nn, double* d_f64, int* d_chk) {
    const int N = e->w.n;
    int chk[4] = {0, 0, 1, 0};
    CK(cudaMemcpyAsync(d_chk, chk, sizeof chk, cudaMemcpyHostToDevice, e->stream));
    CK(ss_launch_validate_agents(n, N, d_i32, d_i32 + nn, d_i32 + 2 * nn, d_i32 + 3 * nn, d_i32 + 4 * nn, d_chk, e->stream));
    CK(cudaMemcpyAsync(chk, d_chk, sizeof chk, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    switch (chk[0]) {
        case 0: break;
        case 3: return set_err(e, SS_ERR_ARG, "agent ids must be >= 1");
        case 4: return set_err(e, SS_ERR_UNSUPPORTED, "vision out of range [1,31]");
        case 5: return set_err(e, SS_ERR_UNSUPPORTED, "2*vision+1 exceeds the grid size");
        case 6: return set_err(e, SS_ERR_ARG, "metabolism out of range [1,65535]");
        default: return set_err(e, SS_ERR_ARG, "agent off the grid");
    }
    if (e->strip_world > 1) {
        int bad = 0;
        CK(cudaMemsetAsync(d_chk, 0, sizeof(int), e->stream));
        CK(ss_launch_validate_rows(n, d_i32 + nn, e->x0, e->rows, d_chk, e->stream));
        CK(cudaMemcpyAsync(&bad, d_chk, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (bad) return set_err(e, SS_ERR_ARG, "agent outside this engine's strip of rows");
    }
    const int32_t max_id = std::max<int32_t>(chk[1], (int32_t)e->global_slots);
    const int vmax = std::max(std::max(chk[2], 1), e->global_vmax);
    if ((uint32_t)max_id > SS_MAX_SLOTS) return set_err(e, SS_ERR_UNSUPPORTED, "agent id exceeds 2^24");
    const uint32_t ns = (uint32_t)std::max(max_id, 1);
    const uint32_t phase = (uint32_t)(e->iteration & 1);
    cudaFree(e->w.agents); cudaFree(e->w.order); cudaFree(e->w.wealth); cudaFree(e->w.sortkeys);
    e->w.agents = nullptr; e->w.order = nullptr; e->w.wealth = nullptr; e->w.sortkeys = nullptr;
    // births take new ids: room for them up front, enlarged between launches when a run needs more
    const bool life = e->cfg.rule_limited_lifespan || e->cfg.rule_procreate;
    const uint32_t cap = life ? ns + (e->life_slack >= 0 ? (uint32_t)e->life_slack : std::max<uint32_t>(65536u, 16u * ns)) : ns;
    e->slot_capacity = cap;
    e->have_life = false;
    const int P = next_pow2((int)cap);
    CK(cudaMalloc(&e->w.agents, sizeof(AgentRec) * cap));
    CK(cudaMalloc(&e->w.order, sizeof(int32_t) * cap));
    CK(cudaMalloc(&e->w.wealth, sizeof(double) * P));
    CK(cudaMalloc(&e->w.sortkeys, sizeof(unsigned long long) * P));
    if (e->cfg.rule_spice) {
        cudaFree(e->sw.aspice); cudaFree(e->sw.sp_endow); cudaFree(e->sw.asp_metab);
        e->sw.aspice = nullptr; e->sw.sp_endow = nullptr; e->sw.asp_metab = nullptr;
        CK(cudaMalloc(&e->sw.aspice, sizeof(double) * cap));
        CK(cudaMalloc(&e->sw.sp_endow, sizeof(double) * cap));
        CK(cudaMalloc(&e->sw.asp_metab, sizeof(int32_t) * cap));
        CK(cudaMemsetAsync(e->sw.aspice, 0, sizeof(double) * cap, e->stream));
        CK(cudaMemsetAsync(e->sw.sp_endow, 0, sizeof(double) * cap, e->stream));
        CK(cudaMemsetAsync(e->sw.asp_metab, 0, sizeof(int32_t) * cap, e->stream));
        e->have_spice = false;
    }
    if (life) {
        cudaFree(e->lw.rec); e->lw.rec = nullptr;
        CK(cudaMalloc(&e->lw.rec, sizeof(LifeRec) * cap));
        CK(cudaMemsetAsync(e->lw.rec, 0, sizeof(LifeRec) * cap, e->stream));
        e->lw.capacity = (int32_t)cap;
        if (!e->lw.id_increment) {
            e->lw.ev_cap = 1 << 20;
            CK(cudaMalloc(&e->lw.id_increment, sizeof(int32_t)));
            CK(cudaMalloc(&e->lw.n_events, sizeof(int32_t)));
            CK(cudaMalloc(&e->lw.status, 2 * sizeof(int32_t)));
            CK(cudaMalloc(&e->lw.events, sizeof(int32_t) * 4 * (size_t)e->lw.ev_cap));
        }
        CK(cudaMemsetAsync(e->lw.n_events, 0, sizeof(int32_t), e->stream));
        CK(cudaMemsetAsync(e->lw.status, 0, 2 * sizeof(int32_t), e->stream));
        CK(cudaMemcpyAsync(e->lw.id_increment, &max_id, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        e->n_events = 0;
    }
    {
        CK(cudaMemsetAsync(e->w.agents, 0, sizeof(AgentRec) * cap, e->stream));
        CK(cudaMemsetAsync(e->w.occw, 0, sizeof(uint32_t) * e->cells, e->stream));
        CK(cudaMemsetAsync(e->d_flags, 0, sizeof(int), e->stream));
        CK(ss_launch_build_agents(e->w, n, phase, d_i32, d_i32 + nn, d_i32 + 2 * nn, d_i32 + 3 * nn, d_i32 + 4 * nn, d_f64,
                                  e->d_flags, e->stream));
        int err = 0;
        CK(cudaMemcpyAsync(&err, e->d_flags, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->w.n_order, &n, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        StepAccum acc;
        memset(&acc, 0, sizeof acc);
        acc.population = (unsigned long long)n;
        CK(cudaMemcpyAsync(e->w.acc, &acc, sizeof acc, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (err == 1) return set_err(e, SS_ERR_ARG, "duplicate agent id");
        if (err == 2) return set_err(e, SS_ERR_ARG, "two agents on one cell");
    }
    if (e->mirror == 2) e->mirror = 0;          // occupancy changed: the cell-word mirror is rebuilt at the next step
    e->w.n_slots = ns;
    e->w.half = ss_order_half_bits(ns);
    e->w.vmax = vmax;
    e->n_uploaded = n;
    e->have_agents = true;
    e->stepped = false;
    return choose_path(e);
}

// Agent.age / maxAge / sex / fertility / sugarEndowment / tags / alive (agent.py:134-151, 204-206) of the n agents
// of the last ss_upload_agents, same order; id_increment = Environment.idIncrement (environment.py:202-204)
extern "C" int ss_upload_agents_life(ss_engine* e, int32_t n, int32_t id_increment, const int32_t* age, const int32_t* max_age,
                                     const int32_t* sex, const int32_t* fert_lo, const int32_t* fert_hi, const double* endowment,
                                     const int32_t* tags, const int32_t* alive) {
    if (!e || n < 0 || (n > 0 && (!age || !max_age || !sex || !fert_lo || !fert_hi || !endowment || !tags)))
        return set_err(e, SS_ERR_ARG, "null argument");
    if (!(e->cfg.rule_limited_lifespan || e->cfg.rule_procreate))
        return set_err(e, SS_ERR_STATE, "the engine was created without limitedLifespan / procreate");
    if (!e->have_agents || n != e->n_uploaded) return set_err(e, SS_ERR_STATE, "ss_upload_agents_life: n differs from the uploaded agents");
    if (id_increment < (int32_t)e->w.n_slots && n > 0) return set_err(e, SS_ERR_ARG, "id_increment below the largest id");
    CK(cudaSetDevice(e->device));
    std::vector<int32_t> order((size_t)std::max(n, 1));
    if (n) CK(cudaMemcpy(order.data(), e->w.order, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    if ((uint32_t)id_increment > e->slot_capacity) {
        const int rc = life_grow(e, (uint32_t)id_increment + (e->life_slack >= 0 ? (uint32_t)e->life_slack : std::max<uint32_t>(65536u, 16u * (uint32_t)n)));
        if (rc) return rc;
    }
    std::vector<LifeRec> recs(e->slot_capacity);
    memset(recs.data(), 0, sizeof(LifeRec) * recs.size());
    for (int i = 0; i < n; i++) {
        LifeRec& r = recs[(size_t)order[i]];
        r.age = age[i]; r.max_age = max_age[i]; r.sex = sex[i]; r.fert_lo = fert_lo[i]; r.fert_hi = fert_hi[i];
        r.tags = tags[i]; r.living = alive ? (alive[i] != 0) : 1; r.endow = endowment[i];
    }
    CK(cudaMemcpyAsync(e->lw.rec, recs.data(), sizeof(LifeRec) * recs.size(), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->lw.id_increment, &id_increment, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->have_life = true;
    return SS_OK;
}

// ... the same columns of the agents in View.agents order (ss_agent_count entries; an agent that reached maxAge is still in
// the list with alive = 0 for one step), and Environment.idIncrement
extern "C" int ss_download_agents_life(ss_engine* e, int32_t* age, int32_t* max_age, int32_t* sex, int32_t* fert_lo, int32_t* fert_hi,
                                       double* endowment, int32_t* tags, int32_t* alive, int32_t* id_increment) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->have_life) return set_err(e, SS_ERR_STATE, "no life columns uploaded");
    CK(cudaSetDevice(e->device));
    int32_t cnt = 0, idinc = 0;
    CK(cudaMemcpy(&cnt, e->w.n_order, sizeof cnt, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&idinc, e->lw.id_increment, sizeof idinc, cudaMemcpyDeviceToHost));
    std::vector<int32_t> order((size_t)std::max(cnt, 1));
    std::vector<LifeRec> recs((size_t)std::max(idinc, 1));
    if (cnt) CK(cudaMemcpy(order.data(), e->w.order, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost));
    if (idinc) CK(cudaMemcpy(recs.data(), e->lw.rec, sizeof(LifeRec) * (size_t)idinc, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; i++) {
        const LifeRec& r = recs[(size_t)order[i]];
        if (age) age[i] = r.age;
        if (max_age) max_age[i] = r.max_age;
        if (sex) sex[i] = r.sex;
        if (fert_lo) fert_lo[i] = r.fert_lo;
        if (fert_hi) fert_hi[i] = r.fert_hi;
        if (endowment) endowment[i] = r.endow;
        if (tags) tags[i] = r.tags;
        if (alive) alive[i] = r.living;
    }
    if (id_increment) *id_increment = idinc;
    return SS_OK;
}

// Environment.grid slots [1] spice and [3] spice capacity (environment.py:24-34, 102-104)
extern "C" int ss_upload_cells_spice(ss_engine* e, const double* spice, const double* spice_cap) {
    if (!e || !spice || !spice_cap) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->cfg.rule_spice) return set_err(e, SS_ERR_STATE, "the engine was created without the spice rule");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(e->sw.spice, spice, e->cells * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->sw.spice_cap, spice_cap, e->cells * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
}
extern "C" int ss_download_cells_spice(ss_engine* e, double* spice, double* spice_cap) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->cfg.rule_spice) return set_err(e, SS_ERR_STATE, "the engine was created without the spice rule");
    CK(cudaSetDevice(e->device));
    if (spice) CK(cudaMemcpyAsync(spice, e->sw.spice, e->cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (spice_cap) CK(cudaMemcpyAsync(spice_cap, e->sw.spice_cap, e->cells * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
}
// Agent.spice / spiceMetabolism / spiceEndowment (= spiceCostForMating; NULL: = spice) of the n agents of the last
// ss_upload_agents, same order (agent.py:127-131, 210-216)
extern "C" int ss_upload_agents_spice(ss_engine* e, int32_t n, const int32_t* spice_metab, const double* spice, const double* endowment) {
    if (!e || n < 0 || (n > 0 && (!spice_metab || !spice))) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->cfg.rule_spice) return set_err(e, SS_ERR_STATE, "the engine was created without the spice rule");
    if (!e->have_agents || n != e->n_uploaded) return set_err(e, SS_ERR_STATE, "ss_upload_agents_spice: n differs from the uploaded agents");
    CK(cudaSetDevice(e->device));
    std::vector<int32_t> order((size_t)std::max(n, 1));
    if (n) CK(cudaMemcpy(order.data(), e->w.order, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    std::vector<double> a(e->slot_capacity, 0.0), b(e->slot_capacity, 0.0);
    std::vector<int32_t> m(e->slot_capacity, 0);
    for (int i = 0; i < n; i++) {
        if (spice_metab[i] < 1 || spice_metab[i] > 65535) return set_err(e, SS_ERR_ARG, "spice metabolism out of range [1,65535]");
        const size_t s = (size_t)order[i];
        a[s] = spice[i]; b[s] = endowment ? endowment[i] : spice[i]; m[s] = spice_metab[i];
    }
    CK(cudaMemcpy(e->sw.aspice, a.data(), sizeof(double) * a.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->sw.sp_endow, b.data(), sizeof(double) * b.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->sw.asp_metab, m.data(), sizeof(int32_t) * m.size(), cudaMemcpyHostToDevice));
    e->have_spice = true;
    return SS_OK;
}
// ... in View.agents order (ss_agent_count entries)
extern "C" int ss_download_agents_spice(ss_engine* e, int32_t* spice_metab, double* spice, double* endowment) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->have_spice) return set_err(e, SS_ERR_STATE, "no spice columns uploaded");
    CK(cudaSetDevice(e->device));
    int32_t cnt = 0;
    CK(cudaMemcpy(&cnt, e->w.n_order, sizeof cnt, cudaMemcpyDeviceToHost));
    const size_t ns = e->w.n_slots;
    std::vector<int32_t> order((size_t)std::max(cnt, 1)), m(ns);
    std::vector<double> a(ns), b(ns);
    if (cnt) CK(cudaMemcpy(order.data(), e->w.order, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(a.data(), e->sw.aspice, sizeof(double) * ns, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), e->sw.sp_endow, sizeof(double) * ns, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(m.data(), e->sw.asp_metab, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; i++) {
        const size_t s = (size_t)order[i];
        if (spice_metab) spice_metab[i] = m[s];
        if (spice) spice[i] = a[s];
        if (endowment) endowment[i] = b[s];
    }
    return SS_OK;
}

// events of the last ss_step call in execution order, 4 ints each: kind | step << 8, a, b, c -- kind 1: agent a mated with
// agent b and created agent c (sugarscape.py:551); 2: a died of starvation; 3: a died of old age (sugarscape.py:597-600)
extern "C" int ss_event_count(ss_engine* e, int32_t* count) {
    if (!e || !count) return set_err(e, SS_ERR_ARG, "null argument");
    *count = e->have_life ? e->n_events : 0;
    return SS_OK;
}
extern "C" int ss_get_events(ss_engine* e, int32_t max_events, int32_t* out) {
    if (!e || (max_events > 0 && !out)) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->have_life) return SS_OK;
    CK(cudaSetDevice(e->device));
    const int32_t k = std::min(max_events, e->n_events);
    if (k > 0) CK(cudaMemcpy(out, e->lw.events, sizeof(int32_t) * 4 * (size_t)k, cudaMemcpyDeviceToHost));
    return SS_OK;
}

extern "C" int ss_set_rng_mt19937(ss_engine* e, const uint32_t state[624], int32_t index) {
    if (!e || !state) return set_err(e, SS_ERR_ARG, "null argument");
    if (index < 0 || index > 624) return set_err(e, SS_ERR_ARG, "MT19937 index out of range");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(e->w.mt, state, 624 * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->w.mti, &index, sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
}

extern "C" int ss_get_rng_mt19937(ss_engine* e, uint32_t state[624], int32_t* index) {
    if (!e || !state || !index) return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(state, e->w.mt, 624 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(index, e->w.mti, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
}

// sugarscape.py:162 random.seed(seed): CPython's random_seed for an int -> init_by_array over the 32-bit words of abs(seed)
extern "C" int ss_seed_mt19937(ss_engine* e, uint64_t seed) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    const int klen = key[1] ? 2 : 1;
    uint32_t mt[624];
    mt[0] = 19650218u;                                                                  // init_genrand
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for (int k = 624 > klen ? 624 : klen; k; k--) {                                     // init_by_array
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (int k = 623; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    return ss_set_rng_mt19937(e, mt, 624);
}

extern "C" int ss_init_world(ss_engine* e, const ss_init_params* p) {
    if (!e || !p) return set_err(e, SS_ERR_ARG, "null argument");
    if (p->n_sites < 0 || p->n_sites > 4 || p->n_groups < 0 || p->n_groups > 4) return set_err(e, SS_ERR_ARG, "at most 4 sites / 4 agent groups");
    if (p->max_metabolism < 1 || p->max_vision < 1 || p->endowment_max < p->endowment_min || p->age_max < p->age_min ||
        p->foresight_hi < p->foresight_lo)
        return set_err(e, SS_ERR_ARG, "empty draw range in ss_init_params");
    for (int s = 0; s < 2; s++)
        if (p->childbearing[s][1] < p->childbearing[s][0] || p->childbearing[s][3] < p->childbearing[s][2])
            return set_err(e, SS_ERR_ARG, "empty fertility range in ss_init_params");
    CK(cudaSetDevice(e->device));
    { const int rc = dr_tidy(e); if (rc) return rc; }
    const int N = e->w.n;
    double D[4] = {1, 1, 1, 1};
    for (int s = 0; s < p->n_sites; s++) {                                              // environment.py:118
        if (p->site_radius[s] <= 0) return set_err(e, SS_ERR_ARG, "site radius must be positive");
        const long long a = std::max(p->site_x[s], N - p->site_x[s]), b = std::max(p->site_y[s], N - p->site_y[s]);
        D[s] = sqrt((double)(a * a + b * b)) * ((double)p->site_radius[s] / (double)N);
    }
    CK(ss_launch_init_sites(e->w, p->n_sites, p->site_x, p->site_y, D, e->cfg.max_capacity, p->grow_to_capacity, e->sm_count,
                            nullptr, nullptr, e->stream));
    if (e->cfg.rule_spice) {                                                            // sugarscape.py:1508-1513
        if (p->n_spice_sites < 0 || p->n_spice_sites > 4) return set_err(e, SS_ERR_ARG, "at most 4 spice sites");
        double Ds[4] = {1, 1, 1, 1};
        for (int s = 0; s < p->n_spice_sites; s++) {
            if (p->spice_site_radius[s] <= 0) return set_err(e, SS_ERR_ARG, "site radius must be positive");
            const long long a = std::max(p->spice_site_x[s], N - p->spice_site_x[s]), b = std::max(p->spice_site_y[s], N - p->spice_site_y[s]);
            Ds[s] = sqrt((double)(a * a + b * b)) * ((double)p->spice_site_radius[s] / (double)N);
        }
        CK(ss_launch_init_sites(e->w, p->n_spice_sites, p->spice_site_x, p->spice_site_y, Ds, e->cfg.max_capacity, p->grow_to_capacity,
                                e->sm_count, e->sw.spice_cap, e->sw.spice, e->stream));
    }
    long long total = 0;
    for (int g = 0; g < p->n_groups; g++) {
        if (p->group_count[g] < 0) return set_err(e, SS_ERR_ARG, "negative agent count");
        total += p->group_count[g];
    }
    if (total > (long long)SS_MAX_SLOTS) return set_err(e, SS_ERR_UNSUPPORTED, "agent id exceeds 2^24");
    const size_t nn = (size_t)std::max<long long>(total, 1);
    int32_t *d_i32 = nullptr, *d_scratch = nullptr;
    double* d_f64 = nullptr;
    int* d_chk = nullptr;
    struct Staging {
        int32_t*& a; double*& b; int*& c; int32_t*& d;
        ~Staging() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d); }
    } staging{d_i32, d_f64, d_chk, d_scratch};
    CK(cudaMalloc(&d_i32, sizeof(int32_t) * 5 * nn));
    CK(cudaMalloc(&d_f64, sizeof(double) * nn));
    CK(cudaMalloc(&d_chk, sizeof(int) * 4));
    const int nsuper = (N + 30) / 32 + 1;
    CK(cudaMalloc(&d_scratch, sizeof(int32_t) * ((size_t)N + nsuper + 2)));
    int32_t* d_result = d_scratch + N + nsuper;
    // with the lifespan / procreation rules the drawn maxAge, sex and fertility (and the group's tags) are kept
    const bool life = e->cfg.rule_limited_lifespan || e->cfg.rule_procreate;
    int32_t* d_life = nullptr;
    struct LifeStaging { int32_t*& p; ~LifeStaging() { cudaFree(p); } } life_staging{d_life};
    if (life) CK(cudaMalloc(&d_life, sizeof(int32_t) * 5 * nn));
    int32_t* d_spice = nullptr;                                                         // spiceMetabolism, spice endowment
    struct SpiceStaging { int32_t*& p; ~SpiceStaging() { cudaFree(p); } } spice_staging{d_spice};
    if (e->cfg.rule_spice) CK(cudaMalloc(&d_spice, sizeof(int32_t) * 2 * nn));
    CK(ss_launch_init_agents(e->w, p, d_scratch, d_scratch + N, d_i32, nn, d_f64, d_life, d_spice, d_result, e->stream));
    int32_t result[2] = {0, 0};
    CK(cudaMemcpyAsync(result, d_result, sizeof result, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    int rc = finish_cells(e);
    if (rc) return rc;
    std::vector<int32_t> hl;
    std::vector<double> hs;
    if (life) {             // (before finish_agents frees nothing of these, but keep the order simple: copy out first)
        hl.resize(5 * nn); hs.resize(nn);
        CK(cudaMemcpy(hl.data(), d_life, sizeof(int32_t) * 5 * nn, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs.data(), d_f64, sizeof(double) * nn, cudaMemcpyDeviceToHost));
    }
    std::vector<int32_t> hsp;
    if (e->cfg.rule_spice) {
        hsp.resize(2 * nn);
        CK(cudaMemcpy(hsp.data(), d_spice, sizeof(int32_t) * 2 * nn, cudaMemcpyDeviceToHost));
    }
    rc = finish_agents(e, result[0], d_i32, nn, d_f64, d_chk);
    if (rc) return rc;
    if (e->cfg.rule_spice) {
        std::vector<double> sp((size_t)std::max(result[0], 1));
        for (int i = 0; i < result[0]; i++) sp[(size_t)i] = (double)hsp[nn + (size_t)i];
        rc = ss_upload_agents_spice(e, result[0], hsp.data(), sp.data(), nullptr);
        if (rc) return rc;
    }
    if (!life) return rc;
    const int32_t n = result[0];
    std::vector<int32_t> age((size_t)std::max(n, 1), 0);                               // Agent.setAge: age = 0 (agent.py:236-238)
    return ss_upload_agents_life(e, n, result[1], age.data(), hl.data(), hl.data() + nn, hl.data() + 2 * nn, hl.data() + 3 * nn,
                                 hs.data(), hl.data() + 4 * nn, nullptr);              // endowment = the drawn initial sugar
}

extern "C" int ss_set_rng_keyed(ss_engine* e, uint64_t seed) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (e->cfg.rng_mode != SS_RNG_KEYED) return set_err(e, SS_ERR_STATE, "engine was created with rng_mode MT19937");
    e->cfg.keyed_seed = seed;
    e->w.cfg.keyed_seed = seed;
    return SS_OK;
}

struct KTimer {
    ss_engine* e; int k;
    KTimer(ss_engine* e_, int k_) : e(e_), k(k_) { if (e->profile) cudaEventRecord(e->ev0, e->stream); }
    ~KTimer() {
        if (!e->profile) return;
        cudaEventRecord(e->ev1, e->stream);
        cudaEventSynchronize(e->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e->ev0, e->ev1);
        e->kt[k] += ms;
    }
};

static int ensure_stats(ss_engine* e, int nsteps) {
    if (nsteps <= e->stats_cap) return SS_OK;
    cudaFree(e->w.stats);
    e->w.stats = nullptr;
    CK(cudaMalloc(&e->w.stats, sizeof(ss_step_stats) * (size_t)nsteps));
    e->stats_cap = nsteps;
    return SS_OK;
}

static void dr_free(ss_engine* e) {
    DrState& d = e->w.dr;
    cudaFree(d.self.mw); cudaFree(d.self.mwx); cudaFree(d.self.dep); cudaFree(d.self.queue); cudaFree(d.self.ctl);
    cudaFree(d.succ); cudaFree(d.pool); cudaFree(d.alive); cudaFree(d.atrisk); cudaFree(d.q1flag); cudaFree(d.q1wait);
    cudaFree(d.self.delta); cudaFree(d.home); cudaFree(d.acc_g); cudaFree(d.hist_g);
    if (e->peers_ipc)
        for (int r = 0; r < DR_MAX_STRIPS; r++)
            for (int k = 0; k < 16; k++) if (e->ipc_ptrs[r][k]) cudaIpcCloseMemHandle(e->ipc_ptrs[r][k]);
    memset(e->ipc_ptrs, 0, sizeof e->ipc_ptrs);
    const int rank = d.rank, world = d.world;
    memset(&d, 0, sizeof d);
    d.rank = rank; d.world = world;
    e->dr_slots = 0; e->peers_connected = 0;
}

// device state of the dependency rounds (ss_world.cuh DrState), sized by the lattice and the slots
static int dr_alloc(ss_engine* e) {
    DrState& d = e->w.dr;
    const uint32_t ns = std::max<uint32_t>(e->w.n_slots, 1u);
    if (!d.self.mw) {
        CK(cudaMalloc(&d.self.mw, e->cells + 64));
        CK(cudaMalloc(&d.self.mwx, e->cells + 64));
        CK(cudaMalloc(&d.self.ctl, DR_CTL_WORDS * sizeof(unsigned int)));
        CK(cudaMemsetAsync(d.self.mw, 0, e->cells + 64, e->stream));
        CK(cudaMemsetAsync(d.self.mwx, 0, e->cells + 64, e->stream));
        CK(cudaMemsetAsync(d.self.ctl, 0, DR_CTL_WORDS * sizeof(unsigned int), e->stream));
    }
    // successors beyond the 14 of a record: a run of the pool; sized for every cell of every conflict zone (the
    // worst case) up to 2^26 entries, beyond that for 16 more per agent
    const int V = std::max(e->w.vmax, 1);
    const uint64_t zone = (uint64_t)(2 * V + 1) * (2 * V + 1) + 4 * V;
    const uint64_t pool64 = std::min<uint64_t>((uint64_t)ns * zone, std::max<uint64_t>((uint64_t)ns * 16u, 1ull << 26)) + 64;
    const uint32_t pool_cap = (uint32_t)std::min<uint64_t>(pool64, 0xfffffff0ull);
    if (ns != e->dr_slots || pool_cap != d.pool_cap) {
        cudaFree(d.self.dep); cudaFree(d.self.queue); cudaFree(d.succ); cudaFree(d.pool); cudaFree(d.alive); cudaFree(d.atrisk); cudaFree(d.q1flag); cudaFree(d.q1wait);
        cudaFree(d.self.delta); cudaFree(d.home); cudaFree(d.acc_g); cudaFree(d.hist_g);
        d.self.delta = nullptr; d.home = nullptr; d.acc_g = nullptr; d.hist_g = nullptr;
        d.q1flag = nullptr;
        d.self.dep = nullptr; d.self.queue = nullptr; d.succ = nullptr; d.pool = nullptr; d.alive = nullptr; d.atrisk = nullptr; d.q1wait = nullptr;
        const size_t words = (size_t)(ns + 31) / 32 + 1;
        CK(cudaMalloc(&d.self.dep, sizeof(uint32_t) * ((size_t)ns / 2 + 2)));
        CK(cudaMalloc(&d.self.queue, sizeof(uint32_t) * ((size_t)ns + 64)));
        CK(cudaMalloc(&d.succ, sizeof(uint4) * 4 * (size_t)ns));
        CK(cudaMalloc(&d.pool, sizeof(uint32_t) * (size_t)pool_cap));
        CK(cudaMalloc(&d.alive, sizeof(uint32_t) * words));
        CK(cudaMalloc(&d.atrisk, sizeof(uint32_t) * words));
        CK(cudaMalloc(&d.q1flag, sizeof(uint32_t) * words));
        CK(cudaMemsetAsync(d.q1flag, 0, sizeof(uint32_t) * words, e->stream));
        CK(cudaMalloc(&d.q1wait, sizeof(uint2) * (size_t)ns));
        CK(cudaMemsetAsync(d.self.dep, 0, sizeof(uint32_t) * ((size_t)ns / 2 + 2), e->stream));
        CK(cudaMemsetAsync(d.alive, 0, sizeof(uint32_t) * words, e->stream));
        CK(cudaMemsetAsync(d.atrisk, 0, sizeof(uint32_t) * words, e->stream));
        CK(cudaMemsetAsync(d.q1wait, 0, sizeof(uint2) * (size_t)ns, e->stream));
        if (e->strip_world > 1) {               // the replicated directory, this strip's delta list, the merged statistics
            d.delta_cap = ns + 64;
            CK(cudaMalloc(&d.self.delta, sizeof(uint32_t) * (size_t)d.delta_cap));
            CK(cudaMalloc(&d.home, (size_t)ns + 16));
            CK(cudaMalloc(&d.acc_g, sizeof(StepAccum)));
            CK(cudaMalloc(&d.hist_g, SS_GINI_BINS * sizeof(uint32_t)));
            CK(cudaMemsetAsync(d.home, 0, (size_t)ns + 16, e->stream));
            CK(cudaMemsetAsync(d.acc_g, 0, sizeof(StepAccum), e->stream));
        }
        e->dr_slots = ns; d.pool_cap = pool_cap;
        e->peers_connected = 0;
        if (e->mirror == 2) e->mirror = 0;
    }
    {   // the predecessor counters take one atomic per edge of the step's DAG: keep them in L2 (access policy window on
        // the engine's stream; the streaming traffic of a step would otherwise evict them between two touches)
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device);
        const size_t bytes = sizeof(uint32_t) * ((size_t)ns / 2 + 2);
        if (e->l2_persist && max_win > 0 && max_persist > 0) {
            const size_t win = std::min<size_t>(bytes, (size_t)max_win);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)max_persist, win));
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = d.self.dep;
            av.accessPolicyWindow.num_bytes = win;
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)win);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            const cudaError_t pe = cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            if (pe != cudaSuccess) cudaGetLastError();
            if (e->trace) {
                size_t lim = 0;
                cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
                fprintf(stderr, "[ss] L2 window: max window %d, max persisting %d, limit now %zu, window %zu bytes, hit ratio %.3f, rc %d\n",
                        max_win, max_persist, lim, win, av.accessPolicyWindow.hitRatio, (int)pe);
            }
        }
    }
    d.queue_cap = ns;
    d.self.occw = e->w.occw; d.self.sugar = e->w.sugar; d.self.poll = e->w.poll; d.self.agents = e->w.agents;
    d.self.acc = reinterpret_cast<unsigned long long*>(e->w.acc); d.self.hist = e->w.hist;
    d.self.x0 = e->x0; d.self.rows = e->rows;
    d.rank = e->strip_rank; d.world = e->strip_world;
    if (!e->peers_connected) { d.lo = d.self; d.hi = d.self; }
    d.all[e->strip_rank] = d.self;
    if (e->peers_connected) {                   // (the neighbours' views stay; this strip's own view is refreshed)
        if (e->strip_world == 1) { d.lo = d.self; d.hi = d.self; }
    }
    e->tma_ok = 0;
    if (e->use_tma) CK(ss_dr_encode_tmap(e->w, e->tmap, &e->tma_ok));
    return SS_OK;
}

static int env_phase(ss_engine* e, cudaStream_t stream = nullptr) {
    if (!stream) stream = e->stream;
    {
        KTimer t(e, KT_GROW);
        // the growback refreshes the int(sugar) mirror that is current (cell words of the rounds / isugar bytes)
        if (e->cfg.rule_grow) {
            if (e->mirror == 2)                                                     // + both layouts of the cell bytes
                CK(ss_launch_dr_grow(e->w, e->iteration, stream == e->stream ? 0 : e->grow_ctas_per_sm * e->sm_count, stream));
            else CK(ss_launch_grow(e->w, e->iteration, e->sm_count, 0, stream));
            e->launches++;
        }
    }
    if (e->cfg.rule_pollution && e->iteration >= e->cfg.diffusion_start_time &&
        e->iteration % e->cfg.diffusion_rate == 0) {                                  // sugarscape.py:661-663
        KTimer t(e, KT_DIFFUSE);
        if (!e->d_handoff) CK(cudaMalloc(&e->d_handoff, ss_diffuse_handoff_bytes(e->w.n)));
        CK(ss_launch_diffuse(e->w, e->d_flags, e->d_handoff, e->diffuse_impl, stream));
        e->launches++;
    }
    return SS_OK;
}

// the main stream waits for what the second stream still runs (end of a stepping call; before anything else touches the state)
static int env_join(ss_engine* e) {
    if (!e->env_in_flight) return SS_OK;
    CK(cudaStreamWaitEvent(e->stream, e->ev_env, 0));
    e->env_in_flight = false;
    return SS_OK;
}

static void pass_geometry(ss_engine* e, int p, PassGeom* out) {
    PassGeom g = e->geom;
    if (e->ntile == 3) {                                  // diagonal shifts by thirds of a window
        const int t = p % 3;
        g.shift_x = g.shift_y = g.align * ((t * e->emin / 3) / g.align);
    } else if (e->ntile == 4) {                           // (0,0) (h,h) (0,h) (h,0)
        const int h = g.align * ((e->emin / 2) / g.align);
        g.shift_x = (p & 1) ? h : 0;
        g.shift_y = ((p & 3) == 1 || (p & 3) == 2) ? h : 0;
    } else {
        g.shift_x = g.shift_y = 0;
    }
    g.bx_lo = 0; g.bx_cnt = g.nwin;
    g.map_read = g.map_write = -1;
    g.ring_per_agent = e->w.log != nullptr;
    *out = g;
}

// the rounds leave harvested cells' f64 sugar and vacated cells' occupancy words to be tidied lazily (ss_rounds.cu):
// do it before anything but the rounds reads them
static int dr_tidy(ss_engine* e) {
    if (!e->dr_dirty) return SS_OK;
    CK(ss_launch_dr_canon(e->w, e->stream));
    e->launches++;
    e->dr_dirty = false;
    return SS_OK;
}

// one timestep as dependency rounds over the whole lattice: no window passes, no host round trips
static int dr_step(ss_engine* e, int stats_index) {
    e->dr_dirty = true;
    if (e->mirror != 2) {           // after an upload or a change of path: cell words, alive / at-risk bitmaps, OCC_RISK flags
        int flag = 0;
        CK(cudaMemsetAsync(e->d_flags, 0, sizeof(int), e->stream));
        CK(ss_launch_dr_mirror(e->w, e->d_flags, e->stream));
        CK(cudaMemcpyAsync(&flag, e->d_flags, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->launches += 3; e->syncs++;
        if (flag) return set_err(e, SS_ERR_UNSUPPORTED, "dependency rounds need cell sugar and capacities < 127");
        e->mirror = 2;
    }
    // the conflict graph of this step reads occupancy, priorities and the alive / at-risk bitmaps -- nothing the previous
    // step's statistics, growback or diffusion touch -- so it runs beside them; the rounds need all of it
    const bool overlap = e->overlap_option && !e->profile && !e->trace;
    // work-list mode of the rounds kernel (one GPU): entries are taken as they are published, so the queue starts empty
    const int async_q = e->async_queue && e->strip_world <= 1;
    if (async_q) { CK(cudaMemsetAsync(e->w.dr.self.queue, 0, sizeof(uint32_t) * ((size_t)e->dr_slots + 64), e->stream)); e->launches++; }
    {
        KTimer t(e, KT_PREPARE);
        if (e->trace) { CK(cudaStreamSynchronize(e->stream)); fprintf(stderr, "[ss] step %lld: build\n", (long long)e->iteration); }
        CK(ss_launch_dr_build(e->w, e->iteration, e->tmap, e->tma_ok && e->use_tma, e->stream));
        { const int rc = env_join(e); if (rc) return rc; }
        CK(cudaMemsetAsync(e->w.hist, 0, SS_GINI_BINS * sizeof(uint32_t), e->stream));
        CK(ss_launch_prepare(e->w, e->iteration, 0, e->stream));
        e->launches += 5;
        if (e->trace) {
            CK(cudaStreamSynchronize(e->stream));
            unsigned ctl[8];
            CK(cudaMemcpy(ctl, e->w.dr.self.ctl, sizeof ctl, cudaMemcpyDeviceToHost));
            fprintf(stderr, "[ss] step %lld: built, ready %u, pool %u, overflow %u; rounds\n", (long long)e->iteration, ctl[0], ctl[6], ctl[7]);
        }
    }
    {
        KTimer t(e, KT_MOVE);
        CK(ss_launch_dr_rounds(e->w, e->iteration, e->env_time, e->sm_count, 0, 0u, 0u, async_q, e->stream));
        if (e->trace) {
            CK(cudaStreamSynchronize(e->stream));
            unsigned ctl[8];
            CK(cudaMemcpy(ctl, e->w.dr.self.ctl, sizeof ctl, cudaMemcpyDeviceToHost));
            fprintf(stderr, "[ss] step %lld: rounds done, queue tail %u, barriers %u\n", (long long)e->iteration, ctl[0], ctl[2]);
        }
        CK(ss_launch_dr_check(e->w, e->stream));
        e->launches += 2; e->passes++;
    }
    cudaStream_t env_stream = e->stream;
    if (overlap) {
        env_stream = e->stream_env;
        CK(cudaEventRecord(e->ev_rounds, e->stream));
        CK(cudaStreamWaitEvent(env_stream, e->ev_rounds, 0));
    }
    {
        KTimer t(e, KT_STATS);
        CK(ss_launch_stats(e->w, stats_index, env_stream));
        e->launches++;
    }
    const int rc = env_phase(e, env_stream);
    if (rc) return rc;
    if (overlap) {
        CK(cudaEventRecord(e->ev_env, env_stream));
        e->env_in_flight = true;
    }
    return SS_OK;
}

static int lattice_step(ss_engine* e, int stats_index) {
    if (e->pass_impl == 3) return dr_step(e, stats_index);
    if (e->mirror != 1) {           // the window passes read the isugar byte mirror
        { const int rc = dr_tidy(e); if (rc) return rc; }
        CK(ss_launch_isugar(e->w, e->d_flags, e->stream));
        e->launches++;
        e->mirror = 1;
    }
    {
        KTimer t(e, KT_PREPARE);
        CK(cudaMemsetAsync(e->w.hist, 0, SS_GINI_BINS * sizeof(uint32_t), e->stream));
        CK(ss_launch_prepare(e->w, e->iteration, 1, e->stream));
        e->launches += e->cfg.rule_can_starve ? 2 : 1;
    }
    {
        KTimer t(e, KT_MOVE);
        const int first_check = e->geom.nwin == 1 ? 1 : std::max(2, e->last_passes - 1);
        int p = 0;
        for (;;) {
            PassGeom g;
            pass_geometry(e, p, &g);
            if (e->w.pendmap && e->use_pendmap) {
                g.map_write = p & 1;
                g.map_read = p > 0 ? ((p - 1) & 1) : -1;
                CK(cudaMemsetAsync(e->w.pendmap + (size_t)g.map_write * e->w.pend_nb * e->w.pend_nb, 0,
                                   sizeof(uint32_t) * (size_t)e->w.pend_nb * e->w.pend_nb, e->stream));
            }
            if (e->pass_impl >= 1) CK(ss_launch_move_solo(e->w, g, e->iteration, e->env_time, e->pass_impl == 2, e->stream));
            else CK(ss_launch_move_pass(e->w, g, e->iteration, e->env_time, e->stream));
            e->launches++; e->passes++; p++;
            if (p >= first_check || e->trace) {
                CK(cudaMemcpyAsync(e->h_pending, &e->w.acc->pending, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   e->stream));
                CK(cudaStreamSynchronize(e->stream));
                e->syncs++;
                if (e->trace) {
                    static double t_prev = 0;
                    struct timespec ts_; clock_gettime(CLOCK_MONOTONIC, &ts_);
                    const double now = ts_.tv_sec * 1e3 + ts_.tv_nsec * 1e-6;
                    fprintf(stderr, "[ss] step %lld pass %d pending %llu  (+%.3f ms)\n", (long long)e->iteration, p, *e->h_pending, now - t_prev);
                    t_prev = now;
                }
                if (*e->h_pending == 0ull) break;
            }
            if (p > 4096) return set_err(e, SS_ERR_STATE, "move passes did not converge");
        }
        e->last_passes = p;
    }
    {
        KTimer t(e, KT_SETTLE);
        CK(ss_launch_settle(e->w, e->iteration, e->env_time, e->stream));
        e->launches++;
    }
    {
        KTimer t(e, KT_STATS);
        CK(ss_launch_stats(e->w, stats_index, e->stream));
        e->launches++;
    }
    return env_phase(e);
}

// enlarge the agent store, the list, the scratch arrays and the life records (births take new ids)
static int life_grow(ss_engine* e, uint32_t cap) {
    const uint32_t old = e->slot_capacity;
    if (cap <= old) return SS_OK;
    if (cap > (1u << 22)) return set_err(e, SS_ERR_UNSUPPORTED, "serial path limited to 4M agent ids");
    const int P = next_pow2((int)cap);
    const bool spice = e->cfg.rule_spice != 0;
    struct Fresh {          // the new arrays; whatever is still here when the function leaves (an error) is freed
        AgentRec* agents = nullptr; int32_t* order = nullptr; double* wealth = nullptr; unsigned long long* sortkeys = nullptr;
        LifeRec* rec = nullptr; double* aspice = nullptr; double* sp_endow = nullptr; int32_t* asp_metab = nullptr;
        ~Fresh() { cudaFree(agents); cudaFree(order); cudaFree(wealth); cudaFree(sortkeys); cudaFree(rec); cudaFree(aspice); cudaFree(sp_endow); cudaFree(asp_metab); }
    } f;
    CK(cudaMalloc(&f.agents, sizeof(AgentRec) * cap));
    CK(cudaMalloc(&f.order, sizeof(int32_t) * cap));
    CK(cudaMalloc(&f.wealth, sizeof(double) * P));
    CK(cudaMalloc(&f.sortkeys, sizeof(unsigned long long) * P));
    CK(cudaMalloc(&f.rec, sizeof(LifeRec) * cap));
    CK(cudaMemsetAsync(f.agents, 0, sizeof(AgentRec) * cap, e->stream));
    CK(cudaMemsetAsync(f.rec, 0, sizeof(LifeRec) * cap, e->stream));
    CK(cudaMemcpyAsync(f.agents, e->w.agents, sizeof(AgentRec) * old, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(f.order, e->w.order, sizeof(int32_t) * old, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(f.rec, e->lw.rec, sizeof(LifeRec) * old, cudaMemcpyDeviceToDevice, e->stream));
    if (spice) {
        CK(cudaMalloc(&f.aspice, sizeof(double) * cap));
        CK(cudaMalloc(&f.sp_endow, sizeof(double) * cap));
        CK(cudaMalloc(&f.asp_metab, sizeof(int32_t) * cap));
        CK(cudaMemsetAsync(f.aspice, 0, sizeof(double) * cap, e->stream));
        CK(cudaMemsetAsync(f.sp_endow, 0, sizeof(double) * cap, e->stream));
        CK(cudaMemsetAsync(f.asp_metab, 0, sizeof(int32_t) * cap, e->stream));
        CK(cudaMemcpyAsync(f.aspice, e->sw.aspice, sizeof(double) * old, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(f.sp_endow, e->sw.sp_endow, sizeof(double) * old, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(f.asp_metab, e->sw.asp_metab, sizeof(int32_t) * old, cudaMemcpyDeviceToDevice, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    std::swap(e->w.agents, f.agents); std::swap(e->w.order, f.order); std::swap(e->w.wealth, f.wealth);
    std::swap(e->w.sortkeys, f.sortkeys); std::swap(e->lw.rec, f.rec);          // (the old arrays leave with `f`)
    if (spice) { std::swap(e->sw.aspice, f.aspice); std::swap(e->sw.sp_endow, f.sp_endow); std::swap(e->sw.asp_metab, f.asp_metab); }
    e->slot_capacity = cap; e->lw.capacity = (int32_t)cap;
    return SS_OK;
}
// after a launch of the serial kernel with the life rules: how many of the *steps it completed (it stops before a step the
// store or the event buffer may not hold -- then they are enlarged), the ids handed out so far, faults
static int life_after_launch(ss_engine* e, int* steps) {
    int32_t st[2] = {0, 0}, idinc = 0, nev = 0;
    CK(cudaMemcpyAsync(st, e->lw.status, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&idinc, e->lw.id_increment, sizeof idinc, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&nev, e->lw.n_events, sizeof nev, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    if (st[1]) return set_err(e, SS_ERR_STATE, "a birth found the agent store or the event buffer full (state unusable)");
    e->w.n_slots = (uint32_t)std::max<int32_t>(idinc, 1);
    e->n_events = nev;
    if (st[0] < *steps) {           // stopped early: more room, the caller launches the rest
        int32_t n_order = 0;
        CK(cudaMemcpy(&n_order, e->w.n_order, sizeof n_order, cudaMemcpyDeviceToHost));
        if ((long long)nev + 10ll * n_order + 64 > e->lw.ev_cap) {
            if (st[0] == 0 && nev == 0) return set_err(e, SS_ERR_UNSUPPORTED, "population too large for the event buffer");
            if (st[0] == 0) return set_err(e, SS_ERR_STATE, "event buffer full: call ss_step with fewer steps at a time and read ss_get_events in between");
        }
        const uint32_t need = (uint32_t)idinc + 8u * (uint32_t)n_order + 64u;       // what the kernel asks for before a step
        const int rc = life_grow(e, e->life_slack >= 0 ? need + (uint32_t)e->life_slack
                                                       : std::max<uint32_t>(2u * e->slot_capacity, need + 65536u));
        if (rc) return rc;
    }
    *steps = st[0];
    return SS_OK;
}

extern "C" int ss_step(ss_engine* e, int32_t nsteps, ss_step_stats* out) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (nsteps < 0) return set_err(e, SS_ERR_ARG, "nsteps < 0");
    if (!e->have_cells || !e->have_agents) return set_err(e, SS_ERR_STATE, "upload cells and agents before stepping");
    if (e->path < 0) return set_err(e, SS_ERR_STATE, "no execution path (upload_agents failed?)");
    if (e->strip_world > 1) return set_err(e, SS_ERR_STATE, "a strip engine steps with ss_strip_step / ss_strip_phase");
    if (nsteps == 0) return SS_OK;
    CK(cudaSetDevice(e->device));
    int rc = ensure_stats(e, nsteps);
    if (rc) return rc;
    if (e->path != PATH_LATTICE) { rc = dr_tidy(e); if (rc) return rc; if (e->mirror == 2) e->mirror = 0; }
    CK(cudaEventRecord(e->evs, e->stream));
    const bool life = e->cfg.rule_limited_lifespan || e->cfg.rule_procreate;
    if (life) {
        if (!e->have_life) return set_err(e, SS_ERR_STATE, "limitedLifespan / procreate: ss_upload_agents_life after ss_upload_agents");
        CK(cudaMemsetAsync(e->lw.n_events, 0, sizeof(int32_t), e->stream));
    }
    if (e->cfg.rule_spice && !e->have_spice) return set_err(e, SS_ERR_STATE, "spice: ss_upload_cells_spice and ss_upload_agents_spice before stepping");
    const LifeWorld no_life{};
    if (e->path == PATH_SERIAL_FUSED) {
        KTimer t(e, KT_SERIAL);
        int done = 0;
        while (done < nsteps) {
            CK(ss_launch_serial(e->w, life ? e->lw : no_life, e->sw, e->iteration, e->env_time, nsteps - done, 1, done, e->stream));
            e->launches++;
            int k = nsteps - done;
            if (life) { rc = life_after_launch(e, &k); if (rc) { e->path = -1; return rc; } }
            e->iteration += k; e->env_time += k; e->steps += k;
            done += k;
        }
    } else {
        for (int t = 0; t < nsteps; t++) {
            if (e->path == PATH_SERIAL_ENV) {
                {
                    KTimer tm(e, KT_SERIAL);
                    for (int k = 0; k == 0;) {
                        CK(ss_launch_serial(e->w, life ? e->lw : no_life, e->sw, e->iteration, e->env_time, 1, 0, t, e->stream));
                        e->launches++;
                        k = 1;
                        if (life) { rc = life_after_launch(e, &k); if (rc) { e->path = -1; return rc; } }
                    }
                }
                rc = env_phase(e);
            } else {
                rc = lattice_step(e, t);
            }
            if (rc) { e->path = -1; return rc; }       // the state is mid-step: unusable until the next upload (SS_ERR_STATE)
            e->iteration++; e->env_time++; e->steps++;
        }
    }
    e->stepped = true;
    if (e->cfg.rule_spice) {            // an operand the restated pow has no answer for would have made the welfare NaN: never silently
        int32_t pf = 0;
        CK(cudaMemcpyAsync(&pf, e->sw.pow_faults, sizeof pf, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        if (pf) { e->path = -1; return set_err(e, SS_ERR_STATE, "welfare: a pow operand outside the restated libm path (state unusable)"); }
    }
    { const int rcj = env_join(e); if (rcj) return rcj; }
    CK(cudaEventRecord(e->eve, e->stream));
    if (out) CK(cudaMemcpyAsync(out, e->w.stats, sizeof(ss_step_stats) * (size_t)nsteps, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, e->evs, e->eve);      // device time of this call, events on the engine's stream
        e->kt[KT_LAST_CALL] = ms;
    }
    return SS_OK;
}

static int fetch_population(ss_engine* e, std::vector<AgentRec>* recs_out, std::vector<int32_t>* order_out) {
    // living agents in the reference's list order (see header)
    const uint32_t ns = e->w.n_slots;
    std::vector<AgentRec> recs(ns);
    CK(cudaMemcpyAsync(recs.data(), e->w.agents, sizeof(AgentRec) * ns, cudaMemcpyDeviceToHost, e->stream));
    std::vector<int32_t> order;
    if (e->path != PATH_LATTICE) {
        int32_t cnt = 0;
        CK(cudaMemcpyAsync(&cnt, e->w.n_order, sizeof cnt, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        order.resize((size_t)cnt);
        if (cnt) CK(cudaMemcpy(order.data(), e->w.order, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost));
    } else {
        CK(cudaStreamSynchronize(e->stream));
        if (!e->stepped) {
            std::vector<int32_t> up((size_t)e->n_uploaded);            // list order of the upload
            if (e->n_uploaded) CK(cudaMemcpy(up.data(), e->w.order, sizeof(int32_t) * up.size(), cudaMemcpyDeviceToHost));
            for (int32_t s : up) if (attr_alive(recs[s].attr)) order.push_back(s);
        } else {
            const OrderKey ok = ss_order_key(ss_step_key(e->cfg.keyed_seed, e->iteration - 1), e->w.half);
            std::vector<uint64_t> keyed, tmp;
            keyed.reserve(ns);
            for (uint32_t s = 0; s < ns; s++)
                if (attr_alive(recs[s].attr)) keyed.push_back(((uint64_t)ss_priority(ok, s) << 32) | s);
            tmp.resize(keyed.size());
            for (int pass = 0; pass < 4; pass++) {                    // LSD radix sort on the 32-bit priority
                const int shift = 32 + 8 * pass;
                size_t cnt[257] = {0};
                for (uint64_t k : keyed) cnt[((k >> shift) & 0xff) + 1]++;
                for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
                for (uint64_t k : keyed) tmp[cnt[(k >> shift) & 0xff]++] = k;
                keyed.swap(tmp);
            }
            order.reserve(keyed.size());
            for (uint64_t k : keyed) order.push_back((int32_t)(k & 0xffffffffu));
        }
    }
    recs_out->swap(recs);
    order_out->swap(order);
    return SS_OK;
}

// strips: the agents standing on this strip (the population counter of the statistics covers all strips)
static unsigned long long strip_population(ss_engine* e) {
    unsigned long long cnt = 0;
    cudaMemsetAsync(e->h_dev_count, 0, sizeof(unsigned long long), e->stream);
    ss_launch_count_alive(e->w, e->h_dev_count, e->stream);
    cudaMemcpyAsync(&cnt, e->h_dev_count, sizeof cnt, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    return cnt;
}
extern "C" int ss_agent_count(ss_engine* e) {
    if (!e || !e->have_agents) return 0;
    cudaSetDevice(e->device);
    if (e->path != PATH_LATTICE) {
        int32_t cnt = 0;
        cudaMemcpy(&cnt, e->w.n_order, sizeof cnt, cudaMemcpyDeviceToHost);
        return cnt;
    }
    if (e->strip_world > 1) return (int)strip_population(e);
    StepAccum acc;
    cudaMemcpy(&acc, e->w.acc, sizeof acc, cudaMemcpyDeviceToHost);
    return (int)acc.population;
}

extern "C" int ss_download_cells(ss_engine* e, double* sugar, double* cap, double* pollution, int32_t* occ_id) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    { const int rc = dr_tidy(e); if (rc) return rc; }
    const size_t cb = e->cells * sizeof(double);
    if (sugar) CK(cudaMemcpyAsync(sugar, e->w.sugar, cb, cudaMemcpyDeviceToHost, e->stream));
    if (cap) CK(cudaMemcpyAsync(cap, e->w.cap, cb, cudaMemcpyDeviceToHost, e->stream));
    if (pollution) CK(cudaMemcpyAsync(pollution, e->w.poll, cb, cudaMemcpyDeviceToHost, e->stream));
    int32_t* d_ids = nullptr;
    struct Guard { int32_t*& p; ~Guard() { cudaFree(p); } } guard{d_ids};       // freed on every exit
    if (occ_id) {
        CK(cudaMalloc(&d_ids, e->cells * sizeof(int32_t)));
        CK(ss_launch_occ_ids(e->w, d_ids, e->stream));
        CK(cudaMemcpyAsync(occ_id, d_ids, e->cells * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SS_OK;
}

extern "C" int ss_download_agents(ss_engine* e, int32_t* id, int32_t* x, int32_t* y, int32_t* vision, int32_t* metab,
                                  double* sugar) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (!e->have_agents) return SS_OK;
    CK(cudaSetDevice(e->device));
    if (e->path == PATH_LATTICE && e->stepped) {
        // keyed order after a step: compaction of the inverse priority map on the device, columns copied straight out
        const uint32_t domain = 1u << (2 * e->w.half);
        const uint32_t nb = ss_export_blocks(domain);
        StepAccum acc;
        CK(cudaMemcpyAsync(&acc, e->w.acc, sizeof acc, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        const size_t pop = e->strip_world > 1 ? (size_t)strip_population(e) : (size_t)acc.population, stride = std::max<size_t>(pop, 1);
        uint32_t* d_cnt = nullptr;
        int32_t* d_cols = nullptr;
        double* d_sugar = nullptr;
        cudaError_t ce;
        if ((ce = cudaMalloc(&d_cnt, sizeof(uint32_t) * ((size_t)nb + 1))) != cudaSuccess ||
            (ce = cudaMalloc(&d_cols, sizeof(int32_t) * 5 * stride)) != cudaSuccess ||
            (ce = cudaMalloc(&d_sugar, sizeof(double) * stride)) != cudaSuccess) {
            cudaFree(d_cnt); cudaFree(d_cols); cudaFree(d_sugar);
            return set_err(e, SS_ERR_CUDA, cudaGetErrorString(ce));
        }
        uint32_t total = 0;
        ce = ss_launch_export_agents(e->w, ss_step_key(e->cfg.keyed_seed, e->iteration - 1), d_cnt, d_cnt + nb, d_cols, stride,
                                     d_sugar, e->stream);
        e->launches += 3;
        int32_t* outs[5] = {id, x, y, vision, metab};
        for (int k = 0; k < 5 && ce == cudaSuccess; k++)
            if (outs[k] && pop) ce = cudaMemcpyAsync(outs[k], d_cols + k * stride, sizeof(int32_t) * pop, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess && sugar && pop) ce = cudaMemcpyAsync(sugar, d_sugar, sizeof(double) * pop, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(&total, d_cnt + nb, sizeof total, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        cudaFree(d_cnt); cudaFree(d_cols); cudaFree(d_sugar);
        if (ce != cudaSuccess) return set_err(e, SS_ERR_CUDA, cudaGetErrorString(ce));
        if ((size_t)total != pop) return set_err(e, SS_ERR_STATE, "agent export count differs from the population counter");
        return SS_OK;
    }
    std::vector<AgentRec> recs;
    std::vector<int32_t> order;
    int rc = fetch_population(e, &recs, &order);
    if (rc) return rc;
    const int N = e->w.n;
    for (size_t i = 0; i < order.size(); i++) {
        const AgentRec& r = recs[order[i]];
        if (id) id[i] = order[i] + 1;
        if (x) x[i] = e->x0 + (int32_t)(r.pos / N);
        if (y) y[i] = (int32_t)(r.pos % N);
        if (vision) vision[i] = attr_vision(r.attr);
        if (metab) metab[i] = attr_metab(r.attr);
        if (sugar) sugar[i] = r.sugar;
    }
    return SS_OK;
}

extern "C" void* ss_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void ss_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" int ss_get_counters(ss_engine* e, int64_t* out, int32_t n) {
    if (!e || !out) return set_err(e, SS_ERR_ARG, "null argument");
    const int64_t v[6] = {e->launches, e->passes, e->steps, e->path, e->syncs, (int64_t)e->geom.emax | ((int64_t)e->pass_impl << 16)};
    for (int i = 0; i < n && i < 6; i++) out[i] = v[i];
    if (n > 6 && !e->w.dbg && e->path == PATH_LATTICE) {
        StepAccum acc;
        cudaMemcpy(&acc, e->w.acc, sizeof acc, cudaMemcpyDeviceToHost);
        const int64_t v2[5] = {(int64_t)acc.watchdog, (int64_t)acc.wd_info[0], (int64_t)acc.wd_info[1], (int64_t)acc.wd_info[2], (int64_t)acc.fault};
        for (int i = 6; i < n && i < 11; i++) out[i] = v2[i - 6];
    }
    if (n > 11) {                  // dependency rounds: grid barriers (= rounds) since the state was allocated
        unsigned int rounds = 0;
        if (e->w.dr.self.ctl) cudaMemcpy(&rounds, e->w.dr.self.ctl + 2, sizeof rounds, cudaMemcpyDeviceToHost);
        out[11] = (int64_t)rounds;
    }
    if (n > 12) out[12] = (int64_t)e->w.n_slots;      // width of the keyed order: the largest Agent.id the engine was sized for
    if (n > 6 && e->w.dbg) {       // instrumentation counters of the move passes (option "debug")
        unsigned long long d[32];
        cudaMemcpy(d, e->w.dbg, sizeof d, cudaMemcpyDeviceToHost);
        for (int i = 6; i < n && i < 38; i++) out[i] = (int64_t)d[i - 6];
    }
    return SS_OK;
}

extern "C" int ss_get_kernel_times(ss_engine* e, double* ms_out, int32_t n) {
    if (!e || !ms_out) return set_err(e, SS_ERR_ARG, "null argument");
    for (int i = 0; i < n && i < KT_COUNT; i++) ms_out[i] = e->kt[i];
    return SS_OK;
}

extern "C" int ss_set_option(ss_engine* e, const char* name, int64_t value) {
    if (!e || !name) return set_err(e, SS_ERR_ARG, "null argument");
    const std::string k(name);
    if (k == "path") {
        e->path_option = (int)value;
        if (e->have_agents) return choose_path(e);
        return SS_OK;
    }
    if (k == "profile") { e->profile = value != 0; return SS_OK; }
    if (k == "reset_times") { for (double& d : e->kt) d = 0; return SS_OK; }
    if (k == "reset_counters") { e->launches = e->passes = e->syncs = 0; return SS_OK; }
    if (k == "diffuse_impl") { e->diffuse_impl = (int)value; return SS_OK; }
    if (k == "debug") {
        if (value && !e->w.dbg) { CK(cudaMalloc(&e->w.dbg, 32 * sizeof(unsigned long long))); CK(cudaMemset(e->w.dbg, 0, 32 * 8)); }
        if (!value && e->w.dbg) { cudaFree(e->w.dbg); e->w.dbg = nullptr; }
        return SS_OK;
    }
    if (k == "pendmap") { e->use_pendmap = (int)value; return SS_OK; }
    if (k == "trace") { e->trace = (int)value; return SS_OK; }
    if (k == "window") { e->window_max = (int)value; if (e->have_agents) return choose_path(e); return SS_OK; }
    if (k == "solo_window") { e->solo_window_max = (int)value; if (e->have_agents) return choose_path(e); return SS_OK; }
    if (k == "tma") { e->use_tma = (int)value; return SS_OK; }
    if (k == "global_slots") { e->global_slots = (uint32_t)value; return SS_OK; }
    if (k == "global_vmax") { e->global_vmax = (int)value; return SS_OK; }
    if (k == "l2_persist") { e->l2_persist = (int)value; if (e->have_agents) return choose_path(e); return SS_OK; }
    if (k == "overlap") { e->overlap_option = (int)value; return SS_OK; }
    if (k == "async_queue") { e->async_queue = (int)value; return SS_OK; }
    if (k == "async_queue_strips") { e->async_queue_strips = (int)value; return SS_OK; }
    if (k == "grow_ctas") { e->grow_ctas_per_sm = (int)value; return SS_OK; }
    if (k == "life_capacity_slack") { e->life_slack = (int)value; return SS_OK; }
    if (k == "pass_impl") { e->pass_impl_option = (int)value; if (e->have_agents) return choose_path(e); return SS_OK; }
    return set_err(e, SS_ERR_ARG, "unknown option: " + k);
}

// ------------------------------------------------------------------ multi-GPU replicas ----
// sharded.py drives one step as: ss_shard_begin_step -> { ss_shard_pass -> all-gather of the
// ranks' turn logs -> ss_shard_apply } until nothing is pending -> ss_shard_end_step.
extern "C" int ss_shard_enable(ss_engine* e, int32_t rank, int32_t world) {
    if (!e || world < 1 || rank < 0 || rank >= world) return set_err(e, SS_ERR_ARG, "bad rank/world");
    if (!e->have_agents || e->path != PATH_LATTICE) return set_err(e, SS_ERR_STATE, "sharding needs the lattice path (upload first)");
    if (e->cfg.rule_pollution && (e->cfg.pollution_a < 0.0 || e->cfg.pollution_b < 0.0))
        return set_err(e, SS_ERR_UNSUPPORTED, "sharded replicas need non-negative pollution coefficients");
    static_assert(sizeof(ShardLogEntry) == SS_SHARD_ENTRY_BYTES, "log entry size");
    CK(cudaSetDevice(e->device));
    e->shard_rank = rank; e->shard_world = world;
    {   // the window size depends on the number of windows per rank (same choice on every rank)
        const int rc = choose_path(e);
        if (rc) return rc;
        if (e->path != PATH_LATTICE) return set_err(e, SS_ERR_STATE, "sharding needs the lattice path");
    }
    if (e->w.log && e->w.log_cap < 8ull * e->w.n_slots + 4096) {       // a larger population was uploaded since: new logs
        cudaFree(e->w.log); cudaFree(e->w.log_count); cudaFree(e->w.clog); cudaFree(e->w.clog_count);
        e->w.log = nullptr; e->w.log_count = nullptr; e->w.clog = nullptr; e->w.clog_count = nullptr;
    }
    if (!e->w.log) {
        e->w.log_cap = 8ull * e->w.n_slots + 4096;     // warps reserve 8 entries at a time: <= 7 unused per valid one
        CK(cudaMalloc(&e->w.log, sizeof(ShardLogEntry) * (size_t)e->w.log_cap));
        CK(cudaMalloc(&e->w.log_count, sizeof(unsigned long long)));
        e->w.clog_cap = e->w.log_cap;
        CK(cudaMalloc(&e->w.clog, sizeof(uint2) * (size_t)e->w.clog_cap));
        CK(cudaMalloc(&e->w.clog_count, sizeof(unsigned long long)));
    }
    CK(cudaMemset(e->w.log_count, 0, sizeof(unsigned long long)));
    CK(cudaMemset(e->w.clog_count, 0, sizeof(unsigned long long)));
    return SS_OK;
}

extern "C" int ss_shard_begin_step(ss_engine* e) {
    if (!e || !e->w.log) return set_err(e, SS_ERR_STATE, "ss_shard_enable first");
    CK(cudaSetDevice(e->device));
    CK(cudaMemsetAsync(e->w.hist, 0, SS_GINI_BINS * sizeof(uint32_t), e->stream));
    CK(cudaMemsetAsync(e->w.log_count, 0, sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(e->w.clog_count, 0, sizeof(unsigned long long), e->stream));
    CK(ss_launch_prepare(e->w, e->iteration, 1, e->stream));
    e->launches += e->cfg.rule_can_starve ? 2 : 1;
    e->shard_pass = 0;
    int rc = ensure_stats(e, 1);
    return rc;
}

// runs pass `shard_pass` on this rank's share of the window rows; returns the ranges it appended to the two logs
extern "C" int ss_shard_pass(ss_engine* e, uint64_t* log_begin, uint64_t* log_end, const void** log_base, uint64_t* clog_begin,
                             uint64_t* clog_end, const void** clog_base) {
    if (!e || !e->w.log || !log_begin || !log_end || !log_base || !clog_begin || !clog_end || !clog_base)
        return set_err(e, SS_ERR_STATE, "ss_shard_enable first");
    CK(cudaSetDevice(e->device));
    unsigned long long before = 0, after = 0, cbefore = 0, cafter = 0;
    CK(cudaMemcpyAsync(&before, e->w.log_count, sizeof before, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&cbefore, e->w.clog_count, sizeof cbefore, cudaMemcpyDeviceToHost, e->stream));
    PassGeom g;
    pass_geometry(e, e->shard_pass, &g);
    const int nw = g.nwin, W = e->shard_world, r = e->shard_rank;
    g.bx_lo = (int)(((long long)r * nw) / W);
    g.bx_cnt = (int)(((long long)(r + 1) * nw) / W) - g.bx_lo;
    if (e->w.pendmap && e->use_pendmap) {
        // coarse pending map, per rank and without communication: a rank knows exactly what its own windows left
        // pending; the 32-row blocks that are not entirely inside its window rows of this pass start as "pending"
        const int nb = e->w.pend_nb, n = e->w.n, nu = n / g.align;
        const size_t words = (size_t)nb * nb;
        g.map_write = e->shard_pass & 1;
        g.map_read = e->shard_pass > 0 ? ((e->shard_pass - 1) & 1) : -1;
        uint32_t* mw = e->w.pendmap + (size_t)g.map_write * words;
        CK(cudaMemsetAsync(mw, 0x01, sizeof(uint32_t) * words, e->stream));
        const long long x_lo = (long long)g.align * (((long long)g.bx_lo * nu) / nw) + g.shift_x;
        const long long x_hi = (long long)g.align * (((long long)(g.bx_lo + g.bx_cnt) * nu) / nw) + g.shift_x;
        const long long first = (x_lo + 31) / 32, last = x_hi / 32;             // block rows [first, last) are mine
        const long long cnt = std::min<long long>(std::max<long long>(last - first, 0), nb);
        const long long f0 = first % nb, c0 = std::min<long long>(cnt, nb - f0);
        if (c0 > 0) CK(cudaMemsetAsync(mw + (size_t)f0 * nb, 0, sizeof(uint32_t) * (size_t)c0 * nb, e->stream));
        if (cnt - c0 > 0) CK(cudaMemsetAsync(mw, 0, sizeof(uint32_t) * (size_t)(cnt - c0) * nb, e->stream));
    }
    if (e->pass_impl >= 1) CK(ss_launch_move_solo(e->w, g, e->iteration, e->env_time, e->pass_impl == 2, e->stream));
    else CK(ss_launch_move_pass(e->w, g, e->iteration, e->env_time, e->stream));
    e->launches++; e->passes++; e->shard_pass++;
    CK(cudaMemcpyAsync(&after, e->w.log_count, sizeof after, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&cafter, e->w.clog_count, sizeof cafter, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    if (after > e->w.log_cap || cafter > e->w.clog_cap) return set_err(e, SS_ERR_STATE, "turn log overflow");
    *log_begin = before; *log_end = after; *log_base = e->w.log;
    *clog_begin = cbefore; *clog_end = cafter; *clog_base = e->w.clog;
    return SS_OK;
}

// replays the entries (device pointers) another rank resolved in one pass on this replica: the clears of both logs, then
// the sets (a cell can be left by a safe agent and entered by an at-risk one, or the other way round, in the same pass)
extern "C" int ss_shard_apply(ss_engine* e, const void* dev_entries, uint64_t n, const void* dev_compact, uint64_t n_compact) {
    if (!e || !e->w.log) return set_err(e, SS_ERR_STATE, "ss_shard_enable first");
    CK(cudaSetDevice(e->device));
    for (int phase = 0; phase < 2; phase++) {
        CK(ss_launch_apply_log(e->w, dev_entries, n, phase, e->stream));
        CK(ss_launch_apply_clog(e->w, dev_compact, n_compact, phase, e->iteration, e->stream));
    }
    e->launches += 2 * ((n ? 1 : 0) + (n_compact ? 1 : 0));
    return SS_OK;
}

// the same for every other rank at once, straight from the all-gathered buffer: rank j's slice starts at base + j * stride
// (48-byte entries at offset 0, 8-byte entries at off_compact), counts = device int64 [world][2] (entries per log)
extern "C" int ss_shard_apply_gathered(ss_engine* e, const void* base, uint64_t stride, uint64_t off_compact, const int64_t* counts,
                                       int32_t world, uint64_t max_full, uint64_t max_compact) {
    if (!e || !e->w.log) return set_err(e, SS_ERR_STATE, "ss_shard_enable first");
    if (!base || !counts || world != e->shard_world) return set_err(e, SS_ERR_ARG, "bad gathered log buffer");
    CK(cudaSetDevice(e->device));
    for (int phase = 0; phase < 2; phase++)
        CK(ss_launch_apply_gathered(e->w, base, stride, off_compact, (const long long*)counts, world, e->shard_rank, max_full,
                                    max_compact, phase, e->iteration, e->stream));
    e->launches += 2 * ((max_full ? 1 : 0) + (max_compact ? 1 : 0));
    return SS_OK;
}

extern "C" int ss_shard_pending(ss_engine* e, uint64_t* pending) {
    if (!e || !pending) return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(e->h_pending, &e->w.acc->pending, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    *pending = *e->h_pending;
    return SS_OK;
}

extern "C" int ss_shard_end_step(ss_engine* e, ss_step_stats* out) {
    if (!e || !e->w.log) return set_err(e, SS_ERR_STATE, "ss_shard_enable first");
    CK(cudaSetDevice(e->device));
    CK(ss_launch_settle(e->w, e->iteration, e->env_time, e->stream));
    CK(ss_launch_stats(e->w, 0, e->stream));
    e->launches += 2;
    int rc = env_phase(e);
    if (rc) return rc;
    e->iteration++; e->env_time++; e->steps++;
    e->stepped = true;
    if (out) CK(cudaMemcpyAsync(out, e->w.stats, sizeof(ss_step_stats), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    return SS_OK;
}


// ------------------------------------------------------------------------------- row strips ----
// One engine per GPU holds a strip of lattice rows and the agents standing on it; the kernels of ss_rounds.cu reach the
// two ring neighbours' rows (halo of the conflict graph, crosses and moves across the border, migrating records) and
// every strip's counters (Q1 notifications) through peer-mapped device pointers.  sugarscape_utilicalc_b200/strips.py.
#define SS_STRIP_NPTR 11
static void strip_ptrs(ss_engine* e, void* p[SS_STRIP_NPTR]) {
    const DrPeer& s = e->w.dr.self;
    void* v[SS_STRIP_NPTR] = {s.occw, s.mw, s.mwx, s.sugar, s.dep, s.queue, s.ctl, s.agents, s.delta, s.acc, s.hist};
    for (int k = 0; k < SS_STRIP_NPTR; k++) p[k] = v[k];
}
static int strip_attach(ss_engine* e, int peer_rank, void* const p[SS_STRIP_NPTR]) {
    const int P = e->strip_world, N = e->w.n;
    DrPeer q;
    memset(&q, 0, sizeof q);
    q.occw = (uint32_t*)p[0]; q.mw = (uint8_t*)p[1]; q.mwx = (uint8_t*)p[2]; q.sugar = (double*)p[3]; q.poll = nullptr;
    q.dep = (uint32_t*)p[4]; q.queue = (uint32_t*)p[5]; q.ctl = (unsigned int*)p[6]; q.agents = (AgentRec*)p[7];
    q.delta = (uint32_t*)p[8]; q.acc = (unsigned long long*)p[9]; q.hist = (uint32_t*)p[10];
    q.x0 = (int)(((long long)peer_rank * N) / P);
    q.rows = (int)(((long long)(peer_rank + 1) * N) / P) - q.x0;
    DrState& d = e->w.dr;
    d.all[peer_rank] = q;
    if (peer_rank == (e->strip_rank + P - 1) % P) d.lo = q;
    if (peer_rank == (e->strip_rank + 1) % P) d.hi = q;
    e->peers_connected++;
    return SS_OK;
}
static int strip_ready(ss_engine* e, const char* what) {
    if (!e) return set_err(e, SS_ERR_ARG, "null argument");
    if (e->strip_world < 2) return set_err(e, SS_ERR_STATE, std::string(what) + ": not a strip engine (ss_create_strip with world > 1)");
    if (!e->have_cells || !e->have_agents || e->path != PATH_LATTICE || e->pass_impl != 3)
        return set_err(e, SS_ERR_STATE, std::string(what) + ": upload the strip's cells and agents first");
    return SS_OK;
}

// the strip's peer-visible arrays as raw device pointers (peers in the same process) ...
extern "C" int ss_strip_export(ss_engine* e, uint64_t* ptrs) {
    int rc = strip_ready(e, "ss_strip_export");
    if (rc) return rc;
    void* p[SS_STRIP_NPTR];
    strip_ptrs(e, p);
    for (int k = 0; k < SS_STRIP_NPTR; k++) ptrs[k] = (uint64_t)(uintptr_t)p[k];
    return SS_OK;
}
// ... and as CUDA IPC handles (one process per GPU): SS_STRIP_NPTR x (64-byte handle of the allocation the array lives in
// + 8-byte offset of the array inside it -- the driver serves small cudaMallocs from shared blocks, and a handle always
// names the whole block)
struct StripIpcEntry { cudaIpcMemHandle_t handle; uint64_t offset; };
static_assert(sizeof(StripIpcEntry) == 72, "IPC entry size");
typedef int (*MemGetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);
extern "C" int ss_strip_ipc_export(ss_engine* e, void* handles) {
    int rc = strip_ready(e, "ss_strip_ipc_export");
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn)
        return set_err(e, SS_ERR_CUDA, "cuMemGetAddressRange is not available");
    void* p[SS_STRIP_NPTR];
    strip_ptrs(e, p);
    StripIpcEntry* out = reinterpret_cast<StripIpcEntry*>(handles);
    for (int k = 0; k < SS_STRIP_NPTR; k++) {
        unsigned long long base = 0;
        size_t size = 0;
        if (((MemGetAddressRangeFn)fn)(&base, &size, (unsigned long long)(uintptr_t)p[k]) != 0) return set_err(e, SS_ERR_CUDA, "cuMemGetAddressRange failed");
        CK(cudaIpcGetMemHandle(&out[k].handle, (void*)(uintptr_t)base));
        out[k].offset = (uint64_t)((unsigned long long)(uintptr_t)p[k] - base);
    }
    return SS_OK;
}
extern "C" int ss_strip_connect(ss_engine* e, int32_t peer_rank, const uint64_t* ptrs) {
    int rc = strip_ready(e, "ss_strip_connect");
    if (rc) return rc;
    if (peer_rank < 0 || peer_rank >= e->strip_world || peer_rank == e->strip_rank || !ptrs) return set_err(e, SS_ERR_ARG, "bad peer");
    void* p[SS_STRIP_NPTR];
    for (int k = 0; k < SS_STRIP_NPTR; k++) p[k] = (void*)(uintptr_t)ptrs[k];
    return strip_attach(e, peer_rank, p);
}
extern "C" int ss_strip_ipc_connect(ss_engine* e, int32_t peer_rank, const void* handles) {
    int rc = strip_ready(e, "ss_strip_ipc_connect");
    if (rc) return rc;
    if (peer_rank < 0 || peer_rank >= e->strip_world || peer_rank == e->strip_rank || !handles) return set_err(e, SS_ERR_ARG, "bad peer");
    CK(cudaSetDevice(e->device));
    void* p[SS_STRIP_NPTR];
    const StripIpcEntry* in = reinterpret_cast<const StripIpcEntry*>(handles);
    for (int k = 0; k < SS_STRIP_NPTR; k++) {
        void* base = nullptr;
        for (int j = 0; j < k && !base; j++)                // several arrays of one block: the block is opened once
            if (memcmp(&in[j].handle, &in[k].handle, sizeof(cudaIpcMemHandle_t)) == 0) base = (char*)p[j] - in[j].offset;
        if (!base) {
            CK(cudaIpcOpenMemHandle(&base, in[k].handle, cudaIpcMemLazyEnablePeerAccess));
            e->ipc_ptrs[peer_rank][k] = base;
        }
        p[k] = (char*)base + in[k].offset;
    }
    e->peers_ipc = true;
    return strip_attach(e, peer_rank, p);
}

// After the uploads: the strip's cell bytes and agent flags, and the replicated directories (alive / at-risk bitmaps,
// the strip that holds each agent) from the agents of the WHOLE world -- every rank passes the same arrays.
extern "C" int ss_strip_init_directory(ss_engine* e, int32_t n_all, const int32_t* id, const int32_t* x, const int32_t* metab,
                                       const double* sugar) {
    int rc = strip_ready(e, "ss_strip_init_directory");
    if (rc) return rc;
    if (n_all < 0 || (n_all > 0 && (!id || !x || !metab || !sugar))) return set_err(e, SS_ERR_ARG, "null argument");
    CK(cudaSetDevice(e->device));
    const size_t nn = (size_t)std::max(n_all, 1);
    int32_t* d_i32 = nullptr;
    double* d_f64 = nullptr;
    struct Staging { int32_t*& a; double*& b; ~Staging() { cudaFree(a); cudaFree(b); } } staging{d_i32, d_f64};
    CK(cudaMalloc(&d_i32, sizeof(int32_t) * 3 * nn));
    CK(cudaMalloc(&d_f64, sizeof(double) * nn));
    const int32_t* cols[3] = {id, x, metab};
    for (int k = 0; k < 3 && n_all > 0; k++) CK(cudaMemcpyAsync(d_i32 + k * nn, cols[k], sizeof(int32_t) * n_all, cudaMemcpyHostToDevice, e->stream));
    if (n_all > 0) CK(cudaMemcpyAsync(d_f64, sugar, sizeof(double) * n_all, cudaMemcpyHostToDevice, e->stream));
    const size_t words = (size_t)(e->dr_slots + 31) / 32 + 1;
    CK(cudaMemsetAsync(e->w.dr.alive, 0, sizeof(uint32_t) * words, e->stream));
    CK(cudaMemsetAsync(e->w.dr.atrisk, 0, sizeof(uint32_t) * words, e->stream));
    CK(ss_launch_dr_directory(e->w, n_all, d_i32, d_i32 + nn, d_i32 + 2 * nn, d_f64, e->stream));
    int flag = 0;
    CK(cudaMemsetAsync(e->d_flags, 0, sizeof(int), e->stream));
    CK(ss_launch_dr_mirror(e->w, e->d_flags, e->stream));
    CK(cudaMemcpyAsync(&flag, e->d_flags, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    StepAccum acc;
    memset(&acc, 0, sizeof acc);
    acc.population = (unsigned long long)n_all;
    CK(cudaMemcpyAsync(e->w.dr.acc_g, &acc, sizeof acc, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->launches += 4; e->syncs++;
    if (flag) return set_err(e, SS_ERR_UNSUPPORTED, "dependency rounds need cell sugar and capacities < 127");
    e->mirror = 2;
    return SS_OK;
}

static int strip_connected(ss_engine* e, const char* what) {
    int rc = strip_ready(e, what);
    if (rc) return rc;
    if (e->peers_connected < e->strip_world - 1) return set_err(e, SS_ERR_STATE, std::string(what) + ": connect every peer strip first");
    if (e->mirror != 2) return set_err(e, SS_ERR_STATE, std::string(what) + ": ss_strip_init_directory first");
    return SS_OK;
}
static int strip_begin_build(ss_engine* e) {
    e->dr_dirty = true;
    CK(cudaMemsetAsync(e->w.hist, 0, SS_GINI_BINS * sizeof(uint32_t), e->stream));
    CK(ss_launch_prepare(e->w, e->iteration, 0, e->stream));
    CK(ss_launch_dr_build(e->w, e->iteration, e->tmap, e->tma_ok && e->use_tma, e->stream));
    e->launches += 5;
    return SS_OK;
}
static int strip_finish(ss_engine* e, int stats_index, cudaStream_t stream = nullptr) {
    if (!stream) stream = e->stream;
    DevWorld g = e->w;                          // the statistics kernel on the merged sums
    g.acc = reinterpret_cast<StepAccum*>(e->w.dr.acc_g);
    g.hist = e->w.dr.hist_g;
    g.outliers = nullptr;                       // (wealth outside the histogram: Gini is reported as NaN on strips)
    {
        KTimer tm(e, KT_STATS);
        CK(ss_launch_stats(g, stats_index, stream));
    }
    e->launches++;
    int rc = env_phase(e, stream);
    if (rc) return rc;
    e->iteration++; e->env_time++; e->steps++;
    e->stepped = true;
    return SS_OK;
}

// Host-driven phases (several strips on one GPU, driven by one process: tests).  0: begin + conflict graph; 1: one round
// over the queue segment [seg_begin, seg_end); 2: after the rounds (statistics merged over the strips, delta lists
// applied); 3: statistics + growback.  Every strip finishes a phase before any strip starts the next one.
// *tail_out = the strip's queue tail after the phase; *stats_out (phase 3) = the step's statistics.
extern "C" int ss_strip_phase(ss_engine* e, int32_t phase, uint32_t seg_begin, uint32_t seg_end, uint32_t* tail_out, ss_step_stats* stats_out) {
    int rc = strip_connected(e, "ss_strip_phase");
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    if (phase == 0) {
        rc = ensure_stats(e, 1);
        if (rc) return rc;
        rc = strip_begin_build(e);
        if (rc) return rc;
    } else if (phase == 1) {
        if (seg_end > seg_begin) {
            CK(ss_launch_dr_rounds(e->w, e->iteration, e->env_time, e->sm_count, 1, seg_begin, seg_end, 0, e->stream));
            e->launches++;
        }
    } else if (phase == 2) {
        CK(ss_launch_dr_check(e->w, e->stream));
        CK(ss_launch_dr_after_rounds(e->w, e->stream));
        e->launches += 3; e->passes++;
    } else if (phase == 3) {
        rc = strip_finish(e, 0);
        if (rc) return rc;
        if (stats_out) CK(cudaMemcpyAsync(stats_out, e->w.stats, sizeof(ss_step_stats), cudaMemcpyDeviceToHost, e->stream));
    } else return set_err(e, SS_ERR_ARG, "bad phase");
    uint32_t tail = 0;
    CK(cudaMemcpyAsync(&tail, e->w.dr.self.ctl, sizeof tail, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    if (tail_out) *tail_out = tail;
    return SS_OK;
}

// One process per GPU: nsteps timesteps with the strips synchronised on the device (barriers over peer memory between
// the phases and between the rounds); no host round trip inside a step.
extern "C" int ss_strip_step(ss_engine* e, int32_t nsteps, ss_step_stats* out) {
    int rc = strip_connected(e, "ss_strip_step");
    if (rc) return rc;
    if (nsteps < 0) return set_err(e, SS_ERR_ARG, "nsteps < 0");
    if (nsteps == 0) return SS_OK;
    CK(cudaSetDevice(e->device));
    rc = ensure_stats(e, nsteps);
    if (rc) return rc;
    CK(cudaEventRecord(e->evs, e->stream));
    // Statistics and growback of step t run on the second stream beside the conflict graph of step t + 1 (which reads the
    // neighbours' border rows for occupancy only); the barrier ahead of the rounds then stands for both "every strip's
    // counters and queue are set up" and "every strip's growback is done: the halos are final".
    const bool overlap = e->overlap_option && !e->profile;
    for (int t = 0; t < nsteps; t++) {
        const int async_q = e->async_queue_strips;                 // work-list mode: no exchange per round, the strips meet after the kernel
        {
            KTimer tm(e, KT_PREPARE);                               // (option "profile": device time per phase, with a host sync each)
            if (async_q) CK(cudaMemsetAsync(e->w.dr.self.queue, 0, sizeof(uint32_t) * ((size_t)e->dr_slots + 64), e->stream));
            rc = strip_begin_build(e);
            if (rc) return rc;
        }
        {
            KTimer tm(e, KT_MOVE);
            rc = env_join(e);                                       // this strip's growback of the previous step
            if (rc) return rc;
            CK(ss_launch_dr_strip_barrier(e->w, e->stream));        // every strip: conflict graph built, previous growback done
            CK(ss_launch_dr_rounds(e->w, e->iteration, e->env_time, e->sm_count, 0, 0u, 0u, async_q, e->stream));
            // every strip has drained its queue and added its sums (the rounds kernel does that after its last turn -- also
            // after the last exchange of the version with one exchange per round)
            CK(ss_launch_dr_strip_barrier(e->w, e->stream));
        }
        {
            KTimer tm(e, KT_SERIAL);                                // ("serial" slot: merge of the statistics + delta lists + barrier)
            CK(ss_launch_dr_check(e->w, e->stream));
            CK(ss_launch_dr_after_rounds(e->w, e->stream));
            CK(ss_launch_dr_strip_barrier(e->w, e->stream));        // everybody has read everybody's sums and delta list
        }
        if (overlap) {
            CK(cudaEventRecord(e->ev_rounds, e->stream));
            CK(cudaStreamWaitEvent(e->stream_env, e->ev_rounds, 0));
            rc = strip_finish(e, t, e->stream_env);                 // statistics kernel + growback
            if (rc) return rc;
            CK(cudaEventRecord(e->ev_env, e->stream_env));
            e->env_in_flight = true;
        } else {
            rc = strip_finish(e, t);                                // (timed as "grow")
            if (rc) return rc;
            KTimer tm(e, KT_SETTLE);
            CK(ss_launch_dr_strip_barrier(e->w, e->stream));        // every strip's growback is done: the halos are final
        }
        e->launches += 7; e->passes++;
    }
    if (overlap) {
        rc = env_join(e);
        if (rc) return rc;
        CK(ss_launch_dr_strip_barrier(e->w, e->stream));            // ... before anybody reads a neighbour's rows again
    }
    CK(cudaEventRecord(e->eve, e->stream));
    if (out) CK(cudaMemcpyAsync(out, e->w.stats, sizeof(ss_step_stats) * (size_t)nsteps, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->syncs++;
    float ms = 0;
    cudaEventElapsedTime(&ms, e->evs, e->eve);
    e->kt[KT_LAST_CALL] = ms;
    return SS_OK;
}
