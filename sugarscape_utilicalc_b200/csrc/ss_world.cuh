// ss_world.cuh -- the device-resident world every kernel receives by value.
#pragma once
#include "ss_device.cuh"

#define SS_GINI_BINS 65536          // wealth histogram bins of the lattice path (integer sugar)

struct StepAccum {                  // per-step reduction targets of the lattice path
    unsigned long long sum_metab;
    unsigned long long sum_vision;
    unsigned long long acted;
    unsigned long long deaths;
    unsigned long long pending;     // agents not yet resolved in the current step
    unsigned long long population;  // living agents
    unsigned long long hist_small;  // agents whose wealth is the 0.01 floor
    unsigned long long inexact;     // wealth values the histogram cannot represent
    unsigned long long fault;       // internal consistency violations (must stay 0)
    unsigned long long watchdog;    // spin-wait watchdog firings since upload (diagnostic; 0 in healthy runs)
    unsigned long long wd_info[3];
    unsigned long long outliers;    // entries of DevWorld::outliers (wealth values outside the histogram, kept one by one)
};
#define SS_GINI_OUTLIERS 4096       // the statistics stay exact with up to this many such values per step

// One resolved turn, as shipped to the other ranks' replicas (multi-GPU, see sharded.py)
struct ShardLogEntry {
    uint32_t slot, old_cell, new_cell, new_word;
    double sugar;                   // at-risk agents: sugar after the turn
    double poll;                    // at-risk agents: pollution of new_cell after the turn
    uint32_t attr, status, turn;
    uint32_t kind_harvest;          // bits 0..7 kind (0 safe, 1 at-risk, 2 skipped), bits 8.. harvest
};

// Dependency rounds over the whole lattice (ss_rounds.cu): per-step dependency state.
//   mw[c]     u8   bit 7 = cell occupied, bits 0..6 = int(sugar[c]) -- the one byte an agent reads per cross cell;
//                  mwx is the same array transposed (index y * rows + x - x0), so that BOTH arms of a cross are
//                  contiguous bytes: two 16-byte loads per arm instead of one 64-byte DRAM atom per cell
//   dep[s]    u16  by slot (two per u32 word; 33 MB at 16.7 M agents: L2-resident): bits 0..14 = unresolved
//                  predecessors, bit 15 = "lose the turn" (Q1: the predecessor in the list order died)
//   succ[s]   64-byte record by slot: [0] = number of successors, [1..14] = successor slots (bits 24..25: the strip
//                  that holds it, 0 self / 1 lo / 2 hi), [15] = first entry of the remaining ones in `pool`
//   queue     slots in the order their agents became ready; round k is the segment appended during round k-1
//   alive / atrisk  bitmaps by slot (list-order predecessor lookup for Q1); q1wait[s] = {slot waiting for s's outcome
//                  | strip << 24, step + 1}
struct DrPeer {                     // one strip of lattice rows [x0, x0 + rows) and its per-step state, as a kernel sees it
    uint32_t* occw;                 // (local or peer-mapped device pointers; cell index (x - x0) * n + y)
    uint8_t* mw;
    uint8_t* mwx;
    double* sugar;
    double* poll;
    uint32_t* dep;
    uint32_t* queue;
    unsigned int* ctl;              // see DR_CTL_*
    AgentRec* agents;
    uint32_t* delta;                // what this strip's turns changed in the replicated directories (alive / at-risk / home)
    unsigned long long* acc;        // StepAccum of the strip (statistics are merged over the strips)
    uint32_t* hist;
    int x0, rows;
};
// control words of a strip (device memory, written by the kernels; peers write [DR_CTL_FLAGS..) over NVLink)
#define DR_CTL_TAIL 0               // queue tail
#define DR_CTL_BAR 1                // grid barrier: arrivals, [2] generation, [3..4] tail snapshots
#define DR_CTL_TAIL0 5              // queue tail after the build
#define DR_CTL_POOL 6               // successor pool tail, [7] pool overflows
#define DR_CTL_DELTA 8              // entries of the delta list
#define DR_CTL_AGENTS 9             // agents the build kernel saw (every one of them must pass through the queue)
#define DR_CTL_MORE 10              // after a round: does any strip have work left (written with the snapshots)
#define DR_CTL_FLAGS 16             // [16 + 2r], [17 + 2r]: what strip r published (round barrier: sequence number, count)
#define DR_CTL_WORDS 64
#define DR_MAX_STRIPS 8
// delta list entries: slot | kind << 24
#define DR_DELTA_DIED 1u
#define DR_DELTA_RISK 2u            // the at-risk bit flipped
#define DR_DELTA_TO_LO 4u           // the agent went to the strip below / above
#define DR_DELTA_TO_HI 8u
#define DR_SUCC_INLINE 14
struct DrState {
    DrPeer self, lo, hi;            // this strip and its ring neighbours (single strip: lo = hi = self)
    DrPeer all[DR_MAX_STRIPS];      // every strip by rank (Q1 notifications, delta lists, statistics)
    int rank, world;
    uint8_t* home;                  // by slot: the strip that holds the agent (replicated; strips only)
    uint32_t delta_cap;
    unsigned long long* acc_g;      // merged StepAccum / histogram (strips: every rank computes the same statistics)
    uint32_t* hist_g;
    uint4* succ;                    // 4 x uint4 per slot
    uint32_t* pool;
    uint32_t pool_cap;
    uint32_t* alive;
    uint32_t* atrisk;
    uint32_t* q1flag;               // bitmap by slot: the list-order predecessor may starve this step
    uint2* q1wait;
    uint32_t queue_cap;
};

struct DevWorld {
    ss_config cfg;
    int n;                          // lattice edge N
    int x0, rows;                   // the lattice rows this engine holds: [x0, x0 + rows) (all of them unless it is a strip)
    uint32_t n_slots;               // max Agent.id
    int half;                       // Feistel half width for the keyed order
    double* sugar;
    double* cap;
    double* poll;
    uint32_t* occw;
    uint8_t* isugar;                // int(sugar) mirror (lattice path)
    AgentRec* agents;
    // serial path (SS_RNG_MT19937, and the universal fallback)
    int32_t* order;                 // persistent list order (slots)
    int32_t* n_order;               // device scalar
    double* wealth;                 // scratch, pow2-padded
    unsigned long long* sortkeys;   // scratch, pow2-padded
    uint32_t* mt;                   // 624 words
    int32_t* mti;
    ss_step_stats* stats;           // per-step outputs, device
    // lattice path
    StepAccum* acc;
    uint32_t* hist;                 // SS_GINI_BINS
    double* outliers;               // SS_GINI_OUTLIERS
    int vmax;                       // max vision over all agents
    unsigned long long* dbg;        // optional instrumentation counters (32 entries) or null
    uint32_t* pendmap;              // 2 x pend_nb^2 counters: agents a launch left pending, per 32x32-cell block
    int pend_nb;
    struct ShardLogEntry* log;      // multi-GPU: turns resolved by this rank (null on a single GPU): at-risk agents, Q1 skips
    unsigned long long* log_count;
    unsigned long long log_cap;
    uint2* clog;                    // ... and the turns of safe agents, 8 bytes each: x = slot | harvest << 24, y = new cell
    unsigned long long* clog_count; //     (old cell, occupancy word and turn stamp follow from the replica's own record)
    unsigned long long clog_cap;
    DrState dr;                     // dependency rounds (ss_rounds.cu)
};

// A wealth value the histogram has no bin for (non-integer -- a fractional uploaded endowment -- or >= SS_GINI_BINS):
// kept individually, merged in value order by ss_stats_kernel.
__device__ inline void ss_note_outlier(const DevWorld& w, double sugar) {
    const unsigned long long at = atomicAdd(&w.acc->outliers, 1ull);
    if (at < SS_GINI_OUTLIERS) w.outliers[at] = sugar;
}

struct PassGeom {                   // window tiling of one move pass
    int nwin;                       // windows per dimension
    int emax;                       // max window extent (smem row stride)
    int ew;                         // bitmap words per row
    int shift_x, shift_y;           // tiling shift
    int ring;                       // 2*vmax: agents closer than this to a window edge are not resolved
    int align;                      // window boundaries are multiples of this many cells (4: vector loads)
    int bx_lo, bx_cnt;              // window rows [bx_lo, bx_lo + bx_cnt) of this launch (multi-GPU: this rank's share)
    int map_read, map_write;        // coarse pending maps (32x32-cell blocks) of the previous / this launch, -1 = unused
    int ring_per_agent;             // warp-per-window scan: ring = vision + vmax per agent instead of 2 vmax for all (fewer passes,
                                    // a longer first pass: pays when every pass costs an exchange, i.e. in the sharded path)
};
