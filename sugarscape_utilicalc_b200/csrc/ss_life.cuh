// ss_life.cuh -- SURVEY 8 f3: what the limitedLifespan / procreate rules add to the agent store (serial path, MT19937 mode).
//
// Agent.age / maxAge / sex / fertility / tags, sugarEndowment (= sugarCostForMating with the (0, 0) cost range, agent.py:204-206)
// and Agent.alive, which is not the same as "in View.agents": an agent that reaches maxAge gets alive = False in incAge
// (agent.py:636-639, sugarscape.py:588-591) but stays in the list and on the lattice, takes one more turn and is removed at
// its end ("died of starvation": age == maxAge is not > maxAge, sugarscape.py:597-600).
#pragma once
#include <stdint.h>

struct LifeRec {
    int32_t age, max_age, sex, fert_lo, fert_hi, tags, living, pad;
    double endow;
};

struct LifeWorld {
    LifeRec* rec;               // by slot (= id - 1), `capacity` entries; null when the rules are off
    int32_t* id_increment;      // Environment.idIncrement: ids handed out so far (device scalar)
    int32_t capacity;           // slots the agent store, the list and the scratch arrays are allocated for
    int32_t* events;            // 4 ints per event, in execution order: kind | step << 8, a, b, c
    int32_t* n_events;          //   kind 1: a mated with b and created c (sugarscape.py:551); 2: a died of starvation;
    int32_t ev_cap;             //   3: a died of old age (sugarscape.py:597-600)
    int32_t* status;            // [0] timesteps this launch completed (it stops before a step the store may not hold),
                                // [1] faults: a birth found the store or the event buffer full (must stay 0)
};

// the spice rule's state (serial path).  Environment.grid slots [1] spice and [3] spice capacity; Agent.spice,
// spiceMetabolism, spiceEndowment (= spiceCostForMating) by slot.
struct SpiceWorld {
    double* spice;              // cells; null when the rule is off
    double* spice_cap;
    double* aspice;             // by slot (same capacity as the agent store)
    double* sp_endow;
    int32_t* asp_metab;
    int32_t* pow_faults;        // operands the restated pow had no answer for (must stay 0)
};
