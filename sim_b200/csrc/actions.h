// sim_b200/csrc/actions.h -- the ten atomic models as TWO decision tables.
//
// The generic interpreter (core.cuh, Engine::ext2 / int2) does not dispatch on the model type: every
// events_ext / events_int of every built-in model is "look at (type, port, phase, queue fill), then push /
// pop jobs, set the phase, set until_next_event, record".  The first half is a table lookup, the second
// half ONE code path shared by all replications of a warp whatever model each of them is serving, so
// lanes at different models of the network stay converged.  Only the three table-based models
// (ParallelGateway collections, Stopwatch jobs, StochasticGate's draw at arrival) keep a body of
// their own ("special").
//
// The tables are built at compile time from the two constexpr functions below, which restate the
// reference's match arms (cited per model).  They are the single statement of the model semantics of
// the generic interpreter; tests/hostsim runs exactly this code against the oracle.
#pragma once
#include <stdint.h>

#include "net.h"

namespace simb {

// ---- events_ext ------------------------------------------------------------------------------------
// key = type * 128 + port class (0, 1, 2, 3 = any other / unknown) * 32 + phase * 8 + queue state
// queue state bits: 1 = empty (len == 0), 2 = full (len == cap), 4 = no room (len + 1 >= cap)
#define XK_EMPTY 1u
#define XK_FULL 2u
#define XK_NOROOM 4u
#define EXT_KEY(type, pc, ph, qs) ((type) * 128u + (pc) * 32u + (ph) * 8u + (qs))

// action word
#define XA_ERR_SHIFT 0      /* 2 bits: 0 none, 1 InvalidMessage (7), 2 InvalidModelState (5) */
#define XA_PUSH (1u << 2)   /* append the job to the model's FIFO ring (+ arrival stamp when MF_STAT) */
#define XA_STORE (1u << 3)  /* Storage: keep the job as the stored value */
#define XA_SETPH (1u << 4)
#define XA_PH_SHIFT 5       /* 2 bits */
#define XA_U_SHIFT 7        /* 3 bits: UA_* */
#define XA_WAIT0 (1u << 10) /* Processor activate: waiting time 0 enters the statistic */
#define XA_SPECIAL_SHIFT 11 /* 2 bits: XS_* */
#define XA_REC1_SHIFT 16    /* 4 bits: SimbAction, 0xF = none */
#define XA_REC2_SHIFT 20
#define XA_AUXPORT (1u << 24) /* the record's aux byte carries the input port index + 1 (0 = unknown port) */
#define XA_NOP_REC (0xFFu << XA_REC1_SHIFT)

enum { UA_KEEP = 0, UA_ZERO = 1, UA_INF = 2, UA_P0 = 3, UA_DRAW = 4 };
enum { XS_NONE = 0, XS_SGATE = 1, XS_PG = 2, XS_STOPWATCH = 3 };

constexpr uint32_t xa(uint32_t flags, uint32_t u, int rec1 = 0xF, int rec2 = 0xF) {
    return flags | (u << XA_U_SHIFT) | ((uint32_t)rec1 << XA_REC1_SHIFT) | ((uint32_t)rec2 << XA_REC2_SHIFT);
}
constexpr uint32_t xa_ph(uint32_t ph) { return XA_SETPH | (ph << XA_PH_SHIFT); }
constexpr uint32_t xa_err(uint32_t code) { return (code << XA_ERR_SHIFT) | XA_NOP_REC; }
constexpr uint32_t xa_special(uint32_t s) { return (s << XA_SPECIAL_SHIFT) | XA_NOP_REC; }

constexpr uint32_t ext_action(uint32_t type, uint32_t pc, uint32_t ph, uint32_t qs) {
    const bool empty = (qs & XK_EMPTY) != 0, full = (qs & XK_FULL) != 0, noroom = (qs & XK_NOROOM) != 0;
    switch (type) {
        case MT_PROCESSOR:  // processor.rs:211-227
            if (pc != 0) return xa_err(1);
            if (empty && full) return xa_err(2);                    // (true, true): capacity 0
            if (full) return xa(0, UA_KEEP, ACT_DROP);              // ignore_job :152-158
            if (empty)                                              // activate :128-150
                return xa(XA_PUSH | xa_ph(1) | XA_WAIT0, UA_DRAW, ACT_ARRIVAL, ACT_PROCESSING_START);
            return xa(XA_PUSH, UA_KEEP, ACT_ARRIVAL);               // add_job :119-126
        case MT_STORAGE:  // storage.rs:151-161
            if (pc == 0) return xa(XA_STORE, UA_KEEP, ACT_ARRIVAL);  // hold_job :106-113
            if (pc == 1) return xa(xa_ph(1), UA_ZERO);               // get_job :101-104
            return xa_err(1);
        case MT_LOAD_BALANCER:  // pass_job load_balancer.rs:78-87 (any port)
            return xa(XA_PUSH | xa_ph(1), UA_ZERO, ACT_ARRIVAL);
        case MT_GATE:  // gate.rs:180-195
            if (pc == 0) {
                if (ph != 1) return xa(XA_PUSH | xa_ph(2), UA_ZERO, ACT_ARRIVAL);  // pass_job :127-137
                return xa(0, UA_KEEP, ACT_ARRIVAL);                                // drop_job :139-145 (Closed)
            }
            if (pc == 1) return xa(xa_ph(0), UA_INF, ACT_ACTIVATION);    // activate :107-115
            if (pc == 2) return xa(xa_ph(1), UA_INF, ACT_DEACTIVATION);  // deactivate :117-125
            return xa_err(1);
        case MT_EXCLUSIVE_GATEWAY:  // pass_job exclusive_gateway.rs:95-108 (any port)
            return xa(XA_PUSH | xa_ph(1) | XA_AUXPORT, UA_ZERO, ACT_ARRIVAL);
        case MT_PARALLEL_GATEWAY:  // parallel_gateway.rs:160-172
            return xa_special(XS_PG);
        case MT_STOCHASTIC_GATE:  // stochastic_gate.rs:163-172
            if (pc != 0) return xa_err(1);
            return xa_special(XS_SGATE);
        case MT_STOPWATCH:  // stopwatch.rs:271-282
            if (pc == 2) return xa(xa_ph(1), UA_ZERO);  // get_job :211-214
            if (pc > 2) return xa_err(1);
            return xa_special(XS_STOPWATCH);
        case MT_BATCHER:  // batcher.rs:192-206 (any port)
            if (!noroom && ph == 2) return xa_err(2);                                // (Release, true)
            if (!noroom && ph == 0) return xa(XA_PUSH | xa_ph(1), UA_P0, ACT_ARRIVAL);  // start_batch :103-112
            if (!noroom) return xa(XA_PUSH, UA_KEEP, ACT_ARRIVAL);                   // add_to_batch :93-101
            return xa(XA_PUSH | xa_ph(2), UA_ZERO, ACT_ARRIVAL);                     // fill_batch :114-123
        default:  // Generator (generator.rs:161-167), Passive (tests/custom.rs:49-56): no-op
            return xa(0, UA_KEEP);
    }
}

// ---- events_int ------------------------------------------------------------------------------------
// key = type * 32 + phase * 8 + queue state
// queue state bits: 1 = empty (len == 0), 2 = len <= cap, 4 = len >= 2 * cap
#define IK_EMPTY 1u
#define IK_LE 2u
#define IK_GE2 4u
#define INT_KEY(type, ph, qs) ((type) * 32u + (ph) * 8u + (qs))

#define IA_ERR_SHIFT 0     /* 2 bits: 0 none, 1 InvalidModelState (5), 2 panic in the reference (34) */
#define IA_SETPH (1u << 4)
#define IA_PH_SHIFT 5
#define IA_U_SHIFT 7       /* 3 bits: UA_* (UA_DRAW = the draw comes before anything else) */
#define IA_NREL_SHIFT 10   /* 2 bits: jobs leaving the ring: NR_* */
#define IA_KIND_SHIFT 12   /* 3 bits: IK_* */

enum { NR_NONE = 0, NR_ONE = 1, NR_ALL = 2, NR_CAP = 3 };
enum {
    KI_PLAIN = 0,       // released jobs leave on port 0
    KI_LB = 1,          // round robin: the port index is advanced first (load_balancer.rs:95-111)
    KI_XGW = 2,         // one index draw per firing selects the port (exclusive_gateway.rs:110-134)
    KI_GEN = 3,         // Generator: initialise, or make job number last_job + 1
    KI_STORED = 4,      // Storage / Stopwatch: send the kept value, if any
    KI_PROC_START = 5,  // Processor: service starts on the job at the head of the queue
    KI_SGATE = 6,       // StochasticGate: the head job passes or is blocked as drawn at its arrival
    KI_PG = 7           // ParallelGateway: first full collection goes to every output port
};

constexpr uint32_t ia(uint32_t flags, uint32_t u, uint32_t nrel = NR_NONE, uint32_t kind = KI_PLAIN) {
    return flags | (u << IA_U_SHIFT) | (nrel << IA_NREL_SHIFT) | (kind << IA_KIND_SHIFT);
}
constexpr uint32_t ia_ph(uint32_t ph) { return IA_SETPH | (ph << IA_PH_SHIFT); }

constexpr uint32_t int_action(uint32_t type, uint32_t ph, uint32_t qs) {
    const bool empty = (qs & IK_EMPTY) != 0, le = (qs & IK_LE) != 0, ge2 = (qs & IK_GE2) != 0;
    switch (type) {
        case MT_GENERATOR:  // generator.rs:169-177
            return ia(0, UA_DRAW, NR_NONE, KI_GEN);
        case MT_PROCESSOR:  // processor.rs:229-238
            if (ph == 0) {
                if (empty) return ia(0, UA_INF);                         // passivate :192-196
                return ia(ia_ph(1), UA_DRAW, NR_NONE, KI_PROC_START);    // process_next :160-175
            }
            if (empty) return 2u << IA_ERR_SHIFT;                        // queue.remove(0) on an empty Vec
            return ia(ia_ph(0), UA_ZERO, NR_ONE);                        // release_job :177-190
        case MT_STORAGE:  // storage.rs:163-171
            if (ph == 1) return ia(ia_ph(0), UA_INF, NR_NONE, KI_STORED);  // release_job :115-130
            return ia(0, UA_INF);
        case MT_LOAD_BALANCER:  // load_balancer.rs:134-142
            if (empty) return ia(ia_ph(0), UA_INF);
            return ia(0, UA_ZERO, NR_ONE, KI_LB);  // send_job :95-111
        case MT_GATE:  // send_jobs gate.rs:149-165
            return ia(ia_ph(0), UA_INF, NR_ALL);
        case MT_EXCLUSIVE_GATEWAY:  // exclusive_gateway.rs:153-161
            if (ph == 1) return ia(ia_ph(0), UA_INF, NR_ALL, KI_XGW);  // send_jobs :110-134
            return ia(0, UA_INF);
        case MT_PARALLEL_GATEWAY:  // parallel_gateway.rs:174-182
            return ia(0, UA_KEEP, NR_NONE, KI_PG);
        case MT_STOCHASTIC_GATE:  // stochastic_gate.rs:174-183
            if (empty) return ia(0, UA_INF);
            return ia(0, UA_ZERO, NR_NONE, KI_SGATE);
        case MT_STOPWATCH:  // stopwatch.rs:284-293
            if (ph == 1) return ia(ia_ph(0), UA_INF, NR_NONE, KI_STORED);  // release_minimum / maximum :216-250
            return ia(0, UA_INF);
        case MT_BATCHER:  // batcher.rs:208-221
            if (le && ge2) return 1u << IA_ERR_SHIFT;
            if (le) return ia(ia_ph(0), UA_INF, NR_ALL);   // release_full_queue :125-141
            if (ge2) return ia(ia_ph(2), UA_ZERO, NR_CAP);  // release_multiple :161-177
            return ia(ia_ph(1), UA_P0, NR_CAP);             // release_partial_queue :143-159
        default:  // MT_PASSIVE never becomes imminent
            return ia(0, UA_KEEP);
    }
}

struct ExtLut {
    uint32_t v[MT_COUNT * 128];
};
struct IntLut {
    uint32_t v[MT_COUNT * 32];
};
constexpr ExtLut make_ext_lut() {
    ExtLut l{};
    for (uint32_t t = 0; t < MT_COUNT; t++)
        for (uint32_t pc = 0; pc < 4; pc++)
            for (uint32_t ph = 0; ph < 4; ph++)
                for (uint32_t qs = 0; qs < 8; qs++) l.v[EXT_KEY(t, pc, ph, qs)] = ext_action(t, pc, ph, qs);
    return l;
}
constexpr IntLut make_int_lut() {
    IntLut l{};
    for (uint32_t t = 0; t < MT_COUNT; t++)
        for (uint32_t ph = 0; ph < 4; ph++)
            for (uint32_t qs = 0; qs < 8; qs++) l.v[INT_KEY(t, ph, qs)] = int_action(t, ph, qs);
    return l;
}
#ifdef __CUDACC__
static __device__ const ExtLut kExtLutDev = make_ext_lut();
static __device__ const IntLut kIntLutDev = make_int_lut();
#endif
static const ExtLut kExtLutHost = make_ext_lut();
static const IntLut kIntLutHost = make_int_lut();

// ---- model table of the generic interpreter ----------------------------------------------------------
// The per-model descriptor fields the handlers read with a PER-LANE model index, as [field][model] u32 rows of
// SIMB_MAX_MODELS entries (one 128-byte line per field) followed by the route tables; read through the L1
// (ld.global.nc).  Indexing the __grid_constant__ descriptor with a lane-varying index would serialise in the
// constant cache.  Built by build_model_table() at post / put.
enum {
    TF_TYPE = 0,   // type | n_in << 8 | n_out << 16 | flags << 24
    TF_ROWS,       // wrow | drow << 16
    TF_OUT,        // out_base | n_w << 16
    TF_WTAB,       // wtab | dist_kind << 16 | dist_err << 24
    TF_RING_CAP,
    TF_RING_W0,
    TF_RING_W1,
    TF_RING_D,
    TF_CAP,
    TF_PREFIX,
    TF_GROUP,      // group_base | coupled_of << 8 (Coupled models: the order events_ext runs in, the parked list)
    TF_RNG,        // rng_row: the model draws from its own stream (0 = the simulation's)
    TF_COUNT
};
// flags in the top byte of a route-table word (the byte that becomes the sequence number of a mailbox slot)
#define RTW_PARK (1u << 24)   /* internal coupling of a Coupled model: the job goes to its parked list */
#define RTW_CONT (1u << 25)   /* second or later delivery of ONE message to a Coupled model: not traced, not counted */
#define RTW_CPL(w) (((w) >> 26) & 3u)
#define SIMB_TAB_ROUTE_BEGIN (TF_COUNT * SIMB_MAX_MODELS)                   /* per out-port: begin | end << 16 */
#define SIMB_TAB_ROUTES (SIMB_TAB_ROUTE_BEGIN + SIMB_MAX_OUT_PORTS)           /* ROUTE_WORD(tgt, port, r, 0) */
#define SIMB_TAB_CONN (SIMB_TAB_ROUTES + SIMB_MAX_ROUTES)                        /* connector index of route r */
#define SIMB_TAB_WORDS (SIMB_TAB_CONN + SIMB_MAX_ROUTES)
// value table: [7][SIMB_MAX_MODELS] 64-bit words: p0, p1, p2 (f64 bits), q0, q1, q2, rng_seed
#define SIMB_VTAB_WORDS (7 * SIMB_MAX_MODELS)

inline void build_model_table(const SimbNetDesc& net, uint32_t* tab, uint64_t* vtab) {
    for (uint32_t i = 0; i < SIMB_TAB_WORDS; i++) tab[i] = 0;
    for (uint32_t i = 0; i < SIMB_VTAB_WORDS; i++) vtab[i] = 0;
    for (uint32_t m = 0; m < net.n_models; m++) {
        const SimbModelDesc& d = net.models[m];
        tab[TF_TYPE * SIMB_MAX_MODELS + m] = d.type | ((uint32_t)d.n_in << 8) | ((uint32_t)d.n_out << 16) | ((uint32_t)d.flags << 24);
        tab[TF_ROWS * SIMB_MAX_MODELS + m] = d.wrow | ((uint32_t)d.drow << 16);
        tab[TF_OUT * SIMB_MAX_MODELS + m] = d.out_base | ((uint32_t)d.n_w << 16);
        tab[TF_WTAB * SIMB_MAX_MODELS + m] = d.wtab | ((uint32_t)d.dist_kind << 16) | ((uint32_t)d.dist_err << 24);
        tab[TF_RING_CAP * SIMB_MAX_MODELS + m] = d.ring_cap;
        tab[TF_RING_W0 * SIMB_MAX_MODELS + m] = d.ring_w0;
        tab[TF_RING_W1 * SIMB_MAX_MODELS + m] = d.ring_w1;
        tab[TF_RING_D * SIMB_MAX_MODELS + m] = d.ring_d;
        tab[TF_CAP * SIMB_MAX_MODELS + m] = d.cap;
        tab[TF_PREFIX * SIMB_MAX_MODELS + m] = d.prefix;
        tab[TF_RNG * SIMB_MAX_MODELS + m] = d.rng_row;
        tab[TF_GROUP * SIMB_MAX_MODELS + m] = net.n_coupled ? (net.group_base[m] | ((uint32_t)net.coupled_of[m] << 8)) : (m | (0xFFu << 8));
        const double p[3] = {d.p0, d.p1, d.p2};
        for (int k = 0; k < 3; k++) {
            uint64_t b = 0;
            __builtin_memcpy(&b, &p[k], 8);
            vtab[k * SIMB_MAX_MODELS + m] = b;
        }
        vtab[3 * SIMB_MAX_MODELS + m] = d.q0;
        vtab[4 * SIMB_MAX_MODELS + m] = d.q1;
        vtab[5 * SIMB_MAX_MODELS + m] = d.q2;
        vtab[6 * SIMB_MAX_MODELS + m] = d.rng_seed;
    }
    for (uint32_t p = 0; p < SIMB_MAX_OUT_PORTS; p++)
        tab[SIMB_TAB_ROUTE_BEGIN + p] = net.route_begin[p] | ((uint32_t)net.route_begin[p + 1] << 16);
    for (uint32_t r = 0; r < net.n_routes && r < SIMB_MAX_ROUTES; r++)
    {
        const uint32_t cf = net.routes[r].connector;
        tab[SIMB_TAB_ROUTES + r] = ROUTE_WORD(net.routes[r].tgt_model, net.routes[r].tgt_port, r, 0) |
                                   ((cf & RT_PARK) ? RTW_PARK : 0u) | ((cf & RT_CONT) ? RTW_CONT : 0u) | (RT_CPL(cf) << 26);
        tab[SIMB_TAB_CONN + r] = RT_CONN(cf);
    }
}

}  // namespace simb
