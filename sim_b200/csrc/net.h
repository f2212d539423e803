// sim_b200/csrc/net.h -- lowered network descriptor shared by the host lowering (lower.cpp) and
// the device interpreter (core.cuh).  Plain-old-data: the whole NetDesc is passed to the step
// kernel as a __grid_constant__ parameter, i.e. it sits in the constant bank for the launch
// ("network topology in constant memory") without any mutable module-level state, so several
// handles with different topologies can coexist in one process.
#pragma once
#include <stdint.h>

#define SIMB_MAX_MODELS 32
#define SIMB_MAX_ROUTES 160     /* < 255: a route index is one byte of the mailbox route word */
#define SIMB_MAX_OUT_PORTS 96   /* total out-ports over all models */
#define SIMB_MAX_WEIGHTS 64     /* total cumulative-weight entries over all ExclusiveGateways */
#define SIMB_MAX_NAMES 255      /* content prefix / literal intern table */
#ifndef SIMB_MAIL_TILE_SLOTS
#define SIMB_MAX_BATCHES 30      /* Schmeiser's k = 30 (sim/src/output_analysis/mod.rs:246-253) */
#define SIMB_MAIL_TILE_SLOTS 8  /* mailbox slots staged in shared memory; deeper slots (rare bursts) stay in HBM */
#endif

enum SimbModelType : uint8_t {
    MT_GENERATOR = 0,        // sim/src/models/generator.rs
    MT_PROCESSOR = 1,        // processor.rs
    MT_STORAGE = 2,          // storage.rs
    MT_LOAD_BALANCER = 3,    // load_balancer.rs
    MT_GATE = 4,             // gate.rs
    MT_EXCLUSIVE_GATEWAY = 5,  // exclusive_gateway.rs
    MT_PARALLEL_GATEWAY = 6,   // parallel_gateway.rs
    MT_STOCHASTIC_GATE = 7,    // stochastic_gate.rs
    MT_STOPWATCH = 8,        // stopwatch.rs
    MT_BATCHER = 9,          // batcher.rs
    MT_PASSIVE = 10,         // sim/tests/custom.rs:18-86 (zero-behaviour sink)
    MT_COUNT = 11
};

enum SimbDistKind : uint8_t {
    DK_EXP = 0, DK_NORMAL = 1, DK_TRIANGULAR = 2, DK_UNIFORM = 3,
    DK_GAMMA = 4,      // p0 = d, p1 = c, p2 = scale; q0 = 0 large shape / 1 small shape (q1 = bits of 1/shape)
    DK_LOGNORMAL = 5,  // p0 = mu, p1 = sigma
    DK_WEIBULL = 6,    // p0 = scale, p1 = 1/shape (rand_distr's roles, i.e. the reference's swapped ones)
    DK_BETA = 7,       // p0 = a, p1 = b as rand_distr's Beta keeps them; q0 bit 0: algorithm BC, bit 1: switched_params
    // the Discrete variables (random_variable.rs:36-50): no built-in model draws them; simb_sample_random_variable does
    DK_GEOMETRIC = 16,    // p0 = p, p1 = pi = (1-p)^(2^k), q0 = k
    DK_POISSON = 17,      // p0 = lambda, p1 = exp(-lambda), p2 = ln lambda, q0 = bits of sqrt(2 lambda), q1 = bits of the
                          // "magic value" lambda ln lambda - log_gamma(1 + lambda)
    DK_UNIFORM_INT = 18,  // q0 = low, q1 = range, q2 = zone (as Index::Uniform)
    DK_NONE = 255
};
// random-variable families of simb_sample_random_variable (the four enums of random_variable.rs:17-63)
enum SimbFamily : uint32_t { FAM_CONTINUOUS = 0, FAM_BOOLEAN = 1, FAM_DISCRETE = 2, FAM_INDEX = 3 };

// ModelDesc.flags
#define MF_STORE_RECORDS 1u
#define MF_STAT 2u        /* Processor: accumulate waiting times (Processing Start - Arrival) */
#define MF_METRIC_MAX 4u  /* Stopwatch metric = Maximum */
#define MF_INDEX_WEIGHTED 8u
#define MF_BERNOULLI_ALWAYS 16u /* p == 1.0: true without drawing */

#define SIMB_PORT_UNKNOWN 0xFFu  /* target port name not among the model's input ports */
#define SIMB_CONTENT_NONE 0xFFFFFFFFu
#define SIMB_NUM_LITERAL 0xFFFFFFu
#define SIMB_CAP_INF 0xFFFFFFFFu

struct SimbModelDesc {
    uint8_t type, n_in, n_out, flags;
    uint8_t dist_kind;
    uint8_t dist_err;   // simb_status to raise at draw time when the parameters are invalid (the
                        // reference constructs the distribution per draw, random_variable.rs:72-87)
    uint16_t wrow;      // first u32 state row private to this model
    uint16_t drow;      // first private f64 state row (Stopwatch best duration), else 0
    uint16_t out_base;  // index of out-port 0 in NetDesc.route_begin
    uint16_t wtab;      // ExclusiveGateway: first entry in NetDesc.wcum
    uint16_t n_w;       // ExclusiveGateway: number of cumulative weights (= #weights - 1)
    uint32_t ring_cap;  // device slots of this model's queue / table (0 = none)
    uint32_t ring_w0;   // slot offset of u32 channel 0 (content) in the ring_w array
    uint32_t ring_w1;   // slot offset of u32 channel 1 (count / pass flag / flags|seq)
    uint32_t ring_d;    // slot offset of the f64 channel (arrival / start time) in ring_d
    uint32_t cap;       // Processor queue_capacity (SIMB_CAP_INF = usize::MAX), Batcher max_batch_size
    uint32_t prefix;    // Generator: (intern id of portsOut.job) << 24
    uint16_t rng_row;   // u32 state row holding the draw index of the model's OWN Philox stream (`rng: Some(..)` of
                        // generator.rs:39 etc.), 0 = the model draws from the simulation's stream
    uint16_t reserved;
    double p0, p1, p2;  // Exp: p0 = 1/lambda; Normal: mean, std_dev; Triangular: min, max, mode;
                        // Uniform: low, scale (already ulp-decremented); Batcher: p0 = max_batch_time
    uint64_t q0, q1, q2;  // Bernoulli: q0 = p_int; Index: q0 = low, q1 = range, q2 = zone
    uint64_t rng_seed;    // seed of the model's own stream (rng_row != 0): Philox key, counter = (block, replication)
};

struct SimbRoute {
    uint8_t tgt_model;
    uint8_t tgt_port;     // index among the target's input ports, SIMB_PORT_UNKNOWN if absent
    uint16_t connector;   // connector index (RT_CONN) and, for networks with Coupled models, the RT_* flags
};
// Coupled models (sim/src/models/coupled.rs) are lowered into the flat model list: components sit where the coupled
// model stood, and its couplings become routes.  A component's out-port route is either a top-level message (over an
// external output coupling and a connector; a Coupled TARGET is expanded over its external input couplings, the
// second and later copies flagged RT_CONT: one Message of the reference, several deliveries) or RT_PARK: an internal
// coupling, kept in the coupled model's parked list until that model next fires (coupled.rs:186-255).
#define RT_CONN(c) ((c) & 0x0FFFu)
#define RT_CPL(c) (((c) >> 12) & 3u)
#define RT_CONT 0x4000u
#define RT_PARK 0x8000u
#define SIMB_MAX_COUPLED 4
struct SimbCoupledDesc {
    uint8_t first, count;  // its components: flat models [first, first + count)
    uint16_t wrow;         // u32 state row: number of parked messages
    uint32_t ring_cap;     // parked-message slots
    uint32_t ring_w0;      // ring_w slots: content
    uint32_t ring_w1;      // ring_w slots: target component | input port << 8
};

struct SimbNetDesc {
    uint32_t n_models, n_routes, n_connectors, max_pending;
    uint32_t bm_warmup, bm_size;  // streaming batch means of the MF_STAT waiting times: jobs deleted at the
                                  // start, jobs per batch (0 = off); at most SIMB_MAX_BATCHES batches
    // state tile layout (rows of N replications each)
    uint16_t n_drows, n_wrows;
    uint16_t drow_time;     // global_time
    uint16_t drow_until;    // until_next_event of model m = drow_until + m
    uint16_t drow_wait;     // waiting-time sum of the MF_STAT processor (0 if none)
    uint16_t drow_bm;       // 3 rows: running sum of the open batch, sum of batch means, sum of squared batch means
    uint16_t wrow_ctl;      // err(8) | epoch(8) | npend(8) | absent(1)
    uint16_t wrow_rng;      // draw index (Philox/external) or 4 words of PCG state
    uint16_t wrow_steps, wrow_events;
    uint16_t wrow_hash;     // 2 words (instrumented handles only; else 0)
    uint16_t wrow_phase;    // 2 bits per model
    uint16_t n_phase_words;
    uint16_t wrow_pend;     // pend_tile x (content, route)
    uint16_t wrow_waitn;    // number of jobs that started service on the MF_STAT processor
    uint16_t uses_exp, uses_normal;
    uint16_t wrow_dmask;    // models whose until_next_event still owes a deferred draw (core.cuh)
    uint16_t pend_tile;     // mailbox slots kept in the state tile (min(max_pending, SIMB_MAIL_TILE_SLOTS));
                            // the rest live in the global overflow array SimbDevState.mail
    SimbModelDesc models[SIMB_MAX_MODELS];
    uint16_t route_begin[SIMB_MAX_OUT_PORTS + 1];
    SimbRoute routes[SIMB_MAX_ROUTES];
    uint64_t wcum[SIMB_MAX_WEIGHTS];
    // Coupled models (appended: the baked shapes' aggregate initialisers leave these zero)
    uint32_t n_coupled;
    SimbCoupledDesc coupled[SIMB_MAX_COUPLED];
    uint8_t group_base[SIMB_MAX_MODELS];  // flat index of the first model of the top-level entry a model belongs to
                                          // (itself for an atomic top-level model): the order events_ext runs in
    uint8_t coupled_of[SIMB_MAX_MODELS];  // coupled model a component belongs to, 0xFF for a top-level atomic model
};

// ctl word fields
#define CTL_ERR(c) ((c) & 0xFFu)
#define CTL_EPOCH(c) (((c) >> 8) & 0xFFu)
#define CTL_NPEND(c) (((c) >> 16) & 0xFFu)
#define CTL_ABSENT(c) (((c) >> 24) & 1u)
#define CTL_MAKE(err, epoch, npend, absent) \
    (((uint32_t)(err) & 0xFFu) | (((uint32_t)(epoch) & 0xFFu) << 8) | (((uint32_t)(npend) & 0xFFu) << 16) | (((uint32_t)(absent) & 1u) << 24))

// route word in a mailbox slot: target model (0xFF = no such model), index among the target's input ports,
// route index (into SimbNetDesc.routes; ROUTE_INJECTED for Simulation::inject_input) and the emission sequence
// number within the step.  The generic interpreter keeps the mailbox SORTED by target model (stable), which is
// the order events_ext is called in (mod.rs:203-221); the sequence number restores the reference's message
// order for get_messages.
#define ROUTE_INJECTED 0xFFu
#define ROUTE_INJECTED_CONT 0xFEu /* further deliveries of one injected message to a Coupled model (not shown) */
#define ROUTE_WORD(tgt, port, r, seq) \
    (((uint32_t)(tgt) & 0xFFu) | (((uint32_t)(port) & 0xFFu) << 8) | (((uint32_t)(r) & 0xFFu) << 16) | (((uint32_t)(seq) & 0xFFu) << 24))
#define ROUTE_TGT(w) ((w) & 0xFFu)
#define ROUTE_PORT(w) (((w) >> 8) & 0xFFu)
#define ROUTE_R(w) (((w) >> 16) & 0xFFu)
#define ROUTE_SEQ(w) ((w) >> 24)

// trace word (shared definition with the oracle, DESIGN.md "trace hash")
#define TRACE_RECORD_WORD(model, action, aux, content) \
    ((1ull << 63) | ((uint64_t)(model) << 48) | ((uint64_t)(action) << 40) | ((uint64_t)(aux) << 32) | (uint64_t)(uint32_t)(content))
#define TRACE_MESSAGE_WORD(connector, content) (((uint64_t)(connector) << 32) | (uint64_t)(uint32_t)(content))
#define TRACE_HASH_INIT 0xcbf29ce484222325ull
#define TRACE_HASH_PRIME 0x100000001B3ull

enum SimbAction : uint8_t {
    ACT_INITIALIZATION = 0, ACT_GENERATION, ACT_ARRIVAL, ACT_PROCESSING_START, ACT_DROP, ACT_DEPARTURE,
    ACT_ACTIVATION, ACT_DEACTIVATION, ACT_PASS, ACT_BLOCK, ACT_START, ACT_STOP, ACT_MINIMUM_FETCH,
    ACT_MAXIMUM_FETCH, ACT_COUNT
};

// device trace record (24 B)
struct SimbTraceRec {
    uint32_t step;
    uint32_t meta;     // kind(1)<<31 | index(16)<<8 ... see pack below
    uint32_t content;
    uint32_t aux;      // action(8) | aux(8)<<8
    double time;
};

// device-side view of one handle's arrays
struct SimbDevState {
    double* d;            // [n_drows][stride]
    uint32_t* w;          // [n_wrows][stride]
    uint32_t* ring_w;     // [ring_w_slots][stride]
    double* ring_d;       // [ring_d_slots][stride]
    uint32_t* mail;       // mailbox overflow: [(max_pending - pend_tile) x (content, route)][stride]
    uint64_t stride;      // padded replication count (multiple of the tile width)
    uint64_t n;           // real replication count
    uint64_t rep_offset;  // global index of local replication 0
    uint64_t seed;
    const uint32_t* tab;   // generic interpreter: model / route table (actions.h, SIMB_TAB_WORDS u32)
    const uint64_t* vtab;  // generic interpreter: distribution value table (SIMB_VTAB_WORDS 64-bit words)
    const uint64_t* ext;  // SIMB_RNG_EXTERNAL: [ext_len][stride]
    uint32_t ext_len;
    uint32_t trace_reps, trace_cap;
    SimbTraceRec* trace;          // [trace_reps][trace_cap]
    uint32_t* trace_count;        // [trace_reps]
    unsigned long long* conn_count;    // [n_connectors]        (instrumented)
    unsigned long long* record_count;  // [n_models][ACT_COUNT] (instrumented)
    unsigned long long* totals;        // [5]: steps, events, still-active replications, failed, ring bytes touched
    double* series;       // waiting-time series of the MF_STAT processor, [series_cap][stride], or null
    uint32_t series_cap;  // samples kept per replication (simb_config.series_capacity)
    uint32_t reserved;
};

// SteadyStateOutput of one replication's series, computed on the device (simb_steady_state_kernel)
struct SimbSteadyState {
    double mean, variance, lower, upper;
    uint32_t deletion_point, batch_size, batch_count;
    int32_t status;
};

struct SimbLaunch {
    uint32_t mode;        // 0 = step_n: run `k` steps; 1 = step_until: run up to `k` steps while clock < until
    uint32_t k;
    uint32_t epoch;       // step_until call epoch (ctl epoch == epoch => finished in this call)
    uint32_t settle;      // 1 = pay every deferred draw before the state is stored (last launch of an API call)
    uint32_t tile0;       // first tile of this grid (a pass over the batch may be split into concurrent grids)
    uint32_t rng_slots;   // baked-shape kernels: Philox draws cached per replication in shared memory (0 = none)
    uint32_t reverse;     // 1 = CTA i takes the grid's tile (gridDim.x - 1 - i): consecutive passes sweep the batch in
                          // opposite directions, so a pass starts on the tiles the previous one left in L2
    uint32_t l2_keep;     // 1 = the state tile copies carry the L2 evict_last hint (streaming passes)
    uint32_t lockstep;    // generic interpreter, A/B switch: 1 = the lanes of a warp wait for each other at every step
                          // boundary (0 = each lane runs through its own steps)
    uint32_t zig_global;  // generic interpreter: the ziggurat x tables are read through the L1 instead of being staged in
                          // shared memory (EngineDims::zig_global: chosen when it lets one more CTA stay resident)
    double until;
};
