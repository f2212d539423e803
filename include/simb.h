/* include/simb.h -- C ABI of libsimb200.so, the B200-native batched-replication engine for
 * ndebuhr/sim ("Sim") discrete event simulations.
 *
 * The reference has no FFI of its own (pure Rust).  Each entry point below
 * replaces one method of the reference's engine facade `Simulation` (sim/src/simulator/mod.rs) or
 * of its string facade `WebSimulation` (sim/src/simulator/web.rs) for the batched case "N
 * independent replications of one Model/Connector network stepped in lockstep", and cites the
 * reference interface it stands in for.  INTEGRATION.md shows the Rust `extern "C"` block a
 * maintainer of the reference would add to bind it.
 *
 * Conventions: plain pointers and sizes only; strings are NUL-terminated UTF-8 borrowed for the
 * duration of the call; every out-buffer is caller-allocated HOST memory with an explicit
 * capacity; a handle is not thread safe (the reference's Simulation is !Send as well,
 * sim/src/input_modeling/dynamic_rng.rs:5).  Every function returns a simb_status
 * (0 = SIMB_OK); simb_last_error_message() gives a human-readable reason for the last non-zero
 * status on the calling thread.  There is no CPU fallback: without a CUDA device every call that
 * needs one fails with SIMB_ERR_CUDA.
 */
#ifndef SIMB_H
#define SIMB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMB_ABI_VERSION 2 /* 2: simb_counters.ring_bytes */

/* ---- status codes: SimulationError variants (sim/src/utils/errors.rs:5-97) keep their names ---- */
typedef enum simb_status {
    SIMB_OK = 0,
    SIMB_ERR_INVALID_MODEL_CONFIGURATION = 1,
    SIMB_ERR_MODEL_NOT_FOUND = 2,
    SIMB_ERR_PORT_NOT_FOUND = 3,
    SIMB_ERR_MODEL_CLONE = 4,
    SIMB_ERR_INVALID_MODEL_STATE = 5,
    SIMB_ERR_EVENT_SCHEDULING = 6,
    SIMB_ERR_INVALID_MESSAGE = 7,
    SIMB_ERR_SERIALIZATION = 8,
    SIMB_ERR_EMPTY_POLYNOMIAL = 9,
    SIMB_ERR_PREREQUISITE_CALC = 10,
    SIMB_ERR_FLOAT_CONV = 11,
    SIMB_ERR_DROPPED_MESSAGE = 12,
    SIMB_ERR_JSON = 13,
    SIMB_ERR_BETA = 14,
    SIMB_ERR_EXP = 15,
    SIMB_ERR_GAMMA = 16,
    SIMB_ERR_NORMAL = 17,
    SIMB_ERR_TRIANGULAR = 18,
    SIMB_ERR_WEIBULL = 19,
    SIMB_ERR_BERNOULLI = 20,
    SIMB_ERR_GEO = 21,
    SIMB_ERR_POISSON = 22,
    SIMB_ERR_WEIGHTED = 23,
    /* no reference analogue */
    SIMB_ERR_CAPACITY = 32,          /* a bounded device queue/table/mailbox overflowed */
    SIMB_ERR_UNSUPPORTED_MODEL = 33, /* custom / Coupled model types cannot run on the device */
    SIMB_ERR_PANIC = 34,             /* the reference would panic here (e.g. Uniform low >= high) */
    SIMB_ERR_CUDA = 35,              /* CUDA runtime failure or no device */
    SIMB_ERR_INVALID_ARGUMENT = 36,
    SIMB_ERR_STREAM_EXHAUSTED = 37   /* SIMB_RNG_EXTERNAL ran out of supplied draws */
} simb_status;

/* ---- uniform u64 source per replication (DynRng, sim/src/input_modeling/dynamic_rng.rs:3-17) --- */
typedef enum simb_rng_kind {
    SIMB_RNG_PHILOX = 0,   /* Philox4x32-10, key = (replication index, seed): the engine's own streams */
    SIMB_RNG_PCG64MCG = 1, /* rand_pcg Pcg64Mcg::new(seed + replication index); replication 0 with
                              seed 42 is the reference's default stream (dynamic_rng.rs:7-9) */
    SIMB_RNG_EXTERNAL = 2  /* caller-supplied u64 draws, see simb_set_rng_stream */
} simb_rng_kind;

#define SIMB_FLAG_INSTRUMENT 1u /* per-replication trace hash, per-connector message counts,
                                   per-(model,action) record counts, trace capture */

typedef struct simb_config {
    uint32_t struct_size;        /* sizeof(simb_config), for ABI evolution */
    uint32_t flags;              /* SIMB_FLAG_* */
    uint64_t replications;       /* replications held by THIS handle (this GPU's shard) */
    uint64_t replication_offset; /* global index of local replication 0 (RNG key = global index) */
    uint64_t seed;
    int32_t rng_kind;            /* simb_rng_kind */
    int32_t device;              /* CUDA device ordinal */
    uint32_t steps_per_launch;   /* K: Simulation::step()s fused per state load (0 -> default) */
    uint32_t queue_capacity;     /* device ring slots for queues the reference leaves unbounded (0 -> 64) */
    uint32_t max_pending;        /* mailbox slots per replication (0 -> derived from the topology) */
    uint32_t trace_replications; /* leading replications whose full event trace is captured */
    uint32_t trace_capacity;     /* events kept per traced replication (0 -> 4096) */
    uint32_t reserved0;
    const char* stat_model_id;   /* Processor whose per-job waiting time is accumulated, or NULL */
    uint32_t batch_warmup;       /* streaming batch means of that statistic: jobs deleted at the start ... */
    uint32_t batch_size;         /* ... and jobs per batch (0 = off); at most 30 batches per replication */
    uint32_t series_capacity;    /* waiting times of stat_model_id kept per replication for simb_steady_state
                                    (0 = off; costs 8 x series_capacity bytes of HBM per replication) */
    uint32_t reserved1;
} simb_config;

typedef struct simb_sim simb_sim; /* opaque handle: owns all device memory of one batched Simulation */

/* Message (sim/src/simulator/coupling.rs:64-71) rendered for one replication */
typedef struct simb_message {
    char source_id[48], source_port[48], target_id[48], target_port[48], content[48];
    double time;
} simb_message;

/* ModelRecord (sim/src/models/mod.rs:48-52) */
typedef struct simb_record {
    double time;
    char action[24];
    char subject[104];
} simb_record;

/* one captured trace event (kind 1 = model record, kind 0 = routed message) */
typedef struct simb_trace_event {
    uint32_t step;    /* 1-based Simulation::step index */
    uint8_t kind;
    uint8_t action;   /* record: action code (SIMB_ACTION_*), message: 0 */
    uint8_t aux;      /* record: 1 + port index when the subject carries " on <port>", else 0 */
    uint8_t reserved;
    uint16_t index;   /* record: model index, message: connector index */
    uint16_t reserved1;
    uint32_t content; /* interned content code, see simb_render_content */
    double time;
} simb_trace_event;

enum {
    SIMB_ACTION_INITIALIZATION = 0, SIMB_ACTION_GENERATION, SIMB_ACTION_ARRIVAL,
    SIMB_ACTION_PROCESSING_START, SIMB_ACTION_DROP, SIMB_ACTION_DEPARTURE, SIMB_ACTION_ACTIVATION,
    SIMB_ACTION_DEACTIVATION, SIMB_ACTION_PASS, SIMB_ACTION_BLOCK, SIMB_ACTION_START,
    SIMB_ACTION_STOP, SIMB_ACTION_MINIMUM_FETCH, SIMB_ACTION_MAXIMUM_FETCH, SIMB_ACTION_COUNT
};

typedef struct simb_counters {
    uint64_t steps;          /* sum over replications of Simulation::step calls executed */
    uint64_t events;         /* sum of events_ext + events_int calls (mod.rs:220,241) */
    uint64_t launches;       /* step-kernel launches issued by this handle */
    double kernel_ms;        /* device time of those launches (CUDA events on the handle's stream) */
    uint64_t state_bytes_per_replication; /* bytes of the SoA state tile one launch loads (= stores) */
    uint64_t failed;         /* replications holding a sticky error */
    uint64_t ring_bytes;     /* bytes of queue / table slots (and mailbox overflow) the step kernels touched in
                                global memory since the last simb_reset_counters: counted on the device, the
                                second term of a launch's algorithmic bytes next to 2 x state bytes */
} simb_counters;

/* SteadyStateOutput of one replication (sim/src/output_analysis/mod.rs:205-345) */
typedef struct simb_steady_state_result {
    double mean, variance, lower, upper; /* batches mean / population variance, confidence interval of the mean */
    uint32_t deletion_point, batch_size, batch_count;
    int32_t status;                      /* 0, or the simb_status the reference's calls would have produced */
} simb_steady_state_result;

typedef struct simb_ci {
    double mean, variance, lower, upper;
    uint64_t n;
} simb_ci;

const char* simb_last_error_message(void);
const char* simb_status_name(int status);
int simb_abi_version(void);
int simb_device_count(int* count);

/* ---- lifecycle ---------------------------------------------------------------------------------
 * simb_post_json   <- Simulation::post (mod.rs:49-56) via the JSON form of WebSimulation::post_json
 *                     (web.rs:23-31); schema = the reference's serde attributes.
 * simb_put_json    <- Simulation::put (mod.rs:82-85): fresh model state, keeps clock, messages, RNG.
 * simb_reset*      <- reset / reset_messages / reset_global_time (mod.rs:131-144).
 * simb_set_rng     <- post_with_rng / set_rng (mod.rs:60-79).
 * simb_set_rng_stream: SIMB_RNG_EXTERNAL draws, draws[i * replications + r] = i-th u64 of local
 *                     replication r ("fed the reference's own uniform draw stream").
 */
int simb_post_json(const char* models_json, const char* connectors_json, const simb_config* config,
                   simb_sim** out);
int simb_put_json(simb_sim* sim, const char* models_json, const char* connectors_json);
int simb_reset(simb_sim* sim);
int simb_reset_messages(simb_sim* sim);
int simb_reset_global_time(simb_sim* sim);
int simb_set_rng(simb_sim* sim, int rng_kind, uint64_t seed);
int simb_set_rng_stream(simb_sim* sim, const uint64_t* draws, uint64_t draws_per_replication);
int simb_destroy(simb_sim* sim);

/* ---- stepping ----------------------------------------------------------------------------------
 * simb_inject_input <- Simulation::inject_input (mod.rs:189-191); mask = one byte per local
 *                      replication (non-zero = inject) or NULL for all.
 * simb_step         <- Simulation::step (mod.rs:198-272), once on every live replication.
 * simb_step_n       <- Simulation::step_n (mod.rs:293-303).
 * simb_step_until   <- Simulation::step_until (mod.rs:277-288): each replication steps until ITS
 *                      clock reaches `until` (the crossing step is executed, as in the reference).
 * The returned Vec<Message> of the reference cannot be materialised for 10^6 replications; the
 * calls return a status and the results are read back through the accessors below.  A
 * replication whose step returned Err keeps that error (sticky) and stops stepping; the call then
 * returns the error of the lowest-indexed failed replication.
 */
int simb_inject_input(simb_sim* sim, const char* target_id, const char* target_port,
                      const char* content, const uint8_t* replica_mask);
int simb_step(simb_sim* sim);
int simb_step_n(simb_sim* sim, uint64_t n);
int simb_step_until(simb_sim* sim, double until);
int simb_synchronize(simb_sim* sim);

/* ---- observation -------------------------------------------------------------------------------
 * simb_get_global_time <- get_global_time (mod.rs:99), one f64 per local replication.
 * simb_get_messages    <- get_messages (mod.rs:94) for one replication.
 * simb_get_status      <- get_status (mod.rs:106-113) for one replication.
 * simb_get_records     <- get_records (mod.rs:118-125): available for traced replications of an
 *                         instrumented handle, filtered by the model's storeRecords flag.
 */
int simb_get_global_time(simb_sim* sim, double* out, uint64_t capacity);
int simb_get_messages(simb_sim* sim, uint64_t replication, simb_message* out, uint64_t capacity,
                      uint64_t* count);
int simb_get_status(simb_sim* sim, const char* model_id, uint64_t replication, char* out,
                    uint64_t capacity);
int simb_get_records(simb_sim* sim, const char* model_id, uint64_t replication, simb_record* out,
                     uint64_t capacity, uint64_t* count);
int simb_get_until_next_event(simb_sim* sim, const char* model_id, double* out, uint64_t capacity);
int simb_get_errors(simb_sim* sim, int32_t* out, uint64_t capacity);
int simb_get_steps(simb_sim* sim, uint32_t* out, uint64_t capacity);
int simb_get_events(simb_sim* sim, uint32_t* out, uint64_t capacity);
int simb_get_draws(simb_sim* sim, uint32_t* out, uint64_t capacity);
int simb_get_counters(simb_sim* sim, simb_counters* out);
int simb_reset_counters(simb_sim* sim);

/* instrumented handles only (SIMB_FLAG_INSTRUMENT) */
int simb_get_trace_hash(simb_sim* sim, uint64_t* out, uint64_t capacity);
int simb_get_message_counts(simb_sim* sim, uint64_t* per_connector, uint64_t capacity);
int simb_get_record_counts(simb_sim* sim, uint64_t* per_model_action, uint64_t capacity); /* [model][SIMB_ACTION_COUNT] */
int simb_get_trace(simb_sim* sim, uint64_t replication, simb_trace_event* out, uint64_t capacity,
                   uint64_t* count);
int simb_render_content(simb_sim* sim, uint32_t content, char* out, uint64_t capacity);
int simb_intern_content(simb_sim* sim, const char* content, uint32_t* out);

/* ---- string API (WebSimulation, sim/src/simulator/web.rs:87-215) --------------------------------
 * The JSON forms of the calls above for ONE observed replication of the batch (every replication
 * still steps).  *json points at a NUL-terminated string owned by the handle, valid until the next
 * *_json call on it.  Text is what serde_json::to_string produces: Vec<Message> with the camelCase
 * keys of coupling.rs:62-71, Vec<ModelRecord> as models/mod.rs:48-52, f64 in shortest round-trip
 * form.  The stepping forms need an instrumented handle whose trace covers the replication
 * (simb_config.trace_replications / trace_capacity).
 * simb_get_messages_json <- get_messages_json (web.rs:87-89)
 * simb_get_records_json  <- get_records_json (web.rs:109-111)
 * simb_inject_input_json <- inject_input_json (web.rs:136-139); message = one serialised Message
 * simb_step_json         <- step_json (web.rs:161-163): the messages Simulation::step returns
 * simb_step_n_json       <- step_n_json (web.rs:207-209): messages of all n steps (mod.rs:293-303)
 * simb_step_until_json   <- step_until_json (web.rs:184-186): messages of the steps that ended
 *                           before `until`; the crossing step's are not returned (mod.rs:277-288)
 * simb_format_f64: the number formatter alone (for tests of the text form)
 */
int simb_get_messages_json(simb_sim* sim, uint64_t replication, const char** json);
int simb_get_records_json(simb_sim* sim, const char* model_id, uint64_t replication, const char** json);
int simb_inject_input_json(simb_sim* sim, const char* message_json, const uint8_t* replica_mask);
int simb_step_json(simb_sim* sim, uint64_t replication, const char** json);
int simb_step_n_json(simb_sim* sim, uint64_t n, uint64_t replication, const char** json);
int simb_step_until_json(simb_sim* sim, double until, uint64_t replication, const char** json);
int simb_format_f64(double value, char* out, uint64_t capacity);

/* ---- whole-simulation documents and checkpoints ------------------------------------------------
 * simb_get_json  <- WebSimulation::get_json (web.rs:43-45): serde_json::to_string_pretty of the Simulation
 *                   { models (each with its live `state` block), connectors, messages, services.globalTime }
 *                   (mod.rs:37-45) of ONE replication of the batch.  The models and connectors of the document can
 *                   be posted again (simb_post_json); like the reference's, the document does not carry the RNG.
 * simb_get_yaml  <- get_yaml (web.rs:69-71), the same document in serde_yaml 0.8 layout.
 * simb_post_yaml <- post_yaml (web.rs:56-64), simb_put_yaml <- put_yaml (web.rs:66-67).
 * simb_yaml_to_json / simb_json_to_yaml: the converters behind the other *_yaml forms of web.rs:91-215
 *                   (get_messages_yaml, get_records_yaml, inject_input_yaml, step*_yaml); *needed receives the
 *                   size of the result including the terminator.
 * simb_state_size / simb_get_state / simb_put_state: the device state of the WHOLE batch (state rows, queue rings,
 *                   mailboxes, RNG positions, statistics, trace buffers, counters) as one opaque blob for exact
 *                   checkpoint / resume: a handle posted with the same models, connectors and simb_config continues
 *                   bit for bit after simb_put_state.  (The reference's serde form skips the RNG, services.rs:11;
 *                   the counter-based streams make the exact form cheap here.)
 */
int simb_get_json(simb_sim* sim, uint64_t replication, const char** json);
int simb_get_yaml(simb_sim* sim, uint64_t replication, const char** yaml);
int simb_post_yaml(const char* models_yaml, const char* connectors_yaml, const simb_config* config, simb_sim** out);
int simb_put_yaml(simb_sim* sim, const char* models_yaml, const char* connectors_yaml);
int simb_yaml_to_json(const char* yaml, char* out, uint64_t capacity, uint64_t* needed);
int simb_json_to_yaml(const char* json, char* out, uint64_t capacity, uint64_t* needed);
int simb_state_size(simb_sim* sim, uint64_t* bytes);
int simb_get_state(simb_sim* sim, void* out, uint64_t capacity, uint64_t* written);
int simb_put_state(simb_sim* sim, const void* data, uint64_t bytes);

/* ---- output_analysis (sim/src/output_analysis/mod.rs) ------------------------------------------
 * Device side: per-replication waiting-time accumulators of `stat_model_id` (the statistic of
 * sim/tests/web.rs:571-597: Processing Start - Arrival per job) and their reduction.
 * simb_get_wait_stats           per-replication (sum, count)
 * simb_stat_partial / _sqdev    this shard's partial sums for a multi-GPU all-reduce: count of
 *                               replications with a defined mean, sum of means, and (second pass,
 *                               given the global mean) sum of squared deviations -- the two-pass
 *                               population variance of output_analysis/mod.rs:23-40.
 * simb_independent_sample_ci    single-GPU convenience: IndependentSample::post + confidence_interval_mean
 *                               (mod.rs:94-125) over the replication means.
 * Host side restatements over caller data: simb_oa_* (IndependentSample, SteadyStateOutput
 * mod.rs:205-345, t_score t_scores.rs:9-30, usize_sqrt utils/mod.rs:84-92).
 */
int simb_get_wait_stats(simb_sim* sim, double* sum, uint32_t* count, uint64_t capacity);
/* per-replication streaming batch means (simb_config.batch_size > 0): mean and population variance of the
 * completed batch means and their number (<= 30) -- the steady-state statistic of BASELINE config 3 with a
 * fixed budget instead of SteadyStateOutput's series-dependent deletion point (mod.rs:224-260) */
int simb_get_batch_means(simb_sim* sim, double* mean, double* variance, uint32_t* batches, uint64_t capacity);
/* SteadyStateOutput::confidence_interval_mean given (batches_mean, batches_variance, batch_count): note the
 * reference's asymmetric degrees of freedom (mod.rs:327 vs :330) */
int simb_oa_batch_means_ci(double mean, double variance, uint64_t batch_count, double alpha, simb_ci* out);
/* SteadyStateOutput::{post, confidence_interval_mean} (mod.rs:205-333) on the device, one result per replication,
 * over the waiting-time series captured with simb_config.series_capacity: the reference's own deletion point
 * (set_to_fixed_budget, mod.rs:224-260), <= 30 batches, asymmetric degrees of freedom.  Bit-equal to
 * simb_oa_steady_state over simb_get_series.  A replication whose series outgrew the capacity reports
 * SIMB_ERR_CAPACITY in its status. */
int simb_steady_state(simb_sim* sim, double alpha, simb_steady_state_result* out, uint64_t capacity);
int simb_get_series(simb_sim* sim, uint64_t replication, double* out, uint64_t capacity, uint64_t* count);
int simb_stat_partial(simb_sim* sim, double* sum_of_means, uint64_t* n);
int simb_stat_partial_sqdev(simb_sim* sim, double mean, double* sum_sqdev);
int simb_independent_sample_ci(simb_sim* sim, double alpha, simb_ci* out);
int simb_oa_confidence_interval(double mean, double variance, uint64_t n, double alpha, simb_ci* out);
int simb_oa_independent_sample(const double* points, uint64_t n, double alpha, simb_ci* out);
int simb_oa_steady_state(const double* series, uint64_t n, double alpha, simb_ci* out,
                         uint64_t* deletion_point, uint64_t* batch_size, uint64_t* batch_count);
int simb_oa_t_score(double alpha, uint64_t df, double* out);
uint64_t simb_usize_sqrt(uint64_t n);

/* ---- input_modeling (sim/src/input_modeling) ------------------------------------------------------
 * simb_sample_random_variable <- {Continuous, Boolean, Discrete, Index}::random_variate (random_variable.rs:65-131) on
 *     the device: `count` variates per replication of ONE random variable, given in the reference's JSON form
 *     (e.g. {"exp":{"lambda":0.5}}, {"bernoulli":{"p":0.3}}, {"poisson":{"lambda":7.0}}, {"geometric":{"p":0.2}},
 *     {"uniform":{"min":7,"max":11}}, {"weightedIndex":{"weights":[1,2,3,4]}}) with family = "continuous" | "boolean" |
 *     "discrete" | "index", drawn from the per-replication Philox streams the engine uses (key = global replication
 *     index, seed; draws numbered from 0).  out[i * replications + r]: continuous -> out_f64, all others -> out_u64
 *     (booleans 0 / 1); draws_consumed[r] (may be NULL) = u64 draws replication r used.  Parameter errors are
 *     reported as the reference raises them, at draw time (BetaError ... WeightedError, Panic for Uniform low >= high).
 * simb_thinning_evaluate      <- Thinning::evaluate (thinning.rs:27-34): the normalised polynomial thinning function,
 *     coefficients from the highest order down (utils/mod.rs:34-41).  (The reference's Generator stores a Thinning and
 *     never evaluates it, generator.rs:98-146; posted models may carry the field, the engine ignores it the same way.)
 */
int simb_sample_random_variable(const char* family, const char* variable_json, uint64_t replications, uint64_t count,
                                uint64_t seed, uint64_t replication_offset, int device, double* out_f64,
                                uint64_t* out_u64, uint32_t* draws_consumed);
int simb_thinning_evaluate(const double* coefficients, uint64_t n, double point, double* out);

/* ---- introspection -----------------------------------------------------------------------------*/
int simb_describe(simb_sim* sim, char* out, uint64_t capacity); /* JSON: layout rows, rings, K, TPB ... */
int simb_model_count(simb_sim* sim, uint64_t* models, uint64_t* connectors);

#ifdef __cplusplus
}
#endif
#endif /* SIMB_H */
