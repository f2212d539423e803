// oracle/oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17, single thread per replication) of the hot path of ndebuhr/sim that
// sim_b200 accelerates: the DEVS coordinator step, the ten built-in atomic models, the
// input_modeling samplers and the output_analysis statistics.  Every function cites the
// reference file:line it follows (paths relative to /root/reference/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product library (libsimb200.so) never links or calls it.
//
// PARITY STATUS: the engine/model/statistics semantics are pinned by the reference's own
// structural tests and known-answer tests (tests/test_oracle_*.py).  The sampling arithmetic lives
// in crates.io dependencies that are not in the reference tree (rand 0.8, rand_distr 0.4,
// rand_pcg 0.3; sim/Cargo.toml:25-28, no Cargo.lock) and is restated from their published
// algorithms; the reference pins it only statistically => "parity unpinned" at that boundary.
#pragma once
#include <cstdint>
#include <cmath>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace orc {

// ---- SimulationError (sim/src/utils/errors.rs:5-97); numeric codes are this project's ----------
enum Err : int {
    OK = 0,
    InvalidModelConfiguration = 1,
    ModelNotFound = 2,
    PortNotFound = 3,
    ModelCloneError = 4,
    InvalidModelState = 5,
    EventSchedulingError = 6,
    InvalidMessage = 7,
    SerializationError = 8,
    EmptyPolynomial = 9,
    PrerequisiteCalcError = 10,
    FloatConvError = 11,
    DroppedMessageError = 12,
    JSONError = 13,
    BetaError = 14,
    ExpError = 15,
    GammaError = 16,
    NormalError = 17,
    TriangularError = 18,
    WeibullError = 19,
    BernoulliError = 20,
    GeoError = 21,
    PoissonError = 22,
    WeightedError = 23,
    // not SimulationError variants: places where the reference panics instead of returning Err
    Panic = 34,
};

// ---- uniform u64 sources (sim/src/input_modeling/dynamic_rng.rs:3-17) ---------------------------
struct Rng {
    virtual ~Rng() {}
    virtual uint64_t next_u64() = 0;
    uint64_t draws = 0;  // number of u64 consumed so far (bookkeeping for tests)
};

// rand_pcg 0.3 Pcg64Mcg == Mcg128Xsl64; the reference default is Pcg64Mcg::new(42)
// (dynamic_rng.rs:7-9).
struct Pcg64Mcg : Rng {
    unsigned __int128 state;
    explicit Pcg64Mcg(unsigned __int128 seed) : state(seed | 1) {}
    uint64_t next_u64() override;
};

// Philox4x32-10 counter stream, one per replication: key = the 64-bit seed, counter = (block, replication).
// u64 number i is words (2*(i&1), 2*(i&1)+1) of block i>>1.  This is the batched engine's own
// per-replication source (north_star); the oracle restates it so both sides can run identical
// replications.
struct Philox : Rng {
    uint32_t key0, key1;
    uint64_t rep = 0;
    uint64_t index = 0;
    Philox(uint64_t seed, uint64_t replication);
    uint64_t next_u64() override;
    static void block(uint32_t key0, uint32_t key1, uint64_t ctr, uint64_t replication, uint32_t out[4]);
    static void block_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
};

// A caller-supplied list of u64 ("fed the reference's own uniform draw stream").
struct ExternalStream : Rng {
    std::vector<uint64_t> values;
    size_t pos = 0;
    bool exhausted = false;
    uint64_t next_u64() override;
};

// ---- random variables (sim/src/input_modeling/random_variable.rs:17-131) -----------------------
enum DistKind : int {
    D_EXP = 0,          // Continuous::Exp{lambda}
    D_NORMAL = 1,       // Continuous::Normal{mean,std_dev}
    D_TRIANGULAR = 2,   // Continuous::Triangular{min,max,mode}
    D_UNIFORM = 3,      // Continuous::Uniform{min,max}
    D_GAMMA = 4,        // Continuous::Gamma{shape,scale}      a=shape, b=scale
    D_LOGNORMAL = 5,    // Continuous::LogNormal{mu,sigma}     a=mu, b=sigma
    D_WEIBULL = 6,      // Continuous::Weibull{shape,scale}    a=shape, b=scale (passed to rand_distr SWAPPED)
    D_BETA = 7,         // Continuous::Beta{alpha,beta}        a=alpha, b=beta
};
struct Continuous {
    int kind = D_EXP;
    double a = 1.0, b = 0.0, c = 0.0;  // Exp: a=lambda; Normal: a=mean,b=std_dev; Tri: a=min,b=max,c=mode; Uniform: a=min,b=max
    int random_variate(Rng& rng, double* out) const;  // random_variable.rs:69-89
};
struct Boolean {  // Boolean::Bernoulli{p}
    double p = 0.5;
    int random_variate(Rng& rng, bool* out) const;  // random_variable.rs:96-101
};
struct Index {  // Index::Uniform{min,max} | Index::WeightedIndex{weights}
    bool weighted = false;
    uint64_t min = 0, max = 1;
    std::vector<uint64_t> weights;
    int random_variate(Rng& rng, uint64_t* out) const;  // random_variable.rs:122-130
};

struct Discrete {  // Discrete::Geometric{p} | Poisson{lambda} | Uniform{min,max}  (random_variable.rs:36-50)
    int kind = 0;  // 0 Geometric, 1 Poisson, 2 Uniform
    double p = 0.5;  // Geometric p / Poisson lambda
    uint64_t min = 0, max = 1;
    int random_variate(Rng& rng, uint64_t* out) const;  // random_variable.rs:108-115
};
// Thinning::evaluate (input_modeling/thinning.rs:27-34) = utils::evaluate_polynomial (utils/mod.rs:34-69)
int thinning_evaluate(const double* coefficients, size_t n, double point, double* out);

// low-level pieces exposed for tests
double standard_f64(Rng& rng);
double open01_f64(Rng& rng);
double sample_exp1(Rng& rng);
double sample_standard_normal(Rng& rng);
const double* zig_exp_x();
const double* zig_exp_f();
const double* zig_norm_x();
const double* zig_norm_f();
uint64_t uniform_u64(Rng& rng, uint64_t low, uint64_t high, int* err);

// ---- coupling (sim/src/simulator/coupling.rs:9-123) --------------------------------------------
struct Connector {
    std::string id, source_id, target_id, source_port, target_port;
};
struct Message {
    std::string source_id, source_port, target_id, target_port;
    double time = 0.0;
    std::string content;
};
struct ModelMessage {  // sim/src/models/mod.rs:42-45
    std::string port_name, content;
};
struct ModelRecord {  // sim/src/models/mod.rs:48-52
    double time;
    std::string action, subject;
};

// ---- trace (this project's definition; shared spec with the CUDA engine) --------
enum Action : int {
    A_INITIALIZATION = 0, A_GENERATION, A_ARRIVAL, A_PROCESSING_START, A_DROP, A_DEPARTURE,
    A_ACTIVATION, A_DEACTIVATION, A_PASS, A_BLOCK, A_START, A_STOP, A_MINIMUM_FETCH,
    A_MAXIMUM_FETCH, A_COUNT
};
const char* action_name(int a);
struct TraceEvent {
    uint32_t step;     // 1-based index of the Simulation::step that produced it
    uint8_t kind;      // 1 = model record, 0 = routed message
    uint16_t index;    // record: model index; message: connector index
    uint8_t action;    // record: Action; message: 0
    uint8_t aux;       // record: 1 + port index when the subject ends in " on <port>", else 0
    uint32_t content;  // interned content code
    double time;
};
struct Tracer {
    bool enabled = false, store = false;
    uint64_t hash = 0xcbf29ce484222325ull;
    uint64_t count = 0;
    uint64_t record_count[32 * 14] = {0};  // [model][action]: would-be ModelRecords (test instrumentation)
    std::vector<TraceEvent> events;
    std::vector<std::string> names;  // prefix / literal intern table
    uint32_t intern(const std::string& s);
    uint32_t content_code(const std::string& content);
    std::string render(uint32_t code) const;
    void push(const TraceEvent& ev);
};
extern const std::string NO_CONTENT;  // marker passed to record() when the subject names no job
static const uint32_t CONTENT_NONE = 0xFFFFFFFFu;
static const uint32_t CONTENT_NUM_LITERAL = 0xFFFFFFu;

// ---- Services (sim/src/simulator/services.rs:9-36) ---------------------------------------------
struct Services {
    Rng* global_rng = nullptr;
    double global_time = 0.0;
    // test instrumentation (not in the reference)
    Tracer* tracer = nullptr;
    int cur_model = -1;
    uint32_t cur_step = 0;
    // event counters of the running Simulation: a Coupled model counts the events_ext / events_int calls it makes on
    // its components instead of its own (one event = one transition of an ATOMIC model)
    uint64_t* n_ext = nullptr;
    uint64_t* n_int = nullptr;
};

// ---- DevsModel + Reportable (sim/src/models/model_trait.rs:37-59) ------------------------------
struct ModelBase {
    bool store_records = false;
    std::vector<ModelRecord> recs;
    // `rng: Option<DynRng>` of Generator / Processor / ExclusiveGateway / StochasticGate (e.g. generator.rs:39,
    // 102-109): a model constructed with its own generator draws from it instead of Services::global_rng.  Here the
    // override is a seed: the model's stream is the Philox stream (own_seed, replication), created at put().
    bool has_own_rng = false;
    uint64_t own_seed = 0;
    std::shared_ptr<Rng> own_rng;
    Rng& rng_of(Services& s) { return own_rng ? *own_rng : *s.global_rng; }
    virtual ~ModelBase() {}
    virtual int events_ext(const ModelMessage& m, Services& s) = 0;
    virtual int events_int(Services& s, std::vector<ModelMessage>& out) = 0;
    virtual void time_advance(double dt) = 0;
    virtual double until_next_event() const = 0;
    virtual std::string status() const = 0;
    virtual const char* type_name() const = 0;
    virtual std::unique_ptr<ModelBase> clone() const = 0;
    void record(Services& s, int action, const std::string& subject, const std::string& content,
                int aux = 0);
};
struct Model {  // sim/src/models/model.rs:13-26
    std::string id;
    std::unique_ptr<ModelBase> inner;
};
// Coupled (sim/src/models/coupled.rs): couplings as plain string tuples
struct ExternalInputCoupling {
    std::string target_id, source_port, target_port;
};
struct ExternalOutputCoupling {
    std::string source_id, source_port, target_port;
};
struct InternalCoupling {
    std::string source_id, target_id, source_port, target_port;
};
std::unique_ptr<ModelBase> make_coupled(const std::vector<std::string>& ports_in, const std::vector<std::string>& ports_out,
                                        std::vector<Model> components, const std::vector<ExternalInputCoupling>& eic,
                                        const std::vector<ExternalOutputCoupling>& eoc,
                                        const std::vector<InternalCoupling>& ic);
bool is_coupled(const ModelBase* m);
// trace indices: atomic models are numbered in document order with the components of a Coupled model in its place
// (the order a flattened network would list them); returns the next free index
int assign_trace_indices(std::vector<Model>& models, int first);

struct Simulation {  // sim/src/simulator/mod.rs:39-304
    std::vector<Model> models;
    std::vector<Connector> connectors;
    std::vector<Message> messages;
    Services services;
    std::unique_ptr<Rng> rng_owner;
    Tracer tracer;
    // counters (test instrumentation): one "event" = one events_ext or events_int call
    uint64_t n_steps = 0, n_ext = 0, n_int = 0;
    std::vector<int> top_trace_index;  // trace index of each top-level model (-1: a Coupled model)
    void bind_model_rngs();            // (re)create the streams of models that carry their own generator
    std::vector<uint64_t> connector_messages;  // per-connector routed message count

    Simulation();
    void put(std::vector<Model> m, std::vector<Connector> c);       // mod.rs:82-85
    void reset();                                                     // mod.rs:131-144
    void inject_input(const Message& m) { messages.push_back(m); }   // mod.rs:189-191
    int step();                                                       // mod.rs:198-272
    int step_until(double until, std::vector<Message>* out);         // mod.rs:277-288
    int step_n(size_t n, std::vector<Message>* out);                 // mod.rs:293-303
    int find_model(const std::string& id) const;
    int connector_index_of(int model_index, const std::string& port, size_t nth) const;
};

// ---- output_analysis (sim/src/output_analysis/mod.rs, t_scores.rs; utils/mod.rs:84-92) ---------
size_t usize_sqrt(size_t n);
int t_score(double alpha, size_t df, double* out);  // Err::Panic where the reference panics
struct ConfidenceInterval {
    double lower, upper;
    double half_width() const { return (upper - lower) / 2.0; }
};
struct IndependentSample {
    std::vector<double> points;
    double mean = 0, variance = 0;
    static IndependentSample post(const std::vector<double>& p);
    int confidence_interval_mean(double alpha, ConfidenceInterval* ci) const;
};
struct SteadyStateOutput {
    std::vector<double> time_series;
    bool has_deletion = false, has_batches = false, has_stats = false;
    size_t deletion_point = 0, batch_size = 0, batch_count = 0;
    std::vector<double> batch_means;
    double batches_mean = 0, batches_variance = 0;
    static SteadyStateOutput post(const std::vector<double>& ts);
    int set_to_fixed_budget();
    int calculate_batch_statistics();
    int confidence_interval_mean(double alpha, ConfidenceInterval* ci);
    int point_estimate_mean(double* out);
};

// ---- model constructors (mirror the reference's `new` signatures) ------------------------------
std::unique_ptr<ModelBase> make_generator(const Continuous& d, const std::string& job_port, bool store);
std::unique_ptr<ModelBase> make_processor(const Continuous& d, int64_t capacity /* <0: usize::MAX */,
                                          const std::string& job_port, const std::string& processed_port, bool store);
std::unique_ptr<ModelBase> make_storage(const std::string& put, const std::string& get,
                                        const std::string& stored, bool store);
std::unique_ptr<ModelBase> make_load_balancer(const std::string& job_port, const std::vector<std::string>& paths, bool store);
std::unique_ptr<ModelBase> make_gate(const std::string& job_in, const std::string& activation,
                                     const std::string& deactivation, const std::string& job_out, bool store);
std::unique_ptr<ModelBase> make_exclusive_gateway(const std::vector<std::string>& in, const std::vector<std::string>& out,
                                                  const Index& weights, bool store);
std::unique_ptr<ModelBase> make_parallel_gateway(const std::vector<std::string>& in, const std::vector<std::string>& out, bool store);
std::unique_ptr<ModelBase> make_stochastic_gate(const Boolean& pass, const std::string& in, const std::string& out, bool store);
std::unique_ptr<ModelBase> make_stopwatch(const std::string& start, const std::string& stop, const std::string& metric_port,
                                          const std::string& job_port, int metric /*0 min, 1 max*/, bool store);
std::unique_ptr<ModelBase> make_batcher(const std::string& in, const std::string& out, double max_time, uint64_t max_size, bool store);
std::unique_ptr<ModelBase> make_passive(const std::string& in);  // sim/tests/custom.rs:18-86
// state injection used by sim/tests/web.rs:30-46 (Processor `state` block)
int preload_processor(ModelBase* m, int phase_active, double until_next_event, const std::vector<std::string>& queue);
int preload_state(ModelBase* m, int phase, double until_next_event, const std::vector<std::string>& jobs);
int preload_parallel_gateway(ModelBase* m, double until_next_event, const std::vector<std::string>& names, const uint64_t* counts);
int preload_stochastic_gate(ModelBase* m, double until_next_event, const std::vector<std::string>& contents, const uint8_t* pass);
int preload_stopwatch(ModelBase* m, int job_fetch, double until_next_event, const std::vector<std::string>& names,
                      const double* start, const double* stop);

}  // namespace orc
