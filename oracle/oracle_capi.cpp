// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface over the CPU oracle so tests (ctypes) and bench.py's cpu_baseline /
// --impl reference legs can drive it.  Nothing in sim_b200/ links or loads this.
#include "oracle.hpp"

#include <atomic>
#include <cstring>
#include <map>
#include <thread>

using namespace orc;

namespace {
struct CoupledDraft {  // a Coupled model under construction (orc_begin_coupled .. orc_end_coupled)
    std::string id;
    std::vector<std::string> ports_in, ports_out;
    std::vector<Model> components;
    std::vector<ExternalInputCoupling> eic;
    std::vector<ExternalOutputCoupling> eoc;
    std::vector<InternalCoupling> ic;
};
struct Handle {
    Simulation sim;
    std::vector<Model> staged_models;
    std::vector<Connector> staged_connectors;
    std::vector<Message> returned;
    std::vector<CoupledDraft> building;  // innermost last; models are added to its components
};
// where orc_add_* / orc_preload_* currently act: the components of the innermost Coupled draft, else the top level
std::vector<Model>& target(Handle* h) { return h->building.empty() ? h->staged_models : h->building.back().components; }
std::vector<std::string> split_nl(const char* s) {
    std::vector<std::string> out;
    if (!s || !*s) return out;
    std::string cur;
    for (const char* p = s; *p; p++) {
        if (*p == '\n') {
            out.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(*p);
        }
    }
    out.push_back(cur);
    return out;
}
void put_str(char* buf, int len, const std::string& s) {
    if (!buf || len <= 0) return;
    size_t n = std::min((size_t)len - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
}
Continuous mk(int kind, double a, double b, double c) {
    Continuous d;
    d.kind = kind;
    d.a = a;
    d.b = b;
    d.c = c;
    return d;
}
void stage(Handle* h, const char* id, std::unique_ptr<ModelBase> m) {
    Model mod;
    mod.id = id;
    mod.inner = std::move(m);
    target(h).push_back(std::move(mod));
}
std::vector<Model> clone_models(const std::vector<Model>& src) {
    std::vector<Model> out;
    for (auto& m : src) {
        Model c;
        c.id = m.id;
        c.inner = m.inner->clone();
        out.push_back(std::move(c));
    }
    return out;
}
// per-job waiting time in processing-start order (the statistic of sim/tests/web.rs:571-597,
// with a deterministic summation order instead of HashMap order)
void wait_from_records(const std::vector<ModelRecord>& recs, double* sum, uint64_t* n) {
    std::map<std::string, double> arrival;
    double s = 0.0;
    uint64_t c = 0;
    for (auto& r : recs) {
        if (r.action == "Arrival") arrival[r.subject] = r.time;
        else if (r.action == "Processing Start") {
            auto it = arrival.find(r.subject);
            if (it != arrival.end()) {
                s += r.time - it->second;
                c++;
                arrival.erase(it);
            }
        }
    }
    *sum = s;
    *n = c;
}
}  // namespace

extern "C" {

void* orc_new() { return new Handle(); }
void orc_free(void* h) { delete (Handle*)h; }

int orc_add_generator(void* h, const char* id, int kind, double a, double b, double c, const char* port_job, int store) {
    stage((Handle*)h, id, make_generator(mk(kind, a, b, c), port_job, store != 0));
    return 0;
}
int orc_add_processor(void* h, const char* id, int kind, double a, double b, double c, int64_t capacity,
                      const char* port_in, const char* port_out, int store) {
    stage((Handle*)h, id, make_processor(mk(kind, a, b, c), capacity, port_in, port_out, store != 0));
    return 0;
}
int orc_add_storage(void* h, const char* id, const char* put, const char* get, const char* stored, int store) {
    stage((Handle*)h, id, make_storage(put, get, stored, store != 0));
    return 0;
}
int orc_add_load_balancer(void* h, const char* id, const char* port_in, const char* paths_nl, int store) {
    stage((Handle*)h, id, make_load_balancer(port_in, split_nl(paths_nl), store != 0));
    return 0;
}
int orc_add_gate(void* h, const char* id, const char* job, const char* act, const char* deact, const char* out, int store) {
    stage((Handle*)h, id, make_gate(job, act, deact, out, store != 0));
    return 0;
}
int orc_add_exclusive_gateway(void* h, const char* id, const char* ins_nl, const char* outs_nl, int weighted,
                              uint64_t mn, uint64_t mx, const uint64_t* w, int nw, int store) {
    Index ix;
    ix.weighted = weighted != 0;
    ix.min = mn;
    ix.max = mx;
    for (int i = 0; i < nw; i++) ix.weights.push_back(w[i]);
    stage((Handle*)h, id, make_exclusive_gateway(split_nl(ins_nl), split_nl(outs_nl), ix, store != 0));
    return 0;
}
int orc_add_parallel_gateway(void* h, const char* id, const char* ins_nl, const char* outs_nl, int store) {
    stage((Handle*)h, id, make_parallel_gateway(split_nl(ins_nl), split_nl(outs_nl), store != 0));
    return 0;
}
int orc_add_stochastic_gate(void* h, const char* id, double p, const char* in, const char* out, int store) {
    Boolean b;
    b.p = p;
    stage((Handle*)h, id, make_stochastic_gate(b, in, out, store != 0));
    return 0;
}
int orc_add_stopwatch(void* h, const char* id, const char* start, const char* stop, const char* metric_port,
                      const char* job_port, int metric, int store) {
    stage((Handle*)h, id, make_stopwatch(start, stop, metric_port, job_port, metric, store != 0));
    return 0;
}
int orc_add_batcher(void* h, const char* id, const char* in, const char* out, double max_time, uint64_t max_size, int store) {
    stage((Handle*)h, id, make_batcher(in, out, max_time, max_size, store != 0));
    return 0;
}
int orc_add_passive(void* h, const char* id, const char* in) {
    stage((Handle*)h, id, make_passive(in));
    return 0;
}
int orc_add_connector(void* h, const char* id, const char* src, const char* tgt, const char* sport, const char* tport) {
    ((Handle*)h)->staged_connectors.push_back(Connector{id, src, tgt, sport, tport});
    return 0;
}
// Generator::new(.., rng: Some(..)) etc.: the last model added draws from its own stream
int orc_set_model_rng(void* hv, const char* id, uint64_t seed) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) {
            m.inner->has_own_rng = true;
            m.inner->own_seed = seed;
            return 0;
        }
    return ModelNotFound;
}
int orc_preload_processor(void* hv, const char* id, int active, double until, const char* queue_nl) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) return preload_processor(m.inner.get(), active, until, split_nl(queue_nl));
    return ModelNotFound;
}
int orc_preload_state(void* hv, const char* id, int phase, double until, const char* jobs_nl) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) return preload_state(m.inner.get(), phase, until, split_nl(jobs_nl));
    return ModelNotFound;
}
int orc_preload_parallel_gateway(void* hv, const char* id, double until, const char* names_nl, const uint64_t* counts) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) return preload_parallel_gateway(m.inner.get(), until, split_nl(names_nl), counts);
    return ModelNotFound;
}
int orc_preload_stochastic_gate(void* hv, const char* id, double until, const char* contents_nl, const uint8_t* pass) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) return preload_stochastic_gate(m.inner.get(), until, split_nl(contents_nl), pass);
    return ModelNotFound;
}
int orc_preload_stopwatch(void* hv, const char* id, int job_fetch, double until, const char* names_nl, const double* start,
                          const double* stop) {
    Handle* h = (Handle*)hv;
    for (auto& m : target(h))
        if (m.id == id) return preload_stopwatch(m.inner.get(), job_fetch, until, split_nl(names_nl), start, stop);
    return ModelNotFound;
}
// Coupled::new (coupled.rs:87-106): the models added between begin and end become its components
int orc_begin_coupled(void* hv, const char* id, const char* ports_in_nl, const char* ports_out_nl) {
    Handle* h = (Handle*)hv;
    CoupledDraft d;
    d.id = id;
    d.ports_in = split_nl(ports_in_nl);
    d.ports_out = split_nl(ports_out_nl);
    h->building.push_back(std::move(d));
    return 0;
}
int orc_coupled_eic(void* hv, const char* target_id, const char* source_port, const char* target_port) {
    Handle* h = (Handle*)hv;
    if (h->building.empty()) return InvalidModelConfiguration;
    h->building.back().eic.push_back(ExternalInputCoupling{target_id, source_port, target_port});
    return 0;
}
int orc_coupled_eoc(void* hv, const char* source_id, const char* source_port, const char* target_port) {
    Handle* h = (Handle*)hv;
    if (h->building.empty()) return InvalidModelConfiguration;
    h->building.back().eoc.push_back(ExternalOutputCoupling{source_id, source_port, target_port});
    return 0;
}
int orc_coupled_ic(void* hv, const char* source_id, const char* target_id, const char* source_port, const char* target_port) {
    Handle* h = (Handle*)hv;
    if (h->building.empty()) return InvalidModelConfiguration;
    h->building.back().ic.push_back(InternalCoupling{source_id, target_id, source_port, target_port});
    return 0;
}
int orc_end_coupled(void* hv) {
    Handle* h = (Handle*)hv;
    if (h->building.empty()) return InvalidModelConfiguration;
    CoupledDraft d = std::move(h->building.back());
    h->building.pop_back();
    stage(h, d.id.c_str(), make_coupled(d.ports_in, d.ports_out, std::move(d.components), d.eic, d.eoc, d.ic));
    return 0;
}
// Simulation::put with fresh clones of the staged models (the reference's replication idiom:
// reset() + put(models.to_vec(), ..), sim/tests/simulations.rs:162-170)
int orc_post(void* hv) {
    Handle* h = (Handle*)hv;
    h->sim.put(clone_models(h->staged_models), h->staged_connectors);
    return 0;
}
int orc_reset(void* hv) {
    ((Handle*)hv)->sim.reset();
    return 0;
}
int orc_set_rng_pcg(void* hv, uint64_t lo, uint64_t hi) {
    Handle* h = (Handle*)hv;
    h->sim.rng_owner = std::make_unique<Pcg64Mcg>(((unsigned __int128)hi << 64) | lo);
    h->sim.services.global_rng = h->sim.rng_owner.get();
    return 0;
}
int orc_set_rng_philox(void* hv, uint64_t seed, uint64_t rep) {
    Handle* h = (Handle*)hv;
    h->sim.rng_owner = std::make_unique<Philox>(seed, rep);
    h->sim.services.global_rng = h->sim.rng_owner.get();
    h->sim.bind_model_rngs();
    return 0;
}
int orc_set_rng_external(void* hv, const uint64_t* v, size_t n) {
    Handle* h = (Handle*)hv;
    auto e = std::make_unique<ExternalStream>();
    e->values.assign(v, v + n);
    h->sim.rng_owner = std::move(e);
    h->sim.services.global_rng = h->sim.rng_owner.get();
    return 0;
}
uint64_t orc_rng_draws(void* hv) { return ((Handle*)hv)->sim.services.global_rng->draws; }
int orc_rng_exhausted(void* hv) {
    auto* e = dynamic_cast<ExternalStream*>(((Handle*)hv)->sim.services.global_rng);
    return e ? (e->exhausted ? 1 : 0) : 0;
}
int orc_inject(void* hv, const char* tgt, const char* tport, const char* content) {
    Handle* h = (Handle*)hv;
    // Message::new(source_id, source_port, target_id, target_port, time, content); the tests inject
    // with source "manual"/"manual" (sim/tests/simulations.rs:640-647)
    h->sim.inject_input(Message{"manual", "manual", tgt, tport, h->sim.services.global_time, content});
    return 0;
}
int orc_step(void* hv) {
    Handle* h = (Handle*)hv;
    h->returned.clear();
    int e = h->sim.step();
    if (!e) h->returned = h->sim.messages;
    return e;
}
int orc_step_n(void* hv, uint64_t n, int collect) {
    Handle* h = (Handle*)hv;
    h->returned.clear();
    return h->sim.step_n((size_t)n, collect ? &h->returned : nullptr);
}
int orc_step_until(void* hv, double until, int collect) {
    Handle* h = (Handle*)hv;
    h->returned.clear();
    return h->sim.step_until(until, collect ? &h->returned : nullptr);
}
static int get_msg(const std::vector<Message>& v, size_t i, char* src, char* sport, char* tgt, char* tport,
                   char* content, int len, double* time) {
    if (i >= v.size()) return -1;
    put_str(src, len, v[i].source_id);
    put_str(sport, len, v[i].source_port);
    put_str(tgt, len, v[i].target_id);
    put_str(tport, len, v[i].target_port);
    put_str(content, len, v[i].content);
    if (time) *time = v[i].time;
    return 0;
}
uint64_t orc_returned_count(void* hv) { return ((Handle*)hv)->returned.size(); }
int orc_returned_get(void* hv, uint64_t i, char* src, char* sport, char* tgt, char* tport, char* content, int len, double* time) {
    return get_msg(((Handle*)hv)->returned, (size_t)i, src, sport, tgt, tport, content, len, time);
}
uint64_t orc_messages_count(void* hv) { return ((Handle*)hv)->sim.messages.size(); }
int orc_messages_get(void* hv, uint64_t i, char* src, char* sport, char* tgt, char* tport, char* content, int len, double* time) {
    return get_msg(((Handle*)hv)->sim.messages, (size_t)i, src, sport, tgt, tport, content, len, time);
}
double orc_global_time(void* hv) { return ((Handle*)hv)->sim.services.global_time; }
int orc_status(void* hv, const char* id, char* buf, int len) {
    Handle* h = (Handle*)hv;
    int i = h->sim.find_model(id);
    if (i < 0) return ModelNotFound;  // mod.rs:106-113
    put_str(buf, len, h->sim.models[i].inner->status());
    return 0;
}
double orc_until_next_event(void* hv, const char* id) {
    Handle* h = (Handle*)hv;
    int i = h->sim.find_model(id);
    return i < 0 ? std::nan("") : h->sim.models[i].inner->until_next_event();
}
int64_t orc_records_count(void* hv, const char* id) {
    Handle* h = (Handle*)hv;
    int i = h->sim.find_model(id);
    if (i < 0) return -ModelNotFound;  // mod.rs:118-125
    return (int64_t)h->sim.models[i].inner->recs.size();
}
int orc_record_get(void* hv, const char* id, uint64_t k, double* time, char* action, char* subject, int len) {
    Handle* h = (Handle*)hv;
    int i = h->sim.find_model(id);
    if (i < 0) return ModelNotFound;
    auto& r = h->sim.models[i].inner->recs;
    if (k >= r.size()) return -1;
    *time = r[k].time;
    put_str(action, len, r[k].action);
    put_str(subject, len, r[k].subject);
    return 0;
}
void orc_trace_enable(void* hv, int store) {
    Handle* h = (Handle*)hv;
    h->sim.tracer.enabled = true;
    h->sim.tracer.store = store != 0;
}
uint64_t orc_trace_hash(void* hv) { return ((Handle*)hv)->sim.tracer.hash; }
uint64_t orc_trace_count(void* hv) { return ((Handle*)hv)->sim.tracer.count; }
// fields: step, kind, index, action, aux, content
int orc_trace_get(void* hv, uint64_t i, uint32_t out[6], double* time) {
    auto& ev = ((Handle*)hv)->sim.tracer.events;
    if (i >= ev.size()) return -1;
    out[0] = ev[i].step;
    out[1] = ev[i].kind;
    out[2] = ev[i].index;
    out[3] = ev[i].action;
    out[4] = ev[i].aux;
    out[5] = ev[i].content;
    *time = ev[i].time;
    return 0;
}
int orc_trace_render(void* hv, uint32_t code, char* buf, int len) {
    put_str(buf, len, ((Handle*)hv)->sim.tracer.render(code));
    return 0;
}
uint32_t orc_trace_intern(void* hv, const char* s) { return ((Handle*)hv)->sim.tracer.content_code(s); }
const char* orc_action_name(int a) { return action_name(a); }
void orc_counters(void* hv, uint64_t out[3]) {
    Handle* h = (Handle*)hv;
    out[0] = h->sim.n_steps;
    out[1] = h->sim.n_ext;
    out[2] = h->sim.n_int;
}
uint64_t orc_connector_messages(void* hv, uint64_t i) {
    Handle* h = (Handle*)hv;
    return i < h->sim.connector_messages.size() ? h->sim.connector_messages[i] : 0;
}
int orc_mean_wait(void* hv, const char* id, double* sum, uint64_t* n) {
    Handle* h = (Handle*)hv;
    int i = h->sim.find_model(id);
    if (i < 0) return ModelNotFound;
    wait_from_records(h->sim.models[i].inner->recs, sum, n);
    return 0;
}

// ---- output_analysis probes --------------------------------------------------------------------
int orc_independent_sample(const double* pts, uint64_t n, double alpha, double out[4]) {
    IndependentSample s = IndependentSample::post(std::vector<double>(pts, pts + n));
    ConfidenceInterval ci{0, 0};
    int e = s.confidence_interval_mean(alpha, &ci);
    out[0] = s.mean;
    out[1] = s.variance;
    out[2] = ci.lower;
    out[3] = ci.upper;
    return e;
}
int orc_steady_state(const double* pts, uint64_t n, double alpha, double out[4], uint64_t outn[3]) {
    SteadyStateOutput s = SteadyStateOutput::post(std::vector<double>(pts, pts + n));
    ConfidenceInterval ci{0, 0};
    int e = s.confidence_interval_mean(alpha, &ci);
    out[0] = s.batches_mean;
    out[1] = s.batches_variance;
    out[2] = ci.lower;
    out[3] = ci.upper;
    outn[0] = s.deletion_point;
    outn[1] = s.batch_size;
    outn[2] = s.batch_count;
    return e;
}
uint64_t orc_usize_sqrt(uint64_t n) { return usize_sqrt((size_t)n); }
int orc_t_score(double alpha, uint64_t df, double* out) { return t_score(alpha, (size_t)df, out); }

// ---- rng / sampler probes ----------------------------------------------------------------------
void orc_pcg_fill(uint64_t lo, uint64_t hi, uint64_t* out, uint64_t n) {
    Pcg64Mcg r(((unsigned __int128)hi << 64) | lo);
    for (uint64_t i = 0; i < n; i++) out[i] = r.next_u64();
}
void orc_philox_block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox::block_raw(ctr, key, out); }
void orc_philox_fill(uint64_t seed, uint64_t rep, uint64_t* out, uint64_t n) {
    Philox r(seed, rep);
    for (uint64_t i = 0; i < n; i++) out[i] = r.next_u64();
}
static std::unique_ptr<Rng> mk_rng(int rng_kind, uint64_t seed, uint64_t rep) {
    if (rng_kind == 0) return std::make_unique<Philox>(seed, rep);
    return std::make_unique<Pcg64Mcg>((unsigned __int128)seed);
}
// draws n variates; returns error code; consumed = u64 draws used
int orc_sample_continuous(int kind, double a, double b, double c, int rng_kind, uint64_t seed, uint64_t rep,
                          double* out, uint64_t n, uint64_t* consumed) {
    auto r = mk_rng(rng_kind, seed, rep);
    Continuous d = mk(kind, a, b, c);
    for (uint64_t i = 0; i < n; i++) {
        int e = d.random_variate(*r, &out[i]);
        if (e) return e;
    }
    if (consumed) *consumed = r->draws;
    return 0;
}
int orc_sample_bernoulli(double p, int rng_kind, uint64_t seed, uint64_t rep, uint8_t* out, uint64_t n, uint64_t* consumed) {
    auto r = mk_rng(rng_kind, seed, rep);
    Boolean d;
    d.p = p;
    for (uint64_t i = 0; i < n; i++) {
        bool b;
        int e = d.random_variate(*r, &b);
        if (e) return e;
        out[i] = b;
    }
    if (consumed) *consumed = r->draws;
    return 0;
}
int orc_sample_discrete(int kind, double p, uint64_t mn, uint64_t mx, int rng_kind, uint64_t seed, uint64_t rep,
                        uint64_t* out, uint64_t n, uint64_t* consumed) {
    auto r = mk_rng(rng_kind, seed, rep);
    Discrete d;
    d.kind = kind;
    d.p = p;
    d.min = mn;
    d.max = mx;
    for (uint64_t i = 0; i < n; i++) {
        int e = d.random_variate(*r, &out[i]);
        if (e) return e;
    }
    if (consumed) *consumed = r->draws;
    return 0;
}
int orc_thinning_evaluate(const double* coefficients, uint64_t n, double point, double* out) {
    return thinning_evaluate(coefficients, (size_t)n, point, out);
}
int orc_sample_index(int weighted, uint64_t mn, uint64_t mx, const uint64_t* w, int nw, int rng_kind, uint64_t seed,
                     uint64_t rep, uint64_t* out, uint64_t n, uint64_t* consumed) {
    auto r = mk_rng(rng_kind, seed, rep);
    Index ix;
    ix.weighted = weighted != 0;
    ix.min = mn;
    ix.max = mx;
    for (int i = 0; i < nw; i++) ix.weights.push_back(w[i]);
    for (uint64_t i = 0; i < n; i++) {
        int e = ix.random_variate(*r, &out[i]);
        if (e) return e;
    }
    if (consumed) *consumed = r->draws;
    return 0;
}
void orc_zig_table(int which, double out[257]) {
    const double* t = which == 0 ? zig_exp_x() : which == 1 ? zig_exp_f() : which == 2 ? zig_norm_x() : zig_norm_f();
    std::memcpy(out, t, 257 * sizeof(double));
}

// ---- batch runner: the reference's replication loop, one replication per host thread -----------
// Each replication r in [rep0, rep0+nreps) gets fresh clones of the staged models, a Philox
// stream keyed (seed, r), and runs step_until(until) (mode 0) or step_n(nsteps) (mode 1).
// Per-replication outputs (any pointer may be null): final global time, trace hash, steps,
// events (ext+int), waiting-time sum / count of `stat_model_id` (needs store_records on it).
// totals[0..2] = steps, ext events, int events over all replications.
// out_conn[n_connectors] / out_records[n_models][14]: totals over the batch of routed messages per connector and of
// would-be ModelRecords per (model, action) (trace_hash must be on for the latter).
int orc_run_batch_counts(void* hv, uint64_t seed, uint64_t rep0, uint64_t nreps, int threads, int mode, double until,
                         uint64_t nsteps, const char* stat_model_id, int trace_hash, uint64_t* totals, double* out_time,
                         uint64_t* out_hash, uint64_t* out_steps, uint64_t* out_events, double* out_wait_sum,
                         uint64_t* out_wait_n, uint64_t* out_conn, uint64_t* out_records);
int orc_run_batch(void* hv, uint64_t seed, uint64_t rep0, uint64_t nreps, int threads, int mode, double until,
                  uint64_t nsteps, const char* stat_model_id, int trace_hash, uint64_t* totals, double* out_time,
                  uint64_t* out_hash, uint64_t* out_steps, uint64_t* out_events, double* out_wait_sum,
                  uint64_t* out_wait_n) {
    return orc_run_batch_counts(hv, seed, rep0, nreps, threads, mode, until, nsteps, stat_model_id, trace_hash, totals,
                                out_time, out_hash, out_steps, out_events, out_wait_sum, out_wait_n, nullptr, nullptr);
}
int orc_run_batch_counts(void* hv, uint64_t seed, uint64_t rep0, uint64_t nreps, int threads, int mode, double until,
                         uint64_t nsteps, const char* stat_model_id, int trace_hash, uint64_t* totals, double* out_time,
                         uint64_t* out_hash, uint64_t* out_steps, uint64_t* out_events, double* out_wait_sum,
                         uint64_t* out_wait_n, uint64_t* out_conn, uint64_t* out_records) {
    Handle* h = (Handle*)hv;
    const size_t n_conn = h->staged_connectors.size(), n_mod = h->staged_models.size();
    std::vector<uint64_t> conn_t((size_t)(threads < 1 ? 1 : threads) * (n_conn + 1), 0);
    std::vector<uint64_t> rec_t((size_t)(threads < 1 ? 1 : threads) * 32 * 14, 0);
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> next{0};
    std::atomic<int> first_err{0};
    std::vector<uint64_t> tot(3 * (size_t)threads, 0);
    std::string stat_id = stat_model_id ? stat_model_id : "";
    auto worker = [&](int t) {
        const uint64_t CHUNK = 16;
        for (;;) {
            uint64_t b = next.fetch_add(CHUNK);
            if (b >= nreps) break;
            uint64_t e = std::min(nreps, b + CHUNK);
            for (uint64_t i = b; i < e; i++) {
                Simulation sim;
                sim.rng_owner = std::make_unique<Philox>(seed, rep0 + i);
                sim.services.global_rng = sim.rng_owner.get();
                if (trace_hash) sim.tracer.enabled = true;
                sim.put(clone_models(h->staged_models), h->staged_connectors);
                int err = mode == 0 ? sim.step_until(until, nullptr) : sim.step_n((size_t)nsteps, nullptr);
                if (err) {
                    int z = 0;
                    first_err.compare_exchange_strong(z, err);
                }
                tot[3 * t + 0] += sim.n_steps;
                tot[3 * t + 1] += sim.n_ext;
                tot[3 * t + 2] += sim.n_int;
                for (size_t c = 0; c < n_conn && c < sim.connector_messages.size(); c++)
                    conn_t[(size_t)t * (n_conn + 1) + c] += sim.connector_messages[c];
                for (size_t k = 0; k < 32 * 14; k++) rec_t[(size_t)t * 32 * 14 + k] += sim.tracer.record_count[k];
                if (out_time) out_time[i] = sim.services.global_time;
                if (out_hash) out_hash[i] = sim.tracer.hash;
                if (out_steps) out_steps[i] = sim.n_steps;
                if (out_events) out_events[i] = sim.n_ext + sim.n_int;
                if (!stat_id.empty() && (out_wait_sum || out_wait_n)) {
                    int mi = sim.find_model(stat_id);
                    double s = 0;
                    uint64_t n = 0;
                    if (mi >= 0) wait_from_records(sim.models[mi].inner->recs, &s, &n);
                    if (out_wait_sum) out_wait_sum[i] = s;
                    if (out_wait_n) out_wait_n[i] = n;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    if (totals) {
        totals[0] = totals[1] = totals[2] = 0;
        for (int t = 0; t < threads; t++)
            for (int k = 0; k < 3; k++) totals[k] += tot[3 * t + k];
    }
    if (out_conn)
        for (size_t c = 0; c < n_conn; c++) {
            out_conn[c] = 0;
            for (int t = 0; t < threads; t++) out_conn[c] += conn_t[(size_t)t * (n_conn + 1) + c];
        }
    if (out_records)
        for (size_t m = 0; m < n_mod && m < 32; m++)
            for (size_t a = 0; a < 14; a++) {
                out_records[m * 14 + a] = 0;
                for (int t = 0; t < threads; t++) out_records[m * 14 + a] += rec_t[(size_t)t * 32 * 14 + m * 14 + a];
            }
    return first_err.load();
}

}  // extern "C"
