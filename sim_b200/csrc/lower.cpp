// sim_b200/csrc/lower.cpp -- JSON (SURVEY.md Appendix E wire schema) -> SimbNetDesc.
// Replaces, for the batched path, Model::deserialize + model_factory::create
// (sim/src/models/model.rs:43-50, model_factory.rs:65-77) and the per-step string matching of
// Simulation::get_message_target_ids/ports (sim/src/simulator/mod.rs:155-182): ids, ports and
// connectors are resolved ONCE into integer routes kept in connector order.
#include "lower.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "json.hpp"

namespace simb {

using json::Value;

static const char* ACTION_NAMES[ACT_COUNT] = {
    "Initialization", "Generation", "Arrival", "Processing Start", "Drop", "Departure",
    "Activation", "Deactivation", "Pass", "Block", "Start", "Stop", "Minimum Fetch", "Maximum Fetch"};
const char* action_name(int a) { return (a >= 0 && a < ACT_COUNT) ? ACTION_NAMES[a] : "?"; }

const char* status_name(int s) {
    switch (s) {
        case 0: return "Ok";
        case 1: return "InvalidModelConfiguration";
        case 2: return "ModelNotFound";
        case 3: return "PortNotFound";
        case 4: return "ModelCloneError";
        case 5: return "InvalidModelState";
        case 6: return "EventSchedulingError";
        case 7: return "InvalidMessage";
        case 8: return "SerializationError";
        case 9: return "EmptyPolynomial";
        case 10: return "PrerequisiteCalcError";
        case 11: return "FloatConvError";
        case 12: return "DroppedMessageError";
        case 13: return "JSONError";
        case 14: return "BetaError";
        case 15: return "ExpError";
        case 16: return "GammaError";
        case 17: return "NormalError";
        case 18: return "TriangularError";
        case 19: return "WeibullError";
        case 20: return "BernoulliError";
        case 21: return "GeoError";
        case 22: return "PoissonError";
        case 23: return "WeightedError";
        case 32: return "CapacityError";
        case 33: return "UnsupportedModel";
        case 34: return "Panic";
        case 35: return "CudaError";
        case 36: return "InvalidArgument";
        case 37: return "StreamExhausted";
    }
    return "Unknown";
}

int HostNet::find_model(const std::string& id) const {
    for (size_t i = 0; i < models.size(); i++)
        if (models[i].coupled < 0 && models[i].id == id) return (int)i;
    return -1;
}
int HostNet::find_top(const std::string& id) const {
    for (size_t i = 0; i < tops.size(); i++)
        if (tops[i].id == id) return (int)i;
    return -1;
}
int HostNet::find_component(int ci, const std::string& id) const {
    if (ci < 0 || ci >= (int)coupled.size()) return -1;
    for (int k = coupled[ci].first; k < coupled[ci].first + coupled[ci].count; k++)
        if (models[k].id == id) return k;
    return -1;
}
std::vector<HostNet::Delivery> HostNet::deliveries(const std::string& id, const std::string& port, bool* bad_coupling) const {
    std::vector<Delivery> out;
    for (size_t t = 0; t < tops.size(); t++) {
        if (tops[t].id != id) continue;
        if (tops[t].coupled < 0) {
            out.push_back(Delivery{tops[t].flat, resolve_in_port(tops[t].flat, port)});
            continue;
        }
        const int cpl = tops[t].coupled;
        for (auto& e : coupled[cpl].eic) {
            if (e.source_port != port) continue;
            const int comp = find_component(cpl, e.target_id);
            if (comp < 0) {  // the reference unwraps a None here (coupled.rs:173-176)
                if (bad_coupling) *bad_coupling = true;
                continue;
            }
            out.push_back(Delivery{comp, resolve_in_port(comp, e.target_port)});
        }
    }
    return out;
}
uint32_t HostNet::intern(const std::string& s, bool* overflow) {
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == s) return (uint32_t)i;
    if (names.size() >= SIMB_MAX_NAMES) {
        if (overflow) *overflow = true;
        return 0;
    }
    names.push_back(s);
    return (uint32_t)(names.size() - 1);
}
// "<prefix> <n>" with n decimal, no leading zero, n < 0xFFFFFF  ->  intern(prefix) << 24 | n
// anything else                                               ->  intern(whole) << 24 | 0xFFFFFF
uint32_t HostNet::content_code(const std::string& content, bool* overflow) {
    size_t sp = content.rfind(' ');
    if (sp != std::string::npos && sp > 0 && sp + 1 < content.size()) {
        size_t nd = content.size() - sp - 1;
        const char* dg = content.c_str() + sp + 1;
        bool ok = nd <= 8 && !(nd > 1 && dg[0] == '0');
        uint64_t n = 0;
        for (size_t i = 0; ok && i < nd; i++) {
            if (dg[i] < '0' || dg[i] > '9') ok = false;
            else n = n * 10 + (uint64_t)(dg[i] - '0');
        }
        if (ok && n < SIMB_NUM_LITERAL) return (intern(content.substr(0, sp), overflow) << 24) | (uint32_t)n;
    }
    return (intern(content, overflow) << 24) | SIMB_NUM_LITERAL;
}
std::string HostNet::render(uint32_t code) const {
    if (code == SIMB_CONTENT_NONE) return "";
    uint32_t p = code >> 24, n = code & 0xFFFFFFu;
    std::string s = p < names.size() ? names[p] : "?";
    if (n == SIMB_NUM_LITERAL) return s;
    return s + " " + std::to_string(n);
}
uint32_t HostNet::resolve_in_port(int m, const std::string& port) const {
    const HostModel& hm = models[m];
    for (size_t i = 0; i < hm.ports_in.size(); i++)
        if (hm.ports_in[i] == port) return (uint32_t)i;  // first match wins, as the if-chains do
    return SIMB_PORT_UNKNOWN;
}

namespace {

struct Ctx {
    std::string* err;
    int fail(int code, const std::string& msg) {
        if (err && err->empty()) *err = msg;
        return code;
    }
};

bool get_str(const Value& obj, const char* key, std::string* out) {
    const Value* v = obj.get(key);
    if (!v || !v->is_string()) return false;
    *out = v->str;
    return true;
}
bool get_num(const Value& obj, const char* key, double* out) {
    const Value* v = obj.get(key);
    if (!v) return false;
    if (v->is_number()) {
        *out = v->num;
        return true;
    }
    // serde_json writes non-finite floats as null; YAML-style spellings are accepted too
    if (v->kind == Value::Null) {
        *out = std::numeric_limits<double>::infinity();
        return true;
    }
    if (v->is_string()) {
        if (v->str == ".inf" || v->str == "inf" || v->str == "Infinity") {
            *out = std::numeric_limits<double>::infinity();
            return true;
        }
    }
    return false;
}
bool get_str_list(const Value& obj, const char* key, std::vector<std::string>* out) {
    const Value* v = obj.get(key);
    if (!v || !v->is_array()) return false;
    for (auto& e : v->arr) {
        if (!e.is_string()) return false;
        out->push_back(e.str);
    }
    return true;
}

// Continuous (random_variable.rs:17-28): {"exp":{"lambda":..}} etc.  Inner field names keep their
// Rust spelling (std_dev).
int lower_continuous(Ctx& c, const Value* v, SimbModelDesc* md) {
    if (!v || !v->is_object() || v->obj.size() != 1) return c.fail(13, "expected a one-variant random variable object");
    const std::string& kind = v->obj[0].first;
    const Value& p = v->obj[0].second;
    md->dist_err = 0;
    if (kind == "exp") {
        double lambda;
        if (!get_num(p, "lambda", &lambda)) return c.fail(13, "exp: missing field `lambda`");
        md->dist_kind = DK_EXP;
        if (!(lambda > 0.0)) md->dist_err = 15;  // Exp::new fails per draw
        md->p0 = 1.0 / lambda;                    // lambda_inverse
        return 0;
    }
    if (kind == "normal") {
        double mean, sd;
        if (!get_num(p, "mean", &mean) || !get_num(p, "std_dev", &sd)) return c.fail(13, "normal: missing field");
        md->dist_kind = DK_NORMAL;
        if (!std::isfinite(sd)) md->dist_err = 17;
        md->p0 = mean;
        md->p1 = sd;
        return 0;
    }
    if (kind == "triangular") {
        double mn, mx, mode;
        if (!get_num(p, "min", &mn) || !get_num(p, "max", &mx) || !get_num(p, "mode", &mode))
            return c.fail(13, "triangular: missing field");
        md->dist_kind = DK_TRIANGULAR;
        if (!(mx >= mode) || !(mode >= mn) || !(mx != mn)) md->dist_err = 18;
        md->p0 = mn;
        md->p1 = mx;
        md->p2 = mode;
        return 0;
    }
    if (kind == "uniform") {
        double low, high;
        if (!get_num(p, "min", &low) || !get_num(p, "max", &high)) return c.fail(13, "uniform: missing field");
        md->dist_kind = DK_UNIFORM;
        // rand 0.8 UniformFloat::new panics unless low < high and both finite
        if (!(low < high) || !std::isfinite(low) || !std::isfinite(high) || !std::isfinite(high - low)) {
            md->dist_err = 34;
            return 0;
        }
        const double max_rand = 1.0 - 2.220446049250313e-16;
        double scale = high - low;
        while (scale * max_rand + low >= high) {
            uint64_t b;
            std::memcpy(&b, &scale, 8);
            b -= 1;
            std::memcpy(&scale, &b, 8);
        }
        md->p0 = low;
        md->p1 = scale;
        return 0;
    }
    if (kind == "gamma") {  // rand_distr 0.4 Gamma::new(shape, scale)
        double shape, scale;
        if (!get_num(p, "shape", &shape) || !get_num(p, "scale", &scale)) return c.fail(13, "gamma: missing field");
        if (!(shape > 0.0) || !(scale > 0.0)) {
            md->dist_kind = DK_GAMMA;
            md->dist_err = 16;
            return 0;
        }
        if (shape == 1.0) {  // GammaRepr::One(Exp::new(1 / scale))
            const double lambda = 1.0 / scale;
            md->dist_kind = DK_EXP;
            if (!(lambda > 0.0)) md->dist_err = 16;
            md->p0 = 1.0 / lambda;
            return 0;
        }
        const bool small = shape < 1.0;
        const double sh = small ? shape + 1.0 : shape;
        const double d = sh - 1.0 / 3.0;
        md->dist_kind = DK_GAMMA;
        md->p0 = d;
        md->p1 = 1.0 / std::sqrt(9.0 * d);
        md->p2 = scale;
        md->q0 = small ? 1 : 0;
        const double inv_shape = 1.0 / shape;
        std::memcpy(&md->q1, &inv_shape, 8);
        return 0;
    }
    if (kind == "logNormal") {
        double mu, sigma;
        if (!get_num(p, "mu", &mu) || !get_num(p, "sigma", &sigma)) return c.fail(13, "logNormal: missing field");
        md->dist_kind = DK_LOGNORMAL;
        if (!std::isfinite(sigma)) md->dist_err = 17;
        md->p0 = mu;
        md->p1 = sigma;
        return 0;
    }
    if (kind == "weibull") {
        // the reference passes (shape, scale) to rand_distr's Weibull::new(scale, shape): swapped roles
        double shape, scale;
        if (!get_num(p, "shape", &shape) || !get_num(p, "scale", &scale)) return c.fail(13, "weibull: missing field");
        md->dist_kind = DK_WEIBULL;
        if (!(shape > 0.0) || !(scale > 0.0)) md->dist_err = 19;
        md->p0 = shape;        // acts as rand_distr's scale
        md->p1 = 1.0 / scale;  // acts as rand_distr's 1/shape
        return 0;
    }
    if (kind == "beta") {  // rand_distr 0.4.3 Beta::new: Cheng's BB (min(alpha, beta) > 1) or BC
        double alpha, beta;
        if (!get_num(p, "alpha", &alpha) || !get_num(p, "beta", &beta)) return c.fail(13, "beta: missing field");
        md->dist_kind = DK_BETA;
        if (!(alpha > 0.0) || !(beta > 0.0)) {
            md->dist_err = 14;
            return 0;
        }
        double a = alpha < beta ? alpha : beta, b = alpha < beta ? beta : alpha;
        bool switched = !(alpha < beta);
        uint64_t flags = 0;
        if (!(a > 1.0)) {  // BC works with a = max, b = min
            std::swap(a, b);
            switched = !switched;
            flags |= 1;
        }
        if (switched) flags |= 2;
        md->p0 = a;
        md->p1 = b;
        md->q0 = flags;
        return 0;
    }
    return c.fail(13, "unknown variant `" + kind + "`");
}

void uniform_int_params(uint64_t low, uint64_t high, SimbModelDesc* md) {
    uint64_t range = (high - 1) - low + 1;
    uint64_t reject = (UINT64_MAX - range + 1) % range;
    md->q0 = low;
    md->q1 = range;
    md->q2 = UINT64_MAX - reject;
}

int lower_index(Ctx& c, const Value* v, SimbModelDesc* md, SimbNetDesc* net, uint32_t* n_wcum) {
    if (!v || !v->is_object() || v->obj.size() != 1) return c.fail(13, "expected a one-variant index random variable");
    const std::string& kind = v->obj[0].first;
    const Value& p = v->obj[0].second;
    md->dist_err = 0;
    if (kind == "uniform") {
        const Value* mn = p.get("min");
        const Value* mx = p.get("max");
        if (!mn || !mx || !mn->is_uint || !mx->is_uint) return c.fail(13, "index uniform: min/max must be unsigned integers");
        if (!(mn->u < mx->u)) md->dist_err = 34;  // Uniform::new asserts low < high
        else uniform_int_params(mn->u, mx->u, md);
        return 0;
    }
    if (kind == "weightedIndex") {
        const Value* w = p.get("weights");
        if (!w || !w->is_array()) return c.fail(13, "weightedIndex: missing field `weights`");
        md->flags |= MF_INDEX_WEIGHTED;
        if (w->arr.empty()) {
            md->dist_err = 23;
            return 0;
        }
        md->wtab = (uint16_t)*n_wcum;
        uint64_t total = 0;
        for (size_t i = 0; i < w->arr.size(); i++) {
            if (!w->arr[i].is_uint) return c.fail(13, "weightedIndex: weights must be unsigned integers");
            if (i > 0) {
                if (*n_wcum >= SIMB_MAX_WEIGHTS) return c.fail(32, "too many gateway weights");
                net->wcum[(*n_wcum)++] = total;
            }
            total += w->arr[i].u;
        }
        md->n_w = (uint16_t)(w->arr.size() - 1);
        if (total == 0) md->dist_err = 23;  // AllWeightsZero
        else uniform_int_params(0, total, md);
        return 0;
    }
    return c.fail(13, "unknown variant `" + kind + "`");
}

// Discrete (random_variable.rs:36-50)
int lower_discrete(Ctx& c, const Value* v, SimbModelDesc* md) {
    if (!v || !v->is_object() || v->obj.size() != 1) return c.fail(13, "expected a one-variant discrete random variable");
    const std::string& kind = v->obj[0].first;
    const Value& p = v->obj[0].second;
    md->dist_err = 0;
    if (kind == "geometric") {  // rand_distr 0.4 Geometric::new
        double pr;
        if (!get_num(p, "p", &pr)) return c.fail(13, "geometric: missing field `p`");
        md->dist_kind = DK_GEOMETRIC;
        if (!std::isfinite(pr) || pr < 0.0 || pr > 1.0) {
            md->dist_err = 21;
            return 0;
        }
        md->p0 = pr;
        md->p1 = pr;
        md->q0 = 0;
        if (!(pr == 0.0 || pr >= 2.0 / 3.0)) {  // smallest k with pi = (1-p)^(2^k) <= 1/2
            uint64_t k = 1;
            double pi = (1.0 - pr) * (1.0 - pr);
            while (pi > 0.5) {
                k++;
                pi = pi * pi;
            }
            md->p1 = pi;
            md->q0 = k;
        }
        return 0;
    }
    if (kind == "poisson") {  // rand_distr 0.4 Poisson::new
        double lambda;
        if (!get_num(p, "lambda", &lambda)) return c.fail(13, "poisson: missing field `lambda`");
        md->dist_kind = DK_POISSON;
        if (!(lambda > 0.0)) {
            md->dist_err = 22;
            return 0;
        }
        md->p0 = lambda;
        md->p1 = std::exp(-lambda);
        md->p2 = std::log(lambda);
        const double sq = std::sqrt(2.0 * lambda);
        // utils::log_gamma(1 + lambda), 6-term Lanczos
        const double x = 1.0 + lambda, tmp = x + 5.5;
        static const double co[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                     -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
        double a = 1.000000000190015, denom = x;
        for (double cf : co) {
            denom = denom + 1.0;
            a = a + (cf / denom);
        }
        const double lg = (x + 0.5) * std::log(tmp) - tmp + std::log(2.5066282746310005 * a / x);
        const double magic = lambda * md->p2 - lg;
        std::memcpy(&md->q0, &sq, 8);
        std::memcpy(&md->q1, &magic, 8);
        return 0;
    }
    if (kind == "uniform") {
        const Value* mn = p.get("min");
        const Value* mx = p.get("max");
        if (!mn || !mx || !mn->is_uint || !mx->is_uint) return c.fail(13, "discrete uniform: min/max must be unsigned integers");
        md->dist_kind = DK_UNIFORM_INT;
        if (!(mn->u < mx->u)) md->dist_err = 34;  // Uniform::new asserts low < high
        else uniform_int_params(mn->u, mx->u, md);
        return 0;
    }
    return c.fail(13, "unknown variant `" + kind + "`");
}

struct PendingModel {
    // lowering scratch per model
    uint32_t n_words = 0, n_dwords = 0;
    bool ring0 = false, ring1 = false, ringd = false;
    uint32_t ring_cap = 0;
    double u0 = std::numeric_limits<double>::infinity();
    uint32_t phase0 = 0;
    std::vector<uint32_t> words0;
    bool own_rng = false;              // `rng: {seed}`: one more private word, the draw index of the model's own stream
    std::vector<std::string> preload;  // queue contents from `state`
    std::vector<uint32_t> preload1;    // second u32 channel of those entries (collection count / pass flag / seq|flags)
    std::vector<double> preload_d;     // f64 channel of those entries (Stopwatch: the time present)
    bool has_best = false;             // Stopwatch `state.jobs` held completed jobs: (name, duration, seq) of the best
    std::string best_name;
    double best_dur = 0.0;
    uint32_t best_seq = 0;
    uint32_t preload_next_port = 0;
};

uint32_t clamp_u32(uint64_t v, uint32_t hi) { return v > hi ? hi : (uint32_t)v; }

}  // namespace

// one random variable of the given family (0 Continuous, 1 Boolean, 2 Discrete, 3 Index) into models[0] (+ wcum)
int lower_random_variable(int family, const char* variable_json, SimbNetDesc* net, std::string* err) {
    Ctx c{err};
    if (!variable_json || !net) return c.fail(36, "null argument");
    Value v;
    json::Parser p(variable_json);
    if (!p.parse(v)) return c.fail(13, "random variable: " + p.err);
    std::memset(net, 0, sizeof *net);
    SimbModelDesc* md = &net->models[0];
    md->dist_kind = DK_NONE;
    uint32_t n_wcum = 0;
    switch (family) {
        case FAM_CONTINUOUS: return lower_continuous(c, &v, md);
        case FAM_DISCRETE: return lower_discrete(c, &v, md);
        case FAM_INDEX: return lower_index(c, &v, md, net, &n_wcum);
        case FAM_BOOLEAN: {  // Boolean::Bernoulli (random_variable.rs:30-34); rand 0.8 Bernoulli::new
            const Value* b = v.get("bernoulli");
            double pr;
            if (!b || !get_num(*b, "p", &pr)) return c.fail(13, "expected {\"bernoulli\": {\"p\": ..}}");
            if (!(pr >= 0.0 && pr < 1.0)) {
                if (pr == 1.0) md->flags |= MF_BERNOULLI_ALWAYS;
                else md->dist_err = 20;
            } else {
                md->q0 = (uint64_t)(pr * (2.0 * (double)(1ull << 63)));
            }
            return 0;
        }
    }
    return c.fail(36, "unknown random variable family");
}

int lower_network(const char* models_json, const char* connectors_json, const LowerOptions& opt, HostNet* out,
                  std::string* err) {
    Ctx c{err};
    if (!models_json || !connectors_json || !out) return c.fail(36, "null argument");
    Value jm, jc;
    {
        json::Parser p(models_json);
        if (!p.parse(jm)) return c.fail(13, "models: " + p.err);
        json::Parser q(connectors_json);
        if (!q.parse(jc)) return c.fail(13, "connectors: " + q.err);
    }
    if (!jm.is_array() || !jc.is_array()) return c.fail(13, "models and connectors must be JSON arrays");
    // (an empty model list is a simulation too: Simulation::post does not look, and every step adds +INF to the clock,
    // mod.rs:225-236)

    HostNet& hn = *out;
    hn = HostNet();
    SimbNetDesc& net = hn.desc;
    std::memset(&net, 0, sizeof net);
    // ---- pass 0: flatten -- the components of a Coupled model (coupled.rs:14-25) take its place in the model list ----
    struct FlatSrc {
        const Value* j;
        int top, coupled;
    };
    std::vector<FlatSrc> flat;
    for (size_t i = 0; i < jm.arr.size(); i++) {
        const Value& j = jm.arr[i];
        if (!j.is_object()) return c.fail(13, "model entries must be objects");
        HostTop top;
        std::string type;
        if (!get_str(j, "id", &top.id)) return c.fail(13, "model: missing field `id`");
        if (!get_str(j, "type", &type)) return c.fail(13, "model " + top.id + ": missing field `type`");
        if (type != "Coupled") {
            top.flat = (int)flat.size();
            flat.push_back(FlatSrc{&j, (int)i, -1});
            hn.tops.push_back(top);
            continue;
        }
        if (hn.coupled.size() >= SIMB_MAX_COUPLED) return c.fail(32, "more than 4 Coupled models");
        HostCoupled hc;
        hc.id = top.id;
        const Value* pin = j.get("portsIn");
        const Value* pout = j.get("portsOut");
        const Value* comps = j.get("components");
        if (!pin || !pout || !get_str_list(*pin, "flowPaths", &hc.ports_in) || !get_str_list(*pout, "flowPaths", &hc.ports_out))
            return c.fail(13, hc.id + ": portsIn.flowPaths / portsOut.flowPaths missing");
        if (!comps || !comps->is_array()) return c.fail(13, hc.id + ": missing field `components`");
        hc.first = (int)flat.size();
        for (auto& cj : comps->arr) {
            std::string ct;
            if (!cj.is_object() || !get_str(cj, "type", &ct)) return c.fail(13, hc.id + ": component entries must be model objects");
            if (ct == "Coupled")
                return c.fail(33, hc.id + ": a Coupled model inside a Coupled model cannot run on the device (one level of coupling is lowered)");
            flat.push_back(FlatSrc{&cj, (int)i, (int)hn.coupled.size()});
        }
        hc.count = (int)flat.size() - hc.first;
        auto couplings = [&](const char* key, bool has_source, bool has_target, std::vector<HostCoupled::Coupling>* dst) -> bool {
            const Value* a = j.get(key);
            if (!a || !a->is_array()) return false;
            for (auto& e : a->arr) {
                HostCoupled::Coupling cp;
                if (!e.is_object() || !get_str(e, "sourcePort", &cp.source_port) || !get_str(e, "targetPort", &cp.target_port)) return false;
                if (has_source && !get_str(e, "sourceID", &cp.source_id)) return false;
                if (has_target && !get_str(e, "targetID", &cp.target_id)) return false;
                dst->push_back(cp);
            }
            return true;
        };
        if (!couplings("externalInputCouplings", false, true, &hc.eic) || !couplings("externalOutputCouplings", true, false, &hc.eoc) ||
            !couplings("internalCouplings", true, true, &hc.ic))
            return c.fail(13, hc.id + ": externalInputCouplings / externalOutputCouplings / internalCouplings missing or malformed");
        if (const Value* cst = j.get("state"))
            if (const Value* pk = cst->get("parkedMessages")) {
                if (!pk->is_array()) return c.fail(13, hc.id + ": state.parkedMessages must be an array");
                for (auto& e : pk->arr) {
                    HostCoupled::Parked pm_;
                    if (!e.is_object() || !get_str(e, "componentId", &pm_.component_id) || !get_str(e, "port", &pm_.port) ||
                        !get_str(e, "content", &pm_.content))
                        return c.fail(13, hc.id + ": state.parkedMessages entries need componentId, port and content");
                    hc.parked.push_back(pm_);
                }
            }
        top.coupled = (int)hn.coupled.size();
        hn.coupled.push_back(hc);
        hn.tops.push_back(top);
    }
    if (flat.size() > SIMB_MAX_MODELS) return c.fail(32, "more than 32 models");
    net.n_models = (uint32_t)flat.size();
    std::vector<PendingModel> pm(flat.size());
    uint32_t n_wcum = 0;
    const uint32_t qcap = opt.queue_capacity ? clamp_u32(opt.queue_capacity, 65535) : 64;
    // transient queues (LoadBalancer, Gate, gateways: they drain within a step or two) default to 16
    // slots; an explicit simb_config.queue_capacity applies to every queue
    const uint32_t tcap = opt.queue_capacity ? (qcap < 4 ? 4 : qcap) : 16;
    bool overflow = false;

    // ---- pass 1: models -------------------------------------------------------------------------
    for (size_t i = 0; i < flat.size(); i++) {
        const Value& j = *flat[i].j;
        if (!j.is_object()) return c.fail(13, "model entries must be objects");
        HostModel hm;
        hm.top = flat[i].top;
        hm.coupled = flat[i].coupled;
        if (!get_str(j, "id", &hm.id)) return c.fail(13, "model: missing field `id`");
        if (!get_str(j, "type", &hm.type)) return c.fail(13, "model " + hm.id + ": missing field `type`");
        if (const Value* sr = j.get("storeRecords")) hm.store_records = sr->kind == Value::Bool && sr->b;
        SimbModelDesc& md = net.models[i];
        PendingModel& p = pm[i];
        md.dist_kind = DK_NONE;
        md.cap = SIMB_CAP_INF;
        if (hm.store_records) md.flags |= MF_STORE_RECORDS;
        const Value* pin = j.get("portsIn");
        const Value* pout = j.get("portsOut");
        const Value* state = j.get("state");
        auto need_ports = [&](bool in, bool outp) -> bool {
            return (!in || (pin && pin->is_object())) && (!outp || (pout && pout->is_object()));
        };
        auto port = [&](const Value* obj, const char* key, std::vector<std::string>* dst) -> bool {
            std::string s;
            if (!obj || !get_str(*obj, key, &s)) return false;
            dst->push_back(s);
            return true;
        };
        const std::string& t = hm.type;
        if (t == "Generator") {  // generator.rs:24-40
            md.type = MT_GENERATOR;
            if (!need_ports(false, true) || !port(pout, "job", &hm.ports_out)) return c.fail(13, hm.id + ": portsOut.job missing");
            int e = lower_continuous(c, j.get("messageInterdepartureTime"), &md);
            if (e) return e;
            md.prefix = hn.intern(hm.ports_out[0], &overflow) << 24;
            p.n_words = 1;  // last_job
            p.words0 = {0};
            p.u0 = 0.0;     // State::default (:60-70)
            if (state && state->is_object()) {
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "Generating" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                double lj;
                if (get_num(*state, "lastJob", &lj)) p.words0[0] = (uint32_t)lj;
            }
        } else if (t == "Processor") {  // processor.rs:24-38
            md.type = MT_PROCESSOR;
            if (!need_ports(true, true) || !port(pin, "job", &hm.ports_in) || !port(pout, "job", &hm.ports_out))
                return c.fail(13, hm.id + ": portsIn.job / portsOut.job missing");
            int e = lower_continuous(c, j.get("serviceTime"), &md);
            if (e) return e;
            uint64_t cap = UINT64_MAX;  // max_usize default (:28-29,40-42)
            if (const Value* qc = j.get("queueCapacity")) {
                if (!qc->is_uint) return c.fail(13, hm.id + ": queueCapacity must be an unsigned integer");
                cap = qc->u;
            }
            p.ring0 = true;
            if (cap <= 65535) {  // honoured exactly (the ring's head / length fields are 16 bits)
                md.cap = (uint32_t)cap;
                p.ring_cap = cap ? (uint32_t)cap : 1;
            } else {
                // larger than any device ring, usize::MAX included: the queue is never "full" (no Drop records); the
                // ring holds simb_config.queue_capacity jobs (default 64) and raises SIMB_ERR_CAPACITY beyond that
                md.cap = SIMB_CAP_INF;
                p.ring_cap = qcap;
            }
            p.n_words = 1;  // head|len
            p.words0 = {0};
            if (state && state->is_object()) {
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "Active" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                get_str_list(*state, "queue", &p.preload);
                if (p.preload.size() > 65535) return c.fail(32, hm.id + ": preloaded queue too long");
                if (p.preload.size() > p.ring_cap) p.ring_cap = (uint32_t)p.preload.size();
            }
        } else if (t == "Storage") {  // storage.rs:15-24
            md.type = MT_STORAGE;
            if (!need_ports(true, true) || !port(pin, "put", &hm.ports_in) || !port(pin, "get", &hm.ports_in) ||
                !port(pout, "stored", &hm.ports_out))
                return c.fail(13, hm.id + ": portsIn.put/get or portsOut.stored missing");
            p.n_words = 1;
            p.words0 = {SIMB_CONTENT_NONE};
            if (state && state->is_object()) {
                std::string ph, job;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "JobFetch" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                if (get_str(*state, "job", &job)) p.preload = {job};  // interned after all generators
            }
        } else if (t == "LoadBalancer") {  // load_balancer.rs:15-24
            md.type = MT_LOAD_BALANCER;
            if (!need_ports(true, true) || !port(pin, "job", &hm.ports_in) || !get_str_list(*pout, "flowPaths", &hm.ports_out))
                return c.fail(13, hm.id + ": portsIn.job / portsOut.flowPaths missing");
            p.ring0 = true;
            p.ring_cap = tcap;
            p.n_words = 2;  // head|len, next_port_out
            p.words0 = {0, 0};
            if (state && state->is_object()) {
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "LoadBalancing" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                double np;
                if (get_num(*state, "nextPortOut", &np)) p.words0[1] = (uint32_t)np;
                get_str_list(*state, "jobs", &p.preload);
                if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
            }
        } else if (t == "Gate") {  // gate.rs:19-28
            md.type = MT_GATE;
            if (!need_ports(true, true) || !port(pin, "job", &hm.ports_in) || !port(pin, "activation", &hm.ports_in) ||
                !port(pin, "deactivation", &hm.ports_in) || !port(pout, "job", &hm.ports_out))
                return c.fail(13, hm.id + ": gate ports missing");
            p.ring0 = true;
            p.ring_cap = tcap;
            p.n_words = 1;
            p.words0 = {0};
            if (state && state->is_object()) {
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "Closed" ? 1 : ph == "Pass" ? 2 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                get_str_list(*state, "jobs", &p.preload);
                if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
            }
        } else if (t == "ExclusiveGateway") {  // exclusive_gateway.rs:20-32
            md.type = MT_EXCLUSIVE_GATEWAY;
            if (!need_ports(true, true) || !get_str_list(*pin, "flowPaths", &hm.ports_in) ||
                !get_str_list(*pout, "flowPaths", &hm.ports_out))
                return c.fail(13, hm.id + ": flowPaths missing");
            int e = lower_index(c, j.get("portWeights"), &md, &net, &n_wcum);
            if (e) return e;
            p.ring0 = true;
            p.ring_cap = tcap;
            p.n_words = 1;
            p.words0 = {0};
            if (state && state->is_object()) {  // exclusive_gateway.rs:55-78
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "Pass" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                get_str_list(*state, "jobs", &p.preload);
                if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
            }
        } else if (t == "ParallelGateway") {  // parallel_gateway.rs:19-28
            md.type = MT_PARALLEL_GATEWAY;
            if (!need_ports(true, true) || !get_str_list(*pin, "flowPaths", &hm.ports_in) ||
                !get_str_list(*pout, "flowPaths", &hm.ports_out))
                return c.fail(13, hm.id + ": flowPaths missing");
            p.ring0 = p.ring1 = true;  // (content, count) table
            p.ring_cap = qcap;
            p.n_words = 1;  // entries
            p.words0 = {0};
            if (state && state->is_object()) {  // parallel_gateway.rs:47-53: collections in the order of the document
                get_num(*state, "untilNextEvent", &p.u0);
                if (const Value* col = state->get("collections")) {
                    if (!col->is_object()) return c.fail(13, hm.id + ": state.collections must be an object");
                    for (auto& kv : col->obj) {
                        if (!kv.second.is_uint) return c.fail(13, hm.id + ": state.collections values must be unsigned integers");
                        p.preload.push_back(kv.first);
                        p.preload1.push_back(clamp_u32(kv.second.u, 0xFFFFFFFFu));
                    }
                    if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
                }
            }
        } else if (t == "StochasticGate") {  // stochastic_gate.rs:19-31
            md.type = MT_STOCHASTIC_GATE;
            if (!need_ports(true, true) || !port(pin, "job", &hm.ports_in) || !port(pout, "job", &hm.ports_out))
                return c.fail(13, hm.id + ": ports missing");
            const Value* pd = j.get("passDistribution");
            const Value* b = pd ? pd->get("bernoulli") : nullptr;
            double pr;
            if (!b || !get_num(*b, "p", &pr)) return c.fail(13, hm.id + ": passDistribution.bernoulli.p missing");
            if (!(pr >= 0.0 && pr < 1.0)) {  // rand 0.8 Bernoulli::new
                if (pr == 1.0) md.flags |= MF_BERNOULLI_ALWAYS;
                else md.dist_err = 20;
            } else {
                md.q0 = (uint64_t)(pr * (2.0 * (double)(1ull << 63)));
            }
            p.ring0 = p.ring1 = true;  // (content, pass)
            p.ring_cap = tcap;
            p.n_words = 1;
            p.words0 = {0};
            if (state && state->is_object()) {  // stochastic_gate.rs:50-73: jobs: [{content, pass}]
                get_num(*state, "untilNextEvent", &p.u0);
                if (const Value* jobs = state->get("jobs")) {
                    if (!jobs->is_array()) return c.fail(13, hm.id + ": state.jobs must be an array");
                    for (auto& jb : jobs->arr) {
                        std::string content;
                        const Value* ps = jb.get("pass");
                        if (!jb.is_object() || !get_str(jb, "content", &content) || !ps || ps->kind != Value::Bool)
                            return c.fail(13, hm.id + ": state.jobs entries need `content` and `pass`");
                        p.preload.push_back(content);
                        p.preload1.push_back(ps->b ? 1u : 0u);
                    }
                    if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
                }
            }
        } else if (t == "Stopwatch") {  // stopwatch.rs:21-32
            md.type = MT_STOPWATCH;
            if (!need_ports(true, true) || !port(pin, "start", &hm.ports_in) || !port(pin, "stop", &hm.ports_in) ||
                !port(pin, "metric", &hm.ports_in) || !port(pout, "job", &hm.ports_out))
                return c.fail(13, hm.id + ": stopwatch ports missing");
            std::string metric;
            if (get_str(j, "metric", &metric)) {
                if (metric == "Maximum") md.flags |= MF_METRIC_MAX;
                else if (metric != "Minimum") return c.fail(13, hm.id + ": unknown metric `" + metric + "`");
            }
            p.ring0 = p.ring1 = p.ringd = true;  // incomplete-job table
            p.ring_cap = qcap;
            p.n_words = 4;  // entries, best content, next seq, best seq
            p.words0 = {0, SIMB_CONTENT_NONE, 0, 0};
            p.n_dwords = 1;  // best duration
            if (state && state->is_object()) {  // stopwatch.rs:66-93: jobs: [{name, start, stop}] (null = None)
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "JobFetch" ? 1 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                if (const Value* jobs = state->get("jobs")) {
                    if (!jobs->is_array()) return c.fail(13, hm.id + ": state.jobs must be an array");
                    uint32_t seq = 0;
                    for (auto& jb : jobs->arr) {
                        std::string name;
                        if (!jb.is_object() || !get_str(jb, "name", &name)) return c.fail(13, hm.id + ": state.jobs entries need `name`");
                        double t0 = 0.0, t1 = 0.0;
                        // Option<f64>: null / absent = None (not the "null = infinity" reading of get_num)
                        const Value* vs = jb.get("start");
                        const Value* ve = jb.get("stop");
                        const bool hs = vs && vs->kind != Value::Null && get_num(jb, "start", &t0);
                        const bool he = ve && ve->kind != Value::Null && get_num(jb, "stop", &t1);
                        if (hs && he) {  // completed: only the best one matters (release_minimum / _maximum :216-250,
                                         // strict compare, first in insertion order)
                            const double dur = t1 - t0;
                            const bool mx = (md.flags & MF_METRIC_MAX) != 0;
                            if (!p.has_best ? (mx ? dur > -std::numeric_limits<double>::infinity() : dur < std::numeric_limits<double>::infinity())
                                            : (mx ? dur > p.best_dur : dur < p.best_dur)) {
                                p.has_best = true;
                                p.best_name = name;
                                p.best_dur = dur;
                                p.best_seq = seq;
                            }
                        } else {  // incomplete: a table entry (name, seq << 2 | has_stop << 1 | has_start, the time present)
                            p.preload.push_back(name);
                            p.preload1.push_back((seq << 2) | (he ? 2u : 0u) | (hs ? 1u : 0u));
                            p.preload_d.push_back(hs ? t0 : he ? t1 : 0.0);
                        }
                        seq++;
                    }
                    p.words0[2] = seq;
                    if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
                }
            }
        } else if (t == "Batcher") {  // batcher.rs:22-33
            md.type = MT_BATCHER;
            if (!need_ports(true, true) || !port(pin, "job", &hm.ports_in) || !port(pout, "job", &hm.ports_out))
                return c.fail(13, hm.id + ": ports missing");
            double mbt;
            const Value* mbs = j.get("maxBatchSize");
            if (!get_num(j, "maxBatchTime", &mbt) || !mbs || !mbs->is_uint) return c.fail(13, hm.id + ": maxBatchTime/maxBatchSize missing");
            md.p0 = mbt;
            md.cap = clamp_u32(mbs->u, 0xFFFFFFFEu);
            p.ring0 = true;
            p.ring_cap = clamp_u32(2ull * (uint64_t)md.cap + 8ull, 65535);
            p.n_words = 1;
            p.words0 = {0};
            if (state && state->is_object()) {  // batcher.rs:46-74
                std::string ph;
                if (get_str(*state, "phase", &ph)) p.phase0 = ph == "Batching" ? 1 : ph == "Release" ? 2 : 0;
                get_num(*state, "untilNextEvent", &p.u0);
                get_str_list(*state, "jobs", &p.preload);
                if (p.preload.size() > p.ring_cap) p.ring_cap = clamp_u32(p.preload.size(), 65535);
            }
        } else if (t == "Passive") {  // sim/tests/custom.rs:18-86
            md.type = MT_PASSIVE;
            if (pin && pin->is_object()) port(pin, "job", &hm.ports_in);
        } else {
            return c.fail(33, "model type `" + t + "` cannot run on the device (custom models are out of scope)");
        }
        if (const Value* rg = j.get("rng")) {  // the model's own generator (generator.rs:39, processor.rs, exclusive_gateway.rs,
                                               // stochastic_gate.rs; #[serde(skip)] in the reference: an extension here)
            if (rg->kind != Value::Null) {
                const Value* sd = rg->get("seed");
                if (!sd || !sd->is_uint) return c.fail(13, hm.id + ": rng must be {\"seed\": <unsigned integer>}");
                if (md.type != MT_GENERATOR && md.type != MT_PROCESSOR && md.type != MT_EXCLUSIVE_GATEWAY && md.type != MT_STOCHASTIC_GATE)
                    return c.fail(13, hm.id + ": this model type has no rng field");
                if (opt.rng_kind != 0) return c.fail(36, hm.id + ": a per-model rng needs the Philox streams (rng_kind 0)");
                md.rng_seed = sd->u;
                p.own_rng = true;
                p.words0.resize(p.n_words, 0u);
                p.words0.push_back(0u);
                p.n_words += 1;
            }
        }
        md.n_in = (uint8_t)(hm.ports_in.size() > 254 ? 254 : hm.ports_in.size());
        md.n_out = (uint8_t)hm.ports_out.size();
        if (hm.ports_out.size() > 64) return c.fail(32, hm.id + ": too many output ports");
        hn.models.push_back(hm);
    }
    // (two posted models may share an id: Simulation::post does not look, and step() hands a message to every model
    // whose id matches, mod.rs:203-221 -- see HostNet::deliveries)
    for (auto& hc : hn.coupled)
        for (int i = hc.first; i < hc.first + hc.count; i++)
            for (int k = i + 1; k < hc.first + hc.count; k++)
                if (hn.models[i].id == hn.models[k].id) return c.fail(1, hc.id + ": duplicate component id `" + hn.models[i].id + "`");
    net.n_coupled = (uint32_t)hn.coupled.size();
    for (size_t i = 0; i < hn.models.size(); i++) {
        const int cp = hn.models[i].coupled;
        net.coupled_of[i] = cp < 0 ? 0xFF : (uint8_t)cp;
        net.group_base[i] = cp < 0 ? (uint8_t)i : (uint8_t)hn.coupled[cp].first;
    }

    // ---- connectors (coupling.rs:7-17) ----------------------------------------------------------
    for (auto& j : jc.arr) {
        if (!j.is_object()) return c.fail(13, "connector entries must be objects");
        HostConnector hc;
        if (!get_str(j, "id", &hc.id) || !get_str(j, "sourceID", &hc.source_id) || !get_str(j, "targetID", &hc.target_id) ||
            !get_str(j, "sourcePort", &hc.source_port) || !get_str(j, "targetPort", &hc.target_port))
            return c.fail(13, "connector: missing field (id, sourceID, targetID, sourcePort, targetPort)");
        hn.connectors.push_back(hc);
    }
    if (hn.connectors.size() > 0xFFFE) return c.fail(32, "too many connectors");
    net.n_connectors = (uint32_t)hn.connectors.size();

    // ---- routes: (model, out-port) -> deliveries, in the reference's order (mod.rs:155-182, coupled.rs:133-164) -------
    uint32_t n_routes = 0, n_outports = 0;
    bool bad_eic = false;
    auto push_route = [&](int tgt, uint32_t port, uint32_t conn) -> bool {
        if (n_routes >= SIMB_MAX_ROUTES) return false;
        SimbRoute r;
        r.tgt_model = tgt < 0 ? 0xFF : (uint8_t)tgt;  // unknown target: never delivered, but it still makes the next
                                                      // step a zero-time one
        r.tgt_port = tgt < 0 ? SIMB_PORT_UNKNOWN : (uint8_t)port;
        r.connector = (uint16_t)conn;
        net.routes[n_routes++] = r;
        return true;
    };
    // the deliveries of one message over top-level connector ci: the first is the Message of the reference (traced,
    // counted), the others -- a second model with the same id, further components of a Coupled model -- ride along
    // flagged RT_CONT; no model of that id: the message is kept for one step and never delivered
    auto send_over = [&](size_t ci) -> bool {
        const HostConnector& hc = hn.connectors[ci];
        const std::vector<HostNet::Delivery> dv = hn.deliveries(hc.target_id, hc.target_port, &bad_eic);
        if (bad_eic) return false;
        if (dv.empty()) return push_route(-1, 0, (uint32_t)ci);
        for (size_t k = 0; k < dv.size(); k++)
            if (!push_route(dv[k].model, dv[k].port, (uint32_t)ci | (k ? RT_CONT : 0u))) return false;
        return true;
    };
    for (size_t i = 0; i < hn.models.size(); i++) {
        SimbModelDesc& md = net.models[i];
        md.out_base = (uint16_t)n_outports;
        for (size_t k = 0; k < hn.models[i].ports_out.size(); k++) {
            if (n_outports >= SIMB_MAX_OUT_PORTS) return c.fail(32, "too many output ports in the network");
            net.route_begin[n_outports++] = (uint16_t)n_routes;
            const std::string& pname = hn.models[i].ports_out[k];
            bool ok = true;
            if (hn.models[i].coupled < 0) {
                for (size_t ci = 0; ci < hn.connectors.size() && ok; ci++)
                    if (hn.connectors[ci].source_id == hn.models[i].id && hn.connectors[ci].source_port == pname) ok = send_over(ci);
            } else {
                const int cpl = hn.models[i].coupled;
                const HostCoupled& C = hn.coupled[cpl];
                for (auto& ic : C.ic) {  // internal couplings: parked until the coupled model next fires (:238-247)
                    if (!ok || ic.source_id != hn.models[i].id || ic.source_port != pname) continue;
                    const int comp = hn.find_component(cpl, ic.target_id);
                    if (comp < 0) return c.fail(1, C.id + ": internal coupling names the unknown component `" + ic.target_id + "`");
                    ok = push_route(comp, hn.resolve_in_port(comp, ic.target_port), RT_PARK | ((uint32_t)cpl << 12));
                }
                for (auto& eo : C.eoc) {  // external output couplings, then the connectors of the coupled model (:248-252)
                    if (eo.source_id != hn.models[i].id || eo.source_port != pname) continue;
                    for (size_t ci = 0; ci < hn.connectors.size() && ok; ci++)
                        if (hn.connectors[ci].source_id == C.id && hn.connectors[ci].source_port == eo.target_port) ok = send_over(ci);
                }
            }
            if (bad_eic) return c.fail(1, "an external input coupling names a component its Coupled model does not have");
            if (!ok) return c.fail(32, "too many routes");
        }
        net.route_begin[n_outports] = (uint16_t)n_routes;
    }
    net.n_routes = n_routes;

    // ---- bursts: the largest number of jobs a model can send on one port / receive in ONE step ---------
    // Most models send one job per firing; a Gate or an ExclusiveGateway sends everything that arrived in the step, a
    // Batcher a whole batch.  Propagated along the routes to a fixed point (cycles: clamped), these bounds size the
    // transient queues and the mailbox, so that e.g. Batcher(20) -> Gate -> Storage runs as it does in the reference
    // (whose Vecs are unbounded) instead of overflowing a 16-slot ring.
    const uint32_t BURST_MAX = 4096;
    std::vector<uint32_t> burst_out(hn.models.size(), 1u), burst_in(hn.models.size(), 0u);
    auto route_source = [&](uint32_t r) -> int {
        for (size_t i = 0; i < hn.models.size(); i++) {
            const SimbModelDesc& md = net.models[i];
            if (r >= net.route_begin[md.out_base] && r < net.route_begin[md.out_base + md.n_out]) return (int)i;
        }
        return -1;
    };
    for (size_t pass = 0; pass < hn.models.size() + 2; pass++) {
        bool changed = false;
        for (size_t i = 0; i < hn.models.size(); i++) {
            uint64_t in = 0;
            for (uint32_t r = 0; r < n_routes; r++)
                if (net.routes[r].tgt_model == i) {
                    const int sm = route_source(r);
                    if (sm >= 0) in += burst_out[sm];
                }
            const uint32_t bi = clamp_u32(in, BURST_MAX);
            uint32_t bo = 1;
            const uint8_t ty = net.models[i].type;
            if (ty == MT_GATE || ty == MT_EXCLUSIVE_GATEWAY) bo = bi ? bi : 1;
            else if (ty == MT_BATCHER) bo = clamp_u32(net.models[i].cap, BURST_MAX);
            if (bi != burst_in[i] || bo != burst_out[i]) changed = true;
            burst_in[i] = bi;
            burst_out[i] = bo;
        }
        if (!changed) break;
    }
    if (!opt.queue_capacity) {  // an explicit simb_config.queue_capacity is taken as given
        for (size_t i = 0; i < hn.models.size(); i++) {
            uint32_t want = 0;
            switch (net.models[i].type) {
                case MT_GATE:
                case MT_EXCLUSIVE_GATEWAY: want = burst_in[i]; break;                         // emptied in the step they fill
                case MT_LOAD_BALANCER:
                case MT_STOCHASTIC_GATE: want = 2u * burst_in[i]; break;                      // one job leaves per step
                case MT_BATCHER: want = clamp_u32((uint64_t)net.models[i].cap + burst_in[i] + 8ull, 65535); break;
                default: break;
            }
            if (want > 65535u) want = 65535u;
            if (want > pm[i].ring_cap) pm[i].ring_cap = want;
            // a ParallelGateway with ONE input completes every collection on arrival and sends it in the same step
            // (parallel_gateway.rs:100-143): its table never holds more than one step's arrivals, not queue_capacity
            if (net.models[i].type == MT_PARALLEL_GATEWAY && net.models[i].n_in == 1) {
                uint32_t cap = std::max<uint32_t>(8u, 2u * burst_in[i]);
                if (cap < pm[i].preload.size()) cap = (uint32_t)pm[i].preload.size();
                if (cap < pm[i].ring_cap) pm[i].ring_cap = cap;
            }
        }
    }
    // mailbox bound: messages one firing of a model can emit = jobs per firing x deliveries per port (a
    // ParallelGateway sends one job on every port); parked messages of a Coupled model: the same sum over its
    // internal couplings
    uint64_t pending_bound = 0;
    std::vector<uint64_t> park_bound(hn.coupled.size(), 0);
    for (size_t i = 0; i < hn.models.size(); i++) {
        const SimbModelDesc& md = net.models[i];
        uint64_t send_max = 0, send_total = 0, park_max = 0, park_total = 0;
        for (uint32_t k = 0; k < md.n_out; k++) {
            uint64_t sends = 0, parks = 0;
            for (uint32_t r = net.route_begin[md.out_base + k]; r < net.route_begin[md.out_base + k + 1]; r++)
                (net.routes[r].connector & RT_PARK) ? parks++ : sends++;
            send_max = std::max(send_max, sends);
            park_max = std::max(park_max, parks);
            send_total += sends;
            park_total += parks;
        }
        const bool pg = md.type == MT_PARALLEL_GATEWAY;
        pending_bound += pg ? send_total : send_max * burst_out[i];
        if (hn.models[i].coupled >= 0) park_bound[hn.models[i].coupled] += pg ? park_total : park_max * burst_out[i];
    }
    // (the mailbox fill and the sequence number are 8-bit fields: 255 messages in flight per replication at most;
    // a replication that needs more reports SIMB_ERR_CAPACITY)
    uint32_t mp = opt.max_pending ? opt.max_pending : (pending_bound ? clamp_u32(pending_bound, 255) : 1);
    if (mp > 255) mp = 255;
    net.max_pending = mp;

    // ---- statistic model ------------------------------------------------------------------------
    if (!opt.stat_model_id.empty()) {
        int sm = hn.find_model(opt.stat_model_id);
        if (sm < 0) return c.fail(2, "stat_model_id `" + opt.stat_model_id + "` is not a model of this simulation");
        if (net.models[sm].type != MT_PROCESSOR) return c.fail(1, "stat_model_id must name a Processor");
        net.models[sm].flags |= MF_STAT;
        pm[sm].ringd = true;
        hn.stat_model = sm;
    }

    // ---- state tile layout ------------------------------------------------------------------------
    uint16_t dr = 0, wr = 0;
    net.drow_time = dr++;
    net.drow_until = dr;
    dr += (uint16_t)net.n_models;
    if (hn.stat_model >= 0) net.drow_wait = dr++;
    if (hn.stat_model >= 0 && opt.batch_size > 0) {
        net.drow_bm = dr;
        dr += 3;
        net.bm_warmup = opt.batch_warmup;
        net.bm_size = opt.batch_size;
    }
    net.wrow_ctl = wr++;
    net.wrow_rng = wr;
    wr += (opt.rng_kind == 1) ? 4 : 1;
    net.wrow_steps = wr++;
    net.wrow_events = wr++;
    if (opt.instrument) {
        net.wrow_hash = wr;
        wr += 2;
    }
    net.wrow_dmask = wr++;
    net.wrow_phase = wr;
    net.n_phase_words = (uint16_t)((net.n_models + 15) / 16);
    wr += net.n_phase_words;
    if (hn.stat_model >= 0) net.wrow_waitn = wr++;
    for (size_t i = 0; i < pm.size(); i++) {
        net.models[i].wrow = wr;
        wr += (uint16_t)pm[i].n_words;
        if (pm[i].own_rng) net.models[i].rng_row = (uint16_t)(wr - 1);
        if (pm[i].n_dwords) {
            net.models[i].drow = dr;
            dr += (uint16_t)pm[i].n_dwords;
        }
    }
    for (size_t k = 0; k < hn.coupled.size(); k++) net.coupled[k].wrow = wr++;  // parked-message count
    net.wrow_pend = wr;
    net.pend_tile = (uint16_t)(net.max_pending < SIMB_MAIL_TILE_SLOTS ? net.max_pending : SIMB_MAIL_TILE_SLOTS);
    wr += (uint16_t)(2 * net.pend_tile);
    net.n_drows = dr;
    net.n_wrows = wr;

    // ---- rings ------------------------------------------------------------------------------------
    uint32_t rw = 0, rd = 0;
    for (size_t i = 0; i < pm.size(); i++) {
        SimbModelDesc& md = net.models[i];
        md.ring_cap = pm[i].ring_cap;
        if (pm[i].ring0) {
            md.ring_w0 = rw;
            rw += md.ring_cap;
        }
        if (pm[i].ring1) {
            md.ring_w1 = rw;
            rw += md.ring_cap;
        }
        if (pm[i].ringd) {
            md.ring_d = rd;
            rd += md.ring_cap;
        }
    }
    for (size_t k = 0; k < hn.coupled.size(); k++) {  // parked messages: (content, target component | port << 8)
        SimbCoupledDesc& cd = net.coupled[k];
        cd.first = (uint8_t)hn.coupled[k].first;
        cd.count = (uint8_t)hn.coupled[k].count;
        // what one firing parks stays until the next firing, which first delivers it: twice the per-firing bound
        cd.ring_cap = opt.queue_capacity ? tcap : clamp_u32(std::max<uint64_t>(16, 2 * park_bound[k]), 65535);
        if (cd.ring_cap < hn.coupled[k].parked.size()) cd.ring_cap = clamp_u32(hn.coupled[k].parked.size(), 65535);
        cd.ring_w0 = rw;
        rw += cd.ring_cap;
        cd.ring_w1 = rw;
        rw += cd.ring_cap;
    }
    hn.ring_w_slots = rw;
    hn.ring_d_slots = rd;

    // ---- initial row values ---------------------------------------------------------------------
    hn.d_init.assign(net.n_drows, 0.0);
    hn.w_init.assign(net.n_wrows, 0u);
    uint32_t phase_words[2] = {0, 0};
    for (size_t i = 0; i < pm.size(); i++) {
        const SimbModelDesc& md = net.models[i];
        hn.d_init[net.drow_until + i] = pm[i].u0;
        phase_words[i / 16] |= (pm[i].phase0 & 3u) << (2 * (i % 16));
        for (size_t k = 0; k < pm[i].words0.size(); k++) hn.w_init[md.wrow + k] = pm[i].words0[k];
        if (md.type == MT_STOPWATCH) hn.d_init[md.drow] = 0.0;
        if (md.type == MT_STORAGE) {
            if (!pm[i].preload.empty()) hn.w_init[md.wrow] = hn.content_code(pm[i].preload[0], &overflow);
        } else if (!pm[i].preload.empty()) {  // `state.queue` / `state.jobs` injected on input (tests/web.rs:30-46)
            for (size_t k = 0; k < pm[i].preload.size(); k++) {
                hn.ring_preload.push_back(RingPreload{md.ring_w0 + (uint32_t)k, hn.content_code(pm[i].preload[k], &overflow)});
                if (k < pm[i].preload1.size()) hn.ring_preload.push_back(RingPreload{md.ring_w1 + (uint32_t)k, pm[i].preload1[k]});
                if (k < pm[i].preload_d.size()) {  // an f64 slot travels as two words (RING_PRELOAD_F64)
                    uint64_t bits;
                    std::memcpy(&bits, &pm[i].preload_d[k], 8);
                    hn.ring_preload.push_back(RingPreload{RING_PRELOAD_F64 | (md.ring_d + (uint32_t)k), (uint32_t)bits});
                    hn.ring_preload.push_back(RingPreload{RING_PRELOAD_F64 | RING_PRELOAD_HI | (md.ring_d + (uint32_t)k), (uint32_t)(bits >> 32)});
                }
            }
            hn.w_init[md.wrow] = (uint32_t)pm[i].preload.size();  // head 0, len n / table entries
        }
        if (md.type == MT_STOPWATCH && pm[i].has_best) {
            hn.w_init[md.wrow + 1] = hn.content_code(pm[i].best_name, &overflow);
            hn.w_init[md.wrow + 3] = pm[i].best_seq;
            hn.d_init[md.drow] = pm[i].best_dur;
        }
        if (net.models[i].dist_kind == DK_EXP) net.uses_exp = 1;
        if (net.models[i].dist_kind == DK_NORMAL || net.models[i].dist_kind == DK_GAMMA ||
            net.models[i].dist_kind == DK_LOGNORMAL)
            net.uses_normal = 1;
    }
    for (size_t k = 0; k < hn.coupled.size(); k++) {  // parked messages given on input
        const SimbCoupledDesc& cd = net.coupled[k];
        uint32_t n = 0;
        for (auto& pmsg : hn.coupled[k].parked) {
            const int comp = hn.find_component((int)k, pmsg.component_id);
            if (comp < 0) return c.fail(1, hn.coupled[k].id + ": a parked message names the unknown component `" + pmsg.component_id + "`");
            hn.ring_preload.push_back(RingPreload{cd.ring_w0 + n, hn.content_code(pmsg.content, &overflow)});
            hn.ring_preload.push_back(RingPreload{cd.ring_w1 + n, (uint32_t)comp | (hn.resolve_in_port(comp, pmsg.port) << 8)});
            n++;
        }
        hn.w_init[cd.wrow] = n;
    }
    for (uint32_t k = 0; k < net.n_phase_words; k++) hn.w_init[net.wrow_phase + k] = phase_words[k];
    if (overflow) return c.fail(32, "content intern table overflow (more than 255 distinct names)");
    return 0;
}

}  // namespace simb
