// tests/hostsim/hostsim.cpp -- UNIT-TEST HARNESS, NOT PRODUCT CODE.
// Compiles the device interpreter (sim_b200/csrc/core.cuh) as plain C++ and runs it one lane at a
// time over the same SoA layout the CUDA kernel uses, so the model logic and the descriptor
// lowering can be checked against the oracle in the CPU-only test tier.  libsimb200.so never
// links this file; the product path has no CPU fallback.
#include <cstring>
#include <string>
#include <vector>

#include "../../sim_b200/csrc/core.cuh"
#include "../../sim_b200/csrc/lower.hpp"

#define SIMB_TABLE_QUAL const
#include "../../sim_b200/csrc/zig_tables.inc"
#include "../../sim_b200/csrc/shapes.inc"

using namespace simb;

namespace {
struct HS {
    HostNet hn;
    SimbDevState st;
    uint64_t n = 0;
    int rng_kind = 0;
    bool instr = false;
    std::vector<double> d, ring_d;
    std::vector<uint32_t> w, ring_w, trace_count, mail;
    std::vector<SimbTraceRec> trace;
    std::vector<unsigned long long> conn_count, record_count, totals;
    std::vector<uint64_t> ext;
    std::vector<uint32_t> tab;
    std::vector<uint64_t> vtab;
    uint32_t epoch = 0;
    uint64_t settle_every = 0;
    uint32_t rng_slots = SIMB_RNG_CACHE;  // Philox draws cached by the baked-shape code path (2 = streaming launches)
    int shape = -1;
};

template <int RNG, bool INSTR, int SM>
void run(HS* h, int mode, uint64_t k, double until) {
    uint64_t rng_cache[SIMB_RNG_CACHE];  // one lane at a time: its column is the whole cache (lane stride applies)
    Tables tab{ZIG_EXP_X, ZIG_EXP_F, ZIG_NORM_X, ZIG_NORM_F, h->tab.data(), h->vtab.data(), rng_cache, h->rng_slots};
    // SM > 0: the literal-index ("baked shape") code path of the interpreter, structure from the host
    // copy of the baked descriptor, values from the lowered network -- as the specialised kernels do
    Engine<RNG, SM> eng(SM > 0 ? *kShapeHost[h->shape] : h->hn.desc, h->hn.desc, h->st, tab);
    for (uint64_t r = 0; r < h->n; r++) {
        Lane ln;
        ln.d = h->d.data() + r;
        ln.w = h->w.data() + r;
        ln.stride = (uint32_t)h->n;
        ln.rep = (uint32_t)r;
        ln.rc = rng_cache;
        ln.rc_stride = 1;
        eng.lane_load(ln, INSTR);
        bool done = false;
        for (uint64_t i = 0; mode == 1 || i < k; i++) {
            bool live = !ln.err && !done;
            if (!live) break;
            eng.template step<INSTR>(ln, true);
            if (mode == 1 && !(ln.time < until)) done = true;
            // h->settle_every > 0: store + reload the lane every few steps, as consecutive kernel
            // launches do (exercises the persisted deferred-draw mask and the Philox prefetch restart)
            if (h->settle_every && (i + 1) % h->settle_every == 0) {
                eng.lane_store(ln, INSTR, 0, 0);
                eng.lane_load(ln, INSTR);
            }
        }
        eng.flush_draws(ln);  // "settle": what the last launch of an API call does
        eng.lane_store(ln, INSTR, done ? h->epoch : 0, 0);
    }
}
void dispatch(HS* h, int mode, uint64_t k, double until) {
    h->epoch = (h->epoch % 255) + 1;
#define RUN(R)                                            \
    if (h->instr) run<R, true, 0>(h, mode, k, until);     \
    else run<R, false, 0>(h, mode, k, until);
    if (h->shape >= 0 && h->rng_kind == 0 && !h->instr) {
        static const int tiers[] = SIMB_SHAPE_TIERS;
        switch (tiers[h->shape]) {
            case 4: run<0, false, 4>(h, mode, k, until); return;
            case 8: run<0, false, 8>(h, mode, k, until); return;
            case 16: run<0, false, 16>(h, mode, k, until); return;
            default: run<0, false, 32>(h, mode, k, until); return;
        }
    }
    if (h->rng_kind == 0) { RUN(0) }
    else if (h->rng_kind == 1) { RUN(1) }
    else { RUN(2) }
#undef RUN
}
}  // namespace

extern "C" {
void* hs_new(const char* models, const char* connectors, uint64_t n, uint64_t rep_offset, uint64_t seed, int rng_kind,
             int instrument, const char* stat_model_id, uint32_t queue_capacity, uint32_t max_pending,
             uint32_t trace_reps, uint32_t trace_cap, uint32_t batch_warmup, uint32_t batch_size, char* errbuf, int errlen,
             int* status) {
    HS* h = new HS();
    LowerOptions opt;
    opt.queue_capacity = queue_capacity;
    opt.max_pending = max_pending;
    opt.instrument = instrument != 0;
    opt.rng_kind = rng_kind;
    if (stat_model_id) opt.stat_model_id = stat_model_id;
    opt.batch_warmup = batch_warmup;
    opt.batch_size = batch_size;
    std::string err;
    int e = lower_network(models, connectors, opt, &h->hn, &err);
    if (status) *status = e;
    if (e) {
        if (errbuf && errlen > 0) {
            std::strncpy(errbuf, err.c_str(), (size_t)errlen - 1);
            errbuf[errlen - 1] = 0;
        }
        delete h;
        return nullptr;
    }
    const SimbNetDesc& net = h->hn.desc;
    h->n = n;
    h->rng_kind = rng_kind;
    h->instr = instrument != 0;
    h->d.assign((size_t)net.n_drows * n, 0.0);
    h->w.assign((size_t)net.n_wrows * n, 0u);
    h->ring_w.assign((size_t)(h->hn.ring_w_slots + 1) * n, 0u);
    h->ring_d.assign((size_t)(h->hn.ring_d_slots + 1) * n, 0.0);
    h->mail.assign((size_t)(2u * (net.max_pending - net.pend_tile) + 1) * n, 0u);
    if (!trace_cap) trace_cap = 4096;
    if (trace_reps > n) trace_reps = (uint32_t)n;
    h->trace.assign((size_t)trace_reps * trace_cap + 1, SimbTraceRec());
    h->trace_count.assign(n + 1, 0u);
    h->conn_count.assign(net.n_connectors + 1, 0ull);
    h->record_count.assign((size_t)net.n_models * ACT_COUNT, 0ull);
    h->totals.assign(4, 0ull);
    for (uint64_t r = 0; r < n; r++) {
        for (uint32_t k = 0; k < net.n_drows; k++) h->d[(size_t)k * n + r] = h->hn.d_init[k];
        for (uint32_t k = 0; k < net.n_wrows; k++) h->w[(size_t)k * n + r] = h->hn.w_init[k];
        if (rng_kind == 1) {  // Pcg64Mcg::new(seed + global replication): state = seed | 1
            unsigned __int128 s = (unsigned __int128)(seed + rep_offset + r) | 1;
            h->w[(size_t)(net.wrow_rng + 0) * n + r] = (uint32_t)s;
            h->w[(size_t)(net.wrow_rng + 1) * n + r] = (uint32_t)(s >> 32);
            h->w[(size_t)(net.wrow_rng + 2) * n + r] = (uint32_t)(s >> 64);
            h->w[(size_t)(net.wrow_rng + 3) * n + r] = (uint32_t)(s >> 96);
        }
        if (instrument) {
            h->w[(size_t)net.wrow_hash * n + r] = (uint32_t)TRACE_HASH_INIT;
            h->w[(size_t)(net.wrow_hash + 1) * n + r] = (uint32_t)(TRACE_HASH_INIT >> 32);
        }
        for (auto& pl : h->hn.ring_preload) {
            if (pl.slot & RING_PRELOAD_F64)
                reinterpret_cast<uint32_t*>(h->ring_d.data())[2 * ((size_t)(pl.slot & 0x3FFFFFFFu) * n + r) + ((pl.slot >> 30) & 1u)] = pl.value;
            else
                h->ring_w[(size_t)pl.slot * n + r] = pl.value;
        }
    }
    SimbDevState& st = h->st;
    std::memset(&st, 0, sizeof st);
    st.d = h->d.data();
    st.w = h->w.data();
    st.ring_w = h->ring_w.data();
    st.ring_d = h->ring_d.data();
    st.mail = h->mail.data();
    st.stride = n;
    st.n = n;
    st.rep_offset = rep_offset;
    st.seed = seed;
    st.trace_reps = trace_reps;
    st.trace_cap = trace_cap;
    st.trace = h->trace.data();
    st.trace_count = h->trace_count.data();
    st.conn_count = h->conn_count.data();
    st.record_count = h->record_count.data();
    st.totals = h->totals.data();
    h->tab.assign(SIMB_TAB_WORDS, 0u);
    h->vtab.assign(SIMB_VTAB_WORDS, 0ull);
    build_model_table(net, h->tab.data(), h->vtab.data());
    st.tab = h->tab.data();
    st.vtab = h->vtab.data();
    return h;
}
void hs_free(void* h) { delete (HS*)h; }
void hs_set_reload_every(void* h, uint64_t n) { ((HS*)h)->settle_every = n; }
void hs_set_rng_slots(void* h, uint32_t n) { ((HS*)h)->rng_slots = n; }
// select the baked-shape code path when the lowered network matches shape `k` structurally; returns 1 on match
int hs_use_shape(void* hv, int k) {
    HS* h = (HS*)hv;
    if (k < 0 || k >= SIMB_N_SHAPES) return 0;
    auto strip = [](SimbNetDesc* d) {
        for (int i = 0; i < SIMB_MAX_MODELS; i++) {
            SimbModelDesc& m = d->models[i];
            m.p0 = m.p1 = m.p2 = 0.0;
            m.q0 = m.q1 = m.q2 = 0;
            m.dist_err = 0;
        }
        std::memset(d->wcum, 0, sizeof d->wcum);
    };
    SimbNetDesc a, b;
    std::memcpy(&a, &h->hn.desc, sizeof a);
    std::memcpy(&b, kShapeHost[k], sizeof b);
    strip(&a);
    strip(&b);
    if (std::memcmp(&a, &b, sizeof a) != 0) return 0;
    h->shape = k;
    return 1;
}
void hs_set_stream(void* hv, const uint64_t* draws, uint64_t per_rep) {
    HS* h = (HS*)hv;
    h->ext.assign(draws, draws + per_rep * h->n);
    h->st.ext = h->ext.data();
    h->st.ext_len = (uint32_t)per_rep;
}
void hs_step_n(void* hv, uint64_t n) { dispatch((HS*)hv, 0, n, 0.0); }
void hs_step_until(void* hv, double until) { dispatch((HS*)hv, 1, 0, until); }
int hs_inject(void* hv, const char* tgt, const char* port, const char* content) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    bool ov = false;
    uint32_t code = h->hn.content_code(content, &ov);
    std::vector<uint32_t> words;  // as simb_inject_input (api.cu): every model of that id, the components of a Coupled one
    for (const HostNet::Delivery& d : h->hn.deliveries(tgt, port))
        words.push_back(ROUTE_WORD(d.model, d.port, words.empty() ? ROUTE_INJECTED : ROUTE_INJECTED_CONT, 0));
    if (words.empty()) words.push_back(ROUTE_WORD(0xFF, SIMB_PORT_UNKNOWN, ROUTE_INJECTED, 0));
    for (uint32_t rw0 : words)
        for (uint64_t r = 0; r < h->n; r++) {
            uint32_t& ctl = h->w[(size_t)net.wrow_ctl * h->n + r];
            uint32_t np = CTL_NPEND(ctl);
            if (np >= net.max_pending) return 32;
            const uint32_t rw = rw0 | (np << 24);
            if (np < net.pend_tile) {
                h->w[(size_t)(net.wrow_pend + 2 * np) * h->n + r] = code;
                h->w[(size_t)(net.wrow_pend + 2 * np + 1) * h->n + r] = rw;
            } else {
                h->mail[(size_t)(2 * (np - net.pend_tile)) * h->n + r] = code;
                h->mail[(size_t)(2 * (np - net.pend_tile) + 1) * h->n + r] = rw;
            }
            ctl = CTL_MAKE(CTL_ERR(ctl), CTL_EPOCH(ctl), np + 1, 0);
        }
    return 0;
}
void hs_get_time(void* hv, double* out) {
    HS* h = (HS*)hv;
    std::memcpy(out, h->d.data() + (size_t)h->hn.desc.drow_time * h->n, h->n * 8);
}
void hs_get_until(void* hv, int m, double* out) {
    HS* h = (HS*)hv;
    std::memcpy(out, h->d.data() + (size_t)(h->hn.desc.drow_until + m) * h->n, h->n * 8);
}
void hs_get_row(void* hv, int which, uint32_t* out) {  // 0 steps, 1 events, 2 draws, 3 err, 4 npend
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    for (uint64_t r = 0; r < h->n; r++) {
        uint32_t ctl = h->w[(size_t)net.wrow_ctl * h->n + r];
        switch (which) {
            case 0: out[r] = h->w[(size_t)net.wrow_steps * h->n + r]; break;
            case 1: out[r] = h->w[(size_t)net.wrow_events * h->n + r]; break;
            case 2: out[r] = h->w[(size_t)net.wrow_rng * h->n + r]; break;
            case 3: out[r] = CTL_ERR(ctl); break;
            default: out[r] = CTL_NPEND(ctl); break;
        }
    }
}
void hs_get_hash(void* hv, uint64_t* out) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    for (uint64_t r = 0; r < h->n; r++)
        out[r] = (uint64_t)h->w[(size_t)net.wrow_hash * h->n + r] | ((uint64_t)h->w[(size_t)(net.wrow_hash + 1) * h->n + r] << 32);
}
void hs_get_wait(void* hv, double* sum, uint32_t* n) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    for (uint64_t r = 0; r < h->n; r++) {
        sum[r] = h->d[(size_t)net.drow_wait * h->n + r];
        n[r] = h->w[(size_t)net.wrow_waitn * h->n + r];
    }
}
void hs_get_batch(void* hv, double* sum_means, double* sum_sq, double* open_sum) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    for (uint64_t r = 0; r < h->n; r++) {
        open_sum[r] = h->d[(size_t)net.drow_bm * h->n + r];
        sum_means[r] = h->d[(size_t)(net.drow_bm + 1) * h->n + r];
        sum_sq[r] = h->d[(size_t)(net.drow_bm + 2) * h->n + r];
    }
}
void hs_get_conn_counts(void* hv, uint64_t* out) {
    HS* h = (HS*)hv;
    for (uint32_t i = 0; i < h->hn.desc.n_connectors; i++) out[i] = h->conn_count[i];
}
void hs_get_record_counts(void* hv, uint64_t* out) {
    HS* h = (HS*)hv;
    for (size_t i = 0; i < h->record_count.size(); i++) out[i] = h->record_count[i];
}
uint32_t hs_trace_count(void* hv, uint64_t rep) { return ((HS*)hv)->trace_count[rep]; }
// out: step, kind, index, action, aux, content
void hs_trace_get(void* hv, uint64_t rep, uint32_t i, uint32_t out[6], double* time) {
    HS* h = (HS*)hv;
    const SimbTraceRec& r = h->trace[(size_t)rep * h->st.trace_cap + i];
    out[0] = r.step;
    out[1] = r.meta >> 31;
    out[2] = r.meta & 0xFFFFu;
    out[3] = r.aux & 0xFFu;
    out[4] = (r.aux >> 8) & 0xFFu;
    out[5] = r.content;
    *time = r.time;
}
void hs_render(void* hv, uint32_t code, char* buf, int len) {
    std::string s = ((HS*)hv)->hn.render(code);
    std::strncpy(buf, s.c_str(), (size_t)len - 1);
    buf[len - 1] = 0;
}
void hs_layout(void* hv, uint32_t out[8]) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    out[0] = net.n_drows;
    out[1] = net.n_wrows;
    out[2] = net.max_pending;
    out[3] = h->hn.ring_w_slots;
    out[4] = h->hn.ring_d_slots;
    out[5] = net.n_routes;
    out[6] = net.n_models;
    out[7] = net.n_connectors;
}
void hs_pending_get(void* hv, uint64_t rep, uint32_t i, uint32_t out[2]) {
    HS* h = (HS*)hv;
    const SimbNetDesc& net = h->hn.desc;
    if (i < net.pend_tile) {
        out[0] = h->w[(size_t)(net.wrow_pend + 2 * i) * h->n + rep];
        out[1] = h->w[(size_t)(net.wrow_pend + 2 * i + 1) * h->n + rep];
    } else {
        out[0] = h->mail[(size_t)(2 * (i - net.pend_tile)) * h->n + rep];
        out[1] = h->mail[(size_t)(2 * (i - net.pend_tile) + 1) * h->n + rep];
    }
}
}
