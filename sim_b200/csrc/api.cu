// sim_b200/csrc/api.cu -- the C ABI of libsimb200.so (include/simb.h): handle management, the
// host drivers of Simulation::step / step_n / step_until over the CUDA engine, observation and
// statistics read-back.  No CPU fallback: every entry that needs the device fails with
// SIMB_ERR_CUDA when CUDA is unavailable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/simb.h"
#include "actions.h"
#include "engine.cuh"
#include "json.hpp"
#include "json_write.hpp"
#include "lower.hpp"
#include "output_analysis.hpp"
#include "yaml.hpp"

using namespace simb;

static thread_local std::string g_last_error;

struct simb_sim {
    HostNet hn;
    LowerOptions opt;
    simb_config cfg;
    std::string stat_model_id;
    SimbDevState st;
    EngineDims dims;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // a pass over the batch is issued as `n_lanes` concurrent grids (contiguous tile ranges) on their own
    // streams, so that one grid's drain overlaps the others' work; lane 0 is `stream`
    static const int MAX_LANES = 4;
    int n_lanes = 1;
    cudaStream_t side[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    // device allocations
    double* d_dinit = nullptr;
    uint32_t* d_winit = nullptr;
    uint32_t* d_pre_slots = nullptr;
    uint32_t* d_pre_vals = nullptr;
    uint32_t n_pre = 0;
    uint64_t* d_ext = nullptr;
    double* d_psum = nullptr;
    unsigned long long* d_pcnt = nullptr;
    unsigned long long* d_overflow = nullptr;
    uint32_t n_red_blocks = 0;
    uint32_t K = 1;
    uint32_t epoch = 0;
    uint64_t launches = 0;
    double kernel_ms = 0.0;
    bool instrument = false;
    int rng_kind = 0;
    uint64_t passes = 0;        // passes issued (parity = sweep direction)
    bool l2_keep = true;        // SIMB_NO_L2_KEEP = A/B switch
    int sm_count = 148;         // cudaDevAttrMultiProcessorCount of the handle's device
    bool sweep_reverse = true;  // alternate the sweep direction of consecutive passes (SIMB_NO_REVERSE = A/B switch)
    std::string json;  // result of the last *_json call (handle-owned, see simb.h)
    json::Value models_doc, connectors_doc;  // the documents of the last post / put (simb_get_json re-emits them)
};

namespace {

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (cudaError_t)(expr);                                                              \
        if (_e != cudaSuccess)                                                                             \
            return fail(SIMB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

// Device memory comes from the stream-ordered pool of the device (cudaMallocAsync) with the release
// threshold raised once, so that re-posting a simulation (the replication idiom: post, run, read, destroy)
// reuses the pool instead of going back to the driver for hundreds of MB every time.
void ensure_pool(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = true;
}

void destroy_streams(simb_sim* s) {
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->fork) cudaEventDestroy(s->fork);
    for (int i = 1; i < simb_sim::MAX_LANES; i++) {
        if (s->join[i]) cudaEventDestroy(s->join[i]);
        if (s->side[i]) cudaStreamDestroy(s->side[i]);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
}

void free_device(simb_sim* s) {
    cudaStream_t stream = s->stream;
    auto F = [stream](void* p) {
        if (p) cudaFreeAsync(p, stream);
    };
    F(s->st.d);
    F(s->st.w);
    F(s->st.ring_w);
    F(s->st.ring_d);
    F(s->st.mail);
    F(s->st.trace);
    F(s->st.trace_count);
    F(s->st.conn_count);
    F(s->st.record_count);
    F(s->st.totals);
    F(s->st.series);
    F((void*)s->st.tab);
    F((void*)s->st.vtab);
    F(s->d_dinit);
    F(s->d_winit);
    F(s->d_pre_slots);
    F(s->d_pre_vals);
    F(s->d_ext);
    F(s->d_psum);
    F(s->d_pcnt);
    F(s->d_overflow);
    std::memset(&s->st, 0, sizeof s->st);
    s->d_dinit = nullptr;
    s->d_winit = nullptr;
    s->d_pre_slots = s->d_pre_vals = nullptr;
    s->d_ext = nullptr;
    s->d_psum = nullptr;
    s->d_pcnt = nullptr;
    s->d_overflow = nullptr;
}

template <class T>
int dev_alloc(cudaStream_t stream, T** p, size_t count, bool zero) {
    size_t bytes = (count ? count : 1) * sizeof(T);
    CUDA_TRY(cudaMallocAsync((void**)p, bytes, stream));
    if (zero) CUDA_TRY(cudaMemsetAsync(*p, 0, bytes, stream));
    return 0;
}

// model / route / value tables of the generic interpreter (actions.h) for s->hn
int upload_tables(simb_sim* s) {
    std::vector<uint32_t> tab(SIMB_TAB_WORDS);
    std::vector<uint64_t> vtab(SIMB_VTAB_WORDS);
    build_model_table(s->hn.desc, tab.data(), vtab.data());
    int r;
    if (!s->st.tab) {
        uint32_t* d_tab = nullptr;
        uint64_t* d_vtab = nullptr;
        if ((r = dev_alloc(s->stream, &d_tab, tab.size(), false))) return r;
        if ((r = dev_alloc(s->stream, &d_vtab, vtab.size(), false))) return r;
        s->st.tab = d_tab;
        s->st.vtab = d_vtab;
    }
    CUDA_TRY(cudaMemcpyAsync((void*)s->st.tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync((void*)s->st.vtab, vtab.data(), vtab.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // the staging vectors go out of scope
    return 0;
}

// allocate + initialise device state for s->hn (already lowered)
int build_device(simb_sim* s, uint32_t keep_mask, const simb_sim* old) {
    (void)old;
    const SimbNetDesc& net = s->hn.desc;
    const uint64_t n = s->cfg.replications;
    int e = engine_configure(net, n, s->rng_kind, s->instrument, &s->dims);
    if (e) return fail(SIMB_ERR_CUDA, std::string("engine_configure: ") + cudaGetErrorString((cudaError_t)e));
    const uint64_t stride = (n + 255) / 256 * 256;
    SimbDevState& st = s->st;
    std::memset(&st, 0, sizeof st);
    st.stride = stride;
    st.n = n;
    st.rep_offset = s->cfg.replication_offset;
    st.seed = s->cfg.seed;
    int r;
    if ((r = dev_alloc(s->stream, &st.d, (size_t)net.n_drows * stride, false))) return r;
    if ((r = dev_alloc(s->stream, &st.w, (size_t)net.n_wrows * stride, false))) return r;
    if ((r = dev_alloc(s->stream, &st.ring_w, (size_t)(s->hn.ring_w_slots + 1) * stride, true))) return r;
    if ((r = dev_alloc(s->stream, &st.ring_d, (size_t)(s->hn.ring_d_slots + 1) * stride, true))) return r;
    if ((r = dev_alloc(s->stream, &st.mail, (size_t)(2u * (net.max_pending - net.pend_tile) + 1) * stride, true))) return r;
    st.trace_reps = 0;
    st.trace_cap = 0;
    if (s->instrument) {
        st.trace_reps = (uint32_t)std::min<uint64_t>(s->cfg.trace_replications, n);
        st.trace_cap = s->cfg.trace_capacity ? s->cfg.trace_capacity : 4096;
        if ((r = dev_alloc(s->stream, &st.trace, (size_t)st.trace_reps * st.trace_cap, true))) return r;
        if ((r = dev_alloc(s->stream, &st.trace_count, (size_t)stride, true))) return r;
        if ((r = dev_alloc(s->stream, &st.conn_count, (size_t)net.n_connectors + 1, true))) return r;
        if ((r = dev_alloc(s->stream, &st.record_count, (size_t)net.n_models * ACT_COUNT, true))) return r;
    }
    if (s->cfg.series_capacity && s->hn.stat_model >= 0) {
        st.series_cap = s->cfg.series_capacity;
        if ((r = dev_alloc(s->stream, &st.series, (size_t)st.series_cap * stride, false))) return r;
    }
    if ((r = dev_alloc(s->stream, &st.totals, 5, true))) return r;
    if ((r = upload_tables(s))) return r;
    if ((r = dev_alloc(s->stream, &s->d_dinit, net.n_drows, false))) return r;
    if ((r = dev_alloc(s->stream, &s->d_winit, net.n_wrows, false))) return r;
    CUDA_TRY(cudaMemcpy(s->d_dinit, s->hn.d_init.data(), net.n_drows * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(s->d_winit, s->hn.w_init.data(), net.n_wrows * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s->n_pre = (uint32_t)s->hn.ring_preload.size();
    if (s->n_pre) {
        std::vector<uint32_t> slots(s->n_pre), vals(s->n_pre);
        for (uint32_t i = 0; i < s->n_pre; i++) {
            slots[i] = s->hn.ring_preload[i].slot;
            vals[i] = s->hn.ring_preload[i].value;
        }
        if ((r = dev_alloc(s->stream, &s->d_pre_slots, s->n_pre, false))) return r;
        if ((r = dev_alloc(s->stream, &s->d_pre_vals, s->n_pre, false))) return r;
        CUDA_TRY(cudaMemcpy(s->d_pre_slots, slots.data(), s->n_pre * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(s->d_pre_vals, vals.data(), s->n_pre * 4, cudaMemcpyHostToDevice));
    }
    s->n_red_blocks = engine_reduce_blocks(n);
    if ((r = dev_alloc(s->stream, &s->d_psum, s->n_red_blocks, true))) return r;
    if ((r = dev_alloc(s->stream, &s->d_pcnt, s->n_red_blocks, true))) return r;
    if ((r = dev_alloc(s->stream, &s->d_overflow, 1, true))) return r;
    CUDA_TRY(engine_launch_init(net, st, s->d_dinit, s->d_winit, s->d_pre_slots, s->d_pre_vals, s->n_pre, s->rng_kind,
                                s->instrument, keep_mask, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int first_error(simb_sim* s, uint64_t failed_hint) {
    if (!failed_hint) return 0;
    std::vector<uint32_t> ctl(s->st.n);
    CUDA_TRY(cudaMemcpy(ctl.data(), s->st.w + (uint64_t)s->hn.desc.wrow_ctl * s->st.stride, s->st.n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < s->st.n; r++)
        if (CTL_ERR(ctl[r]))
            return fail((int)CTL_ERR(ctl[r]), "replication " + std::to_string(r) + " failed with " + status_name((int)CTL_ERR(ctl[r])));
    return 0;
}

// drive the step kernel.  mode 0: exactly n steps; mode 1: until every replication's clock >= until.
// small batches (fewer than two full waves of CTAs per grid) are not split
int grids_per_pass(const simb_sim* s) {
    const uint32_t tiles = (uint32_t)(s->st.stride / s->dims.tile);
    // a grid of fewer than two waves (4 resident CTAs per SM) is not worth splitting
    return tiles >= 2u * 4u * (uint32_t)s->sm_count * (uint32_t)s->n_lanes ? s->n_lanes : 1;
}
// One pass over the batch: the step kernel on every tile, issued as s->n_lanes concurrent grids.  The side
// streams must have been forked from s->stream (pass_fork) and are joined back by pass_join.
int launch_pass(simb_sim* s, const SimbLaunch& la_in, bool count_active, uint64_t* issued) {
    SimbLaunch la = la_in;
    la.reverse = s->sweep_reverse && (s->passes++ & 1u) ? 1u : 0u;
    la.l2_keep = s->l2_keep && la.k <= 8u ? 1u : 0u;  // only passes that re-stream the state soon
    {
        static const bool lockstep = getenv("SIMB_LOCKSTEP") && *getenv("SIMB_LOCKSTEP");  // A/B switch
        la.lockstep = lockstep ? 1u : 0u;
    }
    const SimbNetDesc& net = s->hn.desc;
    const uint32_t tiles = (uint32_t)(s->st.stride / s->dims.tile);
    const int lanes = grids_per_pass(s);
    const uint32_t per = (tiles + lanes - 1) / lanes;
    for (int i = 0; i < lanes; i++) {
        const uint32_t t0 = (uint32_t)i * per;
        if (t0 >= tiles) break;
        const uint32_t nt = std::min(per, tiles - t0);
        CUDA_TRY(engine_launch_step(net, s->st, la, s->dims, s->rng_kind, s->instrument, count_active,
                                    i == 0 ? s->stream : s->side[i], t0, nt));
        (*issued)++;
    }
    return 0;
}
int pass_fork(simb_sim* s) {
    if (s->n_lanes <= 1) return 0;
    CUDA_TRY(cudaEventRecord(s->fork, s->stream));
    for (int i = 1; i < s->n_lanes; i++) CUDA_TRY(cudaStreamWaitEvent(s->side[i], s->fork, 0));
    return 0;
}
int pass_join(simb_sim* s) {
    for (int i = 1; i < s->n_lanes; i++) {
        CUDA_TRY(cudaEventRecord(s->join[i], s->side[i]));
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->join[i], 0));
    }
    return 0;
}

int run_steps(simb_sim* s, int mode, uint64_t n, double until) {
    const SimbNetDesc& net = s->hn.desc;
    (void)net;
    SimbLaunch la;
    std::memset(&la, 0, sizeof la);
    la.mode = (uint32_t)mode;
    la.until = until;
    if (mode == 1) {
        s->epoch = (s->epoch % 255u) + 1u;
        la.epoch = s->epoch;
    }
    unsigned long long tot[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    uint64_t issued = 0;
    int pe = 0;
    if (mode == 0) {
        uint64_t left = n;
        if (n > 0) {
            CUDA_TRY(cudaMemsetAsync(s->st.totals + 2, 0, 2 * sizeof(unsigned long long), s->stream));
            if ((pe = pass_fork(s))) return pe;
        }
        while (left > 0) {
            la.k = (uint32_t)std::min<uint64_t>(left, s->K);
            left -= la.k;
            const bool last = left == 0;
            la.settle = last ? 1u : 0u;  // the state the caller can observe owes no deferred draw
            if ((pe = launch_pass(s, la, last, &issued))) return pe;
        }
        if (n > 0) {
            if ((pe = pass_join(s))) return pe;
            CUDA_TRY(cudaMemcpyAsync(tot, s->st.totals, sizeof tot, cudaMemcpyDeviceToHost, s->stream));
        }
        CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    } else {
        la.k = s->K;
        // launches between termination checks: about 256 steps' worth
        const uint32_t round = std::max<uint32_t>(1u, 256u / std::max<uint32_t>(1u, s->K));
        for (;;) {
            CUDA_TRY(cudaMemsetAsync(s->st.totals + 2, 0, 2 * sizeof(unsigned long long), s->stream));
            if ((pe = pass_fork(s))) return pe;
            for (uint32_t i = 0; i < round; i++) {
                const bool last = i + 1 == round;
                la.settle = last ? 1u : 0u;
                if ((pe = launch_pass(s, la, last, &issued))) return pe;
            }
            if ((pe = pass_join(s))) return pe;
            CUDA_TRY(cudaMemcpyAsync(tot, s->st.totals, sizeof tot, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (tot[2] == 0) break;  // no replication still below `until`
        }
    }
    float ms = 0.f;
    if (issued) {
        CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->kernel_ms += ms;
        s->launches += issued;
    }
    return first_error(s, tot[3]);
}

int column_w(simb_sim* s, uint64_t rep, std::vector<uint32_t>* out) {
    const SimbNetDesc& net = s->hn.desc;
    out->assign(net.n_wrows, 0);
    CUDA_TRY(cudaMemcpy2D(out->data(), 4, s->st.w + rep, s->st.stride * 4, 4, net.n_wrows, cudaMemcpyDeviceToHost));
    return 0;
}
int column_d(simb_sim* s, uint64_t rep, std::vector<double>* out) {
    const SimbNetDesc& net = s->hn.desc;
    out->assign(net.n_drows, 0);
    CUDA_TRY(cudaMemcpy2D(out->data(), 8, s->st.d + rep, s->st.stride * 8, 8, net.n_drows, cudaMemcpyDeviceToHost));
    return 0;
}
void put_str(char* dst, size_t cap, const std::string& v) {
    if (!dst || !cap) return;
    size_t n = std::min(cap - 1, v.size());
    std::memcpy(dst, v.data(), n);
    dst[n] = 0;
}
int copy_row_w(simb_sim* s, uint32_t row, uint32_t* out, uint64_t capacity) {
    if (!out || capacity < s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the replication count");
    CUDA_TRY(cudaMemcpy(out, s->st.w + (uint64_t)row * s->st.stride, s->st.n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int copy_row_d(simb_sim* s, uint32_t row, double* out, uint64_t capacity) {
    if (!out || capacity < s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the replication count");
    CUDA_TRY(cudaMemcpy(out, s->st.d + (uint64_t)row * s->st.stride, s->st.n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// subject string of a record, as the reference formats it per model (file:line in core.cuh)
std::string record_subject(const simb_sim* s, uint32_t model, uint32_t action, uint32_t aux, uint32_t content) {
    const HostModel& hm = s->hn.models[model];
    const uint8_t type = s->hn.desc.models[model].type;
    std::string job = s->hn.render(content);
    if (content == SIMB_CONTENT_NONE) {
        if (action == ACT_INITIALIZATION) return "";
        return "None";  // Storage release with no job / Stopwatch fetch with no completed job
    }
    if (aux) {
        const std::vector<std::string>& ports =
            (action == ACT_ARRIVAL && (type == MT_EXCLUSIVE_GATEWAY || type == MT_PARALLEL_GATEWAY)) ? hm.ports_in : hm.ports_out;
        std::string p = aux - 1 < ports.size() ? ports[aux - 1] : "?";
        return job + " on " + p;
    }
    if (action == ACT_ARRIVAL && type == MT_EXCLUSIVE_GATEWAY) return job + " on ?";
    return job;
}

}  // namespace

extern "C" {

const char* simb_last_error_message(void) { return g_last_error.c_str(); }
const char* simb_status_name(int status) { return status_name(status); }
int simb_abi_version(void) { return SIMB_ABI_VERSION; }
int simb_device_count(int* count) {
    if (!count) return fail(SIMB_ERR_INVALID_ARGUMENT, "null count");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return 0;
}

int simb_post_json(const char* models_json, const char* connectors_json, const simb_config* config, simb_sim** out) {
    if (!out || !config || !models_json || !connectors_json) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    // ABI evolution: earlier callers pass the struct without the trailing batch-means / series fields
    const size_t v1_size = offsetof(simb_config, batch_warmup), v2_size = offsetof(simb_config, series_capacity);
    if (config->struct_size != sizeof(simb_config) && config->struct_size != v1_size && config->struct_size != v2_size)
        return fail(SIMB_ERR_INVALID_ARGUMENT, "simb_config.struct_size mismatch");
    if (config->replications == 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "replications must be > 0");
    if (config->rng_kind < 0 || config->rng_kind > 2) return fail(SIMB_ERR_INVALID_ARGUMENT, "unknown rng_kind");
    simb_sim* s = new simb_sim();
    std::memset(&s->cfg, 0, sizeof s->cfg);
    std::memcpy(&s->cfg, config, config->struct_size);
    if (config->stat_model_id) s->stat_model_id = config->stat_model_id;
    s->cfg.stat_model_id = nullptr;
    s->instrument = (config->flags & SIMB_FLAG_INSTRUMENT) != 0;
    s->rng_kind = config->rng_kind;
    s->K = config->steps_per_launch ? config->steps_per_launch : 128;  // measured: 64..256 is flat, 1 is HBM bound
    s->opt.queue_capacity = config->queue_capacity;
    s->opt.max_pending = config->max_pending;
    s->opt.instrument = s->instrument;
    s->opt.rng_kind = s->rng_kind;
    s->opt.stat_model_id = s->stat_model_id;
    s->opt.batch_warmup = s->cfg.batch_warmup;
    s->opt.batch_size = s->cfg.batch_size;
    std::string err;
    int e = lower_network(models_json, connectors_json, s->opt, &s->hn, &err);
    if (e) {
        delete s;
        return fail(e, err);
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        delete s;
        return fail(SIMB_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                                       "); libsimb200 has no CPU fallback");
    }
    std::memset(&s->st, 0, sizeof s->st);
    auto bail = [&](int code) {
        free_device(s);
        destroy_streams(s);
        delete s;
        return code;
    };
    if (cudaSetDevice(config->device) != cudaSuccess) return bail(fail(SIMB_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(SIMB_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess)
        return bail(fail(SIMB_ERR_CUDA, "cudaEventCreate failed"));
    {
        // measured (profiles/r01_pipeline_experiments.txt): two concurrent half-batch grids hide the grid
        // drain / ramp between dependent launches (+12 % at one step per launch); SIMB_STREAMS is the A/B switch
        const char* nk = getenv("SIMB_NO_L2_KEEP");
        s->l2_keep = !(nk && *nk);
        const char* nr = getenv("SIMB_NO_REVERSE");
        s->sweep_reverse = !(nr && *nr);
        const char* ns = getenv("SIMB_STREAMS");
        int want = ns && *ns ? atoi(ns) : 2;
        s->n_lanes = want < 1 ? 1 : want > simb_sim::MAX_LANES ? simb_sim::MAX_LANES : want;
        bool ok = cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 1; i < s->n_lanes && ok; i++)
            ok = cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&s->join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) return bail(fail(SIMB_ERR_CUDA, "side stream creation failed"));
    }
    ensure_pool(config->device);
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, config->device) == cudaSuccess && sms > 0) s->sm_count = sms;
    }
    e = build_device(s, 0, nullptr);
    if (e) return bail(e);
    {
        json::Parser pm(models_json), pc(connectors_json);
        pm.parse(s->models_doc);
        pc.parse(s->connectors_doc);
    }
    *out = s;
    return 0;
}

// Pending messages across Simulation::put (mod.rs:82-85 keeps `messages`; they are strings, matched against the NEW
// models' ids and ports when delivered, mod.rs:204-216).  The device keeps (content code, route word) pairs resolved
// against the OLD network, so a put that changes ids, ports, connectors or generator prefixes re-resolves them by
// name: content codes through the interned names, targets through (target id, target port); a message whose target
// no longer exists stays in the mailbox and is never delivered, like the reference's.
namespace {
struct CarriedMail {
    std::vector<uint32_t> npend;                  // per replication
    std::vector<std::vector<uint32_t>> content;   // [slot][replication]
    std::vector<std::vector<uint32_t>> route;
    uint32_t max_np = 0;
};
int read_mail(simb_sim* s, CarriedMail* cm) {
    const SimbNetDesc& net = s->hn.desc;
    const uint64_t n = s->st.n, S = s->st.stride;
    std::vector<uint32_t> ctl(n);
    CUDA_TRY(cudaMemcpy(ctl.data(), s->st.w + (uint64_t)net.wrow_ctl * S, n * 4, cudaMemcpyDeviceToHost));
    cm->npend.resize(n);
    for (uint64_t r = 0; r < n; r++) {
        cm->npend[r] = CTL_NPEND(ctl[r]);
        cm->max_np = std::max(cm->max_np, cm->npend[r]);
    }
    cm->content.assign(cm->max_np, std::vector<uint32_t>(n));
    cm->route.assign(cm->max_np, std::vector<uint32_t>(n));
    for (uint32_t j = 0; j < cm->max_np; j++) {
        const uint32_t* c = j < net.pend_tile ? s->st.w + (uint64_t)(net.wrow_pend + 2u * j) * S
                                              : s->st.mail + (uint64_t)(2u * (j - net.pend_tile)) * S;
        CUDA_TRY(cudaMemcpy(cm->content[j].data(), c, n * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(cm->route[j].data(), c + S, n * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}
// Simulation::put keeps `messages` as they are -- strings (mod.rs:82-85) -- and the next step hands each to whatever
// models of the NEW network carry its target id (mod.rs:203-221).  So: every Message in flight (the later deliveries of
// one Message, flagged RT_CONT / ROUTE_INJECTED_CONT, are dropped first) goes back to its (target id, target port)
// names and is expanded into the deliveries those names have in the new network, in the order the messages were sent.
// The interned names of `old` must already be interned in `nw` (name_map).
void remap_mail(const HostNet& old, const HostNet& nw, const std::vector<uint32_t>& name_map, CarriedMail* cm) {
    const std::string lost = "\x01(no such port)";
    auto by_name = [&](const std::string& id, const std::string& port) {
        std::vector<uint32_t> w;
        for (const HostNet::Delivery& d : nw.deliveries(id, port))
            w.push_back(ROUTE_WORD(d.model, d.port, w.empty() ? ROUTE_INJECTED : ROUTE_INJECTED_CONT, 0));
        if (w.empty()) w.push_back(ROUTE_WORD(0xFF, SIMB_PORT_UNKNOWN, ROUTE_INJECTED, 0));
        return w;
    };
    // a message sent over a connector: the same connector of the new network carries it (its deliveries are consecutive
    // routes, the first plain); if there is none it renders as an injected message
    std::vector<std::vector<uint32_t>> of_route(old.desc.n_routes);
    for (uint32_t r = 0; r < old.desc.n_routes; r++) {
        if (old.desc.routes[r].connector & (RT_PARK | RT_CONT)) continue;
        const HostConnector& oc = old.connectors[RT_CONN(old.desc.routes[r].connector)];
        for (uint32_t q = 0; q < nw.desc.n_routes && of_route[r].empty(); q++) {
            if (nw.desc.routes[q].connector & (RT_PARK | RT_CONT)) continue;
            const HostConnector& nc = nw.connectors[RT_CONN(nw.desc.routes[q].connector)];
            if (nc.source_id != oc.source_id || nc.source_port != oc.source_port || nc.target_id != oc.target_id ||
                nc.target_port != oc.target_port)
                continue;
            for (uint32_t k = q; k < nw.desc.n_routes; k++) {
                if (k > q && !(nw.desc.routes[k].connector & RT_CONT)) break;
                of_route[r].push_back(ROUTE_WORD(nw.desc.routes[k].tgt_model, nw.desc.routes[k].tgt_port, k, 0));
            }
        }
        if (of_route[r].empty()) of_route[r] = by_name(oc.target_id, oc.target_port);
    }
    // an injected message: (model, port) indices back to the names inject_input was given
    std::map<uint32_t, std::vector<uint32_t>> of_injected;
    auto injected = [&](uint32_t tgt, uint32_t op) -> const std::vector<uint32_t>& {
        const uint32_t key = tgt << 8 | op;
        auto it = of_injected.find(key);
        if (it != of_injected.end()) return it->second;
        std::string id = lost, port = lost;
        if (tgt < old.models.size()) {
            const HostModel& om = old.models[tgt];
            if (om.coupled < 0) {
                id = om.id;
                if (op < om.ports_in.size()) port = om.ports_in[op];
            } else {
                id = old.coupled[om.coupled].id;
                for (auto& e : old.coupled[om.coupled].eic)
                    if (e.target_id == om.id && old.resolve_in_port((int)tgt, e.target_port) == op) {
                        port = e.source_port;
                        break;
                    }
            }
        }
        return of_injected[key] = by_name(id, port);
    };
    const size_t n = cm->npend.size();
    std::vector<std::vector<uint32_t>> content, route;
    std::vector<uint32_t> npend(n, 0), order;
    uint32_t max_np = 0;
    for (size_t r = 0; r < n; r++) {
        order.clear();
        for (uint32_t j = 0; j < cm->npend[r]; j++) order.push_back(j);
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t x, uint32_t y) { return ROUTE_SEQ(cm->route[x][r]) < ROUTE_SEQ(cm->route[y][r]); });
        uint32_t np = 0;
        for (uint32_t j : order) {
            const uint32_t rw = cm->route[j][r], rr = ROUTE_R(rw);
            if (rr == ROUTE_INJECTED_CONT) continue;
            if (rr != ROUTE_INJECTED && rr < old.desc.n_routes && (old.desc.routes[rr].connector & RT_CONT)) continue;
            uint32_t c = cm->content[j][r];
            if (c != SIMB_CONTENT_NONE) {
                const uint32_t pfx = c >> 24;
                c = ((pfx < name_map.size() ? name_map[pfx] : pfx) << 24) | (c & 0xFFFFFFu);
            }
            const std::vector<uint32_t>& words =
                (rr == ROUTE_INJECTED || rr >= old.desc.n_routes) ? injected(ROUTE_TGT(rw), ROUTE_PORT(rw)) : of_route[rr];
            for (uint32_t w : words) {
                if (np >= content.size()) {
                    content.emplace_back(n, 0u);
                    route.emplace_back(n, 0u);
                }
                content[np][r] = c;
                route[np][r] = w | ((np & 0xFFu) << 24);
                np++;
            }
        }
        npend[r] = np;
        max_np = std::max(max_np, np);
    }
    cm->npend.swap(npend);
    cm->content.swap(content);
    cm->route.swap(route);
    cm->max_np = max_np;
}
int write_mail(simb_sim* s, const CarriedMail& cm) {
    const SimbNetDesc& net = s->hn.desc;
    const uint64_t n = s->st.n, S = s->st.stride;
    std::vector<uint32_t> ctl(n);
    CUDA_TRY(cudaMemcpy(ctl.data(), s->st.w + (uint64_t)net.wrow_ctl * S, n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < n; r++) {
        uint32_t np = cm.npend[r], err = CTL_ERR(ctl[r]);
        if (np > net.max_pending) {  // the new network's mailbox is smaller than what is in flight
            np = net.max_pending;
            if (!err) err = SIMB_ERR_CAPACITY;
        }
        ctl[r] = CTL_MAKE(err, CTL_EPOCH(ctl[r]), np, CTL_ABSENT(ctl[r]));
    }
    CUDA_TRY(cudaMemcpy(s->st.w + (uint64_t)net.wrow_ctl * S, ctl.data(), n * 4, cudaMemcpyHostToDevice));
    for (uint32_t j = 0; j < cm.max_np && j < net.max_pending; j++) {
        uint32_t* c = j < net.pend_tile ? s->st.w + (uint64_t)(net.wrow_pend + 2u * j) * S
                                        : s->st.mail + (uint64_t)(2u * (j - net.pend_tile)) * S;
        CUDA_TRY(cudaMemcpy(c, cm.content[j].data(), n * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c + S, cm.route[j].data(), n * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}
}  // namespace

int simb_put_json(simb_sim* s, const char* models_json, const char* connectors_json) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    // Simulation::put (mod.rs:82-85): new models (fresh state) and connectors; the clock, the
    // pending messages and the RNG position are kept.
    HostNet hn;
    std::string err;
    std::vector<uint32_t> name_map;
    bool same_layout = false, same_meaning = false;
    CarriedMail cm;
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    LowerOptions opt = s->opt;
    for (int attempt = 0; attempt < 2; attempt++) {
        err.clear();
        int e = lower_network(models_json, connectors_json, opt, &hn, &err);
        if (e) return fail(e, err);
        // content names interned so far (generator prefixes, preloaded jobs, injected contents) keep meaning something:
        // intern them into the new table; name_map[i] = new id of old name i
        name_map.assign(s->hn.names.size(), 0);
        bool overflow = false, same_names = true;
        for (size_t i = 0; i < s->hn.names.size(); i++) {
            name_map[i] = hn.intern(s->hn.names[i], &overflow);
            if (name_map[i] != i) same_names = false;
        }
        if (overflow) return fail(SIMB_ERR_CAPACITY, "content intern table overflow (more than 255 distinct names)");
        const SimbNetDesc& o = s->hn.desc;
        const SimbNetDesc& nw = hn.desc;
        same_layout = o.n_models == nw.n_models && o.n_connectors == nw.n_connectors &&  // counter arrays
                      o.n_drows == nw.n_drows && o.n_wrows == nw.n_wrows && o.wrow_pend == nw.wrow_pend &&
                      o.max_pending == nw.max_pending && o.wrow_rng == nw.wrow_rng && o.drow_time == nw.drow_time &&
                      o.wrow_ctl == nw.wrow_ctl && o.pend_tile == nw.pend_tile &&
                      s->hn.ring_w_slots == hn.ring_w_slots && s->hn.ring_d_slots == hn.ring_d_slots &&
                      s->hn.ring_preload.size() == hn.ring_preload.size();
        // do the (content, route) words in the mailboxes mean the same thing in the new network?
        same_meaning = same_layout && same_names && o.n_routes == nw.n_routes && s->hn.connectors.size() == hn.connectors.size();
        for (size_t i = 0; same_meaning && i < hn.models.size(); i++)
            same_meaning = s->hn.models[i].id == hn.models[i].id && s->hn.models[i].ports_in == hn.models[i].ports_in;
        for (size_t i = 0; same_meaning && i < hn.connectors.size(); i++) {
            const HostConnector &a = s->hn.connectors[i], &b = hn.connectors[i];
            same_meaning = a.source_id == b.source_id && a.source_port == b.source_port && a.target_id == b.target_id &&
                           a.target_port == b.target_port;
        }
        for (uint32_t r = 0; same_meaning && r < nw.n_routes; r++)
            same_meaning = o.routes[r].tgt_model == nw.routes[r].tgt_model && o.routes[r].tgt_port == nw.routes[r].tgt_port &&
                           o.routes[r].connector == nw.routes[r].connector;
        if (same_meaning || attempt == 1) break;
        // carry the pending messages over by name; the new mailbox must hold what is in flight (the reference's Vec
        // has no bound): lower once more with that as the explicit bound if the derived one is smaller
        if ((e = read_mail(s, &cm))) return e;
        if (cm.max_np) remap_mail(s->hn, hn, name_map, &cm);
        if (cm.max_np <= hn.desc.max_pending) break;
        // (capacities only: the second lowering has the same models, routes and names, so the remapped words stand)
        opt.max_pending = cm.max_np;
    }
    const SimbNetDesc& o = s->hn.desc;
    const SimbNetDesc& nw = hn.desc;
    int e = 0;
    {
        json::Parser pm(models_json), pc(connectors_json);
        pm.parse(s->models_doc);
        pc.parse(s->connectors_doc);
    }
    if (!same_layout) {
        // different state layout: the clock and the RNG position move to their new rows (device-to-device)
        const uint32_t rr = s->rng_kind == 1 ? 4 : 1;
        double* t_time = nullptr;
        uint32_t* t_rng = nullptr;
        CUDA_TRY(cudaMalloc((void**)&t_time, s->st.stride * 8));
        CUDA_TRY(cudaMalloc((void**)&t_rng, (size_t)rr * s->st.stride * 4));
        CUDA_TRY(cudaMemcpy(t_time, s->st.d + (uint64_t)o.drow_time * s->st.stride, s->st.stride * 8, cudaMemcpyDeviceToDevice));
        CUDA_TRY(cudaMemcpy(t_rng, s->st.w + (uint64_t)o.wrow_rng * s->st.stride, (size_t)rr * s->st.stride * 4, cudaMemcpyDeviceToDevice));
        free_device(s);
        s->hn = hn;
        e = build_device(s, 0, nullptr);
        if (!e) {
            const SimbNetDesc& n2 = s->hn.desc;
            cudaMemcpy(s->st.d + (uint64_t)n2.drow_time * s->st.stride, t_time, s->st.stride * 8, cudaMemcpyDeviceToDevice);
            cudaMemcpy(s->st.w + (uint64_t)n2.wrow_rng * s->st.stride, t_rng, (size_t)rr * s->st.stride * 4, cudaMemcpyDeviceToDevice);
        }
        cudaFree(t_time);
        cudaFree(t_rng);
        if (e) return e;
        return cm.max_np ? write_mail(s, cm) : 0;
    }
    s->hn = hn;
    if ((e = upload_tables(s))) return e;
    CUDA_TRY(cudaMemcpy(s->d_dinit, s->hn.d_init.data(), nw.n_drows * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(s->d_winit, s->hn.w_init.data(), nw.n_wrows * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (s->n_pre) {
        std::vector<uint32_t> slots(s->n_pre), vals(s->n_pre);
        for (uint32_t i = 0; i < s->n_pre; i++) {
            slots[i] = s->hn.ring_preload[i].slot;
            vals[i] = s->hn.ring_preload[i].value;
        }
        CUDA_TRY(cudaMemcpy(s->d_pre_slots, slots.data(), s->n_pre * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(s->d_pre_vals, vals.data(), s->n_pre * 4, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(engine_launch_init(s->hn.desc, s->st, s->d_dinit, s->d_winit, s->d_pre_slots, s->d_pre_vals, s->n_pre, s->rng_kind,
                                s->instrument, 1u | 2u | 4u, s->stream));
    if (s->instrument) {  // records live in the models that were just replaced (mod.rs:82-85): the capture starts over
        if (s->st.trace_count) CUDA_TRY(cudaMemsetAsync(s->st.trace_count, 0, (size_t)s->st.stride * sizeof(uint32_t), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->st.record_count, 0, (size_t)nw.n_models * ACT_COUNT * sizeof(unsigned long long), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->st.conn_count, 0, (size_t)(nw.n_connectors ? nw.n_connectors : 1) * sizeof(unsigned long long), s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (!same_meaning && cm.max_np) return write_mail(s, cm);
    return 0;
}

int simb_reset(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(engine_launch_reset(s->hn.desc, s->st, true, true, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_reset_messages(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(engine_launch_reset(s->hn.desc, s->st, true, false, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_reset_global_time(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(engine_launch_reset(s->hn.desc, s->st, false, true, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_set_rng(simb_sim* s, int rng_kind, uint64_t seed) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    if (rng_kind != s->rng_kind)
        return fail(SIMB_ERR_INVALID_ARGUMENT, "the RNG kind is fixed at post time (it decides the state layout); only the seed can change");
    s->cfg.seed = seed;
    s->st.seed = seed;
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(engine_launch_set_rng(s->hn.desc, s->st, s->rng_kind, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_set_rng_stream(simb_sim* s, const uint64_t* draws, uint64_t per_rep) {
    if (!s || !draws) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (s->rng_kind != SIMB_RNG_EXTERNAL) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_RNG_EXTERNAL");
    if (per_rep > 0xFFFFFFFFull) return fail(SIMB_ERR_INVALID_ARGUMENT, "too many draws per replication");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (s->d_ext) cudaFree(s->d_ext);
    s->d_ext = nullptr;
    CUDA_TRY(cudaMalloc((void**)&s->d_ext, std::max<uint64_t>(1, per_rep * s->st.stride) * 8));
    // caller layout [i][replications] -> device layout [i][stride]
    CUDA_TRY(cudaMemcpy2D(s->d_ext, s->st.stride * 8, draws, s->st.n * 8, s->st.n * 8, per_rep, cudaMemcpyHostToDevice));
    s->st.ext = s->d_ext;
    s->st.ext_len = (uint32_t)per_rep;
    CUDA_TRY(engine_launch_set_rng(s->hn.desc, s->st, s->rng_kind, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_destroy(simb_sim* s) {
    if (!s) return 0;
    cudaSetDevice(s->cfg.device);
    free_device(s);
    if (s->stream) cudaStreamSynchronize(s->stream);
    destroy_streams(s);
    delete s;
    return 0;
}

int simb_inject_input(simb_sim* s, const char* target_id, const char* target_port, const char* content,
                      const uint8_t* replica_mask) {
    if (!s || !target_id || !target_port || !content) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    bool overflow = false;
    const uint32_t code = s->hn.content_code(content, &overflow);
    if (overflow) return fail(SIMB_ERR_CAPACITY, "content intern table overflow");
    // deliveries: every posted model of that id (mod.rs:203-221), for a Coupled model the components its external input
    // couplings name for that port (coupled.rs:102-131); an unknown target is kept and never delivered
    std::vector<uint32_t> words;
    for (const HostNet::Delivery& d : s->hn.deliveries(target_id, target_port))
        words.push_back(ROUTE_WORD(d.model, d.port, words.empty() ? ROUTE_INJECTED : ROUTE_INJECTED_CONT, 0));
    if (words.empty()) words.push_back(ROUTE_WORD(0xFF, SIMB_PORT_UNKNOWN, ROUTE_INJECTED, 0));
    uint8_t* d_mask = nullptr;
    if (replica_mask) {
        CUDA_TRY(cudaMalloc((void**)&d_mask, s->st.n));
        const cudaError_t me = cudaMemcpy(d_mask, replica_mask, s->st.n, cudaMemcpyHostToDevice);
        if (me != cudaSuccess) {
            cudaFree(d_mask);
            return fail(SIMB_ERR_CUDA, std::string("mask upload: ") + cudaGetErrorString(me));
        }
    }
    cudaMemsetAsync(s->d_overflow, 0, sizeof(unsigned long long), s->stream);
    int e = 0;
    for (uint32_t rw : words)  // (the kernel adds the sequence number)
        if (!e) e = engine_launch_inject(s->hn.desc, s->st, code, rw, d_mask, s->d_overflow, s->stream);
    unsigned long long ov = 0;
    cudaMemcpyAsync(&ov, s->d_overflow, sizeof ov, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t ce = cudaStreamSynchronize(s->stream);
    if (d_mask) cudaFree(d_mask);
    if (e || ce != cudaSuccess) return fail(SIMB_ERR_CUDA, "inject kernel failed");
    if (ov) return fail(SIMB_ERR_CAPACITY, std::to_string(ov) + " replications have a full mailbox; raise simb_config.max_pending");
    return 0;
}
int simb_step(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const uint32_t k = s->K;
    s->K = 1;
    int e = run_steps(s, 0, 1, 0.0);
    s->K = k;
    return e;
}
int simb_step_n(simb_sim* s, uint64_t n) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return run_steps(s, 0, n, 0.0);
}
int simb_step_until(simb_sim* s, double until) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return run_steps(s, 1, 0, until);
}
int simb_synchronize(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int simb_get_global_time(simb_sim* s, double* out, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return copy_row_d(s, s->hn.desc.drow_time, out, capacity);
}
int simb_get_until_next_event(simb_sim* s, const char* model_id, double* out, uint64_t capacity) {
    if (!s || !model_id) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    const int m = s->hn.find_model(model_id);
    if (m < 0) return fail(SIMB_ERR_MODEL_NOT_FOUND, std::string("no model `") + model_id + "`");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return copy_row_d(s, s->hn.desc.drow_until + m, out, capacity);
}
int simb_get_messages(simb_sim* s, uint64_t rep, simb_message* out, uint64_t capacity, uint64_t* count) {
    if (!s || !count) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (rep >= s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication index out of range");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const SimbNetDesc& net = s->hn.desc;
    std::vector<uint32_t> w;
    std::vector<double> d;
    int e = column_w(s, rep, &w);
    if (e) return e;
    if ((e = column_d(s, rep, &d))) return e;
    const uint32_t np = CTL_NPEND(w[net.wrow_ctl]);
    *count = np;
    // the mailbox is kept in events_ext order (sorted by target model); the sequence number in the route word
    // restores the order the reference's Vec<Message> has (emission order, injected messages last)
    std::vector<std::pair<uint32_t, uint32_t>> mail(np);
    for (uint32_t j = 0; j < np; j++) {
        if (j < net.pend_tile) {
            mail[j].first = w[net.wrow_pend + 2 * j];
            mail[j].second = w[net.wrow_pend + 2 * j + 1];
        } else {
            const uint32_t* base = s->st.mail + (uint64_t)(2u * (j - net.pend_tile)) * s->st.stride + rep;
            CUDA_TRY(cudaMemcpy(&mail[j].first, base, 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(&mail[j].second, base + s->st.stride, 4, cudaMemcpyDeviceToHost));
        }
    }
    std::stable_sort(mail.begin(), mail.end(),
                     [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) {
                         return ROUTE_SEQ(a.second) < ROUTE_SEQ(b.second);
                     });
    {  // one Message that several models receive (a Coupled model's components, posted models sharing an id) sits in
       // the mailbox once per delivery: show it once
        std::vector<std::pair<uint32_t, uint32_t>> shown;
        for (auto& mw : mail) {
            const uint32_t r = ROUTE_R(mw.second);
            if (r == ROUTE_INJECTED_CONT) continue;
            if (r != ROUTE_INJECTED && r < net.n_routes && (net.routes[r].connector & RT_CONT)) continue;
            shown.push_back(mw);
        }
        mail.swap(shown);
        *count = mail.size();
    }
    const uint32_t np_shown = (uint32_t)mail.size();
    for (uint32_t j = 0; j < np_shown && j < capacity && out; j++) {
        const uint32_t content = mail[j].first, rw = mail[j].second;
        simb_message& m = out[j];
        std::memset(&m, 0, sizeof m);
        const uint32_t conn = ROUTE_R(rw) == ROUTE_INJECTED ? 0xFFFFu : RT_CONN(net.routes[ROUTE_R(rw)].connector);
        if (conn == 0xFFFF) {  // injected: Message::new("manual","manual",..) in the reference's tests
            put_str(m.source_id, sizeof m.source_id, "manual");
            put_str(m.source_port, sizeof m.source_port, "manual");
            const uint32_t tm = ROUTE_TGT(rw);
            if (tm < net.n_models) {
                const HostModel& hm = s->hn.models[tm];
                const uint32_t p = ROUTE_PORT(rw);
                const std::string pname = p < hm.ports_in.size() ? hm.ports_in[p] : std::string();
                if (hm.coupled >= 0) {  // injected into a Coupled model: shown under the model's id and the port whose
                                        // external input coupling led here
                    const HostCoupled& hc = s->hn.coupled[hm.coupled];
                    put_str(m.target_id, sizeof m.target_id, hc.id);
                    for (auto& ec : hc.eic)
                        if (ec.target_id == hm.id && ec.target_port == pname) {
                            put_str(m.target_port, sizeof m.target_port, ec.source_port);
                            break;
                        }
                } else {
                    put_str(m.target_id, sizeof m.target_id, hm.id);
                    put_str(m.target_port, sizeof m.target_port, pname);
                }
            }
        } else if (conn < s->hn.connectors.size()) {
            const HostConnector& c = s->hn.connectors[conn];
            put_str(m.source_id, sizeof m.source_id, c.source_id);
            put_str(m.source_port, sizeof m.source_port, c.source_port);
            put_str(m.target_id, sizeof m.target_id, c.target_id);
            put_str(m.target_port, sizeof m.target_port, c.target_port);
        }
        put_str(m.content, sizeof m.content, s->hn.render(content));
        m.time = d[net.drow_time];  // Message.time = global_time at emission = the current clock
    }
    return 0;
}
int simb_get_status(simb_sim* s, const char* model_id, uint64_t rep, char* out, uint64_t capacity) {
    if (!s || !model_id || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (rep >= s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication index out of range");
    {
        const int tt = s->hn.find_top(model_id);
        if (tt >= 0 && s->hn.tops[tt].coupled >= 0) {  // Coupled::status (coupled.rs:296-302; the branches are the way
                                                       // round the reference has them)
            CUDA_TRY(cudaSetDevice(s->cfg.device));
            std::vector<uint32_t> wc;
            int ec = column_w(s, rep, &wc);
            if (ec) return ec;
            const uint32_t parked = wc[s->hn.desc.coupled[s->hn.tops[tt].coupled].wrow];
            put_str(out, capacity, parked == 0 ? std::string("Processing 0 messages") : std::string("Processing no messages"));
            return 0;
        }
    }
    const int m = s->hn.find_model(model_id);
    if (m < 0) return fail(SIMB_ERR_MODEL_NOT_FOUND, std::string("no model `") + model_id + "`");  // mod.rs:111
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const SimbNetDesc& net = s->hn.desc;
    const SimbModelDesc& md = net.models[m];
    const HostModel& hm = s->hn.models[m];
    std::vector<uint32_t> w;
    int e = column_w(s, rep, &w);
    if (e) return e;
    const uint32_t ph = (w[net.wrow_phase + m / 16] >> (2 * (m % 16))) & 3u;
    auto head_job = [&]() -> std::string {
        const uint32_t head = w[md.wrow] >> 16;
        uint32_t c = 0;
        cudaMemcpy(&c, s->st.ring_w + (uint64_t)(md.ring_w0 + head) * s->st.stride + rep, 4, cudaMemcpyDeviceToHost);
        return s->hn.render(c);
    };
    std::string r;
    switch (md.type) {
        case MT_GENERATOR: r = "Generating " + hm.ports_out[0] + "s"; break;            // generator.rs:189-191
        case MT_PROCESSOR: r = ph == 1 ? "Processing" : "Passive"; break;                // processor.rs:249-256
        case MT_STORAGE:                                                                 // storage.rs:182-188
            r = w[md.wrow] == SIMB_CONTENT_NONE ? "Empty" : "Storing " + s->hn.render(w[md.wrow]);
            break;
        case MT_LOAD_BALANCER: r = "Listening for " + hm.ports_in[0] + "s"; break;       // load_balancer.rs:154-156
        case MT_GATE: r = ph == 0 ? "Open" : ph == 1 ? "Closed" : "Passing " + head_job(); break;  // gate.rs:207-213
        case MT_EXCLUSIVE_GATEWAY: r = ph == 1 ? "Passing " + head_job() : "Passive"; break;       // exclusive_gateway.rs:173-178
        case MT_PARALLEL_GATEWAY: r = "Active"; break;                                   // parallel_gateway.rs:193-195
        case MT_STOCHASTIC_GATE: r = "Gating"; break;                                    // stochastic_gate.rs:194-196
        case MT_STOPWATCH: r = "Measuring durations"; break;  // the running average of stopwatch.rs:306-322 is not kept on the device
        case MT_BATCHER: r = ph == 0 ? "Passive" : ph == 1 ? "Creating batch" : "Releasing batch"; break;  // batcher.rs:232-238
        default: r = "Passive"; break;                                                   // tests/custom.rs:77-79
    }
    put_str(out, capacity, r);
    return 0;
}
int simb_get_errors(simb_sim* s, int32_t* out, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    int e = copy_row_w(s, s->hn.desc.wrow_ctl, reinterpret_cast<uint32_t*>(out), capacity);
    if (e) return e;
    for (uint64_t r = 0; r < s->st.n; r++) out[r] = (int32_t)CTL_ERR((uint32_t)out[r]);
    return 0;
}
int simb_get_steps(simb_sim* s, uint32_t* out, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return copy_row_w(s, s->hn.desc.wrow_steps, out, capacity);
}
int simb_get_events(simb_sim* s, uint32_t* out, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return copy_row_w(s, s->hn.desc.wrow_events, out, capacity);
}
int simb_get_draws(simb_sim* s, uint32_t* out, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    if (s->rng_kind == SIMB_RNG_PCG64MCG) return fail(SIMB_ERR_INVALID_ARGUMENT, "the PCG state does not keep a draw count");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    return copy_row_w(s, s->hn.desc.wrow_rng, out, capacity);
}
int simb_get_counters(simb_sim* s, simb_counters* out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const SimbNetDesc& net = s->hn.desc;
    uint32_t nb = 0;
    CUDA_TRY(engine_launch_reduce(net, s->st, 2, 0.0, s->d_psum, s->d_pcnt, &nb, s->stream));
    std::vector<unsigned long long> a(nb), b(nb);
    unsigned long long ring_bytes = 0;
    CUDA_TRY(cudaMemcpyAsync(&ring_bytes, s->st.totals + 4, 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(a.data(), s->d_psum, nb * 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(b.data(), s->d_pcnt, nb * 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::memset(out, 0, sizeof *out);
    for (uint32_t i = 0; i < nb; i++) {
        out->steps += a[i];
        out->events += b[i];
    }
    out->launches = s->launches;
    out->kernel_ms = s->kernel_ms;
    out->state_bytes_per_replication = (uint64_t)net.n_drows * 8 + (uint64_t)net.n_wrows * 4;
    out->ring_bytes = ring_bytes;
    std::vector<uint32_t> ctl(s->st.n);
    CUDA_TRY(cudaMemcpy(ctl.data(), s->st.w + (uint64_t)net.wrow_ctl * s->st.stride, s->st.n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < s->st.n; r++) out->failed += CTL_ERR(ctl[r]) ? 1 : 0;
    return 0;
}
int simb_reset_counters(simb_sim* s) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    s->launches = 0;
    s->kernel_ms = 0.0;
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaMemsetAsync(s->st.totals + 4, 0, sizeof(unsigned long long), s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int simb_get_trace_hash(simb_sim* s, uint64_t* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->instrument) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_FLAG_INSTRUMENT");
    if (capacity < s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the replication count");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    std::vector<uint32_t> lo(s->st.n), hi(s->st.n);
    int e = copy_row_w(s, s->hn.desc.wrow_hash, lo.data(), s->st.n);
    if (e) return e;
    if ((e = copy_row_w(s, s->hn.desc.wrow_hash + 1, hi.data(), s->st.n))) return e;
    for (uint64_t r = 0; r < s->st.n; r++) out[r] = (uint64_t)lo[r] | ((uint64_t)hi[r] << 32);
    return 0;
}
int simb_get_message_counts(simb_sim* s, uint64_t* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->instrument) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_FLAG_INSTRUMENT");
    if (capacity < s->hn.desc.n_connectors) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the connector count");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaMemcpy(out, s->st.conn_count, s->hn.desc.n_connectors * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int simb_get_record_counts(simb_sim* s, uint64_t* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->instrument) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_FLAG_INSTRUMENT");
    const uint64_t n = (uint64_t)s->hn.desc.n_models * ACT_COUNT;
    if (capacity < n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer too small");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaMemcpy(out, s->st.record_count, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}
static int fetch_trace(simb_sim* s, uint64_t rep, std::vector<SimbTraceRec>* recs) {
    if (!s->instrument) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_FLAG_INSTRUMENT");
    if (rep >= s->st.trace_reps) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication is not traced (simb_config.trace_replications)");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    uint32_t cnt = 0;
    CUDA_TRY(cudaMemcpy(&cnt, s->st.trace_count + rep, 4, cudaMemcpyDeviceToHost));
    if (cnt > s->st.trace_cap)
        return fail(SIMB_ERR_CAPACITY, "trace buffer overflow: " + std::to_string(cnt) + " events > trace_capacity");
    recs->resize(cnt);
    if (cnt) CUDA_TRY(cudaMemcpy(recs->data(), s->st.trace + rep * s->st.trace_cap, (size_t)cnt * sizeof(SimbTraceRec), cudaMemcpyDeviceToHost));
    return 0;
}
int simb_get_trace(simb_sim* s, uint64_t rep, simb_trace_event* out, uint64_t capacity, uint64_t* count) {
    if (!s || !count) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<SimbTraceRec> recs;
    int e = fetch_trace(s, rep, &recs);
    if (e) return e;
    *count = recs.size();
    for (size_t i = 0; i < recs.size() && i < capacity && out; i++) {
        simb_trace_event& t = out[i];
        std::memset(&t, 0, sizeof t);
        t.step = recs[i].step;
        t.kind = (uint8_t)(recs[i].meta >> 31);
        t.index = (uint16_t)(recs[i].meta & 0xFFFFu);
        t.action = (uint8_t)(recs[i].aux & 0xFFu);
        t.aux = (uint8_t)((recs[i].aux >> 8) & 0xFFu);
        t.content = recs[i].content;
        t.time = recs[i].time;
    }
    return 0;
}
int simb_get_records(simb_sim* s, const char* model_id, uint64_t rep, simb_record* out, uint64_t capacity, uint64_t* count) {
    if (!s || !model_id || !count) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    const int m = s->hn.find_model(model_id);
    if (m < 0) return fail(SIMB_ERR_MODEL_NOT_FOUND, std::string("no model `") + model_id + "`");  // mod.rs:123
    std::vector<SimbTraceRec> recs;
    int e = fetch_trace(s, rep, &recs);
    if (e) return e;
    uint64_t k = 0;
    if (s->hn.models[m].store_records) {  // record() is a no-op unless store_records (e.g. processor.rs:198-206)
        for (auto& r : recs) {
            if (!(r.meta >> 31) || (r.meta & 0xFFFFu) != (uint32_t)m) continue;
            if (out && k < capacity) {
                simb_record& o = out[k];
                std::memset(&o, 0, sizeof o);
                o.time = r.time;
                put_str(o.action, sizeof o.action, action_name((int)(r.aux & 0xFFu)));
                put_str(o.subject, sizeof o.subject, record_subject(s, (uint32_t)m, r.aux & 0xFFu, (r.aux >> 8) & 0xFFu, r.content));
            }
            k++;
        }
    }
    *count = k;
    return 0;
}

// ---- WebSimulation string API (sim/src/simulator/web.rs:87-215) for one observed replication ------
namespace {
void message_json(std::string& o, const std::string& sid, const std::string& sport, const std::string& tid,
                  const std::string& tport, double time, const std::string& content) {
    // field order and camelCase names of coupling.rs:62-71
    o += "{\"sourceId\":";
    json::write_string(o, sid);
    o += ",\"sourcePort\":";
    json::write_string(o, sport);
    o += ",\"targetId\":";
    json::write_string(o, tid);
    o += ",\"targetPort\":";
    json::write_string(o, tport);
    o += ",\"time\":";
    json::write_f64(o, time);
    o += ",\"content\":";
    json::write_string(o, content);
    o += '}';
}
// messages the observed replication emitted in trace events [first, end): Simulation::step's return
// value (mod.rs:271) concatenated over the steps of the call (step_n :293-303, step_until :277-288)
int stepped_messages_json(simb_sim* s, uint64_t rep, uint32_t first, bool bounded, double until, const char** out) {
    std::vector<SimbTraceRec> recs;
    int e = fetch_trace(s, rep, &recs);
    if (e) return e;
    std::string& o = s->json;
    o = "[";
    bool any = false;
    for (size_t i = first; i < recs.size(); i++) {
        const SimbTraceRec& r = recs[i];
        if (r.meta >> 31) continue;                 // model record
        if (bounded && !(r.time < until)) continue; // the crossing step's messages are not returned (:281-285)
        const uint32_t conn = r.meta & 0xFFFFu;
        if (conn >= s->hn.connectors.size()) continue;
        const HostConnector& c = s->hn.connectors[conn];
        if (any) o += ',';
        any = true;
        message_json(o, c.source_id, c.source_port, c.target_id, c.target_port, r.time, s->hn.render(r.content));
    }
    o += ']';
    *out = o.c_str();
    return 0;
}
int traced_count(simb_sim* s, uint64_t rep, uint32_t* cnt) {
    if (!s->instrument) return fail(SIMB_ERR_INVALID_ARGUMENT, "handle was not posted with SIMB_FLAG_INSTRUMENT");
    if (rep >= s->st.trace_reps) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication is not traced (simb_config.trace_replications)");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaMemcpy(cnt, s->st.trace_count + rep, 4, cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace

int simb_get_messages_json(simb_sim* s, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint64_t n = 0;
    int e = simb_get_messages(s, rep, nullptr, 0, &n);
    if (e) return e;
    std::vector<simb_message> ms(n ? n : 1);
    if ((e = simb_get_messages(s, rep, ms.data(), n, &n))) return e;
    std::string& o = s->json;
    o = "[";
    for (uint64_t i = 0; i < n; i++) {
        if (i) o += ',';
        message_json(o, ms[i].source_id, ms[i].source_port, ms[i].target_id, ms[i].target_port, ms[i].time, ms[i].content);
    }
    o += ']';
    *out = o.c_str();
    return 0;
}
int simb_get_records_json(simb_sim* s, const char* model_id, uint64_t rep, const char** out) {
    if (!s || !model_id || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint64_t n = 0;
    int e = simb_get_records(s, model_id, rep, nullptr, 0, &n);
    if (e) return e;
    std::vector<simb_record> rs(n ? n : 1);
    if ((e = simb_get_records(s, model_id, rep, rs.data(), n, &n))) return e;
    std::string& o = s->json;
    o = "[";
    for (uint64_t i = 0; i < n; i++) {  // models/mod.rs:48-52
        if (i) o += ',';
        o += "{\"time\":";
        json::write_f64(o, rs[i].time);
        o += ",\"action\":";
        json::write_string(o, rs[i].action);
        o += ",\"subject\":";
        json::write_string(o, rs[i].subject);
        o += '}';
    }
    o += ']';
    *out = o.c_str();
    return 0;
}
int simb_inject_input_json(simb_sim* s, const char* message_json_text, const uint8_t* replica_mask) {
    if (!s || !message_json_text) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value v;
    json::Parser p(message_json_text);
    if (!p.parse(v) || !v.is_object()) return fail(SIMB_ERR_INVALID_ARGUMENT, "message is not a JSON object");
    // serde requires every field of Message (coupling.rs:62-71); web.rs:136-139 unwraps the error
    const char* keys[] = {"sourceId", "sourcePort", "targetId", "targetPort", "content"};
    for (const char* k : keys) {
        const json::Value* f = v.get(k);
        if (!f || !f->is_string()) return fail(SIMB_ERR_INVALID_ARGUMENT, std::string("message: missing field `") + k + "`");
    }
    const json::Value* t = v.get("time");
    if (!t || !t->is_number()) return fail(SIMB_ERR_INVALID_ARGUMENT, "message: missing field `time`");
    return simb_inject_input(s, v.get("targetId")->str.c_str(), v.get("targetPort")->str.c_str(),
                             v.get("content")->str.c_str(), replica_mask);
}
int simb_step_json(simb_sim* s, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint32_t first = 0;
    int e = traced_count(s, rep, &first);
    if (e) return e;
    if ((e = simb_step(s))) return e;
    return stepped_messages_json(s, rep, first, false, 0.0, out);
}
int simb_step_n_json(simb_sim* s, uint64_t n, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint32_t first = 0;
    int e = traced_count(s, rep, &first);
    if (e) return e;
    if ((e = simb_step_n(s, n))) return e;
    return stepped_messages_json(s, rep, first, false, 0.0, out);
}
int simb_step_until_json(simb_sim* s, double until, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint32_t first = 0;
    int e = traced_count(s, rep, &first);
    if (e) return e;
    if ((e = simb_step_until(s, until))) return e;
    return stepped_messages_json(s, rep, first, true, until, out);
}
int simb_format_f64(double v, char* out, uint64_t capacity) {
    if (!out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    std::string o;
    json::write_f64(o, v);
    put_str(out, capacity, o);
    return 0;
}

int simb_render_content(simb_sim* s, uint32_t content, char* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    put_str(out, capacity, s->hn.render(content));
    return 0;
}
int simb_intern_content(simb_sim* s, const char* content, uint32_t* out) {
    if (!s || !content || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    bool overflow = false;
    *out = s->hn.content_code(content, &overflow);
    return overflow ? fail(SIMB_ERR_CAPACITY, "content intern table overflow") : 0;
}


// ---- whole-simulation serialisation (WebSimulation::get_json / get_yaml, web.rs:43,69) -----------------------
namespace {
json::Value jstr(const std::string& v) {
    json::Value x;
    x.kind = json::Value::String;
    x.str = v;
    return x;
}
json::Value jnum(double v) {
    json::Value x;
    if (std::isfinite(v)) {
        x.kind = json::Value::Number;
        x.num = v;
    } else if (std::isinf(v) && v > 0) {  // serde_json writes a non-finite f64 as null; the YAML writer turns the
        x.kind = json::Value::Number;     // number back into `.inf`
        x.num = v;
    }
    return x;
}
json::Value juint(uint64_t v) {
    json::Value x;
    x.kind = json::Value::Number;
    x.num = (double)v;
    x.is_uint = true;
    x.u = v;
    return x;
}
json::Value jarr() {
    json::Value x;
    x.kind = json::Value::Array;
    return x;
}
json::Value jobj() {
    json::Value x;
    x.kind = json::Value::Object;
    return x;
}

// the `state` block of model m for one replication, in the field order of the reference's State structs
int model_state(simb_sim* s, uint64_t rep, uint32_t m, const std::vector<uint32_t>& w, const std::vector<double>& d,
                json::Value* out) {
    const SimbNetDesc& net = s->hn.desc;
    const SimbModelDesc& md = net.models[m];
    const uint32_t ph = (w[net.wrow_phase + m / 16] >> (2 * (m % 16))) & 3u;
    auto ring_w = [&](uint32_t slot, uint32_t* v) -> int {
        CUDA_TRY(cudaMemcpy(v, s->st.ring_w + (uint64_t)slot * s->st.stride + rep, 4, cudaMemcpyDeviceToHost));
        return 0;
    };
    auto ring_d = [&](uint32_t slot, double* v) -> int {
        CUDA_TRY(cudaMemcpy(v, s->st.ring_d + (uint64_t)slot * s->st.stride + rep, 8, cudaMemcpyDeviceToHost));
        return 0;
    };
    auto fifo = [&](json::Value* list) -> int {  // head | len ring -> Vec<String>
        *list = jarr();
        const uint32_t head = w[md.wrow] >> 16, len = w[md.wrow] & 0xFFFFu;
        for (uint32_t i = 0; i < len; i++) {
            uint32_t c = 0, slot = head + i;
            if (slot >= md.ring_cap) slot -= md.ring_cap;
            int e = ring_w(md.ring_w0 + slot, &c);
            if (e) return e;
            list->arr.push_back(jstr(s->hn.render(c)));
        }
        return 0;
    };
    json::Value st = jobj();
    auto phase = [&](std::initializer_list<const char*> names) {
        uint32_t i = 0;
        for (const char* n : names)
            if (i++ == ph) st.obj.push_back({"phase", jstr(n)});
    };
    const json::Value until = jnum(d[net.drow_until + m]);
    json::Value list;
    int e = 0;
    switch (md.type) {
        case MT_GENERATOR:  // generator.rs:51-58 (until_job is write-only in the reference: the last draw; the
                            // device keeps the remaining time only)
            phase({"Initializing", "Generating"});
            st.obj.push_back({"untilNextEvent", until});
            st.obj.push_back({"untilJob", until});
            st.obj.push_back({"lastJob", juint(w[md.wrow])});
            break;
        case MT_PROCESSOR:  // processor.rs:58-64
            phase({"Passive", "Active"});
            st.obj.push_back({"untilNextEvent", until});
            if ((e = fifo(&list))) return e;
            st.obj.push_back({"queue", list});
            break;
        case MT_STORAGE:  // storage.rs:41-47
            phase({"Passive", "JobFetch"});
            st.obj.push_back({"untilNextEvent", until});
            st.obj.push_back({"job", w[md.wrow] == SIMB_CONTENT_NONE ? json::Value() : jstr(s->hn.render(w[md.wrow]))});
            break;
        case MT_LOAD_BALANCER:  // load_balancer.rs:41-48
            phase({"Passive", "LoadBalancing"});
            st.obj.push_back({"untilNextEvent", until});
            st.obj.push_back({"nextPortOut", juint(w[md.wrow + 1])});
            if ((e = fifo(&list))) return e;
            st.obj.push_back({"jobs", list});
            break;
        case MT_GATE:  // gate.rs:45-51
            phase({"Open", "Closed", "Pass"});
            st.obj.push_back({"untilNextEvent", until});
            if ((e = fifo(&list))) return e;
            st.obj.push_back({"jobs", list});
            break;
        case MT_EXCLUSIVE_GATEWAY:  // exclusive_gateway.rs:55-61
            phase({"Passive", "Pass"});
            st.obj.push_back({"untilNextEvent", until});
            if ((e = fifo(&list))) return e;
            st.obj.push_back({"jobs", list});
            break;
        case MT_BATCHER:  // batcher.rs:46-52
            phase({"Passive", "Batching", "Release"});
            st.obj.push_back({"untilNextEvent", until});
            if ((e = fifo(&list))) return e;
            st.obj.push_back({"jobs", list});
            break;
        case MT_PARALLEL_GATEWAY: {  // parallel_gateway.rs:47-53 (insertion order, not HashMap order)
            st.obj.push_back({"untilNextEvent", until});
            json::Value col = jobj();
            for (uint32_t i = 0; i < w[md.wrow]; i++) {
                uint32_t c = 0, n = 0;
                if ((e = ring_w(md.ring_w0 + i, &c)) || (e = ring_w(md.ring_w1 + i, &n))) return e;
                col.obj.push_back({s->hn.render(c), juint(n)});
            }
            st.obj.push_back({"collections", col});
            break;
        }
        case MT_STOCHASTIC_GATE: {  // stochastic_gate.rs:50-64
            st.obj.push_back({"untilNextEvent", until});
            list = jarr();
            const uint32_t head = w[md.wrow] >> 16, len = w[md.wrow] & 0xFFFFu;
            for (uint32_t i = 0; i < len; i++) {
                uint32_t c = 0, pass = 0, slot = head + i;
                if (slot >= md.ring_cap) slot -= md.ring_cap;
                if ((e = ring_w(md.ring_w0 + slot, &c)) || (e = ring_w(md.ring_w1 + slot, &pass))) return e;
                json::Value jb = jobj();
                jb.obj.push_back({"content", jstr(s->hn.render(c))});
                json::Value b;
                b.kind = json::Value::Bool;
                b.b = pass != 0;
                jb.obj.push_back({"pass", b});
                list.arr.push_back(jb);
            }
            st.obj.push_back({"jobs", list});
            break;
        }
        case MT_STOPWATCH: {  // stopwatch.rs:66-93.  The device keeps the jobs that still miss a side plus the best
                              // completed one (DESIGN.md deviations): that one is written as {start 0, stop duration},
                              // which selects the same job when the document is posted again.
            phase({"Passive", "JobFetch"});
            st.obj.push_back({"untilNextEvent", until});
            struct Row {
                uint32_t seq;
                json::Value v;
            };
            std::vector<Row> rows;
            if (w[md.wrow + 1] != SIMB_CONTENT_NONE) {
                json::Value jb = jobj();
                jb.obj.push_back({"name", jstr(s->hn.render(w[md.wrow + 1]))});
                jb.obj.push_back({"start", jnum(0.0)});
                jb.obj.push_back({"stop", jnum(d[md.drow])});
                rows.push_back(Row{w[md.wrow + 3], jb});
            }
            for (uint32_t i = 0; i < w[md.wrow]; i++) {
                uint32_t c = 0, fl = 0;
                double t = 0.0;
                if ((e = ring_w(md.ring_w0 + i, &c)) || (e = ring_w(md.ring_w1 + i, &fl)) || (e = ring_d(md.ring_d + i, &t))) return e;
                json::Value jb = jobj();
                jb.obj.push_back({"name", jstr(s->hn.render(c))});
                jb.obj.push_back({"start", (fl & 1u) ? jnum(t) : json::Value()});
                jb.obj.push_back({"stop", (fl & 2u) ? jnum(t) : json::Value()});
                rows.push_back(Row{fl >> 2, jb});
            }
            std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.seq < b.seq; });
            list = jarr();
            for (auto& r : rows) list.arr.push_back(r.v);
            st.obj.push_back({"jobs", list});
            break;
        }
        default:  // Passive (tests/custom.rs:18-86) carries no state
            *out = json::Value();
            return 0;
    }
    // records: the model's own Vec<ModelRecord> (kept for traced replications of an instrumented handle)
    json::Value recs = jarr();
    if (s->instrument && rep < s->st.trace_reps && s->hn.models[m].store_records) {
        uint64_t n = 0;
        const char* id = s->hn.models[m].id.c_str();
        if ((e = simb_get_records(s, id, rep, nullptr, 0, &n))) return e;
        std::vector<simb_record> rs(n ? n : 1);
        if ((e = simb_get_records(s, id, rep, rs.data(), n, &n))) return e;
        for (uint64_t i = 0; i < n; i++) {
            json::Value r = jobj();
            r.obj.push_back({"time", jnum(rs[i].time)});
            r.obj.push_back({"action", jstr(rs[i].action)});
            r.obj.push_back({"subject", jstr(rs[i].subject)});
            recs.arr.push_back(r);
        }
    }
    st.obj.push_back({"records", recs});
    *out = st;
    return 0;
}

// Simulation { models, connectors, messages, services } (mod.rs:37-45) of one replication
int simulation_doc(simb_sim* s, uint64_t rep, json::Value* doc) {
    if (rep >= s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication index out of range");

    CUDA_TRY(cudaSetDevice(s->cfg.device));
    std::vector<uint32_t> w;
    std::vector<double> d;
    int e = column_w(s, rep, &w);
    if (e) return e;
    if ((e = column_d(s, rep, &d))) return e;
    // Model::serialize (model.rs:28-41): id, type, then the model's own fields -- the posted document's, with `state`
    // replaced by the live state
    auto model_doc = [&](uint32_t m, const json::Value* src, json::Value* out) -> int {
        json::Value mo = jobj();
        mo.obj.push_back({"id", jstr(s->hn.models[m].id)});
        mo.obj.push_back({"type", jstr(s->hn.models[m].type)});
        if (src && src->is_object())
            for (auto& kv : src->obj)
                if (kv.first != "id" && kv.first != "type" && kv.first != "state") mo.obj.push_back(kv);
        json::Value st;
        int e2 = model_state(s, rep, m, w, d, &st);
        if (e2) return e2;
        if (st.kind == json::Value::Object) mo.obj.push_back({"state", st});
        *out = mo;
        return 0;
    };
    json::Value models = jarr();
    for (size_t t = 0; t < s->hn.tops.size(); t++) {
        const HostTop& top = s->hn.tops[t];
        const json::Value* src = (s->models_doc.is_array() && t < s->models_doc.arr.size()) ? &s->models_doc.arr[t] : nullptr;
        json::Value mo;
        if (top.coupled < 0) {
            if ((e = model_doc((uint32_t)top.flat, src, &mo))) return e;
            models.arr.push_back(mo);
            continue;
        }
        // Coupled (coupled.rs:14-25): its fields as posted, the components with their live state, and its own state:
        // the parked messages (:69-84)
        const HostCoupled& hc = s->hn.coupled[top.coupled];
        const SimbCoupledDesc& cd = s->hn.desc.coupled[top.coupled];
        mo = jobj();
        mo.obj.push_back({"id", jstr(hc.id)});
        mo.obj.push_back({"type", jstr("Coupled")});
        const json::Value* comps_src = src ? src->get("components") : nullptr;
        if (src && src->is_object())
            for (auto& kv : src->obj) {
                if (kv.first == "id" || kv.first == "type" || kv.first == "state") continue;
                if (kv.first != "components") {
                    mo.obj.push_back(kv);
                    continue;
                }
                json::Value comps = jarr();
                for (int k = 0; k < hc.count; k++) {
                    json::Value cm;
                    const json::Value* csrc = (comps_src && comps_src->is_array() && (size_t)k < comps_src->arr.size()) ? &comps_src->arr[k] : nullptr;
                    if ((e = model_doc((uint32_t)(hc.first + k), csrc, &cm))) return e;
                    comps.arr.push_back(cm);
                }
                mo.obj.push_back({"components", comps});
            }
        json::Value st = jobj(), parked = jarr();
        for (uint32_t i = 0; i < w[cd.wrow]; i++) {
            uint32_t c = 0, tp = 0;
            CUDA_TRY(cudaMemcpy(&c, s->st.ring_w + (uint64_t)(cd.ring_w0 + i) * s->st.stride + rep, 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(&tp, s->st.ring_w + (uint64_t)(cd.ring_w1 + i) * s->st.stride + rep, 4, cudaMemcpyDeviceToHost));
            const uint32_t comp = tp & 0xFFu, port = (tp >> 8) & 0xFFu;
            json::Value pm = jobj();
            pm.obj.push_back({"componentId", jstr(comp < s->hn.models.size() ? s->hn.models[comp].id : std::string())});
            pm.obj.push_back({"port", jstr(comp < s->hn.models.size() && port < s->hn.models[comp].ports_in.size()
                                               ? s->hn.models[comp].ports_in[port] : std::string())});
            pm.obj.push_back({"content", jstr(s->hn.render(c))});
            parked.arr.push_back(pm);
        }
        st.obj.push_back({"parkedMessages", parked});
        st.obj.push_back({"records", jarr()});
        mo.obj.push_back({"state", st});
        models.arr.push_back(mo);
    }
    json::Value msgs = jarr();
    {
        uint64_t n = 0;
        if ((e = simb_get_messages(s, rep, nullptr, 0, &n))) return e;
        std::vector<simb_message> ms(n ? n : 1);
        if ((e = simb_get_messages(s, rep, ms.data(), n, &n))) return e;
        for (uint64_t i = 0; i < n; i++) {  // coupling.rs:62-71
            json::Value mo = jobj();
            mo.obj.push_back({"sourceId", jstr(ms[i].source_id)});
            mo.obj.push_back({"sourcePort", jstr(ms[i].source_port)});
            mo.obj.push_back({"targetId", jstr(ms[i].target_id)});
            mo.obj.push_back({"targetPort", jstr(ms[i].target_port)});
            mo.obj.push_back({"time", jnum(ms[i].time)});
            mo.obj.push_back({"content", jstr(ms[i].content)});
            msgs.arr.push_back(mo);
        }
    }
    json::Value services = jobj();  // services.rs:9-14: the RNG is #[serde(skip)]
    services.obj.push_back({"globalTime", jnum(d[s->hn.desc.drow_time])});
    *doc = jobj();
    doc->obj.push_back({"models", models});
    doc->obj.push_back({"connectors", s->connectors_doc.is_array() ? s->connectors_doc : jarr()});
    doc->obj.push_back({"messages", msgs});
    doc->obj.push_back({"services", services});
    return 0;
}

// ---- checkpoint: the device state of the whole batch as one blob ---------------------------------------------
struct StateHeader {
    char magic[8];  // "SIMBST02"
    uint64_t fingerprint, n, stride, rep_offset, seed;
    uint32_t rng_kind, instrument, epoch, reserved;
    uint64_t passes;
    uint64_t bytes[12];  // d, w, ring_w, ring_d, mail, series, trace, trace_count, conn_count, record_count, totals, ext
};
uint64_t fnv(const void* p, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001B3ull;
    return h;
}
void state_sections(simb_sim* s, void* ptr[12], uint64_t bytes[12]) {
    const SimbNetDesc& net = s->hn.desc;
    const SimbDevState& st = s->st;
    const uint64_t S = st.stride;
    ptr[0] = st.d;            bytes[0] = (uint64_t)net.n_drows * S * 8;
    ptr[1] = st.w;            bytes[1] = (uint64_t)net.n_wrows * S * 4;
    ptr[2] = st.ring_w;       bytes[2] = (uint64_t)(s->hn.ring_w_slots + 1) * S * 4;
    ptr[3] = st.ring_d;       bytes[3] = (uint64_t)(s->hn.ring_d_slots + 1) * S * 8;
    ptr[4] = st.mail;         bytes[4] = (uint64_t)(2u * (net.max_pending - net.pend_tile) + 1) * S * 4;
    ptr[5] = st.series;       bytes[5] = st.series ? (uint64_t)st.series_cap * S * 8 : 0;
    ptr[6] = st.trace;        bytes[6] = st.trace ? (uint64_t)st.trace_reps * st.trace_cap * sizeof(SimbTraceRec) : 0;
    ptr[7] = st.trace_count;  bytes[7] = st.trace_count ? S * 4 : 0;
    ptr[8] = st.conn_count;   bytes[8] = st.conn_count ? (uint64_t)(net.n_connectors + 1) * 8 : 0;
    ptr[9] = st.record_count; bytes[9] = st.record_count ? (uint64_t)net.n_models * ACT_COUNT * 8 : 0;
    ptr[10] = st.totals;      bytes[10] = 5 * 8;
    ptr[11] = s->d_ext;       bytes[11] = s->d_ext ? (uint64_t)std::max<uint64_t>(1, (uint64_t)st.ext_len * S) * 8 : 0;
}
uint64_t state_fingerprint(simb_sim* s) {
    uint64_t h = fnv(&s->hn.desc, sizeof(SimbNetDesc));
    const uint64_t extra[6] = {s->st.stride, s->hn.ring_w_slots, s->hn.ring_d_slots, (uint64_t)s->rng_kind,
                               (uint64_t)s->instrument, ((uint64_t)s->st.trace_reps << 32) | s->st.trace_cap};
    return fnv(extra, sizeof extra, h);
}
}  // namespace

int simb_get_json(simb_sim* s, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value doc;
    int e = simulation_doc(s, rep, &doc);
    if (e) return e;
    s->json.clear();
    yaml::write_json(s->json, doc, 2);  // serde_json::to_string_pretty (web.rs:44)
    *out = s->json.c_str();
    return 0;
}
int simb_get_yaml(simb_sim* s, uint64_t rep, const char** out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value doc;
    int e = simulation_doc(s, rep, &doc);
    if (e) return e;
    s->json.clear();
    yaml::write_yaml(s->json, doc);  // serde_yaml::to_string (web.rs:70)
    *out = s->json.c_str();
    return 0;
}
static int yaml_pair_to_json(const char* models_yaml, const char* connectors_yaml, std::string* mj, std::string* cj) {
    if (!models_yaml || !connectors_yaml) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value m, c;
    std::string err;
    if (!yaml::to_value(models_yaml, &m, &err)) return fail(SIMB_ERR_JSON, "models: " + err);
    if (!yaml::to_value(connectors_yaml, &c, &err)) return fail(SIMB_ERR_JSON, "connectors: " + err);
    yaml::write_json(*mj, m);
    yaml::write_json(*cj, c);
    return 0;
}
int simb_post_yaml(const char* models_yaml, const char* connectors_yaml, const simb_config* config, simb_sim** out) {
    std::string mj, cj;
    int e = yaml_pair_to_json(models_yaml, connectors_yaml, &mj, &cj);
    if (e) return e;
    return simb_post_json(mj.c_str(), cj.c_str(), config, out);
}
int simb_put_yaml(simb_sim* s, const char* models_yaml, const char* connectors_yaml) {
    std::string mj, cj;
    int e = yaml_pair_to_json(models_yaml, connectors_yaml, &mj, &cj);
    if (e) return e;
    return simb_put_json(s, mj.c_str(), cj.c_str());
}
int simb_yaml_to_json(const char* yaml_text, char* out, uint64_t capacity, uint64_t* needed) {
    if (!yaml_text) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value v;
    std::string err, o;
    if (!yaml::to_value(yaml_text, &v, &err)) return fail(SIMB_ERR_JSON, err);
    yaml::write_json(o, v);
    if (needed) *needed = o.size() + 1;
    if (out && capacity) put_str(out, capacity, o);
    return 0;
}
int simb_json_to_yaml(const char* json_text, char* out, uint64_t capacity, uint64_t* needed) {
    if (!json_text) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    json::Value v;
    json::Parser p(json_text);
    if (!p.parse(v)) return fail(SIMB_ERR_JSON, p.err);
    std::string o;
    yaml::write_yaml(o, v);
    if (needed) *needed = o.size() + 1;
    if (out && capacity) put_str(out, capacity, o);
    return 0;
}

int simb_state_size(simb_sim* s, uint64_t* bytes) {
    if (!s || !bytes) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    void* ptr[12];
    uint64_t sz[12];
    state_sections(s, ptr, sz);
    uint64_t total = sizeof(StateHeader);
    for (int i = 0; i < 12; i++) total += sz[i];
    *bytes = total;
    return 0;
}
int simb_get_state(simb_sim* s, void* out, uint64_t capacity, uint64_t* written) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    uint64_t total = 0;
    simb_state_size(s, &total);
    if (written) *written = total;
    if (capacity < total) return fail(SIMB_ERR_INVALID_ARGUMENT, "state buffer too small (simb_state_size)");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    StateHeader h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, "SIMBST02", 8);
    h.fingerprint = state_fingerprint(s);
    h.n = s->st.n;
    h.stride = s->st.stride;
    h.rep_offset = s->st.rep_offset;
    h.seed = s->st.seed;
    h.rng_kind = (uint32_t)s->rng_kind;
    h.instrument = s->instrument ? 1u : 0u;
    h.epoch = s->epoch;
    h.passes = s->passes;
    void* ptr[12];
    state_sections(s, ptr, h.bytes);
    unsigned char* o = (unsigned char*)out;
    std::memcpy(o, &h, sizeof h);
    o += sizeof h;
    for (int i = 0; i < 12; i++) {
        if (h.bytes[i]) CUDA_TRY(cudaMemcpy(o, ptr[i], h.bytes[i], cudaMemcpyDeviceToHost));
        o += h.bytes[i];
    }
    return 0;
}
int simb_put_state(simb_sim* s, const void* data, uint64_t bytes) {
    if (!s || !data) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes < sizeof(StateHeader)) return fail(SIMB_ERR_INVALID_ARGUMENT, "not a state blob");
    StateHeader h;
    std::memcpy(&h, data, sizeof h);
    if (std::memcmp(h.magic, "SIMBST02", 8) != 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "not a state blob (magic)");
    void* ptr[12];
    uint64_t sz[12];
    state_sections(s, ptr, sz);
    // the external stream is an input, not state: a blob taken with one loaded restores into a handle that has the
    // same stream loaded, or none of it
    if (h.fingerprint != state_fingerprint(s) || h.n != s->st.n || h.rng_kind != (uint32_t)s->rng_kind)
        return fail(SIMB_ERR_INVALID_ARGUMENT, "the state was taken from a different network / batch shape: post the same models, "
                                                "connectors and simb_config first");
    uint64_t total = sizeof h;
    for (int i = 0; i < 12; i++) {
        if (i == 11 && h.bytes[i] != sz[i]) return fail(SIMB_ERR_INVALID_ARGUMENT, "load the same external draw stream (simb_set_rng_stream) before simb_put_state");
        if (h.bytes[i] != sz[i]) return fail(SIMB_ERR_INVALID_ARGUMENT, "state section size mismatch");
        total += sz[i];
    }
    if (bytes < total) return fail(SIMB_ERR_INVALID_ARGUMENT, "truncated state blob");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const unsigned char* in = (const unsigned char*)data + sizeof h;
    for (int i = 0; i < 12; i++) {
        if (sz[i]) CUDA_TRY(cudaMemcpy(ptr[i], in, sz[i], cudaMemcpyHostToDevice));
        in += sz[i];
    }
    s->st.rep_offset = h.rep_offset;  // the streams are keyed by the global replication index and the seed
    s->st.seed = h.seed;
    s->cfg.replication_offset = h.rep_offset;
    s->cfg.seed = h.seed;
    s->epoch = h.epoch;
    s->passes = h.passes;
    return 0;
}

// ---- input_modeling on the device (random_variable.rs:65-131, thinning.rs:27-34) ----------------------------------
int simb_sample_random_variable(const char* family, const char* variable_json, uint64_t replications, uint64_t count,
                                uint64_t seed, uint64_t replication_offset, int device, double* out_f64, uint64_t* out_u64,
                                uint32_t* draws_consumed) {
    if (!family || !variable_json) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    int fam;
    const std::string f = family;
    if (f == "continuous") fam = FAM_CONTINUOUS;
    else if (f == "boolean") fam = FAM_BOOLEAN;
    else if (f == "discrete") fam = FAM_DISCRETE;
    else if (f == "index") fam = FAM_INDEX;
    else return fail(SIMB_ERR_INVALID_ARGUMENT, "family must be continuous, boolean, discrete or index");
    if (replications == 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "replications must be > 0");
    if (fam == FAM_CONTINUOUS ? !out_f64 : !out_u64)
        return fail(SIMB_ERR_INVALID_ARGUMENT, "continuous variates go to out_f64, all others to out_u64");
    SimbNetDesc net;
    std::string err;
    int e = lower_random_variable(fam, variable_json, &net, &err);
    if (e) return fail(e, err);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SIMB_ERR_CUDA, "no CUDA device available; libsimb200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    SimbDevState st;
    std::memset(&st, 0, sizeof st);
    st.n = st.stride = replications;
    st.rep_offset = replication_offset;
    st.seed = seed;
    const size_t total = (size_t)(replications * count);
    double* d_d = nullptr;
    unsigned long long* d_u = nullptr;
    uint32_t* d_draws = nullptr;
    int32_t* d_status = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_d);
        cudaFree(d_u);
        cudaFree(d_draws);
        cudaFree(d_status);
    };
    cudaError_t ce = cudaSuccess;
    if (fam == FAM_CONTINUOUS) ce = cudaMalloc((void**)&d_d, std::max<size_t>(1, total) * 8);
    else ce = cudaMalloc((void**)&d_u, std::max<size_t>(1, total) * 8);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_draws, replications * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_status, 4);
    if (ce == cudaSuccess) ce = cudaMemset(d_status, 0, 4);
    if (ce == cudaSuccess) ce = (cudaError_t)engine_launch_sample(net, st, (uint32_t)fam, count, d_d, d_u, d_draws, d_status, nullptr);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    int32_t status = 0;
    if (ce == cudaSuccess) ce = cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && total)
        ce = fam == FAM_CONTINUOUS ? cudaMemcpy(out_f64, d_d, total * 8, cudaMemcpyDeviceToHost)
                                   : cudaMemcpy(out_u64, d_u, total * 8, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && draws_consumed) ce = cudaMemcpy(draws_consumed, d_draws, replications * 4, cudaMemcpyDeviceToHost);
    cleanup();
    if (ce != cudaSuccess) return fail(SIMB_ERR_CUDA, std::string("simb_sample_random_variable: ") + cudaGetErrorString(ce));
    if (status) return fail(status, std::string("random variable parameters: ") + status_name(status));  // raised at draw time
    return 0;
}
int simb_thinning_evaluate(const double* coefficients, uint64_t n, double point, double* out) {
    if (!out || (n && !coefficients)) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    // Thinning::evaluate -> utils::evaluate_polynomial (utils/mod.rs:34-69): coefficients from the highest order down
    double acc = 0.0;
    for (uint64_t i = 0; i < n; i++) acc = acc * point + coefficients[i];
    *out = acc;
    return 0;
}

// ---- statistics ---------------------------------------------------------------------------------------
int simb_get_wait_stats(simb_sim* s, double* sum, uint32_t* count, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    if (s->hn.stat_model < 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "no stat_model_id was configured");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    int e = 0;
    if (sum && (e = copy_row_d(s, s->hn.desc.drow_wait, sum, capacity))) return e;
    if (count && (e = copy_row_w(s, s->hn.desc.wrow_waitn, count, capacity))) return e;
    return 0;
}
int simb_get_batch_means(simb_sim* s, double* mean, double* variance, uint32_t* batches, uint64_t capacity) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    const SimbNetDesc& net = s->hn.desc;
    if (s->hn.stat_model < 0 || net.bm_size == 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "batch means were not configured (simb_config.batch_size)");
    if (capacity < s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the replication count");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    std::vector<double> sm(s->st.n), sq(s->st.n);
    std::vector<uint32_t> jobs(s->st.n);
    int e;
    if ((e = copy_row_d(s, net.drow_bm + 1, sm.data(), s->st.n))) return e;
    if ((e = copy_row_d(s, net.drow_bm + 2, sq.data(), s->st.n))) return e;
    if ((e = copy_row_w(s, net.wrow_waitn, jobs.data(), s->st.n))) return e;
    for (uint64_t r = 0; r < s->st.n; r++) {
        uint32_t b = jobs[r] > net.bm_warmup ? (jobs[r] - net.bm_warmup) / net.bm_size : 0u;
        if (b > SIMB_MAX_BATCHES) b = SIMB_MAX_BATCHES;
        const double m = b ? sm[r] / (double)b : std::nan("");
        if (batches) batches[r] = b;
        if (mean) mean[r] = m;
        // population variance in one pass over the running sums; cancellation on nearly equal batch means must not
        // produce a (slightly) negative value, whose square root would poison the interval
        if (variance) variance[r] = b ? std::max(0.0, sq[r] / (double)b - m * m) : std::nan("");
    }
    return 0;
}
int simb_steady_state(simb_sim* s, double alpha, simb_steady_state_result* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->st.series) return fail(SIMB_ERR_INVALID_ARGUMENT, "no series was captured (simb_config.series_capacity, stat_model_id)");
    if (capacity < s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "output buffer smaller than the replication count");
    static_assert(sizeof(simb_steady_state_result) == sizeof(SimbSteadyState), "device / ABI result layouts differ");
    double t_lo[31] = {0}, t_hi[31] = {0};
    for (uint64_t b = 1; b <= 30; b++) {  // an unlisted alpha panics in the reference (t_scores.rs:19-22)
        int e = oa::t_score(alpha, b, &t_lo[b]);
        if (e) return fail(e, "alpha is not one of the tabulated levels");
        if (b >= 2) oa::t_score(alpha, b - 1, &t_hi[b]);
    }
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    SimbSteadyState* d_out = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_out, s->st.n * sizeof(SimbSteadyState), s->stream));
    int e = engine_launch_steady_state(s->hn.desc, s->st, t_lo, t_hi, d_out, s->stream);
    if (!e) e = (int)cudaMemcpyAsync(out, d_out, s->st.n * sizeof(SimbSteadyState), cudaMemcpyDeviceToHost, s->stream);
    cudaFreeAsync(d_out, s->stream);
    CUDA_TRY(e);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}
int simb_get_series(simb_sim* s, uint64_t rep, double* out, uint64_t capacity, uint64_t* count) {
    if (!s || !count) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->st.series) return fail(SIMB_ERR_INVALID_ARGUMENT, "no series was captured (simb_config.series_capacity, stat_model_id)");
    if (rep >= s->st.n) return fail(SIMB_ERR_INVALID_ARGUMENT, "replication index out of range");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    uint32_t n = 0;
    CUDA_TRY(cudaMemcpy(&n, s->st.w + (uint64_t)s->hn.desc.wrow_waitn * s->st.stride + rep, 4, cudaMemcpyDeviceToHost));
    *count = n;  // jobs that started service; only the first series_capacity of them are kept
    const uint64_t have = std::min<uint64_t>(n, s->st.series_cap);
    const uint64_t take = std::min<uint64_t>(have, capacity);
    if (out && take)
        CUDA_TRY(cudaMemcpy2D(out, 8, s->st.series + rep, s->st.stride * 8, 8, take, cudaMemcpyDeviceToHost));
    return 0;
}
int simb_oa_batch_means_ci(double mean, double variance, uint64_t batch_count, double alpha, simb_ci* out) {
    if (!out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    double lo, hi;
    int e = oa::batch_means_ci(mean, variance, batch_count, alpha, &lo, &hi);
    if (e) return fail(e, "t_score: unsupported alpha or no completed batch");
    out->mean = mean;
    out->variance = variance;
    out->lower = lo;
    out->upper = hi;
    out->n = batch_count;
    return 0;
}
static int stat_reduce(simb_sim* s, int which, double mean, double* sum, uint64_t* n) {
    if (s->hn.stat_model < 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "no stat_model_id was configured");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    uint32_t nb = 0;
    CUDA_TRY(engine_launch_reduce(s->hn.desc, s->st, which, mean, s->d_psum, s->d_pcnt, &nb, s->stream));
    std::vector<double> ps(nb);
    std::vector<unsigned long long> pc(nb);
    CUDA_TRY(cudaMemcpyAsync(ps.data(), s->d_psum, nb * 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(pc.data(), s->d_pcnt, nb * 8, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    double t = 0.0;
    uint64_t c = 0;
    for (uint32_t i = 0; i < nb; i++) {  // fixed order: results do not depend on scheduling
        t += ps[i];
        c += pc[i];
    }
    if (sum) *sum = t;
    if (n) *n = c;
    return 0;
}
int simb_stat_partial(simb_sim* s, double* sum_of_means, uint64_t* n) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    return stat_reduce(s, 0, 0.0, sum_of_means, n);
}
int simb_stat_partial_sqdev(simb_sim* s, double mean, double* sum_sqdev) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    return stat_reduce(s, 1, mean, sum_sqdev, nullptr);
}
int simb_independent_sample_ci(simb_sim* s, double alpha, simb_ci* out) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    double sum = 0, sq = 0;
    uint64_t n = 0;
    int e = stat_reduce(s, 0, 0.0, &sum, &n);
    if (e) return e;
    if (n == 0) return fail(SIMB_ERR_PREREQUISITE_CALC, "no replication has a defined mean yet");
    const double mean = sum / (double)n;  // sample_mean (output_analysis/mod.rs:23-28)
    if ((e = stat_reduce(s, 1, mean, &sq, nullptr))) return e;
    const double variance = sq / (double)n;  // sample_variance divides by n (mod.rs:32-40)
    return simb_oa_confidence_interval(mean, variance, n, alpha, out);
}
int simb_oa_confidence_interval(double mean, double variance, uint64_t n, double alpha, simb_ci* out) {
    if (!out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    double lo, hi;
    int e = oa::independent_ci(mean, variance, n, alpha, &lo, &hi);
    if (e) return fail(e, "t_score: alpha must be one of .1 .05 .025 .01 .005 .001 .0005 (t_scores.rs:10-22)");
    out->mean = mean;
    out->variance = variance;
    out->lower = lo;
    out->upper = hi;
    out->n = n;
    return 0;
}
int simb_oa_independent_sample(const double* points, uint64_t n, double alpha, simb_ci* out) {
    if (!points || !out || n == 0) return fail(SIMB_ERR_INVALID_ARGUMENT, "null or empty sample");
    double mean, var;
    oa::mean_variance(points, n, &mean, &var);
    return simb_oa_confidence_interval(mean, var, n, alpha, out);
}
int simb_oa_steady_state(const double* series, uint64_t n, double alpha, simb_ci* out, uint64_t* deletion_point,
                         uint64_t* batch_size, uint64_t* batch_count) {
    if (!series || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    oa::SteadyState ss;
    int e = oa::steady_state(series, n, alpha, &ss);
    if (e) return fail(e, e == SIMB_ERR_PANIC ? "the reference panics on this input (series shorter than 2, or unknown alpha)"
                                              : "steady state analysis failed");
    out->mean = ss.mean;
    out->variance = ss.variance;
    out->lower = ss.lower;
    out->upper = ss.upper;
    out->n = ss.batch_count;
    if (deletion_point) *deletion_point = ss.deletion_point;
    if (batch_size) *batch_size = ss.batch_size;
    if (batch_count) *batch_count = ss.batch_count;
    return 0;
}
int simb_oa_t_score(double alpha, uint64_t df, double* out) {
    if (!out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    int e = oa::t_score(alpha, df, out);
    return e ? fail(e, "t_score: unsupported alpha or df == 0") : 0;
}
uint64_t simb_usize_sqrt(uint64_t n) { return oa::usize_sqrt(n); }

int simb_describe(simb_sim* s, char* out, uint64_t capacity) {
    if (!s || !out) return fail(SIMB_ERR_INVALID_ARGUMENT, "null argument");
    const SimbNetDesc& net = s->hn.desc;
    char buf[1024];
    std::snprintf(buf, sizeof buf,
                  "{\"models\": %u, \"connectors\": %u, \"routes\": %u, \"f64_rows\": %u, \"u32_rows\": %u, "
                  "\"state_bytes_per_replication\": %u, \"max_pending\": %u, \"ring_u32_slots\": %u, \"ring_f64_slots\": %u, "
                  "\"tile\": %u, \"smem_bytes\": %u, \"steps_per_launch\": %u, \"stride\": %llu, \"replications\": %llu, "
                  "\"rng_kind\": %d, \"instrument\": %d, \"netdesc_bytes\": %u, \"kernel_shape\": \"%s\", \"grids_per_pass\": %d}",
                  net.n_models, net.n_connectors, net.n_routes, net.n_drows, net.n_wrows, net.n_drows * 8u + net.n_wrows * 4u,
                  net.max_pending, s->hn.ring_w_slots, s->hn.ring_d_slots, s->dims.tile, s->dims.smem_bytes, s->K,
                  (unsigned long long)s->st.stride, (unsigned long long)s->st.n, s->rng_kind, s->instrument ? 1 : 0,
                  (unsigned)sizeof(SimbNetDesc), engine_shape_name(s->dims.shape), grids_per_pass(s));
    put_str(out, capacity, buf);
    return 0;
}
int simb_model_count(simb_sim* s, uint64_t* models, uint64_t* connectors) {
    if (!s) return fail(SIMB_ERR_INVALID_ARGUMENT, "null handle");
    if (models) *models = s->hn.desc.n_models;
    if (connectors) *connectors = s->hn.desc.n_connectors;
    return 0;
}

}  // extern "C"
