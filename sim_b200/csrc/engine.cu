// sim_b200/csrc/engine.cu -- CUDA kernels of the batched-replication engine (sm_100a).
//
// Data layout: every per-replication state word is a ROW of an SoA matrix ([row][replication],
// replications contiguous), f64 rows and u32 rows in two arrays.  One CTA owns a TILE of `tile`
// consecutive replications.  A launch
//   1. pulls the tile (all rows x tile columns) into shared memory with 1-D bulk async copies
//      (TMA, cp.async.bulk ... mbarrier::complete_tx) issued by warp 0,
//   2. lets each thread run up to K Simulation::step()s of its replication out of shared memory
//      (core.cuh; shared-memory columns are bank-conflict free: lane i <-> column i),
//   3. pushes the tile back with bulk async stores.
// Queues (Processor FIFO etc.) stay in global memory as [slot][replication] rings: a step touches
// at most a couple of slots, so staging whole rings would multiply the traffic.
// The topology/parameter block (SimbNetDesc) is a __grid_constant__ kernel parameter, i.e. it is
// read through the constant bank with warp-uniform indices.
#include "engine.cuh"

#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "step_kernel.cuh"

namespace simb {

// ---- small maintenance kernels ----------------------------------------------------------------------
__global__ void simb_init_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                                 const double* __restrict__ d_init, const uint32_t* __restrict__ w_init,
                                 const uint32_t* __restrict__ pre_slots, const uint32_t* __restrict__ pre_vals,
                                 uint32_t n_pre, int rng_kind, int instrument, uint32_t keep_mask) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.stride) return;
    const bool absent = r >= st.n;
    const bool keep_clock = keep_mask & 1u, keep_mail = keep_mask & 2u, keep_rng = keep_mask & 4u;
    uint32_t old_ctl = (keep_mail || keep_rng) ? st.w[(uint64_t)net.wrow_ctl * st.stride + r] : 0u;
    for (uint32_t k = 0; k < net.n_drows; k++) {
        if (keep_clock && k == net.drow_time) continue;
        st.d[(uint64_t)k * st.stride + r] = d_init[k];
    }
    const uint32_t rng_rows = rng_kind == 1 ? 4u : 1u;
    for (uint32_t k = 0; k < net.n_wrows; k++) {
        if (keep_mail && k >= net.wrow_pend && k < net.wrow_pend + 2u * net.pend_tile) continue;
        if (keep_rng && k >= net.wrow_rng && k < net.wrow_rng + rng_rows) continue;
        st.w[(uint64_t)k * st.stride + r] = w_init[k];
    }
    st.w[(uint64_t)net.wrow_ctl * st.stride + r] = CTL_MAKE(0, 0, keep_mail ? CTL_NPEND(old_ctl) : 0u, absent ? 1u : 0u);
    if (!keep_rng) {
        if (rng_kind == 1) {  // Pcg64Mcg::new(seed + global replication index): state = seed | 1
            const uint64_t s = st.seed + st.rep_offset + r;
            st.w[(uint64_t)(net.wrow_rng + 0) * st.stride + r] = (uint32_t)s | 1u;
            st.w[(uint64_t)(net.wrow_rng + 1) * st.stride + r] = (uint32_t)(s >> 32);
            st.w[(uint64_t)(net.wrow_rng + 2) * st.stride + r] = 0u;
            st.w[(uint64_t)(net.wrow_rng + 3) * st.stride + r] = 0u;
        } else {
            st.w[(uint64_t)net.wrow_rng * st.stride + r] = 0u;
        }
    }
    if (instrument) {
        st.w[(uint64_t)net.wrow_hash * st.stride + r] = (uint32_t)TRACE_HASH_INIT;
        st.w[(uint64_t)(net.wrow_hash + 1) * st.stride + r] = (uint32_t)(TRACE_HASH_INIT >> 32);
    }
    for (uint32_t k = 0; k < n_pre; k++) {
        const uint32_t slot = pre_slots[k];
        if (slot & 0x80000000u)  // RING_PRELOAD_F64 (lower.hpp): one half of an f64 slot of ring_d
            reinterpret_cast<uint32_t*>(st.ring_d)[2u * ((uint64_t)(slot & 0x3FFFFFFFu) * st.stride + r) + ((slot >> 30) & 1u)] = pre_vals[k];
        else
            st.ring_w[(uint64_t)slot * st.stride + r] = pre_vals[k];
    }
}

__global__ void simb_reset_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                                  int messages, int clock) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.stride) return;
    if (clock) st.d[(uint64_t)net.drow_time * st.stride + r] = 0.0;
    if (messages) {
        uint32_t c = st.w[(uint64_t)net.wrow_ctl * st.stride + r];
        st.w[(uint64_t)net.wrow_ctl * st.stride + r] = CTL_MAKE(CTL_ERR(c), CTL_EPOCH(c), 0u, CTL_ABSENT(c));
    }
}

__global__ void simb_inject_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                                   uint32_t content, uint32_t route_word, const uint8_t* __restrict__ mask,
                                   unsigned long long* overflow) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.n) return;
    if (mask && !mask[r]) return;
    uint32_t c = st.w[(uint64_t)net.wrow_ctl * st.stride + r];
    uint32_t np = CTL_NPEND(c);
    if (np >= net.max_pending) {
        atomicAdd(overflow, 1ull);
        return;
    }
    route_word |= np << 24;  // sequence number: after everything already pending (the step kernel sorts by target)
    if (np < net.pend_tile) {
        st.w[(uint64_t)(net.wrow_pend + 2u * np) * st.stride + r] = content;
        st.w[(uint64_t)(net.wrow_pend + 2u * np + 1u) * st.stride + r] = route_word;
    } else {
        st.mail[(uint64_t)(2u * (np - net.pend_tile)) * st.stride + r] = content;
        st.mail[(uint64_t)(2u * (np - net.pend_tile) + 1u) * st.stride + r] = route_word;
    }
    st.w[(uint64_t)net.wrow_ctl * st.stride + r] = CTL_MAKE(CTL_ERR(c), CTL_EPOCH(c), np + 1u, CTL_ABSENT(c));
}

__global__ void simb_set_rng_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                                    int rng_kind) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.stride) return;
    if (rng_kind == 1) {
        const uint64_t s = st.seed + st.rep_offset + r;
        st.w[(uint64_t)(net.wrow_rng + 0) * st.stride + r] = (uint32_t)s | 1u;
        st.w[(uint64_t)(net.wrow_rng + 1) * st.stride + r] = (uint32_t)(s >> 32);
        st.w[(uint64_t)(net.wrow_rng + 2) * st.stride + r] = 0u;
        st.w[(uint64_t)(net.wrow_rng + 3) * st.stride + r] = 0u;
    } else {
        st.w[(uint64_t)net.wrow_rng * st.stride + r] = 0u;
    }
}

// deterministic two-stage reduction: fixed per-thread strided order, fixed tree inside the block;
// the host adds the per-block partials in block order.
#define RED_THREADS 256
#define RED_ITEMS 16
__global__ void __launch_bounds__(RED_THREADS)
simb_reduce_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st, int which,
                   double mean, double* __restrict__ psum, unsigned long long* __restrict__ pcnt) {
    __shared__ double sh_s[RED_THREADS];
    __shared__ unsigned long long sh_c[RED_THREADS];
    double s = 0.0;
    unsigned long long c = 0;
    const uint64_t base = (uint64_t)blockIdx.x * RED_THREADS * RED_ITEMS;
    for (int i = 0; i < RED_ITEMS; i++) {
        const uint64_t r = base + (uint64_t)i * RED_THREADS + threadIdx.x;
        if (r >= st.n) continue;
        const uint32_t n = st.w[(uint64_t)net.wrow_waitn * st.stride + r];
        if (n == 0) continue;  // replication mean undefined (NaN in the reference: filtered, web.rs:598)
        const double m = st.d[(uint64_t)net.drow_wait * st.stride + r] / (double)n;
        if (which == 0) s += m;
        else {
            const double dv = m - mean;
            s += dv * dv;
        }
        c += 1ull;
    }
    sh_s[threadIdx.x] = s;
    sh_c[threadIdx.x] = c;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh_s[threadIdx.x] += sh_s[threadIdx.x + off];
            sh_c[threadIdx.x] += sh_c[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        psum[blockIdx.x] = sh_s[0];
        pcnt[blockIdx.x] = sh_c[0];
    }
}
// SteadyStateOutput::{set_to_fixed_budget, calculate_batch_statistics, confidence_interval_mean}
// (sim/src/output_analysis/mod.rs:224-333) of every replication's captured series, one thread per replication;
// the series is [sample][replication], so a warp reads consecutive addresses.  Same operation order as
// output_analysis.cpp (the host restatement), hence bit-equal results.  The MSER-style statistic of the
// backward scan is evaluated twice (minimum over the first half, then the first index that attains it)
// instead of being stored.  t_lo[b] = t(alpha, b), t_hi[b] = t(alpha, b - 1): note the asymmetric degrees
// of freedom of the reference (:327 vs :330).
struct SteadyTScores {
    double t_lo[31], t_hi[31];
};
__global__ void __launch_bounds__(128)
simb_steady_state_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                         const __grid_constant__ SteadyTScores ts, SimbSteadyState* __restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.n) return;
    SimbSteadyState o;
    o.mean = o.variance = o.lower = o.upper = 0.0;
    o.deletion_point = o.batch_size = o.batch_count = 0;
    o.status = 0;
    const uint64_t n = st.w[(uint64_t)net.wrow_waitn * st.stride + r];
    const double* x = st.series + r;
    const uint64_t ld = st.stride;
    if (n > st.series_cap) o.status = 32;  // SIMB_ERR_CAPACITY: the series outgrew simb_config.series_capacity
    else if (n < 2) o.status = 34;         // `len() - 2` underflows (mod.rs:227): panic
    if (o.status) {
        out[r] = o;
        return;
    }
    const double nf = (double)n;
    const uint64_t half = (n - 1) / 2;
    double best = SIMB_INF;
    {
        double s = 0.0, q = 0.0;
        for (uint64_t d = n - 1; d-- > 0;) {
            const double v = x[(d + 1) * ld];
            s += v;
            q += v * v;
            const double w = nf - (double)d;
            const double m = q - (s * s) / (w * w * w);
            if (d < half) best = fmin(best, m);
        }
    }
    uint64_t del = n;
    {
        double s = 0.0, q = 0.0;
        for (uint64_t d = n - 1; d-- > 0;) {
            const double v = x[(d + 1) * ld];
            s += v;
            q += v * v;
            const double w = nf - (double)d;
            const double m = q - (s * s) / (w * w * w);
            if (m == best) del = d;  // scanning downwards: the lowest index wins = the first match of the source
        }
    }
    if (del == n) {
        o.status = 10;  // PrerequisiteCalcError
        out[r] = o;
        return;
    }
    uint64_t batches;
    {  // usize_sqrt (utils/mod.rs:84-92), at most 30 batches (mod.rs:253)
        const uint64_t a = n - del;
        uint64_t xx = a, y = 1;
        while (xx > y) {
            xx = (xx + y) / 2;
            y = a / xx;
        }
        batches = xx > 30 ? 30 : xx;
    }
    if (batches == 0) {
        o.status = 34;
        out[r] = o;
        return;
    }
    const uint64_t size = (n - del) / batches;
    del = n - batches * size;
    double means[30];
    double total = 0.0;
    for (uint64_t b = 0; b < batches; b++) {
        double t = 0.0;
        for (uint64_t i = del + size * b; i < del + size * (b + 1); i++) t = t + x[i * ld];
        means[b] = t / (double)size;
        total = total + means[b];
    }
    const double bm = total / (double)batches;
    double acc = 0.0;
    for (uint64_t b = 0; b < batches; b++) {
        const double dev = means[b] - bm;
        acc = acc + dev * dev;
    }
    const double bv = acc / (double)batches;
    o.deletion_point = (uint32_t)del;
    o.batch_size = (uint32_t)size;
    o.batch_count = (uint32_t)batches;
    o.mean = bm;
    o.variance = bv;
    if (batches == 1) {
        o.lower = o.upper = bm;
    } else {
        o.lower = bm - ts.t_lo[batches] * sqrt(bv) / sqrt((double)batches);
        o.upper = bm + ts.t_hi[batches] * sqrt(bv) / sqrt((double)batches);
    }
    out[r] = o;
}
// input_modeling on the device: every thread (= replication) draws `count` variates of ONE random variable
// (net.models[0], lowered by lower_random_variable) from its own Philox stream -- the streams the engine itself uses:
// key = (global replication index, seed), draws numbered from 0.  out[i * n + r]; draws[r] = u64s consumed.
// The samplers are the Engine's own (core.cuh), so a variate here is bit-identical to the one a model would draw.
__global__ void __launch_bounds__(128)
simb_sample_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st, uint32_t family,
                   uint64_t count, double* __restrict__ out_d, unsigned long long* __restrict__ out_u,
                   uint32_t* __restrict__ draws, int32_t* __restrict__ status) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.n) return;
    Tables tab;
    tab.exp_x = ZIG_EXP_X;  // (read from global memory here: this is not the hot path)
    tab.exp_f = ZIG_EXP_F;
    tab.norm_x = ZIG_NORM_X;
    tab.norm_f = ZIG_NORM_F;
    tab.mt = nullptr;
    tab.vt = nullptr;
    tab.rng_cache = nullptr;
    tab.rng_slots = 0;
    Engine<0, 0, true, false> eng(net, st, tab);
    Lane ln;
    memset(&ln, 0, sizeof ln);
    ln.rep = (uint32_t)r;
    const SimbModelDesc& md = net.models[0];
    const DescP pd{md};
    for (uint64_t i = 0; i < count && !ln.err; i++) {
        if (family == FAM_CONTINUOUS) out_d[i * st.n + r] = eng.sample_continuous_kind(ln, md.dist_kind, pd);
        else if (family == FAM_BOOLEAN) out_u[i * st.n + r] = eng.sample_bernoulli_p(ln, md.flags, pd) ? 1ull : 0ull;
        else if (family == FAM_DISCRETE) out_u[i * st.n + r] = eng.sample_discrete_kind(ln, md.dist_kind, pd);
        else if (md.flags & MF_INDEX_WEIGHTED) out_u[i * st.n + r] = eng.sample_index_p(ln, md.flags, md.n_w, md.wtab, pd);
        else if (pd.dist_err()) ln.err = ln.err ? ln.err : pd.dist_err();
        else out_u[i * st.n + r] = eng.sample_uniform_int(ln, pd.q0(), pd.q1(), pd.q2());  // Index::Uniform is a usize:
                                                                                          // all 64 bits (the models clamp
                                                                                          // it to a port count)
    }
    if (draws) draws[r] = ln.draw_idx;
    if (ln.err) atomicCAS(status, 0, (int32_t)ln.err);
}
// exact u64 sums of the steps / events rows
__global__ void __launch_bounds__(RED_THREADS)
simb_count_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                  unsigned long long* __restrict__ psteps, unsigned long long* __restrict__ pevents) {
    __shared__ unsigned long long sh_a[RED_THREADS];
    __shared__ unsigned long long sh_b[RED_THREADS];
    unsigned long long a = 0, b = 0;
    const uint64_t base = (uint64_t)blockIdx.x * RED_THREADS * RED_ITEMS;
    for (int i = 0; i < RED_ITEMS; i++) {
        const uint64_t r = base + (uint64_t)i * RED_THREADS + threadIdx.x;
        if (r >= st.n) continue;
        a += st.w[(uint64_t)net.wrow_steps * st.stride + r];
        b += st.w[(uint64_t)net.wrow_events * st.stride + r];
    }
    sh_a[threadIdx.x] = a;
    sh_b[threadIdx.x] = b;
    __syncthreads();
    for (int off = RED_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sh_a[threadIdx.x] += sh_a[threadIdx.x + off];
            sh_b[threadIdx.x] += sh_b[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        psteps[blockIdx.x] = sh_a[0];
        pevents[blockIdx.x] = sh_b[0];
    }
}

// ---- host side ------------------------------------------------------------------------------------------
// specialised kernels live in their own translation units (shape_kernel.cu, one per baked shape)
StepKernel simb_shape_kernel_0();
StepKernel simb_shape_kernel_1();
StepKernel simb_shape_kernel_2();
StepKernel simb_shape_kernel_3();
StepKernel simb_shape_stream_kernel_0();
StepKernel simb_shape_stream_kernel_1();
StepKernel simb_shape_stream_kernel_2();
StepKernel simb_shape_stream_kernel_3();

// generic interpreter with the rare continuous variables compiled in (engine_rare.cu)
StepKernel simb_generic_rare_kernel(int rng_kind, bool instrument, bool narrow);

// variant: baked shape -> streaming launch; generic -> narrow tile (see step_kernel.cuh)
static StepKernel pick_kernel(int rng_kind, bool instrument, int shape, bool stream = false, bool rare = true,
                              bool narrow = false, bool basic = false) {
    if (rng_kind == 0) {  // baked shapes exist for Philox streams only (engine_match_shape pairs the instrument flag)
        switch (shape) {
            case 0: return stream ? simb_shape_stream_kernel_0() : simb_shape_kernel_0();
            case 1: return stream ? simb_shape_stream_kernel_1() : simb_shape_kernel_1();
            case 2: return stream ? simb_shape_stream_kernel_2() : simb_shape_kernel_2();
            case 3: return stream ? simb_shape_stream_kernel_3() : simb_shape_kernel_3();
            default: break;
        }
    }
    if (basic && !rare && !instrument && rng_kind == 0)
        return narrow ? simb_step_kernel<0, false, -1, true, false, true> : simb_step_kernel<0, false, -1, false, false, true>;
    if (rare) return simb_generic_rare_kernel(rng_kind, instrument, narrow);
    if (narrow && !instrument && rng_kind == 0) return simb_step_kernel<0, false, -1, true, false>;
    if (instrument) {
        if (rng_kind == 0) return simb_step_kernel<0, true, -1, false, false>;
        if (rng_kind == 1) return simb_step_kernel<1, true, -1, false, false>;
        return simb_step_kernel<2, true, -1, false, false>;
    }
    if (rng_kind == 0) return simb_step_kernel<0, false, -1, false, false>;
    if (rng_kind == 1) return simb_step_kernel<1, false, -1, false, false>;
    return simb_step_kernel<2, false, -1, false, false>;
}
static bool only_basic_models(const SimbNetDesc& net) {
    for (uint32_t m = 0; m < net.n_models; m++) {
        const uint32_t t = net.models[m].type;
        if (t != MT_GENERATOR && t != MT_PROCESSOR && t != MT_STORAGE && t != MT_LOAD_BALANCER && t != MT_PASSIVE) return false;
    }
    return getenv("SIMB_NO_BASIC") == nullptr || !*getenv("SIMB_NO_BASIC");
}
static bool uses_rare_variables(const SimbNetDesc& net) {
    for (uint32_t m = 0; m < net.n_models; m++) {
        if (net.models[m].rng_row) return true;  // a model with its own generator: only the full interpreter draws from it
        const uint32_t k = net.models[m].dist_kind;
        if (k == DK_GAMMA || k == DK_LOGNORMAL || k == DK_WEIBULL || k == DK_BETA) return true;
    }
    // one Message delivered more than once (posted models sharing an id; the components of a Coupled model): the lean
    // builds compile the "not traced, not counted" rule of the later deliveries out
    for (uint32_t r = 0; r < net.n_routes; r++)
        if (net.routes[r].connector & (RT_CONT | RT_PARK)) return true;
    return false;
}

// structure = everything except the distribution values (p0..p2, q0..q2, dist_err, wcum)
static void strip_values(SimbNetDesc* d) {
    for (int i = 0; i < SIMB_MAX_MODELS; i++) {
        SimbModelDesc& m = d->models[i];
        m.p0 = m.p1 = m.p2 = 0.0;
        m.q0 = m.q1 = m.q2 = 0;
        m.dist_err = 0;
    }
    memset(d->wcum, 0, sizeof d->wcum);
}
int engine_match_shape(const SimbNetDesc& net, int rng_kind, bool instrument) {
    if (rng_kind != 0) return -1;
    const char* off = getenv("SIMB_NO_SHAPES");  // A/B switch: a non-empty value forces the generic interpreter
    if (off && *off) return -1;
    static_assert(SIMB_N_SHAPES <= 4, "extend shape_desc<> / pick_kernel / the Makefile for more baked shapes");
    SimbNetDesc a;
    memcpy(&a, &net, sizeof a);
    strip_values(&a);
    static const int shape_is_instr[] = SIMB_SHAPE_INSTR;
    for (int i = 0; i < SIMB_N_SHAPES; i++) {
        if ((shape_is_instr[i] != 0) != instrument) continue;
        SimbNetDesc b;
        memcpy(&b, kShapeHost[i], sizeof b);
        strip_values(&b);
        if (memcmp(&a, &b, sizeof a) == 0) return i;
    }
    return -1;
}
const char* engine_shape_name(int shape) { return (shape >= 0 && shape < SIMB_N_SHAPES) ? kShapeNames[shape] : "generic"; }

int engine_configure(const SimbNetDesc& net, uint64_t n, int rng_kind, bool instrument, EngineDims* dims) {
    dims->shape = engine_match_shape(net, rng_kind, instrument);
    // (Coupled models need the full interpreter: the lean builds compile their parked-message logic out)
    dims->rare = uses_rare_variables(net) || net.n_coupled != 0u;
    dims->basic = only_basic_models(net) && net.n_coupled == 0u && !dims->rare;
    // Tile width = replications per CTA.  Pick the width that keeps the most warps resident per SM
    // (shared memory 227 KB, 64 registers per thread -> at most 32 warps, at most 32 CTAs); among equals
    // prefer the wider tile (fewer, larger bulk copies), and keep half of the SM's unified storage for L1
    // when that costs no occupancy (the ring slots live in L1 between steps).
    uint32_t tile = 256, best_warps = 0;
    bool zig_global = false;
    for (int zg = 0; zg < 2; zg++) {  // second round: the ziggurat tables leave shared memory -- taken only if that keeps
                                      // MORE warps resident
        for (uint32_t cand = 256; cand >= 32; cand >>= 1) {
            const uint32_t smem = smem_plan(net, cand, 0, true, zg == 0).total;
            if (smem > 227u * 1024u) continue;
            uint32_t ctas = (227u * 1024u) / smem;
            const uint32_t regs = (cand <= 128u && !instrument && rng_kind == 0) ? 96u : 64u;  // narrow-tile variant
            const uint32_t by_regs = 65536u / (regs * cand), by_threads = 2048u / cand;
            if (ctas > by_regs) ctas = by_regs;
            if (ctas > by_threads) ctas = by_threads;
            if (ctas > 32u) ctas = 32u;
            const uint32_t warps = ctas * cand / 32u;
            // among equals prefer the wider tile (fewer, larger bulk copies) -- except that 128 beats 256: the narrow-tile
            // instantiation has 96 registers and does not spill (gateway network: 3.7e9 vs 2.8e9 replication-steps/s)
            if (warps > best_warps || (zg == (zig_global ? 1 : 0) && warps == best_warps && cand == 128u && tile == 256u)) {
                best_warps = warps;
                tile = cand;
                zig_global = zg == 1;
            }
        }
    }
    if (const char* zf = getenv("SIMB_ZIG_GLOBAL")) zig_global = *zf == '1';  // A/B switch
    if (best_warps == 0) return (int)cudaErrorInvalidConfiguration;
    if (dims->shape >= 0) tile = SIMB_SHAPE_TILE;  // a baked shape's kernel is compiled for the full tile width
    if (n && n < tile) {
        uint32_t t = 32;
        while (t < n) t <<= 1;
        if (t < tile) tile = t;
    }
    if (const char* force = getenv("SIMB_TILE")) {  // A/B switch: force the tile width (32..256, power of two)
        const uint32_t t = (uint32_t)atoi(force);
        if (t >= 32u && t <= 256u && (t & (t - 1u)) == 0u && smem_plan(net, t, 0, true).total <= 227u * 1024u) tile = t;
    }
    if (dims->shape >= 0 && tile != SIMB_SHAPE_TILE) dims->shape = -1;  // baked-shape kernels assume the full tile width
    // baked-shape kernels add their Philox draw cache: SIMB_RNG_CACHE draws per replication on fused launches
    if (dims->shape >= 0) zig_global = false;
    dims->zig_global = zig_global;
    SmemPlan sp = smem_plan(net, tile, dims->shape >= 0 ? SIMB_RNG_CACHE : 0u, dims->shape < 0, !zig_global);
    if (sp.total > 227u * 1024u) return (int)cudaErrorInvalidConfiguration;
    dims->tile = tile;
    dims->smem_bytes = sp.total;
    if (const char* pad = getenv("SIMB_SMEM_PAD")) {  // occupancy-sensitivity experiments: extra unused bytes per CTA
        if (*pad) dims->smem_bytes += (uint32_t)atoi(pad);
        if (dims->smem_bytes > 227u * 1024u) dims->smem_bytes = 227u * 1024u;
    }
    cudaError_t e = cudaFuncSetAttribute((const void*)pick_kernel(rng_kind, instrument, dims->shape, false, dims->rare, dims->tile <= 128u, dims->basic),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dims->smem_bytes);
    if (e != cudaSuccess) return (int)e;
    if (dims->shape >= 0) {
        e = cudaFuncSetAttribute((const void*)pick_kernel(rng_kind, instrument, dims->shape, true),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dims->smem_bytes);
        if (e != cudaSuccess) return (int)e;
    }
    dims->grid = 0;  // one CTA per tile (profiles/r01_pipeline_experiments.txt: persistent / double-buffered variants lost)
    return (int)cudaSuccess;
}

int engine_launch_step(const SimbNetDesc& net, const SimbDevState& st, const SimbLaunch& la_in, const EngineDims& dims,
                       int rng_kind, bool instrument, bool count_active, cudaStream_t stream, uint32_t tile0,
                       uint32_t n_tiles) {
    const uint32_t all = (uint32_t)(st.stride / dims.tile);
    const uint32_t grid = n_tiles ? n_tiles : all - tile0;
    SimbLaunch la = la_in;
    la.tile0 = tile0;
    const bool streaming = la.k <= SIMB_STREAM_MAX_K;
    la.rng_slots = dims.shape >= 0 ? (streaming ? 2u : SIMB_RNG_CACHE) : 0u;
    la.zig_global = dims.zig_global ? 1u : 0u;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(dims.tile);
    // streaming launches keep one Philox block per replication instead of SIMB_RNG_CACHE draws: less shared
    // memory per CTA, one more resident CTA per SM
    cfg.dynamicSmemBytes = dims.smem_bytes - (dims.shape >= 0 ? (SIMB_RNG_CACHE - la.rng_slots) * 8u * dims.tile : 0u);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // pairs with griddepcontrol.* in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint32_t ca = count_active ? 1u : 0u;
    return (int)cudaLaunchKernelEx(&cfg, pick_kernel(rng_kind, instrument, dims.shape, streaming, dims.rare, dims.tile <= 128u, dims.basic), net, st,
                                   la, ca);
}

static inline uint32_t blocks_for(uint64_t n, uint32_t t) { return (uint32_t)((n + t - 1) / t); }

int engine_launch_init(const SimbNetDesc& net, const SimbDevState& st, const double* d_init, const uint32_t* w_init,
                       const uint32_t* pre_slots, const uint32_t* pre_vals, uint32_t n_pre, int rng_kind, bool instrument,
                       uint32_t keep_mask, cudaStream_t stream) {
    simb_init_kernel<<<blocks_for(st.stride, 256), 256, 0, stream>>>(net, st, d_init, w_init, pre_slots, pre_vals, n_pre,
                                                                       rng_kind, instrument ? 1 : 0, keep_mask);
    return (int)cudaGetLastError();
}
int engine_launch_reset(const SimbNetDesc& net, const SimbDevState& st, bool messages, bool clock, cudaStream_t stream) {
    simb_reset_kernel<<<blocks_for(st.stride, 256), 256, 0, stream>>>(net, st, messages ? 1 : 0, clock ? 1 : 0);
    return (int)cudaGetLastError();
}
int engine_launch_inject(const SimbNetDesc& net, const SimbDevState& st, uint32_t content, uint32_t route_word,
                         const uint8_t* mask, unsigned long long* overflow, cudaStream_t stream) {
    simb_inject_kernel<<<blocks_for(st.n, 256), 256, 0, stream>>>(net, st, content, route_word, mask, overflow);
    return (int)cudaGetLastError();
}
int engine_launch_set_rng(const SimbNetDesc& net, const SimbDevState& st, int rng_kind, cudaStream_t stream) {
    simb_set_rng_kernel<<<blocks_for(st.stride, 256), 256, 0, stream>>>(net, st, rng_kind);
    return (int)cudaGetLastError();
}
int engine_launch_steady_state(const SimbNetDesc& net, const SimbDevState& st, const double* t_lo, const double* t_hi,
                               SimbSteadyState* out, cudaStream_t stream) {
    SteadyTScores ts;
    for (int i = 0; i < 31; i++) {
        ts.t_lo[i] = t_lo[i];
        ts.t_hi[i] = t_hi[i];
    }
    simb_steady_state_kernel<<<blocks_for(st.n, 128), 128, 0, stream>>>(net, st, ts, out);
    return (int)cudaGetLastError();
}
int engine_launch_sample(const SimbNetDesc& net, const SimbDevState& st, uint32_t family, uint64_t count, double* out_d,
                         unsigned long long* out_u, uint32_t* draws, int32_t* status, cudaStream_t stream) {
    simb_sample_kernel<<<blocks_for(st.n, 128), 128, 0, stream>>>(net, st, family, count, out_d, out_u, draws, status);
    return (int)cudaGetLastError();
}
uint32_t engine_reduce_blocks(uint64_t n) { return blocks_for(n, RED_THREADS * RED_ITEMS); }
int engine_launch_reduce(const SimbNetDesc& net, const SimbDevState& st, int which, double mean, double* psum,
                         unsigned long long* pcnt, uint32_t* n_blocks, cudaStream_t stream) {
    const uint32_t nb = engine_reduce_blocks(st.n);
    if (n_blocks) *n_blocks = nb;
    if (which == 2)
        simb_count_kernel<<<nb, RED_THREADS, 0, stream>>>(net, st, reinterpret_cast<unsigned long long*>(psum), pcnt);
    else
        simb_reduce_kernel<<<nb, RED_THREADS, 0, stream>>>(net, st, which, mean, psum, pcnt);
    return (int)cudaGetLastError();
}

}  // namespace simb
