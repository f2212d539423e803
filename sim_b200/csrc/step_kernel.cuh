// sim_b200/csrc/step_kernel.cuh -- the step kernel template (see engine.cu for the design notes).
// Included by engine.cu (generic interpreter instantiations) and by shape_kernel.cu (one translation
// unit per baked network shape, so the specialised kernels compile in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "core.cuh"

#ifndef SIMB_TABLE_QUAL
#define SIMB_TABLE_QUAL __device__ __align__(16) const
#endif
#include "zig_tables.inc"
#include "shapes.inc"
#define ZIG_BYTES 2064u  /* 258 doubles: 257 table entries + 1 pad, a multiple of 16 B for the bulk copy */

namespace simb {

typedef void (*StepKernel)(const SimbNetDesc, const SimbDevState, const SimbLaunch, const uint32_t);

// ---- PTX helpers: mbarrier + bulk async copy (TMA, non-tensor form) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// the same copies with an L2 eviction-priority hint (createpolicy): the state tiles of a streaming pass are
// marked evict_last so that the next pass -- which sweeps the batch in the opposite direction -- finds as much
// of the state as the 126 MB L2 can hold, while the sparsely touched ring slots take the normal priority
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared memory carve-up (all offsets multiples of 16 B): barrier, reduction scratch, ziggurat x tables, then
// the state tile (f64 rows followed by u32 rows).
struct SmemPlan {
    uint32_t off_bar, off_red, off_expx, off_normx, off_tab, off_vtab, off_rng, off_tiles, stage_bytes, d_bytes, total;
};
// generic interpreter: the model / route table without the connector column (instrumented handles read that one
// from global memory) and the value table are copied to shared memory at the start of a launch
#define SIMB_TAB_SMEM_WORDS SIMB_TAB_CONN
// `rng_slots`: Philox draws cached per replication in shared memory (baked-shape kernels; 0 = none)
// `zig`: the ziggurat x tables are staged in shared memory (else read through the L1: 4 KB less per CTA, which buys the
// 16-model networks a fifth resident CTA)
__host__ __device__ inline SmemPlan smem_plan(const SimbNetDesc& net, uint32_t tile, uint32_t rng_slots, bool tables,
                                              bool zig = true) {
    SmemPlan p;
    uint32_t o = 0;
    p.off_bar = o;  // three mbarriers: one per stage + one for the tables
    o += 32;
    p.off_red = o;
    o += 32;
    p.off_expx = o;
    if (net.uses_exp && zig) o += 2064;
    p.off_normx = o;
    if (net.uses_normal && zig) o += 2064;
    o = (o + 15u) & ~15u;
    p.off_tab = o;
    if (tables) o += SIMB_TAB_SMEM_WORDS * 4u;
    o = (o + 15u) & ~15u;
    p.off_vtab = o;
    if (tables) o += SIMB_VTAB_WORDS * 8u;
    o = (o + 127u) & ~127u;
    p.off_rng = o;
    o += rng_slots * 8u * tile;
    p.off_tiles = o;
    p.d_bytes = net.n_drows * tile * 8u;
    p.stage_bytes = (p.d_bytes + net.n_wrows * tile * 4u + 127u) & ~127u;
    o += p.stage_bytes;
    p.total = o;
    return p;
}

// Baked shapes need few registers, so more CTAs fit an SM.  Measured on M/M/1 (profiles/r01_pipeline_experiments.txt):
// fused launches (K = 128) are best at 5 CTAs / SM (48 registers); streaming launches (one or two steps per state
// load) want the most tile loads in flight: 6 CTAs / SM (40 registers) -> two instantiations, chosen per launch.
#ifndef SIMB_SHAPE_MIN_BLOCKS
#define SIMB_SHAPE_MIN_BLOCKS 5
#endif
#ifndef SIMB_SHAPE_STREAM_MIN_BLOCKS
#define SIMB_SHAPE_STREAM_MIN_BLOCKS 6
#endif
#define SIMB_STREAM_MAX_K 2u /* launches with k <= this use the streaming instantiation */
#ifndef SIMB_MIN_BLOCKS
#define SIMB_MIN_BLOCKS 4 /* <= 64 registers: 32 resident warps per SM measured +15% over 3 (24 warps) */
#endif
// SHAPE >= 0: the structure of the network is the baked constant kShapeDev<SHAPE> (shapes.inc), so the
// interpreter's model loops unroll and its dispatch folds; SHAPE < 0: generic interpreter over the
// launch's descriptor.  Either way the distribution parameters come from `net_param`.
template <int SHAPE>
__device__ __forceinline__ const SimbNetDesc& shape_desc(const SimbNetDesc& generic) {
#if SIMB_N_SHAPES > 0
    if (SHAPE == 0) return kShapeDev0;
#endif
#if SIMB_N_SHAPES > 1
    if (SHAPE == 1) return kShapeDev1;
#endif
#if SIMB_N_SHAPES > 2
    if (SHAPE == 2) return kShapeDev2;
#endif
#if SIMB_N_SHAPES > 3
    if (SHAPE == 3) return kShapeDev3;
#endif
    return generic;
}

template <int SHAPE>
__host__ __device__ constexpr int shape_tier() {  // model-count tier of a baked shape (0 = generic)
    constexpr int tiers[] = SIMB_SHAPE_TIERS;
    return SHAPE < 0 ? 0 : tiers[SHAPE < 0 ? 0 : SHAPE];
}

template <int SHAPE>
__host__ __device__ constexpr bool shape_instr() {  // baked shape lowered for an instrumented handle
    constexpr int flags[] = SIMB_SHAPE_INSTR;
    return SHAPE < 0 ? false : flags[SHAPE < 0 ? 0 : SHAPE] != 0;
}

// One CTA = one tile of blockDim.x consecutive replications.  Kernels of baked shapes are only launched
// with SIMB_SHAPE_TILE threads, so their shared-memory row offsets are immediates.
#define SIMB_SHAPE_TILE 256u
// VARIANT: baked shapes -- true = streaming launch (k <= SIMB_STREAM_MAX_K); generic interpreter -- true = narrow
// tile (<= 128 replications per CTA: networks whose state column is large), compiled for up to 96 registers: the
// CTA count of such networks is bounded by shared memory, so the wider register budget is free and removes the
// spills / rematerialisation of the 64-register build (+10-13 % on the 16- and 32-model networks).
template <int RNG, bool INSTR, int SHAPE, bool VARIANT = false, bool RARE = true, bool BASIC = false>
__global__ void __launch_bounds__((SHAPE < 0 && VARIANT) ? 128 : 256,
                                  SHAPE >= 0 ? (VARIANT ? SIMB_SHAPE_STREAM_MIN_BLOCKS : SIMB_SHAPE_MIN_BLOCKS)
                                             : (VARIANT ? 5 : SIMB_MIN_BLOCKS))
simb_step_kernel(const __grid_constant__ SimbNetDesc net_param, const __grid_constant__ SimbDevState st,
                 const __grid_constant__ SimbLaunch la, const uint32_t count_active) {
    const SimbNetDesc& net = shape_desc<SHAPE>(net_param);
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tile = SHAPE >= 0 ? SIMB_SHAPE_TILE : blockDim.x;
    const uint32_t tid = threadIdx.x;
    const uint32_t rng_slots = SHAPE >= 0 ? la.rng_slots : 0u;
    const bool zig_smem = SHAPE >= 0 || la.zig_global == 0u;
    const SmemPlan sp = smem_plan(net, tile, rng_slots, SHAPE < 0, zig_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sp.off_bar);
    unsigned long long* s_red = reinterpret_cast<unsigned long long*>(smem + sp.off_red);
    double* s_expx = reinterpret_cast<double*>(smem + sp.off_expx);
    double* s_normx = reinterpret_cast<double*>(smem + sp.off_normx);
    double* s_d = reinterpret_cast<double*>(smem + sp.off_tiles);
    uint32_t* s_w = reinterpret_cast<uint32_t*>(smem + sp.off_tiles + sp.d_bytes);
    const uint32_t n_rows = net.n_drows + net.n_wrows;
    const uint32_t tile_in_grid = la.reverse ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const uint64_t base = (uint64_t)(tile_in_grid + la.tile0) * tile;  // first replication (padded index) of this tile

    if (tid == 0) {
        mbar_init(bar, 1);
        s_red[0] = 0ull;
        s_red[1] = 0ull;
        s_red[2] = 0ull;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();  // make the barrier init visible to the async proxy
    }
    __syncthreads();
    // Programmatic dependent launch: let the NEXT launch of the stream start scheduling its CTAs now (they
    // become resident as this grid's CTAs retire, hiding launch latency and the tail), and do not touch
    // global memory before the PREVIOUS launch has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- 1. tile load: one bulk copy (TMA) per state row, spread over the lanes of warp 0 ----
    if (tid < 32) {
        if (tid == 0) {
            mbar_expect_tx(bar, net.n_drows * tile * 8u + net.n_wrows * tile * 4u +
                                    (net.uses_exp && zig_smem ? ZIG_BYTES : 0u) + (net.uses_normal && zig_smem ? ZIG_BYTES : 0u));
            // ziggurat x tables ride on the same barrier (random per-lane index -> shared memory, not constant)
            if (net.uses_exp && zig_smem) bulk_g2s(s_expx, ZIG_EXP_X, ZIG_BYTES, bar);
            if (net.uses_normal && zig_smem) bulk_g2s(s_normx, ZIG_NORM_X, ZIG_BYTES, bar);
        }
        __syncwarp();
        const uint64_t pol = l2_policy_evict_last();
#pragma unroll 1
        for (uint32_t r = tid; r < n_rows; r += 32) {
            if (r < net.n_drows) {
                if (la.l2_keep) bulk_g2s_hint(s_d + (size_t)r * tile, st.d + (uint64_t)r * st.stride + base, tile * 8u, bar, pol);
                else bulk_g2s(s_d + (size_t)r * tile, st.d + (uint64_t)r * st.stride + base, tile * 8u, bar);
            } else {
                const uint32_t rr = r - net.n_drows;
                if (la.l2_keep) bulk_g2s_hint(s_w + (size_t)rr * tile, st.w + (uint64_t)rr * st.stride + base, tile * 4u, bar, pol);
                else bulk_g2s(s_w + (size_t)rr * tile, st.w + (uint64_t)rr * st.stride + base, tile * 4u, bar);
            }
        }
    }
    Tables tab;
    tab.mt = reinterpret_cast<const uint32_t*>(smem + sp.off_tab);
    tab.vt = reinterpret_cast<const uint64_t*>(smem + sp.off_vtab);
    if (SHAPE < 0) {  // the tables are written at post / put, long before this launch
        uint32_t* s_mt = reinterpret_cast<uint32_t*>(smem + sp.off_tab);
        uint64_t* s_vt = reinterpret_cast<uint64_t*>(smem + sp.off_vtab);
        for (uint32_t i = tid; i < SIMB_TAB_SMEM_WORDS; i += tile) s_mt[i] = __ldg(st.tab + i);
        for (uint32_t i = tid; i < SIMB_VTAB_WORDS; i += tile) s_vt[i] = __ldg(st.vtab + i);
        __syncthreads();
    }
    tab.exp_x = zig_smem ? s_expx : ZIG_EXP_X;
    tab.exp_f = ZIG_EXP_F;
    tab.norm_x = zig_smem ? s_normx : ZIG_NORM_X;
    tab.norm_f = ZIG_NORM_F;
    tab.rng_cache = reinterpret_cast<uint64_t*>(smem + sp.off_rng);
    tab.rng_slots = rng_slots;
    Engine<RNG, shape_tier<SHAPE>(), RARE, BASIC> eng(net, net_param, st, tab);
    while (!mbar_try_wait(bar, 0)) {
    }

    // ---- 2. per-lane context ----
    Lane ln;
    ln.d = s_d + tid;
    ln.w = s_w + tid;
    ln.rc = tab.rng_cache + tid;
    ln.rc_stride = tile;
    ln.stride = tile;
    ln.rep = (uint32_t)(base + tid);
    const uint32_t ctl = eng.lane_load(ln, INSTR);
    const bool absent = CTL_ABSENT(ctl) != 0;

    // ---- 3. up to K lockstep steps (each warp runs on its own; no block barrier per step) ----
    bool done = absent || (la.mode == 1 && CTL_EPOCH(ctl) == la.epoch);
    bool stepped = false;
    if (SHAPE < 0) {
        // generic interpreter: the lanes advance through their own steps, one micro-iteration at a time (core.cuh)
        typename Engine<RNG, shape_tier<SHAPE>(), RARE, BASIC>::Prog pg;
        eng.prog_init(pg);
        uint32_t k_left = la.k;
        bool act = !done && ln.err == 0 && k_left > 0u;
        bool parked = false;  // lockstep mode: this lane has finished the step the others are still in
        while (__ballot_sync(0xFFFFFFFFu, act) != 0u) {  // until every replication of this warp has done its steps
            stepped = true;
            if (eng.template micro<INSTR>(ln, pg, act && !parked)) {
                k_left--;
                if (la.mode == 1 && !(ln.time < la.until)) done = true;  // step_until: mod.rs:281-285
                act = !done && ln.err == 0 && k_left > 0u;
                parked = la.lockstep != 0u;
            }
            if (la.lockstep && __ballot_sync(0xFFFFFFFFu, act && !parked) == 0u) parked = false;
        }
        // last launch of an API call: pay the deferred draws so the stored state is observable as is -- through the
        // loop's own flush site (one more pass of micro with no lane active)
        // -- and, in any launch, what a replication that has just failed still owes: the handlers that ran before the
        // failure drew in the reference, and dm_ext does not survive the store
        pg.flush = (la.settle || ln.err != 0u) && !absent && (ln.dm_int | ln.dm_ext) != 0u;
        if (__ballot_sync(0xFFFFFFFFu, pg.flush) != 0u) {
            stepped = true;
            eng.template micro<INSTR>(ln, pg, false);
        }
    } else {
        for (uint32_t k = 0; k < la.k; k++) {
            const bool live = !done && ln.err == 0;
            if (__ballot_sync(0xFFFFFFFFu, live) == 0u) break;  // every replication of this warp is finished
            stepped = true;
            eng.template step<INSTR>(ln, live);
            if (live && la.mode == 1 && !(ln.time < la.until)) done = true;  // step_until: mod.rs:281-285
        }
    }
    // last launch of an API call: pay the deferred draws so the stored state is observable as is
    if (SHAPE >= 0 && (la.settle || ln.err != 0u) && !absent && (ln.dm_int | ln.dm_ext)) {
        eng.flush_draws(ln);
        stepped = true;
    }

    // ---- 4. write back ----
    if (stepped) {
        eng.lane_store(ln, INSTR, (la.mode == 1 && done) ? la.epoch : 0u, absent ? 1u : 0u);
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk (async proxy) stores
    }
    {  // queue / table bytes this launch touched in global memory (the counted term of the algorithmic bytes)
        const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, ln.rb);
        if ((tid & 31u) == 0 && wsum) atomicAdd(&s_red[2], (unsigned long long)wsum);
    }
    if (count_active) {
        const uint32_t act = __popc(__ballot_sync(0xFFFFFFFFu, !done && ln.err == 0));
        const uint32_t bad = __popc(__ballot_sync(0xFFFFFFFFu, !absent && ln.err != 0));
        if ((tid & 31u) == 0) {
            if (act) atomicAdd(&s_red[0], (unsigned long long)act);
            if (bad) atomicAdd(&s_red[1], (unsigned long long)bad);
        }
    }
    const bool any_stepped = __syncthreads_or(stepped) != 0;
    if (any_stepped && tid < 32) {
        const uint64_t pol = l2_policy_evict_last();
#pragma unroll 1
        for (uint32_t r = tid; r < n_rows; r += 32) {
            if (r < net.n_drows) {
                if (la.l2_keep) bulk_s2g_hint(st.d + (uint64_t)r * st.stride + base, s_d + (size_t)r * tile, tile * 8u, pol);
                else bulk_s2g(st.d + (uint64_t)r * st.stride + base, s_d + (size_t)r * tile, tile * 8u);
            } else {
                const uint32_t rr = r - net.n_drows;
                if (la.l2_keep) bulk_s2g_hint(st.w + (uint64_t)rr * st.stride + base, s_w + (size_t)rr * tile, tile * 4u, pol);
                else bulk_s2g(st.w + (uint64_t)rr * st.stride + base, s_w + (size_t)rr * tile, tile * 4u);
            }
        }
        bulk_commit();
        bulk_wait_read_all();  // the stores must have read shared memory before the CTA exits
    }
    if (tid == 0 && s_red[2]) atomicAdd(&st.totals[4], s_red[2]);
    if (count_active && tid == 0) {
        if (s_red[0]) atomicAdd(&st.totals[2], s_red[0]);
        if (s_red[1]) atomicAdd(&st.totals[3], s_red[1]);
    }
}

}  // namespace simb
