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

#include "core.cuh"

#define SIMB_TABLE_QUAL __device__ __align__(16) const
#include "zig_tables.inc"
#define ZIG_BYTES 2064u  /* 258 doubles: 257 table entries + 1 pad, a multiple of 16 B for the bulk copy */

namespace simb {

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
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared memory carve-up (all offsets multiples of 16 B)
struct SmemPlan {
    uint32_t off_bar, off_d, off_w, off_expx, off_normx, off_red, total;
};
__host__ __device__ inline SmemPlan smem_plan(const SimbNetDesc& net, uint32_t tile) {
    SmemPlan p;
    uint32_t o = 0;
    p.off_bar = o;
    o += 16;
    p.off_d = o;
    o += net.n_drows * tile * 8u;
    p.off_w = o;
    o += net.n_wrows * tile * 4u;
    o = (o + 15u) & ~15u;
    p.off_expx = o;
    if (net.uses_exp) o += 2064;
    p.off_normx = o;
    if (net.uses_normal) o += 2064;
    p.off_red = o;
    o += 32;
    p.total = o;
    return p;
}

#ifndef SIMB_MIN_BLOCKS
#define SIMB_MIN_BLOCKS 4 /* <= 64 registers: 32 resident warps per SM measured +15% over 3 (24 warps) */
#endif
template <int RNG, bool INSTR>
__global__ void __launch_bounds__(256, SIMB_MIN_BLOCKS)
simb_step_kernel(const __grid_constant__ SimbNetDesc net, const __grid_constant__ SimbDevState st,
                 const __grid_constant__ SimbLaunch la, const uint32_t count_active) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tile = blockDim.x;
    const uint32_t tid = threadIdx.x;
    const SmemPlan sp = smem_plan(net, tile);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sp.off_bar);
    double* s_d = reinterpret_cast<double*>(smem + sp.off_d);
    uint32_t* s_w = reinterpret_cast<uint32_t*>(smem + sp.off_w);
    double* s_expx = reinterpret_cast<double*>(smem + sp.off_expx);
    double* s_normx = reinterpret_cast<double*>(smem + sp.off_normx);
    unsigned long long* s_red = reinterpret_cast<unsigned long long*>(smem + sp.off_red);

    const uint64_t base = (uint64_t)blockIdx.x * tile;  // first replication (padded index) of this tile
    const uint32_t n_rows = net.n_drows + net.n_wrows;

    if (tid == 0) {
        mbar_init(bar, 1);
        s_red[0] = 0ull;
        s_red[1] = 0ull;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();  // make the barrier init visible to the async proxy
    }
    __syncthreads();
    // ---- 1. tile load: one bulk copy per state row, spread over the lanes of warp 0 ----
    if (tid < 32) {
        if (tid == 0) {
            mbar_expect_tx(bar, net.n_drows * tile * 8u + net.n_wrows * tile * 4u +
                                    (net.uses_exp ? ZIG_BYTES : 0u) + (net.uses_normal ? ZIG_BYTES : 0u));
            // ziggurat x tables ride on the same barrier (random per-lane index -> shared memory, not constant)
            if (net.uses_exp) bulk_g2s(s_expx, ZIG_EXP_X, ZIG_BYTES, bar);
            if (net.uses_normal) bulk_g2s(s_normx, ZIG_NORM_X, ZIG_BYTES, bar);
        }
        __syncwarp();
        for (uint32_t r = tid; r < n_rows; r += 32) {
            if (r < net.n_drows) bulk_g2s(s_d + (size_t)r * tile, st.d + (uint64_t)r * st.stride + base, tile * 8u, bar);
            else {
                uint32_t rr = r - net.n_drows;
                bulk_g2s(s_w + (size_t)rr * tile, st.w + (uint64_t)rr * st.stride + base, tile * 4u, bar);
            }
        }
    }
    while (!mbar_try_wait(bar, 0)) {
    }

    // ---- 2. per-lane context ----
    Tables tab;
    tab.exp_x = s_expx;
    tab.exp_f = ZIG_EXP_F;
    tab.norm_x = s_normx;
    tab.norm_f = ZIG_NORM_F;
    Engine<RNG> eng(net, st, tab);
    Lane ln;
    ln.d = s_d + tid;
    ln.w = s_w + tid;
    ln.stride = tile;
    ln.rep_local = base + tid;
    ln.ring_w = st.ring_w + ln.rep_local;
    ln.ring_d = st.ring_d + ln.rep_local;
    ln.ring_stride = st.stride;
    const uint32_t ctl = eng.lane_load(ln, INSTR);
    const bool absent = ln.absent != 0;

    // ---- 3. up to K lockstep steps (each warp runs on its own; no block barrier per step) ----
    bool done = absent || (la.mode == 1 && CTL_EPOCH(ctl) == la.epoch);
    bool stepped = false;
    for (uint32_t k = 0; k < la.k; k++) {
        const bool live = !done && ln.err == 0;
        if (__ballot_sync(0xFFFFFFFFu, live) == 0u) break;  // every replication of this warp is finished
        stepped = true;
        eng.template step<INSTR>(ln, live);
        if (live && la.mode == 1 && !(ln.time < la.until)) done = true;  // step_until: mod.rs:281-285
    }
    // last launch of an API call: pay the deferred draws so the stored state is observable as is
    if (la.settle && !absent && (ln.dm_int | ln.dm_ext)) {
        eng.flush_draws(ln);
        stepped = true;
    }

    // ---- 4. write back ----
    if (stepped) {
        eng.lane_store(ln, INSTR, (la.mode == 1 && done) ? la.epoch : 0u);
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk (async proxy) stores
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
        for (uint32_t r = tid; r < n_rows; r += 32) {
            if (r < net.n_drows) bulk_s2g(st.d + (uint64_t)r * st.stride + base, s_d + (size_t)r * tile, tile * 8u);
            else {
                uint32_t rr = r - net.n_drows;
                bulk_s2g(st.w + (uint64_t)rr * st.stride + base, s_w + (size_t)rr * tile, tile * 4u);
            }
        }
        bulk_commit_wait_read();
    }
    if (count_active && tid == 0) {
        if (s_red[0]) atomicAdd(&st.totals[2], s_red[0]);
        if (s_red[1]) atomicAdd(&st.totals[3], s_red[1]);
    }
}

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
        if (keep_mail && k >= net.wrow_pend && k < net.wrow_pend + 2u * net.max_pending) continue;
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
    for (uint32_t k = 0; k < n_pre; k++) st.ring_w[(uint64_t)pre_slots[k] * st.stride + r] = pre_vals[k];
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
    st.w[(uint64_t)(net.wrow_pend + 2u * np) * st.stride + r] = content;
    st.w[(uint64_t)(net.wrow_pend + 2u * np + 1u) * st.stride + r] = route_word;
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
typedef void (*StepKernel)(const SimbNetDesc, const SimbDevState, const SimbLaunch, const uint32_t);
static StepKernel pick_kernel(int rng_kind, bool instrument) {
    if (instrument) {
        if (rng_kind == 0) return simb_step_kernel<0, true>;
        if (rng_kind == 1) return simb_step_kernel<1, true>;
        return simb_step_kernel<2, true>;
    }
    if (rng_kind == 0) return simb_step_kernel<0, false>;
    if (rng_kind == 1) return simb_step_kernel<1, false>;
    return simb_step_kernel<2, false>;
}

int engine_configure(const SimbNetDesc& net, uint64_t n, int rng_kind, bool instrument, EngineDims* dims) {
    // widest tile that leaves room for at least two CTAs per SM (227 KB usable shared memory)
    uint32_t tile = 256;
    while (tile > 32 && smem_plan(net, tile).total > 100u * 1024u) tile >>= 1;
    if (n && n < tile) {
        uint32_t t = 32;
        while (t < n) t <<= 1;
        if (t < tile) tile = t;
    }
    SmemPlan sp = smem_plan(net, tile);
    if (sp.total > 227u * 1024u) return (int)cudaErrorInvalidConfiguration;
    dims->tile = tile;
    dims->smem_bytes = sp.total;
    dims->grid = 0;
    cudaError_t e = cudaFuncSetAttribute((const void*)pick_kernel(rng_kind, instrument),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.total);
    return (int)e;
}

int engine_launch_step(const SimbNetDesc& net, const SimbDevState& st, const SimbLaunch& la, const EngineDims& dims,
                       int rng_kind, bool instrument, bool count_active, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)(st.stride / dims.tile);
    pick_kernel(rng_kind, instrument)<<<grid, dims.tile, dims.smem_bytes, stream>>>(net, st, la, count_active ? 1u : 0u);
    return (int)cudaGetLastError();
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
