// sim_b200/csrc/engine.cuh -- host-callable launch interface of the CUDA engine (engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "net.h"

namespace simb {

struct EngineDims {
    uint32_t tile;        // replications per CTA (= threads per CTA)
    uint32_t smem_bytes;  // dynamic shared memory per CTA
    uint32_t grid;        // reserved (the launch uses one CTA per tile)
    int shape;            // baked shape id whose specialised kernel is used, or -1 (generic interpreter)
    bool rare;            // the network draws from Beta / Gamma / LogNormal / Weibull (generic kernel variant)
    bool zig_global;      // ziggurat x tables read through the L1 (frees 2-4 KB of shared memory per CTA)
    bool basic;           // only Generator / Processor / Storage / LoadBalancer / Passive models (generic kernel variant)
};
int engine_match_shape(const SimbNetDesc& net, int rng_kind, bool instrument);
const char* engine_shape_name(int shape);

// chooses the tile width for a layout and sets the kernel attributes; returns cudaError_t
int engine_configure(const SimbNetDesc& net, uint64_t stride_hint, int rng_kind, bool instrument, EngineDims* dims);

// one step-kernel launch over tiles [tile0, tile0 + n_tiles) (n_tiles 0 = to the end): every live replication
// of those tiles executes up to la.k Simulation::step()s
int engine_launch_step(const SimbNetDesc& net, const SimbDevState& st, const SimbLaunch& la, const EngineDims& dims,
                       int rng_kind, bool instrument, bool count_active, cudaStream_t stream, uint32_t tile0 = 0,
                       uint32_t n_tiles = 0);

// state initialisation (Simulation::post / put): row constants, RNG state, absent padding lanes.
// keep_mask bit 0: keep clock, bit 1: keep mailbox (ctl.npend + pending rows), bit 2: keep RNG rows
int engine_launch_init(const SimbNetDesc& net, const SimbDevState& st, const double* d_init, const uint32_t* w_init,
                       const uint32_t* preload_slots, const uint32_t* preload_values, uint32_t n_preload, int rng_kind,
                       bool instrument, uint32_t keep_mask, cudaStream_t stream);

// Simulation::reset_messages / reset_global_time (mod.rs:131-144)
int engine_launch_reset(const SimbNetDesc& net, const SimbDevState& st, bool messages, bool clock, cudaStream_t stream);

// Simulation::inject_input (mod.rs:189-191) on every replication whose mask byte is non-zero (mask
// may be null = all).  *overflow (device counter) is incremented for replications whose mailbox is full.
int engine_launch_inject(const SimbNetDesc& net, const SimbDevState& st, uint32_t content, uint32_t route_word,
                         const uint8_t* mask, unsigned long long* overflow, cudaStream_t stream);

// re-seed (Simulation::set_rng): resets draw index / PCG state
int engine_launch_set_rng(const SimbNetDesc& net, const SimbDevState& st, int rng_kind, cudaStream_t stream);

// reductions: per-block partial sums written to `partials` (one double + one u64 count per block)
// which = 0: sum over replications of wait_sum/wait_n where wait_n > 0 (count = number of such)
// which = 1: sum of (wait_sum/wait_n - mean)^2 over the same set
// which = 2: sum of steps (as double) / count = sum of events -- u64 exact in `counts`
int engine_launch_reduce(const SimbNetDesc& net, const SimbDevState& st, int which, double mean, double* partial_sums,
                         unsigned long long* partial_counts, uint32_t* n_blocks, cudaStream_t stream);

// per-replication SteadyStateOutput of the captured series; t_lo[b] = t(alpha, b), t_hi[b] = t(alpha, b - 1), b <= 30
int engine_launch_steady_state(const SimbNetDesc& net, const SimbDevState& st, const double* t_lo, const double* t_hi,
                               SimbSteadyState* out, cudaStream_t stream);

// `count` variates per replication of the random variable lowered into net.models[0] (lower_random_variable), from the
// per-replication Philox streams (st.n, st.rep_offset, st.seed); out[i * st.n + r]
int engine_launch_sample(const SimbNetDesc& net, const SimbDevState& st, uint32_t family, uint64_t count, double* out_d,
                         unsigned long long* out_u, uint32_t* draws, int32_t* status, cudaStream_t stream);

uint32_t engine_reduce_blocks(uint64_t n);

}  // namespace simb
