// sim_b200/csrc/core.cuh -- the per-replication DEVS interpreter: one CUDA thread ("lane") owns
// one replication; all lanes walk the model list in the same order, so the model-type dispatch is
// warp-uniform and only the "is this model imminent / does it have mail" predicate diverges.
//
// Semantics follow sim/src/simulator/mod.rs:198-303 and sim/src/models/*.rs (cited per
// function; SURVEY.md Appendices A-C).  The code is written against a small `Lane` view so the
// SAME source also compiles as plain C++ for tests/hostsim (a unit-test harness that runs the
// interpreter logic on the CPU against the oracle; it is never linked into libsimb200.so).
//
// Arithmetic rules: f64 clock arithmetic is the reference's plain subtract/add (no FMA
// contraction: nvcc -fmad=false, g++ -ffp-contract=off) because scheduling tests `== 0.0`
// exactly (mod.rs:239).
#pragma once
#include "net.h"
#include "actions.h"

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SIMB_HD __host__ __device__ __forceinline__
#else
#define SIMB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SIMB_ON_DEVICE 1
#else
#define SIMB_ON_DEVICE 0
#endif

#if SIMB_ON_DEVICE
#define SIMB_LDG(p) __ldg(p)
#else
#define SIMB_LDG(p) (*(p))
#endif

namespace simb {

#define SIMB_INF (__builtin_huge_val())

// ---- warp helpers (single-lane semantics on the host) -------------------------------------------
SIMB_HD uint32_t warp_or(uint32_t v) {
#if SIMB_ON_DEVICE
    return __reduce_or_sync(0xFFFFFFFFu, v);
#else
    return v;
#endif
}
SIMB_HD bool warp_any(bool p) {
#if SIMB_ON_DEVICE
    return __ballot_sync(0xFFFFFFFFu, p) != 0u;
#else
    return p;
#endif
}
SIMB_HD int lowest_bit(uint32_t v) {
#if SIMB_ON_DEVICE
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
SIMB_HD double bits_f64(uint64_t b) {
#if SIMB_ON_DEVICE
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
SIMB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if SIMB_ON_DEVICE
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// ---- tables --------------------------------------------------------------------------------------
struct Tables {
    const double* exp_x;   // 257 entries (shared memory on the device)
    const double* exp_f;
    const double* norm_x;
    const double* norm_f;
    const uint32_t* mt;   // generic interpreter: model / route table (actions.h; shared memory on the device)
    const uint64_t* vt;   // generic interpreter: distribution value table
    uint64_t* rng_cache;  // baked shapes: Philox draw cache, [rng_slots][lane] (shared memory on the device)
    uint32_t rng_slots;   // prefetched draws per replication there (power of two >= 2)
};
// Optimisations of the baked small-shape kernels (A/B switches; profiles/r02_fused_experiments.txt)
#ifndef SIMB_OPT_NOVOTE
#define SIMB_OPT_NOVOTE 1 /* shapes of <= 4 models: no warp votes around the per-model blocks (some lane almost always needs each) */
#endif
#ifndef SIMB_OPT_UREG
#define SIMB_OPT_UREG 0   /* shapes of <= 4 models: until_next_event in registers for the whole launch (measured: -11 %,
                             the 5-CTAs-per-SM register budget spills) */
#endif
#ifndef SIMB_OPT_SQUEEZE
#define SIMB_OPT_SQUEEZE 0 /* ziggurat wedge: chord / tangent bounds before exp() (measured: -12 %, the fp64 divide and the
                              extra table reads cost more than the exponentials they save) */
#endif
#ifndef SIMB_RNG_CACHE
#define SIMB_RNG_CACHE 8u /* rng_slots of fused launches (streaming launches keep one block: 2) */
#endif

// ---- per-lane view of the replication's state tile ----------------------------------------------
// d/w point at row 0 of this lane's column; consecutive rows are `stride` elements apart
// (device: the shared-memory tile, stride = tile width; hostsim: the global arrays).
struct Lane {
    double* d;
    uint32_t* w;
    uint32_t stride;
    uint32_t rep;  // local replication index (column of the global arrays; rings are addressed through it)
    // registers
    double time;
    uint32_t err, npend, steps, events, draw_idx;
    uint32_t phase_lo, phase_hi;
    uint64_t hash;
    uint64_t pcg_lo, pcg_hi;
    // Philox prefetch, one of two forms (Engine::next_u64): the generic interpreter keeps the next three draws
    // in registers; baked shapes keep tab.rng_slots draws in a column of shared memory
    uint64_t c0, c1, c2;  // generic: c0 is draw number draw_idx
    uint64_t* rc;         // baked: draw number i sits in slot i % rng_slots of this column
    uint32_t rc_stride;   // elements between consecutive slots of that column
    uint32_t ccount;      // cached draws ahead of draw_idx
    uint32_t tmask;       // bit m: the mailbox holds a message for model m
    uint32_t dm_int, dm_ext;  // deferred continuous draws: "until_next_event[m] = sample(dist[m])" still
                              // owed from the previous internal phase / this external phase
    uint32_t traced;
    bool hold;       // the messages being sent come from a component of a Coupled model: they enter the trace when the
                     // whole model has fired (Engine::close_group), as the reference routes them then (mod.rs:241-263)
    double u[4];     // baked shapes of <= 4 models: until_next_event of model m, in registers for the whole launch
    uint64_t kseed;  // RARE builds: Philox key while a model draws from its OWN stream (`own`), see own_begin
    bool own;
    uint32_t rb;  // bytes of queue / table slots (and mailbox overflow) this lane touched in global memory in this
                  // launch: the counted part of the launch's algorithmic bytes (simb_counters.ring_bytes)
    // generic interpreter: models whose until_next_event is not +INF / is exactly 0.0 (kept in step with every
    // write of an until_next_event row, so the time advance and the imminent-model scan visit those only)
    uint32_t amask, zm;
    // ... and the smallest non-zero until_next_event among them, so that most time advances need no search:
    // it is folded forward whenever a value is written and by the subtraction pass itself; `dirty` (a running
    // timer was overwritten, or the state was just loaded) forces one search
    double nmin;
    bool dirty;

    SIMB_HD double& D(uint32_t row) { return d[(size_t)row * stride]; }
    SIMB_HD uint32_t& W(uint32_t row) { return w[(size_t)row * stride]; }
    SIMB_HD uint32_t phase(uint32_t m) const {
        return m < 16 ? (phase_lo >> (2 * m)) & 3u : (phase_hi >> (2 * (m - 16))) & 3u;
    }
    SIMB_HD void set_phase(uint32_t m, uint32_t p) {
        if (m < 16) phase_lo = (phase_lo & ~(3u << (2 * m))) | (p << (2 * m));
        else phase_hi = (phase_hi & ~(3u << (2 * (m - 16)))) | (p << (2 * (m - 16)));
    }
};

// ---- Philox4x32-10 (Salmon et al. SC'11); one block = two u64 draws -------------------------------
// key = the 64-bit seed; counter = (block index, 0, global replication index low, high): every (seed, replication)
// pair has its own stream, with no aliasing between seeds and replication indices.
SIMB_HD void philox_block(uint32_t key0, uint32_t key1, uint32_t ctr_lo, uint32_t rep_lo, uint32_t rep_hi, uint64_t* a,
                          uint64_t* b) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr_lo, c1 = 0u, c2 = rep_lo, c3 = rep_hi, k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    *a = (uint64_t)c0 | ((uint64_t)c1 << 32);
    *b = (uint64_t)c2 | ((uint64_t)c3 << 32);
}

struct U64Pair {
    uint64_t a, b;
};
// out-of-line copy for the rarely taken refill inside next_u64 (arguments and result by value, so the
// caller's Lane stays in registers)
#ifdef __CUDACC__
static __device__ __host__ __noinline__ U64Pair philox_pair_cold(uint32_t key0, uint32_t key1, uint32_t ctr_lo,
                                                                 uint32_t rep_lo, uint32_t rep_hi) {
#else
static inline U64Pair philox_pair_cold(uint32_t key0, uint32_t key1, uint32_t ctr_lo, uint32_t rep_lo, uint32_t rep_hi) {
#endif
    U64Pair r;
    philox_block(key0, key1, ctr_lo, rep_lo, rep_hi, &r.a, &r.b);
    return r;
}

// Literal-index repetition: X(0) X(1) ... (a `#pragma unroll` over the model loop is silently ignored
// once the handlers are inlined into it, and only literal indices let the compiler fold the reads of a
// baked network shape).
#define SIMB_REP4(X) X(0) X(1) X(2) X(3)
#define SIMB_REP8(X) SIMB_REP4(X) X(4) X(5) X(6) X(7)
#define SIMB_REP16(X) SIMB_REP8(X) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)
#define SIMB_REP32(X) SIMB_REP16(X) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)
// run X(m) for every literal m below the tier SM (4, 8, 16 or 32)
#define SIMB_FOR_MODELS(SM, X)                          \
    do {                                                \
        if (SM <= 4) { SIMB_REP4(X) }                   \
        else if (SM <= 8) { SIMB_REP8(X) }              \
        else if (SM <= 16) { SIMB_REP16(X) }            \
        else { SIMB_REP32(X) }                          \
    } while (0)

// Distribution parameters of one model, from the launch's descriptor (baked shapes: the model index is a literal,
// so the reads fold into constant-bank operands) ...
struct DescP {
    const SimbModelDesc& d;
    SIMB_HD double p0() const { return d.p0; }
    SIMB_HD double p1() const { return d.p1; }
    SIMB_HD double p2() const { return d.p2; }
    SIMB_HD uint64_t q0() const { return d.q0; }
    SIMB_HD uint64_t q1() const { return d.q1; }
    SIMB_HD uint64_t q2() const { return d.q2; }
    SIMB_HD uint32_t dist_err() const { return d.dist_err; }
};
// ... or from the value table (generic interpreter: lane-varying model index; a copy in shared memory)
struct TabP {
    const uint64_t* vt;
    uint32_t m, derr;
    SIMB_HD double p0() const { return bits_f64(vt[m]); }
    SIMB_HD double p1() const { return bits_f64(vt[SIMB_MAX_MODELS + m]); }
    SIMB_HD double p2() const { return bits_f64(vt[2 * SIMB_MAX_MODELS + m]); }
    SIMB_HD uint64_t q0() const { return vt[3 * SIMB_MAX_MODELS + m]; }
    SIMB_HD uint64_t q1() const { return vt[4 * SIMB_MAX_MODELS + m]; }
    SIMB_HD uint64_t q2() const { return vt[5 * SIMB_MAX_MODELS + m]; }
    SIMB_HD uint32_t dist_err() const { return derr; }
};

// RNG: 0 Philox, 1 Pcg64Mcg, 2 external stream.
// SM > 0: `net` refers to a compile-time-constant descriptor (a baked network shape, shapes.inc) of at
// most SM models: the model loops are expanded with literal indices so every structural read (types,
// rows, routes) folds to an immediate and the handlers of absent model types disappear; the
// distribution parameters still come from the launch's descriptor `par`.  SM == 0: generic interpreter.
// RARE: the less common continuous variables (Beta, Gamma, LogNormal, Weibull) are compiled in.  The generic
// kernel exists with and without them: they triple its code size, and it is bound by instruction fetch.
// BASIC: only the Generator / Processor / Storage / LoadBalancer / Passive handlers are compiled in (queueing
// networks such as M/M/1 and the tandem configuration); the generic kernel exists in this form too, again
// because its speed follows its code size.
template <int RNG, int SM = 0, bool RARE = true, bool BASIC = false>
struct Engine {
    static constexpr bool STATIC = SM > 0;
    const SimbNetDesc& net;  // structure
    const SimbNetDesc& par;  // values: p0..p2, q0..q2, dist_err, wcum
    const SimbDevState& st;
    Tables tab;

    SIMB_HD Engine(const SimbNetDesc& n, const SimbDevState& s, const Tables& t) : net(n), par(n), st(s), tab(t) {}
    SIMB_HD Engine(const SimbNetDesc& structure, const SimbNetDesc& values, const SimbDevState& s, const Tables& t)
        : net(structure), par(values), st(s), tab(t) {}

    // Philox key = the seed (warp-uniform), counter words 2 / 3 = the global replication index -- derived on demand,
    // not kept in registers
    SIMB_HD uint32_t key0(const Lane& ln) const { return (uint32_t)((RARE && !STATIC && ln.own) ? ln.kseed : st.seed); }
    SIMB_HD uint32_t key1(const Lane& ln) const { return (uint32_t)(((RARE && !STATIC && ln.own) ? ln.kseed : st.seed) >> 32); }
    // ---- a model with its own generator (`rng: Some(..)`, generator.rs:39,102-109; processor.rs, exclusive_gateway.rs,
    // stochastic_gate.rs alike): its draws come from the Philox stream (its seed, replication) whose position is one
    // word of the model's state.  own_begin / own_end bracket such a draw: the simulation's stream is set aside.
    struct SavedRng {
        uint32_t draw_idx, ccount;
        uint64_t c0, c1, c2;
    };
    SIMB_HD SavedRng own_begin(Lane& ln, uint32_t m, uint32_t rng_row) {
        SavedRng sv{ln.draw_idx, ln.ccount, ln.c0, ln.c1, ln.c2};
        ln.draw_idx = ln.W(rng_row);
        ln.ccount = 0;
        ln.kseed = tab.vt[6 * SIMB_MAX_MODELS + m];
        ln.own = true;
        return sv;
    }
    SIMB_HD void own_end(Lane& ln, uint32_t rng_row, const SavedRng& sv) {
        ln.W(rng_row) = ln.draw_idx;
        ln.draw_idx = sv.draw_idx;
        ln.ccount = sv.ccount;
        ln.c0 = sv.c0;
        ln.c1 = sv.c1;
        ln.c2 = sv.c2;
        ln.own = false;
    }
    SIMB_HD uint32_t rep_lo(const Lane& ln) const { return (uint32_t)(st.rep_offset + ln.rep); }
    SIMB_HD uint32_t rep_hi(const Lane& ln) const { return (uint32_t)((st.rep_offset + ln.rep) >> 32); }
    // Philox prefetch.  Draw number i is word pair (i & 1) of block i >> 1, so a cache of upcoming draws is a pure
    // prefetch of the stream: nothing of it is persisted, a launch starts with an empty cache.
    //  * generic interpreter: three draws in registers; a lane whose cache is down to one draw refills at the
    //    step's converged point (call cache_push only with ccount <= 1).
    //  * baked shapes: tab.rng_slots draws per lane in shared memory (slot i % rng_slots holds draw i), topped up at
    //    ONE warp-converged point: when any lane of the warp is down to its last draw, every lane with room for a
    //    block computes its next block, so the ten Philox rounds run once per ~3 warp iterations with most lanes
    //    active (per-lane refills ran them every iteration with a quarter of the lanes active: +17 % on M/M/1).
    SIMB_HD uint64_t& RC(Lane& ln, uint32_t i) { return ln.rc[(size_t)(i & (tab.rng_slots - 1u)) * ln.rc_stride]; }
    SIMB_HD void cache_push(Lane& ln, uint32_t nxt, uint64_t a, uint64_t b) {
        const bool odd = (nxt & 1u) != 0;  // odd position (only right after a state load): the block's second half
        if (STATIC) {
            if (odd) {
                RC(ln, nxt) = b;
            } else {
                RC(ln, nxt) = a;
                RC(ln, nxt + 1u) = b;
            }
        } else {
            if (odd) a = b;
            if (ln.ccount == 0) {
                ln.c0 = a;
                ln.c1 = b;
            } else {
                ln.c1 = a;
                ln.c2 = b;
            }
        }
        ln.ccount += odd ? 1u : 2u;
    }
    // the external stream (rng_kind 2) has run dry: the samplers' rejection loops stop instead of spinning on zeros.
    // (Not "the replication has failed": draws owed by handlers that ran before a failure are still made in full.)
    SIMB_HD bool dry(const Lane& ln) const { return RNG == 2 && ln.draw_idx > st.ext_len; }
    SIMB_HD void philox_refill(Lane& ln) {
        const uint32_t nxt = ln.draw_idx + ln.ccount;
        uint64_t a, b;
        philox_block(key0(ln), key1(ln), nxt >> 1, rep_lo(ln), rep_hi(ln), &a, &b);
        cache_push(ln, nxt, a, b);
    }
    SIMB_HD void philox_refill_cold(Lane& ln) {
        const uint32_t nxt = ln.draw_idx + ln.ccount;
        const U64Pair p = philox_pair_cold(key0(ln), key1(ln), nxt >> 1, rep_lo(ln), rep_hi(ln));
        cache_push(ln, nxt, p.a, p.b);
    }
    // the step's converged refill point (every lane of the warp calls it; `live` lanes take part)
    SIMB_HD void philox_topup(Lane& ln, bool live) {
        if (STATIC) {
            if (warp_any(live && ln.ccount <= 1u)) {
                if (live && ln.ccount + 2u <= tab.rng_slots) philox_refill(ln);
            }
        } else if (live && ln.ccount <= 1u) {
            philox_refill(ln);
        }
    }

    // until_next_event of model m.  Baked shapes of at most four models keep it in registers (every index is a literal
    // there, so ln.u[] is scalarised); everything else reads the state tile.
    static constexpr bool UREG = STATIC && SM <= 4 && SIMB_OPT_UREG;
    SIMB_HD double& U(Lane& ln, uint32_t m) { return UREG ? ln.u[m & 3u] : ln.D(net.drow_until + m); }
    // queue / table storage in global memory, [slot][replication]
    SIMB_HD uint32_t& RW(Lane& ln, uint32_t slot) {
        ln.rb += 4u;
        return st.ring_w[(uint64_t)slot * st.stride + ln.rep];
    }
    SIMB_HD double& RD(Lane& ln, uint32_t slot) {
        ln.rb += 8u;
        return st.ring_d[(uint64_t)slot * st.stride + ln.rep];
    }

    // mailbox slot j: word 0 = content, word 1 = route.  The first net.pend_tile slots are rows of the
    // state tile (shared memory); deeper slots -- reached only by rare bursts -- stay in global memory.
    SIMB_HD uint32_t& mail_word(Lane& ln, uint32_t j, uint32_t which) {
        // no overflow area when the whole mailbox fits the tile (folds away for baked shapes)
        if (net.pend_tile == net.max_pending || j < net.pend_tile) return ln.W(net.wrow_pend + 2u * j + which);
        ln.rb += 4u;
        return st.mail[(uint64_t)(2u * (j - net.pend_tile) + which) * st.stride + ln.rep];
    }

    // ---- uniform u64 source (RngCore::next_u64 on the shared DynRng, services.rs:25-27) ----------
    SIMB_HD uint64_t next_u64(Lane& ln) {
        if (RNG == 0) {
            if (ln.ccount == 0) philox_refill_cold(ln);  // rare: the step tops up at a warp-uniform point
            uint64_t v;
            if (STATIC) {
                v = RC(ln, ln.draw_idx);
            } else {
                v = ln.c0;
                ln.c0 = ln.c1;
                ln.c1 = ln.c2;
            }
            ln.ccount--;
            ln.draw_idx++;
            return v;
        } else if (RNG == 1) {
            // rand_pcg 0.3 Mcg128Xsl64: state *= MULT (mod 2^128); XSL-RR output
            const uint64_t ML = 0x4385DF649FCCF645ull, MH = 0x2360ED051FC65DA4ull;
            uint64_t lo = ln.pcg_lo * ML;
            uint64_t hi = mulhi64(ln.pcg_lo, ML) + ln.pcg_lo * MH + ln.pcg_hi * ML;
            ln.pcg_lo = lo;
            ln.pcg_hi = hi;
            ln.draw_idx++;
            uint32_t rot = (uint32_t)(hi >> 58);
            uint64_t xsl = hi ^ lo;
            return (xsl >> rot) | (xsl << ((64u - rot) & 63u));
        } else {
            uint32_t idx = ln.draw_idx++;
            if (idx >= st.ext_len) {
                if (!ln.err) ln.err = 37;  // SIMB_ERR_STREAM_EXHAUSTED
                return 0;
            }
            return st.ext[(uint64_t)idx * st.stride + ln.rep];
        }
    }

    // ---- rand 0.8 float conversions ---------------------------------------
    SIMB_HD double standard_f64(Lane& ln) {  // Standard: 53 bits -> [0,1)
        return (double)(next_u64(ln) >> 11) * (1.0 / 9007199254740992.0);
    }
    SIMB_HD double open01(Lane& ln) {  // Open01: (0,1)
        return bits_f64((next_u64(ln) >> 12) | 0x3FF0000000000000ull) - (1.0 - 2.220446049250313e-16 / 2.0);
    }
    // rand_distr 0.4 Exp1 (ziggurat, 256 layers)
    SIMB_HD double sample_exp1(Lane& ln) {
        for (;;) {
            uint64_t bits = next_u64(ln);
            uint32_t i = (uint32_t)bits & 0xFFu;
            double u = bits_f64((bits >> 12) | 0x3FF0000000000000ull) - (1.0 - 2.220446049250313e-16 / 2.0);
            double x = u * tab.exp_x[i];
            if (x < tab.exp_x[i + 1]) return x;
            if (i == 0) return 7.69711747013104972 - log(standard_f64(ln));
            double f1 = tab.exp_f[i + 1], f0 = tab.exp_f[i];
            const double y = f1 + (f0 - f1) * standard_f64(ln);
            if (SIMB_OPT_SQUEEZE) {
                // exp(-x) on [x_{i+1}, x_i] lies above its tangent at x_i and below its chord, so most wedge points are
                // decided without the exponential; the narrow band in between (and anything within 1e-9 of a bound, so
                // that rounding can never flip a decision) falls through to the exact test.  Same accept / reject
                // decisions, same draws.
                const double xi = tab.exp_x[i], xj = tab.exp_x[i + 1];
                const double tangent = f0 * (1.0 + (xi - x));
                const double chord = f0 + (f1 - f0) * ((xi - x) / (xi - xj));
                if (y < tangent * (1.0 - 1e-9)) return x;
                if (y > chord * (1.0 + 1e-9)) {
                    if (dry(ln)) return x;
                    continue;
                }
            }
            if (y < exp(-x)) return x;
            if (dry(ln)) return x;  // external stream ran dry: do not spin
        }
    }
    // rand_distr 0.4 StandardNormal (symmetric ziggurat + Marsaglia tail)
    SIMB_HD double sample_standard_normal(Lane& ln) {
        for (;;) {
            uint64_t bits = next_u64(ln);
            uint32_t i = (uint32_t)bits & 0xFFu;
            double u = bits_f64((bits >> 12) | 0x4000000000000000ull) - 3.0;
            double x = u * tab.norm_x[i];
            if (fabs(x) < tab.norm_x[i + 1]) return x;
            if (i == 0) {
                const double R = 3.6541528853610088;
                double tx = 1.0, ty = 0.0;
                while (-2.0 * ty < tx * tx) {
                    double x_ = open01(ln);
                    double y_ = open01(ln);
                    tx = log(x_) / R;
                    ty = log(y_);
                    if (dry(ln)) break;
                }
                return (u < 0.0) ? tx - R : R - tx;
            }
            double f1 = tab.norm_f[i + 1], f0 = tab.norm_f[i];
            if (f1 + (f0 - f1) * standard_f64(ln) < exp(-x * x / 2.0)) return x;
            if (dry(ln)) return x;
        }
    }
    template <class P>
    SIMB_HD double sample_continuous_rare(Lane& ln, uint32_t dist_kind, const P& pd) {
        switch (dist_kind) {
            case DK_GAMMA: {  // rand_distr 0.4 Gamma (Marsaglia-Tsang); shape == 1 is lowered to DK_EXP
                const double d = pd.p0(), cc = pd.p1(), scale = pd.p2();
                double u_small = 0.0;
                if (pd.q0()) u_small = open01(ln);  // GammaSmallShape draws its Open01 first
                double g = 0.0;
                for (;;) {
                    const double x = sample_standard_normal(ln);
                    const double v_cbrt = 1.0 + cc * x;
                    if (v_cbrt <= 0.0) {
                        if (dry(ln)) break;
                        continue;
                    }
                    const double v = v_cbrt * v_cbrt * v_cbrt;
                    const double u = open01(ln);
                    const double x_sqr = x * x;
                    if (u < 1.0 - 0.0331 * x_sqr * x_sqr || log(u) < 0.5 * x_sqr + d * (1.0 - v + log(v))) {
                        g = d * v * scale;
                        break;
                    }
                    if (dry(ln)) break;
                }
                return pd.q0() ? g * pow(u_small, bits_f64(pd.q1())) : g;
            }
            case DK_LOGNORMAL: {
                const double n = sample_standard_normal(ln);
                return exp(pd.p0() + pd.p1() * n);
            }
            case DK_WEIBULL: {  // OpenClosed01, then scale * (-ln x)^(1/shape)
                const double x = (double)((next_u64(ln) >> 11) + 1ull) * (1.0 / 9007199254740992.0);
                return pd.p0() * pow(-log(x), pd.p1());
            }
            case DK_BETA: {  // rand_distr 0.4.3 Beta (Cheng BB / BC); p0 = a, p1 = b, q0 = flags
                const double aa = pd.p0(), bb = pd.p1();
                const double LN4 = 1.3862943611198906, LN5 = 1.6094379124341003;  // ln 4, ln 5
                const double alpha = aa + bb;
                double w = 0.0;
                if (!(pd.q0() & 1u)) {  // BB
                    const double beta = sqrt((alpha - 2.0) / (2.0 * aa * bb - alpha));
                    const double gamma = aa + 1.0 / beta;
                    for (;;) {
                        const double u1 = open01(ln);
                        const double u2 = open01(ln);
                        const double v = beta * log(u1 / (1.0 - u1));
                        w = aa * exp(v);
                        const double z = u1 * u1 * u2;
                        const double r = gamma * v - LN4;
                        const double s_ = aa + r - w;
                        if (s_ + 1.0 + LN5 >= 5.0 * z) break;
                        const double t = log(z);
                        if (s_ >= t) break;
                        if (!(r + alpha * log(alpha / (bb + w)) < t)) break;
                        if (dry(ln)) break;
                    }
                } else {  // BC
                    const double beta = 1.0 / bb;
                    const double delta = 1.0 + aa - bb;
                    const double kappa1 = delta * (1.0 / 18.0 / 4.0 + 3.0 / 18.0 / 4.0 * bb) / (aa * beta - 14.0 / 18.0);
                    const double kappa2 = 0.25 + (0.5 + 0.25 / delta) * bb;
                    for (;;) {
                        double z;
                        const double u1 = open01(ln);
                        const double u2 = open01(ln);
                        if (dry(ln)) break;
                        if (u1 < 0.5) {
                            const double y = u1 * u2;
                            z = u1 * y;
                            if (0.25 * u2 + z - y >= kappa1) continue;
                        } else {
                            z = u1 * u1 * u2;
                            if (z <= 0.25) {
                                const double v = beta * log(u1 / (1.0 - u1));
                                w = aa * exp(v);
                                break;
                            }
                            if (z >= kappa2) continue;
                        }
                        const double v = beta * log(u1 / (1.0 - u1));
                        w = aa * exp(v);
                        if (!(alpha * (log(alpha / (bb + w)) + v) - LN4 < log(z))) break;
                    }
                }
                if (!(pd.q0() & 2u)) return (w == SIMB_INF) ? 1.0 : w / (bb + w);
                return bb / (bb + w);
            }
            default:
                return 0.0;
        }
    }
    // Continuous::random_variate (random_variable.rs:69-89); parameter errors surface per draw
    SIMB_HD double sample_continuous(Lane& ln, const SimbModelDesc& md, const SimbModelDesc& pd) {
        return sample_continuous_kind(ln, md.dist_kind, DescP{pd});
    }
    template <class P>
    SIMB_HD double sample_continuous_kind(Lane& ln, uint32_t dist_kind, const P& pd) {
        if (pd.dist_err()) {
            if (!ln.err) ln.err = pd.dist_err();
            return 0.0;
        }
        switch (dist_kind) {
            case DK_EXP:
                return sample_exp1(ln) * pd.p0();  // Exp1 * lambda_inverse
            case DK_NORMAL: {
                double n = sample_standard_normal(ln);
                return pd.p0() + pd.p1() * n;
            }
            case DK_TRIANGULAR: {
                double mn = pd.p0(), mx = pd.p1(), mode = pd.p2();
                double f = standard_f64(ln);
                double diff_mode_min = mode - mn;
                double range = mx - mn;
                double f_range = f * range;
                if (f_range < diff_mode_min) return mn + sqrt(f_range * diff_mode_min);
                return mx - sqrt((range - f_range) * (mx - mode));
            }
            case DK_UNIFORM: {  // p0 = low, p1 = scale
                double value1_2 = bits_f64((next_u64(ln) >> 12) | 0x3FF0000000000000ull);
                double value0_1 = value1_2 - 1.0;
                return value0_1 * pd.p1() + pd.p0();
            }
            default:
                return RARE ? sample_continuous_rare(ln, dist_kind, pd) : 0.0;
        }
    }
    // ---- Discrete::random_variate (random_variable.rs:108-115); rand_distr 0.4 Geometric / Poisson ------------------
    static SIMB_HD double powi_f64(double a, int b) {  // compiler-rt __powidf2 (f64::powi)
        const bool recip = b < 0;
        double r = 1.0;
        for (;;) {
            if (b & 1) r *= a;
            b /= 2;
            if (b == 0) break;
            a *= a;
        }
        return recip ? 1.0 / r : r;
    }
    static SIMB_HD double log_gamma(double x) {  // rand_distr 0.4 utils::log_gamma (6-term Lanczos)
        const double co[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                              -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
        const double tmp = x + 5.5;
        const double lg = (x + 0.5) * log(tmp) - tmp;
        double a = 1.000000000190015, denom = x;
        for (int i = 0; i < 6; i++) {
            denom = denom + 1.0;
            a = a + (co[i] / denom);
        }
        return lg + log(2.5066282746310005 * a / x);
    }
    template <class P>
    SIMB_HD uint64_t sample_discrete_kind(Lane& ln, uint32_t dist_kind, const P& pd) {
        if (pd.dist_err()) {
            if (!ln.err) ln.err = pd.dist_err();
            return 0;
        }
        if (dist_kind == DK_GEOMETRIC) {
            const double p = pd.p0(), pi = pd.p1();
            if (p >= 2.0 / 3.0) {  // the trivial algorithm
                uint64_t failures = 0;
                for (;;) {
                    const double u = standard_f64(ln);
                    if (u <= p || dry(ln)) break;
                    failures++;
                }
                return failures;
            }
            if (p == 0.0) return 0xFFFFFFFFFFFFFFFFull;
            const uint64_t k = pd.q0();
            uint64_t d = 0;  // D ~ Geo(pi) = Geo(p) / 2^k
            while (standard_f64(ln) < pi && !dry(ln)) d++;
            uint64_t m;  // M ~ Geo(p) % 2^k by rejection
            for (;;) {
                m = next_u64(ln) & ((1ull << k) - 1ull);
                const double p_reject = m <= 0x7FFFFFFFull ? powi_f64(1.0 - p, (int)m) : pow(1.0 - p, (double)m);
                const double u = standard_f64(ln);
                if (u < p_reject || dry(ln)) break;
            }
            return (d << k) + m;
        }
        if (dist_kind == DK_POISSON) {  // Poisson<f64>::sample, then `as u64`
            const double lambda = pd.p0();
            double result;
            if (lambda < 12.0) {  // Knuth
                const double exp_lambda = pd.p1();
                double prod = 1.0;
                result = 0.0;
                while (prod > exp_lambda && !dry(ln)) {
                    prod = prod * standard_f64(ln);
                    result = result + 1.0;
                }
                result = result - 1.0;
            } else {  // rejection from a Cauchy envelope (Numerical Recipes)
                const double log_lambda = pd.p2(), sqrt_2lambda = bits_f64(pd.q0()), magic_val = bits_f64(pd.q1());
                for (;;) {
                    double comp_dev;
                    for (;;) {
                        comp_dev = tan(3.14159265358979323846 * standard_f64(ln));
                        result = sqrt_2lambda * comp_dev + lambda;
                        if (result >= 0.0 || dry(ln)) break;
                    }
                    result = floor(result);
                    const double check = 0.9 * (1.0 + comp_dev * comp_dev) *
                                         exp(result * log_lambda - log_gamma(1.0 + result) - magic_val);
                    if (standard_f64(ln) <= check || dry(ln)) break;
                }
            }
            return result >= 18446744073709551615.0 ? 0xFFFFFFFFFFFFFFFFull : result > 0.0 ? (uint64_t)result : 0ull;
        }
        return sample_uniform_int(ln, pd.q0(), pd.q1(), pd.q2());  // DK_UNIFORM_INT
    }
    // rand 0.8 UniformInt<u64>::sample: widening multiply + zone rejection
    SIMB_HD uint64_t sample_uniform_int(Lane& ln, uint64_t low, uint64_t range, uint64_t zone) {
        for (;;) {
            uint64_t v = next_u64(ln);
            uint64_t hi = mulhi64(v, range), lo = v * range;
            if (lo <= zone) return low + hi;
            if (dry(ln)) return low;
        }
    }
    // Index::random_variate (random_variable.rs:122-130)
    SIMB_HD uint32_t sample_index(Lane& ln, const SimbModelDesc& md, const SimbModelDesc& pd) {
        return sample_index_p(ln, md.flags, md.n_w, md.wtab, DescP{pd});
    }
    template <class P>
    SIMB_HD uint32_t sample_index_p(Lane& ln, uint32_t flags, uint32_t n_w, uint32_t wtab, const P& pd) {
        if (pd.dist_err()) {
            if (!ln.err) ln.err = pd.dist_err();
            return 0;
        }
        uint64_t c = sample_uniform_int(ln, pd.q0(), pd.q1(), pd.q2());
        if (!(flags & MF_INDEX_WEIGHTED)) return (uint32_t)(c > 0xFFFFFFFFull ? 0xFFFFFFFFull : c);
        uint32_t idx = 0;  // first cumulative weight strictly greater than the chosen weight
        while (idx < n_w && par.wcum[wtab + idx] <= c) idx++;
        return idx;
    }
    // Boolean::random_variate (random_variable.rs:96-101)
    SIMB_HD bool sample_bernoulli(Lane& ln, const SimbModelDesc& md, const SimbModelDesc& pd) {
        return sample_bernoulli_p(ln, md.flags, DescP{pd});
    }
    template <class P>
    SIMB_HD bool sample_bernoulli_p(Lane& ln, uint32_t flags, const P& pd) {
        if (pd.dist_err()) {
            if (!ln.err) ln.err = pd.dist_err();
            return false;
        }
        if (flags & MF_BERNOULLI_ALWAYS) return true;  // p == 1.0 consumes no draw
        return next_u64(ln) < pd.q0();
    }

    // ---- deferred continuous draws ------------------------------------------------------------------
    // A handler that must set until_next_event[m] to a fresh variate only notes the debt; the step
    // pays all debts at ONE warp-converged point (after the external phase, before the time advance),
    // so the sampler code is executed once per step instead of once per drawing site.  The uniform
    // stream is consumed in exactly the reference's order: debts are paid oldest first (previous
    // internal phase in model order, then this external phase in model order) and any draw that must
    // happen inline (Bernoulli at arrival, gateway index) first pays the debts that precede it.
    SIMB_HD bool defer_draw(Lane& ln, uint32_t m, const SimbModelDesc& md, bool ext_phase) {
        (void)md;
        const uint32_t de = par.models[m].dist_err;
        if (de) {  // parameter errors surface at draw time (random_variable.rs:72-87)
            if (!ln.err) ln.err = de;
            return false;
        }
        if (ext_phase) ln.dm_ext |= 1u << m;
        else ln.dm_int |= 1u << m;
        return true;
    }
    SIMB_HD void flush_draws(Lane& ln) {
        if (STATIC) {
            // When every drawing model of the baked shape uses the same distribution family (folds to a
            // constant), ONE sampling site serves all debts: the model index only selects the parameters.
            uint32_t kind = DK_NONE;
            bool same = true;
#define SIMB_KIND_ONE(m)                                                                  \
    if ((m) < net.n_models && net.models[m].dist_kind != DK_NONE) {                       \
        if (kind == DK_NONE) kind = net.models[m].dist_kind;                              \
        else if (kind != net.models[m].dist_kind) same = false;                           \
    }
            SIMB_FOR_MODELS(SM, SIMB_KIND_ONE);
#undef SIMB_KIND_ONE
            if (same && kind != DK_NONE) {
                while (ln.dm_int | ln.dm_ext) {
                    uint32_t m;
                    if (ln.dm_int) {
                        m = (uint32_t)lowest_bit(ln.dm_int);
                        ln.dm_int &= ln.dm_int - 1u;
                    } else {
                        m = (uint32_t)lowest_bit(ln.dm_ext);
                        ln.dm_ext &= ln.dm_ext - 1u;
                    }
                    const double dv = sample_continuous_kind(ln, kind, DescP{par.models[m]});
                    if (UREG) {  // (m is a run-time value here: select the register)
                        if (m == 0u) ln.u[0] = dv;
                        else if (m == 1u) ln.u[1] = dv;
                        else if (m == 2u) ln.u[2] = dv;
                        else ln.u[3] = dv;
                    } else {
                        ln.D(net.drow_until + m) = dv;
                    }
                }
                return;
            }
            // otherwise: literal model list, one (foldable) sampling site per model that can draw
#define SIMB_FLUSH_INT(m)                                                                            \
    if ((m) < net.n_models && net.models[m].dist_kind != DK_NONE && (ln.dm_int & (1u << (m)))) {     \
        ln.dm_int &= ~(1u << (m));                                                                   \
        U(ln, (m)) = sample_continuous(ln, net.models[m], par.models[m]);                            \
    }
#define SIMB_FLUSH_EXT(m)                                                                            \
    if ((m) < net.n_models && net.models[m].dist_kind != DK_NONE && (ln.dm_ext & (1u << (m)))) {     \
        ln.dm_ext &= ~(1u << (m));                                                                   \
        U(ln, (m)) = sample_continuous(ln, net.models[m], par.models[m]);                            \
    }
            if (ln.dm_int) SIMB_FOR_MODELS(SM, SIMB_FLUSH_INT);
            if (ln.dm_ext) SIMB_FOR_MODELS(SM, SIMB_FLUSH_EXT);
#undef SIMB_FLUSH_INT
#undef SIMB_FLUSH_EXT
            return;
        }
        while (ln.dm_int | ln.dm_ext) {
            uint32_t m;
            if (ln.dm_int) {
                m = (uint32_t)lowest_bit(ln.dm_int);
                ln.dm_int &= ln.dm_int - 1u;
            } else {
                m = (uint32_t)lowest_bit(ln.dm_ext);
                ln.dm_ext &= ln.dm_ext - 1u;
            }
            // parameter errors were raised when the debt was noted (apply_u), so dist_err is 0 here
            const uint32_t rr = RARE ? TF(TF_RNG, m) : 0u;
            SavedRng sv{};
            if (RARE && rr) sv = own_begin(ln, m, rr);
            const double dv = sample_continuous_kind(ln, (TF(TF_WTAB, m) >> 16) & 0xFFu, TabP{tab.vt, m, 0u});
            if (RARE && rr) own_end(ln, rr, sv);
            ln.D(net.drow_until + m) = dv;
            if (dv == 0.0) ln.zm |= 1u << m;  // a variate of exactly zero makes the model imminent at once
            else ln.nmin = fmin(ln.nmin, dv);
        }
    }

    // ---- instrumentation ---------------------------------------------------------------------------
    template <bool INSTR>
    SIMB_HD void record(Lane& ln, uint32_t m, uint32_t action, uint32_t content, uint32_t aux = 0) {
        if (INSTR) {
            ln.hash = (ln.hash ^ TRACE_RECORD_WORD(m, action, aux, content)) * TRACE_HASH_PRIME;
#if SIMB_ON_DEVICE
            atomicAdd(&st.record_count[m * ACT_COUNT + action], 1ull);
#else
            st.record_count[m * ACT_COUNT + action] += 1ull;
#endif
            if (ln.traced) {
                uint32_t k = st.trace_count[ln.rep];
                if (k < st.trace_cap) {
                    SimbTraceRec r;
                    r.step = ln.steps + 1u;
                    r.meta = (1u << 31) | m;
                    r.content = content;
                    r.aux = action | (aux << 8);
                    r.time = ln.time;
                    st.trace[(uint64_t)ln.rep * st.trace_cap + k] = r;
                }
                st.trace_count[ln.rep] = k + 1u;
            }
        }
    }

    // route one outgoing (port, content) over every matching connector, in connector order
    // (get_message_target_ids/ports, mod.rs:155-182, 243-263)
    template <bool INSTR>
    SIMB_HD void emit(Lane& ln, const SimbModelDesc& md, uint32_t out_port, uint32_t content) {
        uint32_t rb = net.route_begin[md.out_base + out_port], re = net.route_begin[md.out_base + out_port + 1];
        for (uint32_t r = rb; r < re; r++) {
            if (ln.npend >= net.max_pending) {
                if (!ln.err) ln.err = 32;  // SIMB_ERR_CAPACITY: mailbox overflow
                return;
            }
            SimbRoute rt = net.routes[r];
            mail_word(ln, ln.npend, 0) = content;
            mail_word(ln, ln.npend, 1) = ROUTE_WORD(rt.tgt_model, rt.tgt_port, r, ln.npend);
            ln.npend++;
            if (rt.tgt_model < SIMB_MAX_MODELS) ln.tmask |= 1u << rt.tgt_model;
            if (INSTR) {
                ln.hash = (ln.hash ^ TRACE_MESSAGE_WORD(rt.connector, content)) * TRACE_HASH_PRIME;
#if SIMB_ON_DEVICE
                atomicAdd(&st.conn_count[rt.connector], 1ull);
#else
                st.conn_count[rt.connector] += 1ull;
#endif
                if (ln.traced) {
                    uint32_t k = st.trace_count[ln.rep];
                    if (k < st.trace_cap) {
                        SimbTraceRec t;
                        t.step = ln.steps + 1u;
                        t.meta = rt.connector;
                        t.content = content;
                        t.aux = 0;
                        t.time = ln.time;
                        st.trace[(uint64_t)ln.rep * st.trace_cap + k] = t;
                    }
                    st.trace_count[ln.rep] = k + 1u;
                }
            }
        }
    }

    // ---- statistics of the MF_STAT processor (output_analysis on the device) --------------------------
    // One waiting time (Processing Start - Arrival, the series of sim/tests/web.rs:571-597) in service-start
    // order: running sum + count for the IID mean, and -- when configured -- streaming batch means with a
    // fixed warm-up deletion and fixed batch size, at most 30 batches (a defined deviation from
    // SteadyStateOutput, which needs the whole series for its deletion point).
    SIMB_HD void wait_sample(Lane& ln, double wt) {
        const uint32_t j = ln.W(net.wrow_waitn);
        ln.W(net.wrow_waitn) = j + 1u;
        ln.D(net.drow_wait) += wt;
        if (j < st.series_cap) st.series[(uint64_t)j * st.stride + ln.rep] = wt;  // series capture (0 = off)
        if (net.bm_size != 0u && j >= net.bm_warmup) {
            const uint32_t k = j - net.bm_warmup;
            if (k / net.bm_size < SIMB_MAX_BATCHES) {
                double acc = ln.D(net.drow_bm) + wt;
                if ((k + 1u) % net.bm_size == 0u) {  // batch complete: fold its mean into the sums
                    const double mean = acc / (double)net.bm_size;
                    ln.D(net.drow_bm + 1) += mean;
                    ln.D(net.drow_bm + 2) += mean * mean;
                    acc = 0.0;
                }
                ln.D(net.drow_bm) = acc;
            }
        }
    }

    // ---- FIFO ring helpers: queue word = head(16) | len(16) ----------------------------------------
    SIMB_HD uint32_t ring_slot(const SimbModelDesc& md, uint32_t head, uint32_t i) {
        uint32_t s = head + i;
        return s >= md.ring_cap ? s - md.ring_cap : s;
    }
    SIMB_HD bool ring_push(Lane& ln, const SimbModelDesc& md, uint32_t content, uint32_t* slot_out) {
        uint32_t q = ln.W(md.wrow);
        uint32_t head = q >> 16, len = q & 0xFFFFu;
        if (len >= md.ring_cap) {
            if (!ln.err) ln.err = 32;  // SIMB_ERR_CAPACITY: the reference's Vec is unbounded here
            return false;
        }
        uint32_t s = ring_slot(md, head, len);
        RW(ln, md.ring_w0 + s) = content;
        ln.W(md.wrow) = (head << 16) | (len + 1u);
        *slot_out = s;
        return true;
    }
    SIMB_HD uint32_t ring_pop(Lane& ln, const SimbModelDesc& md, uint32_t* slot_out) {
        uint32_t q = ln.W(md.wrow);
        uint32_t head = q >> 16, len = q & 0xFFFFu;
        uint32_t c = RW(ln, md.ring_w0 + head);
        *slot_out = head;
        uint32_t nh = head + 1u;
        if (nh >= md.ring_cap) nh = 0;
        ln.W(md.wrow) = (nh << 16) | (len - 1u);
        return c;
    }

    // ---- per-lane context <-> state tile -----------------------------------------------------------
    // Before: ln.d / ln.w / ln.stride / ln.ring_* / ln.rep are set.  Returns the ctl word.
    SIMB_HD uint32_t lane_load(Lane& ln, bool instr) {
        ln.time = ln.D(net.drow_time);
        const uint32_t ctl = ln.W(net.wrow_ctl);
        ln.err = CTL_ERR(ctl);
        ln.npend = CTL_NPEND(ctl);
        ln.steps = ln.W(net.wrow_steps);
        ln.events = ln.W(net.wrow_events);
        ln.draw_idx = 0;
        ln.pcg_lo = ln.pcg_hi = 0;
        if (RNG == 1) {
            ln.pcg_lo = (uint64_t)ln.W(net.wrow_rng) | ((uint64_t)ln.W(net.wrow_rng + 1) << 32);
            ln.pcg_hi = (uint64_t)ln.W(net.wrow_rng + 2) | ((uint64_t)ln.W(net.wrow_rng + 3) << 32);
        } else {
            ln.draw_idx = ln.W(net.wrow_rng);
        }
        ln.phase_lo = ln.W(net.wrow_phase);
        ln.phase_hi = net.n_phase_words > 1 ? ln.W(net.wrow_phase + 1) : 0u;
        ln.hash = 0;
        if (instr) ln.hash = (uint64_t)ln.W(net.wrow_hash) | ((uint64_t)ln.W(net.wrow_hash + 1) << 32);
        ln.c0 = ln.c1 = ln.c2 = 0;
        ln.ccount = 0;
        ln.traced = instr && ln.rep < st.trace_reps;
        ln.hold = false;
        ln.rb = 0;
        ln.kseed = 0;
        ln.own = false;
        ln.dm_int = ln.W(net.wrow_dmask);
        ln.dm_ext = 0;
        ln.tmask = 0;
        ln.amask = ln.zm = 0;
        ln.nmin = SIMB_INF;
        ln.dirty = true;
        if (STATIC) {
            for (uint32_t j = 0; j < ln.npend; j++) {
                const uint32_t t = ROUTE_TGT(mail_word(ln, j, 1));
                if (t < SIMB_MAX_MODELS) ln.tmask |= 1u << t;
            }
            if (UREG) {
                ln.u[0] = net.n_models > 0u ? ln.D(net.drow_until + 0u) : 0.0;
                ln.u[1] = net.n_models > 1u ? ln.D(net.drow_until + 1u) : 0.0;
                ln.u[2] = net.n_models > 2u ? ln.D(net.drow_until + 2u) : 0.0;
                ln.u[3] = net.n_models > 3u ? ln.D(net.drow_until + 3u) : 0.0;
            }
        } else {
            for (uint32_t m = 0; m < net.n_models; m++) {
                const double u = ln.D(net.drow_until + m);
                if (u != SIMB_INF) ln.amask |= 1u << m;
                if (u == 0.0) ln.zm |= 1u << m;
            }
            ln.amask |= ln.dm_int;  // a model that owes a draw: its until_next_event row is stale
            ln.zm &= ~ln.dm_int;
            // the mailbox must be sorted by target model (inject_input appends; a baked-shape kernel does not sort)
            for (uint32_t i = 1; i < ln.npend; i++) {
                const uint32_t c = mail_word(ln, i, 0), rw = mail_word(ln, i, 1);
                uint32_t j = i;
                while (j > 0u) {
                    const uint32_t prev = mail_word(ln, j - 1u, 1);
                    if (order_key(ROUTE_TGT(prev)) <= order_key(ROUTE_TGT(rw))) break;
                    mail_word(ln, j, 0) = mail_word(ln, j - 1u, 0);
                    mail_word(ln, j, 1) = prev;
                    j--;
                }
                if (j != i) {
                    mail_word(ln, j, 0) = c;
                    mail_word(ln, j, 1) = rw;
                }
            }
        }
        return ctl;
    }
    SIMB_HD void lane_store(Lane& ln, bool instr, uint32_t done_epoch, uint32_t absent) {
        ln.D(net.drow_time) = ln.time;
        if (UREG) {
            if (net.n_models > 0u) ln.D(net.drow_until + 0u) = ln.u[0];
            if (net.n_models > 1u) ln.D(net.drow_until + 1u) = ln.u[1];
            if (net.n_models > 2u) ln.D(net.drow_until + 2u) = ln.u[2];
            if (net.n_models > 3u) ln.D(net.drow_until + 3u) = ln.u[3];
        }
        ln.W(net.wrow_ctl) = CTL_MAKE(ln.err, done_epoch, ln.npend, absent);
        ln.W(net.wrow_steps) = ln.steps;
        ln.W(net.wrow_events) = ln.events;
        if (RNG == 1) {
            ln.W(net.wrow_rng) = (uint32_t)ln.pcg_lo;
            ln.W(net.wrow_rng + 1) = (uint32_t)(ln.pcg_lo >> 32);
            ln.W(net.wrow_rng + 2) = (uint32_t)ln.pcg_hi;
            ln.W(net.wrow_rng + 3) = (uint32_t)(ln.pcg_hi >> 32);
        } else {
            ln.W(net.wrow_rng) = ln.draw_idx;
        }
        ln.W(net.wrow_phase) = ln.phase_lo;
        if (net.n_phase_words > 1) ln.W(net.wrow_phase + 1) = ln.phase_hi;
        if (instr) {
            ln.W(net.wrow_hash) = (uint32_t)ln.hash;
            ln.W(net.wrow_hash + 1) = (uint32_t)(ln.hash >> 32);
        }
        ln.W(net.wrow_dmask) = ln.dm_int;
    }

    // ================================================================================================
    // events_ext per model type
    // ================================================================================================
    template <bool INSTR>
    SIMB_HD void events_ext(Lane& ln, uint32_t m, const SimbModelDesc& md, uint32_t port, uint32_t content) {
        switch (md.type) {
            case MT_GENERATOR:  // generator.rs:161-167: no-op
            case MT_PASSIVE:    // tests/custom.rs:49-56
                break;
            case MT_PROCESSOR: {  // processor.rs:211-227
                if (port != 0) { ln.err = 7; break; }  // InvalidMessage
                uint32_t q = ln.W(md.wrow);
                uint32_t len = q & 0xFFFFu;
                bool empty = len == 0, full = len == md.cap;
                if (empty && full) { ln.err = 5; break; }  // InvalidModelState (capacity 0)
                if (full) {  // ignore_job :152-158
                    record<INSTR>(ln, m, ACT_DROP, content);
                    break;
                }
                uint32_t slot;
                if (!ring_push(ln, md, content, &slot)) break;
                if (md.flags & MF_STAT) RD(ln, md.ring_d + slot) = ln.time;
                if (empty) {  // activate :128-150
                    ln.set_phase(m, 1);
                    if (!defer_draw(ln, m, md, true)) break;  // until_next_event = service time draw
                    record<INSTR>(ln, m, ACT_ARRIVAL, content);
                    record<INSTR>(ln, m, ACT_PROCESSING_START, content);
                    if (md.flags & MF_STAT) wait_sample(ln, ln.time - ln.time);  // Processing Start - Arrival = 0 here
                } else {  // add_job :119-126
                    record<INSTR>(ln, m, ACT_ARRIVAL, content);
                }
                break;
            }
            case MT_STORAGE: {  // storage.rs:151-161
                if (port == 0) {  // hold_job :106-113
                    ln.W(md.wrow) = content;
                    record<INSTR>(ln, m, ACT_ARRIVAL, content);
                } else if (port == 1) {  // get_job :101-104
                    ln.set_phase(m, 1);
                    U(ln, m) = 0.0;
                } else {
                    ln.err = 7;
                }
                break;
            }
            case MT_LOAD_BALANCER: {  // pass_job load_balancer.rs:78-87 (any port)
                ln.set_phase(m, 1);
                U(ln, m) = 0.0;
                uint32_t slot;
                if (!ring_push(ln, md, content, &slot)) break;
                record<INSTR>(ln, m, ACT_ARRIVAL, content);
                break;
            }
            case MT_GATE: {  // gate.rs:180-195
                if (BASIC) break;
                if (port == 0) {
                    if (ln.phase(m) != 1) {  // pass_job :127-137
                        ln.set_phase(m, 2);
                        U(ln, m) = 0.0;
                        uint32_t slot;
                        if (!ring_push(ln, md, content, &slot)) break;
                    }
                    record<INSTR>(ln, m, ACT_ARRIVAL, content);  // also when Closed (drop_job :139-145)
                } else if (port == 1) {  // activate :107-115
                    ln.set_phase(m, 0);
                    U(ln, m) = SIMB_INF;
                    record<INSTR>(ln, m, ACT_ACTIVATION, content);
                } else if (port == 2) {  // deactivate :117-125
                    ln.set_phase(m, 1);
                    U(ln, m) = SIMB_INF;
                    record<INSTR>(ln, m, ACT_DEACTIVATION, content);
                } else {
                    ln.err = 7;
                }
                break;
            }
            case MT_EXCLUSIVE_GATEWAY: {  // pass_job exclusive_gateway.rs:95-108 (any port)
                if (BASIC) break;
                ln.set_phase(m, 1);
                U(ln, m) = 0.0;
                uint32_t slot;
                if (!ring_push(ln, md, content, &slot)) break;
                record<INSTR>(ln, m, ACT_ARRIVAL, content, port == SIMB_PORT_UNKNOWN ? 0u : port + 1u);
                break;
            }
            case MT_PARALLEL_GATEWAY: {  // parallel_gateway.rs:160-172, increment_collection :100-116
                if (BASIC) break;
                if (port == SIMB_PORT_UNKNOWN) { ln.err = 7; break; }
                uint32_t cnt = ln.W(md.wrow);
                uint32_t i = 0;
                for (; i < cnt; i++)
                    if (RW(ln, md.ring_w0 + i) == content) break;
                if (i < cnt) {
                    RW(ln, md.ring_w1 + i) += 1u;
                } else {
                    if (cnt >= md.ring_cap) { ln.err = 32; break; }
                    RW(ln, md.ring_w0 + cnt) = content;
                    RW(ln, md.ring_w1 + cnt) = 1u;
                    ln.W(md.wrow) = cnt + 1u;
                }
                record<INSTR>(ln, m, ACT_ARRIVAL, content, port + 1u);
                U(ln, m) = 0.0;
                break;
            }
            case MT_STOCHASTIC_GATE: {  // stochastic_gate.rs:163-172, receive_job :101-122
                if (BASIC) break;
                if (port != 0) { ln.err = 7; break; }
                U(ln, m) = 0.0;
                if (ln.dm_int | ln.dm_ext) flush_draws(ln);  // earlier draws of the stream come first
                bool pass = sample_bernoulli(ln, md, par.models[m]);  // Bernoulli draw AT ARRIVAL
                if (ln.err) break;
                uint32_t slot;
                if (!ring_push(ln, md, content, &slot)) break;
                RW(ln, md.ring_w1 + slot) = pass ? 1u : 0u;
                record<INSTR>(ln, m, ACT_ARRIVAL, content);
                break;
            }
            case MT_STOPWATCH: {  // stopwatch.rs:271-282
                if (BASIC) break;
                if (port == 2) {  // get_job :211-214
                    ln.set_phase(m, 1);
                    U(ln, m) = 0.0;
                    break;
                }
                if (port > 2) { ln.err = 7; break; }
                record<INSTR>(ln, m, port == 0 ? ACT_START : ACT_STOP, content);
                // matching_or_new_job :137-155 on a bounded table of INCOMPLETE jobs:
                // word0 = entries, word1 = best content, word2 = next insertion seq, word3 = best seq;
                // entry: c0 = name, c1 = seq<<2 | has_stop<<1 | has_start, d = the time present.
                uint32_t cnt = ln.W(md.wrow);
                uint32_t i = 0;
                for (; i < cnt; i++)
                    if (RW(ln, md.ring_w0 + i) == content) break;
                uint32_t mine = port == 0 ? 1u : 2u, other = port == 0 ? 2u : 1u;
                if (i == cnt) {
                    if (cnt >= md.ring_cap) { ln.err = 32; break; }
                    uint32_t seq = ln.W(md.wrow + 2);
                    ln.W(md.wrow + 2) = seq + 1u;
                    RW(ln, md.ring_w0 + cnt) = content;
                    RW(ln, md.ring_w1 + cnt) = (seq << 2) | mine;
                    RD(ln, md.ring_d + cnt) = ln.time;
                    ln.W(md.wrow) = cnt + 1u;
                    break;
                }
                uint32_t fl = RW(ln, md.ring_w1 + i);
                if (!(fl & other)) {  // same side again (or an entry injected with neither side): overwrite the time
                    RD(ln, md.ring_d + i) = ln.time;
                    if (!(fl & mine)) RW(ln, md.ring_w1 + i) = fl | mine;
                    break;
                }
                // both sides now known: duration = stop - start (some_duration :95-100)
                double t_other = RD(ln, md.ring_d + i);
                double dur = port == 0 ? t_other - ln.time : ln.time - t_other;
                uint32_t seq = fl >> 2;
                uint32_t best_c = ln.W(md.wrow + 1), best_seq = ln.W(md.wrow + 3);
                double best_d = ln.D(md.drow);
                bool better;
                if (best_c == SIMB_CONTENT_NONE) better = (md.flags & MF_METRIC_MAX) ? dur > -SIMB_INF : dur < SIMB_INF;
                else if (md.flags & MF_METRIC_MAX) better = dur > best_d || (dur == best_d && seq < best_seq);
                else better = dur < best_d || (dur == best_d && seq < best_seq);
                if (better) {
                    ln.W(md.wrow + 1) = content;
                    ln.W(md.wrow + 3) = seq;
                    ln.D(md.drow) = dur;
                }
                // drop the completed entry (swap with last; lookup is by name, order by seq)
                uint32_t last = cnt - 1u;
                if (i != last) {
                    RW(ln, md.ring_w0 + i) = RW(ln, md.ring_w0 + last);
                    RW(ln, md.ring_w1 + i) = RW(ln, md.ring_w1 + last);
                    RD(ln, md.ring_d + i) = RD(ln, md.ring_d + last);
                }
                ln.W(md.wrow) = last;
                break;
            }
            case MT_BATCHER: {  // batcher.rs:192-206 (any port)
                if (BASIC) break;
                uint32_t q = ln.W(md.wrow);
                uint32_t len = q & 0xFFFFu;
                bool room = (uint64_t)len + 1ull < (uint64_t)md.cap;
                uint32_t ph = ln.phase(m);
                if (room && ph == 2) { ln.err = 5; break; }  // (Release, true) => InvalidModelState
                if (room) {
                    if (ph == 0) {  // start_batch :103-112
                        ln.set_phase(m, 1);
                        U(ln, m) = par.models[m].p0;
                    }  // add_to_batch :93-101 leaves the timer alone
                } else {  // fill_batch :114-123
                    ln.set_phase(m, 2);
                    U(ln, m) = 0.0;
                }
                uint32_t slot;
                if (!ring_push(ln, md, content, &slot)) break;
                record<INSTR>(ln, m, ACT_ARRIVAL, content);
                break;
            }
            default:
                break;
        }
    }

    // ================================================================================================
    // events_int per model type
    // ================================================================================================
    // pop `n` jobs from the FIFO ring and send them out of `out_port`.  All records of one
    // events_int call precede its messages (the reference routes the returned Vec afterwards,
    // mod.rs:240-263), so the trace order is records first, then messages.
    template <bool INSTR>
    SIMB_HD void release_n(Lane& ln, uint32_t m, const SimbModelDesc& md, uint32_t n, uint32_t out_port, uint32_t aux) {
        if (INSTR) {
            uint32_t head = ln.W(md.wrow) >> 16;
            for (uint32_t i = 0; i < n; i++) record<INSTR>(ln, m, ACT_DEPARTURE, RW(ln, md.ring_w0 + ring_slot(md, head, i)), aux);
        }
        for (uint32_t i = 0; i < n && !ln.err; i++) {
            uint32_t slot;
            uint32_t c = ring_pop(ln, md, &slot);
            emit<INSTR>(ln, md, out_port, c);
        }
    }

    template <bool INSTR>
    SIMB_HD void events_int(Lane& ln, uint32_t m, const SimbModelDesc& md) {
        switch (md.type) {
            case MT_GENERATOR: {  // generator.rs:169-177
                if (!defer_draw(ln, m, md, false)) break;  // the draw comes first (:102-109, :129-136)
                if (ln.phase(m) == 1) {  // release_job :98-123
                    uint32_t last = ln.W(md.wrow) + 1u;
                    if (last >= SIMB_NUM_LITERAL) { ln.err = 32; break; }  // content code space exhausted
                    ln.W(md.wrow) = last;
                    uint32_t c = md.prefix | last;
                    record<INSTR>(ln, m, ACT_GENERATION, c);
                    emit<INSTR>(ln, md, 0, c);
                } else {  // initialize_generation :125-146
                    ln.set_phase(m, 1);
                    record<INSTR>(ln, m, ACT_INITIALIZATION, SIMB_CONTENT_NONE);
                }
                break;
            }
            case MT_PROCESSOR: {  // processor.rs:229-238
                uint32_t q = ln.W(md.wrow);
                uint32_t head = q >> 16, len = q & 0xFFFFu;
                if (ln.phase(m) == 0) {
                    if (len == 0) {  // passivate :192-196
                        U(ln, m) = SIMB_INF;
                        break;
                    }
                    ln.set_phase(m, 1);  // process_next :160-175
                    if (!defer_draw(ln, m, md, false)) break;
                    record<INSTR>(ln, m, ACT_PROCESSING_START, RW(ln, md.ring_w0 + head));
                    if (md.flags & MF_STAT) wait_sample(ln, ln.time - RD(ln, md.ring_d + head));
                    break;
                }
                if (len == 0) { ln.err = 34; break; }  // queue.remove(0) on an empty Vec panics
                uint32_t slot;
                uint32_t c = ring_pop(ln, md, &slot);  // release_job :177-190
                ln.set_phase(m, 0);
                U(ln, m) = 0.0;
                record<INSTR>(ln, m, ACT_DEPARTURE, c);
                emit<INSTR>(ln, md, 0, c);
                break;
            }
            case MT_STORAGE: {  // storage.rs:163-171
                U(ln, m) = SIMB_INF;
                if (ln.phase(m) == 1) {  // release_job :115-130
                    ln.set_phase(m, 0);
                    uint32_t c = ln.W(md.wrow);
                    record<INSTR>(ln, m, ACT_DEPARTURE, c);
                    if (c != SIMB_CONTENT_NONE) emit<INSTR>(ln, md, 0, c);
                }
                break;
            }
            case MT_LOAD_BALANCER: {  // load_balancer.rs:134-142
                uint32_t len = ln.W(md.wrow) & 0xFFFFu;
                if (len == 0) {  // passivate
                    ln.set_phase(m, 0);
                    U(ln, m) = SIMB_INF;
                    break;
                }
                if (md.n_out == 0) { ln.err = 34; break; }  // `% 0` panics
                U(ln, m) = 0.0;  // send_job :95-111 (pre-increment)
                uint32_t next = ln.W(md.wrow + 1) + 1u;
                if (next >= md.n_out) next = 0;
                ln.W(md.wrow + 1) = next;
                release_n<INSTR>(ln, m, md, 1, next, next + 1u);
                break;
            }
            case MT_GATE: {  // send_jobs gate.rs:149-165
                if (BASIC) break;
                ln.set_phase(m, 0);
                U(ln, m) = SIMB_INF;
                release_n<INSTR>(ln, m, md, ln.W(md.wrow) & 0xFFFFu, 0, 0);
                break;
            }
            case MT_EXCLUSIVE_GATEWAY: {  // exclusive_gateway.rs:153-161
                if (BASIC) break;
                U(ln, m) = SIMB_INF;
                if (ln.phase(m) == 1) {  // send_jobs :110-134: one draw per firing
                    ln.set_phase(m, 0);
                    if (ln.dm_int | ln.dm_ext) flush_draws(ln);  // earlier draws of the stream come first
                    uint32_t idx = sample_index(ln, md, par.models[m]);
                    if (ln.err) break;
                    if (idx >= md.n_out) { ln.err = 34; break; }  // flow_paths[idx] out of bounds panics
                    release_n<INSTR>(ln, m, md, ln.W(md.wrow) & 0xFFFFu, idx, idx + 1u);
                }
                break;
            }
            case MT_PARALLEL_GATEWAY: {  // parallel_gateway.rs:174-182
                if (BASIC) break;
                uint32_t cnt = ln.W(md.wrow);
                uint32_t i = 0;
                for (; i < cnt; i++)  // canonical choice: earliest-inserted full collection
                    if (RW(ln, md.ring_w1 + i) == md.n_in) break;
                if (i == cnt) {  // passivate
                    U(ln, m) = SIMB_INF;
                    break;
                }
                U(ln, m) = 0.0;  // send_job :118-143
                uint32_t c = RW(ln, md.ring_w0 + i);
                for (uint32_t j = i; j + 1u < cnt; j++) {  // remove, keeping insertion order
                    RW(ln, md.ring_w0 + j) = RW(ln, md.ring_w0 + j + 1u);
                    RW(ln, md.ring_w1 + j) = RW(ln, md.ring_w1 + j + 1u);
                }
                ln.W(md.wrow) = cnt - 1u;
                if (INSTR)
                    for (uint32_t p = 0; p < md.n_out; p++) record<INSTR>(ln, m, ACT_DEPARTURE, c, p + 1u);
                for (uint32_t p = 0; p < md.n_out && !ln.err; p++) emit<INSTR>(ln, md, p, c);
                break;
            }
            case MT_STOCHASTIC_GATE: {  // stochastic_gate.rs:174-183
                if (BASIC) break;
                uint32_t len = ln.W(md.wrow) & 0xFFFFu;
                if (len == 0) {
                    U(ln, m) = SIMB_INF;
                    break;
                }
                U(ln, m) = 0.0;
                uint32_t slot;
                uint32_t c = ring_pop(ln, md, &slot);
                if (RW(ln, md.ring_w1 + slot)) {  // pass_job :129-141
                    record<INSTR>(ln, m, ACT_PASS, c);
                    emit<INSTR>(ln, md, 0, c);
                } else {  // block_job :143-148
                    record<INSTR>(ln, m, ACT_BLOCK, c);
                }
                break;
            }
            case MT_STOPWATCH: {  // stopwatch.rs:284-293
                if (BASIC) break;
                U(ln, m) = SIMB_INF;
                if (ln.phase(m) == 1) {  // release_minimum / release_maximum :216-250
                    ln.set_phase(m, 0);
                    uint32_t c = ln.W(md.wrow + 1);
                    record<INSTR>(ln, m, (md.flags & MF_METRIC_MAX) ? ACT_MAXIMUM_FETCH : ACT_MINIMUM_FETCH, c);
                    if (c != SIMB_CONTENT_NONE) emit<INSTR>(ln, md, 0, c);
                }
                break;
            }
            case MT_BATCHER: {  // batcher.rs:208-221
                if (BASIC) break;
                uint32_t len = ln.W(md.wrow) & 0xFFFFu;
                bool le = len <= md.cap, ge2 = (uint64_t)len >= 2ull * (uint64_t)md.cap;
                if (le && ge2) { ln.err = 5; break; }
                if (le) {  // release_full_queue :125-141
                    ln.set_phase(m, 0);
                    U(ln, m) = SIMB_INF;
                    release_n<INSTR>(ln, m, md, len, 0, 0);
                } else if (ge2) {  // release_multiple :161-177
                    ln.set_phase(m, 2);
                    U(ln, m) = 0.0;
                    release_n<INSTR>(ln, m, md, md.cap, 0, 0);
                } else {  // release_partial_queue :143-159
                    ln.set_phase(m, 1);
                    U(ln, m) = par.models[m].p0;
                    release_n<INSTR>(ln, m, md, md.cap, 0, 0);
                }
                break;
            }
            default:  // MT_PASSIVE never becomes imminent
                break;
        }
    }

    // ================================================================================================
    // Generic interpreter: table-driven handlers (actions.h).  The lane's model index is data, the code
    // path is the same for every model type, so the lanes of a warp stay converged whatever models
    // their replications are serving.
    // ================================================================================================
    SIMB_HD uint32_t TF(uint32_t field, uint32_t m) const { return tab.mt[field * SIMB_MAX_MODELS + m]; }
    SIMB_HD uint32_t ext_lut(uint32_t key) const {
#if SIMB_ON_DEVICE
        return __ldg(&kExtLutDev.v[key]);
#else
        return kExtLutHost.v[key];
#endif
    }
    SIMB_HD uint32_t int_lut(uint32_t key) const {
#if SIMB_ON_DEVICE
        return __ldg(&kIntLutDev.v[key]);
#else
        return kIntLutHost.v[key];
#endif
    }
    struct Prog {
        uint32_t stage;   // 0 = at a step boundary, 1 = external phase, 2 = internal phase
        uint32_t np, ei;  // messages in the mailbox at the start of the step, delivered so far
        uint32_t zmk;     // imminent models not yet fired in this step
        uint32_t nev;
        bool flush;       // a handler that draws inline is waiting for the deferred draws to be paid
        // release in progress (events_int of a queueing model): jobs left | REL_* flags, model | absolute out-port << 8,
        // and the job content when it does not come from the model's ring
        uint32_t rel, rel_mp, rel_one;
        // Coupled models (RARE builds): coupled models that have opened (delivered their parked messages) in this
        // step, the one delivering now (+1), and the scan position in its parked list (component << 16 | index)
        uint32_t copen, cdel, cpos;
        // ... and the one whose components are firing (+1) with the mailbox fill when the first of them began
        uint32_t cgrp, cseq0;
    };
#define REL_RING 0x10000u
#define REL_PORTS 0x20000u
    SIMB_HD void prog_init(Prog& pg) {
        pg.stage = 0;
        pg.np = pg.ei = 0;
        pg.zmk = 0;
        pg.nev = 0;
        pg.flush = false;
        pg.rel = pg.rel_mp = pg.rel_one = 0;
        pg.copen = pg.cdel = pg.cpos = 0;
        pg.cgrp = pg.cseq0 = 0;
    }
    // ---- Coupled models (sim/src/models/coupled.rs; RARE builds only) ---------------------------------------------
    // order key of a mailbox slot: events_ext runs over the TOP-LEVEL models in order, so the components of a Coupled
    // model share their model's place and keep emission order among themselves
    SIMB_HD uint32_t order_key(uint32_t tgt) const {
        if (!RARE || STATIC || net.n_coupled == 0u) return tgt;
        return tgt < SIMB_MAX_MODELS ? (TF(TF_GROUP, tgt) & 0xFFu) : tgt;
    }
    // internal coupling: the job waits in the coupled model's parked list (coupled.rs:238-247)
    SIMB_HD void park(Lane& ln, uint32_t cpl, uint32_t tgt_port, uint32_t content) {
        const SimbCoupledDesc& cd = net.coupled[cpl];
        const uint32_t cnt = ln.W(cd.wrow);
        if (cnt >= cd.ring_cap) {
            if (!ln.err) ln.err = 32;  // SIMB_ERR_CAPACITY: the reference's Vec is unbounded
            return;
        }
        RW(ln, cd.ring_w0 + cnt) = content;
        RW(ln, cd.ring_w1 + cnt) = tgt_port;
        ln.W(cd.wrow) = cnt + 1u;
    }
    // until_next_event[m] := 0 / +INF / p0 / "a fresh variate" (deferred, see flush_draws), keeping the
    // active and imminent masks exact.  false: a distribution parameter error froze the replication.
    template <bool EXT>
    SIMB_HD bool apply_u(Lane& ln, uint32_t m, uint32_t ua) {
        const uint32_t bit = 1u << m;
        if ((ln.amask & ~ln.zm) & bit) ln.dirty = true;  // a running timer is overwritten (Batcher): nmin may be stale
        if (ua == UA_DRAW) {
            const uint32_t de = TF(TF_WTAB, m) >> 24;
            if (de) {  // parameter errors surface at draw time (random_variable.rs:72-87)
                if (!ln.err) ln.err = de;
                return false;
            }
            if (EXT) ln.dm_ext |= bit;
            else ln.dm_int |= bit;
            ln.amask |= bit;
            ln.zm &= ~bit;
            return true;
        }
        double v = 0.0;
        if (ua == UA_INF) v = SIMB_INF;
        if (ua == UA_P0) v = bits_f64(tab.vt[m]);
        ln.D(net.drow_until + m) = v;
        if (ua == UA_INF) ln.amask &= ~bit;
        else ln.amask |= bit;
        if (v == 0.0) {
            ln.zm |= bit;
        } else {
            ln.zm &= ~bit;
            ln.nmin = fmin(ln.nmin, v);  // (+INF leaves it alone)
        }
        return true;
    }

    // first i < cnt with ring_w[base + i] == key (cnt if none).  The table lives in global memory: four entries are
    // requested at once so that a short scan costs one memory round trip instead of one per entry.
    SIMB_HD uint32_t ring_find(Lane& ln, uint32_t base, uint32_t cnt, uint32_t key) {
        for (uint32_t b = 0; b < cnt; b += 4u) {
            const uint32_t v0 = RW(ln, base + b);
            const uint32_t v1 = b + 1u < cnt ? RW(ln, base + b + 1u) : ~key;
            const uint32_t v2 = b + 2u < cnt ? RW(ln, base + b + 2u) : ~key;
            const uint32_t v3 = b + 3u < cnt ? RW(ln, base + b + 3u) : ~key;
            if (v0 == key) return b;
            if (v1 == key) return b + 1u;
            if (v2 == key) return b + 2u;
            if (v3 == key) return b + 3u;
        }
        return cnt;
    }

    // the three models whose events_ext works on a table / draws at arrival
    template <bool INSTR>
    SIMB_HD void ext_special(Lane& ln, uint32_t m, uint32_t sp, uint32_t port, uint32_t content, uint32_t tf,
                             uint32_t rows) {
        const uint32_t wrow = rows & 0xFFFFu, flags = tf >> 24;
        const uint32_t rcap = TF(TF_RING_CAP, m), w0 = TF(TF_RING_W0, m), w1 = TF(TF_RING_W1, m);
        if (sp == XS_SGATE) {  // receive_job stochastic_gate.rs:101-122: Bernoulli draw AT ARRIVAL
            apply_u<true>(ln, m, UA_ZERO);  // (ext2 made sure that no earlier draw of the stream is still owed)
            const uint32_t rr = RARE ? TF(TF_RNG, m) : 0u;
            SavedRng sv{};
            if (RARE && rr) sv = own_begin(ln, m, rr);
            const bool pass = sample_bernoulli_p(ln, flags, TabP{tab.vt, m, TF(TF_WTAB, m) >> 24});
            if (RARE && rr) own_end(ln, rr, sv);
            if (ln.err) return;
            const uint32_t q = ln.W(wrow);
            const uint32_t head = q >> 16, len = q & 0xFFFFu;
            if (len >= rcap) {
                ln.err = 32;
                return;
            }
            uint32_t s = head + len;
            if (s >= rcap) s -= rcap;
            RW(ln, w0 + s) = content;
            RW(ln, w1 + s) = pass ? 1u : 0u;
            ln.W(wrow) = q + 1u;
            record<INSTR>(ln, m, ACT_ARRIVAL, content);
            return;
        }
        if (sp == XS_PG) {  // parallel_gateway.rs:160-172, increment_collection :100-116
            if (port == SIMB_PORT_UNKNOWN) {
                ln.err = 7;
                return;
            }
            const uint32_t cnt = ln.W(wrow);
            const uint32_t i = ring_find(ln, w0, cnt, content);
            if (i < cnt) {
                RW(ln, w1 + i) += 1u;
            } else {
                if (cnt >= rcap) {
                    ln.err = 32;
                    return;
                }
                RW(ln, w0 + cnt) = content;
                RW(ln, w1 + cnt) = 1u;
                ln.W(wrow) = cnt + 1u;
            }
            record<INSTR>(ln, m, ACT_ARRIVAL, content, port + 1u);
            apply_u<true>(ln, m, UA_ZERO);
            return;
        }
        // Stopwatch start / stop (stopwatch.rs:271-282): matching_or_new_job :137-155 on a bounded table of
        // INCOMPLETE jobs: word0 = entries, word1 = best content, word2 = next insertion seq, word3 = best seq;
        // entry: c0 = name, c1 = seq<<2 | has_stop<<1 | has_start, d = the time present.
        const uint32_t rd = TF(TF_RING_D, m), drow = rows >> 16;
        record<INSTR>(ln, m, port == 0 ? ACT_START : ACT_STOP, content);
        const uint32_t cnt = ln.W(wrow);
        const uint32_t i = ring_find(ln, w0, cnt, content);
        const uint32_t mine = port == 0 ? 1u : 2u, other = port == 0 ? 2u : 1u;
        if (i == cnt) {
            if (cnt >= rcap) {
                ln.err = 32;
                return;
            }
            const uint32_t seq = ln.W(wrow + 2);
            ln.W(wrow + 2) = seq + 1u;
            RW(ln, w0 + cnt) = content;
            RW(ln, w1 + cnt) = (seq << 2) | mine;
            RD(ln, rd + cnt) = ln.time;
            ln.W(wrow) = cnt + 1u;
            return;
        }
        const uint32_t fl = RW(ln, w1 + i);
        if (!(fl & other)) {  // same side again (or an entry injected with neither side): overwrite the time
            RD(ln, rd + i) = ln.time;
            if (!(fl & mine)) RW(ln, w1 + i) = fl | mine;
            return;
        }
        // both sides now known: duration = stop - start (some_duration :95-100)
        const double t_other = RD(ln, rd + i);
        const double dur = port == 0 ? t_other - ln.time : ln.time - t_other;
        const uint32_t seq = fl >> 2;
        const uint32_t best_c = ln.W(wrow + 1), best_seq = ln.W(wrow + 3);
        const double best_d = ln.D(drow);
        bool better;
        if (best_c == SIMB_CONTENT_NONE) better = (flags & MF_METRIC_MAX) ? dur > -SIMB_INF : dur < SIMB_INF;
        else if (flags & MF_METRIC_MAX) better = dur > best_d || (dur == best_d && seq < best_seq);
        else better = dur < best_d || (dur == best_d && seq < best_seq);
        if (better) {
            ln.W(wrow + 1) = content;
            ln.W(wrow + 3) = seq;
            ln.D(drow) = dur;
        }
        const uint32_t last = cnt - 1u;  // drop the completed entry (swap with last; lookup is by name, order by seq)
        if (i != last) {
            RW(ln, w0 + i) = RW(ln, w0 + last);
            RW(ln, w1 + i) = RW(ln, w1 + last);
            RD(ln, rd + i) = RD(ln, rd + last);
        }
        ln.W(wrow) = last;
    }

    // events_ext of model m (any type) for one message.  false: nothing was done because the handler draws inline
    // (StochasticGate) while earlier draws of the stream are still owed -- pay them (flush_draws) and call again.
    template <bool INSTR>
    SIMB_HD bool ext2(Lane& ln, uint32_t m, uint32_t port, uint32_t content) {
        const uint32_t tf = TF(TF_TYPE, m);
        const uint32_t type = tf & 0xFFu, flags = tf >> 24;
        const uint32_t rows = TF(TF_ROWS, m);
        const uint32_t wrow = rows & 0xFFFFu;
        const uint32_t q = ln.W(wrow);
        const uint32_t len = q & 0xFFFFu, head = q >> 16;
        const uint32_t cap = TF(TF_CAP, m);
        const uint32_t qs = (len == 0u ? XK_EMPTY : 0u) | (len == cap ? XK_FULL : 0u) |
                            ((uint64_t)len + 1ull >= (uint64_t)cap ? XK_NOROOM : 0u);
        const uint32_t a = ext_lut(EXT_KEY(type, port < 3u ? port : 3u, ln.phase(m), qs));
        if (a & 3u) {
            ln.err = (a & 3u) == 1u ? 7u : 5u;  // InvalidMessage / InvalidModelState
            return true;
        }
        // Owed draws are paid lowest model first, which is the order events_ext ran in -- except inside a Coupled
        // model, whose external input couplings may reach a later component before an earlier one (coupled.rs:166-184):
        // a component that comes after one still owing pays that debt first, so the stream is consumed in the
        // reference's order.
        if (RARE && net.n_coupled != 0u && (ln.dm_ext >> m) > 1u) return false;
        if (!BASIC && ((a >> XA_SPECIAL_SHIFT) & 3u)) {
            const uint32_t sp = (a >> XA_SPECIAL_SHIFT) & 3u;
            if (sp == XS_SGATE && (ln.dm_int | ln.dm_ext)) return false;
            ext_special<INSTR>(ln, m, sp, port, content, tf, rows);
            return true;
        }
        if (a & XA_PUSH) {
            const uint32_t rcap = TF(TF_RING_CAP, m);
            if (len >= rcap) {
                ln.err = 32;  // SIMB_ERR_CAPACITY: the reference's Vec is unbounded here
                return true;
            }
            uint32_t s = head + len;
            if (s >= rcap) s -= rcap;
            RW(ln, TF(TF_RING_W0, m) + s) = content;
            ln.W(wrow) = q + 1u;
            if (flags & MF_STAT) RD(ln, TF(TF_RING_D, m) + s) = ln.time;
        }
        if (a & XA_STORE) ln.W(wrow) = content;
        if (a & XA_SETPH) ln.set_phase(m, (a >> XA_PH_SHIFT) & 3u);
        const uint32_t ua = (a >> XA_U_SHIFT) & 7u;
        if (ua != UA_KEEP && !apply_u<true>(ln, m, ua)) return true;
        if (INSTR) {
            const uint32_t r1 = (a >> XA_REC1_SHIFT) & 0xFu, r2 = (a >> XA_REC2_SHIFT) & 0xFu;
            if (r1 != 0xFu) record<INSTR>(ln, m, r1, content, (a & XA_AUXPORT) ? (port == SIMB_PORT_UNKNOWN ? 0u : port + 1u) : 0u);
            if (r2 != 0xFu) record<INSTR>(ln, m, r2, content);
        }
        if ((a & XA_WAIT0) && (flags & MF_STAT)) wait_sample(ln, ln.time - ln.time);  // Processing Start - Arrival = 0 here
        return true;
    }

    // route one outgoing (out-port, content) over every matching connector, in connector order
    template <bool INSTR>
    SIMB_HD void emit2(Lane& ln, uint32_t out_port_abs, uint32_t content) {
        const uint32_t be = tab.mt[SIMB_TAB_ROUTE_BEGIN + out_port_abs];
        for (uint32_t r = be & 0xFFFFu; r < (be >> 16); r++) {
            const uint32_t tw = tab.mt[SIMB_TAB_ROUTES + r];
            if (RARE && (tw & RTW_PARK)) {  // (before the mailbox bound: a parked job is not a message)
                park(ln, RTW_CPL(tw), tw & 0xFFFFu, content);
                continue;
            }
            if (ln.npend >= net.max_pending) {
                if (!ln.err) ln.err = 32;  // SIMB_ERR_CAPACITY: mailbox overflow
                return;
            }
            const uint32_t rw = (tw & 0x00FFFFFFu) | (ln.npend << 24);
            // keep the mailbox sorted by target model (stable): the order events_ext runs in (mod.rs:203-221)
            uint32_t j = ln.npend;
            const uint32_t key = order_key(ROUTE_TGT(rw));
            while (j > 0u) {
                const uint32_t prev = mail_word(ln, j - 1u, 1);
                if (order_key(ROUTE_TGT(prev)) <= key) break;
                mail_word(ln, j, 0) = mail_word(ln, j - 1u, 0);
                mail_word(ln, j, 1) = prev;
                j--;
            }
            mail_word(ln, j, 0) = content;
            mail_word(ln, j, 1) = rw;
            ln.npend++;
            // one Message of the reference = one trace entry
            if (INSTR && !(RARE && (tw & RTW_CONT)) && !(RARE && ln.hold)) trace_message(ln, r, content);
        }
    }
    // the trace entry / hash contribution / connector count of one Message sent over route r
    SIMB_HD void trace_message(Lane& ln, uint32_t r, uint32_t content) {
        const uint32_t conn = SIMB_LDG(st.tab + SIMB_TAB_CONN + r);
        ln.hash = (ln.hash ^ TRACE_MESSAGE_WORD(conn, content)) * TRACE_HASH_PRIME;
#if SIMB_ON_DEVICE
        atomicAdd(&st.conn_count[conn], 1ull);
#else
        st.conn_count[conn] += 1ull;
#endif
        if (ln.traced) {
            uint32_t k = st.trace_count[ln.rep];
            if (k < st.trace_cap) {
                SimbTraceRec t;
                t.step = ln.steps + 1u;
                t.meta = conn;
                t.content = content;
                t.aux = 0;
                t.time = ln.time;
                st.trace[(uint64_t)ln.rep * st.trace_cap + k] = t;
            }
            st.trace_count[ln.rep] = k + 1u;
        }
    }
    // A Coupled model has fired -- every imminent component has run events_int (coupled.rs:234-255) -- and only now
    // does the reference's step() see the messages it returned and route them (mod.rs:241-263): the trace entries of
    // what its components sent since `cseq0`, in sending order.  A failed component (`?` inside the Coupled model)
    // loses them all.
    template <bool INSTR>
    SIMB_HD void close_group(Lane& ln, Prog& pg) {
        if (INSTR && !ln.err) {
            for (uint32_t q = pg.cseq0; q < ln.npend; q++)
                for (uint32_t j = 0; j < ln.npend; j++) {
                    const uint32_t rw = mail_word(ln, j, 1);
                    if (ROUTE_SEQ(rw) != q) continue;
                    const uint32_t r = ROUTE_R(rw);
                    if (!(tab.mt[SIMB_TAB_ROUTES + r] & RTW_CONT)) trace_message(ln, r, mail_word(ln, j, 0));
                    break;
                }
        }
        pg.cgrp = 0;
        ln.hold = false;
    }

    // events_int of model m (any type).  false: nothing was done because the handler draws inline
    // (ExclusiveGateway) while earlier draws of the stream are still owed -- pay them and call again.
    template <bool INSTR>
    SIMB_HD bool int2(Lane& ln, Prog& pg, uint32_t m) {
        const uint32_t tf = TF(TF_TYPE, m);
        const uint32_t type = tf & 0xFFu, n_out = (tf >> 16) & 0xFFu, flags = tf >> 24;
        const uint32_t rows = TF(TF_ROWS, m);
        const uint32_t wrow = rows & 0xFFFFu;
        const uint32_t q = ln.W(wrow);
        const uint32_t len = q & 0xFFFFu, head = q >> 16;
        const uint32_t cap = TF(TF_CAP, m);
        const uint32_t ph = ln.phase(m);
        const uint32_t qs = (len == 0u ? IK_EMPTY : 0u) | (len <= cap ? IK_LE : 0u) |
                            ((uint64_t)len >= 2ull * (uint64_t)cap ? IK_GE2 : 0u);
        const uint32_t a = int_lut(INT_KEY(type, ph, qs));
        if (a & 3u) {
            ln.err = (a & 3u) == 1u ? 5u : 34u;  // InvalidModelState / a panic of the reference
            return true;
        }
        const uint32_t kind = (a >> IA_KIND_SHIFT) & 7u;
        if (!BASIC && kind == KI_XGW && (ln.dm_int | ln.dm_ext)) return false;
        const uint32_t ob = TF(TF_OUT, m) & 0xFFFFu;
        const uint32_t rcap = TF(TF_RING_CAP, m), w0 = TF(TF_RING_W0, m);
        if (!BASIC && kind == KI_PG) {  // parallel_gateway.rs:174-182
            const uint32_t w1 = TF(TF_RING_W1, m), n_in = (tf >> 8) & 0xFFu;
            const uint32_t cnt = q;
            const uint32_t i = ring_find(ln, w1, cnt, n_in);  // canonical choice: earliest-inserted full collection
            if (i == cnt) {  // passivate
                apply_u<false>(ln, m, UA_INF);
                return true;
            }
            apply_u<false>(ln, m, UA_ZERO);  // send_job :118-143
            const uint32_t c = RW(ln, w0 + i);
            for (uint32_t j = i; j + 1u < cnt; j++) {  // remove, keeping insertion order
                RW(ln, w0 + j) = RW(ln, w0 + j + 1u);
                RW(ln, w1 + j) = RW(ln, w1 + j + 1u);
            }
            ln.W(wrow) = cnt - 1u;
            if (INSTR)
                for (uint32_t p = 0; p < n_out; p++) record<INSTR>(ln, m, ACT_DEPARTURE, c, p + 1u);
            pg.rel = n_out ? (n_out | REL_PORTS) : 0u;
            pg.rel_mp = m | (ob << 8);
            pg.rel_one = c;
            return true;
        }
        if (a & IA_SETPH) ln.set_phase(m, (a >> IA_PH_SHIFT) & 3u);
        const uint32_t ua = (a >> IA_U_SHIFT) & 7u;
        if (ua != UA_KEEP && !apply_u<false>(ln, m, ua)) return true;
        const uint32_t nr = (a >> IA_NREL_SHIFT) & 3u;
        uint32_t n = nr == NR_ONE ? 1u : nr == NR_ALL ? len : nr == NR_CAP ? cap : 0u;
        uint32_t port = 0, aux = 0, one = 0;
        bool from_ring = true;
        if (kind == KI_LB) {  // send_job load_balancer.rs:95-111 (the index is advanced first)
            if (n_out == 0u) {
                ln.err = 34;  // `% 0` panics
                return true;
            }
            uint32_t next = ln.W(wrow + 1u) + 1u;
            if (next >= n_out) next = 0;
            ln.W(wrow + 1u) = next;
            port = next;
            aux = next + 1u;
        } else if (!BASIC && kind == KI_XGW) {  // send_jobs exclusive_gateway.rs:110-134: one draw per firing
            const uint32_t wt = TF(TF_WTAB, m);
            const uint32_t rr = RARE ? TF(TF_RNG, m) : 0u;
            SavedRng sv{};
            if (RARE && rr) sv = own_begin(ln, m, rr);
            const uint32_t idx = sample_index_p(ln, flags, TF(TF_OUT, m) >> 16, wt & 0xFFFFu, TabP{tab.vt, m, wt >> 24});
            if (RARE && rr) own_end(ln, rr, sv);
            if (ln.err) return true;
            if (idx >= n_out) {
                ln.err = 34;  // flow_paths[idx] out of bounds panics
                return true;
            }
            port = idx;
            aux = idx + 1u;
        } else if (kind == KI_GEN) {  // generator.rs:169-177 (the draw was noted first)
            from_ring = false;
            if (ph == 1u) {  // release_job :98-123
                const uint32_t last = q + 1u;
                if (last >= SIMB_NUM_LITERAL) {
                    ln.err = 32;  // content code space exhausted
                    return true;
                }
                ln.W(wrow) = last;
                one = TF(TF_PREFIX, m) | last;
                n = 1;
                record<INSTR>(ln, m, ACT_GENERATION, one);
            } else {  // initialize_generation :125-146
                ln.set_phase(m, 1);
                record<INSTR>(ln, m, ACT_INITIALIZATION, SIMB_CONTENT_NONE);
            }
        } else if (kind == KI_STORED) {  // storage.rs:115-130; stopwatch.rs:216-250
            from_ring = false;
            const bool sw = type == MT_STOPWATCH;
            one = ln.W(wrow + (sw ? 1u : 0u));
            record<INSTR>(ln, m, !sw ? ACT_DEPARTURE : (flags & MF_METRIC_MAX) ? ACT_MAXIMUM_FETCH : ACT_MINIMUM_FETCH, one);
            n = one != SIMB_CONTENT_NONE ? 1u : 0u;
        } else if (kind == KI_PROC_START) {  // process_next processor.rs:160-175
            record<INSTR>(ln, m, ACT_PROCESSING_START, RW(ln, w0 + head));
            if (flags & MF_STAT) wait_sample(ln, ln.time - RD(ln, TF(TF_RING_D, m) + head));
        } else if (!BASIC && kind == KI_SGATE) {  // stochastic_gate.rs:174-183
            from_ring = false;
            one = RW(ln, w0 + head);
            const bool pass = RW(ln, TF(TF_RING_W1, m) + head) != 0u;
            uint32_t nh = head + 1u;
            if (nh >= rcap) nh = 0;
            ln.W(wrow) = (nh << 16) | (len - 1u);
            record<INSTR>(ln, m, pass ? ACT_PASS : ACT_BLOCK, one);  // pass_job :129-141 / block_job :143-148
            n = pass ? 1u : 0u;
        }
        // All records of one events_int call precede its messages (the reference routes the returned Vec
        // afterwards, mod.rs:240-263).
        if (INSTR && from_ring)
            for (uint32_t i = 0; i < n; i++) {
                uint32_t s = head + i;
                if (s >= rcap) s -= rcap;
                record<INSTR>(ln, m, ACT_DEPARTURE, RW(ln, w0 + s), aux);
            }
        // the jobs leave one per micro-iteration (release_one), so a batch does not hold up the other lanes
        pg.rel = n ? (n | (from_ring ? REL_RING : 0u)) : 0u;
        pg.rel_mp = m | ((ob + port) << 8);
        pg.rel_one = one;
        return true;
    }
    // one job of the release in progress: pop it (unless the content is given) and route it
    template <bool INSTR>
    SIMB_HD void release_one(Lane& ln, Prog& pg) {
        uint32_t c = pg.rel_one;
        const uint32_t m = pg.rel_mp & 0xFFu;
        if (pg.rel & REL_RING) {
            const uint32_t wrow = TF(TF_ROWS, m) & 0xFFFFu, rcap = TF(TF_RING_CAP, m);
            const uint32_t qq = ln.W(wrow);
            const uint32_t h = qq >> 16;
            c = RW(ln, TF(TF_RING_W0, m) + h);
            uint32_t nh = h + 1u;
            if (nh >= rcap) nh = 0;
            ln.W(wrow) = (nh << 16) | ((qq & 0xFFFFu) - 1u);
        }
        emit2<INSTR>(ln, pg.rel_mp >> 8, c);
        if (pg.rel & REL_PORTS) pg.rel_mp += 1u << 8;  // ParallelGateway: the same job on the next output port
        pg.rel--;
        if ((pg.rel & 0xFFFFu) == 0u) pg.rel = 0;
    }

    // ---- Simulation::step for one lane, generic interpreter -----------------------------------------
    // Replications are independent, so the lanes of a warp need not agree on WHICH step they are in: the
    // warp runs `micro` in a loop, and in each iteration every lane advances its own step by at most one
    // external event, the time advance, and one internal event.  A replication whose step delivers one
    // message and fires one model completes a whole step per iteration; one with a burst (a fan-out, a
    // batch release) takes more iterations for that step while the other lanes move on to their next
    // steps.  The cost of a launch is then the maximum over lanes of the SUM of their work, not the sum
    // over steps of the maximum (what step-level lockstep pays).
    // One iteration for a lane that is `act`ive (others fall through).  Returns true when the lane has just
    // completed a step (its stage is 0 again).
    template <bool INSTR>
    SIMB_HD bool micro(Lane& ln, Prog& pg, bool act) {
        const uint32_t M = net.n_models;
        // ---- step begin (:199-201); the mailbox is sorted by target model, unknown targets (0xFF) last ----
        if (act && pg.stage == 0u) {
            pg.stage = 1u;
            pg.nev = 0;
            pg.ei = 0;
            pg.np = ln.npend;
            pg.copen = 0;
        }
        // ---- one external event: model order, then message order (:202-223) ----
        if (act && pg.stage == 1u && pg.ei < pg.np && !ln.err) {
            const uint32_t rw = mail_word(ln, pg.ei, 1);
            if (ROUTE_TGT(rw) < M) {
                if (ext2<INSTR>(ln, ROUTE_TGT(rw), ROUTE_PORT(rw), mail_word(ln, pg.ei, 0))) {
                    pg.ei++;
                    pg.nev++;
                } else {
                    pg.flush = true;
                }
            } else {
                pg.ei = pg.np;  // a target id matching no model: never delivered
            }
        }
        // ---- every message delivered (or a failed events_ext, `?` at :222): pay the deferred draws, advance ----
        // ---- the ONE place where deferred draws are paid: at the end of the external phase, and for a handler
        // that is waiting to draw inline ----
        const bool ext_done = act && pg.stage == 1u && (pg.ei >= pg.np || ln.err);
        if (ext_done || pg.flush) {  // (pg.flush without act: the settle pass at the end of a launch)
            pg.flush = false;
            if (RNG == 0 && ln.ccount <= 1u) philox_refill(ln);
            flush_draws(ln);
        }
        if (ext_done) {
            pg.stage = 2u;
            pg.zmk = 0;
            if (!ln.err) {
                // time advance (:225-236) over the models whose until_next_event is not +INF
                double dt = 0.0;
                if (pg.np == 0u) {
                    if (ln.dirty) {  // search the smallest non-zero value (zeros are exactly the bits of zm)
                        double nm = SIMB_INF;
                        for (uint32_t mm = ln.amask & ~ln.zm; mm; mm &= mm - 1u)
                            nm = fmin(nm, ln.D(net.drow_until + (uint32_t)lowest_bit(mm)));
                        ln.nmin = nm;
                        ln.dirty = false;
                    }
                    dt = ln.zm ? fmin(ln.nmin, 0.0) : ln.nmin;
                    if (dt == SIMB_INF) {
                        // nothing is scheduled: the reference subtracts +INF from every model (NaN), the clock becomes +INF
                        for (uint32_t m = 0; m < M; m++)
                            if ((TF(TF_TYPE, m) & 0xFFu) != MT_PASSIVE) {
                                ln.D(net.drow_until + m) = ln.D(net.drow_until + m) - dt;
                                ln.amask |= 1u << m;
                            }
                    } else if (dt != 0.0) {
                        double nm = SIMB_INF;
                        uint32_t z = 0;
                        for (uint32_t mm = ln.amask; mm; mm &= mm - 1u) {
                            const uint32_t m = (uint32_t)lowest_bit(mm);
                            const double u = ln.D(net.drow_until + m) - dt;
                            ln.D(net.drow_until + m) = u;
                            if (u == 0.0) z |= 1u << m;
                            else nm = fmin(nm, u);
                        }
                        ln.zm = z;
                        ln.nmin = nm;
                    }
                }
                ln.time = ln.time + dt;
                ln.npend = 0;  // the mailbox is consumed; events_int refills it from slot 0
                pg.zmk = ln.zm;
            }
        }
        // ---- a Coupled model whose turn has come (coupled.rs:186-255): before any of its components fires, its
        // parked messages are delivered -- component order, then parked order --, the draws they owe are paid, and
        // the components that are imminent NOW are the ones that fire ----
        if (RARE && INSTR && net.n_coupled != 0u && act && pg.cgrp != 0u && pg.stage == 2u && pg.rel == 0u && !pg.flush &&
            !ln.err && (pg.zmk == 0u || (TF(TF_GROUP, (uint32_t)lowest_bit(pg.zmk)) >> 8) != pg.cgrp - 1u))
            close_group<INSTR>(ln, pg);
        if (RARE && net.n_coupled != 0u && act && !ln.err && pg.rel == 0u && !pg.flush) {
            if (pg.stage == 2u && pg.zmk) {
                const uint32_t cpl = TF(TF_GROUP, (uint32_t)lowest_bit(pg.zmk)) >> 8;
                if (cpl < SIMB_MAX_COUPLED && !(pg.copen & (1u << cpl))) {
                    // the Coupled model fires when ITS until_next_event is 0.0, and that is the minimum over its
                    // components (coupled.rs:304-310): one component below zero (a negative variate was drawn) keeps
                    // the whole model from firing in this step
                    const SimbCoupledDesc& cd = net.coupled[cpl];
                    bool below = false;
                    for (uint32_t c = cd.first; c < (uint32_t)cd.first + cd.count; c++)
                        if (ln.D(net.drow_until + c) < 0.0) below = true;
                    if (below) {
                        pg.zmk &= ~((cd.count >= 32u ? 0xFFFFFFFFu : ((1u << cd.count) - 1u)) << cd.first);
                    } else {
                        pg.copen |= 1u << cpl;
                        pg.cdel = cpl + 1u;
                        pg.cpos = (uint32_t)cd.first << 16;
                        pg.stage = 3u;
                    }
                }
            }
            if (pg.stage == 3u) {
                const SimbCoupledDesc& cd = net.coupled[pg.cdel - 1u];
                const uint32_t cnt = ln.W(cd.wrow);
                uint32_t comp = pg.cpos >> 16, idx = pg.cpos & 0xFFFFu;
                bool found = false;
                while (comp < (uint32_t)cd.first + cd.count) {  // next parked message of component `comp`
                    for (; idx < cnt; idx++)
                        if ((RW(ln, cd.ring_w1 + idx) & 0xFFu) == comp) {
                            found = true;
                            break;
                        }
                    if (found) break;
                    comp++;
                    idx = 0;
                }
                if (found) {
                    const uint32_t tp = RW(ln, cd.ring_w1 + idx);
                    if (ext2<INSTR>(ln, comp, (tp >> 8) & 0xFFu, RW(ln, cd.ring_w0 + idx))) {
                        pg.nev++;
                        idx++;
                    } else {
                        pg.flush = true;  // the component draws inline: pay the earlier draws, then deliver again
                    }
                    pg.cpos = (comp << 16) | idx;
                } else {  // all delivered: the list is emptied (:227); pay what the deliveries owe before looking at
                          // the components' until_next_event (:231-233)
                    ln.W(cd.wrow) = 0u;
                    pg.stage = 4u;
                    pg.flush = (ln.dm_int | ln.dm_ext) != 0u;
                }
            } else if (pg.stage == 4u) {
                const SimbCoupledDesc& cd = net.coupled[pg.cdel - 1u];
                const uint32_t mask = (cd.count >= 32u ? 0xFFFFFFFFu : ((1u << cd.count) - 1u)) << cd.first;
                pg.zmk = (pg.zmk & ~mask) | (ln.zm & mask);
                pg.cdel = 0;
                pg.stage = 2u;
            }
        }
        // ---- one internal event, in model order (:237-268) ----
        if (act && pg.stage == 2u && pg.zmk && pg.rel == 0u && !ln.err &&
            !(RARE && net.n_coupled != 0u && (TF(TF_GROUP, (uint32_t)lowest_bit(pg.zmk)) >> 8) < SIMB_MAX_COUPLED &&
              !(pg.copen & (1u << (TF(TF_GROUP, (uint32_t)lowest_bit(pg.zmk)) >> 8))))) {
            const uint32_t m = (uint32_t)lowest_bit(pg.zmk);
            if (RARE && INSTR && net.n_coupled != 0u && pg.cgrp == 0u && (TF(TF_GROUP, m) >> 8) < SIMB_MAX_COUPLED) {
                pg.cgrp = (TF(TF_GROUP, m) >> 8) + 1u;  // the first component of a Coupled model to fire in this step
                pg.cseq0 = ln.npend;
                ln.hold = true;
            }
            if (int2<INSTR>(ln, pg, m)) {
                pg.zmk &= pg.zmk - 1u;
                pg.nev++;
            } else {
                pg.flush = true;
            }
        }
        // ---- one outgoing job of that event ----
        if (act && pg.rel != 0u && !ln.err) release_one<INSTR>(ln, pg);
        // ---- step end ----
        if (act && ((pg.stage == 2u && pg.zmk == 0u && pg.rel == 0u) || (pg.stage >= 2u && ln.err))) {
            // (a failed step may leave draws owed by the handlers that ran before the failure: the reference made
            // them, and the launch pays them before it stores the lane -- step_kernel.cuh, "settle")
            if (RARE && INSTR && pg.cgrp != 0u) close_group<INSTR>(ln, pg);
            pg.stage = 0u;
            pg.rel = 0;
            pg.cdel = 0;
            ln.events += pg.nev;
            if (!ln.err) ln.steps++;
            return true;
        }
        return false;
    }
    template <bool INSTR>
    SIMB_HD void step2(Lane& ln, bool live) {
        if (!live) return;
        Prog pg;
        prog_init(pg);
        while (!micro<INSTR>(ln, pg, true)) {
        }
    }

    // ================================================================================================
    // Simulation::step (mod.rs:198-272) for one lane.  `live` lanes execute;
    // the others only take part in the warp-level votes.
    // ================================================================================================
    template <bool INSTR>
    SIMB_HD void ext_model(Lane& ln, uint32_t m, uint32_t np, uint32_t& nev) {
        const SimbModelDesc& md = net.models[m];
        for (uint32_t j = 0; j < np && !ln.err; j++) {
            const uint32_t rw = mail_word(ln, j, 1);
            if (ROUTE_TGT(rw) != m) continue;
            nev++;
            events_ext<INSTR>(ln, m, md, ROUTE_PORT(rw), mail_word(ln, j, 0));
        }
    }

    template <bool INSTR>
    SIMB_HD void step(Lane& ln, bool live) {
        if (!STATIC) {
            step2<INSTR>(ln, live);
            return;
        }
        const uint32_t M = net.n_models;
        const bool had = live && ln.npend > 0;
        uint32_t nev = 0;
        // ---- external events: model order, then message order (:202-223) ----
        if (warp_any(had)) {
            const uint32_t tmask = had ? ln.tmask : 0u;  // a target id matching no model has no bit: dropped
            const uint32_t np = had ? ln.npend : 0u;
            if (STATIC) {
                // shapes of more than four models: a warp vote skips the blocks of models no lane has mail for
                constexpr bool VOTE = !(SM <= 4 && SIMB_OPT_NOVOTE);
                const uint32_t wm = VOTE ? warp_or(tmask) : 0xFFFFFFFFu;
#define SIMB_EXT_ONE(m)                                                                \
    if ((m) < M && ((wm >> (m)) & 1u)) { /* warp-uniform */                            \
        if (((tmask >> (m)) & 1u) && !ln.err) ext_model<INSTR>(ln, (m), np, nev);      \
    }
                SIMB_FOR_MODELS(SM, SIMB_EXT_ONE);
#undef SIMB_EXT_ONE
            } else {
                // generic interpreter: every lane walks ITS OWN targets in model order (the reference's order
                // for this replication); the model-type dispatch diverges, but a warp step then costs the
                // number of distinct handler TYPES in flight, not the size of the union of target models
                // over 32 replications (measured on the 32-model network: 6 of 32 lanes active before).
                uint32_t tm = tmask;
                while (tm && !ln.err) {
                    const uint32_t m = (uint32_t)lowest_bit(tm);
                    tm &= tm - 1u;
                    ext_model<INSTR>(ln, m, np, nev);
                }
            }
        }
        // ---- pay the deferred draws (and top up the Philox prefetch) at this converged point ----
        if (RNG == 0) philox_topup(ln, live);
        if (live) flush_draws(ln);
        // A failed events_ext aborts the step (`?` at :222); the replication freezes.
        const bool go = live && !ln.err;
        // ---- time advance (:225-236) ----
        double dt = 0.0;
        uint32_t zmask = 0;
        if (STATIC) {
            if (go && !had) {
                dt = SIMB_INF;
#define SIMB_MIN_ONE(m) if ((m) < M) dt = fmin(dt, U(ln, (m)));
                SIMB_FOR_MODELS(SM, SIMB_MIN_ONE);
#undef SIMB_MIN_ONE
            }
            if (go) {
#define SIMB_ADV_ONE(m)                                          \
    if ((m) < M) {                                               \
        double u = U(ln, (m));                                   \
        if (!had && net.models[m].type != MT_PASSIVE) {          \
            u = u - dt;                                          \
            U(ln, (m)) = u;                                      \
        }                                                        \
        if (u == 0.0) zmask |= 1u << (m);                        \
    }
                SIMB_FOR_MODELS(SM, SIMB_ADV_ONE);
#undef SIMB_ADV_ONE
            }
        } else {
            if (go && !had) {
                // The full interpreter keeps these loops rolled: it is bound by instruction fetch, so code size costs
                // more than loop overhead (32-model network +6 %); the small BASIC build gains from unrolling.
                dt = SIMB_INF;
                if (BASIC) {
                    for (uint32_t m = 0; m < M; m++) dt = fmin(dt, ln.D(net.drow_until + m));
                } else {
#pragma unroll 1
                    for (uint32_t m = 0; m < M; m++) dt = fmin(dt, ln.D(net.drow_until + m));
                }
            }
            if (go && BASIC) {
                for (uint32_t m = 0; m < M; m++) {
                    double u = ln.D(net.drow_until + m);
                    if (!had && net.models[m].type != MT_PASSIVE) {
                        u = u - dt;
                        ln.D(net.drow_until + m) = u;
                    }
                    if (u == 0.0) zmask |= 1u << m;
                }
            } else if (go) {
#pragma unroll 1
                for (uint32_t m = 0; m < M; m++) {
                    double u = ln.D(net.drow_until + m);
                    if (!had && net.models[m].type != MT_PASSIVE) {
                        u = u - dt;
                        ln.D(net.drow_until + m) = u;
                    }
                    if (u == 0.0) zmask |= 1u << m;
                }
            }
        }
        if (go) {
            ln.time = ln.time + dt;
            ln.npend = 0;  // the mailbox is consumed; events_int refills it from slot 0
            ln.tmask = 0;
        }
        // ---- internal events in model order (:237-268) ----
        if (STATIC) {
            constexpr bool VOTE = !(SM <= 4 && SIMB_OPT_NOVOTE);
            const uint32_t wm = VOTE ? warp_or(zmask) : 0xFFFFFFFFu;
#define SIMB_INT_ONE(m)                                                                 \
    if ((m) < M && net.models[m].type != MT_PASSIVE && ((wm >> (m)) & 1u)) {            \
        if (((zmask >> (m)) & 1u) && !ln.err) {                                         \
            nev++;                                                                      \
            events_int<INSTR>(ln, (m), net.models[m]);                                  \
        }                                                                               \
    }
            SIMB_FOR_MODELS(SM, SIMB_INT_ONE);
#undef SIMB_INT_ONE
        } else {
            uint32_t zm = zmask;  // per lane, in model order (see the external phase)
            while (zm && !ln.err) {
                const uint32_t m = (uint32_t)lowest_bit(zm);
                zm &= zm - 1u;
                nev++;
                events_int<INSTR>(ln, m, net.models[m]);
            }
        }
        if (go) {
            ln.events += nev;
            if (!ln.err) ln.steps++;
        } else if (live) {
            ln.events += nev;
        }
    }
};

}  // namespace simb
