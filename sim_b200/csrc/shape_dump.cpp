// sim_b200/csrc/shape_dump.cpp -- build-time tool: lower a network (JSON files) and print its
// SimbNetDesc as a C++ aggregate initialiser.  tools/gen_shapes.py uses it to bake the network
// shapes that get specialised step kernels (sim_b200/csrc/shapes.inc).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "lower.hpp"

using namespace simb;

static std::string slurp(const char* path) {
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
static void pd(double v) {
    if (std::isnan(v)) std::printf("__builtin_nan(\"\")");
    else if (std::isinf(v)) std::printf(v > 0 ? "__builtin_huge_val()" : "-__builtin_huge_val()");
    else std::printf("%a", v);
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: shape_dump models.json connectors.json [stat_model_id|-] [instr]\n");
        return 2;
    }
    LowerOptions opt;
    if (argc > 3 && std::strcmp(argv[3], "-") != 0) opt.stat_model_id = argv[3];
    if (argc > 4 && std::strcmp(argv[4], "instr") == 0) opt.instrument = true;
    HostNet hn;
    std::string err;
    int e = lower_network(slurp(argv[1]).c_str(), slurp(argv[2]).c_str(), opt, &hn, &err);
    if (e) {
        std::fprintf(stderr, "lowering failed (%d): %s\n", e, err.c_str());
        return 1;
    }
    const SimbNetDesc& n = hn.desc;
    std::printf("{%uu, %uu, %uu, %uu, %uu, %uu,\n", n.n_models, n.n_routes, n.n_connectors, n.max_pending, n.bm_warmup, n.bm_size);
    std::printf(" %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u,\n", n.n_drows, n.n_wrows, n.drow_time,
                n.drow_until, n.drow_wait, n.drow_bm, n.wrow_ctl, n.wrow_rng, n.wrow_steps, n.wrow_events, n.wrow_hash, n.wrow_phase,
                n.n_phase_words, n.wrow_pend, n.wrow_waitn, n.uses_exp, n.uses_normal, n.wrow_dmask, n.pend_tile);
    std::printf(" {\n");
    for (int i = 0; i < SIMB_MAX_MODELS; i++) {
        const SimbModelDesc& m = n.models[i];
        std::printf("  {%u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %uu, %uu, %uu, %uu, %uu, %uu, %u, 0, ", m.type, m.n_in, m.n_out, m.flags,
                    m.dist_kind, m.dist_err, m.wrow, m.drow, m.out_base, m.wtab, m.n_w, m.ring_cap, m.ring_w0, m.ring_w1, m.ring_d,
                    m.cap, m.prefix, m.rng_row);
        pd(m.p0);
        std::printf(", ");
        pd(m.p1);
        std::printf(", ");
        pd(m.p2);
        std::printf(", %lluull, %lluull, %lluull, %lluull},\n", (unsigned long long)m.q0, (unsigned long long)m.q1,
                    (unsigned long long)m.q2, (unsigned long long)m.rng_seed);
    }
    std::printf(" },\n {");
    for (int i = 0; i <= SIMB_MAX_OUT_PORTS; i++) std::printf("%u,", n.route_begin[i]);
    std::printf("},\n {");
    for (int i = 0; i < SIMB_MAX_ROUTES; i++) std::printf("{%u,%u,%u},", n.routes[i].tgt_model, n.routes[i].tgt_port, n.routes[i].connector);
    std::printf("},\n {");
    for (int i = 0; i < SIMB_MAX_WEIGHTS; i++) std::printf("%lluull,", (unsigned long long)n.wcum[i]);
    std::printf("},\n %uu, {", n.n_coupled);
    for (int i = 0; i < SIMB_MAX_COUPLED; i++)
        std::printf("{%u,%u,%u,%uu,%uu,%uu},", n.coupled[i].first, n.coupled[i].count, n.coupled[i].wrow, n.coupled[i].ring_cap,
                    n.coupled[i].ring_w0, n.coupled[i].ring_w1);
    std::printf("},\n {");
    for (int i = 0; i < SIMB_MAX_MODELS; i++) std::printf("%u,", n.group_base[i]);
    std::printf("},\n {");
    for (int i = 0; i < SIMB_MAX_MODELS; i++) std::printf("%u,", n.coupled_of[i]);
    std::printf("}}\n");
    return 0;
}
