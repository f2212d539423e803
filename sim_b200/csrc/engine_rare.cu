// sim_b200/csrc/engine_rare.cu -- the generic step kernel with the rare continuous variables (Beta, Gamma,
// LogNormal, Weibull) compiled in.  A separate object: these samplers triple the kernel's code
// size, so networks that do not use them run the lean instantiations of engine.cu.
#include "step_kernel.cuh"

namespace simb {
StepKernel simb_generic_rare_kernel(int rng_kind, bool instrument, bool narrow) {
    if (narrow && !instrument && rng_kind == 0) return simb_step_kernel<0, false, -1, true, true>;
    if (instrument) {
        if (rng_kind == 0) return simb_step_kernel<0, true, -1, false, true>;
        if (rng_kind == 1) return simb_step_kernel<1, true, -1, false, true>;
        return simb_step_kernel<2, true, -1, false, true>;
    }
    if (rng_kind == 0) return simb_step_kernel<0, false, -1, false, true>;
    if (rng_kind == 1) return simb_step_kernel<1, false, -1, false, true>;
    return simb_step_kernel<2, false, -1, false, true>;
}
}  // namespace simb
