// sim_b200/csrc/shape_kernel.cu -- compiled once per baked network shape (-DSIMB_SHAPE_ID=k): the step
// kernel specialised for kShapeDev<k> (shapes.inc), Philox streams; the *_instr shapes are the same template with the trace compiled in.
#include "step_kernel.cuh"

#ifndef SIMB_SHAPE_ID
#error "compile with -DSIMB_SHAPE_ID=<baked shape index>"
#endif
#define SIMB_CAT2(a, b) a##b
#define SIMB_CAT(a, b) SIMB_CAT2(a, b)

namespace simb {
#if SIMB_SHAPE_ID < SIMB_N_SHAPES
StepKernel SIMB_CAT(simb_shape_kernel_, SIMB_SHAPE_ID)() { return simb_step_kernel<0, shape_instr<SIMB_SHAPE_ID>(), SIMB_SHAPE_ID, false>; }
StepKernel SIMB_CAT(simb_shape_stream_kernel_, SIMB_SHAPE_ID)() { return simb_step_kernel<0, shape_instr<SIMB_SHAPE_ID>(), SIMB_SHAPE_ID, true>; }
#else
StepKernel SIMB_CAT(simb_shape_kernel_, SIMB_SHAPE_ID)() { return nullptr; }
StepKernel SIMB_CAT(simb_shape_stream_kernel_, SIMB_SHAPE_ID)() { return nullptr; }
#endif
}  // namespace simb
