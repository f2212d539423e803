// sim_b200/csrc/output_analysis.hpp -- host-side restatement of sim::output_analysis
// (sim/src/output_analysis/mod.rs, t_scores.rs) and utils::usize_sqrt (sim/src/utils/mod.rs:84-92),
// kept bit-faithful including the quirks listed in SURVEY.md Appendix D.
#pragma once
#include <stdint.h>

namespace simb {
namespace oa {

uint64_t usize_sqrt(uint64_t n);
// returns 0 or SIMB_ERR_PANIC (the reference unwraps / indexes out of range)
int t_score(double alpha, uint64_t df, double* out);
// sample_mean + sample_variance (population variance, divides by n) in the reference's fold order
void mean_variance(const double* points, uint64_t n, double* mean, double* variance);
// IndependentSample::confidence_interval_mean given (mean, variance, n)
int independent_ci(double mean, double variance, uint64_t n, double alpha, double* lower, double* upper);

// SteadyStateOutput::confidence_interval_mean from (batches_mean, batches_variance, batch_count)
int batch_means_ci(double mean, double variance, uint64_t batch_count, double alpha, double* lower, double* upper);

struct SteadyState {
    uint64_t deletion_point, batch_size, batch_count;
    double mean, variance, lower, upper;
};
// SteadyStateOutput::post + confidence_interval_mean (MSER-style deletion, <= 30 batches)
int steady_state(const double* series, uint64_t n, double alpha, SteadyState* out);

}  // namespace oa
}  // namespace simb
