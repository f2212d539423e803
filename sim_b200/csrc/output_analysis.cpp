// sim_b200/csrc/output_analysis.cpp -- see output_analysis.hpp.
#include "output_analysis.hpp"

#include <cmath>
#include <limits>
#include <vector>

namespace simb {
namespace oa {

static const int ERR_PANIC = 34, ERR_PREREQ = 10;

uint64_t usize_sqrt(uint64_t n) {  // Babylonian iteration, utils/mod.rs:84-92
    uint64_t x = n, y = 1;
    while (x > y) {
        x = (x + y) / 2;
        y = n / x;
    }
    return x;
}

static const double kStudentT[100][7] = {
#include "t_table.inc"
};
static const double kAlphas[7] = {0.1, 0.05, 0.025, 0.01, 0.005, 0.001, 0.0005};
static const double kZ[7] = {1.2816, 1.6449, 1.9600, 2.3263, 2.5758, 3.0902, 3.2905};

int t_score(double alpha, uint64_t df, double* out) {  // t_scores.rs:9-30
    int col = -1;
    for (int i = 0; i < 7; i++)
        if (kAlphas[i] == alpha) col = col < 0 ? i : col;
    if (col < 0) return ERR_PANIC;  // position(..).unwrap() on an unlisted alpha
    if (df > 100) {                 // z row, t_scores.rs:32-34
        *out = kZ[col];
        return 0;
    }
    if (df == 0) return ERR_PANIC;  // [df - 1] with df == 0
    *out = kStudentT[df - 1][col];
    return 0;
}

void mean_variance(const double* p, uint64_t n, double* mean, double* variance) {  // mod.rs:14-40
    double total = 0.0;
    for (uint64_t i = 0; i < n; i++) total = total + p[i];
    const double m = total / (double)n;
    double acc = 0.0;
    for (uint64_t i = 0; i < n; i++) {
        const double dev = p[i] - m;
        acc = acc + dev * dev;
    }
    *mean = m;
    *variance = acc / (double)n;  // divides by n, not n-1 (mod.rs:39)
}

int independent_ci(double mean, double variance, uint64_t n, double alpha, double* lower, double* upper) {  // mod.rs:106-125
    if (n == 1) {
        *lower = *upper = mean;
        return 0;
    }
    double t;
    int e = t_score(alpha, n - 1, &t);
    if (e) return e;
    const double half = t * std::sqrt(variance) / std::sqrt((double)n);
    *lower = mean - half;
    *upper = mean + half;
    return 0;
}

int batch_means_ci(double mean, double variance, uint64_t batches, double alpha, double* lower, double* upper) {
    if (batches == 0) return ERR_PREREQ;
    if (batches == 1) {  // mod.rs:317-322
        *lower = *upper = mean;
        return 0;
    }
    double t_lo, t_hi;  // df = batch_count below, batch_count - 1 above (mod.rs:327 vs :330)
    int e = t_score(alpha, batches, &t_lo);
    if (e) return e;
    if ((e = t_score(alpha, batches - 1, &t_hi))) return e;
    *lower = mean - t_lo * std::sqrt(variance) / std::sqrt((double)batches);
    *upper = mean + t_hi * std::sqrt(variance) / std::sqrt((double)batches);
    return 0;
}

int steady_state(const double* x, uint64_t n, double alpha, SteadyState* out) {
    if (n < 2) return ERR_PANIC;  // `len() - 2` underflows (mod.rs:227)
    // set_to_fixed_budget, mod.rs:224-260: backward scan; note the operator precedence at :233
    std::vector<double> mser(n - 1, 0.0);
    double s = 0.0, q = 0.0;
    const double nf = (double)n;
    for (uint64_t d = n - 1; d-- > 0;) {
        s += x[d + 1];
        q += x[d + 1] * x[d + 1];
        const double w = nf - (double)d;
        mser[d] = q - (s * s) / (w * w * w);
    }
    double best = std::numeric_limits<double>::infinity();
    for (uint64_t i = 0; i < (n - 1) / 2; i++) best = std::fmin(best, mser[i]);
    uint64_t del = n;  // first index (over ALL of mser) equal to the first-half minimum
    for (uint64_t i = 0; i < n - 1; i++)
        if (mser[i] == best) {
            del = i;
            break;
        }
    if (del == n) return ERR_PREREQ;
    uint64_t batches = usize_sqrt(n - del);
    if (batches > 30) batches = 30;
    if (batches == 0) return ERR_PANIC;
    const uint64_t size = (n - del) / batches;
    del = n - batches * size;  // leftovers are dropped from the beginning (:256-258)
    // calculate_batch_statistics, mod.rs:268-296
    std::vector<double> means(batches);
    for (uint64_t b = 0; b < batches; b++) {
        double t = 0.0;
        for (uint64_t i = del + size * b; i < del + size * (b + 1); i++) t = t + x[i];
        means[b] = t / (double)size;
    }
    double bm, bv;
    mean_variance(means.data(), batches, &bm, &bv);
    out->deletion_point = del;
    out->batch_size = size;
    out->batch_count = batches;
    out->mean = bm;
    out->variance = bv;
    if (batches == 1) {
        out->lower = out->upper = bm;
        return 0;
    }
    // confidence_interval_mean, mod.rs:302-333: df = batch_count below, batch_count - 1 above
    double t_lo, t_hi;
    int e = t_score(alpha, batches, &t_lo);
    if (e) return e;
    if ((e = t_score(alpha, batches - 1, &t_hi))) return e;
    // same association as the source: (t * sqrt(var)) / sqrt(n)
    out->lower = bm - t_lo * std::sqrt(bv) / std::sqrt((double)batches);
    out->upper = bm + t_hi * std::sqrt(bv) / std::sqrt((double)batches);
    return 0;
}

}  // namespace oa
}  // namespace simb
