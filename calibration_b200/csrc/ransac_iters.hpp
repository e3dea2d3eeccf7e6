// Host side of the RANSAC kernels' adaptive iteration bound (plain C++: also compiled by tests/host_emul).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/calib_b200.h"

namespace {

// calculate_iterations (ransac.h:64-78) evaluated on the host for every possible inlier count;
// -1 encodes "max_iters".  Same expression, same libm as the reference's own host code.
inline std::vector<int> build_niter_table(int n, const cal_ransac_options& o, int min_samples) {
    std::vector<int> t(n + 1, -1);
    for (int k = 0; k <= n; ++k) {
        const double w = (double)k / (double)n;
        if (o.confidence <= 0.0 || w <= 0.0) continue;
        const double denom = std::log(std::max(1e-12, 1.0 - std::pow(w, (double)min_samples)));
        if (denom >= 0.0) continue;
        const double v = std::ceil(std::log(1.0 - o.confidence) / denom);
        // static_cast<int> of an out-of-range double is what x86 cvttsd2si yields: INT_MIN,
        // which std::clamp then lifts to iters_so_far
        t[k] = (v >= 2147483648.0 || v < -2147483648.0 || v != v) ? INT32_MIN : (int)v;
        if (t[k] == -1) t[k] = -2;  // keep -1 reserved for "max_iters" (cannot occur: the ratio is >= 0)
    }
    return t;
}

// the clamp of calculate_iterations applied to a table entry, as the kernels do after every scored hypothesis
inline int next_iteration_bound(int table_entry, int iters_so_far, int max_iters) {
    int nd = table_entry == -1 ? max_iters : table_entry;
    if (nd < iters_so_far) nd = iters_so_far;
    if (nd > max_iters) nd = max_iters;
    return nd;
}

}  // namespace
