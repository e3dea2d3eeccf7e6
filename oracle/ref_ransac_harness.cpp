// ORACLE — test infrastructure only (never linked into the product).
//
// Pins the oracle's RANSAC loop against the REFERENCE's own code: this translation unit includes the
// reference's calib::ransac<Estimator> template unmodified, from where it lies
// (/root/reference/include/calib/estimation/common/ransac.h:121-194 and its detail:: helpers :40-117),
// and instantiates it with an estimator whose four hooks are the oracle's restated homography pieces
// (the reference's HomographyEstimator needs Eigen, which this image does not have).  Because the
// hooks are the very functions ransac_one() calls, any difference between ref_ransac_homography()
// and orc_ransac_homography() can only come from the loop: the std::sample stream, the degeneracy /
// failed-fit / min_inliers `continue`s, refit_model, is_better_model, calculate_iterations and the
// result bookkeeping.  tests/test_oracle_ref_ransac.py asserts they agree bit for bit.
//
// Built by `make -C oracle ref` into oracle/_ref/ (git-ignored) only where /root/reference exists.
#include "ransac.cpp"  // the oracle's restatement, statics included

#include <array>
#include <optional>
#include <span>

#include "calib/estimation/common/ransac.h"  // the reference's template (resolved through -I /root/reference/include)

namespace {

struct Corr { double x, y, u, v; };
struct HModel { std::array<double, 9> h{1, 0, 0, 0, 1, 0, 0, 0, 1}; };

// same hook set as the reference's HomographyEstimator (linear/homographyestimator.cpp:121-170)
struct OracleHomographyEstimator final {
    using Datum = Corr;
    using Model = HModel;
    static constexpr size_t k_min_samples = 4;

    static std::optional<Model> fit_subset(const std::vector<Datum>& data, std::span<const int> idxs) {
        std::vector<double> sx, sy, su, sv;
        for (int id : idxs) { sx.push_back(data[id].x); sy.push_back(data[id].y); su.push_back(data[id].u); sv.push_back(data[id].v); }
        Model m;
        if (!orc::dlt_homography(sx, sy, su, sv, m.h.data())) return std::nullopt;
        return m;
    }
    static auto fit(const std::vector<Datum>& data, std::span<const int> sample) -> std::optional<Model> {
        if (sample.size() < k_min_samples) return std::nullopt;
        return fit_subset(data, sample);
    }
    static auto residual(const Model& m, const Datum& d) -> double {
        double Hi[9]; orc::mat3_inv(m.h.data(), Hi);
        return orc::transfer_error(m.h.data(), Hi, d.x, d.y, d.u, d.v);
    }
    static auto refit(const std::vector<Datum>& data, std::span<const int> inliers) -> std::optional<Model> {
        if (inliers.size() < k_min_samples) return std::nullopt;
        return fit_subset(data, inliers);
    }
    static auto is_degenerate(const std::vector<Datum>& data, std::span<const int> sample) -> bool {
        std::vector<double> x(data.size()), y(data.size());
        for (size_t i = 0; i < data.size(); ++i) { x[i] = data[i].x; y[i] = data[i].y; }
        return orc::degenerate(x.data(), y.data(), sample.data());
    }
};

struct Pt3 { double x, y, z; };
struct PlaneModel { std::array<double, 4> p{0, 0, 0, 0}; };

// same hook set as the reference's PlaneRansacEstimator (linear/planefit.cpp:9-62), k_min_samples = 3
struct OraclePlaneEstimator final {
    using Datum = Pt3;
    using Model = PlaneModel;
    static constexpr size_t k_min_samples = 3;

    static void soa(const std::vector<Datum>& data, std::vector<double>& x, std::vector<double>& y, std::vector<double>& z) {
        x.resize(data.size()); y.resize(data.size()); z.resize(data.size());
        for (size_t i = 0; i < data.size(); ++i) { x[i] = data[i].x; y[i] = data[i].y; z[i] = data[i].z; }
    }
    static auto fit(const std::vector<Datum>& data, std::span<const int> sample) -> std::optional<Model> {
        if (sample.size() < k_min_samples) return std::nullopt;
        std::vector<double> x, y, z; soa(data, x, y, z);
        Model m;
        if (!orc::plane_from_sample(x.data(), y.data(), z.data(), sample.data(), m.p.data())) return std::nullopt;
        return m;
    }
    static auto residual(const Model& m, const Datum& d) -> double { return orc::plane_residual(m.p.data(), d.x, d.y, d.z); }
    static auto refit(const std::vector<Datum>& data, std::span<const int> inliers) -> std::optional<Model> {
        if (inliers.size() < k_min_samples) return std::nullopt;
        std::vector<double> x, y, z;
        for (int id : inliers) { x.push_back(data[id].x); y.push_back(data[id].y); z.push_back(data[id].z); }
        Model m;
        if (!orc::fit_plane_svd(x, y, z, m.p.data())) return std::nullopt;
        return m;
    }
    static auto is_degenerate(const std::vector<Datum>& data, std::span<const int> sample) -> bool {
        if (sample.size() < k_min_samples) return true;
        std::vector<double> x, y, z; soa(data, x, y, z);
        double tmp[4];
        return !orc::plane_from_sample(x.data(), y.data(), z.data(), sample.data(), tmp);
    }
};

calib::RansacOptions to_opts(const orc_ransac_options* o) {
    calib::RansacOptions opts;
    opts.max_iters = o->max_iters; opts.thresh = o->thresh; opts.min_inliers = o->min_inliers;
    opts.confidence = o->confidence; opts.seed = o->seed; opts.refit_on_inliers = o->refit_on_inliers != 0;
    return opts;
}

}  // namespace

// fit_plane_ransac's packaging (planefit.cpp:86-104) around the reference's loop
extern "C" int ref_ransac_plane(int32_t n, const double* x, const double* y, const double* z, const orc_ransac_options* o,
                                orc_plane_result* res, uint8_t* inlier_mask) {
    std::vector<Pt3> data(n);
    for (int i = 0; i < n; ++i) data[i] = {x[i], y[i], z[i]};
    std::memset(res, 0, sizeof *res);
    res->inlier_rms = std::numeric_limits<double>::infinity();
    if (inlier_mask) std::memset(inlier_mask, 0, n);
    if (data.size() < OraclePlaneEstimator::k_min_samples) return 0;
    const auto best = calib::ransac<OraclePlaneEstimator>(data, to_opts(o));
    if (!best.success) return 0;
    res->success = 1; res->iters = best.iters; res->n_inliers = static_cast<int>(best.inliers.size());
    res->inlier_rms = best.inlier_rms;
    for (int i = 0; i < 4; ++i) res->plane[i] = best.model.p[i];
    if (inlier_mask) for (int id : best.inliers) inlier_mask[id] = 1;
    return 0;
}

extern "C" int ref_ransac_homography(int32_t n, const double* x, const double* y, const double* u, const double* v,
                                     const orc_ransac_options* o, orc_ransac_result* res, uint8_t* inlier_mask) {
    std::vector<Corr> data(n);
    for (int i = 0; i < n; ++i) data[i] = {x[i], y[i], u[i], v[i]};
    const auto best = calib::ransac<OracleHomographyEstimator>(data, to_opts(o));
    std::memset(res, 0, sizeof *res);
    res->success = best.success ? 1 : 0;
    res->iters = best.iters;
    res->n_inliers = static_cast<int>(best.inliers.size());
    res->inlier_rms = best.inlier_rms;
    for (int i = 0; i < 9; ++i) res->hmtx[i] = best.model.h[i];
    if (inlier_mask) { std::memset(inlier_mask, 0, n); for (int id : best.inliers) inlier_mask[id] = 1; }
    return 0;
}
