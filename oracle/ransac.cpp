// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of estimate_homography(view, RansacOptions):
//   ransac<Estimator>            include/calib/estimation/common/ransac.h:54-194
//   HomographyEstimator          src/estimation/linear/homographyestimator.cpp:16-174
//   wrapper + symmetric_rms_px   src/estimation/optim/homography.cpp:18-73
// plus the libstdc++-13 std::sample / std::mt19937_64 stream the reference
// draws its minimal samples from (ransac.h:135,144-145; algorithm read from
// /usr/include/c++/13/bits/stl_algo.h:5841-5907 and uniform_int_dist.h:252-281,
// SURVEY Appendix C).  orc_sample_stream_libstdcxx draws the same stream with
// the toolchain's real std::sample so tests can pin the restatement.
#include <omp.h>

#include <algorithm>
#include <numeric>
#include <random>

#include "oracle_api.h"
#include "oracle_math.hpp"

namespace orc {

// std::mt19937_64 restated (ISO C++ [rand.predef])
struct MT64 {
    uint64_t x[312]; int idx;
    explicit MT64(uint64_t seed) {
        x[0] = seed;
        for (int i = 1; i < 312; ++i) x[i] = 6364136223846793005ULL * (x[i - 1] ^ (x[i - 1] >> 62)) + static_cast<uint64_t>(i);
        idx = 312;
    }
    void twist() {
        const uint64_t UP = 0xFFFFFFFF80000000ULL, LO = 0x7FFFFFFFULL, A = 0xB5026F5AA96619E9ULL;
        for (int i = 0; i < 312; ++i) {
            const uint64_t y = (x[i] & UP) | (x[(i + 1) % 312] & LO);
            x[i] = x[(i + 156) % 312] ^ (y >> 1) ^ ((y & 1) ? A : 0ULL);
        }
        idx = 0;
    }
    uint64_t operator()() {
        if (idx >= 312) twist();
        uint64_t z = x[idx++];
        z ^= (z >> 29) & 0x5555555555555555ULL;
        z ^= (z << 17) & 0x71D67FFFEDA60000ULL;
        z ^= (z << 37) & 0xFFF7EEE000000000ULL;
        z ^= z >> 43;
        return z;
    }
};

// uniform_int_distribution<unsigned long>{0, range-1}(g) for a 64-bit engine:
// Lemire's nearly divisionless method (uniform_int_dist.h:252-281)
static inline uint64_t lemire(MT64& g, uint64_t range) {
    unsigned __int128 product = static_cast<unsigned __int128>(g()) * range;
    uint64_t low = static_cast<uint64_t>(product);
    if (low < range) {
        const uint64_t threshold = -range % range;
        while (low < threshold) { product = static_cast<unsigned __int128>(g()) * range; low = static_cast<uint64_t>(product); }
    }
    return static_cast<uint64_t>(product >> 64);
}

// std::sample(0..N-1, k) — selection sampling, two draws per engine call
// (stl_algo.h:5841-5907 with __gen_two_uniform_ints :3717-3724)
static void sample_indices(MT64& g, int N, int k, int* out) {
    uint64_t uns = static_cast<uint64_t>(N), n = std::min<uint64_t>(k, uns);
    int first = 0, o = 0;
    if (N == 0) return;
    if (UINT64_MAX / uns >= uns) {
        while (n != 0 && uns >= 2) {
            const uint64_t b1 = uns - 1;
            const uint64_t xx = lemire(g, uns * b1);
            const uint64_t p0 = xx / b1, p1 = xx % b1;
            --uns;
            if (p0 < n) { out[o++] = first; --n; }
            ++first;
            if (n == 0) break;
            --uns;
            if (p1 < n) { out[o++] = first; --n; }
            ++first;
        }
    }
    for (; n != 0; ++first) {
        --uns;
        if (lemire(g, uns + 1) < n) { out[o++] = first; --n; }
    }
}

// normalize_points_2d (homographyestimator.cpp:16-43)
static void normalize_points(const std::vector<double>& px, const std::vector<double>& py, std::vector<double>& ox,
                             std::vector<double>& oy, double* T) {
    const size_t n = px.size();
    double sx = 0, sy = 0; for (size_t i = 0; i < n; ++i) { sx += px[i]; sy += py[i]; }
    const double dn = static_cast<double>(std::max<size_t>(1, n));
    const double cx = sx / dn, cy = sy / dn;
    double md = 0; for (size_t i = 0; i < n; ++i) { const double dx = px[i] - cx, dy = py[i] - cy; md += std::sqrt(dx * dx + dy * dy); }
    md /= dn;
    const double sigma = md > 0 ? 1.4142135623730951 / md : 1.0;
    T[0] = sigma; T[1] = 0; T[2] = -sigma * cx; T[3] = 0; T[4] = sigma; T[5] = -sigma * cy; T[6] = 0; T[7] = 0; T[8] = 1;
    ox.resize(n); oy.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const double hx = T[0] * px[i] + T[1] * py[i] + T[2], hy = T[3] * px[i] + T[4] * py[i] + T[5], hz = T[6] * px[i] + T[7] * py[i] + T[8];
        ox[i] = hx / hz; oy[i] = hy / hz;
    }
}

// normalize_and_estimate_homography (homographyestimator.cpp:45-78). The
// reference takes the last right singular vector of Eigen::JacobiSVD(2N x 9);
// here a one-sided Jacobi SVD yields the same vector (up to sign, removed by
// the division by h9).
static bool dlt_homography(const std::vector<double>& sx, const std::vector<double>& sy, const std::vector<double>& du,
                           const std::vector<double>& dv, double* H) {
    std::vector<double> xn, yn, un, vn; double Ts[9], Td[9];
    normalize_points(sx, sy, xn, yn, Ts);
    normalize_points(du, dv, un, vn, Td);
    const int n = static_cast<int>(sx.size());
    const int m = std::max(2 * n, 9);
    std::vector<double> A(static_cast<size_t>(m) * 9, 0.0);
    for (int i = 0; i < n; ++i) {
        const double x = xn[i], y = yn[i], u = un[i], v = vn[i];
        double* r0 = &A[static_cast<size_t>(2 * i) * 9]; double* r1 = r0 + 9;
        r0[0] = -x; r0[1] = -y; r0[2] = -1; r0[6] = u * x; r0[7] = u * y; r0[8] = u;
        r1[3] = -x; r1[4] = -y; r1[5] = -1; r1[6] = v * x; r1[7] = v * y; r1[8] = v;
    }
    std::vector<double> V, sv; jacobi_svd(A, m, 9, V, sv);
    int k = 0; for (int j = 1; j < 9; ++j) if (sv[j] < sv[k]) k = j;
    double Hn[9]; for (int i = 0; i < 9; ++i) Hn[i] = V[i * 9 + k];
    const double h9 = Hn[8]; for (int i = 0; i < 9; ++i) Hn[i] = Hn[i] / h9;
    double Tdi[9], M[9]; mat3_inv(Td, Tdi); mat3_mul(Tdi, Hn, M); mat3_mul(M, Ts, H);
    return std::isfinite(H[0]);
}

// symmetric_transfer_error (homographyestimator.cpp:80-93); Hi = H^-1 hoisted
// out of the per-point call (the reference recomputes the same matrix each time)
static inline double transfer_error(const double* H, const double* Hi, double x, double y, double u, double v) {
    const double qx = H[0] * x + H[1] * y + H[2], qy = H[3] * x + H[4] * y + H[5], qz = H[6] * x + H[7] * y + H[8];
    const double uh = qx / qz, vh = qy / qz;
    const double px = Hi[0] * u + Hi[1] * v + Hi[2], py = Hi[3] * u + Hi[4] * v + Hi[5], pz = Hi[6] * u + Hi[7] * v + Hi[8];
    const double xh = px / pz, yh = py / pz;
    const double e1 = std::sqrt((u - uh) * (u - uh) + (v - vh) * (v - vh));
    const double e2 = std::sqrt((x - xh) * (x - xh) + (y - yh) * (y - yh));
    return std::sqrt(0.5 * (e1 * e1 + e2 * e2));
}

// has_near_collinear_triplet (homographyestimator.cpp:100-119), object coords
static bool degenerate(const double* x, const double* y, const int* s) {
    for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) for (int k = j + 1; k < 4; ++k) {
        const double ax = x[s[i]], ay = y[s[i]];
        const double area = std::fabs((x[s[j]] - ax) * (y[s[k]] - ay) - (y[s[j]] - ay) * (x[s[k]] - ax));
        if (area < 1e-6) return true;
    }
    return false;
}

// calculate_iterations (ransac.h:64-78)
static int calc_iters(double confidence, double w, int min_samples, int iters_so_far, int max_iters) {
    if (confidence <= 0.0 || w <= 0.0) return max_iters;
    const double denom = std::log(std::max(1e-12, 1.0 - std::pow(w, static_cast<double>(min_samples))));
    if (denom >= 0.0) return max_iters;
    const int niter = static_cast<int>(std::ceil(std::log(1.0 - confidence) / denom));
    return std::clamp(niter, iters_so_far, max_iters);
}

struct Scored { std::vector<int> idx; std::vector<double> res; };
static void find_inliers(int n, const double* x, const double* y, const double* u, const double* v, const double* H,
                         double thresh, Scored& out, double& margin) {
    double Hi[9]; mat3_inv(H, Hi);
    out.idx.clear(); out.res.clear();
    for (int i = 0; i < n; ++i) {
        const double r = transfer_error(H, Hi, x[i], y[i], u[i], v[i]);
        if (std::isfinite(r)) margin = std::min(margin, std::fabs(r - thresh));
        if (r <= thresh) { out.idx.push_back(i); out.res.push_back(r); }
    }
}
static double rms(const std::vector<double>& v) {
    if (v.empty()) return std::numeric_limits<double>::infinity();
    double ss = 0; for (double a : v) ss += a * a;
    return std::sqrt(ss / static_cast<double>(v.size()));
}

static void ransac_one(int n, const double* x, const double* y, const double* u, const double* v,
                       const orc_ransac_options& o, const int32_t* sample_idx, orc_ransac_result* res, uint8_t* mask) {
    std::memset(res, 0, sizeof *res);
    res->inlier_rms = std::numeric_limits<double>::infinity();
    res->min_margin = std::numeric_limits<double>::infinity();
    for (int i = 0; i < 9; ++i) res->hmtx[i] = (i % 4 == 0) ? 1.0 : 0.0;  // HomographyResult default Identity
    if (mask) std::memset(mask, 0, n);
    if (n < 4) return;
    MT64 rng(o.seed);
    int dyn = o.max_iters, it = 0;
    std::vector<int> best_inl; bool has_best = false; double best_rms = std::numeric_limits<double>::infinity();
    Scored a, b; int idxs[4];
    std::vector<double> sx(4), sy(4), su(4), sv(4);
    for (it = 0; it < dyn; ++it) {
        if (sample_idx) { for (int k = 0; k < 4; ++k) idxs[k] = sample_idx[4 * it + k]; }
        else sample_indices(rng, n, 4, idxs);
        if (degenerate(x, y, idxs)) continue;
        for (int k = 0; k < 4; ++k) { sx[k] = x[idxs[k]]; sy[k] = y[idxs[k]]; su[k] = u[idxs[k]]; sv[k] = v[idxs[k]]; }
        double H[9];
        if (!dlt_homography(sx, sy, su, sv, H)) continue;
        find_inliers(n, x, y, u, v, H, o.thresh, a, res->min_margin);
        if (static_cast<int>(a.idx.size()) < o.min_inliers) continue;
        double Hf[9]; std::memcpy(Hf, H, sizeof Hf);
        const Scored* fin = &a;
        if (o.refit_on_inliers) {
            b = a;
            // refit_model (ransac.h:97-111) / HomographyEstimator::refit (:146-166)
            if (a.idx.size() >= 4) {
                std::vector<double> rx, ry, ru, rv;
                for (int id : a.idx) { rx.push_back(x[id]); ry.push_back(y[id]); ru.push_back(u[id]); rv.push_back(v[id]); }
                double H2[9];
                if (dlt_homography(rx, ry, ru, rv, H2)) { std::memcpy(Hf, H2, sizeof Hf); find_inliers(n, x, y, u, v, Hf, o.thresh, b, res->min_margin); }
            }
            fin = &b;
        }
        const double frms = rms(fin->res);
        // is_better_model (ransac.h:113-117)
        if (!has_best || fin->idx.size() > best_inl.size() || (fin->idx.size() == best_inl.size() && frms < best_rms)) {
            has_best = true; best_inl = fin->idx; best_rms = frms;
            std::memcpy(res->hmtx, Hf, sizeof Hf); res->iters = it + 1;
        }
        const double ratio = static_cast<double>(fin->idx.size()) / static_cast<double>(n);
        dyn = calc_iters(o.confidence, ratio, 4, it + 1, o.max_iters);
    }
    res->iters_run = it;
    res->success = has_best ? 1 : 0;
    if (has_best) {
        res->n_inliers = static_cast<int>(best_inl.size());
        res->inlier_rms = best_rms;
        if (mask) for (int id : best_inl) mask[id] = 1;
        // symmetric_rms_px (optim/homography.cpp:18-28): sums the (root) residuals
        if (best_inl.empty()) res->symmetric_rms_px = std::numeric_limits<double>::infinity();
        else {
            double Hi[9]; mat3_inv(res->hmtx, Hi); double s = 0;
            for (int id : best_inl) s += transfer_error(res->hmtx, Hi, x[id], y[id], u[id], v[id]);
            res->symmetric_rms_px = std::sqrt(s / (2.0 * static_cast<double>(best_inl.size())));
        }
    }
}

}  // namespace orc
using namespace orc;

extern "C" {

void orc_sample_stream(uint64_t seed, int32_t n, int32_t iters, int32_t* out) {
    MT64 g(seed);
    for (int it = 0; it < iters; ++it) sample_indices(g, n, 4, out + 4 * it);
}

void orc_sample_stream_libstdcxx(uint64_t seed, int32_t n, int32_t iters, int32_t* out) {
    // exactly the calls of ransac.h:131-145
    std::vector<int> all(n), idxs(4);
    std::iota(all.begin(), all.end(), 0);
    std::mt19937_64 rng(seed);
    const size_t k_min_samples = 4;
    for (int it = 0; it < iters; ++it) {
        std::sample(all.begin(), all.end(), idxs.begin(), k_min_samples, rng);
        for (int k = 0; k < 4; ++k) out[4 * it + k] = idxs[k];
    }
}

int orc_ransac_homography(int32_t n, const double* x, const double* y, const double* u, const double* v,
                          const orc_ransac_options* o, const int32_t* sample_idx, orc_ransac_result* res,
                          uint8_t* inlier_mask) {
    ransac_one(n, x, y, u, v, *o, sample_idx, res, inlier_mask);
    return 0;
}

int orc_ransac_homography_batch(int64_t n_problems, int32_t n, const double* x, const double* y, const double* u,
                                const double* v, const orc_ransac_options* o, int seed_per_problem,
                                orc_ransac_result* res, uint8_t* inlier_mask, int num_threads) {
    const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int64_t p = 0; p < n_problems; ++p) {
        orc_ransac_options op = *o;
        if (seed_per_problem) op.seed = o->seed + static_cast<uint64_t>(p);
        ransac_one(n, x + p * n, y + p * n, u + p * n, v + p * n, op, nullptr, res + p, inlier_mask ? inlier_mask + p * n : nullptr);
    }
    return 0;
}

int orc_homography_dlt(int32_t n, const double* x, const double* y, const double* u, const double* v, double* hmtx) {
    std::vector<double> sx(x, x + n), sy(y, y + n), su(u, u + n), sv(v, v + n);
    return dlt_homography(sx, sy, su, sv, hmtx) ? 0 : 1;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// fit_plane_ransac (src/estimation/linear/planefit.cpp:9-62,86-104): the same ransac<> loop with the
// three-point plane estimator; refit = fit_plane_svd (:66-84).
// ---------------------------------------------------------------------------
namespace orc {

// the smallest right singular vector of the centred point matrix has no defined sign (the reference
// takes whatever Eigen::JacobiSVD's V holds, and its tests align signs before comparing,
// planefit_test.cpp:18-20).  Restatement and product both return the sign that makes the normal
// component of largest magnitude positive.  UNPINNED against Eigen.
static void canonical_sign(double* plane) {
    int k = 0;
    for (int i = 1; i < 3; ++i) if (std::fabs(plane[i]) > std::fabs(plane[k])) k = i;
    if (plane[k] < 0.0) for (int i = 0; i < 4; ++i) plane[i] = -plane[i];
}

// fit_plane_svd (planefit.cpp:66-84)
static bool fit_plane_svd(const std::vector<double>& px, const std::vector<double>& py, const std::vector<double>& pz, double* plane) {
    const size_t n = px.size();
    if (n < 3) return false;  // the reference throws std::invalid_argument
    double cx = 0, cy = 0, cz = 0;
    for (size_t i = 0; i < n; ++i) { cx += px[i]; cy += py[i]; cz += pz[i]; }
    const double dn = static_cast<double>(n);
    cx /= dn; cy /= dn; cz /= dn;
    std::vector<double> A(3 * n), V, sv;
    for (size_t i = 0; i < n; ++i) { A[3 * i] = px[i] - cx; A[3 * i + 1] = py[i] - cy; A[3 * i + 2] = pz[i] - cz; }
    jacobi_svd(A, static_cast<int>(n), 3, V, sv);
    int k = 0;  // JacobiSVD sorts descending: V.col(2) belongs to the smallest singular value
    for (int j = 1; j < 3; ++j) if (sv[j] < sv[k]) k = j;
    const double nx = V[0 * 3 + k], ny = V[1 * 3 + k], nz = V[2 * 3 + k];
    const double d = -(nx * cx + ny * cy + nz * cz);
    const double nrm = std::sqrt(nx * nx + ny * ny + nz * nz);
    plane[0] = nx / nrm; plane[1] = ny / nrm; plane[2] = nz / nrm; plane[3] = d / nrm;
    canonical_sign(plane);
    return true;
}

// PlaneRansacEstimator::fit / is_degenerate (planefit.cpp:14-33,52-62): both test the same cross product
static bool plane_from_sample(const double* x, const double* y, const double* z, const int* s, double* plane) {
    const double v1x = x[s[1]] - x[s[0]], v1y = y[s[1]] - y[s[0]], v1z = z[s[1]] - z[s[0]];
    const double v2x = x[s[2]] - x[s[0]], v2y = y[s[2]] - y[s[0]], v2z = z[s[2]] - z[s[0]];
    double nx = v1y * v2z - v1z * v2y, ny = v1z * v2x - v1x * v2z, nz = v1x * v2y - v1y * v2x;
    const double norm = std::sqrt(nx * nx + ny * ny + nz * nz);
    if (norm < 1e-12) return false;
    nx /= norm; ny /= norm; nz /= norm;
    plane[0] = nx; plane[1] = ny; plane[2] = nz; plane[3] = -(nx * x[s[0]] + ny * y[s[0]] + nz * z[s[0]]);
    return true;
}
// PlaneRansacEstimator::residual (planefit.cpp:35-38)
static inline double plane_residual(const double* pl, double x, double y, double z) {
    return std::fabs(pl[0] * x + pl[1] * y + pl[2] * z + pl[3]);
}
static void plane_inliers(int n, const double* x, const double* y, const double* z, const double* pl, double thresh, Scored& out,
                          double& margin) {
    out.idx.clear(); out.res.clear();
    for (int i = 0; i < n; ++i) {
        const double r = plane_residual(pl, x[i], y[i], z[i]);
        if (std::isfinite(r)) margin = std::min(margin, std::fabs(r - thresh));
        if (r <= thresh) { out.idx.push_back(i); out.res.push_back(r); }
    }
}

static void ransac_plane_one(int n, const double* x, const double* y, const double* z, const orc_ransac_options& o,
                             orc_plane_result* res, uint8_t* mask) {
    std::memset(res, 0, sizeof *res);  // PlaneRansacResult: plane = Zero (planefit.h:16)
    res->inlier_rms = std::numeric_limits<double>::infinity();
    res->min_margin = std::numeric_limits<double>::infinity();
    if (mask) std::memset(mask, 0, n);
    if (n < 3) return;
    MT64 rng(o.seed);
    int dyn = o.max_iters, it = 0;
    std::vector<int> best_inl; bool has_best = false; double best_rms = std::numeric_limits<double>::infinity();
    Scored a, b; int idxs[3];
    for (it = 0; it < dyn; ++it) {
        sample_indices(rng, n, 3, idxs);
        double P[4];
        if (!plane_from_sample(x, y, z, idxs, P)) continue;  // is_degenerate, and fit's own test
        plane_inliers(n, x, y, z, P, o.thresh, a, res->min_margin);
        if (static_cast<int>(a.idx.size()) < o.min_inliers) continue;
        double Pf[4]; std::memcpy(Pf, P, sizeof Pf);
        const Scored* fin = &a;
        if (o.refit_on_inliers) {
            b = a;
            if (a.idx.size() >= 3) {  // PlaneRansacEstimator::refit (planefit.cpp:40-50)
                std::vector<double> rx, ry, rz;
                for (int id : a.idx) { rx.push_back(x[id]); ry.push_back(y[id]); rz.push_back(z[id]); }
                double P2[4];
                if (fit_plane_svd(rx, ry, rz, P2)) { std::memcpy(Pf, P2, sizeof Pf); plane_inliers(n, x, y, z, Pf, o.thresh, b, res->min_margin); }
            }
            fin = &b;
        }
        const double frms = rms(fin->res);
        if (!has_best || fin->idx.size() > best_inl.size() || (fin->idx.size() == best_inl.size() && frms < best_rms)) {
            has_best = true; best_inl = fin->idx; best_rms = frms;
            std::memcpy(res->plane, Pf, sizeof Pf); res->iters = it + 1;
        }
        const double ratio = static_cast<double>(fin->idx.size()) / static_cast<double>(n);
        dyn = calc_iters(o.confidence, ratio, 3, it + 1, o.max_iters);
    }
    res->iters_run = it;
    res->success = has_best ? 1 : 0;
    if (has_best) {
        res->n_inliers = static_cast<int>(best_inl.size());
        res->inlier_rms = best_rms;
        if (mask) for (int id : best_inl) mask[id] = 1;
    }
}

}  // namespace orc

extern "C" {

void orc_sample_stream_k(uint64_t seed, int32_t n, int32_t k, int32_t iters, int32_t* out) {
    MT64 g(seed);
    for (int it = 0; it < iters; ++it) sample_indices(g, n, k, out + static_cast<size_t>(k) * it);
}
void orc_sample_stream_k_libstdcxx(uint64_t seed, int32_t n, int32_t k, int32_t iters, int32_t* out) {
    std::vector<int> all(n), idxs(k);
    std::iota(all.begin(), all.end(), 0);
    std::mt19937_64 rng(seed);
    for (int it = 0; it < iters; ++it) {
        std::sample(all.begin(), all.end(), idxs.begin(), static_cast<size_t>(k), rng);
        for (int j = 0; j < k; ++j) out[static_cast<size_t>(k) * it + j] = idxs[j];
    }
}

int orc_fit_plane_svd(int32_t n, const double* x, const double* y, const double* z, double* plane) {
    std::vector<double> px(x, x + n), py(y, y + n), pz(z, z + n);
    return fit_plane_svd(px, py, pz, plane) ? 0 : 1;
}

int orc_ransac_plane(int32_t n, const double* x, const double* y, const double* z, const orc_ransac_options* o,
                     orc_plane_result* res, uint8_t* inlier_mask) {
    ransac_plane_one(n, x, y, z, *o, res, inlier_mask);
    return 0;
}

int orc_ransac_plane_batch(int64_t n_problems, int32_t n, const double* x, const double* y, const double* z,
                           const orc_ransac_options* o, int seed_per_problem, orc_plane_result* res, uint8_t* inlier_mask,
                           int num_threads) {
    const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int64_t p = 0; p < n_problems; ++p) {
        orc_ransac_options op = *o;
        if (seed_per_problem) op.seed = o->seed + static_cast<uint64_t>(p);
        ransac_plane_one(n, x + p * n, y + p * n, z + p * n, op, res + p, inlier_mask ? inlier_mask + p * n : nullptr);
    }
    return 0;
}

}  // extern "C"
