// Columnar observation store: the on-disk side of the hot path's input.
//
// The reference reads detections as JSON (schemas/calib_dataset.schema.json, struct PlanarDetections,
// include/calib/pipeline/dataset.h:15-39) into an nlohmann DOM, then copies them point by point into
// AoS PlanarViews (collect_planar_views, src/pipeline/facades/intrinsics.cpp:38-59; make_planar_view,
// src/pipeline/detail/planar_utils.cpp:45-52).  At C5 scale that is 70 M JSON objects.  Here:
//   * cal_dataset_from_planar_json   one streaming pass over the JSON text (no DOM) that keeps only
//                                    images[].points[].{local_x, local_y, x, y} and writes the columns;
//   * cal_dataset_write / _open      a single file holding exactly the SoA + CSR arrays that
//                                    cal_problem_desc and cal_seed_* consume, 64-byte aligned, mmap-ed
//                                    read-only (optionally page-locked for full-speed H2D) — zero copies
//                                    and zero parsing between the disk and cudaMemcpyAsync.
// File layout (little endian): 64-byte header {magic "CALOBS01", n_views, n_obs, n_cams, 5 x offset},
// then view_offset[int64 n_views+1], view_cam[int32 n_views], obj_x, obj_y, img_u, img_v [f64 n_obs].
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../include/calib_b200.h"

extern "C" void cal_set_last_error_(const char* msg);

namespace {

cal_status dfail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; }

constexpr char kMagic[8] = {'C', 'A', 'L', 'O', 'B', 'S', '0', '1'};
struct Header {
    char magic[8];
    int64_t n_views, n_obs;
    int32_t n_cams, reserved;
    int64_t off_view_offset, off_view_cam, off_x;  // y, u, v follow x, each n_obs doubles (padded to 64 B)
    int64_t stride_obs;                            // bytes between the starts of consecutive observation columns
};
static_assert(sizeof(Header) == 64, "header is one cache line");

int64_t pad64(int64_t b) { return (b + 63) / 64 * 64; }

struct Impl { void* map = nullptr; size_t size = 0; bool pinned = false; };

// ---- minimal streaming JSON scanner (RFC 8259 subset sufficient for the dataset schema) ----
struct Scanner {
    const char* p; const char* end; std::string err;
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool lit(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    bool string(std::string* out) {
        ws();
        if (p >= end || *p != '"') return fail("expected a string");
        ++p;
        if (out) out->clear();
        while (p < end && *p != '"') {
            if (*p == '\\') { if (p + 1 >= end) return fail("bad escape"); if (out) out->push_back(p[1]); p += 2; }
            else { if (out) out->push_back(*p); ++p; }
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool number(double* out) {
        ws();
        // RFC 8259 grammar first (strtod alone would also take nan, inf, hex floats and a leading '+'):
        //   -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?
        const char* q = p;
        if (q < end && *q == '-') ++q;
        if (q >= end || *q < '0' || *q > '9') return fail("expected a number");
        if (*q == '0') ++q; else while (q < end && *q >= '0' && *q <= '9') ++q;
        if (q < end && *q == '.') { ++q; if (q >= end || *q < '0' || *q > '9') return fail("malformed number"); while (q < end && *q >= '0' && *q <= '9') ++q; }
        if (q < end && (*q == 'e' || *q == 'E')) {
            ++q; if (q < end && (*q == '+' || *q == '-')) ++q;
            if (q >= end || *q < '0' || *q > '9') return fail("malformed number");
            while (q < end && *q >= '0' && *q <= '9') ++q;
        }
        char* e = nullptr;
        const double v = std::strtod(p, &e);   // the buffer is NUL-terminated; the grammar above bounds what it reads
        if (e != q) return fail("malformed number");
        if (!std::isfinite(v)) return fail("number out of range");
        p = q; if (out) *out = v;
        return true;
    }
    static constexpr int kMaxDepth = 64;   // nesting of skipped values: bounded recursion on hostile input
    int depth = 0;
    bool skip_value() {  // any JSON value
        if (depth >= kMaxDepth) return fail("nesting too deep");
        struct Level { int& d; explicit Level(int& x) : d(x) { ++d; } ~Level() { --d; } } level(depth);
        ws();
        if (p >= end) return fail("unexpected end of input");
        if (*p == '"') return string(nullptr);
        if (*p == '{') {
            ++p;
            if (lit('}')) return true;
            do { if (!string(nullptr)) return false; if (!lit(':')) return fail("expected ':'"); if (!skip_value()) return false; } while (lit(','));
            return lit('}') || fail("expected '}'");
        }
        if (*p == '[') {
            ++p;
            if (lit(']')) return true;
            do { if (!skip_value()) return false; } while (lit(','));
            return lit(']') || fail("expected ']'");
        }
        if (!std::strncmp(p, "true", 4)) { p += 4; return true; }
        if (!std::strncmp(p, "false", 5)) { p += 5; return true; }
        if (!std::strncmp(p, "null", 4)) { p += 4; return true; }
        return number(nullptr);
    }
};

struct Columns { std::vector<int64_t> off{0}; std::vector<int32_t> cam; std::vector<double> x, y, u, v; };

// images[].points[] of one PlanarDetections document -> views of camera `cam`
bool scan_planar_detections(Scanner& s, int cam, int min_corners, Columns& c) {
    if (!s.lit('{')) return s.fail("dataset: expected an object");
    bool saw_images = false;
    if (s.lit('}')) return s.fail("dataset: 'images' is required");
    do {
        std::string key;
        if (!s.string(&key) || !s.lit(':')) return s.fail("dataset: malformed member");
        if (key != "images") { if (!s.skip_value()) return false; continue; }
        saw_images = true;
        if (!s.lit('[')) return s.fail("images: expected an array");
        if (s.lit(']')) continue;
        do {  // one image
            if (!s.lit('{')) return s.fail("image: expected an object");
            const size_t start = c.x.size();
            if (!s.lit('}')) {
                do {
                    if (!s.string(&key) || !s.lit(':')) return s.fail("image: malformed member");
                    if (key != "points") { if (!s.skip_value()) return false; continue; }
                    if (!s.lit('[')) return s.fail("points: expected an array");
                    if (s.lit(']')) continue;
                    do {  // one point
                        if (!s.lit('{')) return s.fail("point: expected an object");
                        double px = 0, py = 0, lx = 0, ly = 0; int have = 0;
                        if (!s.lit('}')) {
                            do {
                                if (!s.string(&key) || !s.lit(':')) return s.fail("point: malformed member");
                                double* dst = key == "x" ? &px : key == "y" ? &py : key == "local_x" ? &lx : key == "local_y" ? &ly : nullptr;
                                if (dst) { if (!s.number(dst)) return false; have |= key == "x" ? 1 : key == "y" ? 2 : key == "local_x" ? 4 : 8; }
                                else if (!s.skip_value()) return false;
                            } while (s.lit(','));
                            if (!s.lit('}')) return s.fail("point: expected '}'");
                        }
                        if (have != 15) return s.fail("point: x, y, local_x and local_y are required");
                        c.x.push_back(lx); c.y.push_back(ly); c.u.push_back(px); c.v.push_back(py);  // object_xy = local, image_uv = (x, y)
                    } while (s.lit(','));
                    if (!s.lit(']')) return s.fail("points: expected ']'");
                } while (s.lit(','));
                if (!s.lit('}')) return s.fail("image: expected '}'");
            }
            const size_t n = c.x.size() - start;
            if ((int64_t)n < (int64_t)min_corners) { c.x.resize(start); c.y.resize(start); c.u.resize(start); c.v.resize(start); }  // collect_planar_views :45-47
            else { c.off.push_back((int64_t)c.x.size()); c.cam.push_back(cam); }
        } while (s.lit(','));
        if (!s.lit(']')) return s.fail("images: expected ']'");
    } while (s.lit(','));
    if (!s.lit('}')) return s.fail("dataset: expected '}'");
    return saw_images || s.fail("dataset: 'images' is required");
}

}  // namespace

extern "C" cal_status cal_dataset_write(const char* path, int64_t n_views, int32_t n_cams, const int64_t* view_offset,
                                        const int32_t* view_cam, const double* x, const double* y, const double* u, const double* v) {
    if (!path || n_views < 0 || n_cams <= 0 || !view_offset || (n_views > 0 && !view_cam)) return dfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    const int64_t n_obs = view_offset[n_views];
    if (view_offset[0] != 0 || n_obs < 0 || (n_obs > 0 && (!x || !y || !u || !v))) return dfail(CAL_ERR_INVALID_ARGUMENT, "bad observation arrays");
    Header h{};
    std::memcpy(h.magic, kMagic, 8);
    h.n_views = n_views; h.n_obs = n_obs; h.n_cams = n_cams;
    h.off_view_offset = 64;
    h.off_view_cam = h.off_view_offset + pad64((n_views + 1) * 8);
    h.off_x = h.off_view_cam + pad64(n_views * 4);
    h.stride_obs = pad64(n_obs * 8);
    FILE* f = std::fopen(path, "wb");
    if (!f) return dfail(CAL_ERR_RUNTIME, std::string("cannot open ") + path + " for writing: " + std::strerror(errno));
    auto put = [&](const void* p, int64_t bytes, int64_t padded) {
        if (bytes > 0 && std::fwrite(p, 1, (size_t)bytes, f) != (size_t)bytes) return false;
        static const char zeros[64] = {0};
        return padded == bytes || std::fwrite(zeros, 1, (size_t)(padded - bytes), f) == (size_t)(padded - bytes);
    };
    bool ok = put(&h, 64, 64) && put(view_offset, (n_views + 1) * 8, pad64((n_views + 1) * 8)) && put(view_cam, n_views * 4, pad64(n_views * 4)) &&
              put(x, n_obs * 8, h.stride_obs) && put(y, n_obs * 8, h.stride_obs) && put(u, n_obs * 8, h.stride_obs) && put(v, n_obs * 8, h.stride_obs);
    ok = (std::fclose(f) == 0) && ok;
    return ok ? CAL_OK : dfail(CAL_ERR_RUNTIME, std::string("short write to ") + path);
}

extern "C" cal_status cal_dataset_open(const char* path, int pin, cal_dataset* out) {
    if (!path || !out) return dfail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(out, 0, sizeof *out);
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return dfail(CAL_ERR_RUNTIME, std::string("cannot open ") + path + ": " + std::strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 64) { ::close(fd); return dfail(CAL_ERR_RUNTIME, std::string(path) + ": not a CALOBS01 file"); }
    void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) return dfail(CAL_ERR_RUNTIME, std::string("mmap failed for ") + path);
    const Header* h = static_cast<const Header*>(map);
    // every section is checked against the file size with arithmetic that cannot overflow: counts are bounded by
    // the size before they are multiplied, sections must be ordered, aligned and non-overlapping
    const int64_t size = (int64_t)st.st_size;
    bool sane = !std::memcmp(h->magic, kMagic, 8) && h->n_cams > 0 && h->off_view_offset == 64 &&
                h->n_views >= 0 && h->n_views <= (size - 64) / 8 - 1 &&       // (n_views + 1) * 8 fits behind the header
                h->n_obs >= 0 && h->n_obs <= size / 32 &&                     // four columns of n_obs doubles fit
                h->stride_obs >= 0 && h->stride_obs <= size / 4 && h->stride_obs % 8 == 0 && h->stride_obs >= h->n_obs * 8;
    if (sane) {
        const int64_t end_off = h->off_view_offset + (h->n_views + 1) * 8, cam_bytes = h->n_views * 4;
        sane = h->off_view_cam >= end_off && h->off_view_cam % 4 == 0 && h->off_view_cam <= size - cam_bytes &&
               h->off_x >= h->off_view_cam + cam_bytes && h->off_x % 8 == 0 && h->off_x <= size - 4 * h->stride_obs &&
               3 * h->stride_obs + h->n_obs * 8 <= size - h->off_x;
    }
    if (!sane) { munmap(map, (size_t)st.st_size); return dfail(CAL_ERR_RUNTIME, std::string(path) + ": not a CALOBS01 file"); }
    const char* base = static_cast<const char*>(map);
    {   // the CSR offsets and camera ids are what the kernels index with: a corrupt file must not get past here
        const int64_t* vo = reinterpret_cast<const int64_t*>(base + h->off_view_offset);
        const int32_t* vc = reinterpret_cast<const int32_t*>(base + h->off_view_cam);
        bool ok = vo[0] == 0 && vo[h->n_views] == h->n_obs;
        for (int64_t v = 0; ok && v < h->n_views; ++v) ok = vo[v + 1] >= vo[v] && vc[v] >= 0 && vc[v] < h->n_cams;
        if (!ok) {
            munmap(map, (size_t)st.st_size);
            return dfail(CAL_ERR_RUNTIME, std::string(path) + ": view_offset does not span the observations monotonically or view_cam is out of range");
        }
    }
    out->n_views = h->n_views; out->n_obs = h->n_obs; out->n_cams = h->n_cams;
    out->view_offset = reinterpret_cast<const int64_t*>(base + h->off_view_offset);
    out->view_cam = reinterpret_cast<const int32_t*>(base + h->off_view_cam);
    out->obj_x = reinterpret_cast<const double*>(base + h->off_x);
    out->obj_y = reinterpret_cast<const double*>(base + h->off_x + h->stride_obs);
    out->img_u = reinterpret_cast<const double*>(base + h->off_x + 2 * h->stride_obs);
    out->img_v = reinterpret_cast<const double*>(base + h->off_x + 3 * h->stride_obs);
    Impl* im = new Impl; im->map = map; im->size = (size_t)st.st_size;
    if (pin) {  // page-lock the mapping so cudaMemcpyAsync runs at full PCIe speed; optional (needs a CUDA device)
        if (cudaHostRegister(map, im->size, cudaHostRegisterReadOnly) == cudaSuccess) im->pinned = true;
        else cudaGetLastError();
    }
    out->pinned = im->pinned ? 1 : 0;
    out->impl = im;
    return CAL_OK;
}

extern "C" void cal_dataset_close(cal_dataset* d) {
    if (!d || !d->impl) return;
    Impl* im = static_cast<Impl*>(d->impl);
    if (im->pinned) cudaHostUnregister(im->map);
    munmap(im->map, im->size);
    delete im;
    std::memset(d, 0, sizeof *d);
}

extern "C" cal_status cal_dataset_from_planar_json(const char* const* json_paths, int32_t n_cams, int32_t min_corners_per_view,
                                                   const char* out_path, int64_t* n_views_out, int64_t* n_obs_out) {
    if (!json_paths || n_cams <= 0 || !out_path) return dfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    Columns c;
    for (int cam = 0; cam < n_cams; ++cam) {
        if (!json_paths[cam]) return dfail(CAL_ERR_INVALID_ARGUMENT, "null path");
        const int fd = ::open(json_paths[cam], O_RDONLY);
        if (fd < 0) return dfail(CAL_ERR_RUNTIME, std::string("cannot open ") + json_paths[cam] + ": " + std::strerror(errno));
        struct stat st; fstat(fd, &st);
        // strtod needs a terminator: read into a buffer with a trailing NUL (one sequential read, no DOM)
        std::string text; text.resize((size_t)st.st_size);
        size_t got = 0;
        while (got < text.size()) { const ssize_t r = ::read(fd, &text[got], text.size() - got); if (r <= 0) break; got += (size_t)r; }
        ::close(fd);
        if (got != text.size()) return dfail(CAL_ERR_RUNTIME, std::string("short read from ") + json_paths[cam]);
        Scanner s{text.c_str(), text.c_str() + text.size(), {}};
        if (!scan_planar_detections(s, cam, min_corners_per_view, c)) {
            char where[64]; std::snprintf(where, sizeof where, " at byte %lld", (long long)(s.p - text.c_str()));
            return dfail(CAL_ERR_INVALID_ARGUMENT, std::string(json_paths[cam]) + ": " + s.err + where);
        }
        s.ws();
        if (s.p != s.end) return dfail(CAL_ERR_INVALID_ARGUMENT, std::string(json_paths[cam]) + ": trailing characters after the document");
    }
    const int64_t nv = (int64_t)c.cam.size();
    if (n_views_out) *n_views_out = nv;
    if (n_obs_out) *n_obs_out = (int64_t)c.x.size();
    return cal_dataset_write(out_path, nv, n_cams, c.off.data(), c.cam.data(), c.x.data(), c.y.data(), c.u.data(), c.v.data());
}
