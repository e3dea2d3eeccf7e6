// Compile-time work split of K1 (k1_fused.cu): which role (warp) of a tile owns which entry of the local
// system [J | r]^T [J | r], and the order in which the roles emit their per-tile values — the one piece
// of "protocol" the device epilogue and the host-side reduction (k_tile_final's value map) must agree on.
// Host-compilable (constexpr only), so tests/host_emul checks its invariants without a GPU.
#pragma once
#include <stdint.h>

#include <type_traits>
#include <vector>

#include "k1_math.cuh"

namespace calk {

template <int I, int N, class F>
CAL_HD void static_for(F&& f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

// ---------------------------------------------------------------------------
// entry -> (role, slot) tables
// ---------------------------------------------------------------------------
template <int MODEL, int IMODE>
struct K1Roles {
    using LT = Local<MODEL, IMODE>;
    static constexpr int NE = LT::NE, NL = LT::NL, NC = LT::NC, PI = LT::PI;
    static constexpr int NROLE = NE <= 72 ? 1 : (NE <= 144 ? 2 : 3);
    static constexpr int RR = LT::idx(NC, NC);              // |r|^2: accumulated by every role, owned by none
    struct Tbl {
        int role[NE];
        int slot[NE];
        int count[3];
        int col_role[PI > 0 ? PI : 1];  // role that owns the twist x intrinsic-column-j entries
    };
    static constexpr Tbl make() {
        Tbl t{};
        for (int e = 0; e < NE; ++e) { t.role[e] = -1; t.slot[e] = -1; }
        for (int r = 0; r < 3; ++r) t.count[r] = 0;
        const int cap = (NE - 1 + NROLE - 1) / NROLE;
        // twist-twist and twist-residual: role 0
        for (int a = 0; a < 6; ++a) {
            for (int b = a; b < 6; ++b) { const int e = LT::idx(a, b); t.role[e] = 0; t.slot[e] = t.count[0]++; }
            const int e = LT::idx(a, NC); t.role[e] = 0; t.slot[e] = t.count[0]++;
        }
        // twist x intrinsic column j: six entries, never split over roles
        for (int j = 0; j < PI; ++j) {
            int r = -1;
            for (int k = 0; k < NROLE; ++k) if (r < 0 && t.count[k] + 6 <= cap) r = k;
            if (r < 0) { r = 0; for (int k = 1; k < NROLE; ++k) if (t.count[k] < t.count[r]) r = k; }
            t.col_role[j] = r;
            for (int a = 0; a < 6; ++a) { const int e = LT::idx(a, 6 + j); t.role[e] = r; t.slot[e] = t.count[r]++; }
        }
        // everything else fills the roles up
        for (int a = 6; a < NL; ++a)
            for (int b = a; b < NL; ++b) {
                const int e = LT::idx(a, b);
                if (e == RR) continue;
                int r = -1;
                for (int k = 0; k < NROLE; ++k) if (r < 0 && t.count[k] < cap) r = k;
                if (r < 0) { r = 0; for (int k = 1; k < NROLE; ++k) if (t.count[k] < t.count[r]) r = k; }
                t.role[e] = r; t.slot[e] = t.count[r]++;
            }
        return t;
    }
    static constexpr Tbl tbl = make();
    static constexpr int count(int r) { return tbl.count[r]; }
    static constexpr int n_owned_cols(int r) { int n = 0; for (int j = 0; j < PI; ++j) if (tbl.col_role[j] == r) ++n; return n; }
    static constexpr int col_rank(int j) { int n = 0; for (int k = 0; k < j; ++k) if (tbl.col_role[k] == tbl.col_role[j]) ++n; return n; }
    static constexpr bool owns_cols(int r) { for (int j = 0; j < PI; ++j) if (tbl.col_role[j] == r) return true; return false; }

    // ---- order in which a role emits its per-tile values (device epilogue and host map agree on it) ----
    // cam part: the role's entries by slot; role 0 then rr and the cost; bundle view rows:
    // the role's E_vi columns (i major within a column), role 0 then g_v, H_vv, Q.
    static constexpr int n_vals(int r, bool view_rows) {
        int n = tbl.count[r] + (r == 0 ? 2 : 0);
        if (view_rows) {
            for (int j = 0; j < PI; ++j) if (tbl.col_role[j] == r) n += 6;
            if (r == 0) n += 6 + 21 + 36;
        }
        return n;
    }
    static constexpr int val_off(int r, bool view_rows) { int o = 0; for (int k = 0; k < r; ++k) o += n_vals(k, view_rows); return o; }
    static constexpr int nvt(bool view_rows) { return val_off(NROLE, view_rows); }
    // per-camera value index (cam_sums layout of refine_host.cu): [0, NE) local system | NE cost |
    // NE+1.. H_vv(21) g_v(6) Q(36) E_vi(6 PI)
    static void value_map(bool view_rows, std::vector<int32_t>& map) {
        map.assign(nvt(view_rows), -1);
        for (int r = 0; r < NROLE; ++r) {
            int o = val_off(r, view_rows);
            for (int sl = 0; sl < tbl.count[r]; ++sl)
                for (int e = 0; e < NE; ++e) if (tbl.role[e] == r && tbl.slot[e] == sl) map[o++] = e;
            if (r == 0) { map[o++] = RR; map[o++] = NE; }
            if (view_rows) {
                for (int j = 0; j < PI; ++j) if (tbl.col_role[j] == r) for (int i = 0; i < 6; ++i) map[o++] = NE + 64 + PI * i + j;
                if (r == 0) {
                    for (int i = 0; i < 6; ++i) map[o++] = NE + 22 + i;
                    for (int i = 0; i < 21; ++i) map[o++] = NE + 1 + i;
                    for (int i = 0; i < 36; ++i) map[o++] = NE + 28 + i;
                }
            }
        }
    }
};

}  // namespace calk
