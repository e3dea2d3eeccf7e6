// TEST INFRASTRUCTURE — invariants of K1's compile-time role tables (calibration_b200/csrc/k1_roles.hpp), g++ only.
#include <cstdio>
#include <set>
#include <vector>

#include "../../calibration_b200/csrc/k1_roles.hpp"

using namespace calk;

template <int MODEL, int IMODE>
int check(const char* name) {
    using RT = K1Roles<MODEL, IMODE>;
    using LT = Local<MODEL, IMODE>;
    int bad = 0;
    auto fail = [&](const char* what) { std::printf("%s: %s\n", name, what); ++bad; };
    // every entry except rr has exactly one owner and a dense, unique slot within its role
    std::vector<std::set<int>> slots(RT::NROLE);
    for (int e = 0; e < RT::NE; ++e) {
        if (e == RT::RR) { if (RT::tbl.role[e] != -1) fail("rr must be owned by no role"); continue; }
        const int r = RT::tbl.role[e];
        if (r < 0 || r >= RT::NROLE) { fail("entry without a role"); continue; }
        if (!slots[r].insert(RT::tbl.slot[e]).second) fail("duplicate slot");
    }
    int total = 0;
    for (int r = 0; r < RT::NROLE; ++r) {
        if ((int)slots[r].size() != RT::count(r)) fail("slot count mismatch");
        if (!slots[r].empty() && (*slots[r].begin() != 0 || *slots[r].rbegin() != RT::count(r) - 1)) fail("slots not dense");
        total += RT::count(r);
    }
    if (total != RT::NE - 1) fail("entries lost");
    // role-local epilogue: twist x twist and twist x residual with role 0, the six twist entries of an intrinsic column together
    for (int a = 0; a < 6; ++a) {
        for (int b = a; b < 6; ++b) if (RT::tbl.role[LT::idx(a, b)] != 0) fail("twist-twist entry outside role 0");
        if (RT::tbl.role[LT::idx(a, LT::NC)] != 0) fail("twist-residual entry outside role 0");
    }
    for (int j = 0; j < RT::PI; ++j)
        for (int a = 0; a < 6; ++a) if (RT::tbl.role[LT::idx(a, 6 + j)] != RT::tbl.col_role[j]) fail("twist x intrinsic column split over roles");
    // accumulators per role stay within the register budget the kernel was designed for
    for (int r = 0; r < RT::NROLE; ++r) if (RT::count(r) > 72) fail("more than 72 accumulators in one role");
    // the value map is a bijection onto the camera row it fills
    for (int view_rows = 0; view_rows < 2; ++view_rows) {
        std::vector<int32_t> map; RT::value_map(view_rows != 0, map);
        if ((int)map.size() != RT::nvt(view_rows != 0)) fail("value map size");
        std::set<int> seen(map.begin(), map.end());
        if (seen.size() != map.size()) fail("value map not injective");
        const int expect = RT::NE + 1 + (view_rows ? 63 + 6 * RT::PI : 0);
        if ((int)map.size() != expect) fail("value map does not cover the camera row");
        if (*seen.begin() != 0 || *seen.rbegin() != expect - 1) fail("value map range");
        int off = 0;
        for (int r = 0; r < RT::NROLE; ++r) { if (RT::val_off(r, view_rows != 0) != off) fail("role offsets"); off += RT::n_vals(r, view_rows != 0); }
    }
    return bad;
}

extern "C" int roles_check() {
    return check<0, 0>("pinhole / fixed intrinsics") + check<0, 1>("pinhole") + check<0, 2>("pinhole + skew") +
           check<1, 0>("scheimpflug / fixed intrinsics") + check<1, 1>("scheimpflug") + check<1, 2>("scheimpflug + skew");
}
extern "C" int roles_of(int model, int imode) {
    if (model == 0) return imode == 0 ? K1Roles<0, 0>::NROLE : (imode == 1 ? K1Roles<0, 1>::NROLE : K1Roles<0, 2>::NROLE);
    return imode == 0 ? K1Roles<1, 0>::NROLE : (imode == 1 ? K1Roles<1, 1>::NROLE : K1Roles<1, 2>::NROLE);
}
