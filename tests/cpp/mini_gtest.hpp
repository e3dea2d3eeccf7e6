// A few lines of test harness with GoogleTest's spelling (GoogleTest is absent from this image), so the
// C++ host tests read like the reference's tests/unit/*.cpp.  TEST INFRASTRUCTURE.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

namespace mini_gtest {
struct Case { std::string suite, name; std::function<void()> fn; };
inline std::vector<Case>& cases() { static std::vector<Case> c; return c; }
inline int& failures() { static int f = 0; return f; }
struct Registrar { Registrar(const char* s, const char* n, std::function<void()> f) { cases().push_back({s, n, std::move(f)}); } };
struct Abort {};
template <class T>
inline std::string str(const T& v) {
    char b[64];
    if constexpr (std::is_floating_point_v<T>) std::snprintf(b, sizeof b, "%.17g", static_cast<double>(v));
    else std::snprintf(b, sizeof b, "%lld", static_cast<long long>(v));
    return b;
}
inline void report(const char* file, int line, const std::string& what) {
    ++failures();
    std::printf("%s:%d: Failure\n  %s\n", file, line, what.c_str());
}
// run every test whose "Suite.Name" contains one of the filters (all when none); `-` prefix excludes
inline int run(int argc, char** argv) {
    std::vector<std::string> inc, exc;
    bool list = false;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--list")) list = true;
        else if (argv[i][0] == '-' && argv[i][1] != '-') exc.emplace_back(argv[i] + 1);
        else inc.emplace_back(argv[i]);
    }
    int ran = 0, failed = 0;
    std::string failed_names;
    for (auto& c : cases()) {
        const std::string full = c.suite + "." + c.name;
        bool sel = inc.empty();
        for (auto& f : inc) sel = sel || full.find(f) != std::string::npos;
        for (auto& f : exc) sel = sel && full.find(f) == std::string::npos;
        if (!sel) continue;
        if (list) { std::printf("%s\n", full.c_str()); continue; }
        const int before = failures();
        std::printf("[ RUN      ] %s\n", full.c_str());
        try { c.fn(); } catch (const Abort&) {} catch (const std::exception& e) { report("<exception>", 0, std::string("uncaught: ") + e.what()); }
        const bool ok = failures() == before;
        std::printf("%s %s\n", ok ? "[       OK ]" : "[  FAILED  ]", full.c_str());
        std::fflush(stdout);
        ++ran; failed += ok ? 0 : 1;
        if (!ok) failed_names += " " + full;
    }
    if (!list) std::printf("[==========] %d tests ran, %d failed%s%s\n", ran, failed, failed ? ":" : "", failed_names.c_str());
    return failed ? 1 : 0;
}
}  // namespace mini_gtest

#define TEST(suite, name)                                                                        \
    static void suite##_##name##_body();                                                         \
    static mini_gtest::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);     \
    static void suite##_##name##_body()
#define MG_STR(x) #x
#define MG_CHECK(cond, text, fatal)                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            mini_gtest::report(__FILE__, __LINE__, text);                    \
            if (fatal) throw mini_gtest::Abort{};                            \
        }                                                                    \
    } while (0)
#define MG_CMP(a, op, b, fatal)                                                                                         \
    do {                                                                                                                \
        const auto mg_a = (a);                                                                                          \
        const auto mg_b = (b);                                                                                          \
        if (!(mg_a op mg_b)) {                                                                                          \
            mini_gtest::report(__FILE__, __LINE__, std::string(#a " " #op " " #b "  with ") + mini_gtest::str(mg_a) + " vs " + mini_gtest::str(mg_b)); \
            if (fatal) throw mini_gtest::Abort{};                                                                       \
        }                                                                                                               \
    } while (0)
#define EXPECT_TRUE(c) MG_CHECK((c), "expected true: " #c, false)
#define EXPECT_FALSE(c) MG_CHECK(!(c), "expected false: " #c, false)
#define ASSERT_TRUE(c) MG_CHECK((c), "expected true: " #c, true)
#define ASSERT_FALSE(c) MG_CHECK(!(c), "expected false: " #c, true)
#define EXPECT_LT(a, b) MG_CMP(a, <, b, false)
#define EXPECT_LE(a, b) MG_CMP(a, <=, b, false)
#define EXPECT_GT(a, b) MG_CMP(a, >, b, false)
#define EXPECT_GE(a, b) MG_CMP(a, >=, b, false)
#define EXPECT_EQ(a, b) MG_CMP(a, ==, b, false)
#define ASSERT_EQ(a, b) MG_CMP(a, ==, b, true)
#define EXPECT_NEAR(a, b, tol)                                                                                           \
    do {                                                                                                                 \
        const double mg_a = (a), mg_b = (b), mg_t = (tol);                                                               \
        if (!(std::fabs(mg_a - mg_b) <= mg_t))                                                                           \
            mini_gtest::report(__FILE__, __LINE__, std::string("|" #a " - " #b "| <= " #tol "  with ") + mini_gtest::str(mg_a) + " vs " + mini_gtest::str(mg_b) + " diff " + mini_gtest::str(mg_a - mg_b)); \
    } while (0)
#define EXPECT_THROW(stmt, exc)                                                                  \
    do {                                                                                         \
        bool mg_ok = false;                                                                      \
        try { stmt; } catch (const exc&) { mg_ok = true; } catch (...) {}                        \
        if (!mg_ok) mini_gtest::report(__FILE__, __LINE__, "expected " #stmt " to throw " #exc); \
    } while (0)
