// TEST INFRASTRUCTURE: a minimal stand-in for the five googletest facilities the reference's unit test uses
// (TEST, EXPECT_EQ, EXPECT_DOUBLE_EQ with googletest's 4-ULP rule, InitGoogleTest, RUN_ALL_TESTS).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace testing {
struct Registry {
    std::vector<std::pair<std::string, std::function<void()>>> tests;
    int failures = 0;
    static Registry& get()
    {
	static Registry r;
	return r;
    }
};
struct Registrar {
    Registrar(const char* name, std::function<void()> fn) { Registry::get().tests.emplace_back(name, fn); }
};
inline void InitGoogleTest(int*, char**) {}
inline bool almost_equal(double a, double b)
{
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](double x) {
	uint64_t u;
	std::memcpy(&u, &x, 8);
	return (u >> 63) ? ~u + 1 : u | (uint64_t(1) << 63);
    };
    const uint64_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}
}  // namespace testing

#define TEST(suite, name)                                                                      \
    static void suite##_##name##_body();                                                       \
    static ::testing::Registrar suite##_##name##_reg(#suite "." #name, suite##_##name##_body); \
    static void suite##_##name##_body()
#define EXPECT_EQ(a, b)                                                            \
    do {                                                                           \
	if (!((a) == (b))) {                                                       \
	    ::testing::Registry::get().failures++;                                 \
	    std::printf("  FAILED %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); \
	}                                                                          \
    } while (0)
#define EXPECT_DOUBLE_EQ(a, b)                                                                              \
    do {                                                                                                    \
	const double va_ = (a), vb_ = (b);                                                                  \
	if (!::testing::almost_equal(va_, vb_)) {                                                           \
	    ::testing::Registry::get().failures++;                                                          \
	    std::printf("  FAILED %s:%d: %s (%.17g) ~ %s (%.17g)\n", __FILE__, __LINE__, #a, va_, #b, vb_); \
	}                                                                                                   \
    } while (0)
inline int RUN_ALL_TESTS()
{
    auto& r = ::testing::Registry::get();
    int bad = 0;
    for (auto& t : r.tests) {
	const int before = r.failures;
	std::printf("[ RUN      ] %s\n", t.first.c_str());
	try {
	    t.second();
	}
	catch (const std::exception& e) {
	    r.failures++;
	    std::printf("  EXCEPTION %s\n", e.what());
	}
	const bool ok = r.failures == before;
	bad += !ok;
	std::printf("[ %s ] %s\n", ok ? "      OK" : " FAILED ", t.first.c_str());
    }
    std::printf("%zu tests, %d failed\n", r.tests.size(), bad);
    return bad ? 1 : 0;
}
