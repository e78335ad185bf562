// TEST INFRASTRUCTURE — a ~100-line stand-in for googletest (absent from this
// image) so the reference's own *_test.cpp files can be compiled unmodified,
// both against the reference ops.cpp (pins the oracle) and against the CUDA
// drop-in host/ops_cuda.cpp (proves the ops.h boundary).  Supports exactly the
// macros those files use.  Prints gtest-like lines; exit code = #failed tests.
#ifndef LLMI_GTEST_SHIM_H
#define LLMI_GTEST_SHIM_H

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {

struct TestCase {
  const char* suite;
  const char* name;
  std::function<void()> fn;
};

inline std::vector<TestCase>& registry() {
  static std::vector<TestCase> r;
  return r;
}
inline int& current_failures() {
  static int f = 0;
  return f;
}

struct Registrar {
  Registrar(const char* s, const char* n, std::function<void()> f) {
    registry().push_back({s, n, std::move(f)});
  }
};

// Collects an optional `<< message` tail and reports on destruction.
class Failure {
 public:
  Failure(const char* file, int line, const std::string& what) {
    ++current_failures();
    os_ << file << ":" << line << ": Failure\n" << what;
  }
  ~Failure() { std::cout << os_.str() << std::endl; }
  template <typename T>
  Failure& operator<<(const T& v) {
    os_ << v;
    return *this;
  }

 private:
  std::ostringstream os_;
};

// `return Voidify() = Failure(...) << msg;` lets ASSERT_* abort the test body.
struct Voidify {
  void operator=(const Failure&) const {}
};

// Print a value if it is streamable, an enum as its integer, else a placeholder.
template <typename T, typename = void>
struct Printer {
  static void print(std::ostream& os, const T& v) {
    if constexpr (std::is_enum<T>::value) {
      os << static_cast<long long>(v);
    } else {
      (void)v;
      os << "<unprintable>";
    }
  }
};
template <typename T>
struct Printer<T, std::void_t<decltype(std::declval<std::ostream&>()
                                       << std::declval<const T&>())>> {
  static void print(std::ostream& os, const T& v) { os << v; }
};

template <typename A, typename B>
std::string fmt2(const char* ea, const char* eb, const A& a, const B& b,
                 const char* rel) {
  std::ostringstream os;
  os.precision(9);
  os << "Expected: (" << ea << ") " << rel << " (" << eb << "), actual: ";
  Printer<A>::print(os, a);
  os << " vs ";
  Printer<B>::print(os, b);
  os << "\n";
  return os.str();
}

inline bool float_eq_4ulp(float a, float b) {
  if (std::isnan(a) || std::isnan(b)) return false;
  if (a == b) return true;
  int32_t ia, ib;
  memcpy(&ia, &a, 4);
  memcpy(&ib, &b, 4);
  if (ia < 0) ia = int32_t(0x80000000u) - ia;
  if (ib < 0) ib = int32_t(0x80000000u) - ib;
  int64_t d = int64_t(ia) - int64_t(ib);
  return d <= 4 && d >= -4;
}

inline int RunAllTests() {
  int failed = 0;
  for (auto& t : registry()) {
    std::cout << "[ RUN      ] " << t.suite << "." << t.name << std::endl;
    current_failures() = 0;
    try {
      t.fn();
    } catch (const std::exception& e) {
      ++current_failures();
      std::cout << "unexpected exception: " << e.what() << std::endl;
    }
    if (current_failures()) {
      ++failed;
      std::cout << "[  FAILED  ] " << t.suite << "." << t.name << std::endl;
    } else {
      std::cout << "[       OK ] " << t.suite << "." << t.name << std::endl;
    }
  }
  std::cout << "[==========] " << registry().size() << " tests ran, " << failed
            << " failed." << std::endl;
  return failed;
}

}  // namespace testing

#define TEST(suite, name)                                                  \
  static void suite##_##name##_body();                                     \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name,          \
                                                   suite##_##name##_body); \
  static void suite##_##name##_body()

#define LLMI_FAIL_FATAL_(what) \
  return ::testing::Voidify() = ::testing::Failure(__FILE__, __LINE__, what)
#define LLMI_FAIL_SOFT_(what) \
  ::testing::Voidify() = ::testing::Failure(__FILE__, __LINE__, what)

// `switch (0) case 0: default:` guards against dangling-else at the use site.
#define LLMI_EXPECT_(ok, a, b, rel) \
  switch (0)                        \
  case 0:                           \
  default:                          \
    if (ok)                         \
      ;                             \
    else                            \
      LLMI_FAIL_SOFT_(::testing::fmt2(#a, #b, (a), (b), rel))
#define LLMI_ASSERT_(ok, a, b, rel) \
  switch (0)                        \
  case 0:                           \
  default:                          \
    if (ok)                         \
      ;                             \
    else                            \
      LLMI_FAIL_FATAL_(::testing::fmt2(#a, #b, (a), (b), rel))

#define EXPECT_EQ(a, b) LLMI_EXPECT_((a) == (b), a, b, "==")
#define ASSERT_EQ(a, b) LLMI_ASSERT_((a) == (b), a, b, "==")
#define EXPECT_NE(a, b) LLMI_EXPECT_((a) != (b), a, b, "!=")
#define ASSERT_NE(a, b) LLMI_ASSERT_((a) != (b), a, b, "!=")
#define EXPECT_NEAR(a, b, tol) \
  LLMI_EXPECT_(std::fabs(double(a) - double(b)) <= double(tol), a, b, "~=")
#define ASSERT_NEAR(a, b, tol) \
  LLMI_ASSERT_(std::fabs(double(a) - double(b)) <= double(tol), a, b, "~=")
#define EXPECT_FLOAT_EQ(a, b) \
  LLMI_EXPECT_(::testing::float_eq_4ulp((a), (b)), a, b, "=f=")
#define ASSERT_FLOAT_EQ(a, b) \
  LLMI_ASSERT_(::testing::float_eq_4ulp((a), (b)), a, b, "=f=")
#define EXPECT_STREQ(a, b)                                            \
  LLMI_EXPECT_(std::strcmp((a), (b)) == 0, std::string(a), std::string(b), \
               "==")

#endif  // LLMI_GTEST_SHIM_H
