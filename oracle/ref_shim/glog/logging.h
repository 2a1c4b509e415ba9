// oracle/ref_shim/glog/logging.h -- TEST INFRASTRUCTURE.  The CHECK / LOG / VLOG surface of glog
// that the reference sources use, for building oracle/_ref in an image without glog.  A failed
// CHECK prints "Check failed: <expr>" and aborts, like glog.
#ifndef ILQG_REF_SHIM_GLOG
#define ILQG_REF_SHIM_GLOG
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <gflags/gflags.h>  // glog built with gflags support pulls it in; some example sources rely on that

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void SetLogDestination(int, const char*) {}
struct FatalMessage {
  std::ostringstream os;
  FatalMessage(const char* file, int line, const char* what) { os << file << ":" << line << " Check failed: " << what << " "; }
  [[noreturn]] ~FatalMessage() { std::cerr << os.str() << std::endl; std::abort(); }
  std::ostream& stream() { return os; }
};
struct LogMessage {
  std::ostringstream os;
  bool on, fatal;
  LogMessage(bool enabled, bool is_fatal, const char* tag) : on(enabled), fatal(is_fatal) { os << tag << ": "; }
  ~LogMessage() { if (on || fatal) std::cerr << os.str() << std::endl; if (fatal) std::abort(); }
  std::ostream& stream() { return os; }
};
struct Voidify { void operator&(std::ostream&) {} };
template <typename T> T& CheckNotNull(const char* file, int line, const char* what, T& p) {
  if (p == nullptr) FatalMessage(file, line, what);
  return p;
}
template <typename T> T&& CheckNotNull(const char* file, int line, const char* what, T&& p) {
  if (p == nullptr) FatalMessage(file, line, what);
  return static_cast<T&&>(p);
}
namespace shim {
constexpr bool kINFO = false, kWARNING = false, kERROR = true, kFATAL = true;  // which severities print
constexpr bool fINFO = false, fWARNING = false, fERROR = false, fFATAL = true;
}  // namespace shim
}  // namespace google

static bool FLAGS_logtostderr = false;
static int FLAGS_minloglevel = 0;
static int FLAGS_v = 0;
#include <cmath>

#define ILQG_REF_CHECK(cond, text) \
  (cond) ? (void)0 : google::Voidify() & google::FatalMessage(__FILE__, __LINE__, text).stream()
#define CHECK(c) ILQG_REF_CHECK((c), #c)
#define CHECK_EQ(a, b) ILQG_REF_CHECK((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) ILQG_REF_CHECK((a) != (b), #a " != " #b)
#define CHECK_LT(a, b) ILQG_REF_CHECK((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) ILQG_REF_CHECK((a) <= (b), #a " <= " #b)
#define CHECK_GT(a, b) ILQG_REF_CHECK((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) ILQG_REF_CHECK((a) >= (b), #a " >= " #b)
#define CHECK_NEAR(a, b, tol) ILQG_REF_CHECK(std::abs((a) - (b)) <= (tol), #a " near " #b)
#define CHECK_NOTNULL(p) google::CheckNotNull(__FILE__, __LINE__, #p " != nullptr", (p))
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define LOG(sev) google::LogMessage(google::shim::k##sev, google::shim::f##sev, #sev).stream()
#define LOG_IF(sev, c) !(c) ? (void)0 : google::Voidify() & LOG(sev)
#define VLOG(n) true ? (void)0 : google::Voidify() & google::LogMessage(false, false, "V").stream()
#define VLOG_IF(n, c) VLOG(n)
#define LOG_FIRST_N(sev, n) LOG(sev)
#endif
