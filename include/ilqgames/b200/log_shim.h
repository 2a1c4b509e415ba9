// log_shim.h -- glog's CHECK / LOG / VLOG surface for builds without glog.  A failed CHECK prints
// "Check failed: <expr>" to stderr and aborts, which is what the reference's death tests match on.
#ifndef ILQGAMES_B200_LOG_SHIM_H
#define ILQGAMES_B200_LOG_SHIM_H

#ifdef ILQGAMES_B200_USE_GLOG
#include <glog/logging.h>
#else

#include <cstdlib>
#include <iostream>
#include <sstream>

namespace ilqgames_b200_log {
struct Fatal {
  std::ostringstream os;
  Fatal(const char* file, int line, const char* what) { os << file << ":" << line << " Check failed: " << what << " "; }
  [[noreturn]] ~Fatal() { std::cerr << os.str() << std::endl; std::abort(); }
  template <typename T> Fatal& operator<<(const T& v) { os << v; return *this; }
  Fatal& self() { return *this; }
};
struct Sink {
  bool on;
  explicit Sink(bool enabled) : on(enabled) {}
  ~Sink() { if (on) std::cerr << std::endl; }
  template <typename T> Sink& operator<<(const T& v) { if (on) std::cerr << v; return *this; }
};
struct Voidify { void operator&(Fatal&) {} };
}  // namespace ilqgames_b200_log

#define ILQG_CHECK_IMPL(cond, text) \
  (cond) ? (void)0 : ilqgames_b200_log::Voidify() & ilqgames_b200_log::Fatal(__FILE__, __LINE__, text).self()
#define CHECK(c) ILQG_CHECK_IMPL((c), #c)
#define CHECK_EQ(a, b) ILQG_CHECK_IMPL((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) ILQG_CHECK_IMPL((a) != (b), #a " != " #b)
#define CHECK_LT(a, b) ILQG_CHECK_IMPL((a) < (b), #a " < " #b)
#define CHECK_LE(a, b) ILQG_CHECK_IMPL((a) <= (b), #a " <= " #b)
#define CHECK_GT(a, b) ILQG_CHECK_IMPL((a) > (b), #a " > " #b)
#define CHECK_GE(a, b) ILQG_CHECK_IMPL((a) >= (b), #a " >= " #b)
#define CHECK_NOTNULL(p) ILQG_CHECK_IMPL((p) != nullptr, #p " != nullptr")
#define LOG(severity) ilqgames_b200_log::Sink(true) << #severity ": "
#define VLOG(n) ilqgames_b200_log::Sink(false)
#define VLOG_IF(n, c) ilqgames_b200_log::Sink(false)
#define LOG_IF(severity, c) ilqgames_b200_log::Sink((c))

#endif  // ILQGAMES_B200_USE_GLOG
#endif
