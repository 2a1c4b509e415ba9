// oracle/ref_shim/gflags/gflags.h -- TEST INFRASTRUCTURE.  DEFINE_/DECLARE_ flag macros as plain
// globals (no command-line parsing), for building oracle/_ref in an image without gflags.  A flag
// lives in the translation unit that DEFINEs it (several example sources of the reference define
// flags of the same name -- px0, py0, ... -- and only ever read their own).
#ifndef ILQG_REF_SHIM_GFLAGS
#define ILQG_REF_SHIM_GFLAGS
#include <cstdint>
#include <string>
#define DEFINE_double(name, val, txt) namespace { double FLAGS_##name = (val); }
#define DEFINE_bool(name, val, txt) namespace { bool FLAGS_##name = (val); }
#define DEFINE_int32(name, val, txt) namespace { int32_t FLAGS_##name = (val); }
#define DEFINE_int64(name, val, txt) namespace { int64_t FLAGS_##name = (val); }
#define DEFINE_uint64(name, val, txt) namespace { uint64_t FLAGS_##name = (val); }
#define DEFINE_string(name, val, txt) namespace { std::string FLAGS_##name = (val); }
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_int64(name) extern int64_t FLAGS_##name
#define DECLARE_uint64(name) extern uint64_t FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
namespace google { inline void ParseCommandLineFlags(int*, char***, bool) {} }
namespace gflags = google;
#endif
