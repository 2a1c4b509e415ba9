// <gflags/gflags.h> for builds without gflags: DEFINE_* / DECLARE_* become plain globals (no
// command-line parsing), which is all the example sources need from it (src/air_3d_example.cpp:62-66).
// A flag has internal linkage: several example sources define flags of the same name (px0, py0, ...)
// and only ever read their own, so they can share a binary; DECLARE_* across translation units is
// not supported.  Add
// -I<repo>/include/ilqgames/b200/compat to use it.
#ifndef ILQGAMES_B200_COMPAT_GFLAGS_H
#define ILQGAMES_B200_COMPAT_GFLAGS_H
#include <cstdint>
#include <string>
#define DEFINE_double(name, val, txt) namespace { double FLAGS_##name = (val); }
#define DEFINE_bool(name, val, txt) namespace { bool FLAGS_##name = (val); }
#define DEFINE_int32(name, val, txt) namespace { int32_t FLAGS_##name = (val); }
#define DEFINE_string(name, val, txt) namespace { std::string FLAGS_##name = (val); }
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
#endif
