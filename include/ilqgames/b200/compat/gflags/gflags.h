// <gflags/gflags.h> for builds without gflags: DEFINE_* / DECLARE_* become plain globals (no
// command-line parsing), which is all src/air_3d_example.cpp needs from it (:62-66).  Add
// -I<repo>/include/ilqgames/b200/compat to use it.
#ifndef ILQGAMES_B200_COMPAT_GFLAGS_H
#define ILQGAMES_B200_COMPAT_GFLAGS_H
#include <cstdint>
#include <string>
#define DEFINE_double(name, val, txt) double FLAGS_##name = (val)
#define DEFINE_bool(name, val, txt) bool FLAGS_##name = (val)
#define DEFINE_int32(name, val, txt) int32_t FLAGS_##name = (val)
#define DEFINE_string(name, val, txt) std::string FLAGS_##name = (val)
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
#endif
