// <glog/logging.h> for builds without glog: the reference's sources and headers include it
// directly.  Add -I<repo>/include/ilqgames/b200/compat to use it; with the real glog installed
// leave this directory off the include path and define ILQGAMES_B200_USE_GLOG.
#ifndef ILQGAMES_B200_COMPAT_GLOG_LOGGING_H
#define ILQGAMES_B200_COMPAT_GLOG_LOGGING_H
#include <ilqgames/b200/log_shim.h>
#include <gflags/gflags.h>  // glog built with gflags support pulls it in; some example sources rely on that
#endif
