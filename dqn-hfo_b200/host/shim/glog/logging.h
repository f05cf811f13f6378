// glog/logging.h — the glog calls the reference makes (LOG / VLOG / CHECK*, dqn_main.cpp:395-411 log set-up),
// on shim/logging.hpp.  Everything goes to stderr; the per-severity log files of SetLogDestination are not
// reproduced (their names are accepted and ignored).
#pragma once
#include "../logging.hpp"
namespace google {
enum LogSeverity { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
inline void InitGoogleLogging(const char *) {}
inline void InstallFailureSignalHandler() {}
inline void LogToStderr() {}
inline void SetLogDestination(int, const char *) {}
}  // namespace google
namespace fLI { extern int FLAGS_logbuflevel; }
using fLI::FLAGS_logbuflevel;
#ifndef CHECK_NOTNULL
#define CHECK_NOTNULL(p) ::shim::check_notnull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))
#endif
