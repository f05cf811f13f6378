// gflags/gflags.h — the gflags calls the reference makes (dqn_main.cpp:17-62 DEFINE_*, :392-394), on the
// flag table of shim/flags.hpp.
#pragma once
#include "../flags.hpp"
namespace gflags {
void SetUsageMessage(const std::string &usage);
void SetVersionString(const std::string &version);
const char *ProgramUsage();
}  // namespace gflags
