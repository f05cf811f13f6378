// flags.hpp — the handful of gflags macros the reference uses (DEFINE_*/DECLARE_*/FLAGS_*,
// ParseCommandLineFlags), because gflags is not in this image.  Syntax accepted: -name=value,
// --name=value, -name value, -boolname / -noboolname.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <string>

namespace shim {
struct FlagInfo { char type; void *ptr; const char *help; };
inline std::map<std::string, FlagInfo> &flag_registry() { static std::map<std::string, FlagInfo> r; return r; }
struct FlagRegisterer {
  FlagRegisterer(const char *name, char type, void *ptr, const char *help) { flag_registry()[name] = FlagInfo{type, ptr, help}; }
};
inline bool set_flag(const std::string &name, const std::string &value) {
  auto it = flag_registry().find(name);
  if (it == flag_registry().end()) return false;
  switch (it->second.type) {
    case 'i': *static_cast<int32_t *>(it->second.ptr) = (int32_t)std::strtol(value.c_str(), nullptr, 10); break;
    case 'd': *static_cast<double *>(it->second.ptr) = std::strtod(value.c_str(), nullptr); break;
    case 'b': *static_cast<bool *>(it->second.ptr) = !(value == "false" || value == "0" || value == "no"); break;
    case 's': *static_cast<std::string *>(it->second.ptr) = value; break;
  }
  return true;
}
// returns the number of arguments consumed as flags; unknown flags abort like gflags does
int ParseCommandLineFlags(int *argc, char ***argv, bool remove_flags);
int int_flag_or(const char *name, int dflt);
}  // namespace shim

namespace gflags { using shim::ParseCommandLineFlags; }
namespace google { using shim::ParseCommandLineFlags; }

#define DEFINE_int32(name, def, help) int32_t FLAGS_##name = def; static ::shim::FlagRegisterer flagreg_##name(#name, 'i', &FLAGS_##name, help)
#define DEFINE_double(name, def, help) double FLAGS_##name = def; static ::shim::FlagRegisterer flagreg_##name(#name, 'd', &FLAGS_##name, help)
#define DEFINE_bool(name, def, help) bool FLAGS_##name = def; static ::shim::FlagRegisterer flagreg_##name(#name, 'b', &FLAGS_##name, help)
#define DEFINE_string(name, def, help) std::string FLAGS_##name = def; static ::shim::FlagRegisterer flagreg_##name(#name, 's', &FLAGS_##name, help)
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
