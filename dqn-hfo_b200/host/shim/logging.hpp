// logging.hpp — glog look-alikes used by the reference (LOG, VLOG, DLOG, CHECK*).  LOG(FATAL) and
// failed CHECKs abort the process: the reference's error convention is kept.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <utility>

namespace shim {
extern int g_vlog_level;
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class LogMessage {
 public:
  LogMessage(const char *file, int line, int sev) : sev_(sev) {
    static const char tag[] = {'I', 'W', 'E', 'F'};
    const char *base = file;
    for (const char *p = file; *p; ++p) if (*p == '/') base = p + 1;
    os_ << tag[sev] << " " << base << ":" << line << "] ";
  }
  ~LogMessage() {
    os_ << "\n";
    std::cerr << os_.str();
    if (sev_ == FATAL) { std::cerr.flush(); std::abort(); }
  }
  std::ostream &stream() { return os_; }
 private:
  std::ostringstream os_;
  int sev_;
};
struct Voidify { void operator&(std::ostream &) {} };
template <typename T>
T check_notnull(const char *file, int line, const char *what, T &&p) {   // glog's CHECK_NOTNULL: returns its argument
  if (p == nullptr) LogMessage(file, line, FATAL).stream() << "Check failed: " << what;
  return std::forward<T>(p);
}
}  // namespace shim

#define LOG(sev) ::shim::LogMessage(__FILE__, __LINE__, ::shim::sev).stream()
#define VLOG(n) !((n) <= ::shim::g_vlog_level) ? (void)0 : ::shim::Voidify() & LOG(INFO)
#ifdef NDEBUG
#define DLOG(sev) true ? (void)0 : ::shim::Voidify() & LOG(sev)
#else
#define DLOG(sev) LOG(sev)
#endif
#define CHECK(cond) (cond) ? (void)0 : ::shim::Voidify() & LOG(FATAL) << "Check failed: " #cond " "
#define CHECK_OP(a, b, op) ((a) op (b)) ? (void)0 : ::shim::Voidify() & LOG(FATAL) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs " << (b) << ") "
#define CHECK_EQ(a, b) CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP(a, b, <)
#define CHECK_GE(a, b) CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP(a, b, >)
