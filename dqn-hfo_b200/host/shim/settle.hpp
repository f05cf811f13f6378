// settle.hpp — force-included (-include) only when the reference's dqn_main.cpp is compiled: its main() sleeps
// 10 s after starting each agent thread (dqn_main.cpp:425) so that rcssserver sees the players connect in order.
// The in-process environment has no server to wait for, so that wait is a no-op here.  The headers that declare or
// use ::sleep are pulled in first, so only the caller's own call sites see the macro.
#pragma once
#include <unistd.h>
#include <chrono>
#include <thread>
namespace shim {
inline unsigned server_settle(unsigned) { return 0; }
}  // namespace shim
#define sleep(s) ::shim::server_settle(s)
