// common.cuh — shared helpers for libdqn_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace dqnb {

// ---------------------------------------------------------------------------------------------
// error plumbing: every failure lands in a thread-local string read by dqnb_last_error()
// ---------------------------------------------------------------------------------------------
std::string &last_error();

#define DQNB_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
               cudaGetErrorString(_e));                                                     \
      ::dqnb::last_error() = _b;                                                            \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

#define DQNB_FAIL(...)                                                                      \
  do {                                                                                      \
    char _b[512];                                                                           \
    snprintf(_b, sizeof(_b), __VA_ARGS__);                                                  \
    ::dqnb::last_error() = _b;                                                              \
    return -1;                                                                              \
  } while (0)

// ---------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) split for 3xTF32: hi is x rounded to TF32 (10 explicit mantissa bits), so the
// tensor core reads it exactly; lo = x - hi is exact in fp32 and |lo| <= 2^-11 |x|.
// x == hi + lo exactly, so elementwise consumers rebuild x from the two planes.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  union { float f; uint32_t u; } v;
  v.f = x;
  v.u = (v.u + 0x1000u) & 0xFFFFE000u;
  return v.f;
#endif
}

constexpr int kActorOut = 10;   // dqn.hpp:28
constexpr int kActionSize = 4;  // dqn.hpp:20
constexpr int kMiscStride = 16; // replay misc row: act10[10], reward, mc_target, terminal, pad
constexpr float kNegSlope = 0.01f;  // dqn.cpp:300

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Philox4x32-10 counter-based generator (Salmon et al. 2011).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// index n of update `step`: uniform in [0, size) via 64-bit multiply-shift on a 32-bit draw
__host__ __device__ __forceinline__ int32_t sample_index(uint64_t seed, uint64_t step, uint32_t n,
                                                        uint32_t size) {
  uint32_t c[4] = {n, (uint32_t)step, (uint32_t)(step >> 32), 0x5eed5eedu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (int32_t)(((uint64_t)c[0] * (uint64_t)size) >> 32);
}

}  // namespace dqnb
