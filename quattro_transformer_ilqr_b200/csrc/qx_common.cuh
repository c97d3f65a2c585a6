// Shared device/host helpers for libquattro_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/quattro_b200.h"

namespace qx {

// ---- error plumbing (thread-local text, C ABI returns a code) -------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define QX_CHECK_ARG(cond, msg)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      qx::set_error("invalid argument: %s", msg);       \
      return QX_EINVAL;                                 \
    }                                                   \
  } while (0)

#define QX_CHECK_LAUNCH(what)                           \
  do {                                                  \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) return qx::cuda_fail(e__, what); \
  } while (0)

// ---- record layout (see include/quattro_b200.h) ----------------------------------------------
template <int n, int m>
struct Rec {
  // F = [A | B] interleaved row by row: row i holds A[i][0..n) then B[i][0..m)  (pitch n + m)
  static constexpr int F = 0;
  static constexpr int pitch = n + m;
  __host__ __device__ static constexpr int a(int i, int j) { return i * pitch + j; }
  __host__ __device__ static constexpr int b(int i, int j) { return i * pitch + n + j; }
  static constexpr int lx = n * pitch;          // [lx | lu] contiguous
  static constexpr int lu = lx + n;
  static constexpr int lxx = lu + m;
  static constexpr int luu = lxx + n * n;
  static constexpr int lux = luu + m * m;
  static constexpr int size = lux + m * n;      // 416 (quadrotor), 46 (cart-pole)
  static constexpr int term = n + n * n;        // [Vx | Vxx]
  static constexpr int gain = m + m * n;        // [k | K]
};

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One-time launch setup (cudaFuncSetAttribute, SM count) is PER DEVICE and a process may drive several: the caches
// that remember it are arrays indexed by the current device.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  cudaGetDevice(&d);
  return (d < 0 || d >= kMaxDevices) ? 0 : d;
}
inline int device_sm_count() {
  static int sms[kMaxDevices] = {};
  const int d = current_device_slot();
  if (sms[d] == 0) cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
  return sms[d];
}

// Gains are stored with the trajectory index fastest, so that a warp of consecutive trajectories touches whole lines:
//   QX_LAYOUT_SOA (cart-pole)  float  [N][G][B]
//   QX_LAYOUT_REC (quadrotor)  float4 [N][G/4][B]   G = 52: float4 #0 = k, #1..#12 = K row-major, four entries each
// (one 16-B load per thread fetches four entries and a warp's load is 512 contiguous bytes).
template <int LAYOUT>
__host__ __device__ __forceinline__ size_t gain_index(int t, int ch, int b, int B, int G) {
  return (LAYOUT == QX_LAYOUT_REC) ? ((((size_t)t * (G >> 2) + (ch >> 2)) * B + b) << 2) + (ch & 3)
                                   : ((size_t)t * G + ch) * B + b;
}

static __constant__ double kAlphas[QX_NUM_ALPHA] = {1.0, 0.5, 0.25, 0.1, 0.05, 0.01};

}  // namespace qx
