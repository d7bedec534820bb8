// Shared helpers for libcdcmdr (sm_100a).  No torch, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/cdcmdr.h"

namespace cdcmdr {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(const char* what, cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  return 1;
}
inline int fail_msg(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return 2;
}

#define CDC_CHECK(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return ::cdcmdr::fail(#call, e__, __FILE__, __LINE__); \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define CDC_LAUNCHED()                                                       \
  do {                                                                       \
    ::cdcmdr::g_launches.fetch_add(1, std::memory_order_relaxed);            \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) return ::cdcmdr::fail("kernel launch", e__, __FILE__, __LINE__); \
  } while (0)

#define CDC_REQUIRE(cond, msg) \
  do { if (!(cond)) return ::cdcmdr::fail_msg("cdcmdr: " msg " [" #cond "]"); } while (0)

constexpr int kNumSMs = 148;

inline cudaStream_t to_stream(cdcmdr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 64-bit mixer (splitmix64 finaliser): derives the per-step dropout seed from (base seed, step) in cdcmdr_step_tick.
__host__ __device__ __forceinline__ uint32_t mix_hash(uint64_t seed, uint32_t salt, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1) + ((uint64_t)salt << 32 | salt);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

// Stateless dropout mask (nn.Dropout, layer.py:189).  One 32-bit hash decides TWO neighbouring columns (16 bits each):
// element (row, col) is kept iff the (col & 1)-th half of drop_hash_pair(s0, row, col >> 1) is >= p * 2^16.  32-bit
// integer multiplies only, so the GEMM epilogues stay cheaper than the MMAs they overlap.  Every kernel that applies
// dropout uses these helpers; the backward never regenerates the mask (the stored activation is 0 where dropped).
__host__ __device__ __forceinline__ uint32_t drop_s0(uint64_t seed, uint32_t salt) {
  return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0xC2B2AE35u) ^ (salt * 0x27D4EB2Fu);
}
__host__ __device__ __forceinline__ uint32_t drop_mix32(uint32_t x) {           // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
constexpr uint32_t kDropRowMul = 0x9E3779B1u, kDropColMul = 0x85EBCA77u;
__host__ __device__ __forceinline__ uint32_t drop_hash_pair(uint32_t s0, uint32_t row, uint32_t colpair) {
  return drop_mix32(row * kDropRowMul + colpair * kDropColMul + s0);
}
__host__ __device__ __forceinline__ uint32_t drop_thr16(float p) {
  const float t = p * 65536.f + 0.5f;
  return t >= 65535.f ? 65535u : (t <= 0.f ? 0u : (uint32_t)t);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t s0, uint32_t thr16, uint32_t row, uint32_t col) {
  return ((drop_hash_pair(s0, row, col >> 1) >> ((col & 1u) * 16u)) & 0xFFFFu) >= thr16;
}
// The tensor-core GEMM epilogue's generator (3 instructions per element): one mixed 32-bit seed per (row, 32-column block),
// then a 32-bit LCG step per column; element kept iff state >= p * 2^32.
constexpr uint32_t kLcgA = 1664525u, kLcgC = 1013904223u;
__host__ __device__ __forceinline__ uint32_t drop_lcg_seed(uint32_t s0, uint32_t row, uint32_t colblock32) {
  return drop_mix32(row * kDropRowMul + colblock32 * kDropColMul + s0);
}
__host__ __device__ __forceinline__ uint32_t drop_thr32(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
}
__host__ __device__ __forceinline__ bool drop_keep_lcg(uint32_t s0, uint32_t thr32, uint32_t row, uint32_t col) {
  uint32_t x = drop_lcg_seed(s0, row, col >> 5);
  for (uint32_t j = 0; j <= (col & 31u); ++j) x = x * kLcgA + kLcgC;
  return x >= thr32;
}

// (row, col) of flat index i in a [*, C] row-major block.  64-bit div/mod costs ~100 instructions on the GPU; almost every
// index here fits 32 bits, where it is ~20.
__device__ __forceinline__ void split_idx(int64_t i, int64_t C, int64_t& r, int64_t& c) {
  if (((uint64_t)i | (uint64_t)C) <= 0xFFFFFFFFull) {
    const uint32_t q = (uint32_t)i / (uint32_t)C;
    r = q; c = (uint32_t)i - q * (uint32_t)C;
  } else {
    r = i / C; c = i - r * C;
  }
}

__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float(((uint32_t)h) << 16); }
__device__ __forceinline__ uint16_t f32_to_bf16(float f) {
  __nv_bfloat16 b = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);       // one cvt.rn.bf16x2.f32: .x (low half) = lo, .y = hi
  return *reinterpret_cast<const uint32_t*>(&b);
}

template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_act<uint16_t>(const uint16_t* p) { return bf16_to_f32(*p); }
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<uint16_t>(uint16_t* p, float v) { *p = f32_to_bf16(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cdcmdr
