// Shared device helpers for libtaxo_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "taxo_b200.h"

namespace tx {

constexpr int kRowsPerBlock = 64;  // rows owned by one CTA in the partial-reduction kernels (tx_row_blocks)
constexpr int kMaxVocab = 8;       // position_vocab_size is 3 in every reference config (model_zoo.py:140,193)
constexpr int kNumSms = 148;       // B200

void set_error(const char* fmt, ...);

#define TX_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      tx::set_error(__VA_ARGS__);             \
      return TX_ERR_INVALID_ARGUMENT;         \
    }                                         \
  } while (0)

#define TX_LAUNCH_CHECK(name)                                                        \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      tx::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));          \
      return TX_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Counter-based dropout (Philox4x32-10). One call yields the keep decision of 4 consecutive indices.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t <= 0.0) return 0u;
  if (t >= 4294967295.0) return 4294967295u;
  return (uint32_t)t;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// random words of indices 4*idx4 .. 4*idx4+3
__device__ __forceinline__ uint4 drop_words(uint64_t seed, uint32_t stream_id, uint64_t idx4) {
  return philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), stream_id, 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ bool drop_keep1(uint64_t seed, uint32_t stream_id, uint64_t idx, uint32_t thr) {
  const uint4 w = drop_words(seed, stream_id, idx >> 2);
  const uint32_t sel = (uint32_t)(idx & 3);
  const uint32_t r = sel == 0 ? w.x : (sel == 1 ? w.y : (sel == 2 ? w.z : w.w));
  return r >= thr;
}

// ---------------------------------------------------------------------------------------------
// Small vector helpers: VEC = 4 (128-bit) or 1 (scalar) accesses.
// ---------------------------------------------------------------------------------------------
template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  // plain (coherent) load: for buffers written earlier by the same kernel
  __device__ __forceinline__ static Vec load_rw(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    Vec r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec load(const float* p) {
    Vec r;
    r.v[0] = __ldg(p);
    return r;
  }
  __device__ __forceinline__ static Vec load_rw(const float* p) {
    Vec r;
    r.v[0] = *p;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> vzero() {
  Vec<VEC> r;
#pragma unroll
  for (int t = 0; t < VEC; ++t) r.v[t] = 0.f;
  return r;
}

// keep flags of the VEC consecutive elements starting at idx (idx % VEC == 0)
template <int VEC>
__device__ __forceinline__ void drop_keep_vec(uint64_t seed, uint32_t stream_id, uint64_t idx, uint32_t thr,
                                              bool (&keep)[VEC]) {
  if constexpr (VEC == 4) {
    const uint4 w = drop_words(seed, stream_id, idx >> 2);
    keep[0] = w.x >= thr;
    keep[1] = w.y >= thr;
    keep[2] = w.z >= thr;
    keep[3] = w.w >= thr;
  } else {
    keep[0] = drop_keep1(seed, stream_id, idx, thr);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t row_blocks(int64_t n) { return (n + kRowsPerBlock - 1) / kRowsPerBlock; }
inline int grid_for_warps(int64_t n_warp_items, int warps_per_block, int blocks_per_sm) {
  int64_t need = (n_warp_items + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)kNumSms * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace tx
