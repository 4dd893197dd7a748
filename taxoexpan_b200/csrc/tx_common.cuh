// Shared device helpers for libtaxo_sm100 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
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
// Counter-based dropout.  keep(seed, stream, index) is a pure function of its arguments (reproducible in the
// backward pass, by tx_dropout_keep_mask and across kernels): a 4-round Philox-2x32-style network on (index / 4) keyed by
// (seed, stream) yields four 16-bit uniforms, one per element of an aligned group of 4; keep iff u16 >= round(p * 65536).
// (A Philox4x32-10 draw cost ~100 integer instructions per group and made the fused aggregate kernel issue-bound.)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t <= 0.5) return 0u;
  if (t >= 65535.0) return 65535u;
  return (uint32_t)t;
}

// returns {w.x: u16 of elements 0 (low half) and 1 (high half), w.y: elements 2 and 3} of group idx4.
// Four rounds of the Philox-2x32 multiply/xor network on the 64-bit counter idx4 keyed by (seed, stream): each round is one
// 32x32->64 multiply (IMAD.WIDE) and one 3-input xor (LOP3) - ~10 integer instructions per group of 4 elements.
__device__ __forceinline__ uint2 drop_words(uint64_t seed, uint32_t stream_id, uint64_t idx4) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (stream_id * 0x9E3779B1u);
  uint32_t x = (uint32_t)idx4, y = (uint32_t)(idx4 >> 32) ^ 0x85EBCA77u;
  uint64_t p;
  p = (uint64_t)x * 0xD2511F53u; x = (uint32_t)(p >> 32) ^ y ^ k0; y = (uint32_t)p;
  p = (uint64_t)x * 0xCD9E8D57u; x = (uint32_t)(p >> 32) ^ y ^ k1; y = (uint32_t)p;
  p = (uint64_t)x * 0xD2511F53u; x = (uint32_t)(p >> 32) ^ y ^ (k0 + 0x9E3779B9u); y = (uint32_t)p;
  p = (uint64_t)x * 0xCD9E8D57u; x = (uint32_t)(p >> 32) ^ y ^ (k1 + 0xBB67AE85u); y = (uint32_t)p;
  return make_uint2(x, y);
}

__device__ __forceinline__ void drop_keep4(uint64_t seed, uint32_t stream_id, uint64_t idx4, uint32_t thr, bool (&keep)[4]) {
  const uint2 w = drop_words(seed, stream_id, idx4);
  keep[0] = (w.x & 0xFFFFu) >= thr;
  keep[1] = (w.x >> 16) >= thr;
  keep[2] = (w.y & 0xFFFFu) >= thr;
  keep[3] = (w.y >> 16) >= thr;
}

__device__ __forceinline__ bool drop_keep1(uint64_t seed, uint32_t stream_id, uint64_t idx, uint32_t thr) {
  const uint2 w = drop_words(seed, stream_id, idx >> 2);
  const uint32_t sel = (uint32_t)(idx & 3);
  const uint32_t word = sel < 2 ? w.x : w.y;
  const uint32_t r = (sel & 1) ? (word >> 16) : (word & 0xFFFFu);
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
    drop_keep4(seed, stream_id, idx >> 2, thr, keep);
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

// ---------------------------------------------------------------------------------------------
// fp16 hi/lo split of fp32 data for the tcgen05 kind::f16 GEMMs (tx_gemm.cu): x * scale = hi + lo with a per-tensor power-of-two
// scale derived from an UPPER BOUND of |x| so that |x| * scale <= 2^13 (fp16 max is 65504).  hi carries 11 significant bits,
// lo the next 11 (exactly when |x * scale| >= 2^-3, else down to the fp16 subnormal spacing 2^-24): the absolute error is at most
// max(2^-22 |x|, 2^-25 / scale), i.e. <= 2^-22 of the tensor's largest entry whenever the bound is within 2^16 of the true maximum.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float f16_split_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  int e;
  frexpf(bound, &e);                 // bound = m 2^e with m in [0.5, 1)  =>  bound <= 2^e
  int sh = 13 - e;
  sh = sh < -120 ? -120 : (sh > 120 ? 120 : sh);
  return ldexpf(1.f, sh);
}
// returns {hi.x | hi.y << 16, ...}: packs the fp16 hi (or lo) halves of 4 scaled values into two 32-bit words
__device__ __forceinline__ void f16_split4(float4 v, float scale, uint2& hi, uint2& lo) {
  const float c = 65504.f;
  const float x0 = fminf(fmaxf(v.x * scale, -c), c), x1 = fminf(fmaxf(v.y * scale, -c), c);
  const float x2 = fminf(fmaxf(v.z * scale, -c), c), x3 = fminf(fmaxf(v.w * scale, -c), c);
  const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): the hot-path kernels are launched with the programmatic-stream-serialization attribute and
// start with TX_PDL_ENTER() (trigger, then wait) - the trigger lets the NEXT kernel of the stream be scheduled as soon as every CTA of this
// one has started (its launch and prologue then overlap this kernel's tail), the wait blocks until the PREVIOUS kernel has completed and
// its writes are visible.  Every kernel waits before it touches global memory, so the stream's semantics are unchanged; what goes away
// is the 2-3 us of launch latency between ~55 short kernels of a step.  TAXO_PDL=0 launches without the attribute (both instructions
// are then no-ops).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ld.global.nc (__ldg(), and whatever nvcc derives from `const T* __restrict__`) promises data that is read-only for the kernel's
// lifetime, so NVVM and ptxas may schedule such a load in front of griddepcontrol.wait - inline-asm memory clobbers do not order it: a
// kernel that read a device scalar written by its predecessor as its first statement got `LDG.E.CONSTANT` placed before `ACQBULK` and
// used the stale value.  Everything after the wait is therefore made control-dependent on a value neither compiler stage can fold
// (%nsmid, read inside the same asm statement): loads through unknown pointers are never speculated across a branch.
// scripts/check_pdl_sass.py (run by tests/test_abi_cpu.py) disassembles the library and checks that no kernel has a global-memory
// instruction in front of its ACQBULK.
__device__ __forceinline__ bool pdl_wait_guard() {
  unsigned n_sm;
  asm volatile("griddepcontrol.wait;\n\tmov.u32 %0, %%nsmid;" : "=r"(n_sm)::"memory");
  return n_sm != 0;
}
#define TX_PDL_ENTER()          \
  do {                          \
    tx::pdl_trigger();          \
    if (!tx::pdl_wait_guard()) return; \
  } while (0)
// The per-layer calls (tx_layer.cu) clear every device scalar their kernels accumulate into (max|.| outputs, the weight split's
// rendezvous counter) with ONE memset at the start of the call and set this flag, so that tx_absmax / tx_split_f16_weight /
// tx_gemm_nt_f16x3 skip their own cudaMemsetAsync: a memset between two kernels is a full stream dependency that the programmatic
// launch cannot overlap.
extern thread_local int g_preclear;
bool preclear_enabled();                 // TAXO_PRECLEAR=0: the kernels' own memsets stay (A/B knob)
struct PreclearScope {
  int prev;
  PreclearScope() : prev(g_preclear) { if (preclear_enabled()) g_preclear = 1; }
  ~PreclearScope() { g_preclear = prev; }
};
bool pdl_enabled(int group = 0, const char* name = nullptr);
int pdl_set(int enabled);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int group, const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(group, name) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// kernel names with template commas go in parentheses: TX_PDL_LAUNCH((k<1, 2>), grid, block, smem, stream, args...)
#ifndef TX_PDL_GROUP
#define TX_PDL_GROUP 0
#endif
#define TX_PDL_LAUNCH(kernel, grid, block, smem, st, ...) (void)tx::launch_pdl(TX_PDL_GROUP, #kernel, kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)

inline int64_t row_blocks(int64_t n) { return (n + kRowsPerBlock - 1) / kRowsPerBlock; }
inline int grid_for_warps(int64_t n_warp_items, int warps_per_block, int blocks_per_sm) {
  int64_t need = (n_warp_items + warps_per_block - 1) / warps_per_block;
  int64_t cap = (int64_t)kNumSms * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace tx
