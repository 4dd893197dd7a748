// Fused GAT gather-attend-aggregate kernels (sm_100a): the hot path of the hot path.
//
// Forward  (tx_gat_fused_fwd): ONE pass over ft.  A warp owns one (destination row, head): it loads its own row
//   once (128-bit, coalesced), derives the attention half-logits a1 = <ft_j, attn_l>, a2 = <ft_i, attn_r> by
//   warp-shuffle reductions from the rows it gathers (no a1/a2 round trip through HBM, no separate logits pass),
//   runs the edge softmax lane-parallel over the in-edges, and accumulates sum_j alpha~_ij ft_j.  Rows shared by
//   several destinations of an egonet (the anchor row is read by every sibling) are re-read through L1/L2, so
//   DRAM traffic stays at the algorithmic "read ft once, write out once".  The epilogue writes the NEXT layer's
//   input directly (leaky-relu, feat-dropout, position-embedding append, zero padding) and packs the
//   activation-sign / dropout-keep bits into 2 bit-planes (1/16 of the row bytes) for the backward pass.
// Backward (tx_gat_fused_bwd): a CTA owns a TILE of whole graphs (graphs whose first row falls into a window of
//   kTileRows rows; egonets are closed components, so every edge of the tile is internal).  Phase A (per
//   destination): d(alpha~) = <g_i, ft_j>, softmax + leaky-relu backward, ds -> scratch, d(attn_l/r) partials in
//   registers.  __syncthreads.  Phase B (per source): dft_j = sum_i alpha~_ij g_i + da1_j attn_l + da2_j attn_r.
//   g rows touched again in phase B were just read by the same CTA, so they come from L1/L2: DRAM traffic is
//   "read g once, read ft once, write dft once" - the 3 N W of SURVEY.md section 8d.  g is consumed straight from
//   d(z_next): the dropout / leaky-relu derivative factor is rebuilt from the bit-planes on load.
//   No float atomics anywhere: d(attn) partials are reduced per CTA in a fixed order.
#include <math.h>
#include <stdlib.h>

#define TX_PDL_GROUP 5
#include "tx_common.cuh"

namespace tx {

constexpr int kTileRows = 32;   // default backward tile window (rows); tiles are graph-aligned (TAXO_BWD_TILE_ROWS overrides; 16/32 measured, r13-r14)
constexpr int kMaxNV = 4;       // up to 4 float4 per lane per head row -> D' <= 512

template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, int D, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    v[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Bulk L2 prefetch of `bytes` (multiple of 16) starting at a 16-byte aligned global address (cp.async.bulk.prefetch, sm_90+):
// pulls the rows a warp / CTA will touch NEXT into L2 so the demand loads later pay L2, not HBM, latency.
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int NV>
__device__ __forceinline__ float dot_row(const float4 (&a)[NV], const float4* __restrict__ s, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const float4 b = s[lane + 32 * t];
    acc = fmaf(a[t].x, b.x, acc); acc = fmaf(a[t].y, b.y, acc); acc = fmaf(a[t].z, b.z, acc); acc = fmaf(a[t].w, b.w, acc);
  }
  return acc;
}

template <int NV>
__device__ __forceinline__ float dot_rows(const float4 (&a)[NV], const float4 (&b)[NV]) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    acc = fmaf(a[t].x, b[t].x, acc); acc = fmaf(a[t].y, b[t].y, acc); acc = fmaf(a[t].z, b[t].z, acc); acc = fmaf(a[t].w, b[t].w, acc);
  }
  return acc;
}

template <int NV>
__device__ __forceinline__ void axpy_row(float w, const float4 (&x)[NV], float4 (&y)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    y[t].x = fmaf(w, x[t].x, y[t].x); y[t].y = fmaf(w, x[t].y, y[t].y);
    y[t].z = fmaf(w, x[t].z, y[t].z); y[t].w = fmaf(w, x[t].w, y[t].w);
  }
}

// TF32 split of an fp32 value (same rounding as tx_split_tf32): v = hi + lo up to 2^-22 relative
__device__ __forceinline__ float rn_tf32_f(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void store_split4(float* hi_ptr, float* lo_ptr, float4 v) {
  const float4 h = make_float4(rn_tf32_f(v.x), rn_tf32_f(v.y), rn_tf32_f(v.z), rn_tf32_f(v.w));
  *reinterpret_cast<float4*>(hi_ptr) = h;
  *reinterpret_cast<float4*>(lo_ptr) = make_float4(rn_tf32_f(v.x - h.x), rn_tf32_f(v.y - h.y), rn_tf32_f(v.z - h.z), rn_tf32_f(v.w - h.w));
}

struct FusedFwdParams {
  const float* ft; int64_t ldf;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* alpha; float* alpha_d; float* elog;
  float* out; int64_t ldo;
  float* out_lo;           // optional: when set, `out` receives the TF32 hi part and out_lo the lo part of every written value
  // optional fp16-split output (next layer's GEMM operand): x * scale = hi + lo, scale derived from the device bound; replaces `out`
  __half* out16_hi; __half* out16_lo; int64_t ld16; const float* bound; float* scale_out;
  uint32_t* maskbits;      // bytes [n, mask_ld]: byte (h*D + c)/4 of a row = 4 sign bits | 4 keep bits << 4 of columns c..c+3; may be null
  int mask_ld;             // = tx_gat_fused_mask_ld(H, D): H*D/4 rounded up to 16
  // epilogue
  int hidden; float act_slope; const float* next_pos_table; const int32_t* pos; int pos_dim;
  float next_inv_keep; uint32_t next_thr; uint64_t next_seed; uint32_t next_stream;
};

template <int NV>
__device__ __forceinline__ void scale_row(float sc, float4 (&y)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) { y[t].x *= sc; y[t].y *= sc; y[t].z *= sc; y[t].w *= sc; }
}

template <int NV>
__global__ void __launch_bounds__(256, 3) gat_fused_fwd_kernel(const FusedFwdParams p) {
  __shared__ float4 s_l[NV * 32];
  __shared__ float4 s_r[NV * 32];
  const int h = blockIdx.y;
  const int H = p.H, D = p.D;
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool attn_drop = p.attn_thr != 0;
  const float* base = p.ft + (int64_t)h * D;
  const float scale16 = p.out16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.out16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;
  for (int i = warp; i < p.n; i += nwarps) {
    const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
    const int deg = end - beg;
    if (lane == 0 && i + nwarps < p.n) prefetch_l2(base + (int64_t)(i + nwarps) * p.ldf, (uint32_t)D * 4u);
    float4 ra[NV], rb[NV];
    load_row<NV>(base + (int64_t)i * p.ldf, lane, D, ra);            // own row: a2 = <ft_i, attn_r>
    if (deg > 0) load_row<NV>(base + (int64_t)__ldg(p.in_src + beg) * p.ldf, lane, D, rb);   // first neighbour row in flight
    const float a2i = warp_sum(dot_row<NV>(ra, s_r, lane));
    // ---- single pass over the in-edges: logits from the gathered rows + ONLINE edge softmax + aggregation ----
    float m = -INFINITY, l = 0.f, s_mine = 0.f, kw_mine = 1.f;
    float4 acc[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto edge = [&](const float4 (&row)[NV], int k) {
      float s = warp_sum(dot_row<NV>(row, s_l, lane)) + a2i;          // a1[src] + a2[dst]     (model_zoo.py:108)
      s = s > 0.f ? s : s * p.neg_slope;
      float kw = 1.f;
      if (attn_drop)
        kw = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
      if (lane == ((k - beg) & 31)) { s_mine = s; kw_mine = kw; }
      if (deg > 32 && lane == 0) p.elog[(int64_t)k * H + h] = s;     // rare: spill logits for the fix-up pass
      const float m_new = fmaxf(m, s);
      const float sc = expf(m - m_new);                              // 0 on the first edge (m = -inf)
      const float e = expf(s - m_new);
      l = fmaf(l, sc, e);
      if (sc != 1.f) scale_row<NV>(sc, acc);                          // warp-uniform
      axpy_row<NV>(e * kw, row, acc);
      m = m_new;
    };
    for (int k = beg; k < end;) {                                     // ping-pong register buffers: next row in flight
      if (k + 1 < end) load_row<NV>(base + (int64_t)__ldg(p.in_src + k + 1) * p.ldf, lane, D, ra);
      edge(rb, k);
      if (++k >= end) break;
      if (k + 1 < end) load_row<NV>(base + (int64_t)__ldg(p.in_src + k + 1) * p.ldf, lane, D, rb);
      edge(ra, k);
      ++k;
    }
    const float inv_l = deg > 0 ? 1.f / l : 0.f;
    scale_row<NV>(inv_l, acc);
    // ---- per-edge outputs for the backward pass (slot order) ----
    if (deg <= 32) {
      if (lane < deg) {
        const int64_t o = (int64_t)(beg + lane) * H + h;
        const float a = expf(s_mine - m) * inv_l;
        p.elog[o] = s_mine;
        p.alpha[o] = a;
        if (attn_drop) p.alpha_d[o] = a * kw_mine;
      }
    } else {
      __syncwarp();
      for (int k = beg + lane; k < end; k += 32) {
        const int64_t o = (int64_t)k * H + h;
        const float a = expf(p.elog[o] - m) * inv_l;
        p.alpha[o] = a;
        if (attn_drop)
          p.alpha_d[o] = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? a * p.attn_inv_keep : 0.f;
      }
    }
    // ---- epilogue (per-row bases hoisted: the per-vector work is address-free) ----
    const int64_t row_off = (int64_t)i * p.ldo + (p.hidden ? (int64_t)h * D : 0) + lane * 4;
    float* orow = p.out + row_off;
    float* olo = p.out_lo ? p.out_lo + row_off : nullptr;
    uint8_t* mrow = p.maskbits ? reinterpret_cast<uint8_t*>(p.maskbits) + (int64_t)i * p.mask_ld + ((h * D) >> 2) + lane : nullptr;
    const uint64_t idx4_base = (uint64_t)(((int64_t)i * p.ldo + (int64_t)h * D) >> 2) + (uint64_t)lane;   // ldo % 4 == 0, D % 4 == 0
    const bool act = p.hidden && p.act_slope != 1.f;
    const bool drop = p.hidden && p.next_thr != 0;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      if ((lane + 32 * t) * 4 < D) {
        float v[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
        uint32_t code = 0xF0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) code |= v[u] > 0.f ? (1u << u) : 0u;
        if (act) {
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = (code >> u) & 1u ? v[u] : v[u] * p.act_slope;
        }
        if (drop) {
          const uint2 w = drop_words(p.next_seed, p.next_stream, idx4_base + 32u * t);
          const uint32_t r[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool keep = r[u] >= p.next_thr;
            v[u] = keep ? v[u] * p.next_inv_keep : 0.f;
            code &= keep ? 0xFFu : ~(16u << u);
          }
        }
        if (p.out16_hi) {
          uint2 h16, l16;
          f16_split4(make_float4(v[0], v[1], v[2], v[3]), scale16, h16, l16);
          const int64_t o16 = (int64_t)i * p.ld16 + (int64_t)h * D + (lane + 32 * t) * 4;
          *reinterpret_cast<uint2*>(p.out16_hi + o16) = h16;
          *reinterpret_cast<uint2*>(p.out16_lo + o16) = l16;
        } else if (olo) store_split4(orow + 128 * t, olo + 128 * t, make_float4(v[0], v[1], v[2], v[3]));
        else *reinterpret_cast<float4*>(orow + 128 * t) = make_float4(v[0], v[1], v[2], v[3]);
        if (mrow) mrow[32 * t] = (uint8_t)code;     // 4 sign bits | 4 keep bits << 4 of this lane's 4 columns
      }
    }
    // position-embedding append + zero padding (once per row: the warp of the last head)
    if (p.hidden && h == H - 1) {
      const int feat = H * D;
      const int pd = p.pos_dim;
      float* row = p.out + (int64_t)i * p.ldo;
      const float* prow = pd > 0 ? p.next_pos_table + (int64_t)__ldg(p.pos + i) * pd : nullptr;
      const int c_end = p.out16_hi ? (int)p.ld16 : (int)p.ldo;
      for (int c = feat + lane; c < c_end; c += 32) {
        float v = 0.f;
        if (c < feat + pd) {
          v = __ldg(prow + (c - feat));
          if (p.next_thr) v = drop_keep1(p.next_seed, p.next_stream, (uint64_t)((int64_t)i * p.ldo + c), p.next_thr) ? v * p.next_inv_keep : 0.f;
        }
        if (p.out16_hi) {
          const float x = fminf(fmaxf(v * scale16, -65504.f), 65504.f);
          const __half hh = __float2half_rn(x);
          p.out16_hi[(int64_t)i * p.ld16 + c] = hh;
          p.out16_lo[(int64_t)i * p.ld16 + c] = __float2half_rn(x - __half2float(hh));
        } else if (p.out_lo) {
          const float vh = rn_tf32_f(v);
          row[c] = vh;
          p.out_lo[(int64_t)i * p.ldo + c] = rn_tf32_f(v - vh);
        } else {
          row[c] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct FusedBwdParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const uint32_t* maskbits; int mask_ld; int has_keep_plane; float act_slope; float next_inv_keep;
  const float* ft; int64_t ldf;
  const float* alpha; const float* alpha_d; const float* elog;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  const int32_t* out_ptr; const int32_t* out_dst; const int32_t* out_slot;
  const int32_t* node_off; int n_graphs;
  int n; int H; int D; int tile_rows; int stage_meta; int prefetch;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* ds; float* da2;
  float* dft; int64_t ldd;
  float* dft_lo;          // optional: when set, dft receives the TF32 hi part and dft_lo the lo part (feeds the 3xTF32 GEMMs directly)
  float* dattn_partial;   // [gridDim.x, 2, H, D]
};

// g row of node i, head h with the epilogue derivative (dropout keep / (1-p), leaky-relu slope) applied on load
template <int NV>
__device__ __forceinline__ void load_g_row(const FusedBwdParams& p, int i, int h, int lane, float4 (&v)[NV]) {
  const float* row = p.g + (int64_t)i * p.ldg + (int64_t)h * p.g_head_stride;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    float4 x = c < p.D ? __ldg(reinterpret_cast<const float4*>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.maskbits) {
      uint32_t code = c < p.D ? __ldg(reinterpret_cast<const uint8_t*>(p.maskbits) + (int64_t)i * p.mask_ld + ((h * p.D + c) >> 2)) : 0u;
      if (!p.has_keep_plane) code |= 0xF0u;
      const float on = p.next_inv_keep, neg = p.act_slope * p.next_inv_keep;
      x.x *= (code & 16u) ? ((code & 1u) ? on : neg) : 0.f;
      x.y *= (code & 32u) ? ((code & 2u) ? on : neg) : 0.f;
      x.z *= (code & 64u) ? ((code & 4u) ? on : neg) : 0.f;
      x.w *= (code & 128u) ? ((code & 8u) ? on : neg) : 0.f;
    }
    v[t] = x;
  }
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int value) {
  int lo = 0, hi = n;   // first index in [0, n) with a[idx] >= value, or n
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < value) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kHeavyOut = 12;   // sources with more out-edges than this are processed by the whole CTA
constexpr int kStRows = 96;     // tile metadata staged in shared memory when the tile has <= kStRows rows ...
constexpr int kStEdges = 192;   // ... and <= kStEdges edges (else the same code reads the global arrays)

#ifndef TX_BWD_MIN_BLOCKS
#define TX_BWD_MIN_BLOCKS 2   /* 128 registers, no spills: measured 0.71 ms vs 0.85 ms at 3 CTAs/SM with spills (r11) */
#endif
// dynamic smem: s_l, s_r [NV*32] float4 | s_acc [8][2][NV*32] float4 (per-warp d(attn) accumulators) | s_part [8][NV*32]
template <int NV>
__global__ void __launch_bounds__(256, TX_BWD_MIN_BLOCKS) gat_fused_bwd_kernel(const FusedBwdParams p) {
  extern __shared__ float4 smem_f4[];
  float4* s_l = smem_f4;
  float4* s_r = s_l + NV * 32;
  float4* s_acc = s_r + NV * 32;            // [(w*2 + lr) * NV*32 + q]
  float4* s_part = s_acc + 16 * NV * 32;    // [w * NV*32 + q]
  // tile metadata (CSR slices, attention coefficients of this head) and the ds / da2 exchange between the two phases
  __shared__ int s_inptr[kStRows + 1], s_outptr[kStRows + 1], s_insrc[kStEdges], s_outdst[kStEdges], s_outslot[kStEdges];
  __shared__ float s_alpha[kStEdges], s_alphad[kStEdges], s_elog[kStEdges], s_keepw[kStEdges], s_ds[kStEdges], s_da2[kStRows];
  __shared__ int s_bounds[2][2];
  const int h = blockIdx.y;
  const int H = p.H, D = p.D;
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int t = threadIdx.x; t < 16 * NV * 32; t += blockDim.x) s_acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool attn_drop = p.attn_thr != 0;
  float4* my_accl = s_acc + (wid * 2 + 0) * NV * 32;
  float4* my_accr = s_acc + (wid * 2 + 1) * NV * 32;
  const float* fbase = p.ft + (int64_t)h * D;
  const int n_tiles = (p.n + p.tile_rows - 1) / p.tile_rows;
  auto tile_bounds = [&](int tile, int slot) {   // graphs whose first row lies in [tile*R, (tile+1)*R)
    if (tile < n_tiles) {
      const int gb = lower_bound_i32(p.node_off, p.n_graphs + 1, tile * p.tile_rows);
      const int ge = lower_bound_i32(p.node_off, p.n_graphs + 1, (tile + 1) * p.tile_rows);
      s_bounds[slot][0] = __ldg(p.node_off + min(gb, p.n_graphs));
      s_bounds[slot][1] = __ldg(p.node_off + min(ge, p.n_graphs));
    } else {
      s_bounds[slot][0] = s_bounds[slot][1] = 0;
    }
  };
  if (threadIdx.x == 0) tile_bounds(blockIdx.x, 0);
  __syncthreads();
  int slot = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, slot ^= 1) {
    const int r0 = s_bounds[slot][0], r1 = s_bounds[slot][1];
    const int nrows = r1 - r0;
    const int s0 = nrows > 0 ? __ldg(p.in_ptr + r0) : 0, s1 = nrows > 0 ? __ldg(p.in_ptr + r1) : 0;
    const int o0 = nrows > 0 ? __ldg(p.out_ptr + r0) : 0, o1 = nrows > 0 ? __ldg(p.out_ptr + r1) : 0;
    const bool staged = p.stage_meta && nrows <= kStRows && (s1 - s0) <= kStEdges && (o1 - o0) <= kStEdges;
    if (wid == 7) {   // bounds of this CTA's next tile + L2 prefetch of its g / ft rows while this tile is processed
      if (lane == 0) tile_bounds(tile + gridDim.x, slot ^ 1);
      __syncwarp();
      const int q0 = s_bounds[slot ^ 1][0], q1 = p.prefetch ? s_bounds[slot ^ 1][1] : s_bounds[slot ^ 1][0];
      for (int r = q0 + lane; r < q1; r += 32) {
        prefetch_l2(p.g + (int64_t)r * p.ldg + (int64_t)h * p.g_head_stride, (uint32_t)D * 4u);
        prefetch_l2(fbase + (int64_t)r * p.ldf, (uint32_t)D * 4u);
      }
    }
    if (staged) {     // one coalesced sweep instead of 3-4 dependent L2 round trips per row in each phase
      for (int t = threadIdx.x; t <= nrows; t += blockDim.x) {
        s_inptr[t] = __ldg(p.in_ptr + r0 + t);
        s_outptr[t] = __ldg(p.out_ptr + r0 + t);
      }
      for (int t = threadIdx.x; t < s1 - s0; t += blockDim.x) {
        const int64_t o = (int64_t)(s0 + t) * H + h;
        s_insrc[t] = __ldg(p.in_src + s0 + t);
        s_alpha[t] = __ldg(p.alpha + o);
        s_alphad[t] = __ldg(p.alpha_d + o);
        s_elog[t] = __ldg(p.elog + o);
        s_keepw[t] = !attn_drop ? 1.f
                   : (drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + s0 + t) * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f);
      }
      for (int t = threadIdx.x; t < o1 - o0; t += blockDim.x) {
        s_outdst[t] = __ldg(p.out_dst + o0 + t);
        s_outslot[t] = __ldg(p.out_slot + o0 + t);
      }
      __syncthreads();
    }
    auto IN_PTR = [&](int i) { return staged ? s_inptr[i - r0] : __ldg(p.in_ptr + i); };
    auto IN_SRC = [&](int k) { return staged ? s_insrc[k - s0] : __ldg(p.in_src + k); };
    auto OUT_PTR = [&](int j) { return staged ? s_outptr[j - r0] : __ldg(p.out_ptr + j); };
    auto OUT_DST = [&](int k) { return staged ? s_outdst[k - o0] : __ldg(p.out_dst + k); };
    auto OUT_SLOT = [&](int k) { return staged ? s_outslot[k - o0] : __ldg(p.out_slot + k); };
    auto ALPHA = [&](int k) { return staged ? s_alpha[k - s0] : __ldg(p.alpha + (int64_t)k * H + h); };
    auto ALPHAD = [&](int k) { return staged ? s_alphad[k - s0] : __ldg(p.alpha_d + (int64_t)k * H + h); };
    auto ELOG = [&](int k) { return staged ? s_elog[k - s0] : __ldg(p.elog + (int64_t)k * H + h); };
    auto KEEPW = [&](int k) {
      if (!attn_drop) return 1.f;
      if (staged) return s_keepw[k - s0];
      return drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
    };
    auto DS_W = [&](int k, float v) { if (staged) s_ds[k - s0] = v; else p.ds[(int64_t)k * H + h] = v; };
    auto DS_R = [&](int k) { return staged ? s_ds[k - s0] : p.ds[(int64_t)k * H + h]; };
    // ---------------- phase A: per destination: d(alpha~) = <g_i, ft_j>, softmax / leaky-relu backward ----------------
    for (int i = r0 + wid; i < r1; i += 8) {
      const int beg = IN_PTR(i), end = IN_PTR(i + 1);
      const int deg = end - beg;
      float4 gi[NV], ra[NV], rb[NV];
      if (deg > 0) load_row<NV>(fbase + (int64_t)IN_SRC(beg) * p.ldf, lane, D, ra);
      load_g_row<NV>(p, i, h, lane, gi);
      float d_mine = 0.f;
      auto edge = [&](const float4 (&row)[NV], int k) {
        const float d = warp_sum(dot_rows<NV>(gi, row)) * p.g_scale;
        if (lane == ((k - beg) & 31)) d_mine = d;
        if (deg > 32 && lane == 0) DS_W(k, d);
      };
      for (int k = beg; k < end;) {
        if (k + 1 < end) load_row<NV>(fbase + (int64_t)IN_SRC(k + 1) * p.ldf, lane, D, rb);
        edge(ra, k);
        if (++k >= end) break;
        if (k + 1 < end) load_row<NV>(fbase + (int64_t)IN_SRC(k + 1) * p.ldf, lane, D, ra);
        edge(rb, k);
        ++k;
      }
      float a2;
      if (deg <= 32) {
        float da = 0.f, a = 0.f;
        if (lane < deg) {
          da = d_mine * KEEPW(beg + lane);
          a = ALPHA(beg + lane);
        }
        const float tsum = warp_sum(a * da);
        float dsv = 0.f;
        if (lane < deg) {
          const float de = a * (da - tsum);
          dsv = ELOG(beg + lane) > 0.f ? de : de * p.neg_slope;
          DS_W(beg + lane, dsv);
        }
        a2 = warp_sum(dsv);
      } else {
        __syncwarp();
        float tsum = 0.f;
        for (int k = beg + lane; k < end; k += 32) tsum = fmaf(ALPHA(k), DS_R(k) * KEEPW(k), tsum);
        tsum = warp_sum(tsum);
        a2 = 0.f;
        for (int k = beg + lane; k < end; k += 32) {
          const float de = ALPHA(k) * (DS_R(k) * KEEPW(k) - tsum);
          const float dsv = ELOG(k) > 0.f ? de : de * p.neg_slope;
          DS_W(k, dsv);
          a2 += dsv;
        }
        a2 = warp_sum(a2);
      }
      if (lane == 0) {
        if (staged) s_da2[i - r0] = a2; else p.da2[(int64_t)i * H + h] = a2;
      }
    }
    __syncthreads();   // ds / da2 of the whole tile are visible to the CTA
    // ---------------- phase B: per source (light rows: one warp each) ----------------
    for (int j = r0 + wid; j < r1; j += 8) {
      const int beg = OUT_PTR(j), end = OUT_PTR(j + 1);
      if (end - beg > kHeavyOut) continue;
      float4 ra[NV], rb[NV];
      if (beg < end) load_g_row<NV>(p, OUT_DST(beg), h, lane, ra);
      float d1 = 0.f;
      if (beg + lane < end) d1 = DS_R(OUT_SLOT(beg + lane));
      d1 = warp_sum(d1);
      const float d2 = staged ? s_da2[j - r0] : p.da2[(int64_t)j * H + h];
      float4 acc[NV];
      {
        float4 fj[NV];
        load_row<NV>(fbase + (int64_t)j * p.ldf, lane, D, fj);   // L1/L2 hit: read by this CTA in phase A
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int q = lane + 32 * t;
          float4 xl = my_accl[q], xr = my_accr[q];
          xl.x = fmaf(d1, fj[t].x, xl.x); xl.y = fmaf(d1, fj[t].y, xl.y); xl.z = fmaf(d1, fj[t].z, xl.z); xl.w = fmaf(d1, fj[t].w, xl.w);
          xr.x = fmaf(d2, fj[t].x, xr.x); xr.y = fmaf(d2, fj[t].y, xr.y); xr.z = fmaf(d2, fj[t].z, xr.z); xr.w = fmaf(d2, fj[t].w, xr.w);
          my_accl[q] = xl; my_accr[q] = xr;
          const float4 l = s_l[q], r = s_r[q];
          acc[t] = make_float4(fmaf(d1, l.x, d2 * r.x), fmaf(d1, l.y, d2 * r.y), fmaf(d1, l.z, d2 * r.z), fmaf(d1, l.w, d2 * r.w));
        }
      }
      for (int k = beg; k < end;) {
        if (k + 1 < end) load_g_row<NV>(p, OUT_DST(k + 1), h, lane, rb);
        axpy_row<NV>(ALPHAD(OUT_SLOT(k)) * p.g_scale, ra, acc);
        if (++k >= end) break;
        if (k + 1 < end) load_g_row<NV>(p, OUT_DST(k + 1), h, lane, ra);
        axpy_row<NV>(ALPHAD(OUT_SLOT(k)) * p.g_scale, rb, acc);
        ++k;
      }
      float* orow = p.dft + (int64_t)j * p.ldd + (int64_t)h * D;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c = (lane + 32 * t) * 4;
        if (c < D) {
          if (p.dft_lo) store_split4(orow + c, p.dft_lo + (orow - p.dft) + c, acc[t]);
          else *reinterpret_cast<float4*>(orow + c) = acc[t];
        }
      }
    }
    // ---------------- phase B, heavy sources (e.g. the anchor of a large egonet): the whole CTA per row ----------------
    for (int jb = r0; jb < r1; jb += 32) {
     const int jl = min(jb + lane, r1 - 1);
     unsigned heavy = __ballot_sync(0xffffffffu, jb + lane < r1 && (OUT_PTR(jl + 1) - OUT_PTR(jl)) > kHeavyOut);
     while (heavy) {            // identical in every warp of the CTA -> uniform control flow around the barriers below
      const int j = jb + __ffs(heavy) - 1;
      heavy &= heavy - 1;
      const int beg = OUT_PTR(j), end = OUT_PTR(j + 1);
      float4 acc[NV], ra[NV], rb[NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = beg + wid;
      if (k < end) load_g_row<NV>(p, OUT_DST(k), h, lane, ra);
      while (k < end) {
        if (k + 8 < end) load_g_row<NV>(p, OUT_DST(k + 8), h, lane, rb);
        axpy_row<NV>(ALPHAD(OUT_SLOT(k)) * p.g_scale, ra, acc);
        k += 8;
        if (k >= end) break;
        if (k + 8 < end) load_g_row<NV>(p, OUT_DST(k + 8), h, lane, ra);
        axpy_row<NV>(ALPHAD(OUT_SLOT(k)) * p.g_scale, rb, acc);
        k += 8;
      }
#pragma unroll
      for (int t = 0; t < NV; ++t) s_part[wid * NV * 32 + lane + 32 * t] = acc[t];
      __syncthreads();
      if (wid == 0) {
        float d1 = 0.f;
        for (int kk = beg + lane; kk < end; kk += 32) d1 += DS_R(OUT_SLOT(kk));
        d1 = warp_sum(d1);
        const float d2 = staged ? s_da2[j - r0] : p.da2[(int64_t)j * H + h];
        float4 fj[NV];
        load_row<NV>(fbase + (int64_t)j * p.ldf, lane, D, fj);
        float* orow = p.dft + (int64_t)j * p.ldd + (int64_t)h * D;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int q = lane + 32 * t;
          float4 xl = my_accl[q], xr = my_accr[q];
          xl.x = fmaf(d1, fj[t].x, xl.x); xl.y = fmaf(d1, fj[t].y, xl.y); xl.z = fmaf(d1, fj[t].z, xl.z); xl.w = fmaf(d1, fj[t].w, xl.w);
          xr.x = fmaf(d2, fj[t].x, xr.x); xr.y = fmaf(d2, fj[t].y, xr.y); xr.z = fmaf(d2, fj[t].z, xr.z); xr.w = fmaf(d2, fj[t].w, xr.w);
          my_accl[q] = xl; my_accr[q] = xr;
          const float4 l = s_l[q], r = s_r[q];
          float4 o = make_float4(fmaf(d1, l.x, d2 * r.x), fmaf(d1, l.y, d2 * r.y), fmaf(d1, l.z, d2 * r.z), fmaf(d1, l.w, d2 * r.w));
#pragma unroll
          for (int w = 0; w < 8; ++w) {   // fixed order: deterministic
            const float4 x = s_part[w * NV * 32 + q];
            o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
          }
          const int c = q * 4;
          if (c < D) {
            if (p.dft_lo) store_split4(orow + c, p.dft_lo + (orow - p.dft) + c, o);
            else *reinterpret_cast<float4*>(orow + c) = o;
          }
        }
      }
      __syncthreads();
     }
    }
    __syncthreads();   // next tile: its bounds (written by warp 7) are visible, staging buffers and s_part are free
  }
  // ---- d(attn) partials: fixed-order reduction over the 8 per-warp accumulators ----
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * NV * 32; t += blockDim.x) {
    const int lr = t / (NV * 32), q = t % (NV * 32);
    float4 s = s_acc[(0 * 2 + lr) * NV * 32 + q];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 x = s_acc[(w * 2 + lr) * NV * 32 + q];
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    const int c = q * 4;
    if (c < D) {
      float* dst = p.dattn_partial + (((int64_t)blockIdx.x * 2 + lr) * H + h) * D + c;
      *reinterpret_cast<float4*>(dst) = s;
    }
  }
}

// d(P_next) partials only: dpos_partial[b, r, :] = sum_{i in block b, pos_i = r} dz[i, col0 : col0+pd] * keep/(1-p)
// 256 threads = 4 row groups x 64 columns; the row groups are summed in a fixed order (deterministic)
__global__ void __launch_bounds__(256) pos_grad_partials_kernel(const float* __restrict__ dz, int64_t ldz, int col0,
                                                                const int32_t* __restrict__ pos, int n, int pd, int vocab,
                                                                float inv_keep, uint32_t thr, uint64_t seed, uint32_t stream_id,
                                                                float* __restrict__ partial) {
  TX_PDL_ENTER();
  __shared__ float s_acc[3][kMaxVocab][64];
  const int r0 = blockIdx.x * kRowsPerBlock;
  const int r1 = min(n, r0 + kRowsPerBlock);
  const int cl = threadIdx.x & 63, rg = threadIdx.x >> 6;
  for (int cb = 0; cb < pd; cb += 64) {
    const int c = cb + cl;
    float acc[kMaxVocab];
#pragma unroll
    for (int v = 0; v < kMaxVocab; ++v) acc[v] = 0.f;
    if (c < pd) {
      for (int i0 = r0 + rg; i0 < r1; i0 += 16) {        // 4 rows per step: their loads are issued together
        float gv[4];
        int rv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 4 * u;
          gv[u] = i < r1 ? __ldg(dz + (int64_t)i * ldz + col0 + c) : 0.f;
          rv[u] = i < r1 ? __ldg(pos + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 4 * u;
          float g = gv[u];
          if (thr && i < r1) g = drop_keep1(seed, stream_id, (uint64_t)((int64_t)i * ldz + col0 + c), thr) ? g * inv_keep : 0.f;
#pragma unroll
          for (int v = 0; v < kMaxVocab; ++v) acc[v] += (rv[u] == v) ? g : 0.f;
        }
      }
    }
    if (rg > 0) {
#pragma unroll
      for (int v = 0; v < kMaxVocab; ++v) s_acc[rg - 1][v][cl] = acc[v];
    }
    __syncthreads();
    if (rg == 0 && c < pd) {
      for (int v = 0; v < vocab; ++v)
        partial[((int64_t)blockIdx.x * vocab + v) * pd + c] = ((acc[v] + s_acc[0][v][cl]) + s_acc[1][v][cl]) + s_acc[2][v][cl];
    }
    __syncthreads();
  }
}

}  // namespace tx

using namespace tx;

extern "C" {

int tx_gat_fused_supported(int64_t heads, int64_t dim, int32_t mean_heads) {
  if (dim <= 0 || dim % 4 != 0 || dim > 128 * kMaxNV) return 0;
  if (mean_heads && heads != 1) return 0;
  return 1;
}

int64_t tx_gat_fused_mask_ld(int64_t heads, int64_t dim) { return ((heads * dim / 4 + 15) / 16) * 16; }

int64_t tx_gat_fused_mask_words(int64_t n_nodes, int64_t heads, int64_t dim) {
  return n_nodes * tx_gat_fused_mask_ld(heads, dim) / 4;
}

static int bwd_tile_rows() {
  static int cached = 0;
  if (!cached) {
    const char* e = getenv("TAXO_BWD_TILE_ROWS");
    int v = e ? atoi(e) : kTileRows;
    cached = v >= 1 && v <= 4096 ? v : kTileRows;
  }
  return cached;
}

#ifndef TX_BWD_MIN_BLOCKS
#define TX_BWD_MIN_BLOCKS 2   /* 128 registers, no spills: measured 0.71 ms vs 0.85 ms at 3 CTAs/SM with spills (r11) */
#endif
int64_t tx_gat_fused_bwd_blocks(int64_t n_nodes, int64_t heads) {
  const int64_t tiles = (n_nodes + bwd_tile_rows() - 1) / bwd_tile_rows();
  const char* e_c = getenv("TAXO_BWD_CTAS_PER_SM");
  const int64_t per_sm = e_c && atoi(e_c) > 0 ? atoi(e_c) : TX_BWD_MIN_BLOCKS;
  int64_t gx = (per_sm * (int64_t)kNumSms + heads - 1) / heads;
  if (gx > tiles) gx = tiles;
  return gx < 1 ? 1 : gx;
}

static int gat_fused_fwd_impl(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                              const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                              float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                              float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                              float* out_lo, void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out,
                              void* stream);

int tx_gat_fused_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                     const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                     float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                     float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                     float* out_lo, void* stream) {
  return gat_fused_fwd_impl(ft, ldf, attn_l, attn_r, in_ptr, in_src, in_eid, n_nodes, heads, dim, neg_slope, p_attn, attn_seed,
                            attn_stream_id, alpha, alpha_d, elog, out, ldo, epi, maskbits, out_lo, nullptr, nullptr, 0, nullptr, nullptr,
                            stream);
}

int tx_gat_fused_fwd_f16(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                         const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                         float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                         float* alpha_d, float* elog, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits, void* out16_hi,
                         void* out16_lo, int64_t ld16, const float* bound, float* scale_out, void* stream) {
  TX_REQUIRE(epi && !epi->mean_heads, "gat_fused_fwd_f16: only a hidden layer's epilogue (the next layer's input) can be written fp16-split");
  TX_REQUIRE(out16_hi && out16_lo && bound && aligned16(out16_hi) && aligned16(out16_lo) && ld16 % 8 == 0 &&
             ld16 >= heads * dim + epi->pos_dim, "gat_fused_fwd_f16: bad fp16 output buffers");
  return gat_fused_fwd_impl(ft, ldf, attn_l, attn_r, in_ptr, in_src, in_eid, n_nodes, heads, dim, neg_slope, p_attn, attn_seed,
                            attn_stream_id, alpha, alpha_d, elog, nullptr, ldo, epi, maskbits, nullptr, out16_hi, out16_lo, ld16, bound,
                            scale_out, stream);
}

static int gat_fused_fwd_impl(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                              const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                              float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                              float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                              float* out_lo, void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out,
                              void* stream) {
  TX_REQUIRE(epi, "gat_fused_fwd: epilogue required");
  TX_REQUIRE(!out_lo || aligned16(out_lo), "gat_fused_fwd: out_lo must be 16-byte aligned");
  TX_REQUIRE(tx_gat_fused_supported(heads, dim, epi->mean_heads), "gat_fused_fwd: unsupported shape (heads %lld dim %lld); use the general path",
             (long long)heads, (long long)dim);
  TX_REQUIRE(aligned16(ft) && ldf % 4 == 0 && (out16_hi || (out && aligned16(out))) && ldo % 4 == 0 && aligned16(attn_l) && aligned16(attn_r),
             "gat_fused_fwd: 16-byte aligned rows required");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f && epi->p_drop >= 0.f && epi->p_drop < 1.f, "gat_fused_fwd: dropout rates must be in [0,1)");
  TX_REQUIRE(alpha && elog && (p_attn == 0.f || (alpha_d && alpha_d != alpha)), "gat_fused_fwd: alpha/elog/alpha_d buffers");
  const int64_t need = epi->mean_heads ? dim : heads * dim + epi->pos_dim;
  TX_REQUIRE(ldo >= need, "gat_fused_fwd: ldo %lld < %lld", (long long)ldo, (long long)need);
  TX_REQUIRE(epi->pos_dim == 0 || (epi->next_pos_table && epi->pos), "gat_fused_fwd: pos_dim > 0 needs next_pos_table and pos");
  if (n_nodes == 0) return TX_OK;
  FusedFwdParams p;
  p.ft = ft; p.ldf = ldf; p.attn_l = attn_l; p.attn_r = attn_r; p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid;
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.out = out; p.ldo = ldo; p.maskbits = maskbits;
  p.mask_ld = (int)tx_gat_fused_mask_ld(heads, dim);
  p.out_lo = out_lo;
  p.out16_hi = (__half*)out16_hi; p.out16_lo = (__half*)out16_lo; p.ld16 = ld16; p.bound = bound; p.scale_out = scale_out;
  p.hidden = epi->mean_heads ? 0 : 1; p.act_slope = epi->act_slope; p.next_pos_table = epi->next_pos_table; p.pos = epi->pos;
  p.pos_dim = (int)epi->pos_dim; p.next_inv_keep = 1.f / (1.f - epi->p_drop); p.next_thr = drop_threshold(epi->p_drop);
  p.next_seed = epi->seed; p.next_stream = epi->stream_id;
  const int nv = (int)((dim + 127) / 128);
  int gx = grid_for_warps(n_nodes, 8, 6);
  gx = (gx + (int)heads - 1) / (int)heads;
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nv) {
    case 1: gat_fused_fwd_kernel<1><<<grid, 256, 0, st>>>(p); break;
    case 2: gat_fused_fwd_kernel<2><<<grid, 256, 0, st>>>(p); break;
    case 3: gat_fused_fwd_kernel<3><<<grid, 256, 0, st>>>(p); break;
    default: gat_fused_fwd_kernel<4><<<grid, 256, 0, st>>>(p); break;
  }
  TX_LAUNCH_CHECK("tx_gat_fused_fwd");
  return TX_OK;
}

int tx_gat_fused_bwd(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const uint32_t* maskbits,
                     int32_t has_keep_plane, float act_slope, float p_next, const float* ft, int64_t ldf,
                     const float* alpha, const float* alpha_d, const float* elog, const float* attn_l,
                     const float* attn_r, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                     const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_slot, const int32_t* node_off,
                     int64_t n_graphs, int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn,
                     uint64_t attn_seed, uint32_t attn_stream_id, float* ds, float* da2, float* dft, int64_t ldd,
                     float* dft_lo, float* dattn_partial, void* stream) {
  TX_REQUIRE(g_head_stride != 0 || heads == 1, "gat_fused_bwd: a shared g row (head mean) needs heads == 1");
  TX_REQUIRE(dim % 4 == 0 && dim <= 128 * kMaxNV, "gat_fused_bwd: dim must be a multiple of 4 and <= %d", 128 * kMaxNV);
  TX_REQUIRE(aligned16(g) && ldg % 4 == 0 && g_head_stride % 4 == 0 && aligned16(ft) && ldf % 4 == 0 && aligned16(dft) && ldd % 4 == 0 &&
             aligned16(attn_l) && aligned16(attn_r) && aligned16(dattn_partial) && (!maskbits || aligned16(maskbits)),
             "gat_fused_bwd: 16-byte aligned rows required");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f && p_next >= 0.f && p_next < 1.f, "gat_fused_bwd: dropout rates must be in [0,1)");
  TX_REQUIRE(node_off && n_graphs >= 0, "gat_fused_bwd: node_off required");
  if (n_nodes == 0) return TX_OK;
  FusedBwdParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.maskbits = maskbits; p.has_keep_plane = has_keep_plane;
  p.mask_ld = (int)tx_gat_fused_mask_ld(heads, dim);
  p.act_slope = act_slope; p.next_inv_keep = 1.f / (1.f - p_next); p.ft = ft; p.ldf = ldf; p.alpha = alpha;
  p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.attn_l = attn_l; p.attn_r = attn_r;
  p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid; p.out_ptr = out_ptr; p.out_dst = out_dst; p.out_slot = out_slot;
  p.node_off = node_off; p.n_graphs = (int)n_graphs; p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim;
  p.neg_slope = neg_slope; p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed;
  p.attn_stream = attn_stream_id; p.ds = ds; p.da2 = da2; p.dft = dft; p.ldd = ldd; p.dattn_partial = dattn_partial;
  p.dft_lo = dft_lo;
  p.tile_rows = bwd_tile_rows();
  { const char* e = getenv("TAXO_BWD_STAGE"); p.stage_meta = e ? atoi(e) : 1; }
  { const char* e = getenv("TAXO_BWD_PREFETCH"); p.prefetch = e ? atoi(e) : 1; }
  const int nv = (int)((dim + 127) / 128);
  dim3 grid((unsigned)tx_gat_fused_bwd_blocks(n_nodes, heads), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(2 + 16 + 8) * nv * 32 * sizeof(float4);
  static bool attr_set[kMaxNV + 1] = {false, false, false, false, false};
  if (!attr_set[nv]) {
    cudaError_t e = cudaSuccess;
    switch (nv) {
      case 1: e = cudaFuncSetAttribute(gat_fused_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
      case 2: e = cudaFuncSetAttribute(gat_fused_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
      case 3: e = cudaFuncSetAttribute(gat_fused_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
      default: e = cudaFuncSetAttribute(gat_fused_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
    }
    if (e != cudaSuccess) {
      set_error("gat_fused_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
    attr_set[nv] = true;
  }
  switch (nv) {
    case 1: gat_fused_bwd_kernel<1><<<grid, 256, smem, st>>>(p); break;
    case 2: gat_fused_bwd_kernel<2><<<grid, 256, smem, st>>>(p); break;
    case 3: gat_fused_bwd_kernel<3><<<grid, 256, smem, st>>>(p); break;
    default: gat_fused_bwd_kernel<4><<<grid, 256, smem, st>>>(p); break;
  }
  TX_LAUNCH_CHECK("tx_gat_fused_bwd");
  return TX_OK;
}

int tx_pos_grad_partials(const float* dz, int64_t ldz, int64_t col0, const int32_t* pos, int64_t n_nodes,
                         int64_t pos_dim, int64_t vocab, float p_drop, uint64_t seed, uint32_t stream_id,
                         float* partial, void* stream) {
  TX_REQUIRE(vocab <= kMaxVocab, "pos_grad_partials: position vocab %lld > %d", (long long)vocab, kMaxVocab);
  TX_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "pos_grad_partials: p_drop must be in [0,1)");
  TX_REQUIRE(pos && partial && ldz >= col0 + pos_dim, "pos_grad_partials: bad arguments");
  if (n_nodes == 0 || pos_dim == 0) return TX_OK;
  TX_PDL_LAUNCH((pos_grad_partials_kernel), (int)row_blocks(n_nodes), 256, 0, (cudaStream_t)stream, dz, ldz, (int)col0, pos, (int)n_nodes, (int)pos_dim,
                                                                                      (int)vocab, 1.f / (1.f - p_drop), drop_threshold(p_drop),
                                                                                      seed, stream_id, partial);
  TX_LAUNCH_CHECK("tx_pos_grad_partials");
  return TX_OK;
}

}  // extern "C"
