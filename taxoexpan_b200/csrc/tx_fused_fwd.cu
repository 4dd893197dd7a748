// TMA-staged fused GAT forward (sm_100a): gather-attend-aggregate + next-layer epilogue out of shared memory.
//
// Same arithmetic and the same outputs as gat_fused_fwd_kernel (tx_fused.cu; reference model/model_zoo.py:83-96,106-114 and
// :214-216 for the epilogue): a1_j = <ft_j, attn_l>, a2_i = <ft_i, attn_r>, e_ij = leaky_relu(a1_j + a2_i, 0.2), edge softmax over the
// in-edges of i, attention dropout, out_i = sum_j alpha~_ij ft_j, then the NEXT layer's input (leaky-relu 0.01, feat-dropout,
// position-embedding append, sign / keep bytes for the backward pass).  The first kernel gave a warp one (destination, head): it
// recomputed a1 of a shared source (the anchor) for every edge and ran the softmax with warp shuffles - ~1000 instructions per
// (row, head), issue bound at ~36 % of HBM peak.  Here the ft rows of a tile of whole graphs (the tiles of tx_gat_bwd_tiles) are
// bulk-copied into a shared-memory ring one tile ahead by 4 producer warps, and 12 compute warps do
//   step 1  a1 / a2 once per NODE (warp per row, two dot products from the staged row),
//   step 2  the edge softmax, one thread per destination row (scalar work over its <= few in-edges; alpha / logits written here),
//   step 3  aggregation + epilogue, one thread per (row group, pair of float4 columns): no shuffles, each staged row is read
//           once per out-edge.
// DRAM traffic stays the algorithmic "read ft once, write out once".  Deterministic (no atomics).
#include <stdlib.h>

#include "tx_common.cuh"

namespace tx {

constexpr int kFCW = 12;                 // compute warps
constexpr int kFPW = 4;                  // producer warps
constexpr int kFCT = kFCW * 32;
constexpr int kFwdThreads = kFCT + kFPW * 32;
constexpr int kFMR = 112;                // staged tile metadata capacity: rows ...
constexpr int kFME = 224;                // ... and in-edges
constexpr int kFMaxSmem = 227 * 1024;

struct FwdTileMeta {
  int r0, r1, s0, s1, mode, pad0, pad1, pad2;
  int in_ptr[kFMR + 1];                  // tile-local in-edge offsets per row
  int in_src[kFME];                      // tile-local source node per in-edge slot
  int in_eid[kFME];                      // edge ids (attention dropout counter)
  float a1[kFMR], a2[kFMR];              // per-node half logits of this head
  float w[kFME];                         // alpha~ per in-edge slot
};
constexpr int kFwdMetaBytes = (int)((sizeof(FwdTileMeta) + 15) / 16 * 16);

struct StagedFwdParams {
  const float* ft; int64_t ldf;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  const int4* tiles; int n_tiles; int ring_rows;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* alpha; float* alpha_d; float* elog;
  float* out; int64_t ldo; float* out_lo;
  __half* out16_hi; __half* out16_lo; int64_t ld16; const float* bound; float* scale_out;
  uint32_t* maskbits; int mask_ld;
  int hidden; float act_slope; const float* next_pos_table; const int32_t* pos; int pos_dim;
  float next_inv_keep; uint32_t next_thr; uint64_t next_seed; uint32_t next_stream;
};

__device__ __forceinline__ uint32_t fs_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void fbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fbulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fcompute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFCT) : "memory"); }
__device__ __forceinline__ float frn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void ffma4(float w, const float4 x, float4& y) {
  y.x = fmaf(w, x.x, y.x); y.y = fmaf(w, x.y, y.y); y.z = fmaf(w, x.z, y.z); y.w = fmaf(w, x.w, y.w);
}

// epilogue of one float4 (columns c .. c+3 of head h, row i): next layer's input or the plain output row
__device__ __forceinline__ void fwd_store4(const StagedFwdParams& p, float scale16, int i, int h, int q4, float4 a) {
  const int D = p.D;
  float v[4] = {a.x, a.y, a.z, a.w};
  uint32_t code = 0xF0u;
#pragma unroll
  for (int u = 0; u < 4; ++u) code |= v[u] > 0.f ? (1u << u) : 0u;
  if (p.hidden && p.act_slope != 1.f) {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (code >> u) & 1u ? v[u] : v[u] * p.act_slope;
  }
  if (p.hidden && p.next_thr != 0) {
    const uint64_t idx4 = (uint64_t)(((int64_t)i * p.ldo + (int64_t)h * D) >> 2) + (uint64_t)q4;
    const uint2 w = drop_words(p.next_seed, p.next_stream, idx4);
    const uint32_t r[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool keep = r[u] >= p.next_thr;
      v[u] = keep ? v[u] * p.next_inv_keep : 0.f;
      code &= keep ? 0xFFu : ~(16u << u);
    }
  }
  const float4 o = make_float4(v[0], v[1], v[2], v[3]);
  if (p.out16_hi) {
    uint2 h16, l16;
    f16_split4(o, scale16, h16, l16);
    const int64_t o16 = (int64_t)i * p.ld16 + (int64_t)h * D + q4 * 4;
    *reinterpret_cast<uint2*>(p.out16_hi + o16) = h16;
    *reinterpret_cast<uint2*>(p.out16_lo + o16) = l16;
  } else {
    const int64_t off = (int64_t)i * p.ldo + (p.hidden ? (int64_t)h * D : 0) + q4 * 4;
    if (p.out_lo) {
      const float4 hi = make_float4(frn_tf32(o.x), frn_tf32(o.y), frn_tf32(o.z), frn_tf32(o.w));
      *reinterpret_cast<float4*>(p.out + off) = hi;
      *reinterpret_cast<float4*>(p.out_lo + off) = make_float4(frn_tf32(o.x - hi.x), frn_tf32(o.y - hi.y), frn_tf32(o.z - hi.z), frn_tf32(o.w - hi.w));
    } else {
      *reinterpret_cast<float4*>(p.out + off) = o;
    }
  }
  if (p.maskbits) reinterpret_cast<uint8_t*>(p.maskbits)[(int64_t)i * p.mask_ld + ((h * D) >> 2) + q4] = (uint8_t)code;
}

template <int NV>
__global__ void __launch_bounds__(kFwdThreads, 1) gat_fwd_staged_kernel(const StagedFwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // full[2], empty[2]
  FwdTileMeta* meta = reinterpret_cast<FwdTileMeta*>(smem + 64);
  float4* s_l = reinterpret_cast<float4*>(smem + 64 + 2 * kFwdMetaBytes);
  float4* s_r = s_l + NV * 32;
  uint8_t* ring = reinterpret_cast<uint8_t*>(s_r + NV * 32);
  const int H = p.H, D = p.D, D4 = D >> 2;
  const uint32_t rowB = (uint32_t)D * 4u;
  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t full0 = fs_addr(bars), empty0 = fs_addr(bars + 2);
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      fbar_init(full0 + 8 * b, kFPW);
      fbar_init(empty0 + 8 * b, kFCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const float* fbase = p.ft + (int64_t)h * D;
  const bool attn_drop = p.attn_thr != 0;
  const int tile_beg = (int)((int64_t)blockIdx.x * p.n_tiles / gridDim.x), tile_end = (int)((int64_t)(blockIdx.x + 1) * p.n_tiles / gridDim.x);

  if (wid >= kFCW) {
    // =========================== producer warps ===========================
    const int pw = wid - kFCW;
    const int pt = lane * kFPW + pw;
    int prev_rows = 0, it = 0;
    int4 t0 = make_int4(0, 0, 0, 0), t1 = t0;
    if (tile_beg < tile_end) { t0 = __ldg(p.tiles + tile_beg); t1 = __ldg(p.tiles + tile_beg + 1); }
    for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
      const int b = it & 1;
      const int r0 = t0.x, r1 = t1.x, s0 = t0.y, s1 = t1.y;
      const int nrows = r1 - r0, nE = s1 - s0;
      if (tile + 1 < tile_end) { t0 = t1; t1 = __ldg(p.tiles + tile + 2); }
      // mode 1: the ft rows of the tile are staged (even tiles grow from the bottom of the ring, odd tiles from the top);  mode 0:
      // tile too large for the ring / metadata arrays: the compute warps read global memory
      const int mode = (nrows > 0 && nrows <= kFMR && nE <= kFME && nrows <= p.ring_rows) ? 1 : 0;
      const int span = mode ? nrows : 0;
      if (it >= 2) fbar_wait(empty0 + 8 * b, (uint32_t)(((it >> 1) - 1) & 1));
      if (span > 0 && prev_rows + span > p.ring_rows) fbar_wait(empty0 + 8 * (b ^ 1), (uint32_t)(((it - 1) >> 1) & 1));
      FwdTileMeta& m = meta[b];
      const uint32_t full = full0 + 8 * b;
      if (pw == 0 && lane == 0) { m.r0 = r0; m.r1 = r1; m.s0 = s0; m.s1 = s1; m.mode = mode; }
      if (mode) {
        const bool hr = pt <= nrows, he = pt < nE;
        int v_ip = 0, v_src = 0, v_eid = 0;
        if (hr) v_ip = __ldg(p.in_ptr + r0 + pt);
        if (he) {
          v_src = __ldg(p.in_src + s0 + pt);
          if (attn_drop) v_eid = __ldg(p.in_eid + s0 + pt);
        }
        const int my_rows = nrows > pw ? (nrows - pw + kFPW - 1) / kFPW : 0;
        if (lane == 0 && my_rows > 0) fbar_expect_tx(full, (uint32_t)my_rows * rowB);
        __syncwarp();
        for (int r = pt; r < nrows; r += 32 * kFPW) {
          const int slot = b ? p.ring_rows - 1 - r : r;
          fbulk_g2s(fs_addr(ring + (size_t)slot * rowB), fbase + (int64_t)(r0 + r) * p.ldf, rowB, full);
        }
        if (hr) m.in_ptr[pt] = v_ip - s0;
        if (he) { m.in_src[pt] = v_src - r0; m.in_eid[pt] = v_eid; }
        for (int t = pt + 32 * kFPW; t <= nrows; t += 32 * kFPW) m.in_ptr[t] = __ldg(p.in_ptr + r0 + t) - s0;
        for (int t = pt + 32 * kFPW; t < nE; t += 32 * kFPW) {
          m.in_src[t] = __ldg(p.in_src + s0 + t) - r0;
          m.in_eid[t] = attn_drop ? __ldg(p.in_eid + s0 + t) : 0;
        }
      }
      __syncwarp();
      if (lane == 0) fbar_arrive(full);
      prev_rows = span;
    }
    return;
  }

  // =========================== compute warps ===========================
  const int ctid = threadIdx.x;
  const int W = (D4 + 1) >> 1;
  const int groups = kFCT / W;
  const int hgrp = ctid / W, hcol = ctid - hgrp * W;
  const bool col_thread = hgrp < groups;
  const bool has2 = hcol + W < D4;
  const int c1 = has2 ? hcol + W : hcol;
  const float scale16 = p.out16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.out16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;

  int it = 0;
  for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
    const int b = it & 1;
    fbar_wait(full0 + 8 * b, (uint32_t)((it >> 1) & 1));
    FwdTileMeta& m = meta[b];
    const int r0 = m.r0, r1 = m.r1, s0 = m.s0, mode = m.mode;
    const int nrows = r1 - r0;
    if (mode) {
      auto FROW = [&](int j) -> const float4* { return reinterpret_cast<const float4*>(ring + (size_t)(b ? p.ring_rows - 1 - j : j) * rowB); };
      // ---------- step 1: a1 / a2 per node ----------
      for (int j = wid; j < nrows; j += kFCW) {
        const float4* row = FROW(j);
        float x1 = 0.f, x2 = 0.f;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int q = lane + 32 * t;
          if (q < D4) {
            const float4 f = row[q], l = s_l[q], r = s_r[q];
            x1 = fmaf(f.x, l.x, x1); x1 = fmaf(f.y, l.y, x1); x1 = fmaf(f.z, l.z, x1); x1 = fmaf(f.w, l.w, x1);
            x2 = fmaf(f.x, r.x, x2); x2 = fmaf(f.y, r.y, x2); x2 = fmaf(f.z, r.z, x2); x2 = fmaf(f.w, r.w, x2);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ta = __shfl_xor_sync(0xffffffffu, x1, o), tb = __shfl_xor_sync(0xffffffffu, x2, o);
          x1 += ta; x2 += tb;
        }
        if (lane == 0) { m.a1[j] = x1; m.a2[j] = x2; }
      }
      fcompute_sync();
      // ---------- step 2: edge softmax, one thread per destination row ----------
      if (ctid < nrows) {
        const int a = m.in_ptr[ctid], e = m.in_ptr[ctid + 1];
        const float a2i = m.a2[ctid];
        float mx = -INFINITY;
        for (int k = a; k < e; ++k) {
          float s = m.a1[m.in_src[k]] + a2i;                           // a1[src] + a2[dst]     (model_zoo.py:108)
          s = s > 0.f ? s : s * p.neg_slope;
          m.w[k] = s;
          mx = fmaxf(mx, s);
        }
        float l = 0.f;
        for (int k = a; k < e; ++k) l += expf(m.w[k] - mx);
        const float inv_l = e > a ? 1.f / l : 0.f;
        for (int k = a; k < e; ++k) {
          const float s = m.w[k];
          const float al = expf(s - mx) * inv_l;
          float kw = 1.f;
          if (attn_drop) kw = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)m.in_eid[k] * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
          const int64_t o = (int64_t)(s0 + k) * H + h;
          p.elog[o] = s;
          p.alpha[o] = al;
          if (attn_drop) p.alpha_d[o] = al * kw;
          m.w[k] = al * kw;
        }
      }
      fcompute_sync();
      // ---------- step 3: aggregation + epilogue, one thread per (row group, pair of float4 columns) ----------
      if (col_thread) {
        for (int t = hgrp; t < nrows; t += groups) {
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          const int a = m.in_ptr[t], e = m.in_ptr[t + 1];
          for (int k = a; k < e; ++k) {
            const float w = m.w[k];
            const float4* row = FROW(m.in_src[k]);
            ffma4(w, row[hcol], acc0);
            ffma4(w, row[c1], acc1);
          }
          fwd_store4(p, scale16, r0 + t, h, hcol, acc0);
          if (has2) fwd_store4(p, scale16, r0 + t, h, c1, acc1);
        }
      }
    } else if (nrows > 0) {
      // ---------- unstaged tile: same steps from global memory; a1 / a2 / alpha~ go through the alpha / elog buffers ----------
      // step 1+2 fused per destination row by one warp (rows of a huge tile: no shared-memory metadata)
      for (int i = r0 + wid; i < r1; i += kFCW) {
        const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
        auto dot2 = [&](int j, float& y1, float& y2) {
          const float4* row = reinterpret_cast<const float4*>(fbase + (int64_t)j * p.ldf);
          float x1 = 0.f, x2 = 0.f;
#pragma unroll
          for (int t = 0; t < NV; ++t) {
            const int q = lane + 32 * t;
            if (q < D4) {
              const float4 f = __ldg(row + q), l = s_l[q], r = s_r[q];
              x1 = fmaf(f.x, l.x, x1); x1 = fmaf(f.y, l.y, x1); x1 = fmaf(f.z, l.z, x1); x1 = fmaf(f.w, l.w, x1);
              x2 = fmaf(f.x, r.x, x2); x2 = fmaf(f.y, r.y, x2); x2 = fmaf(f.z, r.z, x2); x2 = fmaf(f.w, r.w, x2);
            }
          }
          y1 = warp_sum(x1); y2 = warp_sum(x2);
        };
        float d1, a2i;
        dot2(i, d1, a2i);
        float mx = -INFINITY;
        for (int k = beg; k < end; ++k) {
          float a1j, d2;
          dot2(__ldg(p.in_src + k), a1j, d2);
          float s = a1j + a2i;
          s = s > 0.f ? s : s * p.neg_slope;
          if (lane == 0) p.elog[(int64_t)k * H + h] = s;
          mx = fmaxf(mx, s);
        }
        __syncwarp();
        float l = 0.f;
        for (int k = beg + lane; k < end; k += 32) l += expf(p.elog[(int64_t)k * H + h] - mx);
        l = warp_sum(l);
        const float inv_l = end > beg ? 1.f / l : 0.f;
        for (int k = beg + lane; k < end; k += 32) {
          const int64_t o = (int64_t)k * H + h;
          const float al = expf(p.elog[o] - mx) * inv_l;
          p.alpha[o] = al;
          if (attn_drop)
            p.alpha_d[o] = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? al * p.attn_inv_keep : 0.f;
        }
      }
      fcompute_sync();
      if (col_thread) {
        for (int i = r0 + hgrp; i < r1; i += groups) {
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          for (int k = __ldg(p.in_ptr + i); k < __ldg(p.in_ptr + i + 1); ++k) {
            const float w = p.alpha_d[(int64_t)k * H + h];
            const float4* row = reinterpret_cast<const float4*>(fbase + (int64_t)__ldg(p.in_src + k) * p.ldf);
            ffma4(w, __ldg(row + hcol), acc0);
            ffma4(w, __ldg(row + c1), acc1);
          }
          fwd_store4(p, scale16, i, h, hcol, acc0);
          if (has2) fwd_store4(p, scale16, i, h, c1, acc1);
        }
      }
    }
    // position-embedding append + zero padding (once per row: the CTAs of the last head)
    if (p.hidden && h == H - 1 && nrows > 0) {
      const int feat = H * D, pd = p.pos_dim;
      const int c_end = p.out16_hi ? (int)p.ld16 : (int)p.ldo;
      const int wcols = c_end - feat;
      for (int t = ctid; t < nrows * wcols; t += kFCT) {
        const int i = r0 + t / wcols, c = feat + t % wcols;
        float v = 0.f;
        if (c < feat + pd) {
          v = __ldg(p.next_pos_table + (int64_t)__ldg(p.pos + i) * pd + (c - feat));
          if (p.next_thr) v = drop_keep1(p.next_seed, p.next_stream, (uint64_t)((int64_t)i * p.ldo + c), p.next_thr) ? v * p.next_inv_keep : 0.f;
        }
        if (p.out16_hi) {
          const float x = fminf(fmaxf(v * scale16, -65504.f), 65504.f);
          const __half hh = __float2half_rn(x);
          p.out16_hi[(int64_t)i * p.ld16 + c] = hh;
          p.out16_lo[(int64_t)i * p.ld16 + c] = __float2half_rn(x - __half2float(hh));
        } else if (p.out_lo) {
          const float vh = frn_tf32(v);
          p.out[(int64_t)i * p.ldo + c] = vh;
          p.out_lo[(int64_t)i * p.ldo + c] = frn_tf32(v - vh);
        } else {
          p.out[(int64_t)i * p.ldo + c] = v;
        }
      }
    }
    __syncwarp();
    if (lane == 0) fbar_arrive(empty0 + 8 * b);
  }
}

static int fwd_fixed_smem(int nv) { return 64 + 2 * kFwdMetaBytes + 2 * nv * 32 * 16; }
static int fwd_ring_rows(int64_t dim) {
  const int nv = (int)((dim + 127) / 128);
  int64_t rows = (kFMaxSmem - fwd_fixed_smem(nv)) / (dim * 4);
  if (rows > 2 * kFMR) rows = 2 * kFMR;
  return (int)rows;
}

}  // namespace tx

using namespace tx;

extern "C" {

int tx_gat_fused_fwd_staged(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                            const int32_t* in_src, const int32_t* in_eid, const int32_t* tiles, int64_t n_nodes, int64_t heads,
                            int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                            float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                            float* out_lo, void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out,
                            void* stream) {
  TX_REQUIRE(epi && tiles && aligned16(tiles), "gat_fused_fwd_staged: epilogue and tile table required");
  TX_REQUIRE(tx_gat_fused_supported(heads, dim, epi->mean_heads), "gat_fused_fwd_staged: unsupported shape (heads %lld dim %lld)", (long long)heads, (long long)dim);
  TX_REQUIRE(aligned16(ft) && ldf % 4 == 0 && (out16_hi || (out && aligned16(out))) && ldo % 4 == 0 && aligned16(attn_l) && aligned16(attn_r) &&
             (!out_lo || aligned16(out_lo)), "gat_fused_fwd_staged: 16-byte aligned rows required");
  TX_REQUIRE(!out16_hi || (!epi->mean_heads && out16_lo && bound && aligned16(out16_hi) && aligned16(out16_lo) && ld16 % 8 == 0 &&
                           ld16 >= heads * dim + epi->pos_dim), "gat_fused_fwd_staged: bad fp16 output buffers");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f && epi->p_drop >= 0.f && epi->p_drop < 1.f, "gat_fused_fwd_staged: dropout rates must be in [0,1)");
  TX_REQUIRE(alpha && elog && (p_attn == 0.f || (alpha_d && alpha_d != alpha)), "gat_fused_fwd_staged: alpha/elog/alpha_d buffers");
  const int64_t need = epi->mean_heads ? dim : heads * dim + epi->pos_dim;
  TX_REQUIRE(ldo >= need, "gat_fused_fwd_staged: ldo %lld < %lld", (long long)ldo, (long long)need);
  TX_REQUIRE(epi->pos_dim == 0 || (epi->next_pos_table && epi->pos), "gat_fused_fwd_staged: pos_dim > 0 needs next_pos_table and pos");
  if (n_nodes == 0) return TX_OK;
  StagedFwdParams p;
  p.ft = ft; p.ldf = ldf; p.attn_l = attn_l; p.attn_r = attn_r; p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid;
  p.tiles = reinterpret_cast<const int4*>(tiles); p.n_tiles = (int)tx_gat_bwd_num_tiles(n_nodes, dim); p.ring_rows = fwd_ring_rows(dim);
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.out = out; p.ldo = ldo; p.out_lo = out_lo;
  p.out16_hi = (__half*)out16_hi; p.out16_lo = (__half*)out16_lo; p.ld16 = ld16; p.bound = bound; p.scale_out = scale_out;
  p.maskbits = maskbits; p.mask_ld = (int)tx_gat_fused_mask_ld(heads, dim);
  p.hidden = epi->mean_heads ? 0 : 1; p.act_slope = epi->act_slope; p.next_pos_table = epi->next_pos_table; p.pos = epi->pos;
  p.pos_dim = (int)epi->pos_dim; p.next_inv_keep = 1.f / (1.f - epi->p_drop); p.next_thr = drop_threshold(epi->p_drop);
  p.next_seed = epi->seed; p.next_stream = epi->stream_id;
  const int nv = (int)((dim + 127) / 128);
  const size_t smem = (size_t)fwd_fixed_smem(nv) + (size_t)p.ring_rows * dim * 4;
  TX_REQUIRE(smem <= (size_t)kFMaxSmem && p.ring_rows >= 1, "gat_fused_fwd_staged: shared-memory budget exceeded");
  static bool attr_set[5] = {false, false, false, false, false};
  if (!attr_set[nv]) {
    cudaError_t e = cudaSuccess;
    switch (nv) {
      case 1: e = cudaFuncSetAttribute(gat_fwd_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFMaxSmem); break;
      case 2: e = cudaFuncSetAttribute(gat_fwd_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFMaxSmem); break;
      case 3: e = cudaFuncSetAttribute(gat_fwd_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFMaxSmem); break;
      default: e = cudaFuncSetAttribute(gat_fwd_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFMaxSmem); break;
    }
    if (e != cudaSuccess) {
      set_error("gat_fused_fwd_staged: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
    attr_set[nv] = true;
  }
  dim3 grid((unsigned)tx_gat_fused_bwd_staged_blocks(n_nodes, heads, dim), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nv) {
    case 1: gat_fwd_staged_kernel<1><<<grid, kFwdThreads, smem, st>>>(p); break;
    case 2: gat_fwd_staged_kernel<2><<<grid, kFwdThreads, smem, st>>>(p); break;
    case 3: gat_fwd_staged_kernel<3><<<grid, kFwdThreads, smem, st>>>(p); break;
    default: gat_fwd_staged_kernel<4><<<grid, kFwdThreads, smem, st>>>(p); break;
  }
  TX_LAUNCH_CHECK("tx_gat_fused_fwd_staged");
  return TX_OK;
}

}  // extern "C"
