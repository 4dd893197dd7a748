// TMA-staged fused GAT backward (sm_100a): the rows of a tile of whole graphs are bulk-copied (cp.async.bulk, mbarrier
// complete_tx) into shared memory by 4 producer warps while 12 compute warps work on the previous tile.
//
// Same arithmetic as gat_fused_bwd_kernel (tx_fused.cu; reference: autograd of model_zoo.py:83-96,106-114):
//   per in-edge j->i  : d(alpha~)_ij = <g_i, ft_j>
//   per destination i : softmax + leaky-relu backward -> ds_ij, da2_i = sum_j ds_ij
//   per source j      : dft_j = sum_i alpha~_ij g_i + da1_j attn_l + da2_j attn_r,  d(attn_l/r) += da1_j/da2_j ft_j
// What changed is where the rows come from and how the work is dealt.  The first kernel had every warp chase
// in_ptr -> in_src -> row through L1/L2 with at most two rows in flight per warp and was latency bound at ~18 % of HBM peak
// (profiles/r1_*).  Here
//   * a tile = the graphs whose first row lies in a window of R rows (precomputed by tx_gat_bwd_tiles: row / in-edge / out-edge
//     offsets per tile); every CTA owns a contiguous range of tiles (balanced by rows);
//   * the producer warps issue one bulk copy per g row and per ft row of the NEXT tile (D x 4 bytes each, 16-byte aligned) into a
//     shared-memory ring and stage the tile's CSR slices / attention coefficients next to them with all metadata loads in flight
//     at once; the two pipeline stages grow from opposite ends of ONE ring, so tile sizes may vary freely as long as two
//     consecutive tiles fit together (else the producer waits for the older tile to retire).  A tile with a large graph (more rows
//     than half the ring holds) stages its ft rows over the whole ring for the dot products (g from L2) and then its g rows over
//     them for the per-source step;
//   * compute warps: dot products are dealt per EDGE (two per warp and step), the softmax backward per destination row (one
//     thread each), the per-source step per (row group, pair of float4 columns) - every step is balanced for the 1..57-row
//     egonets, needs no warp shuffles outside the dots, and keeps the d(attn) accumulators in 16 registers per thread;
//   * dft is written (optionally TF32-split) straight to global memory.
// DRAM traffic is the algorithmic "read g once, read ft once, write dft once".  No float atomics; all reductions fixed-order.
// Measured (MAG-CS, L0): 0.735 ms -> 0.334 ms; what bounds it now is the LSU / shared-memory data pipe (every staged byte is
// read ~3.6x), see DESIGN.md section 4.
#include <stdlib.h>

#include <type_traits>

#include "tx_common.cuh"

namespace tx {

#ifndef TX_BWD_CW
#define TX_BWD_CW 12
#endif
constexpr int kCW = TX_BWD_CW;           // compute warps
constexpr int kPW = 4;                   // producer warps (16 warps = 4 per scheduler -> 128 registers each)
constexpr int kCT = kCW * 32;            // compute threads (D = 500: 3 row groups x 125 float4 columns = 375 of 384 busy in phase B)
constexpr int kBwdThreads = kCT + kPW * 32;
constexpr int kMR = 80;                  // staged tile metadata capacity: rows ...
constexpr int kME = 160;                 // ... and in-/out-edges
constexpr int kMaxSmem = 227 * 1024;

struct TileMeta {
  int r0, r1, s0, s1, o0, o1, mode, pad;
  int in_ptr[kMR + 1], out_ptr[kMR + 1];                   // tile-local edge offsets per row
  int in_src[kME], in_dst[kME], out_dst[kME], out_slot[kME];   // tile-local node / in-edge-slot indices
  float alpha[kME], alphad[kME], elog[kME], keepw[kME];
  float dd[kME], ds[kME], da2[kMR];                        // raw dots, d(logit), per-destination sums (compute warps only)
};
constexpr int kMetaBytes = (int)((sizeof(TileMeta) + 15) / 16 * 16);

struct StagedBwdParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const float* ft; int64_t ldf;
  const float* alpha; const float* alpha_d; const float* elog;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  const int32_t* out_ptr; const int32_t* out_dst; const int32_t* out_slot;
  const int4* tiles; int n_tiles; int ring_rows;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* ds; float* da2;
  float* dft; int64_t ldd; float* dft_lo;
  // optional fp16-split output (operand of the weight / input gradient GEMMs): dft * scale = hi + lo; replaces dft / dft_lo
  __half* dft16_hi; __half* dft16_lo; int64_t ld16; const float* bound; float* scale_out;
  float* dattn_partial;   // [gridDim.x, 2, H, D]
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {     // no arrive: the producer arrives once the metadata is staged
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
// one row: global -> shared bulk copy through the TMA engine, completion bytes credited to `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }

// round to nearest (ties away) TF32: the same value as tx_split_tf32's (bits + 0x1000) & ~0x1fff for finite inputs, one instruction
__device__ __forceinline__ float rn_tf32_b(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void store_dft4(float* hi_ptr, float* lo_ptr, float4 v) {
  if (lo_ptr) {
    const float4 h = make_float4(rn_tf32_b(v.x), rn_tf32_b(v.y), rn_tf32_b(v.z), rn_tf32_b(v.w));
    *reinterpret_cast<float4*>(hi_ptr) = h;
    *reinterpret_cast<float4*>(lo_ptr) = make_float4(rn_tf32_b(v.x - h.x), rn_tf32_b(v.y - h.y), rn_tf32_b(v.z - h.z), rn_tf32_b(v.w - h.w));
  } else {
    *reinterpret_cast<float4*>(hi_ptr) = v;
  }
}

// one float4 of dft at (row, column c of head h): fp32, TF32 hi/lo or fp16 hi/lo
__device__ __forceinline__ void store_dft(const StagedBwdParams& p, float scale16, int64_t row, int h, int c, float4 v) {
  if (p.dft16_hi) {
    uint2 h16, l16;
    f16_split4(v, scale16, h16, l16);
    const int64_t o = row * p.ld16 + (int64_t)h * p.D + c;
    *reinterpret_cast<uint2*>(p.dft16_hi + o) = h16;
    *reinterpret_cast<uint2*>(p.dft16_lo + o) = l16;
  } else {
    const int64_t off = row * p.ldd + (int64_t)h * p.D + c;
    store_dft4(p.dft + off, p.dft_lo ? p.dft_lo + off : nullptr, v);
  }
}

template <int NV>
__device__ __forceinline__ void ld_row(const float4* __restrict__ row, int lane, int D4, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) v[t] = (lane + 32 * t) < D4 ? row[lane + 32 * t] : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int NV>
__device__ __forceinline__ float dot4(const float4 (&a)[NV], const float4 (&b)[NV]) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    acc = fmaf(a[t].x, b[t].x, acc); acc = fmaf(a[t].y, b[t].y, acc); acc = fmaf(a[t].z, b[t].z, acc); acc = fmaf(a[t].w, b[t].w, acc);
  }
  return acc;
}
template <int NV>
__device__ __forceinline__ void axpy4(float w, const float4 (&x)[NV], float4 (&y)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    y[t].x = fmaf(w, x[t].x, y[t].x); y[t].y = fmaf(w, x[t].y, y[t].y);
    y[t].z = fmaf(w, x[t].z, y[t].z); y[t].w = fmaf(w, x[t].w, y[t].w);
  }
}
__device__ __forceinline__ void fma4(float w, const float4 x, float4& y) {
  y.x = fmaf(w, x.x, y.x); y.y = fmaf(w, x.y, y.y); y.z = fmaf(w, x.z, y.z); y.w = fmaf(w, x.w, y.w);
}
// two independent warp sums with interleaved shuffles (halves the exposed shuffle latency per pair of edges)
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a += ta; b += tb;
  }
}

#ifdef TX_BWD_PROFILE
// debug build only: per CTA, per staging mode: {tiles, rows, cycles waiting for the producer, dots, softmax backward, phase B}
__device__ long long g_bwd_prof[148 * 4 * 3 * 6];
#define TX_PROF_T(var) const long long var = clock64()
#else
#define TX_PROF_T(var)
#endif

template <int NV>
__global__ void __launch_bounds__(kBwdThreads, 1) gat_bwd_staged_kernel(const StagedBwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                 // full[2], empty[2], mid_full, mid_empty
  TileMeta* meta = reinterpret_cast<TileMeta*>(smem + 64);
  float4* s_l = reinterpret_cast<float4*>(smem + 64 + 2 * kMetaBytes);
  float4* s_r = s_l + NV * 32;
  uint8_t* ring_g = reinterpret_cast<uint8_t*>(s_r + NV * 32);
  const int H = p.H, D = p.D, D4 = D >> 2;
  const uint32_t rowB = (uint32_t)D * 4u;
  uint8_t* ring_f = ring_g + (size_t)p.ring_rows * rowB;
  const int h = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t full0 = s_addr(bars), empty0 = s_addr(bars + 2), mid_full = s_addr(bars + 4), mid_empty = s_addr(bars + 5);

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      bar_init(full0 + 8 * b, kPW);
      bar_init(empty0 + 8 * b, kCW);
    }
    bar_init(mid_full, kPW);
    bar_init(mid_empty, kCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const float* gbase = p.g + (int64_t)h * p.g_head_stride;
  const float* fbase = p.ft + (int64_t)h * D;
  const bool attn_drop = p.attn_thr != 0;
  // every CTA owns a CONTIGUOUS range of tiles: windows hold ~R rows each, so ranges are balanced by rows even where a large graph
  // fills one tile and leaves the following windows empty (round-robin left some CTAs with most of the large tiles)
  const int tile_beg = (int)((int64_t)blockIdx.x * p.n_tiles / gridDim.x), tile_end = (int)((int64_t)(blockIdx.x + 1) * p.n_tiles / gridDim.x);

  if (wid >= kCW) {
    // =========================== producer warps ===========================
    // kPW warps share the work of staging one tile: every thread issues the bulk copies of "its" rows and fetches "its" slice of
    // the tile metadata with all global loads in flight at once (one memory latency per tile instead of a chain of them)
    const int pw = wid - kCW;
    const int pt = lane * kPW + pw;          // rows are dealt round-robin over the warps: row r -> warp r % kPW, lane r / kPW
    int prev_rows = 0;      // ring rows held by the previous tile (0 if it staged nothing)
    int it = 0, n2 = 0;     // n2: mode-2 tiles so far (phase of the mid barriers)
    int4 t0 = make_int4(0, 0, 0, 0), t1 = t0;
    if (tile_beg < tile_end) { t0 = __ldg(p.tiles + tile_beg); t1 = __ldg(p.tiles + tile_beg + 1); }
    for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
      const int b = it & 1;
      const int r0 = t0.x, r1 = t1.x, s0 = t0.y, s1 = t1.y, o0 = t0.z, o1 = t1.z;
      const int nrows = r1 - r0, nE = s1 - s0, nO = o1 - o0;
      if (tile + 1 < tile_end) {     // header of the next tile: in flight while this one is staged
        t0 = t1;
        t1 = __ldg(p.tiles + tile + 2);
      }
      // mode 1: g and ft rows of the tile staged side by side (tile fits half the ring);
      // mode 2: a tile with a large graph: the ft rows are staged over the WHOLE ring for the dot products (g comes from L2, prefetched
      //         here), then - once the compute warps are done with ft - the g rows are staged over them for phase B;
      // mode 0: metadata does not fit: nothing staged, the compute warps read global memory
      const bool meta_ok = nrows > 0 && nrows <= kMR && nE <= kME && nO <= kME;
      const int mode = !meta_ok ? 0 : (nrows <= p.ring_rows ? 1 : (nrows <= 2 * p.ring_rows ? 2 : 0));
      const int span = mode == 1 ? nrows : (mode == 2 ? p.ring_rows + 1 : 0);                   // mode 2 needs the ring exclusively
      if (it >= 2) bar_wait(empty0 + 8 * b, (uint32_t)(((it >> 1) - 1) & 1));                 // tile it-2 retired: stage b is free
      if (span > 0 && prev_rows + span > p.ring_rows) bar_wait(empty0 + 8 * (b ^ 1), (uint32_t)(((it - 1) >> 1) & 1));   // ring too full: tile it-1 too
      TileMeta& m = meta[b];
      const uint32_t full = full0 + 8 * b;
      if (pw == 0 && lane == 0) {
        m.r0 = r0; m.r1 = r1; m.s0 = s0; m.s1 = s1; m.o0 = o0; m.o1 = o1; m.mode = mode;
      }
      if (mode) {
        // ---- metadata loads first (all in flight together), then the row copies, then the shared-memory stores ----
        const bool hr = pt <= nrows, he = pt < nE, ho = pt < nO;
        int v_ip = 0, v_ip1 = 0, v_op = 0, v_src = 0, v_eid = 0, v_od = 0, v_os = 0;
        float v_al = 0.f, v_ad = 0.f, v_el = 0.f;
        if (hr) {
          v_ip = __ldg(p.in_ptr + r0 + pt);
          v_op = __ldg(p.out_ptr + r0 + pt);
          if (pt < nrows) v_ip1 = __ldg(p.in_ptr + r0 + pt + 1);
        }
        if (he) {
          const int64_t o = (int64_t)(s0 + pt) * H + h;
          v_src = __ldg(p.in_src + s0 + pt);
          v_al = __ldg(p.alpha + o);
          v_ad = __ldg(p.alpha_d + o);
          v_el = __ldg(p.elog + o);
          if (attn_drop) v_eid = __ldg(p.in_eid + s0 + pt);
        }
        if (ho) {
          v_od = __ldg(p.out_dst + o0 + pt);
          v_os = __ldg(p.out_slot + o0 + pt);
        }
        // row copies: this warp's rows are pw, pw + kPW, ...; lane 0 posts their bytes first
        const int my_rows = nrows > pw ? (nrows - pw + kPW - 1) / kPW : 0;
        if (lane == 0 && my_rows > 0) bar_expect_tx(full, (uint32_t)my_rows * (mode == 1 ? 2u : 1u) * rowB);
        __syncwarp();
        for (int r = pt; r < nrows; r += 32 * kPW) {
          if (mode == 1) {
            const int slot = b ? p.ring_rows - 1 - r : r;
            bulk_g2s(s_addr(ring_g + (size_t)slot * rowB), gbase + (int64_t)(r0 + r) * p.ldg, rowB, full);
            bulk_g2s(s_addr(ring_f + (size_t)slot * rowB), fbase + (int64_t)(r0 + r) * p.ldf, rowB, full);
          } else {
            bulk_g2s(s_addr(ring_g + (size_t)r * rowB), fbase + (int64_t)(r0 + r) * p.ldf, rowB, full);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gbase + (int64_t)(r0 + r) * p.ldg), "r"(rowB) : "memory");
          }
        }
        // tile metadata, indices made tile-local (every edge of a tile of whole graphs is internal)
        auto put_row = [&](int t, int ip, int ip1, int op) {
          m.in_ptr[t] = ip - s0;
          m.out_ptr[t] = op - o0;
          if (t < nrows) for (int k = ip - s0; k < ip1 - s0; ++k) m.in_dst[k] = t;
        };
        auto put_in = [&](int t, int src, float al, float ad, float el, int eid) {
          m.in_src[t] = src - r0;
          m.alpha[t] = al;
          m.alphad[t] = ad * p.g_scale;
          m.elog[t] = el;
          m.keepw[t] = !attn_drop ? 1.f : (drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)eid * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f);
        };
        if (hr) put_row(pt, v_ip, v_ip1, v_op);
        if (he) put_in(pt, v_src, v_al, v_ad, v_el, v_eid);
        if (ho) { m.out_dst[pt] = v_od - r0; m.out_slot[pt] = v_os - s0; }
        for (int t = pt + 32 * kPW; t <= nrows; t += 32 * kPW)      // rare: tiles with more rows / edges than producer threads
          put_row(t, __ldg(p.in_ptr + r0 + t), t < nrows ? __ldg(p.in_ptr + r0 + t + 1) : 0, __ldg(p.out_ptr + r0 + t));
        for (int t = pt + 32 * kPW; t < nE; t += 32 * kPW) {
          const int64_t o = (int64_t)(s0 + t) * H + h;
          put_in(t, __ldg(p.in_src + s0 + t), __ldg(p.alpha + o), __ldg(p.alpha_d + o), __ldg(p.elog + o), attn_drop ? __ldg(p.in_eid + s0 + t) : 0);
        }
        for (int t = pt + 32 * kPW; t < nO; t += 32 * kPW) {
          m.out_dst[t] = __ldg(p.out_dst + o0 + t) - r0;
          m.out_slot[t] = __ldg(p.out_slot + o0 + t) - s0;
        }
      }
      __syncwarp();
      if (lane == 0) bar_arrive(full);
      if (mode == 2) {     // second stage of a large tile: g rows over the ft rows, as soon as the compute warps release them
        bar_wait(mid_empty, (uint32_t)(n2 & 1));
        const int my_rows = nrows > pw ? (nrows - pw + kPW - 1) / kPW : 0;
        if (lane == 0 && my_rows > 0) bar_expect_tx(mid_full, (uint32_t)my_rows * rowB);
        __syncwarp();
        for (int r = pt; r < nrows; r += 32 * kPW) bulk_g2s(s_addr(ring_g + (size_t)r * rowB), gbase + (int64_t)(r0 + r) * p.ldg, rowB, mid_full);
        __syncwarp();
        if (lane == 0) bar_arrive(mid_full);
        ++n2;
      }
      prev_rows = span;
    }
    return;
  }

  // =========================== compute warps ===========================
  // phase B mapping: thread (hgrp, hcol) owns the float4 columns hcol and hcol + W (W = ceil(D4 / 2)) of the source rows
  // t = hgrp (mod groups) of every tile; it also accumulates d(attn_l), d(attn_r) for its columns over those rows
  float4 hvl[2], hvr[2];
  hvl[0] = hvl[1] = hvr[0] = hvr[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ctid = threadIdx.x;
  const int W = (D4 + 1) >> 1;
  const int groups = kCT / W;
  const int hgrp = ctid / W, hcol = ctid - hgrp * W;
  const bool col_thread = hgrp < groups;
  const bool has2 = hcol + W < D4;
  const float scale16 = p.dft16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.dft16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;
  const int c1 = has2 ? hcol + W : hcol;       // clamped: reads stay in range, the duplicate is never stored

#ifdef TX_BWD_PROFILE
  long long prof[3][6];
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 6; ++c) prof[a][c] = 0;
#endif
  int it = 0, n2 = 0;
  for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
    const int b = it & 1;
    TX_PROF_T(t_start);
    bar_wait(full0 + 8 * b, (uint32_t)((it >> 1) & 1));
    TX_PROF_T(t_ready);
    TileMeta& m = meta[b];
    const int r0 = m.r0, r1 = m.r1, mode = m.mode;
    const int nrows = r1 - r0;
#ifdef TX_BWD_PROFILE
    long long tA = t_ready, tS = t_ready;
#endif

    if (mode != 0) {
      // ---------- staged tile.  Step 1: one raw dot product <g_i, ft_j> per in-edge, two edges per warp and step ----------
      const int nE = m.s1 - m.s0;
      auto FP = [&](int j) -> const float4* {      // ft row of tile-local node j (always in shared memory)
        if (mode == 1) return reinterpret_cast<const float4*>(ring_f + (size_t)(b ? p.ring_rows - 1 - j : j) * rowB);
        return reinterpret_cast<const float4*>(ring_g + (size_t)j * rowB);
      };
      for (int e = wid; e < nE; e += 2 * kCW) {
        const int e2 = e + kCW;
        const bool two = e2 < nE;
        float4 ga[NV], ra[NV], gb[NV], rb[NV];
        if (mode == 1) {
          ld_row<NV>(reinterpret_cast<const float4*>(ring_g + (size_t)(b ? p.ring_rows - 1 - m.in_dst[e] : m.in_dst[e]) * rowB), lane, D4, ga);
          if (two) ld_row<NV>(reinterpret_cast<const float4*>(ring_g + (size_t)(b ? p.ring_rows - 1 - m.in_dst[e2] : m.in_dst[e2]) * rowB), lane, D4, gb);
        } else {
          ld_row<NV>(reinterpret_cast<const float4*>(gbase + (int64_t)(r0 + m.in_dst[e]) * p.ldg), lane, D4, ga);
          if (two) ld_row<NV>(reinterpret_cast<const float4*>(gbase + (int64_t)(r0 + m.in_dst[e2]) * p.ldg), lane, D4, gb);
        }
        ld_row<NV>(FP(m.in_src[e]), lane, D4, ra);
        if (two) ld_row<NV>(FP(m.in_src[e2]), lane, D4, rb);
        float da = dot4<NV>(ga, ra), db = two ? dot4<NV>(gb, rb) : 0.f;
        warp_sum2(da, db);
        if (lane == 0) {
          m.dd[e] = da * p.g_scale;
          if (two) m.dd[e2] = db * p.g_scale;
        }
      }
      compute_sync();
#ifdef TX_BWD_PROFILE
      tA = clock64();
#endif
      // ---------- step 2: edge-softmax + leaky-relu backward, one thread per destination row ----------
      if (ctid < nrows) {
        const int a = m.in_ptr[ctid], e = m.in_ptr[ctid + 1];
        float tsum = 0.f;
        for (int k = a; k < e; ++k) tsum = fmaf(m.alpha[k], m.dd[k] * m.keepw[k], tsum);
        float a2 = 0.f;
        for (int k = a; k < e; ++k) {
          const float de = m.alpha[k] * (m.dd[k] * m.keepw[k] - tsum);         // edge softmax backward
          const float dsv = m.elog[k] > 0.f ? de : de * p.neg_slope;           // leaky-relu(0.2) backward (model_zoo.py:108-109)
          m.ds[k] = dsv;
          a2 += dsv;
        }
        m.da2[ctid] = a2;
      }
      compute_sync();
#ifdef TX_BWD_PROFILE
      tS = clock64();
#endif
      if (mode == 2) {
        // d(attn) needs the ft rows: take them now, then let the producer stage the g rows over them
        if (col_thread) {
          for (int t = hgrp; t < nrows; t += groups) {
            float d1 = 0.f;
            for (int k = m.out_ptr[t]; k < m.out_ptr[t + 1]; ++k) d1 += m.ds[m.out_slot[k]];
            const float d2 = m.da2[t];
            const float4* frow = reinterpret_cast<const float4*>(ring_g + (size_t)t * rowB);
            fma4(d1, frow[hcol], hvl[0]); fma4(d2, frow[hcol], hvr[0]);
            fma4(d1, frow[c1], hvl[1]); fma4(d2, frow[c1], hvr[1]);
          }
        }
        __syncwarp();
        if (lane == 0) bar_arrive(mid_empty);
        bar_wait(mid_full, (uint32_t)(n2 & 1));
        ++n2;
      }
      // ---------- step 3 (phase B): one thread per (source row group, pair of float4 columns) ----------
      if (col_thread) {
        const float4 cl0 = s_l[hcol], cr0 = s_r[hcol], cl1 = s_l[c1], cr1 = s_r[c1];
        for (int t = hgrp; t < nrows; t += groups) {
          float d1 = 0.f;
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          const int ob = m.out_ptr[t], oe = m.out_ptr[t + 1];
          for (int k = ob; k < oe; ++k) {
            const int sl = m.out_slot[k];
            const int di = m.out_dst[k];
            d1 += m.ds[sl];                                     // da1_t = sum over out-edges of ds
            const float w = m.alphad[sl];                       // alpha~ (x g_scale)
            const float4* grow = reinterpret_cast<const float4*>(ring_g + (size_t)((mode == 1 && b) ? p.ring_rows - 1 - di : di) * rowB);
            fma4(w, grow[hcol], acc0);
            fma4(w, grow[c1], acc1);
          }
          const float d2 = m.da2[t];
          if (mode == 1) {
            const float4* frow = reinterpret_cast<const float4*>(ring_f + (size_t)(b ? p.ring_rows - 1 - t : t) * rowB);
            fma4(d1, frow[hcol], hvl[0]); fma4(d2, frow[hcol], hvr[0]);
            fma4(d1, frow[c1], hvl[1]); fma4(d2, frow[c1], hvr[1]);
          }
          fma4(d1, cl0, acc0); fma4(d2, cr0, acc0);
          fma4(d1, cl1, acc1); fma4(d2, cr1, acc1);
          store_dft(p, scale16, r0 + t, h, hcol * 4, acc0);
          if (has2) store_dft(p, scale16, r0 + t, h, c1 * 4, acc1);
        }
      }
    } else if (nrows > 0) {
      // ---------- unstaged tile (metadata does not fit): global memory, warp per destination row, then thread per column ----------
      for (int i = r0 + wid; i < r1; i += kCW) {
        const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
        float4 gi[NV];
        ld_row<NV>(reinterpret_cast<const float4*>(gbase + (int64_t)i * p.ldg), lane, D4, gi);
        for (int k = beg; k < end; ++k) {
          float4 ra[NV];
          ld_row<NV>(reinterpret_cast<const float4*>(fbase + (int64_t)__ldg(p.in_src + k) * p.ldf), lane, D4, ra);
          const float d = warp_sum(dot4<NV>(gi, ra)) * p.g_scale;
          if (lane == 0) p.ds[(int64_t)k * H + h] = d;
        }
        __syncwarp();
        auto KW = [&](int k) {
          return !attn_drop ? 1.f
               : (drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f);
        };
        float tsum = 0.f;
        for (int k = beg + lane; k < end; k += 32) tsum = fmaf(__ldg(p.alpha + (int64_t)k * H + h), p.ds[(int64_t)k * H + h] * KW(k), tsum);
        tsum = warp_sum(tsum);
        float a2 = 0.f;
        for (int k = beg + lane; k < end; k += 32) {
          const float de = __ldg(p.alpha + (int64_t)k * H + h) * (p.ds[(int64_t)k * H + h] * KW(k) - tsum);
          const float dsv = __ldg(p.elog + (int64_t)k * H + h) > 0.f ? de : de * p.neg_slope;
          p.ds[(int64_t)k * H + h] = dsv;
          a2 += dsv;
        }
        a2 = warp_sum(a2);
        if (lane == 0) p.da2[(int64_t)i * H + h] = a2;
      }
      compute_sync();
      if (col_thread) {
        const float4 cl0 = s_l[hcol], cr0 = s_r[hcol], cl1 = s_l[c1], cr1 = s_r[c1];
        for (int j = r0 + hgrp; j < r1; j += groups) {
          float d1 = 0.f;
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          for (int k = __ldg(p.out_ptr + j); k < __ldg(p.out_ptr + j + 1); ++k) {
            const int64_t sl = (int64_t)__ldg(p.out_slot + k) * H + h;
            d1 += p.ds[sl];
            const float w = __ldg(p.alpha_d + sl) * p.g_scale;
            const float4* grow = reinterpret_cast<const float4*>(gbase + (int64_t)__ldg(p.out_dst + k) * p.ldg);
            fma4(w, grow[hcol], acc0);
            fma4(w, grow[c1], acc1);
          }
          const float d2 = p.da2[(int64_t)j * H + h];
          const float4* frow = reinterpret_cast<const float4*>(fbase + (int64_t)j * p.ldf);
          fma4(d1, frow[hcol], hvl[0]); fma4(d2, frow[hcol], hvr[0]);
          fma4(d1, frow[c1], hvl[1]); fma4(d2, frow[c1], hvr[1]);
          fma4(d1, cl0, acc0); fma4(d2, cr0, acc0);
          fma4(d1, cl1, acc1); fma4(d2, cr1, acc1);
          store_dft(p, scale16, j, h, hcol * 4, acc0);
          if (has2) store_dft(p, scale16, j, h, c1 * 4, acc1);
        }
      }
    }
#ifdef TX_BWD_PROFILE
    {
      const long long tE = clock64();
      prof[mode][0] += 1; prof[mode][1] += nrows; prof[mode][2] += t_ready - t_start; prof[mode][3] += tA - t_ready;
      prof[mode][4] += tS - tA; prof[mode][5] += tE - tS;
    }
#endif
    __syncwarp();
    if (lane == 0) bar_arrive(empty0 + 8 * b);     // this warp is done with stage b (rows, metadata, dd / ds / da2)
  }

#ifdef TX_BWD_PROFILE
  if (threadIdx.x == 0)
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 6; ++c) g_bwd_prof[((blockIdx.y * gridDim.x + blockIdx.x) * 3 + a) * 6 + c] = prof[a][c];
#endif
  // ---- d(attn) partials: fixed-order reduction over the row groups (scratch aliases the ring: every tile of this CTA has been
  //      consumed, so no bulk copy is in flight) ----
  compute_sync();
  float4* scratch = reinterpret_cast<float4*>(ring_g);      // [groups][2][D4]
  if (col_thread) {
    scratch[(hgrp * 2 + 0) * D4 + hcol] = hvl[0];
    scratch[(hgrp * 2 + 1) * D4 + hcol] = hvr[0];
    if (has2) {
      scratch[(hgrp * 2 + 0) * D4 + c1] = hvl[1];
      scratch[(hgrp * 2 + 1) * D4 + c1] = hvr[1];
    }
  }
  compute_sync();
  for (int t = ctid; t < 2 * D4; t += kCT) {
    const int lr = t / D4, q = t - lr * D4;
    float4 s = scratch[lr * D4 + q];
    for (int g = 1; g < groups; ++g) {
      const float4 x = scratch[(g * 2 + lr) * D4 + q];
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    *reinterpret_cast<float4*>(p.dattn_partial + (((int64_t)blockIdx.x * 2 + lr) * H + h) * D + q * 4) = s;
  }
}

// tile t = graphs whose first row lies in [t R, (t+1) R): tiles[t] = {first row, first in-edge, first out-edge, 0}
__global__ void bwd_tiles_kernel(const int32_t* __restrict__ node_off, int n_graphs, int n, const int32_t* __restrict__ in_ptr,
                                 const int32_t* __restrict__ out_ptr, int R, int n_tiles, int4* __restrict__ tiles) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  int r = n;
  if (t < n_tiles) {
    int lo = 0, hi = n_graphs + 1;
    const int value = t * R;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(node_off + mid) < value) lo = mid + 1; else hi = mid;
    }
    r = __ldg(node_off + min(lo, n_graphs));
  }
  tiles[t] = make_int4(r, __ldg(in_ptr + r), __ldg(out_ptr + r), 0);
}

static int fixed_smem(int nv) { return 64 + 2 * kMetaBytes + 2 * nv * 32 * 16; }
static int ring_rows_for(int64_t dim) {
  const int nv = (int)((dim + 127) / 128);
  const int64_t row_pair = 2 * dim * 4;
  int64_t rows = (kMaxSmem - fixed_smem(nv)) / row_pair;
  if (rows > 2 * kMR) rows = 2 * kMR;
  return (int)rows;
}

}  // namespace tx

using namespace tx;

extern "C" {

#ifdef TX_BWD_PROFILE
int tx_debug_bwd_prof(long long* host_out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, g_bwd_prof, sizeof(long long) * 148 * 4 * 3 * 6);
}
#endif

int64_t tx_gat_bwd_tile_rows(int64_t dim) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("TAXO_BWD2_TILE_ROWS"); env = e ? atoi(e) : 0; }
  if (env > 0) return env;
  int64_t r = ring_rows_for(dim) / 3;
  if (r > 32) r = 32;
  if (r < 4) r = 4;
  return r;
}

int64_t tx_gat_bwd_num_tiles(int64_t n_nodes, int64_t dim) {
  const int64_t R = tx_gat_bwd_tile_rows(dim);
  return (n_nodes + R - 1) / R;
}

int tx_gat_bwd_tiles(const int32_t* node_off, int64_t n_graphs, int64_t n_nodes, const int32_t* in_ptr, const int32_t* out_ptr,
                     int64_t dim, int32_t* tiles, void* stream) {
  TX_REQUIRE(node_off && in_ptr && out_ptr && tiles && aligned16(tiles), "gat_bwd_tiles: bad arguments");
  TX_REQUIRE(n_nodes >= 0 && n_nodes < INT32_MAX && n_graphs >= 0, "gat_bwd_tiles: bad shape");
  const int R = (int)tx_gat_bwd_tile_rows(dim);
  const int n_tiles = (int)tx_gat_bwd_num_tiles(n_nodes, dim);
  bwd_tiles_kernel<<<(n_tiles + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(node_off, (int)n_graphs, (int)n_nodes, in_ptr, out_ptr, R,
                                                                                  n_tiles, reinterpret_cast<int4*>(tiles));
  TX_LAUNCH_CHECK("tx_gat_bwd_tiles");
  return TX_OK;
}

int64_t tx_gat_fused_bwd_staged_blocks(int64_t n_nodes, int64_t heads, int64_t dim) {
  const int64_t tiles = tx_gat_bwd_num_tiles(n_nodes, dim);
  int64_t gx = ((int64_t)kNumSms + heads - 1) / heads;      // one CTA per SM over all heads
  if (gx > tiles) gx = tiles;
  return gx < 1 ? 1 : gx;
}

int tx_gat_fused_bwd_staged(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft, int64_t ldf,
                            const float* alpha, const float* alpha_d, const float* elog, const float* attn_l,
                            const float* attn_r, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                            const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_slot, const int32_t* tiles,
                            int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed,
                            uint32_t attn_stream_id, float* ds, float* da2, float* dft, int64_t ldd, float* dft_lo,
                            void* dft16_hi, void* dft16_lo, int64_t ld16, const float* bound, float* scale_out,
                            float* dattn_partial, void* stream) {
  TX_REQUIRE(g_head_stride != 0 || heads == 1, "gat_fused_bwd_staged: a shared g row (head mean) needs heads == 1");
  TX_REQUIRE(!dft16_hi || (dft16_lo && bound && aligned16(dft16_hi) && aligned16(dft16_lo) && ld16 % 8 == 0 && ld16 >= heads * dim),
             "gat_fused_bwd_staged: bad fp16 output buffers");
  TX_REQUIRE(dft16_hi || dft, "gat_fused_bwd_staged: an output buffer is required");
  TX_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 512, "gat_fused_bwd_staged: dim must be a multiple of 4 and <= 512");
  TX_REQUIRE(aligned16(g) && ldg % 4 == 0 && g_head_stride % 4 == 0 && aligned16(ft) && ldf % 4 == 0 && (!dft || aligned16(dft)) && ldd % 4 == 0 &&
             aligned16(attn_l) && aligned16(attn_r) && aligned16(dattn_partial) && (!dft_lo || aligned16(dft_lo)) && aligned16(tiles),
             "gat_fused_bwd_staged: 16-byte aligned rows required");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f, "gat_fused_bwd_staged: dropout rate must be in [0,1)");
  TX_REQUIRE(tiles && ds && da2, "gat_fused_bwd_staged: tiles / scratch required");
  if (n_nodes == 0) return TX_OK;
  StagedBwdParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.ft = ft; p.ldf = ldf;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.attn_l = attn_l; p.attn_r = attn_r;
  p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid; p.out_ptr = out_ptr; p.out_dst = out_dst; p.out_slot = out_slot;
  p.tiles = reinterpret_cast<const int4*>(tiles); p.n_tiles = (int)tx_gat_bwd_num_tiles(n_nodes, dim);
  p.ring_rows = ring_rows_for(dim);
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim;
  p.neg_slope = neg_slope; p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed;
  p.attn_stream = attn_stream_id; p.ds = ds; p.da2 = da2; p.dft = dft; p.ldd = ldd; p.dft_lo = dft_lo; p.dattn_partial = dattn_partial;
  p.dft16_hi = (__half*)dft16_hi; p.dft16_lo = (__half*)dft16_lo; p.ld16 = ld16; p.bound = bound; p.scale_out = scale_out;
  const int nv = (int)((dim + 127) / 128);
  size_t ring_bytes = (size_t)p.ring_rows * 2 * dim * 4;
  const size_t scratch_bytes = (size_t)kCT * 4 * 16;   // [groups][2][D4] float4, groups * D4 <= 2 kCT
  if (ring_bytes < scratch_bytes) ring_bytes = scratch_bytes;
  const size_t smem = (size_t)fixed_smem(nv) + ring_bytes;
  TX_REQUIRE(smem <= (size_t)kMaxSmem && p.ring_rows >= 1, "gat_fused_bwd_staged: shared-memory budget exceeded");
  static bool attr_set[5] = {false, false, false, false, false};
  if (!attr_set[nv]) {
    cudaError_t e = cudaSuccess;
    switch (nv) {
      case 1: e = cudaFuncSetAttribute(gat_bwd_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); break;
      case 2: e = cudaFuncSetAttribute(gat_bwd_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); break;
      case 3: e = cudaFuncSetAttribute(gat_bwd_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); break;
      default: e = cudaFuncSetAttribute(gat_bwd_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem); break;
    }
    if (e != cudaSuccess) {
      set_error("gat_fused_bwd_staged: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
    attr_set[nv] = true;
  }
  dim3 grid((unsigned)tx_gat_fused_bwd_staged_blocks(n_nodes, heads, dim), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nv) {
    case 1: gat_bwd_staged_kernel<1><<<grid, kBwdThreads, smem, st>>>(p); break;
    case 2: gat_bwd_staged_kernel<2><<<grid, kBwdThreads, smem, st>>>(p); break;
    case 3: gat_bwd_staged_kernel<3><<<grid, kBwdThreads, smem, st>>>(p); break;
    default: gat_bwd_staged_kernel<4><<<grid, kBwdThreads, smem, st>>>(p); break;
  }
  TX_LAUNCH_CHECK("tx_gat_fused_bwd_staged");
  return TX_OK;
}

}  // extern "C"
