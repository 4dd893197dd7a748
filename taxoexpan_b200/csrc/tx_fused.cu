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

#include "tx_common.cuh"

namespace tx {

constexpr int kTileRows = 16;   // backward tile window (rows); tiles are graph-aligned
constexpr int kMaxNV = 4;       // up to 4 float4 per lane per head row -> D' <= 512

template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, int D, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    v[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
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

struct FusedFwdParams {
  const float* ft; int64_t ldf;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* alpha; float* alpha_d; float* elog;
  float* out; int64_t ldo;
  uint32_t* maskbits;      // [n, H, NV, 8] words: 4 sign planes (one per float4 component) + 4 keep planes; may be null
  // epilogue
  int hidden; float act_slope; const float* next_pos_table; const int32_t* pos; int pos_dim;
  float next_inv_keep; uint32_t next_thr; uint64_t next_seed; uint32_t next_stream;
};

template <int NV>
__global__ void __launch_bounds__(256) gat_fused_fwd_kernel(const FusedFwdParams p) {
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
  for (int i = warp; i < p.n; i += nwarps) {
    const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
    const float* base = p.ft + (int64_t)h * D;
    float4 own[NV];
    load_row<NV>(base + (int64_t)i * p.ldf, lane, D, own);
    const float a2i = warp_sum(dot_row<NV>(own, s_r, lane));
    // ---- pass 1: logits of the in-edges (a1 from the gathered rows), kept lane-distributed ----
    for (int k = beg; k < end; ++k) {
      const int j = __ldg(p.in_src + k);
      float d;
      if (j == i) {
        d = dot_row<NV>(own, s_l, lane);
      } else {
        float4 fj[NV];
        load_row<NV>(base + (int64_t)j * p.ldf, lane, D, fj);
        d = dot_row<NV>(fj, s_l, lane);
      }
      float s = warp_sum(d) + a2i;
      s = s > 0.f ? s : s * p.neg_slope;
      if (lane == 0) p.elog[(int64_t)k * H + h] = s;
    }
    __syncwarp();
    // ---- edge softmax, lane-parallel over the in-edges ----
    float m = -INFINITY;
    for (int c = beg; c < end; c += 32) {
      const int k = c + lane;
      if (k < end) m = fmaxf(m, p.elog[(int64_t)k * H + h]);
    }
    m = warp_max(m);
    float l = 0.f;
    for (int c = beg; c < end; c += 32) {
      const int k = c + lane;
      if (k < end) l += expf(p.elog[(int64_t)k * H + h] - m);
    }
    l = warp_sum(l);
    // ---- pass 2: attention dropout + weighted aggregation ----
    float4 acc[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = beg; c < end; c += 32) {
      const int k = c + lane;
      float ad = 0.f;
      if (k < end) {
        const float a = expf(p.elog[(int64_t)k * H + h] - m) / l;
        p.alpha[(int64_t)k * H + h] = a;
        ad = a;
        if (attn_drop) {
          const bool keep = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr);
          ad = keep ? a * p.attn_inv_keep : 0.f;
          p.alpha_d[(int64_t)k * H + h] = ad;
        }
      }
      const int cnt = min(32, end - c);
      for (int q = 0; q < cnt; ++q) {
        const float w = __shfl_sync(0xffffffffu, ad, q);
        const int j = __ldg(p.in_src + c + q);
        if (j == i) {
          axpy_row<NV>(w, own, acc);
        } else {
          float4 fj[NV];
          load_row<NV>(base + (int64_t)j * p.ldf, lane, D, fj);   // second touch: L1/L2 hit
          axpy_row<NV>(w, fj, acc);
        }
      }
    }
    // ---- epilogue ----
    float* orow = p.out + (int64_t)i * p.ldo + (p.hidden ? (int64_t)h * D : 0);
    const int64_t idx_base = (int64_t)i * p.ldo + (int64_t)h * D;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      const int c = (lane + 32 * t) * 4;
      const bool valid = c < D;
      float v[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
      bool keep[4] = {true, true, true, true};
      bool posv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) posv[u] = v[u] > 0.f;
      if (p.hidden) {
        if (p.act_slope != 1.f) {
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = posv[u] ? v[u] : v[u] * p.act_slope;
        }
        if (p.next_thr && valid) {
          const uint4 w = drop_words(p.next_seed, p.next_stream, (uint64_t)(idx_base + c) >> 2);
          keep[0] = w.x >= p.next_thr; keep[1] = w.y >= p.next_thr; keep[2] = w.z >= p.next_thr; keep[3] = w.w >= p.next_thr;
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = keep[u] ? v[u] * p.next_inv_keep : 0.f;
        }
      }
      if (valid) *reinterpret_cast<float4*>(orow + c) = make_float4(v[0], v[1], v[2], v[3]);
      if (p.maskbits) {
        uint32_t words[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          words[u] = __ballot_sync(0xffffffffu, valid && posv[u]);
          words[4 + u] = __ballot_sync(0xffffffffu, valid && keep[u]);
        }
        if (lane < 8) {
          uint32_t mine = words[0];
#pragma unroll
          for (int u = 1; u < 8; ++u) mine = lane == u ? words[u] : mine;
          p.maskbits[(((int64_t)i * H + h) * NV + t) * 8 + lane] = mine;
        }
      }
    }
    // position-embedding append + zero padding (once per row: the warp of the last head)
    if (p.hidden && h == H - 1) {
      const int feat = H * D;
      const int pd = p.pos_dim;
      float* row = p.out + (int64_t)i * p.ldo;
      const float* prow = pd > 0 ? p.next_pos_table + (int64_t)__ldg(p.pos + i) * pd : nullptr;
      for (int c = feat + lane; c < (int)p.ldo; c += 32) {
        float v = 0.f;
        if (c < feat + pd) {
          v = __ldg(prow + (c - feat));
          if (p.next_thr) v = drop_keep1(p.next_seed, p.next_stream, (uint64_t)((int64_t)i * p.ldo + c), p.next_thr) ? v * p.next_inv_keep : 0.f;
        }
        row[c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct FusedBwdParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const uint32_t* maskbits; int has_keep_plane; float act_slope; float next_inv_keep;
  const float* ft; int64_t ldf;
  const float* alpha; const float* alpha_d; const float* elog;
  const float* attn_l; const float* attn_r;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  const int32_t* out_ptr; const int32_t* out_dst; const int32_t* out_slot;
  const int32_t* node_off; int n_graphs;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* ds; float* da2;
  float* dft; int64_t ldd;
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
      const uint32_t* w = p.maskbits + (((int64_t)i * p.H + h) * NV + t) * 8;
      const uint4 sp = __ldg(reinterpret_cast<const uint4*>(w));
      uint4 kp = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (p.has_keep_plane) kp = __ldg(reinterpret_cast<const uint4*>(w + 4));
      const float on = p.next_inv_keep, neg = p.act_slope * p.next_inv_keep;
      x.x *= ((kp.x >> lane) & 1u) ? (((sp.x >> lane) & 1u) ? on : neg) : 0.f;
      x.y *= ((kp.y >> lane) & 1u) ? (((sp.y >> lane) & 1u) ? on : neg) : 0.f;
      x.z *= ((kp.z >> lane) & 1u) ? (((sp.z >> lane) & 1u) ? on : neg) : 0.f;
      x.w *= ((kp.w >> lane) & 1u) ? (((sp.w >> lane) & 1u) ? on : neg) : 0.f;
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

template <int NV>
__global__ void __launch_bounds__(256, 2) gat_fused_bwd_kernel(const FusedBwdParams p) {
  __shared__ float4 s_red[8][2][NV * 32];   // also holds attn_l / attn_r in slots [0][0], [0][1] during the main loop
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
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool attn_drop = p.attn_thr != 0;
  float4 accl[NV], accr[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) accl[t] = accr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_tiles = (p.n + kTileRows - 1) / kTileRows;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int gb = lower_bound_i32(p.node_off, p.n_graphs + 1, tile * kTileRows);
    const int ge = lower_bound_i32(p.node_off, p.n_graphs + 1, (tile + 1) * kTileRows);
    const int r0 = gb <= p.n_graphs ? __ldg(p.node_off + min(gb, p.n_graphs)) : p.n;
    const int r1 = __ldg(p.node_off + min(ge, p.n_graphs));
    // ---------------- phase A: per destination ----------------
    for (int i = r0 + wid; i < r1; i += 8) {
      const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
      float4 gi[NV];
      load_g_row<NV>(p, i, h, lane, gi);
      const float* fbase = p.ft + (int64_t)h * D;
      for (int k = beg; k < end; ++k) {
        float4 fj[NV];
        load_row<NV>(fbase + (int64_t)__ldg(p.in_src + k) * p.ldf, lane, D, fj);
        const float d = warp_sum(dot_rows<NV>(gi, fj)) * p.g_scale;
        if (lane == 0) p.ds[(int64_t)k * H + h] = d;
      }
      __syncwarp();
      float tsum = 0.f;
      for (int c = beg; c < end; c += 32) {
        const int k = c + lane;
        if (k < end) {
          float da = p.ds[(int64_t)k * H + h];
          if (attn_drop)
            da = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? da * p.attn_inv_keep : 0.f;
          tsum = fmaf(__ldg(p.alpha + (int64_t)k * H + h), da, tsum);
        }
      }
      tsum = warp_sum(tsum);
      float a2 = 0.f;
      for (int c = beg; c < end; c += 32) {
        const int k = c + lane;
        if (k < end) {
          float da = p.ds[(int64_t)k * H + h];
          if (attn_drop)
            da = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr) ? da * p.attn_inv_keep : 0.f;
          const float de = __ldg(p.alpha + (int64_t)k * H + h) * (da - tsum);
          const float dsv = __ldg(p.elog + (int64_t)k * H + h) > 0.f ? de : de * p.neg_slope;
          p.ds[(int64_t)k * H + h] = dsv;
          a2 += dsv;
        }
      }
      a2 = warp_sum(a2);
      if (lane == 0) p.da2[(int64_t)i * H + h] = a2;
      __syncwarp();
      // d(attn_l) += ds_k * ft[src_k] ; d(attn_r) += da2_i * ft[i]
      for (int k = beg; k < end; ++k) {
        float4 fj[NV];
        load_row<NV>(fbase + (int64_t)__ldg(p.in_src + k) * p.ldf, lane, D, fj);   // L1 hit
        axpy_row<NV>(p.ds[(int64_t)k * H + h], fj, accl);
      }
      {
        float4 fi[NV];
        load_row<NV>(fbase + (int64_t)i * p.ldf, lane, D, fi);
        axpy_row<NV>(a2, fi, accr);
      }
    }
    __syncthreads();   // ds / da2 of the whole tile are visible to the CTA
    // ---------------- phase B: per source ----------------
    for (int j = r0 + wid; j < r1; j += 8) {
      const int beg = __ldg(p.out_ptr + j), end = __ldg(p.out_ptr + j + 1);
      float d1 = 0.f;
      for (int c = beg; c < end; c += 32) {
        const int k = c + lane;
        if (k < end) d1 += p.ds[(int64_t)__ldg(p.out_slot + k) * H + h];
      }
      d1 = warp_sum(d1);
      const float d2 = p.da2[(int64_t)j * H + h];
      float4 acc[NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const float4 l = s_l[lane + 32 * t], r = s_r[lane + 32 * t];
        acc[t] = make_float4(fmaf(d1, l.x, d2 * r.x), fmaf(d1, l.y, d2 * r.y), fmaf(d1, l.z, d2 * r.z), fmaf(d1, l.w, d2 * r.w));
      }
      for (int k = beg; k < end; ++k) {
        const float w = __ldg(p.alpha_d + (int64_t)__ldg(p.out_slot + k) * H + h) * p.g_scale;
        float4 gv[NV];
        load_g_row<NV>(p, __ldg(p.out_dst + k), h, lane, gv);
        axpy_row<NV>(w, gv, acc);
      }
      float* orow = p.dft + (int64_t)j * p.ldd + (int64_t)h * D;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c = (lane + 32 * t) * 4;
        if (c < D) *reinterpret_cast<float4*>(orow + c) = acc[t];
      }
    }
    __syncthreads();   // the next tile's phase A overwrites nothing this tile still reads, but keep tiles in lock-step
  }
  // ---- d(attn) partials: fixed-order reduction over the 8 warps of the CTA ----
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    s_red[wid][0][lane + 32 * t] = accl[t];
    s_red[wid][1][lane + 32 * t] = accr[t];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * NV * 32; t += blockDim.x) {
    const int lr = t / (NV * 32), q = t % (NV * 32);
    float4 s = s_red[0][lr][q];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 x = s_red[w][lr][q];
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
__global__ void __launch_bounds__(128) pos_grad_partials_kernel(const float* __restrict__ dz, int64_t ldz, int col0,
                                                                const int32_t* __restrict__ pos, int n, int pd, int vocab,
                                                                float inv_keep, uint32_t thr, uint64_t seed, uint32_t stream_id,
                                                                float* __restrict__ partial) {
  const int r0 = blockIdx.x * kRowsPerBlock;
  const int r1 = min(n, r0 + kRowsPerBlock);
  for (int c = threadIdx.x; c < pd; c += blockDim.x) {
    float acc[kMaxVocab];
#pragma unroll
    for (int v = 0; v < kMaxVocab; ++v) acc[v] = 0.f;
    for (int i = r0; i < r1; ++i) {
      float g = __ldg(dz + (int64_t)i * ldz + col0 + c);
      if (thr) g = drop_keep1(seed, stream_id, (uint64_t)((int64_t)i * ldz + col0 + c), thr) ? g * inv_keep : 0.f;
      const int r = __ldg(pos + i);
#pragma unroll
      for (int v = 0; v < kMaxVocab; ++v) acc[v] += (r == v) ? g : 0.f;
    }
    for (int v = 0; v < vocab; ++v) partial[((int64_t)blockIdx.x * vocab + v) * pd + c] = acc[v];
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

int64_t tx_gat_fused_mask_words(int64_t n_nodes, int64_t heads, int64_t dim) {
  const int64_t nv = (dim + 127) / 128;
  return n_nodes * heads * nv * 8;
}

int64_t tx_gat_fused_bwd_blocks(int64_t n_nodes, int64_t heads) {
  const int64_t tiles = (n_nodes + kTileRows - 1) / kTileRows;
  int64_t gx = (2 * (int64_t)kNumSms + heads - 1) / heads;
  if (gx > tiles) gx = tiles;
  return gx < 1 ? 1 : gx;
}

int tx_gat_fused_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                     const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                     float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                     float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                     void* stream) {
  TX_REQUIRE(epi, "gat_fused_fwd: epilogue required");
  TX_REQUIRE(tx_gat_fused_supported(heads, dim, epi->mean_heads), "gat_fused_fwd: unsupported shape (heads %lld dim %lld); use the general path",
             (long long)heads, (long long)dim);
  TX_REQUIRE(aligned16(ft) && ldf % 4 == 0 && aligned16(out) && ldo % 4 == 0 && aligned16(attn_l) && aligned16(attn_r),
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
  p.hidden = epi->mean_heads ? 0 : 1; p.act_slope = epi->act_slope; p.next_pos_table = epi->next_pos_table; p.pos = epi->pos;
  p.pos_dim = (int)epi->pos_dim; p.next_inv_keep = 1.f / (1.f - epi->p_drop); p.next_thr = drop_threshold(epi->p_drop);
  p.next_seed = epi->seed; p.next_stream = epi->stream_id;
  const int nv = (int)((dim + 127) / 128);
  int gx = grid_for_warps(n_nodes, 8, 8);
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
                     float* dattn_partial, void* stream) {
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
  p.act_slope = act_slope; p.next_inv_keep = 1.f / (1.f - p_next); p.ft = ft; p.ldf = ldf; p.alpha = alpha;
  p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.attn_l = attn_l; p.attn_r = attn_r;
  p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid; p.out_ptr = out_ptr; p.out_dst = out_dst; p.out_slot = out_slot;
  p.node_off = node_off; p.n_graphs = (int)n_graphs; p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim;
  p.neg_slope = neg_slope; p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed;
  p.attn_stream = attn_stream_id; p.ds = ds; p.da2 = da2; p.dft = dft; p.ldd = ldd; p.dattn_partial = dattn_partial;
  const int nv = (int)((dim + 127) / 128);
  dim3 grid((unsigned)tx_gat_fused_bwd_blocks(n_nodes, heads), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nv) {
    case 1: gat_fused_bwd_kernel<1><<<grid, 256, 0, st>>>(p); break;
    case 2: gat_fused_bwd_kernel<2><<<grid, 256, 0, st>>>(p); break;
    case 3: gat_fused_bwd_kernel<3><<<grid, 256, 0, st>>>(p); break;
    default: gat_fused_bwd_kernel<4><<<grid, 256, 0, st>>>(p); break;
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
  pos_grad_partials_kernel<<<(int)row_blocks(n_nodes), 128, 0, (cudaStream_t)stream>>>(dz, ldz, (int)col0, pos, (int)n_nodes, (int)pos_dim,
                                                                                      (int)vocab, 1.f / (1.f - p_drop), drop_threshold(p_drop),
                                                                                      seed, stream_id, partial);
  TX_LAUNCH_CHECK("tx_pos_grad_partials");
  return TX_OK;
}

}  // extern "C"
