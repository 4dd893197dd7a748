// General-CSR kernels of the TaxoExpan propagation + readout path (sm_100a).
//
// One warp owns one destination (forward / destination phase of backward) or one source (source phase)
// node, lanes stride over 128-bit column vectors of the feature row, neighbour reductions are short
// serial loops (egonet in-degrees are 1, 2 or n_gp+1) and per-edge scalar reductions use warp shuffles.
// These kernels are correct for ANY batched graph (the GAT/GCN classes accept arbitrary graphs,
// reference model/model_zoo.py:116-137,169-190); the egonet-resident fast path lives in tx_egonet.cu.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#define TX_PDL_GROUP 0
#include "tx_common.cuh"

namespace tx {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

thread_local int g_preclear = 0;
bool preclear_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TAXO_PRECLEAR"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}
static int g_pdl = -1, g_pdl_mask = 0xFFFF;
static const char *g_pdl_only = nullptr, *g_pdl_skip = nullptr;
static void pdl_init() {
  if (g_pdl >= 0) return;
  const char* e = getenv("TAXO_PDL"); g_pdl = (e && atoi(e) == 0) ? 0 : 1;
  const char* m = getenv("TAXO_PDL_MASK"); g_pdl_mask = m ? atoi(m) : 0xFFFF;
  g_pdl_only = getenv("TAXO_PDL_ONLY"); g_pdl_skip = getenv("TAXO_PDL_SKIP");
}
bool pdl_enabled(int group, const char* name) {
  // TAXO_PDL=0 / tx_pdl_set(0): off; debugging: TAXO_PDL_MASK=<bits> only the source files whose bit is set, TAXO_PDL_ONLY=<substring> only
  // kernels whose name contains it, TAXO_PDL_SKIP=<substring> all but those
  pdl_init();
  if (g_pdl == 0 || !((g_pdl_mask >> group) & 1)) return false;
  if (name && g_pdl_only && !strstr(name, g_pdl_only)) return false;
  if (name && g_pdl_skip && strstr(name, g_pdl_skip)) return false;
  return true;
}
int pdl_set(int enabled) {
  pdl_init();
  const int prev = g_pdl;
  g_pdl = enabled ? 1 : 0;
  return prev;
}

struct Epilogue {  // device copy of tx_gat_epilogue
  int mean_heads;
  float act_slope;
  const float* next_pos_table;
  const int32_t* pos;
  int pos_dim;
  float inv_keep;   // 1/(1-p)
  uint32_t thr;     // keep iff word >= thr
  uint64_t seed;
  uint32_t stream_id;
};

static int make_epilogue(const tx_gat_epilogue* e, Epilogue* d) {
  d->mean_heads = e ? e->mean_heads : 0;
  d->act_slope = e ? e->act_slope : 1.f;
  d->next_pos_table = e ? e->next_pos_table : nullptr;
  d->pos = e ? e->pos : nullptr;
  d->pos_dim = e ? (int)e->pos_dim : 0;
  float p = e ? e->p_drop : 0.f;
  TX_REQUIRE(p >= 0.f && p < 1.f, "epilogue: p_drop must be in [0,1), got %f", p);
  d->inv_keep = 1.f / (1.f - p);
  d->thr = drop_threshold(p);
  d->seed = e ? e->seed : 0;
  d->stream_id = e ? e->stream_id : 0;
  if (d->pos_dim > 0) TX_REQUIRE(d->next_pos_table && d->pos, "epilogue: pos_dim > 0 needs next_pos_table and pos");
  return TX_OK;
}

// Writes VEC activated/dropped values of row i starting at column col of the output row.
template <int VEC>
__device__ __forceinline__ void epilogue_store(const Epilogue& ep, Vec<VEC> v, float* out_row, int64_t row_index_base,
                                               int col, bool activate) {
  if (activate) {
#pragma unroll
    for (int t = 0; t < VEC; ++t) v.v[t] = v.v[t] > 0.f ? v.v[t] : v.v[t] * ep.act_slope;
  }
  if (ep.thr) {
    bool keep[VEC];
    drop_keep_vec<VEC>(ep.seed, ep.stream_id, (uint64_t)(row_index_base + col), ep.thr, keep);
#pragma unroll
    for (int t = 0; t < VEC; ++t) v.v[t] = keep[t] ? v.v[t] * ep.inv_keep : 0.f;
  }
  v.store(out_row + col);
}

// Appends drop(P_next[pos_i]) after the feature columns and zero-fills up to ldo (one warp, all lanes).
template <int VEC>
__device__ __forceinline__ void epilogue_tail(const Epilogue& ep, float* out_row, int64_t row_index_base, int i,
                                              int feat_cols, int ldo, int lane) {
  const int pd = ep.pos_dim;
  const float* prow = pd > 0 ? ep.next_pos_table + (int64_t)ep.pos[i] * pd : nullptr;
  for (int c = feat_cols + lane; c < ldo; c += 32) {
    float v = 0.f;
    if (c < feat_cols + pd) {
      v = __ldg(prow + (c - feat_cols));
      if (ep.thr) v = drop_keep1(ep.seed, ep.stream_id, (uint64_t)(row_index_base + c), ep.thr) ? v * ep.inv_keep : 0.f;
    }
    out_row[c] = v;
  }
}

// =============================================================================================
// concat + dropout (layer-0 input):  z = drop([x || P[pos]])
// =============================================================================================
__global__ void concat_pos_dropout_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ ptab,
                                              const int32_t* __restrict__ pos, int n, int k_in, int pd, float* __restrict__ z,
                                              int ldz, float inv_keep, uint32_t thr, uint64_t seed, uint32_t stream_id) {
  TX_PDL_ENTER();
  const int vec_per_row = ldz >> 2;
  const int64_t total = (int64_t)n * vec_per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / vec_per_row);
    const int c0 = (int)(t - (int64_t)i * vec_per_row) << 2;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      float val = 0.f;
      if (c < k_in) val = __ldg(x + (int64_t)i * ldx + c);
      else if (c < k_in + pd) val = __ldg(ptab + (int64_t)__ldg(pos + i) * pd + (c - k_in));
      v[u] = val;
    }
    if (thr) {
      bool keep[4];
      drop_keep4(seed, stream_id, (uint64_t)(((int64_t)i * ldz + c0) >> 2), thr, keep);
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = keep[u] ? v[u] * inv_keep : 0.f;
    }
    *reinterpret_cast<float4*>(z + (int64_t)i * ldz + c0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// The same, written straight as the fp16 hi / lo operand pair of the first layer's projection GEMM (z itself is never stored): the
// scale comes from the bound max(max|x|, max|P|) / (1 - p) - every CTA derives it from the device scalar x_amax and the (tiny) table -
// and the keep decisions use the SAME counters as the fp32 kernel (row pitch ldz = round4(k_in + pd)), so tx_pos_grad_partials /
// tx_epilogue_bwd rebuild the identical mask in the backward pass.
__global__ void __launch_bounds__(256) concat_pos_dropout_f16_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ ptab,
                                                                     const int32_t* __restrict__ pos, int n, int k_in, int pd, int vocab,
                                                                     const float* __restrict__ x_amax, __half* __restrict__ hi,
                                                                     __half* __restrict__ lo, int ld16, int ldz, float inv_keep,
                                                                     uint32_t thr, uint64_t seed, uint32_t stream_id,
                                                                     float* __restrict__ scale_out) {
  TX_PDL_ENTER();
  __shared__ float s_m[8];
  float m = 0.f;
  for (int t = threadIdx.x; t < vocab * pd; t += blockDim.x) m = fmaxf(m, fabsf(__ldg(ptab + t)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  m = __ldg(x_amax);
#pragma unroll
  for (int w = 0; w < 8; ++w) m = fmaxf(m, s_m[w]);
  const float scale = f16_split_scale(m * inv_keep);
  if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) *scale_out = scale;
  const int vec_per_row = ld16 >> 2;
  const int64_t total = (int64_t)n * vec_per_row;
  const bool x_vec2 = (ldx & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / vec_per_row);
    const int c0 = (int)(t - (int64_t)i * vec_per_row) << 2;
    float v[4];
    if (x_vec2 && c0 + 3 < k_in) {                     // even row pitch: two 64-bit loads (a [N, 250] feature block is 8-byte aligned per row)
      const float2 a = __ldg(reinterpret_cast<const float2*>(x + (int64_t)i * ldx + c0));
      const float2 b = __ldg(reinterpret_cast<const float2*>(x + (int64_t)i * ldx + c0 + 2));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u;
        float val = 0.f;
        if (c < k_in) val = __ldg(x + (int64_t)i * ldx + c);
        else if (c < k_in + pd) val = __ldg(ptab + (int64_t)__ldg(pos + i) * pd + (c - k_in));
        v[u] = val;
      }
    }
    if (thr && c0 < ldz) {
      bool keep[4];
      drop_keep4(seed, stream_id, (uint64_t)(((int64_t)i * ldz + c0) >> 2), thr, keep);
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = keep[u] ? v[u] * inv_keep : 0.f;
    }
    uint2 h, l;
    f16_split4(make_float4(v[0], v[1], v[2], v[3]), scale, h, l);
    *reinterpret_cast<uint2*>(hi + (int64_t)i * ld16 + c0) = h;
    *reinterpret_cast<uint2*>(lo + (int64_t)i * ld16 + c0) = l;
  }
}

// =============================================================================================
// epilogue backward (in place) + position-table gradient partials
// =============================================================================================
__global__ void __launch_bounds__(256) epilogue_bwd_kernel(float* __restrict__ dz, int ldz, const float* __restrict__ z,
                                                           const int32_t* __restrict__ pos, int n, int k_in, int pd,
                                                           int vocab, float slope, float inv_keep, uint32_t thr,
                                                           uint64_t seed, uint32_t stream_id,
                                                           float* __restrict__ dpos_partial) {
  const int r0 = blockIdx.x * kRowsPerBlock;
  const int r1 = min(n, r0 + kRowsPerBlock);
  const bool act = (z != nullptr) && (slope != 1.f);
  // part 1: feature columns [0, k_in), vectors of 4 (ldz % 4 == 0; a vector may straddle k_in)
  const int fvec = (k_in + 3) >> 2;
  if (thr || act) {
    for (int t = threadIdx.x; t < (r1 - r0) * fvec; t += blockDim.x) {
      const int i = r0 + t / fvec;
      const int c0 = (t % fvec) << 2;
      float4 g = *reinterpret_cast<float4*>(dz + (int64_t)i * ldz + c0);
      float gv[4] = {g.x, g.y, g.z, g.w};
      float zv[4] = {1.f, 1.f, 1.f, 1.f};
      if (act) {
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + (int64_t)i * ldz + c0));
        zv[0] = zz.x; zv[1] = zz.y; zv[2] = zz.z; zv[3] = zz.w;
      }
      bool keep[4] = {true, true, true, true};
      if (thr) drop_keep4(seed, stream_id, (uint64_t)(((int64_t)i * ldz + c0) >> 2), thr, keep);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c0 + u < k_in) {
          float f = keep[u] ? inv_keep : 0.f;
          if (act && !(zv[u] > 0.f)) f *= slope;
          gv[u] *= f;
        }
      }
      *reinterpret_cast<float4*>(dz + (int64_t)i * ldz + c0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
  }
  // part 2: dP partials, one thread per position column
  if (pd > 0 && dpos_partial) {
    for (int c = threadIdx.x; c < pd; c += blockDim.x) {
      float acc[kMaxVocab];
#pragma unroll
      for (int v = 0; v < kMaxVocab; ++v) acc[v] = 0.f;
      for (int i = r0; i < r1; ++i) {
        float g = dz[(int64_t)i * ldz + k_in + c];
        if (thr) g = drop_keep1(seed, stream_id, (uint64_t)((int64_t)i * ldz + k_in + c), thr) ? g * inv_keep : 0.f;
        const int r = __ldg(pos + i);
#pragma unroll
        for (int v = 0; v < kMaxVocab; ++v) acc[v] += (r == v) ? g : 0.f;
      }
      for (int v = 0; v < vocab; ++v) dpos_partial[((int64_t)blockIdx.x * vocab + v) * pd + c] = acc[v];
    }
  }
}

// few partial blocks of many elements (split-K GEMM partials): one float4 column per thread, blocks summed in index order
__global__ void __launch_bounds__(256) reduce_partials_wide_kernel(const float* __restrict__ partial, int n_blocks, int64_t m4,
                                                                   float* __restrict__ out) {
  TX_PDL_ENTER();
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m4; t += (int64_t)gridDim.x * blockDim.x) {
    float4 s = __ldg(reinterpret_cast<const float4*>(partial) + t);
    for (int b = 1; b < n_blocks; ++b) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + (int64_t)b * m4 + t);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[t] = s;
  }
}

// the same for a [rows, cols] block of pitched partials written to a (differently) pitched output: out[r, c] = sum_b partial[b, r, c] in
// index order (bit-identical to reduce_partials_wide_kernel).  Lets a weight gradient land directly in its contiguous [rows, cols] home
// (a parameter's .grad / a flat gradient bucket) although the split-K partials have a 16-byte-aligned row pitch.
__global__ void __launch_bounds__(256) reduce_partials_rows_kernel(const float* __restrict__ partial, int n_blocks, int64_t block_stride,
                                                                   int rows, int cols, int64_t ld_in, float* __restrict__ out, int64_t ld_out) {
  TX_PDL_ENTER();
  const int64_t total = (int64_t)rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / cols), c = (int)(t - (int64_t)r * cols);
    const float* src = partial + (int64_t)r * ld_in + c;
    float s = __ldg(src);
    for (int b = 1; b < n_blocks; ++b) s += __ldg(src + (int64_t)b * block_stride);
    out[(int64_t)r * ld_out + c] = s;
  }
}

// partial[b, m] summed over b in a FIXED order: 32 columns per CTA, the blocks dealt round-robin to WY warps (4 independent loads
// in flight per thread), then a fixed-order tree over the warps.  WY = 32 for long block lists (few columns, hundreds of
// blocks: the dependent-load chain per thread was 70+ loads with 8 warps), 8 otherwise.
template <int WY>
__global__ void __launch_bounds__(32 * WY) reduce_partials_kernel(const float* __restrict__ partial, int64_t n_blocks, int64_t m_len,
                                                                  float* __restrict__ out) {
  TX_PDL_ENTER();
  __shared__ float sm[WY][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 32 + x;
  float acc = 0.f;
  if (m < m_len) {
    int64_t b = y;
    for (; b + 3 * WY < n_blocks; b += 4 * WY) {   // 4 independent loads in flight per thread
      const float v0 = partial[b * m_len + m], v1 = partial[(b + WY) * m_len + m];
      const float v2 = partial[(b + 2 * WY) * m_len + m], v3 = partial[(b + 3 * WY) * m_len + m];
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; b < n_blocks; b += WY) acc += partial[b * m_len + m];
  }
  sm[y][x] = acc;
  __syncthreads();
  if (y == 0 && m < m_len) {
    float t = sm[0][x];
#pragma unroll
    for (int r = 1; r < WY; ++r) t += sm[r][x];
    out[m] = t;
  }
}

__global__ void colsum_partials_kernel(const float* __restrict__ x, int64_t ldx, int n_rows, int n_cols,
                                       float* __restrict__ partial) {
  TX_PDL_ENTER();
  const int r0 = blockIdx.x * kRowsPerBlock;
  const int r1 = min(n_rows, r0 + kRowsPerBlock);
  for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
    float acc = 0.f;
    for (int i = r0; i < r1; ++i) acc += __ldg(x + (int64_t)i * ldx + c);
    partial[(int64_t)blockIdx.x * n_cols + c] = acc;
  }
}

// =============================================================================================
// GAT
// =============================================================================================
template <int VEC>
__global__ void __launch_bounds__(256) gat_node_logits_kernel(const float* __restrict__ ft, int64_t ldf,
                                                              const float* __restrict__ attn_l,
                                                              const float* __restrict__ attn_r, int n, int H, int D,
                                                              float* __restrict__ a1, float* __restrict__ a2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nvec = D / VEC;
  for (int64_t w = warp; w < (int64_t)n * H; w += nwarps) {
    const int i = (int)(w / H), h = (int)(w % H);
    const float* row = ft + (int64_t)i * ldf + (int64_t)h * D;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < nvec; c += 32) {
      const Vec<VEC> f = Vec<VEC>::load(row + c * VEC);
      const Vec<VEC> l = Vec<VEC>::load(attn_l + (int64_t)h * D + c * VEC);
      const Vec<VEC> r = Vec<VEC>::load(attn_r + (int64_t)h * D + c * VEC);
#pragma unroll
      for (int t = 0; t < VEC; ++t) {
        s1 = fmaf(f.v[t], l.v[t], s1);
        s2 = fmaf(f.v[t], r.v[t], s2);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      a1[w] = s1;
      a2[w] = s2;
    }
  }
}

struct GatFwdParams {
  const float* ft; int64_t ldf;
  const float* a1; const float* a2;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* alpha; float* alpha_d; float* elog;
  float* out; int64_t ldo;
  Epilogue ep;
};

template <int VEC>
__global__ void __launch_bounds__(256) gat_aggregate_fwd_kernel(const GatFwdParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int H = p.H, D = p.D;
  const int nvec = D / VEC;
  const bool has_drop = p.attn_thr != 0;
  for (int i = warp; i < p.n; i += nwarps) {
    const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
    // ---- stage 1: edge logits, edge softmax, attention dropout (lanes over in-edges) ----
    for (int h = 0; h < H; ++h) {
      const float a2i = __ldg(p.a2 + (int64_t)i * H + h);
      float m = -INFINITY;
      for (int k = beg + lane; k < end; k += 32) {
        float s = __ldg(p.a1 + (int64_t)__ldg(p.in_src + k) * H + h) + a2i;
        s = s > 0.f ? s : s * p.neg_slope;
        p.elog[(int64_t)k * H + h] = s;
        m = fmaxf(m, s);
      }
      m = warp_max(m);
      float l = 0.f;
      for (int k = beg + lane; k < end; k += 32) {
        const float e = expf(p.elog[(int64_t)k * H + h] - m);
        p.alpha[(int64_t)k * H + h] = e;
        l += e;
      }
      l = warp_sum(l);
      for (int k = beg + lane; k < end; k += 32) {
        const float a = p.alpha[(int64_t)k * H + h] / l;
        p.alpha[(int64_t)k * H + h] = a;
        if (has_drop) {
          const bool keep = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr);
          p.alpha_d[(int64_t)k * H + h] = keep ? a * p.attn_inv_keep : 0.f;
        }
      }
    }
    __syncwarp();
    // ---- stage 2: weighted aggregation (lanes over column vectors) + epilogue ----
    const float* wts = has_drop ? p.alpha_d : p.alpha;
    float* orow = p.out + (int64_t)i * p.ldo;
    const int64_t idx_base = (int64_t)i * p.ldo;
    if (!p.ep.mean_heads) {
      for (int h = 0; h < H; ++h) {
        for (int c = lane; c < nvec; c += 32) {
          Vec<VEC> acc = vzero<VEC>();
          const float* col = p.ft + (int64_t)h * D + c * VEC;
          int k = beg;
          for (; k + 4 <= end; k += 4) {
            float w[4];
            Vec<VEC> f[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w[u] = wts[(int64_t)(k + u) * H + h];
              f[u] = Vec<VEC>::load(col + (int64_t)__ldg(p.in_src + k + u) * p.ldf);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(w[u], f[u].v[t], acc.v[t]);
          }
          for (; k < end; ++k) {
            const float w = wts[(int64_t)k * H + h];
            const Vec<VEC> f = Vec<VEC>::load(col + (int64_t)__ldg(p.in_src + k) * p.ldf);
#pragma unroll
            for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(w, f.v[t], acc.v[t]);
          }
          epilogue_store<VEC>(p.ep, acc, orow, idx_base, h * D + c * VEC, p.ep.act_slope != 1.f);
        }
      }
      epilogue_tail<VEC>(p.ep, orow, idx_base, i, H * D, (int)p.ldo, lane);
    } else {
      const float inv_h = 1.f / (float)H;
      for (int c = lane; c < nvec; c += 32) {
        Vec<VEC> tot = vzero<VEC>();
        for (int h = 0; h < H; ++h) {
          Vec<VEC> acc = vzero<VEC>();
          const float* col = p.ft + (int64_t)h * D + c * VEC;
          for (int k = beg; k < end; ++k) {
            const float w = wts[(int64_t)k * H + h];
            const Vec<VEC> f = Vec<VEC>::load(col + (int64_t)__ldg(p.in_src + k) * p.ldf);
#pragma unroll
            for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(w, f.v[t], acc.v[t]);
          }
#pragma unroll
          for (int t = 0; t < VEC; ++t) tot.v[t] += acc.v[t];
        }
        if (H > 1) {
#pragma unroll
          for (int t = 0; t < VEC; ++t) tot.v[t] *= inv_h;
        }
        tot.store(orow + c * VEC);
      }
    }
  }
}

struct GatBwdDstParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const float* ft; int64_t ldf;
  const float* alpha; const float* elog;
  const int32_t* in_ptr; const int32_t* in_src; const int32_t* in_eid;
  int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* ds; float* da2;
};

template <int VEC>
__global__ void __launch_bounds__(256) gat_aggregate_bwd_dst_kernel(const GatBwdDstParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int H = p.H, D = p.D;
  const int nvec = D / VEC;
  for (int i = warp; i < p.n; i += nwarps) {
    const int beg = __ldg(p.in_ptr + i), end = __ldg(p.in_ptr + i + 1);
    for (int h = 0; h < H; ++h) {
      const float* grow = p.g + (int64_t)i * p.ldg + (int64_t)h * p.g_head_stride;
      // d(alpha_d)[k] = <g_i, ft_src(k)>
      for (int k = beg; k < end; ++k) {
        const float* frow = p.ft + (int64_t)__ldg(p.in_src + k) * p.ldf + (int64_t)h * D;
        float s = 0.f;
        for (int c = lane; c < nvec; c += 32) {
          const Vec<VEC> a = Vec<VEC>::load(grow + c * VEC);
          const Vec<VEC> b = Vec<VEC>::load(frow + c * VEC);
#pragma unroll
          for (int t = 0; t < VEC; ++t) s = fmaf(a.v[t], b.v[t], s);
        }
        s = warp_sum(s);
        if (lane == 0) p.ds[(int64_t)k * H + h] = s * p.g_scale;
      }
      __syncwarp();
      // softmax backward over the in-edges, then leaky-relu backward
      float tsum = 0.f;
      for (int k = beg + lane; k < end; k += 32) {
        float da = p.ds[(int64_t)k * H + h];
        if (p.attn_thr) {
          const bool keep = drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)__ldg(p.in_eid + k) * H + h), p.attn_thr);
          da = keep ? da * p.attn_inv_keep : 0.f;
          p.ds[(int64_t)k * H + h] = da;
        }
        tsum = fmaf(__ldg(p.alpha + (int64_t)k * H + h), da, tsum);
      }
      tsum = warp_sum(tsum);
      float a2sum = 0.f;
      for (int k = beg + lane; k < end; k += 32) {
        const float a = __ldg(p.alpha + (int64_t)k * H + h);
        const float de = a * (p.ds[(int64_t)k * H + h] - tsum);
        const float dsv = __ldg(p.elog + (int64_t)k * H + h) > 0.f ? de : de * p.neg_slope;
        p.ds[(int64_t)k * H + h] = dsv;
        a2sum += dsv;
      }
      a2sum = warp_sum(a2sum);
      if (lane == 0) p.da2[(int64_t)i * H + h] = a2sum;
      __syncwarp();
    }
  }
}

struct GatBwdSrcParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const float* alpha_d; const float* ds; const float* da2;
  const float* attn_l; const float* attn_r;
  const int32_t* out_ptr; const int32_t* out_dst; const int32_t* out_slot;
  int n; int H; int D;
  float* da1; float* dft; int64_t ldd;
};

template <int VEC>
__global__ void __launch_bounds__(256) gat_aggregate_bwd_src_kernel(const GatBwdSrcParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int H = p.H, D = p.D;
  const int nvec = D / VEC;
  for (int j = warp; j < p.n; j += nwarps) {
    const int beg = __ldg(p.out_ptr + j), end = __ldg(p.out_ptr + j + 1);
    for (int h = 0; h < H; ++h) {
      float d1 = 0.f;
      for (int k = beg + lane; k < end; k += 32) d1 += __ldg(p.ds + (int64_t)__ldg(p.out_slot + k) * H + h);
      d1 = warp_sum(d1);
      if (lane == 0) p.da1[(int64_t)j * H + h] = d1;
      const float d2 = __ldg(p.da2 + (int64_t)j * H + h);
      for (int c = lane; c < nvec; c += 32) {
        const Vec<VEC> l = Vec<VEC>::load(p.attn_l + (int64_t)h * D + c * VEC);
        const Vec<VEC> r = Vec<VEC>::load(p.attn_r + (int64_t)h * D + c * VEC);
        Vec<VEC> acc;
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(d1, l.v[t], d2 * r.v[t]);
        const float* gcol = p.g + (int64_t)h * p.g_head_stride + c * VEC;
        for (int k = beg; k < end; ++k) {
          const float w = __ldg(p.alpha_d + (int64_t)__ldg(p.out_slot + k) * H + h) * p.g_scale;
          const Vec<VEC> gv = Vec<VEC>::load(gcol + (int64_t)__ldg(p.out_dst + k) * p.ldg);
#pragma unroll
          for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(w, gv.v[t], acc.v[t]);
        }
        acc.store(p.dft + (int64_t)j * p.ldd + (int64_t)h * D + c * VEC);
      }
    }
  }
}

__global__ void __launch_bounds__(256) gat_attn_grad_partials_kernel(const float* __restrict__ ft, int64_t ldf,
                                                                     const float* __restrict__ da1,
                                                                     const float* __restrict__ da2, int n, int H, int D,
                                                                     float* __restrict__ partial) {
  const int r0 = blockIdx.x * kRowsPerBlock;
  const int r1 = min(n, r0 + kRowsPerBlock);
  const int F = H * D;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    const int h = c / D;
    float al = 0.f, ar = 0.f;
    for (int j = r0; j < r1; ++j) {
      const float f = __ldg(ft + (int64_t)j * ldf + c);
      al = fmaf(__ldg(da1 + (int64_t)j * H + h), f, al);
      ar = fmaf(__ldg(da2 + (int64_t)j * H + h), f, ar);
    }
    partial[((int64_t)blockIdx.x * 2 + 0) * F + c] = al;
    partial[((int64_t)blockIdx.x * 2 + 1) * F + c] = ar;
  }
}

// =============================================================================================
// GCN
// =============================================================================================
__global__ void gcn_norm_kernel(const int32_t* __restrict__ in_ptr, int n, float* __restrict__ norm) {
  TX_PDL_ENTER();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int deg = in_ptr[i + 1] - in_ptr[i];
  // torch.pow(degs, -0.5) with inf -> 0 (reference model_zoo.py:158-159)
  norm[i] = deg > 0 ? 1.0f / sqrtf((float)deg) : 0.f;
}

template <int VEC>
__global__ void __launch_bounds__(256) gcn_aggregate_fwd_kernel(const float* __restrict__ y, int64_t ldy,
                                                                const float* __restrict__ norm,
                                                                const float* __restrict__ bias,
                                                                const int32_t* __restrict__ in_ptr,
                                                                const int32_t* __restrict__ in_src, int n, int D,
                                                                float* __restrict__ out, int64_t ldo, const Epilogue ep) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D / VEC;
  for (int i = warp; i < n; i += nwarps) {
    const int beg = __ldg(in_ptr + i), end = __ldg(in_ptr + i + 1);
    const float ni = __ldg(norm + i);
    float* orow = out + (int64_t)i * ldo;
    const int64_t idx_base = (int64_t)i * ldo;
    for (int c = lane; c < nvec; c += 32) {
      Vec<VEC> acc = vzero<VEC>();
      for (int k = beg; k < end; ++k) {
        const int j = __ldg(in_src + k);
        const float nj = __ldg(norm + j);
        const Vec<VEC> f = Vec<VEC>::load(y + (int64_t)j * ldy + c * VEC);
        // (y_j * norm_j) summed, as the reference multiplies before update_all (model_zoo.py:39-41)
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc.v[t] += f.v[t] * nj;
      }
      Vec<VEC> v;
      if (bias) {
        const Vec<VEC> b = Vec<VEC>::load(bias + c * VEC);
#pragma unroll
        for (int t = 0; t < VEC; ++t) v.v[t] = acc.v[t] * ni + b.v[t];
      } else {
#pragma unroll
        for (int t = 0; t < VEC; ++t) v.v[t] = acc.v[t] * ni;
      }
      if (ep.mean_heads) v.store(orow + c * VEC);
      else epilogue_store<VEC>(ep, v, orow, idx_base, c * VEC, ep.act_slope != 1.f);
    }
    if (!ep.mean_heads) epilogue_tail<VEC>(ep, orow, idx_base, i, D, (int)ldo, lane);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) gcn_aggregate_bwd_kernel(const float* __restrict__ g, int64_t ldg,
                                                                const float* __restrict__ norm,
                                                                const int32_t* __restrict__ out_ptr,
                                                                const int32_t* __restrict__ out_dst, int n, int D,
                                                                float* __restrict__ dy, int64_t ldd) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D / VEC;
  for (int j = warp; j < n; j += nwarps) {
    const int beg = __ldg(out_ptr + j), end = __ldg(out_ptr + j + 1);
    const float nj = __ldg(norm + j);
    for (int c = lane; c < nvec; c += 32) {
      Vec<VEC> acc = vzero<VEC>();
      for (int k = beg; k < end; ++k) {
        const int i = __ldg(out_dst + k);
        const float ni = __ldg(norm + i);
        const Vec<VEC> gv = Vec<VEC>::load(g + (int64_t)i * ldg + c * VEC);
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(ni, gv.v[t], acc.v[t]);
      }
#pragma unroll
      for (int t = 0; t < VEC; ++t) acc.v[t] *= nj;
      acc.store(dy + (int64_t)j * ldd + c * VEC);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GCN aggregate with the fused next-layer epilogue PGAT has (round 2): the hidden layer's output goes out as the NEXT GEMM's fp16 hi/lo
// operand pair (the fp32 tensor is never written) together with the sign / keep bytes (1 byte per 4 columns: 4 sign bits | 4 keep
// bits << 4, the layout of tx_gat_fused_mask_ld(1, D)) that let the next layer's d(z) GEMM epilogue apply the derivative of
// "leaky-relu -> dropout" where the gradient is produced - no tx_epilogue_bwd pass, no clone of the gradient - and the backward
// aggregate writes d(y) as an fp16 pair too.  Same arithmetic, dropout counters (row * ldo + col with the LOGICAL fp32 pitch ldo) and
// reference call sites as the fp32 kernels above (model_zoo.py:39-49 followed by :164-165,37 of the next layer).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gcn_aggregate_fwd_f16_kernel(const float* __restrict__ y, int64_t ldy, const float* __restrict__ norm,
                                                                    const float* __restrict__ bias, const int32_t* __restrict__ in_ptr,
                                                                    const int32_t* __restrict__ in_src, int n, int D, int64_t ldo,
                                                                    const Epilogue ep, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                                    int64_t ld16, const float* __restrict__ bound, float* __restrict__ scale_out,
                                                                    uint8_t* __restrict__ mask, int mask_ld) {
  TX_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D >> 2;
  const float scale16 = f16_split_scale(__ldg(bound));
  if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) *scale_out = scale16;
  const bool activate = ep.act_slope != 1.f;
  for (int i = warp; i < n; i += nwarps) {
    const int beg = __ldg(in_ptr + i), end = __ldg(in_ptr + i + 1);
    const float ni = __ldg(norm + i);
    const int64_t idx_base = (int64_t)i * ldo;
    for (int c = lane; c < nvec; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = beg; k < end; ++k) {
        const int j = __ldg(in_src + k);
        const float nj = __ldg(norm + j);
        const float4 f = __ldg(reinterpret_cast<const float4*>(y + (int64_t)j * ldy + c * 4));
        acc.x += f.x * nj; acc.y += f.y * nj; acc.z += f.z * nj; acc.w += f.w * nj;     // (y_j * norm_j) summed, model_zoo.py:39-41
      }
      float v[4] = {acc.x * ni, acc.y * ni, acc.z * ni, acc.w * ni};
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c * 4));
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      uint32_t code = 0xF0u;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool pos = v[u] > 0.f;
        code |= pos ? (1u << u) : 0u;
        if (activate) v[u] = pos ? v[u] : v[u] * ep.act_slope;
      }
      if (ep.thr) {
        bool keep[4];
        drop_keep4(ep.seed, ep.stream_id, (uint64_t)((idx_base + c * 4) >> 2), ep.thr, keep);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = keep[u] ? v[u] * ep.inv_keep : 0.f;
          code &= keep[u] ? 0xFFu : ~(16u << u);
        }
      }
      uint2 h16, l16;
      f16_split4(make_float4(v[0], v[1], v[2], v[3]), scale16, h16, l16);
      *reinterpret_cast<uint2*>(out_hi + (int64_t)i * ld16 + c * 4) = h16;
      *reinterpret_cast<uint2*>(out_lo + (int64_t)i * ld16 + c * 4) = l16;
      if (mask) mask[(int64_t)i * mask_ld + c] = (uint8_t)code;
    }
    // appended drop(P_next[pos_i]) and the zero padding up to ld16
    const int pd = ep.pos_dim;
    const float* prow = pd > 0 ? ep.next_pos_table + (int64_t)ep.pos[i] * pd : nullptr;
    for (int c = D + lane; c < (int)ld16; c += 32) {
      float v = 0.f;
      if (c < D + pd) {
        v = __ldg(prow + (c - D));
        if (ep.thr) v = drop_keep1(ep.seed, ep.stream_id, (uint64_t)(idx_base + c), ep.thr) ? v * ep.inv_keep : 0.f;
      }
      const float x = fminf(fmaxf(v * scale16, -65504.f), 65504.f);
      const __half hh = __float2half_rn(x);
      out_hi[(int64_t)i * ld16 + c] = hh;
      out_lo[(int64_t)i * ld16 + c] = __float2half_rn(x - __half2float(hh));
    }
  }
}

__global__ void __launch_bounds__(256) gcn_aggregate_bwd_f16_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ norm,
                                                                    const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst,
                                                                    int n, int D, __half* __restrict__ dy_hi, __half* __restrict__ dy_lo,
                                                                    int64_t ld16, const float* __restrict__ bound, float* __restrict__ scale_out) {
  TX_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec16 = (int)(ld16 >> 2);
  const float scale16 = f16_split_scale(__ldg(bound));
  if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) *scale_out = scale16;
  for (int j = warp; j < n; j += nwarps) {
    const int beg = __ldg(out_ptr + j), end = __ldg(out_ptr + j + 1);
    const float nj = __ldg(norm + j);
    for (int c = lane; c < nvec16; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c * 4 < D) {
        for (int k = beg; k < end; ++k) {
          const int i = __ldg(out_dst + k);
          const float ni = __ldg(norm + i);
          const float4 gv = __ldg(reinterpret_cast<const float4*>(g + (int64_t)i * ldg + c * 4));
          acc.x = fmaf(ni, gv.x, acc.x); acc.y = fmaf(ni, gv.y, acc.y); acc.z = fmaf(ni, gv.z, acc.z); acc.w = fmaf(ni, gv.w, acc.w);
        }
        acc.x *= nj; acc.y *= nj; acc.z *= nj; acc.w *= nj;
      }
      uint2 h16, l16;
      f16_split4(acc, scale16, h16, l16);
      *reinterpret_cast<uint2*>(dy_hi + (int64_t)j * ld16 + c * 4) = h16;       // columns [D, ld16) are written as zeros
      *reinterpret_cast<uint2*>(dy_lo + (int64_t)j * ld16 + c * 4) = l16;
    }
  }
}

// *out = max(2 ca *a, 2 cb max|b|, ct max|t|): bound of a GCN layer's epilogue output, |norm_i sum_j norm_j y_j + bias| <= sqrt(max
// in-degree) max|y| + max|bias| <= 2 max(.., ..), and of the appended position rows
__global__ void bound_gcn_kernel(const float* a, float ca, const float* b, int64_t nb, float cb, const float* t, int64_t nt, float ct, float* out) {
  TX_PDL_ENTER();
  __shared__ float s_red[8];
  auto block_max = [&](const float* v, int64_t len) -> float {
    float m = 0.f;
    for (int64_t k = threadIdx.x; k < len; k += blockDim.x) m = fmaxf(m, fabsf(v[k]));
    m = warp_max(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s_red[w]);
    return r;
  };
  const float mb = b ? block_max(b, nb) : 0.f, mt = t ? block_max(t, nt) : 0.f;
  if (threadIdx.x == 0) *out = fmaxf(fmaxf(2.f * ca * *a, 2.f * cb * mb), ct * mt);
}

// =============================================================================================
// Readout: one CTA per graph
// =============================================================================================
__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.functional.softplus (beta=1, threshold=20): x if x > 20 else log1p(exp(x))
  return x > 20.f ? x : log1pf(expf(x));
}

template <int VEC>
__global__ void __launch_bounds__(256) readout_fwd_kernel(int kind, const float* __restrict__ h, int64_t ldh,
                                                          const int32_t* __restrict__ pos, const float* __restrict__ pw,
                                                          const int32_t* __restrict__ node_off, int n_graphs, int D,
                                                          float* __restrict__ hg, int64_t ldhg) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D / VEC;
  float sw[3] = {1.f, 1.f, 1.f};
  if (kind == TX_READOUT_WMEAN) {
#pragma unroll
    for (int r = 0; r < 3; ++r) sw[r] = softplus_f(__ldg(pw + r));   // a_i = softplus(w[pos_i]), model_zoo.py:241
  }
  for (int g = warp; g < n_graphs; g += nwarps) {
    const int beg = __ldg(node_off + g), end = __ldg(node_off + g + 1);
    float* orow = hg + (int64_t)g * ldhg;
    if (kind == TX_READOUT_CONCAT) {
      float na = 0.f;
      for (int i = beg + lane; i < end; i += 32) na += (__ldg(pos + i) == 1) ? 1.f : 0.f;
      na = warp_sum(na);
      const float inv_n = 1.f / (float)(end - beg);
      for (int c = lane; c < nvec; c += 32) {
        Vec<VEC> s0 = vzero<VEC>(), s1 = vzero<VEC>(), s2 = vzero<VEC>();
        for (int i = beg; i < end; ++i) {
          const Vec<VEC> v = Vec<VEC>::load(h + (int64_t)i * ldh + c * VEC);
          const int r = __ldg(pos + i);
#pragma unroll
          for (int t = 0; t < VEC; ++t) {
            s0.v[t] += r == 0 ? v.v[t] : 0.f;
            s1.v[t] += r == 1 ? v.v[t] : 0.f;
            s2.v[t] += r == 2 ? v.v[t] : 0.f;
          }
        }
#pragma unroll
        for (int t = 0; t < VEC; ++t) {
          s0.v[t] *= inv_n;          // sum_nodes / normalizer (model_zoo.py:252)
          s1.v[t] /= na;             // mean_nodes with the 0/1 weight (model_zoo.py:254)
          s2.v[t] *= inv_n;
        }
        s0.store(orow + c * VEC);
        s1.store(orow + D + c * VEC);
        s2.store(orow + 2 * D + c * VEC);
      }
      continue;
    }
    float S = 0.f;
    for (int i = beg + lane; i < end; i += 32) {
      const int r = kind == TX_READOUT_WMEAN ? __ldg(pos + i) : 0;
      S += r == 0 ? sw[0] : (r == 1 ? sw[1] : sw[2]);
    }
    S = warp_sum(S);
    for (int c = lane; c < nvec; c += 32) {
      Vec<VEC> acc = vzero<VEC>();
      for (int i = beg; i < end; ++i) {
        const int r = kind == TX_READOUT_WMEAN ? __ldg(pos + i) : 0;
        const float a = r == 0 ? sw[0] : (r == 1 ? sw[1] : sw[2]);
        const Vec<VEC> v = Vec<VEC>::load(h + (int64_t)i * ldh + c * VEC);
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc.v[t] = fmaf(a, v.v[t], acc.v[t]);
      }
#pragma unroll
      for (int t = 0; t < VEC; ++t) acc.v[t] /= S;
      acc.store(orow + c * VEC);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) readout_bwd_kernel(int kind, const float* __restrict__ dhg, int64_t lddhg,
                                                          const float* __restrict__ h, int64_t ldh,
                                                          const float* __restrict__ hg, int64_t ldhg,
                                                          const int32_t* __restrict__ pos, const float* __restrict__ pw,
                                                          const int32_t* __restrict__ node_off, int n_graphs, int D,
                                                          float* __restrict__ dh, int64_t lddh, float* __restrict__ dw_partial) {
  __shared__ float sdw[8][3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D / VEC;
  float sw[3] = {1.f, 1.f, 1.f}, sg[3] = {0.f, 0.f, 0.f};
  if (kind == TX_READOUT_WMEAN) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float w = __ldg(pw + r);
      sw[r] = softplus_f(w);
      sg[r] = w > 20.f ? 1.f : 1.f / (1.f + expf(-w));   // d softplus / dw
    }
  }
  float dw[3] = {0.f, 0.f, 0.f};
  for (int g = warp; g < n_graphs; g += nwarps) {
    const int beg = __ldg(node_off + g), end = __ldg(node_off + g + 1);
    const float* drow = dhg + (int64_t)g * lddhg;
    if (kind == TX_READOUT_CONCAT) {
      float na = 0.f;
      for (int i = beg + lane; i < end; i += 32) na += (__ldg(pos + i) == 1) ? 1.f : 0.f;
      na = warp_sum(na);
      const float inv_n = 1.f / (float)(end - beg);
      for (int i = beg; i < end; ++i) {
        const int r = __ldg(pos + i);
        const bool ok = r >= 0 && r <= 2;
        const float sc = r == 1 ? 1.f / na : inv_n;
        for (int c = lane; c < nvec; c += 32) {
          Vec<VEC> v = ok ? Vec<VEC>::load(drow + (int64_t)r * D + c * VEC) : vzero<VEC>();
#pragma unroll
          for (int t = 0; t < VEC; ++t) v.v[t] *= sc;
          v.store(dh + (int64_t)i * lddh + c * VEC);
        }
      }
      continue;
    }
    float S = 0.f;
    for (int i = beg + lane; i < end; i += 32) {
      const int r = kind == TX_READOUT_WMEAN ? __ldg(pos + i) : 0;
      S += r == 0 ? sw[0] : (r == 1 ? sw[1] : sw[2]);
    }
    S = warp_sum(S);
    const float inv_s = 1.f / S;
    for (int i = beg; i < end; ++i) {
      const int r = kind == TX_READOUT_WMEAN ? __ldg(pos + i) : 0;
      const float a = r == 0 ? sw[0] : (r == 1 ? sw[1] : sw[2]);
      const float sc = a * inv_s;
      float dot = 0.f;
      for (int c = lane; c < nvec; c += 32) {
        const Vec<VEC> d = Vec<VEC>::load(drow + c * VEC);
        Vec<VEC> o;
#pragma unroll
        for (int t = 0; t < VEC; ++t) o.v[t] = d.v[t] * sc;
        o.store(dh + (int64_t)i * lddh + c * VEC);
        if (kind == TX_READOUT_WMEAN) {
          const Vec<VEC> hv = Vec<VEC>::load(h + (int64_t)i * ldh + c * VEC);
          const Vec<VEC> gv = Vec<VEC>::load(hg + (int64_t)g * ldhg + c * VEC);
#pragma unroll
          for (int t = 0; t < VEC; ++t) dot = fmaf(d.v[t], hv.v[t] - gv.v[t], dot);
        }
      }
      if (kind == TX_READOUT_WMEAN) {
        dot = warp_sum(dot) * inv_s;
        dw[0] += r == 0 ? dot * sg[0] : 0.f;
        dw[1] += r == 1 ? dot * sg[1] : 0.f;
        dw[2] += r == 2 ? dot * sg[2] : 0.f;
      }
    }
  }
  if (kind == TX_READOUT_WMEAN && dw_partial) {
    if (lane == 0) { sdw[wid][0] = dw[0]; sdw[wid][1] = dw[1]; sdw[wid][2] = dw[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
      float t = sdw[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += sdw[w][threadIdx.x];
      dw_partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
  }
}

// ---- fast readout path (MEAN / WMEAN, 16-byte aligned rows, D <= 512): whole row of a graph per warp iteration, 4 rows
// in flight per warp (the generic kernels above keep one row in flight and were latency-bound: 105 / 165 us for 75 MB) ----
template <int NV>
__device__ __forceinline__ void ro_load(const float* __restrict__ p, int lane, int D, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    v[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

constexpr int kBigGraph = 12;   // graphs with more rows are shared by the 8 warps of a CTA (one warp per graph otherwise)
// rows a warp keeps in flight, and the CTAs per SM the register budget is held to.  4 rows (r[4][NV] = 64 registers, 2 CTAs per SM)
// vs 2 rows (4 / 3 CTAs per SM), same-box A/B: forward 0.0368 vs 0.0362 ms, backward 0.058 vs 0.067 ms - occupancy is not what bounds
// these kernels, rows in flight per warp help the backward
#ifndef TX_RO_ROWS
#define TX_RO_ROWS 4
#endif
constexpr int kRoRows = TX_RO_ROWS;
constexpr int kRoFwdCtas = TX_RO_ROWS == 2 ? 4 : 2, kRoBwdCtas = TX_RO_ROWS == 2 ? 3 : 2;

// A CTA owns groups of 8 consecutive graphs.  Egonet sizes are skewed (1..57 rows, one large positive egonet per query): with
// one warp per graph the CTA waited ~14 us for the warp that drew the 53-row graph while the others were done after 1-2 us
// (43 us for a 74 MB read).  Small graphs still get one warp each; the large ones of the group are then taken by all 8 warps
// together (rows dealt round-robin, fixed-order reduction over the warps: deterministic).
template <int NV>
__global__ void __launch_bounds__(256, kRoFwdCtas) readout_fwd_fast_kernel(int kind, const float* __restrict__ h, int64_t ldh,
                                                               const int32_t* __restrict__ pos, const float* __restrict__ pw,
                                                               const int32_t* __restrict__ node_off, int n_graphs, int D,
                                                               float* __restrict__ hg, int64_t ldhg) {
  TX_PDL_ENTER();
  __shared__ float4 s_acc[8][NV * 32];
  __shared__ float s_S[8];
  __shared__ int s_size[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float sw[3] = {1.f, 1.f, 1.f};
  if (kind == TX_READOUT_WMEAN) {
#pragma unroll
    for (int r = 0; r < 3; ++r) sw[r] = softplus_f(__ldg(pw + r));
  }
  // rows [beg, end) with stride `step` starting at beg + first: weighted sum into acc / S
  auto accumulate = [&](int beg, int end, int first, int step, float4 (&acc)[NV], float& S) {
    for (int i = beg + first; i < end; i += kRoRows * step) {
      float4 r[kRoRows][NV];
      float a[kRoRows];
      int pr[kRoRows];
#pragma unroll
      for (int u = 0; u < kRoRows; ++u) {               // every load of the batch is issued before anything is consumed (the weight of a
        const bool ok = i + u * step < end;             // row used to be selected right after its position load: the row loads waited for it)
        const int row = ok ? i + u * step : beg;
        pr[u] = kind == TX_READOUT_WMEAN ? __ldg(pos + row) : 0;
        ro_load<NV>(h + (int64_t)row * ldh, lane, D, r[u]);
      }
#pragma unroll
      for (int u = 0; u < kRoRows; ++u) a[u] = i + u * step < end ? (pr[u] == 0 ? sw[0] : (pr[u] == 1 ? sw[1] : sw[2])) : 0.f;
#pragma unroll
      for (int u = 0; u < kRoRows; ++u) {
        S += a[u];
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          acc[t].x = fmaf(a[u], r[u][t].x, acc[t].x); acc[t].y = fmaf(a[u], r[u][t].y, acc[t].y);
          acc[t].z = fmaf(a[u], r[u][t].z, acc[t].z); acc[t].w = fmaf(a[u], r[u][t].w, acc[t].w);
        }
      }
    }
  };
  for (int g0 = blockIdx.x * 8; g0 < n_graphs; g0 += gridDim.x * 8) {
    const int g = g0 + wid;
    int beg = 0, end = 0;
    if (g < n_graphs) { beg = __ldg(node_off + g); end = __ldg(node_off + g + 1); }
    if (lane == 0) s_size[wid] = end - beg;
    if (g < n_graphs && end - beg <= kBigGraph) {
      float S = 0.f;
      float4 acc[NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      accumulate(beg, end, 0, 1, acc, S);
      float* orow = hg + (int64_t)g * ldhg;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c = (lane + 32 * t) * 4;
        if (c < D) *reinterpret_cast<float4*>(orow + c) = make_float4(acc[t].x / S, acc[t].y / S, acc[t].z / S, acc[t].w / S);
      }
    }
    __syncthreads();
    for (int w = 0; w < 8; ++w) {                       // uniform over the CTA
      if (s_size[w] <= kBigGraph) continue;
      const int gb = g0 + w;
      const int b0 = __ldg(node_off + gb), b1 = __ldg(node_off + gb + 1);
      float S = 0.f;
      float4 acc[NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      accumulate(b0, b1, wid, 8, acc, S);
#pragma unroll
      for (int t = 0; t < NV; ++t) s_acc[wid][lane + 32 * t] = acc[t];
      if (lane == 0) s_S[wid] = S;
      __syncthreads();
      if (wid == 0) {
        float St = s_S[0];
#pragma unroll
        for (int v = 1; v < 8; ++v) St += s_S[v];
        float* orow = hg + (int64_t)gb * ldhg;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int c = (lane + 32 * t) * 4;
          if (c < D) {
            float4 a = s_acc[0][lane + 32 * t];
#pragma unroll
            for (int v = 1; v < 8; ++v) {
              const float4 x = s_acc[v][lane + 32 * t];
              a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
            }
            *reinterpret_cast<float4*>(orow + c) = make_float4(a.x / St, a.y / St, a.z / St, a.w / St);
          }
        }
      }
      __syncthreads();
    }
    __syncthreads();                                    // s_size is rewritten by the next group
  }
}

template <int NV>
__global__ void __launch_bounds__(256, kRoBwdCtas) readout_bwd_fast_kernel(int kind, const float* __restrict__ dhg, int64_t lddhg,
                                                               const float* __restrict__ h, int64_t ldh,
                                                               const float* __restrict__ hg, int64_t ldhg,
                                                               const int32_t* __restrict__ pos, const float* __restrict__ pw,
                                                               const int32_t* __restrict__ node_off, int n_graphs, int D,
                                                               float* __restrict__ dh, int64_t lddh, float* __restrict__ dw_partial) {
  TX_PDL_ENTER();
  __shared__ float sdw[8][3];
  __shared__ int s_size[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool wm = kind == TX_READOUT_WMEAN;
  float sw[3] = {1.f, 1.f, 1.f}, sg[3] = {0.f, 0.f, 0.f};
  if (wm) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float w = __ldg(pw + r);
      sw[r] = softplus_f(w);
      sg[r] = w > 20.f ? 1.f : 1.f / (1.f + expf(-w));
    }
  }
  float dw[3] = {0.f, 0.f, 0.f};
  // rows beg + first, + step, ... of graph g: dh rows and this warp's share of d(position weights)
  auto graph_rows = [&](int g, int beg, int end, int first, int step) {
    float4 d[NV], m[NV];
    ro_load<NV>(dhg + (int64_t)g * lddhg, lane, D, d);
    if (wm) ro_load<NV>(hg + (int64_t)g * ldhg, lane, D, m);
    float4 r[kRoRows][NV];
    int pr[kRoRows];
    auto load_batch = [&](int i) {
#pragma unroll
      for (int u = 0; u < kRoRows; ++u) {
        const int row = i + u * step < end ? i + u * step : beg;
        pr[u] = wm ? __ldg(pos + row) : 0;
        if (wm) ro_load<NV>(h + (int64_t)row * ldh, lane, D, r[u]);
      }
    };
    if (beg + first < end) load_batch(beg + first);     // the first rows are on their way while the graph's weight sum is formed
    float S = 0.f;
    for (int i = beg + lane; i < end; i += 32) {
      const int p_ = wm ? __ldg(pos + i) : 0;
      S += p_ == 0 ? sw[0] : (p_ == 1 ? sw[1] : sw[2]);
    }
    S = warp_sum(S);
    const float inv_s = 1.f / S;
    for (int i = beg + first; i < end; i += kRoRows * step) {
      if (i != beg + first) load_batch(i);
#pragma unroll
      for (int u = 0; u < kRoRows; ++u) {
        if (i + u * step < end) {                       // warp-uniform
          const float a = pr[u] == 0 ? sw[0] : (pr[u] == 1 ? sw[1] : sw[2]);
          const float sc = a * inv_s;
          float dot = 0.f;
          float* orow = dh + (int64_t)(i + u * step) * lddh;
#pragma unroll
          for (int t = 0; t < NV; ++t) {
            const int c = (lane + 32 * t) * 4;
            if (c < D) *reinterpret_cast<float4*>(orow + c) = make_float4(d[t].x * sc, d[t].y * sc, d[t].z * sc, d[t].w * sc);
            if (wm) {
              dot = fmaf(d[t].x, r[u][t].x - m[t].x, dot); dot = fmaf(d[t].y, r[u][t].y - m[t].y, dot);
              dot = fmaf(d[t].z, r[u][t].z - m[t].z, dot); dot = fmaf(d[t].w, r[u][t].w - m[t].w, dot);
            }
          }
          if (wm) {
            dot = warp_sum(dot) * inv_s;
            dw[0] += pr[u] == 0 ? dot * sg[0] : 0.f;
            dw[1] += pr[u] == 1 ? dot * sg[1] : 0.f;
            dw[2] += pr[u] == 2 ? dot * sg[2] : 0.f;
          }
        }
      }
    }
  };
  for (int g0 = blockIdx.x * 8; g0 < n_graphs; g0 += gridDim.x * 8) {
    const int g = g0 + wid;
    int beg = 0, end = 0;
    if (g < n_graphs) { beg = __ldg(node_off + g); end = __ldg(node_off + g + 1); }
    if (lane == 0) s_size[wid] = end - beg;
    // task 0: this warp's own graph if it is small; tasks 1..8: the large graphs of the group, rows dealt round-robin to the 8
    // warps (ONE call site of graph_rows: two inlined copies doubled the register count)
    for (int task = 0; task <= 8; ++task) {
      if (task == 1) __syncthreads();                   // s_size of the whole group is visible
      int gg = g, b0 = beg, b1 = end, first = 0, step = 1;
      if (task == 0) {
        if (!(g < n_graphs && end - beg <= kBigGraph)) continue;
      } else {
        if (s_size[task - 1] <= kBigGraph) continue;
        gg = g0 + task - 1;
        b0 = __ldg(node_off + gg); b1 = __ldg(node_off + gg + 1);
        first = wid; step = 8;
      }
      graph_rows(gg, b0, b1, first, step);
    }
    __syncthreads();                                    // s_size is rewritten by the next group
  }
  if (wm && dw_partial) {
    if (lane == 0) { sdw[wid][0] = dw[0]; sdw[wid][1] = dw[1]; sdw[wid][2] = dw[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
      float t = sdw[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += sdw[w][threadIdx.x];
      dw_partial[(int64_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
  }
}

__global__ void dropout_keep_mask_kernel(uint64_t seed, uint32_t stream_id, int64_t first, int64_t n, uint32_t thr,
                                         uint8_t* __restrict__ keep) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  keep[t] = drop_keep1(seed, stream_id, (uint64_t)(first + t), thr) ? 1 : 0;
}

}  // namespace tx

// =================================================================================================
// C ABI
// =================================================================================================
using namespace tx;

extern "C" {

int tx_abi_version(void) { return TX_ABI_VERSION; }
int tx_pdl_set(int enabled) { return tx::pdl_set(enabled); }
const char* tx_last_error(void) { return tx::g_err; }
const char* tx_target_arch(void) { return "sm_100a"; }
int64_t tx_row_blocks(int64_t n_rows) { return row_blocks(n_rows); }

int tx_concat_pos_dropout_fwd(const float* x, int64_t ldx, const float* pos_table, const int32_t* pos,
                              int64_t n_nodes, int64_t k_in, int64_t pos_dim, float* z, int64_t ldz, float p_drop,
                              uint64_t seed, uint32_t stream_id, void* stream) {
  TX_REQUIRE(n_nodes >= 0 && k_in >= 0 && pos_dim >= 0, "concat: negative size");
  TX_REQUIRE(ldz >= k_in + pos_dim && ldz % 4 == 0 && aligned16(z), "concat: z needs ld %% 4 == 0, ld >= k_in+pos_dim, 16B alignment");
  TX_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "concat: p_drop must be in [0,1)");
  TX_REQUIRE(pos_dim == 0 || (pos_table && pos), "concat: pos_dim > 0 needs pos_table and pos");
  TX_REQUIRE(n_nodes * ldz < (int64_t)1 << 40, "concat: too large");
  if (n_nodes == 0) return TX_OK;
  const int64_t total = n_nodes * (ldz / 4);
  const int grid = (int)((total + 255) / 256 < (int64_t)kNumSms * 16 ? (total + 255) / 256 : (int64_t)kNumSms * 16);
  TX_PDL_LAUNCH((concat_pos_dropout_fwd_kernel), grid, 256, 0, (cudaStream_t)stream, x, ldx, pos_table, pos, (int)n_nodes, (int)k_in,
                                                                        (int)pos_dim, z, (int)ldz, 1.f / (1.f - p_drop),
                                                                        drop_threshold(p_drop), seed, stream_id);
  TX_LAUNCH_CHECK("tx_concat_pos_dropout_fwd");
  return TX_OK;
}

int tx_concat_pos_dropout_f16(const float* x, int64_t ldx, const float* pos_table, const int32_t* pos, int64_t n_nodes, int64_t k_in,
                              int64_t pos_dim, int64_t vocab, float p_drop, uint64_t seed, uint32_t stream_id, const float* x_amax,
                              void* hi, void* lo, int64_t ld16, float* scale_out, void* stream) {
  TX_REQUIRE(n_nodes >= 0 && k_in >= 0 && pos_dim >= 0 && vocab >= 0, "concat_f16: negative size");
  TX_REQUIRE(ld16 >= k_in + pos_dim && ld16 % 8 == 0 && aligned16(hi) && aligned16(lo), "concat_f16: outputs need ld %% 8 == 0, ld >= k_in+pos_dim, 16B alignment");
  TX_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "concat_f16: p_drop must be in [0,1)");
  TX_REQUIRE(pos_dim == 0 || (pos_table && pos && vocab > 0), "concat_f16: pos_dim > 0 needs pos_table, pos and vocab");
  TX_REQUIRE(x_amax && scale_out, "concat_f16: x_amax (device max|x|, tx_absmax) and scale_out are required");
  TX_REQUIRE(n_nodes * ld16 < (int64_t)1 << 40 && n_nodes < INT32_MAX, "concat_f16: too large");
  if (n_nodes == 0) return TX_OK;
  const int64_t total = n_nodes * (ld16 / 4);
  const int grid = (int)((total + 255) / 256 < (int64_t)kNumSms * 16 ? (total + 255) / 256 : (int64_t)kNumSms * 16);
  const int64_t ldz = ((k_in + pos_dim + 3) / 4) * 4;
  TX_PDL_LAUNCH((concat_pos_dropout_f16_kernel), grid, 256, 0, (cudaStream_t)stream, x, ldx, pos_table, pos, (int)n_nodes, (int)k_in, (int)pos_dim,
                                                                        (int)(pos_dim > 0 ? vocab : 0), x_amax, (__half*)hi, (__half*)lo,
                                                                        (int)ld16, (int)ldz, 1.f / (1.f - p_drop), drop_threshold(p_drop),
                                                                        seed, stream_id, scale_out);
  TX_LAUNCH_CHECK("tx_concat_pos_dropout_f16");
  return TX_OK;
}

int tx_epilogue_bwd(float* dz, int64_t ldz, const float* z, const int32_t* pos, int64_t n_nodes, int64_t k_in,
                    int64_t pos_dim, int64_t vocab, float slope, float p_drop, uint64_t seed, uint32_t stream_id,
                    float* dpos_partial, void* stream) {
  TX_REQUIRE(ldz % 4 == 0 && aligned16(dz) && (!z || aligned16(z)), "epilogue_bwd: needs ld %% 4 == 0 and 16B alignment");
  TX_REQUIRE(ldz >= k_in + pos_dim, "epilogue_bwd: ldz < k_in + pos_dim");
  TX_REQUIRE(vocab <= kMaxVocab, "epilogue_bwd: position vocab %lld > %d", (long long)vocab, kMaxVocab);
  TX_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "epilogue_bwd: p_drop must be in [0,1)");
  TX_REQUIRE(pos_dim == 0 || !dpos_partial || pos, "epilogue_bwd: pos_dim > 0 needs pos");
  if (n_nodes == 0) return TX_OK;
  epilogue_bwd_kernel<<<(int)row_blocks(n_nodes), 256, 0, (cudaStream_t)stream>>>(
      dz, (int)ldz, z, pos, (int)n_nodes, (int)k_in, (int)pos_dim, (int)vocab, slope, 1.f / (1.f - p_drop),
      drop_threshold(p_drop), seed, stream_id, dpos_partial);
  TX_LAUNCH_CHECK("tx_epilogue_bwd");
  return TX_OK;
}

int tx_reduce_partials(const float* partial, int64_t n_blocks, int64_t m_len, float* out, void* stream) {
  if (m_len <= 0) return TX_OK;
  if (n_blocks <= 64 && m_len >= 4096 && m_len % 4 == 0 && aligned16(partial) && aligned16(out)) {   // split-K partials: few, long
    const int64_t m4 = m_len / 4;
    const int grid = (int)((m4 + 255) / 256 < (int64_t)kNumSms * 16 ? (m4 + 255) / 256 : (int64_t)kNumSms * 16);
    TX_PDL_LAUNCH((reduce_partials_wide_kernel), grid, 256, 0, (cudaStream_t)stream, partial, (int)n_blocks, m4, out);
    TX_LAUNCH_CHECK("tx_reduce_partials");
    return TX_OK;
  }
  if (n_blocks >= 128)
    TX_PDL_LAUNCH((reduce_partials_kernel<32>), (int)((m_len + 31) / 32), 1024, 0, (cudaStream_t)stream, partial, n_blocks, m_len, out);
  else
    TX_PDL_LAUNCH((reduce_partials_kernel<8>), (int)((m_len + 31) / 32), 256, 0, (cudaStream_t)stream, partial, n_blocks, m_len, out);
  TX_LAUNCH_CHECK("tx_reduce_partials");
  return TX_OK;
}

int tx_reduce_partials_rows(const float* partial, int64_t n_blocks, int64_t block_stride, int64_t rows, int64_t cols, int64_t ld_in,
                            float* out, int64_t ld_out, void* stream) {
  TX_REQUIRE(partial && out && n_blocks >= 1 && rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= cols && block_stride >= rows * ld_in &&
             rows < INT32_MAX && cols < INT32_MAX && n_blocks < INT32_MAX, "reduce_partials_rows: bad arguments");
  if (rows == 0 || cols == 0) return TX_OK;
  const int64_t total = rows * cols;
  const int grid = (int)((total + 255) / 256 < (int64_t)kNumSms * 16 ? (total + 255) / 256 : (int64_t)kNumSms * 16);
  TX_PDL_LAUNCH((reduce_partials_rows_kernel), grid, 256, 0, (cudaStream_t)stream, partial, (int)n_blocks, block_stride, (int)rows, (int)cols, ld_in,
                out, ld_out);
  TX_LAUNCH_CHECK("tx_reduce_partials_rows");
  return TX_OK;
}

int tx_colsum_partials(const float* x, int64_t ldx, int64_t n_rows, int64_t n_cols, float* partial, void* stream) {
  if (n_rows == 0 || n_cols == 0) return TX_OK;
  TX_PDL_LAUNCH((colsum_partials_kernel), (int)row_blocks(n_rows), 256, 0, (cudaStream_t)stream, x, ldx, (int)n_rows, (int)n_cols, partial);
  TX_LAUNCH_CHECK("tx_colsum_partials");
  return TX_OK;
}

static bool vec4_ok(const void* p, int64_t ld, int64_t dim) { return aligned16(p) && ld % 4 == 0 && dim % 4 == 0; }

int tx_gat_node_logits(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, int64_t n_nodes,
                       int64_t heads, int64_t dim, float* a1, float* a2, void* stream) {
  TX_REQUIRE(heads > 0 && dim > 0 && ldf >= heads * dim, "gat_node_logits: bad shape");
  if (n_nodes == 0) return TX_OK;
  const int grid = grid_for_warps(n_nodes * heads, 8, 8);
  if (vec4_ok(ft, ldf, dim) && aligned16(attn_l) && aligned16(attn_r))
    gat_node_logits_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(ft, ldf, attn_l, attn_r, (int)n_nodes, (int)heads, (int)dim, a1, a2);
  else
    gat_node_logits_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(ft, ldf, attn_l, attn_r, (int)n_nodes, (int)heads, (int)dim, a1, a2);
  TX_LAUNCH_CHECK("tx_gat_node_logits");
  return TX_OK;
}

int tx_gat_aggregate_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const float* a1,
                         const float* a2, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                         int64_t n_nodes, int64_t n_edges, int64_t heads, int64_t dim, float neg_slope,
                         float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha, float* alpha_d,
                         float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, void* stream) {
  (void)attn_l; (void)attn_r; (void)n_edges;
  TX_REQUIRE(heads > 0 && dim > 0 && ldf >= heads * dim, "gat_aggregate_fwd: bad shape");
  TX_REQUIRE(a1 && a2, "gat_aggregate_fwd: the general-CSR kernel needs a1/a2 (tx_gat_node_logits)");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f, "gat_aggregate_fwd: p_attn must be in [0,1)");
  TX_REQUIRE(alpha && elog && (p_attn == 0.f || (alpha_d && alpha_d != alpha)), "gat_aggregate_fwd: alpha/elog/alpha_d buffers");
  GatFwdParams p;
  if (make_epilogue(epi, &p.ep) != TX_OK) return TX_ERR_INVALID_ARGUMENT;
  const int64_t need = p.ep.mean_heads ? dim : heads * dim + p.ep.pos_dim;
  TX_REQUIRE(ldo >= need, "gat_aggregate_fwd: ldo %lld < %lld", (long long)ldo, (long long)need);
  if (n_nodes == 0) return TX_OK;
  p.ft = ft; p.ldf = ldf; p.a1 = a1; p.a2 = a2; p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid;
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.out = out; p.ldo = ldo;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  if (vec4_ok(ft, ldf, dim) && vec4_ok(out, ldo, dim))
    gat_aggregate_fwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else
    gat_aggregate_fwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  TX_LAUNCH_CHECK("tx_gat_aggregate_fwd");
  return TX_OK;
}

int tx_gat_aggregate_bwd_dst(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft,
                             int64_t ldf, const float* alpha, const float* elog, const int32_t* in_ptr,
                             const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads,
                             int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed,
                             uint32_t attn_stream_id, float* ds, float* da2, void* stream) {
  TX_REQUIRE(heads > 0 && dim > 0, "gat_aggregate_bwd_dst: bad shape");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f, "gat_aggregate_bwd_dst: p_attn must be in [0,1)");
  if (n_nodes == 0) return TX_OK;
  GatBwdDstParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.ft = ft; p.ldf = ldf;
  p.alpha = alpha; p.elog = elog; p.in_ptr = in_ptr; p.in_src = in_src; p.in_eid = in_eid;
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id;
  p.ds = ds; p.da2 = da2;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  if (vec4_ok(g, ldg, dim) && vec4_ok(ft, ldf, dim) && g_head_stride % 4 == 0)
    gat_aggregate_bwd_dst_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else
    gat_aggregate_bwd_dst_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  TX_LAUNCH_CHECK("tx_gat_aggregate_bwd_dst");
  return TX_OK;
}

int tx_gat_aggregate_bwd_src(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale,
                             const float* alpha_d, const float* ds, const float* da2, const float* attn_l,
                             const float* attn_r, const int32_t* out_ptr, const int32_t* out_dst,
                             const int32_t* out_slot, int64_t n_nodes, int64_t heads, int64_t dim, float* da1,
                             float* dft, int64_t ldd, void* stream) {
  TX_REQUIRE(heads > 0 && dim > 0 && ldd >= heads * dim, "gat_aggregate_bwd_src: bad shape");
  if (n_nodes == 0) return TX_OK;
  GatBwdSrcParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.alpha_d = alpha_d; p.ds = ds; p.da2 = da2;
  p.attn_l = attn_l; p.attn_r = attn_r; p.out_ptr = out_ptr; p.out_dst = out_dst; p.out_slot = out_slot;
  p.n = (int)n_nodes; p.H = (int)heads; p.D = (int)dim; p.da1 = da1; p.dft = dft; p.ldd = ldd;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  if (vec4_ok(g, ldg, dim) && vec4_ok(dft, ldd, dim) && g_head_stride % 4 == 0 && aligned16(attn_l) && aligned16(attn_r))
    gat_aggregate_bwd_src_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else
    gat_aggregate_bwd_src_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  TX_LAUNCH_CHECK("tx_gat_aggregate_bwd_src");
  return TX_OK;
}

int tx_gat_attn_grad_partials(const float* ft, int64_t ldf, const float* da1, const float* da2, int64_t n_nodes,
                              int64_t heads, int64_t dim, float* partial, void* stream) {
  TX_REQUIRE(heads > 0 && dim > 0, "gat_attn_grad_partials: bad shape");
  if (n_nodes == 0) return TX_OK;
  gat_attn_grad_partials_kernel<<<(int)row_blocks(n_nodes), 256, 0, (cudaStream_t)stream>>>(ft, ldf, da1, da2, (int)n_nodes,
                                                                                            (int)heads, (int)dim, partial);
  TX_LAUNCH_CHECK("tx_gat_attn_grad_partials");
  return TX_OK;
}

int tx_gcn_norm(const int32_t* in_ptr, int64_t n_nodes, float* norm, void* stream) {
  if (n_nodes == 0) return TX_OK;
  TX_PDL_LAUNCH((gcn_norm_kernel), (int)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream, in_ptr, (int)n_nodes, norm);
  TX_LAUNCH_CHECK("tx_gcn_norm");
  return TX_OK;
}

int tx_gcn_aggregate_fwd(const float* y, int64_t ldy, const float* norm, const float* bias, const int32_t* in_ptr,
                         const int32_t* in_src, int64_t n_nodes, int64_t dim, float* out, int64_t ldo,
                         const tx_gat_epilogue* epi, void* stream) {
  TX_REQUIRE(dim > 0 && ldy >= dim, "gcn_aggregate_fwd: bad shape");
  Epilogue ep;
  if (make_epilogue(epi, &ep) != TX_OK) return TX_ERR_INVALID_ARGUMENT;
  const int64_t need = ep.mean_heads ? dim : dim + ep.pos_dim;
  TX_REQUIRE(ldo >= need, "gcn_aggregate_fwd: ldo %lld < %lld", (long long)ldo, (long long)need);
  if (n_nodes == 0) return TX_OK;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  if (vec4_ok(y, ldy, dim) && vec4_ok(out, ldo, dim) && (!bias || aligned16(bias)))
    gcn_aggregate_fwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(y, ldy, norm, bias, in_ptr, in_src, (int)n_nodes, (int)dim, out, ldo, ep);
  else
    gcn_aggregate_fwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(y, ldy, norm, bias, in_ptr, in_src, (int)n_nodes, (int)dim, out, ldo, ep);
  TX_LAUNCH_CHECK("tx_gcn_aggregate_fwd");
  return TX_OK;
}

int tx_gcn_aggregate_fwd_f16(const float* y, int64_t ldy, const float* norm, const float* bias, const int32_t* in_ptr,
                             const int32_t* in_src, int64_t n_nodes, int64_t dim, int64_t ldo, const tx_gat_epilogue* epi, void* out_hi,
                             void* out_lo, int64_t ld16, const float* bound, float* scale_out, uint32_t* maskbits, void* stream) {
  TX_REQUIRE(dim > 0 && dim % 4 == 0 && ldy >= dim && vec4_ok(y, ldy, dim) && (!bias || aligned16(bias)), "gcn_aggregate_fwd_f16: dim % 4 == 0 and 16-byte aligned rows required");
  TX_REQUIRE(epi && !epi->mean_heads, "gcn_aggregate_fwd_f16: hidden layers only");
  Epilogue ep;
  if (make_epilogue(epi, &ep) != TX_OK) return TX_ERR_INVALID_ARGUMENT;
  TX_REQUIRE(out_hi && out_lo && bound && aligned16(out_hi) && aligned16(out_lo) && ld16 % 8 == 0 && ld16 >= dim + ep.pos_dim && ldo % 4 == 0 &&
             ldo >= dim + ep.pos_dim, "gcn_aggregate_fwd_f16: bad output buffers");
  if (n_nodes == 0) return TX_OK;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  TX_PDL_LAUNCH((gcn_aggregate_fwd_f16_kernel), grid, 256, 0, (cudaStream_t)stream, y, ldy, norm, bias, in_ptr, in_src, (int)n_nodes, (int)dim, ldo, ep,
                                                                      (__half*)out_hi, (__half*)out_lo, ld16, bound, scale_out,
                                                                      reinterpret_cast<uint8_t*>(maskbits), (int)tx_gat_fused_mask_ld(1, dim));
  TX_LAUNCH_CHECK("tx_gcn_aggregate_fwd_f16");
  return TX_OK;
}

int tx_gcn_aggregate_bwd_f16(const float* g, int64_t ldg, const float* norm, const int32_t* out_ptr, const int32_t* out_dst, int64_t n_nodes,
                             int64_t dim, void* dy_hi, void* dy_lo, int64_t ld16, const float* bound, float* scale_out, void* stream) {
  TX_REQUIRE(dim > 0 && dim % 4 == 0 && ldg >= dim && vec4_ok(g, ldg, dim), "gcn_aggregate_bwd_f16: dim % 4 == 0 and 16-byte aligned rows required");
  TX_REQUIRE(dy_hi && dy_lo && bound && aligned16(dy_hi) && aligned16(dy_lo) && ld16 % 8 == 0 && ld16 >= dim, "gcn_aggregate_bwd_f16: bad output buffers");
  if (n_nodes == 0) return TX_OK;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  TX_PDL_LAUNCH((gcn_aggregate_bwd_f16_kernel), grid, 256, 0, (cudaStream_t)stream, g, ldg, norm, out_ptr, out_dst, (int)n_nodes, (int)dim, (__half*)dy_hi,
                                                                      (__half*)dy_lo, ld16, bound, scale_out);
  TX_LAUNCH_CHECK("tx_gcn_aggregate_bwd_f16");
  return TX_OK;
}

int tx_bound_gcn(const float* a, float ca, const float* b, int64_t b_len, float cb, const float* t, int64_t t_len, float ct, float* out, void* stream) {
  TX_REQUIRE(a && out && b_len >= 0 && t_len >= 0, "bound_gcn: bad arguments");
  TX_PDL_LAUNCH((bound_gcn_kernel), 1, 256, 0, (cudaStream_t)stream, a, ca, b_len > 0 ? b : nullptr, b_len, cb, t_len > 0 ? t : nullptr, t_len, ct, out);
  TX_LAUNCH_CHECK("tx_bound_gcn");
  return TX_OK;
}

int tx_gcn_aggregate_bwd(const float* g, int64_t ldg, const float* norm, const int32_t* out_ptr,
                         const int32_t* out_dst, int64_t n_nodes, int64_t dim, float* dy, int64_t ldd, void* stream) {
  TX_REQUIRE(dim > 0 && ldg >= dim && ldd >= dim, "gcn_aggregate_bwd: bad shape");
  if (n_nodes == 0) return TX_OK;
  const int grid = grid_for_warps(n_nodes, 8, 8);
  if (vec4_ok(g, ldg, dim) && vec4_ok(dy, ldd, dim))
    gcn_aggregate_bwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(g, ldg, norm, out_ptr, out_dst, (int)n_nodes, (int)dim, dy, ldd);
  else
    gcn_aggregate_bwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g, ldg, norm, out_ptr, out_dst, (int)n_nodes, (int)dim, dy, ldd);
  TX_LAUNCH_CHECK("tx_gcn_aggregate_bwd");
  return TX_OK;
}

static int readout_grid(int64_t n_graphs) { return grid_for_warps(n_graphs, 8, 8); }
int64_t tx_readout_bwd_blocks(int64_t n_graphs) { return readout_grid(n_graphs); }

int tx_readout_fwd(int32_t kind, const float* h, int64_t ldh, const int32_t* pos, const float* pos_weight,
                   const int32_t* node_off, int64_t n_graphs, int64_t dim, float* hg, int64_t ldhg, void* stream) {
  TX_REQUIRE(kind >= 0 && kind <= 2, "readout_fwd: unknown kind %d", kind);
  TX_REQUIRE(kind == TX_READOUT_MEAN || pos, "readout_fwd: pos required");
  TX_REQUIRE(kind != TX_READOUT_WMEAN || pos_weight, "readout_fwd: pos_weight required for WMEAN");
  TX_REQUIRE(ldhg >= (kind == TX_READOUT_CONCAT ? 3 * dim : dim), "readout_fwd: ldhg too small");
  if (n_graphs == 0) return TX_OK;
  const int grid = readout_grid(n_graphs);
  if (kind != TX_READOUT_CONCAT && dim <= 512 && vec4_ok(h, ldh, dim) && vec4_ok(hg, ldhg, dim)) {
    switch ((int)((dim + 127) / 128)) {
      case 1: TX_PDL_LAUNCH((readout_fwd_fast_kernel<1>), grid, 256, 0, (cudaStream_t)stream, kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg); break;
      case 2: TX_PDL_LAUNCH((readout_fwd_fast_kernel<2>), grid, 256, 0, (cudaStream_t)stream, kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg); break;
      case 3: TX_PDL_LAUNCH((readout_fwd_fast_kernel<3>), grid, 256, 0, (cudaStream_t)stream, kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg); break;
      default: TX_PDL_LAUNCH((readout_fwd_fast_kernel<4>), grid, 256, 0, (cudaStream_t)stream, kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg); break;
    }
    TX_LAUNCH_CHECK("tx_readout_fwd");
    return TX_OK;
  }
  if (vec4_ok(h, ldh, dim) && vec4_ok(hg, ldhg, dim))
    readout_fwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg);
  else
    readout_fwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(kind, h, ldh, pos, pos_weight, node_off, (int)n_graphs, (int)dim, hg, ldhg);
  TX_LAUNCH_CHECK("tx_readout_fwd");
  return TX_OK;
}

int tx_readout_bwd(int32_t kind, const float* dhg, int64_t lddhg, const float* h, int64_t ldh, const float* hg,
                   int64_t ldhg, const int32_t* pos, const float* pos_weight, const int32_t* node_off,
                   int64_t n_graphs, int64_t dim, float* dh, int64_t lddh, float* dw_partial, void* stream) {
  TX_REQUIRE(kind >= 0 && kind <= 2, "readout_bwd: unknown kind %d", kind);
  TX_REQUIRE(kind == TX_READOUT_MEAN || pos, "readout_bwd: pos required");
  TX_REQUIRE(kind != TX_READOUT_WMEAN || (pos_weight && h && hg), "readout_bwd: WMEAN needs pos_weight, h, hg");
  if (n_graphs == 0) return TX_OK;
  const int grid = readout_grid(n_graphs);
  const bool v4 = vec4_ok(dhg, lddhg, dim) && vec4_ok(dh, lddh, dim) && (kind != TX_READOUT_WMEAN || (vec4_ok(h, ldh, dim) && vec4_ok(hg, ldhg, dim)));
  if (v4 && kind != TX_READOUT_CONCAT && dim <= 512) {
    cudaStream_t st = (cudaStream_t)stream;
    switch ((int)((dim + 127) / 128)) {
      case 1: TX_PDL_LAUNCH((readout_bwd_fast_kernel<1>), grid, 256, 0, st, kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off, (int)n_graphs, (int)dim, dh, lddh, dw_partial); break;
      case 2: TX_PDL_LAUNCH((readout_bwd_fast_kernel<2>), grid, 256, 0, st, kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off, (int)n_graphs, (int)dim, dh, lddh, dw_partial); break;
      case 3: TX_PDL_LAUNCH((readout_bwd_fast_kernel<3>), grid, 256, 0, st, kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off, (int)n_graphs, (int)dim, dh, lddh, dw_partial); break;
      default: TX_PDL_LAUNCH((readout_bwd_fast_kernel<4>), grid, 256, 0, st, kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off, (int)n_graphs, (int)dim, dh, lddh, dw_partial); break;
    }
    TX_LAUNCH_CHECK("tx_readout_bwd");
    return TX_OK;
  }
  if (v4)
    readout_bwd_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off,
                                                                  (int)n_graphs, (int)dim, dh, lddh, dw_partial);
  else
    readout_bwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(kind, dhg, lddhg, h, ldh, hg, ldhg, pos, pos_weight, node_off,
                                                                  (int)n_graphs, (int)dim, dh, lddh, dw_partial);
  TX_LAUNCH_CHECK("tx_readout_bwd");
  return TX_OK;
}

int tx_dropout_keep_mask(uint64_t seed, uint32_t stream_id, int64_t first_index, int64_t n, float p_drop,
                         uint8_t* keep, void* stream) {
  TX_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout_keep_mask: p_drop must be in [0,1)");
  if (n <= 0) return TX_OK;
  dropout_keep_mask_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, first_index, n,
                                                                                    drop_threshold(p_drop), keep);
  TX_LAUNCH_CHECK("tx_dropout_keep_mask");
  return TX_OK;
}

}  // extern "C"
