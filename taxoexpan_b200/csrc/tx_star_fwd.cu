// Star-egonet fused GAT forward (sm_100a): the gather-attend-aggregate kernel specialised for the ONLY graph shape the reference
// ever batches - the egonets of data_loader/dataset.py:404-437: nodes [grand-parents.., anchor, siblings..], edges
// [gp -> anchor.., anchor -> sibling.., one self loop per node].  Same arithmetic and outputs as tx_gat_fused_fwd (reference
// model_zoo.py:83-96,106-114 + the next layer's :214-216,82), but nothing is chased through a CSR:
//   * in-edges in closed form from the egonet's (first node, first edge, n_gp, n_sib): grand-parent <- {self}; anchor <- {gp_0.., self};
//     sibling <- {anchor, self}; slots / edge ids as written by tx_star_batch_structure;
//   * every ft row is loaded ONCE per task and both half-logits a1 = <ft, attn_l>, a2 = <ft, attn_r> come from that one load
//     (the general kernel recomputes a1 of the anchor for every sibling and re-gathers the anchor row through L1/L2);
//   * the anchor row and its a1 stay in registers while the siblings stream by: a sibling's output is the closed-form two-way
//     softmax  out = a~_1 ft_anchor + a~_2 ft_self  (no online-softmax rescaling);
//   * work items are (egonet, chunk of C siblings, head), one 16-byte host-built record each; chunk 0 also owns the grand-parents and
//     the anchor.  Warps pull items from a self-resetting atomic queue (egonets have 1..2000 nodes: static dealing leaves a long
//     tail).  Neither a register double buffer for the next row nor a bulk L2 prefetch of the NEXT item's rows (TAXO_STAR_PREFETCH)
//     measured any gain: 24 resident warps per SM hide the loads, the kernel is issue-bound.
// Each output element is produced by exactly one item in a fixed order: results are run-to-run deterministic.
#include <math.h>
#include <stdlib.h>

#define TX_PDL_GROUP 4
#include "tx_common.cuh"

namespace tx {

#ifndef TX_STAR_MIN_BLOCKS
#define TX_STAR_MIN_BLOCKS 3
#endif
constexpr int kStarChunk = 4;           // default siblings per work item (the caller's task table says which chunk size it used)
constexpr int kStarMaxChunks = 128;     // chunk index lives in the top 7 bits of a task word; egonet index in the low 24

struct StarFwdParams {
  const float* ft; int64_t ldf;
  const float* attn_l; const float* attn_r;
  const int32_t* tasks; int n_tasks; int chunk;
  int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* alpha; float* alpha_d; float* elog;
  float* out; int64_t ldo;
  __half* out16_hi; __half* out16_lo; int64_t ld16; const float* bound; float* scale_out;
  uint8_t* maskbits; int mask_ld;
  int hidden; float act_slope; const float* next_pos_table; const int32_t* pos; int pos_dim;
  float next_inv_keep; uint32_t next_thr; uint64_t next_seed; uint32_t next_stream;
  int* queue;                             // [32 * H]: per head (stride 32) {next item, warps retired}; zero before the first launch, self-resetting
  int prefetch;                           // L2 prefetch of the next item's rows (TAXO_STAR_PREFETCH): 0 none (default: measured no gain - the
                                          // kernel is issue-bound, not latency-bound), 1 anchor + first row, 2 all rows
};

struct StarTask { int o, q, a, s, c; };   // first node, first edge / slot, #grand-parents, #siblings, chunk

template <int NV>
__device__ __forceinline__ void star_load_row(const float* __restrict__ p, int lane, int D, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    // D > 128 (NV - 1): only the last column block can be partial
    v[t] = (t < NV - 1 || c < D) ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// both half-logits of one row from one pass over it: dl = <row, attn_l>, dr = <row, attn_r> (warp-reduced, shuffles interleaved)
template <int NV>
__device__ __forceinline__ void star_dots(const float4 (&a)[NV], const float4* __restrict__ s_l, const float4* __restrict__ s_r,
                                          int lane, float& dl, float& dr) {
  float x = 0.f, y = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const float4 l = s_l[lane + 32 * t], r = s_r[lane + 32 * t];
    x = fmaf(a[t].x, l.x, x); x = fmaf(a[t].y, l.y, x); x = fmaf(a[t].z, l.z, x); x = fmaf(a[t].w, l.w, x);
    y = fmaf(a[t].x, r.x, y); y = fmaf(a[t].y, r.y, y); y = fmaf(a[t].z, r.z, y); y = fmaf(a[t].w, r.w, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float tx_ = __shfl_xor_sync(0xffffffffu, x, o), ty_ = __shfl_xor_sync(0xffffffffu, y, o);
    x += tx_; y += ty_;
  }
  dl = x; dr = y;
}

// epilogue of one (row i, head h): next layer's input (leaky-relu, feat-dropout, fp16 hi/lo or fp32) + sign/keep bytes; the warp of
// the last head also appends drop(P_next[pos_i]) and the zero padding.  Same element order / dropout counters as tx_gat_fused_fwd.
// MODE 1 = the training hot path, resolved at compile time: hidden layer, fp16 hi/lo output, leaky-relu and feat-dropout both on,
// sign/keep bytes written.  Per-row pointers are formed once; 1/(1-p) is folded into the (power-of-two) split scale, which leaves
// every written bit unchanged.  MODE 0 = everything decided at run time (output layer, eval mode, fp32 output).
template <int NV, int MODE>
__device__ __forceinline__ void star_epilogue(const StarFwdParams& p, int i, int h, int lane, float scale16, float4 (&acc)[NV]) {
  const int H = p.H, D = p.D;
  if constexpr (MODE == 1) {
    const int64_t o16 = (int64_t)i * p.ld16 + h * D + lane * 4;
    __half* hi = p.out16_hi + o16;
    __half* lo = p.out16_lo + o16;
    uint8_t* mrow = p.maskbits + (int64_t)i * p.mask_ld + ((h * D) >> 2) + lane;
    uint64_t idx4_base = (uint64_t)(((int64_t)i * p.ldo + (int64_t)h * D) >> 2) + (uint64_t)lane;
    // the row's pointers are formed ONCE: without the opaque moves nvcc re-derives all of them (~40 integer instructions) inside
    // every one of the NV column blocks
    asm volatile("" : "+l"(hi), "+l"(lo), "+l"(mrow), "+l"(idx4_base));
    const float slope = p.act_slope, sck = scale16 * p.next_inv_keep;
    const uint32_t thr = p.next_thr;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      if (t < NV - 1 || (lane + 32 * t) * 4 < D) {     // D > 128 (NV - 1): only the last column block can be partial
        float v[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
        uint32_t code = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool pos = v[u] > 0.f;
          code |= pos ? (1u << u) : 0u;
          v[u] = pos ? v[u] : v[u] * slope;
        }
        const uint2 w = drop_words(p.next_seed, p.next_stream, idx4_base + 32u * t);
        const uint32_t r[4] = {w.x & 0xFFFFu, w.x >> 16, w.y & 0xFFFFu, w.y >> 16};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool keep = r[u] >= thr;
          code |= keep ? (16u << u) : 0u;
          v[u] = keep ? v[u] * sck : 0.f;
        }
        // fp16 hi/lo split without the +-65504 clamp of f16_split4: |v| sck <= 2^13 is guaranteed by the caller's bound (max|ft| is
        // measured by the GEMM that produced ft, attention weights are convex), so the clamp is dead code here
        const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
        *reinterpret_cast<uint2*>(hi + 128 * t) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        *reinterpret_cast<uint2*>(lo + 128 * t) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        mrow[32 * t] = (uint8_t)code;
      }
    }
  } else {
  const int64_t row_off = (int64_t)i * p.ldo + (p.hidden ? (int64_t)h * D : 0) + lane * 4;
  float* orow = p.out ? p.out + row_off : nullptr;
  uint8_t* mrow = p.maskbits ? p.maskbits + (int64_t)i * p.mask_ld + ((h * D) >> 2) + lane : nullptr;
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
      } else {
        *reinterpret_cast<float4*>(orow + 128 * t) = make_float4(v[0], v[1], v[2], v[3]);
      }
      if (mrow) mrow[32 * t] = (uint8_t)code;     // 4 sign bits | 4 keep bits << 4 of this lane's 4 columns
    }
  }
  }
  if (p.hidden && h == H - 1) {
    const int feat = H * D;
    const int pd = p.pos_dim;
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
      } else {
        p.out[(int64_t)i * p.ldo + c] = v;
      }
    }
  }
}

template <int NV, int MODE>
__global__ void __launch_bounds__(256, TX_STAR_MIN_BLOCKS) gat_star_fwd_kernel(const StarFwdParams p) {
  TX_PDL_ENTER();
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
  const bool attn_drop = p.attn_thr != 0;
  const float* base = p.ft + (int64_t)h * D;
  const uint32_t rowB = (uint32_t)D * 4u;
  const float scale16 = p.out16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.out16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;
  int* qn = p.queue + 32 * h;              // one 128-byte line per head: the heads' tickets do not serialise on one L2 sector

  // work item record (host-built, one 16-byte load): {first node o, first edge / slot q, n_gp | chunk << 24, n_sib}
  auto decode = [&](int item) -> StarTask {
    const int4 w = __ldg(reinterpret_cast<const int4*>(p.tasks) + item);
    StarTask k;
    k.o = w.x; k.q = w.y; k.a = w.z & 0xFFFFFF; k.c = (w.z >> 24) & 0x7F; k.s = w.w;
    return k;
  };
  // rows of an item: chunk 0 -> [o, o + a + 1 + min(s, C)) (grand-parents, anchor, first siblings: contiguous);
  //                  chunk c -> the anchor row + siblings [c C, min(s, (c + 1) C))
  auto prefetch_item = [&](const StarTask& k) {
    const int k0 = k.c * p.chunk, k1 = min(k.s, k0 + p.chunk);
    const int rb = k.c == 0 ? k.o : k.o + k.a + 1 + k0, re = k.o + k.a + 1 + k1;
    if (p.prefetch == 0) return;
    for (int r = rb + lane; r < (p.prefetch == 1 ? min(re, rb + 1) : re); r += 32)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (int64_t)r * p.ldf), "r"(rowB) : "memory");
    if ((k.c != 0 || p.prefetch == 1) && lane == 31)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (int64_t)(k.o + k.a) * p.ldf), "r"(rowB) : "memory");
  };
  auto keepw = [&](int eid) -> float {
    if (!attn_drop) return 1.f;
    return drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)eid * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
  };
  auto lrelu = [&](float s) -> float { return s > 0.f ? s : s * p.neg_slope; };

  // ONE instance of every expensive code block (row load, dots, epilogue): the row kinds share a single loop body.  (A first version
  // with one inlined epilogue per kind was 85 KB of SASS and spent most of its time in instruction fetch.)
  // Item pipeline: while item n is processed, item n + 1 is decoded and its rows are on their way into L2, and the queue ticket of
  // item n + 2 is in flight (the atomic is issued before the work and read after it).
  int cur = 0, nxt = 0;
  if (lane == 0) { cur = atomicAdd(qn, 1); nxt = atomicAdd(qn, 1); }
  cur = __shfl_sync(0xffffffffu, cur, 0);
  nxt = __shfl_sync(0xffffffffu, nxt, 0);
  StarTask ck = {0, 0, 0, 0, 0}, nk = {0, 0, 0, 0, 0};
  if (cur < p.n_tasks) ck = decode(cur);
  if (nxt < p.n_tasks) { nk = decode(nxt); prefetch_item(nk); }
  while (cur < p.n_tasks) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(qn, 1);

    const int o = ck.o, q = ck.q, a = ck.a, s = ck.s;
    const int self0 = q + a + s;                       // edge id of the first self loop (tx_star_batch_structure)
    const int deg = a + 1;                             // in-degree of the anchor: every grand-parent, then its self loop
    // local rows of this item: chunk 0 -> [0, a + 1 + min(s, C)) = grand-parents, anchor, first siblings; chunk c -> its siblings
    const int k0 = ck.c * p.chunk, k1 = min(s, k0 + p.chunk);
    const int j0 = ck.c == 0 ? 0 : a + 1 + k0, j1 = a + 1 + k1;
    float4 ra[NV];
    star_load_row<NV>(base + (int64_t)(o + a) * p.ldf, lane, D, ra);
    float a1a, a2a;
    star_dots<NV>(ra, s_l, s_r, lane, a1a, a2a);
    float m = -INFINITY, l = 0.f, s_mine = 0.f, kw_mine = 1.f;     // softmax statistics over the anchor's in-edges (chunk 0 only)
    // attention-dropout weight of ONE in-edge of the item per lane (one hash per lane instead of one per row, head and WARP: the
    // row-by-row hashes were 9 % of the kernel's instructions).  chunk 0: [gp self loops (a) | gp -> anchor (a) | anchor self |
    // (anchor -> sibling, sibling self) x siblings of the chunk]; chunk c: the sibling pairs only
    const int e_sib0 = ck.c == 0 ? 2 * a + 1 : 0;
    const bool lanes_kw = e_sib0 + 2 * (k1 - k0) <= 32;
    float kw_lane = 1.f;
    if (attn_drop && lanes_kw) {
      int eid = -1;
      if (ck.c == 0 && lane < a) eid = self0 + lane;
      else if (ck.c == 0 && lane < 2 * a) eid = q + lane - a;
      else if (ck.c == 0 && lane == 2 * a) eid = self0 + a;
      else if (lane >= e_sib0 && lane < e_sib0 + 2 * (k1 - k0)) {
        const int mm = k0 + ((lane - e_sib0) >> 1);
        eid = ((lane - e_sib0) & 1) ? self0 + a + 1 + mm : q + a + mm;
      }
      if (eid >= 0) kw_lane = keepw(eid);
    }
    auto kw_of = [&](int idx, int eid) -> float { return lanes_kw ? __shfl_sync(0xffffffffu, kw_lane, idx) : keepw(eid); };

    for (int j = j0; j < j1; ++j) {
      const bool is_anchor = j == a, is_sib = j > a;
      float4 r[NV];
      float a1r = a1a, a2r = a2a;
      if (is_anchor) {
#pragma unroll
        for (int t = 0; t < NV; ++t) r[t] = ra[t];
      } else {
        // (keeping the next row in flight in a second register buffer measured no gain: 24 resident warps hide the load, and the
        //  copy + 16 registers cost ~4 % of the instructions)
        star_load_row<NV>(base + (int64_t)(o + j) * p.ldf, lane, D, r);
        star_dots<NV>(r, s_l, s_r, lane, a1r, a2r);
      }
      // self loop of local row j (edge id self0 + j): lane j (gp), 2a (anchor) or the sibling's second slot
      const float kw_self = kw_of(is_sib ? e_sib0 + 2 * (j - a - 1 - k0) + 1 : (is_anchor ? 2 * a : j), self0 + j);
      const float s_self = lrelu(a1r + a2r);
      if (!is_sib) {
        // in-edge number min(j, a) of the anchor: gp_j -> anchor (edge id q + j) or the anchor's self loop
        const float sv = is_anchor ? s_self : lrelu(a1r + a2a);
        const float kw = is_anchor ? kw_self : kw_of(a + j, q + j);
        if (lane == (j & 31)) { s_mine = sv; kw_mine = kw; }
        if (deg > 32 && lane == 0) p.elog[(int64_t)(q + a + j) * H + h] = sv;
        const float m_new = fmaxf(m, sv);
        l = fmaf(l, expf(m - m_new), expf(sv - m_new));
        m = m_new;
        if (!is_anchor) {
          // the grand-parent's own output: its only in-edge is its self loop -> alpha = 1 (slot q + j)
          if (lane == 0) {
            const int64_t so = (int64_t)(q + j) * H + h;
            p.elog[so] = s_self;
            p.alpha[so] = 1.f;
            if (attn_drop) p.alpha_d[so] = kw_self;
          }
#pragma unroll
          for (int t = 0; t < NV; ++t) { r[t].x *= kw_self; r[t].y *= kw_self; r[t].z *= kw_self; r[t].w *= kw_self; }
        } else {
          // anchor: out = sum_k alpha~_k ft_gp_k + alpha~_self ft_anchor.  The logits were collected while the grand-parents streamed
          // by; their rows are re-read here (they were loaded moments ago by this warp: L1 / L2 hits) instead of being carried in
          // 16 accumulator registers through the whole item.
          const float inv_l = 1.f / l;
          if (deg > 32) __syncwarp();
          {
            const float w = expf(s_self - m) * inv_l * kw_self;
#pragma unroll
            for (int t = 0; t < NV; ++t) { r[t].x *= w; r[t].y *= w; r[t].z *= w; r[t].w *= w; }
          }
          for (int k = 0; k < a; ++k) {
            float4 nx[NV];
            star_load_row<NV>(base + (int64_t)(o + k) * p.ldf, lane, D, nx);
            float sk, kwk;
            if (k < 32 && deg <= 32) {
              sk = __shfl_sync(0xffffffffu, s_mine, k);
              kwk = __shfl_sync(0xffffffffu, kw_mine, k);
            } else {
              sk = p.elog[(int64_t)(q + a + k) * H + h];
              kwk = keepw(q + k);
            }
            const float w = expf(sk - m) * inv_l * kwk;
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              r[t].x = fmaf(w, nx[t].x, r[t].x); r[t].y = fmaf(w, nx[t].y, r[t].y);
              r[t].z = fmaf(w, nx[t].z, r[t].z); r[t].w = fmaf(w, nx[t].w, r[t].w);
            }
          }
          if (deg <= 32) {
            if (lane < deg) {
              const int64_t so = (int64_t)(q + a + lane) * H + h;
              const float al = expf(s_mine - m) * inv_l;
              p.elog[so] = s_mine;
              p.alpha[so] = al;
              if (attn_drop) p.alpha_d[so] = al * kw_mine;
            }
          } else {
            for (int k = lane; k < deg; k += 32) {
              const int64_t so = (int64_t)(q + a + k) * H + h;
              const float al = expf(p.elog[so] - m) * inv_l;
              p.alpha[so] = al;
              if (attn_drop) p.alpha_d[so] = al * keepw(k < a ? q + k : self0 + a);
            }
          }
        }
      } else {
        // sibling k = j - a - 1: in-edges {anchor -> sib (edge id q + a + k = q + j - 1), self loop}; slots q + 2a + 1 + 2k, + 1
        const float s1 = lrelu(a1a + a2r);
        const float kw1 = kw_of(e_sib0 + 2 * (j - a - 1 - k0), q + j - 1);
        const float mx = fmaxf(s1, s_self);
        const float e1 = expf(s1 - mx), e2 = expf(s_self - mx);
        const float inv = 1.f / (e1 + e2);
        const float al1 = e1 * inv, al2 = e2 * inv;
        if (lane < 2) {
          const int64_t so = (int64_t)(q + 2 * j - 1 + lane) * H + h;      // q + 2a + 1 + 2(j - a - 1) = q + 2j - 1
          const float al = lane ? al2 : al1;
          p.elog[so] = lane ? s_self : s1;
          p.alpha[so] = al;
          if (attn_drop) p.alpha_d[so] = al * (lane ? kw_self : kw1);
        }
        const float w1 = al1 * kw1, w2 = al2 * kw_self;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          r[t].x = fmaf(w1, ra[t].x, w2 * r[t].x); r[t].y = fmaf(w1, ra[t].y, w2 * r[t].y);
          r[t].z = fmaf(w1, ra[t].z, w2 * r[t].z); r[t].w = fmaf(w1, ra[t].w, w2 * r[t].w);
        }
      }
      star_epilogue<NV, MODE>(p, o + j, h, lane, scale16, r);
    }
    cur = nxt;
    ck = nk;
    nxt = __shfl_sync(0xffffffffu, ticket, 0);
    if (nxt < p.n_tasks) { nk = decode(nxt); prefetch_item(nk); }
  }
  // ---- retire: the last warp of this head's queue resets it for the next launch ----
  if (lane == 0) {
    // (this warp's last ticket has been read, i.e. performed, before this point: no fence needed to order the two counters)
    const int total = (int)(gridDim.x * (blockDim.x >> 5));
    if (atomicAdd(qn + 1, 1) == total - 1) {
      qn[0] = 0;
      qn[1] = 0;
    }
  }
}

}  // namespace tx

using namespace tx;

extern "C" {

int64_t tx_gat_star_chunk(void) { return kStarChunk; }
int64_t tx_gat_star_max_chunks(void) { return kStarMaxChunks; }

int tx_gat_star_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* tasks, int64_t n_tasks, int64_t chunk,
                    int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id,
                    float* alpha, float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                    void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out, int32_t* queue, void* stream) {
  TX_REQUIRE(epi, "gat_star_fwd: epilogue required");
  TX_REQUIRE(tx_gat_fused_supported(heads, dim, epi->mean_heads), "gat_star_fwd: unsupported shape (heads %lld dim %lld); use the general path",
             (long long)heads, (long long)dim);
  TX_REQUIRE(tasks && aligned16(tasks) && queue, "gat_star_fwd: 16-byte aligned task records and a queue are required");
  TX_REQUIRE(n_tasks >= 0 && n_tasks < (1ll << 31) && n_nodes >= 0 && n_nodes < (1ll << 31) && chunk >= 1 && chunk < (1 << 20),
             "gat_star_fwd: bad sizes");
  TX_REQUIRE(aligned16(ft) && ldf % 4 == 0 && ldo % 4 == 0 && aligned16(attn_l) && aligned16(attn_r), "gat_star_fwd: 16-byte aligned rows required");
  TX_REQUIRE((out16_hi != nullptr) != (out != nullptr), "gat_star_fwd: exactly one of out / out16_hi");
  TX_REQUIRE(!out || aligned16(out), "gat_star_fwd: out must be 16-byte aligned");
  TX_REQUIRE(!out16_hi || (!epi->mean_heads && out16_lo && bound && aligned16(out16_hi) && aligned16(out16_lo) && ld16 % 8 == 0 &&
                           ld16 >= heads * dim + epi->pos_dim), "gat_star_fwd: bad fp16 output buffers");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f && epi->p_drop >= 0.f && epi->p_drop < 1.f, "gat_star_fwd: dropout rates must be in [0,1)");
  TX_REQUIRE(alpha && elog && (p_attn == 0.f || (alpha_d && alpha_d != alpha)), "gat_star_fwd: alpha/elog/alpha_d buffers");
  const int64_t need = epi->mean_heads ? dim : heads * dim + epi->pos_dim;
  TX_REQUIRE(ldo >= need, "gat_star_fwd: ldo %lld < %lld", (long long)ldo, (long long)need);
  TX_REQUIRE(epi->pos_dim == 0 || (epi->next_pos_table && epi->pos), "gat_star_fwd: pos_dim > 0 needs next_pos_table and pos");
  if (n_tasks == 0 || n_nodes == 0) return TX_OK;
  StarFwdParams p;
  p.ft = ft; p.ldf = ldf; p.attn_l = attn_l; p.attn_r = attn_r; p.tasks = tasks; p.n_tasks = (int)n_tasks; p.chunk = (int)chunk; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.attn_inv_keep = 1.f / (1.f - p_attn); p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.out = out; p.ldo = ldo;
  p.out16_hi = (__half*)out16_hi; p.out16_lo = (__half*)out16_lo; p.ld16 = ld16; p.bound = bound; p.scale_out = scale_out;
  p.maskbits = reinterpret_cast<uint8_t*>(maskbits); p.mask_ld = (int)tx_gat_fused_mask_ld(heads, dim);
  p.hidden = epi->mean_heads ? 0 : 1; p.act_slope = epi->act_slope; p.next_pos_table = epi->next_pos_table; p.pos = epi->pos;
  p.pos_dim = (int)epi->pos_dim; p.next_inv_keep = 1.f / (1.f - epi->p_drop); p.next_thr = drop_threshold(epi->p_drop);
  p.next_seed = epi->seed; p.next_stream = epi->stream_id; p.queue = queue;
  { static int pf = -1; if (pf < 0) { const char* e = getenv("TAXO_STAR_PREFETCH"); pf = e ? atoi(e) : 0; } p.prefetch = pf; }
  const int nv = (int)((dim + 127) / 128);
  int gx = grid_for_warps(n_tasks * heads, 8, TX_STAR_MIN_BLOCKS);
  gx = (gx + (int)heads - 1) / (int)heads;
  if (gx < 1) gx = 1;
  dim3 grid(gx, (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  const bool hot = p.hidden && p.out16_hi && p.maskbits && p.act_slope != 1.f && p.next_thr != 0;
  if (hot) {
    switch (nv) {
      case 1: TX_PDL_LAUNCH((gat_star_fwd_kernel<1, 1>), grid, 256, 0, st, p); break;
      case 2: TX_PDL_LAUNCH((gat_star_fwd_kernel<2, 1>), grid, 256, 0, st, p); break;
      case 3: TX_PDL_LAUNCH((gat_star_fwd_kernel<3, 1>), grid, 256, 0, st, p); break;
      default: TX_PDL_LAUNCH((gat_star_fwd_kernel<4, 1>), grid, 256, 0, st, p); break;
    }
  } else {
    switch (nv) {
      case 1: TX_PDL_LAUNCH((gat_star_fwd_kernel<1, 0>), grid, 256, 0, st, p); break;
      case 2: TX_PDL_LAUNCH((gat_star_fwd_kernel<2, 0>), grid, 256, 0, st, p); break;
      case 3: TX_PDL_LAUNCH((gat_star_fwd_kernel<3, 0>), grid, 256, 0, st, p); break;
      default: TX_PDL_LAUNCH((gat_star_fwd_kernel<4, 0>), grid, 256, 0, st, p); break;
    }
  }
  TX_LAUNCH_CHECK("tx_gat_star_fwd");
  return TX_OK;
}

}  // extern "C"
