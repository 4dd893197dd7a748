// Star-egonet fused GAT backward (sm_100a) - first version, opt-in (TAXO_STAR_BWD=1).  Parity-green on a B200 against the staged
// backward (tests/test_gpu_parity.py::test_star_backward_matches_staged_backward, TAXO_STAR_BWD_TEST=1); measured on the MAG-CS
// benchmark: L0 (H = 4) 0.279 ms vs 0.283 ms staged, L1 (H = 1) 0.145 ms vs 0.097 ms - not yet the default: whole egonets per warp leave
// a tail at H = 1, and the epilogue / pointer arithmetic have not had the forward's treatment.
//
// The forward's recipe (tx_star_fwd.cu) applied to the backward pass: for the egonets of data_loader/dataset.py:404-437 the autograd
// of model_zoo.py:84-96,106-114 has a closed form per egonet (restated and checked against torch autograd in
// oracle/star_backward.py), so neither CSR is read and no tile has to be staged:
//   * grand-parent k (one in-edge, its self loop): alpha = 1, the softmax backward vanishes:
//       d(ft_k) = alpha~_self g_k + alpha~(k->anchor) g_anchor + ds(k->anchor) attn_l
//   * sibling s (in-edges {anchor, self}): closed-form 2x2 softmax backward from <g_s, ft_anchor> and <g_s, ft_s>:
//       d(ft_s) = alpha~_self g_s + ds_self attn_l + (ds_anchor + ds_self) attn_r;   anchor += alpha~(anchor->s) g_s, da1 += ds_anchor
//   * anchor (in-edges {gp_0.., self}): dots <g_anchor, ft_gp_k>, <g_anchor, ft_anchor>, softmax backward over a + 1 logits.
// A warp owns whole egonets; the anchor's g / ft rows and its accumulator stay in registers while the other rows stream by (each
// row's g and ft are read exactly once, plus one re-read of the grand-parents' ft rows through L1).  Ownership is STATIC - warp W of a
// head owns the egonets whose first row lies in [W N / nW, (W + 1) N / nW) - so the per-warp d(attn_l) / d(attn_r) sums
// (shared-memory accumulators, reduced per CTA in a fixed order) are run-to-run deterministic like everything else in the library.
// Same outputs as tx_gat_fused_bwd_staged: dft (fp32 or fp16 hi/lo pair) and dattn_partial [gridDim.x, 2, H, D].
#include <math.h>
#include <stdlib.h>

#include "tx_common.cuh"

namespace tx {

struct StarBwdParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const float* ft; int64_t ldf;
  const float* alpha; const float* alpha_d; const float* elog;
  const float* attn_l; const float* attn_r;
  const int32_t* n_gp; const int32_t* n_sib; const int32_t* node_off; const int32_t* edge_off;
  int n_graphs; int n; int H; int D;
  float neg_slope; float attn_inv_keep; uint32_t attn_thr; uint64_t attn_seed; uint32_t attn_stream;
  float* ds;                     // scratch [E * H]: only anchors with more than 31 grand-parents spill their dots / d(logit) here
  float* dft; int64_t ldd;
  __half* dft16_hi; __half* dft16_lo; int64_t ld16; const float* bound; float* scale_out;
  float* dattn_partial;          // [gridDim.x, 2, H, D]
};

constexpr int kStarBwdWarps = 8;

template <int NV>
__device__ __forceinline__ void sb_load_row(const float* __restrict__ p, int lane, int D, float4 (&v)[NV]) {
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c = (lane + 32 * t) * 4;
    v[t] = (t < NV - 1 || c < D) ? __ldg(reinterpret_cast<const float4*>(p + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NV>
__device__ __forceinline__ float sb_dot(const float4 (&a)[NV], const float4 (&b)[NV]) {
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    acc = fmaf(a[t].x, b[t].x, acc); acc = fmaf(a[t].y, b[t].y, acc); acc = fmaf(a[t].z, b[t].z, acc); acc = fmaf(a[t].w, b[t].w, acc);
  }
  return acc;
}
__device__ __forceinline__ void sb_warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a += ta; b += tb;
  }
}

template <int NV>
__global__ void __launch_bounds__(kStarBwdWarps * 32, 2) gat_star_bwd_kernel(const StarBwdParams p) {
  __shared__ float4 s_l[NV * 32];
  __shared__ float4 s_r[NV * 32];
  __shared__ float4 s_acc[kStarBwdWarps][2][NV * 32];      // per warp: d(attn_l), d(attn_r) of this head
  const int h = blockIdx.y;
  const int H = p.H, D = p.D;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    s_acc[wid][0][lane + 32 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
    s_acc[wid][1][lane + 32 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const bool attn_drop = p.attn_thr != 0;
  const float gs = p.g_scale;
  const float* gbase = p.g + (int64_t)h * p.g_head_stride;
  const float* fbase = p.ft + (int64_t)h * D;
  const float scale16 = p.dft16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.dft16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;

  auto keepw = [&](int eid) -> float {
    if (!attn_drop) return 1.f;
    return drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)eid * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
  };
  auto dslope = [&](float e) -> float { return e > 0.f ? 1.f : p.neg_slope; };
  // first egonet whose first row is >= x (node_off is strictly increasing: every egonet has its anchor)
  auto first_egonet = [&](int64_t x) -> int {
    int lo = 0, hi = p.n_graphs;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(p.node_off + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const int64_t nW = (int64_t)gridDim.x * kStarBwdWarps, W = (int64_t)blockIdx.x * kStarBwdWarps + wid;
  const int eg_beg = first_egonet(W * p.n / nW), eg_end = first_egonet((W + 1) * p.n / nW);

  for (int eg = eg_beg; eg < eg_end; ++eg) {
    const int a = __ldg(p.n_gp + eg), s = __ldg(p.n_sib + eg), o = __ldg(p.node_off + eg), q = __ldg(p.edge_off + eg);
    const int n = a + 1 + s, self0 = q + a + s, A = o + a, deg = a + 1;
    float4 gA[NV], fA[NV], accA[NV];
    sb_load_row<NV>(gbase + (int64_t)A * p.ldg, lane, D, gA);
    sb_load_row<NV>(fbase + (int64_t)A * p.ldf, lane, D, fA);
    // ---- anchor: d(alpha~) of its in-edges {gp_0 .. gp_{a-1}, self} and their weighted sum ----
    float dd_mine = 0.f, al_mine = 0.f, el_mine = 1.f, tsum = 0.f;
    for (int k = 0; k <= a; ++k) {
      float4 rf[NV];
      if (k < a) {
        sb_load_row<NV>(fbase + (int64_t)(o + k) * p.ldf, lane, D, rf);
      } else {
#pragma unroll
        for (int t = 0; t < NV; ++t) rf[t] = fA[t];
      }
      const int64_t so = (int64_t)(q + a + k) * H + h;
      const float d = warp_sum(sb_dot<NV>(gA, rf)) * gs * keepw(k < a ? q + k : self0 + a);
      const float alk = __ldg(p.alpha + so);
      tsum = fmaf(alk, d, tsum);
      if (deg <= 32) {
        if (lane == k) { dd_mine = d; al_mine = alk; el_mine = __ldg(p.elog + so); }
      } else if (lane == 0) {
        p.ds[so] = d;
      }
    }
    float ds_mine = 0.f, da2A = 0.f, ds_aa;
    if (deg <= 32) {
      if (lane < deg) ds_mine = al_mine * (dd_mine - tsum) * dslope(el_mine);
      da2A = warp_sum(ds_mine);
      ds_aa = __shfl_sync(0xffffffffu, ds_mine, a);
    } else {
      __syncwarp();
      for (int k = lane; k < deg; k += 32) {
        const int64_t so = (int64_t)(q + a + k) * H + h;
        const float dsv = __ldg(p.alpha + so) * (p.ds[so] - tsum) * dslope(__ldg(p.elog + so));
        p.ds[so] = dsv;
        da2A += dsv;
      }
      da2A = warp_sum(da2A);
      __syncwarp();
      ds_aa = p.ds[(int64_t)(q + 2 * a) * H + h];
    }
    float da1A = ds_aa;                                   // + sum over siblings of ds(anchor -> sibling), collected below
    {
      const float w = __ldg(p.alpha_d + (int64_t)(q + 2 * a) * H + h) * gs;          // alpha~ of the anchor's self loop
#pragma unroll
      for (int t = 0; t < NV; ++t) accA[t] = make_float4(w * gA[t].x, w * gA[t].y, w * gA[t].z, w * gA[t].w);
    }
    // ---- every row once: grand-parents, siblings, the anchor LAST (it needs the siblings' contributions); ONE loop body ----
    for (int jj = 0; jj < n; ++jj) {
      const int j = jj < a ? jj : (jj < n - 1 ? jj + 1 : a);       // local row: 0..a-1 gp, a anchor, a+1.. siblings
      float4 rg[NV], rf[NV];
      float c1, c2;                                                // da1_j, da2_j: coefficients of attn_l / attn_r and of ft_j in d(attn)
      if (j == a) {
#pragma unroll
        for (int t = 0; t < NV; ++t) { rg[t] = accA[t]; rf[t] = fA[t]; }
        c1 = da1A; c2 = da2A;
      } else {
        sb_load_row<NV>(gbase + (int64_t)(o + j) * p.ldg, lane, D, rg);
        sb_load_row<NV>(fbase + (int64_t)(o + j) * p.ldf, lane, D, rf);
        if (j < a) {
          // grand-parent: alpha~_self g_j + alpha~(j -> anchor) g_anchor; its own softmax backward vanishes (single in-edge)
          const float dsk = deg <= 32 ? __shfl_sync(0xffffffffu, ds_mine, j) : p.ds[(int64_t)(q + a + j) * H + h];
          const float w_self = __ldg(p.alpha_d + (int64_t)(q + j) * H + h) * gs;
          const float w_anch = __ldg(p.alpha_d + (int64_t)(q + a + j) * H + h) * gs;
#pragma unroll
          for (int t = 0; t < NV; ++t) {
            rg[t].x = fmaf(w_self, rg[t].x, w_anch * gA[t].x); rg[t].y = fmaf(w_self, rg[t].y, w_anch * gA[t].y);
            rg[t].z = fmaf(w_self, rg[t].z, w_anch * gA[t].z); rg[t].w = fmaf(w_self, rg[t].w, w_anch * gA[t].w);
          }
          c1 = dsk; c2 = 0.f;
        } else {
          // sibling: in-edges {anchor -> j (slot q + 2j - 1, edge id q + j - 1), self (slot q + 2j, edge id self0 + j)}
          const int64_t s1 = (int64_t)(q + 2 * j - 1) * H + h, s2 = s1 + H;
          float d1 = sb_dot<NV>(rg, fA), d2 = sb_dot<NV>(rg, rf);
          sb_warp_sum2(d1, d2);
          d1 *= gs * keepw(q + j - 1);
          d2 *= gs * keepw(self0 + j);
          const float al1 = __ldg(p.alpha + s1), al2 = __ldg(p.alpha + s2);
          const float ts = fmaf(al1, d1, al2 * d2);
          const float ds1 = al1 * (d1 - ts) * dslope(__ldg(p.elog + s1));
          const float ds2 = al2 * (d2 - ts) * dslope(__ldg(p.elog + s2));
          da1A += ds1;
          const float w1 = __ldg(p.alpha_d + s1) * gs, w2 = __ldg(p.alpha_d + s2) * gs;
#pragma unroll
          for (int t = 0; t < NV; ++t) {
            accA[t].x = fmaf(w1, rg[t].x, accA[t].x); accA[t].y = fmaf(w1, rg[t].y, accA[t].y);
            accA[t].z = fmaf(w1, rg[t].z, accA[t].z); accA[t].w = fmaf(w1, rg[t].w, accA[t].w);
            rg[t].x *= w2; rg[t].y *= w2; rg[t].z *= w2; rg[t].w *= w2;
          }
          c1 = ds2; c2 = ds1 + ds2;
        }
      }
      // common tail: d(ft_j) = rg + c1 attn_l + c2 attn_r;  d(attn_l) += c1 ft_j;  d(attn_r) += c2 ft_j
      const int64_t row = o + j;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        const float4 l = s_l[c4], r = s_r[c4];
        float4 v;
        v.x = fmaf(c1, l.x, fmaf(c2, r.x, rg[t].x)); v.y = fmaf(c1, l.y, fmaf(c2, r.y, rg[t].y));
        v.z = fmaf(c1, l.z, fmaf(c2, r.z, rg[t].z)); v.w = fmaf(c1, l.w, fmaf(c2, r.w, rg[t].w));
        float4 hl = s_acc[wid][0][c4], hr = s_acc[wid][1][c4];
        hl.x = fmaf(c1, rf[t].x, hl.x); hl.y = fmaf(c1, rf[t].y, hl.y); hl.z = fmaf(c1, rf[t].z, hl.z); hl.w = fmaf(c1, rf[t].w, hl.w);
        hr.x = fmaf(c2, rf[t].x, hr.x); hr.y = fmaf(c2, rf[t].y, hr.y); hr.z = fmaf(c2, rf[t].z, hr.z); hr.w = fmaf(c2, rf[t].w, hr.w);
        s_acc[wid][0][c4] = hl;
        s_acc[wid][1][c4] = hr;
        if (t < NV - 1 || c4 * 4 < D) {
          if (p.dft16_hi) {
            uint2 h16, l16;
            f16_split4(v, scale16, h16, l16);
            const int64_t o16 = row * p.ld16 + (int64_t)h * D + c4 * 4;
            *reinterpret_cast<uint2*>(p.dft16_hi + o16) = h16;
            *reinterpret_cast<uint2*>(p.dft16_lo + o16) = l16;
          } else {
            *reinterpret_cast<float4*>(p.dft + row * p.ldd + (int64_t)h * D + c4 * 4) = v;
          }
        }
      }
    }
  }
  // ---- d(attn) partials: the warps' accumulators summed in warp order (fixed) ----
  __syncthreads();
  const int D4 = D >> 2;
  for (int t = threadIdx.x; t < 2 * D4; t += blockDim.x) {
    const int lr = t / D4, c4 = t - lr * D4;
    float4 sum = s_acc[0][lr][c4];
#pragma unroll
    for (int w = 1; w < kStarBwdWarps; ++w) {
      const float4 x = s_acc[w][lr][c4];
      sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
    }
    *reinterpret_cast<float4*>(p.dattn_partial + (((int64_t)blockIdx.x * 2 + lr) * H + h) * D + c4 * 4) = sum;
  }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Team variant (TAXO_STAR_BWD_COOP=<rows>, opt-in, NOT yet run on hardware): same arithmetic, different dealing.  A CTA owns the
// egonets whose first row lies in its slice of the rows; SMALL egonets (fewer than `coop_rows` rows) go to its warps round-robin, a
// LARGE egonet is processed by all 8 warps together: every warp loads the anchor pair, takes the siblings k = warp (mod 8), and the
// warps' anchor accumulators are added through one shared-memory row in warp order (fixed -> deterministic) before warp 0 writes the
// anchor row.  This removes the tail of the whole-egonet-per-warp dealing (a 55-row egonet is 3.5 x the average warp load at H = 1)
// without per-chunk scratch rows or a fix-up pass.  One loop body serves both team sizes.
// ---------------------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kStarBwdWarps * 32, 2) gat_star_bwd_team_kernel(const StarBwdParams p, const int coop_rows) {
  __shared__ float4 s_l[NV * 32];
  __shared__ float4 s_r[NV * 32];
  __shared__ float4 s_acc[kStarBwdWarps][2][NV * 32];      // per warp: d(attn_l), d(attn_r) of this head
  __shared__ float4 s_red[NV * 32];                        // anchor accumulator of a team, added warp by warp
  __shared__ float s_da1;
  const int h = blockIdx.y;
  const int H = p.H, D = p.D;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    s_acc[wid][0][lane + 32 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
    s_acc[wid][1][lane + 32 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const bool attn_drop = p.attn_thr != 0;
  const float gs = p.g_scale;
  const float* gbase = p.g + (int64_t)h * p.g_head_stride;
  const float* fbase = p.ft + (int64_t)h * D;
  const float scale16 = p.dft16_hi ? f16_split_scale(__ldg(p.bound)) : 1.f;
  if (p.dft16_hi && p.scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.scale_out = scale16;

  auto keepw = [&](int eid) -> float {
    if (!attn_drop) return 1.f;
    return drop_keep1(p.attn_seed, p.attn_stream, (uint64_t)((int64_t)eid * H + h), p.attn_thr) ? p.attn_inv_keep : 0.f;
  };
  auto dslope = [&](float e) -> float { return e > 0.f ? 1.f : p.neg_slope; };
  auto first_egonet = [&](int64_t x) -> int {
    int lo = 0, hi = p.n_graphs;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(p.node_off + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  // the CTA's egonets (identical in every warp: the loop below is CTA-uniform, so the barriers of the large egonets are safe)
  const int eg_beg = first_egonet((int64_t)blockIdx.x * p.n / gridDim.x), eg_end = first_egonet((int64_t)(blockIdx.x + 1) * p.n / gridDim.x);
  int n_small = 0;

  for (int eg = eg_beg; eg < eg_end; ++eg) {
    const int a = __ldg(p.n_gp + eg), s = __ldg(p.n_sib + eg), o = __ldg(p.node_off + eg), q = __ldg(p.edge_off + eg);
    const int n = a + 1 + s, self0 = q + a + s, A = o + a, deg = a + 1;
    const bool big = n >= coop_rows;                     // CTA-uniform
    int team = 1, rank = 0;
    if (big) { team = kStarBwdWarps; rank = wid; }
    else if ((n_small++ % kStarBwdWarps) != wid) continue;     // a small egonet belongs to one warp; the others move on (no barrier here)

    float4 gA[NV], fA[NV], accA[NV];
    sb_load_row<NV>(gbase + (int64_t)A * p.ldg, lane, D, gA);
    sb_load_row<NV>(fbase + (int64_t)A * p.ldf, lane, D, fA);
    float ds_mine = 0.f, da2A = 0.f, da1A = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) accA[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rank == 0) {
      // ---- anchor: d(alpha~) of its in-edges {gp_0 .. gp_{a-1}, self}, softmax backward (as in gat_star_bwd_kernel) ----
      float dd_mine = 0.f, al_mine = 0.f, el_mine = 1.f, tsum = 0.f;
      for (int k = 0; k <= a; ++k) {
        float4 rf[NV];
        if (k < a) {
          sb_load_row<NV>(fbase + (int64_t)(o + k) * p.ldf, lane, D, rf);
        } else {
#pragma unroll
          for (int t = 0; t < NV; ++t) rf[t] = fA[t];
        }
        const int64_t so = (int64_t)(q + a + k) * H + h;
        const float d = warp_sum(sb_dot<NV>(gA, rf)) * gs * keepw(k < a ? q + k : self0 + a);
        const float alk = __ldg(p.alpha + so);
        tsum = fmaf(alk, d, tsum);
        if (deg <= 32) {
          if (lane == k) { dd_mine = d; al_mine = alk; el_mine = __ldg(p.elog + so); }
        } else if (lane == 0) {
          p.ds[so] = d;
        }
      }
      float ds_aa;
      if (deg <= 32) {
        if (lane < deg) ds_mine = al_mine * (dd_mine - tsum) * dslope(el_mine);
        da2A = warp_sum(ds_mine);
        ds_aa = __shfl_sync(0xffffffffu, ds_mine, a);
      } else {
        __syncwarp();
        for (int k = lane; k < deg; k += 32) {
          const int64_t so = (int64_t)(q + a + k) * H + h;
          const float dsv = __ldg(p.alpha + so) * (p.ds[so] - tsum) * dslope(__ldg(p.elog + so));
          p.ds[so] = dsv;
          da2A += dsv;
        }
        da2A = warp_sum(da2A);
        __syncwarp();
        ds_aa = p.ds[(int64_t)(q + 2 * a) * H + h];
      }
      da1A = ds_aa;
      const float w = __ldg(p.alpha_d + (int64_t)(q + 2 * a) * H + h) * gs;            // alpha~ of the anchor's self loop
#pragma unroll
      for (int t = 0; t < NV; ++t) accA[t] = make_float4(w * gA[t].x, w * gA[t].y, w * gA[t].z, w * gA[t].w);
    }
    // rows of this warp: rank 0 takes the grand-parents, every rank the siblings k = rank (mod team); the anchor row is a second
    // pass of the SAME loop body (rank 0 only), after the team's anchor accumulators have been combined
    const int cnt_gp = rank == 0 ? a : 0;
    const int cnt_sib = s > rank ? (s - rank + team - 1) / team : 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int it_beg = pass == 0 ? 0 : cnt_gp + cnt_sib;
      const int it_end = pass == 0 ? cnt_gp + cnt_sib : (rank == 0 ? cnt_gp + cnt_sib + 1 : cnt_gp + cnt_sib);
      for (int it = it_beg; it < it_end; ++it) {
        const int j = it < cnt_gp ? it : (it < cnt_gp + cnt_sib ? a + 1 + rank + (it - cnt_gp) * team : a);
        float4 rg[NV], rf[NV];
        float c1, c2;
        if (j == a) {
#pragma unroll
          for (int t = 0; t < NV; ++t) { rg[t] = accA[t]; rf[t] = fA[t]; }
          c1 = da1A; c2 = da2A;
        } else {
          sb_load_row<NV>(gbase + (int64_t)(o + j) * p.ldg, lane, D, rg);
          sb_load_row<NV>(fbase + (int64_t)(o + j) * p.ldf, lane, D, rf);
          if (j < a) {
            const float dsk = deg <= 32 ? __shfl_sync(0xffffffffu, ds_mine, j) : p.ds[(int64_t)(q + a + j) * H + h];
            const float w_self = __ldg(p.alpha_d + (int64_t)(q + j) * H + h) * gs;
            const float w_anch = __ldg(p.alpha_d + (int64_t)(q + a + j) * H + h) * gs;
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              rg[t].x = fmaf(w_self, rg[t].x, w_anch * gA[t].x); rg[t].y = fmaf(w_self, rg[t].y, w_anch * gA[t].y);
              rg[t].z = fmaf(w_self, rg[t].z, w_anch * gA[t].z); rg[t].w = fmaf(w_self, rg[t].w, w_anch * gA[t].w);
            }
            c1 = dsk; c2 = 0.f;
          } else {
            const int64_t s1 = (int64_t)(q + 2 * j - 1) * H + h, s2 = s1 + H;
            float d1 = sb_dot<NV>(rg, fA), d2 = sb_dot<NV>(rg, rf);
            sb_warp_sum2(d1, d2);
            d1 *= gs * keepw(q + j - 1);
            d2 *= gs * keepw(self0 + j);
            const float al1 = __ldg(p.alpha + s1), al2 = __ldg(p.alpha + s2);
            const float ts = fmaf(al1, d1, al2 * d2);
            const float ds1 = al1 * (d1 - ts) * dslope(__ldg(p.elog + s1));
            const float ds2 = al2 * (d2 - ts) * dslope(__ldg(p.elog + s2));
            da1A += ds1;
            const float w1 = __ldg(p.alpha_d + s1) * gs, w2 = __ldg(p.alpha_d + s2) * gs;
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              accA[t].x = fmaf(w1, rg[t].x, accA[t].x); accA[t].y = fmaf(w1, rg[t].y, accA[t].y);
              accA[t].z = fmaf(w1, rg[t].z, accA[t].z); accA[t].w = fmaf(w1, rg[t].w, accA[t].w);
              rg[t].x *= w2; rg[t].y *= w2; rg[t].z *= w2; rg[t].w *= w2;
            }
            c1 = ds2; c2 = ds1 + ds2;
          }
        }
        const int64_t row = o + j;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int c4 = lane + 32 * t;
          const float4 l = s_l[c4], r = s_r[c4];
          float4 v;
          v.x = fmaf(c1, l.x, fmaf(c2, r.x, rg[t].x)); v.y = fmaf(c1, l.y, fmaf(c2, r.y, rg[t].y));
          v.z = fmaf(c1, l.z, fmaf(c2, r.z, rg[t].z)); v.w = fmaf(c1, l.w, fmaf(c2, r.w, rg[t].w));
          float4 hl = s_acc[wid][0][c4], hr = s_acc[wid][1][c4];
          hl.x = fmaf(c1, rf[t].x, hl.x); hl.y = fmaf(c1, rf[t].y, hl.y); hl.z = fmaf(c1, rf[t].z, hl.z); hl.w = fmaf(c1, rf[t].w, hl.w);
          hr.x = fmaf(c2, rf[t].x, hr.x); hr.y = fmaf(c2, rf[t].y, hr.y); hr.z = fmaf(c2, rf[t].z, hr.z); hr.w = fmaf(c2, rf[t].w, hr.w);
          s_acc[wid][0][c4] = hl;
          s_acc[wid][1][c4] = hr;
          if (t < NV - 1 || c4 * 4 < D) {
            if (p.dft16_hi) {
              uint2 h16, l16;
              f16_split4(v, scale16, h16, l16);
              const int64_t o16 = row * p.ld16 + (int64_t)h * D + c4 * 4;
              *reinterpret_cast<uint2*>(p.dft16_hi + o16) = h16;
              *reinterpret_cast<uint2*>(p.dft16_lo + o16) = l16;
            } else {
              *reinterpret_cast<float4*>(p.dft + row * p.ldd + (int64_t)h * D + c4 * 4) = v;
            }
          }
        }
      }
      if (pass == 0 && big) {
        // combine the team: accA and da1A of the 8 warps, added in warp order through one shared row (CTA-uniform branch)
        for (int w = 0; w < kStarBwdWarps; ++w) {
          if (wid == w) {
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              const int c4 = lane + 32 * t;
              float4 x = accA[t];
              if (w > 0) { const float4 y = s_red[c4]; x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
              s_red[c4] = x;
            }
            if (lane == 0) s_da1 = w > 0 ? s_da1 + da1A : da1A;
          }
          __syncthreads();
        }
        if (rank == 0) {
#pragma unroll
          for (int t = 0; t < NV; ++t) accA[t] = s_red[lane + 32 * t];
          da1A = s_da1;
        }
        __syncthreads();                                  // s_red / s_da1 are free again before the next large egonet
      }
    }
  }
  // ---- d(attn) partials: the warps' accumulators summed in warp order (fixed) ----
  __syncthreads();
  const int D4 = D >> 2;
  for (int t = threadIdx.x; t < 2 * D4; t += blockDim.x) {
    const int lr = t / D4, c4 = t - lr * D4;
    float4 sum = s_acc[0][lr][c4];
#pragma unroll
    for (int w = 1; w < kStarBwdWarps; ++w) {
      const float4 x = s_acc[w][lr][c4];
      sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
    }
    *reinterpret_cast<float4*>(p.dattn_partial + (((int64_t)blockIdx.x * 2 + lr) * H + h) * D + c4 * 4) = sum;
  }
}

}  // namespace tx

using namespace tx;

extern "C" {

int64_t tx_gat_star_bwd_blocks(int64_t n_nodes, int64_t heads) {
  if (heads < 1) heads = 1;
  int64_t gx = (2 * (int64_t)kNumSms + heads - 1) / heads;          // two CTAs per SM over all heads
  const int64_t need = (n_nodes + kStarBwdWarps - 1) / kStarBwdWarps;
  if (gx > need) gx = need;
  return gx < 1 ? 1 : gx;
}

int tx_gat_star_bwd(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft, int64_t ldf,
                    const float* alpha, const float* alpha_d, const float* elog, const float* attn_l, const float* attn_r,
                    const int32_t* n_gp, const int32_t* n_sib, const int32_t* node_off, const int32_t* edge_off, int64_t n_graphs,
                    int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed,
                    uint32_t attn_stream_id, float* ds, float* dft, int64_t ldd, void* dft16_hi, void* dft16_lo, int64_t ld16,
                    const float* bound, float* scale_out, float* dattn_partial, void* stream) {
  TX_REQUIRE(g_head_stride != 0 || heads == 1, "gat_star_bwd: a shared g row (head mean) needs heads == 1");
  TX_REQUIRE(!dft16_hi || (dft16_lo && bound && aligned16(dft16_hi) && aligned16(dft16_lo) && ld16 % 8 == 0 && ld16 >= heads * dim),
             "gat_star_bwd: bad fp16 output buffers");
  TX_REQUIRE(dft16_hi || (dft && aligned16(dft) && ldd % 4 == 0 && ldd >= heads * dim), "gat_star_bwd: an output buffer is required");
  TX_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 512, "gat_star_bwd: dim must be a multiple of 4 and <= 512");
  TX_REQUIRE(aligned16(g) && ldg % 4 == 0 && g_head_stride % 4 == 0 && aligned16(ft) && ldf % 4 == 0 && aligned16(attn_l) &&
             aligned16(attn_r) && aligned16(dattn_partial), "gat_star_bwd: 16-byte aligned rows required");
  TX_REQUIRE(p_attn >= 0.f && p_attn < 1.f, "gat_star_bwd: dropout rate must be in [0,1)");
  TX_REQUIRE(n_gp && n_sib && node_off && edge_off && alpha && elog && ds && dattn_partial, "gat_star_bwd: null pointer");
  TX_REQUIRE(n_graphs >= 0 && n_graphs < (1ll << 31) && n_nodes >= 0 && n_nodes < (1ll << 31), "gat_star_bwd: bad sizes");
  if (n_nodes == 0 || n_graphs == 0) return TX_OK;
  StarBwdParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.ft = ft; p.ldf = ldf;
  p.alpha = alpha; p.alpha_d = alpha_d ? alpha_d : alpha; p.elog = elog; p.attn_l = attn_l; p.attn_r = attn_r;
  p.n_gp = n_gp; p.n_sib = n_sib; p.node_off = node_off; p.edge_off = edge_off; p.n_graphs = (int)n_graphs; p.n = (int)n_nodes;
  p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope; p.attn_inv_keep = 1.f / (1.f - p_attn);
  p.attn_thr = drop_threshold(p_attn); p.attn_seed = attn_seed; p.attn_stream = attn_stream_id; p.ds = ds;
  p.dft = dft; p.ldd = ldd; p.dft16_hi = (__half*)dft16_hi; p.dft16_lo = (__half*)dft16_lo; p.ld16 = ld16; p.bound = bound;
  p.scale_out = scale_out; p.dattn_partial = dattn_partial;
  const int nv = (int)((dim + 127) / 128);
  dim3 grid((unsigned)tx_gat_star_bwd_blocks(n_nodes, heads), (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  static int coop = -1;                                   // TAXO_STAR_BWD_COOP=<rows>: team variant for egonets with >= rows nodes
  if (coop < 0) { const char* e = getenv("TAXO_STAR_BWD_COOP"); coop = e ? atoi(e) : 0; if (coop < 0) coop = 0; }
  if (coop > 0) {
    switch (nv) {
      case 1: gat_star_bwd_team_kernel<1><<<grid, kStarBwdWarps * 32, 0, st>>>(p, coop); break;
      case 2: gat_star_bwd_team_kernel<2><<<grid, kStarBwdWarps * 32, 0, st>>>(p, coop); break;
      case 3: gat_star_bwd_team_kernel<3><<<grid, kStarBwdWarps * 32, 0, st>>>(p, coop); break;
      default: gat_star_bwd_team_kernel<4><<<grid, kStarBwdWarps * 32, 0, st>>>(p, coop); break;
    }
    TX_LAUNCH_CHECK("tx_gat_star_bwd (team)");
    return TX_OK;
  }
  switch (nv) {
    case 1: gat_star_bwd_kernel<1><<<grid, kStarBwdWarps * 32, 0, st>>>(p); break;
    case 2: gat_star_bwd_kernel<2><<<grid, kStarBwdWarps * 32, 0, st>>>(p); break;
    case 3: gat_star_bwd_kernel<3><<<grid, kStarBwdWarps * 32, 0, st>>>(p); break;
    default: gat_star_bwd_kernel<4><<<grid, kStarBwdWarps * 32, 0, st>>>(p); break;
  }
  TX_LAUNCH_CHECK("tx_gat_star_bwd");
  return TX_OK;
}

}  // extern "C"
