// Star-egonet fused GAT backward (sm_100a), second generation: the default backward of the hot path for EgonetBatch structures.
//
// Arithmetic: the autograd of reference model_zoo.py:84-96,106-114 in the closed form the egonets of data_loader/dataset.py:404-437
// allow (restated and checked against torch autograd in oracle/star_backward.py):
//   * grand-parent k (one in-edge, its self loop: alpha = 1, its softmax backward vanishes):
//       d(ft_k) = alpha~_self g_k + alpha~(k->anchor) g_anchor + ds(k->anchor) attn_l
//   * sibling s (in-edges {anchor, self}): 2x2 softmax backward from <g_s, ft_anchor> and <g_s, ft_s>:
//       d(ft_s) = alpha~_self g_s + ds_self attn_l + (ds_anchor + ds_self) attn_r;   anchor += alpha~(anchor->s) g_s, da1_anchor += ds_anchor
//   * anchor (in-edges {gp_0.., self}): dots <g_anchor, ft_gp_k>, <g_anchor, ft_anchor>, softmax backward over a + 1 logits.
//   With alpha~ = alpha * keep / (1 - p) saved by the forward, ds = (alpha~ d - alpha sum_k alpha~_k d_k) * lrelu'(e): no dropout hash here.
//
// How it runs (what the first version - whole egonets per warp, rows through LDG, d(attn) in shared memory - lacked):
//   * WARP-PRIVATE TMA RING: every warp owns 3 shared-memory slots {64-byte header, g row, ft row}.  Lane 0 issues the bulk copies
//     (cp.async.bulk + mbarrier complete_tx) of the row(s) TWO steps ahead of the one being processed, six lanes add the step's
//     attention scalars (alpha, alpha~, logits) with 4-byte cp.async on the same mbarrier, so every byte a step needs is in flight long
//     before it is touched and the rows cost no registers while they fly.  The first version had 16 warps x one row pair in flight
//     only while the warp was not computing: 38 % of the HBM rate, long-scoreboard 4.8 per issue.
//   * WORK ITEMS = (egonet, chunk of C siblings) from the forward's task table, pulled from a self-resetting atomic queue (tickets two
//     items ahead).  An egonet with more than C siblings is shared by several warps: every chunk writes its part of the anchor's
//     accumulator to a scratch row, bumps the egonet's counter, and the LAST chunk to arrive adds the parts IN CHUNK ORDER and writes
//     the anchor's row - deterministic whoever finishes last, no second launch, no whole-egonet tail (an egonet of 57 rows was 3.5 x
//     the average warp load of the output layer).
//   * d(attn_l), d(attn_r) leave the hot loop: the kernel only emits da1_j, da2_j per (row, head).  With the fp16-pair output they go
//     into 2 H extra columns of the d(ft) operand, so the weight-gradient GEMM dW = d(ft)^T z that follows returns
//     v_h = z^T da_h as 2 H extra rows for free, and d(attn_l)[h] = W_h v_h (tx_attn_grad_from_v) because ft_h = z W_h^T.
//   * fp16-pair scale: the rigorous bound of |d ft| (tx_bound_dft) is ~2^13 above the true maximum, which costs small rows their
//     `lo` bits.  The kernel therefore runs with an OPTIMISTIC scale (a small multiple of max|g|) and records any value that leaves the
//     fp16 range in a device flag; a second launch with the rigorous bound exits at once unless the flag is set.  Never saturates
//     silently, and the common case keeps 22 significant bits for every row within 2^13 of the largest.
// Every output element is produced by exactly one step in a fixed order: results are run-to-run deterministic.
#include <math.h>
#include <stdlib.h>

#define TX_PDL_GROUP 3
#include "tx_common.cuh"

namespace tx {

constexpr int kSbWarps = 16;           // one CTA per SM
constexpr int kSbStages = 3;
constexpr int kSbHdrBytes = 64;
enum : int { kOpEnd = 0, kOpAnchor0 = 1, kOpAnchorC = 2, kOpGpDot = 3, kOpGpOut = 4, kOpSib = 5 };

struct StarBwdParams {
  const float* g; int64_t ldg; int64_t g_head_stride; float g_scale;
  const float* ft; int64_t ldf;
  const float* alpha; const float* alpha_d; const float* elog;
  const float* attn_l; const float* attn_r;
  const int32_t* tasks; int n_tasks; int chunk;
  int H; int D;
  float neg_slope;
  float* ds;                     // scratch [E * H]: only anchors with more than 31 grand-parents spill their dots / d(logit) here
  float* da1; float* da2;        // optional fp32 [N * H]: coefficients of attn_l / attn_r per (row, head)
  float* dft; int64_t ldd;
  __half* dft16_hi; __half* dft16_lo; int64_t ld16; int tail16;   // tail16 = ld16 - H * D columns: [da1 x H | da2 x H | zeros]
  const float* bounds;           // [0] rigorous bound of |dft|, [1] optimistic bound, [2] c (da is stored as da * c), device scalars
  int* flag;                     // != 0: a value left the fp16 range under the optimistic scale
  int* reruns;                   // optional statistics: number of second launches that had to do the work
  float* scale_out;
  float* partial;                // [n_tasks, H, NV * 128 + 4]: anchor parts of egonets with several chunks
  int* counters;                 // [n_tasks * H], zero before the first launch, self-resetting
  int* queue;                    // [32 * H]: per head {next item, warps retired}; zero before the first launch, self-resetting
  int pass;                      // 0: optimistic scale; 1: rigorous scale, runs only if *flag
};

__device__ __forceinline__ uint32_t sb_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sb_bar_arrive_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_bar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void sb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void sb_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// the executing thread's earlier cp.async copies arrive on `bar` when they have landed (counted in the barrier's expected arrivals)
__device__ __forceinline__ void sb_cp_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

struct SbItem { int o, q, a, s, c, ticket; };

template <int NV, bool F16OUT>
__global__ void __launch_bounds__(kSbWarps * 32, 1) gat_star_bwd_kernel(const StarBwdParams p) {
  TX_PDL_ENTER();
  constexpr int kRowB = NV * 512;                          // one staged row (zero-padded columns are never copied)
  constexpr int kSlotB = kSbHdrBytes + 2 * kRowB;
  constexpr int kPartLd = NV * 128 + 4;                    // floats per partial row: accumulator | da1 part, da2 (chunk 0)
  if (p.pass == 1 && *reinterpret_cast<volatile int*>(p.flag) == 0) return;      // the optimistic launch was exact: nothing to redo

  extern __shared__ __align__(128) unsigned char sb_smem[];
  float4* s_l = reinterpret_cast<float4*>(sb_smem);
  float4* s_r = s_l + NV * 32;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sb_smem + 2 * kRowB);
  unsigned char* s_slots = sb_smem + 2 * kRowB + 512;      // 16 warps x 3 barriers x 8 B = 384 B, padded to 512

  const int h = blockIdx.y;
  const int H = p.H, D = p.D;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < NV * 32; t += blockDim.x) {
    const int c = t * 4;
    s_l[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_l + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    s_r[t] = c < D ? __ldg(reinterpret_cast<const float4*>(p.attn_r + (int64_t)h * D + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t bar0 = sb_saddr(s_bar + wid * kSbStages);
  unsigned char* my_slots = s_slots + (size_t)wid * kSbStages * kSlotB;
  const uint32_t slot0 = sb_saddr(my_slots);
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kSbStages; ++st) sb_bar_init(bar0 + 8 * st, 33);      // lane 0's arrive.expect_tx + 32 cp.async arrivals
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const float gs = p.g_scale;
  const float* gbase = p.g + (int64_t)h * p.g_head_stride;
  const float* fbase = p.ft + (int64_t)h * D;
  const uint32_t rowB = (uint32_t)D * 4u;
  const int C = p.chunk;
  const float bound = F16OUT ? __ldg(p.bounds + (p.pass == 0 ? 1 : 0)) : 1.f;
  const float scale16 = F16OUT ? f16_split_scale(bound) : 1.f;
  const float da_scale = F16OUT ? scale16 * __ldg(p.bounds + 2) : 1.f;
  const float gsS = p.g_scale * scale16;                   // attention weights of g rows carry the output scale
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    if (F16OUT && p.scale_out) *p.scale_out = scale16;
    if (p.pass == 1 && p.reruns) atomicAdd(p.reruns, 1);
  }
  int* qn = p.queue + 32 * h;
  float vmax = 0.f;                                        // largest |scaled value| this lane has written (fp16 range check)

  auto decode = [&](int item) -> SbItem {
    const int4 w = __ldg(reinterpret_cast<const int4*>(p.tasks) + item);
    SbItem k;
    k.o = w.x; k.q = w.y; k.a = w.z & 0xFFFFFF; k.c = (w.z >> 24) & 0x7F; k.s = w.w; k.ticket = item;
    return k;
  };
  auto n_ops_of = [&](const SbItem& k) -> int {
    return k.c == 0 ? 1 + 2 * k.a + min(k.s, C) : 1 + min(C, k.s - k.c * C);
  };
  auto dslope = [&](float e) -> float { return e > 0.f ? 1.f : p.neg_slope; };

  // ---- producer side: the item queue (tickets two items ahead) and the step cursor ----
  int pending = 0;
  SbItem fk = {0, 0, 0, 0, 0, 0}, nk = {0, 0, 0, 0, 0, 0};
  {
    int c0 = 0, c1 = 0;
    if (lane == 0) { c0 = atomicAdd(qn, 1); c1 = atomicAdd(qn, 1); pending = atomicAdd(qn, 1); }
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    c1 = __shfl_sync(0xffffffffu, c1, 0);
    fk.ticket = c0; nk.ticket = c1;
    if (c0 < p.n_tasks) fk = decode(c0);
    if (c1 < p.n_tasks) nk = decode(c1);
  }
  int f_t = 0, f_n = fk.ticket < p.n_tasks ? n_ops_of(fk) : 0;
  bool f_done = false;
  // The steps of an item are a pure function of its record; every lane works out ONE of them (step t0 + lane) when the item starts and
  // a staging step just picks its lane's descriptor with two shuffles: {kind | row << 3, first attention slot}.
  int my_kr = 0, my_so = 0;
  auto build_steps = [&](int t0) {
    const int t = t0 + lane;
    int kind = kOpEnd, row = 0, so = 0;
    if (fk.ticket < p.n_tasks && t < f_n) {
      if (t == 0) {
        kind = fk.c == 0 ? kOpAnchor0 : kOpAnchorC; row = fk.o + fk.a; so = fk.q + 2 * fk.a;
      } else if (fk.c == 0 && t <= fk.a) {
        kind = kOpGpDot; row = fk.o + t - 1; so = fk.q + fk.a + t - 1;
      } else if (fk.c == 0 && t <= 2 * fk.a) {
        kind = kOpGpOut; row = fk.o + t - fk.a - 1; so = fk.q + t - fk.a - 1;
      } else {
        const int j = fk.c == 0 ? t - fk.a : fk.a + 1 + fk.c * C + (t - 1);      // local row of the sibling
        kind = kOpSib; row = fk.o + j; so = fk.q + 2 * j - 1;
      }
    }
    my_kr = kind | (row << 3);
    my_so = so;
  };
  build_steps(0);
  // this lane's attention scalar of a step: lanes 0..5 -> {alpha[so], alpha[so + 1], elog[so], elog[so + 1], alpha_d[so], alpha_d[so + 1]}
  // (slot units, times H, plus h); which lanes take part depends on the kind (one byte per kind, bit = lane)
  const float* my_arr = (lane < 2 ? p.alpha : (lane < 4 ? p.elog : p.alpha_d)) + h;
  const uint64_t kLaneMask = (0x15ull << (8 * kOpAnchor0)) | (0x15ull << (8 * kOpGpDot)) | (0x10ull << (8 * kOpGpOut)) | (0x3Full << (8 * kOpSib));

  // stage step number `it` (slot it % 3): header, scalars, row copies.  Warp-uniform control flow; every lane arrives once.
  auto fetch = [&](int it) {
    if (f_done) return;
    const int st = it % kSbStages;
    const uint32_t bar = bar0 + 8 * st;
    const uint32_t slot = slot0 + st * kSlotB;
    if (f_t == f_n) {                                      // next item
      fk = nk;
      f_t = 0;
      const int t2 = __shfl_sync(0xffffffffu, pending, 0);
      nk.ticket = t2;
      if (t2 < p.n_tasks) nk = decode(t2);
      // (no new ticket once the queue is exhausted: every atomic of this warp has been read, i.e. performed, before it retires)
      if (lane == 0 && fk.ticket < p.n_tasks) pending = atomicAdd(qn, 1);
      f_n = fk.ticket < p.n_tasks ? n_ops_of(fk) : 0;
      build_steps(0);
    } else if ((f_t & 31) == 0) {
      build_steps(f_t);
    }
    const int kr = __shfl_sync(0xffffffffu, my_kr, f_t & 31), so = __shfl_sync(0xffffffffu, my_so, f_t & 31);
    const int kind = kr & 7, row = kr >> 3;
    ++f_t;
    f_done = kind == kOpEnd;
    if (lane < 8 && ((kLaneMask >> (8 * kind + lane)) & 1ull))
      sb_cp4(slot + 32 + 4 * lane, my_arr + (int64_t)(so + (kind == kOpSib ? (lane & 1) : 0)) * H);
    sb_cp_arrive(bar);
    if (lane == 0) {
      const bool has_g = (0x32 >> kind) & 1, has_f = (0x2E >> kind) & 1;       // kinds {1, 4, 5} carry g, {1, 2, 3, 5} carry ft
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(kind), "r"(row), "r"(fk.ticket), "r"(0) : "memory");
      if (kind == kOpAnchor0 || kind == kOpAnchorC)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 16), "r"(fk.o), "r"(fk.q), "r"(fk.a | (fk.c << 24)), "r"(fk.s) : "memory");
      sb_bar_arrive_tx(bar, ((has_g ? 1u : 0u) + (has_f ? 1u : 0u)) * rowB);
      if (has_g) sb_bulk_g2s(slot + kSbHdrBytes, gbase + (int64_t)row * p.ldg, rowB, bar);
      if (has_f) sb_bulk_g2s(slot + kSbHdrBytes + kRowB, fbase + (int64_t)row * p.ldf, rowB, bar);
    }
  };

  // d(ft_row) S = rg + (c1 S) attn_l + (c2 S) attn_r with rg ALREADY times S = the fp16-pair scale (1 for fp32 output; the callers fold
  // it into the attention weights), written as fp16 hi/lo pair or fp32, and the row's (da1, da2) = (c1, c2)
  auto write_row = [&](int row, const float4 (&rg)[NV], float c1, float c2) {
    const int64_t r64 = row;
    const float c1s = c1 * scale16, c2s = c2 * scale16;
    if constexpr (F16OUT) {
      __half* hi = p.dft16_hi + r64 * p.ld16 + (int64_t)h * D + lane * 4;
      __half* lo = p.dft16_lo + r64 * p.ld16 + (int64_t)h * D + lane * 4;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        if (t < NV - 1 || c4 * 4 < D) {
          const float4 l = s_l[c4], r = s_r[c4];
          const float x0 = fmaf(c1s, l.x, fmaf(c2s, r.x, rg[t].x)), x1 = fmaf(c1s, l.y, fmaf(c2s, r.y, rg[t].y));
          const float x2 = fmaf(c1s, l.z, fmaf(c2s, r.z, rg[t].z)), x3 = fmaf(c1s, l.w, fmaf(c2s, r.w, rg[t].w));
          vmax = fmaxf(fmaxf(vmax, fabsf(x0)), fmaxf(fabsf(x1), fmaxf(fabsf(x2), fabsf(x3))));
          const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
          *reinterpret_cast<uint2*>(hi + 128 * t) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(lo + 128 * t) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
      }
      // tail columns [H D, ld16): da1 of head h at H D + h, da2 at H D + H + h (both times c, scaled like the row), zeros after 2 H
      if (lane < 2) {
        const float x = (lane ? c2 : c1) * da_scale;
        vmax = fmaxf(vmax, fabsf(x));
        const __half xh = __float2half_rn(x);
        const int64_t o = r64 * p.ld16 + (int64_t)H * D + lane * H + h;
        p.dft16_hi[o] = xh;
        p.dft16_lo[o] = __float2half_rn(x - __half2float(xh));
      } else if (h == 0 && lane - 2 < p.tail16 - 2 * H) {
        const int64_t o = r64 * p.ld16 + (int64_t)H * D + 2 * H + (lane - 2);
        p.dft16_hi[o] = __float2half_rn(0.f);
        p.dft16_lo[o] = __float2half_rn(0.f);
      }
    } else {
      float* out = p.dft + r64 * p.ldd + (int64_t)h * D + lane * 4;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        if (t < NV - 1 || c4 * 4 < D) {
          const float4 l = s_l[c4], r = s_r[c4];
          float4 v;
          v.x = fmaf(c1s, l.x, fmaf(c2s, r.x, rg[t].x)); v.y = fmaf(c1s, l.y, fmaf(c2s, r.y, rg[t].y));
          v.z = fmaf(c1s, l.z, fmaf(c2s, r.z, rg[t].z)); v.w = fmaf(c1s, l.w, fmaf(c2s, r.w, rg[t].w));
          *reinterpret_cast<float4*>(out + 128 * t) = v;
        }
      }
    }
    if (p.da1 && lane == 0) {
      p.da1[r64 * H + h] = c1;
      p.da2[r64 * H + h] = c2;
    }
  };

  // ---- consumer state of the current item ----
  SbItem ck = {0, 0, 0, 0, 0, 0};
  int ops_left = 0;
  float4 gA[NV], fA[NV], accA[NV];
  float dd_mine = 0.f, al_mine = 0.f, el_mine = 1.f, ad_mine = 0.f, ds_mine = 0.f, tsum = 0.f, da1A = 0.f, da2A = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) gA[t] = fA[t] = accA[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  // softmax + leaky-relu backward over the anchor's a + 1 in-edges once all their dots are known
  auto finish_anchor_softmax = [&]() {
    const int a = ck.a, deg = a + 1;
    if (deg <= 32) {
      ds_mine = lane < deg ? (ad_mine * dd_mine - al_mine * tsum) * dslope(el_mine) : 0.f;
      da2A = warp_sum(ds_mine);
      da1A = __shfl_sync(0xffffffffu, ds_mine, a);
    } else {
      __syncwarp();
      float acc = 0.f;
      for (int k = lane; k < deg; k += 32) {
        const int64_t so = (int64_t)(ck.q + a + k) * H + h;
        const float dsv = (__ldg(p.alpha_d + so) * p.ds[so] - __ldg(p.alpha + so) * tsum) * dslope(__ldg(p.elog + so));
        p.ds[so] = dsv;
        acc += dsv;
      }
      da2A = warp_sum(acc);
      __syncwarp();
      da1A = p.ds[(int64_t)(ck.q + 2 * a) * H + h];
    }
  };

  fetch(0);
  fetch(1);
  for (int it = 0;; ++it) {
    __syncwarp();                                          // every lane is done with the slot that is staged next
    fetch(it + 2);
    const int st = it % kSbStages;
    sb_bar_wait(bar0 + 8 * st, (uint32_t)((it / kSbStages) & 1));
    const unsigned char* slot = my_slots + (size_t)st * kSlotB;
    const uint4 hd = *reinterpret_cast<const uint4*>(slot);
    const int kind = (int)hd.x, row = (int)hd.y;
    if (kind == kOpEnd) break;
    const float4 sc0 = *reinterpret_cast<const float4*>(slot + 32);       // alpha[so], alpha[so+1], elog[so], elog[so+1]
    const float2 sc1 = *reinterpret_cast<const float2*>(slot + 48);       // alpha_d[so], alpha_d[so+1]
    const float4* sg = reinterpret_cast<const float4*>(slot + kSbHdrBytes);
    const float4* sf = reinterpret_cast<const float4*>(slot + kSbHdrBytes + kRowB);
    const int D4 = D >> 2;

    if (kind == kOpAnchor0 || kind == kOpAnchorC) {
      const uint4 rec = *reinterpret_cast<const uint4*>(slot + 16);
      ck.o = (int)rec.x; ck.q = (int)rec.y; ck.a = (int)(rec.z & 0xFFFFFF); ck.c = (int)((rec.z >> 24) & 0x7F); ck.s = (int)rec.w;
      ck.ticket = (int)hd.z;
      ops_left = n_ops_of(ck);
      da1A = 0.f; da2A = 0.f; tsum = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        fA[t] = (t < NV - 1 || c4 < D4) ? sf[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (kind == kOpAnchor0) {
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          const int c4 = lane + 32 * t;
          gA[t] = (t < NV - 1 || c4 < D4) ? sg[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // the anchor's self loop: in-edge number a of its softmax
        const float d = warp_sum(sb_dot<NV>(gA, fA)) * gs;
        const float al = sc0.x, el = sc0.z, ad = sc1.x;
        tsum = ad * d;
        if (ck.a + 1 <= 32) {
          if (lane == ck.a) { dd_mine = d; al_mine = al; el_mine = el; ad_mine = ad; }
        } else if (lane == 0) {
          p.ds[(int64_t)(ck.q + 2 * ck.a) * H + h] = d;
        }
        const float w = ad * gsS;
#pragma unroll
        for (int t = 0; t < NV; ++t) accA[t] = make_float4(w * gA[t].x, w * gA[t].y, w * gA[t].z, w * gA[t].w);
        if (ck.a == 0) finish_anchor_softmax();
      } else {
#pragma unroll
        for (int t = 0; t < NV; ++t) accA[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (kind == kOpGpDot) {
      const int k = row - ck.o;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        if (t < NV - 1 || c4 < D4) {
          const float4 f = sf[c4];
          acc = fmaf(gA[t].x, f.x, acc); acc = fmaf(gA[t].y, f.y, acc); acc = fmaf(gA[t].z, f.z, acc); acc = fmaf(gA[t].w, f.w, acc);
        }
      }
      const float d = warp_sum(acc) * gs;
      const float al = sc0.x, el = sc0.z, ad = sc1.x;
      tsum = fmaf(ad, d, tsum);
      if (ck.a + 1 <= 32) {
        if (lane == k) { dd_mine = d; al_mine = al; el_mine = el; ad_mine = ad; }
      } else if (lane == 0) {
        p.ds[(int64_t)(ck.q + ck.a + k) * H + h] = d;
      }
      if (k == ck.a - 1) finish_anchor_softmax();
    } else if (kind == kOpGpOut) {
      const int j = row - ck.o;
      float dsk, w_anch;
      if (ck.a + 1 <= 32) {
        dsk = __shfl_sync(0xffffffffu, ds_mine, j);
        w_anch = __shfl_sync(0xffffffffu, ad_mine, j) * gsS;
      } else {
        const int64_t so = (int64_t)(ck.q + ck.a + j) * H + h;
        dsk = p.ds[so];
        w_anch = __ldg(p.alpha_d + so) * gsS;
      }
      const float w_self = sc1.x * gsS;
      float4 rg[NV];
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        const float4 x = (t < NV - 1 || c4 < D4) ? sg[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        rg[t].x = fmaf(w_self, x.x, w_anch * gA[t].x); rg[t].y = fmaf(w_self, x.y, w_anch * gA[t].y);
        rg[t].z = fmaf(w_self, x.z, w_anch * gA[t].z); rg[t].w = fmaf(w_self, x.w, w_anch * gA[t].w);
      }
      write_row(row, rg, dsk, 0.f);
    } else {   // kOpSib
      float4 rg[NV];
      float d1 = 0.f, d2 = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int c4 = lane + 32 * t;
        if (t < NV - 1 || c4 < D4) {
          rg[t] = sg[c4];
          const float4 f = sf[c4];
          d1 = fmaf(rg[t].x, fA[t].x, d1); d1 = fmaf(rg[t].y, fA[t].y, d1); d1 = fmaf(rg[t].z, fA[t].z, d1); d1 = fmaf(rg[t].w, fA[t].w, d1);
          d2 = fmaf(rg[t].x, f.x, d2); d2 = fmaf(rg[t].y, f.y, d2); d2 = fmaf(rg[t].z, f.z, d2); d2 = fmaf(rg[t].w, f.w, d2);
        } else {
          rg[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      sb_warp_sum2(d1, d2);
      d1 *= gs; d2 *= gs;
      const float al1 = sc0.x, al2 = sc0.y, el1 = sc0.z, el2 = sc0.w, ad1 = sc1.x, ad2 = sc1.y;
      const float ts = fmaf(ad1, d1, ad2 * d2);
      const float ds1 = (ad1 * d1 - al1 * ts) * dslope(el1);
      const float ds2 = (ad2 * d2 - al2 * ts) * dslope(el2);
      da1A += ds1;
      const float w1 = ad1 * gsS, w2 = ad2 * gsS;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        accA[t].x = fmaf(w1, rg[t].x, accA[t].x); accA[t].y = fmaf(w1, rg[t].y, accA[t].y);
        accA[t].z = fmaf(w1, rg[t].z, accA[t].z); accA[t].w = fmaf(w1, rg[t].w, accA[t].w);
        rg[t].x *= w2; rg[t].y *= w2; rg[t].z *= w2; rg[t].w *= w2;
      }
      write_row(row, rg, ds2, ds1 + ds2);
    }

    if (--ops_left == 0) {
      // ---- end of the item: the anchor's row ----
      const int n_chunks = ck.s > C ? (ck.s + C - 1) / C : 1;
      if (n_chunks == 1) {
        write_row(ck.o + ck.a, accA, da1A, da2A);
      } else {
        float* part = p.partial + ((int64_t)ck.ticket * H + h) * kPartLd;
#pragma unroll
        for (int t = 0; t < NV; ++t) __stcg(reinterpret_cast<float4*>(part) + lane + 32 * t, accA[t]);
        if (lane == 0) __stcg(reinterpret_cast<float4*>(part) + NV * 32, make_float4(da1A, da2A, 0.f, 0.f));
        __syncwarp();                                      // the warp's stores happen before lane 0's fence (cumulativity)
        const int first = ck.ticket - ck.c;
        int prev = 0;
        if (lane == 0) {
          __threadfence();
          prev = atomicAdd(p.counters + (int64_t)first * H + h, 1);
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == n_chunks - 1) {
          // last chunk to arrive: add the parts in chunk order (fixed) and write the anchor's row
          __threadfence();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int t = 0; t < NV; ++t) accA[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int cc = 0; cc < n_chunks; ++cc) {
            const float4* pr = reinterpret_cast<const float4*>(p.partial + ((int64_t)(first + cc) * H + h) * kPartLd);
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              const float4 x = __ldcg(pr + lane + 32 * t);
              accA[t].x += x.x; accA[t].y += x.y; accA[t].z += x.z; accA[t].w += x.w;
            }
            const float4 sc = __ldcg(pr + NV * 32);
            s1 += sc.x;
            if (cc == 0) s2 = sc.y;
          }
          if (lane == 0) p.counters[(int64_t)first * H + h] = 0;
          write_row(ck.o + ck.a, accA, s1, s2);
        }
      }
    }
  }
  if (F16OUT && vmax > 65504.f) *p.flag = 1;           // (benign race: every writer stores the same value)
  // ---- retire: the last warp of this head's queue resets it for the next launch ----
  if (lane == 0) {
    const int total = (int)(gridDim.x * kSbWarps);
    if (atomicAdd(qn + 1, 1) == total - 1) {
      qn[0] = 0;
      qn[1] = 0;
    }
  }
}

// d(attn_l)[h, :] = W_h v[h, :] / c,  d(attn_r)[h, :] = W_h v[H + h, :] / c   (W_h = rows h D .. h D + D - 1 of the layer's fc weight,
// v = z^T [da1 c | da2 c] = the 2 H extra rows of the weight-gradient GEMM): one warp per output element pair, fixed-order sums.
__global__ void attn_grad_from_v_kernel(const float* __restrict__ w, int64_t ldw, const float* __restrict__ v, int64_t ldv, int H, int D, int K,
                                        const float* __restrict__ c_ptr, float* __restrict__ dal, float* __restrict__ dar) {
  TX_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= H * D) return;
  const int h = row / D;
  const float* wr = w + (int64_t)row * ldw;
  const float* v1 = v + (int64_t)h * ldv;
  const float* v2 = v + (int64_t)(H + h) * ldv;
  float a = 0.f, b = 0.f;
  int k = lane;
  for (; k + 7 * 32 < K; k += 8 * 32) {          // 24 independent loads in flight (K = 2050: the plain loop was 64 dependent round trips)
    float x[8], y1[8], y2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = __ldg(wr + k + 32 * u); y1[u] = __ldg(v1 + k + 32 * u); y2[u] = __ldg(v2 + k + 32 * u); }
#pragma unroll
    for (int u = 0; u < 8; ++u) { a = fmaf(x[u], y1[u], a); b = fmaf(x[u], y2[u], b); }
  }
  for (; k < K; k += 32) {
    const float x = __ldg(wr + k);
    a = fmaf(x, __ldg(v1 + k), a);
    b = fmaf(x, __ldg(v2 + k), b);
  }
  sb_warp_sum2(a, b);
  if (lane == 0) {
    const float inv = 1.f / __ldg(c_ptr);
    dal[row] = a * inv;
    dar[row] = b * inv;
  }
}

template <int NV>
static size_t sb_smem_bytes() { return (size_t)2 * NV * 512 + 512 + (size_t)kSbWarps * kSbStages * (kSbHdrBytes + 2 * NV * 512); }

template <int NV, bool F16OUT>
static int sb_launch2(const StarBwdParams& p, dim3 grid, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = sb_smem_bytes<NV>();
  if (!attr_done) {
    if (cudaFuncSetAttribute(gat_star_bwd_kernel<NV, F16OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("gat_star_bwd: cannot reserve %zu bytes of shared memory", smem);
      return TX_ERR_CUDA;
    }
    attr_done = true;
  }
  TX_PDL_LAUNCH((gat_star_bwd_kernel<NV, F16OUT>), grid, kSbWarps * 32, smem, st, p);
  return TX_OK;
}
template <int NV>
static int sb_launch(const StarBwdParams& p, dim3 grid, cudaStream_t st) {
  return p.dft16_hi ? sb_launch2<NV, true>(p, grid, st) : sb_launch2<NV, false>(p, grid, st);
}

}  // namespace tx

using namespace tx;

extern "C" {

int64_t tx_gat_star_bwd_partial_floats(int64_t n_tasks, int64_t heads, int64_t dim) {
  return n_tasks * heads * (((dim + 127) / 128) * 128 + 4);
}

int tx_gat_star_bwd(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft, int64_t ldf,
                    const float* alpha, const float* alpha_d, const float* elog, const float* attn_l, const float* attn_r,
                    const int32_t* tasks, int64_t n_tasks, int64_t chunk, int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope,
                    float* ds, float* da1, float* da2, float* dft, int64_t ldd, void* dft16_hi, void* dft16_lo, int64_t ld16,
                    const float* bounds, int32_t* flag, int32_t* reruns, float* scale_out, float* partial, int32_t* counters,
                    int32_t* queue, void* stream) {
  TX_REQUIRE(g_head_stride != 0 || heads == 1, "gat_star_bwd: a shared g row (head mean) needs heads == 1");
  TX_REQUIRE(!dft16_hi || (dft16_lo && bounds && flag && aligned16(dft16_hi) && aligned16(dft16_lo) && ld16 % 8 == 0 &&
                           ld16 >= heads * dim + 2 * heads && ld16 - heads * dim <= 2 * heads + 30),
             "gat_star_bwd: bad fp16 output buffers (ld16 must hold heads * dim + 2 * heads columns)");
  TX_REQUIRE(dft16_hi || (dft && da1 && da2 && aligned16(dft) && ldd % 4 == 0 && ldd >= heads * dim),
             "gat_star_bwd: an output buffer is required (fp32 output also needs da1 / da2)");
  TX_REQUIRE((da1 == nullptr) == (da2 == nullptr), "gat_star_bwd: da1 and da2 go together");
  TX_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 512, "gat_star_bwd: dim must be a multiple of 4 and <= 512");
  TX_REQUIRE(aligned16(g) && ldg % 4 == 0 && g_head_stride % 4 == 0 && aligned16(ft) && ldf % 4 == 0 && aligned16(attn_l) &&
             aligned16(attn_r) && aligned16(partial), "gat_star_bwd: 16-byte aligned rows required");
  TX_REQUIRE(tasks && aligned16(tasks) && alpha && alpha_d && elog && ds && partial && counters && queue, "gat_star_bwd: null pointer");
  TX_REQUIRE(n_tasks >= 0 && n_tasks < (1ll << 31) && n_nodes >= 0 && n_nodes < (1ll << 31) && chunk >= 1 && chunk < (1 << 20) &&
             heads >= 1 && heads <= 64, "gat_star_bwd: bad sizes");
  if (n_nodes == 0 || n_tasks == 0) return TX_OK;
  StarBwdParams p;
  p.g = g; p.ldg = ldg; p.g_head_stride = g_head_stride; p.g_scale = g_scale; p.ft = ft; p.ldf = ldf;
  p.alpha = alpha; p.alpha_d = alpha_d; p.elog = elog; p.attn_l = attn_l; p.attn_r = attn_r;
  p.tasks = tasks; p.n_tasks = (int)n_tasks; p.chunk = (int)chunk; p.H = (int)heads; p.D = (int)dim; p.neg_slope = neg_slope;
  p.ds = ds; p.da1 = da1; p.da2 = da2; p.dft = dft; p.ldd = ldd;
  p.dft16_hi = (__half*)dft16_hi; p.dft16_lo = (__half*)dft16_lo; p.ld16 = ld16; p.tail16 = (int)(ld16 - heads * dim);
  p.bounds = bounds; p.flag = flag; p.reruns = reruns; p.scale_out = scale_out; p.partial = partial; p.counters = counters; p.queue = queue;
  const int nv = (int)((dim + 127) / 128);
  int gx = (int)((kNumSms + heads - 1) / heads);
  const int64_t need = (n_tasks + kSbWarps - 1) / kSbWarps;
  if (gx > need) gx = (int)need;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)heads);
  cudaStream_t st = (cudaStream_t)stream;
  const int passes = dft16_hi ? 2 : 1;
  for (int pass = 0; pass < passes; ++pass) {
    p.pass = pass;
    int rc;
    switch (nv) {
      case 1: rc = sb_launch<1>(p, grid, st); break;
      case 2: rc = sb_launch<2>(p, grid, st); break;
      case 3: rc = sb_launch<3>(p, grid, st); break;
      default: rc = sb_launch<4>(p, grid, st); break;
    }
    if (rc != TX_OK) return rc;
    TX_LAUNCH_CHECK("tx_gat_star_bwd");
  }
  return TX_OK;
}

int tx_attn_grad_from_v(const float* weight, int64_t ldw, const float* v, int64_t ldv, int64_t heads, int64_t dim, int64_t k,
                        const float* c, float* dattn_l, float* dattn_r, void* stream) {
  TX_REQUIRE(weight && v && c && dattn_l && dattn_r && heads >= 1 && dim >= 1 && k >= 1 && ldw >= k && ldv >= k, "attn_grad_from_v: bad arguments");
  const int64_t rows = heads * dim;
  TX_PDL_LAUNCH((attn_grad_from_v_kernel), (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, weight, ldw, v, ldv, (int)heads, (int)dim, (int)k, c,
                                                                                        dattn_l, dattn_r);
  TX_LAUNCH_CHECK("tx_attn_grad_from_v");
  return TX_OK;
}

}  // extern "C"
