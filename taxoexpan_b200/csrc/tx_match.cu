// Matching + InfoNCE epilogue: the step right after the readout.
//
//   tx_match_rowdot_fwd/bwd : scores[g] = f(<u_g, q_g>), u = hg W (the bilinear form's projection, a GEMM on tx_gemm.cu),
//                             f = identity (BIM, model_zoo.py:301-313) or exp (LBM, model_zoo.py:316-328).  Replaces the
//                             row-dot half of nn.Bilinear + torch.exp and their autograd kernels.
//   tx_info_nce_fwd/bwd     : loss = sum_q [logsumexp(scores[q, :]) - scores[q, target_q]] = F.cross_entropy(.., reduction="sum")
//                             on the [n_queries, 1 + negative_size] reshape of trainer/trainer.py:52-56 (loss.py:52-57).
// All of it is HBM-trivial (G x r floats); what it buys is launch count: ~12 small torch kernels per step become 5.
// A warp owns one row (egonet or query); reductions are warp shuffles in a fixed order, the loss is summed by
// tx_reduce_partials in a fixed order: results are run-to-run deterministic like the rest of the library.
#include <math.h>

#include <initializer_list>

#define TX_PDL_GROUP 1
#include "tx_common.cuh"

namespace tx {

template <int VEC>
struct VecN;
template <>
struct VecN<4> { using type = float4; };
template <>
struct VecN<2> { using type = float2; };
template <>
struct VecN<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ float dot_vec(const typename VecN<VEC>::type& a, const typename VecN<VEC>::type& b) {
  if constexpr (VEC == 4) return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
  else if constexpr (VEC == 2) return fmaf(a.x, b.x, a.y * b.y);
  else return a * b;
}
template <int VEC>
__device__ __forceinline__ typename VecN<VEC>::type scale_vec(float s, const typename VecN<VEC>::type& a) {
  if constexpr (VEC == 4) return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
  else if constexpr (VEC == 2) return make_float2(s * a.x, s * a.y);
  else return s * a;
}

template <int VEC>
__global__ void __launch_bounds__(256) match_rowdot_fwd_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ q,
                                                               int64_t ldq, int n_rows, int r, int apply_exp,
                                                               float* __restrict__ scores) {
  TX_PDL_ENTER();
  using V = typename VecN<VEC>::type;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int rv = r / VEC;
  for (int g = warp; g < n_rows; g += nwarps) {
    const V* ur = reinterpret_cast<const V*>(u + (int64_t)g * ldu);
    const V* qr = reinterpret_cast<const V*>(q + (int64_t)g * ldq);
    float acc = 0.f;
    for (int c = lane; c < rv; c += 32) acc += dot_vec<VEC>(__ldg(ur + c), __ldg(qr + c));
    acc = warp_sum(acc);
    if (lane == 0) scores[g] = apply_exp ? expf(acc) : acc;
  }
}

// d(t) = d(score) * (exp: score, identity: 1);  du_g = d(t) q_g;  dq_g = d(t) u_g
template <int VEC>
__global__ void __launch_bounds__(256) match_rowdot_bwd_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ q,
                                                               int64_t ldq, const float* __restrict__ scores,
                                                               const float* __restrict__ dscores, int n_rows, int r, int apply_exp,
                                                               float* __restrict__ du, int64_t lddu, float* __restrict__ dq,
                                                               int64_t lddq) {
  TX_PDL_ENTER();
  using V = typename VecN<VEC>::type;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int rv = r / VEC;
  for (int g = warp; g < n_rows; g += nwarps) {
    float dt = __ldg(dscores + g);
    if (apply_exp) dt *= __ldg(scores + g);
    const V* ur = reinterpret_cast<const V*>(u + (int64_t)g * ldu);
    const V* qr = reinterpret_cast<const V*>(q + (int64_t)g * ldq);
    if (du) {
      V* o = reinterpret_cast<V*>(du + (int64_t)g * lddu);
      for (int c = lane; c < rv; c += 32) o[c] = scale_vec<VEC>(dt, __ldg(qr + c));
    }
    if (dq) {
      V* o = reinterpret_cast<V*>(dq + (int64_t)g * lddq);
      for (int c = lane; c < rv; c += 32) o[c] = scale_vec<VEC>(dt, __ldg(ur + c));
    }
  }
}

// one warp per query: m = max_j x_j, s = sum_j exp(x_j - m) (log_softmax's own formulation), loss_q = m + log s - x_target
__global__ void __launch_bounds__(256) info_nce_fwd_kernel(const float* __restrict__ scores, int n_queries, int group,
                                                           const int32_t* __restrict__ target, float* __restrict__ loss_q,
                                                           float* __restrict__ lse_q) {
  TX_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int qi = warp; qi < n_queries; qi += nwarps) {
    const float* row = scores + (int64_t)qi * group;
    float m = -INFINITY;
    for (int j = lane; j < group; j += 32) m = fmaxf(m, __ldg(row + j));
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < group; j += 32) s += expf(__ldg(row + j) - m);
    s = warp_sum(s);
    if (lane == 0) {
      const int t = target ? __ldg(target + qi) : 0;
      const float lse = m + logf(s);
      loss_q[qi] = (t >= 0 && t < group) ? lse - __ldg(row + t) : NAN;     // out-of-range class index: a visible NaN, never an OOB read
      if (lse_q) lse_q[qi] = lse;
    }
  }
}

// d(scores[q, j]) = d(loss) * (softmax_j - [j == target_q]); softmax_j = exp(x_j - lse_q)
__global__ void __launch_bounds__(256) info_nce_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ lse_q,
                                                           int64_t total, int group, const int32_t* __restrict__ target,
                                                           const float* __restrict__ dloss, float* __restrict__ dscores) {
  TX_PDL_ENTER();
  const float gl = __ldg(dloss);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int qi = (int)(i / group), j = (int)(i - (int64_t)qi * group);
    const int t = target ? __ldg(target + qi) : 0;
    const float p = expf(__ldg(scores + i) - __ldg(lse_q + qi));
    dscores[i] = gl * (p - (j == t ? 1.f : 0.f));
  }
}

static int pick_vec(int64_t r, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  int vec = 4;
  auto lower = [&](bool ok4, bool ok2) { if (!ok4 && vec == 4) vec = 2; if (!ok2) vec = 1; };
  lower(r % 4 == 0, r % 2 == 0);
  for (int64_t ld : lds) lower(ld % 4 == 0, ld % 2 == 0);
  for (const void* p : ptrs) {
    if (!p) continue;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    lower((a & 15u) == 0, (a & 7u) == 0);
  }
  return vec;
}

}  // namespace tx

using namespace tx;

extern "C" {

int tx_match_rowdot_fwd(const float* u, int64_t ldu, const float* q, int64_t ldq, int64_t n_rows, int64_t r, int32_t apply_exp,
                        float* scores, void* stream) {
  TX_REQUIRE(n_rows >= 0 && r >= 0 && n_rows < (1ll << 31) && r < (1ll << 31), "tx_match_rowdot_fwd: bad sizes (%lld rows, r = %lld)",
             (long long)n_rows, (long long)r);
  if (n_rows == 0) return TX_OK;
  TX_REQUIRE(u && q && scores, "tx_match_rowdot_fwd: null pointer");
  TX_REQUIRE(ldu >= r && ldq >= r, "tx_match_rowdot_fwd: leading dimension smaller than r");
  const int grid = grid_for_warps(n_rows, 8, 8);
  const int vec = pick_vec(r, {ldu, ldq}, {u, q});
  cudaStream_t s = (cudaStream_t)stream;
  if (vec == 4) TX_PDL_LAUNCH((match_rowdot_fwd_kernel<4>), grid, 256, 0, s, u, ldu, q, ldq, (int)n_rows, (int)r, apply_exp, scores);
  else if (vec == 2) TX_PDL_LAUNCH((match_rowdot_fwd_kernel<2>), grid, 256, 0, s, u, ldu, q, ldq, (int)n_rows, (int)r, apply_exp, scores);
  else TX_PDL_LAUNCH((match_rowdot_fwd_kernel<1>), grid, 256, 0, s, u, ldu, q, ldq, (int)n_rows, (int)r, apply_exp, scores);
  TX_LAUNCH_CHECK("tx_match_rowdot_fwd");
  return TX_OK;
}

int tx_match_rowdot_bwd(const float* u, int64_t ldu, const float* q, int64_t ldq, const float* scores, const float* dscores,
                        int64_t n_rows, int64_t r, int32_t apply_exp, float* du, int64_t lddu, float* dq, int64_t lddq,
                        void* stream) {
  TX_REQUIRE(n_rows >= 0 && r >= 0 && n_rows < (1ll << 31) && r < (1ll << 31), "tx_match_rowdot_bwd: bad sizes (%lld rows, r = %lld)",
             (long long)n_rows, (long long)r);
  if (n_rows == 0 || (!du && !dq)) return TX_OK;
  TX_REQUIRE(u && q && dscores && (scores || !apply_exp), "tx_match_rowdot_bwd: null pointer");
  TX_REQUIRE(ldu >= r && ldq >= r && (!du || lddu >= r) && (!dq || lddq >= r), "tx_match_rowdot_bwd: leading dimension smaller than r");
  const int grid = grid_for_warps(n_rows, 8, 8);
  const int vec = pick_vec(r, {ldu, ldq, du ? lddu : 4, dq ? lddq : 4}, {u, q, du, dq});
  cudaStream_t s = (cudaStream_t)stream;
  if (vec == 4)
    TX_PDL_LAUNCH((match_rowdot_bwd_kernel<4>), grid, 256, 0, s, u, ldu, q, ldq, scores, dscores, (int)n_rows, (int)r, apply_exp, du, lddu, dq, lddq);
  else if (vec == 2)
    TX_PDL_LAUNCH((match_rowdot_bwd_kernel<2>), grid, 256, 0, s, u, ldu, q, ldq, scores, dscores, (int)n_rows, (int)r, apply_exp, du, lddu, dq, lddq);
  else
    TX_PDL_LAUNCH((match_rowdot_bwd_kernel<1>), grid, 256, 0, s, u, ldu, q, ldq, scores, dscores, (int)n_rows, (int)r, apply_exp, du, lddu, dq, lddq);
  TX_LAUNCH_CHECK("tx_match_rowdot_bwd");
  return TX_OK;
}

int tx_info_nce_fwd(const float* scores, int64_t n_queries, int64_t group, const int32_t* target, float* loss_per_query,
                    float* lse_per_query, float* loss, void* stream) {
  TX_REQUIRE(n_queries >= 0 && group >= 1 && n_queries < (1ll << 31) && group < (1ll << 31) && n_queries * group < (1ll << 40),
             "tx_info_nce_fwd: bad sizes (%lld queries x %lld)", (long long)n_queries, (long long)group);
  TX_REQUIRE(loss, "tx_info_nce_fwd: null loss pointer");
  if (n_queries == 0) {
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("tx_info_nce_fwd: memset failed: %s", cudaGetErrorString(e)); return TX_ERR_CUDA; }
    return TX_OK;
  }
  TX_REQUIRE(scores && loss_per_query, "tx_info_nce_fwd: null pointer");
  TX_PDL_LAUNCH((info_nce_fwd_kernel), grid_for_warps(n_queries, 8, 8), 256, 0, (cudaStream_t)stream, scores, (int)n_queries, (int)group, target,
                                                                                        loss_per_query, lse_per_query);
  TX_LAUNCH_CHECK("tx_info_nce_fwd");
  return tx_reduce_partials(loss_per_query, n_queries, 1, loss, stream);
}

int tx_info_nce_bwd(const float* scores, const float* lse_per_query, int64_t n_queries, int64_t group, const int32_t* target,
                    const float* dloss, float* dscores, void* stream) {
  TX_REQUIRE(n_queries >= 0 && group >= 1 && n_queries < (1ll << 31) && group < (1ll << 31) && n_queries * group < (1ll << 40),
             "tx_info_nce_bwd: bad sizes (%lld queries x %lld)", (long long)n_queries, (long long)group);
  if (n_queries == 0) return TX_OK;
  TX_REQUIRE(scores && lse_per_query && dloss && dscores, "tx_info_nce_bwd: null pointer");
  const int64_t total = n_queries * group;
  const int64_t need = (total + 255) / 256, cap = (int64_t)kNumSms * 8;
  TX_PDL_LAUNCH((info_nce_bwd_kernel), (int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream, scores, lse_per_query, total, (int)group,
                                                                                       target, dloss, dscores);
  TX_LAUNCH_CHECK("tx_info_nce_bwd");
  return TX_OK;
}

}  // extern "C"
