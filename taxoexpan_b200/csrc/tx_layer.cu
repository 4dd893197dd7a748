// One C-ABI call per GAT layer and direction ("caller-owned workspace, tx_workspace_bytes"): the launch sequence of
// the default hot path - fp16-pair tensor-core GEMMs (tx_gemm.cu) + star-egonet fused forward / backward (tx_star_fwd.cu,
// tx_star_bwd.cu) + their small helpers - enqueued from native code into ONE caller-owned workspace, instead of ~15 ctypes calls
// and ~30 torch allocations per layer and direction from Python (the host needed 1.75 ms to enqueue a 2.0 ms step; round-1 verdict).
// Same kernels, same arguments, same order as the per-kernel path of taxoexpan_b200/functional.py (which remains for every
// configuration this path does not cover): results are bit-identical.  Reference call sites: GATLayer.forward and its autograd,
// model/model_zoo.py:80-114, inside the PGAT / GAT stacks of :183-190,210-220.
//
// Nothing here allocates device memory, synchronises or reads device data on the host.  The optional per-launch timing
// (tx_prof_enable) is a measurement aid for bench.py: it creates CUDA events and brackets every launch of these calls.
#include <string.h>

#include <string>
#include <vector>

#include "tx_common.cuh"

namespace tx {

static inline int64_t r4(int64_t k) { return (k + 3) / 4 * 4; }
static inline int64_t r8(int64_t k) { return (k + 7) / 8 * 8; }

// bump allocator over the caller's workspace (256-byte aligned pieces)
struct Carver {
  unsigned char* base;
  size_t off;
  explicit Carver(void* p) : base(reinterpret_cast<unsigned char*>(p)), off(0) {}
  template <typename T>
  T* take(int64_t count) {
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += ((size_t)(count > 0 ? count : 0) * sizeof(T) + 255) / 256 * 256;
    return r;
  }
};

// ---- optional per-launch CUDA-event timing ----
struct ProfEntry { std::string name, tag; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
static int64_t g_sub_launches = 0;
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(const char* name, const char* tag, cudaStream_t s) : st(s), on(g_prof_on) {
    ++g_sub_launches;
    if (!on) return;
    g_prof.push_back(ProfEntry{name, tag ? tag : "", prof_event(), prof_event()});
    cudaEventRecord(g_prof.back().e0, st);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_prof.back().e1, st); }
};

#define TX_SUB(call)                 \
  do {                               \
    int rc_ = (call);                \
    if (rc_ != TX_OK) return rc_;    \
  } while (0)

struct FwdLayout {
  __half *z_hi, *z_lo; float *z_scale, *z_amax;
  __half *w_hi, *w_lo, *wt_hi, *wt_lo; float* w_scal;
  float *ft, *ft_amax, *alpha, *alpha_d, *elog;
  uint32_t* maskbits; __half *out_hi, *out_lo; float *out_scale, *bound, *scal;
  size_t bytes;
};
constexpr int kScalars = 16;     // floats in a layout's scalar block (64 bytes)

static FwdLayout carve_fwd(const tx_gat_layer_desc& d, int split_input, void* ws) {
  Carver c(ws);
  FwdLayout L;
  const int64_t F = d.heads * d.dim, K = d.k;
  L.z_hi = c.take<__half>(split_input ? d.n * r8(K) : 0);
  L.z_lo = c.take<__half>(split_input ? d.n * r8(K) : 0);
  L.scal = c.take<float>(kScalars);          // every device scalar of the call in one block, cleared by one memset
  L.w_scal = L.scal; L.z_scale = L.scal + 4; L.z_amax = L.scal + 5; L.ft_amax = L.scal + 6; L.out_scale = L.scal + 7; L.bound = L.scal + 8;
  L.w_hi = c.take<__half>(F * r8(K));
  L.w_lo = c.take<__half>(F * r8(K));
  L.wt_hi = c.take<__half>(K * r8(F));
  L.wt_lo = c.take<__half>(K * r8(F));
  L.ft = c.take<float>(d.n * F);
  L.alpha = c.take<float>(d.e * d.heads);
  L.elog = c.take<float>(d.e * d.heads);
  L.alpha_d = d.p_attn > 0.f ? c.take<float>(d.e * d.heads) : L.alpha;
  const bool mask = d.hidden && (d.act_slope != 1.f || d.p_next > 0.f);
  L.maskbits = c.take<uint32_t>(mask ? tx_gat_fused_mask_words(d.n, d.heads, d.dim) : 0);
  if (!mask) L.maskbits = nullptr;
  const int64_t ld16 = r8(F + d.pos_dim);
  L.out_hi = c.take<__half>(d.hidden ? d.n * ld16 : 0);
  L.out_lo = c.take<__half>(d.hidden ? d.n * ld16 : 0);
  L.bytes = c.off;
  return L;
}

struct BwdLayout {
  float* pos_partial; float* bounds; __half *d_hi, *d_lo; float *d_scale, *g_amax, *star_partial, *ds, *tn_partial, *dz_amax, *scal;
  int64_t splits;
  size_t bytes;
};

static BwdLayout carve_bwd(const tx_gat_layer_desc& d, void* ws) {
  Carver c(ws);
  BwdLayout L;
  const int64_t F = d.heads * d.dim, K = d.k, M = F + 2 * d.heads;
  L.pos_partial = c.take<float>(d.hidden && d.pos_dim > 0 ? tx_row_blocks(d.n) * d.vocab * d.pos_dim : 0);
  L.scal = c.take<float>(kScalars);
  L.bounds = L.scal; L.d_scale = L.scal + 4; L.g_amax = L.scal + 5; L.dz_amax = L.scal + 6;
  const int64_t ld16 = r8(M);
  L.d_hi = c.take<__half>(d.n * ld16);
  L.d_lo = c.take<__half>(d.n * ld16);
  L.star_partial = c.take<float>(tx_gat_star_bwd_partial_floats(d.n_tasks_bwd, d.heads, d.dim));
  L.ds = c.take<float>(d.e * d.heads);
  L.splits = tx_gemm_tn_f16_splits(M, K, d.n);
  L.tn_partial = c.take<float>(L.splits > 1 ? L.splits * M * r4(K) : 0);
  L.bytes = c.off;
  return L;
}

// process-wide (backward runs on autograd's thread, not the caller's)
static cudaEvent_t g_after_star_event = nullptr;
static int64_t g_after_star_count = 0;

static int check_desc(const tx_gat_layer_desc* d, const char* who) {
  TX_REQUIRE(d, "%s: null descriptor", who);
  TX_REQUIRE(d->n > 0 && d->e > 0 && d->k > 0 && d->heads >= 1 && d->dim > 0 && d->dim % 4 == 0 && d->pos_dim >= 0, "%s: bad sizes", who);
  TX_REQUIRE(d->hidden || d->heads == 1, "%s: the output layer (head mean) needs heads == 1 on this path", who);
  TX_REQUIRE(d->weight && d->attn_l && d->attn_r && d->ldw >= d->k, "%s: parameters missing", who);
  TX_REQUIRE(d->tasks_fwd && d->tasks_bwd && d->queue && d->counters, "%s: star task tables / queue / counters missing", who);
  TX_REQUIRE(!(d->hidden && d->pos_dim > 0) || (d->next_pos_table && d->pos && d->vocab > 0), "%s: next position table / positions missing", who);
  return TX_OK;
}

}  // namespace tx

using namespace tx;

extern "C" {

int64_t tx_gat_layer_fwd_bytes(const tx_gat_layer_desc* d, int32_t split_input) {
  if (!d) return -1;
  return (int64_t)carve_fwd(*d, split_input, nullptr).bytes;
}
int64_t tx_gat_layer_bwd_bytes(const tx_gat_layer_desc* d) {
  if (!d) return -1;
  return (int64_t)carve_bwd(*d, nullptr).bytes;
}

int tx_gat_layer_fwd(const tx_gat_layer_desc* d, const float* z, int64_t ldz, const tx_gat_layer_state* prev, void* workspace,
                     tx_gat_layer_state* state, float* out, void* stream) {
  TX_SUB(check_desc(d, "gat_layer_fwd"));
  TX_REQUIRE(workspace && state && aligned16(workspace), "gat_layer_fwd: workspace / state missing");
  TX_REQUIRE((z != nullptr) != (prev != nullptr), "gat_layer_fwd: exactly one of z (fp32 input) / prev (the previous layer's fp16-pair output)");
  TX_REQUIRE(d->hidden || out, "gat_layer_fwd: the output layer needs `out`");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = d->n, K = d->k, H = d->heads, D = d->dim, F = H * D, pd = d->hidden ? d->pos_dim : 0;
  const FwdLayout L = carve_fwd(*d, z != nullptr, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("gat_layer_fwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  tx_gat_layer_state& S = *state;
  memset(&S, 0, sizeof(S));
  if (z) {
    { ProfScope ps("tx_absmax", d->tag, st); TX_SUB(tx_absmax(z, ldz, n, K, L.z_amax, stream)); }
    { ProfScope ps("tx_split_f16", d->tag, st); TX_SUB(tx_split_f16(z, ldz, n, K, L.z_amax, L.z_hi, L.z_lo, r8(K), L.z_scale, stream)); }
    S.z_hi = L.z_hi; S.z_lo = L.z_lo; S.z_scale = L.z_scale; S.ldz16 = r8(K);
  } else {
    TX_REQUIRE(prev->out_hi && prev->out_lo && prev->out_scale && prev->ld16_out >= r8(K), "gat_layer_fwd: the previous layer published no fp16 pair");
    S.z_hi = prev->out_hi; S.z_lo = prev->out_lo; S.z_scale = prev->out_scale; S.ldz16 = prev->ld16_out;
  }
  // the weights are split ONCE per step, in both orientations: [F, K] for this GEMM, [K, F] for the input-gradient GEMM
  { ProfScope ps("tx_split_f16_weight", d->tag, st);
    TX_SUB(tx_split_f16_weight(d->weight, d->ldw, F, K, L.w_hi, L.w_lo, r8(K), L.wt_hi, L.wt_lo, r8(F), L.w_scal, L.w_scal + 2, stream)); }
  S.wt_hi = L.wt_hi; S.wt_lo = L.wt_lo; S.w_scale = L.w_scal + 2; S.ldwt = r8(F);
  // (tx_gemm_nt_f16x3 clears its amax_out itself)
  { ProfScope ps("gemm_fwd", d->tag, st);                                       // ft = fc(h), model_zoo.py:83
    TX_SUB(tx_gemm_nt_f16x3(S.z_hi, S.z_lo, S.ldz16, L.w_hi, L.w_lo, r8(K), S.z_scale, S.w_scale, L.ft, F, n, F, K, nullptr, L.ft_amax, stream)); }
  S.ft = L.ft; S.ft_amax = L.ft_amax; S.alpha = L.alpha; S.alpha_d = L.alpha_d; S.elog = L.elog;
  tx_gat_epilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.mean_heads = d->hidden ? 0 : 1;
  epi.act_slope = d->act_slope;
  epi.next_pos_table = pd > 0 ? d->next_pos_table : nullptr;
  epi.pos = pd > 0 ? d->pos : nullptr;
  epi.pos_dim = pd;
  epi.p_drop = d->hidden ? d->p_next : 0.f;
  epi.seed = d->next_seed;
  epi.stream_id = d->next_stream;
  if (d->hidden) {
    // |z_next| <= max|ft| / ((1 - p_attn)(1 - p_next)) (attention weights are convex), appended rows <= max|P| / (1 - p_next)
    const int64_t ld16 = r8(F + pd);
    { ProfScope ps("tx_bound_max2", d->tag, st);
      TX_SUB(tx_bound_max2(L.ft_amax, 1.f / ((1.f - d->p_attn) * (1.f - d->p_next)), pd > 0 ? d->next_pos_table : nullptr,
                           pd > 0 ? d->vocab * pd : 0, 1.f / (1.f - d->p_next), L.bound, stream)); }
    { ProfScope ps("tx_gat_star_fwd", d->tag, st);
      TX_SUB(tx_gat_star_fwd(L.ft, F, d->attn_l, d->attn_r, d->tasks_fwd, d->n_tasks_fwd, d->chunk_fwd, n, H, D, d->neg_slope, d->p_attn,
                             d->attn_seed, d->attn_stream, L.alpha, L.alpha_d, L.elog, nullptr, r4(F + pd), &epi, L.maskbits, L.out_hi,
                             L.out_lo, ld16, L.bound, L.out_scale, d->queue, stream)); }
    S.out_hi = L.out_hi; S.out_lo = L.out_lo; S.out_scale = L.out_scale; S.ld16_out = ld16; S.maskbits = L.maskbits;
  } else {
    ProfScope ps("tx_gat_star_fwd", d->tag, st);
    TX_SUB(tx_gat_star_fwd(L.ft, F, d->attn_l, d->attn_r, d->tasks_fwd, d->n_tasks_fwd, d->chunk_fwd, n, H, D, d->neg_slope, d->p_attn,
                           d->attn_seed, d->attn_stream, L.alpha, L.alpha_d, L.elog, out, D, &epi, nullptr, nullptr, nullptr, 0, nullptr,
                           nullptr, d->queue, stream));
  }
  S.heads = H; S.dim = D; S.act_slope = d->act_slope; S.p_next = d->hidden ? d->p_next : 0.f;
  return TX_OK;
}

int tx_gat_layer_bwd(const tx_gat_layer_desc* d, const tx_gat_layer_state* state, const tx_gat_layer_state* prev, const float* dout,
                     int64_t ldg, const float* g_amax, void* workspace, float* dz, float* dw_ext, float* dw_main, float* dattn_l,
                     float* dattn_r, float* dtab, float** dz_amax_out, void* stream) {
  TX_SUB(check_desc(d, "gat_layer_bwd"));
  TX_REQUIRE(state && workspace && aligned16(workspace) && dout && dw_ext && dattn_l && dattn_r, "gat_layer_bwd: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const tx_gat_layer_state& S = *state;
  const int64_t n = d->n, K = d->k, H = d->heads, D = d->dim, F = H * D, M = F + 2 * H, pd = d->hidden ? d->pos_dim : 0;
  const BwdLayout L = carve_bwd(*d, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("gat_layer_bwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  if (d->hidden && pd > 0 && dtab) {
    // gradient of the position rows appended by this layer's epilogue (model_zoo.py:214-215)
    { ProfScope ps("tx_pos_grad_partials", d->tag, st);
      TX_SUB(tx_pos_grad_partials(dout, ldg, F, d->pos, n, pd, d->vocab, d->p_next, d->next_seed, d->next_stream, L.pos_partial, stream)); }
    { ProfScope ps("tx_reduce_partials", d->tag, st); TX_SUB(tx_reduce_partials(L.pos_partial, tx_row_blocks(n), d->vocab * pd, dtab, stream)); }
  }
  const int64_t g_head_stride = d->hidden ? D : 0;
  const float g_scale = d->hidden ? 1.f : 1.f / (float)H;
  if (!g_amax) {
    ProfScope ps("tx_absmax", d->tag, st);
    TX_SUB(tx_absmax(dout, ldg, n, d->hidden ? F : D, L.g_amax, stream));
    g_amax = L.g_amax;
  }
  // |dft| bounds (rigorous and optimistic), see taxoexpan_b200/functional.py and tx_bound_dft
  const float deg = (float)(d->max_out_deg > 1 ? d->max_out_deg : 1);
  const float slope = fabsf(d->neg_slope) > 1.f ? fabsf(d->neg_slope) : 1.f;
  const float keep = 1.f - d->p_attn;
  { ProfScope ps("tx_bound_dft", d->tag, st);
    TX_SUB(tx_bound_dft(g_amax, S.ft_amax, d->attn_l, d->attn_r, F, g_scale * deg / keep, g_scale * 2.f * (deg + 1.f) * (float)D * slope / keep,
                        g_scale * d->dft_optimism / keep, L.bounds, stream)); }
  const int64_t ld16 = r8(M);
  { ProfScope ps("tx_gat_star_bwd", d->tag, st);
    TX_SUB(tx_gat_star_bwd(dout, ldg, g_head_stride, g_scale, S.ft, F, S.alpha, S.alpha_d, S.elog, d->attn_l, d->attn_r, d->tasks_bwd,
                           d->n_tasks_bwd, d->chunk_bwd, n, H, D, d->neg_slope, L.ds, nullptr, nullptr, nullptr, F, L.d_hi, L.d_lo, ld16,
                           L.bounds, reinterpret_cast<int32_t*>(L.bounds + 3), d->reruns, L.d_scale, L.star_partial, d->counters, d->queue,
                           stream)); }
  if (g_after_star_event) {     // see tx_set_after_star_bwd_event: a collective may start here, beside the GEMMs that leave SMs idle
    cudaEventRecord(g_after_star_event, st);
    ++g_after_star_count;
  }
  // dW_fk = d(ft)^T z with the 2 H attention-coefficient columns riding along, then d(attn) = W_h v_h / c
  const int64_t ldc = r4(K);
  { ProfScope ps("gemm_dw", d->tag, st);
    TX_SUB(tx_gemm_tn_f16x3(L.d_hi, L.d_lo, ld16, S.z_hi, S.z_lo, S.ldz16, L.d_scale, S.z_scale, L.splits > 1 ? L.tn_partial : dw_ext, ldc,
                            M * ldc, M, K, n, L.splits, stream));
    if (L.splits > 1 && dw_main) {
      // the F weight rows go straight to their contiguous [F, K] home (a parameter's .grad), the 2 H attention rows to dw_ext
      TX_SUB(tx_reduce_partials_rows(L.tn_partial, L.splits, M * ldc, F, K, ldc, dw_main, K, stream));
      TX_SUB(tx_reduce_partials_rows(L.tn_partial + F * ldc, L.splits, M * ldc, 2 * H, K, ldc, dw_ext + F * ldc, ldc, stream));
    } else {
      if (L.splits > 1) TX_SUB(tx_reduce_partials(L.tn_partial, L.splits, M * ldc, dw_ext, stream));
      if (dw_main && cudaMemcpy2DAsync(dw_main, K * sizeof(float), dw_ext, ldc * sizeof(float), K * sizeof(float), F, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("gat_layer_bwd: copy of the weight gradient failed");
        return TX_ERR_CUDA;
      }
    } }
  { ProfScope ps("tx_attn_grad_from_v", d->tag, st);
    TX_SUB(tx_attn_grad_from_v(d->weight, d->ldw, dw_ext + F * ldc, ldc, H, D, K, L.bounds + 2, dattn_l, dattn_r, stream)); }
  if (dz) {
    // d(z)[:, c0a:K] = d(ft) W[:, c0a:K]; the epilogue applies the derivative of the previous layer's leaky-relu / dropout
    const int64_t c0 = d->dz_from < K ? d->dz_from : K;
    const int64_t c0a = (c0 / 8) * 8;
    const int64_t ldz = r4(K);
    if (K > c0a) {
      tx_gemm_epilogue epi;
      const tx_gemm_epilogue* pe = nullptr;
      if (prev && prev->maskbits && c0a == 0) {
        memset(&epi, 0, sizeof(epi));
        epi.act_mask = reinterpret_cast<const uint8_t*>(prev->maskbits);
        epi.heads = prev->heads; epi.dim = prev->dim; epi.mask_stride = tx_gat_fused_mask_ld(prev->heads, prev->dim); epi.col0 = 0;
        epi.act_slope = prev->act_slope; epi.p_drop = prev->p_next; epi.has_keep_plane = prev->p_next > 0.f ? 1 : 0;
        pe = &epi;
      }
      ProfScope ps("gemm_dz", d->tag, st);
      TX_SUB(tx_gemm_nt_f16x3(L.d_hi, L.d_lo, ld16, reinterpret_cast<const __half*>(S.wt_hi) + c0a * S.ldwt,
                              reinterpret_cast<const __half*>(S.wt_lo) + c0a * S.ldwt, S.ldwt, L.d_scale, S.w_scale, dz + c0a, ldz, n, K - c0a,
                              F, pe, L.dz_amax, stream));
    }
    if (dz_amax_out) *dz_amax_out = L.dz_amax;
  }
  return TX_OK;
}

// =================================================================================================================================
// GCN layer (model_zoo.py:34-50)
// =================================================================================================================================
}  // extern "C"

namespace tx {

struct GcnFwdLayout {
  __half *z_hi, *z_lo; float *z_scale, *z_amax;
  __half *w_hi, *w_lo, *wt_hi, *wt_lo; float* w_scal;
  float *y, *y_amax; uint32_t* maskbits; __half *out_hi, *out_lo; float *out_scale, *bound, *scal;
  size_t bytes;
};
static GcnFwdLayout carve_gcn_fwd(const tx_gcn_layer_desc& d, int split_input, void* ws) {
  Carver c(ws);
  GcnFwdLayout L;
  const int64_t D = d.dim, K = d.k;
  L.z_hi = c.take<__half>(split_input ? d.n * r8(K) : 0);
  L.z_lo = c.take<__half>(split_input ? d.n * r8(K) : 0);
  L.scal = c.take<float>(kScalars);
  L.w_scal = L.scal; L.z_scale = L.scal + 4; L.z_amax = L.scal + 5; L.y_amax = L.scal + 6; L.out_scale = L.scal + 7; L.bound = L.scal + 8;
  L.w_hi = c.take<__half>(K * r8(D));        // [K, D]: the d(z) operand
  L.w_lo = c.take<__half>(K * r8(D));
  L.wt_hi = c.take<__half>(D * r8(K));       // [D, K]: the forward operand
  L.wt_lo = c.take<__half>(D * r8(K));
  L.y = c.take<float>(d.n * r4(D));
  const bool mask = d.hidden && (d.act_slope != 1.f || d.p_next > 0.f);
  L.maskbits = c.take<uint32_t>(mask ? tx_gat_fused_mask_words(d.n, 1, D) : 0);
  if (!mask) L.maskbits = nullptr;
  const int64_t ld16 = r8(D + d.pos_dim);
  L.out_hi = c.take<__half>(d.hidden ? d.n * ld16 : 0);
  L.out_lo = c.take<__half>(d.hidden ? d.n * ld16 : 0);
  L.bytes = c.off;
  return L;
}
struct GcnBwdLayout {
  float *pos_partial, *col_partial, *bound, *g_amax; __half *d_hi, *d_lo; float *d_scale, *tn_partial, *dz_amax, *scal;
  int64_t splits;
  size_t bytes;
};
static GcnBwdLayout carve_gcn_bwd(const tx_gcn_layer_desc& d, void* ws) {
  Carver c(ws);
  GcnBwdLayout L;
  const int64_t D = d.dim, K = d.k;
  L.pos_partial = c.take<float>(d.hidden && d.pos_dim > 0 ? tx_row_blocks(d.n) * d.vocab * d.pos_dim : 0);
  L.col_partial = c.take<float>(d.bias ? tx_row_blocks(d.n) * D : 0);
  L.scal = c.take<float>(kScalars);
  L.bound = L.scal; L.g_amax = L.scal + 1; L.d_scale = L.scal + 2; L.dz_amax = L.scal + 3;
  L.d_hi = c.take<__half>(d.n * r8(D));
  L.d_lo = c.take<__half>(d.n * r8(D));
  L.splits = tx_gemm_tn_f16_splits(D, K, d.n);
  L.tn_partial = c.take<float>(L.splits > 1 ? L.splits * D * r4(K) : 0);
  L.bytes = c.off;
  return L;
}
static int check_gcn_desc(const tx_gcn_layer_desc* d, const char* who) {
  TX_REQUIRE(d, "%s: null descriptor", who);
  TX_REQUIRE(d->n > 0 && d->k > 0 && d->dim > 0 && d->dim % 4 == 0 && d->pos_dim >= 0, "%s: bad sizes", who);
  TX_REQUIRE(d->weight && d->ldw >= d->dim && d->norm && d->in_ptr && d->in_src && d->out_ptr && d->out_dst, "%s: parameters / structure missing", who);
  TX_REQUIRE(!(d->hidden && d->pos_dim > 0) || (d->next_pos_table && d->pos && d->vocab > 0), "%s: next position table / positions missing", who);
  return TX_OK;
}

}  // namespace tx

extern "C" {

int64_t tx_gcn_layer_fwd_bytes(const tx_gcn_layer_desc* d, int32_t split_input) { return d ? (int64_t)carve_gcn_fwd(*d, split_input, nullptr).bytes : -1; }
int64_t tx_gcn_layer_bwd_bytes(const tx_gcn_layer_desc* d) { return d ? (int64_t)carve_gcn_bwd(*d, nullptr).bytes : -1; }

int tx_gcn_layer_fwd(const tx_gcn_layer_desc* d, const float* z, int64_t ldz, const tx_gat_layer_state* prev, void* workspace,
                     tx_gat_layer_state* state, float* out, void* stream) {
  TX_SUB(check_gcn_desc(d, "gcn_layer_fwd"));
  TX_REQUIRE(workspace && state && aligned16(workspace), "gcn_layer_fwd: workspace / state missing");
  TX_REQUIRE((z != nullptr) != (prev != nullptr), "gcn_layer_fwd: exactly one of z (fp32 input) / prev (the previous layer's fp16-pair output)");
  TX_REQUIRE(d->hidden || out, "gcn_layer_fwd: the output layer needs `out`");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = d->n, K = d->k, D = d->dim, pd = d->hidden ? d->pos_dim : 0;
  const GcnFwdLayout L = carve_gcn_fwd(*d, z != nullptr, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("gcn_layer_fwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  tx_gat_layer_state& S = *state;
  memset(&S, 0, sizeof(S));
  if (z) {
    { ProfScope ps("tx_absmax", d->tag, st); TX_SUB(tx_absmax(z, ldz, n, K, L.z_amax, stream)); }
    { ProfScope ps("tx_split_f16", d->tag, st); TX_SUB(tx_split_f16(z, ldz, n, K, L.z_amax, L.z_hi, L.z_lo, r8(K), L.z_scale, stream)); }
    S.z_hi = L.z_hi; S.z_lo = L.z_lo; S.z_scale = L.z_scale; S.ldz16 = r8(K);
  } else {
    TX_REQUIRE(prev->out_hi && prev->out_lo && prev->out_scale && prev->ld16_out >= r8(K), "gcn_layer_fwd: the previous layer published no fp16 pair");
    S.z_hi = prev->out_hi; S.z_lo = prev->out_lo; S.z_scale = prev->out_scale; S.ldz16 = prev->ld16_out;
  }
  { ProfScope ps("tx_split_f16_weight", d->tag, st);
    TX_SUB(tx_split_f16_weight(d->weight, d->ldw, K, D, L.w_hi, L.w_lo, r8(D), L.wt_hi, L.wt_lo, r8(K), L.w_scal, L.w_scal + 2, stream)); }
  S.wt_hi = L.w_hi; S.wt_lo = L.w_lo; S.w_scale = L.w_scal + 2; S.ldwt = r8(D);       // [K, D]: rows c0a.. are the d(z) operand
  { ProfScope ps("gemm_fwd", d->tag, st);                                       // y = torch.mm(h, W), model_zoo.py:37
    TX_SUB(tx_gemm_nt_f16x3(S.z_hi, S.z_lo, S.ldz16, L.wt_hi, L.wt_lo, r8(K), S.z_scale, S.w_scale, L.y, r4(D), n, D, K, nullptr, L.y_amax, stream)); }
  S.ft = L.y; S.ft_amax = L.y_amax;
  tx_gat_epilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.mean_heads = d->hidden ? 0 : 1;
  epi.act_slope = d->act_slope;
  epi.next_pos_table = pd > 0 ? d->next_pos_table : nullptr;
  epi.pos = pd > 0 ? d->pos : nullptr;
  epi.pos_dim = pd;
  epi.p_drop = d->hidden ? d->p_next : 0.f;
  epi.seed = d->next_seed;
  epi.stream_id = d->next_stream;
  if (d->hidden) {
    const float keep = 1.f - d->p_next;
    const int64_t ld16 = r8(D + pd);
    { ProfScope ps("tx_bound_gcn", d->tag, st);
      TX_SUB(tx_bound_gcn(L.y_amax, sqrtf((float)(d->max_in_deg > 1 ? d->max_in_deg : 1)) / keep, d->bias, d->bias ? D : 0, 1.f / keep,
                          pd > 0 ? d->next_pos_table : nullptr, pd > 0 ? d->vocab * pd : 0, 1.f / keep, L.bound, stream)); }
    { ProfScope ps("tx_gcn_aggregate_fwd", d->tag, st);
      TX_SUB(tx_gcn_aggregate_fwd_f16(L.y, r4(D), d->norm, d->bias, d->in_ptr, d->in_src, n, D, r4(D + pd), &epi, L.out_hi, L.out_lo, ld16,
                                      L.bound, L.out_scale, L.maskbits, stream)); }
    S.out_hi = L.out_hi; S.out_lo = L.out_lo; S.out_scale = L.out_scale; S.ld16_out = ld16; S.maskbits = L.maskbits;
  } else {
    ProfScope ps("tx_gcn_aggregate_fwd", d->tag, st);
    TX_SUB(tx_gcn_aggregate_fwd(L.y, r4(D), d->norm, d->bias, d->in_ptr, d->in_src, n, D, out, D, &epi, stream));
  }
  S.heads = 1; S.dim = D; S.act_slope = d->act_slope; S.p_next = d->hidden ? d->p_next : 0.f;
  return TX_OK;
}

int tx_gcn_layer_bwd(const tx_gcn_layer_desc* d, const tx_gat_layer_state* state, const tx_gat_layer_state* prev, const float* dout,
                     int64_t ldg, const float* g_amax, void* workspace, float* dz, float* dwt, float* dbias, float* dtab,
                     float** dz_amax_out, void* stream) {
  TX_SUB(check_gcn_desc(d, "gcn_layer_bwd"));
  TX_REQUIRE(state && workspace && aligned16(workspace) && dout && dwt, "gcn_layer_bwd: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const tx_gat_layer_state& S = *state;
  const int64_t n = d->n, K = d->k, D = d->dim, pd = d->hidden ? d->pos_dim : 0;
  const GcnBwdLayout L = carve_gcn_bwd(*d, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("gcn_layer_bwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  if (d->hidden && pd > 0 && dtab) {
    { ProfScope ps("tx_pos_grad_partials", d->tag, st);
      TX_SUB(tx_pos_grad_partials(dout, ldg, D, d->pos, n, pd, d->vocab, d->p_next, d->next_seed, d->next_stream, L.pos_partial, stream)); }
    { ProfScope ps("tx_reduce_partials", d->tag, st); TX_SUB(tx_reduce_partials(L.pos_partial, tx_row_blocks(n), d->vocab * pd, dtab, stream)); }
  }
  if (d->bias && dbias) {                                                      // d(bias) = column sums of the gradient (model_zoo.py:47)
    { ProfScope ps("tx_colsum_partials", d->tag, st); TX_SUB(tx_colsum_partials(dout, ldg, n, D, L.col_partial, stream)); }
    { ProfScope ps("tx_reduce_partials", d->tag, st); TX_SUB(tx_reduce_partials(L.col_partial, tx_row_blocks(n), D, dbias, stream)); }
  }
  if (!g_amax) {
    ProfScope ps("tx_absmax", d->tag, st);
    TX_SUB(tx_absmax(dout, ldg, n, D, L.g_amax, stream));
    g_amax = L.g_amax;
  }
  { ProfScope ps("tx_bound_gcn", d->tag, st);                                  // |dy_j| <= out-degree max|g|
    TX_SUB(tx_bound_gcn(g_amax, 0.5f * (float)(d->max_out_deg > 1 ? d->max_out_deg : 1), nullptr, 0, 0.f, nullptr, 0, 0.f, L.bound, stream)); }
  const int64_t ld16 = r8(D);
  { ProfScope ps("tx_gcn_aggregate_bwd", d->tag, st);
    TX_SUB(tx_gcn_aggregate_bwd_f16(dout, ldg, d->norm, d->out_ptr, d->out_dst, n, D, L.d_hi, L.d_lo, ld16, L.bound, L.d_scale, stream)); }
  const int64_t ldc = r4(K);
  { ProfScope ps("gemm_dw", d->tag, st);                                       // dW^T [D, K] = d(y)^T z
    TX_SUB(tx_gemm_tn_f16x3(L.d_hi, L.d_lo, ld16, S.z_hi, S.z_lo, S.ldz16, L.d_scale, S.z_scale, L.splits > 1 ? L.tn_partial : dwt, ldc, D * ldc,
                            D, K, n, L.splits, stream));
    if (L.splits > 1) TX_SUB(tx_reduce_partials(L.tn_partial, L.splits, D * ldc, dwt, stream)); }
  if (dz) {
    const int64_t c0 = d->dz_from < K ? d->dz_from : K;
    const int64_t c0a = (c0 / 8) * 8;
    const int64_t ldz = r4(K);
    if (K > c0a) {
      tx_gemm_epilogue epi;
      const tx_gemm_epilogue* pe = nullptr;
      if (prev && prev->maskbits && c0a == 0) {
        memset(&epi, 0, sizeof(epi));
        epi.act_mask = reinterpret_cast<const uint8_t*>(prev->maskbits);
        epi.heads = prev->heads; epi.dim = prev->dim; epi.mask_stride = tx_gat_fused_mask_ld(prev->heads, prev->dim); epi.col0 = 0;
        epi.act_slope = prev->act_slope; epi.p_drop = prev->p_next; epi.has_keep_plane = prev->p_next > 0.f ? 1 : 0;
        pe = &epi;
      }
      ProfScope ps("gemm_dz", d->tag, st);
      TX_SUB(tx_gemm_nt_f16x3(L.d_hi, L.d_lo, ld16, reinterpret_cast<const __half*>(S.wt_hi) + c0a * S.ldwt,
                              reinterpret_cast<const __half*>(S.wt_lo) + c0a * S.ldwt, S.ldwt, L.d_scale, S.w_scale, dz + c0a, ldz, n, K - c0a,
                              D, pe, L.dz_amax, stream));
    }
    if (dz_amax_out) *dz_amax_out = L.dz_amax;
  }
  return TX_OK;
}

// =================================================================================================================================
// Readout + bilinear matching (model.py:85-86)
// =================================================================================================================================
}  // extern "C"

namespace tx {

struct HeadFwdLayout { float *hg, *hg_amax; __half *hg_hi, *hg_lo; float* hg_scale; __half *w_hi, *w_lo, *wt_hi, *wt_lo; float *w_scal, *u, *scal; size_t bytes; };
static HeadFwdLayout carve_head_fwd(const tx_head_desc& d, void* ws) {
  Carver c(ws);
  HeadFwdLayout L;
  const int64_t l = d.dim, r = d.r;
  L.hg = c.take<float>(d.g * l);
  L.scal = c.take<float>(kScalars);
  L.w_scal = L.scal; L.hg_amax = L.scal + 4; L.hg_scale = L.scal + 5;
  L.hg_hi = c.take<__half>(d.g * r8(l));
  L.hg_lo = c.take<__half>(d.g * r8(l));
  L.w_hi = c.take<__half>(l * r8(r));          // [l, r]: the d(hg) operand
  L.w_lo = c.take<__half>(l * r8(r));
  L.wt_hi = c.take<__half>(r * r8(l));         // [r, l]: the forward operand
  L.wt_lo = c.take<__half>(r * r8(l));
  L.u = c.take<float>(d.g * r4(r));
  L.bytes = c.off;
  return L;
}
struct HeadBwdLayout { float *du, *du_amax; __half *du_hi, *du_lo; float *du_scale, *dhg, *tn_partial, *pw_partial, *dhg_amax, *scal; int64_t splits; size_t bytes; };
static HeadBwdLayout carve_head_bwd(const tx_head_desc& d, void* ws) {
  Carver c(ws);
  HeadBwdLayout L;
  const int64_t l = d.dim, r = d.r;
  L.du = c.take<float>(d.g * r);
  L.scal = c.take<float>(kScalars);
  L.du_amax = L.scal; L.du_scale = L.scal + 1; L.dhg_amax = L.scal + 2;
  L.du_hi = c.take<__half>(d.g * r8(r));
  L.du_lo = c.take<__half>(d.g * r8(r));
  L.dhg = c.take<float>(d.g * r4(l));
  L.splits = tx_gemm_tn_f16_splits(l, r, d.g);
  L.tn_partial = c.take<float>(L.splits > 1 ? L.splits * l * r4(r) : 0);
  L.pw_partial = c.take<float>(d.kind == TX_READOUT_WMEAN ? tx_readout_bwd_blocks(d.g) * 3 : 0);
  L.bytes = c.off;
  return L;
}
static int check_head_desc(const tx_head_desc* d, const char* who) {
  TX_REQUIRE(d, "%s: null descriptor", who);
  TX_REQUIRE(d->n > 0 && d->g > 0 && d->dim > 0 && d->r > 0, "%s: bad sizes", who);
  TX_REQUIRE(d->kind == TX_READOUT_MEAN || d->kind == TX_READOUT_WMEAN, "%s: mean / weighted-mean readout only", who);
  TX_REQUIRE(d->w && d->ldw >= d->r && d->node_off && (d->kind == TX_READOUT_MEAN || (d->pos && d->pos_weight)), "%s: parameters / structure missing", who);
  return TX_OK;
}

}  // namespace tx

extern "C" {

int64_t tx_head_fwd_bytes(const tx_head_desc* d) { return d ? (int64_t)carve_head_fwd(*d, nullptr).bytes : -1; }
int64_t tx_head_bwd_bytes(const tx_head_desc* d) { return d ? (int64_t)carve_head_bwd(*d, nullptr).bytes : -1; }

int tx_head_fwd(const tx_head_desc* d, const float* h, int64_t ldh, const float* q, int64_t ldq, void* workspace, tx_head_state* state,
                float* scores, void* stream) {
  TX_SUB(check_head_desc(d, "head_fwd"));
  TX_REQUIRE(h && q && workspace && state && scores && aligned16(workspace), "head_fwd: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t G = d->g, l = d->dim, r = d->r;
  const HeadFwdLayout L = carve_head_fwd(*d, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("head_fwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  { ProfScope ps("tx_readout_fwd", d->tag, st);
    TX_SUB(tx_readout_fwd(d->kind, h, ldh, d->pos, d->pos_weight, d->node_off, G, l, L.hg, l, stream)); }
  { ProfScope ps("tx_absmax", d->tag, st); TX_SUB(tx_absmax(L.hg, l, G, l, L.hg_amax, stream)); }
  { ProfScope ps("tx_split_f16", d->tag, st); TX_SUB(tx_split_f16(L.hg, l, G, l, L.hg_amax, L.hg_hi, L.hg_lo, r8(l), L.hg_scale, stream)); }
  { ProfScope ps("tx_split_f16_weight", d->tag, st);
    TX_SUB(tx_split_f16_weight(d->w, d->ldw, l, r, L.w_hi, L.w_lo, r8(r), L.wt_hi, L.wt_lo, r8(l), L.w_scal, L.w_scal + 2, stream)); }
  { ProfScope ps("tx_gemm_nt_f16x3", d->tag, st);                               // u = hg W
    TX_SUB(tx_gemm_nt_f16x3(L.hg_hi, L.hg_lo, r8(l), L.wt_hi, L.wt_lo, r8(l), L.hg_scale, L.w_scal + 2, L.u, r4(r), G, r, l, nullptr, nullptr, stream)); }
  { ProfScope ps("tx_match_rowdot_fwd", d->tag, st); TX_SUB(tx_match_rowdot_fwd(L.u, r4(r), q, ldq, G, r, d->apply_exp, scores, stream)); }
  tx_head_state& S = *state;
  S.hg = L.hg; S.hg_hi = L.hg_hi; S.hg_lo = L.hg_lo; S.hg_scale = L.hg_scale; S.w_hi = L.w_hi; S.w_lo = L.w_lo; S.w_scale = L.w_scal + 2;
  S.u = L.u; S.scores = scores;
  return TX_OK;
}

int tx_head_bwd(const tx_head_desc* d, const tx_head_state* state, const float* h, int64_t ldh, const float* q, int64_t ldq,
                const float* dscores, void* workspace, float* dh, float* dw, float* dw_main, float* dpos_weight, float** dh_amax_out,
                void* stream) {
  TX_SUB(check_head_desc(d, "head_bwd"));
  TX_REQUIRE(state && h && q && dscores && workspace && dh && (dw || dw_main) && aligned16(workspace), "head_bwd: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const tx_head_state& S = *state;
  const int64_t G = d->g, l = d->dim, r = d->r;
  const HeadBwdLayout L = carve_head_bwd(*d, workspace);
  if (cudaMemsetAsync(L.scal, 0, kScalars * sizeof(float), st) != cudaSuccess) { set_error("head_bwd: memset failed"); return TX_ERR_CUDA; }
  PreclearScope preclear;
  { ProfScope ps("tx_match_rowdot_bwd", d->tag, st);
    TX_SUB(tx_match_rowdot_bwd(S.u, r4(r), q, ldq, S.scores, dscores, G, r, d->apply_exp, L.du, r, nullptr, 0, stream)); }
  { ProfScope ps("tx_absmax", d->tag, st); TX_SUB(tx_absmax(L.du, r, G, r, L.du_amax, stream)); }
  { ProfScope ps("tx_split_f16", d->tag, st); TX_SUB(tx_split_f16(L.du, r, G, r, L.du_amax, L.du_hi, L.du_lo, r8(r), L.du_scale, stream)); }
  { ProfScope ps("tx_gemm_nt_f16x3", d->tag, st);                               // d(hg) = d(u) W^T
    TX_SUB(tx_gemm_nt_f16x3(L.du_hi, L.du_lo, r8(r), S.w_hi, S.w_lo, r8(r), L.du_scale, S.w_scale, L.dhg, r4(l), G, l, r, nullptr, nullptr, stream)); }
  const int64_t ldc = r4(r);
  { ProfScope ps("tx_gemm_tn_f16x3", d->tag, st);                               // dW = hg^T d(u)
    // dw: [l, round4(r)] pitched; dw_main (optional): the contiguous [l, r] home of the gradient (a parameter's .grad)
    TX_REQUIRE(dw || L.splits > 1, "head_bwd: dw (pitched) is required when the weight-gradient GEMM does not run split-K");
    TX_SUB(tx_gemm_tn_f16x3(S.hg_hi, S.hg_lo, r8(l), L.du_hi, L.du_lo, r8(r), S.hg_scale, L.du_scale, L.splits > 1 ? L.tn_partial : dw, ldc, l * ldc,
                            l, r, G, L.splits, stream));
    if (L.splits > 1 && dw_main) {
      TX_SUB(tx_reduce_partials_rows(L.tn_partial, L.splits, l * ldc, l, r, ldc, dw_main, r, stream));
    } else {
      if (L.splits > 1) TX_SUB(tx_reduce_partials(L.tn_partial, L.splits, l * ldc, dw, stream));
      if (dw_main && cudaMemcpy2DAsync(dw_main, r * sizeof(float), dw, ldc * sizeof(float), r * sizeof(float), l, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("head_bwd: copy of the weight gradient failed");
        return TX_ERR_CUDA;
      }
    } }
  { ProfScope ps("tx_readout_bwd", d->tag, st);
    TX_SUB(tx_readout_bwd(d->kind, L.dhg, r4(l), h, ldh, S.hg, l, d->pos, d->pos_weight, d->node_off, G, l, dh, l,
                          d->kind == TX_READOUT_WMEAN && dpos_weight ? L.pw_partial : nullptr, stream)); }
  if (d->kind == TX_READOUT_WMEAN && dpos_weight) {
    ProfScope ps("tx_reduce_partials", d->tag, st);
    TX_SUB(tx_reduce_partials(L.pw_partial, tx_readout_bwd_blocks(G), 3, dpos_weight, stream));
  }
  // every readout weight a_i / S, 1 / n_g is <= 1, so max|d(h)| <= max|d(hg)|: a 16 MB reduction instead of one over the 74 MB d(h)
  { ProfScope ps("tx_absmax", d->tag, st); TX_SUB(tx_absmax(L.dhg, r4(l), G, l, L.dhg_amax, stream)); }
  if (dh_amax_out) *dh_amax_out = L.dhg_amax;
  return TX_OK;
}

int tx_set_after_star_bwd_event(void* cuda_event) {
  g_after_star_event = reinterpret_cast<cudaEvent_t>(cuda_event);
  return TX_OK;
}
int64_t tx_after_star_bwd_event_count(void) { return g_after_star_count; }

// ---- launch accounting and per-launch timing of the calls above ----
int64_t tx_layer_launches(int32_t reset) {
  const int64_t v = g_sub_launches;
  if (reset) g_sub_launches = 0;
  return v;
}
void tx_prof_enable(int32_t on) { g_prof_on = on != 0; }
void tx_prof_clear(void) {
  for (auto& e : g_prof) { g_event_pool.push_back(e.e0); g_event_pool.push_back(e.e1); }
  g_prof.clear();
}
int64_t tx_prof_count(void) { return (int64_t)g_prof.size(); }
// after the stream has been synchronised: name / tag (<= 63 characters each) and milliseconds of entry i
int tx_prof_get(int64_t i, char* name64, char* tag64, float* ms) {
  TX_REQUIRE(i >= 0 && i < (int64_t)g_prof.size() && name64 && tag64 && ms, "prof_get: bad index");
  const ProfEntry& e = g_prof[(size_t)i];
  strncpy(name64, e.name.c_str(), 63); name64[63] = 0;
  strncpy(tag64, e.tag.c_str(), 63); tag64[63] = 0;
  if (cudaEventElapsedTime(ms, e.e0, e.e1) != cudaSuccess) { set_error("prof_get: events not complete"); return TX_ERR_CUDA; }
  return TX_OK;
}

}  // extern "C"
