/*
 * taxo_b200.h -- C ABI of libtaxo_sm100.so: TaxoExpan's position-enhanced graph propagation + readout
 * hot path as hand-written CUDA for sm_100a (B200).
 *
 * The reference (mickeysjm/TaxoExpan) has no FFI: its boundary is Python duck typing
 * (model/model.py:16-67,83-86).  The arithmetic of this path lives in DGL 0.4.0's message-passing kernels
 * and torch ops called from model/model_zoo.py; every entry point below cites the reference call site
 * (file:line under the reference root) it replaces.  taxoexpan_b200/model_zoo.py mirrors the reference's
 * module surface on top of this ABI; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; float = IEEE fp32; indices int32.
 *   - caller owns every buffer (inputs, outputs, workspace); the library never allocates, frees or
 *     synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - returns 0 on success, a negative TX_ERR_* code otherwise; tx_last_error() gives the message of the
 *     calling thread's last failure.
 *   - row-major matrices with an explicit leading dimension `ld*` counted in floats.  Kernels use 128-bit
 *     accesses when pointers are 16-byte aligned and ld % 4 == 0, scalar accesses otherwise.
 *   - graphs are CSR by DESTINATION (in_ptr[N+1], in_src[E], in_eid[E]: in-edges of node i in edge-id
 *     order) and CSR by SOURCE (out_ptr[N+1], out_dst[E], out_slot[E]: out_slot = position of that edge in
 *     the destination-sorted order).  Per-edge arrays (alpha, elog, ...) are stored in destination-sorted
 *     order ("slots").
 *   - dropout: keep(seed, stream_id, index) is a counter-based hash draw (16-bit uniform >= round(p * 65536);
 *     tx_common.cuh), a pure function of its arguments: reproducible in the backward pass and by
 *     tx_dropout_keep_mask; drop(x) = keep ? x / (1 - p) : 0 (torch.nn.Dropout
 *     semantics, reference model/model_zoo.py:57-64).  index = row * ld + col for feature matrices and
 *     eid * H + head for attention coefficients.
 */
#ifndef TAXO_B200_H_
#define TAXO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TX_ABI_VERSION 1

#define TX_OK 0
#define TX_ERR_INVALID_ARGUMENT (-1)
#define TX_ERR_CUDA (-2)
#define TX_ERR_UNSUPPORTED (-3)

/* activation applied by an aggregate epilogue: v > 0 ? v : slope * v.  slope = 1 means "no activation"
 * (output layers, reference model_zoo.py:126,152,189,219); hidden layers use F.leaky_relu's default 0.01
 * (model/model.py:25,30,35,40). */

int tx_abi_version(void);
const char* tx_last_error(void);
/* Name of the device kernels are compiled for ("sm_100a"). */
const char* tx_target_arch(void);
/* Programmatic dependent launch of the hot-path kernels (launch + prologue of kernel k+1 overlap the tail of kernel k; every kernel
 * waits for its predecessor before it touches global memory, so results do not change): on by default, TAXO_PDL=0 in the environment or
 * tx_pdl_set(0) switches it off.  Returns the previous setting. */
int tx_pdl_set(int enabled);

/* ------------------------------------------------------------------------------------------------
 * Graph structure (replaces DGL's graph index built by dgl.batch, data_loader/data_loaders.py:25, and the
 * lazy CSR copies DGL makes on first kernel use).
 * tx_build_csr: counting sort of E (src,dst) pairs by key (stable: ties keep edge-id order).
 *   by_dst: key = dst  -> ptr = in_ptr, nbr = in_src, aux = in_eid
 *   by_src: key = src  -> ptr = out_ptr, nbr = out_dst, aux = out_slot  (needs slot_of_eid from the by_dst pass)
 * workspace: tx_csr_workspace_bytes(N, E) bytes ((N + 1 + E) int32), initialised by the call.
 * ------------------------------------------------------------------------------------------------ */
int tx_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t* bytes);
int tx_build_csr_by_dst(const int32_t* src, const int32_t* dst, int64_t n_nodes, int64_t n_edges,
                        int32_t* in_ptr, int32_t* in_src, int32_t* in_eid, int32_t* slot_of_eid,
                        void* workspace, void* stream);
int tx_build_csr_by_src(const int32_t* src, const int32_t* dst, const int32_t* slot_of_eid, int64_t n_nodes,
                        int64_t n_edges, int32_t* out_ptr, int32_t* out_dst, int32_t* out_slot, void* workspace,
                        void* stream);
/* Closed-form structure of a batch of star egonets laid out as data_loader/dataset.py:404-437 does
 * (nodes [gp.., anchor, sib..]; edges [gp->anchor.., anchor->sib.., self loops]) from per-egonet counts.
 * node_off/edge_off are exclusive prefix sums [G+1] of n = n_gp+1+n_sib and e = 2n-1.  Writes pos, the
 * edge list in edge-id order and both CSRs without any sort. Any output pointer may be NULL. */
int tx_star_batch_structure(const int32_t* n_gp, const int32_t* n_sib, const int32_t* node_off,
                            const int32_t* edge_off, int64_t n_graphs, int32_t* pos, int32_t* src, int32_t* dst,
                            int32_t* in_ptr, int32_t* in_src, int32_t* in_eid, int32_t* out_ptr, int32_t* out_dst,
                            int32_t* out_slot, void* stream);
/* out[i, :d] = table[ids[i], :d] (ids clamped to [0, n_table)): a batch's feature rows from the node-embedding table resident in HBM
 * (the reference collates them per egonet from g_full.ndata['x'] on the host, data_loader/dataset.py:157,429-431). */
int tx_gather_rows(const float* table, int64_t ldt, int64_t n_table, const int32_t* ids, int64_t n, int64_t d, float* out, int64_t ldo,
                   void* stream);
/* The plan of a star-egonet batch from the counts alone (one launch; egonets in size-class order, longest work items first): node_off / edge_off [G + 1] (exclusive scans of n = n_gp + 1 + n_sib
 * and 2 n - 1, the layout of dataset.py:404-437 batched as data_loaders.py:25 does) and the work-item tables of tx_gat_star_fwd /
 * tx_gat_star_bwd: one 16-byte record {first node, first edge, n_gp | chunk << 24, n_sib} per (egonet, chunk of chunk_fwd resp.
 * chunk_bwd siblings), sum_k max(1, ceil(n_sib_k / chunk)) records each (the caller sizes them; either table may be NULL). */
int tx_star_batch_plan(const int32_t* n_gp, const int32_t* n_sib, int64_t n_graphs, int64_t chunk_fwd, int64_t chunk_bwd, int32_t* node_off,
                       int32_t* edge_off, int32_t* tasks_fwd, int32_t* tasks_bwd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Position-embedding concat + feature dropout: z = drop([x || P[pos]])
 * replaces nn.Embedding(3,pos_dim)(positions) + torch.cat((h,p),1) + feat_drop / GCN dropout
 * (model/model_zoo.py:146,165-166,201,214-215,82,35-36).  pos_dim = 0 -> z = drop(x) (GAT / GCN stacks).
 * z has ldz >= k_in + pos_dim columns; columns beyond k_in + pos_dim are zero-filled.
 * ------------------------------------------------------------------------------------------------ */
int tx_concat_pos_dropout_fwd(const float* x, int64_t ldx, const float* pos_table, const int32_t* pos,
                              int64_t n_nodes, int64_t k_in, int64_t pos_dim, float* z, int64_t ldz, float p_drop,
                              uint64_t seed, uint32_t stream_id, void* stream);
/* The same, emitted directly as the fp16 hi / lo operand pair (x * scale = hi + lo, scale = a power of two from the bound
 * max(*x_amax, max|pos_table|) / (1 - p_drop); *x_amax = max|x| on the device, e.g. from tx_absmax) that the first layer's projection
 * GEMM reads - z itself is never stored.  Keep decisions are those of tx_concat_pos_dropout_fwd with ldz = round4(k_in + pos_dim).
 * hi / lo: [n_nodes, ld16] halves, ld16 % 8 == 0, columns >= k_in + pos_dim zero-filled; *scale_out receives the scale. */
int tx_concat_pos_dropout_f16(const float* x, int64_t ldx, const float* pos_table, const int32_t* pos, int64_t n_nodes, int64_t k_in,
                              int64_t pos_dim, int64_t vocab, float p_drop, uint64_t seed, uint32_t stream_id, const float* x_amax,
                              void* hi, void* lo, int64_t ld16, float* scale_out, void* stream);

/* Backward of an "activation -> [.|| P[pos]] -> dropout" epilogue, in place:
 *   dz[i,c] *= keep(i,c)/(1-p) * (c < k_in ? act'(z[i,c]) : 1)          (act' = 1 where z > 0 else slope)
 *   dpos_partial[b, r, :] = sum over rows i of block b with pos_i = r of dz[i, k_in : k_in+pos_dim]
 * z is the saved forward output of the epilogue (its sign is the activation's sign); pass z = NULL or
 * slope = 1 for "no activation" (the layer-0 concat).  dpos_partial has n_blocks = tx_row_blocks(n_nodes)
 * blocks of 3*pos_dim floats, reduced by tx_reduce_partials. */
int64_t tx_row_blocks(int64_t n_rows);
int tx_epilogue_bwd(float* dz, int64_t ldz, const float* z, const int32_t* pos, int64_t n_nodes, int64_t k_in,
                    int64_t pos_dim, int64_t vocab, float slope, float p_drop, uint64_t seed, uint32_t stream_id,
                    float* dpos_partial, void* stream);
/* out[m] = sum_b partial[b*m_len + m]  (fixed order: deterministic). */
int tx_reduce_partials(const float* partial, int64_t n_blocks, int64_t m_len, float* out, void* stream);
/* The same for pitched [rows, cols] partials (partial[b] starts at b * block_stride, row pitch ld_in) into an output of row pitch ld_out:
 * out[r, c] = sum_b partial[b, r, c], summed in index order (bit-identical to tx_reduce_partials).  Used to land a split-K weight gradient
 * directly in its contiguous home (a parameter's .grad inside a flat gradient bucket) without a strided copy afterwards. */
int tx_reduce_partials_rows(const float* partial, int64_t n_blocks, int64_t block_stride, int64_t rows, int64_t cols, int64_t ld_in,
                            float* out, int64_t ld_out, void* stream);
/* colsum_partial[b, c] = sum over rows of block b of x[i, c]  (GCN bias gradient, model_zoo.py:47). */
int tx_colsum_partials(const float* x, int64_t ldx, int64_t n_rows, int64_t n_cols, float* partial, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAT layer (model/model_zoo.py:80-114), everything after the dense projection ft = fc(h).
 * ------------------------------------------------------------------------------------------------ */
/* a1[i,h] = <ft[i,h,:], attn_l[h,:]>, a2[i,h] = <ft[i,h,:], attn_r[h,:]>           model_zoo.py:84-85 */
int tx_gat_node_logits(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, int64_t n_nodes,
                       int64_t heads, int64_t dim, float* a1, float* a2, void* stream);

typedef struct tx_gat_epilogue {
  /* mean_heads != 0: out[i, :dim] = mean_h agg[i,h,:]  (output layer, model_zoo.py:189,219); no activation.
   * mean_heads == 0: out[i, h*dim + c] = drop(act(agg[i,h,c])) and, when pos_dim > 0,
   *                  out[i, heads*dim + c] = drop(next_pos_table[pos_i, c])  -- i.e. the NEXT layer's
   *                  input z = feat_drop(cat(leaky_relu(flatten(out)), p)) written in one pass
   *                  (model_zoo.py:214-216 followed by :82 of the next layer). Columns up to ldo are zeroed. */
  int32_t mean_heads;
  float act_slope;
  const float* next_pos_table; /* [vocab, pos_dim] or NULL */
  const int32_t* pos;          /* [N] or NULL */
  int64_t pos_dim;
  float p_drop;                /* dropout on the written row (next layer's feat_drop); 0 = none */
  uint64_t seed;
  uint32_t stream_id;
} tx_gat_epilogue;

/* Fused edge attention + edge softmax + attention dropout + weighted aggregation (+ epilogue):
 *   e = leaky_relu(a1[src] + a2[dst], neg_slope)                                    model_zoo.py:90,106-109
 *   alpha = softmax over the in-edges of each destination, per head                 model_zoo.py:112
 *   alpha_d = attn_drop(alpha)                                                      model_zoo.py:114
 *   agg[i,h,:] = sum_e alpha_d[e,h] * ft[src_e,h,:]                                 model_zoo.py:95
 * alpha / elog (post-leaky logits) are written per slot [E, heads] for the backward pass; alpha_d may alias
 * alpha when p_attn == 0.  a1/a2 may be NULL: the kernel then derives the logits from the ft rows itself. */
int tx_gat_aggregate_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const float* a1,
                         const float* a2, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                         int64_t n_nodes, int64_t n_edges, int64_t heads, int64_t dim, float neg_slope,
                         float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha, float* alpha_d,
                         float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, void* stream);

/* Backward, destination phase: for every destination i and head h, with g = d(agg) [N, heads*dim] (ldg):
 *   dalpha_d[e] = <g[i,h,:], ft[src_e,h,:]>;  dalpha = dalpha_d * keep/(1-p)
 *   de = alpha * (dalpha - sum_in(alpha * dalpha));  ds = de * (elog > 0 ? 1 : neg_slope)
 *   writes ds[slot, h] and da2[i,h] = sum_in ds.
 * g_scale multiplies g on load (1/heads for the output layer's mean over heads, model_zoo.py:219). */
int tx_gat_aggregate_bwd_dst(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft,
                             int64_t ldf, const float* alpha, const float* elog, const int32_t* in_ptr,
                             const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads,
                             int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed,
                             uint32_t attn_stream_id, float* ds, float* da2, void* stream);
/* Backward, source phase: dft[j,h,:] = sum_out alpha_d[slot,h] * g[dst,h,:] + da1[j,h]*attn_l[h,:] + da2[j,h]*attn_r[h,:]
 * with da1[j,h] = sum_out ds[slot,h] (written to da1). */
int tx_gat_aggregate_bwd_src(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale,
                             const float* alpha_d, const float* ds, const float* da2, const float* attn_l,
                             const float* attn_r, const int32_t* out_ptr, const int32_t* out_dst,
                             const int32_t* out_slot, int64_t n_nodes, int64_t heads, int64_t dim, float* da1,
                             float* dft, int64_t ldd, void* stream);
/* dattn_l[h,:] = sum_j da1[j,h] ft[j,h,:], dattn_r[h,:] = sum_j da2[j,h] ft[j,h,:] as per-block partials
 * [n_blocks, 2, heads*dim] (n_blocks = tx_row_blocks(n_nodes)); reduce with tx_reduce_partials. */
int tx_gat_attn_grad_partials(const float* ft, int64_t ldf, const float* da1, const float* da2, int64_t n_nodes,
                              int64_t heads, int64_t dim, float* partial, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused fast path for batched graphs (the egonet batches of data_loader/dataset.py:404-437): same arithmetic as
 * tx_gat_node_logits + tx_gat_aggregate_fwd, and tx_epilogue_bwd + tx_gat_aggregate_bwd_dst/_src +
 * tx_gat_attn_grad_partials, in ONE forward and ONE backward kernel with algorithmic DRAM traffic (read ft once, write out
 * once; read g, ft once, write dft once).  Needs dim % 4 == 0, dim <= 512 and, for the output layer
 * (epi->mean_heads), heads == 1 -- tx_gat_fused_supported() says whether a shape qualifies; otherwise use the general
 * kernels above.  The backward kernel needs every edge to stay inside one graph of node_off (true for any dgl.batch).
 *   maskbits: tx_gat_fused_mask_words(N, heads, dim) uint32 words written by the forward epilogue (sign of the
 *             pre-activation and dropout keep bits) and consumed by the backward kernel, which reads g straight from
 *             d(z_next) (ldg = ld of z_next, g_head_stride = dim) and applies keep/(1-p_next) * leaky' on load.
 *             NULL = no activation/dropout to undo (g is used as is).
 *   dattn_partial: [tx_gat_fused_bwd_blocks(N, heads), 2, heads, dim] -> reduce with tx_reduce_partials to
 *             [d attn_l (heads*dim) | d attn_r (heads*dim)].
 *   ds [E*heads], da2 [N*heads]: scratch.
 *   out_lo / dft_lo (optional, same shape and pitch as out / dft): when given, out / dft receive the TF32 "hi" part and
 *             out_lo / dft_lo the "lo" part of every value (the split of tx_split_tf32 done in the producer's registers), so the
 *             3xTF32 GEMMs consume them without a separate split pass.
 * ------------------------------------------------------------------------------------------------ */
int tx_gat_fused_supported(int64_t heads, int64_t dim, int32_t mean_heads);
int64_t tx_gat_fused_mask_words(int64_t n_nodes, int64_t heads, int64_t dim);
int64_t tx_gat_fused_mask_ld(int64_t heads, int64_t dim); /* bytes per row of maskbits: heads*dim/4 rounded up to 16 */
int64_t tx_gat_fused_bwd_blocks(int64_t n_nodes, int64_t heads);
int tx_gat_fused_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                     const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                     float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                     float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                     float* out_lo, void* stream);
int tx_gat_fused_bwd(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const uint32_t* maskbits,
                     int32_t has_keep_plane, float act_slope, float p_next, const float* ft, int64_t ldf,
                     const float* alpha, const float* alpha_d, const float* elog, const float* attn_l,
                     const float* attn_r, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                     const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_slot, const int32_t* node_off,
                     int64_t n_graphs, int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn,
                     uint64_t attn_seed, uint32_t attn_stream_id, float* ds, float* da2, float* dft, int64_t ldd,
                     float* dft_lo, float* dattn_partial, void* stream);
/* TMA-staged variant of tx_gat_fused_bwd (same arithmetic, same outputs; replaces the same reference autograd of
 * model_zoo.py:83-96,106-114) for g that already carries the epilogue derivative (no maskbits): a producer warp bulk-copies
 * the g / ft rows of the next tile of whole graphs into shared memory (cp.async.bulk + mbarrier) while 16 compute warps run
 * the two phases on the current tile out of shared memory.
 *   tiles: int32 [(tx_gat_bwd_num_tiles(N, dim) + 1) * 4] written by tx_gat_bwd_tiles (per tile: first row, first in-edge,
 *          first out-edge, 0; tile t = graphs whose first row lies in [t R, (t+1) R), R = tx_gat_bwd_tile_rows(dim)); depends on
 *          the batch structure and dim only - build once per batch.
 *   dattn_partial: [tx_gat_fused_bwd_staged_blocks(N, heads, dim), 2, heads, dim]. */
int64_t tx_gat_bwd_tile_rows(int64_t dim);
int64_t tx_gat_bwd_num_tiles(int64_t n_nodes, int64_t dim);
int tx_gat_bwd_tiles(const int32_t* node_off, int64_t n_graphs, int64_t n_nodes, const int32_t* in_ptr, const int32_t* out_ptr,
                     int64_t dim, int32_t* tiles, void* stream);
int64_t tx_gat_fused_bwd_staged_blocks(int64_t n_nodes, int64_t heads, int64_t dim);
int tx_gat_fused_bwd_staged(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft, int64_t ldf,
                            const float* alpha, const float* alpha_d, const float* elog, const float* attn_l,
                            const float* attn_r, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                            const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_slot, const int32_t* tiles,
                            int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed,
                            uint32_t attn_stream_id, float* ds, float* da2, float* dft, int64_t ldd, float* dft_lo,
                            void* dft16_hi, void* dft16_lo, int64_t ld16, const float* bound, float* scale_out,
                            float* dattn_partial, void* stream);
/* fp16-split outputs of the fused kernels (operands of tx_gemm_*_f16x3; x * scale = hi + lo, hi / lo fp16 [N, ld16], ld16 % 8 == 0,
 * scale derived on the device from *bound, an upper bound of the written magnitudes, and stored to *scale_out):
 *   tx_gat_fused_bwd_staged: dft16_hi != NULL replaces dft / dft_lo;  bound from tx_bound_dft.
 *   tx_gat_fused_fwd_f16   : a hidden layer's epilogue (the next layer's input) replaces `out`; bound = max(max|ft| / ((1 - p_attn)
 *                            (1 - p_drop)), max|next_pos_table| / (1 - p_drop)) (attention weights are convex).  `ldo` stays the
 *                            LOGICAL fp32 pitch round4(heads*dim + pos_dim): it indexes the dropout counters. */
int tx_gat_fused_fwd_f16(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                         const int32_t* in_src, const int32_t* in_eid, int64_t n_nodes, int64_t heads, int64_t dim,
                         float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                         float* alpha_d, float* elog, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits, void* out16_hi,
                         void* out16_lo, int64_t ld16, const float* bound, float* scale_out, void* stream);
/* TMA-staged variant of tx_gat_fused_fwd / tx_gat_fused_fwd_f16 (same arithmetic and outputs, reference model_zoo.py:83-96,106-114):
 * the ft rows of a tile of whole graphs (tiles of tx_gat_bwd_tiles) are bulk-copied into shared memory one tile ahead; a1 / a2 are
 * computed once per node, the edge softmax by one thread per destination row, aggregation + epilogue by one thread per (row
 * group, pair of float4 columns).  Exactly one of {out (+ optional TF32 out_lo), out16_hi/out16_lo (+ bound, scale_out)} is written. */
int tx_gat_fused_fwd_staged(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* in_ptr,
                            const int32_t* in_src, const int32_t* in_eid, const int32_t* tiles, int64_t n_nodes, int64_t heads,
                            int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id, float* alpha,
                            float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                            float* out_lo, void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out,
                            void* stream);
/* Star-egonet variant of tx_gat_fused_fwd / tx_gat_fused_fwd_f16 (same arithmetic and outputs; reference model_zoo.py:83-96,106-114)
 * for batches whose structure is the closed form of tx_star_batch_structure (data_loader/dataset.py:404-437): no CSR is read, every
 * ft row is loaded once per work item, the anchor row stays in registers while its siblings stream by.
 *   tasks: int32 [4 * n_tasks] (16-byte aligned), one record per (egonet, chunk c):
 *          {node_off, edge_off, n_gp | c << 24, n_sib} of that egonet (the quantities of tx_star_batch_structure); chunk c covers
 *          siblings [c C, (c+1) C) with C = `chunk` (tx_gat_star_chunk() is the recommended value); chunk 0 (present for every
 *          egonet) also owns the grand-parents and the anchor; an egonet has max(1, ceil(n_sib / C)) <= tx_gat_star_max_chunks()
 *          chunks and n_gp < 2^24.
 *   queue: int32 [32 * heads] work-queue counters (one 128-byte line per head), ZERO before the first launch; the kernel leaves them zero again (one buffer per
 *          stream: concurrent launches must not share it).
 *   Exactly one of {out (fp32), out16_hi/out16_lo (+ bound, scale_out)} is written; alpha / alpha_d / elog per slot as usual. */
int64_t tx_gat_star_chunk(void);
int64_t tx_gat_star_max_chunks(void);
int tx_gat_star_fwd(const float* ft, int64_t ldf, const float* attn_l, const float* attn_r, const int32_t* tasks, int64_t n_tasks, int64_t chunk,
                    int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope, float p_attn, uint64_t attn_seed, uint32_t attn_stream_id,
                    float* alpha, float* alpha_d, float* elog, float* out, int64_t ldo, const tx_gat_epilogue* epi, uint32_t* maskbits,
                    void* out16_hi, void* out16_lo, int64_t ld16, const float* bound, float* scale_out, int32_t* queue, void* stream);
/* Star-egonet fused GAT backward (tx_star_bwd.cu; default for EgonetBatch structures, TAXO_STAR_BWD=0 -> tx_gat_fused_bwd_staged):
 * the autograd of model_zoo.py:84-96,106-114 in closed form per egonet (restated in oracle/star_backward.py).  Work items = the
 * (egonet, chunk of `chunk` siblings) records of tx_gat_star_fwd's task table pulled from a self-resetting queue; rows reach each warp
 * through its own TMA ring in shared memory; an egonet with several chunks combines its anchor row through `partial`
 * ([tx_gat_star_bwd_partial_floats(n_tasks, heads, dim)] floats) and `counters` ([n_tasks * heads] int32, zero before the first
 * launch, left zero) in chunk order.  Outputs: d(ft) either as fp32 (dft, ldd; then da1 / da2 [N * heads] are required) or as the fp16
 * hi/lo pair [N, ld16] with ld16 >= heads * dim + 2 * heads: columns heads*dim + h and heads*dim + heads + h carry c * da1[., h] and
 * c * da2[., h] (the per-node coefficients of attn_l / attn_r), so that the weight-gradient GEMM over heads*dim + 2*heads columns also
 * returns v = z^T [c da1 | c da2] and d(attn_l)[h] = W_h v[h] / c, d(attn_r)[h] = W_h v[heads + h] / c (tx_attn_grad_from_v; W_h = rows
 * h*dim.. of the layer's fc weight, because ft_h = z W_h^T).  bounds = the 4 floats written by tx_bound_dft: the kernel first runs with
 * the optimistic scale (bounds[1]); a value outside the fp16 range sets *flag (= bounds + 3, reset by tx_bound_dft) and the second
 * launch, which otherwise exits at once, redoes the pass with the rigorous scale (bounds[0]); *reruns (optional) counts those.
 * ds: scratch [E * heads] (anchors with more than 31 grand-parents). */
int64_t tx_gat_star_bwd_partial_floats(int64_t n_tasks, int64_t heads, int64_t dim);
int tx_gat_star_bwd(const float* g, int64_t ldg, int64_t g_head_stride, float g_scale, const float* ft, int64_t ldf,
                    const float* alpha, const float* alpha_d, const float* elog, const float* attn_l, const float* attn_r,
                    const int32_t* tasks, int64_t n_tasks, int64_t chunk, int64_t n_nodes, int64_t heads, int64_t dim, float neg_slope,
                    float* ds, float* da1, float* da2, float* dft, int64_t ldd, void* dft16_hi, void* dft16_lo, int64_t ld16,
                    const float* bounds, int32_t* flag, int32_t* reruns, float* scale_out, float* partial, int32_t* counters,
                    int32_t* queue, void* stream);
int tx_attn_grad_from_v(const float* weight, int64_t ldw, const float* v, int64_t ldv, int64_t heads, int64_t dim, int64_t k,
                        const float* c, float* dattn_l, float* dattn_r, void* stream);
/* dpos_partial[b, r, :] = sum over rows i of block b (tx_row_blocks) with pos_i = r of dz[i, col0 : col0+pos_dim] * keep/(1-p)
 * (gradient of the appended position-embedding block, reference model_zoo.py:214-215). */
int tx_pos_grad_partials(const float* dz, int64_t ldz, int64_t col0, const int32_t* pos, int64_t n_nodes,
                         int64_t pos_dim, int64_t vocab, float p_drop, uint64_t seed, uint32_t stream_id,
                         float* partial, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GCN layer (model/model_zoo.py:34-50) after y = dropout(h) @ W:
 *   out[i,:] = act(norm_i * sum_in norm_src * y[src,:] + bias)      :39-49, copy_src/sum at :41
 * with the same "next layer input" epilogue as the GAT kernel (heads = 1). norm = in_degree^-0.5 (inf -> 0),
 * model_zoo.py:157-161, computed by tx_gcn_norm from in_ptr.
 * ------------------------------------------------------------------------------------------------ */
int tx_gcn_norm(const int32_t* in_ptr, int64_t n_nodes, float* norm, void* stream);
int tx_gcn_aggregate_fwd(const float* y, int64_t ldy, const float* norm, const float* bias, const int32_t* in_ptr,
                         const int32_t* in_src, int64_t n_nodes, int64_t dim, float* out, int64_t ldo,
                         const tx_gat_epilogue* epi, void* stream);
/* The same aggregate with the fused next-layer epilogue of the GAT path: the hidden layer's output goes out as the NEXT GEMM's fp16 hi/lo
 * operand pair [N, ld16] (x * scale = hi + lo, scale from *bound, see tx_bound_gcn; the fp32 tensor is never written) plus the sign /
 * keep bytes (tx_gat_fused_mask_words(N, 1, dim) words; may be NULL) consumed by the next layer's d(z) GEMM epilogue
 * (tx_gemm_epilogue), and d(y) of the backward aggregate as an fp16 pair as well (columns [dim, ld16) zero).  ldo = the LOGICAL fp32
 * pitch round4(dim + pos_dim): it indexes the dropout counters exactly like tx_gcn_aggregate_fwd. */
int tx_gcn_aggregate_fwd_f16(const float* y, int64_t ldy, const float* norm, const float* bias, const int32_t* in_ptr,
                             const int32_t* in_src, int64_t n_nodes, int64_t dim, int64_t ldo, const tx_gat_epilogue* epi, void* out_hi,
                             void* out_lo, int64_t ld16, const float* bound, float* scale_out, uint32_t* maskbits, void* stream);
int tx_gcn_aggregate_bwd_f16(const float* g, int64_t ldg, const float* norm, const int32_t* out_ptr, const int32_t* out_dst, int64_t n_nodes,
                             int64_t dim, void* dy_hi, void* dy_lo, int64_t ld16, const float* bound, float* scale_out, void* stream);
/* *out = max(2 ca *a, 2 cb max|b[0..b_len)|, ct max|t[0..t_len)|): bound of a GCN epilogue's output (a = max|y|, b = bias, t = the next
 * position table) or, with b = t = NULL, of d(y) (a = max|g|). */
int tx_bound_gcn(const float* a, float ca, const float* b, int64_t b_len, float cb, const float* t, int64_t t_len, float ct, float* out, void* stream);
/* dy[j,:] = norm_j * sum_out norm_dst * g[dst,:] */
int tx_gcn_aggregate_bwd(const float* g, int64_t ldg, const float* norm, const int32_t* out_ptr,
                         const int32_t* out_dst, int64_t n_nodes, int64_t dim, float* dy, int64_t ldd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Readout (model/model_zoo.py:227-258; dgl.mean_nodes / dgl.sum_nodes).
 *   TX_READOUT_MEAN      hg[g,:] = mean_i h[i,:]                                       :231-232
 *   TX_READOUT_WMEAN     a_i = softplus(w[pos_i]); hg = sum a_i h_i / sum a_i            :240-242
 *   TX_READOUT_CONCAT    [sum_{pos=0} h / n_g || sum_{pos=1} h / #anchors || sum_{pos=2} h / n_g]   :248-258
 * node_off[G+1] = exclusive prefix sum of batch_num_nodes.  hg is [G, dim] ([G, 3*dim] for CONCAT).
 * ------------------------------------------------------------------------------------------------ */
#define TX_READOUT_MEAN 0
#define TX_READOUT_WMEAN 1
#define TX_READOUT_CONCAT 2
int tx_readout_fwd(int32_t kind, const float* h, int64_t ldh, const int32_t* pos, const float* pos_weight,
                   const int32_t* node_off, int64_t n_graphs, int64_t dim, float* hg, int64_t ldhg, void* stream);
/* dh[i,:] and, for WMEAN, dw_partial[b, 0..2] for b < tx_readout_bwd_blocks(n_graphs) (reduce with tx_reduce_partials):
 *   S = sum a; dh_i = a_i/S * dhg_g; da_i = <dhg_g, h_i - hg_g>/S; dw[pos_i] += da_i * sigmoid(w[pos_i]). */
int64_t tx_readout_bwd_blocks(int64_t n_graphs);
int tx_readout_bwd(int32_t kind, const float* dhg, int64_t lddhg, const float* h, int64_t ldh, const float* hg,
                   int64_t ldhg, const int32_t* pos, const float* pos_weight, const int32_t* node_off,
                   int64_t n_graphs, int64_t dim, float* dh, int64_t lddh, float* dw_partial, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense projection on tcgen05 tensor cores with fp32-faithful accuracy (3xTF32 split, fp32 accumulation in TMEM):
 *   C[m, n] = sum_k A[m, k] * B[n, k]          (A: [M, K] row-major, B: [N, K] row-major, C: [M, ldc])
 * replaces self.fc(h) / torch.mm(h, W) and their autograd GEMMs (reference model/model_zoo.py:83,37), which the
 * reference runs as cuBLAS/MKL fp32.  Operands are passed PRE-SPLIT: x = hi + lo with hi = rn_tf32(x) and
 * lo = rn_tf32(x - hi) (tx_split_tf32), row pitch a multiple of 4 floats, 16-byte aligned.  Columns
 * [N, round4(N)) of C are written as zeros when ldc allows (padded activations buffers).
 * ------------------------------------------------------------------------------------------------ */
int tx_split_tf32(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ldo, void* stream);
/* Weight gradient form: C[m, n] = sum_r A[r, m] * B[r, n]  (A: [R, M], B: [R, N] row-major, reduction over the R rows =
 * nodes; dW = dft^T . z, the autograd GEMM of model_zoo.py:83,37).  The reduction is split over `splits`
 * (tx_gemm_tn_splits) CTAs per output tile; split s writes its partial [M, ldc] at c_partial + s * split_stride; sum them
 * with tx_reduce_partials (fixed order, deterministic). */
int64_t tx_gemm_tn_splits(int64_t m, int64_t n, int64_t r);
int tx_gemm_tn_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                      float* c_partial, int64_t ldc, int64_t split_stride, int64_t m, int64_t n, int64_t r, int64_t splits,
                      void* stream);
int tx_gemm_nt_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                      float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, void* stream);
/* Same with a fused output transform: for columns c < heads*dim, C[row, c] *= keep ? (positive ? 1 : act_slope) / (1 - p_drop) : 0,
 * decoded from the sign/keep bytes written by tx_gat_fused_fwd (maskbits viewed as bytes; mask_stride =
 * tx_gat_fused_mask_ld(heads, dim) bytes per row).  Used for d(z_next) = d(ft_next) . W_next: the backward of "leaky_relu -> feat_drop"
 * (reference autograd of model_zoo.py:215-216,82) is applied where the gradient is produced, so tx_gat_fused_bwd reads it as is. */
typedef struct tx_gemm_epilogue {
  const uint8_t* act_mask;
  int64_t heads, dim, mask_stride, col0;
  float act_slope, p_drop;
  int32_t has_keep_plane;
} tx_gemm_epilogue;
int tx_gemm_nt_tf32x3_ex(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                         float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, const tx_gemm_epilogue* epilogue, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Test / parity utility: materialise the keep-mask the kernels use (1 = keep) for n indices
 * index = first_index .. first_index + n - 1, so the CPU oracle can be run with the same mask.
 * ------------------------------------------------------------------------------------------------ */
int tx_dropout_keep_mask(uint64_t seed, uint32_t stream_id, int64_t first_index, int64_t n, float p_drop,
                         uint8_t* keep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp16-split variant of the dense projections (same reference call sites: model/model_zoo.py:83,37 and their autograd GEMMs):
 * every fp32 operand x is passed as TWO fp16 arrays with x * scale = hi + lo (hi = rn_f16(x scale), lo = rn_f16(x scale - hi)) and a
 * per-tensor power-of-two `scale` held in a DEVICE float; the product is accumulated as A_lo.B_hi + A_hi.B_lo + A_hi.B_hi by
 * tcgen05.mma.kind::f16 (fp32 accumulation in tensor memory, chunked promotion as for the TF32 form) and the epilogue multiplies by
 * 1 / (scale_a scale_b).  22 significant bits per operand like the 3xTF32 form, at twice the tensor rate and half the operand
 * bytes.  The scale comes from an UPPER BOUND of max|x| (device float, tx_absmax or an analytic bound combined on the device by
 * tx_bound_max2 / tx_bound_dft): |x| scale <= 2^13; accuracy is full while the bound is within 2^16 of the true maximum.
 * Nothing here synchronises with the host.  Row pitches of the fp16 arrays are multiples of 8 elements (16 bytes).
 * ------------------------------------------------------------------------------------------------ */
int tx_absmax(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* out, void* stream);  /* *out = max |x[i, c]| */
/* *out = max(*a ca, max|b[0..b_len)| cb); b (a small parameter table, reduced by the same launch) may be NULL */
int tx_bound_max2(const float* a, float ca, const float* b, int64_t b_len, float cb, float* out, void* stream);
/* out[0] = *g_amax (c_direct + c_attn *ft_amax c), c = max(max|attn_l|, max|attn_r|, 2^-20) over attn_l/attn_r[0..attn_len): rigorous
 * bound of |dft| written by the fused GAT backward (dft_j = sum_i alpha~_ij g_i + da1_j attn_l + da2_j attn_r with |d alpha~| <= dim
 * max|g| max|ft|);  out[1] = min(out[0], *g_amax c_optimistic) (c_optimistic <= 0: out[0]): the scale tx_gat_star_bwd tries first;
 * out[2] = c;  out[3] = 0 as int32: tx_gat_star_bwd's fp16-range flag.  out: 4 floats, 16-byte aligned. */
int tx_bound_dft(const float* g_amax, const float* ft_amax, const float* attn_l, const float* attn_r, int64_t attn_len, float c_direct,
                 float c_attn, float c_optimistic, float* out, void* stream);
/* hi, lo: fp16 [rows, ldo] (columns >= cols zero); *scale_out (optional) = the scale derived from *bound. */
int tx_split_f16(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* bound, void* hi, void* lo, int64_t ldo,
                 float* scale_out, void* stream);
/* One launch for a weight matrix w [rows, cols]: max|w|, then the split both row-major (hi, lo: [rows, ld]) and transposed
 * (hi_t, lo_t: [cols, ld_t]) with the same scale -- the forward projection and the input-gradient GEMM read the same weights in the
 * two orientations (model_zoo.py:83 and its autograd).  scratch2: 2 device floats of workspace. */
int tx_split_f16_weight(const float* w, int64_t ldw, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld, void* hi_t, void* lo_t,
                        int64_t ld_t, float* scratch2, float* scale_out, void* stream);
/* C[m, n] = sum_k A[m, k] B[n, k] / (scale_a scale_b); optional fused mask epilogue (as tx_gemm_nt_tf32x3_ex) and optional
 * *amax_out = max |C| (device float, atomically maximised; zeroed by the call). */
int tx_gemm_nt_f16x3(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                     const float* scale_a, const float* scale_b, float* c, int64_t ldc, int64_t m, int64_t n, int64_t k,
                     const tx_gemm_epilogue* epilogue, float* amax_out, void* stream);
/* Weight-gradient form C[m, n] = sum_r A[r, m] B[r, n] / (scale_a scale_b), split-K partials as tx_gemm_tn_tf32x3. */
int64_t tx_gemm_tn_f16_splits(int64_t m, int64_t n, int64_t r);
int tx_gemm_tn_f16x3(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                     const float* scale_a, const float* scale_b, float* c_partial, int64_t ldc, int64_t split_stride, int64_t m,
                     int64_t n, int64_t r, int64_t splits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One call per GAT layer and direction (tx_layer.cu): the launch sequence of the default hot path - GATLayer.forward and its autograd,
 * model/model_zoo.py:80-114 inside the stacks of :183-190,210-220, on the fp16-pair GEMMs and the star-egonet kernels above - enqueued
 * from native code into ONE caller-owned workspace (tx_gat_layer_fwd_bytes / tx_gat_layer_bwd_bytes; 16-byte aligned; the forward
 * workspace must stay alive until the layer's backward has run, the backward workspace until the layer below has run its backward).
 * Same kernels, arguments and order as the per-kernel calls; nothing allocates, synchronises or reads device data on the host.
 *   forward : ft = z W^T (z: fp32 [N, ldz], split here, or the previous layer's published fp16 pair `prev`), star forward; a hidden layer
 *             publishes the next layer's input pair + sign/keep bytes in `state`, the output layer writes `out` [N, dim].
 *   backward: d(pos table) of the appended rows (hidden), |dft| bounds, star backward, dW (+ 2 heads extra rows) -> dw_ext
 *             [(heads*dim + 2*heads), round4(k)], d(attn_l | attn_r) -> dattn [2 * heads*dim], d(z)[:, dz_from8 : k] -> dz [N, round4(k)]
 *             (dz may be NULL) with the previous layer's activation / dropout derivative applied from `prev` (NULL for the first layer);
 *             g_amax: device bound of max|dout| or NULL (measured here); *dz_amax_out: device max|dz| for the layer below.
 * ------------------------------------------------------------------------------------------------ */
typedef struct tx_gat_layer_desc {
  int64_t n, e, k, heads, dim, pos_dim, vocab, dz_from, max_out_deg;
  int32_t hidden;                         /* 1: hidden layer (emits the next layer's input), 0: output layer (head mean, heads == 1) */
  float neg_slope, p_attn, act_slope, p_next, dft_optimism;
  uint64_t attn_seed, next_seed;
  uint32_t attn_stream, next_stream;
  const int32_t* tasks_fwd; int64_t n_tasks_fwd, chunk_fwd;     /* tx_gat_star_fwd's task table */
  const int32_t* tasks_bwd; int64_t n_tasks_bwd, chunk_bwd;     /* tx_gat_star_bwd's task table */
  const int32_t* pos;
  int32_t *queue, *counters, *reruns;
  const float* weight; int64_t ldw;       /* fc weight [heads*dim, k] */
  const float *attn_l, *attn_r, *next_pos_table;
  char tag[16];                           /* label of the per-launch timings (tx_prof_get) */
} tx_gat_layer_desc;
typedef struct tx_gat_layer_state {       /* filled by tx_gat_layer_fwd; pointers into its workspace (or the previous layer's) */
  void *z_hi, *z_lo; float* z_scale; int64_t ldz16;
  void *wt_hi, *wt_lo; float* w_scale; int64_t ldwt;
  float *ft, *ft_amax, *alpha, *alpha_d, *elog;
  void *out_hi, *out_lo; float* out_scale; int64_t ld16_out; uint32_t* maskbits;
  int64_t heads, dim; float act_slope, p_next;
} tx_gat_layer_state;
int64_t tx_gat_layer_fwd_bytes(const tx_gat_layer_desc* d, int32_t split_input);
int64_t tx_gat_layer_bwd_bytes(const tx_gat_layer_desc* d);
int tx_gat_layer_fwd(const tx_gat_layer_desc* d, const float* z, int64_t ldz, const tx_gat_layer_state* prev, void* workspace,
                     tx_gat_layer_state* state, float* out, void* stream);
/* dw_ext: [heads*dim + 2*heads, round4(k)] (weight rows, then the attention-coefficient rows); dw_main (optional): contiguous
 * [heads*dim, k] - when given the weight rows are written there as well / instead (directly by the split-K reduction), e.g. a
 * parameter's .grad inside a flat gradient bucket; dattn_l / dattn_r: [heads*dim] each; dtab (optional): [vocab, pos_dim]. */
int tx_gat_layer_bwd(const tx_gat_layer_desc* d, const tx_gat_layer_state* state, const tx_gat_layer_state* prev, const float* dout,
                     int64_t ldg, const float* g_amax, void* workspace, float* dz, float* dw_ext, float* dw_main, float* dattn_l,
                     float* dattn_r, float* dtab, float** dz_amax_out, void* stream);
/* The same for a GCN layer (GCNLayer.forward and its autograd, model/model_zoo.py:34-50 inside the stacks of :128-137,155-167): split /
 * weight split / y = z W GEMM / bound / tx_gcn_aggregate_fwd(_f16), and d(position table), d(bias), bound, tx_gcn_aggregate_bwd_f16,
 * dW^T (-> dwt [dim, round4(k)]) , d(z).  Layers hand each other the tx_gat_layer_state (ft = y; alpha / elog unused; heads = 1). */
typedef struct tx_gcn_layer_desc {
  int64_t n, k, dim, pos_dim, vocab, dz_from, max_in_deg, max_out_deg;
  int32_t hidden;                         /* 1: hidden layer (emits the next layer's input), 0: output layer (fp32 [N, dim]) */
  float act_slope, p_next;
  uint64_t next_seed;
  uint32_t next_stream;
  const int32_t *in_ptr, *in_src, *out_ptr, *out_dst, *pos;
  const float* norm;
  const float* weight; int64_t ldw;       /* weight [k, dim] (torch.mm(h, W), model_zoo.py:37) */
  const float *bias, *next_pos_table;
  char tag[16];
} tx_gcn_layer_desc;
int64_t tx_gcn_layer_fwd_bytes(const tx_gcn_layer_desc* d, int32_t split_input);
int64_t tx_gcn_layer_bwd_bytes(const tx_gcn_layer_desc* d);
int tx_gcn_layer_fwd(const tx_gcn_layer_desc* d, const float* z, int64_t ldz, const tx_gat_layer_state* prev, void* workspace,
                     tx_gat_layer_state* state, float* out, void* stream);
int tx_gcn_layer_bwd(const tx_gcn_layer_desc* d, const tx_gat_layer_state* state, const tx_gat_layer_state* prev, const float* dout,
                     int64_t ldg, const float* g_amax, void* workspace, float* dz, float* dwt, float* dbias, float* dtab,
                     float** dz_amax_out, void* stream);
/* Readout + bilinear matching as one call per direction (TaxoExpan.forward lines model/model.py:85-86 and their autograd):
 * hg = readout(h) (MeanReadout / WeightedMeanReadout, model_zoo.py:227-242), u = hg W (the projection half of nn.Bilinear(l, r, 1),
 * model_zoo.py:301-328, on the fp16-pair GEMMs), scores = <u, q> (exp'd for LBM); backward: d(u), d(hg) = d(u) W^T, dW = hg^T d(u),
 * readout backward -> d(h) [N, dim], d(position weights) [3] (WMEAN), and max|d(hg)| (>= max|d(h)|: the bound the output layer's
 * backward needs) at *dh_amax_out.  The forward workspace must stay alive until the backward has run. */
typedef struct tx_head_desc {
  int64_t n, g, dim, r;                   /* nodes, graphs, readout width l = dim, query width r */
  int32_t kind, apply_exp;                /* TX_READOUT_MEAN / TX_READOUT_WMEAN; 1 = LBM */
  const int32_t *pos, *node_off;
  const float* pos_weight;                /* [3] (WMEAN) or NULL */
  const float* w; int64_t ldw;            /* match.W.weight[0]: [dim, r] */
  char tag[16];
} tx_head_desc;
typedef struct tx_head_state {
  float* hg; void *hg_hi, *hg_lo; float* hg_scale;
  void *w_hi, *w_lo; float* w_scale;
  float* u; float* scores;
} tx_head_state;
int64_t tx_head_fwd_bytes(const tx_head_desc* d);
int64_t tx_head_bwd_bytes(const tx_head_desc* d);
int tx_head_fwd(const tx_head_desc* d, const float* h, int64_t ldh, const float* q, int64_t ldq, void* workspace, tx_head_state* state,
                float* scores, void* stream);
int tx_head_bwd(const tx_head_desc* d, const tx_head_state* state, const float* h, int64_t ldh, const float* q, int64_t ldq,
                const float* dscores, void* workspace, float* dh, float* dw, float* dw_main, float* dpos_weight, float** dh_amax_out,
                void* stream);      /* dw: [dim, round4(r)] pitched; dw_main (optional): contiguous [dim, r] */
/* Multi-GPU hook: tx_gat_layer_bwd records `cuda_event` (a cudaEvent_t, or NULL to stop) on its stream right after it has launched the
 * star backward, and counts the records.  A gradient bucket lets the all-reduce of the segments that were complete before that point
 * wait on this event from a side stream: the collective then runs beside the weight- / input-gradient GEMMs of the layer (which leave
 * SMs idle) instead of beside the persistent one-CTA-per-SM star backward (which it delayed by 0.04-0.07 ms on 4 / 8 GPUs). */
int tx_set_after_star_bwd_event(void* cuda_event);
int64_t tx_after_star_bwd_event_count(void);
/* measurement aid (bench.py): kernel launches issued by the two calls above since the last reset, and optional CUDA-event timing of
 * each of them (creates events; read after synchronising the stream) */
int64_t tx_layer_launches(int32_t reset);
void tx_prof_enable(int32_t on);
void tx_prof_clear(void);
int64_t tx_prof_count(void);
int tx_prof_get(int64_t i, char* name64, char* tag64, float* ms);

/* ------------------------------------------------------------------------------------------------
 * Matching + InfoNCE epilogue: the step right after the readout.
 * tx_match_rowdot: scores[g] = f(<u[g,:], q[g,:]>) with u = hg . W[0] (the projection half of nn.Bilinear(l, r, 1, bias=False),
 *   a GEMM of this library), f = identity for BIM (model/model_zoo.py:301-313) and exp for LBM (:316-328; apply_exp = 1).
 *   Backward: dt = dscores[g] * (apply_exp ? scores[g] : 1); du[g,:] = dt q[g,:]; dq[g,:] = dt u[g,:] (du or dq may be NULL).
 * tx_info_nce: loss = sum_q (logsumexp_j scores[q, j] - scores[q, target[q]]) = F.cross_entropy(scores.reshape(n_queries, group),
 *   target, reduction="sum") (model/loss.py:52-57 on the reshape of trainer/trainer.py:52-55).  target may be NULL (= all zeros,
 *   the only layout the reference produces: one positive first, then negative_size negatives, data_loader/dataset.py:308-313);
 *   a class index outside [0, group) yields a NaN loss.  loss_per_query / lse_per_query: [n_queries] device floats of caller-owned
 *   workspace (lse is kept for the backward pass); *loss is summed in a fixed order (deterministic).
 *   Backward: dscores[q, j] = *dloss * (exp(scores[q, j] - lse[q]) - [j == target[q]]).
 * ------------------------------------------------------------------------------------------------ */
int tx_match_rowdot_fwd(const float* u, int64_t ldu, const float* q, int64_t ldq, int64_t n_rows, int64_t r, int32_t apply_exp,
                        float* scores, void* stream);
int tx_match_rowdot_bwd(const float* u, int64_t ldu, const float* q, int64_t ldq, const float* scores, const float* dscores,
                        int64_t n_rows, int64_t r, int32_t apply_exp, float* du, int64_t lddu, float* dq, int64_t lddq,
                        void* stream);
int tx_info_nce_fwd(const float* scores, int64_t n_queries, int64_t group, const int32_t* target, float* loss_per_query,
                    float* lse_per_query, float* loss, void* stream);
int tx_info_nce_bwd(const float* scores, const float* lse_per_query, int64_t n_queries, int64_t group, const int32_t* target,
                    const float* dloss, float* dscores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAXO_B200_H_ */
