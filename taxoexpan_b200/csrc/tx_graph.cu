// Graph-structure kernels: CSR construction for arbitrary edge lists and the closed-form structure of a
// batch of star egonets (reference data_loader/dataset.py:404-437 + dgl.batch, data_loaders.py:25).
// Integer work only; results are bit-exact against oracle/taxo_oracle.py::csr_by_dst.
#define TX_PDL_GROUP 6
#include "tx_common.cuh"

namespace tx {

__global__ void histogram_kernel(const int32_t* __restrict__ key, int64_t n_edges, int32_t* __restrict__ count_plus1) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(count_plus1 + key[e] + 1, 1);
}

// Single-CTA exclusive scan of counts stored at ptr[1..n] (ptr[0] = 0) -> ptr becomes the CSR row pointer.
// 1024 threads, each owning a contiguous chunk; not on the per-step path for egonet batches.
__global__ void __launch_bounds__(1024) scan_kernel(int32_t* __restrict__ ptr, int64_t n) {
  __shared__ int32_t sums[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + 1023) / 1024;
  const int64_t b = 1 + (int64_t)t * chunk, e = min(n + 1, b + chunk);
  int32_t s = 0;
  for (int64_t i = b; i < e; ++i) s += ptr[i];
  sums[t] = s;
  __syncthreads();
  // inclusive scan of the 1024 chunk sums (Hillis-Steele)
  for (int off = 1; off < 1024; off <<= 1) {
    int32_t v = t >= off ? sums[t - off] : 0;
    __syncthreads();
    sums[t] += v;
    __syncthreads();
  }
  int32_t run = t > 0 ? sums[t - 1] : 0;
  for (int64_t i = b; i < e; ++i) {
    run += ptr[i];
    ptr[i] = run;
  }
  if (t == 0) ptr[0] = 0;
}

// Scatter edges into their rows with an atomic cursor (arbitrary order inside a row) ...
__global__ void fill_kernel(const int32_t* __restrict__ key, int64_t n_edges, const int32_t* __restrict__ ptr,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ eid_out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = key[e];
    const int slot = ptr[k] + atomicAdd(cursor + k, 1);
    eid_out[slot] = (int32_t)e;
  }
}

// ... then restore edge-id order inside every row (stable counting sort overall). One thread per row;
// rows of egonet batches have <= expand_factor + 2 entries.
__global__ void sort_rows_kernel(const int32_t* __restrict__ ptr, int64_t n_nodes, int32_t* __restrict__ eid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int b = ptr[i], e = ptr[i + 1];
  for (int a = b + 1; a < e; ++a) {  // insertion sort
    const int32_t v = eid[a];
    int c = a - 1;
    while (c >= b && eid[c] > v) {
      eid[c + 1] = eid[c];
      --c;
    }
    eid[c + 1] = v;
  }
}

__global__ void gather_by_dst_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ in_eid, int64_t n_edges,
                                     int32_t* __restrict__ in_src, int32_t* __restrict__ slot_of_eid) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_edges; s += (int64_t)gridDim.x * blockDim.x) {
    const int e = in_eid[s];
    in_src[s] = src[e];
    if (slot_of_eid) slot_of_eid[e] = (int32_t)s;
  }
}

__global__ void gather_by_src_kernel(const int32_t* __restrict__ dst, const int32_t* __restrict__ slot_of_eid,
                                     const int32_t* __restrict__ out_eid, int64_t n_edges, int32_t* __restrict__ out_dst,
                                     int32_t* __restrict__ out_slot) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_edges; s += (int64_t)gridDim.x * blockDim.x) {
    const int e = out_eid[s];
    out_dst[s] = dst[e];
    out_slot[s] = slot_of_eid[e];
  }
}

// One warp per egonet; lanes over its nodes. See the slot algebra in DESIGN.md ("closed-form egonet CSR").
__global__ void __launch_bounds__(256) star_batch_structure_kernel(
    const int32_t* __restrict__ n_gp, const int32_t* __restrict__ n_sib, const int32_t* __restrict__ node_off,
    const int32_t* __restrict__ edge_off, int n_graphs, int32_t* __restrict__ pos, int32_t* __restrict__ src,
    int32_t* __restrict__ dst, int32_t* __restrict__ in_ptr, int32_t* __restrict__ in_src, int32_t* __restrict__ in_eid,
    int32_t* __restrict__ out_ptr, int32_t* __restrict__ out_dst, int32_t* __restrict__ out_slot) {
  TX_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int g = warp; g < n_graphs; g += nwarps) {
    const int a = n_gp[g], s = n_sib[g], o = node_off[g], q = edge_off[g];
    const int n = a + 1 + s;
    const int self0 = q + a + s;  // edge id of the first self loop
    for (int t = lane; t < n; t += 32) {
      const int node = o + t;
      if (t < a) {  // grand-parent t: in = {self}; out = {-> anchor, self}
        if (pos) pos[node] = 0;
        if (src) { src[q + t] = node; dst[q + t] = o + a; }
        if (in_ptr) { in_ptr[node] = q + t; in_src[q + t] = node; in_eid[q + t] = self0 + t; }
        if (out_ptr) {
          out_ptr[node] = q + 2 * t;
          out_dst[q + 2 * t] = o + a;      out_slot[q + 2 * t] = q + a + t;
          out_dst[q + 2 * t + 1] = node;   out_slot[q + 2 * t + 1] = q + t;
        }
      } else if (t == a) {  // anchor: in = {gp_0..gp_{a-1}, self}; out = {-> sib_0.., self}
        if (pos) pos[node] = 1;
        if (in_ptr) {
          in_ptr[node] = q + a;
          for (int k = 0; k < a; ++k) { in_src[q + a + k] = o + k; in_eid[q + a + k] = q + k; }
          in_src[q + 2 * a] = node; in_eid[q + 2 * a] = self0 + a;
        }
        if (out_ptr) {
          out_ptr[node] = q + 2 * a;
          out_dst[q + 2 * a + s] = node; out_slot[q + 2 * a + s] = q + 2 * a;
        }
      } else {  // sibling k: in = {anchor, self}; out = {self}
        const int k = t - a - 1;
        if (pos) pos[node] = 2;
        if (src) { src[q + a + k] = o + a; dst[q + a + k] = node; }
        const int slot = q + 2 * a + 1 + 2 * k;
        if (in_ptr) {
          in_ptr[node] = slot;
          in_src[slot] = o + a;      in_eid[slot] = q + a + k;
          in_src[slot + 1] = node;   in_eid[slot + 1] = self0 + t;
        }
        if (out_ptr) {
          out_dst[q + 2 * a + k] = node; out_slot[q + 2 * a + k] = slot;          // anchor -> sib_k
          out_ptr[node] = q + 2 * a + s + 1 + k;
          out_dst[q + 2 * a + s + 1 + k] = node; out_slot[q + 2 * a + s + 1 + k] = slot + 1;
        }
      }
      if (src) { src[self0 + t] = node; dst[self0 + t] = node; }
    }
    if (g == n_graphs - 1 && lane == 0) {
      const int E = q + 2 * n - 1;
      if (in_ptr) in_ptr[o + n] = E;
      if (out_ptr) out_ptr[o + n] = E;
    }
  }
}


// Plan of a star-egonet batch from the per-egonet counts alone (one CTA, tiles of 1024 egonets): node / edge offsets (exclusive scans of
// n = a + 1 + s and e = 2 n - 1) and the work-item tables of the star kernels - one 16-byte record {first node, first edge,
// n_gp | chunk << 24, n_sib} per (egonet, chunk of C siblings), C = chunk_fwd for tx_gat_star_fwd and chunk_bwd for tx_gat_star_bwd.  An
// egonet's records are consecutive; the EGONETS are ordered by size class - more than C siblings first, then 1..C siblings, then none
// (stable inside a class) - so that the work queues of the star kernels hand out their longest items first and the last warps to
// finish hold items of a few rows, not of 17 (the tail was 27 % of the output layer's backward).  Moves ~0.4 ms of numpy (two cumsums,
// two row-repeats per table) per batch off the host.
__global__ void __launch_bounds__(1024, 1) star_batch_plan_kernel(const int32_t* __restrict__ n_gp, const int32_t* __restrict__ n_sib, int G,
                                                                   int chunk_fwd, int chunk_bwd, int32_t* __restrict__ node_off,
                                                                   int32_t* __restrict__ edge_off, int4* __restrict__ tasks_fwd,
                                                                   int4* __restrict__ tasks_bwd) {
  TX_PDL_ENTER();
  constexpr int NQ = 7;                                   // scanned quantities: nodes, records of the 3 classes x {fwd, bwd}
  constexpr int PER = 8;                                  // consecutive egonets per thread and round (8192 per round)
  __shared__ int s_warp[NQ][32];
  __shared__ int s_carry[NQ];
  __shared__ int s_tot[NQ];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto cls = [](int s, int chunk) -> int { return s > chunk ? 0 : (s > 0 ? 1 : 2); };
  if (threadIdx.x < NQ) { s_carry[threadIdx.x] = 0; s_tot[threadIdx.x] = 0; }
  __syncthreads();
  // ---- pass 1 (only when the batch does not fit one round): records per class = the classes' base offsets in the tables ----
  const bool one_round = G <= 1024 * PER;
  if (!one_round) {
    int acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0;
    for (int k = threadIdx.x; k < G; k += 1024) {
      const int s = n_sib[k];
      if (tasks_fwd) acc[1 + cls(s, chunk_fwd)] += max(1, (s + chunk_fwd - 1) / chunk_fwd);
      if (tasks_bwd) acc[4 + cls(s, chunk_bwd)] += max(1, (s + chunk_bwd - 1) / chunk_bwd);
    }
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
      int x = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) atomicAdd(&s_tot[q], x);            // integer sums: order does not matter
    }
    __syncthreads();
  }
  // ---- pass 2: every thread owns PER consecutive egonets of a round: local sums, ONE block-wide scan, local prefix walk ----
  for (int base = 0; base < G; base += 1024 * PER) {
    const int k0 = base + threadIdx.x * PER;
    int a[PER], s[PER], tot[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) tot[q] = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = k0 + i;
      a[i] = k < G ? n_gp[k] : 0;
      s[i] = k < G ? n_sib[k] : 0;
      if (k < G) {
        tot[0] += a[i] + 1 + s[i];
        if (tasks_fwd) tot[1 + cls(s[i], chunk_fwd)] += max(1, (s[i] + chunk_fwd - 1) / chunk_fwd);
        if (tasks_bwd) tot[4 + cls(s[i], chunk_bwd)] += max(1, (s[i] + chunk_bwd - 1) / chunk_bwd);
      }
    }
    int incl[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      int x = tot[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      incl[q] = x;
      if (lane == 31) s_warp[q][wid] = x;
    }
    __syncthreads();
    if (wid < NQ) {
      int x = s_warp[wid][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      s_warp[wid][lane] = x;                                // inclusive scan of the warp totals
      if (one_round && lane == 31) s_tot[wid] = x;          // a single round: its totals ARE the class sizes
    }
    __syncthreads();
    int run[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) run[q] = s_carry[q] + (wid ? s_warp[q][wid - 1] : 0) + incl[q] - tot[q];
    const int base_f[3] = {0, s_tot[1], s_tot[1] + s_tot[2]};
    const int base_b[3] = {0, s_tot[4], s_tot[4] + s_tot[5]};
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int k = k0 + i;
      if (k < G) {
        const int o = run[0], e = 2 * o - k;               // sum of (2 n - 1) over the earlier egonets
        node_off[k] = o;
        edge_off[k] = e;
        run[0] += a[i] + 1 + s[i];
        if (tasks_fwd) {
          const int c3 = cls(s[i], chunk_fwd), cnt = max(1, (s[i] + chunk_fwd - 1) / chunk_fwd), first = base_f[c3] + run[1 + c3];
          for (int c = 0; c < cnt; ++c) tasks_fwd[first + c] = make_int4(o, e, a[i] | (c << 24), s[i]);
          run[1 + c3] += cnt;
        }
        if (tasks_bwd) {
          const int c3 = cls(s[i], chunk_bwd), cnt = max(1, (s[i] + chunk_bwd - 1) / chunk_bwd), first = base_b[c3] + run[4 + c3];
          for (int c = 0; c < cnt; ++c) tasks_bwd[first + c] = make_int4(o, e, a[i] | (c << 24), s[i]);
          run[4 + c3] += cnt;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < NQ) s_carry[threadIdx.x] += s_warp[threadIdx.x][31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    node_off[G] = s_carry[0];
    edge_off[G] = 2 * s_carry[0] - G;
  }
}

// out[i, :] = table[ids[i], :]: the per-step feature rows of a batch taken from the RESIDENT node-embedding table (the reference keeps
// g_full.ndata['x'] in host memory and collates rows per egonet, dataset.py:157,429-431; here a step ships node ids, not rows).
template <int VEC>
__global__ void gather_rows_kernel(const float* __restrict__ table, int64_t ldt, const int32_t* __restrict__ ids, int64_t n, int d, int64_t n_table,
                                   float* __restrict__ out, int64_t ldo) {
  TX_PDL_ENTER();
  const int per_row = d / VEC;
  const int64_t total = n * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / per_row;
    const int c = (int)(t - i * per_row) * VEC;
    int64_t r = ids[i];
    r = r < 0 ? 0 : (r >= n_table ? n_table - 1 : r);      // ids are validated by the caller; never read out of bounds
    if (VEC == 4) {
      *reinterpret_cast<float4*>(out + i * ldo + c) = __ldg(reinterpret_cast<const float4*>(table + r * ldt + c));
    } else {
      out[i * ldo + c] = __ldg(table + r * ldt + c);
    }
  }
}

}  // namespace tx

using namespace tx;

extern "C" {

int tx_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t* bytes) {
  TX_REQUIRE(bytes, "csr_workspace_bytes: null output");
  *bytes = (n_nodes + 1) * 4 + n_edges * 4;  // row cursor + edge-id scratch for the by-src pass
  return TX_OK;
}

static int build_rows(const int32_t* key, int64_t n_nodes, int64_t n_edges, int32_t* ptr, int32_t* eid_sorted,
                      int32_t* cursor, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(ptr, 0, (n_nodes + 1) * sizeof(int32_t), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(cursor, 0, (n_nodes + 1) * sizeof(int32_t), st);
  if (e != cudaSuccess) {
    set_error("build_csr: memset failed: %s", cudaGetErrorString(e));
    return TX_ERR_CUDA;
  }
  if (n_edges > 0) {
    const int grid = (int)((n_edges + 255) / 256 < (int64_t)kNumSms * 8 ? (n_edges + 255) / 256 : (int64_t)kNumSms * 8);
    histogram_kernel<<<grid, 256, 0, st>>>(key, n_edges, ptr);
    scan_kernel<<<1, 1024, 0, st>>>(ptr, n_nodes);
    fill_kernel<<<grid, 256, 0, st>>>(key, n_edges, ptr, cursor, eid_sorted);
    sort_rows_kernel<<<(int)((n_nodes + 127) / 128), 128, 0, st>>>(ptr, n_nodes, eid_sorted);
  }
  TX_LAUNCH_CHECK("tx_build_csr");
  return TX_OK;
}

int tx_build_csr_by_dst(const int32_t* src, const int32_t* dst, int64_t n_nodes, int64_t n_edges, int32_t* in_ptr,
                        int32_t* in_src, int32_t* in_eid, int32_t* slot_of_eid, void* workspace, void* stream) {
  TX_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_nodes < INT32_MAX && n_edges < INT32_MAX, "build_csr: sizes must fit int32");
  TX_REQUIRE(workspace && in_ptr, "build_csr: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = build_rows(dst, n_nodes, n_edges, in_ptr, in_eid, (int32_t*)workspace, st);
  if (rc != TX_OK) return rc;
  if (n_edges > 0) {
    const int grid = (int)((n_edges + 255) / 256 < (int64_t)kNumSms * 8 ? (n_edges + 255) / 256 : (int64_t)kNumSms * 8);
    gather_by_dst_kernel<<<grid, 256, 0, st>>>(src, in_eid, n_edges, in_src, slot_of_eid);
    TX_LAUNCH_CHECK("tx_build_csr_by_dst");
  }
  return TX_OK;
}

int tx_build_csr_by_src(const int32_t* src, const int32_t* dst, const int32_t* slot_of_eid, int64_t n_nodes,
                        int64_t n_edges, int32_t* out_ptr, int32_t* out_dst, int32_t* out_slot, void* workspace,
                        void* stream) {
  TX_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_nodes < INT32_MAX && n_edges < INT32_MAX, "build_csr: sizes must fit int32");
  TX_REQUIRE(workspace && out_ptr && slot_of_eid, "build_csr: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* cursor = (int32_t*)workspace;
  int32_t* out_eid = cursor + (n_nodes + 1);
  int rc = build_rows(src, n_nodes, n_edges, out_ptr, out_eid, cursor, st);
  if (rc != TX_OK) return rc;
  if (n_edges > 0) {
    const int grid = (int)((n_edges + 255) / 256 < (int64_t)kNumSms * 8 ? (n_edges + 255) / 256 : (int64_t)kNumSms * 8);
    gather_by_src_kernel<<<grid, 256, 0, st>>>(dst, slot_of_eid, out_eid, n_edges, out_dst, out_slot);
    TX_LAUNCH_CHECK("tx_build_csr_by_src");
  }
  return TX_OK;
}

int tx_star_batch_structure(const int32_t* n_gp, const int32_t* n_sib, const int32_t* node_off,
                            const int32_t* edge_off, int64_t n_graphs, int32_t* pos, int32_t* src, int32_t* dst,
                            int32_t* in_ptr, int32_t* in_src, int32_t* in_eid, int32_t* out_ptr, int32_t* out_dst,
                            int32_t* out_slot, void* stream) {
  TX_REQUIRE(n_graphs >= 0 && n_graphs < INT32_MAX, "star_batch_structure: bad graph count");
  TX_REQUIRE((src == nullptr) == (dst == nullptr), "star_batch_structure: src and dst go together");
  TX_REQUIRE(!in_ptr || (in_src && in_eid), "star_batch_structure: in_ptr needs in_src and in_eid");
  TX_REQUIRE(!out_ptr || (out_dst && out_slot), "star_batch_structure: out_ptr needs out_dst and out_slot");
  if (n_graphs == 0) return TX_OK;
  const int grid = grid_for_warps(n_graphs, 8, 8);
  TX_PDL_LAUNCH((star_batch_structure_kernel), grid, 256, 0, (cudaStream_t)stream, n_gp, n_sib, node_off, edge_off, (int)n_graphs, pos, src,
                                                                     dst, in_ptr, in_src, in_eid, out_ptr, out_dst, out_slot);
  TX_LAUNCH_CHECK("tx_star_batch_structure");
  return TX_OK;
}

int tx_gather_rows(const float* table, int64_t ldt, int64_t n_table, const int32_t* ids, int64_t n, int64_t d, float* out, int64_t ldo,
                   void* stream) {
  TX_REQUIRE(table && ids && out && n >= 0 && d > 0 && n_table > 0 && ldt >= d && ldo >= d && d < INT32_MAX, "gather_rows: bad arguments");
  if (n == 0) return TX_OK;
  const bool vec = d % 4 == 0 && ldt % 4 == 0 && ldo % 4 == 0 && aligned16(table) && aligned16(out);
  const int64_t total = n * (vec ? d / 4 : d);
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)kNumSms * 16) grid = (int64_t)kNumSms * 16;
  if (vec) TX_PDL_LAUNCH((gather_rows_kernel<4>), (unsigned)grid, 256, 0, (cudaStream_t)stream, table, ldt, ids, n, (int)d, n_table, out, ldo);
  else TX_PDL_LAUNCH((gather_rows_kernel<1>), (unsigned)grid, 256, 0, (cudaStream_t)stream, table, ldt, ids, n, (int)d, n_table, out, ldo);
  TX_LAUNCH_CHECK("tx_gather_rows");
  return TX_OK;
}

int tx_star_batch_plan(const int32_t* n_gp, const int32_t* n_sib, int64_t n_graphs, int64_t chunk_fwd, int64_t chunk_bwd, int32_t* node_off,
                       int32_t* edge_off, int32_t* tasks_fwd, int32_t* tasks_bwd, void* stream) {
  TX_REQUIRE(n_graphs >= 0 && n_graphs < INT32_MAX && n_gp && n_sib && node_off && edge_off, "star_batch_plan: bad arguments");
  TX_REQUIRE((!tasks_fwd || (chunk_fwd >= 1 && aligned16(tasks_fwd))) && (!tasks_bwd || (chunk_bwd >= 1 && aligned16(tasks_bwd))),
             "star_batch_plan: task tables need a chunk size >= 1 and 16-byte alignment");
  TX_PDL_LAUNCH((star_batch_plan_kernel), 1, 1024, 0, (cudaStream_t)stream, n_gp, n_sib, (int)n_graphs, (int)chunk_fwd, (int)chunk_bwd, node_off, edge_off,
                                                               reinterpret_cast<int4*>(tasks_fwd), reinterpret_cast<int4*>(tasks_bwd));
  TX_LAUNCH_CHECK("tx_star_batch_plan");
  return TX_OK;
}

}  // extern "C"
