// Micro-benchmark: how fast can ONE producer warp per SM stream "tiles" of R rows x D floats (row pitch LD floats) of two arrays
// into shared memory, double buffered, with (a) one 1-D cp.async.bulk per row, (b) 2-D tensor TMA boxes, (c) 16-byte cp.async
// (LDGSTS)?  Consumers only wait on the full barrier and release the stage.  Prints GB/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench scripts/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

constexpr int STAGES = 2;

// variant 0: 1-D bulk per row (lanes issue);  1: 2-D tensor boxes {128 floats, R rows} x ceil(D/128);  2: cp.async 16 B;
// variant 3: 1-D bulk per row, but STAGES deep only limited by smem (same as 0; PW producer warps split the rows)
template <int VARIANT>
__global__ void __launch_bounds__(256, 1) stream_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t ld, int n_rows, int R,
                                                        int D, int heads, const __grid_constant__ CUtensorMap map_a,
                                                        const __grid_constant__ CUtensorMap map_b, int pw, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2 * STAGES];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = s_addr(bars), empty0 = s_addr(bars + STAGES);
  const uint32_t rowB = (uint32_t)D * 4u;
  const int n_chunks = (D + 127) / 128;
  const uint32_t stage_bytes = 2u * (uint32_t)R * (uint32_t)n_chunks * 512u;
  const int h = blockIdx.y;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      bar_init(full0 + 8 * s, VARIANT == 2 ? 32 * pw : pw);
      bar_init(empty0 + 8 * s, 8 - pw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_tiles = (n_rows + R - 1) / R;
  float acc = 0.f;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1;
    const int r0 = tile * R;
    const int nr = min(R, n_rows - r0);
    uint8_t* st = smem + (size_t)s * stage_bytes;
    if (wid < pw) {   // producer warps
      if (it >= STAGES) bar_wait(empty0 + 8 * s, ph ^ 1);
      const uint32_t full = full0 + 8 * s;
      if (VARIANT == 0) {
        int mine = 0;
        for (int r = wid * 32 + lane; r < nr; r += 32 * pw) ++mine;
        // each producer warp's lane 0 posts the bytes of its warp's rows
        int wrows = 0;
        for (int r = wid * 32; r < nr; r += 32 * pw) wrows += min(32, nr - r);
        if (lane == 0) bar_expect_tx(full, (uint32_t)wrows * 2u * rowB);
        __syncwarp();
        for (int r = wid * 32 + lane; r < nr; r += 32 * pw) {
          bulk_g2s(s_addr(st + (size_t)r * rowB), a + (int64_t)(r0 + r) * ld + (int64_t)h * D, rowB, full);
          bulk_g2s(s_addr(st + (size_t)(R + r) * rowB), b + (int64_t)(r0 + r) * ld + (int64_t)h * D, rowB, full);
        }
        __syncwarp();
        if (lane == 0) bar_arrive(full);
      } else if (VARIANT == 1) {
        if (wid == 0 && lane == 0) {
          uint32_t bytes = 0;
          for (int c = 0; c < n_chunks; ++c) bytes += 2u * (uint32_t)R * 512u;
          bar_expect_tx(full, bytes);     // OOB rows are zero-filled and still counted
          uint32_t off = 0;
          for (int c = 0; c < n_chunks; ++c) {
            const uint32_t cb = (uint32_t)R * 512u;
            tma_2d(s_addr(st + off), &map_a, full, h * D + 128 * c, r0);
            tma_2d(s_addr(st + off + cb), &map_b, full, h * D + 128 * c, r0);
            off += 2 * cb;
          }
        }
        __syncwarp();
        if (lane == 0) bar_arrive(full);
      } else {
        const int v_per_row = D / 4;
        const int total = nr * v_per_row;
        for (int t = wid * 32 + lane; t < total; t += 32 * pw) {
          const int r = t / v_per_row, c = t - r * v_per_row;
          cp16(s_addr(st + (size_t)r * rowB + c * 16), a + (int64_t)(r0 + r) * ld + (int64_t)h * D + c * 4);
          cp16(s_addr(st + (size_t)(R + r) * rowB + c * 16), b + (int64_t)(r0 + r) * ld + (int64_t)h * D + c * 4);
        }
        cp_arrive_noinc(full);
      }
    } else {          // consumers: wait, touch one value, release
      bar_wait(full0 + 8 * s, ph);
      acc += reinterpret_cast<const float*>(st)[threadIdx.x];
      __syncwarp();
      if (lane == 0) bar_arrive(empty0 + 8 * s);
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st));
  EncodeTiledFn fn = (EncodeTiledFn)p;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

template <int V>
static void run(const char* name, const float* a, const float* b, int64_t ld, int n, int R, int D, int H, int pw, float* sink) {
  CUtensorMap ma, mb;
  make_map(&ma, a, n, ld, ld, R, 128);
  make_map(&mb, b, n, ld, ld, R, 128);
  const int n_chunks = (D + 127) / 128;
  const size_t smem = (size_t)STAGES * 2 * R * n_chunks * 512 + 1024;
  CK(cudaFuncSetAttribute(stream_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(148 / H, H);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) stream_kernel<V><<<grid, 256, smem>>>(a, b, ld, n, R, D, H, ma, mb, pw, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int i = 0; i < reps; ++i) stream_kernel<V><<<grid, 256, smem>>>(a, b, ld, n, R, D, H, ma, mb, pw, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double bytes = 2.0 * n * H * D * 4.0;
  printf("%-34s R=%3d pw=%d smem=%6zu B : %.3f ms  %.0f GB/s\n", name, R, pw, smem, ms, bytes / ms / 1e6);
}

int main() {
  const int n = 37039 * 4, H = 4, D = 500;   // 4x the bench batch so that the arrays (2 x 1.2 GB) exceed L2
  const int64_t ld = (int64_t)H * D;
  float *a, *b, *sink;
  CK(cudaMalloc(&a, (size_t)n * ld * 4)); CK(cudaMalloc(&b, (size_t)n * ld * 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(a, 0, (size_t)n * ld * 4)); CK(cudaMemset(b, 0, (size_t)n * ld * 4));
  for (int R : {8, 16, 24}) {
    run<0>("1-D bulk per row", a, b, ld, n, R, D, H, 1, sink);
    run<0>("1-D bulk per row", a, b, ld, n, R, D, H, 2, sink);
    run<0>("1-D bulk per row", a, b, ld, n, R, D, H, 4, sink);
    run<1>("2-D tensor boxes {128 x R}", a, b, ld, n, R, D, H, 1, sink);
    run<2>("cp.async 16 B", a, b, ld, n, R, D, H, 1, sink);
    run<2>("cp.async 16 B", a, b, ld, n, R, D, H, 2, sink);
    run<2>("cp.async 16 B", a, b, ld, n, R, D, H, 4, sink);
  }
  return 0;
}
