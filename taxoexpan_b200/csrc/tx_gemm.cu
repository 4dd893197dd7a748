// fp32-faithful dense projection on the 5th-generation tensor cores (sm_100a): C[M,N] = A[M,K] . B[N,K]^T
//
// The reference path is fp32 end to end and the parity bar is 1e-5, so a single TF32 pass (2^-11 relative error per
// product) is not acceptable.  Every fp32 operand x is split as x = hi + lo with hi = rn_tf32(x) (a valid TF32
// value) and lo = rn_tf32(x - hi) (the difference is exact in fp32, then rounded to TF32), and the product is accumulated as
//     A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (3xTF32, fp32 accumulation in tensor memory),
// which leaves ~2^-21 relative error per product.
//
// Kernel anatomy (one CTA per 128 x BN output tile, 192 threads):
//   warp 0     : TMA producer - cp.async.bulk.tensor.2d of the four operand tiles (A_hi, A_lo, B_hi, B_lo; K-major,
//                32 fp32 = 128 B per row, SWIZZLE_128B) into a STAGES-deep shared-memory ring, mbarrier expect_tx.
//   warp 1     : MMA issuer - one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8),
//                12 per k-block; accumulator lives in TMEM (BN columns x 128 lanes); tcgen05.commit frees the stage.
//   warps 2..5 : epilogue - tcgen05.ld 32x32b.x32 (TMEM -> registers, one row per thread), 128-bit global stores.
// Operand tails: TMA zero-fills out-of-bounds rows / K columns, the epilogue masks rows >= M and columns >= ldc.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#define TX_PDL_GROUP 2
#include "tx_common.cuh"

namespace tx {

constexpr int kBM = 128;
constexpr int kBK = 32;                      // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int kABytes = kBM * kBK * 4;       // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// multicast variants for a 2-CTA cluster: the TMA box lands in BOTH CTAs' shared memory (same offsets) and completes bytes on
// both CTAs' mbarriers; the commit arrives on the barrier at the same offset in every CTA of the mask
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1) |
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups) | [46,48) version = 1 |
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// same with 64-byte rows (16 fp32 per k-block): SWIZZLE_64B (layout type 4), 8 rows x 64 B = 512 B between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t smem_addr) {
  return BK == 32 ? umma_desc_k_sw128(smem_addr) : umma_desc_k_sw64(smem_addr);
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32 (F16 = false) or kind::f16 (F16 = true), cta_group::1
template <bool F16>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// MN-major operands (reduction index slow in memory).  For 32-bit (tf32) MN-major operands the tensor core accepts ONE
// shared-memory layout: SWIZZLE_128B_BASE32B (CUTLASS sm100_smem_selector: "for mn-major tf32 operands, SW128_32B is the
// only available smem layout"; UMMA::LayoutType 1, Swizzle<2,5,2>): 128-byte rows of 32 consecutive MN elements, atoms
// of 4 k-rows (512 B) whose 32-byte chunks are XOR-permuted by the row index - exactly what a {32 cols, kBK rows} TMA box
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.  Canonical form ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in floats:
//   LBO = bytes between consecutive 32-float MN blocks (one TMA box = kBK * 128 B), SBO = bytes between 4-row k groups (512).
__host__ __device__ __forceinline__ uint64_t umma_desc_mn_bits(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint64_t bits) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | bits;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Optional fused output transform of the NT kernel: C[row, c] *= keep ? (positive ? on : neg) : 0 for c < feat_cols, decoded
// from the sign/keep bytes the fused GAT forward wrote (1 byte per aligned group of 4 columns: byte index = row * stride + c / 4,
// stride = tx_gat_fused_mask_ld, a multiple of 16, so a thread's 32-column chunks are aligned 8-byte runs).  This is the backward of
// "leaky-relu -> dropout" applied right where d(z) is produced (reference: autograd of model_zoo.py:215-216 and :82).
struct GemmEpilogue {
  const uint8_t* mask;
  int heads, dim, stride, feat_cols, has_keep;
  float on, neg;
  // fp16-split operands carry per-tensor power-of-two scales: C = acc / (scale_a * scale_b) (device scalars; null = 1)
  const float* scale_a;
  const float* scale_b;
  float* amax_out;      // optional: atomicMax of |C| over everything stored (device scalar, zeroed by the launcher)
};

constexpr int kChunkK = 64;      // K extent accumulated inside the tensor core before promotion to fp32 registers (24 MMAs)
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM lane quadrant)
constexpr int kPairThreads = 576;  // cta_group::2 kernel: warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (4 per TMEM lane quadrant, 64 columns
                                   // each: the accumulator drain + global stores of a 128 x 256 tile were the bound of the short-K GEMMs)

// C[M, N] (+ split offset) = sum_k A . B with 3xTF32 products and CHUNKED PROMOTION:
//   the tensor core adds each MMA into its fp32 TMEM accumulator with truncation (measured: ~0.5 ulp lost per MMA, a
//   relative bias of ~2e-8 x #MMAs, 1.5e-5 at K = 2050), so the accumulator is drained every kChunk k-blocks
//   (24 MMAs) into fp32 REGISTERS of the epilogue warps (round-to-nearest adds) while the MMA warp fills the other
//   TMEM buffer.  2 x BN columns of TMEM, double buffered.
// TN = false: A [M, K], B [N, K] row-major (K-major operands): y = z W^T, dz = dy W.
// TN = true : A [R, M], B [R, N] row-major (MN-major operands, reduction over rows): dW = dy^T z; grid.z = split index.
// CL = 2: clusters of two CTAs along M share every B tile - each CTA fetches half of it and TMA-multicasts it to both, which cuts
// the L2 -> SM operand traffic per CTA from (A + B) to (A + B/2) (the 3xTF32 kernel is L2-bandwidth bound: 96 KB per k-block).
// F16 = true: the operands are fp16 hi / lo pairs (x * scale = hi + lo, per-tensor power-of-two scale): same 128-byte smem rows
// (64 fp16 instead of 32 tf32 per k-block row), tcgen05.mma.kind::f16 (UMMA_K = 16, i.e. the same 32 bytes per step) at twice
// the tensor rate and half the operand bytes per flop; MN-major operands use the plain SWIZZLE_128B layout (64 elements per row).
template <int BN, int STAGES, bool TN, int BK, int CL, bool F16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   float* __restrict__ C, int64_t ldc, int64_t split_stride, int M, int n_store, int k_blocks_total,
                   int k_blocks_per_split, uint64_t mn_desc_bits, const GemmEpilogue epi, int gran) {
  static_assert(BK == 32 || (BK == 16 && !TN), "BK = 16 (64-byte rows, SWIZZLE_64B) is implemented for the K-major form only");
  constexpr int kBK = BK;                     // 4-byte units per k-block row of a K-major tile (32 = 128 bytes)
  constexpr int BKE = F16 ? 2 * BK : BK;      // reduction ELEMENTS per k-block
  constexpr int BOXC = F16 ? 64 : 32;         // TN: MN elements per box row (128 bytes)
  constexpr int kABytes = kBM * BK * 4;
  constexpr int kChunk = kChunkK / BK;
  constexpr int B_BYTES = BN * kBK * 4;
  constexpr int STAGE_BYTES = 2 * kABytes + 2 * B_BYTES;
  constexpr int BOX_BYTES = BKE * 128;       // TN: one {BOXC cols, BKE rows} box
  static_assert(!(TN && F16) || BN % 64 == 0, "fp16 MN-major tiles are made of 64-column boxes");
  constexpr int HALF = ((BN / 32 + 1) / 2) * 32;   // columns owned by an epilogue warp of group 0 (group 1: BN - HALF), chunks of 32
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 32 in [64, 256]");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CL);           // the stage is reusable once the MMAs of every CTA writing into it retired
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 8);           // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_trigger();
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  if (warp == 1) {   // one warp allocates both accumulator buffers (power of two >= 32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();              // the peer's barriers are initialised before anything is multicast into it
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  if (!pdl_wait_guard()) __trap();             // (never taken, see tx_common.cuh) barriers / TMEM were set up while the previous kernel drained
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  // LAST tile of a row of tiles: the MMA is issued only as wide as the columns that exist, rounded up to `gran` (the instruction's N
  // granularity - 16 - or, for MN-major B, whole TMA boxes; 0 = always BN): N = 300 with BN = 192 costs 192 + 128 columns instead of
  // 2 x 192, and (TN, no multicast) the boxes past nd are not even fetched.  Rows / boxes past N that are fetched are zero-filled by TMA.
  // (The same narrowing in the cta_group::2 pair kernel bought nothing - its wide GEMMs are not limited by MMA issue slots, see
  // profiles/r2_gemm_bound_analysis.txt - and the extra live values cost it 3-5 % at its 96-register cap, so it stays fixed-width.)
  const int nd = gran > 0 ? min(BN, ((n_store - n0 + gran - 1) / gran) * gran) : BN;
  const int kb0 = blockIdx.z * k_blocks_per_split;
  const int nkb = max(0, min(k_blocks_per_split, k_blocks_total - kb0));
  const int n_chunks = (nkb + kChunk - 1) / kChunk;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);                     // slot released by the MMA warp
        const uint32_t full = full0 + 8 * s;
        const int nb_dyn = (TN && CL == 1) ? nd / BOXC : BN / BOXC;       // B boxes this tile fetches
        mbar_expect_tx(full, (TN && CL == 1) ? 2 * kABytes + 2 * nb_dyn * BOX_BYTES : STAGE_BYTES);
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
        const int kk = (kb0 + kb) * BKE;
        if constexpr (!TN) {
          tma_load_2d(base, &map_a_hi, full, kk, m0);
          tma_load_2d(base + kABytes, &map_a_lo, full, kk, m0);
          if constexpr (CL == 1) {
            tma_load_2d(base + 2 * kABytes, &map_b_hi, full, kk, n0);
            tma_load_2d(base + 2 * kABytes + B_BYTES, &map_b_lo, full, kk, n0);
          } else {   // my half of the B rows (the B maps have box rows = BN / 2), delivered to both CTAs
            const uint32_t off = crank * (uint32_t)(B_BYTES / 2);
            tma_load_2d_mc(base + 2 * kABytes + off, &map_b_hi, full, kk, n0 + (int)crank * (BN / 2), (uint16_t)3);
            tma_load_2d_mc(base + 2 * kABytes + B_BYTES + off, &map_b_lo, full, kk, n0 + (int)crank * (BN / 2), (uint16_t)3);
          }
        } else {
#pragma unroll
          for (int b = 0; b < kBM / BOXC; ++b) {
            tma_load_2d(base + b * BOX_BYTES, &map_a_hi, full, m0 + BOXC * b, kk);
            tma_load_2d(base + kABytes + b * BOX_BYTES, &map_a_lo, full, m0 + BOXC * b, kk);
          }
          constexpr int NB = BN / BOXC, NB0 = (NB + 1) / 2;        // CL = 2: rank 0 fetches boxes [0, NB0), rank 1 the rest
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if (CL == 1) {
              if (b < nb_dyn) {
                tma_load_2d(base + 2 * kABytes + b * BOX_BYTES, &map_b_hi, full, n0 + BOXC * b, kk);
                tma_load_2d(base + 2 * kABytes + B_BYTES + b * BOX_BYTES, &map_b_lo, full, n0 + BOXC * b, kk);
              }
            } else if ((b < NB0) == (crank == 0)) {
              tma_load_2d_mc(base + 2 * kABytes + b * BOX_BYTES, &map_b_hi, full, n0 + BOXC * b, kk, (uint16_t)3);
              tma_load_2d_mc(base + 2 * kABytes + B_BYTES + b * BOX_BYTES, &map_b_lo, full, n0 + BOXC * b, kk, (uint16_t)3);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer =====
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6) = 1, a/b_format TF32 [7,10),[10,13) = 2,
      // a_major [15], b_major [16] (0 = K-major, 1 = MN-major), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
      const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (TN ? ((1u << 15) | (1u << 16)) : 0u) |
                             ((uint32_t)(nd >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);   // a/b_format: 0 = F16, 2 = TF32
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        mbar_wait(tempty0 + 8 * buf, ((ch >> 1) & 1) ^ 1);     // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
        const int kb_end = min(nkb, (ch + 1) * kChunk);
        for (int kb = ch * kChunk; kb < kb_end; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);                        // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          uint64_t a_hi, a_lo, b_hi, b_lo;
          if constexpr (!TN) {
            a_hi = umma_desc_k<BK>(base); a_lo = umma_desc_k<BK>(base + kABytes);
            b_hi = umma_desc_k<BK>(base + 2 * kABytes); b_lo = umma_desc_k<BK>(base + 2 * kABytes + B_BYTES);
          } else {
            a_hi = umma_desc_mn(base, mn_desc_bits); a_lo = umma_desc_mn(base + kABytes, mn_desc_bits);
            b_hi = umma_desc_mn(base + 2 * kABytes, mn_desc_bits); b_lo = umma_desc_mn(base + 2 * kABytes + B_BYTES, mn_desc_bits);
          }
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            // UMMA_K = 8 tf32 / 16 fp16.  K-major: +32 bytes inside the 128-byte swizzle row; MN-major: the next 8 / 16 reduction
            // rows (+1024 / +2048 bytes)
            const uint64_t adv = TN ? (uint64_t)((k * (F16 ? 2048 : 1024)) >> 4) : (uint64_t)(2 * k);
            const uint32_t first = (kb == ch * kChunk && k == 0) ? 0u : 1u;
            umma_tf32<F16>(tacc, a_lo + adv, b_hi + adv, idesc, first);
            umma_tf32<F16>(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32<F16>(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          if (CL == 1) tcgen05_commit(empty0 + 8 * s);         // frees the smem stage once these MMAs retire
          else tcgen05_commit_mc(empty0 + 8 * s, (uint16_t)3);  // ... in both CTAs (each writes half of B into the other)
        }
        tcgen05_commit(tfull0 + 8 * buf);                      // chunk accumulator complete
      }
    }
  } else {             // ===== epilogue warps: drain chunk accumulators into fp32 registers, then store =====
    const int q = warp & 3;                                    // a warp may only touch TMEM lanes [32 q, 32 q + 32)
    const int hsel = (warp - 2) >> 2;                          // which part of the BN columns: [0, HALF) or [HALF, BN)
    const int n_chunks32 = max(0, min(hsel == 0 ? HALF / 32 : (BN - HALF) / 32, (nd - hsel * HALF + 31) / 32));   // columns >= nd were not computed
    float acc[HALF];
#pragma unroll
    for (int c = 0; c < HALF; ++c) acc[c] = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int buf = ch & 1;
      mbar_wait(tfull0 + 8 * buf, (ch >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN + hsel * HALF) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < HALF / 32; ++c) {
        if (c < n_chunks32) {                                  // warp-uniform
          uint32_t r[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    const int row = m0 + q * 32 + lane;
    const float inv_scale = (epi.scale_a ? 1.f / __ldg(epi.scale_a) : 1.f) * (epi.scale_b ? 1.f / __ldg(epi.scale_b) : 1.f);   // powers of two: exact
    float vmax = 0.f;
    if (row < M) {
      float* crow = C + (int64_t)blockIdx.z * split_stride + (int64_t)row * ldc + n0 + hsel * HALF;
      uint32_t mw[HALF / 16];                                  // 4 mask bytes (16 columns) per word
      const bool masked = !TN && epi.mask != nullptr;
      if (masked) {
#pragma unroll
        for (int w = 0; w < HALF / 16; ++w) {
          const int col = n0 + hsel * HALF + w * 16;
          mw[w] = (w < n_chunks32 * 2 && col < epi.feat_cols)
                      ? __ldg(reinterpret_cast<const uint32_t*>(epi.mask + (int64_t)row * epi.stride + (col >> 2))) : 0xFFFFFFFFu;
        }
      }
#pragma unroll
      for (int j = 0; j < HALF / 4; ++j) {
        const int col = n0 + hsel * HALF + j * 4;
        if (j < n_chunks32 * 8 && col < n_store) {
          float4 v = make_float4(acc[4 * j] * inv_scale, acc[4 * j + 1] * inv_scale, acc[4 * j + 2] * inv_scale, acc[4 * j + 3] * inv_scale);
          if (masked && col < epi.feat_cols) {
            uint32_t code = (mw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            if (!epi.has_keep) code |= 0xF0u;
            v.x *= (code & 16u) ? ((code & 1u) ? epi.on : epi.neg) : 0.f;
            v.y *= (code & 32u) ? ((code & 2u) ? epi.on : epi.neg) : 0.f;
            v.z *= (code & 64u) ? ((code & 4u) ? epi.on : epi.neg) : 0.f;
            v.w *= (code & 128u) ? ((code & 8u) ? epi.on : epi.neg) : 0.f;
          }
          *reinterpret_cast<float4*>(crow + j * 4) = v;
          vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
      }
    }
    if (epi.amax_out) {      // warp-uniform branch
      vmax = warp_max(vmax);
      if (lane == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(epi.amax_out), __float_as_uint(vmax));   // non-negative floats order like uints
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();              // nobody exits while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// =====================================================================================================================
// cta_group::2 variant: a PAIR of CTAs (cluster {1,2,1}) computes a 256 x BN tile with one tcgen05.mma.cta_group::2 stream
// issued by the leader CTA.  Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows or BN/64 boxes); the
// tensor cores of both SMs read both halves, so per SM and k-block the operand traffic is 64 KB instead of 96 KB from L2
// AND the shared-memory read traffic of the MMAs drops from A + B to A + B/2 (the single-CTA kernel is shared-memory /
// L2 bandwidth bound at ~55 % tensor-pipe utilisation).  Accumulator rows [0,128) live in the leader's TMEM, [128,256) in the
// peer's; each CTA's epilogue warps drain their own TMEM (chunked promotion as above).
//   barriers:  full[s]  (leader)  <- TMA bytes of BOTH CTAs (peer's loads signal the leader's barrier: peer bit cleared)
//              empty[s] (both)    <- leader's tcgen05.commit.cta_group::2 multicast
//              tfull[b] (both)    <- leader's commit multicast after a chunk
//              tempty[b](leader)  <- 8 + 8 epilogue warps (the peer's arrive remotely through mapa)
// =====================================================================================================================
template <bool F16>
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (same offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar) {   // arrive on the barrier at this offset in CTA rank 0 of the cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(bar));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int BN, int STAGES, bool TN, bool F16, bool SLAB>
__global__ void __launch_bounds__(kPairThreads, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                        const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                        float* __restrict__ C, int64_t ldc, int64_t split_stride, int M, int n_store, int k_blocks_total,
                        int k_blocks_per_split, uint64_t mn_desc_bits, const GemmEpilogue epi, int n_tiles_n, int n_m_pairs,
                        int n_work, int kChunk) {
  constexpr int BK = 32;
  constexpr int BKE = F16 ? 2 * BK : BK;                  // reduction elements per k-block
  constexpr int BOXC = F16 ? 64 : 32;                     // TN: MN elements per box row
  constexpr int kABytes = kBM * BK * 4;
  constexpr int BH = BN / 2;                              // B rows (NT) / columns (TN) staged by one CTA
  static_assert(BN % 64 == 0, "the pair kernel splits B into two halves of whole 32-wide blocks");
  static_assert(!(TN && F16) || BH % 64 == 0, "fp16 MN-major halves are made of 64-column boxes");
  constexpr int BH_BYTES = BH * BK * 4;
  constexpr int STAGE_BYTES = 2 * kABytes + 2 * BH_BYTES;
  constexpr int BOX_BYTES = BKE * 128;
  constexpr int SL = BN / 4;                              // columns drained / stored by one epilogue warp (4 warps per TMEM lane quadrant)
  static_assert(BN % 128 == 0, "the pair kernel's 16 epilogue warps take BN / 4 columns each, in chunks of 32");
  constexpr uint32_t TMEM_COLS = 512;
  constexpr int NBUF = 512 / BN;                          // accumulator buffers in TMEM: 2 (BN = 256) or 4 (BN = 128).  With 4 the MMA
                                                          // stream can run a whole short-K tile ahead of an epilogue that is stuck in its stores
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);
  float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);    // SLAB: 16 epilogue warps x [32][36] floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + NBUF);
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 32);          // 16 epilogue warps of each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_trigger();
  if (warp == 1) {   // both CTAs, same warp id: paired TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  if (!pdl_wait_guard()) __trap();             // (never taken, see tx_common.cuh) barriers / TMEM were set up while the previous kernel drained
  // PERSISTENT: a 1-D grid of 2 x n_clusters CTAs (cluster {2,1,1}); cluster c walks the work items c, c + n_clusters, ... of the
  // linear list (n tile fastest, then M-tile pair, then k split), so that the epilogue of one tile (accumulator drain + global
  // stores) overlaps the TMA / MMA stream of the next: the smem ring and the two TMEM buffers simply keep rotating across tiles.
  const int n_clusters = (int)(gridDim.x >> 1), cluster_id = (int)(blockIdx.x >> 1);
  auto work_coords = [&](int w, int& m0, int& n0, int& kb0, int& nkb, int& z) {
    const int n_idx = w % n_tiles_n, rest = w / n_tiles_n;
    const int mp = rest % n_m_pairs;
    z = rest / n_m_pairs;
    m0 = (mp * 2 + (int)crank) * kBM;
    n0 = n_idx * BN;
    kb0 = z * k_blocks_per_split;
    nkb = max(0, min(k_blocks_per_split, k_blocks_total - kb0));
  };

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer (both CTAs) =====
      int gkb = 0;       // k-blocks issued so far by this CTA (ring position)
      for (int w = cluster_id; w < n_work; w += n_clusters) {
      int m0, n0, kb0, nkb, z;
      work_coords(w, m0, n0, kb0, nkb, z);
      for (int kb = 0; kb < nkb; ++kb, ++gkb) {
        const int s = gkb % STAGES;
        const uint32_t ph = (gkb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t full = full0 + 8 * s;
        if (leader) mbar_expect_tx(full, 2 * STAGE_BYTES);       // bytes of both CTAs land on the leader's barrier
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
        const int kk = (kb0 + kb) * BKE;
        if constexpr (!TN) {
          tma_load_2d_pair(base, &map_a_hi, full, kk, m0);
          tma_load_2d_pair(base + kABytes, &map_a_lo, full, kk, m0);
          tma_load_2d_pair(base + 2 * kABytes, &map_b_hi, full, kk, n0 + (int)crank * BH);
          tma_load_2d_pair(base + 2 * kABytes + BH_BYTES, &map_b_lo, full, kk, n0 + (int)crank * BH);
        } else {
#pragma unroll
          for (int b = 0; b < kBM / BOXC; ++b) {
            tma_load_2d_pair(base + b * BOX_BYTES, &map_a_hi, full, m0 + BOXC * b, kk);
            tma_load_2d_pair(base + kABytes + b * BOX_BYTES, &map_a_lo, full, m0 + BOXC * b, kk);
          }
#pragma unroll
          for (int b = 0; b < BH / BOXC; ++b) {
            tma_load_2d_pair(base + 2 * kABytes + b * BOX_BYTES, &map_b_hi, full, n0 + (int)crank * BH + BOXC * b, kk);
            tma_load_2d_pair(base + 2 * kABytes + BH_BYTES + b * BOX_BYTES, &map_b_lo, full, n0 + (int)crank * BH + BOXC * b, kk);
          }
        }
      }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {   // ===== MMA issuer (leader CTA only) =====
      const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (TN ? ((1u << 15) | (1u << 16)) : 0u) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);       // M = 256 across the pair
      int gkb = 0, gch = 0;     // k-blocks / chunks consumed so far (ring and TMEM buffer positions)
      for (int w = cluster_id; w < n_work; w += n_clusters) {
      int m0, n0, kb0, nkb, z;
      work_coords(w, m0, n0, kb0, nkb, z);
      const int n_chunks = (nkb + kChunk - 1) / kChunk;
      for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
        const int buf = gch % NBUF;
        mbar_wait(tempty0 + 8 * buf, ((gch / NBUF) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
        const int kb_end = min(nkb, (ch + 1) * kChunk);
        for (int kb = ch * kChunk; kb < kb_end; ++kb, ++gkb) {
          const int s = gkb % STAGES;
          const uint32_t ph = (gkb / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
          uint64_t a_hi, a_lo, b_hi, b_lo;
          if constexpr (!TN) {
            a_hi = umma_desc_k_sw128(base); a_lo = umma_desc_k_sw128(base + kABytes);
            b_hi = umma_desc_k_sw128(base + 2 * kABytes); b_lo = umma_desc_k_sw128(base + 2 * kABytes + BH_BYTES);
          } else {
            a_hi = umma_desc_mn(base, mn_desc_bits); a_lo = umma_desc_mn(base + kABytes, mn_desc_bits);
            b_hi = umma_desc_mn(base + 2 * kABytes, mn_desc_bits); b_lo = umma_desc_mn(base + 2 * kABytes + BH_BYTES, mn_desc_bits);
          }
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = TN ? (uint64_t)((k * (F16 ? 2048 : 1024)) >> 4) : (uint64_t)(2 * k);
            const uint32_t first = (kb == ch * kChunk && k == 0) ? 0u : 1u;
            umma_tf32_pair<F16>(tacc, a_lo + adv, b_hi + adv, idesc, first);
            umma_tf32_pair<F16>(tacc, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_tf32_pair<F16>(tacc, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          tcgen05_commit_pair(empty0 + 8 * s);                   // stage free in both CTAs
        }
        tcgen05_commit_pair(tfull0 + 8 * buf);                   // chunk accumulator complete in both CTAs
      }
      }
    }
  } else {             // ===== epilogue warps (both CTAs): own TMEM lanes = own 128 rows =====
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;                          // column slice [hsel SL, (hsel + 1) SL)
    const float inv_scale = (epi.scale_a ? 1.f / __ldg(epi.scale_a) : 1.f) * (epi.scale_b ? 1.f / __ldg(epi.scale_b) : 1.f);   // powers of two: exact
    float vmax = 0.f;
    int gch = 0;
    for (int w = cluster_id; w < n_work; w += n_clusters) {
    int m0, n0, kb0, nkb, z;
    work_coords(w, m0, n0, kb0, nkb, z);
    const int n_chunks = (nkb + kChunk - 1) / kChunk;
    float acc[SL];
#pragma unroll
    for (int c = 0; c < SL; ++c) acc[c] = 0.f;
    // sign / keep bytes of this thread's row: requested before the accumulator chunks arrive, so the loads are off the store path
    const int row = m0 + q * 32 + lane;
    uint32_t mw[SL / 16];
    const bool masked = !TN && epi.mask != nullptr;
    if (masked && row < M) {
#pragma unroll
      for (int w = 0; w < SL / 16; ++w) {
        const int col = n0 + hsel * SL + w * 16;
        mw[w] = (col < epi.feat_cols)
                    ? __ldg(reinterpret_cast<const uint32_t*>(epi.mask + (int64_t)row * epi.stride + (col >> 2))) : 0xFFFFFFFFu;
      }
    }
    for (int ch = 0; ch < n_chunks; ++ch, ++gch) {
      const int buf = gch % NBUF;
      mbar_wait(tfull0 + 8 * buf, (gch / NBUF) & 1);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN + hsel * SL) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < SL / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(tacc + (uint32_t)(c * 16), r);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty0 + 8 * buf); else mbar_arrive_cta0(tempty0 + 8 * buf);
      }
    }
    if constexpr (SLAB) {
      // Stores through a per-warp shared-memory slab (32 rows x 32 columns, pitch 36 floats: conflict-free both ways).  A thread owns
      // one ROW of the accumulator, so storing straight from registers makes every STG touch 32 different lines with 16 bytes each
      // (half sectors); ncu shows the epilogue warps stalled on exactly those stores (the next write of the store's source register
      // waits until the LSU has drained it).  After the transpose one STG writes 4 rows x 128 contiguous bytes: full lines.
      float* slab = stg + (warp - 2) * (32 * 36);
      const int row_base = m0 + q * 32;
#pragma unroll
      for (int g2 = 0; g2 < SL / 32; ++g2) {
        const int colg = n0 + hsel * SL + g2 * 32;
        if (colg < n_store) {                                     // warp-uniform
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = g2 * 8 + u;
            const int col = colg + u * 4;
            float4 v = make_float4(acc[4 * j] * inv_scale, acc[4 * j + 1] * inv_scale, acc[4 * j + 2] * inv_scale, acc[4 * j + 3] * inv_scale);
            if (row < M && col < n_store) {
              if (masked && col < epi.feat_cols) {
                uint32_t code = (mw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                if (!epi.has_keep) code |= 0xF0u;
                v.x *= (code & 16u) ? ((code & 1u) ? epi.on : epi.neg) : 0.f;
                v.y *= (code & 32u) ? ((code & 2u) ? epi.on : epi.neg) : 0.f;
                v.z *= (code & 64u) ? ((code & 4u) ? epi.on : epi.neg) : 0.f;
                v.w *= (code & 128u) ? ((code & 8u) ? epi.on : epi.neg) : 0.f;
              }
              vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
            *reinterpret_cast<float4*>(slab + lane * 36 + u * 4) = v;
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int rl = 4 * t + (lane >> 3), c4 = lane & 7;
            const int grow = row_base + rl, col = colg + c4 * 4;
            if (grow < M && col < n_store)
              *reinterpret_cast<float4*>(C + (int64_t)z * split_stride + (int64_t)grow * ldc + col) = *reinterpret_cast<const float4*>(slab + rl * 36 + c4 * 4);
          }
          __syncwarp();
        }
      }
    } else if (row < M) {
      float* crow = C + (int64_t)z * split_stride + (int64_t)row * ldc + n0 + hsel * SL;
#pragma unroll
      for (int j = 0; j < SL / 4; ++j) {
        const int col = n0 + hsel * SL + j * 4;
        if (col < n_store) {
          float4 v = make_float4(acc[4 * j] * inv_scale, acc[4 * j + 1] * inv_scale, acc[4 * j + 2] * inv_scale, acc[4 * j + 3] * inv_scale);
          if (masked && col < epi.feat_cols) {
            uint32_t code = (mw[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            if (!epi.has_keep) code |= 0xF0u;
            v.x *= (code & 16u) ? ((code & 1u) ? epi.on : epi.neg) : 0.f;
            v.y *= (code & 32u) ? ((code & 2u) ? epi.on : epi.neg) : 0.f;
            v.z *= (code & 64u) ? ((code & 4u) ? epi.on : epi.neg) : 0.f;
            v.w *= (code & 128u) ? ((code & 8u) ? epi.on : epi.neg) : 0.f;
          }
          *reinterpret_cast<float4*>(crow + j * 4) = v;
          vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
      }
    }
    }   // next work item
    if (epi.amax_out) {      // warp-uniform branch
      vmax = warp_max(vmax);
      if (lane == 0 && vmax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(epi.amax_out), __float_as_uint(vmax));   // non-negative floats order like uints
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// hi = rn_tf32(x), lo = rn_tf32(x - hi) (x - hi is exact in fp32); round-to-nearest keeps the split error signed and
// unbiased (truncation left a systematic 1e-6 drift over K = 2050).  Outputs padded to ldo with zeros.
__device__ __forceinline__ float rn_tf32(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);   // round half away in magnitude, low 13 bits cleared
}

__global__ void split_tf32_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols, float* __restrict__ hi,
                                  float* __restrict__ lo, int ldo) {
  const int vec_per_row = ldo >> 2;
  const int64_t total = (int64_t)rows * vec_per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / vec_per_row);
    const int c0 = (int)(t - (int64_t)i * vec_per_row) << 2;
    float h[4], l[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      const float v = c < cols ? __ldg(x + (int64_t)i * ldx + c) : 0.f;
      const float vh = rn_tf32(v);
      h[u] = vh;
      l[u] = rn_tf32(v - vh);
    }
    *reinterpret_cast<float4*>(hi + (int64_t)i * ldo + c0) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + (int64_t)i * ldo + c0) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// ---- fp16 split path: per-tensor bounds and scales live in DEVICE scalars, so nothing here synchronises with the host ----
__global__ void absmax_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols, float* __restrict__ out) {
  TX_PDL_ENTER();
  float m = 0.f;
  if (ldx == cols && aligned16(x)) {                            // a contiguous block (e.g. [N, 250] features): flat 128-bit loads + tail
    const int64_t total = (int64_t)rows * cols, n4 = total >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = __ldg(x4 + t);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(total - (n4 << 2))) m = fmaxf(m, fabsf(__ldg(x + (n4 << 2) + threadIdx.x)));
  } else if ((ldx & 3) == 0 && (cols & 3) == 0 && aligned16(x)) {      // 128-bit loads, one division per 4 elements
    const int vpr = cols >> 2;
    const int64_t total = (int64_t)rows * vpr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      const int i = (int)(t / vpr);
      const int c = (int)(t - (int64_t)i * vpr) << 2;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)i * ldx + c));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
      const int i = (int)(t / cols);
      const int c = (int)(t - (int64_t)i * cols);
      m = fmaxf(m, fabsf(__ldg(x + (int64_t)i * ldx + c)));
    }
  }
  // one atomic per CTA: ~9.5 k same-address atomics (one per warp) serialised in L2 and cost 10 us on a 16 MB input
  __shared__ float s_m[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s_m[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
  }
}

// max |v[0 .. len)| over one 256-thread block (small parameter vectors / tables), returned to every thread
__device__ __forceinline__ float block_absmax(const float* __restrict__ v, int64_t len, float* s_red) {
  float m = 0.f;
  if (v) for (int64_t t = threadIdx.x; t < len; t += blockDim.x) m = fmaxf(m, fabsf(__ldg(v + t)));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s_red[w]);
  __syncthreads();
  return r;
}
// out = max(a ca, max|b[0..b_len)| cb)  (b may be null)
__global__ void bound_max2_kernel(const float* a, float ca, const float* b, int64_t b_len, float cb, float* out) {
  TX_PDL_ENTER();
  __shared__ float s_red[8];
  const float mb = block_absmax(b, b_len, s_red);
  if (threadIdx.x == 0) *out = fmaxf(*a * ca, mb * cb);
}
// bounds of |dft| (see tx_bound_dft): out[0] = g (c_direct + c_attn ft c) with c = max(max|attn_l|, max|attn_r|, 2^-20) - rigorous;
// out[1] = min(out[0], g c_optimistic) - the scale the star backward tries first; out[2] = c; out[3] = 0 (its fp16-range flag)
__global__ void bound_dft_kernel(const float* g, const float* ft, const float* al, const float* ar, int64_t len, float c_direct, float c_attn,
                                 float c_optimistic, float* out) {
  TX_PDL_ENTER();
  __shared__ float s_red[8];
  const float ml = block_absmax(al, len, s_red), mr = block_absmax(ar, len, s_red);
  if (threadIdx.x == 0) {
    const float c = fmaxf(fmaxf(ml, mr), 9.5367431640625e-07f);
    const float rigorous = *g * (c_direct + c_attn * *ft * c);
    out[0] = rigorous;
    out[1] = c_optimistic > 0.f ? fminf(rigorous, *g * c_optimistic) : rigorous;
    out[2] = c;
    reinterpret_cast<int*>(out)[3] = 0;
  }
}

__global__ void split_f16_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols, const float* __restrict__ bound,
                                 __half* __restrict__ hi, __half* __restrict__ lo, int ldo, float* __restrict__ scale_out) {
  TX_PDL_ENTER();
  const float scale = f16_split_scale(__ldg(bound));
  if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) *scale_out = scale;
  const int vec_per_row = ldo >> 2;
  const int64_t total = (int64_t)rows * vec_per_row;
  const bool vec_in = (ldx & 3) == 0 && aligned16(x);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / vec_per_row);
    const int c0 = (int)(t - (int64_t)i * vec_per_row) << 2;
    float4 v;
    if (vec_in && c0 + 3 < cols) {
      v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)i * ldx + c0));
    } else {
      v.x = c0 < cols ? __ldg(x + (int64_t)i * ldx + c0) : 0.f;
      v.y = c0 + 1 < cols ? __ldg(x + (int64_t)i * ldx + c0 + 1) : 0.f;
      v.z = c0 + 2 < cols ? __ldg(x + (int64_t)i * ldx + c0 + 2) : 0.f;
      v.w = c0 + 3 < cols ? __ldg(x + (int64_t)i * ldx + c0 + 3) : 0.f;
    }
    uint2 h, l;
    f16_split4(v, scale, h, l);
    *reinterpret_cast<uint2*>(hi + (int64_t)i * ldo + c0) = h;
    *reinterpret_cast<uint2*>(lo + (int64_t)i * ldo + c0) = l;
  }
}

// One launch for a WEIGHT matrix w [rows, cols]: max|w| (grid-wide, through a device counter: the grid is at most kSplitWCtasPerSm
// CTAs per SM and the launch bounds guarantee that many fit, so every CTA is resident and the spin below cannot deadlock), then
// the fp16 hi/lo split both row-major [rows, ld] (forward
// projection operand) and transposed [cols, ldt] (input-gradient operand) with the same scale.
constexpr int kSplitWCtasPerSm = 4;
__global__ void __launch_bounds__(256, kSplitWCtasPerSm) split_f16_weight_kernel(const float* __restrict__ w, int64_t ldw, int rows, int cols,
                                                               __half* __restrict__ hi, __half* __restrict__ lo, int ld,
                                                               __half* __restrict__ hit, __half* __restrict__ lot, int ldt,
                                                               float* __restrict__ amax, unsigned int* __restrict__ counter,
                                                               float* __restrict__ scale_out) {
  TX_PDL_ENTER();
  const int64_t total = (int64_t)rows * cols;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  float m = 0.f;
  if (ldw == cols && (total & 3) == 0 && aligned16(w)) {       // a contiguous parameter tensor: flat 128-bit loads, no index arithmetic
    const float4* w4 = reinterpret_cast<const float4*>(w);
    for (int64_t t = tid; t < (total >> 2); t += nth) {
      const float4 v = __ldg(w4 + t);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    for (int64_t t = tid; t < total; t += nth) {
      const int i = (int)(t / cols), c = (int)(t - (int64_t)i * cols);
      m = fmaxf(m, fabsf(__ldg(w + (int64_t)i * ldw + c)));
    }
  }
  __shared__ float s_m[8];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {                   // one atomicMax per CTA (same-address atomics serialise in L2), then the arrival
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
    if (m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < gridDim.x) __nanosleep(32);
  }
  __syncthreads();
  const float scale = f16_split_scale(__ldcg(amax));
  if (scale_out && tid == 0) *scale_out = scale;
  const float c65 = 65504.f;
  // 32 x 32 tiles: coalesced reads, row-major outputs written directly, transposed outputs through a shared-memory tile.
  // The tile grid covers the padded extents (ld columns, ldt transposed columns): padding is written as zeros.
  __shared__ __half s_hi[32][34], s_lo[32][34];
  const int tr = (max(rows, ldt) + 31) / 32, tc = (max(cols, ld) + 31) / 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;      // 8 warps x 4 rows each
  for (int tile = blockIdx.x; tile < tr * tc; tile += gridDim.x) {
    const int i0 = (tile / tc) * 32, c0 = (tile % tc) * 32;
    const int c = c0 + lane;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int il = wy * 4 + r, i = i0 + il;
      const float x = (i < rows && c < cols) ? fminf(fmaxf(__ldg(w + (int64_t)i * ldw + c) * scale, -c65), c65) : 0.f;
      const __half h = __float2half_rn(x);
      const __half l = __float2half_rn(x - __half2float(h));
      if (i < rows && c < ld) { hi[(int64_t)i * ld + c] = h; lo[(int64_t)i * ld + c] = l; }
      s_hi[il][lane] = h; s_lo[il][lane] = l;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int cl = wy * 4 + r, cc = c0 + cl, i = i0 + lane;      // transposed row cc, transposed column i
      if (cc < cols && i < ldt) { hit[(int64_t)cc * ldt + i] = s_hi[lane][cl]; lot[(int64_t)cc * ldt + i] = s_lo[lane][cl]; }
    }
    __syncthreads();
  }
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (floats); box = [box_rows, 32 cols], 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int box_cols = 32, bool f16 = false) {
  if (rows <= 0 || cols <= 0) {
    set_error("gemm: empty operand");
    return TX_ERR_INVALID_ARGUMENT;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("gemm: cuTensorMapEncodeTiled driver entry point unavailable");
    return TX_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * (f16 ? 2 : 4)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm: cuTensorMapEncodeTiled failed with CUresult %d (rows %lld cols %lld ld %lld)", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return TX_ERR_CUDA;
  }
  return TX_OK;
}

// Granularity of the last tile's MMA width in the single-CTA kernel: `natural` = the finest the operand layout allows (16 columns for
// K-major B; whole boxes for MN-major B).  TAXO_GEMM_DYN_N=0 switches the narrowing off, =<g> forces a coarser multiple of `natural`.
static int dyn_n_gran(int natural) {
  static int v = -2;
  if (v == -2) { const char* e = getenv("TAXO_GEMM_DYN_N"); v = e ? atoi(e) : -1; }
  if (v < 0) return natural;
  if (v == 0) return 0;
  return ((v + natural - 1) / natural) * natural;
}

template <int BN, int STAGES, bool TN, int BK = 32, int CL = 1, bool F16 = false>
static int launch_gemm(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                       float* c, int64_t ldc, int64_t split_stride, int64_t M, int64_t N, int64_t K, int splits, cudaStream_t st,
                       const GemmEpilogue& epi = GemmEpilogue{nullptr, 0, 1, 0, 0, 0, 1.f, 1.f, nullptr, nullptr, nullptr}) {
  alignas(64) CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  constexpr int kBK = BK;
  constexpr int kABytes = kBM * BK * 4;
  if (!TN) {   // A [M, K], B [N, K]: box = {BK k-cols, tile rows}; 128-byte (BK = 32) or 64-byte (BK = 16) swizzle
    const CUtensorMapSwizzle swk = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    constexpr int BKE = F16 ? 2 * BK : BK;
    if ((rc = make_map(&ma_hi, a_hi, M, K, lda, kBM, swk, BKE, F16)) != TX_OK) return rc;
    if ((rc = make_map(&ma_lo, a_lo, M, K, lda, kBM, swk, BKE, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, BN / CL, swk, BKE, F16)) != TX_OK) return rc;   // CL = 2: each CTA fetches half of the rows
    if ((rc = make_map(&mb_lo, b_lo, N, K, ldb, BN / CL, swk, BKE, F16)) != TX_OK) return rc;
  } else {     // A [K, M], B [K, N]: box = {32 m/n-cols, kBK reduction rows}
    // (debug knobs TAXO_TN_SWIZZLE / _LAYOUT / _LBO / _SBO override the layout constants below)
    const char* e_sw = getenv("TAXO_TN_SWIZZLE");
    // fp16: plain SWIZZLE_128B boxes of {64 cols, 64 reduction rows}; tf32: SWIZZLE_128B_ATOM_32B boxes of {32 cols, 32 rows}
    const CUtensorMapSwizzle sw = F16 ? CU_TENSOR_MAP_SWIZZLE_128B : (e_sw ? (CUtensorMapSwizzle)atoi(e_sw) : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    constexpr int KROWS = F16 ? 2 * BK : BK, BOXC = F16 ? 64 : 32;
    if ((rc = make_map(&ma_hi, a_hi, K, M, lda, KROWS, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&ma_lo, a_lo, K, M, lda, KROWS, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_hi, b_hi, K, N, ldb, KROWS, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_lo, b_lo, K, N, ldb, KROWS, sw, BOXC, F16)) != TX_OK) return rc;
  }
  constexpr int STAGE_BYTES = 2 * kABytes + 2 * BN * kBK * 4;
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers + tmem slot */;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, STAGES, TN, BK, CL, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%zu B smem) failed: %s", SMEM, cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
    attr_done = true;
  }
  constexpr int kBKE = F16 ? 2 * BK : BK;      // reduction elements per k-block
  const int kbt = (int)((K + kBKE - 1) / kBKE);
  const int kbs = (kbt + splits - 1) / splits;
  const int64_t n_store = ((N + 3) / 4) * 4;
  const unsigned m_tiles = (unsigned)((M + kBM - 1) / kBM);
  dim3 grid((unsigned)((N + BN - 1) / BN), (m_tiles + CL - 1) / CL * CL, (unsigned)splits);   // padded M tiles are fully masked
  const char* e_l = getenv("TAXO_TN_LAYOUT");
  const char* e_lbo = getenv("TAXO_TN_LBO");
  const char* e_sbo = getenv("TAXO_TN_SBO");
  // MN-major descriptors: LBO = bytes between consecutive boxes (one box = k-rows x 128 B), SBO = bytes between k-row groups of the
  // swizzle atom (tf32 / SWIZZLE_128B_BASE32B: 4 rows = 512 B; fp16 / SWIZZLE_128B: 8 rows = 1024 B)
  const uint64_t mn_bits = F16 ? umma_desc_mn_bits((uint32_t)(kBKE * 128), 1024u, 2u)
                               : umma_desc_mn_bits(e_lbo ? (uint32_t)atoi(e_lbo) : (uint32_t)(kBK * 128), e_sbo ? (uint32_t)atoi(e_sbo) : 512u,
                                                   e_l ? (uint32_t)atoi(e_l) : 1u);
  const int gran = dyn_n_gran(TN ? (F16 ? 64 : 32) : 16);
  if (CL == 1) {
    TX_PDL_LAUNCH((gemm_tf32x3_kernel<BN, STAGES, TN, BK, CL, F16>), grid, kGemmThreads, SMEM, st, ma_hi, ma_lo, mb_hi, mb_lo, c, ldc, split_stride, (int)M,
                  (int)n_store, kbt, kbs, mn_bits, epi, gran);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = CL;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled(TX_PDL_GROUP) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tf32x3_kernel<BN, STAGES, TN, BK, CL, F16>, ma_hi, ma_lo, mb_hi, mb_lo, c, ldc, split_stride,
                                       (int)M, (int)n_store, kbt, kbs, mn_bits, epi, gran);
    if (e != cudaSuccess) {
      set_error("gemm: cluster launch failed: %s", cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
  }
  TX_LAUNCH_CHECK(TN ? "tx_gemm_tn_tf32x3" : "tx_gemm_nt_tf32x3");
  return TX_OK;
}

template <int BN, int STAGES, bool TN, bool F16 = false, bool SLAB = false>
static int launch_gemm_pair(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                            float* c, int64_t ldc, int64_t split_stride, int64_t M, int64_t N, int64_t K, int splits, cudaStream_t st,
                            const GemmEpilogue& epi = GemmEpilogue{nullptr, 0, 1, 0, 0, 0, 1.f, 1.f, nullptr, nullptr, nullptr}) {
  alignas(64) CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  constexpr int BK = 32;
  constexpr int BKE = F16 ? 2 * BK : BK;
  if (!TN) {
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    if ((rc = make_map(&ma_hi, a_hi, M, K, lda, kBM, sw, BKE, F16)) != TX_OK) return rc;
    if ((rc = make_map(&ma_lo, a_lo, M, K, lda, kBM, sw, BKE, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, BN / 2, sw, BKE, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_lo, b_lo, N, K, ldb, BN / 2, sw, BKE, F16)) != TX_OK) return rc;
  } else {
    const CUtensorMapSwizzle sw = F16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    constexpr int BOXC = F16 ? 64 : 32;
    if ((rc = make_map(&ma_hi, a_hi, K, M, lda, BKE, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&ma_lo, a_lo, K, M, lda, BKE, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_hi, b_hi, K, N, ldb, BKE, sw, BOXC, F16)) != TX_OK) return rc;
    if ((rc = make_map(&mb_lo, b_lo, K, N, ldb, BKE, sw, BOXC, F16)) != TX_OK) return rc;
  }
  constexpr int STAGE_BYTES = 2 * kBM * BK * 4 + 2 * (BN / 2) * BK * 4;
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + (SLAB ? 16 * 32 * 36 * 4 : 0);
  static_assert(SMEM <= 227 * 1024, "pair kernel: shared memory budget");
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<BN, STAGES, TN, F16, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) {
      set_error("gemm(pair): cudaFuncSetAttribute(%zu B smem) failed: %s", SMEM, cudaGetErrorString(e));
      return TX_ERR_CUDA;
    }
    attr_done = true;
  }
  const int kbt = (int)((K + BKE - 1) / BKE);
  const int kbs = (kbt + splits - 1) / splits;
  const int64_t n_store = ((N + 3) / 4) * 4;
  const unsigned m_tiles = (unsigned)((M + kBM - 1) / kBM);
  const int n_tiles_n = (int)((N + BN - 1) / BN), n_m_pairs = (int)((m_tiles + 1) / 2);   // a padded odd M tile is fully masked
  const int64_t n_work64 = (int64_t)n_tiles_n * n_m_pairs * splits;
  if (n_work64 >= INT32_MAX) {
    set_error("gemm(pair): too many tiles");
    return TX_ERR_INVALID_ARGUMENT;
  }
  const int n_work = (int)n_work64;
  static int persist = -1;     // TAXO_GEMM_PERSIST=0 -> one cluster per tile (the pre-persistent behaviour)
  if (persist < 0) { const char* e_p = getenv("TAXO_GEMM_PERSIST"); persist = (e_p && atoi(e_p) == 0) ? 0 : 1; }
  const int n_clusters = persist ? (n_work < kNumSms / 2 ? n_work : kNumSms / 2) : n_work;
  // k-blocks accumulated inside the tensor core between promotions to fp32 registers (see gemm_tf32x3_kernel): 2 by default (24 MMAs);
  // TAXO_GEMM_CHUNK overrides (longer chunks = fewer TMEM drains, more accumulated truncation bias)
  static int chunk_env = -1;
  if (chunk_env < 0) { const char* e_c = getenv("TAXO_GEMM_CHUNK"); chunk_env = e_c ? atoi(e_c) : 0; }
  const int chunk_kb = chunk_env > 0 ? chunk_env : kChunkK / BK;
  dim3 grid(2u * (unsigned)n_clusters, 1, 1);
  const uint64_t mn_bits = F16 ? umma_desc_mn_bits((uint32_t)(BKE * 128), 1024u, 2u) : umma_desc_mn_bits((uint32_t)(BK * 128), 512u, 1u);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled(TX_PDL_GROUP) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tf32x3_pair_kernel<BN, STAGES, TN, F16, SLAB>, ma_hi, ma_lo, mb_hi, mb_lo, c, ldc, split_stride,
                                     (int)M, (int)n_store, kbt, kbs, mn_bits, epi, n_tiles_n, n_m_pairs, n_work, chunk_kb);
  if (e != cudaSuccess) {
    set_error("gemm(pair): cluster launch failed: %s", cudaGetErrorString(e));
    return TX_ERR_CUDA;
  }
  return TX_OK;
}

}  // namespace tx

using namespace tx;

extern "C" {

static bool use_pair() {      // cta_group::2 pair kernel for 256-wide tiles (default); TAXO_GEMM_PAIR=0 -> single-CTA MMAs
  static int v = -1;
  if (v < 0) { const char* e = getenv("TAXO_GEMM_PAIR"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

static bool use_cluster() {   // TAXO_GEMM_CLUSTER=1 -> no 2-CTA multicast clusters
  static int v = -1;
  if (v < 0) { const char* e = getenv("TAXO_GEMM_CLUSTER"); v = (e && atoi(e) == 1) ? 0 : 1; }
  return v != 0;
}

// tile width: 64 / 128 for narrow outputs, else 256 unless 160-wide tiles waste fewer MMA columns (e.g. N = 300: 2 x 160 vs 2 x 256)
static int pick_bn(int64_t n) {
  if (n <= 64) return 64;
  if (n <= 128) return 128;
  const int64_t w256 = ((n + 255) / 256) * 256, w160 = ((n + 159) / 160) * 160;
  return w160 * 10 < w256 * 8 ? 160 : 256;      // switch only for a > 20 % saving
}

int64_t tx_gemm_tn_splits(int64_t m, int64_t n, int64_t r) {
  const int64_t bn = pick_bn(n);
  const int64_t tiles = ((m + kBM - 1) / kBM) * ((n + bn - 1) / bn);
  int64_t s = kNumSms / (tiles > 0 ? tiles : 1);
  const int64_t kbt = (r + kBK - 1) / kBK;
  if (s > kbt / 8) s = kbt / 8;          // keep at least 8 k-blocks per split
  if (s > 32) s = 32;
  return s < 1 ? 1 : s;
}

int tx_gemm_tn_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                      float* c_partial, int64_t ldc, int64_t split_stride, int64_t m, int64_t n, int64_t r, int64_t splits,
                      void* stream) {
  TX_REQUIRE(m > 0 && n > 0 && r > 0 && m < INT32_MAX && n < INT32_MAX && r < INT32_MAX, "gemm_tn: bad shape");
  TX_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && lda >= m && ldb >= n, "gemm_tn: operand row pitch must be a multiple of 4 floats and >= M / N");
  TX_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(c_partial), "gemm_tn: 16-byte aligned pointers required");
  TX_REQUIRE(ldc % 4 == 0 && ldc >= ((n + 3) / 4) * 4 && split_stride % 4 == 0 && split_stride >= m * ldc, "gemm_tn: bad ldc / split_stride");
  TX_REQUIRE(splits >= 1 && splits <= 65535, "gemm_tn: bad split count");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_pair() && m > kBM && n > 128 && pick_bn(n) == 256)
    return launch_gemm_pair<256, 3, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
  if (use_cluster() && m > kBM) {
    if (n > 128 && pick_bn(n) == 160) return launch_gemm<160, 3, true, 32, 2>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
    if (n > 128) return launch_gemm<256, 2, true, 32, 2>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
  }
  if (n > 128 && pick_bn(n) == 160) return launch_gemm<160, 3, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
  if (n > 128) return launch_gemm<256, 2, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
  if (n > 64) return launch_gemm<128, 3, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
  return launch_gemm<64, 4, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st);
}

int tx_split_tf32(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ldo, void* stream) {
  TX_REQUIRE(ldo % 4 == 0 && ldo >= cols && aligned16(hi) && aligned16(lo), "split_tf32: outputs need ld %% 4 == 0, ld >= cols, 16B alignment");
  TX_REQUIRE(rows >= 0 && cols >= 0 && rows < INT32_MAX && ldo < INT32_MAX, "split_tf32: bad shape");
  if (rows == 0 || cols == 0) return TX_OK;
  const int64_t total = rows * (ldo / 4);
  const int grid = (int)((total + 255) / 256 < (int64_t)kNumSms * 16 ? (total + 255) / 256 : (int64_t)kNumSms * 16);
  split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, (int)rows, (int)cols, hi, lo, (int)ldo);
  TX_LAUNCH_CHECK("tx_split_tf32");
  return TX_OK;
}

int tx_gemm_nt_tf32x3(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                      float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, void* stream) {
  return tx_gemm_nt_tf32x3_ex(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, m, n, k, nullptr, stream);
}

int tx_gemm_nt_tf32x3_ex(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb,
                         float* c, int64_t ldc, int64_t m, int64_t n, int64_t k, const tx_gemm_epilogue* e, void* stream) {
  TX_REQUIRE(m > 0 && n > 0 && k > 0 && m < INT32_MAX && n < INT32_MAX && k < INT32_MAX, "gemm: bad shape %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
  TX_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && lda >= k && ldb >= k, "gemm: operand row pitch must be a multiple of 4 floats (16 B) and >= K");
  TX_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(c), "gemm: 16-byte aligned pointers required");
  TX_REQUIRE(ldc % 4 == 0 && ldc >= n, "gemm: ldc must be a multiple of 4 and >= N");
  cudaStream_t st = (cudaStream_t)stream;
  TX_REQUIRE(ldc >= ((n + 3) / 4) * 4, "gemm: ldc must hold round4(N) columns");
  GemmEpilogue epi{nullptr, 0, 1, 0, 0, 0, 1.f, 1.f, nullptr, nullptr, nullptr};
  if (e && e->act_mask) {
    TX_REQUIRE(e->heads > 0 && e->dim > 0 && e->dim % 4 == 0 && e->mask_stride % 16 == 0 && e->mask_stride >= e->heads * e->dim / 4 &&
               aligned16(e->act_mask), "gemm epilogue: bad mask geometry");
    TX_REQUIRE(e->col0 % 4 == 0 && e->col0 >= 0, "gemm epilogue: col0 must be a non-negative multiple of 4");
    TX_REQUIRE(e->col0 == 0, "gemm epilogue: a column offset is not supported with a fused mask");
    TX_REQUIRE(e->p_drop >= 0.f && e->p_drop < 1.f, "gemm epilogue: p_drop must be in [0,1)");
    epi.mask = e->act_mask; epi.heads = (int)e->heads; epi.dim = (int)e->dim; epi.stride = (int)e->mask_stride;
    epi.feat_cols = (int)(e->heads * e->dim); epi.has_keep = e->has_keep_plane;
    epi.on = 1.f / (1.f - e->p_drop); epi.neg = e->act_slope * epi.on;
  }
  if (use_pair() && m > kBM && n > 128 && pick_bn(n) == 256)
    return launch_gemm_pair<256, 3, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  if (use_cluster() && m > kBM && n > 128) {
    if (pick_bn(n) == 160) return launch_gemm<160, 3, false, 32, 2>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
    return launch_gemm<256, 2, false, 32, 2>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  }
  if (n > 128 && pick_bn(n) == 160) return launch_gemm<160, 3, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  {
    // 256-wide tiles: 64-byte k-blocks (BK = 16) give a 4-deep TMA ring in the same 192 KB instead of 2 stages of BK = 32
    static int bk16 = -1;
    if (bk16 < 0) { const char* e_bk = getenv("TAXO_NT_BK"); bk16 = (e_bk && atoi(e_bk) == 16) ? 1 : 0; }   // measured slower (r20): off by default
    if (n > 128 && bk16) return launch_gemm<256, 4, false, 16>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  }
  if (n > 128) return launch_gemm<256, 2, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  if (n > 64) return launch_gemm<128, 3, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  return launch_gemm<64, 4, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
}


// ------------------------------------------------------------------------------------------------
// fp16-split (3 x kind::f16) GEMMs
// ------------------------------------------------------------------------------------------------
int tx_absmax(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* out, void* stream) {
  TX_REQUIRE(out && rows >= 0 && cols >= 0 && rows < INT32_MAX && cols < INT32_MAX && (rows == 0 || cols == 0 || x), "absmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_preclear && cudaMemsetAsync(out, 0, sizeof(float), st) != cudaSuccess) { set_error("absmax: memset failed"); return TX_ERR_CUDA; }
  if (rows == 0 || cols == 0) return TX_OK;
  const int64_t total = rows * cols;
  const int grid = (int)((total + 1023) / 1024 < (int64_t)kNumSms * 8 ? (total + 1023) / 1024 : (int64_t)kNumSms * 8);
  TX_PDL_LAUNCH((absmax_kernel), grid, 256, 0, st, x, ldx, (int)rows, (int)cols, out);
  TX_LAUNCH_CHECK("tx_absmax");
  return TX_OK;
}

int tx_bound_max2(const float* a, float ca, const float* b, int64_t b_len, float cb, float* out, void* stream) {
  TX_REQUIRE(a && out && b_len >= 0, "bound_max2: bad arguments");
  TX_PDL_LAUNCH((bound_max2_kernel), 1, 256, 0, (cudaStream_t)stream, a, ca, b, b_len, cb, out);
  TX_LAUNCH_CHECK("tx_bound_max2");
  return TX_OK;
}

int tx_bound_dft(const float* g_amax, const float* ft_amax, const float* attn_l, const float* attn_r, int64_t attn_len, float c_direct,
                 float c_attn, float c_optimistic, float* out, void* stream) {
  TX_REQUIRE(g_amax && ft_amax && attn_l && attn_r && attn_len >= 0 && out && aligned16(out), "bound_dft: bad arguments (out: 4 floats, 16-byte aligned)");
  TX_PDL_LAUNCH((bound_dft_kernel), 1, 256, 0, (cudaStream_t)stream, g_amax, ft_amax, attn_l, attn_r, attn_len, c_direct, c_attn, c_optimistic, out);
  TX_LAUNCH_CHECK("tx_bound_dft");
  return TX_OK;
}

int tx_split_f16(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* bound, void* hi, void* lo, int64_t ldo,
                 float* scale_out, void* stream) {
  TX_REQUIRE(ldo % 8 == 0 && ldo >= cols && aligned16(hi) && aligned16(lo), "split_f16: outputs need ld %% 8 == 0, ld >= cols, 16B alignment");
  TX_REQUIRE(rows >= 0 && cols >= 0 && rows < INT32_MAX && ldo < INT32_MAX && bound, "split_f16: bad arguments");
  if (rows == 0 || cols == 0) return TX_OK;
  const int64_t total = rows * (ldo / 4);
  const int grid = (int)((total + 255) / 256 < (int64_t)kNumSms * 16 ? (total + 255) / 256 : (int64_t)kNumSms * 16);
  TX_PDL_LAUNCH((split_f16_kernel), grid, 256, 0, (cudaStream_t)stream, x, ldx, (int)rows, (int)cols, bound, (__half*)hi, (__half*)lo, (int)ldo, scale_out);
  TX_LAUNCH_CHECK("tx_split_f16");
  return TX_OK;
}

int tx_split_f16_weight(const float* w, int64_t ldw, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld, void* hi_t, void* lo_t,
                        int64_t ld_t, float* scratch2, float* scale_out, void* stream) {
  TX_REQUIRE(w && hi && lo && hi_t && lo_t && scratch2 && rows > 0 && cols > 0 && rows < INT32_MAX && cols < INT32_MAX, "split_f16_weight: bad arguments");
  TX_REQUIRE(ld % 8 == 0 && ld >= cols && ld_t % 8 == 0 && ld_t >= rows && aligned16(hi) && aligned16(lo) && aligned16(hi_t) && aligned16(lo_t),
             "split_f16_weight: outputs need ld %% 8 == 0 and 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_preclear && cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), st) != cudaSuccess) { set_error("split_f16_weight: memset failed"); return TX_ERR_CUDA; }
  const int64_t tiles = ((std::max(rows, ld_t) + 31) / 32) * ((std::max(cols, ld) + 31) / 32);      // 32 x 32 tiles of the split phase
  int64_t grid = tiles;
  if (grid > (int64_t)kNumSms * kSplitWCtasPerSm) grid = (int64_t)kNumSms * kSplitWCtasPerSm;     // the grid-wide rendezvous needs every CTA resident
  if (grid < 1) grid = 1;
  TX_PDL_LAUNCH((split_f16_weight_kernel), (int)grid, 256, 0, st, w, ldw, (int)rows, (int)cols, (__half*)hi, (__half*)lo, (int)ld, (__half*)hi_t, (__half*)lo_t,
                                                     (int)ld_t, scratch2, reinterpret_cast<unsigned int*>(scratch2 + 1), scale_out);
  TX_LAUNCH_CHECK("tx_split_f16_weight");
  return TX_OK;
}

static int pick_bn_tn16(int64_t n) {     // MN-major fp16 tiles are made of 64-column boxes
  if (n <= 64) return 64;
  if (n <= 128) return 128;
  const int64_t w256 = ((n + 255) / 256) * 256, w192 = ((n + 191) / 192) * 192;
  return w192 * 10 < w256 * 8 ? 192 : 256;
}

int64_t tx_gemm_tn_f16_splits(int64_t m, int64_t n, int64_t r) {
  const int64_t bn = pick_bn_tn16(n);
  const int64_t tiles = ((m + kBM - 1) / kBM) * ((n + bn - 1) / bn);
  int64_t s = kNumSms / (tiles > 0 ? tiles : 1);
  const int64_t kbt = (r + 63) / 64;
  if (s > kbt / 8) s = kbt / 8;          // keep at least 8 k-blocks per split
  if (s > 32) s = 32;
  return s < 1 ? 1 : s;
}

int tx_gemm_tn_f16x3(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                     const float* scale_a, const float* scale_b, float* c_partial, int64_t ldc, int64_t split_stride, int64_t m,
                     int64_t n, int64_t r, int64_t splits, void* stream) {
  TX_REQUIRE(m > 0 && n > 0 && r > 0 && m < INT32_MAX && n < INT32_MAX && r < INT32_MAX, "gemm_tn_f16: bad shape");
  TX_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= m && ldb >= n, "gemm_tn_f16: operand row pitch must be a multiple of 8 halves and >= M / N");
  TX_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(c_partial), "gemm_tn_f16: 16-byte aligned pointers required");
  TX_REQUIRE(ldc % 4 == 0 && ldc >= ((n + 3) / 4) * 4 && split_stride % 4 == 0 && split_stride >= m * ldc, "gemm_tn_f16: bad ldc / split_stride");
  TX_REQUIRE(splits >= 1 && splits <= 65535, "gemm_tn_f16: bad split count");
  cudaStream_t st = (cudaStream_t)stream;
  GemmEpilogue epi{nullptr, 0, 1, 0, 0, 0, 1.f, 1.f, scale_a, scale_b, nullptr};
  const int bn = pick_bn_tn16(n);
  if (use_pair() && m > kBM && bn == 256)
    return launch_gemm_pair<256, 3, true, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st, epi);
  if (bn == 256) return launch_gemm<256, 2, true, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st, epi);
  if (bn == 192) return launch_gemm<192, 2, true, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st, epi);
  if (bn == 128) return launch_gemm<128, 3, true, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st, epi);
  return launch_gemm<64, 4, true, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c_partial, ldc, split_stride, m, n, r, (int)splits, st, epi);
}

int tx_gemm_nt_f16x3(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo, int64_t ldb,
                     const float* scale_a, const float* scale_b, float* c, int64_t ldc, int64_t m, int64_t n, int64_t k,
                     const tx_gemm_epilogue* e, float* amax_out, void* stream) {
  TX_REQUIRE(m > 0 && n > 0 && k > 0 && m < INT32_MAX && n < INT32_MAX && k < INT32_MAX, "gemm_f16: bad shape %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
  TX_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= k && ldb >= k, "gemm_f16: operand row pitch must be a multiple of 8 halves (16 B) and >= K");
  TX_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo) && aligned16(c), "gemm_f16: 16-byte aligned pointers required");
  TX_REQUIRE(ldc % 4 == 0 && ldc >= ((n + 3) / 4) * 4, "gemm_f16: ldc must be a multiple of 4 and hold round4(N) columns");
  cudaStream_t st = (cudaStream_t)stream;
  GemmEpilogue epi{nullptr, 0, 1, 0, 0, 0, 1.f, 1.f, scale_a, scale_b, amax_out};
  if (e && e->act_mask) {
    TX_REQUIRE(e->heads > 0 && e->dim > 0 && e->dim % 4 == 0 && e->mask_stride % 16 == 0 && e->mask_stride >= e->heads * e->dim / 4 &&
               aligned16(e->act_mask), "gemm_f16 epilogue: bad mask geometry");
    TX_REQUIRE(e->col0 == 0, "gemm_f16 epilogue: a column offset is not supported with a fused mask");
    TX_REQUIRE(e->p_drop >= 0.f && e->p_drop < 1.f, "gemm_f16 epilogue: p_drop must be in [0,1)");
    epi.mask = e->act_mask; epi.heads = (int)e->heads; epi.dim = (int)e->dim; epi.stride = (int)e->mask_stride;
    epi.feat_cols = (int)(e->heads * e->dim); epi.has_keep = e->has_keep_plane;
    epi.on = 1.f / (1.f - e->p_drop); epi.neg = e->act_slope * epi.on;
  }
  if (amax_out && !g_preclear && cudaMemsetAsync(amax_out, 0, sizeof(float), st) != cudaSuccess) { set_error("gemm_f16: memset failed"); return TX_ERR_CUDA; }
  if (use_pair() && m > kBM && n > 128 && pick_bn(n) == 256) {
    // experiment knob: 128-wide tiles leave room for 4 accumulator buffers in TMEM, so the MMA stream can run a whole short-K tile
    // ahead of the epilogue.  Measured on the two wide-output GEMMs (K = 300 / 500): no gain (0.226 vs 0.229 ms, 0.290 vs 0.262 ms),
    // so it is off unless TAXO_GEMM_NARROW_K=<K limit> asks for it
    static int narrow_k = -1;
    if (narrow_k < 0) { const char* e_n = getenv("TAXO_GEMM_NARROW_K"); narrow_k = e_n ? atoi(e_n) : 0; }
    if (k <= narrow_k) return launch_gemm_pair<128, 4, false, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
    // short reductions with wide outputs (fwd L0: K = 300, dz L1: K = 500) are bound by the stores of C: 2-stage ring + store slabs
    static int slab_k = -1;
    if (slab_k < 0) { const char* e_s = getenv("TAXO_GEMM_SLAB_K"); slab_k = e_s ? atoi(e_s) : 640; }
    if (k <= slab_k) return launch_gemm_pair<256, 2, false, true, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
    return launch_gemm_pair<256, 3, false, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  }
  if (use_cluster() && m > kBM && n > 128) {
    if (pick_bn(n) == 160) return launch_gemm<160, 3, false, 32, 2, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
    return launch_gemm<256, 2, false, 32, 2, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  }
  if (n > 128 && pick_bn(n) == 160) return launch_gemm<160, 3, false, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  if (n > 128) return launch_gemm<256, 2, false, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  if (n > 64) return launch_gemm<128, 3, false, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
  return launch_gemm<64, 4, false, 32, 1, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, c, ldc, 0, m, n, k, 1, st, epi);
}

}  // extern "C"
