// TMA-fed instantiation of the FP64 tile GEMM (T,N form: both operands contraction-contiguous -- the form every product
// of potrf / trtri_t / lauum_oop_t / the prediction solve takes, csrc/blocked.hpp).
//
// Same tile (128 x 64 x 16), same warp layout (4 warps of 64 x 32, 32 DMMA.8x8x4 accumulator tiles per warp), same
// arithmetic order as dgemm128_kernel<true, true, 8, 2> (dgemm_sm100.cuh), but the operand ring is filled by the TMA
// unit instead of 12 LDGSTS per thread and k-tile:
//   * one elected thread issues two cp.async.bulk.tensor.3d copies per k-tile (A: 16 x 128 doubles, B: 16 x 64) that
//     complete on a per-stage "full" mbarrier; consumers release a stage through an "empty" mbarrier (one arrive per
//     warp), so there is no __syncthreads in the main loop and warps drift by up to a stage;
//   * tiles land DENSE in shared memory (128-byte rows, SWIZZLE_128B: the 16-byte chunk c of row r sits at chunk
//     c ^ (r & 7)) -- 24 KB per stage instead of 36 KB with the padded LDGSTS layout, so the ring is 4 deep;
//   * fragment loads stay conflict-free LDS.128 because the eight row groups of an accumulator tile are PERMUTED:
//     lane group g reads tile row pi(g) = ((g & 1) << 2) | (g >> 1), which makes the two row groups of a quarter-warp
//     differ in bit 2 of (r & 7) and hence cover all eight 16-byte bank groups.  The permutation only relabels which
//     row / column of C an accumulator element is (epilogue index maps below).
// The 3-D tensor maps (k, row, batch member) are encoded on the host per launch (cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint; a few microseconds, hidden behind the asynchronous launch queue).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dgemm_sm100.cuh"
#include "kbuild_tma.cuh"

namespace gpr {

constexpr int TG_BM = 128, TG_BN = 64, TG_BK = 16, TG_STAGES = 4, TG_THREADS = 128;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 8, TG_B_BYTES = TG_BN * TG_BK * 8;
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr size_t TG_SMEM_BYTES = (size_t)TG_STAGES * TG_STAGE_BYTES + 1024 /* alignment slack */ + 2 * TG_STAGES * 8;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"((uint64_t)map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ int tg_pi(int g) { return ((g & 1) << 2) | (g >> 1); }

template <bool EPI>
__global__ void __launch_bounds__(TG_THREADS, 2)
dgemm128_tma_kernel(const GemmParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ unsigned char tg_smem_raw[];
  // SWIZZLE_128B destinations must be 1024-byte aligned
  unsigned char* tg_smem = reinterpret_cast<unsigned char*>(((uintptr_t)tg_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tg_smem + (size_t)TG_STAGES * TG_STAGE_BYTES);
  uint64_t* empty = full + TG_STAGES;

  // CTA rasterisation, identical to dgemm128_kernel (16 row tiles x all column tiles per pass)
  constexpr int GM = 16;
  int tile_m, tile_n;
  {
    const int Mx = gridDim.x, Ny = gridDim.y;
    const int pid = blockIdx.x + blockIdx.y * Mx;
    const int group = pid / (GM * Ny), first_m = group * GM;
    const int gsize = min(Mx - first_m, GM);
    const int rem = pid - group * GM * Ny;
    tile_m = first_m + rem % gsize;
    tile_n = rem / gsize;
  }
  // KUPTO: the contraction length grows with the column tile -> issue the longest tiles first (shorter tail)
  if (p.flags & GEMM_MAP_KUPTO) tile_n = gridDim.y - 1 - tile_n;
  const int blk_n = (tile_n * TG_BN) >> 7;
  if ((p.flags & GEMM_UPPER_ONLY) && tile_m > blk_n) return;
  if ((p.flags & GEMM_SKIP_TILE00) && tile_m == 0 && blk_n == 0) return;
  int gt = 0;   // global tile column (tile-mapped forms of the block-cyclic multi-GPU drivers, csrc/dist_blocked.hpp)
  if (p.flags & (GEMM_MAP_UPPER | GEMM_MAP_KUPTO | GEMM_MAP_BROWS)) {
    gt = p.col_gtile[blk_n];
    if ((p.flags & GEMM_MAP_UPPER) && p.row_gtile0 + tile_m > gt) return;
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  const int pg = tg_pi(g);
  const long long m0 = (long long)tile_m * TG_BM, n0 = (long long)tile_n * TG_BN;
  const int bz = blockIdx.z;
  const int kt0 = (p.flags & GEMM_K_FROM_N) ? blk_n * (128 / TG_BK) : 0;
  int KT = p.K / TG_BK - kt0;
  if (p.flags & GEMM_MAP_KUPTO) KT = min(KT, (gt - p.k_gtile0 + 1) * (128 / TG_BK));
  const int nB = (p.flags & GEMM_MAP_BROWS) ? gt * 128 + (int)(n0 & 127) : (int)n0;   // first op(B) column of this tile

  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TG_THREADS / 32); }
    mbar_fence_init();
  }
  __syncthreads();

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (p.beta != 0.0) {   // pull the C tile towards L2 while the main loop runs
    const double* cbase = p.C + bz * p.sC + m0 + n0 * p.ldc;
#pragma unroll
    for (int q = 0; q < TG_BN * 8 / TG_THREADS; ++q) {
      const int L = tid + TG_THREADS * q;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(cbase + (long long)(L >> 3) * p.ldc + (L & 7) * 16));
    }
  }

  // prologue: fill the ring
  if (tid == 0) {
    for (int s = 0; s < TG_STAGES - 1 && s < KT; ++s) {
      unsigned char* st = tg_smem + (size_t)s * TG_STAGE_BYTES;
      mbar_expect_tx(&full[s], TG_STAGE_BYTES);
      tma_load_3d(st, &mapA, &full[s], (kt0 + s) * TG_BK, (int)m0, bz);
      tma_load_3d(st + TG_A_BYTES, &mapB, &full[s], (kt0 + s) * TG_BK, nB, bz);
    }
  }

  // per-lane constants of the swizzled fragment addresses: row = base + 8 i + pi(g), 16-byte chunk (4 sp + t) ^ pi(g)
  const unsigned a_lane = (unsigned)((64 * wm + pg) * 128), b_lane = (unsigned)((32 * wn + pg) * 128);
  const unsigned ch0 = (unsigned)(((0 + t) ^ pg) << 4), ch1 = (unsigned)(((4 + t) ^ pg) << 4);

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % TG_STAGES;
    if (warp == 0) {
      if (lane == 0) {
        const int nk = kt + TG_STAGES - 1;
        if (nk < KT) {
          const int ns = nk % TG_STAGES;
          if (nk >= TG_STAGES) mbar_wait(&empty[ns], (unsigned)((nk / TG_STAGES - 1) & 1));   // all warps are done with k-tile nk - STAGES
          unsigned char* st = tg_smem + (size_t)ns * TG_STAGE_BYTES;
          mbar_expect_tx(&full[ns], TG_STAGE_BYTES);
          tma_load_3d(st, &mapA, &full[ns], (kt0 + nk) * TG_BK, (int)m0, bz);
          tma_load_3d(st + TG_A_BYTES, &mapB, &full[ns], (kt0 + nk) * TG_BK, nB, bz);
        }
      }
      __syncwarp();
    }
    mbar_wait(&full[s], (unsigned)((kt / TG_STAGES) & 1));
    const unsigned char* As = tg_smem + (size_t)s * TG_STAGE_BYTES;
    const unsigned char* Bs = As + TG_A_BYTES;
#pragma unroll
    for (int sp = 0; sp < 2; ++sp) {
      const unsigned ch = sp ? ch1 : ch0;
      double a[8][2], b[4][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(As + a_lane + i * 1024 + ch);
        a[i][0] = v.x; a[i][1] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(Bs + b_lane + j * 1024 + ch);
        b[j][0] = v.x; b[j][1] = v.y;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i][h], b[j][h]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // epilogue: element (row, col) of the tile lives in acc[i][j][v] with row = 64 wm + 8 i + pi(g), col = 32 wn + 8 j + pi(2 t + v)
  const bool diag_tile = ((p.flags & GEMM_UPPER_ONLY) && (tile_m == blk_n)) ||
                         ((p.flags & GEMM_MAP_UPPER) && (p.row_gtile0 + tile_m == gt));
  const int coff = (int)(n0 & 127);
  const double alpha = p.alpha, beta = p.beta;
  double* Cp = p.C + bz * p.sC + m0 + n0 * p.ldc;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double old[8][2];
    double ep[EPI ? 8 : 1][2];
    if constexpr (EPI) {     // Hadamard epilogue C = alpha * acc .* E + beta * C (split-predict mean, split_predict.jl:14-16)
      const double* Ep = p.E + bz * p.sC + m0 + n0 * p.lde;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const double* Ecol = Ep + (long long)(32 * wn + 8 * j + tg_pi(2 * t + v)) * p.lde;
#pragma unroll
        for (int i = 0; i < 8; ++i) ep[i][v] = Ecol[64 * wm + 8 * i + pg];
      }
    }
    if (beta != 0.0) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int col = 32 * wn + 8 * j + tg_pi(2 * t + v);
        const double* Ccol = Cp + (long long)col * p.ldc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 64 * wm + 8 * i + pg;
          old[i][v] = (diag_tile && row > col + coff) ? 0.0 : Ccol[row];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) old[i][0] = old[i][1] = 0.0;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int col = 32 * wn + 8 * j + tg_pi(2 * t + v);
      double* Ccol = Cp + (long long)col * p.ldc;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 64 * wm + 8 * i + pg;
        if (diag_tile && row > col + coff) continue;
        if constexpr (EPI) Ccol[row] = alpha * acc[i][j][v] * ep[i][v] + beta * old[i][v];
        else Ccol[row] = alpha * acc[i][j][v] + beta * old[i][v];
      }
    }
  }
}

// 3-D FP64 tensor map {k, rows, batch} of a k-contiguous operand: element (k, r, z) at base[k + r * ld + z * sz]
inline bool tma_make_map3_f64(CUtensorMap* map, const double* base, uint64_t kdim, uint64_t rows, uint64_t ld, uint64_t batch, uint64_t sz,
                              uint32_t box_rows) {
  PFN_encodeTiled fn = tma_encode_fn();
  if (!fn) return false;
  if (batch <= 1) { batch = 1; sz = ld * 2; }   // a single member: the stride is never applied, it only has to be well formed
  if (((uintptr_t)base & 15) || ((ld * 8) & 15) || ((sz * 8) & 15)) return false;
  cuuint64_t dims[3] = {kdim, rows, batch};
  cuuint64_t strides[2] = {ld * 8, sz * 8};
  cuuint32_t box[3] = {TG_BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline cudaError_t gemm_tma_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(dgemm128_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dgemm128_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_BYTES);
  return e;
}

// true if this product can run on the TMA-fed kernel
inline bool gemm_tma_supported(char transA, char transB, int M, int N, int K, const double* A, const double* B, const double* C, int flags,
                               int batch, long long sA, long long sB) {
  if (!(transA == 'T' || transA == 't') || (transB == 'T' || transB == 't')) return false;
  if (flags & ~(GEMM_UPPER_ONLY | GEMM_K_FROM_N | GEMM_SKIP_TILE00 | GEMM_MAP_UPPER | GEMM_MAP_KUPTO | GEMM_MAP_BROWS)) return false;
  if ((const double*)C == A || (const double*)C == B) return false;            // in-place forms need the full-width tile
  if (batch > 1 && (sA <= 0 || sB <= 0)) return false;
  if (M < 128 || N < 64 || K < 16) return false;
  return tma_encode_fn() != nullptr;
}

inline cudaError_t launch_dgemm128_tma(cudaStream_t st, int M, int N, int K, double alpha, const double* A, long long lda, const double* B,
                                       long long ldb, double beta, double* C, long long ldc, int flags, int batch, long long sA,
                                       long long sB, long long sC, const int* col_gtile = nullptr, int row_gtile0 = 0, int k_gtile0 = 0,
                                       const double* E = nullptr, long long lde = 0) {
  if ((flags & (GEMM_MAP_UPPER | GEMM_MAP_KUPTO | GEMM_MAP_BROWS)) && !col_gtile) return cudaErrorInvalidValue;
  CUtensorMap mA, mB;
  // BROWS: the op(B) columns are addressed by GLOBAL tile column, beyond the N local columns of C -> open row extent
  const uint64_t rowsB = (flags & GEMM_MAP_BROWS) ? ((uint64_t)1 << 31) : (uint64_t)N;
  if (!tma_make_map3_f64(&mA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, (uint64_t)batch, (uint64_t)sA, TG_BM) ||
      !tma_make_map3_f64(&mB, B, (uint64_t)K, rowsB, (uint64_t)ldb, (uint64_t)batch, (uint64_t)sB, TG_BN))
    return cudaErrorInvalidValue;
  GemmParams p{M, N, K, alpha, beta, A, lda, B, ldb, C, ldc, flags, sA, sB, sC, col_gtile, row_gtile0, k_gtile0, E, lde};
  dim3 grid(M / TG_BM, N / TG_BN, batch), block(TG_THREADS);
  if (E) dgemm128_tma_kernel<true><<<grid, block, TG_SMEM_BYTES, st>>>(p, mA, mB);
  else dgemm128_tma_kernel<false><<<grid, block, TG_SMEM_BYTES, st>>>(p, mA, mB);
  return cudaGetLastError();
}

}  // namespace gpr
