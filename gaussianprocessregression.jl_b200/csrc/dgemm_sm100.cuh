// FP64 tile GEMM for sm_100a on the DMMA pipe (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// This is the work-horse of the factor / solve / inverse path that replaces the
// reference's LAPACK calls (dpotrf / dpotrs / dtrsm / dsyrk / dgemm; see
// SURVEY.md 2.3 B1-B3, B8, B9, B11 -> /root/reference/src/cost.jl:77-111,
// /root/reference/src/predict.jl:29-95, /root/reference/src/split_predict.jl:10-19).
//
//   C(MxN) = alpha * op(A)(MxK) * op(B)(KxN) + beta * C          (column major)
//
// All dimensions are multiples of the tile (M,N % 128 == 0, K % 16 == 0): the
// library pads every matrix to 128 (identity on the padded diagonal), so the
// kernel has no edge handling at all.
//
// Operand storage forms (what "contiguous" means in global memory):
//   A_KC = true  : op(A)[m,k] = A[k + m*lda]   (transA = 'T', k contiguous)
//   A_KC = false : op(A)[m,k] = A[m + k*lda]   (transA = 'N', m contiguous)
//   B_KC = true  : op(B)[k,n] = B[k + n*ldb]   (transB = 'N', k contiguous)
//   B_KC = false : op(B)[k,n] = B[n + k*ldb]   (transB = 'T', n contiguous)
//
// Three tile configurations (template MI = 8x8 accumulator tiles per warp along m, WN = warps along n):
//   MI = 8, WN = 4 : CTA tile 128x128x16,  8 warps (2 x 4) of 64x32, 4-stage ring, 1 CTA / SM
//   MI = 8, WN = 2 : CTA tile 128x64x16,   4 warps (2 x 2) of 64x32, 3-stage ring, 2 CTAs / SM
//   MI = 4, WN = 4 : CTA tile 128x128x16, 16 warps (4 x 4) of 32x32, 4-stage ring, 1 CTA / SM (4 warps per
//                    scheduler hide barrier / LDS / cp.async waits of each other)
// Operands are staged global->shared with cp.async (16 B, LDGSTS) and read back as conflict-free LDS.128
// fragments:
//   k-contiguous tile : smem[row][24]      (row stride 24 doubles)
//   m-contiguous tile : smem[k][rows + 2]  (row stride 130 / 66 doubles)
// Inside a k-block of 8 the lane with threadID_in_group t owns k = 2t (first
// DMMA) and k = 2t+1 (second DMMA) so one LDS.128 feeds two DMMAs; A and B use
// the same assignment so the contraction is consistent.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;    // granularity required of N by the callers (both tile configurations divide it)
constexpr int GEMM_BK = 16;
constexpr int GEMM_LDK = 24;    // k-contiguous smem row stride (doubles)

template <int MI, int WN> struct GemmCfg {
  static constexpr int WARPS_M = 16 / MI;             // 128 rows / (8*MI rows per warp)
  static constexpr int BN = 32 * WN;
  static constexpr int THREADS = 32 * WARPS_M * WN;
  static constexpr int STAGES = (WN == 4) ? 4 : 3;
  static constexpr int MIN_CTAS = (WN == 4) ? 1 : 2;
};

enum GemmFlags : int {
  GEMM_UPPER_ONLY = 1,   // C is a diagonal-anchored symmetric block: only tiles/elements with row <= col are computed/stored
  GEMM_K_FROM_N = 2,     // tile column block J contracts over k >= 128*J only (triangular product W W^T, W upper)
  GEMM_C_ALIASES_A = 4,  // C overwrites the A operand in place (needs the full-width 128x128 tile)
  // Tile-mapped forms of the block-cyclic multi-GPU drivers (csrc/dist_blocked.hpp): the local 128-tile column
  // jt of C stands for the GLOBAL tile column gt = col_gtile[jt].
  GEMM_MAP_UPPER = 8,    // only elements with (row_gtile0 + tile row, row in tile) <= (gt, col in tile) exist
  GEMM_MAP_KUPTO = 16,   // the tile column contracts over k < (gt - k_gtile0 + 1) * 128 only
  GEMM_SKIP_TILE00 = 64, // the 128x128 tile at (0, 0) of C is neither read nor written (it is produced elsewhere, concurrently)
  GEMM_MAP_BROWS = 32,   // the op(B) columns of the tile column start at gt * 128 instead of at the local column
};

struct GemmParams {
  int M, N, K;
  double alpha, beta;
  const double* A; long long lda;
  const double* B; long long ldb;
  double* C; long long ldc;
  int flags;
  long long sA, sB, sC;   // strided batch (blockIdx.z)
  const int* col_gtile;   // GEMM_MAP_*: global 128-tile column per local 128-tile column (device memory)
  int row_gtile0, k_gtile0;
  const double* E; long long lde;   // EPI instantiation: C = alpha * (op(A) op(B)) .* E + beta * C  (E laid out like C)
};

// smem layout of one operand tile with ROWS rows (m or n extent) and 16 k
template <bool KC, int ROWS> struct GemmTile {
  static constexpr int LD = KC ? GEMM_LDK : (ROWS + 2);
  static constexpr int ELEMS = KC ? (ROWS * GEMM_LDK) : (GEMM_BK * (ROWS + 2));
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Stage one ROWS x 16 operand tile (kt-th k-tile) into shared memory.
//   KC : element (r,k) at P[k + r*ld]   -> smem[r*LDK + k]
//   !KC: element (r,k) at P[r + k*ld]   -> smem[k*(ROWS+2) + r]
template <bool KC, int ROWS, int THREADS>
__device__ __forceinline__ void gemm_load_tile(double* smem, const double* __restrict__ P, long long ld, int kt, int tid) {
  constexpr int PER_THREAD = ROWS * 8 / THREADS;
#pragma unroll
  for (int q = 0; q < PER_THREAD; ++q) {
    const int c = tid + THREADS * q;
    if (KC) {
      const int r = c >> 3, kc = c & 7;
      cp_async16(smem + r * GEMM_LDK + 2 * kc, P + (long long)kt * GEMM_BK + 2 * kc + (long long)r * ld);
    } else {
      const int k = c / (ROWS / 2), rc = c % (ROWS / 2);
      cp_async16(smem + k * (ROWS + 2) + 2 * rc, P + 2 * rc + ((long long)kt * GEMM_BK + k) * ld);
    }
  }
}

// Tile-local row of accumulator sub-tile i (0..7) for lane group g (A operand),
// and tile-local column of accumulator sub-tile j (0..3) for in-tile column q
// (B operand).  The maps differ per storage form so that fragment loads are
// LDS.128 and bank-conflict free.
template <bool KC, int MI> __device__ __forceinline__ int gemm_row_of(int wm, int i, int g) {
  return KC ? (8 * MI * wm + 8 * i + g) : (8 * MI * wm + 16 * (i >> 1) + 2 * g + (i & 1));
}
template <bool KC> __device__ __forceinline__ int gemm_col_of(int wn, int j, int q) {
  return KC ? (32 * wn + 8 * j + q) : (32 * wn + 16 * (j >> 1) + 2 * q + (j & 1));
}

// EPI: Hadamard epilogue C = alpha * acc .* E + beta * C (T,N form only) -- the split-predict mean
// mu += A_k .* (B_k (Diagonal(wt) C_k)) of /root/reference/src/split_predict.jl:14-16 in one pass, no BCw round trip.
template <bool A_KC, bool B_KC, int MI, int WN, bool EPI = false>
__global__ void __launch_bounds__(GemmCfg<MI, WN>::THREADS, GemmCfg<MI, WN>::MIN_CTAS) dgemm128_kernel(const GemmParams p) {
  static_assert(!EPI || (A_KC && B_KC), "the Hadamard epilogue exists for the T,N form only");
  extern __shared__ __align__(16) double gemm_smem[];
  using Cfg = GemmCfg<MI, WN>;
  constexpr int BN = Cfg::BN;
  constexpr int THREADS = Cfg::THREADS;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int WROWS = 8 * MI;      // rows per warp
  constexpr int SA = GemmTile<A_KC, GEMM_BM>::ELEMS;
  constexpr int SB = GemmTile<B_KC, BN>::ELEMS;
  constexpr int LDA_S = GemmTile<A_KC, GEMM_BM>::LD;
  constexpr int LDB_S = GemmTile<B_KC, BN>::LD;

  // CTA rasterisation.  The hardware issues CTAs in x-fastest order; taken literally (tile_m = blockIdx.x) one wave
  // of ~300 CTAs would span every row tile of a single column tile, i.e. stream the WHOLE A operand from DRAM once per
  // column tile (ncu, N = 32768: 1.3 TB of DRAM reads per evaluation against 26 GB algorithmic).  Instead consecutive
  // CTAs walk GM row tiles x all column tiles, so a wave covers a near-square patch of C and its operand slabs stay in L2.
  constexpr int GM = (WN == 2) ? 16 : 12;
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
  const int blk_n = (tile_n * BN) >> 7;   // 128-block column of this tile (tile_m is already a 128-block row)
  if ((p.flags & GEMM_UPPER_ONLY) && tile_m > blk_n) return;
  if ((p.flags & GEMM_SKIP_TILE00) && tile_m == 0 && blk_n == 0) return;
  int gt = 0;   // global tile column (mapped forms)
  if (p.flags & (GEMM_MAP_UPPER | GEMM_MAP_KUPTO | GEMM_MAP_BROWS)) {
    gt = p.col_gtile[blk_n];
    if ((p.flags & GEMM_MAP_UPPER) && p.row_gtile0 + tile_m > gt) return;
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;

  const long long m0 = (long long)tile_m * GEMM_BM, n0 = (long long)tile_n * BN;
  const long long bz = blockIdx.z;
  const double* Ap = p.A + bz * p.sA + (A_KC ? m0 * p.lda : m0);
  const long long nB = (p.flags & GEMM_MAP_BROWS) ? (long long)gt * 128 + (n0 & 127) : n0;   // first op(B) column of this tile
  const double* Bp = p.B + bz * p.sB + (B_KC ? nB * p.ldb : nB);
  const int kt0 = (p.flags & GEMM_K_FROM_N) ? blk_n * (128 / GEMM_BK) : 0;   // first k-tile of this tile

  double acc[MI][4][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  int KT = p.K / GEMM_BK - kt0;
  if (p.flags & GEMM_MAP_KUPTO) KT = min(KT, (gt - p.k_gtile0 + 1) * (128 / GEMM_BK));

  // Pull the C tile towards L2 while the main loop runs (the epilogue reads it when beta != 0):
  // BN columns x 1 KiB = 8 lines of 128 B per column.
  if (p.beta != 0.0) {
    const double* cbase = p.C + bz * p.sC + m0 + n0 * p.ldc;
#pragma unroll
    for (int q = 0; q < BN * 8 / THREADS; ++q) {
      const int L = tid + THREADS * q;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(cbase + (long long)(L >> 3) * p.ldc + (L & 7) * 16));
    }
  }

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      gemm_load_tile<A_KC, GEMM_BM, THREADS>(gemm_smem + s * (SA + SB), Ap, p.lda, kt0 + s, tid);
      gemm_load_tile<B_KC, BN, THREADS>(gemm_smem + s * (SA + SB) + SA, Bp, p.ldb, kt0 + s, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) {
        int s = nk % STAGES;
        gemm_load_tile<A_KC, GEMM_BM, THREADS>(gemm_smem + s * (SA + SB), Ap, p.lda, kt0 + nk, tid);
        gemm_load_tile<B_KC, BN, THREADS>(gemm_smem + s * (SA + SB) + SA, Bp, p.ldb, kt0 + nk, tid);
      }
      cp_async_commit();
    }
    const double* As = gemm_smem + (kt % STAGES) * (SA + SB);
    const double* Bs = As + SA;

#pragma unroll
    for (int sp = 0; sp < 2; ++sp) {   // two k-blocks of 8 per 16-wide tile
      double a[MI][2], b[4][2];
      if (A_KC) {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const double2 v = *reinterpret_cast<const double2*>(As + (WROWS * wm + 8 * i + g) * LDA_S + 8 * sp + 2 * t);
          a[i][0] = v.x; a[i][1] = v.y;
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int pi = 0; pi < MI / 2; ++pi) {
            const double2 v = *reinterpret_cast<const double2*>(As + (8 * sp + 2 * t + h) * LDA_S + WROWS * wm + 16 * pi + 2 * g);
            a[2 * pi][h] = v.x; a[2 * pi + 1][h] = v.y;
          }
      }
      if (B_KC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 v = *reinterpret_cast<const double2*>(Bs + (32 * wn + 8 * j + g) * LDB_S + 8 * sp + 2 * t);
          b[j][0] = v.x; b[j][1] = v.y;
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int pj = 0; pj < 2; ++pj) {
            const double2 v = *reinterpret_cast<const double2*>(Bs + (8 * sp + 2 * t + h) * LDB_S + 32 * wn + 16 * pj + 2 * g);
            b[2 * pj][h] = v.x; b[2 * pj + 1][h] = v.y;
          }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i][h], b[j][h]);
    }
  }
  cp_async_wait<0>();

  // epilogue: C = alpha*acc + beta*C (element (row, col) of the tile lives in
  // acc[i][j][v] with row = row_of(i,g), col = col_of(j, 2t+v)).  The old C
  // values are fetched in batches of 16 independent loads BEFORE any store of
  // the batch is issued (a load-after-store to the same array cannot be
  // hoisted by the compiler, which would serialise 64 DRAM round trips).
  // On the diagonal 128-block of an UPPER_ONLY product only row <= col (inside the block) is touched.
  const bool diag_tile = ((p.flags & GEMM_UPPER_ONLY) && (tile_m == blk_n)) ||
                         ((p.flags & GEMM_MAP_UPPER) && (p.row_gtile0 + tile_m == gt));
  const int coff = (int)(n0 & 127);   // column offset of this tile inside its 128-block (0 or 64)
  const double alpha = p.alpha, beta = p.beta;
  double* Cp = p.C + bz * p.sC + m0 + n0 * p.ldc;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double old[MI][2];
    double ep[EPI ? MI : 1][2];
    if constexpr (EPI) {
      const double* Ep = p.E + bz * p.sC + m0 + n0 * p.lde;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const double* Ecol = Ep + (long long)gemm_col_of<B_KC>(wn, j, 2 * t + v) * p.lde;
#pragma unroll
        for (int i = 0; i < (EPI ? MI : 1); ++i) ep[i][v] = Ecol[gemm_row_of<true, MI>(wm, i, g)];
      }
    }
    if (beta != 0.0) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int col = gemm_col_of<B_KC>(wn, j, 2 * t + v);
        const double* Ccol = Cp + (long long)col * p.ldc;
        if (A_KC) {
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            const int row = gemm_row_of<true, MI>(wm, i, g);
            old[i][v] = (diag_tile && row > col + coff) ? 0.0 : Ccol[row];
          }
        } else {
#pragma unroll
          for (int pi = 0; pi < MI / 2; ++pi) {
            const int row = gemm_row_of<false, MI>(wm, 2 * pi, g);
            if (!diag_tile) {
              const double2 o = *reinterpret_cast<const double2*>(Ccol + row);
              old[2 * pi][v] = o.x; old[2 * pi + 1][v] = o.y;
            } else {
              old[2 * pi][v] = (row <= col + coff) ? Ccol[row] : 0.0;
              old[2 * pi + 1][v] = (row + 1 <= col + coff) ? Ccol[row + 1] : 0.0;
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < MI; ++i) old[i][0] = old[i][1] = 0.0;
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int col = gemm_col_of<B_KC>(wn, j, 2 * t + v);
      double* Ccol = Cp + (long long)col * p.ldc;
      if (A_KC) {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int row = gemm_row_of<true, MI>(wm, i, g);
          if (diag_tile && row > col + coff) continue;
          if constexpr (EPI) Ccol[row] = alpha * acc[i][j][v] * ep[i][v] + beta * old[i][v];
          else Ccol[row] = alpha * acc[i][j][v] + beta * old[i][v];
        }
      } else {
#pragma unroll
        for (int pi = 0; pi < MI / 2; ++pi) {
          const int row = gemm_row_of<false, MI>(wm, 2 * pi, g);   // even row; row+1 is sub-tile 2*pi+1
          const double r0 = alpha * acc[2 * pi][j][v] + beta * old[2 * pi][v];
          const double r1 = alpha * acc[2 * pi + 1][j][v] + beta * old[2 * pi + 1][v];
          if (!diag_tile) {
            *reinterpret_cast<double2*>(Ccol + row) = make_double2(r0, r1);
          } else {
            if (row <= col + coff) Ccol[row] = r0;
            if (row + 1 <= col + coff) Ccol[row + 1] = r1;
          }
        }
      }
    }
  }
}

template <bool A_KC, bool B_KC, int MI, int WN, bool EPI = false> constexpr size_t gemm_smem_bytes() {
  return (size_t)GemmCfg<MI, WN>::STAGES *
         (GemmTile<A_KC, GEMM_BM>::ELEMS + GemmTile<B_KC, GemmCfg<MI, WN>::BN>::ELEMS) * sizeof(double);
}

template <bool A_KC, bool B_KC, int MI, int WN> inline cudaError_t gemm_set_attr() {
  return cudaFuncSetAttribute(dgemm128_kernel<A_KC, B_KC, MI, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)gemm_smem_bytes<A_KC, B_KC, MI, WN>());
}
template <int MI, int WN> inline cudaError_t gemm_set_attr_cfg() {
  cudaError_t e = gemm_set_attr<true, true, MI, WN>();
  if (e == cudaSuccess) e = gemm_set_attr<false, true, MI, WN>();
  if (e == cudaSuccess) e = gemm_set_attr<false, false, MI, WN>();
  return e;
}
// once per context / device
inline cudaError_t gemm_setup_attributes() {
  cudaError_t e = gemm_set_attr_cfg<8, 4>();
  if (e == cudaSuccess) e = gemm_set_attr_cfg<8, 2>();
  if (e == cudaSuccess) e = gemm_set_attr_cfg<4, 4>();
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(dgemm128_kernel<true, true, 8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gemm_smem_bytes<true, true, 8, 2>());
  return e;
}

template <bool A_KC, bool B_KC, int MI, int WN>
inline cudaError_t gemm_launch_cfg(cudaStream_t st, const GemmParams& p, int batch) {
  dim3 grid(p.M / GEMM_BM, p.N / GemmCfg<MI, WN>::BN, batch), block(GemmCfg<MI, WN>::THREADS);
  dgemm128_kernel<A_KC, B_KC, MI, WN><<<grid, block, gemm_smem_bytes<A_KC, B_KC, MI, WN>(), st>>>(p);
  return cudaGetLastError();
}
template <int MI, int WN>
inline cudaError_t gemm_launch_form(cudaStream_t st, const GemmParams& p, int batch, bool aT, bool bT) {
  if (aT) return gemm_launch_cfg<true, true, MI, WN>(st, p, batch);
  if (!bT) return gemm_launch_cfg<false, true, MI, WN>(st, p, batch);
  return gemm_launch_cfg<false, false, MI, WN>(st, p, batch);
}

// Tile configuration `cfg`: 0 = automatic, 1 = 128x128 / 8 warps, 2 = 128x64 / 4 warps x 2 CTAs, 3 = 128x128 / 16 warps
// (per-context option "gemm_cfg": tuning and tests; passed in by the caller, there is no process-global state).
// transA/transB: 'N' or 'T' (BLAS meaning, column major).  Supported: TN, NN, NT.
inline cudaError_t launch_dgemm128(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha,
                                   const double* A, long long lda, const double* B, long long ldb, double beta,
                                   double* C, long long ldc, int flags, int batch = 1, long long sA = 0,
                                   long long sB = 0, long long sC = 0, const int* col_gtile = nullptr,
                                   int row_gtile0 = 0, int k_gtile0 = 0, int cfg = 0, const double* E = nullptr,
                                   long long lde = 0) {
  if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
  if ((M % GEMM_BM) || (N % GEMM_BN) || (K % GEMM_BK) || K <= 0 || batch > 65535) return cudaErrorInvalidValue;
  if ((flags & GEMM_K_FROM_N) && K < N) return cudaErrorInvalidValue;
  if ((flags & (GEMM_MAP_UPPER | GEMM_MAP_KUPTO | GEMM_MAP_BROWS)) && !col_gtile) return cudaErrorInvalidValue;
  GemmParams p{M, N, K, alpha, beta, A, lda, B, ldb, C, ldc, flags, sA, sB, sC, col_gtile, row_gtile0, k_gtile0, E, lde};
  const bool aT = (transA == 'T' || transA == 't'), bT = (transB == 'T' || transB == 't');
  if (aT && bT) return cudaErrorNotSupported;
  if (E) {   // Hadamard epilogue: T,N form, 128x64 tile
    if (!aT || bT || flags) return cudaErrorNotSupported;
    dim3 grid(M / GEMM_BM, N / GemmCfg<8, 2>::BN, batch), block(GemmCfg<8, 2>::THREADS);
    dgemm128_kernel<true, true, 8, 2, true><<<grid, block, gemm_smem_bytes<true, true, 8, 2>(), st>>>(p);
    return cudaGetLastError();
  }
  // C written over the A operand: a CTA must own every k column it reads -> full-width 128x128 tile.
  const bool alias_a = (flags & GEMM_C_ALIASES_A) || (const double*)C == A;
  // default: 128x64 tiles, two co-resident CTAs per SM (measured on B200 at 4096^3: TN 33.8 / NN 34.3 / NT 31.4
  // TFLOP/s vs 31.5 / 30.8 / 29.8 for the single-CTA 128x128 tile; cuBLAS DGEMM 35.4)
  if (cfg < 1 || cfg > 3) cfg = 2;
  if (alias_a && cfg == 2) cfg = 1;
  if (cfg == 1) return gemm_launch_form<8, 4>(st, p, batch, aT, bT);
  if (cfg == 2) return gemm_launch_form<8, 2>(st, p, batch, aT, bT);
  return gemm_launch_form<4, 4>(st, p, batch, aT, bT);
}

}  // namespace gpr
