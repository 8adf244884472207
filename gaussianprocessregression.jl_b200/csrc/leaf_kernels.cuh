// 128x128 diagonal-block kernels of the blocked factorization (one CTA each),
// plus small utility kernels on the factor.
//
// potrf_leaf: the unblocked dpotrf('U') of the reference
// (/root/reference/src/cost.jl:77,87,104; /root/reference/src/predict.jl:31)
// restricted to one 128x128 diagonal block held in shared memory, followed by
// the explicit inverse of that block (used as the "trsm by GEMM" operand of
// every blocked solve; csrc/blocked.hpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpr {

constexpr int LEAF_N = 128;
constexpr int LEAF_LDS = 129;                                  // padded smem leading dimension of the factor
constexpr int LEAF_THREADS = 256;
constexpr int LEAF_NB = 32;                                    // inner block size
constexpr int LEAF_WLD = 33;                                   // padded leading dimension of a 32x32 inverse block
constexpr int LEAF_WBLK = LEAF_NB * LEAF_WLD;                  // doubles per inverse block
constexpr int LEAF_NWBLK = 10;                                 // upper 4x4 block triangle
constexpr size_t LEAF_SMEM_BYTES =
    (size_t)(LEAF_N * LEAF_LDS + LEAF_NWBLK * LEAF_WBLK + 2 * LEAF_NB + LEAF_N + 8) * sizeof(double);

__device__ __forceinline__ int leaf_wblk(int I, int J) { return (J * (J + 1) / 2 + I) * LEAF_WBLK; }   // I <= J

// A: 128x128 block (column major, lda) on the diagonal of the global matrix at
// global row/col offset goff.  On exit the upper triangle of the block holds U
// (U^T U = A); the strict lower triangle is left untouched (it keeps K, which
// the reference's dpotrf('U') also leaves in place: test/test_loss.jl:46).
// dinv: 128x128 (ld 128), receives inv(U) with explicit zeros below the diagonal.
// info: first failing pivot (1-based global index), LAPACK style; 0 = ok.
//
// Blocked with NB = 32 inside one CTA:
//   per block step J: (1) warp 0 factors the 32x32 diagonal block with one column per lane in registers
//   (pivot broadcast by shuffle, scaled pivot row exchanged through a 2 x 32 double buffer, one
//   __syncwarp per pivot), (2) forward substitution of the 32-row panel, one column per thread,
//   (3) rank-32 update of the trailing upper triangle.  The inverse is formed block-wise: the four
//   diagonal 32x32 inverses by four warps (branch-free back substitution in registers), then the six
//   off-diagonal blocks in three rounds of small block products.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, long long lda, double* __restrict__ dinv, long long* info, long long goff) {
  extern __shared__ __align__(16) double leaf_smem[];
  double* S = leaf_smem;                                   // S(r,c) = S[c*LEAF_LDS + r]
  double* W = S + LEAF_N * LEAF_LDS;                        // 10 inverse blocks, block (I,J) at leaf_wblk(I,J), (r,c) at [c*33 + r]
  double* rowbuf = W + LEAF_NWBLK * LEAF_WBLK;              // 2 x 32
  double* rinvs = rowbuf + 2 * LEAF_NB;                     // 128: 1 / U_kk
  int* flag = reinterpret_cast<int*>(rinvs + LEAF_N);       // failure flag
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) *flag = 0;
  // load the block: 64 elements per thread, 8 independent loads in flight
#pragma unroll 1
  for (int q0 = 0; q0 < 64; q0 += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + LEAF_THREADS * (q0 + q);
      v[q] = A[(idx & 127) + (long long)(idx >> 7) * lda];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + LEAF_THREADS * (q0 + q);
      S[(idx >> 7) * LEAF_LDS + (idx & 127)] = v[q];
    }
  }
  __syncthreads();

#pragma unroll 1
  for (int J = 0; J < 4; ++J) {
    const int o = J * LEAF_NB;
    // ---- (1) diagonal block, warp 0: lane j owns column j (rows 0..j used)
    if (warp == 0) {
      double a[LEAF_NB];
#pragma unroll
      for (int i = 0; i < LEAF_NB; ++i) a[i] = S[(o + lane) * LEAF_LDS + o + i];
      bool bad = false;
#pragma unroll
      for (int k = 0; k < LEAF_NB; ++k) {
        const double piv = __shfl_sync(0xffffffffu, a[k], k);
        if (!(piv > 0.0)) {          // uniform across the warp; also catches NaN
          if (lane == 0) { if (*info == 0) *info = goff + o + k + 1; *flag = 1; }
          bad = true;
          break;
        }
        const double rinv = rsqrt(piv);
        if (lane == k) { a[k] = piv * rinv; rinvs[o + k] = rinv; }
        else if (lane > k) a[k] *= rinv;                     // u_kj, row k of U in column j
        double* buf = rowbuf + (k & 1) * LEAF_NB;
        if (lane > k) buf[lane] = a[k];
        __syncwarp();
        if (lane > k) {
          const double ukj = a[k];
#pragma unroll
          for (int i = k + 1; i < LEAF_NB; ++i)
            if (i <= lane) a[i] -= buf[i] * ukj;             // a_ij -= u_ki u_kj
        }
      }
      if (!bad) {
#pragma unroll
        for (int i = 0; i < LEAF_NB; ++i)
          if (i <= lane) S[(o + lane) * LEAF_LDS + o + i] = a[i];
      }
    }
    __syncthreads();
    if (*flag) break;
    // ---- (2) panel: U(o.., c) = U_JJ^-T A(o.., c), one column per thread
    const int rem = LEAF_N - o - LEAF_NB;
    if (tid < rem) {
      const int c = o + LEAF_NB + tid;
      double b[LEAF_NB];
#pragma unroll
      for (int r = 0; r < LEAF_NB; ++r) b[r] = S[c * LEAF_LDS + o + r];
#pragma unroll
      for (int k = 0; k < LEAF_NB; ++k) {
        const double xk = b[k] * rinvs[o + k];
        b[k] = xk;
#pragma unroll
        for (int r = k + 1; r < LEAF_NB; ++r) b[r] -= S[(o + r) * LEAF_LDS + o + k] * xk;   // U(k,r)
      }
#pragma unroll
      for (int r = 0; r < LEAF_NB; ++r) S[c * LEAF_LDS + o + r] = b[r];
    }
    __syncthreads();
    // ---- (3) trailing update of the upper triangle: S(i,c) -= sum_k U(o+k,i) U(o+k,c)
    {
      const int c = tid & 127, part = tid >> 7;
      if (c >= o + LEAF_NB) {
        double u[LEAF_NB];
#pragma unroll
        for (int k = 0; k < LEAF_NB; ++k) u[k] = S[c * LEAF_LDS + o + k];
        for (int i = o + LEAF_NB + part; i <= c; i += 2) {
          const double* ui = S + i * LEAF_LDS + o;
          double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
          for (int k = 0; k < LEAF_NB; k += 4) {
            d0 += ui[k] * u[k]; d1 += ui[k + 1] * u[k + 1]; d2 += ui[k + 2] * u[k + 2]; d3 += ui[k + 3] * u[k + 3];
          }
          S[c * LEAF_LDS + i] -= (d0 + d1) + (d2 + d3);
        }
      }
    }
    __syncthreads();
  }
  const bool failed = (*flag != 0);

  // write U (upper part only)
  if (!failed) {
    for (int idx = tid; idx < LEAF_N * LEAF_N; idx += LEAF_THREADS) {
      const int r = idx & (LEAF_N - 1), c = idx >> 7;
      if (r <= c) A[r + (long long)c * lda] = S[c * LEAF_LDS + r];
    }
  }

  if (!failed) {
    // ---- inverse, diagonal blocks: warp J, lane j -> column j of inv(U_JJ); rows k > j come out as exact zeros
    if (warp < 4) {
      const int o = warp * LEAF_NB;
      double sacc[LEAF_NB];
#pragma unroll
      for (int i = 0; i < LEAF_NB; ++i) sacc[i] = (i == lane) ? 1.0 : 0.0;
      double* wd = W + leaf_wblk(warp, warp) + lane * LEAF_WLD;
#pragma unroll
      for (int k = LEAF_NB - 1; k >= 0; --k) {
        const double wk = sacc[k] * rinvs[o + k];
        wd[k] = wk;
#pragma unroll
        for (int i = 0; i < k; ++i) sacc[i] -= S[(o + k) * LEAF_LDS + o + i] * wk;   // U(i,k)
      }
    }
    __syncthreads();
    // ---- off-diagonal blocks, three rounds: W_IJ = W_II * ( - sum_{K=I+1..J} U_IK W_KJ )
    const int r = lane;              // row inside the block
    const int cbase = warp;          // columns cbase + 8*m
#pragma unroll 1
    for (int round = 1; round <= 3; ++round) {
      const int npairs = 4 - round;  // (I, I+round), I = 0..npairs-1
      double acc[3][4];
      // phase A: T = - sum_K U_IK W_KJ  -> written to the (I,J) block as a temporary
#pragma unroll
      for (int pidx = 0; pidx < 3; ++pidx) {
        if (pidx < npairs) {
          const int I = pidx, Jb = pidx + round;
          double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
          for (int Kb = I + 1; Kb <= Jb; ++Kb) {
            const double* Urow = S + (Kb * LEAF_NB) * LEAF_LDS + I * LEAF_NB + r;      // U_IK(r, kk) = Urow[kk*LDS]
            const double* Wk = W + leaf_wblk(Kb, Jb);                                   // W_KJ(kk, c) = Wk[c*33 + kk]
#pragma unroll 8
            for (int kk = 0; kk < LEAF_NB; ++kk) {
              const double uv = Urow[kk * LEAF_LDS];
              t0 -= uv * Wk[(cbase) * LEAF_WLD + kk];
              t1 -= uv * Wk[(cbase + 8) * LEAF_WLD + kk];
              t2 -= uv * Wk[(cbase + 16) * LEAF_WLD + kk];
              t3 -= uv * Wk[(cbase + 24) * LEAF_WLD + kk];
            }
          }
          acc[pidx][0] = t0; acc[pidx][1] = t1; acc[pidx][2] = t2; acc[pidx][3] = t3;
        }
      }
#pragma unroll
      for (int pidx = 0; pidx < 3; ++pidx) {
        if (pidx < npairs) {
          double* Tb = W + leaf_wblk(pidx, pidx + round);
#pragma unroll
          for (int m = 0; m < 4; ++m) Tb[(cbase + 8 * m) * LEAF_WLD + r] = acc[pidx][m];
        }
      }
      __syncthreads();
      // phase B: W_IJ = W_II * T
#pragma unroll
      for (int pidx = 0; pidx < 3; ++pidx) {
        if (pidx < npairs) {
          const int I = pidx, Jb = pidx + round;
          const double* Wii = W + leaf_wblk(I, I);          // W_II(r, kk) = Wii[kk*33 + r]
          const double* Tb = W + leaf_wblk(I, Jb);          // T(kk, c)    = Tb[c*33 + kk]
          double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll 8
          for (int kk = 0; kk < LEAF_NB; ++kk) {
            const double wv = Wii[kk * LEAF_WLD + r];
            t0 += wv * Tb[(cbase) * LEAF_WLD + kk];
            t1 += wv * Tb[(cbase + 8) * LEAF_WLD + kk];
            t2 += wv * Tb[(cbase + 16) * LEAF_WLD + kk];
            t3 += wv * Tb[(cbase + 24) * LEAF_WLD + kk];
          }
          acc[pidx][0] = t0; acc[pidx][1] = t1; acc[pidx][2] = t2; acc[pidx][3] = t3;
        }
      }
      __syncthreads();
#pragma unroll
      for (int pidx = 0; pidx < 3; ++pidx) {
        if (pidx < npairs) {
          double* Tb = W + leaf_wblk(pidx, pidx + round);
#pragma unroll
          for (int m = 0; m < 4; ++m) Tb[(cbase + 8 * m) * LEAF_WLD + r] = acc[pidx][m];
        }
      }
      __syncthreads();
    }
  }
  // dinv (ld 128): inverse in the upper block triangle, explicit zeros elsewhere
  for (int idx = tid; idx < LEAF_N * LEAF_N; idx += LEAF_THREADS) {
    const int rr = idx & (LEAF_N - 1), cc = idx >> 7;
    double v = 0.0;
    if (!failed && (rr >> 5) <= (cc >> 5)) v = W[leaf_wblk(rr >> 5, cc >> 5) + (cc & 31) * LEAF_WLD + (rr & 31)];
    dinv[idx] = v;
  }
}

// ---------------------------------------------------------------------------
// Vector solves (alpha = K^-1 y when y is a vector: ldiv!(alpha, kchol, y), src/cost.jl:79,89,106).
// Memory-bound matrix-vector kernels over blocks of the factor; every reduction has a fixed order.
// ---------------------------------------------------------------------------
// y(M) += alpha * A^T x,  op(A)[m,k] = A[k + m*lda]  (k contiguous): one warp per output element.
__global__ void __launch_bounds__(256) gemv_t_kernel(int M, int K, double alpha, const double* __restrict__ A,
                                                     long long lda, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const double* col = A + (long long)m * lda;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = 2 * lane;
  for (; k + 192 < K; k += 256) {
    const double2 a0 = *reinterpret_cast<const double2*>(col + k), a1 = *reinterpret_cast<const double2*>(col + k + 64);
    const double2 a2 = *reinterpret_cast<const double2*>(col + k + 128), a3 = *reinterpret_cast<const double2*>(col + k + 192);
    const double2 x0 = *reinterpret_cast<const double2*>(x + k), x1 = *reinterpret_cast<const double2*>(x + k + 64);
    const double2 x2 = *reinterpret_cast<const double2*>(x + k + 128), x3 = *reinterpret_cast<const double2*>(x + k + 192);
    s0 += a0.x * x0.x + a0.y * x0.y; s1 += a1.x * x1.x + a1.y * x1.y;
    s2 += a2.x * x2.x + a2.y * x2.y; s3 += a3.x * x3.x + a3.y * x3.y;
  }
  for (; k < K; k += 64) {
    const double2 a0 = *reinterpret_cast<const double2*>(col + k);
    const double2 x0 = *reinterpret_cast<const double2*>(x + k);
    s0 += a0.x * x0.x + a0.y * x0.y;
  }
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[m] += alpha * s;
}

// y(M) += alpha * A x,  op(A)[m,k] = A[m + k*lda]  (m contiguous): a CTA owns 64 rows, its 8 warps split k.
__global__ void __launch_bounds__(256) gemv_n_kernel(int M, int K, double alpha, const double* __restrict__ A,
                                                     long long lda, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 64 + 2 * lane;
  double sx0 = 0.0, sy0 = 0.0, sx1 = 0.0, sy1 = 0.0;
  if (m0 < M) {
    const double* p = A + m0;
    int k = warp;
    for (; k + 24 < K; k += 32) {
      const double2 a0 = *reinterpret_cast<const double2*>(p + (long long)k * lda);
      const double2 a1 = *reinterpret_cast<const double2*>(p + (long long)(k + 8) * lda);
      const double2 a2 = *reinterpret_cast<const double2*>(p + (long long)(k + 16) * lda);
      const double2 a3 = *reinterpret_cast<const double2*>(p + (long long)(k + 24) * lda);
      const double x0 = x[k], x1 = x[k + 8], x2 = x[k + 16], x3 = x[k + 24];
      sx0 += a0.x * x0; sy0 += a0.y * x0; sx1 += a1.x * x1; sy1 += a1.y * x1;
      sx0 += a2.x * x2; sy0 += a2.y * x2; sx1 += a3.x * x3; sy1 += a3.y * x3;
    }
    for (; k < K; k += 8) {
      const double2 a0 = *reinterpret_cast<const double2*>(p + (long long)k * lda);
      const double x0 = x[k];
      sx0 += a0.x * x0; sy0 += a0.y * x0;
    }
  }
  red[warp][2 * lane] = sx0 + sx1;
  red[warp][2 * lane + 1] = sy0 + sy1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int m = blockIdx.x * 64 + threadIdx.x;
    if (m < M) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      y[m] += alpha * s;
    }
  }
}

// v(128) <- s * op(Dinv) v with Dinv a 128x128 block (ld 128); trans != 0: op = transpose.  One CTA.
__global__ void __launch_bounds__(256, 1) leaf_mv_kernel(const double* __restrict__ dinv, double* __restrict__ v,
                                                         double s, int trans) {
  extern __shared__ __align__(16) double mv_smem[];
  double* S = mv_smem;                       // S(r,c) = S[c*129 + r]
  double* vs = S + LEAF_N * LEAF_LDS;        // 128
  double* part = vs + LEAF_N;                // 256
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int q0 = 0; q0 < 64; q0 += 8) {
    double t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) t[q] = dinv[tid + LEAF_THREADS * (q0 + q)];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + LEAF_THREADS * (q0 + q);
      S[(idx >> 7) * LEAF_LDS + (idx & 127)] = t[q];
    }
  }
  if (tid < LEAF_N) vs[tid] = v[tid];
  __syncthreads();
  const int m = tid & 127, half = tid >> 7;
  double a0 = 0.0, a1 = 0.0;
  if (trans) {
#pragma unroll 8
    for (int k = half * 64; k < half * 64 + 64; k += 2) { a0 += S[m * LEAF_LDS + k] * vs[k]; a1 += S[m * LEAF_LDS + k + 1] * vs[k + 1]; }
  } else {
#pragma unroll 8
    for (int k = half * 64; k < half * 64 + 64; k += 2) { a0 += S[k * LEAF_LDS + m] * vs[k]; a1 += S[(k + 1) * LEAF_LDS + m] * vs[k + 1]; }
  }
  part[tid] = a0 + a1;
  __syncthreads();
  if (tid < LEAF_N) v[tid] = s * (part[tid] + part[tid + 128]);
}
constexpr size_t LEAF_MV_SMEM_BYTES = (size_t)(LEAF_N * LEAF_LDS + LEAF_N + 256) * sizeof(double);

// dst (128x128 block of a column-major matrix, ldd; batch member blockIdx.x at dst + z*stride) <- Dinv block
// (src + z*dstride, ld 128): the upper part only, or (full) the whole block with the zeros below the diagonal.
__global__ void copy_dinv_128_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                     long long stride, long long dstride, int full) {
  dst += (long long)blockIdx.x * stride;
  src += (long long)blockIdx.x * dstride;
  for (int idx = threadIdx.x; idx < LEAF_N * LEAF_N; idx += blockDim.x) {
    const int r = idx & (LEAF_N - 1), c = idx >> 7;
    if (full || r <= c) dst[r + (long long)c * ldd] = src[idx];
  }
}

// dst (128x128 block, ldd; batch member blockIdx.x at dst + z*stride) <- TRANSPOSE of the Dinv block (src + z*dstride):
// the whole block, i.e. inv(U_leaf)^T in the lower part and explicit zeros above the diagonal.
__global__ void copy_dinv_128_t_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                       long long stride, long long dstride) {
  __shared__ double t[32][33];
  dst += (long long)blockIdx.x * stride;
  src += (long long)blockIdx.x * dstride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads = 32 x 8
  for (int tile = 0; tile < 16; ++tile) {
    const int br = tile & 3, bc = tile >> 2;                 // source tile (rows br, cols bc)
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) t[ty + 8 * q][tx] = src[(br * 32 + tx) + (bc * 32 + ty + 8 * q) * LEAF_N];   // t[c][r]
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[(bc * 32 + tx) + (long long)(br * 32 + ty + 8 * q) * ldd] = t[tx][ty + 8 * q];   // dst(c, r) = src(r, c)
  }
}

// dst(upper) <- src(upper), n x n column major (32 x 8 threads per 32 x 32 tile)
__global__ void copy_upper_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src, long long lds,
                                  long long n) {
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x > blockIdx.y) return;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const long long j = (long long)blockIdx.y * 32 + r;
    if (i < n && j < n && i <= j) dst[i + j * ldd] = src[i + j * lds];
  }
}

// A(lower) <- A(upper)^T for an n x n column-major matrix (used before K^-1 is
// handed to the host as a full symmetric matrix; reference: tc.K⁻¹ is dense,
// /root/reference/src/cost.jl:90-92).
__global__ void symmetrize_from_upper_kernel(double* __restrict__ A, long long ld, long long n) {
  __shared__ double tile[32][33];
  const long long bi = blockIdx.x, bj = blockIdx.y;   // tile (bi, bj) of the upper part, bi <= bj
  if (bi > bj) return;
  const int tx = threadIdx.x, ty = threadIdx.y;       // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long i = bi * 32 + tx, j = bj * 32 + r;
    tile[r][tx] = (i < n && j < n) ? A[i + j * ld] : 0.0;   // tile[r][tx] = A(i = bi*32+tx, j = bj*32+r)
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // write A(jj, ii) with jj = bj*32 + tx (row), ii = bi*32 + r (col); value = A(ii, jj) = tile[tx][r]
    const long long jj = bj * 32 + tx, ii = bi * 32 + r;
    if (jj < n && ii < n && jj > ii) A[jj + ii * ld] = tile[tx][r];
  }
}

// out[0] = sum_i 2*log(U_ii) over the first n diagonal entries; out[1] = dot(y, alpha) over n.
// Single CTA, fixed reduction order (deterministic).
__global__ void __launch_bounds__(1024, 1)
logdet_dot_kernel(const double* __restrict__ U, long long ld, long long n, const double* __restrict__ y,
                  const double* __restrict__ alpha, double* __restrict__ out) {
  __shared__ double s0[1024], s1[1024];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    a += log(U[i + i * ld]);
    b += y[i] * alpha[i];
  }
  s0[threadIdx.x] = a; s1[threadIdx.x] = b;
  __syncthreads();
  for (int st = 512; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) { s0[threadIdx.x] += s0[threadIdx.x + st]; s1[threadIdx.x] += s1[threadIdx.x + st]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = 2.0 * s0[0]; out[1] = s1[0]; }
}

// dst (rows_pad x cols_pad, ld) <- zero padded copy of src (rows x cols, lds)
__global__ void pad_copy_kernel(double* __restrict__ dst, long long ldd, long long rows_pad, long long cols_pad,
                                const double* __restrict__ src, long long lds, long long rows, long long cols) {
  const long long total = rows_pad * cols_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx % rows_pad, c = idx / rows_pad;
    dst[r + c * ldd] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
  }
}

// A <- A^T in place, n x n column major, n % 32 == 0.  One CTA (32 x 8 threads) per pair of 32 x 32 tiles
// (bi <= bj, linearised over the upper block triangle); HBM-bound: 16 n^2 bytes.
__global__ void __launch_bounds__(256) transpose_inplace_kernel(double* __restrict__ A, long long ld, long long nt) {
  __shared__ double t0[32][33], t1[32][33];
  long long lin = blockIdx.x;
  long long bj = (long long)((sqrt(8.0 * (double)lin + 1.0) - 1.0) * 0.5);
  while (bj * (bj + 1) / 2 > lin) --bj;
  while ((bj + 1) * (bj + 2) / 2 <= lin) ++bj;
  const long long bi = lin - bj * (bj + 1) / 2;
  const int tx = threadIdx.x, ty = threadIdx.y;
  double* P = A + bi * 32 + bj * 32 * ld;   // tile (bi, bj)
  double* Q = A + bj * 32 + bi * 32 * ld;   // tile (bj, bi)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = ty + 8 * r;
    t0[c][tx] = P[tx + (long long)c * ld];
    if (bi != bj) t1[c][tx] = Q[tx + (long long)c * ld];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = ty + 8 * r;
    Q[tx + (long long)c * ld] = t0[tx][c];
    if (bi != bj) P[tx + (long long)c * ld] = t1[tx][c];
  }
}

// ---------------------------------------------------------------------------
// out = A y for a symmetric A of which only the UPPER triangle is stored (the lower part of the buffer is never
// read): alpha = K^-1 y from the explicit inverse on the gradient path, replacing the two triangular sweeps
// (ldiv!(alpha, kchol, y), src/cost.jl:89,106) when K^-1 is formed anyway.  Two deterministic passes over the
// triangle (8 n^2 bytes): strictly-upper column parts (one warp per column) and row parts including the diagonal
// (a CTA owns 64 rows, its 8 warps split the columns).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) symv_upper_cols_kernel(long long n, const double* __restrict__ A, long long ld,
                                                              const double* __restrict__ y, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long j = blockIdx.x * 8LL + warp;
  if (j >= n) return;
  const double* col = A + j * ld;
  double s0 = 0.0, s1 = 0.0;
  const long long jeven = j & ~1LL;                  // rows [0, jeven) in double2 steps, row jeven (< j) separately
  for (long long i = 2 * lane; i < jeven; i += 64) {
    const double2 a = *reinterpret_cast<const double2*>(col + i);
    const double2 v = *reinterpret_cast<const double2*>(y + i);
    s0 += a.x * v.x; s1 += a.y * v.y;
  }
  if (lane == 0 && jeven < j) s0 += col[jeven] * y[jeven];
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) out[j] = s;                          // sum_{i < j} A_ij y_i
}

// out[i] += sum_{j >= i} A_ij y_j
__global__ void __launch_bounds__(256) symv_upper_rows_kernel(long long n, const double* __restrict__ A, long long ld,
                                                              const double* __restrict__ y, double* __restrict__ out) {
  __shared__ double red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r0 = blockIdx.x * 64LL;
  const long long i0 = r0 + 2 * lane;                 // this lane's two rows
  double sx = 0.0, sy = 0.0;
  if (i0 < n) {
    const double* p = A + i0;
    // columns r0 .. r0+63 (the diagonal 64-block): element-wise test j >= i
    for (long long j = r0 + warp; j < r0 + 64 && j < n; j += 8) {
      const double yj = y[j];
      if (j >= i0) sx += p[j * ld] * yj;
      if (j >= i0 + 1 && i0 + 1 < n) sy += p[1 + j * ld] * yj;
    }
    // columns right of the diagonal block: both rows valid, 16-byte loads
    for (long long j = r0 + 64 + warp; j < n; j += 8) {
      const double2 a = *reinterpret_cast<const double2*>(p + j * ld);
      const double yj = y[j];
      sx += a.x * yj; sy += a.y * yj;
    }
  }
  red[warp][2 * lane] = sx; red[warp][2 * lane + 1] = sy;
  __syncthreads();
  if (threadIdx.x < 64) {
    const long long i = r0 + threadIdx.x;
    if (i < n) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      out[i] += s;
    }
  }
}

}  // namespace gpr
