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
constexpr int LEAF_LDS = 129;                                  // padded smem leading dimension
constexpr int LEAF_THREADS = 256;
constexpr int LEAF_PACKED = LEAF_N * (LEAF_N + 1) / 2;
constexpr size_t LEAF_SMEM_BYTES = (size_t)(LEAF_N * LEAF_LDS + LEAF_PACKED) * sizeof(double);

// A: 128x128 block (column major, lda) on the diagonal of the global matrix at
// global row/col offset goff.  On exit the upper triangle of the block holds U
// (U^T U = A); the strict lower triangle is left untouched (it keeps K, which
// the reference's dpotrf('U') also leaves in place: test/test_loss.jl:46).
// dinv: 128x128 (ld 128), receives inv(U) with explicit zeros below the diagonal.
// info: first failing pivot (1-based global index), LAPACK style; 0 = ok.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, long long lda, double* __restrict__ dinv, long long* info, long long goff) {
  extern __shared__ __align__(16) double leaf_smem[];
  double* S = leaf_smem;                          // S(r,c) = S[c*LEAF_LDS + r]
  double* Wp = leaf_smem + LEAF_N * LEAF_LDS;     // packed upper inverse: W(i,j) = Wp[j*(j+1)/2 + i]
  const int tid = threadIdx.x;

  for (int idx = tid; idx < LEAF_N * LEAF_N; idx += LEAF_THREADS) {
    const int r = idx & (LEAF_N - 1), c = idx >> 7;
    S[c * LEAF_LDS + r] = A[r + (long long)c * lda];
  }
  __syncthreads();

  // right-looking U^T U factorization of the upper triangle
  const int j = tid & (LEAF_N - 1);
  const int half = tid >> 7;
  bool failed = false;
  for (int k = 0; k < LEAF_N; ++k) {
    const double piv = S[k * LEAF_LDS + k];
    if (!(piv > 0.0)) {   // also catches NaN
      if (tid == 0 && *info == 0) *info = goff + k + 1;
      failed = true;
      break;
    }
    const double d = sqrt(piv);
    const double inv = 1.0 / d;
    __syncthreads();   // everyone has read the pivot
    if (half == 0) {
      if (j > k) S[j * LEAF_LDS + k] *= inv;
      else if (j == k) S[k * LEAF_LDS + k] = d;
    }
    __syncthreads();
    if (j > k) {
      const double ukj = S[j * LEAF_LDS + k];
      for (int i = k + 1 + half; i <= j; i += 2) S[j * LEAF_LDS + i] -= S[i * LEAF_LDS + k] * ukj;
    }
    __syncthreads();
  }
  __syncthreads();

  // write U (upper part only)
  for (int idx = tid; idx < LEAF_N * LEAF_N; idx += LEAF_THREADS) {
    const int r = idx & (LEAF_N - 1), c = idx >> 7;
    if (r <= c) A[r + (long long)c * lda] = S[c * LEAF_LDS + r];
  }

  // inverse of the upper factor, column j owned by thread j (row sweep from the bottom)
  if (!failed && tid < LEAF_N) {
    const int cj = tid;
    double* w = Wp + cj * (cj + 1) / 2;
    for (int i = LEAF_N - 1; i >= 0; --i) {   // same i in every lane: U(i,k) reads are broadcasts
      if (i > cj) continue;
      double s0 = (i == cj) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = i + 1;
      for (; k + 3 <= cj; k += 4) {
        s0 -= S[k * LEAF_LDS + i] * w[k];
        s1 -= S[(k + 1) * LEAF_LDS + i] * w[k + 1];
        s2 -= S[(k + 2) * LEAF_LDS + i] * w[k + 2];
        s3 -= S[(k + 3) * LEAF_LDS + i] * w[k + 3];
      }
      for (; k <= cj; ++k) s0 -= S[k * LEAF_LDS + i] * w[k];
      w[i] = ((s0 + s1) + (s2 + s3)) / S[i * LEAF_LDS + i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < LEAF_N * LEAF_N; idx += LEAF_THREADS) {
    const int r = idx & (LEAF_N - 1), c = idx >> 7;
    double v = 0.0;
    if (!failed && r <= c) v = Wp[c * (c + 1) / 2 + r];
    dinv[idx] = v;
  }
}

// dst (128x128 block of a column-major matrix, ldd): upper part <- src (ld 128)
__global__ void copy_upper_128_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src) {
  for (int idx = threadIdx.x; idx < LEAF_N * LEAF_N; idx += blockDim.x) {
    const int r = idx & (LEAF_N - 1), c = idx >> 7;
    if (r <= c) dst[r + (long long)c * ldd] = src[idx];
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

}  // namespace gpr
