// Backend-agnostic blocked (recursive) dense algorithms on an upper-triangular
// column-major factor.  Everything is expressed through three primitives of a
// backend BE:
//
//   be.gemm(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, flags, batch, sA, sB, sC)
//        strided batch: member z uses A + z*sA, B + z*sB, C + z*sC
//        flags: BLK_UPPER_ONLY (C diagonal-anchored, only row <= col computed/stored),
//               BLK_K_FROM_N   (tile column block J contracts over k >= 128*J only: the triangular product W W^T)
//               BLK_SKIP_TILE00 (the 128x128 tile (0,0) of C is left alone)
//   be.potrf_leaf(A, lda, Dinv_blk, global_row_offset)   128x128 diagonal block:
//        A(upper) <- chol(A) (U^T U = A) and Dinv_blk <- inv(U) (full 128x128, zeros below the diagonal)
//   be.transpose_inplace(A, ld, n)   A <- A^T (n x n, in place)
//   be.copy_dinv_128_t(dst, ldd, src128, batch, stride, dstride)   the same with the TRANSPOSED Dinv block, full block
//   be.copy_dinv_128(dst, ldd, src128, batch, stride, dstride, full)
//        dst (128 block; member z at dst + z*stride) <- Dinv block (src + z*dstride); upper part only, or the
//        full block incl. the explicit zeros below the diagonal when full
//
// The product (libgpr_sm100a.so) instantiates this with the CUDA backend
// (DMMA GEMM + leaf kernels, gpr_api.cu).  tests/hostlogic instantiates it
// with a plain-loop CPU backend so that the index arithmetic of the
// recursions is checked in CI without a GPU; that CPU backend is test
// infrastructure and is never linked into the product library.
//
// Replaces the LAPACK calls of the reference path:
//   dpotrf('U')  /root/reference/src/cost.jl:77,87,104   /root/reference/src/predict.jl:31
//   dpotrs       /root/reference/src/cost.jl:79,89,106,107-109   /root/reference/src/predict.jl:32
//   dtrsm(R,U,N) /root/reference/src/predict.jl:84,90
// and computes K^-1 as trtri + lauum (2N^3/3) instead of potrs on the identity (2N^3).
//
// All sizes are multiples of 128 (the library pads to 128 with an identity
// diagonal).  A triangular factor is addressed as (T, ldt, n, b0): pointer to
// its top-left element, leading dimension, order, and the index of its first
// 128-block inside the global factor (selects the Dinv blocks).
#pragma once
#include <stdint.h>

namespace gpr {

constexpr int LEAF = 128;
constexpr int BLK_UPPER_ONLY = 1;   // == GEMM_UPPER_ONLY
constexpr int BLK_K_FROM_N = 2;     // == GEMM_K_FROM_N
constexpr int BLK_SKIP_TILE00 = 64; // == GEMM_SKIP_TILE00

template <class BE>
struct Blocked {
  BE& be;
  double* dinv;   // nblocks x (128*128), block b at dinv + b*128*128, ld 128

  Blocked(BE& b, double* d) : be(b), dinv(d) {}

  double* dinv_blk(int64_t b) const { return dinv + b * (int64_t)(LEAF * LEAF); }
  static int64_t split(int64_t n) { return ((n / LEAF) / 2) * LEAF; }

  // A(upper) <- U with U^T U = A.  A: n x n at the diagonal, global block b0.
  // Right-looking recursion that carries the whole row panel to the right of the block ("mr" further
  // columns of the same rows): factor the block, apply U^-T to the panel, update everything below/right.
  // Every launch is as wide as the remaining matrix, so only the last few block columns produce
  // launches with fewer tiles than SMs.
  void potrf(double* A, int64_t ld, int64_t n, int64_t b0) { potrf_panel(A, ld, n, 0, b0); }

  // leaf_lookahead: factor the NEXT 128x128 diagonal block on the backend's side queue as soon as its own tile of
  // the trailing update is done, while the main queue runs the rest of that update (the 256 single-CTA leaf
  // factorizations of an N = 32768 problem are otherwise a serial 21 ms chain during which the GPU idles).
  // Needs be.fork() / be.join() / be.side(on); must stay off when the caller already runs on the side queue.
  bool leaf_lookahead = false;

  void potrf_panel(double* A, int64_t ld, int64_t n, int64_t mr, int64_t b0, bool first_leaf_done = false) {
    if (n == LEAF) {
      if (first_leaf_done) be.join();                   // factored ahead on the side queue
      else be.potrf_leaf(A, ld, dinv_blk(b0), b0 * LEAF);
      if (mr > 0) {
        double* P = A + (int64_t)LEAF * ld;   // row panel right of the block: 128 x mr
        be.gemm('T', 'N', LEAF, mr, LEAF, 1.0, dinv_blk(b0), LEAF, P, ld, 0.0, P, ld, 0, 1, 0, 0, 0);
      }
      return;
    }
    const int64_t n1 = split(n), n2 = n - n1;
    double* A12 = A + n1 * ld;           // n1 x (n2 + mr): rest of the row panel
    double* A22 = A + n1 + n1 * ld;      // (n2) x (n2 + mr), diagonal anchored
    potrf_panel(A, ld, n1, n2 + mr, b0, first_leaf_done);
    if (!leaf_lookahead) {
      be.gemm('T', 'N', n2, n2 + mr, n1, -1.0, A12, ld, A12, ld, 1.0, A22, ld, BLK_UPPER_ONLY, 1, 0, 0, 0);
      potrf_panel(A22, ld, n2, mr, b0 + n1 / LEAF);
      return;
    }
    // side queue: tile (0,0) of A22 gets its update and is factored; main queue: the rest of the update
    be.fork();
    be.side(true);
    be.gemm('T', 'N', LEAF, LEAF, n1, -1.0, A12, ld, A12, ld, 1.0, A22, ld, BLK_UPPER_ONLY, 1, 0, 0, 0);
    be.potrf_leaf(A22, ld, dinv_blk(b0 + n1 / LEAF), (b0 + n1 / LEAF) * LEAF);
    be.side(false);
    be.gemm('T', 'N', n2, n2 + mr, n1, -1.0, A12, ld, A12, ld, 1.0, A22, ld, BLK_UPPER_ONLY | BLK_SKIP_TILE00, 1, 0, 0, 0);
    potrf_panel(A22, ld, n2, mr, b0 + n1 / LEAF, true);
  }

  // B (n x m) <- s * T^-T B,   T upper n x n, s = +-1
  // Strided batch (all trsm/trmm drivers): `bt` independent problems, member z uses T + z*sT, B + z*sB and the
  // Dinv blocks shifted by z*sD doubles.
  void trsm_LUT(const double* T, int64_t ldt, int64_t n, int64_t b0, double* B, int64_t ldb, int64_t m, double s,
                int64_t bt = 1, int64_t sT = 0, int64_t sB = 0, int64_t sD = 0) {
    if (n == LEAF) { be.gemm('T', 'N', LEAF, m, LEAF, s, dinv_blk(b0), LEAF, B, ldb, 0.0, B, ldb, 0, bt, sD, sB, sB); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    const double* Tb = T + n1 * ldt;
    const double* Tc = T + n1 + n1 * ldt;
    trsm_LUT(T, ldt, n1, b0, B, ldb, m, s, bt, sT, sB, sD);
    be.gemm('T', 'N', n2, m, n1, -s, Tb, ldt, B, ldb, 1.0, B + n1, ldb, 0, bt, sT, sB, sB);
    trsm_LUT(Tc, ldt, n2, b0 + n1 / LEAF, B + n1, ldb, m, s, bt, sT, sB, sD);
  }

  // B (n x m) <- s * T^-1 B
  void trsm_LUN(const double* T, int64_t ldt, int64_t n, int64_t b0, double* B, int64_t ldb, int64_t m, double s,
                int64_t bt = 1, int64_t sT = 0, int64_t sB = 0, int64_t sD = 0) {
    if (n == LEAF) { be.gemm('N', 'N', LEAF, m, LEAF, s, dinv_blk(b0), LEAF, B, ldb, 0.0, B, ldb, 0, bt, sD, sB, sB); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    const double* Tb = T + n1 * ldt;
    const double* Tc = T + n1 + n1 * ldt;
    trsm_LUN(Tc, ldt, n2, b0 + n1 / LEAF, B + n1, ldb, m, s, bt, sT, sB, sD);
    be.gemm('N', 'N', n1, m, n2, -s, Tb, ldt, B + n1, ldb, 1.0, B, ldb, 0, bt, sT, sB, sB);
    trsm_LUN(T, ldt, n1, b0, B, ldb, m, s, bt, sT, sB, sD);
  }

  // B (m x n) <- s * B T^-1
  void trsm_RUN(const double* T, int64_t ldt, int64_t n, int64_t b0, double* B, int64_t ldb, int64_t m, double s,
                int64_t bt = 1, int64_t sT = 0, int64_t sB = 0, int64_t sD = 0) {
    if (n == LEAF) { be.gemm('N', 'N', m, LEAF, LEAF, s, B, ldb, dinv_blk(b0), LEAF, 0.0, B, ldb, 0, bt, sB, sD, sB); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    const double* Tb = T + n1 * ldt;
    const double* Tc = T + n1 + n1 * ldt;
    double* B2 = B + n1 * ldb;
    trsm_RUN(T, ldt, n1, b0, B, ldb, m, s, bt, sT, sB, sD);
    be.gemm('N', 'N', m, n2, n1, -s, B, ldb, Tb, ldt, 1.0, B2, ldb, 0, bt, sB, sT, sB);
    trsm_RUN(Tc, ldt, n2, b0 + n1 / LEAF, B2, ldb, m, s, bt, sT, sB, sD);
  }

  // B (m x n) <- B T^T,  T upper n x n whose diagonal 128-blocks are the Dinv blocks
  // (i.e. T is the already inverted factor W = U^-1).
  void trmm_RUT(const double* T, int64_t ldt, int64_t n, int64_t b0, double* B, int64_t ldb, int64_t m) {
    if (n == LEAF) { be.gemm('N', 'T', m, LEAF, LEAF, 1.0, B, ldb, dinv_blk(b0), LEAF, 0.0, B, ldb, 0, 1, 0, 0, 0); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    const double* Tb = T + n1 * ldt;
    const double* Tc = T + n1 + n1 * ldt;
    double* B2 = B + n1 * ldb;
    trmm_RUT(T, ldt, n1, b0, B, ldb, m);
    be.gemm('N', 'T', m, n1, n2, 1.0, B2, ldb, Tb, ldt, 1.0, B, ldb, 0, 1, 0, 0, 0);
    trmm_RUT(Tc, ldt, n2, b0 + n1 / LEAF, B2, ldb, m);
  }

  // W(upper) <- inv(U), in place.  Needs the Dinv blocks produced by potrf.
  // Level-synchronous: the two half-size inversions of a node are independent, so all nodes of one depth
  // are processed as ONE strided batch (members `stride` doubles apart along the diagonal, Dinv blocks
  // `dstride` apart).  Every launch then covers the whole matrix width regardless of depth.
  // full_diag: write the diagonal 128-blocks of W with explicit zeros below the diagonal (needed by the
  // out-of-place W W^T product; only legal when W is not the buffer whose strict lower triangle keeps K).
  void trtri(double* W, int64_t ld, int64_t n, int64_t b0, bool full_diag = false) {
    trtri_batched(W, ld, n, b0, 1, 0, 0, full_diag);
  }
  void trtri_batched(double* W, int64_t ld, int64_t n, int64_t b0, int64_t bt, int64_t stride, int64_t dstride,
                     bool full_diag) {
    if (n == LEAF) { be.copy_dinv_128(W, ld, dinv_blk(b0), bt, stride, dstride, full_diag); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    double* W12 = W + n1 * ld;
    double* W22 = W + n1 + n1 * ld;
    trsm_LUN(W, ld, n1, b0, W12, ld, n2, 1.0, bt, stride, stride, dstride);                    // U11^-1 U12
    trsm_RUN(W22, ld, n2, b0 + n1 / LEAF, W12, ld, n1, -1.0, bt, stride, stride, dstride);     // -(.) U22^-1
    if (n1 == n2 && (bt == 1 || stride == 2 * n1 * (ld + 1))) {
      // children of all members are equally spaced: merge them into one batch of 2*bt
      trtri_batched(W, ld, n1, b0, 2 * bt, n1 * (ld + 1), (n1 / LEAF) * (int64_t)(LEAF * LEAF), full_diag);
    } else {
      trtri_batched(W, ld, n1, b0, bt, stride, dstride, full_diag);
      trtri_batched(W22, ld, n2, b0 + n1 / LEAF, bt, stride, dstride, full_diag);
    }
  }

  // Z(lower) <- U^-T = (U^-1)^T, out of place: U is only read (upper blocks + Dinv leaves), Z's lower triangle incl. full
  // diagonal 128-blocks is written, its upper part is never touched.  Bottom-up over the halving tree,
  //     Z = [ Z11 0 ; Z21 Z22 ],   Z21 = -U22^-T (U12^T Z11),
  // so every product has BOTH operands contraction-contiguous (T,N form, the fastest one of the DMMA kernel; the
  // top-down trtri above runs in the N,N form), the triangular factor Z11 costs half a GEMM (BLK_K_FROM_N), and the
  // result is already the transposed factor lauum_oop_t wants: no copy of U, no transpose pass.
  // Level-synchronous like trtri_batched: all nodes of one depth form one strided batch.
  void trtri_t(const double* U, int64_t ldu, double* Z, int64_t ldz, int64_t n, int64_t b0) {
    trtri_t_batched(U, ldu, Z, ldz, n, b0, 1, 0, 0, 0);
  }
  void trtri_t_batched(const double* U, int64_t ldu, double* Z, int64_t ldz, int64_t n, int64_t b0, int64_t bt, int64_t sU,
                       int64_t sZ, int64_t dstride) {
    if (n == LEAF) { be.copy_dinv_128_t(Z, ldz, dinv_blk(b0), bt, sZ, dstride); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    if (n1 == n2 && (bt == 1 || (sU == 2 * n1 * (ldu + 1) && sZ == 2 * n1 * (ldz + 1)))) {
      trtri_t_batched(U, ldu, Z, ldz, n1, b0, 2 * bt, n1 * (ldu + 1), n1 * (ldz + 1), (n1 / LEAF) * (int64_t)(LEAF * LEAF));
    } else {
      trtri_t_batched(U, ldu, Z, ldz, n1, b0, bt, sU, sZ, dstride);
      trtri_t_batched(U + n1 * (ldu + 1), ldu, Z + n1 * (ldz + 1), ldz, n2, b0 + n1 / LEAF, bt, sU, sZ, dstride);
    }
    const double* U12 = U + n1 * ldu;            // n1 x n2
    const double* U22 = U + n1 * (ldu + 1);
    double* Z21 = Z + n1;                        // n2 x n1
    be.gemm('T', 'N', n2, n1, n1, 1.0, U12, ldu, Z, ldz, 0.0, Z21, ldz, BLK_K_FROM_N, bt, sU, sZ, sZ);     // U12^T Z11
    trsm_LUT(U22, ldu, n2, b0 + n1 / LEAF, Z21, ldz, n1, -1.0, bt, sU, sZ, dstride);                         // -U22^-T (.)
  }

  // C(upper) <- W W^T out of place, W upper triangular with clean diagonal blocks (trtri(..., full_diag = true)):
  // C_IJ = sum_{K >= J} W_IK W_JK^T, one fully parallel launch (no recursion, no small launches).
  void lauum_oop(const double* W, int64_t ldw, int64_t n, double* C, int64_t ldc) {
    be.gemm('N', 'T', n, n, n, 1.0, W, ldw, W, ldw, 0.0, C, ldc, BLK_UPPER_ONLY | BLK_K_FROM_N, 1, 0, 0, 0);
  }

  // Same product from the TRANSPOSED factor: Wt(lower) = W^T (be.transpose_inplace of a trtri(..., full_diag = true)
  // result), C_IJ = sum_{K >= J} Wt_KI^T Wt_KJ.  Both operands are then contraction-contiguous, the fastest form of
  // the DMMA kernel (ncu, 4096^3: DMMA pipe 92.9 % active against 85.6 % for the N,T form of lauum_oop).
  void lauum_oop_t(const double* Wt, int64_t ldw, int64_t n, double* C, int64_t ldc) {
    be.gemm('T', 'N', n, n, n, 1.0, Wt, ldw, Wt, ldw, 0.0, C, ldc, BLK_UPPER_ONLY | BLK_K_FROM_N, 1, 0, 0, 0);
  }

  // W(upper) <- W W^T (upper part), in place, W = inv(U) as left by trtri.
  void lauum(double* W, int64_t ld, int64_t n, int64_t b0) {
    if (n == LEAF) {
      be.gemm('N', 'T', LEAF, LEAF, LEAF, 1.0, dinv_blk(b0), LEAF, dinv_blk(b0), LEAF, 0.0, W, ld, BLK_UPPER_ONLY, 1, 0, 0, 0);
      return;
    }
    const int64_t n1 = split(n), n2 = n - n1;
    double* W12 = W + n1 * ld;
    double* W22 = W + n1 + n1 * ld;
    lauum(W, ld, n1, b0);
    be.gemm('N', 'T', n1, n1, n2, 1.0, W12, ld, W12, ld, 1.0, W, ld, BLK_UPPER_ONLY, 1, 0, 0, 0);
    trmm_RUT(W22, ld, n2, b0 + n1 / LEAF, W12, ld, n1);
    lauum(W22, ld, n2, b0 + n1 / LEAF);
  }

  // Vector right-hand side (the loss path, y is a vector): same recursions with matrix-vector
  // primitives, memory bound (reads the triangle once per sweep).
  //   be.gemv(tA, M, K, alpha, A, lda, x, y):  y(M) += alpha * op(A)(M x K) * x(K)
  //   be.leaf_mv(tA, dinv_blk, v, s):          v(128) <- s * op(Dinv) * v
  void trsv_LUT(const double* T, int64_t ldt, int64_t n, int64_t b0, double* v, double s) {
    if (n == LEAF) { be.leaf_mv('T', dinv_blk(b0), v, s); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    trsv_LUT(T, ldt, n1, b0, v, s);
    be.gemv('T', n2, n1, -s, T + n1 * ldt, ldt, v, v + n1);
    trsv_LUT(T + n1 + n1 * ldt, ldt, n2, b0 + n1 / LEAF, v + n1, s);
  }
  void trsv_LUN(const double* T, int64_t ldt, int64_t n, int64_t b0, double* v, double s) {
    if (n == LEAF) { be.leaf_mv('N', dinv_blk(b0), v, s); return; }
    const int64_t n1 = split(n), n2 = n - n1;
    trsv_LUN(T + n1 + n1 * ldt, ldt, n2, b0 + n1 / LEAF, v + n1, s);
    be.gemv('N', n1, n2, -s, T + n1 * ldt, ldt, v + n1, v);
    trsv_LUN(T, ldt, n1, b0, v, s);
  }
  // v (n) <- (U^T U)^-1 v
  void potrsv(const double* U, int64_t ld, int64_t n, double* v) {
    trsv_LUT(U, ld, n, 0, v, 1.0);
    trsv_LUN(U, ld, n, 0, v, 1.0);
  }

  // B (n x m) <- (U^T U)^-1 B
  void potrs(const double* U, int64_t ld, int64_t n, double* B, int64_t ldb, int64_t m) {
    trsm_LUT(U, ld, n, 0, B, ldb, m, 1.0);
    trsm_LUN(U, ld, n, 0, B, ldb, m, 1.0);
  }
};

}  // namespace gpr
