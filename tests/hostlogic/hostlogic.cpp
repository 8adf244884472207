// TEST INFRASTRUCTURE ONLY -- never linked into libgpr_sm100a.so.
// Plain-loop CPU backend for csrc/blocked.hpp so that the recursion / index
// arithmetic of the blocked potrf / trsm / trtri / lauum / potrs drivers can be
// checked in the CPU test-suite (pytest -m "not gpu").  The product library
// instantiates the same template with the CUDA backend only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../gaussianprocessregression.jl_b200/csrc/blocked.hpp"
#include "../../gaussianprocessregression.jl_b200/csrc/dist_blocked.hpp"

namespace {

struct CpuBE {
  long long info = 0;
  long long gemm_calls = 0;
  void activate() {}
  void fork() {}
  void join() {}
  void side(bool) {}
  // tile-mapped GEMM of csrc/dist_blocked.hpp (same predicate the CUDA kernel evaluates per 128-tile)
  void gemm_map(char tA, char tB, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int flags, const gpr::TileMap& map) {
    gemm_calls++;
    std::vector<double> tmp((size_t)M * N, 0.0);
    std::vector<char> live((size_t)M * N, 0);
    for (int64_t n = 0; n < N; ++n) {
      const int64_t gt = map.col_gtile[n / gpr::LEAF];
      int64_t kend = K;
      if (flags & gpr::BLK_MAP_KUPTO) kend = std::min<int64_t>(K, (gt - map.k_gtile0 + 1) * gpr::LEAF);
      const int64_t brow = (flags & gpr::BLK_MAP_BROWS) ? gt * gpr::LEAF + n % gpr::LEAF : n;   // op(B) column
      for (int64_t m = 0; m < M; ++m) {
        if (flags & gpr::BLK_MAP_UPPER) {
          const int64_t rt = map.row_gtile0 + m / gpr::LEAF;
          if (rt > gt || (rt == gt && m % gpr::LEAF > n % gpr::LEAF)) continue;
        }
        double s = 0.0;
        for (int64_t k = 0; k < kend; ++k) {
          const double a = (tA == 'T') ? A[k + m * lda] : A[m + k * lda];
          const double b = (tB == 'T') ? B[brow + k * ldb] : B[k + brow * ldb];
          s += a * b;
        }
        tmp[m + n * M] = s;
        live[m + n * M] = 1;
      }
    }
    for (int64_t n = 0; n < N; ++n)
      for (int64_t m = 0; m < M; ++m) {
        if (!live[m + n * M]) continue;
        double r = alpha * tmp[m + n * M];
        if (beta != 0.0) r += beta * C[m + n * ldc];
        C[m + n * ldc] = r;
      }
  }
  void gemm(char tA, char tB, int64_t M, int64_t N, int64_t K, double alpha, const double* A0, int64_t lda,
            const double* B0, int64_t ldb, double beta, double* C0, int64_t ldc, int flags, int64_t batch = 1,
            int64_t sA = 0, int64_t sB = 0, int64_t sC = 0) {
    gemm_calls++;
    for (int64_t z = 0; z < batch; ++z) {
      const double* A = A0 + z * sA;
      const double* B = B0 + z * sB;
      double* C = C0 + z * sC;
      // like the GPU kernel: a tile reads all of its operands before it stores, so C may alias A or B
      std::vector<double> tmp((size_t)M * N);
      for (int64_t n = 0; n < N; ++n)
        for (int64_t m = 0; m < M; ++m) {
          double s = 0.0;
          const int64_t kbeg = (flags & gpr::BLK_K_FROM_N) ? (n / gpr::LEAF) * gpr::LEAF : 0;
          for (int64_t k = kbeg; k < K; ++k) {
            const double a = (tA == 'T') ? A[k + m * lda] : A[m + k * lda];
            const double b = (tB == 'T') ? B[n + k * ldb] : B[k + n * ldb];
            s += a * b;
          }
          tmp[m + n * M] = s;
        }
      for (int64_t n = 0; n < N; ++n)
        for (int64_t m = 0; m < M; ++m) {
          if ((flags & gpr::BLK_UPPER_ONLY) && (m / gpr::LEAF > n / gpr::LEAF || (m / gpr::LEAF == n / gpr::LEAF && m > n))) continue;
          if ((flags & gpr::BLK_SKIP_TILE00) && m < gpr::LEAF && n < gpr::LEAF) continue;
          double r = alpha * tmp[m + n * M];
          if (beta != 0.0) r += beta * C[m + n * ldc];
          C[m + n * ldc] = r;
        }
    }
  }
  void potrf_leaf(double* A, int64_t lda, double* dinv, int64_t goff) {
    const int n = gpr::LEAF;
    for (int k = 0; k < n; ++k) {
      double p = A[k + k * lda];
      if (!(p > 0.0)) { if (!info) info = goff + k + 1; memset(dinv, 0, sizeof(double) * n * n); return; }
      const double d = std::sqrt(p);
      A[k + k * lda] = d;
      for (int j = k + 1; j < n; ++j) A[k + j * lda] /= d;
      for (int j = k + 1; j < n; ++j)
        for (int i = k + 1; i <= j; ++i) A[i + j * lda] -= A[k + i * lda] * A[k + j * lda];
    }
    for (int j = 0; j < n; ++j) {
      for (int i = n - 1; i >= 0; --i) {
        if (i > j) { dinv[i + j * n] = 0.0; continue; }
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = i + 1; k <= j; ++k) s -= A[i + k * lda] * dinv[k + j * n];
        dinv[i + j * n] = s / A[i + i * lda];
      }
    }
  }
  void gemv(char tA, int64_t M, int64_t K, double alpha, const double* A, int64_t lda, const double* x, double* y) {
    for (int64_t m = 0; m < M; ++m) {
      double s = 0.0;
      for (int64_t k = 0; k < K; ++k) s += ((tA == 'T') ? A[k + m * lda] : A[m + k * lda]) * x[k];
      y[m] += alpha * s;
    }
  }
  void leaf_mv(char tA, const double* dinv, double* v, double s) {
    const int n = gpr::LEAF;
    double tmp[gpr::LEAF];
    for (int m = 0; m < n; ++m) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += ((tA == 'T') ? dinv[k + m * n] : dinv[m + k * n]) * v[k];
      tmp[m] = s * acc;
    }
    for (int m = 0; m < n; ++m) v[m] = tmp[m];
  }
  void transpose_inplace(double* A, int64_t ld, int64_t n) {
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i < j; ++i) std::swap(A[i + j * ld], A[j + i * ld]);
  }
  void copy_dinv_128_t(double* dst, int64_t ldd, const double* src, int64_t batch, int64_t stride, int64_t dstride) {
    for (int64_t z = 0; z < batch; ++z)
      for (int c = 0; c < gpr::LEAF; ++c)
        for (int r = 0; r < gpr::LEAF; ++r) dst[z * stride + c + (int64_t)r * ldd] = src[z * dstride + r + c * gpr::LEAF];
  }
  void copy_dinv_128(double* dst, int64_t ldd, const double* src, int64_t batch, int64_t stride, int64_t dstride,
                     bool full) {
    for (int64_t z = 0; z < batch; ++z)
      for (int c = 0; c < gpr::LEAF; ++c)
        for (int r = 0; r < gpr::LEAF; ++r)
          if (full || r <= c) dst[z * stride + r + c * ldd] = src[z * dstride + r + c * gpr::LEAF];
  }
};

// All G ranks live in this process; a collective is a set of plain copies between the ranks' arrays.
struct CpuComm {
  gpr::DistLayout lay;
  std::vector<gpr::DistRank<CpuBE>>* ranks = nullptr;
  long long barriers = 0;
  void barrier() { barriers++; }
  void bcast_diag(int64_t k, bool with_owner, int b) {
    const int o = lay.owner(k);
    const int64_t nb = lay.nb, kb = k / lay.G;
    const auto& S = (*ranks)[o];
    for (auto& R : *ranks) {
      if (R.r == o && !with_owner) continue;
      for (int64_t c = 0; c < nb; ++c)
        for (int64_t p = 0; p < nb; ++p) R.Ukk[b][p + c * nb] = S.L[k * nb + p + (kb * nb + c) * S.ld];
      if (R.r != o)
        memcpy(R.dinv + k * lay.tpb() * 128 * 128, S.dinv + k * lay.tpb() * 128 * 128, sizeof(double) * lay.tpb() * 128 * 128);
    }
  }
  void gather_rowpanel(int64_t k, int b) {
    const int64_t nb = lay.nb;
    for (auto& R : *ranks)
      for (int64_t J = k + 1; J < lay.nblk; ++J) {
        const auto& S = (*ranks)[lay.owner(J)];
        for (int64_t c = 0; c < nb; ++c)
          for (int64_t p = 0; p < nb; ++p)
            R.panel[b][p + ((J - k - 1) * nb + c) * nb] = S.L[k * nb + p + ((J / lay.G) * nb + c) * S.ld];
      }
  }
  void gather_rowpanel_t(int64_t k, int b) {
    const int64_t nb = lay.nb, Krem = lay.Np - (k + 1) * nb;
    for (auto& R : *ranks)
      for (int64_t J = k + 1; J < lay.nblk; ++J) {
        const auto& S = (*ranks)[lay.owner(J)];
        for (int64_t c = 0; c < nb; ++c)
          for (int64_t p = 0; p < nb; ++p)
            R.panel[b][((J - k - 1) * nb + c) + p * Krem] = S.L[k * nb + p + ((J / lay.G) * nb + c) * S.ld];
      }
  }
  void bcast_colpanel(int64_t k, int b) {
    const int64_t nb = lay.nb;
    const auto& S = (*ranks)[lay.owner(k)];
    for (auto& R : *ranks)
      for (int64_t c = 0; c < nb; ++c)
        for (int64_t i = 0; i < (k + 1) * nb; ++i) R.panel[b][c + i * nb] = S.L[i + ((k / lay.G) * nb + c) * S.ld];
  }
};

}  // namespace

extern "C" {

// Distributed (block-cyclic over G simulated ranks) factorization of A (n x n, n % nb == 0) with right-hand
// sides Y (n x nyp).  mode 0: potrf (A upper <- U, Y <- U^-T Y), 1: + trtri (A upper <- U^-1, Y <- -A^-1 Y),
// 2: + lauum (A upper <- A^-1).  The strict lower triangle of A comes back zero.
long long hl_dist_factor(double* A, int64_t n, int64_t nb, int G, double* Y, int64_t nyp, int mode) {
  gpr::DistLayout lay;
  lay.G = G; lay.Np = n; lay.nb = nb; lay.nblk = n / nb; lay.nyp = nyp;
  std::vector<CpuBE> bes(G);
  std::vector<gpr::DistRank<CpuBE>> ranks(G);
  std::vector<std::vector<double>> Ls(G), dinvs(G), Ukks(2 * G), panels(2 * G);
  std::vector<std::vector<int>> gts(G);
  for (int r = 0; r < G; ++r) {
    Ls[r].assign((size_t)n * std::max<int64_t>(lay.lcols(r), 1), 0.0);
    dinvs[r].assign((size_t)n * 128, 0.0);
    for (int b = 0; b < 2; ++b) {
      Ukks[2 * r + b].assign((size_t)nb * nb, 0.0);
      panels[2 * r + b].assign((size_t)n * nb, 0.0);
    }
    gts[r].resize(std::max<int64_t>(lay.ltiles(r), 1));
    for (int64_t t = 0; t < lay.ltiles(r); ++t) gts[r][t] = lay.gtile(r, t);
    for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
      const int64_t J = lay.gblock(r, lb);
      for (int64_t c = 0; c < nb; ++c)
        for (int64_t i = 0; i <= J * nb + c; ++i) Ls[r][i + (lb * nb + c) * n] = A[i + (J * nb + c) * n];   // upper only
    }
    if (r == lay.y_owner())
      for (int64_t c = 0; c < nyp; ++c) memcpy(&Ls[r][(size_t)(lay.ycol0(r) + c) * n], Y + c * n, sizeof(double) * n);
    ranks[r] = gpr::DistRank<CpuBE>{r, &bes[r], Ls[r].data(), n, dinvs[r].data(), {Ukks[2 * r].data(), Ukks[2 * r + 1].data()},
                                    {panels[2 * r].data(), panels[2 * r + 1].data()}, gts[r].data()};
  }
  CpuComm comm;
  comm.lay = lay; comm.ranks = &ranks;
  gpr::DistBlocked<CpuBE, CpuComm> db(lay, ranks, comm);
  db.potrf();
  if (mode >= 1) db.trtri();
  if (mode >= 2) db.lauum();
  long long info = 0;
  for (int r = 0; r < G; ++r) {
    if (bes[r].info && (!info || bes[r].info < info)) info = bes[r].info;
    for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
      const int64_t J = lay.gblock(r, lb);
      for (int64_t c = 0; c < nb; ++c) memcpy(A + (J * nb + c) * n, &Ls[r][(size_t)(lb * nb + c) * n], sizeof(double) * n);
    }
    if (r == lay.y_owner())
      for (int64_t c = 0; c < nyp; ++c) memcpy(Y + c * n, &Ls[r][(size_t)(lay.ycol0(r) + c) * n], sizeof(double) * n);
  }
  return info;
}

// A: n x n column major (n % 128 == 0), factored in place.  mode 0: potrf, 1: +trtri, 2: +lauum.
long long hl_factor(double* A, int64_t n, int mode, long long* gemm_calls) {
  CpuBE be;
  std::vector<double> dinv((size_t)n * 128);
  gpr::Blocked<CpuBE> blk(be, dinv.data());
  blk.leaf_lookahead = (mode & 16) != 0;   // the ordering the CUDA product uses (side-queue calls are no-ops here)
  mode &= 15;
  blk.potrf(A, n, n, 0);
  if (mode == 5) {   // the product path: lower(Z) = U^-T bottom-up (T,N products), C = Z^T Z; Z's upper part stays poisoned
    std::vector<double> Z((size_t)n * n, std::nan("")), C((size_t)n * n, 0.0);
    blk.trtri_t(A, n, Z.data(), n, n, 0);
    blk.lauum_oop_t(Z.data(), n, n, C.data(), n);
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i <= j; ++i) A[i + j * n] = C[i + j * n];
  } else if (mode == 3 || mode == 4) {   // out-of-place inverse: W = copy of U with clean diagonal blocks, C = W W^T written back to A (upper)
    std::vector<double> W((size_t)n * n), C((size_t)n * n, 0.0);
    memcpy(W.data(), A, sizeof(double) * n * n);
    blk.trtri(W.data(), n, n, 0, true);
    if (mode == 3) blk.lauum_oop(W.data(), n, n, C.data(), n);
    else { be.transpose_inplace(W.data(), n, n); blk.lauum_oop_t(W.data(), n, n, C.data(), n); }   // the product path
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i <= j; ++i) A[i + j * n] = C[i + j * n];
  } else {
    if (mode >= 1) blk.trtri(A, n, n, 0);
    if (mode >= 2) blk.lauum(A, n, n, 0);
  }
  if (gemm_calls) *gemm_calls = be.gemm_calls;
  return be.info;
}

// potrf(A) then B (n x m) <- A^-1 B
long long hl_potrs(double* A, int64_t n, double* B, int64_t m) {
  CpuBE be;
  std::vector<double> dinv((size_t)n * 128);
  gpr::Blocked<CpuBE> blk(be, dinv.data());
  blk.potrf(A, n, n, 0);
  blk.potrs(A, n, n, B, n, m);
  return be.info;
}

// potrf(A) then v (n) <- A^-1 v through the vector (trsv) recursions
long long hl_potrsv(double* A, int64_t n, double* v) {
  CpuBE be;
  std::vector<double> dinv((size_t)n * 128);
  gpr::Blocked<CpuBE> blk(be, dinv.data());
  blk.potrf(A, n, n, 0);
  blk.potrsv(A, n, n, v);
  return be.info;
}

// potrf(A) then B (m x n) <- B U^-1
long long hl_trsm_run(double* A, int64_t n, double* B, int64_t m) {
  CpuBE be;
  std::vector<double> dinv((size_t)n * 128);
  gpr::Blocked<CpuBE> blk(be, dinv.data());
  blk.potrf(A, n, n, 0);
  blk.trsm_RUN(A, n, n, 0, B, m, m, 1.0);
  return be.info;
}

}  // extern "C"
