// Block-cyclic multi-GPU drivers for the factor / inverse path at sizes where one factorization no longer
// fits (or is too slow on) a single device: BASELINE.json config 5 (N = 131072, 137 GB of FP64 K), SURVEY.md 8e.
//
// Replaces, across G ranks, the same LAPACK calls as csrc/blocked.hpp:
//   dpotrf('U') + dpotrs on y     /root/reference/src/cost.jl:104-106
//   K^-1 (dpotrs on the identity) /root/reference/src/cost.jl:107-109   (here trtri + lauum, 2N^3/3)
//
// Layout.  The padded Np x Np matrix is cut into block columns of width nb (a multiple of 128); block column J
// lives on rank J mod G as local block J / G (1-D block-cyclic).  The right-hand sides y (padded to nyp columns)
// form one more, narrower block column with index nblk, owned by rank nblk mod G and stored after that rank's
// matrix columns.  Every rank keeps its columns in ONE column-major array L (Np rows, leading dimension ld), so
// the row panel U(k, :) of step k is a contiguous-by-column slab of every rank's L and all trailing updates of a
// step are a single GEMM launch per rank.  Only the upper triangle is ever referenced; the strict lower
// triangle of L must be zero (the distributed covariance build writes it that way).
//
// Algorithms (U^T U = K, upper factor, all in place in L):
//   potrf   right-looking over block rows k: owner factors the nb x nb diagonal block (recursive panel code of
//           blocked.hpp, which also applies U_kk^-T to the owner's part of the row panel) -> U_kk and its Dinv
//           leaves go to every rank -> every rank finishes its part of the row panel -> the row panel is
//           gathered on every rank -> one rank-nb update of the local trailing columns (tiles above the
//           diagonal only).  The y columns ride along, so z = U^-T y falls out of the same updates.
//   trtri   bottom-up over block rows k: W(k, j) = -U_kk^-1 * sum_{k < m <= j} U(k, m) W(m, j) for the local
//           columns j > k (row panel of U gathered, one GEMM with a per-column contraction limit).  In the y
//           columns the same recurrence with the old content added is the back substitution: they end up
//           holding -alpha = -K^-1 y.
//   lauum   K^-1 = W W^T accumulated column panel by column panel: step k sends W(0:k, k) to every rank, which
//           adds W(i, k) W(j, k)^T to its columns j <= k (rows i <= j).
// Communication per phase and rank: 4 N^2 bytes received in total, in N / nb steps.
//
// The drivers are lock-step SPMD over the ranks held by THIS process: `ranks` holds all G ranks in the
// single-process multi-device mode (gpr_mgpu_*, peer copies over NVLink) and exactly one rank when every GPU
// has its own process.  All cross-rank data movement goes through COMM; all arithmetic through BE
// (the CUDA backend in the product, a plain-loop CPU backend in tests/hostlogic).
//
//   BE (in addition to the members blocked.hpp needs):
//     be.gemm_map(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, flags, map)
//        flags: BLK_MAP_UPPER  element (m, n) exists iff (map.row_gtile0 + m/128, m%128) <= (gt(n), n%128),
//                              gt(n) = map.col_gtile[n / 128]  (global 128-tile column of local tile n/128)
//               BLK_MAP_KUPTO  column tile contracts over k < (gt(n) - map.k_gtile0 + 1) * 128 only
//               BLK_MAP_BROWS  (tB = 'T') rows of B for column tile n/128 start at gt(n) * 128 instead of n
//   COMM (collective: called once per step, acts on every local rank):
//     comm.bcast_diag(k, with_owner)   Ukk (nb x nb, ld nb) and dinv blocks k*tpb.. <- owner's diagonal block k
//     comm.gather_rowpanel(k)          panel[p + ((J-k-1)*nb + c)*nb] <- U(k*nb + p, J*nb + c), J = k+1 .. nblk-1
//     comm.bcast_colpanel(k)           panel[i + c*Np] <- L_owner(i, block k col c), i < (k+1)*nb
//     comm.barrier()                   every rank's queued work is ordered before every rank's later work
//   The collectives themselves do not synchronise: the drivers place a barrier between producing a block on
//   one rank and reading it from another, and between the last remote read of a block and overwriting it
//   (a stream-ordered transport such as NCCL makes barrier() a no-op).  be.activate() selects the rank's device.
#pragma once
#include <stdint.h>

#include <vector>

#include "blocked.hpp"

namespace gpr {

constexpr int BLK_MAP_UPPER = 8;    // == GEMM_MAP_UPPER
constexpr int BLK_MAP_KUPTO = 16;   // == GEMM_MAP_KUPTO
constexpr int BLK_MAP_BROWS = 32;   // == GEMM_MAP_BROWS
constexpr int GTILE_RHS = 1 << 22;  // global tile index of the y columns: beyond every matrix tile

struct TileMap {
  const int* col_gtile;   // device (or host, CPU backend) array: global 128-tile column of each local 128-tile column
  int row_gtile0;         // global 128-tile row of the first row of C
  int k_gtile0;           // global 128-tile row of the first contraction index
};

struct DistLayout {
  int G = 1;
  int64_t Np = 0, nb = 0, nblk = 0, nyp = 0;
  int tpb() const { return (int)(nb / LEAF); }
  int owner(int64_t J) const { return (int)(J % G); }
  int y_owner() const { return (int)(nblk % G); }
  int64_t nloc(int r) const { return nblk > r ? (nblk - r + G - 1) / G : 0; }          // matrix blocks owned by r
  int64_t count_le(int r, int64_t k) const { return k >= r ? (k - r) / G + 1 : 0; }    // owned blocks J <= k
  int64_t ycol0(int r) const { return nloc(r) * nb; }
  int64_t lcols(int r) const { return nloc(r) * nb + (r == y_owner() ? nyp : 0); }
  int64_t ltiles(int r) const { return lcols(r) / LEAF; }
  int gtile(int r, int64_t jt) const {
    const int64_t t = tpb();
    if (jt >= nloc(r) * t) return GTILE_RHS + (int)(jt - nloc(r) * t);
    return (int)(((jt / t) * G + r) * t + jt % t);
  }
};

template <class BE>
struct DistRank {
  int r = 0;
  BE* be = nullptr;
  double* L = nullptr;       // Np x lcols(r), ld
  int64_t ld = 0;
  double* dinv = nullptr;    // Np/128 leaves, global leaf index (every rank holds all leaves it ever needs)
  double* Ukk = nullptr;     // nb x nb
  double* panel = nullptr;   // Np * nb doubles: row panel (nb x Np) or column panel (Np x nb)
  const int* gtile = nullptr;   // ltiles(r) entries, readable by BE
};

template <class BE, class COMM>
struct DistBlocked {
  DistLayout lay;
  std::vector<DistRank<BE>>& ranks;
  COMM& comm;

  DistBlocked(const DistLayout& l, std::vector<DistRank<BE>>& rk, COMM& c) : lay(l), ranks(rk), comm(c) {}

  // L(upper) <- U, y columns <- U^-T y
  void potrf() {
    const int64_t nb = lay.nb, Np = lay.Np;
    const int tpb = lay.tpb();
    for (int64_t k = 0; k < lay.nblk; ++k) {
      const int o = lay.owner(k);
      for (auto& R : ranks) {
        if (R.r != o) continue;
        R.be->activate();
        const int64_t kb = k / lay.G;
        Blocked<BE> blk(*R.be, R.dinv);
        blk.potrf_panel(R.L + k * nb + kb * nb * R.ld, R.ld, nb, lay.lcols(o) - (kb + 1) * nb, k * tpb);
      }
      if (lay.G > 1) {
        comm.barrier();
        comm.bcast_diag(k, false);
      }
      for (auto& R : ranks) {
        if (R.r == o) continue;
        const int64_t c0 = lay.count_le(R.r, k) * nb, m = lay.lcols(R.r) - c0;
        if (m <= 0) continue;
        R.be->activate();
        Blocked<BE> blk(*R.be, R.dinv);
        blk.trsm_LUT(R.Ukk, nb, nb, k * tpb, R.L + k * nb + c0 * R.ld, R.ld, m, 1.0);
      }
      const int64_t Mrows = Np - (k + 1) * nb;
      if (Mrows <= 0) break;   // last block row: nothing below it
      comm.barrier();
      comm.gather_rowpanel(k);
      for (auto& R : ranks) {
        const int64_t c0 = lay.count_le(R.r, k) * nb, m = lay.lcols(R.r) - c0;
        if (m <= 0) continue;
        R.be->activate();
        TileMap map{R.gtile + c0 / LEAF, (int)((k + 1) * tpb), 0};
        R.be->gemm_map('T', 'N', Mrows, m, nb, -1.0, R.panel, nb, R.L + k * nb + c0 * R.ld, R.ld, 1.0,
                       R.L + (k + 1) * nb + c0 * R.ld, R.ld, BLK_MAP_UPPER, map);
      }
    }
  }

  // L(upper) <- W = U^-1 (diagonal 128-blocks with explicit zeros below the diagonal), y columns <- -U^-1 (y columns)
  void trtri() {
    const int64_t nb = lay.nb, Np = lay.Np;
    const int tpb = lay.tpb();
    comm.barrier();
    for (int64_t k = lay.nblk - 1; k >= 0; --k) {
      const int o = lay.owner(k);
      const int64_t Krem = Np - (k + 1) * nb;
      comm.bcast_diag(k, true);
      if (Krem > 0) comm.gather_rowpanel(k);
      comm.barrier();   // every rank holds its copies before row panel k / the diagonal block are overwritten
      for (auto& R : ranks) {
        const int64_t c0 = lay.count_le(R.r, k) * nb;
        const int64_t mreg = lay.nloc(R.r) * nb - c0, m = lay.lcols(R.r) - c0;
        R.be->activate();
        double* rowp = R.L + k * nb + c0 * R.ld;
        const double* below = R.L + (k + 1) * nb + c0 * R.ld;
        if (Krem > 0 && mreg > 0) {
          TileMap map{R.gtile + c0 / LEAF, 0, (int)((k + 1) * tpb)};
          R.be->gemm_map('N', 'N', nb, mreg, Krem, 1.0, R.panel, nb, below, R.ld, 0.0, rowp, R.ld, BLK_MAP_KUPTO, map);
        }
        if (Krem > 0 && m > mreg)
          R.be->gemm('N', 'N', nb, m - mreg, Krem, 1.0, R.panel, nb, below + mreg * R.ld, R.ld, 1.0, rowp + mreg * R.ld, R.ld, 0, 1, 0, 0, 0);
        if (m > 0) {
          Blocked<BE> blk(*R.be, R.dinv);
          blk.trsm_LUN(R.Ukk, nb, nb, k * tpb, rowp, R.ld, m, -1.0);
        }
        if (R.r == o) {
          Blocked<BE> blk(*R.be, R.dinv);
          blk.trtri(R.L + k * nb + (k / lay.G) * nb * R.ld, R.ld, nb, k * tpb, true);
        }
      }
    }
  }

  // L(upper) <- W W^T (matrix columns only)
  void lauum() {
    const int64_t nb = lay.nb, Np = lay.Np;
    comm.barrier();
    for (int64_t k = 0; k < lay.nblk; ++k) {
      const int o = lay.owner(k);
      comm.bcast_colpanel(k);
      comm.barrier();   // the owner overwrites block column k below
      for (auto& R : ranks) {
        const int64_t cnt = lay.count_le(R.r, k);
        const int64_t nacc = (R.r == o) ? cnt - 1 : cnt;
        const int64_t Mrows = (k + 1) * nb;
        R.be->activate();
        if (nacc > 0) {
          TileMap map{R.gtile, 0, 0};
          R.be->gemm_map('N', 'T', Mrows, nacc * nb, nb, 1.0, R.panel, Np, R.panel, Np, 1.0, R.L, R.ld,
                         BLK_MAP_UPPER | BLK_MAP_BROWS, map);
        }
        if (R.r == o) {
          TileMap map{R.gtile + (cnt - 1) * lay.tpb(), 0, 0};
          R.be->gemm_map('N', 'T', Mrows, nb, nb, 1.0, R.panel, Np, R.panel, Np, 0.0, R.L + (cnt - 1) * nb * R.ld, R.ld,
                         BLK_MAP_UPPER | BLK_MAP_BROWS, map);
        }
      }
    }
  }
};

}  // namespace gpr
