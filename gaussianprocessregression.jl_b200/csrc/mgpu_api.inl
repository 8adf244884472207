// Single-process multi-device NLML + gradient (gpr_mgpu_*, include/gpr_sm100a.h): BASELINE.json config 5.
// Included at the end of gpr_api.cu (same translation unit: shares gpr_ctx, CudaBE, the kernels).
//
// One rank per entry of the device list; a device may appear more than once ("virtual ranks": the whole
// distributed code path then runs on a single GPU, which is how tests/test_gpu_parity.py covers it on a
// 1-GPU box).  The algorithms are csrc/dist_blocked.hpp; three transports move the panels between ranks:
//   0  LocalComm   all ranks in this process: pull kernels / copy engines over peer memory (NVLink), CUDA events
//   1  PackedComm  all ranks in this process, but through the pack -> collective -> unpack data path of the NCCL
//                  transport with plain device copies as the "collective" (1-GPU test of that data path)
//   2  PackedComm  one rank per PROCESS (torchrun), NCCL broadcast / all-gather / all-reduce (gpr_dist_create)
namespace {

constexpr int MGPU_MAX_RANKS = 16;

struct PeerPtrs { const double* p[MGPU_MAX_RANKS]; };

// dst (rows x cols, ldd) <- src (rows x cols, lds); rows even, both 16-byte aligned.  src may be peer memory.
__global__ void __launch_bounds__(256) copy2d_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                                     long long lds, long long rows, long long cols) {
  const long long r2 = rows >> 1;
  for (long long c = blockIdx.y; c < cols; c += gridDim.y)
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < r2; p += (long long)gridDim.x * blockDim.x)
      reinterpret_cast<double2*>(dst + c * ldd)[p] = reinterpret_cast<const double2*>(src + c * lds)[p];
}

// dst (cols x rows, ldd) <- transpose of src (rows x cols, lds); rows, cols % 32 == 0.  32 x 8 threads per 32 x 32 tile;
// src may be peer memory (coalesced 256-byte reads over NVLink, local transposed writes).
__global__ void __launch_bounds__(256) copy2d_transpose_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                                               long long lds, long long row_tiles, long long col_tiles) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (long long lin = blockIdx.x; lin < row_tiles * col_tiles; lin += gridDim.x) {
    const long long br = lin % row_tiles, bc = lin / row_tiles;
    const double* S = src + br * 32 + bc * 32 * lds;
    double* D = dst + bc * 32 + br * 32 * ldd;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) t[ty + 8 * r][tx] = S[tx + (long long)(ty + 8 * r) * lds];   // t[c][r]
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) D[tx + (long long)(ty + 8 * r) * ldd] = t[tx][ty + 8 * r];    // dst(c = tx, r = ty + 8r)
  }
}

// Row panel k of the block-cyclic factor, gathered in global column order on the calling rank:
//   dst[p + ((J-k-1)*nb + c)*nb] = L_{owner(J)}[k*nb + p + ((J/G)*nb + c)*ld],  J = k+1+blockIdx.x
// (owner(J): snake order of DistLayout::owner)
__global__ void __launch_bounds__(256) gather_rowpanel_kernel(double* __restrict__ dst, const PeerPtrs src, int G, long long ld,
                                                              long long k, long long nb) {
  const long long J = k + 1 + blockIdx.x;
  const int pos = (int)(J % G);
  const double* S = src.p[((J / G) & 1) ? G - 1 - pos : pos] + k * nb + (J / G) * nb * ld;
  double* D = dst + (long long)blockIdx.x * nb * nb;
  const long long r2 = nb >> 1;
  for (long long c = blockIdx.y; c < nb; c += gridDim.y)
    for (long long p = threadIdx.x; p < r2; p += blockDim.x)
      reinterpret_cast<double2*>(D + c * nb)[p] = reinterpret_cast<const double2*>(S + c * ld)[p];
}

// Row panel pulled by the copy engines arrives rank-major in `stage` (source rank s: its blocks J > k in local order at
// block offset off[s]); put the blocks into global column order: dst block (J - k - 1) <- stage block off[s] + J/G - first[s].
struct GatherMap { int off[MGPU_MAX_RANKS]; int first[MGPU_MAX_RANKS]; };
__global__ void __launch_bounds__(256) unpermute_rowpanel_kernel(double* __restrict__ dst, const double* __restrict__ stage,
                                                                 const GatherMap gm, int G, long long k, long long nb) {
  const long long J = k + 1 + blockIdx.x;
  const int pos = (int)(J % G);
  const int s = ((J / G) & 1) ? G - 1 - pos : pos;
  const double2* S = reinterpret_cast<const double2*>(stage + ((long long)gm.off[s] + J / G - gm.first[s]) * nb * nb);
  double2* D = reinterpret_cast<double2*>(dst + (long long)blockIdx.x * nb * nb);
  const long long n2 = nb * nb / 2;
  for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.y * blockDim.x) D[i] = S[i];
}
// The same with every nb x nb block TRANSPOSED into a (Krem x nb, ld = ldt) panel: dst[(J-k-1)*nb + c + p*ldt] = block(p, c).
// Block blockIdx.x; 32 x 32 tiles over blockIdx.y (32 x 8 threads).
__global__ void __launch_bounds__(256) unpermute_rowpanel_t_kernel(double* __restrict__ dst, long long ldt, const double* __restrict__ stage,
                                                                   const GatherMap gm, int G, long long k, long long nb) {
  __shared__ double t[32][33];
  const long long J = k + 1 + blockIdx.x;
  const int pos = (int)(J % G);
  const int s = ((J / G) & 1) ? G - 1 - pos : pos;
  const double* S = stage + ((long long)gm.off[s] + J / G - gm.first[s]) * nb * nb;   // (p, c) at S[p + c*nb]
  double* D = dst + (long long)blockIdx.x * nb;                                          // (c, p) at D[c + p*ldt]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long nt = nb / 32;
  for (long long tile = blockIdx.y; tile < nt * nt; tile += gridDim.y) {
    const long long bp = tile % nt, bc = tile / nt;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) t[ty + 8 * q][tx] = S[(bp * 32 + tx) + (bc * 32 + ty + 8 * q) * nb];     // t[c][p]
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) D[(bc * 32 + tx) + (bp * 32 + ty + 8 * q) * ldt] = t[tx][ty + 8 * q];   // dst(c = tx, p = ty + 8q)
  }
}

// out[0] = sum over the local matrix columns of log(L[gcol(c), c]) (the rank's share of log det U); one CTA.
__global__ void __launch_bounds__(1024, 1) dist_logdiag_kernel(const double* __restrict__ L, long long ld, long long ncols,
                                                               long long nb, int G, int rank, double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (long long c = threadIdx.x; c < ncols; c += blockDim.x) {
    const long long lb = c / nb;
    const long long g = (lb * G + ((lb & 1) ? G - 1 - rank : rank)) * nb + c % nb;   // DistLayout::gblock
    s += log(L[g + c * ld]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// alpha[i] = -X[i] (the y columns hold -K^-1 y after the trtri sweep); out[1] = dot(y, alpha) over n.  One CTA.
__global__ void __launch_bounds__(1024, 1) dist_alpha_kernel(const double* __restrict__ X, const double* __restrict__ y,
                                                             long long n, long long np, double* __restrict__ alpha,
                                                             double* __restrict__ out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (long long i = threadIdx.x; i < np; i += blockDim.x) {
    const double a = (i < n) ? -X[i] : 0.0;
    alpha[i] = a;
    if (i < n) s += y[i] * a;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[1] = red[0];
}

// tot[s] = sum_b partial[b][s] in block order (deterministic)
__global__ void grad_partial_sum_kernel(const double* __restrict__ partial, int nblocks, int P, double* __restrict__ tot) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > P) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * (P + 1) + s];
  tot[s] = v;
}

struct MRank {
  int rank = 0;              // global rank of this entry (== its index unless transport 2)
  gpr_ctx* ctx = nullptr;
  CudaBE be{nullptr};
  double *L = nullptr, *dinv = nullptr, *Ukk[2] = {nullptr, nullptr}, *panel[2] = {nullptr, nullptr};
  double* stage = nullptr;   // landing buffer of the copy-engine pulls (Np * nb doubles)
  int* gtile = nullptr;
  // model part
  double *x = nullptr, *y = nullptr, *hp = nullptr, *alpha = nullptr, *gpart = nullptr, *scal = nullptr;   // scal: [logdiag, y.alpha, tot[P+1]...]
  int gr_blocks = 0;
  cudaEvent_t ev = nullptr;
  long long* d_flag = nullptr;   // transport 2: all-reduced factorization status
};

struct NcclApi;

}  // namespace

struct gpr_mgpu {
  int G = 0;                 // ranks of the distributed factorization (world size)
  int64_t nb = 1024;
  std::vector<MRank> rk;     // the ranks held by this process: all G (transports 0, 1) or one (transport 2)
  int transport = 0;
  NcclApi* nccl = nullptr;
  void* comm = nullptr;      // ncclComm_t (transport 2)
  std::string err;
  cudaEvent_t t0 = nullptr, t1 = nullptr, tot0 = nullptr, tot1 = nullptr;
  // how the panels of the NEXT step travel during trtri / lauum: 0 = on the main queue before the step's GEMMs (no
  // overlap), 1 = side queue, SM-driven peer reads, 2 = side queue, copy engines + local re-layout.
  // Measured at N = 131072 on 8 x B200 (profiles/README.md), phase time in ms:
  //   trtri  mode 0: 3760   mode 1: 3779   mode 2: 3688     lauum  mode 0: 3735   mode 1: 3552   mode 2: 3140
  // (at 2 GPUs the three modes are within 0.5 % of each other), hence the defaults.
  int prefetch_trtri = 2, prefetch_lauum = 2;
  struct gpr_mgpu_model* model = nullptr;   // the one live model of this context (released with it if still alive)
};

namespace {

int mfail(gpr_mgpu* mg, int code, const std::string& msg) {
  if (mg) mg->err = msg; else g_create_error = msg;
  return code;
}
#define MCK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) {                                                                               \
      char b_[512];                                                                                        \
      snprintf(b_, sizeof b_, "CUDA error %d (%s) at %s [mgpu_api.inl:%d]", (int)e_, cudaGetErrorString(e_), #call, __LINE__); \
      return mfail(mg, e_ == cudaErrorMemoryAllocation ? GPR_ERR_MEMORY : GPR_ERR_CUDA, b_);               \
    }                                                                                                      \
  } while (0)

// COMM interface of csrc/dist_blocked.hpp (virtual: the transport is chosen at run time) plus the two exchanges of the
// model level: the alpha vector from the y owner to everybody and the sum of a few doubles over the ranks.
struct CommBase {
  int dma_mode = 2;
  virtual ~CommBase() {}
  virtual void barrier() = 0;
  virtual void bcast_diag(int64_t k, bool with_owner, int b) = 0;
  virtual void gather_rowpanel(int64_t k, int b) = 0;
  virtual void gather_rowpanel_t(int64_t k, int b) = 0;
  virtual void bcast_colpanel(int64_t k, int b) = 0;
  virtual void bcast_alpha(int root, int64_t count) = 0;                     // MRank::alpha
  virtual int sum_scal(int64_t off, int64_t count, double* host_out) = 0;    // sum over ranks of MRank::scal[off ..], rank order
};

// NCCL through dlopen: libgpr_sm100a.so keeps linking the CUDA runtime only; inside a torch process the already loaded
// libnccl.so.2 (torch's bundled build) is the one that answers.
}  // namespace
#include <dlfcn.h>
#include <nccl.h>
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api(std::string& err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
      *(void**)&api.GetUniqueId = dlsym(api.lib, "ncclGetUniqueId");
      *(void**)&api.CommInitRank = dlsym(api.lib, "ncclCommInitRank");
      *(void**)&api.CommDestroy = dlsym(api.lib, "ncclCommDestroy");
      *(void**)&api.Broadcast = dlsym(api.lib, "ncclBroadcast");
      *(void**)&api.AllGather = dlsym(api.lib, "ncclAllGather");
      *(void**)&api.AllReduce = dlsym(api.lib, "ncclAllReduce");
      *(void**)&api.GetErrorString = dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib) { err = "libnccl.so.2 could not be loaded (dlopen)"; return nullptr; }
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllGather || !api.AllReduce) {
    err = "libnccl.so.2 lacks a required entry point";
    return nullptr;
  }
  return &api;
}

// sum over the local ranks of MRank::scal[off .. off + count), in rank order (deterministic), on the host
int host_sum_scal(gpr_mgpu* mg, int64_t off, int64_t count, double* host_out) {
  std::vector<double> part((size_t)count);
  for (int64_t i = 0; i < count; ++i) host_out[i] = 0.0;
  for (auto& R : mg->rk) {
    MCK(cudaSetDevice(R.ctx->device));
    MCK(cudaMemcpyAsync(part.data(), R.scal + off, sizeof(double) * count, cudaMemcpyDeviceToHost, R.ctx->stream));
    MCK(cudaStreamSynchronize(R.ctx->stream));
    for (int64_t i = 0; i < count; ++i) host_out[i] += part[i];
  }
  return GPR_OK;
}

// COMM for ranks that all live in this process (transport 0)
struct LocalComm : CommBase {
  gpr_mgpu* mg;
  DistLayout lay;
  int64_t ld;
  LocalComm(gpr_mgpu* m, const DistLayout& l, int64_t ld_) : mg(m), lay(l), ld(ld_) {}
  void act(int r) { cudaSetDevice(mg->rk[r].ctx->device); }
  void bcast_alpha(int root, int64_t count) override {
    barrier();
    for (int r = 0; r < mg->G; ++r)
      if (r != root) copy2d(r, mg->rk[r].alpha, count, mg->rk[root].alpha, count, count, 1);
  }
  int sum_scal(int64_t off, int64_t count, double* host_out) override { return host_sum_scal(mg, off, count, host_out); }
  void barrier() override {
    if (mg->G == 1) return;
    for (int r = 0; r < mg->G; ++r) { act(r); mg->rk[r].be.note(cudaEventRecord(mg->rk[r].ev, mg->rk[r].ctx->stream)); }
    for (int r = 0; r < mg->G; ++r) {
      act(r);
      for (int s = 0; s < mg->G; ++s)
        if (s != r) mg->rk[r].be.note(cudaStreamWaitEvent(mg->rk[r].ctx->stream, mg->rk[s].ev, 0));
    }
  }
  void copy2d(int r, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    MRank& R = mg->rk[r];
    act(r);
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((rows / 2 + 255) / 256, 64)), (unsigned)std::min<int64_t>(cols, 1024));
    copy2d_kernel<<<grid, 256, 0, R.ctx->stream>>>(dst, ldd, src, lds, rows, cols);
    R.be.note(cudaGetLastError());
    R.ctx->launches++;
  }
  // A rank that currently issues to its SIDE queue is prefetching while its main queue runs GEMMs that fill every SM
  // (2 CTAs x 250 registers): an SM-driven pull would have to take CTA slots away from them and hold them for the
  // length of an NVLink read (measured on 8 x B200: trtri + 300 ms).  Those transfers therefore go through the copy
  // engines (cudaMemcpy2DAsync between peers) into a staging buffer; only the cheap local re-layout is a kernel.
  bool on_side(int r) const { return dma_mode == 2 && mg->rk[r].ctx->side_stream && mg->rk[r].ctx->stream == mg->rk[r].ctx->side_stream; }
  void dma2d(int r, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    MRank& R = mg->rk[r];
    act(r);
    R.be.note(cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * rows, (size_t)cols,
                                cudaMemcpyDefault, R.ctx->stream));
  }
  void bcast_diag(int64_t k, bool with_owner, int b) override {
    const int o = lay.owner(k);
    const int64_t nb = lay.nb, kb = k / lay.G;
    const MRank& S = mg->rk[o];
    const int64_t dl = (int64_t)lay.tpb() * 128 * 128;
    for (int r = 0; r < mg->G; ++r) {
      if (r == o && !with_owner) continue;
      if (on_side(r)) {
        dma2d(r, mg->rk[r].Ukk[b], nb, S.L + k * nb + kb * nb * ld, ld, nb, nb);
        if (r != o) dma2d(r, mg->rk[r].dinv + k * dl, dl, S.dinv + k * dl, dl, dl, 1);
      } else {
        copy2d(r, mg->rk[r].Ukk[b], nb, S.L + k * nb + kb * nb * ld, ld, nb, nb);
        if (r != o) copy2d(r, mg->rk[r].dinv + k * dl, dl, S.dinv + k * dl, dl, dl, 1);
      }
    }
  }
  void gather_rowpanel(int64_t k, int b) override { gather_impl(k, b, false); }
  void gather_rowpanel_t(int64_t k, int b) override { gather_impl(k, b, true); }
  void gather_impl(int64_t k, int b, bool transposed) {
    const int64_t nrem = lay.nblk - k - 1;
    if (nrem <= 0) return;
    const int64_t nb = lay.nb, Krem = nrem * nb;
    PeerPtrs pp{};
    for (int s = 0; s < mg->G; ++s) pp.p[s] = mg->rk[s].L;
    GatherMap gm{};
    int off = 0;
    for (int s = 0; s < mg->G; ++s) {
      gm.first[s] = (int)lay.count_le(s, k);
      gm.off[s] = off;
      off += (int)(lay.nloc(s) - lay.count_le(s, k));
    }
    for (int r = 0; r < mg->G; ++r) {
      MRank& R = mg->rk[r];
      act(r);
      if (on_side(r)) {   // copy engines -> staging (rank-major), then a local re-layout
        for (int s = 0; s < mg->G; ++s) {
          const int64_t cnt = lay.nloc(s) - gm.first[s];
          if (cnt > 0) dma2d(r, R.stage + (int64_t)gm.off[s] * nb * nb, nb, mg->rk[s].L + k * nb + (int64_t)gm.first[s] * nb * ld, ld, nb, cnt * nb);
        }
        if (transposed) {
          dim3 grid((unsigned)nrem, (unsigned)std::max<int64_t>(1, std::min<int64_t>((nb / 32) * (nb / 32), 2048 / nrem)));
          unpermute_rowpanel_t_kernel<<<grid, 256, 0, R.ctx->stream>>>(R.panel[b], Krem, R.stage, gm, mg->G, k, nb);
        } else {
          dim3 grid((unsigned)nrem, (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, 1024 / nrem)));
          unpermute_rowpanel_kernel<<<grid, 256, 0, R.ctx->stream>>>(R.panel[b], R.stage, gm, mg->G, k, nb);
        }
      } else {            // SM-driven peer reads straight into global column order (+ a local transpose)
        dim3 grid((unsigned)nrem, (unsigned)std::min<int64_t>(nb, std::max<int64_t>(4, 2048 / nrem)));
        gather_rowpanel_kernel<<<grid, 256, 0, R.ctx->stream>>>(transposed ? R.stage : R.panel[b], pp, mg->G, ld, k, nb);
        if (transposed) {
          R.be.note(cudaGetLastError());
          R.ctx->launches++;
          const int64_t tiles = (nb / 32) * (Krem / 32);
          copy2d_transpose_kernel<<<(unsigned)std::min<int64_t>(tiles, 148 * 16), dim3(32, 8), 0, R.ctx->stream>>>(
              R.panel[b], Krem, R.stage, nb, nb / 32, Krem / 32);
        }
      }
      R.be.note(cudaGetLastError());
      R.ctx->launches++;
    }
  }
  void bcast_colpanel(int64_t k, int b) override {   // transposed: panel[b] is nb x (k+1)*nb, ld nb
    const MRank& S = mg->rk[lay.owner(k)];
    const int64_t rows = (k + 1) * lay.nb, cols = lay.nb;
    const double* src = S.L + (k / lay.G) * lay.nb * ld;
    for (int r = 0; r < mg->G; ++r) {
      MRank& R = mg->rk[r];
      act(r);
      const double* tsrc = src;
      int64_t tld = ld;
      if (on_side(r)) {   // copy engine -> staging (rows x cols, ld rows), then a local transpose
        dma2d(r, R.stage, rows, src, ld, rows, cols);
        tsrc = R.stage; tld = rows;
      }
      const int64_t tiles = (rows / 32) * (cols / 32);
      copy2d_transpose_kernel<<<(unsigned)std::min<int64_t>(tiles, 148 * 16), dim3(32, 8), 0, R.ctx->stream>>>(
          R.panel[b], cols, tsrc, tld, rows / 32, cols / 32);
      R.be.note(cudaGetLastError());
      R.ctx->launches++;
    }
  }
};

// COMM through pack -> collective -> unpack (transports 1 and 2).  Everything is issued on the main queue: NCCL wants
// one consistent order of collectives per communicator, so the panels are not prefetched here.
struct PackedComm : CommBase {
  gpr_mgpu* mg;
  DistLayout lay;
  int64_t ld;
  PackedComm(gpr_mgpu* m, const DistLayout& l, int64_t ld_) : mg(m), lay(l), ld(ld_) {}
  MRank* local(int rank) { for (auto& R : mg->rk) if (R.rank == rank) return &R; return nullptr; }
  void sync_all() { for (auto& R : mg->rk) { cudaSetDevice(R.ctx->device); R.be.note(cudaStreamSynchronize(R.ctx->stream)); } }
  void nccl_note(MRank& R, ncclResult_t rc) { if (rc != ncclSuccess) R.be.note(cudaErrorUnknown); }
  void copy2d(MRank& R, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    cudaSetDevice(R.ctx->device);
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((rows / 2 + 255) / 256, 64)), (unsigned)std::min<int64_t>(cols, 1024));
    copy2d_kernel<<<grid, 256, 0, R.ctx->stream>>>(dst, ldd, src, lds, rows, cols);
    R.be.note(cudaGetLastError());
    R.ctx->launches++;
  }
  // buf(R): the same logical buffer on every rank
  template <class F> void bcast(int root, F buf, size_t count) {
    if (mg->transport == 2) {
      for (auto& R : mg->rk) {
        cudaSetDevice(R.ctx->device);
        nccl_note(R, mg->nccl->Broadcast(buf(R), buf(R), count, ncclDouble, root, (ncclComm_t)mg->comm, R.ctx->stream));
      }
      return;
    }
    sync_all();
    MRank* S = local(root);
    for (auto& R : mg->rk)
      if (&R != S) { cudaSetDevice(R.ctx->device); R.be.note(cudaMemcpyAsync(buf(R), buf(*S), sizeof(double) * count, cudaMemcpyDefault, R.ctx->stream)); }
    sync_all();
  }
  // in place: rank r's contribution sits at stage + r * slot on rank r
  void allgather_stage(size_t slot) {
    if (mg->transport == 2) {
      for (auto& R : mg->rk) {
        cudaSetDevice(R.ctx->device);
        nccl_note(R, mg->nccl->AllGather(R.stage + (size_t)R.rank * slot, R.stage, slot, ncclDouble, (ncclComm_t)mg->comm, R.ctx->stream));
      }
      return;
    }
    sync_all();
    for (auto& D : mg->rk)
      for (auto& S : mg->rk)
        if (&D != &S) {
          cudaSetDevice(D.ctx->device);
          D.be.note(cudaMemcpyAsync(D.stage + (size_t)S.rank * slot, S.stage + (size_t)S.rank * slot, sizeof(double) * slot, cudaMemcpyDefault, D.ctx->stream));
        }
    sync_all();
  }
  void barrier() override {}
  void bcast_diag(int64_t k, bool, int b) override {
    const int o = lay.owner(k);
    const int64_t nb = lay.nb, kb = k / lay.G, dl = (int64_t)lay.tpb() * 128 * 128;
    if (MRank* S = local(o)) copy2d(*S, S->Ukk[b], nb, S->L + k * nb + kb * nb * ld, ld, nb, nb);   // pack
    bcast(o, [b](MRank& R) { return R.Ukk[b]; }, (size_t)(nb * nb));
    bcast(o, [k, dl](MRank& R) { return R.dinv + k * dl; }, (size_t)dl);
  }
  void gather_rowpanel(int64_t k, int b) override { gather_impl(k, b, false); }
  void gather_rowpanel_t(int64_t k, int b) override { gather_impl(k, b, true); }
  void gather_impl(int64_t k, int b, bool transposed) {
    const int64_t nrem = lay.nblk - k - 1;
    if (nrem <= 0) return;
    const int64_t nb = lay.nb;
    int64_t maxcnt = 0;
    GatherMap gm{};
    for (int s = 0; s < mg->G; ++s) {
      gm.first[s] = (int)lay.count_le(s, k);
      maxcnt = std::max<int64_t>(maxcnt, lay.nloc(s) - gm.first[s]);
    }
    for (int s = 0; s < mg->G; ++s) gm.off[s] = (int)(s * maxcnt);
    const size_t slot = (size_t)(maxcnt * nb * nb);
    for (auto& R : mg->rk) {   // pack this rank's blocks of the row panel
      const int64_t first = gm.first[R.rank], cnt = lay.nloc(R.rank) - first;
      if (cnt > 0) copy2d(R, R.stage + (size_t)R.rank * slot, nb, R.L + k * nb + first * nb * ld, ld, nb, cnt * nb);
    }
    allgather_stage(slot);
    for (auto& R : mg->rk) {   // rank-major -> global column order (optionally transposed)
      cudaSetDevice(R.ctx->device);
      if (transposed) {
        dim3 grid((unsigned)nrem, (unsigned)std::max<int64_t>(1, std::min<int64_t>((nb / 32) * (nb / 32), 2048 / nrem)));
        unpermute_rowpanel_t_kernel<<<grid, 256, 0, R.ctx->stream>>>(R.panel[b], nrem * nb, R.stage, gm, mg->G, k, nb);
      } else {
        dim3 grid((unsigned)nrem, (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, 1024 / nrem)));
        unpermute_rowpanel_kernel<<<grid, 256, 0, R.ctx->stream>>>(R.panel[b], R.stage, gm, mg->G, k, nb);
      }
      R.be.note(cudaGetLastError());
      R.ctx->launches++;
    }
  }
  void bcast_colpanel(int64_t k, int b) override {
    const int o = lay.owner(k);
    const int64_t rows = (k + 1) * lay.nb, cols = lay.nb;
    if (MRank* S = local(o)) copy2d(*S, S->stage, rows, S->L + (k / lay.G) * lay.nb * ld, ld, rows, cols);   // pack
    bcast(o, [](MRank& R) { return R.stage; }, (size_t)(rows * cols));
    for (auto& R : mg->rk) {
      cudaSetDevice(R.ctx->device);
      const int64_t tiles = (rows / 32) * (cols / 32);
      copy2d_transpose_kernel<<<(unsigned)std::min<int64_t>(tiles, 148 * 16), dim3(32, 8), 0, R.ctx->stream>>>(
          R.panel[b], cols, R.stage, rows, rows / 32, cols / 32);
      R.be.note(cudaGetLastError());
      R.ctx->launches++;
    }
  }
  void bcast_alpha(int root, int64_t count) override { bcast(root, [](MRank& R) { return R.alpha; }, (size_t)count); }
  int sum_scal(int64_t off, int64_t count, double* host_out) override {
    if (mg->transport == 2) {
      for (auto& R : mg->rk) {
        cudaSetDevice(R.ctx->device);
        nccl_note(R, mg->nccl->AllReduce(R.scal + off, R.scal + off, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)mg->comm, R.ctx->stream));
      }
    }
    return host_sum_scal(mg, off, count, host_out);   // transport 2: one local rank holding the global sum
  }
};

// dense state of one distributed factorization
struct MDense {
  DistLayout lay;
  int64_t ld = 0;
};

void mdense_free(gpr_mgpu* mg) {
  for (auto& R : mg->rk) {
    if (!R.ctx) continue;
    cudaSetDevice(R.ctx->device);
    cudaStreamSynchronize(R.ctx->stream);
    cudaFree(R.L); cudaFree(R.dinv); cudaFree(R.gtile); cudaFree(R.stage); R.stage = nullptr;
    for (int b = 0; b < 2; ++b) { cudaFree(R.Ukk[b]); cudaFree(R.panel[b]); R.Ukk[b] = R.panel[b] = nullptr; }
    cudaFree(R.x); cudaFree(R.y); cudaFree(R.hp); cudaFree(R.alpha); cudaFree(R.gpart); cudaFree(R.scal); cudaFree(R.d_flag); R.d_flag = nullptr;
    R.L = R.dinv = R.x = R.y = R.hp = R.alpha = R.gpart = R.scal = nullptr;
    R.gtile = nullptr;
  }
}

int mdense_alloc(gpr_mgpu* mg, MDense& md, int64_t Np, int64_t nyp) {
  DistLayout& lay = md.lay;
  lay.G = mg->G; lay.nb = mg->nb; lay.Np = Np; lay.nblk = Np / mg->nb; lay.nyp = nyp;
  md.ld = Np;
  for (auto& R : mg->rk) {
    const int r = R.rank;
    MCK(cudaSetDevice(R.ctx->device));
    const int64_t lc = std::max<int64_t>(lay.lcols(r), 128);
    MCK(cudaMalloc(&R.L, sizeof(double) * Np * lc));
    MCK(cudaMalloc(&R.dinv, sizeof(double) * Np * 128));
    for (int b = 0; b < 2; ++b) {
      MCK(cudaMalloc(&R.Ukk[b], sizeof(double) * lay.nb * lay.nb));
      MCK(cudaMalloc(&R.panel[b], sizeof(double) * Np * lay.nb));
    }
    MCK(cudaMalloc(&R.stage, sizeof(double) * (Np + (int64_t)mg->G * lay.nb) * lay.nb));   // + one block per rank: padded all-gather slots
    MCK(cudaMalloc(&R.d_flag, sizeof(long long)));
    const int64_t lt = std::max<int64_t>(lay.ltiles(r), 1);
    MCK(cudaMalloc(&R.gtile, sizeof(int) * lt));
    std::vector<int> gt((size_t)lt, 0);
    for (int64_t t = 0; t < lay.ltiles(r); ++t) gt[t] = lay.gtile(r, t);
    MCK(cudaMemcpyAsync(R.gtile, gt.data(), sizeof(int) * lt, cudaMemcpyHostToDevice, R.ctx->stream));
    MCK(cudaMemsetAsync(R.ctx->d_info, 0, sizeof(long long), R.ctx->stream));
    MCK(cudaStreamSynchronize(R.ctx->stream));
  }
  return GPR_OK;
}

std::vector<DistRank<CudaBE>> mdense_ranks(gpr_mgpu* mg, const MDense& md) {
  std::vector<DistRank<CudaBE>> v;
  for (auto& R : mg->rk)
    v.push_back(DistRank<CudaBE>{R.rank, &R.be, R.L, md.ld, R.dinv, {R.Ukk[0], R.Ukk[1]}, {R.panel[0], R.panel[1]}, R.gtile});
  return v;
}

int msync_all(gpr_mgpu* mg, const char* where, long long* info) {
  long long first = 0;
  for (auto& R : mg->rk) {
    MCK(cudaSetDevice(R.ctx->device));
    long long h = 0;
    MCK(cudaMemcpyAsync(&h, R.ctx->d_info, sizeof h, cudaMemcpyDeviceToHost, R.ctx->stream));
    MCK(cudaStreamSynchronize(R.ctx->stream));
    if (R.ctx->pending != cudaSuccess) {
      cudaError_t e = R.ctx->pending; R.ctx->pending = cudaSuccess;
      char b[256];
      snprintf(b, sizeof b, "CUDA / NCCL error %d (%s) in %s on rank %d", (int)e, cudaGetErrorString(e), where, R.rank);
      return mfail(mg, GPR_ERR_CUDA, b);
    }
    if (h && (!first || h < first)) first = h;
  }
  if (mg->transport == 2 && info) {
    // every process must take the same branch on a failed factorization: smallest failing pivot over all ranks
    // (max of 2^62 - pivot; 0 = no failure)
    MRank& R = mg->rk[0];
    long long enc = first ? (1LL << 62) - first : 0;
    MCK(cudaMemcpyAsync(R.d_flag, &enc, sizeof enc, cudaMemcpyHostToDevice, R.ctx->stream));
    if (mg->nccl->AllReduce(R.d_flag, R.d_flag, 1, ncclInt64, ncclMax, (ncclComm_t)mg->comm, R.ctx->stream) != ncclSuccess)
      return mfail(mg, GPR_ERR_CUDA, "ncclAllReduce(status) failed");
    MCK(cudaMemcpyAsync(&enc, R.d_flag, sizeof enc, cudaMemcpyDeviceToHost, R.ctx->stream));
    MCK(cudaStreamSynchronize(R.ctx->stream));
    first = enc ? (1LL << 62) - enc : 0;
  }
  if (info) *info = first;
  return GPR_OK;
}

std::unique_ptr<CommBase> make_comm(gpr_mgpu* mg, const DistLayout& lay, int64_t ld) {
  if (mg->transport == 0) return std::unique_ptr<CommBase>(new LocalComm(mg, lay, ld));
  return std::unique_ptr<CommBase>(new PackedComm(mg, lay, ld));
}

}  // namespace

struct gpr_mgpu_model {
  gpr_mgpu* mg = nullptr;
  KSpec spec;
  int ncomp = 0, D = 0, P = 0, nk = 0, ny = 1, train_axis = 1;
  int64_t N = 0, Np = 0, nyp = 128;
  MDense md;
  bool have_inverse = false, have_alpha = false;
  double ms[GPR_T_COUNT] = {0};
};

extern "C" {

int gpr_mgpu_create(int nranks, const int* devices, int64_t nb, gpr_mgpu** out) {
  gpr_mgpu* mg = nullptr;
  if (!out || !devices) return mfail(nullptr, GPR_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (nranks < 1 || nranks > MGPU_MAX_RANKS) return mfail(nullptr, GPR_ERR_ARG, "number of ranks must be in 1..16");
  if (nb < 128 || nb % 128) return mfail(nullptr, GPR_ERR_ARG, "panel width nb must be a positive multiple of 128");
  mg = new gpr_mgpu();
  live_add(mg);
  mg->G = nranks; mg->nb = nb;
  mg->rk.resize(nranks);
  for (int r = 0; r < nranks; ++r) {
    int rc = gpr_ctx_create(devices[r], &mg->rk[r].ctx);
    if (rc) { std::string e = g_create_error; gpr_mgpu_destroy(mg); g_create_error = e; return rc; }
    mg->rk[r].be = CudaBE{mg->rk[r].ctx};
    mg->rk[r].rank = r;
    if (cudaEventCreateWithFlags(&mg->rk[r].ev, cudaEventDisableTiming) != cudaSuccess) {
      gpr_mgpu_destroy(mg);
      return mfail(nullptr, GPR_ERR_CUDA, "cudaEventCreate failed");
    }
  }
  // peer access between every pair of distinct devices (NVLink / NVSwitch on the 8 x B200 box)
  for (int a = 0; a < nranks; ++a)
    for (int b = 0; b < nranks; ++b) {
      const int da = devices[a], db = devices[b];
      if (da == db) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, da, db);
      if (!can) { gpr_mgpu_destroy(mg); return mfail(nullptr, GPR_ERR_UNSUPPORTED, "devices " + std::to_string(da) + " and " + std::to_string(db) + " have no peer access"); }
      cudaSetDevice(da);
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { gpr_mgpu_destroy(mg); return mfail(nullptr, GPR_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
    }
  cudaSetDevice(devices[0]);
  cudaEventCreate(&mg->t0); cudaEventCreate(&mg->t1);
  cudaEventCreate(&mg->tot0); cudaEventCreate(&mg->tot1);
  *out = mg;
  return GPR_OK;
}

int gpr_dist_unique_id(void* id128) {
  if (!id128) return GPR_ERR_ARG;
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) return mfail(nullptr, GPR_ERR_UNSUPPORTED, err);
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (api->GetUniqueId(&id) != ncclSuccess) return mfail(nullptr, GPR_ERR_CUDA, "ncclGetUniqueId failed");
  memcpy(id128, &id, sizeof id);
  return GPR_OK;
}

int gpr_dist_create(int device, int rank, int world, const void* id128, int64_t nb, gpr_mgpu** out) {
  if (!out || !id128) return mfail(nullptr, GPR_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (world < 1 || world > MGPU_MAX_RANKS || rank < 0 || rank >= world) return mfail(nullptr, GPR_ERR_ARG, "rank / world out of range (world <= 16)");
  if (nb < 128 || nb % 128) return mfail(nullptr, GPR_ERR_ARG, "panel width nb must be a positive multiple of 128");
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) return mfail(nullptr, GPR_ERR_UNSUPPORTED, err);
  gpr_mgpu* mg = new gpr_mgpu();
  live_add(mg);
  mg->G = world; mg->nb = nb; mg->transport = 2; mg->nccl = api;
  mg->prefetch_trtri = mg->prefetch_lauum = 0;
  mg->rk.resize(1);
  int rc = gpr_ctx_create(device, &mg->rk[0].ctx);
  if (rc) { std::string e = g_create_error; gpr_mgpu_destroy(mg); g_create_error = e; return rc; }
  mg->rk[0].be = CudaBE{mg->rk[0].ctx};
  mg->rk[0].rank = rank;
  cudaEventCreateWithFlags(&mg->rk[0].ev, cudaEventDisableTiming);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm = nullptr;
  ncclResult_t nr = api->CommInitRank(&comm, world, id, rank);
  if (nr != ncclSuccess) {
    std::string e = std::string("ncclCommInitRank failed: ") + (api->GetErrorString ? api->GetErrorString(nr) : "?");
    gpr_mgpu_destroy(mg);
    return mfail(nullptr, GPR_ERR_CUDA, e);
  }
  mg->comm = comm;
  cudaEventCreate(&mg->t0); cudaEventCreate(&mg->t1);
  cudaEventCreate(&mg->tot0); cudaEventCreate(&mg->tot1);
  *out = mg;
  return GPR_OK;
}

int gpr_mgpu_destroy(gpr_mgpu* mg) {
  if (!mg || !live_take(mg)) return GPR_OK;
  if (mg->model) { gpr_mgpu_model* mm = mg->model; mg->model = nullptr; if (live_take(mm)) delete mm; }
  mdense_free(mg);
  if (mg->comm && mg->nccl) { cudaSetDevice(mg->rk[0].ctx->device); mg->nccl->CommDestroy((ncclComm_t)mg->comm); mg->comm = nullptr; }
  for (auto& R : mg->rk) {
    if (!R.ctx) continue;
    cudaSetDevice(R.ctx->device);
    if (R.ev) cudaEventDestroy(R.ev);
    gpr_ctx_destroy(R.ctx);
  }
  if (mg->t0) cudaEventDestroy(mg->t0);
  if (mg->t1) cudaEventDestroy(mg->t1);
  if (mg->tot0) cudaEventDestroy(mg->tot0);
  if (mg->tot1) cudaEventDestroy(mg->tot1);
  delete mg;
  return GPR_OK;
}

int gpr_mgpu_set_option(gpr_mgpu* mg, const char* name, int64_t value) {
  if (!mg || !name) return GPR_ERR_ARG;
  if (!strcmp(name, "ozaki") || !strcmp(name, "ozaki_min") || !strcmp(name, "ozaki_phases") || !strcmp(name, "ozaki_kchunk") ||
      !strcmp(name, "ozaki_lauum") || !strcmp(name, "ozaki_lauum_map")) {   // INT8 route of the tile-mapped products
    for (auto& R : mg->rk)
      if (R.ctx && gpr_ctx_set_option(R.ctx, name, value) != GPR_OK) return mfail(mg, GPR_ERR_ARG, R.ctx->err);
    return GPR_OK;
  }
  if (value < 0 || value > 2) return mfail(mg, GPR_ERR_ARG, "option value must be 0, 1 or 2");
  if (!strcmp(name, "gemm_tma") || !strcmp(name, "kbuild_gram")) {   // forwarded to every rank's context
    for (auto& R : mg->rk) if (R.ctx) gpr_ctx_set_option(R.ctx, name, value);
    return GPR_OK;
  }
  if (!strcmp(name, "transport")) {
    if (mg->transport == 2 || value == 2) return mfail(mg, GPR_ERR_ARG, "transport 2 (NCCL, one process per rank) is chosen by gpr_dist_create only");
    if (mg->rk[0].L) return mfail(mg, GPR_ERR_STATE, "set the transport before creating a model");
    mg->transport = (int)value;
    if (value != 0) mg->prefetch_trtri = mg->prefetch_lauum = 0;
    return GPR_OK;
  }
  if (mg->transport != 0 && value != 0) return mfail(mg, GPR_ERR_UNSUPPORTED, "the packed transports issue every collective on the main queue (no prefetch)");
  if (!strcmp(name, "prefetch_trtri")) { mg->prefetch_trtri = (int)value; return GPR_OK; }
  if (!strcmp(name, "prefetch_lauum")) { mg->prefetch_lauum = (int)value; return GPR_OK; }
  return mfail(mg, GPR_ERR_ARG, std::string("unknown option ") + name);
}

const char* gpr_mgpu_last_error(gpr_mgpu* mg) { return mg ? mg->err.c_str() : g_create_error.c_str(); }

int64_t gpr_mgpu_launch_count(gpr_mgpu* mg) {
  int64_t n = 0;
  if (mg) for (auto& R : mg->rk) if (R.ctx) n += R.ctx->launches;
  return n;
}

int gpr_mgpu_model_create(gpr_mgpu* mg, const int* comp_types, int ncomp, int D, int64_t N, const double* x, const double* y,
                          int ny, int train_axis, gpr_mgpu_model** out) {
  if (!mg) return GPR_ERR_ARG;
  if (!out || !x || !y) return mfail(mg, GPR_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (N < 1) return mfail(mg, GPR_ERR_ARG, "x and y size mismatch.");
  if (ny < 1 || train_axis < 1 || train_axis > ny) return mfail(mg, GPR_ERR_ARG, "train_axis out of range");
  if (mg->rk[0].L) return mfail(mg, GPR_ERR_STATE, "this multi-GPU context already holds a model");
  gpr_mgpu_model* m = new gpr_mgpu_model();
  m->mg = mg;
  int rc = make_spec(mg->rk[0].ctx, comp_types, ncomp, D, &m->spec, &m->P, &m->nk);
  if (rc) { mg->err = mg->rk[0].ctx->err; delete m; return rc; }
  if (m->P > 90) { delete m; return mfail(mg, GPR_ERR_UNSUPPORTED, "more than 90 hyper-parameters are not supported"); }
  m->ncomp = ncomp; m->D = D; m->N = N; m->ny = ny; m->train_axis = train_axis;
  m->Np = round_up(N, mg->nb); m->nyp = round_up(ny, 128);
  rc = mdense_alloc(mg, m->md, m->Np, m->nyp);
  if (rc) { mdense_free(mg); delete m; return rc; }
  for (auto& R : mg->rk) {
    cudaError_t e = cudaSetDevice(R.ctx->device);
    auto A = [&](double** p, size_t elems) { if (e == cudaSuccess) e = cudaMalloc(p, elems * sizeof(double)); };
    A(&R.x, (size_t)D * N);
    A(&R.y, (size_t)N * ny);
    A(&R.hp, (size_t)m->P);
    A(&R.alpha, (size_t)m->Np);
    A(&R.scal, (size_t)(m->P + 3));
    R.gr_blocks = R.ctx->sm_count * 4;
    A(&R.gpart, (size_t)R.gr_blocks * (m->P + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(R.x, x, sizeof(double) * D * N, cudaMemcpyHostToDevice, R.ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(R.y, y, sizeof(double) * N * ny, cudaMemcpyHostToDevice, R.ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(R.ctx->stream);
    if (e != cudaSuccess) {
      mdense_free(mg); delete m;
      return mfail(mg, e == cudaErrorMemoryAllocation ? GPR_ERR_MEMORY : GPR_ERR_CUDA, std::string("model allocation / upload: ") + cudaGetErrorString(e));
    }
  }
  live_add(m);
  mg->model = m;
  *out = m;
  return GPR_OK;
}

int gpr_mgpu_model_destroy(gpr_mgpu_model* m) {
  if (!m || !live_take(m)) return GPR_OK;    // already released together with its gpr_mgpu
  m->mg->model = nullptr;
  mdense_free(m->mg);
  delete m;
  return GPR_OK;
}

/* loss_grad! / log_loss_grad! (src/cost.jl:50-70) with K block-cyclic over the ranks */
int gpr_mgpu_nlml_grad(gpr_mgpu_model* m, const double* hp_in, int P, int log_scale, double eps, double* F, double* G,
                       int64_t* info) {
  if (!m) return GPR_ERR_ARG;
  gpr_mgpu* mg = m->mg;
  if (!hp_in) return mfail(mg, GPR_ERR_ARG, "hp is NULL");
  if (P != m->P) return mfail(mg, GPR_ERR_ARG, "Parameter size mismatch.");
  std::vector<double> hp(hp_in, hp_in + P);
  if (log_scale) for (auto& v : hp) v = std::exp(v);
  if (info) *info = 0;
  const DistLayout& lay = m->md.lay;
  const int64_t N = m->N, Np = m->Np, nb = lay.nb, ld = m->md.ld;
  const int Gn = mg->G;
  for (int i = 0; i < GPR_T_COUNT; ++i) m->ms[i] = 0.0;
  auto tick = [&]() { cudaSetDevice(mg->rk[0].ctx->device); cudaEventRecord(mg->t0, mg->rk[0].ctx->stream); };
  auto tock = [&](int slot) -> int {
    // phase time = rank 0's stream from tick to the point where every rank has finished the phase
    for (auto& R : mg->rk) { cudaSetDevice(R.ctx->device); MCK(cudaStreamSynchronize(R.ctx->stream)); }
    cudaSetDevice(mg->rk[0].ctx->device);
    MCK(cudaEventRecord(mg->t1, mg->rk[0].ctx->stream));
    MCK(cudaEventSynchronize(mg->t1));
    float t = 0.f; cudaEventElapsedTime(&t, mg->t0, mg->t1);
    m->ms[slot] += t;
    return GPR_OK;
  };
  cudaEvent_t tot0 = mg->tot0, tot1 = mg->tot1;
  cudaSetDevice(mg->rk[0].ctx->device);
  cudaEventRecord(tot0, mg->rk[0].ctx->stream);

  // ---- covariance build into the block-cyclic layout (upper triangle, zero below), y columns
  tick();
  for (auto& R : mg->rk) {
    const int r = R.rank;
    gpr_ctx* ctx = R.ctx;
    MCK(cudaSetDevice(ctx->device));
    MCK(cudaMemcpyAsync(R.hp, hp.data(), sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    MCK(cudaMemsetAsync(ctx->d_info, 0, sizeof(long long), ctx->stream));
    MCK(cudaMemsetAsync(R.scal, 0, sizeof(double) * (P + 3), ctx->stream));
    for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
      const int64_t J = lay.gblock(r, lb);
      const int64_t cvalid = std::max<int64_t>(0, std::min<int64_t>(nb, N - J * nb));
      KBuildArgs a{};
      a.out = R.L + lb * nb * ld; a.ldo = ld; a.R = N; a.C = cvalid; a.Rp = Np; a.Cp = nb;
      a.x1 = R.x; a.x2 = R.x + (cvalid > 0 ? J * nb * m->D : 0); a.D = m->D; a.hp = R.hp; a.spec = m->spec;
      a.eps = eps; a.same = 1; a.add_noise = 1; a.pad_identity = 1; a.sigma_one = 0; a.row_scale = nullptr;
      a.diag_shift = J * nb; a.zero_lower = 1;
      int rc = launch_kbuild(ctx, DM_EUCLID, a, R.x);
      if (rc) return mfail(mg, rc, ctx->err);
    }
    if (r == lay.y_owner()) {
      const int64_t total = Np * m->nyp;
      pad_copy_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 65535), 256, 0, ctx->stream>>>(
          R.L + lay.ycol0(r) * ld, ld, Np, m->nyp, R.y, N, N, m->ny);
      ctx->launches++;
      MCK(cudaGetLastError());
    }
  }
  { int rc = tock(GPR_T_KBUILD); if (rc) return rc; }

  auto ranks = mdense_ranks(mg, m->md);
  std::unique_ptr<CommBase> comm_p = make_comm(mg, lay, ld);
  CommBase& comm = *comm_p;
  DistBlocked<CudaBE, CommBase> db(lay, ranks, comm);
  db.prefetch_trtri = mg->prefetch_trtri != 0;
  db.prefetch_lauum = mg->prefetch_lauum != 0;

  // INT8-tensor-core route of the rank-nb trailing updates of potrf (same automatic rule as the single-GPU path)
  {
    double sig2 = 0.0, noise2 = 0.0;
    bool have_noise = false;
    for (int c = 0; c < m->ncomp; ++c) {
      const double v = hp[m->spec.hp_off[c]];
      if (m->spec.type[c] == KT_NOISE) { if (!have_noise) { noise2 = v * v; have_noise = true; } }
      else sig2 += v * v;
    }
    const double floor_ = noise2 + m->nk * eps;
    const int auto_digits = (floor_ > 0.0 && ((double)N * sig2 + floor_) / floor_ <= 1e8) ? 8 : 0;
    for (auto& R : mg->rk) R.ctx->oz_active = R.ctx->ozaki >= 0 ? R.ctx->ozaki : auto_digits;
  }
  auto set_phase = [&](int ph) { for (auto& R : mg->rk) R.ctx->oz_cur = ph; };

  // ---- potrf (+ forward substitution of y), log det
  tick();
  set_phase(1);
  db.potrf();
  set_phase(8);
  for (auto& R : mg->rk) {
    const int r = R.rank;
    MCK(cudaSetDevice(R.ctx->device));
    dist_logdiag_kernel<<<1, 1024, 0, R.ctx->stream>>>(R.L, ld, lay.nloc(r) * nb, nb, Gn, r, R.scal);
    R.ctx->launches++;
    MCK(cudaGetLastError());
  }
  { int rc = tock(GPR_T_POTRF); if (rc) return rc; }
  long long h_info = 0;
  { int rc = msync_all(mg, "potrf", &h_info); if (rc) return rc; }
  if (info) *info = h_info;
  m->have_inverse = false;
  if (h_info != 0) {
    char buf[128];
    snprintf(buf, sizeof buf, "matrix is not positive definite; Cholesky failed at pivot %lld", h_info);
    return mfail(mg, GPR_ERR_NOT_POSDEF, buf);
  }

  const int yo = lay.y_owner();
  MRank* Ry = nullptr;                       // the local rank that owns the y columns (if it lives in this process)
  for (auto& R : mg->rk) if (R.rank == yo) Ry = &R;
  if (!G) {
    // loss only: y^T K^-1 y = |z|^2 with z = U^-T y, which the potrf sweep left in the y columns -- no back
    // substitution and no inverse needed (alpha is not formed: gpr_mgpu_fetch(ALPHA) then reports a state error)
    tick();
    if (Ry) {
      MCK(cudaSetDevice(Ry->ctx->device));
      sumsq_kernel<<<1, 1024, 0, Ry->ctx->stream>>>(Ry->L + (lay.ycol0(yo) + (m->train_axis - 1)) * ld, N, Ry->scal + 1);
      Ry->ctx->launches++;
      MCK(cudaGetLastError());
    }
    m->have_alpha = false;
    { int rc = tock(GPR_T_POTRS); if (rc) return rc; }
  } else {
    // ---- trtri (+ back substitution): W = U^-1 in place, y columns = -alpha
    tick();
    comm.dma_mode = mg->prefetch_trtri;
    set_phase(2);
    db.trtri();
    set_phase(8);
    { int rc = tock(GPR_T_TRTRI); if (rc) return rc; }
    tick();
    if (Ry) {
      MCK(cudaSetDevice(Ry->ctx->device));
      const double* X = Ry->L + (lay.ycol0(yo) + (m->train_axis - 1)) * ld;
      dist_alpha_kernel<<<1, 1024, 0, Ry->ctx->stream>>>(X, Ry->y + (int64_t)(m->train_axis - 1) * N, N, Np, Ry->alpha, Ry->scal);
      Ry->ctx->launches++;
      MCK(cudaGetLastError());
    }
    comm.bcast_alpha(yo, Np);
    m->have_alpha = true;
    { int rc = tock(GPR_T_POTRS); if (rc) return rc; }
  }

  double Fv = 0.0;
  {
    double sc[2];   // [sum of the ranks' log-diagonal shares, y . alpha (non-zero on the y owner only)]
    int rc = comm.sum_scal(0, 2, sc);
    if (rc) return rc;
    Fv = 0.5 * (sc[1] + 2.0 * sc[0] + (double)N * std::log(2.0 * M_PI));   // src/loss_grad.jl:40
  }

  if (G) {
    // ---- K^-1 = W W^T in place, then the fused all-hyper-parameter reduction over the local columns
    tick();
    comm.dma_mode = mg->prefetch_lauum;
    set_phase(4);
    db.lauum();
    set_phase(8);
    { int rc = tock(GPR_T_LAUUM); if (rc) return rc; }
    m->have_inverse = true;
    tick();
    std::vector<double> tot((size_t)P + 1, 0.0);
    for (auto& R : mg->rk) {
      const int r = R.rank;
      gpr_ctx* ctx = R.ctx;
      MCK(cudaSetDevice(ctx->device));
      GradArgs a{};
      a.Kinv = R.L; a.ld = ld; a.alpha = R.alpha; a.x = R.x; a.D = m->D; a.N = N; a.hp = R.hp; a.spec = m->spec; a.P = P;
      a.eps = eps; a.partial = R.gpart; a.G = Gn; a.rank = r; a.nbt = (int)(nb / GR_TILE); a.lcol_tiles = lay.nloc(r) * (nb / GR_TILE);
      const size_t smem = ((size_t)(P + 1) * GR_THREADS + 2 * (size_t)m->D * GR_TILE + 2 * GR_TILE + 32) * sizeof(double);
      grad_reduce_kernel<true><<<R.gr_blocks, GR_THREADS, smem, ctx->stream>>>(a);
      ctx->launches++;
      MCK(cudaGetLastError());
      grad_partial_sum_kernel<<<(P + 1 + 127) / 128, 128, 0, ctx->stream>>>(R.gpart, R.gr_blocks, P, R.scal + 2);
      ctx->launches++;
      MCK(cudaGetLastError());
    }
    { int rc = comm.sum_scal(2, P + 1, tot.data()); if (rc) return rc; }   // fixed rank order (or NCCL's fixed tree): deterministic
    // sigma: -acc/|sigma| ; l_d: +l_d * acc ; noise: -sigma_n * acc_diag  (loss_grad.jl:43-52, deriv_covar.jl:23,26,31)
    for (int c = 0; c < m->spec.ncomp; ++c) {
      const int off = m->spec.hp_off[c];
      if (m->spec.type[c] == KT_NOISE) G[off] = -hp[off] * tot[P];
      else {
        G[off] = -tot[off] / std::fabs(hp[off]);
        for (int d = 0; d < m->D; ++d) G[off + 1 + d] = hp[off + 1 + d] * tot[off + 1 + d];
      }
    }
    if (log_scale) for (int p = 0; p < P; ++p) G[p] *= hp[p];   // src/cost.jl:65
    { int rc = tock(GPR_T_GRAD); if (rc) return rc; }
  }
  if (F) *F = Fv;
  cudaSetDevice(mg->rk[0].ctx->device);
  cudaEventRecord(tot1, mg->rk[0].ctx->stream);
  cudaEventSynchronize(tot1);
  { float t = 0.f; cudaEventElapsedTime(&t, tot0, tot1); m->ms[GPR_T_TOTAL] = t; }
  return msync_all(mg, "nlml_grad", nullptr);
}

int gpr_mgpu_timings(gpr_mgpu_model* m, double* ms, int n) {
  if (!m || !ms) return GPR_ERR_ARG;
  for (int i = 0; i < n && i < GPR_T_COUNT; ++i) ms[i] = m->ms[i];
  return GPR_OK;
}

/* which: GPR_FETCH_ALPHA (N) or GPR_FETCH_KINV (N x N, full symmetric; needs a preceding gradient evaluation) */
int gpr_mgpu_fetch(gpr_mgpu_model* m, int which, double* out) {
  if (!m || !out) return GPR_ERR_ARG;
  gpr_mgpu* mg = m->mg;
  const DistLayout& lay = m->md.lay;
  const int64_t N = m->N, nb = lay.nb, ld = m->md.ld;
  if (which == GPR_FETCH_KINV && mg->transport == 2)
    return mfail(mg, GPR_ERR_UNSUPPORTED, "fetch K^-1: with one process per rank every process holds its own block columns only");
  if (which == GPR_FETCH_ALPHA) {
    if (!m->have_alpha) return mfail(mg, GPR_ERR_STATE, "fetch alpha: the last evaluation was loss-only (no back substitution)");
    MRank& R = mg->rk[0];
    MCK(cudaSetDevice(R.ctx->device));
    MCK(cudaMemcpyAsync(out, R.alpha, sizeof(double) * N, cudaMemcpyDeviceToHost, R.ctx->stream));
    MCK(cudaStreamSynchronize(R.ctx->stream));
    return GPR_OK;
  }
  if (which != GPR_FETCH_KINV) return mfail(mg, GPR_ERR_ARG, "unknown fetch selector");
  if (!m->have_inverse) return mfail(mg, GPR_ERR_STATE, "fetch K^-1: cache holds no inverse");
  for (auto& R : mg->rk) {
    const int r = R.rank;
    MCK(cudaSetDevice(R.ctx->device));
    for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
      const int64_t J = lay.gblock(r, lb);
      const int64_t cv = std::max<int64_t>(0, std::min<int64_t>(nb, N - J * nb));
      if (cv > 0)
        MCK(cudaMemcpy2DAsync(out + J * nb * N, sizeof(double) * N, R.L + lb * nb * ld, sizeof(double) * ld, sizeof(double) * N, cv,
                              cudaMemcpyDeviceToHost, R.ctx->stream));
    }
    MCK(cudaStreamSynchronize(R.ctx->stream));
  }
  for (int64_t j = 0; j < N; ++j)
    for (int64_t i = j + 1; i < N; ++i) out[i + j * N] = out[j + i * N];
  return GPR_OK;
}

/* diagnostics: distributed factorization of a host SPD matrix A (N x N, upper triangle referenced) with right-hand
 * sides Y (N x ny, may be NULL with ny = 0).  mode 0: potrf (upper(A) <- U, Y <- U^-T Y), 1: + trtri (upper(A) <- U^-1,
 * Y <- -A^-1 Y), 2: + lauum (upper(A) <- A^-1).  ms[0..2] = phase times. */
int gpr_mgpu_dbg_factor(gpr_mgpu* mg, double* A, int64_t N, double* Y, int ny, int mode, int64_t* info, double* ms) {
  if (!mg) return GPR_ERR_ARG;
  if (!A || N < 1 || (ny > 0 && !Y)) return mfail(mg, GPR_ERR_ARG, "NULL or empty argument");
  if (mg->rk[0].L) return mfail(mg, GPR_ERR_STATE, "this multi-GPU context already holds a model");
  const int64_t nb = mg->nb, Np = round_up(N, nb), nyp = ny > 0 ? round_up(ny, 128) : 0;
  MDense md;
  int rc = mdense_alloc(mg, md, Np, nyp);
  if (rc) { mdense_free(mg); return rc; }
  const DistLayout& lay = md.lay;
  std::vector<double> col((size_t)Np * nb);
  auto body = [&]() -> int {
    for (auto& R : mg->rk) {
      const int r = R.rank;
      MCK(cudaSetDevice(R.ctx->device));
      for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
        const int64_t J = lay.gblock(r, lb);
        std::fill(col.begin(), col.end(), 0.0);
        for (int64_t c = 0; c < nb; ++c) {
          const int64_t gc = J * nb + c;
          if (gc < N) for (int64_t i = 0; i <= gc; ++i) col[i + c * Np] = A[i + gc * N];
          else col[gc + c * Np] = 1.0;
        }
        MCK(cudaMemcpyAsync(R.L + lb * nb * md.ld, col.data(), sizeof(double) * Np * nb, cudaMemcpyHostToDevice, R.ctx->stream));
        MCK(cudaStreamSynchronize(R.ctx->stream));
      }
      if (r == lay.y_owner() && nyp > 0) {
        std::vector<double> yb((size_t)Np * nyp, 0.0);
        for (int c = 0; c < ny; ++c) memcpy(&yb[(size_t)c * Np], Y + (int64_t)c * N, sizeof(double) * N);
        MCK(cudaMemcpyAsync(R.L + lay.ycol0(r) * md.ld, yb.data(), sizeof(double) * Np * nyp, cudaMemcpyHostToDevice, R.ctx->stream));
        MCK(cudaStreamSynchronize(R.ctx->stream));
      }
    }
    auto ranks = mdense_ranks(mg, md);
    std::unique_ptr<CommBase> comm_p = make_comm(mg, lay, md.ld);
    CommBase& comm = *comm_p;
    DistBlocked<CudaBE, CommBase> db(lay, ranks, comm);
    db.prefetch_trtri = mg->prefetch_trtri != 0;
    db.prefetch_lauum = mg->prefetch_lauum != 0;
    double t[3] = {0, 0, 0};
    for (int ph = 0; ph <= mode && ph < 3; ++ph) {
      cudaSetDevice(mg->rk[0].ctx->device);
      cudaEventRecord(mg->t0, mg->rk[0].ctx->stream);
      comm.dma_mode = ph == 1 ? mg->prefetch_trtri : mg->prefetch_lauum;
      if (ph == 0) db.potrf(); else if (ph == 1) db.trtri(); else db.lauum();
      for (auto& R : mg->rk) { cudaSetDevice(R.ctx->device); MCK(cudaStreamSynchronize(R.ctx->stream)); }
      cudaSetDevice(mg->rk[0].ctx->device);
      cudaEventRecord(mg->t1, mg->rk[0].ctx->stream);
      MCK(cudaEventSynchronize(mg->t1));
      float f = 0.f; cudaEventElapsedTime(&f, mg->t0, mg->t1);
      t[ph] = f;
    }
    if (ms) { ms[0] = t[0]; ms[1] = t[1]; ms[2] = t[2]; }
    long long h_info = 0;
    int rc2 = msync_all(mg, "dbg_factor", &h_info);
    if (rc2) return rc2;
    if (info) *info = h_info;
    for (auto& R : mg->rk) {
      const int r = R.rank;
      MCK(cudaSetDevice(R.ctx->device));
      for (int64_t lb = 0; lb < lay.nloc(r); ++lb) {
        const int64_t J = lay.gblock(r, lb);
        const int64_t cv = std::max<int64_t>(0, std::min<int64_t>(nb, N - J * nb));
        if (cv > 0)
          MCK(cudaMemcpy2DAsync(A + J * nb * N, sizeof(double) * N, R.L + lb * nb * md.ld, sizeof(double) * md.ld, sizeof(double) * N, cv,
                                cudaMemcpyDeviceToHost, R.ctx->stream));
      }
      if (r == lay.y_owner() && ny > 0)
        MCK(cudaMemcpy2DAsync(Y, sizeof(double) * N, R.L + lay.ycol0(r) * md.ld, sizeof(double) * md.ld, sizeof(double) * N, ny,
                              cudaMemcpyDeviceToHost, R.ctx->stream));
      MCK(cudaStreamSynchronize(R.ctx->stream));
    }
    return h_info ? GPR_ERR_NOT_POSDEF : GPR_OK;
  };
  rc = body();
  mdense_free(mg);
  return rc;
}

}  // extern "C"
