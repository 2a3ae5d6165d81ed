// SyncBatchNorm statistics exchange over NVLink peer memory (replaces the all_gather / all_reduce that
// torch.nn.SyncBatchNorm issues per layer, reference configs/base.yaml:22 -> torch/nn/modules/_functions.py:39-170).
//
// Every rank owns one "symmetric" buffer that all peers have mapped (torch.distributed._symmetric_memory or
// cudaIpc); VtbSyncBn carries the peer-mapped base pointers.  Buffer layout (bytes):
//   [0,4)        sequence counter of this rank (device-resident, so replay from a CUDA graph stays consistent)
//   [256,512)    flags[2][kMaxRanks] uint32: flags[parity][r] = last sequence number rank r has published to me
//   [1024, ...)  slots[2][kMaxRanks][kSlotDoubles] double: slot[parity][r] = rank r's per-channel sums
// One exchange = one single-block kernel, fused with the work around it:
//   1. reduce my partial rows to fp64 per-channel sums           (what vtb_bn_stats_reduce does)
//   2. PUSH them into slot[parity][my_rank] of every peer (posted NVLink writes), fence.sys, publish the flag
//   3. wait until all world flags[parity][*] carry this sequence number (spin on LOCAL memory)
//   4. sum the world slots in rank order (same order on every rank -> bit-identical statistics everywhere)
//   5. finalise                                                   (what vtb_bn_finalize / _bwd_finalize do)
// The parity double-buffering is sufficient: a rank can only reach sequence s+2 after every peer has signalled s+1,
// which a peer does after its kernel s has finished reading.  ~one NVLink write latency instead of an NCCL call.
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/vtb.h"
#include "common.cuh"

namespace vtb {

constexpr int kMaxRanks = VTB_SYNC_MAX_RANKS;
constexpr int kSlotDoubles = VTB_SYNC_MAX_CHANNELS * 2;
constexpr size_t kFlagsOff = 256, kSlotsOff = 1024;

struct Peers {
  unsigned char* base[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_volatile_d2(const double* p) {
  double2 v;
  asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

// steps 1-4; returns (through smem-free per-thread loop) the GLOBAL sums of the channels this thread owns via callback
template <typename Finalize>
__device__ __forceinline__ void exchange_and_finalize(const float* __restrict__ partial, int rows, int c, const Peers& peers,
                                                      int rank, int world, double* local_out, Finalize fin) {
  __shared__ unsigned int s_seq;
  unsigned char* mine = peers.base[rank];
  if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned int*>(mine) + 1u;
  __syncthreads();
  const unsigned int seq = s_seq;
  const unsigned int par = seq & 1u;
  // 1 + 2: local sums -> every peer's slot[par][rank]
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double s = 0.0, q = 0.0;
    const float* src = partial + (size_t)ch * 2;
    for (int r = 0; r < rows; r += 8) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (r + u < rows) ? __ldg(reinterpret_cast<const float2*>(src + (size_t)(r + u) * c * 2)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s += (double)v[u].x;
        q += (double)v[u].y;
      }
    }
    if (local_out) {
      local_out[ch * 2] = s;
      local_out[ch * 2 + 1] = q;
    }
    for (int p = 0; p < world; ++p) {
      double* slot = reinterpret_cast<double*>(peers.base[p] + kSlotsOff) + ((size_t)par * kMaxRanks + rank) * kSlotDoubles;
      *reinterpret_cast<double2*>(slot + ch * 2) = make_double2(s, q);
    }
  }
  __threadfence_system();
  __syncthreads();
  // publish, then wait for everybody (thread r handles peer r)
  if ((int)threadIdx.x < world) {
    unsigned int* pf = reinterpret_cast<unsigned int*>(peers.base[threadIdx.x] + kFlagsOff) + par * kMaxRanks + rank;
    st_release_sys(pf, seq);
    const unsigned int* mf = reinterpret_cast<const unsigned int*>(mine + kFlagsOff) + par * kMaxRanks + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mf) != seq) {
      if (clock64() - t0 > 20000000000LL) {   // ~10 s: a peer died; fail loudly instead of hanging the box
        printf("vtb: SyncBN exchange timeout (rank %d waiting for rank %d, seq %u)\n", rank, (int)threadIdx.x, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  // 4 + 5
  const double* slots = reinterpret_cast<const double*>(mine + kSlotsOff) + (size_t)par * kMaxRanks * kSlotDoubles;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int r = 0; r < world; ++r) {
      const double2 v = ld_volatile_d2(slots + (size_t)r * kSlotDoubles + ch * 2);
      s += v.x;
      q += v.y;
    }
    fin(ch, s, q);
  }
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned int*>(mine) = seq;
}

__global__ void __launch_bounds__(1024)
bn_sync_finalize_kernel(const float* __restrict__ partial, int rows, int c, Peers peers, int rank, int world,
                        double count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        float momentum, float* running_mean, float* running_var, long long* nbt, float* mean_out,
                        float* invstd_out, float* scale, float* shift) {
  pdl_wait();
  pdl_trigger();
  exchange_and_finalize(partial, rows, c, peers, rank, world, nullptr, [&](int ch, double s, double q) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    mean_out[ch] = (float)mean;
    invstd_out[ch] = invstd;
    const float sc = gamma[ch] * invstd;
    scale[ch] = sc;
    shift[ch] = beta[ch] - (float)mean * sc;
    if (running_mean) {
      const double unbiased = count > 1 ? var * (count / (count - 1.0)) : var;
      running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
      running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
    if (nbt && ch == 0) *nbt += 1;
  });
}

__global__ void __launch_bounds__(1024)
bn_sync_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int c, Peers peers, int rank, int world,
                            double count, float* dgamma, float* dbeta, int accumulate, float* __restrict__ coef,
                            double* __restrict__ local_scratch) {
  pdl_wait();
  pdl_trigger();
  exchange_and_finalize(partial, rows, c, peers, rank, world, local_scratch, [&](int ch, double s, double q) {
    // parameter gradients come from the LOCAL sums (the gradient all-reduce averages them afterwards, like DDP);
    // the dx formula needs the GLOBAL means
    const double ls = local_scratch[ch * 2], lq = local_scratch[ch * 2 + 1];
    if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)lq : (float)lq;
    if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)ls : (float)ls;
    coef[ch * 2] = (float)(s / count);
    coef[ch * 2 + 1] = (float)(q / count);
  });
}

static bool sync_ok(const VtbSyncBn* s, int c) {
  if (!s || s->world < 1 || s->world > kMaxRanks || s->rank < 0 || s->rank >= s->world || c <= 0 ||
      c > VTB_SYNC_MAX_CHANNELS)
    return false;
  for (int r = 0; r < s->world; ++r)
    if (!s->peer_buffers[r]) return false;
  return true;
}
static Peers make_peers(const VtbSyncBn* s) {
  Peers p;
  for (int r = 0; r < kMaxRanks; ++r) p.base[r] = r < s->world ? (unsigned char*)s->peer_buffers[r] : nullptr;
  return p;
}

}  // namespace vtb

using namespace vtb;

extern "C" {

size_t vtb_bn_sync_buffer_bytes(void) { return kSlotsOff + (size_t)2 * kMaxRanks * kSlotDoubles * sizeof(double); }

int vtb_bn_sync_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                         const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                         float* shift, void* stream) {
  if (!partial || rows <= 0 || !sync_ok(sync, c) || count <= 0 || !gamma || !beta || !mean || !invstd || !scale ||
      !shift || ((running_mean == nullptr) != (running_var == nullptr)))
    return fail(VTB_EINVAL, "vtb_bn_sync_finalize: bad arguments");
  const int threads = std::min(1024, ((c + 31) / 32) * 32);
  count_launch(1);
  return check_cuda((int)launch_pdl(bn_sync_finalize_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream, partial,
                                    rows, c, make_peers(sync), sync->rank, sync->world, count, gamma, beta, eps,
                                    momentum, running_mean, running_var, num_batches_tracked, mean, invstd, scale,
                                    shift),
                    "bn_sync_finalize_kernel");
}

int vtb_bn_sync_bwd_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                             float* dgamma, float* dbeta, int accumulate, float* coef, double* local_scratch,
                             void* stream) {
  if (!partial || rows <= 0 || !sync_ok(sync, c) || count <= 0 || !coef || !local_scratch)
    return fail(VTB_EINVAL, "vtb_bn_sync_bwd_finalize: bad arguments");
  const int threads = std::min(1024, ((c + 31) / 32) * 32);
  count_launch(1);
  return check_cuda((int)launch_pdl(bn_sync_bwd_finalize_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream,
                                    partial, rows, c, make_peers(sync), sync->rank, sync->world, count, dgamma, dbeta,
                                    accumulate, coef, local_scratch),
                    "bn_sync_bwd_finalize_kernel");
}

}  // extern "C"
