// Stand-alone SyncBatchNorm exchange kernels (see syncbn.cuh for the protocol): replace the all_gather / all_reduce that
// torch.nn.SyncBatchNorm issues per layer (reference configs/base.yaml:22 -> torch/nn/modules/_functions.py:39-170) when
// the exchange is not fused into the producing kernel.
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/vtb.h"
#include "common.cuh"
#include "syncbn.cuh"

namespace vtb {

typedef SyncPeers Peers;

// reduce my partial rows, exchange, hand the GLOBAL sums of every channel to `fin` (single 1024-thread block)
template <typename Finalize>
__device__ __forceinline__ void exchange_and_finalize(const float* __restrict__ partial, int rows, int c, const Peers& peers,
                                                      double* local_out, Finalize fin) {
  __shared__ unsigned int s_seq;
  if (threadIdx.x == 0) s_seq = sync_read_seq(peers);
  __syncthreads();
  const unsigned int seq = s_seq;
  // local sums -> every peer's slot.  All threads share the row reduction: thread = (row group g, channel), 8 independent
  // loads in flight each; the groups are combined through shared memory in a fixed order.
  __shared__ double2 s_part[1024];
  for (int cb = 0; cb < c; cb += blockDim.x) {          // channel blocks (one for c <= 1024)
    const int cw = min((int)blockDim.x, c - cb);
    const int G = max(1, (int)blockDim.x / cw);
    const int g = threadIdx.x / cw, lc = threadIdx.x - g * cw;
    if (g < G) {
      double s = 0.0, q = 0.0;
      const float* src = partial + (size_t)(cb + lc) * 2;
      for (int r = g; r < rows; r += 8 * G) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r + u * G;
          v[u] = (rr < rows) ? __ldg(reinterpret_cast<const float2*>(src + (size_t)rr * c * 2)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          s += (double)v[u].x;
          q += (double)v[u].y;
        }
      }
      s_part[g * cw + lc] = make_double2(s, q);
    }
    __syncthreads();
    if ((int)threadIdx.x < cw) {
      double s = 0.0, q = 0.0;
      for (int gg = 0; gg < G; ++gg) {
        const double2 v = s_part[gg * cw + threadIdx.x];
        s += v.x;
        q += v.y;
      }
      const int ch = cb + threadIdx.x;
      if (local_out) {
        local_out[ch * 2] = s;
        local_out[ch * 2 + 1] = q;
      }
      sync_push(peers, seq, ch, s, q);
    }
    __syncthreads();
  }
  if (!peers.tagged) {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < peers.world) sync_signal_wait(peers, kSyncFlagsStd, 0, seq, threadIdx.x);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const double2 v = sync_gather(peers, seq, ch);
    fin(ch, v.x, v.y);
  }
  __syncthreads();
  if (threadIdx.x == 0) sync_write_seq(peers, seq);
}

__global__ void __launch_bounds__(1024)
bn_sync_finalize_kernel(const float* __restrict__ partial, int rows, int c, Peers peers,
                        double count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        float momentum, float* running_mean, float* running_var, long long* nbt, float* mean_out,
                        float* invstd_out, float* scale, float* shift) {
  pdl_wait();
  pdl_trigger();
  exchange_and_finalize(partial, rows, c, peers, nullptr, [&](int ch, double s, double q) {
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
bn_sync_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int c, Peers peers,
                            double count, float* dgamma, float* dbeta, int accumulate, float* __restrict__ coef,
                            double* __restrict__ local_scratch) {
  pdl_wait();
  pdl_trigger();
  exchange_and_finalize(partial, rows, c, peers, local_scratch, [&](int ch, double s, double q) {
    // parameter gradients come from the LOCAL sums (the gradient all-reduce averages them afterwards, like DDP);
    // the dx formula needs the GLOBAL means
    const double ls = local_scratch[ch * 2], lq = local_scratch[ch * 2 + 1];
    if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)lq : (float)lq;
    if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)ls : (float)ls;
    coef[ch * 2] = (float)(s / count);
    coef[ch * 2 + 1] = (float)(q / count);
  });
}

static bool sync_ok(const VtbSyncBn* s, int c) { return sync_args_ok(s, c); }
static Peers make_peers(const VtbSyncBn* s) { return make_sync_peers(s); }

}  // namespace vtb

using namespace vtb;

extern "C" {

size_t vtb_bn_sync_buffer_bytes(void) { return kSyncBufferBytes; }

int vtb_bn_sync_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                         const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                         float* shift, void* stream) {
  if (!partial || rows <= 0 || !sync_ok(sync, c) || count <= 0 || !gamma || !beta || !mean || !invstd || !scale ||
      !shift || ((running_mean == nullptr) != (running_var == nullptr)))
    return fail(VTB_EINVAL, "vtb_bn_sync_finalize: bad arguments");
  const int threads = 1024;
  count_launch(1);
  return check_cuda((int)launch_pdl(bn_sync_finalize_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream, partial,
                                    rows, c, make_peers(sync), count, gamma, beta, eps,
                                    momentum, running_mean, running_var, num_batches_tracked, mean, invstd, scale,
                                    shift),
                    "bn_sync_finalize_kernel");
}

int vtb_bn_sync_bwd_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                             float* dgamma, float* dbeta, int accumulate, float* coef, double* local_scratch,
                             void* stream) {
  if (!partial || rows <= 0 || !sync_ok(sync, c) || count <= 0 || !coef || !local_scratch)
    return fail(VTB_EINVAL, "vtb_bn_sync_bwd_finalize: bad arguments");
  const int threads = 1024;
  count_launch(1);
  return check_cuda((int)launch_pdl(bn_sync_bwd_finalize_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream,
                                    partial, rows, c, make_peers(sync), count, dgamma, dbeta,
                                    accumulate, coef, local_scratch),
                    "bn_sync_bwd_finalize_kernel");
}

}  // extern "C"
