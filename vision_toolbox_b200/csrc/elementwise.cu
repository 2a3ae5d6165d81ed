// HBM-bound pieces of the ConvNormAct path: BatchNorm statistics finalisation, normalise(+ReLU)(+residual)
// materialisation, BatchNorm/ReLU backward (reduce + apply), gradient fan-in adds, layout conversion.
// All activations are NHWC bf16 views (pointer + pixel pitch); one thread handles 8 channels (16 bytes).
#include <algorithm>
#include <cstdlib>
#include <cstdint>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/vtb.h"
#include "common.cuh"
#include "syncbn.cuh"

namespace vtb {

__device__ __forceinline__ float lo16(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float hi16(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = lo16(u.x); f[1] = hi16(u.x); f[2] = lo16(u.y); f[3] = hi16(u.y);
  f[4] = lo16(u.z); f[5] = hi16(u.z); f[6] = lo16(u.w); f[7] = hi16(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack2(f[0], f[1]); u.y = pack2(f[2], f[3]); u.z = pack2(f[4], f[5]); u.w = pack2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

static int ew_grid(long long vecs, int block) {
  const int sms = std::max(1, num_sms());
  return (int)std::max<long long>(1, std::min<long long>((vecs + block - 1) / block, (long long)sms * 16));
}

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics: partial rows -> (optionally) double sums -> mean/invstd/scale/shift + running stats
// ---------------------------------------------------------------------------------------------
// one warp per channel
__global__ void bn_stats_reduce_kernel(const float* __restrict__ partial, int rows, int c, double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s = 0, q = 0;
  for (int r = lane; r < rows; r += 32) {
    const float2 v = *reinterpret_cast<const float2*>(partial + ((size_t)r * c + ch) * 2);
    s += v.x;
    q += v.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    sums[ch * 2] = s;
    sums[ch * 2 + 1] = q;
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ partial, int rows, const double* __restrict__ sums_in,
                                   double count, int c, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, long long* nbt, float* mean_out, float* invstd_out,
                                   float* scale, float* shift) {
  pdl_wait();
  pdl_trigger();
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s = 0, q = 0;
  if (sums_in) {
    s = sums_in[ch * 2];
    q = sums_in[ch * 2 + 1];
  } else {
#pragma unroll 4
    for (int r = lane; r < rows; r += 32) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * c + ch) * 2));
      s += v.x;
      q += v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma[ch], b = beta[ch];
    mean_out[ch] = (float)mean;
    invstd_out[ch] = invstd;
    const float sc = g * invstd;
    scale[ch] = sc;
    shift[ch] = b - (float)mean * sc;
    if (running_mean) {
      const double unbiased = count > 1 ? var * (count / (count - 1.0)) : var;
      running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
      running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
    if (nbt && ch == 0) *nbt += 1;
  }
}

__global__ void bn_eval_affine_kernel(int c, const float* gamma, const float* beta, const float* rm, const float* rv,
                                      float eps, float* scale, float* shift) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float sc = gamma[ch] / sqrtf(rv[ch] + eps);
  scale[ch] = sc;
  shift[ch] = beta[ch] - rm[ch] * sc;
}

// ---------------------------------------------------------------------------------------------
// Thread geometry of the per-pixel passes: a block is `ppi` pixel lanes x `cv` channel vectors (8 channels = 16 bytes
// each), so a thread's channels - and with them its per-channel parameters - are loop invariant: they are loaded once
// into registers and the pixel loop contains no division and no parameter loads.  The pixel loop is unrolled kEwUnroll
// times with all loads issued before the first use (memory-level parallelism is what reaches HBM bandwidth here).
// ---------------------------------------------------------------------------------------------
constexpr int kEwUnroll = 4;
struct EwGeom {
  int cv;      // channel vectors per block (<= 64); blockDim.x = ppi * cv
  int ppi;     // pixel lanes per block
  int chunks;  // blocks along the channel axis (gridDim.y)
  int rows;    // blocks along the pixel axis (gridDim.x)
};
static EwGeom ew_geom(long long pixels, int c8, int max_threads, int blocks_per_sm) {
  EwGeom g;
  g.cv = std::min(c8, 64);
  g.chunks = (c8 + g.cv - 1) / g.cv;
  g.cv = (c8 + g.chunks - 1) / g.chunks;            // even split (c8 = 20/24/28 for VoVNet's 160/192/224 channels)
  g.ppi = std::max(1, max_threads / g.cv);
  const int sms = std::max(1, num_sms());
  const long long want = (pixels + (long long)g.ppi * kEwUnroll - 1) / ((long long)g.ppi * kEwUnroll);
  g.rows = (int)std::max<long long>(1, std::min<long long>(want, std::max(1, sms * blocks_per_sm / g.chunks)));
  return g;
}

// ---------------------------------------------------------------------------------------------
// out = [relu](y*scale + shift) [+ residual]     (bf16 rounding after the affine+relu, then after the add:
// the reference adds two bf16 tensors, darknet.py:28 / vovnet.py:61)
// ---------------------------------------------------------------------------------------------
template <bool RELU, bool RES>
__global__ void __launch_bounds__(512)
bn_act_kernel(const __nv_bfloat16* __restrict__ y, int ldy, long long pixels, int c8, int cv,
              const float* __restrict__ scale, const float* __restrict__ shift,
              const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_wait();
  pdl_trigger();
  const int vl = threadIdx.x % cv, pl = threadIdx.x / cv, ppi = blockDim.x / cv;
  const int v = blockIdx.y * cv + vl;
  if (v >= c8) return;
  const int ch = v * 8;
  float sc[8], sh[8];
  load8f(scale + ch, sc);
  load8f(shift + ch, sh);
  const long long step = (long long)gridDim.x * ppi;
  for (long long pix = (long long)blockIdx.x * ppi + pl; pix < pixels; pix += step * kEwUnroll) {
    uint4 yv[kEwUnroll], rv[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long q = pix + u * step;
      if (q < pixels) {
        yv[u] = __ldg(reinterpret_cast<const uint4*>(y + q * ldy + ch));
        if (RES) rv[u] = __ldg(reinterpret_cast<const uint4*>(res + q * ldr + ch));
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long q = pix + u * step;
      if (q < pixels) {
        float f[8];
        unpack8(yv[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], sc[j], sh[j]);
          if (RELU) f[j] = fmaxf(f[j], 0.f);
        }
        if (RES) {
          float r[8];
          unpack8(rv[u], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = rbf(f[j]) + r[j];
        }
        *reinterpret_cast<uint4*>(out + q * ldo + ch) = pack8(f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(+ReLU) backward, stage 1: per-channel sums of dz and dz*xhat
//   z = y*scale+shift, dz = dout * (z > 0) (ReLU mask recomputed from the saved conv output), xhat = (y-mean)*invstd
// one partial row per block along the pixel axis; the pixel lanes of a block are combined in a fixed order
// ---------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, int lddo, const __nv_bfloat16* __restrict__ y, int ldy,
                     long long pixels, int c8, int cv, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, float* __restrict__ partial, int c) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float red[];   // [blockDim.x][17]
  const int t = threadIdx.x;
  const int vl = t % cv, pl = t / cv, ppi = blockDim.x / cv;
  const int v = blockIdx.y * cv + vl;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (v < c8) {
    const int ch = v * 8;
    float sc[8], sh[8], mu[8], is[8];
    load8f(scale + ch, sc);
    load8f(shift + ch, sh);
    load8f(mean + ch, mu);
    load8f(invstd + ch, is);
    const long long step = (long long)gridDim.x * ppi;
    for (long long pix = (long long)blockIdx.x * ppi + pl; pix < pixels; pix += step * kEwUnroll) {
      uint4 gv[kEwUnroll], yv[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long q = pix + u * step;
        if (q < pixels) {
          gv[u] = __ldg(reinterpret_cast<const uint4*>(dout + q * lddo + ch));
          yv[u] = __ldg(reinterpret_cast<const uint4*>(y + q * ldy + ch));
        }
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long q = pix + u * step;
        if (q < pixels) {
          float g[8], yy[8];
          unpack8(gv[u], g);
          unpack8(yv[u], yy);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = fmaf(yy[j], sc[j], sh[j]);
            const float dz = (!RELU || z > 0.f) ? g[j] : 0.f;
            s1[j] += dz;
            s2[j] = fmaf(dz, (yy[j] - mu[j]) * is[j], s2[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[t * 17 + j] = s1[j];
    red[t * 17 + 8 + j] = s2[j];
  }
  __syncthreads();
  // outputs: cv vectors x 16 values; sum over the pixel lanes in fixed order (deterministic)
  for (int o = t; o < cv * 16; o += blockDim.x) {
    const int vl2 = o / 16, j = o % 16;
    const int vg = blockIdx.y * cv + vl2;
    if (vg >= c8) continue;
    float acc = 0.f;
    for (int q = 0; q < ppi; ++q) acc += red[(q * cv + vl2) * 17 + j];
    const int ch = vg * 8 + (j & 7);
    partial[((size_t)blockIdx.x * c + ch) * 2 + (j >> 3)] = acc;
  }
}

// stage 2: partial rows (or cross-rank-reduced double sums) -> dgamma/dbeta (local sums) and the two means
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, const double* __restrict__ sums_in,
                                       double count, int c, float* dgamma, float* dbeta, int accumulate,
                                       const double* __restrict__ local_sums, float* __restrict__ coef,
                                       double* __restrict__ sums_out) {
  pdl_wait();
  pdl_trigger();
  const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  double s = 0, q = 0;
  if (sums_in) {
    s = sums_in[ch * 2];
    q = sums_in[ch * 2 + 1];
  } else {
#pragma unroll 4
    for (int r = lane; r < rows; r += 32) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * c + ch) * 2));
      s += v.x;
      q += v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
  }
  if (lane == 0) {
    if (sums_out) {  // SyncBN step 1: publish local sums only
      sums_out[ch * 2] = s;
      sums_out[ch * 2 + 1] = q;
      return;
    }
    // parameter gradients use the LOCAL sums (DDP averages them afterwards), the dx formula the global ones
    const double ls = local_sums ? local_sums[ch * 2] : s;
    const double lq = local_sums ? local_sums[ch * 2 + 1] : q;
    if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)lq : (float)lq;
    if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)ls : (float)ls;
    coef[ch * 2] = (float)(s / count);
    coef[ch * 2 + 1] = (float)(q / count);
  }
}

// stage 3: dy = scale * (dz - mean_dz - xhat * mean_dzxhat)
template <bool RELU>
__global__ void __launch_bounds__(512)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, int lddo, const __nv_bfloat16* __restrict__ y, int ldy,
                    long long pixels, int c8, int cv, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ coef,
                    __nv_bfloat16* __restrict__ dy, int lddy) {
  pdl_wait();
  pdl_trigger();
  const int vl = threadIdx.x % cv, pl = threadIdx.x / cv, ppi = blockDim.x / cv;
  const int v = blockIdx.y * cv + vl;
  if (v >= c8) return;
  const int ch = v * 8;
  // per-channel constants: dy = sc*dz - k0 - k1*(y - mu)   with k0 = sc*coef0, k1 = sc*invstd*coef1
  float sc[8], sh[8], mu[8], k0[8], k1[8];
  {
    float is[8], cf[16];
    load8f(scale + ch, sc);
    load8f(shift + ch, sh);
    load8f(mean + ch, mu);
    load8f(invstd + ch, is);
    load8f(coef + ch * 2, cf);
    load8f(coef + ch * 2 + 8, cf + 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k0[j] = sc[j] * cf[2 * j];
      k1[j] = sc[j] * is[j] * cf[2 * j + 1];
    }
  }
  const long long step = (long long)gridDim.x * ppi;
  for (long long pix = (long long)blockIdx.x * ppi + pl; pix < pixels; pix += step * kEwUnroll) {
    uint4 gv[kEwUnroll], yv[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long q = pix + u * step;
      if (q < pixels) {
        gv[u] = __ldg(reinterpret_cast<const uint4*>(dout + q * lddo + ch));
        yv[u] = __ldg(reinterpret_cast<const uint4*>(y + q * ldy + ch));
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const long long q = pix + u * step;
      if (q < pixels) {
        float g[8], yy[8], o[8];
        unpack8(gv[u], g);
        unpack8(yv[u], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(yy[j], sc[j], sh[j]);
          const float dz = (!RELU || z > 0.f) ? g[j] : 0.f;
          o[j] = fmaf(sc[j], dz, -k0[j]) - k1[j] * (yy[j] - mu[j]);
        }
        *reinterpret_cast<uint4*>(dy + q * lddy + ch) = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(+ReLU) backward in ONE cooperative launch: stage 1 (partial sums) -> grid barrier over the blocks of a
// channel chunk -> every block finalises its chunk's (mean dz, mean dz*xhat) from the partial rows (fp64, fixed order)
// -> stage 3 (apply).  Saves two launches per layer and the second pass finds dout / y in L2 for mid-sized layers.
// sync[4 * chunk + {0,1,2}]: zero-initialised arrive / depart / ready counters, sync[255]: finished chunk exchanges; all
// left zero again (self-cleaning).  Under SyncBN (sp.world > 1) block 0 of every chunk exchanges the chunk's sums with
// all ranks over NVLink peer memory (syncbn.cuh) and broadcasts the global means to its sibling blocks through
// coef_g = partial + gridDim.x * c * 2.
// ---------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(512)
bn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dout, int lddo, const __nv_bfloat16* __restrict__ y, int ldy,
                    long long pixels, int c8, int cv, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, float* __restrict__ partial, int c, double count,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                    unsigned int* __restrict__ sync, __nv_bfloat16* __restrict__ dy, int lddy, SyncPeers sp) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float red[];   // stage 1: [blockDim.x][17] floats; stage 2: doubles [G][pairs][4] then coef [cv*8][2]
  const int t = threadIdx.x;
  const int vl = t % cv, pl = t / cv, ppi = blockDim.x / cv;
  const int v = blockIdx.y * cv + vl;
  const bool active = v < c8;
  const int ch = v * 8;
  const long long step = (long long)gridDim.x * ppi;
  float sc[8], sh[8], mu[8], is[8];
  if (active) {
    load8f(scale + ch, sc);
    load8f(shift + ch, sh);
    load8f(mean + ch, mu);
    load8f(invstd + ch, is);
  }
  // ---------------- stage 1
  {
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (active) {
      for (long long pix = (long long)blockIdx.x * ppi + pl; pix < pixels; pix += step * kEwUnroll) {
        uint4 gv[kEwUnroll], yv[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
          const long long q = pix + u * step;
          if (q < pixels) {
            gv[u] = __ldg(reinterpret_cast<const uint4*>(dout + q * lddo + ch));
            yv[u] = __ldg(reinterpret_cast<const uint4*>(y + q * ldy + ch));
          }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
          const long long q = pix + u * step;
          if (q < pixels) {
            float g[8], yy[8];
            unpack8(gv[u], g);
            unpack8(yv[u], yy);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float z = fmaf(yy[j], sc[j], sh[j]);
              const float dz = (!RELU || z > 0.f) ? g[j] : 0.f;
              s1[j] += dz;
              s2[j] = fmaf(dz, (yy[j] - mu[j]) * is[j], s2[j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[t * 17 + j] = s1[j];
      red[t * 17 + 8 + j] = s2[j];
    }
    __syncthreads();
    for (int o = t; o < cv * 16; o += blockDim.x) {
      const int vl2 = o / 16, j = o % 16;
      const int vg = blockIdx.y * cv + vl2;
      if (vg >= c8) continue;
      float acc = 0.f;
      for (int q = 0; q < ppi; ++q) acc += red[(q * cv + vl2) * 17 + j];
      partial[((size_t)blockIdx.x * c + vg * 8 + (j & 7)) * 2 + (j >> 3)] = acc;
    }
  }
  // ---------------- barrier over the gridDim.x blocks of this channel chunk
  unsigned int* cnt = sync + 4 * blockIdx.y;   // arrive, depart, ready
  unsigned int seq = 0;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    if (sp.world > 1) seq = sync_read_seq(sp);   // before any exchange of this launch can complete
    atomicAdd(cnt, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(cnt) < gridDim.x) {
      if (clock64() - t0 > 4000000000LL) __trap();   // co-residency is guaranteed by the cooperative launch
    }
    __threadfence();
  }
  __syncthreads();
  // ---------------- stage 2: this chunk's sums, every block redundantly (thread = channel pair x row group)
  {
    const int chans = min(cv * 8, c - blockIdx.y * cv * 8);
    const int pairs = chans / 2;
    const int G = max(1, (int)blockDim.x / pairs);
    const int pr = t % pairs, g = t / pairs;
    double* dsum = reinterpret_cast<double*>(red);     // [G][pairs][4]
    float* cf = reinterpret_cast<float*>(dsum + (size_t)G * pairs * 4);   // [chans][2]
    float* coef_g = partial + (size_t)gridDim.x * c * 2;
    __shared__ unsigned int s_seq;
    if (t == 0) s_seq = seq;
    if (g < G) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const float* src = partial + ((size_t)blockIdx.y * cv * 8 + pr * 2) * 2;
      const int R = (int)gridDim.x;
      for (int r = g; r < R; r += 8 * G) {   // 8 independent loads in flight per thread
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r + u * G;
          q[u] = (rr < R) ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)rr * c * 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { a0 += q[u].x; a1 += q[u].y; a2 += q[u].z; a3 += q[u].w; }
      }
      double* d = dsum + ((size_t)g * pairs + pr) * 4;
      d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
    }
    __syncthreads();
    double a[4] = {0, 0, 0, 0};
    if (t < pairs) {
      for (int gg = 0; gg < G; ++gg) {
        const double* d = dsum + ((size_t)gg * pairs + t) * 4;
        a[0] += d[0]; a[1] += d[1]; a[2] += d[2]; a[3] += d[3];
      }
      if (blockIdx.x == 0) {   // parameter gradients from the LOCAL sums (the gradient all-reduce averages them)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int gc = blockIdx.y * cv * 8 + t * 2 + h;
          if (dgamma) dgamma[gc] = accumulate ? dgamma[gc] + (float)a[2 * h + 1] : (float)a[2 * h + 1];
          if (dbeta) dbeta[gc] = accumulate ? dbeta[gc] + (float)a[2 * h] : (float)a[2 * h];
        }
      }
    }
    if (sp.world > 1) {
      if (blockIdx.x == 0) {
        const unsigned int sq = s_seq;
        if (t < pairs) {
          const int gc = blockIdx.y * cv * 8 + t * 2;
          sync_push(sp, sq, gc, a[0], a[1]);
          sync_push(sp, sq, gc + 1, a[2], a[3]);
        }
        if (!sp.tagged) {
          __threadfence_system();
          __syncthreads();
          if (t < sp.world) sync_signal_wait(sp, kSyncFlagsBwd, blockIdx.y, sq, t);
          __syncthreads();
        }
        if (t < pairs) {
          const int gc = blockIdx.y * cv * 8 + t * 2;
          const double2 v0 = sync_gather(sp, sq, gc), v1 = sync_gather(sp, sq, gc + 1);
          *reinterpret_cast<float4*>(coef_g + (size_t)gc * 2) =
              make_float4((float)(v0.x / count), (float)(v0.y / count), (float)(v1.x / count), (float)(v1.y / count));
        }
        __threadfence();
        __syncthreads();
        if (t == 0) {
          *reinterpret_cast<volatile unsigned int*>(cnt + 2) = 1u;   // ready: siblings may read coef_g
          if (atomicAdd(sync + 255, 1u) == gridDim.y - 1) {          // last chunk exchange of this launch
            sync[255] = 0u;
            sync_write_seq(sp, sq);
          }
        }
      } else if (t == 0) {
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned int*>(cnt + 2) == 0u) {
          if (clock64() - t0 > 130000000000LL) __trap();
        }
        __threadfence();
      }
      __syncthreads();
      for (int lc = t; lc < chans; lc += blockDim.x) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(coef_g + ((size_t)blockIdx.y * cv * 8 + lc) * 2));
        cf[lc * 2] = v.x;
        cf[lc * 2 + 1] = v.y;
      }
    } else if (t < pairs) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cf[(t * 2 + h) * 2] = (float)(a[2 * h] / count);
        cf[(t * 2 + h) * 2 + 1] = (float)(a[2 * h + 1] / count);
      }
    }
    __syncthreads();
    if (t == 0 && atomicAdd(cnt + 1, 1u) == gridDim.x - 1) {   // last block out: nobody spins or reads any more
      cnt[0] = 0u;
      cnt[1] = 0u;
      cnt[2] = 0u;
    }
    // ---------------- stage 3: dy = sc*dz - k0 - k1*(y - mu)
    if (!active) return;
    float k0[8], k1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k0[j] = sc[j] * cf[(vl * 8 + j) * 2];
      k1[j] = sc[j] * is[j] * cf[(vl * 8 + j) * 2 + 1];
    }
    for (long long pix = (long long)blockIdx.x * ppi + pl; pix < pixels; pix += step * kEwUnroll) {
      uint4 gv[kEwUnroll], yv[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long q = pix + u * step;
        if (q < pixels) {
          gv[u] = __ldg(reinterpret_cast<const uint4*>(dout + q * lddo + ch));
          yv[u] = __ldg(reinterpret_cast<const uint4*>(y + q * ldy + ch));
        }
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        const long long q = pix + u * step;
        if (q < pixels) {
          float g[8], yy[8], o[8];
          unpack8(gv[u], g);
          unpack8(yv[u], yy);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = fmaf(yy[j], sc[j], sh[j]);
            const float dz = (!RELU || z > 0.f) ? g[j] : 0.f;
            o[j] = fmaf(sc[j], dz, -k0[j]) - k1[j] * (yy[j] - mu[j]);
          }
          *reinterpret_cast<uint4*>(dy + q * lddy + ch) = pack8(o);
        }
      }
    }
  }
}

// dst (+)= src   (bf16 views; fan-out copies / fan-in adds of activation gradients)
template <bool ADD>
__global__ void grad_add_kernel(__nv_bfloat16* __restrict__ dst, int ldd, const __nv_bfloat16* __restrict__ src,
                                int lds, long long pixels, int c8) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8) * 8;
    uint4 s = __ldg(reinterpret_cast<const uint4*>(src + pix * lds + ch));
    if (ADD) {
      float a[8], b[8];
      unpack8(s, a);
      unpack8(*reinterpret_cast<const uint4*>(dst + pix * ldd + ch), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      s = pack8(a);
    }
    *reinterpret_cast<uint4*>(dst + pix * ldd + ch) = s;
  }
}

// NCHW fp32 image -> NHWC bf16 with channels zero-padded to cpad (multiple of 8)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int c, long long hw,
                                    __nv_bfloat16* __restrict__ out, int cpad) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * hw;
  const int v8 = cpad / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, p = i - img * hw;
    for (int v = 0; v < v8; ++v) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = v * 8 + j;
        f[j] = ch < c ? __ldg(x + (img * c + ch) * hw + p) : 0.f;
      }
      *reinterpret_cast<uint4*>(out + i * cpad + v * 8) = pack8(f);
    }
  }
}

// Image (NCHW fp32, c <= 4 channels) -> the GATHERED operand of its first convolution:
//   out[n][ho][wo][t * c + ci] = x[n][ci][ho * s - p + kh][wo * s - p + kw],  t = kh * k + kw   (bf16; zero outside the
// image and for columns >= k * k * c).  The first convolution then is a 1x1 GEMM over kp = round_up(k*k*c, 16) columns:
// ONE TMA request per tile instead of one per tap - the few-channel stem is bound by the per-SM im2col request cadence
// (~450 cycles per request whatever its size), not by bytes.  One thread per output pixel, 16-byte stores.
// KS > 0: compile-time kernel size with 3 image channels (fully unrolled: the k*k*3 loads of a pixel are independent and
// all in flight together); KS == 0: runtime k / c.
template <int KS>
__global__ void __launch_bounds__(256)
im2col_input_kernel(const float* __restrict__ x, int n, int c, int h, int w, int k, int stride, int pad, int ho, int wo,
                    __nv_bfloat16* __restrict__ out, int kp) {
  __shared__ uint4 stage[KS > 0 ? 8 * 128 : 1];
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * ho * wo;
  const long long hw = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % wo);
    const long long t0 = i / wo;
    const int oh = (int)(t0 % ho);
    const long long img = t0 / ho;
    const int ih0 = oh * stride - pad, iw0 = ow * stride - pad;
    if (KS > 0) {
      constexpr int KR = (KS > 0 ? KS * KS * 3 : 1), KP = (KR + 15) / 16 * 16;
      const float* xi = x + img * 3 * hw;
      float f[KP];
#pragma unroll
      for (int q = 0; q < KP; ++q) f[q] = 0.f;
#pragma unroll
      for (int kh = 0; kh < KS; ++kh) {
        const int ih = ih0 + kh;
        const bool rok = ih >= 0 && ih < h;
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          const int iw = iw0 + kw;
          if (rok && iw >= 0 && iw < w) {
            const float* src = xi + (long long)ih * w + iw;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) f[(kh * KS + kw) * 3 + ci] = __ldg(src + ci * hw);
          }
        }
      }
      // a warp's 32 pixels are 32 * KP * 2 contiguous bytes of the output: stage the rows in shared memory (XOR-swizzled
      // 16-byte chunks) and write them back with fully coalesced 512-byte store instructions
      constexpr int CH = KP / 8;   // 16-byte chunks per pixel
      const int ln = threadIdx.x & 31, wrp = threadIdx.x >> 5;
      const long long warp_base = i - ln;
      if (CH == 4 && warp_base + 31 < total) {
        uint4* st = stage + wrp * 128;
#pragma unroll
        for (int v = 0; v < CH; ++v) st[ln * 4 + (v ^ ((ln >> 1) & 3))] = pack8(f + v * 8);
        __syncwarp();
        uint4* o16 = reinterpret_cast<uint4*>(out + warp_base * KP);
#pragma unroll
        for (int v = 0; v < CH; ++v) {
          const int q = v * 32 + ln, pix = q >> 2, chunk = q & 3;
          o16[q] = st[pix * 4 + (chunk ^ ((pix >> 1) & 3))];
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int v = 0; v < CH; ++v) *reinterpret_cast<uint4*>(out + i * KP + v * 8) = pack8(f + v * 8);
      }
    } else {
      const int kreal = k * k * c;
      const float* xi = x + img * c * hw;
      int ci = 0, kw = 0, kh = 0;
      for (int v = 0; v < kp / 8; ++v) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float val = 0.f;
          if (v * 8 + j < kreal) {
            const int ih = ih0 + kh, iw = iw0 + kw;
            if (ih >= 0 && ih < h && iw >= 0 && iw < w) val = __ldg(xi + ci * hw + (long long)ih * w + iw);
            if (++ci == c) { ci = 0; if (++kw == k) { kw = 0; ++kh; } }
          }
          f[j] = val;
        }
        *reinterpret_cast<uint4*>(out + i * kp + v * 8) = pack8(f);
      }
    }
  }
}

// Feature-pyramid fuse (reference necks.py:69-79): out = [a +] resize(b) with nn.Upsample(scale_factor=2 | 0.5, "nearest"):
//   up   (top-down):  b is (hb, wb) = (h/2, w/2), out[y][x] reads b[y >> 1][x >> 1]
//   down (bottom-up): b is (hb, wb) with h = hb/2, w = wb/2 (floor), out[y][x] reads b[2y][2x]
// a == nullptr: plain resize (the "concat" fuse writes it into a channel slice).  bf16 in / out, the add in fp32.
template <bool UP, bool HAS_A>
__global__ void __launch_bounds__(256)
resize2_add_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb, int n, int h,
                   int w, int c8, int hb, int wb, __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * h * w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % w), y = (int)((pix / w) % h);
    const long long img = pix / ((long long)w * h);
    const int sy = UP ? (y >> 1) : (y << 1), sx = UP ? (x >> 1) : (x << 1);
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((img * hb + sy) * wb + sx) * ldb + v * 8)), f);
    if (HAS_A) {
      float g[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a + pix * lda + v * 8)), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j];
    }
    *reinterpret_cast<uint4*>(out + pix * ldo + v * 8) = pack8(f);
  }
}

// gradient of the resized operand: gb (+)= resize^T(gout).  up: every b pixel gathers its 2x2 block of gout (fp32 sum,
// one rounding, as upsample_nearest2d_backward does); down: b pixels (2y, 2x) take gout[y][x], all others get zero.
template <bool UP, bool ADD>
__global__ void __launch_bounds__(256)
resize2_add_bwd_kernel(const __nv_bfloat16* __restrict__ gout, int ldg, int n, int h, int w, int c8,
                       __nv_bfloat16* __restrict__ gb, int ldgb, int hb, int wb) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * hb * wb * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % wb), y = (int)((pix / wb) % hb);
    const long long img = pix / ((long long)wb * hb);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (UP) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int oy = 2 * y + dy, ox = 2 * x + dx;
          if (oy < h && ox < w) {
            float g[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(gout + ((img * h + oy) * w + ox) * ldg + v * 8)), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += g[j];
          }
        }
    } else if ((y & 1) == 0 && (x & 1) == 0 && (y >> 1) < h && (x >> 1) < w) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(gout + ((img * h + (y >> 1)) * w + (x >> 1)) * ldg + v * 8)), f);
    }
    __nv_bfloat16* dst = gb + pix * ldgb + v * 8;
    if (ADD) {
      float o[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = rbf(f[j]) + o[j];
    }
    *reinterpret_cast<uint4*>(dst) = pack8(f);
  }
}

// RandomMixup / RandomCutmix of the reference trainer (extras.py:14-109) on the device, decision and parameters read from
// DEVICE memory (no host synchronisation): prm = {mode, lambda, x1, y1, x2, y2}; image i is paired with image i-1 (the
// reference rolls the batch by one).  mode 1: out = x*lambda + x_prev*(1-lambda) with the reference's rounding sequence
// (two products, one sum - no fused multiply-add); mode 2: the box [y1,y2) x [x1,x2) is taken from x_prev; else: copy.
__global__ void __launch_bounds__(256)
mix_images_kernel(const float* __restrict__ x, float* __restrict__ out, int n, long long chw, int h, int w,
                  const float* __restrict__ prm) {
  pdl_wait();
  pdl_trigger();
  const int mode = (int)prm[0];
  const float lam = prm[1], oml = 1.0f - lam;
  const int x1 = (int)prm[2], y1 = (int)prm[3], x2 = (int)prm[4], y2 = (int)prm[5];
  const long long total = (long long)n * chw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / chw, r = i - img * chw;
    const long long j = (img == 0 ? (long long)(n - 1) : img - 1) * chw + r;
    float v = x[i];
    if (mode == 1) {
      v = __fadd_rn(__fmul_rn(v, lam), __fmul_rn(x[j], oml));
    } else if (mode == 2) {
      const int col = (int)(r % w), row = (int)((r / w) % h);
      if (col >= x1 && col < x2 && row >= y1 && row < y2) v = x[j];
    }
    out[i] = v;
  }
}

// weight gradient of the gathered-operand convolution, [cout][k*k*c] in (tap, ci) column order -> OIHW
__global__ void dw_from_col_kernel(const float* __restrict__ dw_col, int cout, int c, int kk, float* __restrict__ dw,
                                   int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int total = cout * c * kk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % kk, ci = (i / kk) % c, co = i / (kk * c);
    const float v = dw_col[(long long)co * kk * c + t * c + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

}  // namespace vtb

using namespace vtb;

#define VIEW_OK(ptr, ld, c) ((ptr) != nullptr && (ld) >= (c) && (ld) % 8 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0)

extern "C" {

int vtb_bn_bwd_rows(long long pixels, int c) {
  if (pixels <= 0 || c <= 0 || c % 8) return fail(VTB_EINVAL, "vtb_bn_bwd_rows: bad arguments");
  return ew_geom(pixels, c / 8, 512, 2).rows;
}

int vtb_bn_stats_reduce(const float* partial, int rows, int c, double* sums, void* stream) {
  if (!partial || !sums || rows <= 0 || c <= 0) return fail(VTB_EINVAL, "vtb_bn_stats_reduce: bad arguments");
  launch_pdl(bn_stats_reduce_kernel, dim3((c * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, partial, rows, c, sums);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_stats_reduce_kernel");
}

int vtb_bn_finalize(const float* partial, int rows, const double* sums, double count, int c, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                    void* stream) {
  if ((partial == nullptr) == (sums == nullptr) || c <= 0 || count <= 0 || !gamma || !beta || !mean || !invstd ||
      !scale || !shift || (partial && rows <= 0) || ((running_mean == nullptr) != (running_var == nullptr)))
    return fail(VTB_EINVAL, "vtb_bn_finalize: bad arguments");
  launch_pdl(bn_finalize_kernel, dim3((c * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, 
      partial, rows, sums, count, c, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, mean,
      invstd, scale, shift);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_finalize_kernel");
}

int vtb_bn_eval_affine(int c, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream) {
  if (c <= 0 || !gamma || !beta || !running_mean || !running_var || !scale || !shift)
    return fail(VTB_EINVAL, "vtb_bn_eval_affine: bad arguments");
  launch_pdl(bn_eval_affine_kernel, dim3((c + 255) / 256), dim3(256), 0, (cudaStream_t)stream, c, gamma, beta, running_mean, running_var,
                                                                           eps, scale, shift);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_eval_affine_kernel");
}

int vtb_bn_act(const void* y, int ldy, long long pixels, int c, const float* scale, const float* shift, int relu,
               const void* residual, int ldr, void* out, int ldo, void* stream) {
  if (pixels <= 0 || c <= 0 || c % 8 || !VIEW_OK(y, ldy, c) || !VIEW_OK(out, ldo, c) || !scale || !shift ||
      (residual && !VIEW_OK(residual, ldr, c)))
    return fail(VTB_EINVAL, "vtb_bn_act: bad arguments");
  const int c8 = c / 8;
  const EwGeom g = ew_geom(pixels, c8, 512, 4);
  const dim3 grid(g.rows, g.chunks);
  const int bs = g.ppi * g.cv;
  const __nv_bfloat16* yy = (const __nv_bfloat16*)y;
  const __nv_bfloat16* rr = (const __nv_bfloat16*)residual;
  __nv_bfloat16* oo = (__nv_bfloat16*)out;
  cudaStream_t st = (cudaStream_t)stream;
  if (relu && rr) launch_pdl(bn_act_kernel<true, true>, dim3(grid), dim3(bs), 0, st, yy, ldy, pixels, c8, g.cv, scale, shift, rr, ldr, oo, ldo);
  else if (relu) launch_pdl(bn_act_kernel<true, false>, dim3(grid), dim3(bs), 0, st, yy, ldy, pixels, c8, g.cv, scale, shift, rr, ldr, oo, ldo);
  else if (rr) launch_pdl(bn_act_kernel<false, true>, dim3(grid), dim3(bs), 0, st, yy, ldy, pixels, c8, g.cv, scale, shift, rr, ldr, oo, ldo);
  else launch_pdl(bn_act_kernel<false, false>, dim3(grid), dim3(bs), 0, st, yy, ldy, pixels, c8, g.cv, scale, shift, rr, ldr, oo, ldo);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_act_kernel");
}

int vtb_bn_bwd_reduce(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c,
                      const float* scale, const float* shift, const float* mean, const float* invstd, int relu,
                      float* partial, void* stream) {
  if (pixels <= 0 || c <= 0 || c % 8 || !VIEW_OK(dout, lddo, c) || !VIEW_OK(y, ldy, c) || !scale || !shift || !mean ||
      !invstd || !partial)
    return fail(VTB_EINVAL, "vtb_bn_bwd_reduce: bad arguments");
  const int c8 = c / 8;
  const EwGeom g = ew_geom(pixels, c8, 512, 2);
  const dim3 grid(g.rows, g.chunks);
  const int bs = g.ppi * g.cv;
  const size_t sm = (size_t)bs * 17 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (relu)
    launch_pdl(bn_bwd_reduce_kernel<true>, dim3(grid), dim3(bs), sm, st, (const __nv_bfloat16*)dout, lddo, (const __nv_bfloat16*)y, ldy,
                                                     pixels, c8, g.cv, scale, shift, mean, invstd, partial, c);
  else
    launch_pdl(bn_bwd_reduce_kernel<false>, dim3(grid), dim3(bs), sm, st, (const __nv_bfloat16*)dout, lddo, (const __nv_bfloat16*)y, ldy,
                                                      pixels, c8, g.cv, scale, shift, mean, invstd, partial, c);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_bwd_reduce_kernel");
}

int vtb_bn_bwd_finalize(const float* partial, int rows, const double* sums, const double* local_sums, double count,
                        int c, float* dgamma, float* dbeta, int accumulate, float* coef, double* sums_out,
                        void* stream) {
  if ((partial == nullptr) == (sums == nullptr) || c <= 0 || count <= 0 || (!coef && !sums_out) ||
      (partial && rows <= 0))
    return fail(VTB_EINVAL, "vtb_bn_bwd_finalize: bad arguments");
  launch_pdl(bn_bwd_finalize_kernel, dim3((c * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, 
      partial, rows, sums, count, c, dgamma, dbeta, accumulate, local_sums, coef, sums_out);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_bwd_finalize_kernel");
}

int vtb_bn_bwd_apply(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int relu, const float* coef,
                     void* dy, int lddy, void* stream) {
  if (pixels <= 0 || c <= 0 || c % 8 || !VIEW_OK(dout, lddo, c) || !VIEW_OK(y, ldy, c) || !VIEW_OK(dy, lddy, c) ||
      !scale || !shift || !mean || !invstd || !coef)
    return fail(VTB_EINVAL, "vtb_bn_bwd_apply: bad arguments");
  const int c8 = c / 8;
  const EwGeom g = ew_geom(pixels, c8, 512, 4);
  const dim3 grid(g.rows, g.chunks);
  const int bs = g.ppi * g.cv;
  cudaStream_t st = (cudaStream_t)stream;
  if (relu)
    launch_pdl(bn_bwd_apply_kernel<true>, dim3(grid), dim3(bs), 0, st, (const __nv_bfloat16*)dout, lddo, (const __nv_bfloat16*)y, ldy,
                                                   pixels, c8, g.cv, scale, shift, mean, invstd, coef,
                                                   (__nv_bfloat16*)dy, lddy);
  else
    launch_pdl(bn_bwd_apply_kernel<false>, dim3(grid), dim3(bs), 0, st, (const __nv_bfloat16*)dout, lddo, (const __nv_bfloat16*)y, ldy,
                                                    pixels, c8, g.cv, scale, shift, mean, invstd, coef,
                                                    (__nv_bfloat16*)dy, lddy);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "bn_bwd_apply_kernel");
}

// geometry of the fused backward: all blocks must be co-resident (cooperative launch), one 512-thread block per SM
static EwGeom bwd_fused_geom(long long pixels, int c8) {
  EwGeom g = ew_geom(pixels, c8, 512, 1);
  // stage 2 reads gridDim.x rows per channel pair with blockDim / pairs row groups: keep that <= ~40 loads per thread
  const int pairs = g.cv * 4;
  const int G = std::max(1, (g.ppi * g.cv) / pairs);
  g.rows = std::max(1, std::min(g.rows, 40 * G));
  return g;
}

int vtb_bn_bwd_fused_rows(long long pixels, int c) {
  if (pixels <= 0 || c <= 0 || c % 16) return fail(VTB_EINVAL, "vtb_bn_bwd_fused_rows: bad arguments");
  return bwd_fused_geom(pixels, c / 8).rows;
}

int vtb_bn_bwd_fused(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int relu, double count,
                     float* partial, float* dgamma, float* dbeta, int accumulate, unsigned int* sync, void* dy,
                     int lddy, const VtbSyncBn* peers, void* stream) {
  if (pixels <= 0 || c <= 0 || c % 16 || !VIEW_OK(dout, lddo, c) || !VIEW_OK(y, ldy, c) || !VIEW_OK(dy, lddy, c) ||
      !scale || !shift || !mean || !invstd || !partial || !sync || count <= 0)
    return fail(VTB_EINVAL, "vtb_bn_bwd_fused: bad arguments");
  int c8 = c / 8;
  EwGeom g = bwd_fused_geom(pixels, c8);
  // SyncBN: every block of this launch stays resident while the chunk's block 0 waits for the peers' sums.  A launch
  // that holds EVERY SM can close a cross-rank wait cycle with a collective on another stream (rank A: this kernel
  // resident, spinning for rank B; rank B: its NCCL kernel resident, waiting for A's NCCL kernel, which finds no SM
  // on A; B's copy of this kernel cannot become co-resident next to B's NCCL blocks) - observed at 8 GPUs when one rank
  // ran ahead of the others.  Leave VTB_SM_RESERVE SMs (default 16; create the NCCL communicator with max_ctas <= that,
  // parallel.nccl_pg_options) to whatever else must be able to start.
  static const int reserve = [] {
    const char* v = getenv("VTB_SM_RESERVE");
    return v ? std::max(0, atoi(v)) : 16;
  }();
  if (peers != nullptr) {
    const int cap = std::max(1, (std::max(1, num_sms()) - reserve) / g.chunks);
    g.rows = std::min(g.rows, cap);
  }
  if (g.chunks > 60) return fail(VTB_EINVAL, "vtb_bn_bwd_fused: too many channels");
  if (peers && !sync_args_ok(peers, c)) return fail(VTB_EINVAL, "vtb_bn_bwd_fused: bad SyncBN peers");
  SyncPeers sp = make_sync_peers(peers);
  int cv = g.cv;
  const dim3 grid(g.rows, g.chunks);
  const dim3 block(g.ppi * g.cv);
  const int chans = g.cv * 8, pairs = chans / 2, G = std::max(1, (int)block.x / pairs);
  const size_t sm = std::max<size_t>((size_t)block.x * 17 * sizeof(float),
                                     (size_t)G * pairs * 4 * sizeof(double) + (size_t)chans * 2 * sizeof(float));
  const void* fn = relu ? (const void*)bn_bwd_fused_kernel<true> : (const void*)bn_bwd_fused_kernel<false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[relu ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return check_cuda((int)e, "bn_bwd_fused_kernel attribute");
    attr_set[relu ? 1 : 0] = true;
  }
  if (sm > 96 * 1024) return fail(VTB_EINVAL, "vtb_bn_bwd_fused: shared memory");
  const __nv_bfloat16* dout_p = (const __nv_bfloat16*)dout;
  const __nv_bfloat16* y_p = (const __nv_bfloat16*)y;
  __nv_bfloat16* dy_p = (__nv_bfloat16*)dy;
  void* args[] = {&dout_p, &lddo, &y_p, &ldy, &pixels, &c8, &cv, &scale, &shift, &mean, &invstd, &partial, &c, &count,
                  &dgamma, &dbeta, &accumulate, &sync, &dy_p, &lddy, &sp};
  count_launch(1);
  // Single-GPU: an ordinary launch with programmatic stream serialization.  The grid is at most one 512-thread block per
  // SM, every kernel that can still be running when these blocks are scheduled (the previous launch of this stream, a
  // weight-gradient GEMM on the side stream) finishes without waiting for this one, so every block becomes resident and
  // the hand-rolled grid barrier cannot deadlock - while the launch itself is cheaper than a cooperative one and its
  // blocks may take SMs as the previous kernel drains.  With SyncBN peers (cross-rank spin inside the kernel, NCCL
  // kernels on another stream) the co-residency GUARANTEE of the cooperative launch is kept; it is also the faster one
  // there (2 GPUs, weight gradients on the side stream: 14.77 ms per step against 14.96 ms with the ordinary launch,
  // whose early-resident blocks spin on SMs the weight-gradient CTAs could use; profiles/r02_bench_2gpu_t11*.json).
  // VTB_BWD_COOP=1 forces the cooperative launch everywhere, VTB_BWD_COOP=0 the ordinary one (needs the SM reserve).
  static const int coop_env = getenv("VTB_BWD_COOP") ? atoi(getenv("VTB_BWD_COOP")) : -1;
  const bool coop = coop_env == 1 || (peers != nullptr && !(coop_env == 0 && reserve > 0));
  if (!coop) {
    cudaError_t e;
    if (relu)
      e = launch_pdl(bn_bwd_fused_kernel<true>, grid, block, sm, (cudaStream_t)stream, dout_p, lddo, y_p, ldy, pixels, c8, cv,
                     scale, shift, mean, invstd, partial, c, count, dgamma, dbeta, accumulate, sync, dy_p, lddy, sp);
    else
      e = launch_pdl(bn_bwd_fused_kernel<false>, grid, block, sm, (cudaStream_t)stream, dout_p, lddo, y_p, ldy, pixels, c8, cv,
                     scale, shift, mean, invstd, partial, c, count, dgamma, dbeta, accumulate, sync, dy_p, lddy, sp);
    return check_cuda((int)e, "bn_bwd_fused_kernel");
  }
  return check_cuda((int)cudaLaunchCooperativeKernel(fn, grid, block, args, sm, (cudaStream_t)stream),
                    "bn_bwd_fused_kernel");
}

int vtb_grad_add(void* dst, int ldd, const void* src, int lds, long long pixels, int c, int accumulate, void* stream) {
  if (pixels <= 0 || c <= 0 || c % 8 || !VIEW_OK(dst, ldd, c) || !VIEW_OK(src, lds, c))
    return fail(VTB_EINVAL, "vtb_grad_add: bad arguments");
  const int c8 = c / 8;
  const int grid = ew_grid(pixels * c8, 256);
  if (accumulate)
    launch_pdl(grad_add_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (__nv_bfloat16*)dst, ldd, (const __nv_bfloat16*)src,
                                                                  lds, pixels, c8);
  else
    launch_pdl(grad_add_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (__nv_bfloat16*)dst, ldd, (const __nv_bfloat16*)src,
                                                                   lds, pixels, c8);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "grad_add_kernel");
}

int vtb_nchw_to_nhwc(const float* x, int n, int c, int h, int w, void* out, int cpad, void* stream) {
  if (!x || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || cpad < c || cpad % 8)
    return fail(VTB_EINVAL, "vtb_nchw_to_nhwc: bad arguments");
  const long long total = (long long)n * h * w;
  launch_pdl(nchw_to_nhwc_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, n, c, (long long)h * w,
                                                                             (__nv_bfloat16*)out, cpad);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "nchw_to_nhwc_kernel");
}

int vtb_im2col_input(const float* x, int n, int c, int h, int w, int k, int stride, int pad, void* out, int kp,
                     void* stream) {
  if (!x || !out || n <= 0 || c <= 0 || c > 4 || h <= 0 || w <= 0 || k <= 0 || stride <= 0 || pad < 0 || kp % 8 ||
      kp < k * k * c || (reinterpret_cast<uintptr_t>(out) & 15) || h + 2 * pad < k || w + 2 * pad < k)
    return fail(VTB_EINVAL, "vtb_im2col_input: bad arguments");
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  const long long total = (long long)n * ho * wo;
  const dim3 grid(ew_grid(total, 256));
  if (c == 3 && k == 3 && kp == 32)
    launch_pdl(im2col_input_kernel<3>, grid, dim3(256), 0, (cudaStream_t)stream, x, n, c, h, w, k, stride, pad, ho, wo,
               (__nv_bfloat16*)out, kp);
  else
    launch_pdl(im2col_input_kernel<0>, grid, dim3(256), 0, (cudaStream_t)stream, x, n, c, h, w, k, stride, pad, ho, wo,
               (__nv_bfloat16*)out, kp);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "im2col_input_kernel");
}

static bool resize_ok(int h, int w, int hb, int wb, int up) {
  return up ? (h == 2 * hb && w == 2 * wb) : (h == hb / 2 && w == wb / 2 && h > 0 && w > 0);
}

int vtb_resize2_add(const void* a, int lda, const void* b, int ldb, int n, int h, int w, int c, int hb, int wb, int up,
                    void* out, int ldo, void* stream) {
  if (n <= 0 || c <= 0 || c % 8 || !resize_ok(h, w, hb, wb, up) || !VIEW_OK(b, ldb, c) || !VIEW_OK(out, ldo, c) ||
      (a && !VIEW_OK(a, lda, c)))
    return fail(VTB_EINVAL, "vtb_resize2_add: bad arguments (nearest x2 needs h == 2*hb, x0.5 needs h == hb/2)");
  const int c8 = c / 8;
  const dim3 grid(ew_grid((long long)n * h * w * c8, 256));
  const __nv_bfloat16 *aa = (const __nv_bfloat16*)a, *bb = (const __nv_bfloat16*)b;
  __nv_bfloat16* oo = (__nv_bfloat16*)out;
  cudaStream_t st = (cudaStream_t)stream;
  if (up && a) launch_pdl(resize2_add_kernel<true, true>, grid, dim3(256), 0, st, aa, lda, bb, ldb, n, h, w, c8, hb, wb, oo, ldo);
  else if (up) launch_pdl(resize2_add_kernel<true, false>, grid, dim3(256), 0, st, aa, lda, bb, ldb, n, h, w, c8, hb, wb, oo, ldo);
  else if (a) launch_pdl(resize2_add_kernel<false, true>, grid, dim3(256), 0, st, aa, lda, bb, ldb, n, h, w, c8, hb, wb, oo, ldo);
  else launch_pdl(resize2_add_kernel<false, false>, grid, dim3(256), 0, st, aa, lda, bb, ldb, n, h, w, c8, hb, wb, oo, ldo);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "resize2_add_kernel");
}

int vtb_resize2_add_bwd(const void* gout, int ldg, int n, int h, int w, int c, void* gb, int ldgb, int hb, int wb, int up,
                        int accumulate, void* stream) {
  if (n <= 0 || c <= 0 || c % 8 || !resize_ok(h, w, hb, wb, up) || !VIEW_OK(gout, ldg, c) || !VIEW_OK(gb, ldgb, c))
    return fail(VTB_EINVAL, "vtb_resize2_add_bwd: bad arguments");
  const int c8 = c / 8;
  const dim3 grid(ew_grid((long long)n * hb * wb * c8, 256));
  const __nv_bfloat16* gg = (const __nv_bfloat16*)gout;
  __nv_bfloat16* bb = (__nv_bfloat16*)gb;
  cudaStream_t st = (cudaStream_t)stream;
  if (up && accumulate) launch_pdl(resize2_add_bwd_kernel<true, true>, grid, dim3(256), 0, st, gg, ldg, n, h, w, c8, bb, ldgb, hb, wb);
  else if (up) launch_pdl(resize2_add_bwd_kernel<true, false>, grid, dim3(256), 0, st, gg, ldg, n, h, w, c8, bb, ldgb, hb, wb);
  else if (accumulate) launch_pdl(resize2_add_bwd_kernel<false, true>, grid, dim3(256), 0, st, gg, ldg, n, h, w, c8, bb, ldgb, hb, wb);
  else launch_pdl(resize2_add_bwd_kernel<false, false>, grid, dim3(256), 0, st, gg, ldg, n, h, w, c8, bb, ldgb, hb, wb);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "resize2_add_bwd_kernel");
}

int vtb_mix_images(const float* x, float* out, int n, int c, int h, int w, const float* params_device, void* stream) {
  if (!x || !out || x == out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || !params_device)
    return fail(VTB_EINVAL, "vtb_mix_images: bad arguments (out of place only)");
  const long long chw = (long long)c * h * w;
  launch_pdl(mix_images_kernel, dim3(ew_grid((long long)n * chw, 256)), dim3(256), 0, (cudaStream_t)stream, x, out, n, chw, h, w,
             params_device);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "mix_images_kernel");
}

int vtb_dw_from_col(const float* dw_col, int cout, int c, int kk, float* dw_oihw, int accumulate, void* stream) {
  if (!dw_col || !dw_oihw || cout <= 0 || c <= 0 || kk <= 0) return fail(VTB_EINVAL, "vtb_dw_from_col: bad arguments");
  launch_pdl(dw_from_col_kernel, dim3((cout * c * kk + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dw_col, cout, c, kk,
             dw_oihw, accumulate);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "dw_from_col_kernel");
}

}  // extern "C"
