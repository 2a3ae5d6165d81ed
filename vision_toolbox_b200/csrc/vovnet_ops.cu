// VoVNet-only HBM-bound ops: MaxPool2d(3,2,1) (vovnet.py:94) and the eSE channel gate (vovnet.py:20-28),
// forward and backward, on NHWC bf16 views.
#include <algorithm>
#include <cstdint>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/vtb.h"
#include "common.cuh"

namespace vtb {

__device__ __forceinline__ float v_lo16(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float v_hi16(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t v_pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float v_rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void v_unpack8(const uint4& u, float* f) {
  f[0] = v_lo16(u.x); f[1] = v_hi16(u.x); f[2] = v_lo16(u.y); f[3] = v_hi16(u.y);
  f[4] = v_lo16(u.z); f[5] = v_hi16(u.z); f[6] = v_lo16(u.w); f[7] = v_hi16(u.w);
}
__device__ __forceinline__ uint4 v_pack8(const float* f) {
  uint4 u;
  u.x = v_pack2(f[0], f[1]); u.y = v_pack2(f[2], f[3]); u.z = v_pack2(f[4], f[5]); u.w = v_pack2(f[6], f[7]);
  return u;
}

static int vgrid(long long work, int block) {
  const int sms = std::max(1, num_sms());
  return (int)std::max<long long>(1, std::min<long long>((work + block - 1) / block, (long long)sms * 16));
}

// ------------------------------------------------------------------ MaxPool2d(k=3, s=2, p=1)
// idx (optional): per output element the window position r*3+s of its FIRST maximum (uint8, [pixel][c]) - what
// aten::max_pool2d_with_indices saves for backward
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int c8, int ho,
                                   int wo, __nv_bfloat16* __restrict__ out, int ldo, unsigned char* __restrict__ idx) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * ho * wo * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long t = i / c8;
    const int ow = (int)(t % wo);
    t /= wo;
    const int oh = (int)(t % ho);
    const int img = (int)(t / ho);
    float m[8];
    unsigned int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; arg[j] = 255u; }
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= h) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= w) continue;
        float f[8];
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)img * h + ih) * w + iw) * ldx + v * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > m[j] || f[j] != f[j]) { m[j] = f[j]; arg[j] = r * 3 + s; }
      }
    }
    const long long opix = ((long long)img * ho + oh) * wo + ow;
    *reinterpret_cast<uint4*>(out + opix * ldo + v * 8) = v_pack8(m);
    if (idx != nullptr) {
      uint2 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(idx + (opix * c8 + v) * 8) = pk;
    }
  }
}

// dx[pixel] (+)= sum over the <=4 windows containing it of dout[window] * [argmax(window) == pixel]
// (first maximum in row-major window order wins ties, as in aten max_pool2d_with_indices)
template <bool ADD>
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int c8, int ho,
                                   int wo, const __nv_bfloat16* __restrict__ dout, int lddo,
                                   __nv_bfloat16* __restrict__ dx, int lddx) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * h * w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long t = i / c8;
    const int iw0 = (int)(t % w);
    t /= w;
    const int ih0 = (int)(t % h);
    const int img = (int)(t / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oh_lo = max(0, (ih0) / 2), oh_hi = min(ho - 1, (ih0 + 1) / 2);
    const int ow_lo = max(0, (iw0) / 2), ow_hi = min(wo - 1, (iw0 + 1) / 2);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        // is (ih0, iw0) inside this window?
        const int r0 = ih0 - (oh * 2 - 1), s0 = iw0 - (ow * 2 - 1);
        if (r0 < 0 || r0 > 2 || s0 < 0 || s0 > 2) continue;
        float m[8];
        int arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; arg[j] = -1; }
        for (int r = 0; r < 3; ++r) {
          const int ih = oh * 2 - 1 + r;
          if (ih < 0 || ih >= h) continue;
          for (int s = 0; s < 3; ++s) {
            const int iw = ow * 2 - 1 + s;
            if (iw < 0 || iw >= w) continue;
            float f[8];
            v_unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)img * h + ih) * w + iw) * ldx + v * 8)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (f[j] > m[j] || f[j] != f[j]) { m[j] = f[j]; arg[j] = r * 3 + s; }
          }
        }
        float g[8];
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(dout + (((long long)img * ho + oh) * wo + ow) * lddo + v * 8)), g);
        const int me = r0 * 3 + s0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (arg[j] == me) acc[j] += g[j];
      }
    __nv_bfloat16* dst = dx + (((long long)img * h + ih0) * w + iw0) * lddx + v * 8;
    if (ADD) {
      float o[8];
      v_unpack8(*reinterpret_cast<const uint4*>(dst), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = v_rbf(acc[j]) + o[j];
    }
    *reinterpret_cast<uint4*>(dst) = v_pack8(acc);
  }
}

// same result from the saved argmax indices: <= 4 index + gradient loads per input pixel instead of 36 input loads
template <bool ADD>
__global__ void maxpool_bwd_idx_kernel(const unsigned char* __restrict__ idx, int n, int h, int w, int c8, int ho, int wo,
                                       const __nv_bfloat16* __restrict__ dout, int lddo,
                                       __nv_bfloat16* __restrict__ dx, int lddx) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * h * w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % c8);
    long long t = i / c8;
    const int iw0 = (int)(t % w);
    t /= w;
    const int ih0 = (int)(t % h);
    const int img = (int)(t / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oh_lo = ih0 / 2, oh_hi = min(ho - 1, (ih0 + 1) / 2);
    const int ow_lo = iw0 / 2, ow_hi = min(wo - 1, (iw0 + 1) / 2);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int r0 = ih0 - (oh * 2 - 1), s0 = iw0 - (ow * 2 - 1);
        if (r0 < 0 || r0 > 2 || s0 < 0 || s0 > 2) continue;
        const unsigned int me = r0 * 3 + s0;
        const long long opix = ((long long)img * ho + oh) * wo + ow;
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + (opix * c8 + v) * 8));
        float g[8];
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(dout + opix * lddo + v * 8)), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned int a = ((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 255u;
          if (a == me) acc[j] += g[j];
        }
      }
    __nv_bfloat16* dst = dx + (((long long)img * h + ih0) * w + iw0) * lddx + v * 8;
    if (ADD) {
      float o[8];
      v_unpack8(*reinterpret_cast<const uint4*>(dst), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = v_rbf(acc[j]) + o[j];
    }
    *reinterpret_cast<uint4*>(dst) = v_pack8(acc);
  }
}

// ------------------------------------------------------------------ small fp32 GEMM for the eSE 1x1 "conv" on (N,C,1,1)
// C[i][j] = sum_k A(i,k) * B(k,j), A(i,k) = a[i*sai + k*sak], B(k,j) = b[k*sbk + j*sbj]; operands optionally rounded to
// bf16 first (autocast runs nn.Conv2d in bf16).  32x32 tiles, 16-deep k steps, 256 threads x (2x2) results.
// EPI 0: plain store (scaled by alpha); 1: z / gate epilogue of the eSE forward; 2: accumulate-or-store (dW);
// 3: acc * alpha + bias[j] (the classifier head's nn.Linear)
template <bool RA, bool RB, int EPI>
__global__ void __launch_bounds__(256)
small_gemm_kernel(const float* __restrict__ a, long long sai, long long sak, const float* __restrict__ b, long long sbk,
                  long long sbj, int M, int N, int K, float alpha, float* __restrict__ out, long long ldc,
                  const float* __restrict__ bias, float* __restrict__ out2, int accumulate) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[16][33], Bs[16][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 512; e += 256) {
      // choose the fast index so that global reads are as contiguous as the operand allows
      int kk, ii;
      if (sak == 1) { kk = e & 15; ii = e >> 4; } else { ii = e & 31; kk = e >> 5; }
      const int gi = i0 + ii, gk = k0 + kk;
      float v = (gi < M && gk < K) ? __ldg(a + gi * sai + gk * sak) : 0.f;
      As[kk][ii] = RA ? v_rbf(v) : v;
      int kb, jj;
      if (sbk == 1) { kb = e & 15; jj = e >> 4; } else { jj = e & 31; kb = e >> 5; }
      const int gj = j0 + jj, gkb = k0 + kb;
      float u = (gj < N && gkb < K) ? __ldg(b + gkb * sbk + gj * sbj) : 0.f;
      Bs[kb][jj] = RB ? v_rbf(u) : u;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
      const float b0 = Bs[kk][tx * 2], b1 = Bs[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int gi = i0 + ty * 2 + u, gj = j0 + tx * 2 + w;
      if (gi >= M || gj >= N) continue;
      float* dst = out + gi * ldc + gj;
      if (EPI == 0) {
        *dst = acc[u][w] * alpha;
      } else if (EPI == 1) {
        const float zz = v_rbf(acc[u][w] + v_rbf(bias[gj]));
        *dst = zz;
        out2[gi * ldc + gj] = v_rbf(fminf(fmaxf(zz * (1.f / 6.f) + 0.5f, 0.f), 1.f));
      } else if (EPI == 3) {
        *dst = fmaf(acc[u][w], alpha, bias[gj]);
      } else {
        *dst = accumulate ? *dst + acc[u][w] : acc[u][w];
      }
    }
}
// ------------------------------------------------------------------ fp32 GEMM for the classifier head
// C[i][j] = alpha * sum_k A(i,k) * B(k,j) with strided operands as above; BM x 64 tiles (BM = 32 | 64), 16-deep k steps,
// 256 threads x (BM/16 x 4) results: 8 | 16 FMAs per pair of 16-byte shared-memory loads (the 32x32 kernel above issues
// one load per FMA), next k step's operands prefetched into registers while the current one is multiplied.
// alpha_dev (optional): device scalar multiplied into alpha (the upstream gradient of the loss).
// EPI 0: store; 2: accumulate-or-store; 3: + bias[j]
template <int BM, int EPI>
__global__ void __launch_bounds__(256)
tile_gemm_kernel(const float* __restrict__ a, long long sai, long long sak, const float* __restrict__ b, long long sbk,
                 long long sbj, int M, int N, int K, float alpha, const float* __restrict__ alpha_dev,
                 float* __restrict__ out, long long ldc, const float* __restrict__ bias, int accumulate) {
  constexpr int BN = 64, BK = 16, RM = BM / 16, LA = BM * BK / 256, LB = BN * BK / 256;
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  float ra[LA], rb[LB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int e = t + u * 256;
      const int kk = a_kfast ? (e % BK) : (e / BM), ii = a_kfast ? (e / BK) : (e % BM);
      const int gi = i0 + ii, gk = k0 + kk;
      ra[u] = (gi < M && gk < K) ? __ldg(a + gi * sai + gk * sak) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int e = t + u * 256;
      const int kk = b_kfast ? (e % BK) : (e / BN), jj = b_kfast ? (e / BK) : (e % BN);
      const int gj = j0 + jj, gk = k0 + kk;
      rb[u] = (gj < N && gk < K) ? __ldg(b + gk * sbk + gj * sbj) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int e = t + u * 256;
      const int kk = a_kfast ? (e % BK) : (e / BM), ii = a_kfast ? (e / BK) : (e % BM);
      As[kk][ii] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int e = t + u * 256;
      const int kk = b_kfast ? (e % BK) : (e / BN), jj = b_kfast ? (e / BK) : (e % BN);
      Bs[kk][jj] = rb[u];
    }
  };
  float acc[RM][4];
#pragma unroll
  for (int u = 0; u < RM; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) acc[u][w] = 0.f;
  fetch(0);
  stash();
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[RM];
      if constexpr (RM == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        av[0] = a4.x; av[1] = a4.y; av[RM - 2] = a4.z; av[RM - 1] = a4.w;
      } else {
        const float2 a2 = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
        av[0] = a2.x; av[1] = a2.y;
      }
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
      for (int u = 0; u < RM; ++u) {
        acc[u][0] = fmaf(av[u], b4.x, acc[u][0]);
        acc[u][1] = fmaf(av[u], b4.y, acc[u][1]);
        acc[u][2] = fmaf(av[u], b4.z, acc[u][2]);
        acc[u][3] = fmaf(av[u], b4.w, acc[u][3]);
      }
    }
    __syncthreads();
    if (more) {
      stash();
      __syncthreads();
    }
  }
  const float al = alpha * (alpha_dev ? __ldg(alpha_dev) : 1.f);
#pragma unroll
  for (int u = 0; u < RM; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int gi = i0 + ty * RM + u, gj = j0 + tx * 4 + w;
      if (gi >= M || gj >= N) continue;
      float* dst = out + gi * ldc + gj;
      const float v = acc[u][w] * al;
      if (EPI == 3) *dst = v + bias[gj];
      else if (EPI == 2) *dst = accumulate ? *dst + v : v;
      else *dst = v;
    }
}
// db[j] (+)= (*gscale) * sum_i dl[i][j]   (classifier head bias gradient)
__global__ void head_dbias_kernel(const float* __restrict__ dl, int n, int k, const float* __restrict__ gscale,
                                  float* __restrict__ db, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += dl[(long long)i * k + j];
  acc *= gscale ? __ldg(gscale) : 1.f;
  db[j] = accumulate ? db[j] + acc : acc;
}
// db[co] (+)= sum_n dz[n][co]
__global__ void ese_dbias_kernel(const float* __restrict__ dz, int n, int c, float* __restrict__ db, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= c) return;
  float acc = 0.f;
  for (int img = 0; img < n; ++img) acc += dz[(long long)img * c + co];
  db[co] = accumulate ? db[co] + acc : acc;
}

// ------------------------------------------------------------------ eSE
// per-(image, channel) sums over H*W of a (optionally a*b) : block = 32 channel-vectors x 8 pixel lanes
template <bool PRODUCT>
__global__ void hw_reduce_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b,
                                 int ldb, int hw, int c8, float* __restrict__ out, int c, float mul) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int v = blockIdx.x * 32 + tx;
  const int img = blockIdx.y;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (v < c8) {
    for (int p = ty; p < hw; p += 8) {
      float f[8];
      v_unpack8(__ldg(reinterpret_cast<const uint4*>(a + ((long long)img * hw + p) * lda + v * 8)), f);
      if (PRODUCT) {
        float g[8];
        v_unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((long long)img * hw + p) * ldb + v * 8)), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = fmaf(f[j], g[j], s[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx][j] = s[j];
  __syncthreads();
  if (ty == 0 && v < c8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
      for (int q = 0; q < 8; ++q) acc += red[q][tx][j];
      out[(long long)img * c + v * 8 + j] = acc * mul;
    }
  }
}

// z[n][co] = bf16(sum_ci bf16(W[co][ci]) * bf16(pool[n][ci]) + bf16(b[co])); gate = bf16(hardsigmoid(z))
// one warp per (n, co); mirrors the bf16 autocast rounding points of nn.Conv2d(C, C, 1) + nn.Hardsigmoid
__global__ void ese_fc_fwd_kernel(const float* __restrict__ pool, const float* __restrict__ W,
                                  const float* __restrict__ bias, int n, int c, float* __restrict__ z,
                                  float* __restrict__ gate) {
  pdl_wait();
  pdl_trigger();
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)n * c) return;
  const int img = (int)(wid / c), co = (int)(wid % c);
  float acc = 0.f;
  for (int ci = lane; ci < c; ci += 32) acc = fmaf(v_rbf(W[(long long)co * c + ci]), v_rbf(pool[(long long)img * c + ci]), acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float zz = v_rbf(acc + v_rbf(bias[co]));
    z[wid] = zz;
    gate[wid] = v_rbf(fminf(fmaxf(zz * (1.f / 6.f) + 0.5f, 0.f), 1.f));
  }
}

// out = x * gate[n][c] (+ residual)
template <bool RES>
__global__ void ese_scale_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gate, int hw,
                                 long long pixels, int c8, const __nv_bfloat16* __restrict__ res, int ldr,
                                 __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c8;
  const int c = c8 * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8) * 8;
    const long long img = pix / hw;
    float f[8];
    v_unpack8(__ldg(reinterpret_cast<const uint4*>(x + pix * ldx + ch)), f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + img * c + ch));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + img * c + ch) + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= g[j];
    if (RES) {
      float r[8];
      v_unpack8(__ldg(reinterpret_cast<const uint4*>(res + pix * ldr + ch)), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = v_rbf(f[j]) + r[j];
    }
    *reinterpret_cast<uint4*>(out + pix * ldo + ch) = v_pack8(f);
  }
}

// dz[n][co] = dgate * hardsigmoid'(z)
__global__ void ese_dz_kernel(const float* __restrict__ dgate, const float* __restrict__ z, long long total,
                              float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float t = z[i] * (1.f / 6.f) + 0.5f;
  dz[i] = (t > 0.f && t < 1.f) ? dgate[i] * (1.f / 6.f) : 0.f;
}
// dpool[n][ci] = sum_co dz[n][co] * bf16(W[co][ci]) ; scaled by 1/hw for the mean
__global__ void ese_dpool_kernel(const float* __restrict__ dz, const float* __restrict__ W, int n, int c, float inv_hw,
                                 float* __restrict__ dpool) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * c) return;
  const int img = (int)(i / c), ci = (int)(i % c);
  float acc = 0.f;
  for (int co = 0; co < c; ++co) acc = fmaf(dz[(long long)img * c + co], v_rbf(W[(long long)co * c + ci]), acc);
  dpool[i] = acc * inv_hw;
}
// dW[co][ci] (+)= sum_n dz[n][co]*bf16(pool[n][ci]); db[co] (+)= sum_n dz[n][co]
__global__ void ese_dw_kernel(const float* __restrict__ dz, const float* __restrict__ pool, int n, int c,
                              float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)c * c) return;
  const int co = (int)(i / c), ci = (int)(i % c);
  float acc = 0.f, accb = 0.f;
  for (int img = 0; img < n; ++img) {
    const float d = dz[(long long)img * c + co];
    acc = fmaf(d, v_rbf(pool[(long long)img * c + ci]), acc);
    accb += d;
  }
  dW[i] = accumulate ? dW[i] + acc : acc;
  if (ci == 0) db[co] = accumulate ? db[co] + accb : accb;
}
// dx (+)= dout * gate + dpool[n][c]
template <bool ADD>
__global__ void ese_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dout, int lddo, const float* __restrict__ gate,
                                  const float* __restrict__ dpool, int hw, long long pixels, int c8,
                                  __nv_bfloat16* __restrict__ dx, int lddx) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c8;
  const int c = c8 * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8) * 8;
    const long long img = pix / hw;
    float f[8];
    v_unpack8(__ldg(reinterpret_cast<const uint4*>(dout + pix * lddo + ch)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], gate[img * c + ch + j], dpool[img * c + ch + j]);
    __nv_bfloat16* dst = dx + pix * lddx + ch;
    if (ADD) {
      float o[8];
      v_unpack8(*reinterpret_cast<const uint4*>(dst), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = v_rbf(f[j]) + o[j];
    }
    *reinterpret_cast<uint4*>(dst) = v_pack8(f);
  }
}

// ------------------------------------------------------------------ classifier head (reference classifier.py:59-64, 92)
// AdaptiveAvgPool2d(1) + Flatten + Linear(C, K) + cross_entropy(label_smoothing) on the last feature map.
// One block per sample: log-sum-exp over the K logits (fixed reduction tree -> deterministic), the sample's loss
//   (1 - eps) * (lse - z_y) + eps * (lse - mean_j z_j)          [torch.nn.functional.cross_entropy, label_smoothing = eps]
// and d(mean loss) / d logits = (softmax - (1 - eps) * onehot - eps / K) / N.
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();   // sh may still be read by the previous reduction
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, sh[w]) : r + sh[w];
  return r;
}

__global__ void __launch_bounds__(256)
head_ce_kernel(const float* __restrict__ logits, int k, const long long* __restrict__ labels, float eps, float inv_n,
               float* __restrict__ dlogits, float* __restrict__ row_loss) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  const int i = blockIdx.x;
  const float* z = logits + (long long)i * k;
  float m = -INFINITY, sz = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float v = z[j];
    m = fmaxf(m, v);
    sz += v;
  }
  m = block_reduce(m, sh, true);
  sz = block_reduce(sz, sh, false);
  float se = 0.f;
  for (int j = threadIdx.x; j < k; j += blockDim.x) se += expf(z[j] - m);
  se = block_reduce(se, sh, false);
  const float lse = m + logf(se);
  const int y = (int)labels[i];
  const float on = 1.f - eps, off = eps / (float)k;
  for (int j = threadIdx.x; j < k; j += blockDim.x)
    dlogits[(long long)i * k + j] = (expf(z[j] - lse) - (j == y ? on : 0.f) - off) * inv_n;
  if (threadIdx.x == 0) row_loss[i] = on * (lse - z[y]) + eps * (lse - sz / (float)k);
}

// loss = mean of the per-sample losses (one block, fixed order)
__global__ void __launch_bounds__(256) head_loss_mean_kernel(const float* __restrict__ row_loss, int n, float* __restrict__ loss) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += row_loss[i];
  s = block_reduce(s, sh, false);
  if (threadIdx.x == 0) *loss = s / (float)n;
}

// df[n][pixel][c] = dpooled[n][c]   (gradient of the spatial mean; dpooled already carries the 1/hw factor)
__global__ void head_broadcast_kernel(const float* __restrict__ dpooled, int hw, long long pixels, int c8,
                                      __nv_bfloat16* __restrict__ df, int lddf) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c8;
  const int c = c8 * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / c8;
    const int ch = (int)(i - pix * c8) * 8;
    const long long img = pix / hw;
    const float4 a = __ldg(reinterpret_cast<const float4*>(dpooled + img * c + ch));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dpooled + img * c + ch) + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(df + pix * lddf + ch) = v_pack8(f);
  }
}

}  // namespace vtb

using namespace vtb;
#define VVIEW_OK(ptr, ld, c) ((ptr) != nullptr && (ld) >= (c) && (ld) % 8 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0)

extern "C" {

int vtb_maxpool3s2_fwd(const void* x, int ldx, int n, int h, int w, int c, void* out, int ldo, void* idx, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || !VVIEW_OK(x, ldx, c) || !VVIEW_OK(out, ldo, c))
    return fail(VTB_EINVAL, "vtb_maxpool3s2_fwd: bad arguments");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  launch_pdl(maxpool_fwd_kernel, dim3(vgrid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ldx, n, h, w, c / 8,
                                                                          ho, wo, (__nv_bfloat16*)out, ldo, (unsigned char*)idx);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "maxpool_fwd_kernel");
}

int vtb_maxpool3s2_bwd(const void* x, int ldx, int n, int h, int w, int c, const void* dout, int lddo, void* dx,
                       int lddx, int accumulate, const void* idx, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || (!idx && !VVIEW_OK(x, ldx, c)) || !VVIEW_OK(dout, lddo, c) ||
      !VVIEW_OK(dx, lddx, c))
    return fail(VTB_EINVAL, "vtb_maxpool3s2_bwd: bad arguments");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * h * w * (c / 8);
  if (idx) {
    if (accumulate)
      launch_pdl(maxpool_bwd_idx_kernel<true>, dim3(vgrid(total, 256)), dim3(256), 0, (cudaStream_t)stream,
                 (const unsigned char*)idx, n, h, w, c / 8, ho, wo, (const __nv_bfloat16*)dout, lddo, (__nv_bfloat16*)dx, lddx);
    else
      launch_pdl(maxpool_bwd_idx_kernel<false>, dim3(vgrid(total, 256)), dim3(256), 0, (cudaStream_t)stream,
                 (const unsigned char*)idx, n, h, w, c / 8, ho, wo, (const __nv_bfloat16*)dout, lddo, (__nv_bfloat16*)dx, lddx);
    count_launch(1);
    return check_cuda((int)cudaGetLastError(), "maxpool_bwd_idx_kernel");
  }
  if (accumulate)
    launch_pdl(maxpool_bwd_kernel<true>, dim3(vgrid(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, ldx, n, h, w, c / 8, ho, wo, (const __nv_bfloat16*)dout, lddo, (__nv_bfloat16*)dx, lddx);
  else
    launch_pdl(maxpool_bwd_kernel<false>, dim3(vgrid(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, ldx, n, h, w, c / 8, ho, wo, (const __nv_bfloat16*)dout, lddo, (__nv_bfloat16*)dx, lddx);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "maxpool_bwd_kernel");
}

int vtb_ese_fwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* bias,
                const void* residual, int ldr, void* out, int ldo, float* pool, float* z, float* gate, void* stream) {
  if (n <= 0 || hw <= 0 || c <= 0 || c % 8 || !VVIEW_OK(x, ldx, c) || !VVIEW_OK(out, ldo, c) || !weight || !bias ||
      !pool || !z || !gate || (residual && !VVIEW_OK(residual, ldr, c)))
    return fail(VTB_EINVAL, "vtb_ese_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int c8 = c / 8;
  launch_pdl(hw_reduce_kernel<false>, dim3(dim3((c8 + 31) / 32, n)), dim3(256), 0, st, (const __nv_bfloat16*)x, ldx, nullptr, 0, hw, c8,
                                                                   pool, c, 1.f / hw);
  // z[n][co] = bf16(sum_ci bf16(pool[n][ci]) * bf16(W[co][ci]) + bf16(b[co])), gate = bf16(hardsigmoid(z))
  launch_pdl(small_gemm_kernel<true, true, 1>, dim3((c + 31) / 32, (n + 31) / 32), dim3(256), 0, st, (const float*)pool,
             (long long)c, 1LL, weight, 1LL, (long long)c, n, c, c, 1.f, z, (long long)c, bias, gate, 0);
  const long long pixels = (long long)n * hw;
  if (residual)
    launch_pdl(ese_scale_kernel<true>, dim3(vgrid(pixels * c8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)x, ldx, gate, hw, pixels, c8,
                                                                    (const __nv_bfloat16*)residual, ldr,
                                                                    (__nv_bfloat16*)out, ldo);
  else
    launch_pdl(ese_scale_kernel<false>, dim3(vgrid(pixels * c8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)x, ldx, gate, hw, pixels, c8,
                                                                     nullptr, 0, (__nv_bfloat16*)out, ldo);
  count_launch(3);
  return check_cuda((int)cudaGetLastError(), "ese forward kernels");
}

int vtb_ese_bwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* pool, const float* z,
                const float* gate, const void* dout, int lddo, void* dx, int lddx, int accumulate_dx, float* dweight,
                float* dbias, int accumulate_dw, float* scratch, void* stream) {
  if (n <= 0 || hw <= 0 || c <= 0 || c % 8 || !VVIEW_OK(x, ldx, c) || !VVIEW_OK(dout, lddo, c) ||
      !VVIEW_OK(dx, lddx, c) || !weight || !pool || !z || !gate || !dweight || !dbias || !scratch)
    return fail(VTB_EINVAL, "vtb_ese_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int c8 = c / 8;
  const long long nc = (long long)n * c;
  float* dgate = scratch;          // [n][c]
  float* dz = scratch + nc;        // [n][c]
  float* dpool = scratch + 2 * nc; // [n][c]
  launch_pdl(hw_reduce_kernel<true>, dim3(dim3((c8 + 31) / 32, n)), dim3(256), 0, st, (const __nv_bfloat16*)dout, lddo,
                                                                  (const __nv_bfloat16*)x, ldx, hw, c8, dgate, c, 1.f);
  launch_pdl(ese_dz_kernel, dim3((unsigned)((nc + 255) / 256)), dim3(256), 0, st, dgate, z, nc, dz);
  // dpool[n][ci] = (1/hw) sum_co dz[n][co] * bf16(W[co][ci])
  launch_pdl(small_gemm_kernel<false, true, 0>, dim3((c + 31) / 32, (n + 31) / 32), dim3(256), 0, st, (const float*)dz,
             (long long)c, 1LL, weight, (long long)c, 1LL, n, c, c, 1.f / hw, dpool, (long long)c, (const float*)nullptr,
             (float*)nullptr, 0);
  // dW[co][ci] (+)= sum_n dz[n][co] * bf16(pool[n][ci]);  db[co] (+)= sum_n dz[n][co]
  launch_pdl(small_gemm_kernel<false, true, 2>, dim3((c + 31) / 32, (c + 31) / 32), dim3(256), 0, st, (const float*)dz, 1LL,
             (long long)c, pool, (long long)c, 1LL, c, c, n, 1.f, dweight, (long long)c, (const float*)nullptr,
             (float*)nullptr, accumulate_dw);
  launch_pdl(ese_dbias_kernel, dim3((c + 255) / 256), dim3(256), 0, st, (const float*)dz, n, c, dbias, accumulate_dw);
  const long long pixels = (long long)n * hw;
  if (accumulate_dx)
    launch_pdl(ese_bwd_dx_kernel<true>, dim3(vgrid(pixels * c8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)dout, lddo, gate, dpool, hw,
                                                                     pixels, c8, (__nv_bfloat16*)dx, lddx);
  else
    launch_pdl(ese_bwd_dx_kernel<false>, dim3(vgrid(pixels * c8, 256)), dim3(256), 0, st, (const __nv_bfloat16*)dout, lddo, gate, dpool, hw,
                                                                      pixels, c8, (__nv_bfloat16*)dx, lddx);
  count_launch(6);
  return check_cuda((int)cudaGetLastError(), "ese backward kernels");
}

int vtb_head_ce_fwd(const void* f, int ldf, int n, int hw, int c, const float* weight, const float* bias, int k,
                    const long long* labels, float label_smoothing, float* pooled, float* logits, float* dlogits,
                    float* row_loss, float* loss, void* stream) {
  if (!VVIEW_OK(f, ldf, c) || c % 8 || n <= 0 || hw <= 0 || k <= 0 || !weight || !bias || !labels || !pooled || !logits ||
      !dlogits || !row_loss || !loss || label_smoothing < 0.f || label_smoothing >= 1.f)
    return fail(VTB_EINVAL, "vtb_head_ce_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int c8 = c / 8;
  launch_pdl(hw_reduce_kernel<false>, dim3((c8 + 31) / 32, n), dim3(256), 0, st, (const __nv_bfloat16*)f, ldf,
             (const __nv_bfloat16*)nullptr, 0, hw, c8, pooled, c, 1.f / (float)hw);
  // logits[n][k] = pooled[n][:] . weight[k][:] + bias[k]
  launch_pdl(tile_gemm_kernel<32, 3>, dim3((k + 63) / 64, (n + 31) / 32), dim3(256), 0, st, (const float*)pooled, (long long)c,
             1LL, weight, 1LL, (long long)c, n, k, c, 1.f, (const float*)nullptr, logits, (long long)k, bias, 0);
  launch_pdl(head_ce_kernel, dim3(n), dim3(256), 0, st, (const float*)logits, k, labels, label_smoothing, 1.f / (float)n, dlogits,
             row_loss);
  launch_pdl(head_loss_mean_kernel, dim3(1), dim3(256), 0, st, (const float*)row_loss, n, loss);
  count_launch(4);
  return check_cuda((int)cudaGetLastError(), "vtb_head_ce_fwd");
}

int vtb_head_ce_bwd(const float* pooled, const float* dlogits, const float* weight, int n, int hw, int c, int k,
                    const float* gscale, float* dweight, float* dbias, int accumulate, void* df, int lddf, float* scratch,
                    void* stream) {
  if (!pooled || !dlogits || !weight || n <= 0 || hw <= 0 || c <= 0 || c % 8 || k <= 0 || !dweight || !dbias || !scratch ||
      (df && !VVIEW_OK(df, lddf, c)))
    return fail(VTB_EINVAL, "vtb_head_ce_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // the upstream gradient of the scalar loss (gscale, a device scalar) rides in the GEMMs' alpha: no scaling pass
  float* dpooled = scratch + (long long)n * k;   // [n][c]  (the first n*k floats of scratch are unused since round 2)
  // dweight[k][c] (+)= g * sum_n dlogits[n][k] * pooled[n][c]
  launch_pdl(tile_gemm_kernel<64, 2>, dim3((c + 63) / 64, (k + 63) / 64), dim3(256), 0, st, dlogits, 1LL, (long long)k, pooled,
             (long long)c, 1LL, k, c, n, 1.f, gscale, dweight, (long long)c, (const float*)nullptr, accumulate);
  launch_pdl(head_dbias_kernel, dim3((k + 127) / 128), dim3(128), 0, st, dlogits, n, k, gscale, dbias, accumulate);
  int launches = 2;
  if (df != nullptr) {
    // dpooled[n][c] = (g / hw) * sum_k dlogits[n][k] * weight[k][c]
    launch_pdl(tile_gemm_kernel<32, 0>, dim3((c + 63) / 64, (n + 31) / 32), dim3(256), 0, st, dlogits, (long long)k, 1LL, weight,
               (long long)c, 1LL, n, c, k, 1.f / (float)hw, gscale, dpooled, (long long)c, (const float*)nullptr, 0);
    const long long pixels = (long long)n * hw;
    launch_pdl(head_broadcast_kernel, dim3(vgrid(pixels * (c / 8), 256)), dim3(256), 0, st, (const float*)dpooled, hw, pixels,
               c / 8, (__nv_bfloat16*)df, lddf);
    launches += 2;
  }
  count_launch(launches);
  return check_cuda((int)cudaGetLastError(), "vtb_head_ce_bwd");
}

}  // extern "C"