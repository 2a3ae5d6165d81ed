// fp32 parity mode ("fp32 mode" of the north star: forward maps / gradients within 1e-4 relative of the reference run
// WITHOUT autocast).  Same dataflow as the bf16 tensor-core path - NHWC views with a pixel pitch, concat slices,
// residual epilogues, BatchNorm statistics in fp64 - but every tensor is fp32 and every contraction is a plain FMA
// loop on the CUDA cores with a fixed summation order (deterministic, no atomics).  This mode exists to pin the
// semantics of the path against the reference at fp32 accuracy; throughput is not its purpose.
//
// Reference call sites: vision_toolbox/components.py:26-39 (Conv2d / BatchNorm2d / ReLU), backbones/darknet.py:28,53,
// backbones/vovnet.py:20-28,55-61,94.
#include <algorithm>
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/vtb.h"
#include "common.cuh"

namespace vtb {

static int f32_grid(long long work, int block) {
  const int sms = std::max(1, num_sms());
  return (int)std::max<long long>(1, std::min<long long>((work + block - 1) / block, (long long)sms * 16));
}

struct F32Geom {
  int n, h, w, cin, cout, k, stride, pad, ho, wo, cin_real;
};

// ------------------------------------------------------------------------------------------------
// One tiled SIMT GEMM (64 x 64 x 16 tiles, 256 threads x 4x4 results) with three operand gathers:
//   MODE 0  fprop : C[m = out pixel][n = cout]      = sum_{k = (tap, ci)}  x[pixel(m, tap)][ci]   * w[n][ci][tap]
//   MODE 1  dgrad : C[m = in pixel ][n = cin]       = sum_{k = (tap, co)}  dy[pixel'(m, tap)][co] * w[co][n][tap]
//   MODE 2  wgrad : C[m = cout     ][n = (tap, ci)] = sum_{k = out pixel}  dy[k][m]               * x[pixel(k, tap)][ci]
//                   (K split over blockIdx.z, one fp32 partial tile per split, summed in order by f32_wgrad_reduce_kernel)
// Each result: fp32 FMA chains of kBK terms added into an fp64 running sum (error ~ 2^-24 of the term magnitudes).
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBN = 64, kBK = 16;

template <int MODE>
__device__ __forceinline__ float f32_fetch_a(const F32Geom& g, const float* __restrict__ src, int ld, long long m, long long k,
                                             long long M, long long K) {
  if (m >= M || k >= K) return 0.f;
  if (MODE == 0) {
    const int tap = (int)(k / g.cin), ci = (int)(k - (long long)tap * g.cin);
    const int kh = tap / g.k, kw = tap - kh * g.k;
    const int wo = (int)(m % g.wo);
    const long long t = m / g.wo;
    const int ho = (int)(t % g.ho);
    const long long img = t / g.ho;
    const int ih = ho * g.stride - g.pad + kh, iw = wo * g.stride - g.pad + kw;
    if (ih < 0 || ih >= g.h || iw < 0 || iw >= g.w) return 0.f;
    return __ldg(src + ((img * g.h + ih) * g.w + iw) * ld + ci);
  } else if (MODE == 1) {
    const int tap = (int)(k / g.cout), co = (int)(k - (long long)tap * g.cout);
    const int kh = tap / g.k, kw = tap - kh * g.k;
    const int iw = (int)(m % g.w);
    const long long t = m / g.w;
    const int ih = (int)(t % g.h);
    const long long img = t / g.h;
    const int th = ih + g.pad - kh, tw = iw + g.pad - kw;
    if (th < 0 || tw < 0 || th % g.stride || tw % g.stride) return 0.f;
    const int oh = th / g.stride, ow = tw / g.stride;
    if (oh >= g.ho || ow >= g.wo) return 0.f;
    return __ldg(src + ((img * g.ho + oh) * g.wo + ow) * ld + co);
  } else {
    return __ldg(src + k * ld + m);   // dy[pixel k][channel m]
  }
}

template <int MODE>
__device__ __forceinline__ float f32_fetch_b(const F32Geom& g, const float* __restrict__ src, int ld, long long k, long long n,
                                             long long K, long long N) {
  if (k >= K || n >= N) return 0.f;
  const int kk2 = g.k * g.k;
  if (MODE == 0) {
    const int tap = (int)(k / g.cin), ci = (int)(k - (long long)tap * g.cin);
    if (ci >= g.cin_real) return 0.f;
    return __ldg(src + ((long long)n * g.cin_real + ci) * kk2 + tap);
  } else if (MODE == 1) {
    const int tap = (int)(k / g.cout), co = (int)(k - (long long)tap * g.cout);
    if (n >= g.cin_real) return 0.f;
    return __ldg(src + ((long long)co * g.cin_real + n) * kk2 + tap);
  } else {
    const int tap = (int)(n / g.cin_real), ci = (int)(n - (long long)tap * g.cin_real);
    const int kh = tap / g.k, kw = tap - kh * g.k;
    const int wo = (int)(k % g.wo);
    const long long t = k / g.wo;
    const int ho = (int)(t % g.ho);
    const long long img = t / g.ho;
    const int ih = ho * g.stride - g.pad + kh, iw = wo * g.stride - g.pad + kw;
    if (ih < 0 || ih >= g.h || iw < 0 || iw >= g.w) return 0.f;
    return __ldg(src + ((img * g.h + ih) * g.w + iw) * ld + ci);
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
f32_conv_gemm_kernel(const F32Geom g, const float* __restrict__ a_src, int lda, const float* __restrict__ b_src, int ldb,
                     float* __restrict__ out, int ldo, int accumulate, long long M, long long N, long long K,
                     long long k_per_split) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[kBK][kBM + 4], Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * kBM, n0 = (long long)blockIdx.y * kBN;
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  const long long kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  // Two-level accumulation: fp32 FMA chains of kBK terms, added into fp64 running sums.  One sequential fp32 chain over the
  // whole K (up to 9216 here) loses ~sqrt(K) ulps; at the real layer sizes that flips enough ReLU masks against the
  // reference to move gradients by 1e-3 (tests/test_real_shapes.py) - oneDNN's blocked accumulation does not.
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (long long k0 = kbeg; k0 < kend; k0 += kBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int mm, kk;
      if (MODE == 2) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }   // fast index = contiguous axis
      As[kk][mm] = f32_fetch_a<MODE>(g, a_src, lda, m0 + mm, k0 + kk, M, kend);
      int nn, kb;
      if (MODE == 2) { nn = e & 63; kb = e >> 6; } else { kb = e & 15; nn = e >> 4; }
      Bs[kb][nn] = f32_fetch_b<MODE>(g, b_src, ldb, k0 + kb, n0 + nn, kend, N);
    }
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float r = (float)acc[i][j];
      if (MODE == 2) {
        out[((long long)blockIdx.z * M + m) * N + n] = r;   // partial tile of this split
      } else {
        float* dst = out + m * ldo + n;
        *dst = accumulate ? *dst + r : r;
      }
    }
  }
}

// dw_oihw[co][ci][tap] (+)= sum_split ws[split][co][tap*cin_real + ci]   (fixed order)
__global__ void f32_wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int cout, int cin_real, int taps,
                                        float* __restrict__ dw, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)cout * cin_real * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long t = i / taps;
    const int ci = (int)(t % cin_real);
    const int co = (int)(t / cin_real);
    const long long ncols = (long long)taps * cin_real;
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += (double)ws[((long long)s * cout + co) * ncols + (long long)tap * cin_real + ci];
    dw[i] = accumulate ? dw[i] + (float)acc : (float)acc;
  }
}

static int f32_wgrad_splits(const F32Geom& g) {
  const long long pixels = (long long)g.n * g.ho * g.wo;
  const long long tiles = (long long)((g.cout + kBM - 1) / kBM) * ((g.k * g.k * g.cin + kBN - 1) / kBN);   // padded cin: same answer for the workspace query and the launch
  long long s = (4LL * 148 + tiles - 1) / tiles;
  s = std::min<long long>(s, (pixels + 255) / 256);
  s = std::min<long long>(s, 128);
  return (int)std::max<long long>(1, s);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics / backward sums: two levels, fp64, fixed order.
//   level 1: block (32 channels x 8 pixel lanes) over one pixel range -> partial[r][c][2] (double)
//   level 2: sums[c][2] = sum_r partial[r][c][2]
// BWD = false: (y, y*y);  BWD = true: dz = dout * [bn(y) > 0 if relu], (dz, dz * xhat)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_bn(float y, float mean, float invstd, float gamma, float beta) {
  return (y - mean) * invstd * gamma + beta;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
f32_bn_sums_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ dout, int lddo, long long pixels, int c,
                   const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int relu, int rows, double* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[8][32][2];
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + tx;
  const int r = blockIdx.y;
  const long long per = (pixels + rows - 1) / rows;
  const long long p0 = r * per, p1 = (p0 + per < pixels) ? p0 + per : pixels;
  double s = 0.0, q = 0.0;
  if (ch < c) {
    float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
    if (BWD) { mu = mean[ch]; is = invstd[ch]; ga = gamma[ch]; be = beta[ch]; }
    for (long long p = p0 + tyy; p < p1; p += 8) {
      const float v = __ldg(y + p * ldy + ch);
      if (BWD) {
        float dz = __ldg(dout + p * lddo + ch);
        if (relu && !(f32_bn(v, mu, is, ga, be) > 0.f)) dz = 0.f;
        s += (double)dz;
        q += (double)dz * (double)((v - mu) * is);
      } else {
        s += (double)v;
        q += (double)v * (double)v;
      }
    }
  }
  red[tyy][tx][0] = s;
  red[tyy][tx][1] = q;
  __syncthreads();
  if (tyy == 0 && ch < c) {
    double ss = 0.0, qq = 0.0;
    for (int l = 0; l < 8; ++l) { ss += red[l][tx][0]; qq += red[l][tx][1]; }
    partial[((long long)r * c + ch) * 2] = ss;
    partial[((long long)r * c + ch) * 2 + 1] = qq;
  }
}

__global__ void f32_bn_sums_reduce_kernel(const double* __restrict__ partial, int rows, int c, double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over c*2
  if (i >= c * 2) return;
  double acc = 0.0;
  for (int r = 0; r < rows; ++r) acc += partial[(long long)r * c * 2 + i];
  sums[i] = acc;
}

static int f32_bn_rows(long long pixels, int c) {
  const long long cb = (c + 31) / 32;
  long long rows = (8LL * 148 + cb - 1) / cb;
  rows = std::min<long long>(rows, (pixels + 63) / 64);
  rows = std::min<long long>(rows, 512);
  return (int)std::max<long long>(1, rows);
}

// out = [relu]((y - mean) * invstd * gamma + beta) [+ residual]
__global__ void f32_bn_act_kernel(const float* __restrict__ y, int ldy, long long pixels, int c, const float* __restrict__ mean,
                                  const float* __restrict__ invstd, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int relu, const float* __restrict__ res, int ldr,
                                  float* __restrict__ out, int ldo) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i - p * c);
    float v = f32_bn(__ldg(y + p * ldy + ch), mean[ch], invstd[ch], gamma[ch], beta[ch]);
    if (relu) v = v > 0.f ? v : 0.f;
    if (res != nullptr) v += __ldg(res + p * ldr + ch);
    out[p * ldo + ch] = v;
  }
}

// dy = gamma * invstd * (dz - coef0 - xhat * coef1)
__global__ void f32_bn_bwd_apply_kernel(const float* __restrict__ dout, int lddo, const float* __restrict__ y, int ldy,
                                        long long pixels, int c, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, int relu, const float* __restrict__ coef,
                                        float* __restrict__ dy, int lddy) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i - p * c);
    const float v = __ldg(y + p * ldy + ch);
    const float mu = mean[ch], is = invstd[ch], ga = gamma[ch];
    float dz = __ldg(dout + p * lddo + ch);
    if (relu && !(f32_bn(v, mu, is, ga, beta[ch]) > 0.f)) dz = 0.f;
    const float xhat = (v - mu) * is;
    dy[p * lddy + ch] = ga * is * (dz - coef[ch * 2] - xhat * coef[ch * 2 + 1]);
  }
}

__global__ void f32_grad_add_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, long long pixels,
                                    int c, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i - p * c);
    const float v = __ldg(src + p * lds + ch);
    float* d = dst + p * ldd + ch;
    *d = accumulate ? *d + v : v;
  }
}

__global__ void f32_nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int c, long long hw, float* __restrict__ out,
                                        int cpad) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * hw * cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpad);
    const long long t = i / cpad;
    const long long p = t % hw;
    const long long img = t / hw;
    out[i] = ch < c ? __ldg(x + (img * c + ch) * hw + p) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// VoVNet ops in fp32
// ------------------------------------------------------------------------------------------------
__global__ void f32_maxpool_fwd_kernel(const float* __restrict__ x, int ldx, int n, int h, int w, int c, int ho, int wo,
                                       float* __restrict__ out, int ldo, unsigned char* __restrict__ idx) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * ho * wo * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int ow = (int)(t % wo);
    t /= wo;
    const int oh = (int)(t % ho);
    const long long img = t / ho;
    float m = -INFINITY;
    unsigned int arg = 255u;
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= h) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= w) continue;
        const float f = __ldg(x + ((img * h + ih) * w + iw) * ldx + ch);
        if (f > m || f != f) { m = f; arg = r * 3 + s; }
      }
    }
    const long long opix = (img * ho + oh) * wo + ow;
    out[opix * ldo + ch] = m;
    if (idx != nullptr) idx[opix * c + ch] = (unsigned char)arg;
  }
}

__global__ void f32_maxpool_bwd_kernel(const unsigned char* __restrict__ idx, int n, int h, int w, int c, int ho, int wo,
                                       const float* __restrict__ dout, int lddo, float* __restrict__ dx, int lddx,
                                       int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)n * h * w * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long t = i / c;
    const int iw0 = (int)(t % w);
    t /= w;
    const int ih0 = (int)(t % h);
    const long long img = t / h;
    float acc = 0.f;
    const int oh_lo = ih0 / 2, oh_hi = min(ho - 1, (ih0 + 1) / 2);
    const int ow_lo = iw0 / 2, ow_hi = min(wo - 1, (iw0 + 1) / 2);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int r0 = ih0 - (oh * 2 - 1), s0 = iw0 - (ow * 2 - 1);
        if (r0 < 0 || r0 > 2 || s0 < 0 || s0 > 2) continue;
        const long long opix = (img * ho + oh) * wo + ow;
        if (idx[opix * c + ch] == (unsigned char)(r0 * 3 + s0)) acc += __ldg(dout + opix * lddo + ch);
      }
    float* d = dx + ((img * h + ih0) * w + iw0) * lddx + ch;
    *d = accumulate ? *d + acc : acc;
  }
}

// out[img][ch] = mul * sum_hw a (* b): block = 32 channels x 8 pixel lanes, fixed-order combine
template <bool PRODUCT>
__global__ void __launch_bounds__(256)
f32_hw_reduce_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int hw, int c,
                     float* __restrict__ out, float mul) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + tx;
  const long long img = blockIdx.y;
  float s = 0.f;
  if (ch < c)
    for (int p = tyy; p < hw; p += 8) {
      const float f = __ldg(a + (img * hw + p) * lda + ch);
      s = PRODUCT ? fmaf(f, __ldg(b + (img * hw + p) * ldb + ch), s) : s + f;
    }
  red[tyy][tx] = s;
  __syncthreads();
  if (tyy == 0 && ch < c) {
    float acc = 0.f;
    for (int l = 0; l < 8; ++l) acc += red[l][tx];
    out[img * c + ch] = acc * mul;
  }
}

// z[n][co] = sum_ci W[co][ci] * pool[n][ci] + b[co]; gate = hardsigmoid(z).  One warp per (n, co).
__global__ void f32_ese_fc_kernel(const float* __restrict__ pool, const float* __restrict__ W, const float* __restrict__ bias,
                                  int n, int c, float* __restrict__ z, float* __restrict__ gate) {
  pdl_wait();
  pdl_trigger();
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)n * c) return;
  const int img = (int)(wid / c), co = (int)(wid % c);
  float acc = 0.f;
  for (int ci = lane; ci < c; ci += 32) acc = fmaf(W[(long long)co * c + ci], pool[(long long)img * c + ci], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float zz = acc + bias[co];
    z[wid] = zz;
    gate[wid] = fminf(fmaxf(zz * (1.f / 6.f) + 0.5f, 0.f), 1.f);
  }
}

__global__ void f32_ese_scale_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gate, int hw,
                                     long long pixels, int c, const float* __restrict__ res, int ldr, float* __restrict__ out,
                                     int ldo) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i - p * c);
    float v = __ldg(x + p * ldx + ch) * gate[(p / hw) * c + ch];
    if (res != nullptr) v += __ldg(res + p * ldr + ch);
    out[p * ldo + ch] = v;
  }
}

// dz = dgate * hardsigmoid'(z)
__global__ void f32_ese_dz_kernel(const float* __restrict__ dgate, const float* __restrict__ z, long long total,
                                  float* __restrict__ dz) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float zz = z[i];
  dz[i] = (zz > -3.f && zz < 3.f) ? dgate[i] * (1.f / 6.f) : 0.f;
}
// dpool[n][ci] = inv_hw * sum_co dz[n][co] * W[co][ci]
__global__ void f32_ese_dpool_kernel(const float* __restrict__ dz, const float* __restrict__ W, int n, int c, float inv_hw,
                                     float* __restrict__ dpool) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * c) return;
  const int img = (int)(i / c), ci = (int)(i % c);
  float acc = 0.f;
  for (int co = 0; co < c; ++co) acc = fmaf(dz[(long long)img * c + co], W[(long long)co * c + ci], acc);
  dpool[i] = acc * inv_hw;
}
// dW[co][ci] (+)= sum_n dz[n][co] * pool[n][ci]; db[co] (+)= sum_n dz[n][co]
__global__ void f32_ese_dw_kernel(const float* __restrict__ dz, const float* __restrict__ pool, int n, int c,
                                  float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)c * c) return;
  const int co = (int)(i / c), ci = (int)(i % c);
  float acc = 0.f, accb = 0.f;
  for (int img = 0; img < n; ++img) {
    const float d = dz[(long long)img * c + co];
    acc = fmaf(d, pool[(long long)img * c + ci], acc);
    accb += d;
  }
  dW[i] = accumulate ? dW[i] + acc : acc;
  if (ci == 0) db[co] = accumulate ? db[co] + accb : accb;
}
// dx (+)= dout * gate + dpool
__global__ void f32_ese_bwd_dx_kernel(const float* __restrict__ dout, int lddo, const float* __restrict__ gate,
                                      const float* __restrict__ dpool, int hw, long long pixels, int c, float* __restrict__ dx,
                                      int lddx, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long total = pixels * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c;
    const int ch = (int)(i - p * c);
    const long long img = p / hw;
    const float v = fmaf(__ldg(dout + p * lddo + ch), gate[img * c + ch], dpool[img * c + ch]);
    float* d = dx + p * lddx + ch;
    *d = accumulate ? *d + v : v;
  }
}

static bool f32_geom(const VtbConv* c, int cin_real, F32Geom& g, const char* who) {
  if (!c || c->n <= 0 || c->h <= 0 || c->w <= 0 || c->cin <= 0 || c->cout <= 0 || c->k <= 0 || c->stride <= 0 ||
      c->pad < 0 || cin_real <= 0 || cin_real > c->cin) {
    fail(VTB_EINVAL, "%s: bad convolution geometry", who);
    return false;
  }
  g.n = c->n; g.h = c->h; g.w = c->w; g.cin = c->cin; g.cout = c->cout; g.k = c->k; g.stride = c->stride; g.pad = c->pad;
  g.ho = (c->h + 2 * c->pad - c->k) / c->stride + 1;
  g.wo = (c->w + 2 * c->pad - c->k) / c->stride + 1;
  g.cin_real = cin_real;
  if (g.ho <= 0 || g.wo <= 0) {
    fail(VTB_EINVAL, "%s: empty output", who);
    return false;
  }
  return true;
}

}  // namespace vtb

using namespace vtb;
#define FVIEW_OK(ptr, ld, c) ((ptr) != nullptr && (ld) >= (c) && (reinterpret_cast<uintptr_t>(ptr) & 3) == 0)

extern "C" {

int vtb_f32_nchw_to_nhwc(const float* x, int n, int c, int h, int w, float* out, int cpad, void* stream) {
  if (!x || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0 || cpad < c) return fail(VTB_EINVAL, "vtb_f32_nchw_to_nhwc: bad arguments");
  const long long total = (long long)n * h * w * cpad;
  launch_pdl(f32_nchw_to_nhwc_kernel, dim3(f32_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, n, c, (long long)h * w, out,
             cpad);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_nchw_to_nhwc_kernel");
}

int vtb_f32_conv_fprop(const VtbConv* c, const float* x, int ldx, const float* w_oihw, int cin_real, float* y, int ldy,
                       void* stream) {
  F32Geom g;
  if (!f32_geom(c, cin_real, g, "vtb_f32_conv_fprop")) return VTB_EINVAL;
  if (!FVIEW_OK(x, ldx, g.cin) || !FVIEW_OK(y, ldy, g.cout) || !w_oihw) return fail(VTB_EINVAL, "vtb_f32_conv_fprop: bad views");
  const long long M = (long long)g.n * g.ho * g.wo, N = g.cout, K = (long long)g.k * g.k * g.cin;
  dim3 grid((unsigned)((M + kBM - 1) / kBM), (unsigned)((N + kBN - 1) / kBN), 1);
  launch_pdl(f32_conv_gemm_kernel<0>, grid, dim3(256), 0, (cudaStream_t)stream, g, x, ldx, w_oihw, 0, y, ldy, 0, M, N, K, K);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_conv_gemm_kernel<fprop>");
}

int vtb_f32_conv_dgrad(const VtbConv* c, const float* dy, int lddy, const float* w_oihw, int cin_real, float* dx, int lddx,
                       int accumulate, void* stream) {
  F32Geom g;
  if (!f32_geom(c, cin_real, g, "vtb_f32_conv_dgrad")) return VTB_EINVAL;
  if (!FVIEW_OK(dy, lddy, g.cout) || !FVIEW_OK(dx, lddx, g.cin) || !w_oihw) return fail(VTB_EINVAL, "vtb_f32_conv_dgrad: bad views");
  const long long M = (long long)g.n * g.h * g.w, N = g.cin, K = (long long)g.k * g.k * g.cout;
  dim3 grid((unsigned)((M + kBM - 1) / kBM), (unsigned)((N + kBN - 1) / kBN), 1);
  launch_pdl(f32_conv_gemm_kernel<1>, grid, dim3(256), 0, (cudaStream_t)stream, g, dy, lddy, w_oihw, 0, dx, lddx, accumulate, M, N, K,
             K);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_conv_gemm_kernel<dgrad>");
}

size_t vtb_f32_conv_wgrad_workspace_bytes(const VtbConv* c) {
  F32Geom g;
  if (!f32_geom(c, c ? c->cin : 0, g, "vtb_f32_conv_wgrad_workspace_bytes")) return 0;
  return (size_t)f32_wgrad_splits(g) * g.cout * g.k * g.k * g.cin * sizeof(float);
}

int vtb_f32_conv_wgrad(const VtbConv* c, const float* dy, int lddy, const float* x, int ldx, void* workspace, float* dw_oihw,
                       int cin_real, int accumulate, void* stream) {
  F32Geom g;
  if (!f32_geom(c, cin_real, g, "vtb_f32_conv_wgrad")) return VTB_EINVAL;
  if (!FVIEW_OK(dy, lddy, g.cout) || !FVIEW_OK(x, ldx, g.cin) || !workspace || !dw_oihw)
    return fail(VTB_EINVAL, "vtb_f32_conv_wgrad: bad views");
  const long long M = g.cout, N = (long long)g.k * g.k * g.cin_real, K = (long long)g.n * g.ho * g.wo;
  const int splits = f32_wgrad_splits(g);
  long long per = (K + splits - 1) / splits;
  per = (per + kBK - 1) / kBK * kBK;
  dim3 grid((unsigned)((M + kBM - 1) / kBM), (unsigned)((N + kBN - 1) / kBN), (unsigned)splits);
  launch_pdl(f32_conv_gemm_kernel<2>, grid, dim3(256), 0, (cudaStream_t)stream, g, dy, lddy, x, ldx, (float*)workspace, 0, 0, M, N, K,
             per);
  launch_pdl(f32_wgrad_reduce_kernel, dim3(f32_grid(M * N, 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)workspace, splits,
             g.cout, g.cin_real, g.k * g.k, dw_oihw, accumulate);
  count_launch(2);
  return check_cuda((int)cudaGetLastError(), "f32_conv_gemm_kernel<wgrad>");
}

int vtb_f32_bn_rows(long long pixels, int c) {
  if (pixels <= 0 || c <= 0) return fail(VTB_EINVAL, "vtb_f32_bn_rows: bad arguments");
  return f32_bn_rows(pixels, c);
}

int vtb_f32_bn_stats(const float* y, int ldy, long long pixels, int c, double* partial, double* sums, void* stream) {
  if (!FVIEW_OK(y, ldy, c) || pixels <= 0 || !partial || !sums) return fail(VTB_EINVAL, "vtb_f32_bn_stats: bad arguments");
  const int rows = f32_bn_rows(pixels, c);
  launch_pdl(f32_bn_sums_kernel<false>, dim3((c + 31) / 32, rows), dim3(256), 0, (cudaStream_t)stream, y, ldy, (const float*)nullptr, 0,
             pixels, c, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, 0, rows, partial);
  launch_pdl(f32_bn_sums_reduce_kernel, dim3((c * 2 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const double*)partial, rows, c,
             sums);
  count_launch(2);
  return check_cuda((int)cudaGetLastError(), "f32_bn_sums_kernel");
}

int vtb_f32_bn_act(const float* y, int ldy, long long pixels, int c, const float* mean, const float* invstd, const float* gamma,
                   const float* beta, int relu, const float* residual, int ldr, float* out, int ldo, void* stream) {
  if (!FVIEW_OK(y, ldy, c) || !FVIEW_OK(out, ldo, c) || pixels <= 0 || !mean || !invstd || !gamma || !beta ||
      (residual && ldr < c))
    return fail(VTB_EINVAL, "vtb_f32_bn_act: bad arguments");
  launch_pdl(f32_bn_act_kernel, dim3(f32_grid(pixels * c, 256)), dim3(256), 0, (cudaStream_t)stream, y, ldy, pixels, c, mean, invstd,
             gamma, beta, relu, residual, ldr, out, ldo);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_bn_act_kernel");
}

int vtb_f32_bn_bwd_reduce(const float* dout, int lddo, const float* y, int ldy, long long pixels, int c, const float* mean,
                          const float* invstd, const float* gamma, const float* beta, int relu, double* partial, double* sums,
                          void* stream) {
  if (!FVIEW_OK(dout, lddo, c) || !FVIEW_OK(y, ldy, c) || pixels <= 0 || !mean || !invstd || !gamma || !beta || !partial || !sums)
    return fail(VTB_EINVAL, "vtb_f32_bn_bwd_reduce: bad arguments");
  const int rows = f32_bn_rows(pixels, c);
  launch_pdl(f32_bn_sums_kernel<true>, dim3((c + 31) / 32, rows), dim3(256), 0, (cudaStream_t)stream, y, ldy, dout, lddo, pixels, c, mean,
             invstd, gamma, beta, relu, rows, partial);
  launch_pdl(f32_bn_sums_reduce_kernel, dim3((c * 2 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const double*)partial, rows, c,
             sums);
  count_launch(2);
  return check_cuda((int)cudaGetLastError(), "f32_bn_sums_kernel<bwd>");
}

int vtb_f32_bn_bwd_apply(const float* dout, int lddo, const float* y, int ldy, long long pixels, int c, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, int relu, const float* coef, float* dy,
                         int lddy, void* stream) {
  if (!FVIEW_OK(dout, lddo, c) || !FVIEW_OK(y, ldy, c) || !FVIEW_OK(dy, lddy, c) || pixels <= 0 || !mean || !invstd || !gamma ||
      !beta || !coef)
    return fail(VTB_EINVAL, "vtb_f32_bn_bwd_apply: bad arguments");
  launch_pdl(f32_bn_bwd_apply_kernel, dim3(f32_grid(pixels * c, 256)), dim3(256), 0, (cudaStream_t)stream, dout, lddo, y, ldy, pixels,
             c, mean, invstd, gamma, beta, relu, coef, dy, lddy);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_bn_bwd_apply_kernel");
}

int vtb_f32_grad_add(void* dst, int ldd, const void* src, int lds, long long pixels, int c, int accumulate, void* stream) {
  if (!FVIEW_OK(dst, ldd, c) || !FVIEW_OK(src, lds, c) || pixels <= 0) return fail(VTB_EINVAL, "vtb_f32_grad_add: bad arguments");
  launch_pdl(f32_grad_add_kernel, dim3(f32_grid(pixels * c, 256)), dim3(256), 0, (cudaStream_t)stream, (float*)dst, ldd,
             (const float*)src, lds, pixels, c, accumulate);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_grad_add_kernel");
}

int vtb_f32_maxpool3s2_fwd(const void* x, int ldx, int n, int h, int w, int c, void* out, int ldo, void* idx, void* stream) {
  if (!FVIEW_OK(x, ldx, c) || !FVIEW_OK(out, ldo, c) || n <= 0 || h <= 0 || w <= 0) return fail(VTB_EINVAL, "vtb_f32_maxpool3s2_fwd: bad arguments");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * ho * wo * c;
  launch_pdl(f32_maxpool_fwd_kernel, dim3(f32_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)x, ldx, n, h, w, c,
             ho, wo, (float*)out, ldo, (unsigned char*)idx);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_maxpool_fwd_kernel");
}

int vtb_f32_maxpool3s2_bwd(const void* x, int ldx, int n, int h, int w, int c, const void* dout, int lddo, void* dx, int lddx,
                           int accumulate, const void* idx, void* stream) {
  (void)x; (void)ldx;
  if (!idx) return fail(VTB_EINVAL, "vtb_f32_maxpool3s2_bwd: the saved argmax indices are required in fp32 mode");
  if (!FVIEW_OK(dout, lddo, c) || !FVIEW_OK(dx, lddx, c) || n <= 0 || h <= 0 || w <= 0) return fail(VTB_EINVAL, "vtb_f32_maxpool3s2_bwd: bad arguments");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const long long total = (long long)n * h * w * c;
  launch_pdl(f32_maxpool_bwd_kernel, dim3(f32_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const unsigned char*)idx, n, h, w,
             c, ho, wo, (const float*)dout, lddo, (float*)dx, lddx, accumulate);
  count_launch(1);
  return check_cuda((int)cudaGetLastError(), "f32_maxpool_bwd_kernel");
}

int vtb_f32_ese_fwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* bias, const void* residual,
                    int ldr, void* out, int ldo, float* pool, float* z, float* gate, void* stream) {
  if (!FVIEW_OK(x, ldx, c) || !FVIEW_OK(out, ldo, c) || n <= 0 || hw <= 0 || !weight || !bias || !pool || !z || !gate ||
      (residual && ldr < c))
    return fail(VTB_EINVAL, "vtb_f32_ese_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long pixels = (long long)n * hw;
  launch_pdl(f32_hw_reduce_kernel<false>, dim3((c + 31) / 32, n), dim3(256), 0, st, (const float*)x, ldx, (const float*)nullptr, 0, hw,
             c, pool, 1.f / (float)hw);
  launch_pdl(f32_ese_fc_kernel, dim3((unsigned)(((long long)n * c * 32 + 255) / 256)), dim3(256), 0, st, (const float*)pool, weight,
             bias, n, c, z, gate);
  launch_pdl(f32_ese_scale_kernel, dim3(f32_grid(pixels * c, 256)), dim3(256), 0, st, (const float*)x, ldx, (const float*)gate, hw,
             pixels, c, (const float*)residual, ldr, (float*)out, ldo);
  count_launch(3);
  return check_cuda((int)cudaGetLastError(), "f32_ese_fwd");
}

int vtb_f32_ese_bwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* pool, const float* z,
                    const float* gate, const void* dout, int lddo, void* dx, int lddx, int accumulate_dx, float* dweight,
                    float* dbias, int accumulate_dw, float* scratch, void* stream) {
  if (!FVIEW_OK(x, ldx, c) || !FVIEW_OK(dout, lddo, c) || !FVIEW_OK(dx, lddx, c) || n <= 0 || hw <= 0 || !weight || !pool || !z ||
      !gate || !dweight || !dbias || !scratch)
    return fail(VTB_EINVAL, "vtb_f32_ese_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long pixels = (long long)n * hw, nc = (long long)n * c;
  float* dgate = scratch;
  float* dz = scratch + nc;
  float* dpool = scratch + 2 * nc;
  launch_pdl(f32_hw_reduce_kernel<true>, dim3((c + 31) / 32, n), dim3(256), 0, st, (const float*)dout, lddo, (const float*)x, ldx, hw, c,
             dgate, 1.f);
  launch_pdl(f32_ese_dz_kernel, dim3((unsigned)((nc + 255) / 256)), dim3(256), 0, st, (const float*)dgate, z, nc, dz);
  launch_pdl(f32_ese_dpool_kernel, dim3((unsigned)((nc + 255) / 256)), dim3(256), 0, st, (const float*)dz, weight, n, c,
             1.f / (float)hw, dpool);
  launch_pdl(f32_ese_dw_kernel, dim3((unsigned)(((long long)c * c + 255) / 256)), dim3(256), 0, st, (const float*)dz, pool, n, c, dweight,
             dbias, accumulate_dw);
  launch_pdl(f32_ese_bwd_dx_kernel, dim3(f32_grid(pixels * c, 256)), dim3(256), 0, st, (const float*)dout, lddo, gate,
             (const float*)dpool, hw, pixels, c, (float*)dx, lddx, accumulate_dx);
  count_launch(5);
  return check_cuda((int)cudaGetLastError(), "f32_ese_bwd");
}

}  // extern "C"
