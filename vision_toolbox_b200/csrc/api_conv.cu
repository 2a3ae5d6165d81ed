// C-ABI entry points for the convolution family (include/vtb.h): geometry -> tensor maps + kernel params.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/vtb.h"
#include "common.cuh"
#include "igemm.cuh"
#include "tmap.cuh"

namespace vtb {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_cuda(int e, const char* what) {
  if (e == 0) return VTB_OK;
  return fail(VTB_ECUDA, "%s: %s", what, cudaGetErrorString((cudaError_t)e));
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("VTB_PDL");
    return !(v && v[0] == '0');
  }();
  return on;
}

int num_sms() {
  static int sms = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
  }();
  return sms;
}

static int chunk_of(int c) { return (c % 64 == 0) ? 64 : (c % 32 == 0 ? 32 : 16); }
// largest multiple of 16 that divides c and is <= 256
static int block_of(int c) {
  for (int b = 256; b >= 16; b -= 16)
    if (c % b == 0) return b;
  return 0;
}
static bool conv_ok(const VtbConv* c) {
  return c && c->n > 0 && c->h > 0 && c->w > 0 && c->cin > 0 && c->cout > 0 && c->cin % 16 == 0 && c->cout % 16 == 0 &&
         c->k >= 1 && c->k * c->k <= kMaxTaps && c->stride >= 1 && c->stride <= 2 && c->pad >= 0 &&
         (c->h + 2 * c->pad - c->k) >= 0 && (c->w + 2 * c->pad - c->k) >= 0;
}
static void out_hw(const VtbConv* c, int* ho, int* wo) {
  *ho = (c->h + 2 * c->pad - c->k) / c->stride + 1;
  *wo = (c->w + 2 * c->pad - c->k) / c->stride + 1;
}
// development overrides (tools/bench_conv): VTB_BLOCK_M / VTB_BLOCK_N / VTB_STAGES, read once per process
static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
static unsigned long long* g_dbg = nullptr;
// Tiling decisions of one conv_igemm_kernel launch (GEMM: M pixels x ncols channels).
struct ConvTiling {
  int block_m, block_n, panel_w, stages, panel_bufs, grid, n_blocks, stat_rows, ksplit;
};
// M pixels x ncols channels, K = total_k16 slices of 16 elements
static ConvTiling plan_conv_tiling(long long M, int ncols, int total_k16) {
  ConvTiling t;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  t.block_n = block_of(ncols);
  static const int o_bn = env_int("VTB_BLOCK_N");
  if (o_bn >= 16 && ncols % o_bn == 0 && o_bn <= 256) t.block_n = o_bn;
  // the epilogue keeps per-panel statistics in registers: at most kMaxPanels panels per accumulator
  while (t.block_n > 16 && t.block_n / std::min(32, chunk_of(t.block_n)) > kMaxPanels) {
    int b = t.block_n - 16;
    while (b > 16 && ncols % b) b -= 16;
    t.block_n = b;
  }
  // staged output panels: at most 32 columns (the 16 epilogue warps each stage 32 rows x 64 B), and at least two
  // panels per accumulator so that all four epilogue groups have work
  t.panel_w = std::min(32, chunk_of(t.block_n));
  // (narrow tiles, block_n <= 64, run the tile-parallel epilogue: whole 32-column panels, 64-byte row segments)
  static const bool tp_ok = !(getenv("VTB_TILE_PAR") && atoi(getenv("VTB_TILE_PAR")) == 0);
  if (!(tp_ok && t.block_n <= 64) && t.block_n / t.panel_w < 2 && t.panel_w > 16) t.panel_w /= 2;
  t.n_blocks = ncols / t.block_n;
  // 256-row tiles halve the weight traffic per FLOP and double the bytes per TMA request; use them whenever
  // there are enough of them to occupy most of the machine
  // Waves x tile time decides: a 256-row tile costs ~1.36x a 128-row one (the im2col request cadence does not depend on
  // the box height), so 128-row tiles only win while all of them still fit ONE wave (profiles/r02_convs_vovnet_tiles.txt:
  // 3x3 192->192 @14^2, 98 / 196 tiles: 20.3 us with 256 rows, 27.1 us with 128; 224->224 @7^2, 25 / 49 tiles: 24.6 / 18.1)
  const long long tiles256 = ((M + 255) / 256) * t.n_blocks;
  const long long tiles128 = ((M + 127) / 128) * t.n_blocks;
  const double cost256 = 1.36 * (double)((tiles256 + sms - 1) / sms), cost128 = (double)((tiles128 + sms - 1) / sms);
  t.block_m = (cost256 <= cost128) ? 256 : 128;
  static const bool old_rule = getenv("VTB_BLOCKM_RULE") && !strcmp(getenv("VTB_BLOCKM_RULE"), "r1");   // A/B: round-1 rule
  if (old_rule) t.block_m = (tiles256 * 4 >= (long long)sms * 3) ? 256 : 128;
  static const int o_bm = env_int("VTB_BLOCK_M");
  if (o_bm == 128 || o_bm == 256) t.block_m = o_bm;
  const long long tiles = ((M + t.block_m - 1) / t.block_m) * t.n_blocks;
  long long grid = std::min<long long>(tiles, sms);
  grid = std::max<long long>(t.n_blocks, grid / t.n_blocks * t.n_blocks);
  t.grid = (int)grid;
  const int per_stage = t.block_m * kStageK * 2 + ((t.block_n * kStageK * 2 + 1023) / 1024) * 1024;
  t.panel_bufs = 1;
  int s = (kSmemBudget - (kEpiWarps * 2048 + 256 + 1024)) / per_stage;
  t.stages = std::max(2, std::min(s, 8));
  // independent accumulation chains (see igemm.cu): up to 4 per tile; a long K loop prefers chains over TMEM
  // double buffering, a short one (1x1 convolutions) keeps two sets so the epilogue overlaps the next tile
  {
    const int halves = t.block_m / kBlockM;
    const int acc_stride = (t.block_n + 31) & ~31;
    int ks = 4 / halves;
    while (ks > 1 && (halves * ks * acc_stride > 512 || ks > total_k16)) ks >>= 1;
    // two TMEM sets (the epilogue of tile i overlaps the MMAs of tile i+1) beat extra chains: the kernel is bound by the
    // operand feed, not by the MMA issue rate (VTB_KSPLIT sweeps in profiles/r01_conv_kernel_bench.txt)
    while (ks > 1 && 2 * halves * ks * acc_stride > 512) ks >>= 1;
    static const int o_ks = env_int("VTB_KSPLIT");
    if (o_ks > 0 && halves * o_ks * acc_stride <= 512 && o_ks <= total_k16) ks = o_ks;
    t.ksplit = ks;
  }
  static const int o_stages = env_int("VTB_STAGES");
  if (o_stages > 0) t.stages = std::min(t.stages, o_stages);
  t.stat_rows = t.grid / t.n_blocks;  // one row per CTA
  return t;
}
static void apply_tiling(ConvIgemmParams& p, const ConvTiling& t) {
  p.dbg = g_dbg;
  p.block_m = t.block_m;
  p.block_n = t.block_n;
  p.panel_w = t.panel_w;
  p.num_stages = t.stages;
  p.panel_bufs = t.panel_bufs;
  p.ksplit = t.ksplit;
}

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin_real, int cin, int kk,
                                   __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const long long total = (long long)cout * kk * cin;
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int t = (int)((i / cin) % kk);
    const int co = (int)(i / ((long long)cin * kk));
    const float v = (ci < cin_real) ? w[((long long)co * cin_real + ci) * kk + t] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    wf[i] = b;
    if (wd) wd[((long long)ci * kk + t) * cout + co] = b;
  }
}

// The same re-pack for EVERY convolution of a plan in one launch (VtbPackJob table in device memory): block b serves the
// job with the largest first_block <= b and, inside it, one tile of 32 output channels x CI_T input channels (all taps).
// The fp32 master is read in its own order (for one output channel the CI_T x kk block is contiguous), staged in shared
// memory, and written out twice with 64-byte runs: wf[co][tap][ci] (ci fastest) and wd[ci][tap][co] (co fastest).
constexpr int kPackCoT = 32;
__host__ __device__ inline int pack_ci_tile(int kk) { return kk == 1 ? 256 : (kk <= 9 ? 32 : 8); }
constexpr int kPackSmemElems = 36 * kPackCoT * 9;   // >= kk * 32 * (CI_T + 1): kk 1 (CI_T 256), kk <= 9 (CI_T 32), kk <= 36 (CI_T 8)

// one 32 x CI_T tile of one job; KK > 0: compile-time tap count (divisions become shifts / multiplies), 0: runtime
// SGD: J.g != NULL -> the master weight is first UPDATED in place (torch.optim.SGD: g += wd * w; m = mu * m + g;
// w -= lr * m; reference classifier.py:141-169) and the bf16 operands are cut from the new value in the same pass.
template <int KK, bool SGD>
__device__ __forceinline__ void pack_tile(const VtbPackJob& J, int b, __nv_bfloat16* tile, float lr, float mu) {
  const int kk = KK > 0 ? KK : J.kk;
  const int cit = pack_ci_tile(kk), pitch = cit + 1;
  const int n_ci_tiles = (J.cin + cit - 1) / cit;
  const int co0 = (b / n_ci_tiles) * kPackCoT, ci0 = (b % n_ci_tiles) * cit;
  const int per_co = cit * kk;
  const float* __restrict__ w = J.w;
  for (int e = threadIdx.x; e < kPackCoT * per_co; e += blockDim.x) {
    const int co_l = e / per_co, r = e - co_l * per_co;
    const int ci_l = r / kk, t = r - ci_l * kk;
    const int co = co0 + co_l, ci = ci0 + ci_l;
    float v = 0.f;
    if (co < J.cout && ci < J.cin_real) {
      const long long idx = ((long long)co * J.cin_real + ci) * kk + t;
      if (SGD && J.g != nullptr) {
        float* wm = const_cast<float*>(w);
        const float w0 = wm[idx];
        const float g = fmaf(J.weight_decay, w0, J.g[idx]);
        const float m = fmaf(mu, J.m[idx], g);
        J.m[idx] = m;
        v = fmaf(-lr, m, w0);
        wm[idx] = v;
      } else {
        v = __ldg(w + idx);
      }
    }
    tile[(t * kPackCoT + co_l) * pitch + ci_l] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  __nv_bfloat16* __restrict__ wf = reinterpret_cast<__nv_bfloat16*>(J.wf);
  __nv_bfloat16* __restrict__ wd = reinterpret_cast<__nv_bfloat16*>(J.wd);
  for (int e = threadIdx.x; e < kPackCoT * per_co; e += blockDim.x) {
    const int ci_l = e % cit, q = e / cit;
    const int t = q % kk, co_l = q / kk;
    const int co = co0 + co_l, ci = ci0 + ci_l;
    if (co < J.cout && ci < J.cin) wf[(long long)co * J.wf_ld + t * J.cin + ci] = tile[(t * kPackCoT + co_l) * pitch + ci_l];
  }
  if (wd != nullptr) {
    for (int e = threadIdx.x; e < kPackCoT * per_co; e += blockDim.x) {
      const int co_l = e % kPackCoT, q = e / kPackCoT;
      const int t = q % kk, ci_l = q / kk;
      const int co = co0 + co_l, ci = ci0 + ci_l;
      if (co < J.cout && ci < J.cin)
        wd[((long long)ci * kk + t) * J.wd_ld + J.wd_co_off + co] = tile[(t * kPackCoT + co_l) * pitch + ci_l];
    }
  }
  __syncthreads();   // the tile buffer is reused by this block's next tile
}

constexpr int kPackMaxJobs = 256;
template <bool SGD>
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const VtbPackJob* __restrict__ jobs, int njobs,
                                                                  long long total_tiles, const float* __restrict__ hyper) {
  __shared__ __nv_bfloat16 tile[kPackSmemElems];   // [tap][co_l][CI_T + 1]
  __shared__ long long first[kPackMaxJobs];        // first tile of every job (the search below stays in shared memory)
  pdl_wait();
  pdl_trigger();
  const float lr = SGD ? hyper[0] : 0.f, mu = SGD ? hyper[1] : 0.f;
  for (int j = threadIdx.x; j < njobs; j += blockDim.x) first[j] = jobs[j].first_block;
  __syncthreads();
  for (long long tix = blockIdx.x; tix < total_tiles; tix += gridDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (first[mid] <= tix) lo = mid; else hi = mid - 1;
    }
    const VtbPackJob J = jobs[lo];
    const int b = (int)(tix - first[lo]);
    if (J.kk == 1) pack_tile<1, SGD>(J, b, tile, lr, mu);
    else if (J.kk == 9) pack_tile<9, SGD>(J, b, tile, lr, mu);
    else pack_tile<0, SGD>(J, b, tile, lr, mu);
  }
}

// SGD with momentum over a table of plain tensors (BatchNorm weights / biases, the classifier head): block b serves 1024
// consecutive elements of the job with the largest first_block <= b.
constexpr int kSgdBlockElems = 1024;
__global__ void __launch_bounds__(256) sgd_step_kernel(const VtbSgdJob* __restrict__ jobs, int njobs, long long total_blocks,
                                                      const float* __restrict__ hyper) {
  __shared__ long long first[kPackMaxJobs];
  pdl_wait();
  pdl_trigger();
  const float lr = hyper[0], mu = hyper[1];
  for (int j = threadIdx.x; j < njobs; j += blockDim.x) first[j] = jobs[j].first_block;
  __syncthreads();
  for (long long bix = blockIdx.x; bix < total_blocks; bix += gridDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (first[mid] <= bix) lo = mid; else hi = mid - 1;
    }
    const VtbSgdJob J = jobs[lo];
    const long long e0 = (bix - first[lo]) * kSgdBlockElems;
#pragma unroll
    for (int u = 0; u < kSgdBlockElems / 256; ++u) {
      const long long i = e0 + u * 256 + threadIdx.x;
      if (i < J.n) {
        const float w0 = J.w[i];
        const float g = fmaf(J.weight_decay, w0, J.g[i]);
        const float m = fmaf(mu, J.m[i], g);
        J.m[i] = m;
        J.w[i] = fmaf(-lr, m, w0);
      }
    }
  }
}

// dw[co][ci][t] (+)= sum_split ws[split][co][t*cin + ci].  One block per (output channel, slice of EW <= 64 input
// channels).  The 256 threads are EW channel lanes x SG split groups: every group sums its share of the splits with
// coalesced reads (many independent loads in flight: the split count, not the tile, carries the parallelism for small
// layers), the groups are combined through shared memory in a fixed order (deterministic), and the slice is written out
// in OIHW order, which is contiguous for the block ((ci, t) fastest).
template <int EW, bool DEEP>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int cout, int cin, int cin_real, int kk,
                    float* __restrict__ dw, int accumulate, float* __restrict__ dw2, int split) {
  constexpr int SG = 256 / EW;
  extern __shared__ float red_sm[];   // [SG][kk][EW + 1]
  const int co = blockIdx.x;
  const int c0 = blockIdx.y * EW;
  const int cw = min(EW, cin - c0);              // channels of the (padded) workspace row handled here
  const int cr = min(cw, cin_real - c0);         // of which real (the image stem pads 3 -> 16)
  pdl_wait();
  pdl_trigger();
  if (cr <= 0) return;
  const int e = threadIdx.x % EW, sg = threadIdx.x / EW;
  constexpr int pitch = EW + 1;
  const size_t split_stride = (size_t)cout * kk * cin;
  if (e < cw) {
    // the (tap, split) pairs are dealt round-robin to the SG split groups (fixed mapping -> deterministic), so every
    // group has loads to issue even when the layer needed a single split (the big 3x3 layers: a pure transpose)
    const float* src = ws + (size_t)co * kk * cin + c0 + e;
    float* mine = red_sm + (size_t)sg * kk * pitch + e;
    for (int t = 0; t < kk; ++t) mine[t * pitch] = 0.f;
    const int total = kk * splits;
    int q = sg;
    // DEEP (many splits: the small layers): 16, then 4 independent loads in flight per thread - the sums are latency-bound
    // (partials sit in L2) and there are few blocks, so memory-level parallelism is what shortens the kernel; layers with
    // few splits have thousands of blocks and prefer the lighter variant (registers -> occupancy)
    for (; DEEP && q + 15 * SG < total; q += 16 * SG) {
      float v[16];
      int tt[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int qq = q + u * SG;
        const int t = qq / splits, sp = qq - t * splits;
        tt[u] = t;
        v[u] = __ldg(src + (size_t)sp * split_stride + (size_t)t * cin);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) mine[tt[u] * pitch] += v[u];
    }
    for (; q + 3 * SG < total; q += 4 * SG) {
      float v[4];
      int tt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int qq = q + u * SG;
        const int t = qq / splits, sp = qq - t * splits;
        tt[u] = t;
        v[u] = __ldg(src + (size_t)sp * split_stride + (size_t)t * cin);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) mine[tt[u] * pitch] += v[u];
    }
    for (; q < total; q += SG) {
      const int t = q / splits, sp = q - t * splits;
      mine[t * pitch] += __ldg(src + (size_t)sp * split_stride + (size_t)t * cin);
    }
  }
  __syncthreads();
  // output channels >= split belong to the second weight tensor of a side-by-side pair (dw2, indexed from 0)
  float* dst = (dw2 != nullptr && co >= split) ? dw2 + ((size_t)(co - split) * cin_real + c0) * kk
                                               : dw + ((size_t)co * cin_real + c0) * kk;
  constexpr int ngroups = SG;
  for (int j = threadIdx.x; j < cr * kk; j += 256) {
    const int c = j / kk, t = j - c * kk;
    float v = 0.f;
    for (int g = 0; g < ngroups; ++g) v += red_sm[(g * kk + t) * pitch + c];
    dst[j] = accumulate ? dst[j] + v : v;
  }
}

struct WgradPlan {
  int ca, cc, ma, kpix, boxes_per_tap, total_boxes, boxes_per_tile, n_cols, n_tiles, m_tiles, ksplit, acc_stride, splits,
      kblocks, stages;
  long long mpix;
};
static WgradPlan plan_wgrad(const VtbConv* c) {
  WgradPlan w;
  int ho, wo;
  out_hw(c, &ho, &wo);
  const int sms = std::max(1, num_sms() > 0 ? num_sms() : 148);
  w.mpix = (long long)c->n * ho * wo;
  w.ca = chunk_of(c->cout);
  w.cc = chunk_of(c->cin);
  w.ma = std::min(128, c->cout);
  const int taps = c->k * c->k;
  w.boxes_per_tap = c->cin / w.cc;
  w.total_boxes = taps * w.boxes_per_tap;
  // column tiling of the flattened tap*cin axis: tiles of <= 256 columns, as even as the box width allows
  // 128-column tiles for the big 3x3 layers: with N = 128 the four K-split accumulator chains fit TMEM, a stage is
  // 64 KB at 128 pixels (3 stages) and the measured optimum of the kpix x boxes sweep (profiles/r01_wgrad_sweep.txt)
  const int max_boxes = (taps > 1 && taps * c->cin >= 512 && w.cc == 64) ? 128 / w.cc : 256 / w.cc;
  int best_tiles = (w.total_boxes + max_boxes - 1) / max_boxes, best_bpt = 0, best_waste = 1 << 30;
  for (int nt = best_tiles; nt <= best_tiles + 2; ++nt) {
    const int bpt = (w.total_boxes + nt - 1) / nt;
    const int tiles = (w.total_boxes + bpt - 1) / bpt;
    const int waste = tiles * bpt - w.total_boxes;
    if (waste < best_waste) { best_waste = waste; best_bpt = bpt; best_tiles = tiles; }
  }
  if (taps == 1) {
    // 1x1 layers: the output is tiny (cout x cin) and the pixel axis carries all the parallelism, so the split count -
    // and with it the fp32 partial traffic (splits x cout x cin x 4 B written, then read by the reduce kernel) - is what
    // the tile width decides.  Measured (profiles/r02_wgrad_boxes_sweep.txt): the optimum keeps >= ~2500 pixels per split;
    // e.g. 256->256 @11^2: 2 tiles x 74 splits 27.4 us, 8 tiles x 18 splits 15.5 us.  Widest tile that still does.
    const int kpix_guess = 128;
    const long long kblocks = (w.mpix + kpix_guess - 1) / kpix_guess;
    int pick = 0;
    long long pick_px = -1;
    for (int b = std::min(max_boxes, w.total_boxes); b >= 1; --b) {
      const int nt = (w.total_boxes + b - 1) / b;
      if ((nt * b - w.total_boxes) * 8 >= w.total_boxes) continue;   // > 12.5 % of the MMA columns would be padding
      const int tiles = nt * ((c->cout + 127) / 128);
      long long sp = std::max(1, sms / tiles);
      sp = std::min<long long>(sp, std::max<long long>(1, kblocks / 2));
      const long long px = w.mpix / sp;
      if (px >= 2500) { pick = b; break; }
      if (px > pick_px) { pick_px = px; pick = b; }                  // nothing qualifies: the fewest splits, widest first
    }
    best_bpt = std::max(1, pick);
    best_tiles = (w.total_boxes + best_bpt - 1) / best_bpt;
  }
  static const int o_bpt = env_int("VTB_WG_BOXES");
  if (o_bpt > 0 && o_bpt <= max_boxes) { best_bpt = std::min(o_bpt, w.total_boxes); best_tiles = (w.total_boxes + best_bpt - 1) / best_bpt; }
  w.boxes_per_tile = best_bpt;
  w.n_tiles = best_tiles;
  w.n_cols = best_bpt * w.cc;
  w.m_tiles = (c->cout + 127) / 128;
  // independent accumulation chains: as many as TMEM holds (an MMA chain into one accumulator has ~170 cycles latency)
  w.acc_stride = (w.n_cols + 31) & ~31;
  int ks = 8;
  while (ks > 1 && ks * w.acc_stride > 512) ks >>= 1;
  static const int o_ks = env_int("VTB_WG_KSPLIT");
  if (o_ks > 0 && o_ks * w.acc_stride <= 512) ks = o_ks;
  w.ksplit = ks;
  // pixels per stage: the largest box that still leaves >= 2 stages (fewer, larger TMA requests: the per-SM TMA
  // ingest rate grows with the box height, profiles/r01_tma_bw.txt)
  const int budget = kSmemBudget - 256 - 1024;
  auto stage_bytes = [&](int kp) { return ((kp * w.ma * 2 + 1023) / 1024 + (kp * w.n_cols * 2 + 1023) / 1024) * 1024; };
  int kpix = 256;
  while (kpix > 64 && budget / stage_bytes(kpix) < 2) kpix >>= 1;
  static const int o_kp = env_int("VTB_WG_KPIX");
  if ((o_kp == 64 || o_kp == 128 || o_kp == 256) && budget / stage_bytes(o_kp) >= 1) kpix = o_kp;
  while (kpix > 64 && w.mpix < (long long)kpix * 2) kpix >>= 1;
  w.kpix = kpix;
  w.stages = std::max(1, std::min(8, budget / stage_bytes(kpix)));
  static const int o_st = env_int("VTB_WG_STAGES");
  if (o_st > 0) w.stages = std::min(w.stages, o_st);
  w.kblocks = (int)((w.mpix + kpix - 1) / kpix);
  const int ctas = w.n_tiles * w.m_tiles;
  int splits = std::max(1, sms / ctas);
  splits = std::min(splits, std::max(1, w.kblocks / 2));
  w.splits = splits;
  return w;
}

}  // namespace vtb

using namespace vtb;

extern "C" {

const char* vtb_last_error(void) { return g_err; }
// development aid, not part of include/vtb.h: device buffer of [grid][8] role wait-cycle counters (nullptr = off)
void vtb_debug_counters(unsigned long long* dev_buf) { g_dbg = dev_buf; }
int vtb_version(void) { return 100; }
int vtb_num_sms(void) { return num_sms(); }
long long vtb_launch_count(void) { return g_launches.load(); }

int vtb_conv_out_hw(const VtbConv* c, int* ho, int* wo) {
  if (!conv_ok(c) || !ho || !wo) return fail(VTB_EINVAL, "vtb_conv_out_hw: bad geometry");
  out_hw(c, ho, wo);
  return VTB_OK;
}

int vtb_conv_stats_rows(const VtbConv* c) {
  if (!conv_ok(c)) return fail(VTB_EINVAL, "vtb_conv_stats_rows: bad geometry");
  int ho, wo;
  out_hw(c, &ho, &wo);
  return plan_conv_tiling((long long)c->n * ho * wo, c->cout, c->k * c->k * c->cin / 16).stat_rows;
}

size_t vtb_conv_wgrad_workspace_bytes(const VtbConv* c) {
  if (!conv_ok(c)) return 0;
  const WgradPlan w = plan_wgrad(c);
  return (size_t)w.splits * c->cout * c->k * c->k * c->cin * sizeof(float);
}

int vtb_conv_tiling_info(const VtbConv* c, int op, int* info) {
  if (!conv_ok(c) || !info || op < 0 || op > 2) return fail(VTB_EINVAL, "vtb_conv_tiling_info: bad arguments");
  int ho, wo;
  out_hw(c, &ho, &wo);
  if (op == 2) {
    const WgradPlan w = plan_wgrad(c);
    const int v[8] = {w.kpix, w.n_cols, w.n_tiles * w.m_tiles, w.splits, w.kblocks, w.ksplit, w.stages, 0};
    for (int i = 0; i < 8; ++i) info[i] = v[i];
    return VTB_OK;
  }
  long long M = (long long)c->n * ho * wo;
  int ncols = c->cout, k16 = c->k * c->k * c->cin / 16;
  if (op == 1) {
    ncols = c->cin;
    if (c->stride == 1) {
      M = (long long)c->n * c->h * c->w;
      k16 = c->k * c->k * c->cout / 16;
    } else {   // phase (1,1): the most taps
      M = (long long)c->n * (c->h / 2) * (c->w / 2);
      int nr = 0;
      for (int r = 0; r < c->k; ++r) nr += (((1 + c->pad - r) % 2 + 2) % 2 == 0);
      k16 = nr * nr * c->cout / 16;
      if (M <= 0) M = 1;
    }
  }
  const ConvTiling t = plan_conv_tiling(M, ncols, k16);
  const int halves = t.block_m / kBlockM, acc_stride = (t.block_n + 31) & ~31;
  const long long tiles = ((M + t.block_m - 1) / t.block_m) * t.n_blocks;
  const int v[8] = {t.block_m, t.block_n, t.grid, (int)tiles, (int)((tiles + t.grid - 1) / t.grid),
                    (2 * halves * t.ksplit * acc_stride <= 512) ? 2 : 1, t.ksplit, t.stages};
  for (int i = 0; i < 8; ++i) info[i] = v[i];
  return VTB_OK;
}

int vtb_pack_weight(const VtbConv* c, const float* w_oihw, int cin_real, void* wf, void* wd, void* stream) {
  if (!conv_ok(c) || !w_oihw || !wf || cin_real <= 0 || cin_real > c->cin)
    return fail(VTB_EINVAL, "vtb_pack_weight: bad arguments");
  const long long total = (long long)c->cout * c->k * c->k * c->cin;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 4096);
  count_launch(1);
  return check_cuda((int)launch_pdl(pack_weight_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, w_oihw, c->cout,
                                    cin_real, c->cin, c->k * c->k, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd),
                    "pack_weight_kernel");
}

long long vtb_pack_job_blocks(int cout, int cin, int kk) {
  if (cout <= 0 || cin <= 0 || kk <= 0) return fail(VTB_EINVAL, "vtb_pack_job_blocks: bad arguments");
  if (kk > 36) return fail(VTB_EINVAL, "vtb_pack_job_blocks: kernels larger than 6x6 are not supported");
  const int cit = pack_ci_tile(kk);
  return (long long)((cout + kPackCoT - 1) / kPackCoT) * ((cin + cit - 1) / cit);
}

int vtb_pack_weights(const VtbPackJob* jobs_device, int njobs, long long total_blocks, void* stream) {
  if (!jobs_device || njobs <= 0 || njobs > kPackMaxJobs || total_blocks <= 0)
    return fail(VTB_EINVAL, "vtb_pack_weights: bad arguments (at most 256 jobs per launch)");
  count_launch(1);
  const long long grid = std::min<long long>(total_blocks, (long long)std::max(1, num_sms()) * 8);
  return check_cuda((int)launch_pdl(pack_weights_batched_kernel<false>, dim3((unsigned)grid), dim3(256), 0,
                                    (cudaStream_t)stream, jobs_device, njobs, total_blocks, (const float*)nullptr),
                    "pack_weights_batched_kernel");
}

int vtb_sgd_pack_weights(const VtbPackJob* jobs_device, int njobs, long long total_blocks, const float* hyper_device,
                         void* stream) {
  if (!jobs_device || njobs <= 0 || njobs > kPackMaxJobs || total_blocks <= 0 || !hyper_device)
    return fail(VTB_EINVAL, "vtb_sgd_pack_weights: bad arguments (at most 256 jobs per launch)");
  count_launch(1);
  const long long grid = std::min<long long>(total_blocks, (long long)std::max(1, num_sms()) * 8);
  return check_cuda((int)launch_pdl(pack_weights_batched_kernel<true>, dim3((unsigned)grid), dim3(256), 0,
                                    (cudaStream_t)stream, jobs_device, njobs, total_blocks, hyper_device),
                    "pack_weights_batched_kernel(sgd)");
}

long long vtb_sgd_job_blocks(long long n) { return n <= 0 ? 0 : (n + kSgdBlockElems - 1) / kSgdBlockElems; }

int vtb_sgd_step(const VtbSgdJob* jobs_device, int njobs, long long total_blocks, const float* hyper_device, void* stream) {
  if (!jobs_device || njobs <= 0 || njobs > kPackMaxJobs || total_blocks <= 0 || !hyper_device)
    return fail(VTB_EINVAL, "vtb_sgd_step: bad arguments (at most 256 jobs per launch)");
  count_launch(1);
  const long long grid = std::min<long long>(total_blocks, (long long)std::max(1, num_sms()) * 8);
  return check_cuda((int)launch_pdl(sgd_step_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, jobs_device,
                                    njobs, total_blocks, hyper_device),
                    "sgd_step_kernel");
}

static int fprop_impl(const VtbConv* c, const void* x, int ldx, const void* wf, void* y, int ldy, float* stats_partial,
                      const float* scale, const float* shift, int relu, const void* residual, int ldr,
                      const VtbBnTrain* bn, void* stream) {
  if (!conv_ok(c) || !x || !wf || !y) return fail(VTB_EINVAL, "vtb_conv_fprop: bad arguments");
  if (bn && (!stats_partial || !bn->tickets || !bn->gamma || !bn->beta || !bn->mean || !bn->invstd || !bn->scale ||
             !bn->shift || bn->count <= 0 || ((bn->running_mean == nullptr) != (bn->running_var == nullptr))))
    return fail(VTB_EINVAL, "vtb_conv_fprop_bn: bad BatchNorm arguments");
  if (ldx < c->cin || ldy < c->cout || ldx % 8 || ldy % 8) return fail(VTB_EINVAL, "vtb_conv_fprop: bad pitch");
  if ((scale == nullptr) != (shift == nullptr)) return fail(VTB_EINVAL, "vtb_conv_fprop: scale/shift must pair");
  if (residual && (ldr < c->cout || ldr % 8)) return fail(VTB_EINVAL, "vtb_conv_fprop: bad residual pitch");
  if (!driver_api().ok) return fail(VTB_ENODEV, "cuTensorMapEncode* not available from this driver");
  int ho, wo;
  out_hw(c, &ho, &wo);
  const int taps = c->k * c->k;
  ConvIgemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = c->n * ho * wo;
  p.Wq = wo;
  p.Hp = ho;
  p.stride = c->stride;
  p.lower_w = p.lower_h = -c->pad;
  p.ntaps = taps;
  p.cin = c->cin;
  p.kc = chunk_of(c->cin);
  p.cout = c->cout;
  const ConvTiling tl = plan_conv_tiling(p.M, c->cout, taps * c->cin / 16);
  apply_tiling(p, tl);
  p.a_tiled = (c->k == 1 && c->stride == 1 && c->pad == 0) ? 1 : 0;
  for (int r = 0; r < c->k; ++r)
    for (int s = 0; s < c->k; ++s) {
      const int t = r * c->k + s;
      p.tap_ow[t] = (uint16_t)s;
      p.tap_oh[t] = (uint16_t)r;
      p.tap_kofs[t] = t * c->cin;
    }
  p.store_mode = kStoreTma;
  p.out = (__nv_bfloat16*)y;
  p.ldo = ldy;
  p.stats_partial = stats_partial;
  if (bn) {
    p.tickets = bn->tickets;
    p.bn_count = bn->count;
    p.bn_gamma = bn->gamma;
    p.bn_beta = bn->beta;
    p.bn_eps = bn->eps;
    p.bn_momentum = bn->momentum;
    p.bn_running_mean = bn->running_mean;
    p.bn_running_var = bn->running_var;
    p.bn_nbt = bn->num_batches_tracked;
    p.bn_mean = bn->mean;
    p.bn_invstd = bn->invstd;
    p.bn_scale = bn->scale;
    p.bn_shift = bn->shift;
    p.bn_split = bn->split;
    if (bn->split > 0) {
      if (bn->split >= c->cout || !bn->gamma2 || !bn->beta2 || ((bn->running_mean2 == nullptr) != (bn->running_var2 == nullptr)) ||
          ((bn->running_mean == nullptr) != (bn->running_mean2 == nullptr)))
        return fail(VTB_EINVAL, "vtb_conv_fprop_bn: bad side-by-side BatchNorm arguments");
      p.bn_gamma2 = bn->gamma2;
      p.bn_beta2 = bn->beta2;
      p.bn_running_mean2 = bn->running_mean2;
      p.bn_running_var2 = bn->running_var2;
      p.bn_nbt2 = bn->num_batches_tracked2;
    }
    if (bn->sync && !sync_args_ok(bn->sync, c->cout)) return fail(VTB_EINVAL, "vtb_conv_fprop_bn: bad SyncBN peers");
    p.sync = make_sync_peers(bn->sync);
    if (bn->act_out != nullptr) {   // fused normalise (+ReLU, + residual) in the same launch
      if (bn->split > 0 || bn->act_ld < c->cout || bn->act_ld % 8 || (reinterpret_cast<uintptr_t>(bn->act_out) & 15) ||
          (bn->act_residual && (bn->act_ldr < c->cout || bn->act_ldr % 8 ||
                                (reinterpret_cast<uintptr_t>(bn->act_residual) & 15))) ||
          c->cout / tl.block_n > 64 || tl.grid > std::max(1, num_sms()))
        return fail(VTB_EINVAL, "vtb_conv_fprop_bn: bad fused-normalise arguments");
      p.fn_out = (__nv_bfloat16*)bn->act_out;
      p.fn_ldo = bn->act_ld;
      relu = bn->act_relu;
      residual = bn->act_residual;
      ldr = bn->act_ldr;
    }
  }
  p.scale = scale;
  p.shift = shift;
  p.relu = relu;
  p.residual = (const __nv_bfloat16*)residual;
  p.ldr = ldr;
  const int upper = c->pad - (c->k - 1);
  CUtensorMap tmA, tmB, tmD;
  if (p.a_tiled) {
    if (!tmap_tiled_2d(&tmA, x, c->cin, p.M, (uint64_t)ldx * 2, p.kc, p.block_m, p.kc * 2))
      return fail(VTB_ECUDA, "vtb_conv_fprop: tensor map for x failed");
  } else if (!tmap_im2col_nhwc(&tmA, x, c->cin, c->w, c->h, c->n, ldx, -c->pad, -c->pad, upper, upper, p.kc,
                               p.block_m, c->stride, p.kc * 2))
    return fail(VTB_ECUDA, "vtb_conv_fprop: im2col tensor map for x failed");
  if (!tmap_tiled_2d(&tmB, wf, (uint64_t)taps * c->cin, c->cout, (uint64_t)taps * c->cin * 2, p.kc, p.block_n,
                     p.kc * 2))
    return fail(VTB_ECUDA, "vtb_conv_fprop: tensor map for weights failed");
  if (!tmap_tiled_2d(&tmD, y, c->cout, p.M, (uint64_t)ldy * 2, p.panel_w, 32, p.panel_w * 2))
    return fail(VTB_ECUDA, "vtb_conv_fprop: tensor map for y failed");
  count_launch(1);
  return check_cuda(launch_conv_igemm(tmA, tmB, tmD, p, tl.grid, (cudaStream_t)stream), "conv_igemm_kernel(fprop)");
}

int vtb_conv_fprop(const VtbConv* c, const void* x, int ldx, const void* wf, void* y, int ldy, float* stats_partial,
                   const float* scale, const float* shift, int relu, const void* residual, int ldr, void* stream) {
  return fprop_impl(c, x, ldx, wf, y, ldy, stats_partial, scale, shift, relu, residual, ldr, nullptr, stream);
}

int vtb_conv_fprop_bn(const VtbConv* c, const void* x, int ldx, const void* wf, void* y, int ldy, float* stats_partial,
                      const VtbBnTrain* bn, void* stream) {
  if (!bn) return fail(VTB_EINVAL, "vtb_conv_fprop_bn: bn must not be NULL");
  return fprop_impl(c, x, ldx, wf, y, ldy, stats_partial, nullptr, nullptr, 0, nullptr, 0, bn, stream);
}

// tiling of the dgrad launch(es) of a geometry: stride 1 -> one launch; stride 2 -> one per output-parity phase
static int dgrad_phase_tiling(const VtbConv* c, int ph, int pq, ConvTiling* tl, int* nr_out, int* ns_out) {
  const int k = c->k, pad = c->pad;
  if (c->stride == 1) {
    *tl = plan_conv_tiling((long long)c->n * c->h * c->w, c->cin, k * k * c->cout / 16);
    return 1;
  }
  const int Hph = (c->h - ph + 1) / 2, Wph = (c->w - pq + 1) / 2;
  if (Hph <= 0 || Wph <= 0) return 0;
  int nr = 0, ns = 0;
  for (int r = 0; r < k; ++r) nr += (((ph + pad - r) % 2 + 2) % 2 == 0);
  for (int s = 0; s < k; ++s) ns += (((pq + pad - s) % 2 + 2) % 2 == 0);
  if (nr_out) *nr_out = nr;
  if (ns_out) *ns_out = ns;
  if (nr == 0 || ns == 0) return -1;
  *tl = plan_conv_tiling((long long)c->n * Hph * Wph, c->cin, nr * ns * c->cout / 16);
  return 1;
}

// copies the BatchNorm-backward statistics request into the kernel parameters; `panel_w`: a staged panel must not
// straddle the boundary between the two producer layers
static int apply_dgrad_bn(ConvIgemmParams& p, const VtbConv* c, const VtbDgradBn* bn, int panel_w) {
  if (!bn->partial || bn->count <= 0 || bn->split < 0 || bn->split >= c->cin || (bn->split % panel_w) != 0)
    return fail(VTB_EINVAL, "vtb_conv_dgrad_bn: bad statistics arguments (split must be a multiple of %d)", panel_w);
  const int nl = bn->split > 0 ? 2 : 1;
  for (int l = 0; l < nl; ++l) {
    const VtbBnBwdLayer& L = bn->layer[l];
    const int lc = (nl == 1) ? c->cin : (l == 0 ? bn->split : c->cin - bn->split);
    if (!L.y || L.ldy < lc || L.ldy % 8 || (reinterpret_cast<uintptr_t>(L.y) & 15) || !L.scale || !L.shift || !L.mean ||
        !L.invstd || !L.coef)
      return fail(VTB_EINVAL, "vtb_conv_dgrad_bn: bad producer layer %d", l);
    p.bwd_y[l] = (const __nv_bfloat16*)L.y;
    p.bwd_ldy[l] = L.ldy;
    p.bwd_scale[l] = L.scale;
    p.bwd_shift[l] = L.shift;
    p.bwd_mean[l] = L.mean;
    p.bwd_invstd[l] = L.invstd;
    p.bwd_relu[l] = L.relu;
    p.bwd_dgamma[l] = L.dgamma;
    p.bwd_dbeta[l] = L.dbeta;
    p.bwd_coef[l] = L.coef;
  }
  p.bwd_split = bn->split;
  p.stats_partial = bn->partial;
  p.bn_count = bn->count;
  if (bn->sync && !sync_args_ok(bn->sync, c->cin)) return fail(VTB_EINVAL, "vtb_conv_dgrad_bn: bad SyncBN peers");
  p.sync = make_sync_peers(bn->sync);
  return VTB_OK;
}

static int dgrad_impl(const VtbConv* c, const void* dy, int lddy, const void* wd, void* dx, int lddx, int accumulate,
                      const VtbDgradBn* bn, void* stream) {
  if (!conv_ok(c) || !dy || !wd || !dx) return fail(VTB_EINVAL, "vtb_conv_dgrad: bad arguments");
  if (bn && !bn->tickets) return fail(VTB_EINVAL, "vtb_conv_dgrad_bn: tickets must not be NULL");
  if (lddy < c->cout || lddx < c->cin || lddy % 8 || lddx % 8) return fail(VTB_EINVAL, "vtb_conv_dgrad: bad pitch");
  if (!driver_api().ok) return fail(VTB_ENODEV, "cuTensorMapEncode* not available from this driver");
  int ho, wo;
  out_hw(c, &ho, &wo);
  const int k = c->k, pad = c->pad;
  ConvIgemmParams p;
  memset(&p, 0, sizeof(p));
  // GEMM: dX[pixels][cin] = im2col(dY)[pixels][taps*cout] * Wd[cin][taps*cout]^T
  p.cin = c->cout;   // channels per tap of the activation operand (dY)
  p.cout = c->cin;   // GEMM N
  p.kc = chunk_of(c->cout);
  CUtensorMap tmA, tmB, tmD;

  if (c->stride == 1) {
    if (ho != c->h + 2 * pad - k + 1) return fail(VTB_EINVAL, "vtb_conv_dgrad: internal shape error");
    const int lo = -(k - 1 - pad);
    const int up = -pad;  // positions = ho + up - lo = ho + k - 1 - 2*pad = h
    p.M = c->n * c->h * c->w;
    const ConvTiling tl = plan_conv_tiling(p.M, c->cin, k * k * c->cout / 16);
    apply_tiling(p, tl);
    if (!tmap_tiled_2d(&tmB, wd, (uint64_t)k * k * c->cout, c->cin, (uint64_t)k * k * c->cout * 2, p.kc, p.block_n,
                       p.kc * 2))
      return fail(VTB_ECUDA, "vtb_conv_dgrad: tensor map for weights failed");
    p.a_tiled = (k == 1 && pad == 0) ? 1 : 0;
    p.Wq = c->w;
    p.Hp = c->h;
    p.stride = 1;
    p.lower_w = p.lower_h = lo;
    p.ntaps = k * k;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        const int t = r * k + s;
        p.tap_ow[t] = (uint16_t)(k - 1 - s);
        p.tap_oh[t] = (uint16_t)(k - 1 - r);
        p.tap_kofs[t] = t * c->cout;
      }
    p.store_mode = accumulate ? kStoreTmaAdd : kStoreTma;
    p.out = (__nv_bfloat16*)dx;
    p.ldo = lddx;
    if (p.a_tiled) {
      if (!tmap_tiled_2d(&tmA, dy, c->cout, p.M, (uint64_t)lddy * 2, p.kc, p.block_m, p.kc * 2))
        return fail(VTB_ECUDA, "vtb_conv_dgrad: tensor map for dy failed");
    } else if (!tmap_im2col_nhwc(&tmA, dy, c->cout, wo, ho, c->n, lddy, lo, lo, up, up, p.kc, p.block_m, 1, p.kc * 2))
      return fail(VTB_ECUDA, "vtb_conv_dgrad: im2col tensor map for dy failed");
    if (!tmap_tiled_2d(&tmD, dx, c->cin, p.M, (uint64_t)lddx * 2, p.panel_w, 32, p.panel_w * 2))
      return fail(VTB_ECUDA, "vtb_conv_dgrad: tensor map for dx failed");
    if (bn) {
      if (int e = apply_dgrad_bn(p, c, bn, p.panel_w)) return e;
      p.tickets = bn->tickets;
    }
    count_launch(1);
    return check_cuda(launch_conv_igemm(tmA, tmB, tmD, p, tl.grid, (cudaStream_t)stream),
                      "conv_igemm_kernel(dgrad s1)");
  }

  // statistics rows of the phase launches are stacked; the last launch reduces all of them
  int last_ph = -1, total_rows = 0;
  if (bn) {
    for (int ph = 0; ph < 2; ++ph)
      for (int pq = 0; pq < 2; ++pq) {
        ConvTiling tl;
        if (dgrad_phase_tiling(c, ph, pq, &tl, nullptr, nullptr) > 0) {
          last_ph = ph * 2 + pq;
          total_rows += tl.stat_rows;
        }
      }
  }
  int row0 = 0;

  // stride 2: one launch per output-parity phase (ph, pw); each is a dense stride-1 walk over dY.
  for (int ph = 0; ph < 2; ++ph) {
    for (int pq = 0; pq < 2; ++pq) {
      const int Hph = (c->h - ph + 1) / 2, Wph = (c->w - pq + 1) / 2;
      if (Hph <= 0 || Wph <= 0) continue;
      // taps r with (ph + pad - r) even: ho = h' + (ph + pad - r)/2
      int rs[8], dhs[8], nr = 0, ss[8], dws[8], ns = 0;
      for (int r = 0; r < k; ++r)
        if (((ph + pad - r) % 2 + 2) % 2 == 0) {
          rs[nr] = r;
          dhs[nr++] = (ph + pad - r) / 2;  // exact: numerator even
        }
      for (int s = 0; s < k; ++s)
        if (((pq + pad - s) % 2 + 2) % 2 == 0) {
          ss[ns] = s;
          dws[ns++] = (pq + pad - s) / 2;
        }
      ConvIgemmParams q = p;
      q.M = c->n * Hph * Wph;
      const ConvTiling tl = plan_conv_tiling(q.M, c->cin, nr * ns * c->cout / 16);
      apply_tiling(q, tl);
      if (!tmap_tiled_2d(&tmB, wd, (uint64_t)k * k * c->cout, c->cin, (uint64_t)k * k * c->cout * 2, q.kc, q.block_n,
                         q.kc * 2))
        return fail(VTB_ECUDA, "vtb_conv_dgrad: tensor map for weights failed");
      tmD = tmB;  // unused in scatter mode, but must be a valid map
      q.Wq = Wph;
      q.Hp = Hph;
      q.stride = 1;
      q.store_mode = accumulate ? kStoreScatterAdd : kStoreScatter;
      q.out = (__nv_bfloat16*)dx;
      q.OH = c->h;
      q.OW = c->w;
      q.os = 2;
      q.os_w = 2;
      q.oph = ph;
      q.opw = pq;
      q.ldo = lddx;
      if (nr == 0 || ns == 0) {
        // no contributing taps: this phase of dx is zero. Not reachable for k=3,p=1 / k=1,p=0 handled below.
        return fail(VTB_EINVAL, "vtb_conv_dgrad: stride-2 geometry with an empty phase is unsupported");
      }
      int lo_h = dhs[0], lo_w = dws[0];
      for (int i = 0; i < nr; ++i) lo_h = std::min(lo_h, dhs[i]);
      for (int i = 0; i < ns; ++i) lo_w = std::min(lo_w, dws[i]);
      q.lower_h = lo_h;
      q.lower_w = lo_w;
      q.ntaps = nr * ns;
      for (int i = 0; i < nr; ++i)
        for (int j = 0; j < ns; ++j) {
          const int t = i * ns + j;
          q.tap_oh[t] = (uint16_t)(dhs[i] - lo_h);
          q.tap_ow[t] = (uint16_t)(dws[j] - lo_w);
          q.tap_kofs[t] = (rs[i] * k + ss[j]) * c->cout;
        }
      const int up_h = Hph - ho + lo_h, up_w = Wph - wo + lo_w;
      if (!tmap_im2col_nhwc(&tmA, dy, c->cout, wo, ho, c->n, lddy, lo_w, lo_h, up_w, up_h, q.kc, q.block_m, 1,
                            q.kc * 2))
        return fail(VTB_ECUDA, "vtb_conv_dgrad: im2col tensor map for dy (phase %d,%d) failed", ph, pq);
      if (bn) {
        if (int e = apply_dgrad_bn(q, c, bn, q.panel_w)) return e;
        q.stats_row0 = row0;
        row0 += tl.stat_rows;
        if (ph * 2 + pq == last_ph) {
          q.tickets = bn->tickets;
          q.fin_rows = total_rows;
        }
      }
      count_launch(1);
      int e = launch_conv_igemm(tmA, tmB, tmD, q, tl.grid, (cudaStream_t)stream);
      if (e) return check_cuda(e, "conv_igemm_kernel(dgrad s2)");
    }
  }
  return VTB_OK;
}

int vtb_conv_dgrad(const VtbConv* c, const void* dy, int lddy, const void* wd, void* dx, int lddx, int accumulate,
                   void* stream) {
  return dgrad_impl(c, dy, lddy, wd, dx, lddx, accumulate, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------
// Stride-2 dgrad (k = 3, pad = 1, even H and W) as ONE dense GEMM over 2x2 "super-pixels" instead of four phase launches
// that each stream the whole dY:   dX[n][2i+a][2j+b][ci] = sum_{dh,dw in {0,1}} sum_co dY[n][i+dh][j+dw][co] * W'[(a,b,ci)][(dh,dw,co)]
// with W'[(a,b,ci)][(dh,dw,co)] = W[co][ci][r(a,dh)][s(b,dw)], r(0,0) = 1, r(1,1) = 0, r(1,0) = 2, (0,1): no such tap -> 0
// (7 of the 16 (a,b,dh,dw) blocks are zero: 16/9 of the minimal FLOPs - worth it only where the layer is HBM-bound, i.e.
// few channels).  M = N*(H/2)*(W/2) super-pixels, GEMM N = 4*Cin as two n-blocks (a = 0, 1) of 2*Cin columns (b, ci):
// with lddx == Cin those are the 2*Cin contiguous elements of pixels (2i+a, 2j) and (2i+a, 2j+1).
// ---------------------------------------------------------------------------------------------
__global__ void pack_dgrad_s2_kernel(const __nv_bfloat16* __restrict__ wd, int cin, int cout, __nv_bfloat16* __restrict__ wp) {
  pdl_wait();
  pdl_trigger();
  const long long total = 16LL * cin * cout;
  const int K = 4 * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kcol = (int)(i % K);
    const int row = (int)(i / K);
    const int co = kcol % cout, t = kcol / cout, dh = t >> 1, dw = t & 1;
    const int ci = row % cin, ab = row / cin, a = ab >> 1, b = ab & 1;
    const int r = (a == 0) ? (dh == 0 ? 1 : -1) : (dh == 1 ? 0 : 2);
    const int sx = (b == 0) ? (dw == 0 ? 1 : -1) : (dw == 1 ? 0 : 2);
    wp[i] = (r < 0 || sx < 0) ? __float2bfloat16(0.f) : wd[((size_t)ci * 9 + (r * 3 + sx)) * cout + co];
  }
}

static bool dgrad_s2_ok(const VtbConv* c) {
  return conv_ok(c) && c->k == 3 && c->stride == 2 && c->pad == 1 && c->h % 2 == 0 && c->w % 2 == 0 && c->cin <= 64 &&
         block_of(2 * c->cin) == 2 * c->cin;
}

size_t vtb_conv_dgrad_s2_workspace_bytes(const VtbConv* c) {
  return dgrad_s2_ok(c) ? (size_t)16 * c->cin * c->cout * sizeof(__nv_bfloat16) : 0;
}

int vtb_conv_dgrad_s2(const VtbConv* c, const void* dy, int lddy, const void* wd, void* workspace, void* dx, int lddx,
                      int accumulate, void* stream) {
  if (!dgrad_s2_ok(c) || !dy || !wd || !workspace || !dx) return fail(VTB_EINVAL, "vtb_conv_dgrad_s2: unsupported geometry or bad arguments");
  if (lddy < c->cout || lddy % 8 || lddx != c->cin) return fail(VTB_EINVAL, "vtb_conv_dgrad_s2: bad pitch (dx must be dense: lddx == cin)");
  if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) return fail(VTB_EINVAL, "vtb_conv_dgrad_s2: workspace must be 128-byte aligned");
  if (!driver_api().ok) return fail(VTB_ENODEV, "cuTensorMapEncode* not available from this driver");
  int ho, wo;
  out_hw(c, &ho, &wo);   // == h/2, w/2
  cudaStream_t st = (cudaStream_t)stream;
  const long long welems = 16LL * c->cin * c->cout;
  launch_pdl(pack_dgrad_s2_kernel, dim3((unsigned)std::min<long long>((welems + 255) / 256, 1024)), dim3(256), 0, st,
             (const __nv_bfloat16*)wd, c->cin, c->cout, (__nv_bfloat16*)workspace);
  ConvIgemmParams p;
  memset(&p, 0, sizeof(p));
  p.cin = c->cout;          // channels per window tap of the activation operand (dY)
  p.cout = 4 * c->cin;      // GEMM N
  p.kc = chunk_of(c->cout);
  p.M = c->n * ho * wo;
  ConvTiling tl = plan_conv_tiling(p.M, 2 * c->cin, 4 * c->cout / 16);
  if (tl.block_n != 2 * c->cin) return fail(VTB_EINVAL, "vtb_conv_dgrad_s2: internal tiling error");
  {
    const int sms = std::max(1, num_sms() > 0 ? num_sms() : 148);
    const long long tiles = ((p.M + tl.block_m - 1) / tl.block_m) * 2;
    tl.n_blocks = 2;
    tl.grid = (int)std::max<long long>(2, std::min<long long>(tiles, sms) / 2 * 2);
  }
  apply_tiling(p, tl);
  p.a_tiled = 0;
  p.Wq = wo;
  p.Hp = ho;
  p.stride = 1;
  p.lower_w = p.lower_h = 0;
  p.ntaps = 4;
  for (int t = 0; t < 4; ++t) {
    p.tap_oh[t] = (uint16_t)(t >> 1);
    p.tap_ow[t] = (uint16_t)(t & 1);
    p.tap_kofs[t] = t * c->cout;
  }
  p.store_mode = accumulate ? kStoreScatterAdd : kStoreScatter;
  p.out = (__nv_bfloat16*)dx;
  p.OH = c->h;
  p.OW = c->w / 2;          // super-pixel columns
  p.os = 2;                 // row step; the n-block index supplies the row parity (merge_n)
  p.os_w = 1;
  p.oph = 0;
  p.opw = 0;
  p.ldo = 2 * lddx;         // one super-pixel = two pixels
  p.merge_n = 1;
  CUtensorMap tmA, tmB, tmD;
  if (!tmap_tiled_2d(&tmB, workspace, (uint64_t)4 * c->cout, (uint64_t)4 * c->cin, (uint64_t)4 * c->cout * 2, p.kc, p.block_n, p.kc * 2))
    return fail(VTB_ECUDA, "vtb_conv_dgrad_s2: tensor map for the merged weights failed");
  tmD = tmB;   // unused in scatter mode, but must be a valid map
  // window taps (dh, dw) in {0,1}^2 over dY, rows / columns past the edge read as zero (lower corner 0, upper corner 0)
  if (!tmap_im2col_nhwc(&tmA, dy, c->cout, wo, ho, c->n, lddy, 0, 0, 0, 0, p.kc, p.block_m, 1, p.kc * 2))
    return fail(VTB_ECUDA, "vtb_conv_dgrad_s2: im2col tensor map for dy failed");
  count_launch(2);
  return check_cuda(launch_conv_igemm(tmA, tmB, tmD, p, tl.grid, st), "conv_igemm_kernel(dgrad s2 merged)");
}

int vtb_conv_dgrad_bn(const VtbConv* c, const void* dy, int lddy, const void* wd, void* dx, int lddx, int accumulate,
                      const VtbDgradBn* bn, void* stream) {
  if (!bn) return fail(VTB_EINVAL, "vtb_conv_dgrad_bn: bn must not be NULL");
  return dgrad_impl(c, dy, lddy, wd, dx, lddx, accumulate, bn, stream);
}

int vtb_conv_dgrad_stats_rows(const VtbConv* c) {
  if (!conv_ok(c)) return fail(VTB_EINVAL, "vtb_conv_dgrad_stats_rows: bad geometry");
  int rows = 0;
  for (int ph = 0; ph < (c->stride == 1 ? 1 : 2); ++ph)
    for (int pq = 0; pq < (c->stride == 1 ? 1 : 2); ++pq) {
      ConvTiling tl;
      if (dgrad_phase_tiling(c, ph, pq, &tl, nullptr, nullptr) > 0) rows += tl.stat_rows;
    }
  return rows;
}

// panel width of the dgrad epilogue: VtbDgradBn.split must be a multiple of it
int vtb_conv_dgrad_panel_w(const VtbConv* c) {
  if (!conv_ok(c)) return fail(VTB_EINVAL, "vtb_conv_dgrad_panel_w: bad geometry");
  ConvTiling tl;
  int w = 0;
  for (int ph = 0; ph < (c->stride == 1 ? 1 : 2); ++ph)
    for (int pq = 0; pq < (c->stride == 1 ? 1 : 2); ++pq)
      if (dgrad_phase_tiling(c, ph, pq, &tl, nullptr, nullptr) > 0) w = std::max(w, tl.panel_w);
  return w;
}

static int wgrad_impl(const VtbConv* c, const void* dy, int lddy, const void* x, int ldx, void* workspace,
                      float* dw_oihw, float* dw2, int split, int cin_real, int accumulate, void* stream) {
  if (!conv_ok(c) || !dy || !x || !workspace || !dw_oihw || cin_real <= 0 || cin_real > c->cin)
    return fail(VTB_EINVAL, "vtb_conv_wgrad: bad arguments");
  if (lddy < c->cout || ldx < c->cin || lddy % 8 || ldx % 8) return fail(VTB_EINVAL, "vtb_conv_wgrad: bad pitch");
  if (!driver_api().ok) return fail(VTB_ENODEV, "cuTensorMapEncode* not available from this driver");
  int ho, wo;
  out_hw(c, &ho, &wo);
  const WgradPlan w = plan_wgrad(c);
  WgradIgemmParams p;
  memset(&p, 0, sizeof(p));
  p.Mpix = (int)w.mpix;
  p.Wq = wo;
  p.Hp = ho;
  p.stride = c->stride;
  p.lower_w = p.lower_h = -c->pad;
  p.cout = c->cout;
  p.cin = c->cin;
  p.ntaps = c->k * c->k;
  p.cc = w.cc;
  p.ca = w.ca;
  p.ma = w.ma;
  p.kpix = w.kpix;
  p.boxes_per_tap = w.boxes_per_tap;
  p.total_boxes = w.total_boxes;
  p.boxes_per_tile = w.boxes_per_tile;
  p.n_cols = w.n_cols;
  p.n_tiles = w.n_tiles;
  p.ksplit = w.ksplit;
  p.acc_stride = (uint32_t)w.acc_stride;
  p.splits = w.splits;
  p.kblocks = w.kblocks;
  p.num_stages = w.stages;
  for (int r = 0; r < c->k; ++r)
    for (int s = 0; s < c->k; ++s) {
      p.tap_ow[r * c->k + s] = (uint16_t)s;
      p.tap_oh[r * c->k + s] = (uint16_t)r;
    }
  p.ws = (float*)workspace;
  p.dbg = g_dbg;
  const int upper = c->pad - (c->k - 1);
  CUtensorMap tmDY, tmX;
  if (!tmap_tiled_2d(&tmDY, dy, c->cout, (uint64_t)w.mpix, (uint64_t)lddy * 2, w.ca, w.kpix, w.ca * 2))
    return fail(VTB_ECUDA, "vtb_conv_wgrad: tensor map for dy failed");
  // (a tiled [pixels][cin] map for the X operand of 1x1 layers was measured: no difference, profiles/r02_wgrad_1x1_xtiled_sweep.txt)
  if (!tmap_im2col_nhwc(&tmX, x, c->cin, c->w, c->h, c->n, ldx, -c->pad, -c->pad, upper, upper, w.cc, w.kpix,
                        c->stride, w.cc * 2))
    return fail(VTB_ECUDA, "vtb_conv_wgrad: im2col tensor map for x failed");
  count_launch(2);
  int e = launch_wgrad_igemm(tmDY, tmX, p, w.m_tiles * w.n_tiles, (cudaStream_t)stream);
  if (e) return check_cuda(e, "wgrad_igemm_kernel");
  static const int no_reduce = env_int("VTB_WG_NOREDUCE");   // development: time the GEMM alone
  if (no_reduce) return VTB_OK;
  const int ew = c->cin <= 16 ? 16 : (c->cin <= 32 ? 32 : 64);
  const dim3 rgrid(c->cout, (c->cin + ew - 1) / ew);
  const size_t rsmem = (size_t)(256 / ew) * p.ntaps * (ew + 1) * sizeof(float);
  const float* wsf = (const float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t re;
  const bool deep = (long long)p.ntaps * w.splits >= 16 * (256 / ew);
#define VTB_REDUCE(EWV, DEEPV)                                                                                              \
  launch_pdl(wgrad_reduce_kernel<EWV, DEEPV>, rgrid, dim3(256), rsmem, st, wsf, w.splits, c->cout, c->cin, cin_real, p.ntaps, \
             dw_oihw, accumulate, dw2, split)
  if (ew == 16) re = deep ? VTB_REDUCE(16, true) : VTB_REDUCE(16, false);
  else if (ew == 32) re = deep ? VTB_REDUCE(32, true) : VTB_REDUCE(32, false);
  else re = deep ? VTB_REDUCE(64, true) : VTB_REDUCE(64, false);
#undef VTB_REDUCE
  return check_cuda((int)re, "wgrad_reduce_kernel");
}

int vtb_conv_wgrad(const VtbConv* c, const void* dy, int lddy, const void* x, int ldx, void* workspace,
                   float* dw_oihw, int cin_real, int accumulate, void* stream) {
  return wgrad_impl(c, dy, lddy, x, ldx, workspace, dw_oihw, nullptr, 0, cin_real, accumulate, stream);
}

int vtb_conv_wgrad_pair(const VtbConv* c, const void* dy, int lddy, const void* x, int ldx, void* workspace,
                        float* dw_a, float* dw_b, int split, int cin_real, int accumulate, void* stream) {
  if (!c || !dw_b || split <= 0 || split >= c->cout) return fail(VTB_EINVAL, "vtb_conv_wgrad_pair: bad split");
  return wgrad_impl(c, dy, lddy, x, ldx, workspace, dw_a, dw_b, split, cin_real, accumulate, stream);
}

}  // extern "C"
