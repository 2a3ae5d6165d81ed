// Implicit-GEMM convolution kernels for sm_100a (tcgen05 + TMEM + TMA), shared declarations.
//
//  conv_igemm_kernel : D[M pixels][Cout] = im2col(Act)[M][taps*Cin] * Wp[Cout][taps*Cin]^T
//      used for fprop (Act = X), stride-1 dgrad (Act = dY, flipped/transposed weights) and the four
//      parity phases of a stride-2 dgrad (Act = dY, tap subsets, strided scatter of the result).
//  wgrad_igemm_kernel: dW[Cout][tap][Cin] = sum_pixels dY[pix][Cout] * im2col(X)[pix][tap][Cin]
//      (both operands MN-major, split over the pixel axis).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vtb {

constexpr int kBlockM = 128;       // output pixels per tile (UMMA M)
constexpr int kStageK = 64;        // K elements (bf16) per pipeline stage
constexpr int kMaxTaps = 36;       // 6x6 filter
constexpr int kNumThreads = 192;   // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;

enum StoreMode : int {
  kStoreTma = 0,        // bf16 tile via TMA store
  kStoreTmaAdd = 1,     // bf16 tile via TMA reduce-add (gradient fan-in)
  kStoreScatter = 2,    // bf16 rows written by threads to a strided pixel lattice (stride-2 dgrad phases)
  kStoreScatterAdd = 3,
};

struct ConvIgemmParams {
  // GEMM / pixel space of the output of this launch
  int M;            // number of output pixels handled by this launch (Nimg*Hp*Wq)
  int Wq, Hp;       // output-pixel lattice (q fastest, then p, then image)
  int stride;       // traversal stride of the im2col walk
  int lower_w, lower_h;  // coordinate of the base pixel for q=0 / p=0 (== pixelBoxLowerCorner)
  int ntaps;
  int cin;          // channels per tap of the activation operand
  int kc;           // channels per TMA sub-load (16/32/64), divides cin
  int block_n;      // UMMA N (multiple of 16, <= 256), divides cout
  int cout;
  int num_stages;
  uint16_t tap_ow[kMaxTaps];
  uint16_t tap_oh[kMaxTaps];
  int tap_kofs[kMaxTaps];   // first K index of this tap inside the packed weight matrix
  // epilogue
  int store_mode;
  int panel_w;      // columns per staged output panel (16/32/64), divides block_n
  // scatter mode: pixel (img,p,q) -> out + ((img*OH + p*os + oph)*OW + q*os + opw)*ldo
  __nv_bfloat16* out;
  int OH, OW, os, oph, opw, ldo;
  // per-channel batch statistics of the bf16-rounded result: [gridDim.x][groups][cout][2] (sum, sumsq)
  float* stats_partial;
  // optional fused per-channel affine + ReLU (+ residual) epilogue (eval-mode folded BN)
  const float* scale;
  const float* shift;
  int relu;
  const __nv_bfloat16* residual;
  int ldr;
};

struct WgradIgemmParams {
  int Mpix;          // total pixels of dY (GEMM K)
  int Wq, Hp;        // dY pixel lattice
  int stride, lower_w, lower_h;
  int cout, cin;
  int ntaps;         // taps handled per n-tile group table below
  int cc;            // channels per B sub-tile box (16/32/64)
  int ca;            // channels per A (dY) box (16/32/64)
  int sub_n;         // UMMA N per MMA (= min(cin, 256) slice width)
  int subs_per_tile; // number of (tap, c0) sub-tiles per CTA n-tile; subs_per_tile*sub_n <= 512
  int n_tiles;       // CTA n-tiles
  int total_subs;    // ntaps * (cin / sub_n)
  int splits;        // split-K factor over pixel blocks
  int kblocks;       // ceil(Mpix / 64)
  int num_stages;
  uint16_t tap_ow[kMaxTaps];
  uint16_t tap_oh[kMaxTaps];
  float* ws;         // [splits][cout][ntaps*cin] fp32 partials
};

size_t conv_igemm_smem_bytes(int block_n, int num_stages);
size_t wgrad_igemm_smem_bytes(int subs_per_tile, int sub_n, int num_stages);

// host launchers (igemm.cu); return cudaError_t as int
int launch_conv_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                      const ConvIgemmParams& p, int grid, cudaStream_t stream);
int launch_wgrad_igemm(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradIgemmParams& p, int grid_x,
                       cudaStream_t stream);

}  // namespace vtb
