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

#include "syncbn.cuh"

namespace vtb {

constexpr int kBlockM = 128;       // rows per UMMA (M); a conv tile is 1 or 2 such halves (block_m 128 / 256)
constexpr int kStageK = 64;        // K elements (bf16) per pipeline stage
constexpr int kMaxTaps = 36;       // 6x6 filter
constexpr int kWgProducers = 8;    // wgrad kernel: TMA producer warps
constexpr int kWgThreads = (kWgProducers + 1 + 4) * 32;  // + MMA warp + 4 epilogue warps
constexpr int kConvThreads = 640;  // conv kernel: warp0/3 A producers, warp1 MMA, warp2 B producer, warps 4-19 epilogue (4 groups)
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = 128;   // threads per epilogue group
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448;
constexpr int kMaxPanels = 8;      // staged output panels per accumulator (block_n / panel_w)

enum StoreMode : int {   // every mode stages the bf16 tile in shared memory and writes 16-byte row segments
  kStoreTma = 0,        // dense rows
  kStoreTmaAdd = 1,     // dense rows, read-modify-write (gradient fan-in)
  kStoreScatter = 2,    // rows go to a strided pixel lattice (stride-2 dgrad phases)
  kStoreScatterAdd = 3,
};

struct ConvIgemmParams {
  // GEMM / pixel space of the output of this launch
  int M;            // number of output pixels handled by this launch (Nimg*Hp*Wq)
  int block_m;      // 128 or 256 output pixels per tile (256 = two UMMA halves sharing one B tile)
  int a_tiled;      // 1: A operand is a plain [pixels][cin] matrix loaded in tiled mode (1x1 stride-1), 0: im2col
  int panel_bufs;   // staged-output buffers per epilogue group (1 or 2)
  int ksplit;       // accumulators per 128-row half, fed round-robin with successive K slices (1, 2 or 4)
  // values derived on the host (fill_derived) so that no kernel role has to keep division results in registers
  int halves, n_blocks, num_m_blocks, m_step, chunks_per_tap, total_chunks, subs_per_stage, num_k_stages;
  uint32_t acc_stride, set_cols, nsets;                          // TMEM layout
  // tile-parallel epilogue (narrow tiles, two TMEM sets): epilogue groups {0,1} drain the even tiles of the CTA (set 0),
  // groups {2,3} the odd ones (set 1) - two tiles' TMEM -> staging -> global chains overlap instead of one tile being
  // cut into panels too small to hide the chain's latency
  int tile_par;
  uint32_t a_stage, b_stage, b_off, panel_off, bar_off;          // shared-memory layout (bytes from the aligned base)
  uint32_t a_sub_bytes, a_half_bytes, b_sub_bytes;
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
  // scatter mode: pixel (img,p,q) -> out + ((img*OH + p*os + oph)*OW + q*os_w + opw)*ldo   (os: row step, os_w: column step)
  // merge_n != 0 (stride-2 dgrad as ONE GEMM over 2x2 super-pixels, api_conv.cu): n-block b writes output row parity
  // oph + b and its block_n columns start at column 0 of the (two-pixel) super-pixel, i.e. at out - b * block_n
  __nv_bfloat16* out;
  int OH, OW, os, os_w, oph, opw, ldo, merge_n;
  // per-channel batch statistics of the bf16-rounded result: [rows][cout][2] (sum, sumsq),
  // rows = gridDim.x / n_blocks: one per CTA; every (row, channel) is written exactly once
  float* stats_partial;
  // optional in-kernel BatchNorm finalisation by the last CTA of every n-block (tickets != nullptr): [n_blocks]
  // zero-initialised counters (left zero again), the per-channel element count and the BatchNorm tensors
  unsigned int* tickets;
  double bn_count;
  const float* bn_gamma;
  const float* bn_beta;
  float bn_eps, bn_momentum;
  float* bn_running_mean;
  float* bn_running_var;
  long long* bn_nbt;
  float* bn_mean;
  float* bn_invstd;
  float* bn_scale;
  float* bn_shift;
  // two BatchNorm layers side by side on the channel axis (two convolutions of the same input run as ONE launch: CSP
  // conv1 | conv2): channels >= bn_split use the second set of parameter tensors (indexed from 0); 0 = single layer
  int bn_split;
  const float* bn_gamma2;
  const float* bn_beta2;
  float* bn_running_mean2;
  float* bn_running_var2;
  long long* bn_nbt2;
  // Fused normalise (fprop with in-kernel finalisation only; fn_out != nullptr): once the coefficients of its n-block are
  // final, every CTA re-reads the raw tiles it produced (L2) and writes fn_out = [relu](y*scale+shift) [+ residual] - the
  // unit's vtb_bn_act pass without its launch.  Uses `relu`, `residual`, `ldr` below and tickets[128 + n_blk] ("ready"),
  // tickets[192 + n_blk] ("departed"); needs every CTA of the launch co-resident (grid <= SMs, one CTA per SM).
  __nv_bfloat16* fn_out;
  int fn_ldo;
  // SyncBN: peer-mapped exchange buffers (world <= 1: single-GPU statistics); bn_count is then the GLOBAL count and
  // tickets[64] counts the finished n-block exchanges of the launch
  SyncPeers sync;
  // BatchNorm(+ReLU) BACKWARD statistics (dgrad launches, bwd_y[0] != nullptr): this launch completes the gradient `g` of
  // a tensor that was produced by one ConvNormAct (or two, side by side on the channel axis: channels >= bwd_split belong
  // to the second, indexed from 0).  The epilogue reads the producer's raw conv output y at the pixels it stores and
  // accumulates, per channel, sum(dz) and sum(dz * y) with dz = g * (y*scale + shift > 0) into stats_partial
  // rows [stats_row0, stats_row0 + m_step); with `tickets` the last CTA of each n-block reduces rows [0, fin_rows)
  // (a stride-2 dgrad is four launches: only the last one finalises), exchanges them under SyncBN and writes
  // coef[c][2] = (mean dz, mean dz*xhat), dgamma = sum dz*xhat, dbeta = sum dz (local sums): the standalone BatchNorm
  // backward of that layer then is a single apply pass (vtb_bn_bwd_apply).
  int bwd_split;
  const __nv_bfloat16* bwd_y[2];
  int bwd_ldy[2];
  const float* bwd_scale[2];
  const float* bwd_shift[2];
  const float* bwd_mean[2];
  const float* bwd_invstd[2];
  int bwd_relu[2];
  float* bwd_dgamma[2];
  float* bwd_dbeta[2];
  float* bwd_coef[2];
  int stats_row0, fin_rows;
  // optional fused per-channel affine + ReLU (+ residual) epilogue (eval-mode folded BN)
  const float* scale;
  const float* shift;
  int relu;
  const __nv_bfloat16* residual;
  int ldr;
  // development aid (tools/bench_conv): per-CTA [8] cycle counters of the time each role spent waiting; usually null
  unsigned long long* dbg;
};

struct WgradIgemmParams {
  int Mpix;          // total pixels of dY (GEMM K)
  int Wq, Hp;        // dY pixel lattice
  int stride, lower_w, lower_h;
  int cout, cin;
  int ntaps;
  int ca;            // channels per dY box (16/32/64)
  int cc;            // channels per X box (16/32/64); the flattened tap*cin column axis is tiled in boxes of cc
  int ma;            // dY channels staged per stage: min(128, cout)
  int kpix;          // pixels (GEMM K) per pipeline stage = TMA box rows (64/128/256)
  int boxes_per_tap; // cin / cc
  int total_boxes;   // ntaps * boxes_per_tap
  int boxes_per_tile;// X boxes per CTA tile: n_cols = boxes_per_tile * cc <= 256
  int n_cols;
  int n_tiles;       // CTA tiles along the column axis
  int ksplit;        // TMEM accumulators fed round-robin (power of two, ksplit * acc_stride <= 512)
  uint32_t acc_stride;
  int splits;        // split-K factor over pixel blocks
  int kblocks;       // ceil(Mpix / kpix)
  int num_stages;
  uint16_t tap_ow[kMaxTaps];
  uint16_t tap_oh[kMaxTaps];
  float* ws;         // [splits][cout][ntaps*cin] fp32 partials
  unsigned long long* dbg;   // development aid (tools/bench_conv): per-CTA [16] cycle counters; usually null
};

size_t conv_igemm_smem_bytes(int block_m, int block_n, int num_stages, int panel_bufs);
size_t wgrad_igemm_smem_bytes(int ma, int n_cols, int kpix, int num_stages);

// completes the derived fields of ConvIgemmParams from block_m/block_n/num_stages/ksplit/kc/cin/ntaps/M/cout and grid
void fill_derived(ConvIgemmParams& p, int grid);

// host launchers (igemm.cu); return cudaError_t as int
int launch_conv_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                      const ConvIgemmParams& p, int grid, cudaStream_t stream);
int launch_wgrad_igemm(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradIgemmParams& p, int grid_x,
                       cudaStream_t stream);

}  // namespace vtb
