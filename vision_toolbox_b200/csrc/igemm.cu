// tcgen05 implicit-GEMM convolution kernels (see igemm.cuh for the math each one computes).
//
// Both kernels are warp-specialised: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer
// (one lane), warps 2..5 = epilogue (TMEM -> registers -> smem/global).  Operands are staged in shared
// memory by TMA (im2col mode for the activation operand, so a tile of 128 consecutive output pixels may
// cross row and image boundaries and padding is zero-filled by hardware), accumulators live in TMEM.
#include "common.cuh"
#include "igemm.cuh"
#include "syncbn.cuh"
#include "ptx.cuh"

namespace vtb {

static __host__ __device__ inline int round_up_int(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (host + device agree through these helpers)
// ------------------------------------------------------------------------------------------------
struct ConvSmem {
  uint32_t a_off, b_off, panel_off, bar_off, total;
  uint32_t a_stage, b_stage;
};
static __host__ __device__ inline ConvSmem conv_smem_layout(int block_m, int block_n, int num_stages, int panel_bufs) {
  ConvSmem s;
  s.a_stage = block_m * kStageK * 2;
  s.b_stage = round_up_int(block_n * kStageK * 2, 1024);
  s.a_off = 0;
  s.b_off = num_stages * s.a_stage;
  s.panel_off = s.b_off + num_stages * s.b_stage;
  s.bar_off = s.panel_off + kEpiWarps * 2048;  // one 32-row x 64 B staging buffer per epilogue warp
  s.total = s.bar_off + 256;
  return s;
}
size_t conv_igemm_smem_bytes(int block_m, int block_n, int num_stages, int panel_bufs) {
  return conv_smem_layout(block_m, block_n, num_stages, panel_bufs).total + 1024;  // + slack for 1024B alignment
}

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }


// Per-warp drain of a staged bf16 panel of 32 rows x (8*CH) columns (row pitch 16*CH bytes, XOR-swizzled 16-byte
// chunks): lane = rg * CH + chunk reads chunk `chunk` of the CH rows of row group rg with 16-byte loads and writes them to
// global memory at pixel index pix[r] (< 0: row outside the tensor) - a warp instruction covers whole 16*CH-byte row
// segments -> coalesced; dense tiles pass consecutive pixel indices, stride-2 dgrad phases a strided lattice.
//   ADD    : read-modify-write for gradient fan-in, rounding like two bf16 tensors being added
//   MODE 1 : BatchNorm forward statistics of the stored values: S = sum v, Q = sum v^2
//   MODE 2 : BatchNorm(+ReLU) backward statistics against the producer's raw output y (BwdCols): the stored value is the
//            final gradient g; dz = g * (y*scale + shift > 0); S = sum dz, Q = sum dz * y (the finaliser subtracts
//            mean * S in double: two per-channel constants in registers instead of three)
// A recursive-halving butterfly across the row groups then leaves the 32-row (S, Q) totals of
//   CH == 4: column chunk*8 + (rg&1)*4 + ((rg>>1)&1)*2 + (rg>>2)      CH == 2: same formula on rg bits 0..2; lanes with
//   rg bit 3 clear hold the result                                                                   -> (o0, o1)
struct BwdCols {            // per-panel view of the producer layer: pointers already offset to the panel's first column
  const __nv_bfloat16* y;
  int ldy;
  const float* scale;
  const float* shift;
  int relu;
};
// `early` (MODE 2): the producer's y rows, loaded by the caller BEFORE staging the panel (panel_early_load) so that their
// latency overlaps the TMEM -> shared-memory staging.
template <int CH, bool ADD, int MODE>
__device__ __forceinline__ void panel_early_load(int lane, const __nv_bfloat16* gcol, int ld, const int (&pix)[4],
                                                 const BwdCols& bw, uint4 (&early)[4]) {
  const int chunk = lane % CH;
  if (MODE == 2) {
#pragma unroll
    for (int r = 0; r < CH; ++r)
      early[r] = (pix[r] >= 0) ? __ldg(reinterpret_cast<const uint4*>(bw.y + (size_t)pix[r] * bw.ldy + chunk * 8))
                               : make_uint4(0u, 0u, 0u, 0u);
  }
}
// lane-level butterfly over the row-group bits of the lane index (see panel_drain): 8 (S, Q) pairs per lane -> the totals
// of one column per lane
template <int CH>
__device__ __forceinline__ void stats_butterfly(int lane, float (&S)[8], float (&Q)[8]) {
  constexpr int RG = 32 / CH;
  int n = 8;
#pragma unroll
  for (int step = 0; (1 << step) < RG; ++step) {
    const int partner_bit = CH << step;
    const bool upper = (lane & partner_bit) != 0;
    if (n > 1) {
      const int hn = n / 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < hn) {
          // send the half I do not keep, receive the partner's copy of the half I keep
          const float sendS = upper ? S[j] : S[j + hn], sendQ = upper ? Q[j] : Q[j + hn];
          const float keepS = upper ? S[j + hn] : S[j], keepQ = upper ? Q[j + hn] : Q[j];
          S[j] = keepS + __shfl_xor_sync(0xffffffffu, sendS, partner_bit);
          Q[j] = keepQ + __shfl_xor_sync(0xffffffffu, sendQ, partner_bit);
        }
      }
      n = hn;
    } else {
      S[0] += __shfl_xor_sync(0xffffffffu, S[0], partner_bit);
      Q[0] += __shfl_xor_sync(0xffffffffu, Q[0], partner_bit);
    }
  }
}

// DEFER: the caller keeps this lane's 8 (S, Q) pairs in registers across ALL tiles of the CTA (accS / accQ) and runs the
// butterfly once at the end - possible when every drain of the warp covers the same columns (one local panel).
// Addressing is prepared by the caller ONCE per CTA / per tile (the drain is issue-bound: every integer instruction saved
// here is saved 8 x per 256 x 128 tile and warp): dsm0 / dsm1 = shared-memory addresses of this lane's chunk in its first
// / third staged row (the swizzle XOR is folded in; rows 1 and 3 sit one row pitch behind), rowoff[r] = element offset of
// row r in the destination (pixel index * pitch; 0xFFFFFFFF: row outside the tensor).
template <int CH, bool ADD, int MODE, bool DEFER = false>
__device__ __forceinline__ void panel_drain(uint32_t dsm0, uint32_t dsm1, int lane, __nv_bfloat16* gcol,
                                            const uint32_t (&rowoff)[4], const int (&pix)[4], const BwdCols& bw,
                                            const uint4 (&early)[4], float& o0, float& o1, float (&accS)[8],
                                            float (&accQ)[8]) {
  constexpr int ROWS = CH;             // rows per group
  constexpr uint32_t PITCH = 16 * CH;
  const int chunk = lane % CH;
  __nv_bfloat16* gbase = gcol + chunk * 8;
  // sums of the lane's 8 channels as packed fp32 pairs: element 2j in the low word, 2j+1 in the high word
  uint64_t S2[4], Q2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    S2[j] = DEFER ? f2_pack(accS[2 * j], accS[2 * j + 1]) : 0ull;
    Q2[j] = DEFER ? f2_pack(accQ[2 * j], accQ[2 * j + 1]) : 0ull;
  }
  float sc[MODE == 2 ? 8 : 1], sh[MODE == 2 ? 8 : 1];
  if (MODE == 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(bw.scale + chunk * 8) + h);
      const float4 b = __ldg(reinterpret_cast<const float4*>(bw.shift + chunk * 8) + h);
      sc[4 * h] = a.x; sc[4 * h + 1] = a.y; sc[4 * h + 2] = a.z; sc[4 * h + 3] = a.w;
      sh[4 * h] = b.x; sh[4 * h + 1] = b.y; sh[4 * h + 2] = b.z; sh[4 * h + 3] = b.w;
    }
  }
  // ADD: the old destination rows are read here, every read of a batch issued before its first use so that the round
  // trips overlap (the drain is latency-bound per warp); MODE 2 + ADD: two rows at a time (register budget)
  constexpr int RB = (ADD && MODE == 2) ? (ROWS > 2 ? 2 : ROWS) : ROWS;
#pragma unroll
  for (int r0 = 0; r0 < ROWS; r0 += RB) {
    uint4 old[ADD ? RB : 1];
    if (ADD) {
#pragma unroll
      for (int r = 0; r < RB; ++r)
        old[r] = (rowoff[r0 + r] != 0xFFFFFFFFu) ? *reinterpret_cast<const uint4*>(gbase + rowoff[r0 + r])
                                                  : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const uint32_t ro = rowoff[r0 + r];
      const int px = (ro != 0xFFFFFFFFu) ? 0 : -1;
      uint32_t w[4];
      const uint32_t sa = (((r0 + r) & 2) ? dsm1 : dsm0) + ((r0 + r) & 1) * PITCH;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(sa));
      if (px >= 0) {
        uint4* gp = reinterpret_cast<uint4*>(gbase + ro);
        if (ADD) {
          const uint4 ov = old[r];
          const uint32_t oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(bf16lo(w[j]) + bf16lo(oo[j]), bf16hi(w[j]) + bf16hi(oo[j]));
        }
        *gp = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (MODE == 1) {   // rows outside the tensor hold zeros (TMA zero-fills the operand rows): they add nothing
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t v2 = f2_pack(bf16lo(w[j]), bf16hi(w[j]));
          S2[j] = f2_add(S2[j], v2);
          Q2[j] = f2_fma(v2, v2, Q2[j]);
        }
      }
      if (MODE == 2) {
        if (px >= 0) {
          const uint4 yy4 = early[r0 + r];
          const uint32_t yy[4] = {yy4.x, yy4.y, yy4.z, yy4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float dzv[2], yv2[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int e = 2 * j + hh;
              const float g = hh ? bf16hi(w[j]) : bf16lo(w[j]);
              const float y = hh ? bf16hi(yy[j]) : bf16lo(yy[j]);
              const float z = fmaf(y, sc[e], sh[e]);
              dzv[hh] = (!bw.relu || z > 0.f) ? g : 0.f;
              yv2[hh] = y;
            }
            const uint64_t d2 = f2_pack(dzv[0], dzv[1]);
            S2[j] = f2_add(S2[j], d2);
            Q2[j] = f2_fma(d2, f2_pack(yv2[0], yv2[1]), Q2[j]);
          }
        }
      }
    }
  }
  float S[8], Q[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f2_unpack(S2[j], S[2 * j], S[2 * j + 1]);
    f2_unpack(Q2[j], Q[2 * j], Q[2 * j + 1]);
  }
  if (DEFER) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      accS[j] = S[j];
      accQ[j] = Q[j];
    }
    return;
  }
  if (MODE != 0) stats_butterfly<CH>(lane, S, Q);
  o0 = S[0];
  o1 = Q[0];
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// mbarrier wait that optionally accounts the stalled cycles (development counters, ConvIgemmParams::dbg)
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool timed, long long& acc) {
  if (!timed) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

// ------------------------------------------------------------------------------------------------
// Fused normalise (+ReLU, + residual) tail of a training fprop (ConvIgemmParams::fn_out): called by the 16 epilogue warps
// after the statistics / finalisation code.  The coefficients of an n-block are final once its last CTA has written them;
// every CTA of the n-block then applies them to the tiles it produced itself (still in L2).  __noinline__: its registers
// must not count against the main loop's.  Everything it needs comes BY VALUE: a reference to the kernel's parameter
// block turns every field access into a generic load (ncu: the top stall of the first version).
// ------------------------------------------------------------------------------------------------
struct FnArgs {
  const __nv_bfloat16* y;          // raw conv output (this launch's `out`)
  const __nv_bfloat16* res;        // residual or nullptr
  __nv_bfloat16* out;              // destination view
  const float* scale;
  const float* shift;
  unsigned int* tickets;
  int ldy, ldr, ldo, relu, block_m, block_n, M, nmb, mstep;   // nmb / mstep: num_m_blocks / m_step of the launch
};
__device__ __noinline__ void fused_normalise_tail(FnArgs a, int et, int n_blk, int n0, int m_first, bool last, float2* slots) {
  constexpr int NT = kEpiWarps * 32;
  volatile unsigned int* ready = a.tickets + 128 + n_blk;
  unsigned int* depart = a.tickets + 192 + n_blk;
  if (last) {
    __threadfence();                         // scale / shift (threads et < block_n) visible device-wide
    named_bar_sync(1, NT);
    if (et == 0) *ready = 1u;
  } else if (et == 0) {
    const long long t0 = clock64();
    while (*ready == 0u) {
      if (clock64() - t0 > 200000000000LL) __trap();   // ~100 s: a co-residency assumption was violated
    }
    __threadfence();
  }
  named_bar_sync(1, NT);
  // All tiles of this CTA form ONE index space, walked with FU independent 16-byte loads in flight per thread: a per-tile
  // loop would expose one memory round trip per tile (16 KB tiles on the few-channel layers).
  const int cpr = a.block_n / 8;                                         // 16-byte chunks per tile row
  const int my_tiles = (a.nmb - m_first + a.mstep - 1) / a.mstep;
  const int rows_total = my_tiles * a.block_m;                           // rows of all my tiles, tile after tile
  const bool has_res = a.res != nullptr;
  const bool relu = a.relu != 0;
  auto apply = [&](const uint4& yq, const uint4& rq, const float (&sc)[8], const float (&sh)[8], uint32_t pix, int cc) {
    const uint32_t yw[4] = {yq.x, yq.y, yq.z, yq.w};
    const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f0 = fmaf(bf16lo(yw[j]), sc[2 * j], sh[2 * j]), f1 = fmaf(bf16hi(yw[j]), sc[2 * j + 1], sh[2 * j + 1]);
      if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
      if (has_res) {   // the reference adds two bf16 tensors: round first, then add
        f0 = __bfloat162float(__float2bfloat16_rn(f0)) + bf16lo(rw[j]);
        f1 = __bfloat162float(__float2bfloat16_rn(f1)) + bf16hi(rw[j]);
      }
      ow[j] = pack_bf16x2(f0, f1);
    }
    *reinterpret_cast<uint4*>(a.out + (size_t)pix * a.ldo + n0 + cc) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  // row index over my tiles -> pixel (block_m is 128 or 256)
  const int bm_shift = (a.block_m == 256) ? 8 : 7;
  auto pixel_of = [&](int rowg) -> uint32_t {
    const int tl = rowg >> bm_shift;
    return (uint32_t)((m_first + tl * a.mstep) * a.block_m + (rowg & (a.block_m - 1)));
  };
  if ((cpr & (cpr - 1)) == 0 && cpr <= NT) {
    // power-of-two tile width: the channel chunk of a thread is loop-invariant (coefficients in registers) and its rows
    // advance by a constant step - no division in the loop
    const int cc = (et & (cpr - 1)) * 8;
    const int rstep = NT / cpr;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldcg(a.scale + n0 + cc + j);
      sh[j] = __ldcg(a.shift + n0 + cc + j);
    }
    const __nv_bfloat16* ycol = a.y + n0 + cc;
    if (!has_res) {
      constexpr int FU = 8;
      for (int r0 = et / cpr; r0 < rows_total; r0 += rstep * FU) {
        uint4 yv[FU];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int rg = r0 + u * rstep;
          const uint32_t pix = pixel_of(rg);
          yv[u] = (rg < rows_total && pix < (uint32_t)a.M) ? __ldcg(reinterpret_cast<const uint4*>(ycol + (size_t)pix * a.ldy))
                                                           : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int rg = r0 + u * rstep;
          const uint32_t pix = pixel_of(rg);
          if (rg < rows_total && pix < (uint32_t)a.M) apply(yv[u], make_uint4(0u, 0u, 0u, 0u), sc, sh, pix, cc);
        }
      }
    } else {
      constexpr int FU = 4;
      const __nv_bfloat16* rcol = a.res + n0 + cc;
      for (int r0 = et / cpr; r0 < rows_total; r0 += rstep * FU) {
        uint4 yv[FU], rv[FU];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int rg = r0 + u * rstep;
          const uint32_t pix = pixel_of(rg);
          const bool ok = rg < rows_total && pix < (uint32_t)a.M;
          yv[u] = ok ? __ldcg(reinterpret_cast<const uint4*>(ycol + (size_t)pix * a.ldy)) : make_uint4(0u, 0u, 0u, 0u);
          rv[u] = ok ? __ldg(reinterpret_cast<const uint4*>(rcol + (size_t)pix * a.ldr)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
          const int rg = r0 + u * rstep;
          const uint32_t pix = pixel_of(rg);
          if (rg < rows_total && pix < (uint32_t)a.M) apply(yv[u], rv[u], sc, sh, pix, cc);
        }
      }
    }
  } else {
    // general tile width (160 / 192 / 224 channels): coefficients staged in shared memory, one item at a time
    float2* cf = slots;   // [block_n] (scale, shift)
    for (int cidx = et; cidx < a.block_n; cidx += NT)
      cf[cidx] = make_float2(__ldcg(a.scale + n0 + cidx), __ldcg(a.shift + n0 + cidx));
    named_bar_sync(1, NT);
    const int total_items = rows_total * cpr;
    for (int gi = et; gi < total_items; gi += NT) {
      const int rg = gi / cpr;
      const int cc = (gi - rg * cpr) * 8;
      const uint32_t pix = pixel_of(rg);
      if (pix >= (uint32_t)a.M) continue;
      const uint4 yq = __ldcg(reinterpret_cast<const uint4*>(a.y + (size_t)pix * a.ldy + n0 + cc));
      const uint4 rq = has_res ? __ldg(reinterpret_cast<const uint4*>(a.res + (size_t)pix * a.ldr + n0 + cc))
                               : make_uint4(0u, 0u, 0u, 0u);
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 c = cf[cc + j];
        sc[j] = c.x;
        sh[j] = c.y;
      }
      apply(yq, rq, sc, sh, pix, cc);
    }
  }
  // the last CTA of the n-block to get here re-arms the flags (nobody spins on `ready` any more)
  named_bar_sync(1, NT);
  if (et == 0 && atomicAdd(depart, 1u) == (unsigned int)(a.mstep - 1)) {
    *depart = 0u;
    *ready = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// conv_igemm_kernel
//
// Persistent, warp-specialised. One tile = block_m (128 | 256) output pixels x block_n channels; a CTA keeps ONE
// n-block for its whole life (tile = m_blk * n_blocks + n_blk with n_blk = blockIdx.x % n_blocks), so the weight
// tile stays hot in L2 and the per-channel statistics can be accumulated in registers across tiles.
//
// TMA issue rate, not bandwidth, bounds the operand feed on B200 (profiles/r01_tma_bw.txt: one im2col request
// stream moves a <=256-row box per ~625 cycles whatever its size), hence: 256-row A boxes, and THREE independent
// producer threads (A even stages / A odd stages / B).
//   warp 0, 3 : A producers (im2col or tiled TMA), alternating pipeline stages
//   warp 2    : B (weight) producer
//   warp 1    : TMEM owner + MMA issuer (two M=128 UMMAs per K step when block_m == 256, same B descriptor)
//   warps 4-7 / 8-11 : two epilogue groups, each draining its own 128-row accumulator
// TMEM: back-to-back MMAs into ONE accumulator serialise on a ~170-cycle dependency (profiles/r01_mma_rate.txt), so
// every 128-row half owns `ksplit` accumulators fed round-robin with successive K slices (summed in the epilogue):
// chains = halves * ksplit independent accumulation chains of block_n columns form one "set"; 512 / set columns
// (max 2) sets give double buffering against the epilogue when the tile is narrow.
// ------------------------------------------------------------------------------------------------
// EPI selects the epilogue flavour at compile time (each instantiation only carries the registers it needs):
//   kEpiStats   dense store + BatchNorm statistics (training fprop)
//   kEpiAffine  dense store of [relu](acc*scale+shift)[+residual] (eval-mode fused fprop)
//   kEpiPlain   store or read-modify-write accumulate, dense or strided pixel lattice (dgrad, fprop without statistics)
//   kEpiBwd     kEpiPlain + BatchNorm(+ReLU) backward statistics of the producer layer(s) (ConvIgemmParams::bwd_*)
//   kEpiStats1  kEpiStats for tiles whose every epilogue warp owns ONE panel (block_n <= 64 with 256-row tiles, <= 128
//               with 128-row tiles): the per-lane sums stay in registers across tiles, one butterfly at the end
enum EpiKind : int { kEpiStats = 0, kEpiAffine = 1, kEpiPlain = 2, kEpiBwd = 3, kEpiStats1 = 4 };

template <bool TIMED, int EPI, int PW>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmD, const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.num_stages;
  // everything below is either a kernel parameter (constant bank, costs no register) or one add away from one
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.b_off;
  const uint32_t panel_base = base + p.panel_off;
  const uint32_t bar_base = base + p.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 4 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool timed = TIMED;   // development counters (ConvIgemmParams::dbg); compiled out of the production kernel
  const long long t_start = timed ? clock64() : 0;
  if (timed && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[blockIdx.x * 16 + 11] = gt;   // CTA start (ns)
  }

#define halves p.halves
#define ksplit p.ksplit
#define acc_stride p.acc_stride
#define set_cols p.set_cols
#define nsets p.nsets
#define num_m_blocks p.num_m_blocks
#define n_blocks p.n_blocks
#define m_step p.m_step
#define kc p.kc
#define chunks_per_tap p.chunks_per_tap
#define total_chunks p.total_chunks
#define subs_per_stage p.subs_per_stage
#define num_k_stages p.num_k_stages
#define a_sub_bytes p.a_sub_bytes
#define a_half_bytes p.a_half_bytes
#define b_sub_bytes p.b_sub_bytes
  const int n_blk = blockIdx.x % n_blocks;
  const int m_first = blockIdx.x / n_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 2);   // one arrive.expect_tx from the A producer of the stage, one from the B producer
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), p.tile_par ? kEpiWarps / 2 : kEpiWarps);   // one arrival per epilogue warp that drains the set
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
  pdl_trigger();

  if (warp == 0 || warp == 3) {
    // ======================= A producers =======================
    // elect.sync (not `lane == 0`): ptxas then knows a single thread is active and emits straight-line
    // UTMALDG / UTCHMMA with uniform-register operands instead of a per-instruction divergence ("waterfall") loop
    if (elect_one()) {
      const uint32_t who = (warp == 0) ? 0u : 1u;
      uint32_t stage = 0, phase = 0, g = 0;
      long long waited = 0;
      for (int m_blk = m_first; m_blk < num_m_blocks; m_blk += m_step) {
        const int m0 = m_blk * p.block_m;
        const int q0 = m0 % p.Wq;
        const int t = m0 / p.Wq;
        const int p0 = t % p.Hp;
        const int img = t / p.Hp;
        const int bw = p.lower_w + q0 * p.stride;
        const int bh = p.lower_h + p0 * p.stride;
        for (int ks = 0; ks < num_k_stages; ++ks, ++g) {
          if ((g & 1u) == who) {
            mbar_wait_t(empty_bar(stage), phase ^ 1u, timed, waited);
            const int j0 = ks * subs_per_stage;
            const int nsub = min(subs_per_stage, total_chunks - j0);
            mbar_expect_tx(full_bar(stage), nsub * a_sub_bytes);
            const uint32_t a_st = a_base + stage * p.a_stage;
            for (int sub = 0; sub < nsub; ++sub) {
              const int j = j0 + sub;
              const int tap = j / chunks_per_tap;
              const int c0 = (j - tap * chunks_per_tap) * kc;
              if (p.a_tiled)
                tma_load_2d(a_st + sub * a_sub_bytes, &tmA, full_bar(stage), c0, m0);
              else
                tma_load_im2col_4d(a_st + sub * a_sub_bytes, &tmA, full_bar(stage), c0, bw, bh, img, p.tap_ow[tap],
                                   p.tap_oh[tap]);
            }
          }
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
      }
      if (timed) p.dbg[blockIdx.x * 16 + who] = waited;
    }
  } else if (warp == 2) {
    // ======================= B producer =======================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      long long waited = 0;
      for (int m_blk = m_first; m_blk < num_m_blocks; m_blk += m_step) {
        for (int ks = 0; ks < num_k_stages; ++ks) {
          mbar_wait_t(empty_bar(stage), phase ^ 1u, timed, waited);
          const int j0 = ks * subs_per_stage;
          const int nsub = min(subs_per_stage, total_chunks - j0);
          mbar_expect_tx(full_bar(stage), nsub * b_sub_bytes);
          const uint32_t b_st = b_base + stage * p.b_stage;
          for (int sub = 0; sub < nsub; ++sub) {
            const int j = j0 + sub;
            const int tap = j / chunks_per_tap;
            const int c0 = (j - tap * chunks_per_tap) * kc;
            tma_load_2d(b_st + sub * b_sub_bytes, &tmB, full_bar(stage), p.tap_kofs[tap] + c0, n_blk * p.block_n);
          }
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
      }
      if (timed) p.dbg[blockIdx.x * 16 + 2] = waited;
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      // The issuing thread is on the critical path (one tcgen05.mma every ~64-150 cycles): keep the per-MMA
      // instruction count minimal - descriptors are a constant high word OR-ed with (smem address >> 4).
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, 0);
      const uint32_t ltype = (kc == 64) ? 2u : (kc == 32 ? 4u : 6u);
      const uint64_t desc_hi = make_smem_desc(0, 16, 8u * kc * 2u, ltype);
      const uint32_t ks_mask = (uint32_t)ksplit - 1u;          // ksplit is a power of two
      const uint32_t half_cols = (uint32_t)ksplit * acc_stride;  // TMEM column distance between the two halves
      const uint32_t a_sub16 = a_sub_bytes >> 4, b_sub16 = b_sub_bytes >> 4, a_half16 = a_half_bytes >> 4;
      const int k16_per_sub = kc / 16;
      uint32_t stage = 0, phase = 0;
      uint32_t lt = 0;  // local tile counter
      long long w_full = 0, w_tempty = 0;
      for (int m_blk = m_first; m_blk < num_m_blocks; m_blk += m_step, ++lt) {
        const uint32_t set = (nsets == 2u) ? (lt & 1u) : 0u;
        const uint32_t use = (nsets == 2u) ? (lt >> 1) : lt;
        mbar_wait_t(tempty_bar(set), (use & 1u) ^ 1u, timed, w_tempty);
        tc_fence_after();
        const uint32_t d_set = tmem_base + set * set_cols;
        uint32_t kk = 0;  // running K-slice (16 elements) counter of this tile
        for (int ks = 0; ks < num_k_stages; ++ks) {
          mbar_wait_t(full_bar(stage), phase, timed, w_full);
          tc_fence_after();
          const int nsub = min(subs_per_stage, total_chunks - ks * subs_per_stage);
          uint32_t a16 = (a_base + stage * p.a_stage) >> 4;
          uint32_t b16 = (b_base + stage * p.b_stage) >> 4;
          for (int sub = 0; sub < nsub; ++sub, a16 += a_sub16, b16 += b_sub16) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              if (k16 < k16_per_sub) {
                const uint64_t bdesc = desc_hi | (uint64_t)(b16 + 2u * k16);
                const uint32_t d0 = d_set + (kk & ks_mask) * acc_stride;
                const uint32_t accumulate = (kk > ks_mask) ? 1u : 0u;
                umma_bf16(d0, desc_hi | (uint64_t)(a16 + 2u * k16), bdesc, idesc, accumulate);
                if (halves == 2)
                  umma_bf16(d0 + half_cols, desc_hi | (uint64_t)(a16 + a_half16 + 2u * k16), bdesc, idesc, accumulate);
                ++kk;
              }
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(set));
      }
      if (timed) {
        p.dbg[blockIdx.x * 16 + 3] = w_full;
        p.dbg[blockIdx.x * 16 + 4] = w_tempty;
        p.dbg[blockIdx.x * 16 + 15] = clock64() - t_start;   // MMA issuer done (all MMAs issued)
      }
    }
  } else {
    // ======================= epilogue (4 groups x 4 warps, every warp independent) =======================
    // The drain is latency-bound per warp (dependent TMEM -> cvt -> smem -> global chains), so it is spread over 16
    // warps: group g owns (half g>>1, panels g&1, g&1 + 2, ...) of a 256-row tile or panels g, g+4, .. of a 128-row one.
    // A warp owns 32 accumulator rows (its TMEM lane quarter): TMEM -> registers -> its own swizzled 32-row smem
    // panel -> its own TMA store.  No block-level barriers: only __syncwarp and per-thread bulk-group waits.
    const int grp = (warp - 4) >> 2;            // epilogue group 0..3
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;        // row inside the 128-row half == TMEM lane
    constexpr int pw = PW;                      // staged panel width: 32 or 16 columns (compile time: the drain is issue-bound)
    constexpr uint32_t pitch = PW * 2;
    constexpr uint32_t smask = (PW == 32) ? 3u : 1u;
    constexpr int CH = PW / 8;                  // 16-byte chunks per staged row = rows each lane drains per panel
    constexpr bool do_stats = (EPI == kEpiStats || EPI == kEpiBwd || EPI == kEpiStats1);   // per-channel (S, Q) sums ride in the drain
    float accS[8], accQ[8];   // kEpiStats1: this lane's sums over every tile of the CTA
#pragma unroll
    for (int j = 0; j < 8; ++j) accS[j] = accQ[j] = 0.f;
    const bool scatter = (p.store_mode == kStoreScatter || p.store_mode == kStoreScatterAdd);
    const bool add = (p.store_mode == kStoreTmaAdd || p.store_mode == kStoreScatterAdd);
    const int npanels = p.block_n / pw;         // host guarantees npanels <= kMaxPanels
    const int n0 = n_blk * p.block_n;
    // this warp's staging buffer: 32 rows x 64 B (panels are at most 32 columns wide)
    const uint32_t my_panels = panel_base + (uint32_t)(warp - 4) * 2048u;
    // running (S, Q) totals of this lane's column for the whole CTA lifetime: up to 4 local panels x 2 values
    // (scalars, not an array: a runtime-indexed array would be demoted to local memory = L2 round trips here)
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    // SyncBN: the sequence number of this launch must be read before any exchange of the launch can complete
    unsigned int my_seq = 0;
    if (do_stats && p.sync.world > 1 && p.tickets != nullptr && threadIdx.x == 128) my_seq = sync_read_seq(p.sync);
    uint32_t lt = 0;
    long long w_tfull = 0, t_ld = 0, t_cvt = 0, t_drain = 0;
    // work split: with two halves each group drains its own half; with one half the groups take alternate panels
    // tile-parallel mode: the group pair gp = grp >> 1 owns every second tile (and TMEM set gp); inside the pair the two
    // groups split the halves of a 256-row tile, or alternate panels of a 128-row one
    const bool tp = p.tile_par != 0;
    const int gp = tp ? (grp >> 1) : 0;
    const int my_half = (halves == 2) ? (tp ? (grp & 1) : (grp >> 1)) : 0;
    const int pi_first = (halves == 2) ? (tp ? 0 : (grp & 1)) : (tp ? (grp & 1) : grp);
    const int pi_step = (halves == 2) ? (tp ? 1 : 2) : (tp ? 2 : 4);
    // rows of the staged panel this lane writes out (panel_drain): row group rg, rows rg*CH .. rg*CH + CH-1
    constexpr int drain_ch = CH;
    const int drain_row0 = quarter * 32 + (lane / drain_ch) * drain_ch;
    // loop-invariant shared-memory addresses of this lane (swizzle XOR folded in):
    //   staging writes: row `lane`, 16-byte chunk c -> st_base | ((c << 4) ^ st_x)
    //   drain reads   : rows rg*CH + {0,1} from dsm0 (+ pitch), rows rg*CH + {2,3} from dsm1 (+ pitch)
    const uint32_t st_base = my_panels + (uint32_t)lane * pitch;
    const uint32_t st_x = ((((uint32_t)lane * pitch) >> 7) & smask) << 4;
    uint32_t dsm0, dsm1;
    {
      const uint32_t rg = (uint32_t)lane / CH, chunk = (uint32_t)lane % CH;
      const uint32_t row0 = rg * CH;
      dsm0 = my_panels + row0 * pitch + ((chunk ^ (((row0 * pitch) >> 7) & smask)) << 4);
      dsm1 = my_panels + (row0 + 2) * pitch + ((chunk ^ ((((row0 + 2) * pitch) >> 7) & smask)) << 4);
    }
    // TMEM: this warp's lane quarter and accumulator half
    const uint32_t tm_h = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(my_half * ksplit) * acc_stride;
    const int lt_step = tp ? 2 : 1;
    lt = (uint32_t)gp;
    for (int m_blk = m_first + gp * m_step; m_blk < num_m_blocks; m_blk += m_step * lt_step, lt += lt_step) {
      const int h = my_half;
      const uint32_t set = (nsets == 2u) ? (lt & 1u) : 0u;
      const uint32_t use = (nsets == 2u) ? (lt >> 1) : lt;
      const int m0 = m_blk * p.block_m + h * kBlockM;
      // pixel index (in the output tensor's lattice) of the rows this lane stores; < 0: outside the tensor
      int pix[4];
      uint32_t rowoff[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int mm = m0 + drain_row0 + r;
        int px = -1;
        if (r < drain_ch && mm < p.M) {
          px = mm;
          if (scatter) {
            const int q = mm % p.Wq;
            const int t = mm / p.Wq;
            const int pp = t % p.Hp;
            const int img = t / p.Hp;
            px = (img * p.OH + pp * p.os + p.oph + (p.merge_n ? n_blk : 0)) * p.OW + q * p.os_w + p.opw;
          }
        }
        pix[r] = px;
        rowoff[r] = (px >= 0) ? (uint32_t)px * (uint32_t)p.ldo : 0xFFFFFFFFu;   // host: pixels * pitch < 2^32
      }
      if (EPI == kEpiBwd) {
        // the rows this lane will read in its drains (producer y / old destination): pull their lines into L2 while the
        // tile's MMAs are still running, so the loads below are L2 hits
        if ((lane % drain_ch) == 0) {
          for (int pi = pi_first; pi < npanels; pi += pi_step) {
            const int colp = n0 + pi * pw;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if (pix[r] >= 0) {
                if (add) prefetch_l2(p.out + (size_t)pix[r] * p.ldo + colp - (p.merge_n ? n_blk * p.block_n : 0));
                if (EPI == kEpiBwd) {
                  const int L = (p.bwd_split > 0 && colp >= p.bwd_split) ? 1 : 0;
                  prefetch_l2(p.bwd_y[L] + (size_t)pix[r] * p.bwd_ldy[L] + (colp - (L ? p.bwd_split : 0)));
                }
              }
            }
          }
        }
      }
      mbar_wait_t(tfull_bar(set), use & 1u, timed, w_tfull);
      tc_fence_after();
      if (pi_first >= npanels) {  // nothing to drain for this group: release the set straight away
        tc_fence_before();
        if (lane == 0) mbar_arrive(tempty_bar(set));
      }
      const uint32_t acc0 = tm_h + set * set_cols;

      const int m = m0 + row;
      for (int pi = pi_first; pi < npanels; pi += pi_step) {
        const uint32_t taddr = acc0 + pi * pw;
        long long tp_ld = 0, tp_cvt = 0;
        const int colp = n0 + pi * pw;            // first column of this panel
        __nv_bfloat16* gcol = p.out + colp - (p.merge_n ? n_blk * p.block_n : 0);
        BwdCols bw;
        bw.y = nullptr; bw.ldy = 0; bw.scale = bw.shift = nullptr; bw.relu = 0;
        uint4 early[4];
        if (EPI == kEpiBwd) {
          // the producer layer this panel's channels belong to (a panel never straddles bwd_split: host-checked)
          const int L = (p.bwd_split > 0 && colp >= p.bwd_split) ? 1 : 0;
          const int lc = colp - (L ? p.bwd_split : 0);
          bw.y = p.bwd_y[L] + lc;
          bw.ldy = p.bwd_ldy[L];
          bw.scale = p.bwd_scale[L] + lc;
          bw.shift = p.bwd_shift[L] + lc;
          bw.relu = p.bwd_relu[L];
          panel_early_load<CH, false, 2>(lane, gcol, p.ldo, pix, bw, early);
        }
        // the panel is drained in pieces of 16 columns to keep the register footprint small: this kernel runs with
        // ~no L1 (all of it is shared memory), so a spilled register costs an L2 round trip
        constexpr int npieces = PW / 16;
#pragma unroll
        for (int pc = 0; pc < npieces; ++pc) {
          const long long tq0 = timed ? clock64() : 0;
          uint32_t v[16];
          float f[16];
          const uint32_t ta = taddr + pc * 16;
          tmem_ld16(ta, v);
          if (ksplit >= 2) {   // sum the K-interleaved partial accumulators; loads are issued before the single wait
            uint32_t v2[16];
            tmem_ld16(ta + acc_stride, v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + __uint_as_float(v2[i]);
            for (int ksel = 2; ksel < ksplit; ++ksel) {
              tmem_ld16(ta + ksel * acc_stride, v2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += __uint_as_float(v2[i]);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          }
          if (pi + pi_step >= npanels && pc == npieces - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(set));
          }
          const long long tq1 = timed ? clock64() : 0;
          {
            const int ch = pc;   // 16-column chunk inside the panel
            const int col0 = n0 + pi * pw + ch * 16;
            if (EPI == kEpiAffine && p.scale != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], __ldg(p.scale + col0 + i), __ldg(p.shift + col0 + i));
            }
            if (EPI == kEpiAffine && p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (EPI == kEpiAffine && p.residual != nullptr && m < p.M) {
              const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + (size_t)m * p.ldr + col0);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint4 rv = __ldg(r4 + hh);
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // the reference adds two bf16 tensors: round first, then add
                  f[hh * 8 + 2 * i] = __bfloat162float(__float2bfloat16_rn(f[hh * 8 + 2 * i])) + bf16lo(rr[i]);
                  f[hh * 8 + 2 * i + 1] =
                      __bfloat162float(__float2bfloat16_rn(f[hh * 8 + 2 * i + 1])) + bf16hi(rr[i]);
                }
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t sa = st_base | ((uint32_t)((ch * 2 + hh) << 4) ^ st_x);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sa), "r"(pk[4 * hh]),
                           "r"(pk[4 * hh + 1]), "r"(pk[4 * hh + 2]), "r"(pk[4 * hh + 3])
                           : "memory");
            }
          }
          if (timed) {
            tp_ld += tq1 - tq0;
            tp_cvt += clock64() - tq1;
          }
        }
        const long long tp2 = timed ? clock64() : 0;
        {
          __syncwarp();
          float o0 = 0.f, o1 = 0.f;
          if (EPI == kEpiBwd) {
            if (add) panel_drain<CH, true, 2>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
            else panel_drain<CH, false, 2>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
          } else if (EPI == kEpiStats) {
            panel_drain<CH, false, 1>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
          } else if (EPI == kEpiStats1) {
            panel_drain<CH, false, 1, true>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
          } else if (EPI == kEpiPlain && add) {
            panel_drain<CH, true, 0>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
          } else {
            panel_drain<CH, false, 0>(dsm0, dsm1, lane, gcol, rowoff, pix, bw, early, o0, o1, accS, accQ);
          }
          if (do_stats && EPI != kEpiStats1) {
            const int lpi = (pi - pi_first) / pi_step;   // local panel index of this group, < 4
            if (lpi == 0) { a0.x += o0; a0.y += o1; }
            else if (lpi == 1) { a0.z += o0; a0.w += o1; }
            else if (lpi == 2) { a1.x += o0; a1.y += o1; }
            else { a1.z += o0; a1.w += o1; }
          }
          __syncwarp();   // the staging buffer is rewritten by the next panel
        }
        if (timed) {
          t_ld += tp_ld; t_cvt += tp_cvt; t_drain += clock64() - tp2;
        }
      }
    }
    if (do_stats) {
      // Combine the 4 (x2 halves) lane-quarter slices of the CTA in shared memory (fixed order), so that the CTA
      // contributes ONE statistics row; then, if the caller asked for it (p.tickets), the last CTA of this n-block
      // to finish turns the rows into mean / invstd / scale / shift and the running-statistics update (forward) or
      // into the BatchNorm-backward coefficients and parameter gradients of the producer layer (kEpiBwd) -
      // BatchNorm's finalisation costs no extra launch.
      if (EPI == kEpiStats1) {   // the one deferred butterfly: per-lane sums of all tiles -> this lane's column totals
        stats_butterfly<CH>(lane, accS, accQ);
        a0.x = accS[0];
        a0.y = accQ[0];
      }
      const int et = threadIdx.x - 128;             // 0..511 over the 16 epilogue warps
      const int nslices = halves * 4 * (tp ? 2 : 1);
      float2* slots = reinterpret_cast<float2*>(smem_raw + (panel_base - smem_u32(smem_raw)));   // [8][block_n]
      named_bar_sync(1, kEpiWarps * 32);            // every warp is done with its staging buffer
      const int q8 = (gp * halves + my_half) * 4 + quarter;
      for (int pi = pi_first; pi < npanels; pi += pi_step) {
        const int lpi = (pi - pi_first) / pi_step;
        const float4 o4 = (lpi < 2) ? a0 : a1;
        const float2 o = (lpi & 1) ? make_float2(o4.z, o4.w) : make_float2(o4.x, o4.y);
        // column owned by this lane after the butterfly (see panel_drain): row-group bit k selects half 4>>k
        int col = -1;
        if (pw == 32) {
          const int rg = lane >> 2;
          col = (lane & 3) * 8 + (rg & 1) * 4 + ((rg >> 1) & 1) * 2 + (rg >> 2);
        } else if ((lane >> 4) == 0) {
          const int rg = lane >> 1;
          col = (lane & 1) * 8 + (rg & 1) * 4 + ((rg >> 1) & 1) * 2 + ((rg >> 2) & 1);
        }
        if (col >= 0) slots[q8 * p.block_n + pi * pw + col] = o;
      }
      named_bar_sync(1, kEpiWarps * 32);
      float* rowp = p.stats_partial + (size_t)(p.stats_row0 + m_first) * p.cout * 2;
      if (et < p.block_n) {
        float sx = 0.f, sq = 0.f;
        for (int q = 0; q < nslices; ++q) {
          const float2 v = slots[q * p.block_n + et];
          sx += v.x;
          sq += v.y;
        }
        *reinterpret_cast<float2*>(rowp + (size_t)(n0 + et) * 2) = make_float2(sx, sq);
      }
      if (p.tickets != nullptr) {
        if (et < p.block_n) __threadfence();
        named_bar_sync(1, kEpiWarps * 32);
        uint32_t* flag = reinterpret_cast<uint32_t*>(slots);
        if (et == 0) {
          const unsigned int t = atomicAdd(p.tickets + n_blk, 1u);
          *flag = (t == (unsigned int)(m_step - 1)) ? 1u : 0u;
        }
        named_bar_sync(1, kEpiWarps * 32);
        const bool last = (*flag != 0u);
        named_bar_sync(1, kEpiWarps * 32);          // everyone has read the flag before the slots are reused
        if (last) {
          __threadfence();
          // G row groups x block_n columns; each thread sums rows g, g+G, .. of its column in double
          const int G = (kEpiWarps * 32) / p.block_n;
          const int g = et / p.block_n, col = et - g * p.block_n;
          const int nrows = p.fin_rows;
          double* dsl = reinterpret_cast<double*>(slots);   // [G][block_n][2]
          if (g < G) {
            double sx = 0.0, sq = 0.0;
            const float* src = p.stats_partial + (size_t)(n0 + col) * 2;
            const size_t rstride = (size_t)p.cout * 2;
            // 16 independent loads in flight per thread (a dependent add right behind each load would serialise them; this
            // loop is the serial tail of the launch: every other CTA has left, the next kernel waits for the coefficients)
            for (int r = g; r < nrows; r += 16 * G) {
              float2 v[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const int rr = r + u * G;
                v[u] = (rr < nrows) ? __ldcg(reinterpret_cast<const float2*>(src + (size_t)rr * rstride)) : make_float2(0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                sx += (double)v[u].x;
                sq += (double)v[u].y;
              }
            }
            dsl[(g * p.block_n + col) * 2] = sx;
            dsl[(g * p.block_n + col) * 2 + 1] = sq;
          }
          named_bar_sync(1, kEpiWarps * 32);
          double sx = 0.0, sq = 0.0;
          const int ch = n0 + et;
          if (et < p.block_n) {
            for (int gg = 0; gg < G; ++gg) {
              sx += dsl[(gg * p.block_n + et) * 2];
              sq += dsl[(gg * p.block_n + et) * 2 + 1];
            }
          }
          // kEpiBwd: the producer layer of this channel; its sums so far are S = sum dz, Q = sum dz*y
          const int BL = (EPI == kEpiBwd && p.bwd_split > 0 && ch >= p.bwd_split) ? 1 : 0;
          const int bc = ch - (BL ? p.bwd_split : 0);
          if (EPI == kEpiBwd && et < p.block_n) {
            sq = (sq - (double)p.bwd_mean[BL][bc] * sx) * (double)p.bwd_invstd[BL][bc];   // sum dz * xhat
            // parameter gradients come from the LOCAL sums (the gradient all-reduce averages them across ranks)
            if (p.bwd_dgamma[BL] != nullptr) p.bwd_dgamma[BL][bc] = (float)sq;
            if (p.bwd_dbeta[BL] != nullptr) p.bwd_dbeta[BL][bc] = (float)sx;
          }
          if (p.sync.world > 1) {
            // SyncBN: exchange this n-block's sums with all ranks over NVLink peer memory (syncbn.cuh)
            named_bar_sync(1, kEpiWarps * 32);          // dsl has been consumed; reuse its first word for the seq
            if (et == 0) *reinterpret_cast<volatile unsigned int*>(slots) = my_seq;
            named_bar_sync(1, kEpiWarps * 32);
            const unsigned int seq = *reinterpret_cast<volatile unsigned int*>(slots);
            if (et < p.block_n) sync_push(p.sync, seq, ch, sx, sq);
            if (!p.sync.tagged) {   // flag protocol: make the pushes visible, publish / await the per-rank flags
              __threadfence_system();
              named_bar_sync(1, kEpiWarps * 32);
              if (et < p.sync.world) sync_signal_wait(p.sync, kSyncFlagsConv, n_blk, seq, et);
              named_bar_sync(1, kEpiWarps * 32);
            }
            if (et < p.block_n) {
              const double2 v = sync_gather(p.sync, seq, ch);
              sx = v.x;
              sq = v.y;
            }
            if (et == 0) {   // the last n-block to finish its exchange completes this launch's sequence number
              const unsigned int done = atomicAdd(p.tickets + 64, 1u);
              if (done == (unsigned int)(n_blocks - 1)) {
                p.tickets[64] = 0u;
                sync_write_seq(p.sync, seq);
              }
            }
          }
          if (EPI == kEpiBwd) {
            if (et < p.block_n)
              *reinterpret_cast<float2*>(p.bwd_coef[BL] + (size_t)bc * 2) =
                  make_float2((float)(sx / p.bn_count), (float)(sq / p.bn_count));
          } else if (et < p.block_n) {
            const double mean = sx / p.bn_count;
            double var = sq / p.bn_count - mean * mean;
            if (var < 0) var = 0;
            const float invstd = (float)(1.0 / sqrt(var + (double)p.bn_eps));
            p.bn_mean[ch] = (float)mean;
            p.bn_invstd[ch] = invstd;
            // parameter tensors of the layer this channel belongs to (second layer of a side-by-side pair: bn_split)
            const bool second = p.bn_split > 0 && ch >= p.bn_split;
            const int pc = second ? ch - p.bn_split : ch;
            const float* gam = second ? p.bn_gamma2 : p.bn_gamma;
            const float* bet = second ? p.bn_beta2 : p.bn_beta;
            float* rmean = second ? p.bn_running_mean2 : p.bn_running_mean;
            float* rvar = second ? p.bn_running_var2 : p.bn_running_var;
            long long* nbt = second ? p.bn_nbt2 : p.bn_nbt;
            const float sc = gam[pc] * invstd;
            p.bn_scale[ch] = sc;
            p.bn_shift[ch] = bet[pc] - (float)mean * sc;
            if (rmean != nullptr) {
              const double unbiased = p.bn_count > 1 ? var * (p.bn_count / (p.bn_count - 1.0)) : var;
              rmean[pc] = (1.f - p.bn_momentum) * rmean[pc] + p.bn_momentum * (float)mean;
              rvar[pc] = (1.f - p.bn_momentum) * rvar[pc] + p.bn_momentum * (float)unbiased;
            }
            if (pc == 0 && nbt != nullptr) *nbt += 1;
          }
          if (et == 0) p.tickets[n_blk] = 0u;       // self-cleaning: ready for the next launch
        }
        if (EPI != kEpiBwd && p.fn_out != nullptr) {
          FnArgs fa;
          fa.y = p.out; fa.res = p.residual; fa.out = p.fn_out; fa.scale = p.bn_scale; fa.shift = p.bn_shift;
          fa.tickets = p.tickets; fa.ldy = p.ldo; fa.ldr = p.ldr; fa.ldo = p.fn_ldo; fa.relu = p.relu;
          fa.block_m = p.block_m; fa.block_n = p.block_n; fa.M = p.M; fa.nmb = num_m_blocks; fa.mstep = m_step;
          fused_normalise_tail(fa, et, n_blk, n0, m_first, last, reinterpret_cast<float2*>(slots));
        }
      }
    }
    if (timed && lane == 0) atomicMax(p.dbg + blockIdx.x * 16 + 14, (unsigned long long)(clock64() - t_start));   // slowest epilogue warp
    if (timed && quarter == 0 && lane == 0 && grp < 2) p.dbg[blockIdx.x * 16 + 5 + grp] = w_tfull;
    if (timed && warp == 4 && lane == 0) {
      p.dbg[blockIdx.x * 16 + 8] = t_ld;
      p.dbg[blockIdx.x * 16 + 9] = t_cvt;
      p.dbg[blockIdx.x * 16 + 10] = t_drain;
    }
  }
  if (timed && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 7] = clock64() - t_start;

  tc_fence_before();
  __syncthreads();
  if (timed && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[blockIdx.x * 16 + 12] = gt;   // every role of the CTA is done (ns)
    p.dbg[blockIdx.x * 16 + 13] = clock64() - t_start;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

#undef halves
#undef ksplit
#undef acc_stride
#undef set_cols
#undef nsets
#undef num_m_blocks
#undef n_blocks
#undef m_step
#undef kc
#undef chunks_per_tap
#undef total_chunks
#undef subs_per_stage
#undef num_k_stages
#undef a_sub_bytes
#undef a_half_bytes
#undef b_sub_bytes

void fill_derived(ConvIgemmParams& p, int grid) {
  const ConvSmem L = conv_smem_layout(p.block_m, p.block_n, p.num_stages, p.panel_bufs);
  p.halves = p.block_m / kBlockM;
  p.n_blocks = p.cout / p.block_n;
  p.num_m_blocks = (p.M + p.block_m - 1) / p.block_m;
  p.m_step = grid / p.n_blocks;
  p.chunks_per_tap = p.cin / p.kc;
  p.total_chunks = p.ntaps * p.chunks_per_tap;
  p.subs_per_stage = kStageK / p.kc;
  p.num_k_stages = (p.total_chunks + p.subs_per_stage - 1) / p.subs_per_stage;
  p.acc_stride = (uint32_t)((p.block_n + 31) & ~31);
  p.set_cols = (uint32_t)(p.halves * p.ksplit) * p.acc_stride;
  p.nsets = (2u * p.set_cols <= 512u) ? 2u : 1u;
  {
    static const bool tp_ok = !(getenv("VTB_TILE_PAR") && atoi(getenv("VTB_TILE_PAR")) == 0);
    // narrow tiles only: each epilogue warp then owns <= 2 panels per tile (its running statistics live in 8 registers)
    p.tile_par = (tp_ok && p.nsets == 2u && p.block_n <= 64 && p.block_n / p.panel_w <= (p.halves == 2 ? 2 : 4)) ? 1 : 0;
  }
  p.a_stage = L.a_stage;
  p.b_stage = L.b_stage;
  p.b_off = L.b_off;
  p.panel_off = L.panel_off;
  p.bar_off = L.bar_off;
  p.a_sub_bytes = (uint32_t)p.block_m * p.kc * 2;
  p.a_half_bytes = (uint32_t)kBlockM * p.kc * 2;
  p.b_sub_bytes = (uint32_t)p.block_n * p.kc * 2;
}

// ------------------------------------------------------------------------------------------------
// wgrad_igemm_kernel
//
// dW[Cout][tap*Cin] = sum over pixels dY[pix][Cout]^T * im2col(X)[pix][tap*Cin]: both operands MN-major (the GEMM-K
// axis = pixels is the strided one), split over the pixel axis, fp32 partials per split (deterministic).
// One CTA owns a 128 (Cout) x n_cols (<= 256 flattened tap*Cin columns) tile and ONE UMMA N = n_cols; successive
// 16-pixel K slices go round-robin to `ksplit` TMEM accumulators (independent dependency chains, summed in the epilogue).
// Operand feed: per stage `kpix` pixels = a_boxes dY boxes [kpix][ca] + b_boxes im2col X boxes [kpix][cc].  A single
// thread sustains only one TMA request per ~300 (tiled) / ~600 (im2col) cycles (profiles/r01_tma_bw.txt), so the requests
// of a stage are dealt round-robin to kWgProducers producer warps; every producer arrives on the stage's full barrier
// with the byte count of the boxes it issued.
//   warps 0..NP-1 : TMA producers     warp NP : TMEM owner + MMA issuer     warps NP+1..NP+4 : epilogue
// ------------------------------------------------------------------------------------------------
struct WgradSmem {
  uint32_t a_stage, b_stage, a_off, b_off, bar_off, total;
};
static __host__ __device__ inline WgradSmem wgrad_smem_layout(int ma, int n_cols, int kpix, int num_stages) {
  WgradSmem s;
  s.a_stage = round_up_int(kpix * ma * 2, 1024);
  s.b_stage = round_up_int(kpix * n_cols * 2, 1024);
  s.a_off = 0;
  s.b_off = num_stages * s.a_stage;
  s.bar_off = s.b_off + num_stages * s.b_stage;
  s.total = s.bar_off + 256;
  return s;
}
size_t wgrad_igemm_smem_bytes(int ma, int n_cols, int kpix, int num_stages) {
  return wgrad_smem_layout(ma, n_cols, kpix, num_stages).total + 1024;
}

template <bool TIMED>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ WgradIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.num_stages;
  const int kpix = p.kpix;
  const WgradSmem L = wgrad_smem_layout(p.ma, p.n_cols, kpix, S);
  const uint32_t a_base = base + L.a_off;
  const uint32_t b_base = base + L.b_off;
  const uint32_t bar_base = base + L.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool timed = TIMED;   // development counters (WgradIgemmParams::dbg), compiled out of the production kernel
  const long long t_start = timed ? clock64() : 0;
  unsigned long long* dbg = timed ? p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;

  // tile decode: blockIdx.x = m_tile * n_tiles + n_tile, blockIdx.y = split
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int split = blockIdx.y;
  const int box0 = n_tile * p.boxes_per_tile;                       // first (tap, c0) box of this tile
  const int b_boxes = min(p.boxes_per_tile, p.total_boxes - box0);  // real X boxes of this tile
  const int kb_per = (p.kblocks + p.splits - 1) / p.splits;
  const int kb0 = split * kb_per;
  const int kb1 = min(p.kblocks, kb0 + kb_per);
  const int nkb = max(0, kb1 - kb0);
  const int co0 = m_tile * 128;
  const int a_boxes = min(128, p.cout - co0) / p.ca;                // real dY boxes along Cout
  const uint32_t a_box_bytes = kpix * p.ca * 2;
  const uint32_t b_box_bytes = kpix * p.cc * 2;
  const int reqs = a_boxes + b_boxes;                               // TMA requests per stage

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), kWgProducers);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == kWgProducers) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();
  pdl_trigger();

  if (warp < kWgProducers) {
    // ======================= TMA producers =======================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      uint32_t g = 0;  // running request counter at the start of the stage (same value in every producer)
      long long waited = 0;
      for (int kb = kb0; kb < kb1; ++kb, g += reqs) {
        mbar_wait_t(empty_bar(stage), phase ^ 1u, timed, waited);
        // requests r of this stage with (g + r) % NP == warp are mine
        int r = (int)((warp + kWgProducers - (g % kWgProducers)) % kWgProducers);
        int mine_a = 0, mine_b = 0;
        for (int q = r; q < reqs; q += kWgProducers) (q < a_boxes) ? ++mine_a : ++mine_b;
        const uint32_t bytes = mine_a * a_box_bytes + mine_b * b_box_bytes;
        if (bytes) mbar_expect_tx(full_bar(stage), bytes);
        else mbar_arrive(full_bar(stage));
        if (bytes) {
          const int pix0 = kb * kpix;
          const int q0 = pix0 % p.Wq;
          const int t = pix0 / p.Wq;
          const int p0 = t % p.Hp;
          const int img = t / p.Hp;
          const int bw = p.lower_w + q0 * p.stride;
          const int bh = p.lower_h + p0 * p.stride;
          const uint32_t a_st = a_base + stage * L.a_stage;
          const uint32_t b_st = b_base + stage * L.b_stage;
          for (; r < reqs; r += kWgProducers) {
            if (r < a_boxes) {
              tma_load_2d(a_st + r * a_box_bytes, &tmDY, full_bar(stage), co0 + r * p.ca, pix0);
            } else {
              const int b = r - a_boxes;
              const int box = box0 + b;
              const int tap = box / p.boxes_per_tap;
              const int c0 = (box - tap * p.boxes_per_tap) * p.cc;
              tma_load_im2col_4d(b_st + b * b_box_bytes, &tmX, full_bar(stage), c0, bw, bh, img, p.tap_ow[tap],
                                 p.tap_oh[tap]);
            }
          }
        }
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
      }
      if (timed && warp == 0) dbg[0] = waited;
    }
  } else if (warp == kWgProducers) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t n_mma = (uint32_t)b_boxes * p.cc;   // UMMA N of this tile (multiple of 16, <= 256)
      const uint32_t idesc = make_idesc_bf16(128, n_mma, 1, 1);
      const uint32_t lta = (p.ca == 64) ? 2u : (p.ca == 32 ? 4u : 6u);
      const uint32_t ltb = (p.cc == 64) ? 2u : (p.cc == 32 ? 4u : 6u);
      const uint32_t sbo_a = 8u * p.ca * 2u, sbo_b = 8u * p.cc * 2u;
      // MN-major operands: LBO = distance between the channel boxes, SBO = distance between 8-pixel groups
      const uint64_t adesc_hi = make_smem_desc(0, a_box_bytes, sbo_a, lta);
      const uint64_t bdesc_hi = make_smem_desc(0, b_box_bytes, sbo_b, ltb);
      const uint32_t a_k16 = (2u * sbo_a) >> 4, b_k16 = (2u * sbo_b) >> 4;
      const uint32_t ks_mask = (uint32_t)p.ksplit - 1u;
      const int k16s = kpix / 16;
      uint32_t stage = 0, phase = 0, kk = 0;
      long long w_full = 0;
      if (timed) dbg[6] = clock64() - t_start;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_t(full_bar(stage), phase, timed, w_full);
        if (timed && kb == kb0) dbg[5] = w_full;
        tc_fence_after();
        uint32_t a16 = (a_base + stage * L.a_stage) >> 4;
        uint32_t b16 = (b_base + stage * L.b_stage) >> 4;
        for (int k16 = 0; k16 < k16s; ++k16, ++kk, a16 += a_k16, b16 += b_k16)
          umma_bf16(tmem_base + (kk & ks_mask) * p.acc_stride, adesc_hi | (uint64_t)a16, bdesc_hi | (uint64_t)b16, idesc,
                    (kk > ks_mask) ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar);
      if (timed) dbg[1] = w_full;
    }
  }
  if (warp != kWgProducers) {
    // ======================= epilogue: TMEM -> fp32 partial tile of this split =======================
    // Every warp but the MMA issuer drains: a warp may touch the TMEM lane quarter (warp % 4), so the 8 producer warps
    // (idle once their last load is issued) and the 4 epilogue warps form 3 parts per quarter that interleave the
    // 16-column pieces.
    __syncwarp();
    const int quarter = warp & 3;
    const int part = (warp < kWgProducers) ? (warp >> 2) : 2;
    constexpr int nparts = kWgProducers / 4 + 1;
    const bool dbg_warp = (warp == kWgProducers + 4) && lane == 0;
    const int row = quarter * 32 + lane;
    const int co = co0 + row;
    const int ktot = p.ntaps * p.cin;
    const int n_mma = b_boxes * p.cc;
    const int ks_used = min(p.ksplit, nkb * (kpix / 16));   // accumulators that received at least one MMA
    long long w_tf = 0;
    if (nkb > 0) {
      mbar_wait_t(tfull_bar, 0, timed, w_tf);
      tc_fence_after();
    }
    const long long t_epi = timed ? clock64() : 0;
    if (timed && dbg_warp) dbg[3] = w_tf;
    if (co0 + quarter * 32 < p.cout) {   // warp-uniform: skip lane quarters that hold no real output channel
      float* wrow = p.ws + ((size_t)split * p.cout + min(co, p.cout - 1)) * ktot + (size_t)box0 * p.cc;
      for (int ch = part; ch < n_mma / 16; ch += nparts) {
        uint32_t v[16];
        if (nkb > 0) {
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + ch * 16;
          tmem_ld16(ta, v);
          if (ks_used >= 2) {   // both loads in flight before the wait
            uint32_t v2[16];
            tmem_ld16(ta + p.acc_stride, v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
            for (int s = 2; s < ks_used; ++s) {
              tmem_ld16(ta + s * p.acc_stride, v2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
            }
          } else {
            tmem_ld_wait();
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (co < p.cout) {
          float4* dst = reinterpret_cast<float4*>(wrow + ch * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                 __uint_as_float(v[4 * i + 3]));
        }
      }
    }
    if (timed && dbg_warp) dbg[4] = clock64() - t_epi;
  }
  tc_fence_before();
  __syncthreads();
  if (timed && threadIdx.x == 0) dbg[2] = clock64() - t_start;
  if (warp == kWgProducers) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <bool TIMED, int EPI, int PW>
static int launch_conv_variant_pw(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                               const ConvIgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<TIMED, EPI, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  return (int)launch_pdl(conv_igemm_kernel<TIMED, EPI, PW>, dim3(grid), dim3(kConvThreads), smem, stream, tmA, tmB, tmD, p);
}
template <bool TIMED, int EPI>
static int launch_conv_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                               const ConvIgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (p.panel_w == 32) return launch_conv_variant_pw<TIMED, EPI, 32>(tmA, tmB, tmD, p, grid, smem, stream);
  return launch_conv_variant_pw<TIMED, EPI, 16>(tmA, tmB, tmD, p, grid, smem, stream);
}

int launch_conv_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                      const ConvIgemmParams& p_in, int grid, cudaStream_t stream) {
  ConvIgemmParams p = p_in;
  fill_derived(p, grid);
  const size_t smem = conv_igemm_smem_bytes(p.block_m, p.block_n, p.num_stages, p.panel_bufs);
  if (p.fin_rows <= 0) p.fin_rows = p.m_step;
  // one local panel per epilogue warp; 1x1 layers only (measured: the 3x3 32->32 layer loses 6 % to the extra registers)
  const int local_panels = p.tile_par ? ((p.block_n / p.panel_w + (p.halves == 2 ? 0 : 1)) / (p.halves == 2 ? 1 : 2))
                                      : ((p.block_n / p.panel_w + (p.halves == 2 ? 1 : 3)) / (p.halves == 2 ? 2 : 4));
  const bool one_panel = p.a_tiled && local_panels <= 1;
  static const bool defer_ok = !(getenv("VTB_STATS_DEFER") && atoi(getenv("VTB_STATS_DEFER")) == 0);
  const int epi = p.bwd_y[0] != nullptr ? kEpiBwd
                  : (p.stats_partial != nullptr ? ((one_panel && defer_ok) ? kEpiStats1 : kEpiStats)
                                                : ((p.scale != nullptr || p.relu || p.residual != nullptr) ? kEpiAffine : kEpiPlain));
  if (p.dbg != nullptr) {
    switch (epi) {
      case kEpiStats: return launch_conv_variant<true, kEpiStats>(tmA, tmB, tmD, p, grid, smem, stream);
      case kEpiStats1: return launch_conv_variant<true, kEpiStats1>(tmA, tmB, tmD, p, grid, smem, stream);
      case kEpiAffine: return launch_conv_variant<true, kEpiAffine>(tmA, tmB, tmD, p, grid, smem, stream);
      case kEpiPlain: return launch_conv_variant<true, kEpiPlain>(tmA, tmB, tmD, p, grid, smem, stream);
      default: return launch_conv_variant<true, kEpiBwd>(tmA, tmB, tmD, p, grid, smem, stream);
    }
  }
  switch (epi) {
    case kEpiStats: return launch_conv_variant<false, kEpiStats>(tmA, tmB, tmD, p, grid, smem, stream);
    case kEpiStats1: return launch_conv_variant<false, kEpiStats1>(tmA, tmB, tmD, p, grid, smem, stream);
    case kEpiAffine: return launch_conv_variant<false, kEpiAffine>(tmA, tmB, tmD, p, grid, smem, stream);
    case kEpiPlain: return launch_conv_variant<false, kEpiPlain>(tmA, tmB, tmD, p, grid, smem, stream);
    default: return launch_conv_variant<false, kEpiBwd>(tmA, tmB, tmD, p, grid, smem, stream);
  }
}

int launch_wgrad_igemm(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradIgemmParams& p, int grid_x,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const size_t smem = wgrad_igemm_smem_bytes(p.ma, p.n_cols, p.kpix, p.num_stages);
  if (p.dbg != nullptr)
    return (int)launch_pdl(wgrad_igemm_kernel<true>, dim3(grid_x, p.splits), dim3(kWgThreads), smem, stream, tmDY, tmX, p);
  return (int)launch_pdl(wgrad_igemm_kernel<false>, dim3(grid_x, p.splits), dim3(kWgThreads), smem, stream, tmDY, tmX, p);
}

}  // namespace vtb
