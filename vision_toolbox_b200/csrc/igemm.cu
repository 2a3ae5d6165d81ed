// tcgen05 implicit-GEMM convolution kernels (see igemm.cuh for the math each one computes).
//
// Both kernels are warp-specialised: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer
// (one lane), warps 2..5 = epilogue (TMEM -> registers -> smem/global).  Operands are staged in shared
// memory by TMA (im2col mode for the activation operand, so a tile of 128 consecutive output pixels may
// cross row and image boundaries and padding is zero-filled by hardware), accumulators live in TMEM.
#include "igemm.cuh"
#include "ptx.cuh"

namespace vtb {

static __host__ __device__ inline int round_up_int(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (host + device agree through these helpers)
// ------------------------------------------------------------------------------------------------
struct ConvSmem {
  uint32_t a_off, b_off, panel_off, bar_off, total;
  uint32_t b_stage;
};
static __host__ __device__ inline ConvSmem conv_smem_layout(int block_n, int num_stages) {
  ConvSmem s;
  s.b_stage = round_up_int(block_n * kStageK * 2, 1024);
  s.a_off = 0;
  s.b_off = num_stages * (kBlockM * kStageK * 2);
  s.panel_off = s.b_off + num_stages * s.b_stage;
  s.bar_off = s.panel_off + 2 * (kBlockM * 128);
  s.total = s.bar_off + 256;
  return s;
}
size_t conv_igemm_smem_bytes(int block_n, int num_stages) {
  return conv_smem_layout(block_n, num_stages).total + 1024;  // + slack for manual 1024B alignment
}

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }

// ------------------------------------------------------------------------------------------------
// conv_igemm_kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmD, const __grid_constant__ ConvIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const ConvSmem L = conv_smem_layout(p.block_n, p.num_stages);
  const int S = p.num_stages;

  const uint32_t a_base = base + L.a_off;
  const uint32_t b_base = base + L.b_off;
  const uint32_t panel_base = base + L.panel_off;
  const uint32_t bar_base = base + L.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to aligned base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_blocks = (p.M + kBlockM - 1) / kBlockM;
  const int n_blocks = p.cout / p.block_n;
  const int num_tiles = num_m_blocks * n_blocks;
  const int kc = p.kc;
  const int chunks_per_tap = p.cin / kc;
  const int total_chunks = p.ntaps * chunks_per_tap;
  const int subs_per_stage = kStageK / kc;
  const int num_k_stages = (total_chunks + subs_per_stage - 1) / subs_per_stage;
  const uint32_t a_sub_bytes = kBlockM * kc * 2;
  const uint32_t b_sub_bytes = p.block_n * kc * 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks;
        const int n_blk = tile - m_blk * n_blocks;
        const int m0 = m_blk * kBlockM;
        const int q0 = m0 % p.Wq;
        const int t = m0 / p.Wq;
        const int p0 = t % p.Hp;
        const int img = t / p.Hp;
        const int bw = p.lower_w + q0 * p.stride;
        const int bh = p.lower_h + p0 * p.stride;
        for (int ks = 0; ks < num_k_stages; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const int j0 = ks * subs_per_stage;
          const int nsub = min(subs_per_stage, total_chunks - j0);
          mbar_expect_tx(full_bar(stage), nsub * (a_sub_bytes + b_sub_bytes));
          const uint32_t a_st = a_base + stage * (kBlockM * kStageK * 2);
          const uint32_t b_st = b_base + stage * L.b_stage;
          for (int sub = 0; sub < nsub; ++sub) {
            const int j = j0 + sub;
            const int tap = j / chunks_per_tap;
            const int c0 = (j - tap * chunks_per_tap) * kc;
            tma_load_im2col_4d(a_st + sub * a_sub_bytes, &tmA, full_bar(stage), c0, bw, bh, img, p.tap_ow[tap],
                               p.tap_oh[tap]);
            tma_load_2d(b_st + sub * b_sub_bytes, &tmB, full_bar(stage), p.tap_kofs[tap] + c0, n_blk * p.block_n);
          }
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, 0);
      const uint32_t ltype = (kc == 64) ? 2u : (kc == 32 ? 4u : 6u);
      const uint32_t sbo = 8u * kc * 2u;
      uint32_t stage = 0, phase = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const uint32_t use = (uint32_t)(lt >> 1);
        mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        for (int ks = 0; ks < num_k_stages; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const int j0 = ks * subs_per_stage;
          const int nsub = min(subs_per_stage, total_chunks - j0);
          const uint32_t a_st = a_base + stage * (kBlockM * kStageK * 2);
          const uint32_t b_st = b_base + stage * L.b_stage;
          for (int sub = 0; sub < nsub; ++sub) {
            for (int k16 = 0; k16 < kc / 16; ++k16) {
              const uint64_t adesc = make_smem_desc(a_st + sub * a_sub_bytes + k16 * 32, 16, sbo, ltype);
              const uint64_t bdesc = make_smem_desc(b_st + sub * b_sub_bytes + k16 * 32, 16, sbo, ltype);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ======================= epilogue (4 warps) =======================
    const int et = threadIdx.x - 64;            // 0..127
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;        // tile row == TMEM lane
    const int pw = p.panel_w;
    const int nchunk = pw / 16;
    const uint32_t pitch = pw * 2;
    const uint32_t smask = (pw == 64) ? 7u : (pw == 32 ? 3u : 1u);
    const bool tma_mode = (p.store_mode == kStoreTma || p.store_mode == kStoreTmaAdd);
    const bool do_stats = (p.stats_partial != nullptr) && tma_mode;
    const int pairs = pw / 2;
    const int groups = kEpiThreads / pairs;
    const int rows_per_group = kBlockM / groups;
    const int my_pair = et % pairs;
    const int my_grp = et / pairs;
    uint32_t pcount = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const uint32_t use = (uint32_t)(lt >> 1);
      const int m_blk = tile / n_blocks;
      const int n_blk = tile - m_blk * n_blocks;
      const int m0 = m_blk * kBlockM;
      const int n0 = n_blk * p.block_n;
      mbar_wait(tfull_bar(acc), use & 1u);
      tc_fence_after();

      // scatter-mode addressing for this thread's pixel
      __nv_bfloat16* out_row = nullptr;
      const int m = m0 + row;
      if (!tma_mode && m < p.M) {
        const int q = m % p.Wq;
        const int t = m / p.Wq;
        const int pp = t % p.Hp;
        const int img = t / p.Hp;
        out_row = p.out + ((size_t)((size_t)img * p.OH + (size_t)pp * p.os + p.oph) * p.OW + (size_t)q * p.os + p.opw) *
                              (size_t)p.ldo;
      }
      const int npanels = p.block_n / pw;
      for (int pi = 0; pi < npanels; ++pi) {
        const uint32_t buf = pcount & 1u;
        ++pcount;
        const uint32_t panel = panel_base + buf * (kBlockM * 128);
        if (tma_mode) {
          if (et == 0) tma_store_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
        }
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + pi * pw;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          if (ch < nchunk) tmem_ld16(taddr + ch * 16, v + ch * 16);
        tmem_ld_wait();
        if (pi == npanels - 1) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (ch < nchunk) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[ch * 16 + i]);
            const int col0 = n0 + pi * pw + ch * 16;
            if (p.scale != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], __ldg(p.scale + col0 + i), __ldg(p.shift + col0 + i));
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.residual != nullptr && m < p.M) {
              const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + (size_t)m * p.ldr + col0);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 rv = __ldg(r4 + h);
                const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // the reference adds two bf16 tensors: round first, then add
                  f[h * 8 + 2 * i] = __bfloat162float(__float2bfloat16_rn(f[h * 8 + 2 * i])) + bf16lo(rr[i]);
                  f[h * 8 + 2 * i + 1] = __bfloat162float(__float2bfloat16_rn(f[h * 8 + 2 * i + 1])) + bf16hi(rr[i]);
                }
              }
            }
            if (tma_mode) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint32_t off = swz(row * pitch + (ch * 2 + h) * 16, smask);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(panel + off), "r"(pk[4 * h]),
                             "r"(pk[4 * h + 1]), "r"(pk[4 * h + 2]), "r"(pk[4 * h + 3])
                             : "memory");
              }
            } else if (out_row != nullptr) {
              uint4* o4 = reinterpret_cast<uint4*>(out_row + col0);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (p.store_mode == kStoreScatterAdd) {
                  const uint4 ov = o4[h];
                  const uint32_t oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    f[h * 8 + 2 * i] = __bfloat162float(__float2bfloat16_rn(f[h * 8 + 2 * i])) + bf16lo(oo[i]);
                    f[h * 8 + 2 * i + 1] = __bfloat162float(__float2bfloat16_rn(f[h * 8 + 2 * i + 1])) + bf16hi(oo[i]);
                  }
                }
                uint4 sv;
                sv.x = pack_bf16x2(f[h * 8 + 0], f[h * 8 + 1]);
                sv.y = pack_bf16x2(f[h * 8 + 2], f[h * 8 + 3]);
                sv.z = pack_bf16x2(f[h * 8 + 4], f[h * 8 + 5]);
                sv.w = pack_bf16x2(f[h * 8 + 6], f[h * 8 + 7]);
                o4[h] = sv;
              }
            }
          }
        }
        if (tma_mode) {
          fence_proxy_async_smem();
          named_bar_sync(2, kEpiThreads);
          if (et == 0) {
            if (p.store_mode == kStoreTma)
              tma_store_2d(&tmD, panel, n0 + pi * pw, m0);
            else
              tma_reduce_add_2d(&tmD, panel, n0 + pi * pw, m0);
            tma_store_commit();
          }
          if (do_stats) {
            // column sums over this thread's row group, read back from the staged bf16 panel
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            const int r0 = my_grp * rows_per_group;
            const uint32_t cofs = (my_pair >> 2) * 16 + (my_pair & 3) * 4;
#pragma unroll 8
            for (int r = 0; r < rows_per_group; ++r) {
              uint32_t w;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(panel + swz((r0 + r) * pitch + cofs, smask)));
              const float a = bf16lo(w), b = bf16hi(w);
              s0 += a;
              q0 = fmaf(a, a, q0);
              s1 += b;
              q1 = fmaf(b, b, q1);
            }
            float4* dst = reinterpret_cast<float4*>(
                p.stats_partial + (((size_t)blockIdx.x * groups + my_grp) * p.cout + n0 + pi * pw + 2 * my_pair) * 2);
            float4 cur = *dst;
            cur.x += s0;
            cur.y += q0;
            cur.z += s1;
            cur.w += q1;
            *dst = cur;
          }
        }
      }
    }
    if (tma_mode && et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad_igemm_kernel
// ------------------------------------------------------------------------------------------------
struct WgradSmem {
  uint32_t a_stage, b_stage, a_off, b_off, bar_off, total;
};
static __host__ __device__ inline WgradSmem wgrad_smem_layout(int subs_per_tile, int sub_n, int kpix, int num_stages) {
  WgradSmem s;
  s.a_stage = kpix * 128 * 2;
  s.b_stage = round_up_int(subs_per_tile * sub_n * kpix * 2, 1024);
  s.a_off = 0;
  s.b_off = num_stages * s.a_stage;
  s.bar_off = s.b_off + num_stages * s.b_stage;
  s.total = s.bar_off + 256;
  return s;
}
size_t wgrad_igemm_smem_bytes(int subs_per_tile, int sub_n, int num_stages) {
  return wgrad_smem_layout(subs_per_tile, sub_n, kStageK, num_stages).total + 1024;
}

__global__ void __launch_bounds__(kNumThreads, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ WgradIgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int kpix = kStageK;
  const int S = p.num_stages;
  const WgradSmem L = wgrad_smem_layout(p.subs_per_tile, p.sub_n, kpix, S);
  const uint32_t a_base = base + L.a_off;
  const uint32_t b_base = base + L.b_off;
  const uint32_t bar_base = base + L.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * S);
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 1);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode: blockIdx.x = (m_tile * n_tiles + n_tile), blockIdx.y = split
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int split = blockIdx.y;
  const int sub0 = n_tile * p.subs_per_tile;
  const int nsubs = min(p.subs_per_tile, p.total_subs - sub0);
  const int kb_per = (p.kblocks + p.splits - 1) / p.splits;
  const int kb0 = split * kb_per;
  const int kb1 = min(p.kblocks, kb0 + kb_per);
  const int nkb = max(0, kb1 - kb0);
  const int subs_per_tap = p.cin / p.sub_n;
  const int co0 = m_tile * 128;
  const int a_boxes = min(128, p.cout - co0) / p.ca;     // real dY boxes along Cout
  const int b_boxes = p.sub_n / p.cc;                     // boxes per sub-tile along Cin
  const uint32_t a_box_bytes = kpix * p.ca * 2;
  const uint32_t b_box_bytes = kpix * p.cc * 2;
  const uint32_t b_sub_bytes = kpix * p.sub_n * 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmDY);
      tma_prefetch_desc(&tmX);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const int pix0 = kb * kpix;
        const int q0 = pix0 % p.Wq;
        const int t = pix0 / p.Wq;
        const int p0 = t % p.Hp;
        const int img = t / p.Hp;
        const int bw = p.lower_w + q0 * p.stride;
        const int bh = p.lower_h + p0 * p.stride;
        mbar_expect_tx(full_bar(stage), a_boxes * a_box_bytes + nsubs * b_sub_bytes);
        const uint32_t a_st = a_base + stage * L.a_stage;
        const uint32_t b_st = b_base + stage * L.b_stage;
        for (int b = 0; b < a_boxes; ++b)
          tma_load_2d(a_st + b * a_box_bytes, &tmDY, full_bar(stage), co0 + b * p.ca, pix0);
        for (int s = 0; s < nsubs; ++s) {
          const int sub = sub0 + s;
          const int tap = sub / subs_per_tap;
          const int c0 = (sub - tap * subs_per_tap) * p.sub_n;
          for (int b = 0; b < b_boxes; ++b)
            tma_load_im2col_4d(b_st + s * b_sub_bytes + b * b_box_bytes, &tmX, full_bar(stage), c0 + b * p.cc, bw, bh,
                               img, p.tap_ow[tap], p.tap_oh[tap]);
        }
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.sub_n, 1, 1);
      const uint32_t lta = (p.ca == 64) ? 2u : (p.ca == 32 ? 4u : 6u);
      const uint32_t ltb = (p.cc == 64) ? 2u : (p.cc == 32 ? 4u : 6u);
      const uint32_t sbo_a = 8u * p.ca * 2u, sbo_b = 8u * p.cc * 2u;
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_st = a_base + stage * L.a_stage;
        const uint32_t b_st = b_base + stage * L.b_stage;
        for (int s = 0; s < nsubs; ++s) {
          for (int k16 = 0; k16 < kpix / 16; ++k16) {
            const uint64_t adesc = make_smem_desc(a_st + k16 * 2 * sbo_a, a_box_bytes, sbo_a, lta);
            const uint64_t bdesc = make_smem_desc(b_st + s * b_sub_bytes + k16 * 2 * sbo_b, b_box_bytes, sbo_b, ltb);
            umma_bf16(tmem_base + s * p.sub_n, adesc, bdesc, idesc, (kb > kb0 || k16 > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(stage));
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int co = co0 + row;
    const int ktot = p.ntaps * p.cin;
    if (nkb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    float* wrow = p.ws + ((size_t)split * p.cout + co) * ktot;
    for (int s = 0; s < nsubs; ++s) {
      const int sub = sub0 + s;
      const int tap = sub / subs_per_tap;
      const int c0 = (sub - tap * subs_per_tap) * p.sub_n;
      for (int ch = 0; ch < p.sub_n / 16; ++ch) {
        uint32_t v[16];
        if (nkb > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + s * p.sub_n + ch * 16, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (co < p.cout) {
          float4* dst = reinterpret_cast<float4*>(wrow + tap * p.cin + c0 + ch * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                 __uint_as_float(v[4 * i + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_conv_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                      const ConvIgemmParams& p, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const size_t smem = conv_igemm_smem_bytes(p.block_n, p.num_stages);
  conv_igemm_kernel<<<grid, kNumThreads, smem, stream>>>(tmA, tmB, tmD, p);
  return (int)cudaGetLastError();
}

int launch_wgrad_igemm(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradIgemmParams& p, int grid_x,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const size_t smem = wgrad_igemm_smem_bytes(p.subs_per_tile, p.sub_n, p.num_stages);
  wgrad_igemm_kernel<<<dim3(grid_x, p.splits), kNumThreads, smem, stream>>>(tmDY, tmX, p);
  return (int)cudaGetLastError();
}

}  // namespace vtb
