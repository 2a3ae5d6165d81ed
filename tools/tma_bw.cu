// Microbenchmark: how fast can one SM pull operand tiles through TMA on B200?
// Persistent CTAs, P producer threads each cycling over S smem stages; a consumer warp only waits on the
// "full" barrier and frees the slot. Reports bytes/cycle/SM and chip GB/s for tiled vs im2col boxes of
// 128 rows x {32,64,128} bytes (the operand shapes of csrc/igemm.cu), 1 or 2 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bw tools/tma_bw.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vision_toolbox_b200/csrc/ptx.cuh"
#include "../vision_toolbox_b200/csrc/tmap.cuh"

using namespace vtb;

struct Params {
  int mode;        // 0 tiled 2D, 1 im2col 4D
  int row_bytes;   // 32/64/128
  int rows;        // rows per box (128 or 256)
  int stages;
  int producers;   // producer threads (each owns stages/producers slots)
  int iters;       // loads per producer
  int total_rows;  // rows of the source matrix (tiled) / pixels
  int W, H;        // im2col image dims
  int taps;        // im2col: cycle through taps*taps filter offsets
  long long* cycles;
};

__global__ void __launch_bounds__(256, 1)
tma_bw_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.rows * p.row_bytes;
  const uint32_t stage_stride = (stage_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = base + p.stages * stage_stride;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = p.stages / p.producers;
  const long long t0 = clock64();
  if (warp < p.producers) {
    if (elect_one()) {
      const int s0 = warp * per;
      uint32_t phase = 0;
      int st = 0;
      // spread CTAs over the source so the working set streams from L2/HBM
      long long row = ((long long)blockIdx.x * 977 + warp * 131) * p.rows % (p.total_rows - p.rows);
      for (int it = 0; it < p.iters; ++it) {
        const int s = s0 + st;
        mbar_wait(empty_bar(s), phase ^ 1u);
        mbar_expect_tx(full_bar(s), stage_bytes);
        if (p.mode == 0) {
          tma_load_2d(base + s * stage_stride, &tm, full_bar(s), 0, (int)row);
        } else {
          const int pix = (int)row;
          const int q = pix % p.W, t = pix / p.W, hh = t % p.H, img = t / p.H;
          const int tap = it % (p.taps * p.taps);
          const int pad = p.taps / 2;
          tma_load_im2col_4d(base + s * stage_stride, &tm, full_bar(s), 0, q - pad, hh - pad, img,
                             (uint16_t)(tap % p.taps), (uint16_t)(tap / p.taps));
        }
        row += (long long)gridDim.x * p.rows;
        if (row >= p.total_rows - p.rows) row -= (p.total_rows - p.rows);
        if (++st == per) { st = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 4 + p.producers) {
    if (lane == 0) {
      const int s0 = (warp - 4) * per;
      uint32_t phase = 0;
      int st = 0;
      for (int it = 0; it < p.iters; ++it) {
        const int s = s0 + st;
        mbar_wait(full_bar(s), phase);
        mbar_arrive(empty_bar(s));
        if (++st == per) { st = 0; phase ^= 1u; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  // source: 512 MB of bf16 viewed either as [rows][C] or NHWC image batch
  const size_t bytes = 1024ull << 20;
  void* src;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%-8s %5s %5s %3s %3s %4s | %9s %9s %8s\n", "mode", "rowB", "rows", "S", "P", "cta", "B/cyc/SM", "GB/s", "cyc/row");
  for (size_t ws_mb : {32})
  for (int mode = 0; mode < 2; ++mode)
    for (int rb : {32, 64, 128})
      for (int rows : {128, 256, 512, 1024})
        for (int ctas_per_sm : {1})
          for (int producers : {1, 2, 3}) {
            if (mode == 0 && rows > 256) continue;
            if (rows > 256 && rb == 128 && rows * rb * 3 > 200000) continue;
            Params p{};
            p.mode = mode;
            p.row_bytes = rb;
            p.rows = rows;
            const int C = rb / 2;
            p.stages = 6;
            while (p.stages * (((rows * rb) + 1023) & ~1023) > 200000) --p.stages;
            p.stages = (p.stages / producers) * producers;
            if (p.stages < producers) continue;
            p.producers = producers;
            p.iters = 4000 / producers;
            p.W = 64; p.H = 64; p.taps = 3;
            const long long total_pix = (long long)((ws_mb << 20) / rb);
            const int nimg = (int)(total_pix / (p.W * p.H));
            p.total_rows = nimg * p.W * p.H;
            p.cycles = cyc;
            CUtensorMap tm;
            bool ok;
            if (mode == 0)
              ok = tmap_tiled_2d(&tm, src, C, p.total_rows, rb, C, rows, rb);
            else
              ok = tmap_im2col_nhwc(&tm, src, C, p.W, p.H, nimg, C, -1, -1, -1, -1, C, rows, 1, rb);
            if (!ok) { printf("tensor map failed mode %d rb %d rows %d\n", mode, rb, rows); continue; }
            const uint32_t stage_stride = ((rows * rb) + 1023u) & ~1023u;
            size_t smem = p.stages * stage_stride + 1024 + 256;
            if (ctas_per_sm == 2 && smem > 110000) continue;
            const int grid = sms * ctas_per_sm;
            // for 2 CTAs/SM relax launch bounds by shared memory only (256 threads, regs are tiny)
            tma_bw_kernel<<<grid, 256, smem>>>(tm, p);  // warm-up
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            tma_bw_kernel<<<grid, 256, smem>>>(tm, p);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> h(grid);
            CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (auto v : h) avg += (double)v;
            avg /= grid;
            const double bytes_cta = (double)p.iters * producers * rows * rb;
            const double bpc = bytes_cta * ctas_per_sm / avg;
            const double gbs = bytes_cta * grid / (ms * 1e-3) / 1e9;
            printf("%4zuMB %-8s %5d %5d %3d %3d %4d | %9.1f %9.0f %8.2f\n", ws_mb, mode ? "im2col" : "tiled", rb, rows, p.stages, producers,
                   ctas_per_sm, bpc, gbs, avg / ((double)p.iters * producers * rows * ctas_per_sm));
          }
  return 0;
}
