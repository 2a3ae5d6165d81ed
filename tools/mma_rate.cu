// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, cta_group::1) issued from one thread with operands
// resident in shared memory (no TMA in the loop). Variables: N, number of independent accumulators, whether the K
// slices walk inside one 128B-swizzled tile (the conv kernel's pattern) and how many CTAs run per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vision_toolbox_b200/csrc/ptx.cuh"

using namespace vtb;

struct P {
  int m;         // UMMA M (64 | 128)
  int n;         // UMMA N
  int naccs;     // independent accumulators cycled through
  int iters;     // k-stages
  int stages;    // distinct smem stages cycled through (A: 16 KB each, B: n*128 B each)
  int commit_every;  // tcgen05.commit + wait every this many stages (0 = only at the end)
  long long* out;
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t a_stage = 128 * 128;
  const uint32_t b_stage = p.n * 128;
  const uint32_t a_base = base, b_base = base + p.stages * a_stage;
  // fill operands with small finite numbers
  for (uint32_t i = threadIdx.x; i < (p.stages * (a_stage + b_stage)) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(p.m, p.n, 0, 0);
    const uint32_t stride = 512 / p.naccs;
    uint32_t parity = 0;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t a_st = a_base + (it % p.stages) * a_stage;
      const uint32_t b_st = b_base + (it % p.stages) * b_stage;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint64_t adesc = make_smem_desc(a_st + k16 * 32, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(b_st + k16 * 32, 16, 1024, 2);
        for (int a = 0; a < p.naccs; ++a) umma_bf16(tmem_base + a * stride, adesc, bdesc, idesc, it > 0 || k16 > 0);
      }
      if (p.commit_every && (it + 1) % p.commit_every == 0) {
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1u;
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), parity);
    p.out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* out;
  CK(cudaMalloc(&out, 8 * 1024));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048));
  printf("%4s %5s %6s %6s %7s | %10s %10s %9s\n", "N", "naccs", "stages", "commit", "grid", "cyc/MMA", "floor", "TF/s@chip");
  const int warm = getenv("WARM") ? atoi(getenv("WARM")) : 0;
  for (int m : {128})
  for (int n : {128, 256})
    for (int naccs : {1, 2, 4})
      for (int stages : {4})
        for (int commit_every : {0})
          for (int grid : {sms}) {
            if (naccs * n > 512) continue;
            P p{m, n, naccs, 2000, stages, commit_every, out};
            const size_t smem = stages * (128 * 128 + n * 128) + 2048;
            if (smem > 220000) continue;
            for (int w = 0; w < 1 + warm; ++w) mma_rate_kernel<<<grid, 128, smem>>>(p);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            mma_rate_kernel<<<grid, 128, smem>>>(p);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            std::vector<long long> h(grid);
            CK(cudaMemcpy(h.data(), out, 8 * grid, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : h) avg += (double)v; avg /= grid;
            const double mmas = (double)p.iters * 4 * naccs;
            const double flops = mmas * 2.0 * m * n * 16 * grid;
            printf("M%3d %4d %5d %6d %6d %7d | %10.1f %10.1f %9.1f\n", m, n, naccs, stages, commit_every, grid, avg / mmas, 128.0 * n / 256.0,
                   flops / (ms * 1e-3) / 1e12);
          }
  return 0;
}
