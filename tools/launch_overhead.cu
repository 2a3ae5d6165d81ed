// Microbenchmark: per-launch cost of a chain of dependent kernels replayed from a CUDA graph, as a function of the
// launch configuration (threads, dynamic shared memory, TMEM allocation, PDL attribute, kernel parameter size).
//   tools/launch_overhead            (prints a table)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(3);} } while (0)

struct Big { char pad[1280]; };

template <bool TMEM, bool PDLWAIT>
__global__ void __launch_bounds__(640, 1) chain_kernel(float* buf, int spin, const __grid_constant__ Big big) {
  extern __shared__ unsigned char smem[];
  __shared__ unsigned int slot;
  if (TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (PDLWAIT) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  // a dependent read-modify-write so that consecutive launches really depend on each other
  if (threadIdx.x == 0) {
    float v = buf[blockIdx.x] + big.pad[0];
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    buf[blockIdx.x] = v + 1.f;
    smem[0] = 1;
  }
  __syncthreads();
  if (TMEM && threadIdx.x < 32) {
    unsigned t = slot;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
  }
}

template <typename K>
static float run(K kernel, int grid, int threads, size_t smem, bool pdl, int spin, float* buf) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t cs; CK(cudaStreamCreate(&cs));
  Big big = {};
  cudaGraph_t graph; cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < 50; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = cs;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, buf, spin, big));
  }
  CK(cudaStreamEndCapture(cs, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  CK(cudaGraphLaunch(exec, cs)); CK(cudaStreamSynchronize(cs));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, cs);
  for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(exec, cs));
  cudaEventRecord(e1, cs);
  CK(cudaStreamSynchronize(cs));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); cudaStreamDestroy(cs);
  return ms * 1e3f / 500;
}

int main() {
  float* buf; CK(cudaMalloc(&buf, 4096)); CK(cudaMemset(buf, 0, 4096));
  const int spin = 2000;   // ~1 us of "work" per CTA
  printf("per-launch time (us) of a 50-kernel dependent chain in a CUDA graph; each CTA spins %d cycles\n", spin);
  printf("%-44s %8s %8s\n", "config", "no PDL", "PDL");
  struct Cfg { const char* name; int grid, threads; size_t smem; bool tmem; } cfgs[] = {
      {"148 CTAs x 640 thr, 227 KB smem, TMEM 512", 148, 640, 231424, true},
      {"148 CTAs x 640 thr, 227 KB smem, no TMEM", 148, 640, 231424, false},
      {"148 CTAs x 640 thr, 100 KB smem, TMEM 512", 148, 640, 102400, true},
      {"148 CTAs x 640 thr, 100 KB smem, no TMEM", 148, 640, 102400, false},
      {"148 CTAs x 640 thr,   0 KB smem, no TMEM", 148, 640, 0, false},
      {"148 CTAs x 256 thr,   0 KB smem, no TMEM", 148, 256, 0, false},
      {"592 CTAs x 512 thr,   0 KB smem, no TMEM", 592, 512, 0, false},
  };
  for (auto& c : cfgs) {
    float a, b;
    if (c.tmem) { a = run(chain_kernel<true, false>, c.grid, c.threads, c.smem, false, spin, buf); b = run(chain_kernel<true, true>, c.grid, c.threads, c.smem, true, spin, buf); }
    else { a = run(chain_kernel<false, false>, c.grid, c.threads, c.smem, false, spin, buf); b = run(chain_kernel<false, true>, c.grid, c.threads, c.smem, true, spin, buf); }
    printf("%-44s %8.2f %8.2f\n", c.name, a, b);
  }
  return 0;
}
