// Shared host-side helpers of libvtb_b200 (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>

namespace vtb {
int fail(int code, const char* fmt, ...);
int check_cuda(int cuda_error, const char* what);
void count_launch(int n);
int num_sms();
// VTB_PDL=0 disables programmatic dependent launch (A/B switch, read once)
bool pdl_enabled();

// Programmatic dependent launch (PDL): every kernel of this library is launched with programmatic stream serialization
// and executes pdl_wait() before its first access to global memory (read OR write), then pdl_trigger().  The next
// kernel's thread blocks may therefore be scheduled - and run their prologue: barrier init, TMEM allocation, tensor-map
// prefetch - while this kernel drains; they block in pdl_wait() until this grid has completed and flushed.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
}  // namespace vtb
