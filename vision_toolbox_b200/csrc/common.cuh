// Shared host-side helpers of libvtb_b200 (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>

namespace vtb {
int fail(int code, const char* fmt, ...);
int check_cuda(int cuda_error, const char* what);
void count_launch(int n);
int num_sms();
}  // namespace vtb
