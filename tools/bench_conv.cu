// Times the C-ABI convolution entry points on one geometry and prints where each kernel role waited.
//   tools/bench_conv n h w cin cout k stride pad [iters]
// Env overrides (development): VTB_BLOCK_M, VTB_BLOCK_N, VTB_STAGES.
// Build: make -C vision_toolbox_b200/csrc bench_conv
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../include/vtb.h"

extern "C" void vtb_debug_counters(unsigned long long* dev_buf);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(3);} } while (0)
#define CV(x) do { int r_ = (x); if (r_ != 0) { printf("vtb error %d: %s line %d\n", r_, vtb_last_error(), __LINE__); exit(4);} } while (0)

static void fill(__nv_bfloat16* d, size_t n, unsigned seed) {
  std::vector<__nv_bfloat16> h(n);
  unsigned s = seed;
  for (size_t i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    h[i] = __float2bfloat16(((s >> 8) & 0xFFFF) / 65536.0f - 0.5f);
  }
  CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
}

int main(int argc, char** argv) {
  if (argc < 9) { printf("usage: bench_conv n h w cin cout k stride pad [iters]\n"); return 1; }
  VtbConv c = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8])};
  const int iters = argc > 9 ? atoi(argv[9]) : 20;
  int ho, wo;
  CV(vtb_conv_out_hw(&c, &ho, &wo));
  const size_t inpix = (size_t)c.n * c.h * c.w, outpix = (size_t)c.n * ho * wo;
  const int kk = c.k * c.k;
  __nv_bfloat16 *x, *y, *dy, *dx, *wf, *wd;
  float *w, *stats, *dw;
  void* ws;
  unsigned long long* dbg;
  const int srows = vtb_conv_stats_rows(&c);
  CK(cudaMalloc(&x, inpix * c.cin * 2)); CK(cudaMalloc(&dx, inpix * c.cin * 2));
  CK(cudaMalloc(&y, outpix * c.cout * 2)); CK(cudaMalloc(&dy, outpix * c.cout * 2));
  CK(cudaMalloc(&w, (size_t)c.cout * c.cin * kk * 4)); CK(cudaMalloc(&dw, (size_t)c.cout * c.cin * kk * 4));
  CK(cudaMalloc(&wf, (size_t)c.cout * c.cin * kk * 2)); CK(cudaMalloc(&wd, (size_t)c.cout * c.cin * kk * 2));
  CK(cudaMalloc(&stats, (size_t)srows * c.cout * 8));
  CK(cudaMalloc(&ws, vtb_conv_wgrad_workspace_bytes(&c) + 16));
  CK(cudaMalloc(&dbg, 1024 * 16 * 8));
  fill(x, inpix * c.cin, 1); fill(dy, outpix * c.cout, 2);
  { std::vector<float> hw((size_t)c.cout * c.cin * kk); unsigned s = 7; for (auto& v : hw) { s = s * 1664525u + 1013904223u; v = (((s >> 8) & 0xFFFF) / 65536.0f - 0.5f) * 0.1f; }
    CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice)); }
  CV(vtb_pack_weight(&c, w, c.cin, wf, wd, 0));
  // L2 flush buffer between timed iterations (activations of the big layers exceed L2 anyway)
  void* flush; const size_t flush_bytes = 256ull << 20; CK(cudaMalloc(&flush, flush_bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double flops = 2.0 * outpix * c.cout * kk * c.cin;
  const double bytes = 2.0 * (inpix * c.cin + outpix * c.cout) + 2.0 * kk * c.cin * c.cout;
  const char* names[5] = {"fprop", "dgrad", "wgrad", "dgradbn", "dgr+acc"};
  // fake producer layer for the dgrad that carries BatchNorm-backward statistics (vtb_conv_dgrad_bn)
  __nv_bfloat16* yprod; float *bnp, *partial_d; unsigned int* tickets;
  CK(cudaMalloc(&yprod, inpix * c.cin * 2)); fill(yprod, inpix * c.cin, 5);
  CK(cudaMalloc(&bnp, (size_t)c.cin * 8 * 4));
  { std::vector<float> hb((size_t)c.cin * 8, 0.5f); CK(cudaMemcpy(bnp, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice)); }
  const int drows = vtb_conv_dgrad_stats_rows(&c);
  CK(cudaMalloc(&partial_d, (size_t)(drows > 0 ? drows : 1) * c.cin * 8));
  CK(cudaMalloc(&tickets, 4096)); CK(cudaMemset(tickets, 0, 4096));
  VtbDgradBn dbn; memset(&dbn, 0, sizeof(dbn));
  dbn.layer[0].y = yprod; dbn.layer[0].ldy = c.cin; dbn.layer[0].scale = bnp; dbn.layer[0].shift = bnp + c.cin;
  dbn.layer[0].mean = bnp + 2 * c.cin; dbn.layer[0].invstd = bnp + 3 * c.cin; dbn.layer[0].relu = 1;
  dbn.layer[0].dgamma = bnp + 4 * c.cin; dbn.layer[0].dbeta = bnp + 5 * c.cin; dbn.layer[0].coef = bnp + 6 * c.cin;
  dbn.count = (double)inpix; dbn.partial = partial_d; dbn.tickets = tickets;
  const bool flush_l2 = getenv("VTB_FLUSH") != nullptr;
  for (int which = 0; which < 5; ++which) {
    auto run = [&]() {
      if (which == 0) CV(vtb_conv_fprop(&c, x, c.cin, wf, y, c.cout, stats, nullptr, nullptr, 0, nullptr, 0, 0));
      if (which == 1) CV(vtb_conv_dgrad(&c, dy, c.cout, wd, dx, c.cin, 0, 0));
      if (which == 2) CV(vtb_conv_wgrad(&c, dy, c.cout, x, c.cin, ws, dw, c.cin, 0, 0));
      if (which == 3) CV(vtb_conv_dgrad_bn(&c, dy, c.cout, wd, dx, c.cin, 0, &dbn, 0));
      if (which == 4) CV(vtb_conv_dgrad(&c, dy, c.cout, wd, dx, c.cin, 1, 0));
    };
    vtb_debug_counters(nullptr);
    {  // warm up until the clocks have ramped (idle GPUs sit far below boost for the first tens of ms)
      cudaEventRecord(e0);
      float el = 0;
      const float warm_ms = getenv("VTB_WARM_MS") ? (float)atof(getenv("VTB_WARM_MS")) : 250.f;
      while (el < warm_ms) {
        for (int i = 0; i < 10; ++i) run();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&el, e0, e1);
      }
    }
    float total = 0, best = 1e30f;
    for (int i = 0; i < iters; ++i) {
      if (flush_l2) CK(cudaMemsetAsync(flush, i, flush_bytes));
      cudaEventRecord(e0); run(); cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      total += ms; if (ms < best) best = ms;
    }
    double avg = total / iters;
    if (getenv("VTB_GRAPH")) {
      // GPU-side time per launch with the host out of the picture: 20 back-to-back launches replayed from a CUDA graph
      cudaStream_t cs; CK(cudaStreamCreate(&cs));
      auto run_s = [&](cudaStream_t st) {
        if (which == 0) CV(vtb_conv_fprop(&c, x, c.cin, wf, y, c.cout, stats, nullptr, nullptr, 0, nullptr, 0, st));
        if (which == 1) CV(vtb_conv_dgrad(&c, dy, c.cout, wd, dx, c.cin, 0, st));
        if (which == 2) CV(vtb_conv_wgrad(&c, dy, c.cout, x, c.cin, ws, dw, c.cin, 0, st));
        if (which == 3) CV(vtb_conv_dgrad_bn(&c, dy, c.cout, wd, dx, c.cin, 0, &dbn, st));
        if (which == 4) CV(vtb_conv_dgrad(&c, dy, c.cout, wd, dx, c.cin, 1, st));
      };
      cudaGraph_t graph; cudaGraphExec_t exec;
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < 20; ++i) run_s(cs);
      CK(cudaStreamEndCapture(cs, &graph));
      CK(cudaGraphInstantiate(&exec, graph, 0));
      CK(cudaGraphLaunch(exec, cs)); CK(cudaStreamSynchronize(cs));
      cudaEventRecord(e0, cs);
      for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(exec, cs));
      cudaEventRecord(e1, cs);
      CK(cudaStreamSynchronize(cs));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%-6s graph replay: %8.1f us per launch (host-free, back to back)\n", names[which], ms * 1e3 / 100);
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph); cudaStreamDestroy(cs);
    }
    printf("%-6s %4d->%4d k%ds%d %3dx%-3d n%-3d | avg %8.1f us  best %8.1f us | %7.1f TF/s %7.1f GB/s\n", names[which], c.cin, c.cout,
           c.k, c.stride, c.h, c.w, c.n, avg * 1e3, best * 1e3, flops / (avg * 1e-3) / 1e12, bytes / (avg * 1e-3) / 1e9);
    if (which == 2) {
      CK(cudaMemset(dbg, 0, 1024 * 128));
      vtb_debug_counters(dbg);
      run();
      CK(cudaDeviceSynchronize());
      vtb_debug_counters(nullptr);
      std::vector<unsigned long long> h(1024 * 16);
      CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
      double s[16] = {0}; int n = 0;
      for (int b = 0; b < 1024; ++b) if (h[b * 16 + 2]) { ++n; for (int j = 0; j < 16; ++j) s[j] += (double)h[b * 16 + j]; }
      if (n) printf("       ctas %d | cycles/CTA %.0f | setup %.0f | first-full %.0f | producer0 on empty %.0f | MMA on full %.0f | epi on tmem-full %.0f, drain %.0f\n",
                    n, s[2] / n, s[6] / n, s[5] / n, s[0] / n, s[1] / n, s[3] / n, s[4] / n);
    }
    if (which != 2 && !(which != 0 && c.stride == 2)) {
      CK(cudaMemset(dbg, 0, 1024 * 128));
      vtb_debug_counters(dbg);
      run();
      CK(cudaDeviceSynchronize());
      vtb_debug_counters(nullptr);
      std::vector<unsigned long long> h(1024 * 16);
      CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
      double s[16] = {0}; int n = 0;
      for (int b = 0; b < 1024; ++b) if (h[b * 16 + 7]) { ++n; for (int j = 0; j < 16; ++j) s[j] += (double)h[b * 16 + j]; }
      {
        unsigned long long lo = ~0ull, hi = 0; double cyc = 0;
        for (int b = 0; b < 1024; ++b) if (h[b * 16 + 7]) { lo = h[b * 16 + 11] < lo ? h[b * 16 + 11] : lo; hi = h[b * 16 + 12] > hi ? h[b * 16 + 12] : hi; cyc += (double)h[b * 16 + 13]; }
        double mma = 0, epi = 0;
        for (int b = 0; b < 1024; ++b) if (h[b * 16 + 7]) { mma += (double)h[b * 16 + 15]; epi += (double)h[b * 16 + 14]; }
        if (n) printf("       kernel span (first CTA start -> last CTA done): %.2f us | full CTA lifetime %.0f cycles avg | MMA issuer done at %.0f, slowest epilogue warp done at %.0f\n", (hi - lo) * 1e-3, cyc / n, mma / n, epi / n);
      }
      if (n) printf("       ctas %d | cycles/CTA %.0f | waits: A0 %.0f A1 %.0f B %.0f (on empty) | MMA on full %.0f, on tmem-empty %.0f | epi0 %.0f epi1 %.0f (on tmem-full) | epi warp4: tmem-ld %.0f cvt+sts %.0f drain %.0f\n",
                    n, s[7] / n, s[0] / n, s[1] / n, s[2] / n, s[3] / n, s[4] / n, s[5] / n, s[6] / n, s[8] / n, s[9] / n, s[10] / n);
    }
  }
  if (getenv("VTB_INTERLEAVE")) {
    // Do kernels run slower when DIFFERENT kernels alternate (instruction caches, L2 state) than when one kernel replays
    // back to back?  (a) fprop -> dgrad -> wgrad round-robin; (b) fprop alternating with a small unrelated kernel.
    cudaStream_t cs; CK(cudaStreamCreate(&cs));
    __nv_bfloat16 *ga, *gb; CK(cudaMalloc(&ga, 1 << 20)); CK(cudaMalloc(&gb, 1 << 20));
    CK(cudaMemset(ga, 0, 1 << 20)); CK(cudaMemset(gb, 0, 1 << 20));
    for (int mode = 0; mode < 3; ++mode) {
      cudaGraph_t graph; cudaGraphExec_t exec;
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      for (int i = 0; i < 20; ++i) {
        if (mode == 0) {
          CV(vtb_conv_fprop(&c, x, c.cin, wf, y, c.cout, stats, nullptr, nullptr, 0, nullptr, 0, cs));
          CV(vtb_conv_dgrad(&c, dy, c.cout, wd, dx, c.cin, 0, cs));
          CV(vtb_conv_wgrad(&c, dy, c.cout, x, c.cin, ws, dw, c.cin, 0, cs));
        } else if (mode == 1) {
          CV(vtb_conv_fprop(&c, x, c.cin, wf, y, c.cout, stats, nullptr, nullptr, 0, nullptr, 0, cs));
          CV(vtb_grad_add(ga, 64, gb, 64, 4096, 64, 0, cs));
        } else {
          CV(vtb_grad_add(ga, 64, gb, 64, 4096, 64, 0, cs));
        }
      }
      CK(cudaStreamEndCapture(cs, &graph));
      CK(cudaGraphInstantiate(&exec, graph, 0));
      CK(cudaGraphLaunch(exec, cs)); CK(cudaStreamSynchronize(cs));
      cudaEventRecord(e0, cs);
      for (int r = 0; r < 5; ++r) CK(cudaGraphLaunch(exec, cs));
      cudaEventRecord(e1, cs);
      CK(cudaStreamSynchronize(cs));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const char* what[3] = {"fprop+dgrad+wgrad round-robin", "fprop + small unrelated kernel", "small unrelated kernel alone"};
      printf("interleave: %-32s %8.1f us per iteration\n", what[mode], ms * 1e3 / 100);
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    }
    cudaStreamDestroy(cs);
  }
  return 0;
}
