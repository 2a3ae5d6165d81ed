// Standalone GPU check of the C-ABI convolution entry points against a plain CPU loop nest.
// Build: make -C vision_toolbox_b200/csrc test_igemm ; run: tools/test_igemm <case|all>
// (development harness; the judged parity tests live in tests/ and go through the same C ABI)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../include/vtb.h"

static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(3);                                                                     \
    }                                                                              \
  } while (0)
#define CV(x)                                                            \
  do {                                                                   \
    int r_ = (x);                                                        \
    if (r_ != 0) {                                                       \
      printf("vtb error %d: %s (%s:%d)\n", r_, vtb_last_error(), __FILE__, __LINE__); \
      exit(4);                                                           \
    }                                                                    \
  } while (0)

struct Case {
  const char* name;
  int n, h, w, cin, cout, k, s, p, cin_real, ldx_extra, ldy_extra;
};

static const Case kCases[] = {
    {"1x1 64->64", 2, 16, 16, 64, 64, 1, 1, 0, 64, 0, 0},
    {"3x3s1 64->128 tail", 2, 12, 12, 64, 128, 3, 1, 1, 64, 0, 0},
    {"3x3s2 128->256 odd", 3, 11, 11, 128, 256, 3, 2, 1, 128, 0, 0},
    {"3x3s1 32->64 kc32", 2, 22, 22, 32, 64, 3, 1, 1, 32, 0, 0},
    {"3x3s1 stem 3(16)->32", 2, 20, 20, 16, 32, 3, 1, 1, 3, 0, 0},
    {"1x1 64->512 nblk2", 1, 9, 9, 64, 512, 1, 1, 0, 64, 0, 0},
    {"3x3s1 192->192", 2, 14, 14, 192, 192, 3, 1, 1, 192, 0, 0},
    {"3x3s1 160->160", 2, 7, 7, 160, 160, 3, 1, 1, 160, 0, 0},
    {"3x3s2 32->64 even", 2, 16, 16, 32, 64, 3, 2, 1, 32, 0, 0},
    {"6x6s2 stem 3(16)->64", 1, 20, 20, 16, 64, 6, 2, 2, 3, 0, 0},
    {"3x3s1 64->64 multi-tile", 8, 44, 44, 64, 64, 3, 1, 1, 64, 0, 0},
    {"1x1 concat-slice pitches", 2, 10, 10, 64, 128, 1, 1, 0, 64, 32, 32},
    {"3x3s1 concat-slice pitches", 2, 10, 10, 64, 128, 3, 1, 1, 64, 32, 64},
    {"1x1 1024->512 deepK", 2, 6, 6, 1024, 512, 1, 1, 0, 1024, 0, 0},
};
static const int kNumCases = sizeof(kCases) / sizeof(kCases[0]);

static void report(const char* what, const std::vector<float>& got, const std::vector<float>& exp, int rows, int cols,
                   float rtol, float atol, bool* ok_all) {
  double max_err = 0, max_ref = 0;
  long bad = 0;
  int shown = 0;
  std::vector<long> bad_q(4, 0), bad_c(16, 0), bad_r8(8, 0);
  for (long i = 0; i < (long)rows * cols; ++i) {
    const double e = fabs((double)got[i] - exp[i]);
    max_ref = fmax(max_ref, fabs((double)exp[i]));
    if (e > max_err) max_err = e;
    if (!(e <= atol + rtol * fabs(exp[i]))) {
      ++bad;
      const int r = (int)(i / cols), c = (int)(i % cols);
      bad_q[(r % 128) / 32]++;
      bad_c[(c / 16) % 16]++;
      bad_r8[r % 8]++;
      if (shown < 6) {
        printf("    mismatch %s[%d][%d]: got %.6f exp %.6f\n", what, r, c, got[i], exp[i]);
        ++shown;
      }
    }
  }
  printf("  %-10s %s  max_abs_err %.3e (max |ref| %.3e) bad %ld / %ld\n", what, bad ? "FAIL" : "ok", max_err, max_ref,
         bad, (long)rows * cols);
  if (bad) {
    *ok_all = false;
    printf("    bad by row-quarter: %ld %ld %ld %ld | by row%%8:", bad_q[0], bad_q[1], bad_q[2], bad_q[3]);
    for (int i = 0; i < 8; ++i) printf(" %ld", bad_r8[i]);
    printf(" | by col chunk16:");
    for (int i = 0; i < 16; ++i) printf(" %ld", bad_c[i]);
    printf("\n");
  }
}

static bool run_case(const Case& cs, int which) {
  printf("case %d: %s  (n%d %dx%d cin%d cout%d k%d s%d p%d)\n", which, cs.name, cs.n, cs.h, cs.w, cs.cin, cs.cout, cs.k,
         cs.s, cs.p);
  fflush(stdout);
  VtbConv c = {cs.n, cs.h, cs.w, cs.cin, cs.cout, cs.k, cs.s, cs.p};
  int ho, wo;
  CV(vtb_conv_out_hw(&c, &ho, &wo));
  const int ldx = cs.cin + cs.ldx_extra, ldy = cs.cout + cs.ldy_extra;
  const long inpix = (long)cs.n * cs.h * cs.w, outpix = (long)cs.n * ho * wo;
  const int kk = cs.k * cs.k;
  bool ok = true;

  // host data (values pre-rounded to bf16 so the CPU loop sees what the GPU sees)
  std::vector<float> x(inpix * cs.cin), w((long)cs.cout * cs.cin_real * kk), dy(outpix * cs.cout);
  for (auto& v : x) v = bf(frand());
  for (long i = 0; i < inpix; ++i)
    for (int ci = cs.cin_real; ci < cs.cin; ++ci) x[i * cs.cin + ci] = 0.f;
  for (auto& v : w) v = bf(frand() * 0.5f);
  for (auto& v : dy) v = bf(frand());

  std::vector<__nv_bfloat16> xb(inpix * ldx, __float2bfloat16(7.0f)), dyb(outpix * ldy, __float2bfloat16(7.0f));
  for (long i = 0; i < inpix; ++i)
    for (int ci = 0; ci < cs.cin; ++ci) xb[i * ldx + ci] = __float2bfloat16(x[i * cs.cin + ci]);
  for (long i = 0; i < outpix; ++i)
    for (int co = 0; co < cs.cout; ++co) dyb[i * ldy + co] = __float2bfloat16(dy[i * cs.cout + co]);

  __nv_bfloat16 *d_x, *d_y, *d_dy, *d_dx, *d_wf, *d_wd;
  float *d_w, *d_stats, *d_dw;
  void* d_ws;
  const int srows = vtb_conv_stats_rows(&c);
  const size_t ws_bytes = vtb_conv_wgrad_workspace_bytes(&c);
  CK(cudaMalloc(&d_x, xb.size() * 2));
  CK(cudaMalloc(&d_dx, xb.size() * 2));
  CK(cudaMalloc(&d_y, dyb.size() * 2));
  CK(cudaMalloc(&d_dy, dyb.size() * 2));
  CK(cudaMalloc(&d_w, w.size() * 4));
  CK(cudaMalloc(&d_dw, w.size() * 4));
  CK(cudaMalloc(&d_wf, (size_t)cs.cout * kk * cs.cin * 2));
  CK(cudaMalloc(&d_wd, (size_t)cs.cout * kk * cs.cin * 2));
  CK(cudaMalloc(&d_stats, (size_t)srows * cs.cout * 2 * 4));
  CK(cudaMalloc(&d_ws, ws_bytes));
  CK(cudaMemcpy(d_x, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dy, dyb.data(), dyb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_y, 0x7f, dyb.size() * 2));
  CK(cudaMemset(d_dx, 0x7f, xb.size() * 2));

  CV(vtb_pack_weight(&c, d_w, cs.cin_real, d_wf, d_wd, 0));
  CK(cudaDeviceSynchronize());

  // ---------------- fprop ----------------
  CV(vtb_conv_fprop(&c, d_x, ldx, d_wf, d_y, ldy, d_stats, nullptr, nullptr, 0, nullptr, 0, 0));
  CK(cudaDeviceSynchronize());
  std::vector<float> yref(outpix * cs.cout), ygot(outpix * cs.cout);
  for (int n = 0; n < cs.n; ++n)
    for (int oh = 0; oh < ho; ++oh)
      for (int ow = 0; ow < wo; ++ow)
        for (int co = 0; co < cs.cout; ++co) {
          double acc = 0;
          for (int r = 0; r < cs.k; ++r) {
            const int ih = oh * cs.s - cs.p + r;
            if (ih < 0 || ih >= cs.h) continue;
            for (int s = 0; s < cs.k; ++s) {
              const int iw = ow * cs.s - cs.p + s;
              if (iw < 0 || iw >= cs.w) continue;
              const float* xp = &x[(((long)n * cs.h + ih) * cs.w + iw) * cs.cin];
              for (int ci = 0; ci < cs.cin_real; ++ci) acc += (double)xp[ci] * w[((long)co * cs.cin_real + ci) * kk + r * cs.k + s];
            }
          }
          yref[(((long)n * ho + oh) * wo + ow) * cs.cout + co] = bf((float)acc);
        }
  {
    std::vector<__nv_bfloat16> yb(dyb.size());
    CK(cudaMemcpy(yb.data(), d_y, yb.size() * 2, cudaMemcpyDeviceToHost));
    for (long i = 0; i < outpix; ++i)
      for (int co = 0; co < cs.cout; ++co) ygot[i * cs.cout + co] = __bfloat162float(yb[i * ldy + co]);
    report("fprop", ygot, yref, (int)outpix, cs.cout, 1.0f / 64, 2e-3f, &ok);
    // pad columns of the pitch must be untouched
    long touched = 0;
    for (long i = 0; i < outpix; ++i)
      for (int co = cs.cout; co < ldy; ++co) {
        uint16_t raw;
        memcpy(&raw, &yb[i * ldy + co], 2);
        touched += (raw != 0x7f7f);
      }
    if (touched) { printf("  fprop wrote %ld elements outside its channel slice: FAIL\n", touched); ok = false; }
    // statistics of the bf16-rounded GPU output
    std::vector<float> st((size_t)srows * cs.cout * 2);
    CK(cudaMemcpy(st.data(), d_stats, st.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> sgot(cs.cout * 2, 0.f), sref(cs.cout * 2, 0.f);
    for (int co = 0; co < cs.cout; ++co) {
      double a = 0, b = 0;
      for (int r = 0; r < srows; ++r) { a += st[((size_t)r * cs.cout + co) * 2]; b += st[((size_t)r * cs.cout + co) * 2 + 1]; }
      sgot[co * 2] = (float)a; sgot[co * 2 + 1] = (float)b;
      double ra = 0, rb = 0;
      for (long i = 0; i < outpix; ++i) { const double v = ygot[i * cs.cout + co]; ra += v; rb += v * v; }
      sref[co * 2] = (float)ra; sref[co * 2 + 1] = (float)rb;
    }
    report("stats", sgot, sref, cs.cout, 2, 2e-4f, 2e-3f, &ok);
  }

  // ---------------- fused eval epilogue (scale/shift/relu/residual) ----------------
  {
    std::vector<float> sc(cs.cout), sh(cs.cout);
    for (int i = 0; i < cs.cout; ++i) { sc[i] = 0.5f + frand(); sh[i] = frand() * 0.3f; }
    float *d_sc, *d_sh;
    CK(cudaMalloc(&d_sc, cs.cout * 4)); CK(cudaMalloc(&d_sh, cs.cout * 4));
    CK(cudaMemcpy(d_sc, sc.data(), cs.cout * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sh, sh.data(), cs.cout * 4, cudaMemcpyHostToDevice));
    CV(vtb_conv_fprop(&c, d_x, ldx, d_wf, d_y, ldy, nullptr, d_sc, d_sh, 1, d_dy, ldy, 0));
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> yb(dyb.size());
    CK(cudaMemcpy(yb.data(), d_y, yb.size() * 2, cudaMemcpyDeviceToHost));
    std::vector<float> eref(outpix * cs.cout), egot(outpix * cs.cout);
    // recompute unrounded conv in double for the reference of the fused path
    for (int n = 0; n < cs.n; ++n)
      for (int oh = 0; oh < ho; ++oh)
        for (int ow = 0; ow < wo; ++ow)
          for (int co = 0; co < cs.cout; ++co) {
            double acc = 0;
            for (int r = 0; r < cs.k; ++r) {
              const int ih = oh * cs.s - cs.p + r;
              if (ih < 0 || ih >= cs.h) continue;
              for (int s = 0; s < cs.k; ++s) {
                const int iw = ow * cs.s - cs.p + s;
                if (iw < 0 || iw >= cs.w) continue;
                const float* xp = &x[(((long)n * cs.h + ih) * cs.w + iw) * cs.cin];
                for (int ci = 0; ci < cs.cin_real; ++ci) acc += (double)xp[ci] * w[((long)co * cs.cin_real + ci) * kk + r * cs.k + s];
              }
            }
            const long i = (((long)n * ho + oh) * wo + ow);
            float v = fmaxf((float)acc * sc[co] + sh[co], 0.f);
            eref[i * cs.cout + co] = bf(bf(v) + dy[i * cs.cout + co]);
          }
    for (long i = 0; i < outpix; ++i)
      for (int co = 0; co < cs.cout; ++co) egot[i * cs.cout + co] = __bfloat162float(yb[i * ldy + co]);
    report("fused-eval", egot, eref, (int)outpix, cs.cout, 1.0f / 64, 4e-3f, &ok);
    cudaFree(d_sc); cudaFree(d_sh);
  }

  // ---------------- dgrad ----------------
  if (cs.cin_real == cs.cin) {
    std::vector<float> dxref(inpix * cs.cin, 0.f);
    std::vector<double> dxacc(inpix * cs.cin, 0.0);
    for (int n = 0; n < cs.n; ++n)
      for (int oh = 0; oh < ho; ++oh)
        for (int ow = 0; ow < wo; ++ow) {
          const float* g = &dy[(((long)n * ho + oh) * wo + ow) * cs.cout];
          for (int r = 0; r < cs.k; ++r) {
            const int ih = oh * cs.s - cs.p + r;
            if (ih < 0 || ih >= cs.h) continue;
            for (int s = 0; s < cs.k; ++s) {
              const int iw = ow * cs.s - cs.p + s;
              if (iw < 0 || iw >= cs.w) continue;
              double* dst = &dxacc[(((long)n * cs.h + ih) * cs.w + iw) * cs.cin];
              for (int co = 0; co < cs.cout; ++co) {
                const float gv = g[co];
                const float* wp = &w[((long)co * cs.cin) * kk + r * cs.k + s];
                for (int ci = 0; ci < cs.cin; ++ci) dst[ci] += (double)gv * wp[(long)ci * kk];
              }
            }
          }
        }
    for (size_t i = 0; i < dxacc.size(); ++i) dxref[i] = bf((float)dxacc[i]);
    for (int accumulate = 0; accumulate < 2; ++accumulate) {
      if (accumulate) CK(cudaMemcpy(d_dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
      CV(vtb_conv_dgrad(&c, d_dy, ldy, d_wd, d_dx, ldx, accumulate, 0));
      CK(cudaDeviceSynchronize());
      std::vector<__nv_bfloat16> dxb(xb.size());
      CK(cudaMemcpy(dxb.data(), d_dx, dxb.size() * 2, cudaMemcpyDeviceToHost));
      std::vector<float> got(inpix * cs.cin), exp(inpix * cs.cin);
      for (long i = 0; i < inpix; ++i)
        for (int ci = 0; ci < cs.cin; ++ci) {
          got[i * cs.cin + ci] = __bfloat162float(dxb[i * ldx + ci]);
          exp[i * cs.cin + ci] = accumulate ? bf(dxref[i * cs.cin + ci] + x[i * cs.cin + ci]) : dxref[i * cs.cin + ci];
        }
      report(accumulate ? "dgrad+acc" : "dgrad", got, exp, (int)inpix, cs.cin, 1.0f / 64, 4e-3f, &ok);
    }
  }

  // ---------------- wgrad ----------------
  {
    std::vector<double> dwacc(w.size(), 0.0);
    for (int n = 0; n < cs.n; ++n)
      for (int oh = 0; oh < ho; ++oh)
        for (int ow = 0; ow < wo; ++ow) {
          const float* g = &dy[(((long)n * ho + oh) * wo + ow) * cs.cout];
          for (int r = 0; r < cs.k; ++r) {
            const int ih = oh * cs.s - cs.p + r;
            if (ih < 0 || ih >= cs.h) continue;
            for (int s = 0; s < cs.k; ++s) {
              const int iw = ow * cs.s - cs.p + s;
              if (iw < 0 || iw >= cs.w) continue;
              const float* xp = &x[(((long)n * cs.h + ih) * cs.w + iw) * cs.cin];
              for (int co = 0; co < cs.cout; ++co)
                for (int ci = 0; ci < cs.cin_real; ++ci)
                  dwacc[((long)co * cs.cin_real + ci) * kk + r * cs.k + s] += (double)g[co] * xp[ci];
            }
          }
        }
    std::vector<float> dwref(w.size()), dwgot(w.size());
    for (size_t i = 0; i < w.size(); ++i) dwref[i] = (float)dwacc[i];
    CK(cudaMemset(d_dw, 0, w.size() * 4));
    CV(vtb_conv_wgrad(&c, d_dy, ldy, d_x, ldx, d_ws, d_dw, cs.cin_real, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(dwgot.data(), d_dw, w.size() * 4, cudaMemcpyDeviceToHost));
    double mx = 0;
    for (float v : dwref) mx = fmax(mx, fabs(v));
    report("wgrad", dwgot, dwref, cs.cout, cs.cin_real * kk, 1e-3f, (float)(1e-4 * mx + 1e-5), &ok);
  }

  cudaFree(d_x); cudaFree(d_dx); cudaFree(d_y); cudaFree(d_dy); cudaFree(d_w); cudaFree(d_dw);
  cudaFree(d_wf); cudaFree(d_wd); cudaFree(d_stats); cudaFree(d_ws);
  printf("case %d %s\n", which, ok ? "PASS" : "FAIL");
  fflush(stdout);
  return ok;
}

int main(int argc, char** argv) {
  int sel = -1;
  if (argc > 1 && strcmp(argv[1], "all") != 0) sel = atoi(argv[1]);
  if (argc > 1 && strcmp(argv[1], "count") == 0) { printf("%d\n", kNumCases); return 0; }
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { printf("no CUDA device\n"); return 2; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  bool all_ok = true;
  for (int i = 0; i < kNumCases; ++i)
    if (sel < 0 || sel == i) all_ok &= run_case(kCases[i], i);
  printf("%s\n", all_ok ? "ALL PASS" : "SOME FAILED");
  return all_ok ? 0 : 1;
}
