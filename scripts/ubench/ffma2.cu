// ffma2.cu — latency / issue rate of FFMA, FFMA2 (scalar-broadcast form) and of the PGS row chain
// (FMNMX -> FADD -> FFMA2 -> FMNMX ...) for ONE warp per SM sub-partition.  Diagnostic for DESIGN.md §5.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 f2(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
template <int MODE>
__global__ void __launch_bounds__(32) k(float* out, int iters, float x, float y, long long* cyc) {
  float t = threadIdx.x * 1e-3f;
  float2 a0 = make_float2(t, t + 1), a1 = make_float2(t + 2, t + 3), a2 = make_float2(t + 4, t + 5), a3 = make_float2(t + 6, t + 7), a4 = make_float2(t + 8, t + 9), a5 = make_float2(t + 10, t + 11);
  float2 w0 = make_float2(x, y), w1 = make_float2(y, x), w2 = make_float2(x, x), w3 = make_float2(y, y), w4 = make_float2(x * y, y), w5 = make_float2(x, x * y);
  float s = x, lam = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 64; i++) {
      if (MODE == 0) { a0.x = fmaf(a0.x, x, y); }                                  // dependent FFMA
      if (MODE == 1) { a0 = f2(a0, s, w0); }                                        // dependent FFMA2
      if (MODE == 2) { a0 = f2(w0, s, a0); a1 = f2(w1, s, a1); a2 = f2(w2, s, a2); a3 = f2(w3, s, a3); a4 = f2(w4, s, a4); a5 = f2(w5, s, a5); }   // 6 independent FFMA2 (one PGS row update)
      if (MODE == 3) { a0.x = fmaf(w0.x, s, a0.x); a0.y = fmaf(w0.y, s, a0.y); a1.x = fmaf(w1.x, s, a1.x); a1.y = fmaf(w1.y, s, a1.y); a2.x = fmaf(w2.x, s, a2.x); a2.y = fmaf(w2.y, s, a2.y);
                       a3.x = fmaf(w3.x, s, a3.x); a3.y = fmaf(w3.y, s, a3.y); a4.x = fmaf(w4.x, s, a4.x); a4.y = fmaf(w4.y, s, a4.y); a5.x = fmaf(w5.x, s, a5.x); a5.y = fmaf(w5.y, s, a5.y); }   // 12 independent FFMA
      if (MODE == 4) {   // PGS normal-row chain: clamp -> delta -> packed update of 6 pairs, next row reads a different component
        float g = (i & 1) ? a0.y : a1.x;
        float ln = fmaxf(g, 0.f); float dl = lam - ln; lam = ln;
        a0 = f2(w0, dl, a0); a1 = f2(w1, dl, a1); a2 = f2(w2, dl, a2); a3 = f2(w3, dl, a3); a4 = f2(w4, dl, a4); a5 = f2(w5, dl, a5);
      }
      if (MODE == 5) {   // same with scalar FFMAs
        float g = (i & 1) ? a0.y : a1.x;
        float ln = fmaxf(g, 0.f); float dl = lam - ln; lam = ln;
        a0.x = fmaf(w0.x, dl, a0.x); a0.y = fmaf(w0.y, dl, a0.y); a1.x = fmaf(w1.x, dl, a1.x); a1.y = fmaf(w1.y, dl, a1.y); a2.x = fmaf(w2.x, dl, a2.x); a2.y = fmaf(w2.y, dl, a2.y);
        a3.x = fmaf(w3.x, dl, a3.x); a3.y = fmaf(w3.y, dl, a3.y); a4.x = fmaf(w4.x, dl, a4.x); a4.y = fmaf(w4.y, dl, a4.y); a5.x = fmaf(w5.x, dl, a5.x); a5.y = fmaf(w5.y, dl, a5.y);
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + lam;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* what, int grid) {
  float* out; long long* cyc; cudaMalloc(&out, grid * 32 * 4); cudaMalloc(&cyc, grid * 8);
  int iters = 2000;
  k<MODE><<<grid, 32>>>(out, iters, 0.999f, 0.5f, cyc);
  k<MODE><<<grid, 32>>>(out, iters, 0.999f, 0.5f, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; i++) s += h[i];
  printf("{\"what\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_unrolled_step\": %.2f}\n", what, grid / 592, s / grid / (64.0 * iters));
  cudaFree(out); cudaFree(cyc); delete[] h;
}
int main() {
  for (int grid : {592, 1184}) {
    run<0>("dependent FFMA", grid); run<1>("dependent FFMA2 (broadcast)", grid); run<2>("6 independent FFMA2", grid); run<3>("12 independent FFMA", grid);
    run<4>("PGS row: FMNMX+FADD+6 FFMA2", grid); run<5>("PGS row: FMNMX+FADD+12 FFMA", grid);
  }
  return 0;
}
