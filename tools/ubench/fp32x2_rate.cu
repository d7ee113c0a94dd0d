// Throughput of scalar FADD/FMUL versus packed FADD2/FMUL2 on sm_100a (development aid:
// decides whether the packed forms are worth using in the blur kernels; --fmad=false).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o fp32x2_rate fp32x2_rate.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
  const float2 c = make_float2(seed * 0.5f, seed * 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }       // 2 FADD
        if (MODE == 1) { a[i] = add2(a[i], c); }                                                     // 1 FADD2
        if (MODE == 2) { a[i].x = __fmul_rn(a[i].x, c.x); a[i].y = __fmul_rn(a[i].y, c.y); }        // 2 FMUL
        if (MODE == 3) { a[i] = mul2(a[i], c); }                                                     // 1 FMUL2
        if (MODE == 4) { float2 p = make_float2(__fmul_rn(a[i].y, c.x), __fmul_rn(a[i].x, c.y)); a[i] = add2(a[i], p); }  // 2 FMUL + FADD2
        if (MODE == 5) { float px = __fmul_rn(a[i].y, c.x), py = __fmul_rn(a[i].x, c.y); a[i].x = __fadd_rn(a[i].x, px); a[i].y = __fadd_rn(a[i].y, py); }  // 2 FMUL + 2 FADD
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double lane_ops_per_iter) {
  const int ctas = 148 * 8, iters = 2000;
  float* out;
  cudaMalloc(&out, sizeof(float) * ctas * 256);
  k<MODE><<<ctas, 256>>>(out, 1.0f, 10);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<ctas, 256>>>(out, 1.0f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = lane_ops_per_iter * iters * ctas * 256.0;
  printf("%-22s %8.3f ms  %8.2f Tflop-lane-ops/s\n", name, ms, ops / ms / 1e9);
  cudaFree(out);
}

int main() {
  run<0>("FADD x2 (scalar)", 128);
  run<1>("FADD2", 128);
  run<2>("FMUL x2 (scalar)", 128);
  run<3>("FMUL2", 128);
  run<4>("2 FMUL + FADD2", 256);
  run<5>("2 FMUL + 2 FADD", 256);
  return 0;
}
