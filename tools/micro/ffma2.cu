// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2, sm_100a) issue throughput.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b),
                        c = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(c));
  d = *reinterpret_cast<float2*>(&r);
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float2 acc[8];
  float2 a = make_float2(s + threadIdx.x, s * 0.5f), b = make_float2(1.0001f, 0.9999f);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a.x), "f"(b.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a.y), "f"(b.y));
      } else {
        ffma2(acc[i], a, b);
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 1 << 16;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 256>>>(out, iters, 1.f); else k<1><<<148 * 4, 256>>>(out, iters, 1.f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 4 * 256 * iters * 16.0;
      printf("mode %s: %.3f ms  %.2f TFLOP/s fp32 (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9,
             fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  }
  return 0;
}
