// microbenchmark: legacy mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16) and FFMA peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tf32(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
  float c[32]; for (int j = 0; j < 32; ++j) c[j] = j;
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = fmaf(c[j], b, a);
  }
  float s = 0; for (int j = 0; j < 32; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {
    int threads = wps * 32 > 1024 ? 1024 : wps * 32, blocks = 148 * (wps * 32 / threads);
    float ms = timeit([&] { k_tf32<<<blocks, threads>>>(out, iters); });
    double fl = (double)blocks * (threads / 32) * iters * 8 * (2.0 * 16 * 8 * 8);
    printf("tf32 mma.sync m16n8k8  warps/SM=%2d: %.1f TFLOP/s\n", wps, fl / ms / 1e9);
    ms = timeit([&] { k_bf16<<<blocks, threads>>>(out, iters); });
    fl = (double)blocks * (threads / 32) * iters * 8 * (2.0 * 16 * 8 * 16);
    printf("bf16 mma.sync m16n8k16 warps/SM=%2d: %.1f TFLOP/s\n", wps, fl / ms / 1e9);
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters); });
    fl = (double)blocks * threads * iters * 32 * 2.0;
    printf("ffma                   warps/SM=%2d: %.1f TFLOP/s\n", wps, fl / ms / 1e9);
  }
  return 0;
}
