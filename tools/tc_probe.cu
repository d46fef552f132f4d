// tc_probe.cu -- standalone bring-up of the tcgen05 pieces the tensor-core forward kernel needs (sm_100a):
//   TMEM alloc / dealloc, tcgen05.st (registers -> TMEM A operand), tcgen05.mma kind::tf32 with A in TMEM and B in
//   shared memory (K-major, 128B swizzle), tcgen05.commit -> mbarrier, tcgen05.ld (TMEM accumulator -> registers).
// Computes D[128][128] = A[128][128] * B[128][128]^T (B row-major [n][k], i.e. an nn.Linear weight) once with plain TF32
// and once with the 3xTF32 split (hi*hi + lo*hi + hi*lo), and prints the max error against an FP64 reference.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int M = 128, N = 128, K = 128;
constexpr int kPanelBytes = N * 128;  // one K-block of 32 floats for all N rows

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                      // layout type SWIZZLE_128B
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both K-major, M = 128, N
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

#define TMEM_ST8(addr, v)                                                                                       \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),      \
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), \
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory")

#define TMEM_LD8(addr, v)                                                                                              \
  do {                                                                                                                 \
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                        \
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(addr));         \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");                                                       \
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);     \
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);     \
  } while (0)

// mode 0: single-pass TF32, mode 1: 3xTF32
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* b_hi = smem;                              // 4 panels x 16 KB
  uint8_t* b_lo = smem + 4 * kPanelBytes;            // 4 panels x 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * kPanelBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) mbar_init(bar, 1);
  // B operand: row n, K-block p, 16-byte chunk c (4 floats) -> panel p, row n (128 B), physical chunk c ^ (n & 7)
  for (int i = t; i < N * K / 4; i += 128) {
    const int n = i / (K / 4), c4 = i % (K / 4);  // c4: chunk along K (0..31)
    const int p = c4 >> 3, c = c4 & 7;
    const float4 v = *reinterpret_cast<const float4*>(B + n * K + 4 * c4);
    float4 hi, lo;
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = tf32_hi(v.x - hi.x); lo.y = tf32_hi(v.y - hi.y); lo.z = tf32_hi(v.z - hi.z); lo.w = tf32_hi(v.w - hi.w);
    const int off = p * kPanelBytes + n * 128 + ((c ^ (n & 7)) << 4);
    *reinterpret_cast<float4*>(b_hi + off) = (mode == 0) ? v : hi;
    *reinterpret_cast<float4*>(b_lo + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes
  const uint32_t col_ahi = 0, col_alo = 128, col_d = 256;

  // A operand: thread t owns row t; A_hi in columns [0,128), A_lo in [128,256)
  for (int k0 = 0; k0 < K; k0 += 8) {
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = A[t * K + k0 + j];
      hi[j] = (mode == 0) ? x : tf32_hi(x);
      lo[j] = tf32_hi(x - tf32_hi(x));
    }
    TMEM_ST8(lane_base + col_ahi + k0, hi);
    TMEM_ST8(lane_base + col_alo + k0, lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  if (t == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(N);
    uint32_t acc = 0;
    const int n_terms = mode == 0 ? 1 : 3;
    for (int term = 0; term < n_terms; ++term) {
      // small terms first: lo*hi, hi*lo, then hi*hi
      const bool a_lo = (n_terms == 3 && term == 0), b_lo_sel = (n_terms == 3 && term == 1);
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t a_addr = tmem + (a_lo ? col_alo : col_ahi) + ks * 8;
        const uint8_t* bbase = (b_lo_sel ? b_lo : b_hi) + (ks >> 2) * kPanelBytes + (ks & 3) * 32;
        mma_tf32_ts(tmem + col_d, a_addr, make_desc(smem_u32(bbase)), idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    TMEM_LD8(lane_base + col_d + n0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) D[t * N + n0 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  std::vector<float> hA(M * K), hB(N * K), hD(M * N);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& x : hA) x = rnd() * 3.f;
  for (auto& x : hB) x = rnd();
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 8 * kPanelBytes + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<double> ref(M * N);
  double scale = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)hA[m * K + k] * hB[n * K + k];
      ref[m * N + n] = acc;
      scale = fmax(scale, fabs(acc));
    }
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(dD, 0, hD.size() * 4));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs(hD[i] - ref[i]));
    printf("mode %d (%s): max |err| = %.3e  (max |ref| = %.3f, rel %.3e)  D[0][0]=%.6f ref=%.6f  D[127][127]=%.6f ref=%.6f\n", mode,
           mode == 0 ? "1xTF32" : "3xTF32", maxerr, scale, maxerr / scale, hD[0], ref[0], hD[M * N - 1], ref[M * N - 1]);
  }
  return 0;
}
