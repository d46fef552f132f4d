// tc_forward.cu -- tcgen05 / TMEM implementation of the forward-only MLP pass (sm_100a).
//
// Same contract as mlp_forward_kernel (model.act's network pass, the target-network / target-critic pass of the
// learners; marlbase/dqn/model.py:99,132-134, marlbase/ac/model.py:148-149,190-193) but the three GEMMs of each 128-row
// tile run on the 5th-generation tensor cores:
//   * activations never touch shared memory: row r of a tile = TMEM lane r, 16 warps each own a lane quarter x column quarter;
//     the A operand of every layer is written with tcgen05.st, the accumulator is read back with tcgen05.ld, bias + ReLU
//     happen in registers;
//   * weights are the B operand, resident in shared memory as a pre-packed image (K-major, 128-byte swizzle, one
//     16-KB panel per 32 input features) built by pack_weights_kernel and kept current by the optimiser step (pack_param);
//   * FP32 parity (<= 1e-5) is kept with the error-compensated 3xTF32 split: every operand is stored as
//     hi = tf32(x) and lo = tf32(x - hi) and each product is accumulated as lo*hi + hi*lo + hi*hi in the FP32 TMEM
//     accumulator (measured 4e-7 relative on a 128x128x128 product, tools/tc_probe.cu).
// One elected thread issues tcgen05.mma (kind::tf32, cta_group::1, M = 128, N = 128 or 16); completion reaches the
// other threads through tcgen05.commit -> mbarrier.
#include "tc_common.cuh"

namespace marl {

size_t tc_image_bytes() { return kImageBytes; }
size_t tc_bwd_image_bytes() { return kBwdImageBytes; }

// Whole images from the flat parameters: the padding (observation columns >= in, head rows >= out) is zeroed, everything else goes
// through pack_param (tc_common.cuh).
__global__ void pack_weights_kernel(const float* __restrict__ theta, NetLayout lay, int n_nets, uint8_t* __restrict__ image, uint8_t* __restrict__ bwd_image) {
  const int net = blockIdx.y;
  if (net >= n_nets) return;
  const float* th = theta + (size_t)net * lay.P;
  uint8_t* img = image + (size_t)net * kImageBytes;
  uint8_t* bwd = bwd_image ? bwd_image + (size_t)net * kBwdImageBytes : nullptr;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kHidden * 32 && (i & 31) >= lay.in) {  // W1 [128][32]: zero columns beyond in
    const int o = panel_offset(i >> 5, i & 31, kPanelBytes);
    *reinterpret_cast<float*>(img + kOffW1Hi + o) = 0.f; *reinterpret_cast<float*>(img + kOffW1Lo + o) = 0.f;
  }
  if (i < kHeadRows * kHidden && (i >> 7) >= lay.out) {  // W3 [16][128]: zero rows beyond out (also in the FP32 copy [8][128])
    const int o = panel_offset(i >> 7, i & 127, kHeadPanelBytes);
    *reinterpret_cast<float*>(img + kOffW3Hi + o) = 0.f; *reinterpret_cast<float*>(img + kOffW3Lo + o) = 0.f;
    if (i < kOutPad * kHidden) reinterpret_cast<float*>(img + kOffW3F)[i] = 0.f;
  }
  if (i < kHeadRows && i >= lay.out) reinterpret_cast<float*>(img + kOffB3)[i] = 0.f;
  if (i < lay.P) pack_param(lay, i, th[i], img, bwd);
}


// 3xTF32: D (+)= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi over KSTEPS steps of 8 features; issued by one thread.
// Fully unrolled: every operand address is base + compile-time constant (the descriptor's address field counts
// 16-byte units, so advancing by b bytes is desc + (b >> 4)); the issuing thread spends a handful of instructions per
// MMA and the tensor pipe, not the issue loop, sets the pace.
template <int KSTEPS, int N, int PANEL_BYTES>
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t d_col, uint32_t b_hi_addr, uint32_t b_lo_addr, int ksteps_rt) {
  const uint32_t idesc = idesc_tf32(N);
  const uint64_t dhi = kmajor_desc(b_hi_addr), dlo = kmajor_desc(b_lo_addr);
  const uint32_t d = tmem + d_col, ahi = tmem + kColAHi, alo = tmem + kColALo;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      if (ks < ksteps_rt) {
        constexpr int kDummy = 0; (void)kDummy;
        const uint32_t off16 = (uint32_t)(((ks >> 2) * PANEL_BYTES + (ks & 3) * 32) >> 4);
        mma_tf32_ts(d, (term == 0 ? alo : ahi) + ks * 8, (term == 1 ? dlo : dhi) + off16, idesc, (term | ks) ? 1u : 0u);
      }
    }
  }
}

// 16 warps: warp w owns TMEM lane quarter (w & 3) -- rows 32 (w & 3) .. +31 of the tile -- and column quarter (w >> 2) of the
// 128 hidden features, so that four warps per SM sub-partition overlap their TMEM / shared / global latencies.
__global__ void __launch_bounds__(kTrThreads, 1) tc_forward_kernel(FwdParams p, const uint8_t* __restrict__ images) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);  // swizzle atoms need 1024-byte alignment
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kImageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lq = warp & 3, cq = warp >> 2, r = 32 * lq + (t & 31), c0 = 32 * cq;
  int net, row_begin, row_end;
  cta_rows(p.plan, net, row_begin, row_end);
  if (row_begin >= row_end) { pdl_wait(); return; }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) { mbar_init(bar, 1); mbar_init(bar + 2, 1); mbar_init(bar + 3, 1); mbar_init(bar + 4, 1); fence_mbar_init(); }
  pdl_wait();   // nothing above touches global memory (PDL contract, common.cuh)
  pdl_launch_dependents();
  // the weight image is already in shared-memory layout: three TMA bulk copies (cp.async.bulk -> mbarrier, issued by one thread) in the
  // order the first tile needs them (W1 + biases, W2, W3), so that its first layer does not wait for the whole image
  if (t == 0) tma_forward_image(smem_u32(smem), images + (size_t)net * kImageBytes, bar + 2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t lane_base = tmem + ((uint32_t)(32 * lq) << 16);
  const float* b1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* b2 = reinterpret_cast<const float*>(smem + kOffB2);
  const float* b3 = reinterpret_cast<const float*>(smem + kOffB3);
  const int D = p.src.D, out = p.lay.out;
  const int k1steps = (D + 7) >> 3;
  const bool x_active = cq < k1steps;   // column quarter cq stages observation columns [8 cq, 8 cq + 8)
  int image_groups_pending = 3;         // block-uniform
  uint32_t parity = 0;
  // this thread's 8 observation columns of its row (prefetched one tile ahead) and where the row's outputs go
  // (two steps, a barrier apart, so that neither waits on a load it has just issued: A = decode + episode index, B = the columns)
  float xin[8], xnext[8];
  size_t dst_row = 0, dst_next = 0;
  struct RowKey { int agent, unit, off, ep; bool valid; };
  auto fetch_a = [&](int vr0, int nrows, RowKey& k, size_t& dst) {
    k.agent = 0; k.unit = 0; k.off = 0; k.ep = 0; k.valid = r < nrows;
    if (k.valid) {
      decode_row(p.plan, net, vr0 + r, k.agent, k.unit, k.off);
      dst = p.src.mode == 0 ? ((size_t)k.unit * p.src.N + k.agent) : (((size_t)k.agent * p.plan.units_per_agent + k.unit) * p.plan.unit_rows + k.off);
      if (p.src.mode != 0) k.ep = p.src.idx[k.unit];
    }
  };
  auto fetch_b = [&](const RowKey& k, float (&x)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.f;
    if (k.valid && x_active) {
      const TrajView& tv = p.src.traj;
      const float* src = p.src.mode == 0 ? p.src.dense + ((size_t)k.unit * p.src.N + k.agent) * D
                                         : tv.obs + (((size_t)k.ep * tv.N + k.agent) * (size_t)(tv.T + 1) + k.off) * D;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (8 * cq + j < D) ? src[8 * cq + j] : 0.f;
    }
  };
  RowKey key_nxt;
  fetch_a(row_begin, min(kTileRows, row_end - row_begin), key_nxt, dst_row);
  fetch_b(key_nxt, xin);

  for (int vr0 = row_begin; vr0 < row_end; vr0 += kTileRows) {
    const int nrows = min(kTileRows, row_end - vr0);
    const bool has_next = vr0 + kTileRows < row_end;
    if (has_next) fetch_a(vr0 + kTileRows, min(kTileRows, row_end - vr0 - kTileRows), key_nxt, dst_next);
    // ---- input row -> A operand (hi / lo), zero padded to k1steps * 8 features -----------------------------------
    if (x_active) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tf32_split(xin[j], hi[j], lo[j]);
      tmem_st8(lane_base + kColAHi + 8 * cq, hi);
      tmem_st8(lane_base + kColALo + 8 * cq, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (image_groups_pending == 3) { mbar_wait(bar + 2, 0); image_groups_pending = 2; }   // W1 + biases have landed
    tc_fence_before();
    __syncthreads();
    // ---- layer 1 -------------------------------------------------------------------------------------------------
    if (t == 0) {
      tc_fence_after();
      issue_layer<kMaxObsDim / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW1Hi, smem_base + kOffW1Lo, k1steps);
      mma_commit(bar);
    }
    // prefetch the next tile's rows: the loads stay in flight under this tile's epilogues and MMAs
    if (has_next) fetch_b(key_nxt, xnext);
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
    // ---- bias + ReLU, next A operand (twice: after layer 1 and after layer 2): this thread's 32 columns --------------------
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const float* bias = (layer == 0 ? b1 : b2) + c0;
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(lane_base + kColD + c0, ra);
      tmem_ld16_issue(lane_base + kColD + c0 + 16, rb);
      tmem_ld_wait(ra);
      tmem_ld_wait(rb);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t (&acc)[16] = half ? rb : ra;
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float h = fmaxf(__uint_as_float(acc[j]) + bias[16 * half + j], 0.f);
          tf32_split(h, hi[j], lo[j]);
        }
        tmem_st16(lane_base + kColAHi + c0 + 16 * half, hi);
        tmem_st16(lane_base + kColALo + c0 + 16 * half, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (image_groups_pending == 2 - layer) {   // W2 before the layer-2 MMAs, W3 before the head's
        mbar_wait(bar + 3 + layer, 0);
        image_groups_pending = 1 - layer;
      }
      tc_fence_before();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        if (layer == 0) issue_layer<kHidden / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW2Hi, smem_base + kOffW2Lo, kHidden / 8);
        else issue_layer<kHidden / 8, kHeadRows, kHeadPanelBytes>(tmem, kColDHead, smem_base + kOffW3Hi, smem_base + kOffW3Lo, kHidden / 8);
        mma_commit(bar);
      }
      mbar_wait(bar, parity); parity ^= 1;
      tc_fence_after();
    }
    // ---- head epilogue: outputs to global (column quarter 0) ---------------------------------------------------------------
    if (cq == 0) {
      float v[16];
      tmem_ld16(lane_base + kColDHead, v);
      if (r < nrows) {
        float* dst = p.out + dst_row * out;
#pragma unroll
        for (int o = 0; o < kOutPad; ++o) if (o < out) dst[o] = v[o] + b3[o];
      }
    }
    dst_row = dst_next;
#pragma unroll
    for (int j = 0; j < 8; ++j) xin[j] = xnext[j];
    // the next tile's first MMAs write D / read A only after the __syncthreads that follows its operand staging, by which time
    // every warp has finished reading this tile's accumulators
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- launchers ----------------------------------------------------------------------------------------------------------
int tc_forward_init() {
  MARL_CUDA_TRY(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  return MARL_OK;
}

int launch_pack_weights(const float* theta, const NetLayout& lay, int n_nets, uint8_t* image, cudaStream_t st, uint8_t* bwd_image) {
  dim3 grid((lay.P + 255) / 256, n_nets);   // P > 128 * 128 >= every padded extent
  pack_weights_kernel<<<grid, 256, 0, st>>>(theta, lay, n_nets, image, bwd_image);
  MARL_CUDA_TRY(cudaGetLastError());
  return MARL_OK;
}

int launch_tc_forward(const FwdParams& p, const uint8_t* images, cudaStream_t st) {
  MARL_CUDA_TRY(launch_pdl(tc_forward_kernel, dim3(p.plan.cta_begin[p.plan.n_nets]), dim3(kTrThreads), kTcSmemBytes, st, p, images));
  return MARL_OK;
}

}  // namespace marl
