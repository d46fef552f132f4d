// tc_train.cu -- tensor-core (tcgen05 / TMEM, 3xTF32) training pass of the DQN-family learner as a three-kernel pipeline.
//
// Same arithmetic as train_kernel<KP, kHeadDqn> (QNetwork._compute_loss + backward, marlbase/dqn/model.py:118-168), split where
// one SM's shared memory / TMEM cannot hold every operand twice (hi / lo) at once:
//   tc_dqn_fwd_kernel   online forward (A operand in TMEM, weights = K-major image), TD head in registers, dH2 = (dq x W3) * relu';
//                        stores H1, H2, dH2 (FP32 rows) and dq for the next two kernels
//   tc_dh1_kernel       dH1 = (dH2 x W2) * relu'(H1): A = dH2 in TMEM, B = W2 as an MN-major operand (SWIZZLE_128B_BASE32B image)
//   tc_dw_kernel        dW2 | db2, dW1 | db1, dW3 (+ db3): row-streaming weight-gradient GEMMs, both operands MN-major from shared
//                        memory, accumulators resident in TMEM across all the CTA's rows, flushed once into the per-CTA partial
// The partials feed the same grad_reduce_kernel / adam_kernel as the FP32 path.
#include "tc_common.cuh"

namespace marl {

struct TcTrainParams {
  RowPlan plan; RowSource src; NetLayout lay;
  const uint8_t* images;      // forward images [n_nets][kImageBytes]
  const uint8_t* bwd_images;  // backward images [n_nets][kBwdImageBytes]
  float* q_out;               // [rows][out] online outputs (optional)
  float* h1g; float* h2g; float* dh2g; float* dh1g; float* dqg;  // [rows][128] x4, [rows][8]
  const float* tq; const float* td_ext; float gamma; int double_q;
  float* scratch; int scratch_pitch; float* loss_part;
};

__device__ __forceinline__ size_t dst_of(const RowPlan& plan, const RowSource& src, int net, int vr, int& agent, int& unit, int& off) {
  decode_row(plan, net, vr, agent, unit, off);
  return src.mode == 0 ? ((size_t)unit * src.N + agent) : (((size_t)agent * plan.units_per_agent + unit) * plan.unit_rows + off);
}

__device__ __forceinline__ void store16(float* dst, const float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void load16(const float* src, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 x = reinterpret_cast<const float4*>(src)[j];
    v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
  }
}

// issue helpers (one thread): 3xTF32, TS form, compile-time unrolled
template <int KSTEPS, int N, int PANEL_BYTES>
__device__ __forceinline__ void issue_kmajor(uint32_t tmem, uint32_t d_col, uint32_t b_hi, uint32_t b_lo, int ksteps_rt) {
  const uint32_t idesc = idesc_tf32(N);
  const uint64_t dhi = kmajor_desc(b_hi), dlo = kmajor_desc(b_lo);
#pragma unroll
  for (int term = 0; term < 3; ++term)
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      if (ks < ksteps_rt)
        mma_tf32_ts(tmem + d_col, tmem + (term == 0 ? kColALo : kColAHi) + ks * 8,
                    (term == 1 ? dlo : dhi) + (uint32_t)(((ks >> 2) * PANEL_BYTES + (ks & 3) * 32) >> 4), idesc, (term | ks) ? 1u : 0u);
}

// =====================================================================================================================
// 1. online forward + TD head + dH2
// =====================================================================================================================
__global__ void __launch_bounds__(kTcThreads, 1) tc_dqn_fwd_kernel(TcTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kImageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* qs = reinterpret_cast<float*>(smem + kImageBytes + 64);   // [128][8] outputs of this tile, then loss reduction scratch
  float* carry = qs + kTileRows * kOutPad;                         // [8] outputs of the first row of the previously processed (higher) tile
  const int t = threadIdx.x, warp = t >> 5;
  int net, row_begin, row_end;
  cta_rows(p.plan, net, row_begin, row_end);
  float st[4] = {0.f, 0.f, 0.f, 0.f};
  if (row_begin >= row_end) {
    if (t < 4) p.loss_part[4 * blockIdx.x + t] = 0.f;
    return;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) mbar_init(bar, 1);
  {
    const uint8_t* src = p.images + (size_t)net * kImageBytes;
    const uint32_t dst = smem_u32(smem);
#pragma unroll 4
    for (int i = t; i < kImageBytes / 16; i += kTcThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * (size_t)i) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, smem_base = smem_u32(smem), lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const float* b1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* b2 = reinterpret_cast<const float*>(smem + kOffB2);
  const float* b3 = reinterpret_cast<const float*>(smem + kOffB3);
  const float* w3f = reinterpret_cast<const float*>(smem + kOffW3F);
  const int D = p.src.D, A = p.lay.out, T = p.src.traj.T, B = p.plan.units_per_agent;
  const int k1steps = (D + 7) >> 3;
  uint32_t parity = 0;

  // tiles from the top of the chunk downwards (the double-Q argmax needs the next row's outputs)
  for (int vr_hi = row_end; vr_hi > row_begin; vr_hi -= kTileRows) {
    const int vr0 = max(row_begin, vr_hi - kTileRows), nrows = vr_hi - vr0;
    // ---- this thread's row: source pointer, destination row, loss-head scalars ---------------------------------------
    int agent = 0, b = 0, tt = 0, act = 0; float rew = 0.f, filled = 0.f, done1 = 0.f;
    size_t dst_row = 0;
    float xin[kMaxObsDim];
    {
      const float* src = nullptr;
      if (t < nrows) {
        dst_row = dst_of(p.plan, p.src, net, vr0 + t, agent, b, tt);
        src = row_ptr(p.src, agent, b, tt);
        if (tt < T) {
          const TrajView& tv = p.src.traj;
          const size_t ep = (size_t)p.src.idx[b];
          act = tv.act[(ep * tv.N + agent) * T + tt];
          rew = tv.rew[(ep * tv.N + agent) * T + tt];
          filled = (float)tv.filled[ep * T + tt];
          done1 = (float)tv.done[ep * (T + 1) + tt + 1];
        }
      }
#pragma unroll
      for (int j = 0; j < kMaxObsDim; ++j) xin[j] = (src != nullptr && j < D) ? src[j] : 0.f;
    }
#pragma unroll
    for (int k0 = 0; k0 < kMaxObsDim; k0 += 16) {
      if (k0 < k1steps * 8) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { hi[j] = tf32_rn(xin[k0 + j]); lo[j] = tf32_rn(xin[k0 + j] - hi[j]); }
        tmem_st16(lane_base + kColAHi + k0, hi);
        tmem_st16(lane_base + kColALo + k0, lo);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      issue_kmajor<kMaxObsDim / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW1Hi, smem_base + kOffW1Lo, k1steps);
      mma_commit(bar);
    }
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const float* bias = layer == 0 ? b1 : b2;
      float* hg = (layer == 0 ? p.h1g : p.h2g) + dst_row * kHidden;
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(lane_base + kColD, ra);
      tmem_ld_wait(ra);
#pragma unroll
      for (int c = 0; c < kHidden / 16; ++c) {
        uint32_t (&cur)[16] = (c & 1) ? rb : ra;
        uint32_t (&nxt)[16] = (c & 1) ? ra : rb;
        if (c + 1 < kHidden / 16) tmem_ld16_issue(lane_base + kColD + 16 * (c + 1), nxt);
        float h[16], hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          h[j] = fmaxf(__uint_as_float(cur[j]) + bias[16 * c + j], 0.f);
          hi[j] = tf32_rn(h[j]); lo[j] = tf32_rn(h[j] - hi[j]);
        }
        tmem_st16(lane_base + kColAHi + 16 * c, hi);
        tmem_st16(lane_base + kColALo + 16 * c, lo);
        if (t < nrows) store16(hg + 16 * c, h);
        if (c + 1 < kHidden / 16) tmem_ld_wait(nxt);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        if (layer == 0) issue_kmajor<kHidden / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW2Hi, smem_base + kOffW2Lo, kHidden / 8);
        else issue_kmajor<kHidden / 8, kHeadRows, kHeadPanelBytes>(tmem, kColDHead, smem_base + kOffW3Hi, smem_base + kOffW3Lo, kHidden / 8);
        mma_commit(bar);
      }
      mbar_wait(bar, parity); parity ^= 1;
      tc_fence_after();
    }
    // ---- outputs of this tile -> shared (next-row exchange), TD head ----------------------------------------------------
    float q[kOutPad];
    {
      float v[16];
      tmem_ld16(lane_base + kColDHead, v);
#pragma unroll
      for (int o = 0; o < kOutPad; ++o) q[o] = o < A ? v[o] + b3[o] : 0.f;
    }
    *reinterpret_cast<float4*>(qs + t * kOutPad) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(qs + t * kOutPad + 4) = make_float4(q[4], q[5], q[6], q[7]);
    __syncthreads();
    float dq[kOutPad];
#pragma unroll
    for (int o = 0; o < kOutPad; ++o) dq[o] = 0.f;
    if (t < nrows) {
      if (p.q_out) for (int o = 0; o < A; ++o) p.q_out[dst_row * A + o] = q[o];
      if (tt < T) {
        float g;
        if (p.td_ext) {
          g = p.td_ext[(size_t)b * T + tt];
        } else {
          const float* qn = (t + 1 < nrows) ? (qs + (t + 1) * kOutPad) : carry;
          const float* tq = p.tq + (((size_t)agent * B + b) * (T + 1) + tt + 1) * A;
          float tsel;
          if (p.double_q) {
            int best = 0; float bv = qn[0];
            for (int o = 1; o < A; ++o) if (qn[o] > bv) { bv = qn[o]; best = o; }
            tsel = tq[best];
          } else {
            tsel = tq[0];
            for (int o = 1; o < A; ++o) tsel = fmaxf(tsel, tq[o]);
          }
          const float y = rew + p.gamma * tsel * (1.f - done1);
          const float delta = q[act] - y;
          st[0] += delta * delta * filled;
          if (agent == 0) st[1] += filled;
          g = 2.f * delta * filled;
        }
#pragma unroll
        for (int o = 0; o < kOutPad; ++o) dq[o] = (o == act) ? g : 0.f;
      }
      float* dqd = p.dqg + dst_row * kOutPad;
      *reinterpret_cast<float4*>(dqd) = make_float4(dq[0], dq[1], dq[2], dq[3]);
      *reinterpret_cast<float4*>(dqd + 4) = make_float4(dq[4], dq[5], dq[6], dq[7]);
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int o = 0; o < kOutPad; ++o) carry[o] = q[o];
    }
    // ---- dH2[r][j] = (sum_a dq[r][a] W3[a][j]) * (H2[r][j] > 0): H2's hi part is still in the A region of TMEM -----------------
    if (true) {
      float* dst = p.dh2g + dst_row * kHidden;
#pragma unroll 1
      for (int c = 0; c < kHidden / 16; ++c) {
        float hh[16], g[16];
        tmem_ld16(lane_base + kColAHi + 16 * c, hh);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float s = 0.f;
#pragma unroll
          for (int o = 0; o < kOutPad; ++o) s = fmaf(dq[o], w3f[o * kHidden + 16 * c + j], s);
          g[j] = hh[j] > 0.f ? s : 0.f;
        }
        if (t < nrows) store16(dst + 16 * c, g);
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // ---- per-CTA loss statistics -----------------------------------------------------------------------------------------
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) qs[k * kTcThreads + t] = st[k];
  __syncthreads();
  for (int s = kTcThreads / 2; s > 0; s >>= 1) {
    if (t < s) {
#pragma unroll
      for (int k = 0; k < 4; ++k) qs[k * kTcThreads + t] += qs[k * kTcThreads + t + s];
    }
    __syncthreads();
  }
  if (t < 4) p.loss_part[4 * blockIdx.x + t] = qs[t * kTcThreads];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// =====================================================================================================================
// 2. dH1 = (dH2 x W2) * relu'(H1)
// =====================================================================================================================
__global__ void __launch_bounds__(kTcThreads, 1) tc_dh1_kernel(TcTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBwdImageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5;
  int net, row_begin, row_end;
  cta_rows(p.plan, net, row_begin, row_end);
  if (row_begin >= row_end) return;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) mbar_init(bar, 1);
  {
    const uint8_t* src = p.bwd_images + (size_t)net * kBwdImageBytes;
    const uint32_t dst = smem_u32(smem);
#pragma unroll 4
    for (int i = t; i < kBwdImageBytes / 16; i += kTcThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * (size_t)i) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, smem_base = smem_u32(smem), lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc = idesc_tf32_major(kHidden, 0, 1);
  uint32_t parity = 0;
  for (int vr0 = row_begin; vr0 < row_end; vr0 += kTileRows) {
    const int nrows = min(kTileRows, row_end - vr0);
    size_t dst_row = 0;
    if (t < nrows) { int a, u, o; dst_row = dst_of(p.plan, p.src, net, vr0 + t, a, u, o); }
    const float* gin = p.dh2g + dst_row * kHidden;
#pragma unroll 1
    for (int c = 0; c < kHidden / 16; ++c) {
      float v[16], hi[16], lo[16];
      if (t < nrows) load16(gin + 16 * c, v);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { hi[j] = tf32_rn(v[j]); lo[j] = tf32_rn(v[j] - hi[j]); }
      tmem_st16(lane_base + kColAHi + 16 * c, hi);
      tmem_st16(lane_base + kColALo + 16 * c, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      // D[r][j1] = sum_{j2} dH2[r][j2] W2[j2][j1]: k-step = 8 rows of the [k = j2][n = j1] image (1024 bytes), panels 16 KB apart
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int ks = 0; ks < kHidden / 8; ++ks)
          mma_tf32_ts(tmem + kColD, tmem + (term == 0 ? kColALo : kColAHi) + ks * 8,
                      mnmajor_desc(smem_base + (term == 1 ? 4 * kPanelBytes : 0) + ks * 1024, kPanelBytes), idesc, (term | ks) ? 1u : 0u);
      mma_commit(bar);
    }
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
    const float* h1 = p.h1g + dst_row * kHidden;
    float* gout = p.dh1g + dst_row * kHidden;
#pragma unroll 1
    for (int c = 0; c < kHidden / 16; ++c) {
      float d[16], h[16];
      tmem_ld16(lane_base + kColD + 16 * c, d);
      if (t < nrows) {
        load16(h1 + 16 * c, h);
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = h[j] > 0.f ? d[j] : 0.f;
        store16(gout + 16 * c, d);
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// =====================================================================================================================
// 3. weight gradients: row-streaming TN GEMMs, accumulators in TMEM
// =====================================================================================================================
constexpr int kDwThreads = 256;
constexpr int kChunkRows = 32;
constexpr int kChunkPanel = kChunkRows * 128;        // one 32-feature panel of a 32-row chunk
constexpr int kOpBytes = 4 * kChunkPanel;            // a [32 rows][128 features] operand (hi or lo)
// shared-memory map of one chunk (bytes): dH2 hi|lo, H1 hi (+ ones panel) | lo (+ zero panel), dH1 hi|lo, H2 hi|lo, X hi|lo, dq hi|lo
constexpr int kSDh2 = 0, kSH1 = kSDh2 + 2 * kOpBytes, kSDh1 = kSH1 + 2 * (kOpBytes + kChunkPanel), kSH2 = kSDh1 + 2 * kOpBytes;
constexpr int kSX = kSH2 + 2 * kOpBytes, kSDq = kSX + 2 * kChunkPanel, kSEnd = kSDq + 2 * kChunkPanel;
constexpr int kDwSmemBytes = kSEnd + 256 + 1024;
// TMEM columns: dW2 | db2 [0,160), dW1 | db1 [160,192), dW3^T [192,208)
constexpr uint32_t kColW2 = 0, kColW1 = 160, kColW3 = 192;

__device__ __forceinline__ void stage4(uint8_t* hi_img, uint8_t* lo_img, int r, int col, float4 v) {
  float4 h, l;
  h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
  l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
  const int off = mn_offset(r, col, kChunkPanel);
  *reinterpret_cast<float4*>(hi_img + off) = h;
  *reinterpret_cast<float4*>(lo_img + off) = l;
}

__global__ void __launch_bounds__(kDwThreads, 1) tc_dw_kernel(TcTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSEnd);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  long long* rowmap = reinterpret_cast<long long*>(smem + kSEnd + 16);  // [32] destination row (-1: padding), fits in 256 bytes
  const int t = threadIdx.x, warp = t >> 5;
  int net, row_begin, row_end;
  cta_rows(p.plan, net, row_begin, row_end);
  float* gs = p.scratch + (size_t)blockIdx.x * p.scratch_pitch;
  if (row_begin >= row_end) {
    for (int i = t; i < p.lay.P; i += kDwThreads) gs[i] = 0.f;
    return;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (t == 0) mbar_init(bar, 1);
  // constant panels: ones column (n = 128) behind H1 hi, zeros behind H1 lo; X / dq panels are rewritten per chunk
  for (int i = t; i < kChunkPanel / 4; i += kDwThreads) {
    reinterpret_cast<float*>(smem + kSH1 + kOpBytes)[i] = 0.f;
    reinterpret_cast<float*>(smem + kSH1 + 2 * kOpBytes + kChunkPanel)[i] = 0.f;
  }
  __syncthreads();
  if (t < kChunkRows) *reinterpret_cast<float*>(smem + kSH1 + kOpBytes + mn_offset(t, 0, kChunkPanel)) = 1.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, sb = smem_u32(smem);
  const int D = p.src.D, A = p.lay.out;
  const uint32_t id_w2 = idesc_tf32_major(160, 1, 1), id_w1 = idesc_tf32_major(32, 1, 1), id_w3 = idesc_tf32_major(16, 1, 1);
  uint32_t parity = 0, acc = 0;
  float db3 = 0.f;

  for (int vr0 = row_begin; vr0 < row_end; vr0 += kChunkRows) {
    const int nrows = min(kChunkRows, row_end - vr0);
    if (t < kChunkRows) {
      long long d = -1;
      if (t < nrows) { int a, u, o; d = (long long)dst_of(p.plan, p.src, net, vr0 + t, a, u, o); }
      rowmap[t] = d;
    }
    __syncthreads();
    // ---- stage the four [32][128] operands: 1024 float4 each, 4 per thread ------------------------------------------------
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = t + q * kDwThreads, r = i >> 5, c4 = i & 31;
      const long long d = rowmap[r];
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
      if (d >= 0) {
        v0 = reinterpret_cast<const float4*>(p.dh2g + d * kHidden)[c4];
        v1 = reinterpret_cast<const float4*>(p.h1g + d * kHidden)[c4];
        v2 = reinterpret_cast<const float4*>(p.dh1g + d * kHidden)[c4];
        v3 = reinterpret_cast<const float4*>(p.h2g + d * kHidden)[c4];
      }
      stage4(smem + kSDh2, smem + kSDh2 + kOpBytes, r, 4 * c4, v0);
      stage4(smem + kSH1, smem + kSH1 + kOpBytes + kChunkPanel, r, 4 * c4, v1);
      stage4(smem + kSDh1, smem + kSDh1 + kOpBytes, r, 4 * c4, v2);
      stage4(smem + kSH2, smem + kSH2 + kOpBytes, r, 4 * c4, v3);
    }
    // ---- X (| ones at column D) and dq panels ------------------------------------------------------------------------------
    for (int i = t; i < kChunkRows * 32; i += kDwThreads) {
      const int r = i >> 5, k = i & 31;
      float x = 0.f;
      if (r < nrows) {
        if (k < D) { int a, u, o; decode_row(p.plan, net, vr0 + r, a, u, o); x = row_ptr(p.src, a, u, o)[k]; }
        else if (k == D) x = 1.f;
      }
      const float hi = tf32_rn(x);
      *reinterpret_cast<float*>(smem + kSX + mn_offset(r, k, kChunkPanel)) = hi;
      *reinterpret_cast<float*>(smem + kSX + kChunkPanel + mn_offset(r, k, kChunkPanel)) = tf32_rn(x - hi);
      float g = 0.f;
      const long long d = rowmap[r];
      if (k < kOutPad && d >= 0) g = p.dqg[d * kOutPad + k];
      const float gh = tf32_rn(g);
      *reinterpret_cast<float*>(smem + kSDq + mn_offset(r, k, kChunkPanel)) = gh;
      *reinterpret_cast<float*>(smem + kSDq + kChunkPanel + mn_offset(r, k, kChunkPanel)) = tf32_rn(g - gh);
    }
    if (t < A) {
      for (int r = 0; r < nrows; ++r) db3 += p.dqg[rowmap[r] * kOutPad + t];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t a_sel = term == 0 ? 1u : 0u, b_sel = term == 1 ? 1u : 0u;  // lo*hi, hi*lo, hi*hi
#pragma unroll
        for (int ks = 0; ks < kChunkRows / 8; ++ks) {
          const uint32_t ko = ks * 1024;
          const uint32_t accf = (acc | term | ks) ? 1u : 0u;
          // dW2[j2][j1 | 1] += dH2^T x [H1 | 1]
          mma_tf32_ss(tmem + kColW2, mnmajor_desc(sb + kSDh2 + a_sel * kOpBytes + ko, kChunkPanel),
                      mnmajor_desc(sb + kSH1 + b_sel * (kOpBytes + kChunkPanel) + ko, kChunkPanel), id_w2, accf);
          // dW1[j1][i | 1] += dH1^T x [X | 1]
          mma_tf32_ss(tmem + kColW1, mnmajor_desc(sb + kSDh1 + a_sel * kOpBytes + ko, kChunkPanel),
                      mnmajor_desc(sb + kSX + b_sel * kChunkPanel + ko, kChunkPanel), id_w1, accf);
          // dW3^T[j][a] += H2^T x dq
          mma_tf32_ss(tmem + kColW3, mnmajor_desc(sb + kSH2 + a_sel * kOpBytes + ko, kChunkPanel),
                      mnmajor_desc(sb + kSDq + b_sel * kChunkPanel + ko, kChunkPanel), id_w3, accf);
        }
      }
      mma_commit(bar);
    }
    acc = 1;
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
  }
  // ---- flush: lane j of warps 0..3 owns output feature j -----------------------------------------------------------------------
  if (t < kHidden) {
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float* w2row = gs + p.lay.w2 + t * kHidden;
#pragma unroll 1
    for (int c = 0; c < kHidden / 16; ++c) {
      float v[16];
      tmem_ld16(lane_base + kColW2 + 16 * c, v);
      store16(w2row + 16 * c, v);
    }
    float v[16];
    tmem_ld16(lane_base + kColW2 + kHidden, v);
    gs[p.lay.b2 + t] = v[0];
    float w[16], w_hi[16];
    tmem_ld16(lane_base + kColW1, w);
    tmem_ld16(lane_base + kColW1 + 16, w_hi);
    for (int i = 0; i < p.lay.in; ++i) gs[p.lay.w1 + t * p.lay.in + i] = i < 16 ? w[i] : w_hi[i - 16];
    gs[p.lay.b1 + t] = D < 16 ? w[D] : w_hi[D - 16];
    tmem_ld16(lane_base + kColW3, v);
    for (int a = 0; a < A; ++a) gs[p.lay.w3 + a * kHidden + t] = v[a];
  }
  if (t < A) gs[p.lay.b3 + t] = db3;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// =====================================================================================================================
// launchers
// =====================================================================================================================
constexpr int kFwdTrainSmem = kImageBytes + 64 + (kTileRows * kOutPad + 16) * 4 + 1024;
constexpr int kDh1Smem = kBwdImageBytes + 64 + 1024;

int tc_train_init() {
  MARL_CUDA_TRY(cudaFuncSetAttribute(tc_dqn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdTrainSmem));
  MARL_CUDA_TRY(cudaFuncSetAttribute(tc_dh1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDh1Smem));
  MARL_CUDA_TRY(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes));
  return MARL_OK;
}

// all three kernels walk the same episode-aligned row split, so the per-CTA partials line up with ReduceParams::cta_begin
int launch_tc_dqn_train(const TrainParams& tp, const TcBuffers& buf, cudaStream_t st) {
  MARL_REQUIRE(tp.lay.in < kMaxObsDim, "tensor-core backward: observation width %d needs a spare column for the bias trick (max %d)", tp.lay.in, kMaxObsDim - 1);
  TcTrainParams p; memset(&p, 0, sizeof(p));
  p.plan = tp.plan; p.src = tp.src; p.lay = tp.lay; p.images = buf.image; p.bwd_images = buf.bwd_image; p.q_out = nullptr;
  p.h1g = buf.h1; p.h2g = buf.h2; p.dh2g = buf.dh2; p.dh1g = buf.dh1; p.dqg = buf.dq;
  p.tq = tp.tq; p.td_ext = tp.td_ext; p.gamma = tp.gamma; p.double_q = tp.double_q;
  p.scratch = tp.scratch; p.scratch_pitch = tp.scratch_pitch; p.loss_part = tp.loss_part;
  const int grid = tp.plan.cta_begin[tp.plan.n_nets];
  tc_dqn_fwd_kernel<<<grid, kTcThreads, kFwdTrainSmem, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  tc_dh1_kernel<<<grid, kTcThreads, kDh1Smem, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  tc_dw_kernel<<<grid, kDwThreads, kDwSmemBytes, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  return MARL_OK;
}

}  // namespace marl
