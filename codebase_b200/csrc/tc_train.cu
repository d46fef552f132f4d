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

constexpr int kRowRec = 16;  // floats per row record

struct TcTrainParams {
  RowPlan plan; RowSource src; NetLayout lay;
  const uint8_t* images;      // forward images [n_nets][kImageBytes]
  const uint8_t* bwd_images;  // backward images [n_nets][kBwdImageBytes]
  float* q_out;               // [rows][out] online outputs (optional)
  // H1, H2, dH2, dH1: [32 float4 column chunks][rows][4] -- chunk-major, so that a warp whose lanes are 32 consecutive rows writes or
  // reads 512 contiguous bytes per instruction (row-major rows of 512 B cost one cache line per lane and instruction)
  float* h1g; float* h2g; float* dh2g; float* dh1g; size_t rows;
  float* dqg;  // [rows][kRowRec] row records: dq[8] | ReLU mask of H1 (4 words) | observation offset (int64) | pad
  const float* tq; const float* td_ext; float gamma; int double_q;
  float* scratch; int scratch_pitch; float* loss_part;
  int debug;
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
// Thread layout of the three kernels: 16 warps.  Warp w works on TMEM lane quarter lq = w & 3 (the hardware restricts a warp to
// lanes 32 (w % 4) .. +31) and on column quarter cq = w >> 2 of the 128 hidden features, so every SM sub-partition holds four
// warps whose TMEM / global latencies overlap (with one warp per sub-partition the kernels sat at 15 % issue utilisation).
// =====================================================================================================================

__device__ __forceinline__ void split16(const float (&h)[16], float (&hi)[16], float (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) { hi[j] = tf32_rn(h[j]); lo[j] = tf32_rn(h[j] - hi[j]); }
}

// =====================================================================================================================
// 1. online forward + TD head + dH2
// =====================================================================================================================
__global__ void __launch_bounds__(kTrThreads, 1) tc_dqn_fwd_kernel(TcTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kImageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* qs = reinterpret_cast<float*>(smem + kImageBytes + 64);   // [128][8] outputs of this tile, then loss reduction scratch
  float* carry = qs + kTileRows * kOutPad;                         // [8] outputs of the first row of the previously processed (higher) tile
  float* gsm = carry + kOutPad;                                    // [128] d loss / d q[act] of each row
  int* actsm = reinterpret_cast<int*>(gsm + kTileRows);            // [128] action of each row
  const int t = threadIdx.x, warp = t >> 5, lq = warp & 3, cq = warp >> 2, r = 32 * lq + (t & 31), c0 = 32 * cq;
  int net, row_begin, row_end;
  cta_rows(p.plan, net, row_begin, row_end);
  float st[2] = {0.f, 0.f};
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
    // three cp.async groups in the order the first tile needs them: W1 + biases + W3 copy, W2, W3
    auto copy = [&](int begin, int end) {
      for (int i = begin / 16 + t; i < end / 16; i += kTrThreads)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * (size_t)i) : "memory");
    };
    copy(kOffW1Hi, kOffW2Hi); copy(kOffB1, kImageBytes);
    asm volatile("cp.async.commit_group;" ::: "memory");
    copy(kOffW2Hi, kOffW3Hi);
    asm volatile("cp.async.commit_group;" ::: "memory");
    copy(kOffW3Hi, kOffB1);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, smem_base = smem_u32(smem), lane_base = tmem + ((uint32_t)(32 * lq) << 16);
  int image_groups_pending = 3;   // block-uniform: groups not yet waited for
  const float* b1 = reinterpret_cast<const float*>(smem + kOffB1);
  const float* b2 = reinterpret_cast<const float*>(smem + kOffB2);
  const float* b3 = reinterpret_cast<const float*>(smem + kOffB3);
  const float* w3f = reinterpret_cast<const float*>(smem + kOffW3F);
  const int D = p.src.D, A = p.lay.out, T = p.src.traj.T, B = p.plan.units_per_agent;
  const int k1steps = (D + 7) >> 3;
  const bool x_active = 8 * cq < 8 * k1steps;   // column quarter cq stages observation columns [8 cq, 8 cq + 8)
  uint32_t parity = 0;

  // this thread's row of a tile: destination row, its 8 observation columns, and (column quarter 0) the loss-head scalars;
  // fetched one tile ahead
  struct RowIn { size_t dst; long long xoff; int agent, b, tt, act; float rew; uint8_t filled, done1; float x[8]; };
  const float* obs_base = p.src.mode == 0 ? p.src.dense : p.src.traj.obs;
  auto fetch = [&](int vr0, int nrows, RowIn& ri) {
    ri.dst = 0; ri.xoff = 0; ri.agent = 0; ri.b = 0; ri.tt = 0; ri.act = 0; ri.rew = 0.f; ri.filled = 0; ri.done1 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) ri.x[j] = 0.f;
    if (r < nrows) {
      ri.dst = dst_of(p.plan, p.src, net, vr0 + r, ri.agent, ri.b, ri.tt);
      const float* src = row_ptr(p.src, ri.agent, ri.b, ri.tt);
      ri.xoff = (long long)(src - obs_base);
      if (x_active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ri.x[j] = (8 * cq + j < D) ? src[8 * cq + j] : 0.f;
      }
      if (cq == 0 && ri.tt < T) {
        const TrajView& tv = p.src.traj;
        const size_t ep = (size_t)p.src.idx[ri.b];
        ri.act = tv.act[(ep * tv.N + ri.agent) * T + ri.tt];
        ri.rew = tv.rew[(ep * tv.N + ri.agent) * T + ri.tt];
        ri.filled = tv.filled[ep * T + ri.tt];   // raw bytes: converting here would wait for the loads inside the prefetch
        ri.done1 = tv.done[ep * (T + 1) + ri.tt + 1];
      }
    }
  };
  RowIn cur, nxt;
  fetch(max(row_begin, row_end - kTileRows), row_end - max(row_begin, row_end - kTileRows), cur);

  // tiles from the top of the chunk downwards (the double-Q argmax needs the next row's outputs)
  for (int vr_hi = row_end; vr_hi > row_begin; vr_hi -= kTileRows) {
    const int vr0 = max(row_begin, vr_hi - kTileRows), nrows = vr_hi - vr0;
    if (x_active) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { hi[j] = tf32_rn(cur.x[j]); lo[j] = tf32_rn(cur.x[j] - hi[j]); }
      tmem_st8(lane_base + kColAHi + 8 * cq, hi);
      tmem_st8(lane_base + kColALo + 8 * cq, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (image_groups_pending == 3) { asm volatile("cp.async.wait_group 2;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); image_groups_pending = 2; }
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      if (!(p.debug & 4)) issue_kmajor<kMaxObsDim / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW1Hi, smem_base + kOffW1Lo, k1steps);
      mma_commit(bar);
    }
    const size_t dst_row = cur.dst;
    const int agent = cur.agent, b = cur.b, tt = cur.tt, act = cur.act;
    const float rew = cur.rew; const uint8_t filled_u8 = cur.filled, done1_u8 = cur.done1; const long long xoff = cur.xoff;
    // the next (lower) tile's rows: the loads stay in flight under this tile's MMAs and epilogues
    if (vr0 > row_begin) { const int nv0 = max(row_begin, vr0 - kTileRows); fetch(nv0, vr0 - nv0, nxt); }
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
    uint32_t h2mask = 0;
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const float* bias = (layer == 0 ? b1 : b2) + c0;
      float4* hg = reinterpret_cast<float4*>(layer == 0 ? p.h1g : p.h2g) + dst_row;
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(lane_base + kColD + c0, ra);
      tmem_ld16_issue(lane_base + kColD + c0 + 16, rb);
      tmem_ld_wait(ra);
      tmem_ld_wait(rb);
      uint32_t mask = 0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t (&acc)[16] = half ? rb : ra;
        float h[16], hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          h[j] = fmaxf(__uint_as_float(acc[j]) + bias[16 * half + j], 0.f);
          mask |= (h[j] > 0.f ? 1u : 0u) << (16 * half + j);
        }
        split16(h, hi, lo);
        tmem_st16(lane_base + kColAHi + c0 + 16 * half, hi);
        tmem_st16(lane_base + kColALo + c0 + 16 * half, lo);
        if (r < nrows && !(p.debug & 1)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) hg[(size_t)(8 * cq + 4 * half + j) * p.rows] = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        }
      }
      h2mask = mask;
      if (image_groups_pending == 2 - layer) {   // W2 before the layer-2 MMAs, W3 before the head's
        if (layer == 0) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        image_groups_pending = 1 - layer;
      }
      if (layer == 0 && r < nrows) reinterpret_cast<uint32_t*>(p.dqg + dst_row * kRowRec)[8 + cq] = mask;   // ReLU mask of H1 for tc_dh1_kernel
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncthreads();
      if (t == 0) {
        tc_fence_after();
        if (p.debug & 4) {}
        else if (layer == 0) issue_kmajor<kHidden / 8, kHidden, kPanelBytes>(tmem, kColD, smem_base + kOffW2Hi, smem_base + kOffW2Lo, kHidden / 8);
        else issue_kmajor<kHidden / 8, kHeadRows, kHeadPanelBytes>(tmem, kColDHead, smem_base + kOffW3Hi, smem_base + kOffW3Lo, kHidden / 8);
        mma_commit(bar);
      }
      mbar_wait(bar, parity); parity ^= 1;
      tc_fence_after();
    }
    // ---- outputs of this tile -> shared (next-row exchange), TD head: column quarter 0 (threads 0..127, r == t) ------------
    float q[kOutPad];
    if (cq == 0) {
      float v[16];
      tmem_ld16(lane_base + kColDHead, v);
#pragma unroll
      for (int o = 0; o < kOutPad; ++o) q[o] = o < A ? v[o] + b3[o] : 0.f;
      *reinterpret_cast<float4*>(qs + r * kOutPad) = make_float4(q[0], q[1], q[2], q[3]);
      *reinterpret_cast<float4*>(qs + r * kOutPad + 4) = make_float4(q[4], q[5], q[6], q[7]);
    }
    __syncthreads();
    if (cq == 0) {
      float g = 0.f;
      if (r < nrows && !(p.debug & 8)) {
        if (p.q_out) for (int o = 0; o < A; ++o) p.q_out[dst_row * A + o] = q[o];
        if (tt < T) {
          if (p.td_ext) {
            g = p.td_ext[(size_t)b * T + tt];
          } else {
            const float* qn = (r + 1 < nrows) ? (qs + (r + 1) * kOutPad) : carry;
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
            const float filled = (float)filled_u8, done1 = (float)done1_u8;
            const float y = rew + p.gamma * tsel * (1.f - done1);
            float qa = q[0];
#pragma unroll
            for (int o = 1; o < kOutPad; ++o) qa = (o == act) ? q[o] : qa;
            const float delta = qa - y;
            st[0] += delta * delta * filled;
            if (agent == 0) st[1] += filled;
            g = 2.f * delta * filled;
          }
        }
        float dq[kOutPad];
#pragma unroll
        for (int o = 0; o < kOutPad; ++o) dq[o] = (o == act && tt < T) ? g : 0.f;
        float* dqd = p.dqg + dst_row * kRowRec;
        *reinterpret_cast<float4*>(dqd) = make_float4(dq[0], dq[1], dq[2], dq[3]);
        *reinterpret_cast<float4*>(dqd + 4) = make_float4(dq[4], dq[5], dq[6], dq[7]);
        *reinterpret_cast<long long*>(dqd + 12) = xoff;
      }
      gsm[r] = g; actsm[r] = act;
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int o = 0; o < kOutPad; ++o) carry[o] = q[o];
    }
    // ---- dH2[r][j] = dq[r][act] W3[act][j] (H2[r][j] > 0): dq has one non-zero per row, the ReLU mask is still in registers ----
    if (r < nrows && !(p.debug & 2)) {
      const float g = gsm[r];
      const float4* wrow = reinterpret_cast<const float4*>(w3f + actsm[r] * kHidden + c0);
      float4* dst = reinterpret_cast<float4*>(p.dh2g) + dst_row;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float4 w = wrow[jj];
        float4 d;
        d.x = (h2mask >> (4 * jj)) & 1u ? g * w.x : 0.f;
        d.y = (h2mask >> (4 * jj + 1)) & 1u ? g * w.y : 0.f;
        d.z = (h2mask >> (4 * jj + 2)) & 1u ? g * w.z : 0.f;
        d.w = (h2mask >> (4 * jj + 3)) & 1u ? g * w.w : 0.f;
        dst[(size_t)(8 * cq + jj) * p.rows] = d;
      }
    }
    cur = nxt;
  }
  // ---- per-CTA loss statistics (threads 0..127 hold them) ---------------------------------------------------------------
  __syncthreads();
  if (t < kTileRows) { qs[t] = st[0]; qs[kTileRows + t] = st[1]; }
  __syncthreads();
  for (int s = kTileRows / 2; s > 0; s >>= 1) {
    if (t < s) { qs[t] += qs[t + s]; qs[kTileRows + t] += qs[kTileRows + t + s]; }
    __syncthreads();
  }
  if (t < 4) p.loss_part[4 * blockIdx.x + t] = t < 2 ? qs[t * kTileRows] : 0.f;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// =====================================================================================================================
// 2. dH1 = (dH2 x W2) * relu'(H1)
// =====================================================================================================================
__global__ void __launch_bounds__(kTrThreads, 1) tc_dh1_kernel(TcTrainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBwdImageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lq = warp & 3, cq = warp >> 2, r = 32 * lq + (t & 31), c0 = 32 * cq;
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
    for (int i = t; i < kBwdImageBytes / 16; i += kTrThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16 * (size_t)i) : "memory");
  }
  // this thread's 32 columns of a row (8 float4), fetched one tile ahead; the weight image lands meanwhile
  auto row_of = [&](int vr0) -> long long {
    long long d = -1;
    if (vr0 + r < row_end && r < kTileRows) { int a, u, o; d = (long long)dst_of(p.plan, p.src, net, vr0 + r, a, u, o); }
    return d;
  };
  auto load8 = [&](const float* base, long long d, float4 (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = d >= 0 ? reinterpret_cast<const float4*>(base)[(size_t)(8 * cq + j) * p.rows + d] : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  long long d_cur = row_of(row_begin), d_nxt = -1;
  float4 in_cur[8], in_nxt[8];
  load8(p.dh2g, d_cur, in_cur);
  asm volatile("cp.async.wait_all;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, smem_base = smem_u32(smem), lane_base = tmem + ((uint32_t)(32 * lq) << 16);
  const uint32_t idesc = idesc_tf32_major(kHidden, 0, 1);
  uint32_t parity = 0;
  for (int vr0 = row_begin; vr0 < row_end; vr0 += kTileRows) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[16], hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = in_cur[4 * half + j];
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
      }
      split16(v, hi, lo);
      tmem_st16(lane_base + kColAHi + c0 + 16 * half, hi);
      tmem_st16(lane_base + kColALo + c0 + 16 * half, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      // D[r][j1] = sum_{j2} dH2[r][j2] W2[j2][j1]: k-step = 8 rows of the [k = j2][n = j1] image (1024 bytes), panels 16 KB apart
      if (!(p.debug & 64)) {
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int ks = 0; ks < kHidden / 8; ++ks)
          mma_tf32_ts(tmem + kColD, tmem + (term == 0 ? kColALo : kColAHi) + ks * 8,
                      mnmajor_desc(smem_base + (term == 1 ? 4 * kPanelBytes : 0) + ks * 1024, kPanelBytes), idesc, (term | ks) ? 1u : 0u);
      }
      mma_commit(bar);
    }
    // H1's ReLU mask (one word per thread, written by tc_dqn_fwd_kernel) and the next tile's dH2 are fetched while the MMAs run
    const uint32_t m1 = d_cur >= 0 ? reinterpret_cast<const uint32_t*>(p.dqg + d_cur * kRowRec)[8 + cq] : 0u;
    if (vr0 + kTileRows < row_end) { d_nxt = row_of(vr0 + kTileRows); load8(p.dh2g, d_nxt, in_nxt); }
    mbar_wait(bar, parity); parity ^= 1;
    tc_fence_after();
    {
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(lane_base + kColD + c0, ra);
      tmem_ld16_issue(lane_base + kColD + c0 + 16, rb);
      tmem_ld_wait(ra);
      tmem_ld_wait(rb);
      if (d_cur >= 0) {
        float4* gout = reinterpret_cast<float4*>(p.dh1g) + d_cur;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t (&acc)[16] = j < 4 ? ra : rb;
          const int o = 4 * (j & 3);
          float4 d;
          d.x = (m1 >> (4 * j)) & 1u ? __uint_as_float(acc[o]) : 0.f; d.y = (m1 >> (4 * j + 1)) & 1u ? __uint_as_float(acc[o + 1]) : 0.f;
          d.z = (m1 >> (4 * j + 2)) & 1u ? __uint_as_float(acc[o + 2]) : 0.f; d.w = (m1 >> (4 * j + 3)) & 1u ? __uint_as_float(acc[o + 3]) : 0.f;
          gout[(size_t)(8 * cq + j) * p.rows] = d;
        }
      }
    }
    d_cur = d_nxt;
#pragma unroll
    for (int j = 0; j < 8; ++j) in_cur[j] = in_nxt[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// =====================================================================================================================
// 3. weight gradients: row-streaming TN GEMMs, accumulators in TMEM
// =====================================================================================================================
constexpr int kDwThreads = 512;
constexpr int kChunkRows = 16;                       // rows per staged chunk (two 8-row k-steps); two chunk buffers alternate
constexpr int kChunkPanel = kChunkRows * 128;        // one 32-feature panel of a chunk
constexpr int kOpBytes = 4 * kChunkPanel;            // a [16 rows][128 features] operand (hi or lo)
// shared-memory map of one chunk buffer (bytes): dH2 hi|lo, H1 hi (+ ones panel) | lo (+ zero panel), dH1 hi|lo, H2 hi|lo, X hi|lo, dq hi|lo
constexpr int kSDh2 = 0, kSH1 = kSDh2 + 2 * kOpBytes, kSDh1 = kSH1 + 2 * (kOpBytes + kChunkPanel), kSH2 = kSDh1 + 2 * kOpBytes;
constexpr int kSX = kSH2 + 2 * kOpBytes, kSDq = kSX + 2 * kChunkPanel, kSEnd = kSDq + 2 * kChunkPanel;
constexpr int kDwSmemBytes = 2 * kSEnd + 64 + 1024;  // two buffers + barriers / TMEM slot + alignment slack
static_assert(kSEnd % 1024 == 0, "chunk buffers must keep the 1024-byte swizzle alignment");
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
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kSEnd);   // [0], [1]: chunk buffer consumed; [2]: everything done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  // staging: warp = (4-row group, 32-feature panel), lane = (float4 column within the panel, row within the group): shared-memory
  // stores of the swizzled MN-major image stay conflict-free and every global load instruction reads eight 64-byte segments
  const int t = threadIdx.x, warp = t >> 5, rr = 4 * (warp >> 2) + (t & 3), c4 = 8 * (warp & 3) + ((t & 31) >> 2);
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
  if (t == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); }
  // constant panels of both buffers: ones column (n = 128) behind H1 hi, zeros behind H1 lo
  for (int i = t; i < 2 * (kChunkPanel / 4); i += kDwThreads) {
    uint8_t* bufp = smem + (i / (kChunkPanel / 4)) * kSEnd;
    const int w = i % (kChunkPanel / 4);
    reinterpret_cast<float*>(bufp + kSH1 + kOpBytes)[w] = 0.f;
    reinterpret_cast<float*>(bufp + kSH1 + 2 * kOpBytes + kChunkPanel)[w] = 0.f;
  }
  __syncthreads();
  if (t < 2 * kChunkRows) *reinterpret_cast<float*>(smem + (t / kChunkRows) * kSEnd + kSH1 + kOpBytes + mn_offset(t % kChunkRows, 0, kChunkPanel)) = 1.0f;
  const int D = p.src.D, A = p.lay.out;
  const int n_chunks = (row_end - row_begin + kChunkRows - 1) / kChunkRows, rpa = p.plan.units_per_agent * p.plan.unit_rows;
  const float* obs_base = p.src.mode == 0 ? p.src.dense : p.src.traj.obs;
  // a chunk's data in registers: this thread's float4 of each [16][128] operand and its element of the [16][32] X / dq panels
  struct Pre { float4 v[4]; float xv, gv; };
  auto issue_loads = [&](int chunk, Pre& pre) {
    const int vr = row_begin + chunk * kChunkRows + rr;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    pre.v[0] = z; pre.v[1] = z; pre.v[2] = z; pre.v[3] = z; pre.xv = 0.f; pre.gv = 0.f;
    if (vr < row_end) {
      size_t d;
      if (p.src.mode == 0) { int a, u, o; d = dst_of(p.plan, p.src, net, vr, a, u, o); }
      else { const int slot = vr / rpa; d = (size_t)p.plan.slot_agent[p.plan.slot_begin[net] + slot] * rpa + (vr - slot * rpa); }  // dst_of(), one division
      const float* rec = p.dqg + d * kRowRec;
      const float* xp = obs_base + *reinterpret_cast<const long long*>(rec + 12);
      const size_t hi = (size_t)c4 * p.rows + d;
      pre.v[0] = reinterpret_cast<const float4*>(p.dh2g)[hi];
      pre.v[1] = reinterpret_cast<const float4*>(p.h1g)[hi];
      pre.v[2] = reinterpret_cast<const float4*>(p.dh1g)[hi];
      pre.v[3] = reinterpret_cast<const float4*>(p.h2g)[hi];
      pre.xv = c4 < D ? xp[c4] : (c4 == D ? 1.f : 0.f);   // [X | 1]: the ones column carries db1
      if (c4 < kOutPad) pre.gv = rec[c4];
    }
  };
  float db3 = 0.f;  // this thread's share of sum_r dq[r][c4]
  auto stage = [&](uint8_t* bufp, const Pre& pre) {
    if (p.debug & 32) { db3 += pre.v[0].x + pre.v[1].x + pre.v[2].x + pre.v[3].x + pre.xv + pre.gv; return; }
    stage4(bufp + kSDh2, bufp + kSDh2 + kOpBytes, rr, 4 * c4, pre.v[0]);
    stage4(bufp + kSH1, bufp + kSH1 + kOpBytes + kChunkPanel, rr, 4 * c4, pre.v[1]);
    stage4(bufp + kSDh1, bufp + kSDh1 + kOpBytes, rr, 4 * c4, pre.v[2]);
    stage4(bufp + kSH2, bufp + kSH2 + kOpBytes, rr, 4 * c4, pre.v[3]);
    const float xh = tf32_rn(pre.xv), gh = tf32_rn(pre.gv);
    *reinterpret_cast<float*>(bufp + kSX + mn_offset(rr, c4, kChunkPanel)) = xh;
    *reinterpret_cast<float*>(bufp + kSX + kChunkPanel + mn_offset(rr, c4, kChunkPanel)) = tf32_rn(pre.xv - xh);
    *reinterpret_cast<float*>(bufp + kSDq + mn_offset(rr, c4, kChunkPanel)) = gh;
    *reinterpret_cast<float*>(bufp + kSDq + kChunkPanel + mn_offset(rr, c4, kChunkPanel)) = tf32_rn(pre.gv - gh);
    db3 += pre.gv;
  };
  Pre pre0, pre1;
  issue_loads(0, pre0);
  issue_loads(1, pre1);   // past the end: zeros, no loads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot, sb = smem_u32(smem);
  const uint32_t id_w2 = idesc_tf32_major(160, 1, 1), id_w1 = idesc_tf32_major(32, 1, 1), id_w3 = idesc_tf32_major(16, 1, 1);
  // one thread: the 18 MMAs of a chunk (3xTF32 terms x 2 k-steps x 3 GEMMs), then a commit onto the buffer's barrier
  auto issue_mmas = [&](uint32_t base, bool first, uint64_t* bar) {
    tc_fence_after();
    if (p.debug & 16) { mma_commit(bar); return; }
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t a_sel = term == 0 ? 1u : 0u, b_sel = term == 1 ? 1u : 0u;  // lo*hi, hi*lo, hi*hi
#pragma unroll
      for (int ks = 0; ks < kChunkRows / 8; ++ks) {
        const uint32_t ko = ks * 1024;
        const uint32_t accf = (!first || term || ks) ? 1u : 0u;
        // dW2[j2][j1 | 1] += dH2^T x [H1 | 1]
        mma_tf32_ss(tmem + kColW2, mnmajor_desc(base + kSDh2 + a_sel * kOpBytes + ko, kChunkPanel),
                    mnmajor_desc(base + kSH1 + b_sel * (kOpBytes + kChunkPanel) + ko, kChunkPanel), id_w2, accf);
        // dW1[j1][i | 1] += dH1^T x [X | 1]
        mma_tf32_ss(tmem + kColW1, mnmajor_desc(base + kSDh1 + a_sel * kOpBytes + ko, kChunkPanel),
                    mnmajor_desc(base + kSX + b_sel * kChunkPanel + ko, kChunkPanel), id_w1, accf);
        // dW3^T[j][a] += H2^T x dq
        mma_tf32_ss(tmem + kColW3, mnmajor_desc(base + kSH2 + a_sel * kOpBytes + ko, kChunkPanel),
                    mnmajor_desc(base + kSDq + b_sel * kChunkPanel + ko, kChunkPanel), id_w3, accf);
      }
    }
    mma_commit(bar);
  };
  uint32_t ph0 = 0, ph1 = 0;
  for (int c = 0; c < n_chunks; c += 2) {
    // even chunk -> buffer 0 (its previous user, chunk c - 2, must have been consumed)
    if (c >= 2) { mbar_wait(bars, ph0); ph0 ^= 1; tc_fence_after(); }
    stage(smem, pre0);
    issue_loads(c + 2, pre0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (t == 0) issue_mmas(sb, c == 0, bars);
    if (c + 1 < n_chunks) {
      if (c >= 2) { mbar_wait(bars + 1, ph1); ph1 ^= 1; tc_fence_after(); }
      stage(smem + kSEnd, pre1);
      issue_loads(c + 3, pre1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncthreads();
      if (t == 0) issue_mmas(sb + kSEnd, false, bars + 1);
    }
  }
  if (t == 0) mma_commit(bars + 2);   // completes when every MMA issued above has
  mbar_wait(bars + 2, 0);
  tc_fence_after();
  // db3[a] = sum over the threads whose panel column c4 is a
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // operands are dead
  if (t < kOutPad) red[t] = 0.f;
  __syncthreads();
  if (c4 < kOutPad) atomicAdd(red + c4, db3);   // 64 threads hold partial sums of the 8 dq columns
  __syncthreads();
  if (t < A) gs[p.lay.b3 + t] = red[t];
  // ---- flush: lane j of lane quarter lq owns output feature j; the column quarters share its accumulator columns -----------
  {
    const int lq = warp & 3, cq = warp >> 2, j = 32 * lq + (t & 31);
    const uint32_t lane_base = tmem + ((uint32_t)(32 * lq) << 16);
    float v[16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld16(lane_base + kColW2 + 32 * cq + 16 * half, v);
      store16(gs + p.lay.w2 + j * kHidden + 32 * cq + 16 * half, v);
    }
    if (cq == 0) {
      tmem_ld16(lane_base + kColW2 + kHidden, v);
      gs[p.lay.b2 + j] = v[0];
    } else if (cq == 1) {
      float w_hi[16];
      tmem_ld16(lane_base + kColW1, v);
      tmem_ld16(lane_base + kColW1 + 16, w_hi);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = i < 16 ? v[i & 15] : w_hi[i & 15];
        if (i < D) gs[p.lay.w1 + j * D + i] = x;
        else if (i == D) gs[p.lay.b1 + j] = x;
      }
    } else if (cq == 2) {
      tmem_ld16(lane_base + kColW3, v);
#pragma unroll
      for (int a = 0; a < kOutPad; ++a) if (a < A) gs[p.lay.w3 + a * kHidden + j] = v[a];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// =====================================================================================================================
// launchers
// =====================================================================================================================
constexpr int kFwdTrainSmem = kImageBytes + 64 + (kTileRows * kOutPad + kOutPad + 2 * kTileRows + 8) * 4 + 1024;
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
  p.h1g = buf.h1; p.h2g = buf.h2; p.dh2g = buf.dh2; p.dh1g = buf.dh1; p.dqg = buf.dq; p.rows = buf.rows;
  p.tq = tp.tq; p.td_ext = tp.td_ext; p.gamma = tp.gamma; p.double_q = tp.double_q;
  p.scratch = tp.scratch; p.scratch_pitch = tp.scratch_pitch; p.loss_part = tp.loss_part; p.debug = tc_debug_bits();
  const int grid = tp.plan.cta_begin[tp.plan.n_nets];
  tc_dqn_fwd_kernel<<<grid, kTrThreads, kFwdTrainSmem, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  tc_dh1_kernel<<<grid, kTrThreads, kDh1Smem, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  tc_dw_kernel<<<grid, kDwThreads, kDwSmemBytes, st>>>(p);
  MARL_CUDA_TRY(cudaGetLastError());
  return MARL_OK;
}

}  // namespace marl
