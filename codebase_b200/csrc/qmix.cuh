// qmix.cuh -- the QMIX mixing network (marlbase/dqn/model.py:272-340) and its use in QMixNetwork._compute_loss (386-431).
//
//   Q_tot(q, s) = elu(q . |W1(s)| + b1(s)) . |w_final(s)| + V(s)      q: the agents' chosen (or target) Q-values, s: their observations concatenated
//   W1 = Linear(He -> N*E) o ReLU o Linear(S -> He), w_final likewise (-> E), b1 = Linear(S -> E), V = Linear(E -> 1) o ReLU o Linear(S -> E)
//
// The agents' networks stay on the tensor-core training pass: the mixer only replaces VDN's sum, i.e. it turns the agents' Q-values of every sampled
// (episode, step) into a TD error and hands dL/dq_a back through td_ext (per agent).  Work split:
//   qmix_mix_kernel    one thread per (episode b, step t).  Target: double-Q pick per agent at t + 1, target mixer on the state at t + 1.  Online:
//                      mixer on the state at t, delta = Q_tot - (r + gamma (1 - done) Q_tot_target), dL/dq_a -> td[a][b][t], and the back-propagated
//                      values at the OUTPUT of each of the mixer's seven linear layers next to those layers' inputs -> a per-sample record,
//                      stored field-major ([field][sample]: coalesced for this kernel's writes and the next kernel's reads).
//                      The mixer's parameters sit in shared memory (49 KB at N = 2, S = 30); all lanes read the same weight (broadcast).
//   qmix_wgrad_kernel  dW = sum over samples of (output gradient) x (input): 32 x 32 tiles of every layer's [O][I + 1 (bias)] matrix, the samples in
//                      kQmixChunks chunks -> partial sums (every parameter belongs to exactly one tile).
//   qmix_reduce_kernel the chunks in fixed order -> gradient; the filled count next to it (Adam's 1 / filled.sum()).
// The mixer's parameters take the shared Adam step WITHOUT gradient clipping: the reference clips self.critic.parameters() only (dqn/model.py:169-170).
#pragma once
#include "learner.cuh"

namespace marl {

constexpr int kQmixThreads = 128, kQmixEmbedMax = 64, kQmixHypMax = 64, kQmixNEMax = 512, kQmixStateMax = 256, kQmixChunks = 32, kQmixMaxTiles = 512;

struct QmixLayout {   // offsets (floats) into the mixer's flat parameter vector (reference state_dict order) and into a sample's record
  int N, S, E, He, n;
  int w1a, b1a, w1b, b1b, wfa, bfa, wfb, bfb, wb, bb, wva, bva, wvb, bvb;
  int r_x, r_h1, r_h2, r_hv, r_dz1, r_draw1, r_dzf, r_drawf, r_dhb, r_dzv, r_dv, R;
};

inline QmixLayout qmix_layout(int N, int S, int E, int He) {
  QmixLayout L; L.N = N; L.S = S; L.E = E; L.He = He;
  int o = 0;
  L.w1a = o; o += He * S; L.b1a = o; o += He; L.w1b = o; o += N * E * He; L.b1b = o; o += N * E;
  L.wfa = o; o += He * S; L.bfa = o; o += He; L.wfb = o; o += E * He; L.bfb = o; o += E;
  L.wb = o; o += E * S; L.bb = o; o += E; L.wva = o; o += E * S; L.bva = o; o += E; L.wvb = o; o += E; L.bvb = o; o += 1;
  L.n = o;
  int r = 0;
  L.r_x = r; r += S; L.r_h1 = r; r += He; L.r_h2 = r; r += He; L.r_hv = r; r += E;
  L.r_dz1 = r; r += He; L.r_draw1 = r; r += N * E; L.r_dzf = r; r += He; L.r_drawf = r; r += E; L.r_dhb = r; r += E; L.r_dzv = r; r += E; L.r_dv = r; r += 1;
  L.R = r;
  return L;
}

struct QmixTile { int o0, i0, O, I, doff, ioff, woff, boff; };   // a 32 x 32 tile of one linear layer's weight-gradient matrix ([O][I], bias = column I)

struct QmixParams {
  QmixLayout L;
  const float* q; const float* tq;   // [N][B][T+1][A] online / target Q-values of every gathered row
  TrajView traj; const int32_t* idx; int B, A, D; float gamma; int double_q;
  const float* mix; const float* mix_tgt;
  float* rec;         // [R][B*T]
  float* td;          // [N][B][T] = dL/dq_a (un-normalised: x 2 delta filled)
  float* loss_part;   // [gridDim][4]
};

// rows of `w` ([O][I], I arbitrary) times the thread's x[I]: four outputs per pass so that a load of x feeds four FMAs
__device__ __forceinline__ void qmix_rows_x(const float* __restrict__ w, const float* __restrict__ bias, const float* x, int O, int I, float* out, bool relu) {
  for (int o = 0; o < O; o += 4) {
    float a0 = bias[o], a1 = bias[o + 1], a2 = bias[o + 2], a3 = bias[o + 3];
    const float* w0 = w + (size_t)o * I;
    for (int i = 0; i < I; ++i) {
      const float xv = x[i];
      a0 = fmaf(w0[i], xv, a0); a1 = fmaf(w0[I + i], xv, a1); a2 = fmaf(w0[2 * I + i], xv, a2); a3 = fmaf(w0[3 * I + i], xv, a3);
    }
    if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
    out[o] = a0; out[o + 1] = a1; out[o + 2] = a2; out[o + 3] = a3;
  }
}
// the same with I a multiple of 4 and 16-byte aligned rows (the hypernetworks' second layers)
__device__ __forceinline__ void qmix_rows_h(const float* __restrict__ w, const float* __restrict__ bias, const float* h, int O, int I, float* out) {
  for (int o = 0; o < O; o += 4) {
    float a0 = bias[o], a1 = bias[o + 1], a2 = bias[o + 2], a3 = bias[o + 3];
    const float4* w0 = reinterpret_cast<const float4*>(w + (size_t)o * I);
    const int I4 = I >> 2;
    for (int i = 0; i < I4; ++i) {
      const float h0 = h[4 * i], h1 = h[4 * i + 1], h2 = h[4 * i + 2], h3 = h[4 * i + 3];
      const float4 u0 = w0[i], u1 = w0[I4 + i], u2 = w0[2 * I4 + i], u3 = w0[3 * I4 + i];
      a0 = fmaf(u0.x, h0, a0); a0 = fmaf(u0.y, h1, a0); a0 = fmaf(u0.z, h2, a0); a0 = fmaf(u0.w, h3, a0);
      a1 = fmaf(u1.x, h0, a1); a1 = fmaf(u1.y, h1, a1); a1 = fmaf(u1.z, h2, a1); a1 = fmaf(u1.w, h3, a1);
      a2 = fmaf(u2.x, h0, a2); a2 = fmaf(u2.y, h1, a2); a2 = fmaf(u2.z, h2, a2); a2 = fmaf(u2.w, h3, a2);
      a3 = fmaf(u3.x, h0, a3); a3 = fmaf(u3.y, h1, a3); a3 = fmaf(u3.z, h2, a3); a3 = fmaf(u3.w, h3, a3);
    }
    out[o] = a0; out[o + 1] = a1; out[o + 2] = a2; out[o + 3] = a3;
  }
}
// dh[I] = sum_o d[o] w[o][I] (transposed product), I a multiple of 4; then the ReLU mask of the layer below (h > 0)
__device__ __forceinline__ void qmix_rows_t(const float* __restrict__ w, const float* d, int O, int I, const float* h, float* out) {
  for (int i = 0; i < I; i += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int o = 0; o < O; ++o) {
      const float dv = d[o];
      const float4 u = *reinterpret_cast<const float4*>(w + (size_t)o * I + i);
      a0 = fmaf(u.x, dv, a0); a1 = fmaf(u.y, dv, a1); a2 = fmaf(u.z, dv, a2); a3 = fmaf(u.w, dv, a3);
    }
    out[i] = h[i] > 0.f ? a0 : 0.f; out[i + 1] = h[i + 1] > 0.f ? a1 : 0.f; out[i + 2] = h[i + 2] > 0.f ? a2 : 0.f; out[i + 3] = h[i + 3] > 0.f ? a3 : 0.f;
  }
}

__device__ __forceinline__ float qmix_sgn(float x) { return (float)(x > 0.f) - (float)(x < 0.f); }   // torch.abs' gradient (0 at 0)

// Mixer forward of one sample with the parameters `w` (shared memory).  Leaves h1, h2, hv, raw1 (W1 before abs), rawf, pre (before elu) for the backward.
__device__ __forceinline__ float qmix_forward(const float* __restrict__ w, const QmixLayout& L, const float* x, const float* qa,
                                              float* h1, float* h2, float* hv, float* raw1, float* rawf, float* pre) {
  qmix_rows_x(w + L.w1a, w + L.b1a, x, L.He, L.S, h1, true);
  qmix_rows_x(w + L.wfa, w + L.bfa, x, L.He, L.S, h2, true);
  qmix_rows_x(w + L.wva, w + L.bva, x, L.E, L.S, hv, true);
  qmix_rows_x(w + L.wb, w + L.bb, x, L.E, L.S, pre, false);
  qmix_rows_h(w + L.w1b, w + L.b1b, h1, L.N * L.E, L.He, raw1);
  qmix_rows_h(w + L.wfb, w + L.bfb, h2, L.E, L.He, rawf);
  float y = w[L.bvb];
  for (int e = 0; e < L.E; ++e) {
    float p = pre[e];
    for (int a = 0; a < L.N; ++a) p = fmaf(qa[a], fabsf(raw1[a * L.E + e]), p);
    pre[e] = p;
    const float hid = p > 0.f ? p : expm1f(p);
    y = fmaf(hid, fabsf(rawf[e]), y);
    y = fmaf(w[L.wvb + e], hv[e], y);
  }
  return y;
}

__global__ void __launch_bounds__(kQmixThreads) qmix_mix_kernel(QmixParams p) {
  extern __shared__ __align__(16) float qw[];
  __shared__ float red[2 * kQmixThreads];
  const QmixLayout& L = p.L;
  const int T = p.traj.T, Sn = p.B * T, s = blockIdx.x * kQmixThreads + threadIdx.x;
  const bool live = s < Sn;
  const int b = live ? s / T : 0, t = live ? s - b * T : 0;
  const size_t ep = (size_t)p.idx[b];
  float x[kQmixStateMax], qa[MARL_MAX_AGENTS], h1[kQmixHypMax], h2[kQmixHypMax], hv[kQmixEmbedMax], pre[kQmixEmbedMax], rawf[kQmixEmbedMax], raw1[kQmixNEMax];
  // ---- target: r + gamma (1 - done) Q_tot'(double-Q picks at t + 1, state at t + 1) ----
  for (int i = threadIdx.x; i < L.n; i += kQmixThreads) qw[i] = p.mix_tgt[i];
  __syncthreads();
  float ytgt = 0.f;
  if (live) {
    for (int a = 0; a < L.N; ++a) {
      const size_t row = ((size_t)a * p.B + b) * (T + 1) + t + 1;
      const float* q1 = p.q + row * p.A; const float* t1 = p.tq + row * p.A;
      if (p.double_q) {
        int best = 0; float bv = q1[0];
        for (int o = 1; o < p.A; ++o) if (q1[o] > bv) { bv = q1[o]; best = o; }
        qa[a] = t1[best];
      } else {
        float m = t1[0];
        for (int o = 1; o < p.A; ++o) m = fmaxf(m, t1[o]);
        qa[a] = m;
      }
      const float* ob = p.traj.obs + ((ep * L.N + a) * (T + 1) + t + 1) * p.D;
      for (int d = 0; d < p.D; ++d) x[a * p.D + d] = ob[d];
    }
    ytgt = qmix_forward(qw, L, x, qa, h1, h2, hv, raw1, rawf, pre);
  }
  __syncthreads();
  // ---- online: Q_tot of the chosen actions, TD error, gradients at every linear layer's output ----
  for (int i = threadIdx.x; i < L.n; i += kQmixThreads) qw[i] = p.mix[i];
  __syncthreads();
  float loss = 0.f, fill = 0.f;
  if (live) {
    for (int a = 0; a < L.N; ++a) {
      const size_t row = ((size_t)a * p.B + b) * (T + 1) + t;
      qa[a] = p.q[row * p.A + p.traj.act[(ep * L.N + a) * T + t]];
      const float* ob = p.traj.obs + ((ep * L.N + a) * (T + 1) + t) * p.D;
      for (int d = 0; d < p.D; ++d) x[a * p.D + d] = ob[d];
    }
    const float y = qmix_forward(qw, L, x, qa, h1, h2, hv, raw1, rawf, pre);
    const float filled = (float)p.traj.filled[ep * T + t];
    const float ret = p.traj.rew[(ep * L.N + 0) * T + t] + p.gamma * ytgt * (1.f - (float)p.traj.done[ep * (T + 1) + t + 1]);
    const float delta = y - ret, dy = 2.f * delta * filled;
    loss = delta * delta * filled; fill = filled;
    float* rc = p.rec + s;
    for (int i = 0; i < L.S; ++i) rc[(size_t)(L.r_x + i) * Sn] = x[i];
    for (int j = 0; j < L.He; ++j) { rc[(size_t)(L.r_h1 + j) * Sn] = h1[j]; rc[(size_t)(L.r_h2 + j) * Sn] = h2[j]; }
    rc[(size_t)L.r_dv * Sn] = dy;
    // per embedding unit: V's hidden layer, w_final, the ELU; dpre overwrites pre, d_rawf overwrites rawf
    for (int e = 0; e < L.E; ++e) {
      const float pe = pre[e], hid = pe > 0.f ? pe : expm1f(pe), rf = rawf[e];
      const float dp = dy * fabsf(rf) * (pe > 0.f ? 1.f : hid + 1.f);
      const float drf = dy * hid * qmix_sgn(rf);
      rc[(size_t)(L.r_hv + e) * Sn] = hv[e];
      rc[(size_t)(L.r_dzv + e) * Sn] = hv[e] > 0.f ? dy * qw[L.wvb + e] : 0.f;
      rc[(size_t)(L.r_drawf + e) * Sn] = drf;
      rc[(size_t)(L.r_dhb + e) * Sn] = dp;
      pre[e] = dp; rawf[e] = drf;
    }
    for (int a = 0; a < L.N; ++a) {
      float dq = 0.f;
      for (int e = 0; e < L.E; ++e) {
        const float r1 = raw1[a * L.E + e], dp = pre[e];
        dq = fmaf(dp, fabsf(r1), dq);
        const float d1 = dp * qa[a] * qmix_sgn(r1);
        raw1[a * L.E + e] = d1;
        rc[(size_t)(L.r_draw1 + a * L.E + e) * Sn] = d1;
      }
      p.td[((size_t)a * p.B + b) * T + t] = dq;
    }
    qmix_rows_t(qw + L.wfb, rawf, L.E, L.He, h2, hv);          // hv, pre: scratch from here on
    for (int j = 0; j < L.He; ++j) rc[(size_t)(L.r_dzf + j) * Sn] = hv[j];
    qmix_rows_t(qw + L.w1b, raw1, L.N * L.E, L.He, h1, hv);
    for (int j = 0; j < L.He; ++j) rc[(size_t)(L.r_dz1 + j) * Sn] = hv[j];
  }
  red[threadIdx.x] = loss; red[kQmixThreads + threadIdx.x] = fill;
  __syncthreads();
  for (int k = kQmixThreads / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) { red[threadIdx.x] += red[threadIdx.x + k]; red[kQmixThreads + threadIdx.x] += red[kQmixThreads + threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { p.loss_part[4 * blockIdx.x] = red[0]; p.loss_part[4 * blockIdx.x + 1] = red[kQmixThreads]; p.loss_part[4 * blockIdx.x + 2] = 0.f; p.loss_part[4 * blockIdx.x + 3] = 0.f; }
}

__global__ void __launch_bounds__(256) qmix_wgrad_kernel(const float* __restrict__ rec, int Sn, const QmixTile* __restrict__ tiles, int chunk_len, float* part, int n) {
  __shared__ float As[32][33], Bs[32][33];
  const QmixTile tl = tiles[blockIdx.x];
  const int chunk = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int s_begin = chunk * chunk_len, s_end = min(Sn, s_begin + chunk_len);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s0 = s_begin; s0 < s_end; s0 += 32) {
    const int s = s0 + tx;
    const bool valid = s < s_end;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int o = tl.o0 + ty + 8 * k, i = tl.i0 + ty + 8 * k;
      As[tx][ty + 8 * k] = (valid && o < tl.O) ? rec[(size_t)(tl.doff + o) * Sn + s] : 0.f;
      Bs[tx][ty + 8 * k] = !valid ? 0.f : i < tl.I ? rec[(size_t)(tl.ioff + i) * Sn + s] : i == tl.I ? 1.f : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int ss = 0; ss < 32; ++ss) {
      const float bv = Bs[ss][tx];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(As[ss][ty + 8 * k], bv, acc[k]);
    }
    __syncthreads();
  }
  float* dst = part + (size_t)chunk * n;
  const int i = tl.i0 + tx;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o = tl.o0 + ty + 8 * k;
    if (o < tl.O) {
      if (i < tl.I) dst[tl.woff + o * tl.I + i] = acc[k];
      else if (i == tl.I) dst[tl.boff + o] = acc[k];
    }
  }
}

// grad[0..n) = the chunks' partial sums in fixed order; grad[n..n+4) = (loss numerator, filled count, 0, 0) of this update
__global__ void __launch_bounds__(256) qmix_reduce_kernel(const float* __restrict__ part, int chunks, int n, float* grad, const float* __restrict__ loss_part, int n_loss_parts) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j < n) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * n + j];
    grad[j] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, l = threadIdx.x & 31;
    float x = 0.f;
    for (int c = l; c < n_loss_parts; c += 32) x += loss_part[4 * c + which];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, off);
    if (l == 0) { grad[n + which] = x; grad[n + 2 + which] = 0.f; }
  }
}

// tiles of the seven linear layers (host side)
inline int qmix_tiles(const QmixLayout& L, QmixTile* out) {
  struct Lay { int O, I, doff, ioff, woff, boff; };
  const Lay lays[7] = {
      {L.He, L.S, L.r_dz1, L.r_x, L.w1a, L.b1a}, {L.N * L.E, L.He, L.r_draw1, L.r_h1, L.w1b, L.b1b}, {L.He, L.S, L.r_dzf, L.r_x, L.wfa, L.bfa},
      {L.E, L.He, L.r_drawf, L.r_h2, L.wfb, L.bfb}, {L.E, L.S, L.r_dhb, L.r_x, L.wb, L.bb}, {L.E, L.S, L.r_dzv, L.r_x, L.wva, L.bva},
      {1, L.E, L.r_dv, L.r_hv, L.wvb, L.bvb}};
  int n = 0;
  for (const Lay& l : lays)
    for (int o0 = 0; o0 < l.O; o0 += 32)
      for (int i0 = 0; i0 <= l.I; i0 += 32) {   // column I is the bias
        if (n == kQmixMaxTiles) return -1;
        out[n++] = QmixTile{o0, i0, l.O, l.I, l.doff, l.ioff, l.woff, l.boff};
      }
  return n;
}

}  // namespace marl
