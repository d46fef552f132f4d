// learner.cuh -- kernels shared by the DQN-family and actor-critic learners: gathered MLP forward, fused
// forward+loss-head+backward training pass, deterministic gradient reduction, clip + Adam + target update.
#pragma once
#include "mlp.cuh"

namespace marl {

struct TrajView {  // device view of marl_traj_view
  const float* obs; const int32_t* act; const float* rew; const uint8_t* done; const uint8_t* filled;
  int capacity, N, T, D;
};

// Which rows a launch covers and how CTAs split them.  A "unit" is an indivisible run of rows that must stay inside
// one CTA: one sampled episode (T+1 rows) for training passes, one row for plain inference.
struct RowPlan {
  int n_nets;
  int cta_begin[MARL_MAX_AGENTS + 1];   // CTAs [cta_begin[k], cta_begin[k+1]) work on net k
  int slot_begin[MARL_MAX_AGENTS + 1];  // agents of net k = slot_agent[slot_begin[k] .. slot_begin[k+1])
  int slot_agent[MARL_MAX_AGENTS];
  int unit_rows;                        // rows per unit
  int units_per_agent;                  // B (episodes) or E (envs)
};

struct RowSource {
  int mode;  // 0: dense obs float[E][N][D];  1: gather from the trajectory store through episode indices
  const float* dense; int E, N, D;
  TrajView traj; const int32_t* idx;  // idx[B] ring slots (device)
};

__device__ __forceinline__ void cta_rows(const RowPlan& p, int& net, int& row_begin, int& row_end) {
  net = 0;
  while (net + 1 < p.n_nets && (int)blockIdx.x >= p.cta_begin[net + 1]) ++net;
  const int ncta = p.cta_begin[net + 1] - p.cta_begin[net], c = (int)blockIdx.x - p.cta_begin[net];
  const long long units = (long long)(p.slot_begin[net + 1] - p.slot_begin[net]) * p.units_per_agent;
  row_begin = (int)(units * c / ncta) * p.unit_rows;
  row_end = (int)(units * (c + 1) / ncta) * p.unit_rows;
}

// virtual row of net -> (agent, unit index within agent, offset within unit)
__device__ __forceinline__ void decode_row(const RowPlan& p, int net, int vr, int& agent, int& unit, int& off) {
  const int rpa = p.units_per_agent * p.unit_rows;
  const int slot = vr / rpa, rem = vr - slot * rpa;
  agent = p.slot_agent[p.slot_begin[net] + slot];
  unit = rem / p.unit_rows;
  off = rem - unit * p.unit_rows;
}

__device__ __forceinline__ const float* row_ptr(const RowSource& s, int agent, int unit, int off) {
  if (s.mode == 0) return s.dense + ((size_t)unit * s.N + agent) * s.D;
  const size_t ep = (size_t)s.idx[unit];
  return s.traj.obs + ((ep * s.traj.N + agent) * (size_t)(s.traj.T + 1) + off) * s.traj.D;
}

// Fill the [128][KP] input tile with rows [vr0, vr0 + nrows) of `net`, zero padded in both directions.
template <int KP>
__device__ __forceinline__ void gather_tile(float* X, const RowPlan& p, const RowSource& s, int net, int vr0, int nrows) {
  for (int i = threadIdx.x; i < kTileRows * KP; i += kMlpThreads) {
    const int r = i / KP, k = i - r * KP;
    float v = 0.f;
    if (r < nrows && k < s.D) {
      int agent, unit, off;
      decode_row(p, net, vr0 + r, agent, unit, off);
      v = row_ptr(s, agent, unit, off)[k];
    }
    at1<KP>(X, r, k) = v;
  }
}

struct FwdParams {
  RowPlan plan; RowSource src;
  const float* theta;   // [n_nets][P]
  NetLayout lay;
  float* out;           // mode 0: [E][N][out]; mode 1: [N][B][T+1][out]
};

struct DqnTrainParams {
  RowPlan plan; RowSource src;
  const float* theta; NetLayout lay;
  const float* tq;        // target-net Q-values of every gathered row, [N][B][T+1][out] (from the forward kernel)
  const float* td_ext;    // VDN: precomputed 2*delta*filled per (b, t), [B][T]; NULL for independent learners
  float gamma; int double_q;
  float* scratch;         // [gridDim][pitch] per-CTA gradient sums (un-normalised); pitch = P rounded up to 4 floats
  int scratch_pitch;
  float* loss_part;       // [gridDim][2] = (sum delta^2*filled, sum filled counted on agent 0 only)
};

struct ReduceParams {
  const float* scratch; const float* loss_part; int n_nets; int cta_begin[MARL_MAX_AGENTS + 1]; int P; int scratch_pitch;
  int n_loss_parts;
  float* grad;  // [n_nets*P + 2]: gradient sums, then loss_sum, filled_sum
};

struct AdamParams {
  float* theta; float* theta_tgt; float* m; float* v; const float* grad; int n;  // n = n_nets*P trainable floats
  float lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_clip;  // grad_clip <= 0: off
  int target_mode;  // 0 none, 1 hard copy, 2 polyak
  float tau;
  float* loss_out;  // [2]: mean loss, grad norm
};

// host-side launchers (defined next to the kernels in learner_kernels.cu); return MARL_* codes
int learner_kernels_init(int in_dim);                       // opt in to > 48 KB dynamic shared memory
int launch_mlp_forward(const FwdParams& p, cudaStream_t st);
int launch_dqn_train(const DqnTrainParams& p, cudaStream_t st);
int launch_grad_reduce(const ReduceParams& p, cudaStream_t st);
int launch_adam(const AdamParams& p, cudaStream_t st);

template <int KP>
constexpr size_t forward_smem_bytes() { return sizeof(float) * (WeightSmem<KP>::kFloats + kTileRows * KP + 2 * kTileRows * kHidden + kTileRows * kOutPad); }
template <int KP>
constexpr size_t train_smem_bytes() { return forward_smem_bytes<KP>() + sizeof(float) * 16; }

}  // namespace marl
