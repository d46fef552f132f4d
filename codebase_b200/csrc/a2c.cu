// a2c.cu -- C ABI of the independent actor-critic learner (marl_a2c_*), host-side orchestration of the fused kernels.
//
// Replaces marlbase/ac/model.py A2CNetwork (22-246): act's actor forward (147-153), get_value (155-163),
// update (189-246: target-critic pass, compute_nstep_returns utils/utils.py:38-63, evaluate_actions 165-182,
// policy-gradient + entropy + value losses, Adam, target sync), for independent or shared per-agent networks with a
// decentralised critic (critic.centralised: False, configs/algorithm/ia2c.yaml:18).
#include "learner.cuh"
#include "retms.cuh"
#include <math.h>
#include <vector>

namespace marl {

constexpr int kMaxNStep = 64;

struct NStepParams {
  const float* vt;     // [N][P][T+1] target-critic values
  TrajView traj; const int32_t* idx; int N, P, n_steps;
  float gpow[kMaxNStep + 1];  // float32(gamma ** k), the Python-float powers of utils/utils.py:57-60
  float* ret;          // [N][P][T]
  const float* ret_ms; // standardise_returns (ac/model.py:195-196): [mean[N] | var[N]] of the running return statistics, or NULL
};

// compute_nstep_returns (utils/utils.py:38-63): G_t = sum_{k<n} g^k r_{t+k}(1-d_{t+k}) + g^n V(o_{t+n})(1-d_{t+n}), cut (no
// bootstrap) where t+k reaches the end of the stored episode; d_t = dones[t] is the terminal flag of observation t.
__global__ void nstep_returns_kernel(NStepParams p) {
  const int T = p.traj.T, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N * p.P * T) return;
  const int a = i / (p.P * T), rem = i - a * p.P * T, b = rem / T, t = rem - b * T;
  const size_t ep = (size_t)p.idx[b];
  float acc = 0.f;
  for (int k = 0; k <= p.n_steps; ++k) {
    const int tt = t + k;
    if (tt >= T) break;
    const float d = (float)p.traj.done[ep * (T + 1) + tt];
    float src;
    if (k == p.n_steps) {
      src = p.vt[((size_t)a * p.P + b) * (T + 1) + tt];
      if (p.ret_ms) src = __fadd_rn(__fmul_rn(src, sqrtf(p.ret_ms[p.N + a])), p.ret_ms[a]);   // next_value * sqrt(var) + mean
    } else {
      src = p.traj.rew[(ep * p.N + a) * T + tt];
    }
    acc += (p.gpow[k] * src) * (1.f - d);
  }
  p.ret[i] = acc;
}

// log-probabilities of the taken actions under the collecting policy (ac/model.py:281-292), from the actor outputs of every gathered row:
// old_logp[a][b][t] = log_softmax(logits[a][b][t][:])[act[a][b][t]] -- the formulas of head_a2c_actor (learner_kernels.cu)
struct OldLogpParams { const float* logits; TrajView traj; const int32_t* idx; int N, P, A; float* out; };
__global__ void old_logp_kernel(OldLogpParams p) {
  const int T = p.traj.T, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N * p.P * T) return;
  const int a = i / (p.P * T), rem = i - a * p.P * T, b = rem / T, t = rem - b * T;
  const float* q = p.logits + (((size_t)a * p.P + b) * (T + 1) + t) * p.A;
  const int act = p.traj.act[((size_t)p.idx[b] * p.N + a) * T + t];
  float m = q[0];
  for (int o = 1; o < p.A; ++o) m = fmaxf(m, q[o]);
  float s = 0.f;
  for (int o = 0; o < p.A; ++o) s += expf(q[o] - m);
  p.out[i] = q[act] - (m + logf(s));
}

// mean over the epochs of the six device metrics (ac/model.py:352)
__global__ void mean_metrics_kernel(const float* per_epoch, int n_epochs, float* out) {
  const int k = threadIdx.x;
  if (k >= 6) return;
  float s = 0.f;
  for (int e = 0; e < n_epochs; ++e) s += per_epoch[6 * e + k];
  out[k] = k == 4 ? per_epoch[4] : s / (float)n_epochs;   // [4] = filled count (the same every epoch)
}

// centralised critic (ac/model.py:62-65,156-157): the joint observation of every (env, step) = the agents' observations side by side
struct JointParams { TrajView traj; const int32_t* idx; int P; float* out; };   // out: [P][T+1][N * D]
__global__ void joint_obs_kernel(JointParams p) {
  const int T1 = p.traj.T + 1, ND = p.traj.N * p.traj.D;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.P * T1 * ND) return;
  const int c = (int)(i % ND), t = (int)((i / ND) % T1), b = (int)(i / ((size_t)ND * T1));
  const int j = c / p.traj.D, d = c - j * p.traj.D;
  p.out[i] = p.traj.obs[(((size_t)p.idx[b] * p.traj.N + j) * T1 + t) * p.traj.D + d];
}

__global__ void iota_kernel(int32_t* x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = i;
}

}  // namespace marl

using namespace marl;

struct marl_a2c {
  NetSet actor, critic;
  marl_a2c_hp hp;
  int device = 0, n_sm = 148, max_envs = 0, max_T = 0;
  int64_t n_actor = 0, n_critic = 0, n_params = 0;
  int scratch_pitch = 0;
  float *theta = nullptr, *theta_tgt = nullptr, *m = nullptr, *v = nullptr, *grad = nullptr;
  float *scratch = nullptr, *loss_part = nullptr, *vt = nullptr, *ret = nullptr, *adv = nullptr, *metrics = nullptr;
  int32_t* idx = nullptr;
  uint8_t* image = nullptr;  // packed weight images for the tensor-core forward path
  int64_t opt_steps = 0;
  float *logits_all = nullptr, *old_logp = nullptr, *epoch_metrics = nullptr;   // PPO (allocated on first use)
  // standardise_returns: RunningMeanStd(shape=(n_agents,)) -- mean[N] | var[N] (float32), count (a Python float in the reference), partial sums
  int standardise = 0; float* ret_ms = nullptr; double* ret_count = nullptr; double* ret_part = nullptr;
  int centralised = 0; float* joint = nullptr;   // critic.centralised: joint observations of the batch [P][T+1][N * D]
};
constexpr int kMaxPpoEpochs = 64;

extern "C" {

int marl_a2c_destroy(marl_a2c* h) {
  if (!h) return MARL_OK;
  cudaSetDevice(h->device);
  cudaFree(h->theta); cudaFree(h->theta_tgt); cudaFree(h->m); cudaFree(h->v); cudaFree(h->grad); cudaFree(h->scratch); cudaFree(h->loss_part);
  cudaFree(h->vt); cudaFree(h->ret); cudaFree(h->adv); cudaFree(h->metrics); cudaFree(h->idx); cudaFree(h->image);
  cudaFree(h->logits_all); cudaFree(h->old_logp); cudaFree(h->epoch_metrics); cudaFree(h->ret_ms); cudaFree(h->ret_count); cudaFree(h->ret_part); cudaFree(h->joint);
  delete h;
  return MARL_OK;
}

int marl_a2c_create(const marl_mlp_cfg* actor, const marl_mlp_cfg* critic, const marl_a2c_hp* hp, int32_t max_envs, int32_t max_T, int32_t device, marl_a2c** out) {
  MARL_REQUIRE(hp && out, "marl_a2c_create: NULL argument");
  *out = nullptr;
  if (int rc = check_mlp_cfg(actor, "marl_a2c_create(actor)")) return rc;
  if (int rc = check_mlp_cfg(critic, "marl_a2c_create(critic)")) return rc;
  MARL_REQUIRE(actor->n_agents == critic->n_agents && (actor->in_dim == critic->in_dim || critic->in_dim == actor->n_agents * actor->in_dim),
               "marl_a2c_create: the critic's input width must be the actor's (%d) or, for a centralised critic, n_agents x it (%d)", actor->in_dim, actor->n_agents * actor->in_dim);
  MARL_REQUIRE(critic->out_dim == 1, "marl_a2c_create: the critic outputs one state value per agent");
  MARL_REQUIRE(max_envs >= 1 && max_T >= 1, "marl_a2c_create: max_envs/max_T must be >= 1");
  MARL_REQUIRE(hp->n_steps >= 1 && hp->n_steps <= kMaxNStep, "marl_a2c_create: n_steps %d out of range (1..%d)", hp->n_steps, kMaxNStep);
  if (int rc = check_device(device)) return rc;
  marl_a2c* h = new marl_a2c();
  h->actor = to_netset(actor); h->critic = to_netset(critic);
  h->hp = *hp; h->device = device; h->max_envs = max_envs; h->max_T = max_T;
  h->centralised = (critic->in_dim != actor->in_dim || (actor->n_agents == 1 && false)) ? 1 : 0;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device); h->n_sm = prop.multiProcessorCount;
  h->n_actor = (int64_t)actor->n_nets * h->actor.lay.P; h->n_critic = (int64_t)critic->n_nets * h->critic.lay.P; h->n_params = h->n_actor + h->n_critic;
  const int pmax = h->actor.lay.P > h->critic.lay.P ? h->actor.lay.P : h->critic.lay.P;
  h->scratch_pitch = (pmax + 3) & ~3;
  const size_t rows = (size_t)actor->n_agents * max_envs * (max_T + 1);
  int rc = 0;
  rc |= dev_alloc_zero(&h->theta, h->n_params); rc |= dev_alloc_zero(&h->theta_tgt, h->n_critic);
  rc |= dev_alloc_zero(&h->m, h->n_params); rc |= dev_alloc_zero(&h->v, h->n_params); rc |= dev_alloc_zero(&h->grad, h->n_params + 4);
  rc |= dev_alloc_zero(&h->scratch, (size_t)h->n_sm * h->scratch_pitch); rc |= dev_alloc_zero(&h->loss_part, 4 * (size_t)h->n_sm);
  rc |= dev_alloc_zero(&h->vt, rows); rc |= dev_alloc_zero(&h->ret, rows); rc |= dev_alloc_zero(&h->adv, rows); rc |= dev_alloc_zero(&h->metrics, 8);
  rc |= dev_alloc_zero(reinterpret_cast<float**>(&h->idx), max_envs);
  rc |= dev_alloc_zero(reinterpret_cast<float**>(&h->image), (size_t)(actor->n_nets > critic->n_nets ? actor->n_nets : critic->n_nets) * tc_image_bytes() / 4 + 4);
  if (rc) { marl_a2c_destroy(h); return MARL_ENOMEM; }
  iota_kernel<<<(max_envs + 255) / 256, 256>>>(h->idx, max_envs);
  if (h->centralised && dev_alloc_zero(&h->joint, (size_t)max_envs * (max_T + 1) * critic->in_dim)) { marl_a2c_destroy(h); return MARL_ENOMEM; }
  if (int rc2 = learner_kernels_init(actor->in_dim)) { marl_a2c_destroy(h); return rc2; }
  if (int rc2 = learner_kernels_init(critic->in_dim)) { marl_a2c_destroy(h); return rc2; }
  if (int rc2 = tc_forward_init()) { marl_a2c_destroy(h); return rc2; }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("marl_a2c_create: device error during setup"); marl_a2c_destroy(h); return MARL_ECUDA; }
  *out = h;
  return MARL_OK;
}

/* theta = [actor nets | critic nets] (flat, reference state_dict order per net), theta_tgt = target critic. */
int marl_a2c_param_ptrs(marl_a2c* h, float** theta, float** theta_tgt, float** adam_m, float** adam_v, float** grad, int64_t* n_actor, int64_t* n_critic) {
  MARL_REQUIRE(h != nullptr, "marl_a2c_param_ptrs: NULL handle");
  if (theta) *theta = h->theta; if (theta_tgt) *theta_tgt = h->theta_tgt; if (adam_m) *adam_m = h->m; if (adam_v) *adam_v = h->v;
  if (grad) *grad = h->grad; if (n_actor) *n_actor = h->n_actor; if (n_critic) *n_critic = h->n_critic;
  return MARL_OK;
}

int marl_a2c_scratch_ptrs(marl_a2c* h, float** target_values, float** returns, float** advantages) {
  MARL_REQUIRE(h != nullptr, "marl_a2c_scratch_ptrs: NULL handle");
  if (target_values) *target_values = h->vt; if (returns) *returns = h->ret; if (advantages) *advantages = h->adv;
  return MARL_OK;
}

int marl_a2c_sync_target(marl_a2c* h, void* stream) {  // soft_update(1.0), ac/model.py:101
  MARL_REQUIRE(h != nullptr, "marl_a2c_sync_target: NULL handle");
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  MARL_CUDA_TRY(cudaMemcpyAsync(h->theta_tgt, h->theta + h->n_actor, h->n_critic * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MARL_OK;
}

static int a2c_dense_forward(marl_a2c* h, const NetSet& ns, const float* theta, const float* obs, int n_envs, float* out, void* stream) {
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  const RowPlan plan = make_plan(ns, n_envs, 1, h->n_sm, 32);
  RowSource src; memset(&src, 0, sizeof(src));
  src.mode = 0; src.dense = obs; src.E = n_envs; src.N = ns.n_agents; src.D = ns.in;
  if (ns.in != h->actor.in) { src.mode = 3; src.joint = obs; }   // centralised critic: obs float[E][N][D] read as the joint rows float[E][N * D]
  return forward_any(ns, plan, src, theta, h->image, out, (cudaStream_t)stream);
}

/* actor forward of A2CNetwork.act (ac/model.py:148-150): obs float[E][N][in] -> logits float[E][N][n_actions] */
int marl_a2c_forward_actor(marl_a2c* h, const float* obs, int32_t n_envs, float* logits_out, void* stream) {
  MARL_REQUIRE(h && obs && logits_out && n_envs >= 1, "marl_a2c_forward_actor: bad argument");
  return a2c_dense_forward(h, h->actor, h->theta, obs, n_envs, logits_out, stream);
}

/* get_value (ac/model.py:155-163): values float[E][N][1] from the critic or the target critic */
int marl_a2c_forward_critic(marl_a2c* h, const float* obs, int32_t n_envs, int32_t use_target, float* values_out, void* stream) {
  MARL_REQUIRE(h && obs && values_out && n_envs >= 1, "marl_a2c_forward_critic: bad argument");
  return a2c_dense_forward(h, h->critic, use_target ? h->theta_tgt : h->theta + h->n_actor, obs, n_envs, values_out, stream);
}

struct A2cPass { RowSource src, csrc; RowPlan cplan, aplan; };   // csrc: the critic's rows (== src unless the critic is centralised)

// target-critic pass + n-step returns (ac/model.py:190-201): everything of an update that does not depend on the trainable parameters
static int a2c_prepare(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, cudaStream_t st, A2cPass& ps) {
  MARL_REQUIRE(h && batch, "marl_a2c_update: NULL argument");
  MARL_REQUIRE(n_envs >= 1 && n_envs <= h->max_envs && n_envs <= batch->capacity, "marl_a2c_update: n_envs %d out of range", n_envs);
  MARL_REQUIRE(batch->T >= 1 && batch->T <= h->max_T, "marl_a2c_update: T %d exceeds max_T %d", batch->T, h->max_T);
  MARL_REQUIRE(batch->n_agents == h->actor.n_agents && batch->obs_dim == h->actor.in, "marl_a2c_update: batch shape mismatch");
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  const int T = batch->T, N = h->actor.n_agents;
  const int min_units = (64 + T) / (T + 1) > 0 ? (64 + T) / (T + 1) : 1;
  memset(&ps.src, 0, sizeof(ps.src));
  ps.src.mode = 1; ps.src.traj = to_view(batch); ps.src.idx = h->idx; ps.src.N = N; ps.src.D = h->actor.in;
  ps.cplan = make_plan(h->critic, n_envs, T + 1, h->n_sm, min_units);
  ps.aplan = make_plan(h->actor, n_envs, T + 1, h->n_sm, min_units);
  ps.csrc = ps.src;
  if (h->centralised) {   // get_value (ac/model.py:156-157): every agent's critic reads the concatenated observations
    JointParams jp; jp.traj = ps.src.traj; jp.idx = h->idx; jp.P = n_envs; jp.out = h->joint;
    const size_t n = (size_t)n_envs * (T + 1) * h->critic.in;
    joint_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(jp);
    MARL_CUDA_TRY(cudaGetLastError());
    ps.csrc.mode = 2; ps.csrc.joint = h->joint; ps.csrc.D = h->critic.in;
  }
  // 1. target critic on all T+1 observations (ac/model.py:190-193)
  if (int rc = forward_any(h->critic, ps.cplan, ps.csrc, h->theta_tgt, h->image, h->vt, st)) return rc;
  // 2. n-step returns (ac/model.py:198-201)
  NStepParams np; np.vt = h->vt; np.traj = ps.src.traj; np.idx = h->idx; np.N = N; np.P = n_envs; np.n_steps = h->hp.n_steps; np.ret = h->ret;
  np.ret_ms = h->standardise ? h->ret_ms : nullptr;
  for (int k = 0; k <= h->hp.n_steps; ++k) np.gpow[k] = (float)pow((double)h->hp.gamma, (double)k);
  nstep_returns_kernel<<<(N * n_envs * T + 255) / 256, 256, 0, st>>>(np);
  MARL_CUDA_TRY(cudaGetLastError());
  if (h->standardise) {  // ac/model.py:202-204
    RetMsParams rp; rp.ret = h->ret; rp.N = N; rp.P = n_envs; rp.T = T; rp.part = h->ret_part; rp.ret_ms = h->ret_ms; rp.count = h->ret_count;
    MARL_CUDA_TRY(ret_ms_step(rp, st));
  }
  return MARL_OK;
}

/* cfg.standardise_returns (ac/model.py:112-114): switches the RunningMeanStd over the n-step returns on (mean 0, var 1, count 1e-4) or off. */
int marl_a2c_standardise_returns(marl_a2c* h, int32_t enable) {
  MARL_REQUIRE(h != nullptr, "marl_a2c_standardise_returns: NULL handle");
  MARL_REQUIRE(h->actor.n_agents <= 32, "marl_a2c_standardise_returns: at most 32 agents");
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  if (enable && !h->ret_ms) {
    const int N = h->actor.n_agents;
    std::vector<float> init(2 * N, 0.f);
    for (int a = 0; a < N; ++a) init[N + a] = 1.f;
    const double c0 = 1e-4;
    MARL_CUDA_TRY(cudaMalloc(&h->ret_ms, 2 * N * sizeof(float))); MARL_CUDA_TRY(cudaMalloc(&h->ret_count, sizeof(double)));
    MARL_CUDA_TRY(cudaMalloc(&h->ret_part, (size_t)kRetBlocks * N * 2 * sizeof(double)));
    MARL_CUDA_TRY(cudaMemcpy(h->ret_ms, init.data(), 2 * N * sizeof(float), cudaMemcpyHostToDevice));
    MARL_CUDA_TRY(cudaMemcpy(h->ret_count, &c0, sizeof(double), cudaMemcpyHostToDevice));
  }
  h->standardise = enable ? 1 : 0;
  return MARL_OK;
}
/* the running statistics as device pointers: ret_ms float[2 n_agents] = mean | var, count double[1] (NULL before the first enable) */
int marl_a2c_ret_ms_ptrs(marl_a2c* h, float** ret_ms, double** count) {
  MARL_REQUIRE(h != nullptr, "marl_a2c_ret_ms_ptrs: NULL handle");
  if (ret_ms) *ret_ms = h->ret_ms; if (count) *count = h->ret_count;
  return MARL_OK;
}

// critic and actor training passes -> grad[] (un-normalised sums) + loss statistics; old_logp != NULL: PPO's clipped surrogate
static int a2c_gradients(marl_a2c* h, const A2cPass& ps, cudaStream_t st, const float* old_logp, float ppo_clip) {
  // 3. critic: forward, value loss, backward; leaves advantage = returns - V for the actor pass
  TrainParams tp; memset(&tp, 0, sizeof(tp));
  tp.plan = ps.cplan; tp.src = ps.csrc; tp.theta = h->theta + h->n_actor; tp.lay = h->critic.lay; tp.scratch = h->scratch; tp.scratch_pitch = h->scratch_pitch;
  tp.loss_part = h->loss_part; tp.returns = h->ret; tp.adv_out = h->adv; tp.value_coef = h->hp.value_loss_coef;
  if (int rc = launch_train(tp, kHeadA2cCritic, st)) return rc;
  ReduceParams rp; memset(&rp, 0, sizeof(rp));  // (sumsq_part stays NULL: two passes write different gradient slices)
  rp.scratch = h->scratch; rp.loss_part = h->loss_part; rp.n_nets = h->critic.n_nets; rp.P = h->critic.lay.P; rp.scratch_pitch = h->scratch_pitch;
  memcpy(rp.cta_begin, ps.cplan.cta_begin, sizeof(rp.cta_begin));
  rp.n_loss_parts = ps.cplan.cta_begin[ps.cplan.n_nets]; rp.grad = h->grad + h->n_actor; rp.stats = h->grad + h->n_params; rp.stats_accumulate = 0;
  if (int rc = launch_grad_reduce(rp, st)) return rc;
  // 4. actor: forward, log-softmax, policy-gradient (or clipped surrogate) + entropy loss, backward
  tp.plan = ps.aplan; tp.src = ps.src; tp.theta = h->theta; tp.lay = h->actor.lay; tp.adv = h->adv; tp.entropy_coef = h->hp.entropy_coef;
  tp.old_logp = old_logp; tp.ppo_clip = ppo_clip;
  if (int rc = launch_train(tp, kHeadA2cActor, st)) return rc;
  rp.n_nets = h->actor.n_nets; rp.P = h->actor.lay.P; memcpy(rp.cta_begin, ps.aplan.cta_begin, sizeof(rp.cta_begin));
  rp.n_loss_parts = ps.aplan.cta_begin[ps.aplan.n_nets]; rp.grad = h->grad; rp.stats_accumulate = 1;
  return launch_grad_reduce(rp, st);
}

int marl_a2c_update_grads(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, void* stream) {
  A2cPass ps;
  if (int rc = a2c_prepare(h, batch, n_envs, (cudaStream_t)stream, ps)) return rc;
  return a2c_gradients(h, ps, (cudaStream_t)stream, nullptr, 0.f);
}

/* metrics_out device float[6] = (policy-gradient term, grad norm, entropy, value_loss, filled count, 0):
 * actor_loss = m[0] - entropy_coef*m[2]; loss = actor_loss + value_loss_coef*m[3]  (ac/model.py:216-226,241-246) */
static int a2c_apply(marl_a2c* h, int64_t step, float* metrics_out, void* stream, bool target_update) {
  MARL_REQUIRE(h != nullptr, "marl_a2c_update_apply: NULL handle");
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  h->opt_steps += 1;
  AdamParams ap; memset(&ap, 0, sizeof(ap));
  ap.theta = h->theta; ap.theta_tgt = h->theta_tgt; ap.m = h->m; ap.v = h->v; ap.grad = h->grad; ap.n = (int)h->n_params;
  ap.tgt_begin = (int)h->n_actor; ap.tgt_n = (int)h->n_critic;
  ap.lr = h->hp.lr; ap.beta1 = h->hp.beta1; ap.beta2 = h->hp.beta2; ap.eps = h->hp.eps; ap.grad_clip = h->hp.grad_clip;
  ap.bc1 = (float)(1.0 - pow((double)h->hp.beta1, (double)h->opt_steps));
  ap.bc2_sqrt = (float)sqrt(1.0 - pow((double)h->hp.beta2, (double)h->opt_steps));
  const float tu = h->hp.target_update_interval_or_tau;  // ac/model.py:233-239: `step` counts environment steps
  ap.tau = tu;
  if (target_update && tu > 1.0f && fmod((double)step, (double)tu) == 0.0) ap.target_mode = 1;
  else if (target_update && tu < 1.0f) ap.target_mode = 2;
  ap.loss_out = metrics_out ? metrics_out : h->metrics;
  return launch_adam(ap, (cudaStream_t)stream);
}

int marl_a2c_update_apply(marl_a2c* h, int64_t step, float* metrics_out, void* stream) { return a2c_apply(h, step, metrics_out, stream, true); }

/* PPONetwork.update (ac/model.py:265-352) on the same handle: returns and the collecting policy's log-probabilities once, then `num_epochs`
 * optimisation steps on the same batch with the clipped surrogate (-min(r adv, clip(r, 1 -+ ppo_clip) adv) - entropy_coef H + value_loss_coef
 * value loss; clip_grad_norm_ over all parameters when hp.grad_clip > 0), the target critic synchronised after the last epoch.
 * metrics_out: device float[6] as marl_a2c_update, averaged over the epochs ([0] = the surrogate term). */
int marl_ppo_update(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, int64_t step, int32_t num_epochs, float ppo_clip, float* metrics_out, void* stream) {
  MARL_REQUIRE(h != nullptr && num_epochs >= 1 && num_epochs <= kMaxPpoEpochs, "marl_ppo_update: num_epochs %d out of range (1..%d)", (int)num_epochs, kMaxPpoEpochs);
  MARL_REQUIRE(ppo_clip > 0.f, "marl_ppo_update: ppo_clip must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  A2cPass ps;
  if (int rc = a2c_prepare(h, batch, n_envs, st, ps)) return rc;
  if (!h->logits_all) {
    const size_t rows = (size_t)h->actor.n_agents * h->max_envs * (h->max_T + 1);
    int rc = dev_alloc_zero(&h->logits_all, rows * h->actor.out) | dev_alloc_zero(&h->old_logp, rows) | dev_alloc_zero(&h->epoch_metrics, 6 * kMaxPpoEpochs);
    if (rc) return MARL_ENOMEM;
  }
  // log-probabilities of the taken actions under the collecting policy = the current actor (ac/model.py:281-292)
  if (int rc = forward_any(h->actor, ps.aplan, ps.src, h->theta, h->image, h->logits_all, st)) return rc;
  OldLogpParams op; op.logits = h->logits_all; op.traj = ps.src.traj; op.idx = h->idx; op.N = h->actor.n_agents; op.P = n_envs; op.A = h->actor.out; op.out = h->old_logp;
  old_logp_kernel<<<(op.N * n_envs * batch->T + 255) / 256, 256, 0, st>>>(op);
  MARL_CUDA_TRY(cudaGetLastError());
  for (int e = 0; e < num_epochs; ++e) {
    if (int rc = a2c_gradients(h, ps, st, h->old_logp, ppo_clip)) return rc;
    if (int rc = a2c_apply(h, step, h->epoch_metrics + 6 * e, stream, e == num_epochs - 1)) return rc;
  }
  mean_metrics_kernel<<<1, 32, 0, st>>>(h->epoch_metrics, num_epochs, metrics_out ? metrics_out : h->metrics);
  MARL_CUDA_TRY(cudaGetLastError());
  return MARL_OK;
}

int marl_a2c_update(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, int64_t step, float* metrics_out, void* stream) {
  if (int rc = marl_a2c_update_grads(h, batch, n_envs, stream)) return rc;
  return marl_a2c_update_apply(h, step, metrics_out, stream);
}

}  // extern "C"
