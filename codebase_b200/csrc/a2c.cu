// a2c.cu -- C ABI of the independent actor-critic learner (marl_a2c_*), host-side orchestration of the fused kernels.
//
// Replaces marlbase/ac/model.py A2CNetwork (22-246): act's actor forward (147-153), get_value (155-163),
// update (189-246: target-critic pass, compute_nstep_returns utils/utils.py:38-63, evaluate_actions 165-182,
// policy-gradient + entropy + value losses, Adam, target sync), for independent or shared per-agent networks with a
// decentralised critic (critic.centralised: False, configs/algorithm/ia2c.yaml:18).
#include "learner.cuh"
#include <math.h>
#include <vector>

namespace marl {

constexpr int kMaxNStep = 64;

struct NStepParams {
  const float* vt;     // [N][P][T+1] target-critic values
  TrajView traj; const int32_t* idx; int N, P, n_steps;
  float gpow[kMaxNStep + 1];  // float32(gamma ** k), the Python-float powers of utils/utils.py:57-60
  float* ret;          // [N][P][T]
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
    const float src = (k == p.n_steps) ? p.vt[((size_t)a * p.P + b) * (T + 1) + tt] : p.traj.rew[(ep * p.N + a) * T + tt];
    acc += (p.gpow[k] * src) * (1.f - d);
  }
  p.ret[i] = acc;
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
};

extern "C" {

int marl_a2c_destroy(marl_a2c* h) {
  if (!h) return MARL_OK;
  cudaSetDevice(h->device);
  cudaFree(h->theta); cudaFree(h->theta_tgt); cudaFree(h->m); cudaFree(h->v); cudaFree(h->grad); cudaFree(h->scratch); cudaFree(h->loss_part);
  cudaFree(h->vt); cudaFree(h->ret); cudaFree(h->adv); cudaFree(h->metrics); cudaFree(h->idx); cudaFree(h->image);
  delete h;
  return MARL_OK;
}

int marl_a2c_create(const marl_mlp_cfg* actor, const marl_mlp_cfg* critic, const marl_a2c_hp* hp, int32_t max_envs, int32_t max_T, int32_t device, marl_a2c** out) {
  MARL_REQUIRE(hp && out, "marl_a2c_create: NULL argument");
  *out = nullptr;
  if (int rc = check_mlp_cfg(actor, "marl_a2c_create(actor)")) return rc;
  if (int rc = check_mlp_cfg(critic, "marl_a2c_create(critic)")) return rc;
  MARL_REQUIRE(actor->n_agents == critic->n_agents && actor->in_dim == critic->in_dim, "marl_a2c_create: actor / critic shapes differ (a centralised critic is not implemented)");
  MARL_REQUIRE(critic->out_dim == 1, "marl_a2c_create: the critic outputs one state value per agent");
  MARL_REQUIRE(max_envs >= 1 && max_T >= 1, "marl_a2c_create: max_envs/max_T must be >= 1");
  MARL_REQUIRE(hp->n_steps >= 1 && hp->n_steps <= kMaxNStep, "marl_a2c_create: n_steps %d out of range (1..%d)", hp->n_steps, kMaxNStep);
  if (int rc = check_device(device)) return rc;
  marl_a2c* h = new marl_a2c();
  h->actor = to_netset(actor); h->critic = to_netset(critic);
  h->hp = *hp; h->device = device; h->max_envs = max_envs; h->max_T = max_T;
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
  if (int rc2 = learner_kernels_init(actor->in_dim)) { marl_a2c_destroy(h); return rc2; }
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

int marl_a2c_update_grads(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, void* stream) {
  MARL_REQUIRE(h && batch, "marl_a2c_update_grads: NULL argument");
  MARL_REQUIRE(n_envs >= 1 && n_envs <= h->max_envs && n_envs <= batch->capacity, "marl_a2c_update_grads: n_envs %d out of range", n_envs);
  MARL_REQUIRE(batch->T >= 1 && batch->T <= h->max_T, "marl_a2c_update_grads: T %d exceeds max_T %d", batch->T, h->max_T);
  MARL_REQUIRE(batch->n_agents == h->actor.n_agents && batch->obs_dim == h->actor.in, "marl_a2c_update_grads: batch shape mismatch");
  MARL_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int T = batch->T, N = h->actor.n_agents;
  const int min_units = (64 + T) / (T + 1) > 0 ? (64 + T) / (T + 1) : 1;
  RowSource src; memset(&src, 0, sizeof(src));
  src.mode = 1; src.traj = to_view(batch); src.idx = h->idx; src.N = N; src.D = h->actor.in;
  const RowPlan cplan = make_plan(h->critic, n_envs, T + 1, h->n_sm, min_units);
  const RowPlan aplan = make_plan(h->actor, n_envs, T + 1, h->n_sm, min_units);
  // 1. target critic on all T+1 observations (ac/model.py:190-193)
  if (int rc = forward_any(h->critic, cplan, src, h->theta_tgt, h->image, h->vt, st)) return rc;
  // 2. n-step returns (ac/model.py:198-201)
  NStepParams np; np.vt = h->vt; np.traj = src.traj; np.idx = h->idx; np.N = N; np.P = n_envs; np.n_steps = h->hp.n_steps; np.ret = h->ret;
  for (int k = 0; k <= h->hp.n_steps; ++k) np.gpow[k] = (float)pow((double)h->hp.gamma, (double)k);
  nstep_returns_kernel<<<(N * n_envs * T + 255) / 256, 256, 0, st>>>(np);
  MARL_CUDA_TRY(cudaGetLastError());
  // 3. critic: forward, value loss, backward; leaves advantage = returns - V for the actor pass
  TrainParams tp; memset(&tp, 0, sizeof(tp));
  tp.plan = cplan; tp.src = src; tp.theta = h->theta + h->n_actor; tp.lay = h->critic.lay; tp.scratch = h->scratch; tp.scratch_pitch = h->scratch_pitch;
  tp.loss_part = h->loss_part; tp.returns = h->ret; tp.adv_out = h->adv; tp.value_coef = h->hp.value_loss_coef;
  if (int rc = launch_train(tp, kHeadA2cCritic, st)) return rc;
  ReduceParams rp; memset(&rp, 0, sizeof(rp));  // (sumsq_part stays NULL: two passes write different gradient slices)
  rp.scratch = h->scratch; rp.loss_part = h->loss_part; rp.n_nets = h->critic.n_nets; rp.P = h->critic.lay.P; rp.scratch_pitch = h->scratch_pitch;
  memcpy(rp.cta_begin, cplan.cta_begin, sizeof(rp.cta_begin));
  rp.n_loss_parts = cplan.cta_begin[cplan.n_nets]; rp.grad = h->grad + h->n_actor; rp.stats = h->grad + h->n_params; rp.stats_accumulate = 0;
  if (int rc = launch_grad_reduce(rp, st)) return rc;
  // 4. actor: forward, log-softmax, policy-gradient + entropy loss, backward
  tp.plan = aplan; tp.theta = h->theta; tp.lay = h->actor.lay; tp.adv = h->adv; tp.entropy_coef = h->hp.entropy_coef;
  if (int rc = launch_train(tp, kHeadA2cActor, st)) return rc;
  rp.n_nets = h->actor.n_nets; rp.P = h->actor.lay.P; memcpy(rp.cta_begin, aplan.cta_begin, sizeof(rp.cta_begin));
  rp.n_loss_parts = aplan.cta_begin[aplan.n_nets]; rp.grad = h->grad; rp.stats_accumulate = 1;
  return launch_grad_reduce(rp, st);
}

/* metrics_out device float[6] = (policy-gradient term, grad norm, entropy, value_loss, filled count, 0):
 * actor_loss = m[0] - entropy_coef*m[2]; loss = actor_loss + value_loss_coef*m[3]  (ac/model.py:216-226,241-246) */
int marl_a2c_update_apply(marl_a2c* h, int64_t step, float* metrics_out, void* stream) {
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
  if (tu > 1.0f && fmod((double)step, (double)tu) == 0.0) ap.target_mode = 1;
  else if (tu < 1.0f) ap.target_mode = 2;
  ap.loss_out = metrics_out ? metrics_out : h->metrics;
  return launch_adam(ap, (cudaStream_t)stream);
}

int marl_a2c_update(marl_a2c* h, const marl_traj_view* batch, int32_t n_envs, int64_t step, float* metrics_out, void* stream) {
  if (int rc = marl_a2c_update_grads(h, batch, n_envs, stream)) return rc;
  return marl_a2c_update_apply(h, step, metrics_out, stream);
}

}  // extern "C"
