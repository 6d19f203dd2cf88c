// qlearn.cu — the discrete-action learner steps of the reference's dqn / sqn families on the GPU:
//   DDQN  algos/dqn/actor_learner.py:27-67   one Q network  q = mlp(x) -> [B, nA]; target r + gamma (1-d) q_targ(x2)[argmax_a q_main(x2)]
//   SQN   algos/sqn/actor_learner.py:27-73, algos/sqn/core.py:30-79   two Q networks; target
//         r + gamma (1-d) ( min(max_a q1_targ(x2), max_a q2_targ(x2)) - alpha * sum_a softmax(q1_main(x2)/alpha) log_softmax(q1_main(x2)/alpha) )
// Both minimise 0.5 * mean((backup - q_k(x)[a])^2) summed over their networks with tf.train.AdamOptimizer (TF1 form:
// lr_t = lr sqrt(1-b2^t)/(1-b1^t), epsilon added to sqrt(v)) and then polyak-average EVERY main variable into its target.
//
// They reuse the SAC1 learner's building blocks: the grouped fp32 GEMM (sac_gemm.cuh: virtual [x|1] operand so a layer is
// one [K+1, N] block = kernel rows + bias row, relu / relu-mask epilogues, split-K weight gradients), one row-wise loss
// kernel, and one reduce + Adam + polyak pass over the flat parameter buffer.  A step is ONE CUDA graph of 10 kernels:
//   k_ql_ingest (the caller's batch -> the handle's buffers; its arguments are patched per step) -> L1, L2, L3 (all forward
//   passes of a layer in ONE grouped launch) -> k_ql_loss -> B3 {dW3, dH2} -> B2 {dW2, dH1} -> B1 {dW1} -> k_ql_adam ->
//   k_ql_emit (losses and Q values to the caller's arrays; patched per step).
// The flat parameter layout IS the reference's variable order (per network: dense/kernel, dense/bias, dense_1/kernel, ...),
// so get / set weights are plain copies.
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>

#include "sac_gemm.cuh"

namespace ddrl {
namespace {

constexpr int QL_MAX_NETS = 2, QL_MAX_PASSES = 5, QL_MAX_ACT = 64;

struct QlLossArgs {
  int B, nA, mode, nnets;            // mode 0: DDQN, 1: SQN
  float gamma, alpha;
  const float *acts, *rews, *done;
  const float* Q[QL_MAX_PASSES];     // [B, nA] per forward pass
  float* dQ[QL_MAX_NETS];            // [B, nA] gradient of the loss wrt q_k(x)
  double* partial;                   // [nnets][blocks] per-CTA loss sums
  unsigned int* ticket;
  float* out_loss;                   // [nnets + 1]: per-network losses, then their sum (device scalars)
};

// one thread per row; per-CTA partial sums in fp64, summed in CTA order by the last CTA (deterministic)
__global__ void __launch_bounds__(256) k_ql_loss(const QlLossArgs a) {
  __shared__ double s_sum[QL_MAX_NETS][8];
  __shared__ bool s_last;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  double l[QL_MAX_NETS] = {0.0, 0.0};
  if (row < a.B) {
    const int nA = a.nA;
    const int act = (int)a.acts[row];            // tf.cast(a_ph, tf.int32)
    float target;
    if (a.mode == 0) {
      // DDQN: argmax of the ONLINE network at x2 (first maximum, like tf.argmax), evaluated by the target network
      const float* qx2 = a.Q[1] + (size_t)row * nA;
      int best = 0;
      for (int j = 1; j < nA; ++j) if (qx2[j] > qx2[best]) best = j;
      target = a.Q[2][(size_t)row * nA + best];
    } else {
      // SQN: min over the two target networks of max_a q, minus alpha * sum softmax * log_softmax of q1_main(x2) / alpha
      const float* q1x2 = a.Q[2] + (size_t)row * nA;
      const float *t1 = a.Q[3] + (size_t)row * nA, *t2 = a.Q[4] + (size_t)row * nA;
      float m1 = t1[0], m2 = t2[0], zmax = q1x2[0] / a.alpha;
      for (int j = 1; j < nA; ++j) { m1 = fmaxf(m1, t1[j]); m2 = fmaxf(m2, t2[j]); zmax = fmaxf(zmax, q1x2[j] / a.alpha); }
      float se = 0.0f;
      for (int j = 0; j < nA; ++j) se += expf(q1x2[j] / a.alpha - zmax);
      const float lse = zmax + logf(se);
      float ent = 0.0f;                          // sum_a exp(pi_log) * pi_log  (= -entropy)
      for (int j = 0; j < nA; ++j) { const float pl = q1x2[j] / a.alpha - lse; ent += expf(pl) * pl; }
      target = fminf(m1, m2) - a.alpha * ent;
    }
    const float backup = a.rews[row] + a.gamma * (1.0f - a.done[row]) * target;
    for (int k = 0; k < a.nnets; ++k) {
      const float qa = (act >= 0 && act < nA) ? a.Q[k][(size_t)row * nA + act] : 0.0f;   // one_hot of an out-of-range index is all zero
      const float diff = backup - qa;
      l[k] = 0.5 * (double)diff * (double)diff;
      float* dq = a.dQ[k] + (size_t)row * nA;
      for (int j = 0; j < nA; ++j) dq[j] = (j == act) ? -diff / (float)a.B : 0.0f;
    }
  }
  // CTA reduction (warp shuffle, then 8 warps), per network
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < a.nnets; ++k) {
    double v = l[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s_sum[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < a.nnets; ++k) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += s_sum[k][w];
      a.partial[(size_t)k * gridDim.x + blockIdx.x] = v;
    }
    __threadfence();
    s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float total = 0.0f;
    for (int k = 0; k < a.nnets; ++k) {
      double v = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) v += ((volatile double*)a.partial)[(size_t)k * gridDim.x + b];
      const float lk = (float)(v / (double)a.B);
      a.out_loss[k] = lk;
      total += lk;
    }
    a.out_loss[a.nnets] = total;
    *a.ticket = 0u;
  }
}

// split-K partial gradients -> Adam (TF1) -> polyak, one pass over the flat buffers
__global__ void __launch_bounds__(256) k_ql_adam(int64_t P, int S, const float* __restrict__ Gp, float* G, const float* __restrict__ lr_dev,
                                                 float polyak, float* W, float* Wt, float* Mo, float* Vo) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float lr_t = *lr_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float g = 0.0f;
    for (int s = 0; s < S; ++s) g += Gp[(size_t)s * P + i];
    G[i] = g;
    const float m = b1 * Mo[i] + (1.0f - b1) * g;
    const float v = b2 * Vo[i] + (1.0f - b2) * g * g;
    const float w = W[i] - lr_t * m / (sqrtf(v) + eps);
    Mo[i] = m; Vo[i] = v; W[i] = w;
    Wt[i] = polyak * Wt[i] + (1.0f - polyak) * w;
  }
}

// First and last node of the step's graph: their arguments are the only per-step values of a step (the caller's batch
// arrays, this step's bias-corrected learning rate, the caller's output arrays) and are patched into the instantiated
// graph before every launch; everything in between reads and writes buffers owned by the handle.
struct QlIngest {
  const float *obs1, *obs2, *acts, *rews, *done;
  float *X1, *X2, *ACTS, *R, *DN, *lr_dev;
  float lr_t;
  int B, D;
};
__global__ void __launch_bounds__(256) k_ql_ingest(const QlIngest a) {
  const int64_t n = (int64_t)a.B * a.D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { a.X1[i] = a.obs1[i]; a.X2[i] = a.obs2[i]; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += stride) { a.ACTS[i] = a.acts[i]; a.R[i] = a.rews[i]; a.DN[i] = a.done[i]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.lr_dev = a.lr_t;
}
struct QlEmit {
  const float* loss; float* out_loss; int nloss;
  const float* q[QL_MAX_NETS]; float* out_q; int nnets; int64_t per_net;
};
__global__ void __launch_bounds__(256) k_ql_emit(const QlEmit a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a.out_loss && tid < a.nloss) a.out_loss[tid] = a.loss[tid];
  if (a.out_q)
    for (int k = 0; k < a.nnets; ++k)
      for (int64_t i = tid; i < a.per_net; i += stride) a.out_q[(size_t)k * a.per_net + i] = a.q[k][i];
}

__global__ void k_ql_copy(int64_t n, const float* __restrict__ src, float* dst, float* dst2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    dst[i] = v;
    if (dst2) dst2[i] = v;
  }
}

struct QlGroup { GemmGroup grp; int tiles = 0; };

}  // namespace
}  // namespace ddrl

using namespace ddrl;

struct ddrl_ql {
  int device = 0, D = 0, nA = 0, h1 = 0, h2 = 0, maxB = 0, nnets = 1, mode = 0, sms = 148, Smax = 1;
  float gamma = 0.99f, polyak = 0.995f, lr = 1e-3f, alpha = 0.1f;
  int64_t Pnet = 0, P = 0, o1 = 0, o2 = 0, o3 = 0;      // per-network block offsets: [D+1,h1], [h1+1,h2], [h2+1,nA]
  float *W = nullptr, *Wt = nullptr, *Mo = nullptr, *Vo = nullptr, *G = nullptr, *Gp = nullptr;
  float *H1[QL_MAX_PASSES] = {}, *H2[QL_MAX_PASSES] = {}, *Q[QL_MAX_PASSES] = {};
  float *dQ[QL_MAX_NETS] = {}, *dH2[QL_MAX_NETS] = {}, *dH1[QL_MAX_NETS] = {};
  float *X1 = nullptr, *X2 = nullptr, *ACTS = nullptr, *R = nullptr, *DN = nullptr, *lr_dev = nullptr;   // the step's own copy of the batch
  float* loss = nullptr;
  double* partial = nullptr;
  unsigned int* ticket = nullptr;
  int64_t t = 0;
  bool use_graph = true;                // DDRL_NO_GRAPH=1: plain stream launches (profilers)
  cudaStream_t cap_stream = nullptr;
  struct StepGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; cudaGraphNode_t ingest = nullptr, emit = nullptr; int64_t kernels = 0; };
  std::map<int, StepGraph> graphs;      // per batch size
  std::vector<void*> allocs;
};

namespace {

int ql_alloc(ddrl_ql* h, float** p, size_t nfloats) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, nfloats * sizeof(float) + 16);
  if (e != cudaSuccess) return fail(DDRL_ENOMEM, "cudaMalloc(%zu floats) failed: %s", nfloats, cudaGetErrorString(e));
  if ((e = cudaMemset(q, 0, nfloats * sizeof(float) + 16)) != cudaSuccess)
    return fail(DDRL_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = (float*)q;
  return 0;
}

GemmProb ql_prob(const float* a, int lda, int aw, int a_trans, const float* Bm, int ldb, int b_trans, float* C, int ldc, int M, int N,
                 int K, int epi = EPI_NONE, const float* mask = nullptr, int ldmask = 0) {
  GemmProb p{};
  p.a0 = Seg{a, lda, aw}; p.a1 = Seg{nullptr, 0, 0};
  p.a_ones = 1; p.a_trans = a_trans;
  p.B = Bm; p.ldb = ldb; p.b_trans = b_trans;
  p.C = C; p.ldc = ldc; p.c_split_stride = 0;
  p.M = M; p.N = N; p.K = K;
  p.epi = epi; p.mask = mask; p.ldmask = ldmask;
  p.splits = 1; p.k_per_split = K;
  p.cfg = N <= 16 ? 1 : 0;
  return p;
}

int ql_launch(std::vector<GemmProb>& v, cudaStream_t s) {
  if ((int)v.size() > GEMM_MAX_PROBS) return fail(DDRL_EINVAL, "too many GEMM problems in one stage (%d)", (int)v.size());
  GemmGroup g{};
  int t = 0;
  g.nprob = (int)v.size();
  for (size_t i = 0; i < v.size(); ++i) {
    GemmProb& p = v[i];
    const int BM = p.cfg == 0 ? 64 : 128, BN = p.cfg == 0 ? 64 : 16;
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    p.tile_begin = t;
    t += p.tiles_m * p.tiles_n * p.splits;
    g.p[i] = p;
  }
  if (t == 0) return 0;
  DDRL_CUDA(launch_pdl(gemm_grouped_f32, dim3(t), dim3(256), 0, s, g));
  DDRL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int ddrl_ql_create(int device, int obs_dim, int n_actions, int h1, int h2, int max_batch, int n_nets, float gamma, float polyak,
                   float lr, float alpha, ddrl_ql_t* out) {
  if (!out) return fail(DDRL_EINVAL, "ddrl_ql_create: out is NULL");
  *out = nullptr;
  if (obs_dim < 1 || n_actions < 1 || n_actions > QL_MAX_ACT || h1 < 1 || h2 < 1 || max_batch < 1)
    return fail(DDRL_EINVAL, "ddrl_ql_create: obs_dim=%d n_actions=%d (<= %d) h1=%d h2=%d max_batch=%d", obs_dim, n_actions,
                QL_MAX_ACT, h1, h2, max_batch);
  if (n_nets != 1 && n_nets != 2) return fail(DDRL_EINVAL, "ddrl_ql_create: n_nets must be 1 (DDQN) or 2 (SQN)");
  if (n_nets == 2 && !(alpha > 0.0f)) return fail(DDRL_EINVAL, "ddrl_ql_create: SQN needs alpha > 0 (algos/sqn/hyperparams.py:26)");
  int ndev = 0;
  DDRL_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DDRL_EINVAL, "ddrl_ql_create: device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_ql_create: cannot select device %d", device);
  ddrl_ql* h = new ddrl_ql();
  h->device = device; h->D = obs_dim; h->nA = n_actions; h->h1 = h1; h->h2 = h2; h->maxB = max_batch;
  h->nnets = n_nets; h->mode = n_nets == 2 ? 1 : 0;
  h->gamma = gamma; h->polyak = polyak; h->lr = lr; h->alpha = alpha;
  h->sms = sm_count(device);
  h->o1 = 0;
  h->o2 = (int64_t)(obs_dim + 1) * h1;
  h->o3 = h->o2 + (int64_t)(h1 + 1) * h2;
  h->Pnet = h->o3 + (int64_t)(h2 + 1) * n_actions;
  h->P = h->Pnet * n_nets;
  h->Smax = (max_batch + 255) / 256;
  const size_t P = (size_t)h->P, M = (size_t)max_batch;
  const int npass = h->mode == 0 ? 3 : 5;
  int rc = 0;
  auto A_ = [&](float** p, size_t n) { if (!rc) rc = ql_alloc(h, p, n); };
  A_(&h->W, P); A_(&h->Wt, P); A_(&h->Mo, P); A_(&h->Vo, P); A_(&h->G, P); A_(&h->Gp, P * h->Smax);
  for (int p = 0; p < npass; ++p) { A_(&h->H1[p], M * h1); A_(&h->H2[p], M * h2); A_(&h->Q[p], M * n_actions); }
  for (int k = 0; k < n_nets; ++k) { A_(&h->dQ[k], M * n_actions); A_(&h->dH2[k], M * h2); A_(&h->dH1[k], M * h1); }
  A_(&h->X1, M * obs_dim); A_(&h->X2, M * obs_dim); A_(&h->ACTS, M); A_(&h->R, M); A_(&h->DN, M); A_(&h->lr_dev, 1);
  { const char* e = getenv("DDRL_NO_GRAPH"); h->use_graph = !(e && e[0] == '1'); }
  A_(&h->loss, 4);
  float* tmp = nullptr;
  A_(&tmp, 2 * (size_t)QL_MAX_NETS * ((M + 255) / 256) + 2);
  h->partial = reinterpret_cast<double*>(tmp);
  A_(&tmp, 1);
  h->ticket = reinterpret_cast<unsigned int*>(tmp);
  if (rc) { ddrl_ql_destroy(h); return rc; }
  *out = h;
  return 0;
}

int ddrl_ql_destroy(ddrl_ql_t h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->graphs) { if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec); if (kv.second.graph) cudaGraphDestroy(kv.second.graph); }
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

int64_t ddrl_ql_param_count(ddrl_ql_t h) { return h ? h->P : -1; }

int ddrl_ql_set_weights(ddrl_ql_t h, const float* d_flat, int also_target, void* stream) {
  if (!h || !d_flat) return fail(DDRL_EINVAL, "ddrl_ql_set_weights: NULL argument");
  DeviceGuard guard(h->device);
  const int blocks = (int)std::min<int64_t>((h->P + 255) / 256, h->sms * 8);
  k_ql_copy<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->P, d_flat, h->W, also_target ? h->Wt : nullptr);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_ql_get_weights(ddrl_ql_t h, float* d_flat, int which, void* stream) {
  if (!h || !d_flat) return fail(DDRL_EINVAL, "ddrl_ql_get_weights: NULL argument");
  const float* src = which == 0 ? h->W : which == 1 ? h->Wt : which == 2 ? h->Mo : which == 3 ? h->Vo : which == 4 ? h->G : nullptr;
  if (!src) return fail(DDRL_EINVAL, "ddrl_ql_get_weights: which=%d not in 0..4", which);
  DeviceGuard guard(h->device);
  const int blocks = (int)std::min<int64_t>((h->P + 255) / 256, h->sms * 8);
  k_ql_copy<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->P, src, d_flat, nullptr);
  DDRL_LAUNCH_CHECK();
  return 0;
}

// forward passes of one layer for every (input, network, main|target) pass of the step, in one grouped launch
static int ql_forward(ddrl_ql* h, int B, const float* const* x_of_pass, const int* net_of_pass, const int* targ_of_pass, int npass,
                      cudaStream_t s) {
  const int D = h->D, h1 = h->h1, h2 = h->h2, nA = h->nA;
  std::vector<GemmProb> v;
  auto Wof = [&](int p) { return (targ_of_pass[p] ? h->Wt : h->W) + (int64_t)net_of_pass[p] * h->Pnet; };
  for (int p = 0; p < npass; ++p) v.push_back(ql_prob(x_of_pass[p], D, D, 0, Wof(p) + h->o1, h1, 0, h->H1[p], h1, B, h1, D + 1, EPI_RELU));
  int rc = ql_launch(v, s);
  if (rc) return rc;
  v.clear();
  for (int p = 0; p < npass; ++p) v.push_back(ql_prob(h->H1[p], h1, h1, 0, Wof(p) + h->o2, h2, 0, h->H2[p], h2, B, h2, h1 + 1, EPI_RELU));
  if ((rc = ql_launch(v, s))) return rc;
  v.clear();
  for (int p = 0; p < npass; ++p) v.push_back(ql_prob(h->H2[p], h2, h2, 0, Wof(p) + h->o3, nA, 0, h->Q[p], nA, B, nA, h2 + 1));
  return ql_launch(v, s);
}

// everything of a step between the ingest and the emit kernels: reads the handle's copy of the batch, static pointers only
static int ql_step_body(ddrl_ql* h, int B, cudaStream_t s) {
  const int D = h->D, h1 = h->h1, h2 = h->h2, nA = h->nA, K = h->nnets;
  // forward passes.  DDQN: main@x, main@x2, target@x2.  SQN: q1 main@x, q2 main@x, q1 main@x2, q1 target@x2, q2 target@x2
  const float* xs[QL_MAX_PASSES];
  int net[QL_MAX_PASSES], targ[QL_MAX_PASSES], npass;
  if (h->mode == 0) {
    npass = 3;
    xs[0] = h->X1; xs[1] = h->X2; xs[2] = h->X2;
    net[0] = net[1] = net[2] = 0;
    targ[0] = 0; targ[1] = 0; targ[2] = 1;
  } else {
    npass = 5;
    xs[0] = h->X1; xs[1] = h->X1; xs[2] = h->X2; xs[3] = h->X2; xs[4] = h->X2;
    net[0] = 0; net[1] = 1; net[2] = 0; net[3] = 0; net[4] = 1;
    targ[0] = targ[1] = targ[2] = 0; targ[3] = targ[4] = 1;
  }
  int rc = ql_forward(h, B, xs, net, targ, npass, s);
  if (rc) return rc;
  QlLossArgs la{};
  la.B = B; la.nA = nA; la.mode = h->mode; la.nnets = K; la.gamma = h->gamma; la.alpha = h->alpha;
  la.acts = h->ACTS; la.rews = h->R; la.done = h->DN;
  for (int p = 0; p < npass; ++p) la.Q[p] = h->Q[p];
  for (int k = 0; k < K; ++k) la.dQ[k] = h->dQ[k];
  la.partial = h->partial; la.ticket = h->ticket; la.out_loss = h->loss;
  k_ql_loss<<<(B + 255) / 256, 256, 0, s>>>(la);
  DDRL_LAUNCH_CHECK();
  // backward of the differentiated passes (pass k = network k at x): weight gradients as split-K partials over the batch
  const int S = (B + 255) / 256, kps = 256;
  auto wg = [&](GemmProb p) { p.splits = S; p.k_per_split = kps; p.c_split_stride = h->P; return p; };
  std::vector<GemmProb> v;
  for (int k = 0; k < K; ++k) {
    float* Gk = h->Gp + (int64_t)k * h->Pnet;
    const float* Wk = h->W + (int64_t)k * h->Pnet;
    v.push_back(wg(ql_prob(h->H2[k], h2, h2, 1, h->dQ[k], nA, 0, Gk + h->o3, nA, h2 + 1, nA, B)));                       // d[W3;b3] = [H2|1]^T dQ
    GemmProb dg = ql_prob(h->dQ[k], nA, nA, 0, Wk + h->o3, nA, 1, h->dH2[k], h2, B, h2, nA, EPI_MASK, h->H2[k], h2);      // dH2 = dQ W3^T . relu'
    dg.a_ones = 0;
    v.push_back(dg);
  }
  if ((rc = ql_launch(v, s))) return rc;
  v.clear();
  for (int k = 0; k < K; ++k) {
    float* Gk = h->Gp + (int64_t)k * h->Pnet;
    const float* Wk = h->W + (int64_t)k * h->Pnet;
    v.push_back(wg(ql_prob(h->H1[k], h1, h1, 1, h->dH2[k], h2, 0, Gk + h->o2, h2, h1 + 1, h2, B)));                      // d[W2;b2] = [H1|1]^T dH2
    GemmProb dg = ql_prob(h->dH2[k], h2, h2, 0, Wk + h->o2, h2, 1, h->dH1[k], h1, B, h1, h2, EPI_MASK, h->H1[k], h1);     // dH1 = dH2 W2^T . relu'
    dg.a_ones = 0;
    v.push_back(dg);
  }
  if ((rc = ql_launch(v, s))) return rc;
  v.clear();
  for (int k = 0; k < K; ++k)
    v.push_back(wg(ql_prob(xs[k], D, D, 1, h->dH1[k], h1, 0, h->Gp + (int64_t)k * h->Pnet + h->o1, h1, D + 1, h1, B)));   // d[W1;b1] = [x|1]^T dH1
  if ((rc = ql_launch(v, s))) return rc;
  const int blocks = (int)std::min<int64_t>((h->P + 255) / 256, h->sms * 8);
  k_ql_adam<<<blocks, 256, 0, s>>>(h->P, S, h->Gp, h->G, h->lr_dev, h->polyak, h->W, h->Wt, h->Mo, h->Vo);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_ql_step(ddrl_ql_t h, const float* d_obs1, const float* d_obs2, const float* d_acts, const float* d_rews,
                 const float* d_done, int batch, float* d_out_loss, float* d_out_q, void* stream) {
  if (!h) return fail(DDRL_EINVAL, "ddrl_ql_step: NULL handle");
  if (batch < 1 || batch > h->maxB) return fail(DDRL_EINVAL, "ddrl_ql_step: batch=%d not in [1, %d]", batch, h->maxB);
  if (!d_obs1 || !d_obs2 || !d_acts || !d_rews || !d_done) return fail(DDRL_EINVAL, "ddrl_ql_step: NULL batch array");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int B = batch;
  h->t += 1;
  const double t = (double)h->t;
  QlIngest in{d_obs1, d_obs2, d_acts, d_rews, d_done, h->X1, h->X2, h->ACTS, h->R, h->DN, h->lr_dev,
              (float)((double)h->lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t))), B, h->D};
  QlEmit out{};
  out.loss = h->loss; out.out_loss = d_out_loss; out.nloss = h->nnets + 1;
  for (int k = 0; k < h->nnets; ++k) out.q[k] = h->Q[k];       // the reference fetches q (DDQN) / q1, q2 (SQN) of the sampled states
  out.out_q = d_out_q; out.nnets = h->nnets; out.per_net = (int64_t)B * h->nA;
  const dim3 gin((unsigned)std::min<int64_t>(((int64_t)B * h->D + 255) / 256, h->sms * 4)), gout((unsigned)std::min<int64_t>((out.per_net + 255) / 256, h->sms * 2));
  auto enqueue = [&](cudaStream_t st) -> int {
    k_ql_ingest<<<gin, 256, 0, st>>>(in);
    DDRL_LAUNCH_CHECK();
    int rc = ql_step_body(h, B, st);
    if (rc) return rc;
    k_ql_emit<<<gout, 256, 0, st>>>(out);
    DDRL_LAUNCH_CHECK();
    return 0;
  };
  if (!h->use_graph) return enqueue(s);
  // the whole step is one CUDA graph per batch size (stream-order launches cost ~1.6 us each and complete on a 2.05 us
  // grid, tools/probes/tick_probe.cu); the first and last node carry this call's pointers and learning rate
  ddrl_ql::StepGraph& g = h->graphs[B];
  if (!g.exec) {
    if (!h->cap_stream) DDRL_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const int64_t before = g_launches.load();
    DDRL_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue(h->cap_stream);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g.graph);
    g.kernels = g_launches.load() - before;
    g_launches.store(before);      // captured, not executed
    if (rc) { if (g.graph) cudaGraphDestroy(g.graph); h->graphs.erase(B); return rc; }
    if (e != cudaSuccess) { h->graphs.erase(B); return fail(DDRL_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); }
    size_t n = 0;
    DDRL_CUDA(cudaGraphGetNodes(g.graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    DDRL_CUDA(cudaGraphGetNodes(g.graph, nodes.data(), &n));
    for (cudaGraphNode_t nd : nodes) {
      cudaGraphNodeType ty;
      cudaKernelNodeParams kp{};
      if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
      if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
      if (kp.func == (void*)k_ql_ingest) g.ingest = nd;
      if (kp.func == (void*)k_ql_emit) g.emit = nd;
    }
    if (!g.ingest || !g.emit) { cudaGraphDestroy(g.graph); h->graphs.erase(B); return fail(DDRL_ECUDA, "captured step lacks its ingest / emit node"); }
    e = cudaGraphInstantiate(&g.exec, g.graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g.graph); h->graphs.erase(B); return fail(DDRL_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
  } else {
    void* a_in[] = {&in};
    void* a_out[] = {&out};
    cudaKernelNodeParams np{};
    np.blockDim = dim3(256); np.sharedMemBytes = 0; np.extra = nullptr;
    np.func = (void*)k_ql_ingest; np.gridDim = gin; np.kernelParams = a_in;
    DDRL_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.ingest, &np));
    np.func = (void*)k_ql_emit; np.gridDim = gout; np.kernelParams = a_out;
    DDRL_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, g.emit, &np));
  }
  DDRL_CUDA(cudaGraphLaunch(g.exec, s));
  g_launches.fetch_add(g.kernels, std::memory_order_relaxed);
  return 0;
}

int ddrl_ql_forward(ddrl_ql_t h, const float* d_obs, int n, int net, float* d_out_q, void* stream) {
  if (!h || !d_obs || !d_out_q) return fail(DDRL_EINVAL, "ddrl_ql_forward: NULL argument");
  if (n < 1 || n > h->maxB) return fail(DDRL_EINVAL, "ddrl_ql_forward: n=%d not in [1, %d]", n, h->maxB);
  if (net < 0 || net >= h->nnets) return fail(DDRL_EINVAL, "ddrl_ql_forward: net=%d not in [0, %d)", net, h->nnets);
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const float* xs[1] = {d_obs};
  const int nets[1] = {net}, targ[1] = {0};
  int rc = ql_forward(h, n, xs, nets, targ, 1, s);
  if (rc) return rc;
  DDRL_CUDA(cudaMemcpyAsync(d_out_q, h->Q[0], (size_t)n * h->nA * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // extern "C"
