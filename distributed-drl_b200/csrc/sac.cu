// sac.cu — SAC1 learner step on the GPU (parity mode, fp32).
//
// Replaces Learner.train / get_weights / set_weights of algos/sac1/actor_learner.py:20-142 (graph in
// :27-101, nets in algos/sac1/core.py:15-121), i.e. one `sess.run(step_ops)`:
//   forward  3 policy passes (main@x, main@x2, target@x2) and 5 Q passes
//            (main Q1,Q2@(x,a); main Q1@(x,pi(x)); target Q1,Q2@(x2,pi_targ(x2)))
//   losses   pi_loss, q1_loss, q2_loss with the Bellman target  r + gamma (1-d)(min Q_targ - alpha logp2)
//   backward pi-loss wrt main/pi (through Q1's inputs only), value-loss wrt main/q1,q2 — both at theta_t
//   update   TF1 Adam(pi), TF1 Adam(q), then polyak over all main tensors (policy included) with the
//            post-update weights; optional entropy-alpha Adam step.
//
// Parameters, Adam moments, target weights and gradients are FLAT fp32 buffers with one contiguous
// [K+1, N] block per dense layer (kernel rows + bias row; see sac_gemm.cuh), so the optimiser, the
// target averaging, the gradient all-reduce and the parameter broadcast are each one pass / one
// collective over one buffer.  The policy's mu and log_std heads share one [h2+1, 2A] block.
// The step after the prologue is captured once per batch size into a CUDA graph (tensor-core GEMM stages and
// row-wise kernels on the main stream, bias / skinny gradients on a forked side stream); the prologue kernel is
// launched directly because it carries the per-step values (input pointers, Adam step numbers and bias-corrected
// rates computed on the host, Philox counter) as by-value arguments and publishes them in device memory.
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "sac_gemm.cuh"
#include "sac_gemm_tc.cuh"

namespace ddrl {

// ------------------------------------------------------------------------------------------------
// device-resident per-step state
// ------------------------------------------------------------------------------------------------
struct StepDyn {   // per-step values: by-value argument of the prologue kernel, which runs OUTSIDE the captured graph
  const float *obs1, *obs2, *acts, *rews, *done;  // external batch
  const float* noise;                              // external [3,B,A] noise or NULL (Philox)
  float *out_scalars, *out_q1, *out_q2, *out_logp; // nullable
  unsigned long long seed;
  float grad_scale;
  // advanced on the host (the handle mirrors the step count): Adam step numbers, TF1 Adam's bias-corrected
  // rates lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) (tensorflow/python/training/adam.py, 1.x), noise counter
  int t_pi, t_q;
  float lr_pi, lr_q;
  unsigned long long noise_counter;
  // fused sample -> update (ddrl_sac_step_from_buffer): the batch is gathered straight from the replay ring by the
  // prologue, with the same Philox index stream as ddrl_rb_sample (philox_index(row, seed, counter, stream, size))
  const float* ring;                               // NULL: external batch arrays above
  int ring_row_f;
  unsigned int ring_stream;
  unsigned long long ring_seed, ring_counter, ring_size;
};
struct StepState {  // persistent
  StepDyn dyn;
  int t_pi, t_q, t_alpha;
  unsigned long long noise_counter;
  float log_alpha, alpha_m, alpha_v;
  float alpha_cur;      // alpha used by this step (pre-update)
  int auto_alpha;
  float alpha_const;
  float lr, lr_pi, lr_q;  // base rate and TF1 Adam's bias-corrected rates for this step
  unsigned long long* trace;  // DDRL_DP_TRACE=1: %globaltimer stamps of the step's exchange (tools/dp_trace.py), else null
};
// step trace slots (two-kernel data-parallel exchange): 0 reduce kernel start, 1 its flag published, 2 optimiser kernel
// start, 3 all flags seen, 4 optimiser kernel end, 6 this step's prologue start, 7 the previous step's
__device__ __forceinline__ void step_stamp(const StepState* st, int slot) {
  if (st->trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (slot == 6) st->trace[7] = st->trace[6];
    st->trace[slot] = t;
  }
}

// copy the external batch into the learner's own buffers and materialise the noise
// tensor-core mode: the three concatenated inputs [x|a], [x|a1], [x2|a3] as pre-split hi/lo planes
// ([2][maxB][pitch]; the action columns of the last two are filled by the policy-head kernel)
struct XaOut {
  float *xa_d, *xa_f, *xa_g;   // nullptr: FFMA mode (plain X / X2 / ACT copies instead)
  int pitch;
  long long plane;
};
__device__ __forceinline__ void put_split(float* hi_plane, long long plane, size_t idx, float v) {
  float hi, lo;
  tc::split_tf32(v, hi, lo);
  hi_plane[idx] = hi;
  hi_plane[plane + idx] = lo;
}
__device__ __forceinline__ void d_prologue(int vb, int vgrid, StepState* st, const StepDyn& d, int B, int D, int A,
                                           float* X, float* X2, float* ACT, float* R, float* DN, float* NOISE, XaOut xa) {
  if (vb == 0 && threadIdx.x == 0) {
    // publish the step's values for the kernels of the captured graph that follow
    st->dyn = d;
    st->t_pi = d.t_pi; st->t_q = d.t_q; st->lr_pi = d.lr_pi; st->lr_q = d.lr_q; st->noise_counter = d.noise_counter;
    st->alpha_cur = st->auto_alpha ? expf(st->log_alpha) : st->alpha_const;
  }
  const int64_t tid = (int64_t)vb * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)vgrid * blockDim.x;
  if (d.ring) {
    // one warp per sampled row: packed ring row [obs1 | obs2 | acts | rew | done] -> the learner's input buffers
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = tid >> 5, nwarps = nthr >> 5;
    for (int64_t r = warp0; r < B; r += nwarps) {
      const int64_t idx = philox_index((uint64_t)r, d.ring_seed, d.ring_counter, d.ring_stream, d.ring_size);
      const float* src = d.ring + (size_t)idx * d.ring_row_f;
      for (int c = lane; c < 2 * D + A + 2; c += 32) {
        const float v = src[c];
        if (c < D) {
          if (xa.xa_d) { put_split(xa.xa_d, xa.plane, (size_t)r * xa.pitch + c, v); put_split(xa.xa_f, xa.plane, (size_t)r * xa.pitch + c, v); }
          else X[(size_t)r * D + c] = v;
        } else if (c < 2 * D) {
          if (xa.xa_d) put_split(xa.xa_g, xa.plane, (size_t)r * xa.pitch + (c - D), v);
          else X2[(size_t)r * D + (c - D)] = v;
        } else if (c < 2 * D + A) {
          if (xa.xa_d) put_split(xa.xa_d, xa.plane, (size_t)r * xa.pitch + D + (c - 2 * D), v);
          else ACT[(size_t)r * A + (c - 2 * D)] = v;
        } else if (c == 2 * D + A) R[r] = v;
        else DN[r] = v;
      }
    }
  } else if (d.obs1 && xa.xa_d) {
    for (int64_t i = tid; i < (int64_t)B * D; i += nthr) {
      const int64_t r = i / D, c = i - r * D;
      const size_t o = (size_t)r * xa.pitch + c;
      const float v1 = d.obs1[i], v2 = d.obs2[i];
      float hi, lo;
      tc::split_tf32(v1, hi, lo);
      xa.xa_d[o] = hi; xa.xa_d[xa.plane + o] = lo;
      xa.xa_f[o] = hi; xa.xa_f[xa.plane + o] = lo;
      put_split(xa.xa_g, xa.plane, o, v2);
    }
    for (int64_t i = tid; i < (int64_t)B * A; i += nthr) {
      const int64_t r = i / A, c = i - r * A;
      put_split(xa.xa_d, xa.plane, (size_t)r * xa.pitch + D + c, d.acts[i]);
    }
    for (int64_t i = tid; i < B; i += nthr) { R[i] = d.rews[i]; DN[i] = d.done[i]; }
  } else if (d.obs1) {
    for (int64_t i = tid; i < (int64_t)B * D; i += nthr) { X[i] = d.obs1[i]; X2[i] = d.obs2[i]; }
    for (int64_t i = tid; i < (int64_t)B * A; i += nthr) ACT[i] = d.acts[i];
    for (int64_t i = tid; i < B; i += nthr) { R[i] = d.rews[i]; DN[i] = d.done[i]; }
  }
  const int64_t n = 3LL * B * A;
  if (d.noise) {
    for (int64_t i = tid; i < n; i += nthr) NOISE[i] = d.noise[i];
  } else {
    // Philox4x32-10 -> two Box-Muller pairs per counter (oracle/replay_oracle.py: philox_normals)
    const unsigned long long ctr = d.noise_counter;
    for (int64_t q = tid; q < (n + 3) / 4; q += nthr) {
      const Philox4 p = philox4x32_10((uint32_t)q, (uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5AC1u,
                                      (uint32_t)d.seed, (uint32_t)(d.seed >> 32));
      const float u0 = ((float)p.x + 0.5f) * 2.3283064365386963e-10f;
      const float u1 = ((float)p.y + 0.5f) * 2.3283064365386963e-10f;
      const float u2 = ((float)p.z + 0.5f) * 2.3283064365386963e-10f;
      const float u3 = ((float)p.w + 0.5f) * 2.3283064365386963e-10f;
      const float r0 = sqrtf(-2.0f * logf(fmaxf(u0, 1e-30f))), r1 = sqrtf(-2.0f * logf(fmaxf(u2, 1e-30f)));
      float s0, c0, s1, c1;
      sincospif(2.0f * u1, &s0, &c0);
      sincospif(2.0f * u3, &s1, &c1);
      const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < n) NOISE[4 * q + j] = z[j];
    }
  }
}
__global__ void __launch_bounds__(256) k_prologue(StepState* st, const __grid_constant__ StepDyn dyn, int B, int D, int A, float* X,
                                                  float* X2, float* ACT, float* R, float* DN, float* NOISE, XaOut xa) {
  pdl_trigger();
  pdl_wait();
  step_stamp(st, 6);
  d_prologue(blockIdx.x, gridDim.x, st, dyn, B, D, A, X, X2, ACT, R, DN, NOISE, xa);
}

// ------------------------------------------------------------------------------------------------
// policy head, element-wise part (algos/sac1/core.py:45-46,71-87,104-106).  The float32 operation
// ORDER of the reference graph is kept (separate mul/add, (pi - mu), 1 - pi^2): the graph is
// ill-conditioned where std << |mu| or |u| is large, and only the same sequence of roundings gives
// the reference's numbers there.
// ------------------------------------------------------------------------------------------------
struct PolEl {
  float ls_t, log_std, std, u, diff, den, z, pi, omp, clipped;
};
__device__ __forceinline__ PolEl policy_elem(float mu, float ls_pre, float eps) {
  PolEl e;
  e.ls_t = tanhf(ls_pre);
  e.log_std = __fadd_rn(-20.0f, __fmul_rn(11.0f, __fadd_rn(e.ls_t, 1.0f)));  // LOG_STD_MIN + 0.5*(MAX-MIN)*(t+1)
  e.std = expf(e.log_std);
  e.u = __fadd_rn(mu, __fmul_rn(eps, e.std));                                // pi = mu + noise * std
  e.diff = __fsub_rn(e.u, mu);
  e.den = __fadd_rn(e.std, 1e-8f);
  e.z = __fdiv_rn(e.diff, e.den);
  e.pi = tanhf(e.u);
  e.omp = __fsub_rn(1.0f, __fmul_rn(e.pi, e.pi));
  e.clipped = fminf(fmaxf(e.omp, 0.0f), 1.0f);
  return e;
}
__device__ __forceinline__ float logp_term(const PolEl& e) {
  // -0.5 * (z^2 + 2 log_std + log(2 pi))  -  log(clip(1 - pi^2, 0, 1) + 1e-6)
  const float pre = __fmul_rn(-0.5f, __fadd_rn(__fadd_rn(__fmul_rn(e.z, e.z), __fmul_rn(2.0f, e.log_std)),
                                               1.8378770664093453f));
  return __fsub_rn(pre, logf(__fadd_rn(e.clipped, 1e-6f)));
}

// ------------------------------------------------------------------------------------------------
// Row-wise kernels: one warp per batch row.  The skinny layers (policy heads N = 2A, Q heads N = 1,
// their dgrads with K = 2A / 1 / A) are dot products against a few KB of weights — they are fused with
// the element-wise math that consumes them instead of being launched as 1-tile-wide GEMMs.
// ------------------------------------------------------------------------------------------------
constexpr int ROW_WARPS = 8;      // warps (= rows) per CTA
constexpr int MAX_HEAD = 64;      // 2A <= 64

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[j] = sum_k x[k] * Wt(k, j) (+ bias row), j in [0, n): weights row-major [K+1, ldw] when !w_trans
// (forward: W[k*ldw + j]), or [n, ldw] when w_trans (dgrad against W^T: W[j*ldw + k]).  Results land in
// sout[0..n) (shared, per warp), identical in every lane.
__device__ __forceinline__ void warp_dots(const float* __restrict__ x, int K, const float* __restrict__ W, int ldw, int n,
                                          bool w_trans, bool bias, float* sout, int lane,
                                          const float* __restrict__ x_lo = nullptr) {
  for (int j0 = 0; j0 < n; j0 += 8) {
    float acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0f;
    for (int k = lane; k < K; k += 32) {
      const float xv = x_lo ? x[k] + x_lo[k] : x[k];   // hi + lo is exact
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        if (j < n) acc[jj] = fmaf(xv, w_trans ? W[(size_t)j * ldw + k] : W[(size_t)k * ldw + j], acc[jj]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float v = warp_sum(acc[jj]);
      const int j = j0 + jj;
      if (j < n) {
        if (bias) v += W[(size_t)K * ldw + j];
        if (lane == 0) sout[j] = v;
      }
    }
  }
  __syncwarp();
}

// policy heads + squashed-Gaussian sample / log-likelihood for the policy passes
// (pass 0: main pi(x) -> A1, LOGP1, HD;  pass 1: main pi(x2) -> LOGP2;  pass 2: target pi(x2) -> A3)
// A launch covers `npass` of them: slot s of the grid (rows [s B, (s+1) B)) is logical pass pmap.p[s].  The step runs
// passes 0 and 2 after the first forward stage (their actions feed the second one) and pass 1 after the second
// forward stage, where main pi(x2) now lives so that both forward stages fit on the chip in one wave.
struct PassMap { int n, p[3]; };
// One THREAD per (row, head output): the row of H2 sits in shared memory, the weight column W[:, j] is read
// with the same address across the rows of a CTA (L1 broadcast) and consecutive j are consecutive floats.
// R = 256 / 2A rows per CTA.  ldh = row pitch of the head block (2A rounded up to 4 floats).
constexpr int HEADS_THREADS = 256;
constexpr size_t HEADS_SMEM_MAX = 96 * 1024;
// RB = rows per thread (register blocking: one weight load feeds RB rows).  A CTA covers R = RB * (256 / 2A) rows.
template <int RB>
__global__ void __launch_bounds__(HEADS_THREADS) k_policy_heads_fwd(
    int B, int A, int h2, int ldh, int R, float act_scale, const float* __restrict__ H2a, const float* __restrict__ H2b,
    const float* __restrict__ H2c, const float* __restrict__ Whead, const float* __restrict__ Whead_t,
    const float* __restrict__ NOISE, float* HD, float* A1, float* A3, float* LOGP1, float* LOGP2, XaOut xa, int D,
    PassMap pmap) {
  extern __shared__ __align__(16) float hsm[];
  pdl_trigger();
  pdl_wait();
  const int n = 2 * A, ldx = (h2 + 3) / 4 * 4 + 4;
  float* xs = hsm;                    // [R][ldx]
  float* s_out = xs + R * ldx;        // [R][n]   head pre-activations
  float* s_pre = s_out + R * n;       // [R][A]   gaussian log-likelihood terms
  float* s_sq = s_pre + R * A;        // [R][A]   squash correction terms
  const int tid = threadIdx.x;
  const int total = pmap.n * B, g0 = blockIdx.x * R;
  const bool vec = (h2 & 3) == 0;
  {
    // stage the R rows of H2 (a CTA's rows may straddle two passes)
    const int q = vec ? h2 >> 2 : h2;
    for (int r = tid / q, c = tid - (tid / q) * q; r < R; ) {
      const int g = g0 + r;
      if (g < total) {
        const int slot = g / B, row = g - slot * B, pass = pmap.p[slot];
        const float* src = (pass == 0 ? H2a : pass == 1 ? H2b : H2c) + (size_t)row * h2;
        if (vec) *reinterpret_cast<float4*>(xs + r * ldx + 4 * c) = *reinterpret_cast<const float4*>(src + 4 * c);
        else xs[r * ldx + c] = src[c];
      } else if (vec) {
        *reinterpret_cast<float4*>(xs + r * ldx + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        xs[r * ldx + c] = 0.0f;
      }
      c += HEADS_THREADS;
      while (c >= q) { c -= q; ++r; }
    }
  }
  __syncthreads();
  const int groups = R / RB;                       // row groups of RB rows; thread = (row group, output j)
  const int rg = tid / n, j = tid - rg * n;
  if (rg < groups) {
    // all RB rows of a thread use the same weight column: groups start at multiples of RB and the host picks
    // RB = 4 only when B % 4 == 0, so a group never straddles a pass boundary
    const int g_first = g0 + rg * RB;
    const int pass_w = pmap.p[min(g_first, total - 1) / B];
    const float* W = (pass_w == 2 ? Whead_t : Whead) + j;
    const float* x = xs + rg * RB * ldx;
    float acc[RB][2];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) acc[rb][0] = acc[rb][1] = 0.0f;
    int k = 0;
    if (vec) {
      for (; k + 8 <= h2; k += 8) {
        float wv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = W[(size_t)(k + i) * ldh];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          const float4 xa4 = *reinterpret_cast<const float4*>(x + rb * ldx + k);
          const float4 xb4 = *reinterpret_cast<const float4*>(x + rb * ldx + k + 4);
          acc[rb][0] = fmaf(xa4.w, wv[3], fmaf(xa4.z, wv[2], fmaf(xa4.y, wv[1], fmaf(xa4.x, wv[0], acc[rb][0]))));
          acc[rb][1] = fmaf(xb4.w, wv[7], fmaf(xb4.z, wv[6], fmaf(xb4.y, wv[5], fmaf(xb4.x, wv[4], acc[rb][1]))));
        }
      }
    }
    for (; k < h2; ++k) {
      const float wk = W[(size_t)k * ldh];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) acc[rb][0] = fmaf(x[rb * ldx + k], wk, acc[rb][0]);
    }
    const float bias = W[(size_t)h2 * ldh];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int r = rg * RB + rb, g = g0 + r;
      if (g < total) {
        const float v = (acc[rb][0] + acc[rb][1]) + bias;
        s_out[r * n + j] = v;
        if (pmap.p[g / B] == 0) HD[(size_t)(g % B) * n + j] = v;
      }
    }
  }
  __syncthreads();
  // element-wise part: one thread per (row, action)
  for (int i = tid; i < R * A; i += HEADS_THREADS) {
    const int r = i / A, ja = i - r * A, g = g0 + r;
    if (g >= total) continue;
    const int slot = g / B, row = g - slot * B, pass = pmap.p[slot];
    const float eps = NOISE[((size_t)pass * B + row) * A + ja];
    const PolEl e = policy_elem(s_out[r * n + ja], s_out[r * n + A + ja], eps);
    s_pre[i] = __fmul_rn(-0.5f, __fadd_rn(__fadd_rn(__fmul_rn(e.z, e.z), __fmul_rn(2.0f, e.log_std)), 1.8378770664093453f));
    s_sq[i] = logf(__fadd_rn(e.clipped, 1e-6f));
    const float act = __fmul_rn(e.pi, act_scale);
    if (pass == 0) {
      A1[(size_t)row * A + ja] = act;
      if (xa.xa_f) put_split(xa.xa_f, xa.plane, (size_t)row * xa.pitch + D + ja, act);
    } else if (pass == 2) {
      A3[(size_t)row * A + ja] = act;
      if (xa.xa_g) put_split(xa.xa_g, xa.plane, (size_t)row * xa.pitch + D + ja, act);
    }
  }
  __syncthreads();
  for (int r = tid; r < R; r += HEADS_THREADS) {
    const int g = g0 + r;
    if (g >= total) continue;
    const int pass = pmap.p[g / B];
    if (pass == 2) continue;                        // the target policy needs no log-likelihood
    // reduce_sum over the action axis, in index order, of the two terms separately (core.py:32,86)
    float gauss = 0.0f, squash = 0.0f;
    for (int t = 0; t < A; ++t) {
      gauss = __fadd_rn(gauss, s_pre[r * A + t]);
      squash = __fadd_rn(squash, s_sq[r * A + t]);
    }
    (pass == 0 ? LOGP1 : LOGP2)[g % B] = __fsub_rn(gauss, squash);
  }
}

// 128-bit helpers for the warp-per-row kernels below: lane handles 4 consecutive features per step of 128
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Narrow policy heads (2A <= 16): one WARP per row, lanes split the h2 features four at a time, the head block
// rows (ldh = 4 * Q floats, zero padded) are read as Q 128-bit loads per feature; then 2A warp reductions.
template <int Q>
__device__ __forceinline__ void heads_row_dots(const float* __restrict__ x, int K, const float* __restrict__ W, int n, float* sout,
                                               int lane) {
  constexpr int LDH = 4 * Q;
  float acc[LDH];
#pragma unroll
  for (int j = 0; j < LDH; ++j) acc[j] = 0.0f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 x4 = ld4(x + k);
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
    float4 w[4][Q];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < Q; ++q) w[i][q] = ld4(W + (size_t)(k + i) * LDH + 4 * q);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        acc[4 * q] = fmaf(xv[i], w[i][q].x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(xv[i], w[i][q].y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(xv[i], w[i][q].z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(xv[i], w[i][q].w, acc[4 * q + 3]);
      }
  }
#pragma unroll
  for (int j = 0; j < LDH; ++j) {
    const float v = warp_sum(acc[j]);
    if (j < n && lane == 0) sout[j] = v + W[(size_t)K * LDH + j];
  }
  __syncwarp();
}
// one row of the narrow policy head + squashed-Gaussian sample / log-likelihood; all lanes return logp, lanes < A the
// scaled action in *act_out, lanes < 2A find the head pre-activations in sout
__device__ __forceinline__ float heads_row_eval(const float* __restrict__ x, int h2, int ldh, int A, const float* __restrict__ W,
                                                const float* __restrict__ eps, float act_scale, float* sout, int lane,
                                                float* act_out) {
  switch (ldh) {
    case 4: heads_row_dots<1>(x, h2, W, 2 * A, sout, lane); break;
    case 8: heads_row_dots<2>(x, h2, W, 2 * A, sout, lane); break;
    case 12: heads_row_dots<3>(x, h2, W, 2 * A, sout, lane); break;
    default: heads_row_dots<4>(x, h2, W, 2 * A, sout, lane); break;
  }
  float pre = 0.0f, sq = 0.0f;
  *act_out = 0.0f;
  if (lane < A) {
    const PolEl e = policy_elem(sout[lane], sout[A + lane], eps[lane]);
    pre = __fmul_rn(-0.5f, __fadd_rn(__fadd_rn(__fmul_rn(e.z, e.z), __fmul_rn(2.0f, e.log_std)), 1.8378770664093453f));
    sq = logf(__fadd_rn(e.clipped, 1e-6f));
    *act_out = __fmul_rn(e.pi, act_scale);
  }
  // reduce_sum over the action axis, in index order, of the two terms separately (core.py:32,86)
  float gauss = 0.0f, squash = 0.0f;
  for (int t = 0; t < A; ++t) {
    gauss = __fadd_rn(gauss, __shfl_sync(0xffffffffu, pre, t));
    squash = __fadd_rn(squash, __shfl_sync(0xffffffffu, sq, t));
  }
  return __fsub_rn(gauss, squash);
}
__global__ void __launch_bounds__(ROW_WARPS * 32) k_policy_heads_rows(
    int B, int A, int h2, int ldh, float act_scale, const float* __restrict__ H2a, const float* __restrict__ H2b,
    const float* __restrict__ H2c, const float* __restrict__ Whead, const float* __restrict__ Whead_t,
    const float* __restrict__ NOISE, float* HD, float* A1, float* A3, float* LOGP1, float* LOGP2, XaOut xa, int D,
    PassMap pmap) {
  __shared__ float s_out[ROW_WARPS][16];
  pdl_trigger();
  pdl_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * ROW_WARPS + w;
  if (g >= pmap.n * B) return;
  const int slot = g / B, row = g - slot * B, pass = pmap.p[slot];
  const float* x = (pass == 0 ? H2a : pass == 1 ? H2b : H2c) + (size_t)row * h2;
  float act_val;
  const float logp = heads_row_eval(x, h2, ldh, A, pass == 2 ? Whead_t : Whead, NOISE + ((size_t)pass * B + row) * A, act_scale,
                                    s_out[w], lane, &act_val);
  if (lane < A) {
    if (pass == 0) {
      A1[(size_t)row * A + lane] = act_val;
      if (xa.xa_f) put_split(xa.xa_f, xa.plane, (size_t)row * xa.pitch + D + lane, act_val);
    } else if (pass == 2) {
      A3[(size_t)row * A + lane] = act_val;
      if (xa.xa_g) put_split(xa.xa_g, xa.plane, (size_t)row * xa.pitch + D + lane, act_val);
    }
  }
  if (pass == 0) {
    if (lane < 2 * A) HD[(size_t)row * 2 * A + lane] = s_out[w][lane];
    if (lane == 0) LOGP1[row] = logp;
  } else if (pass == 1 && lane == 0) {
    LOGP2[row] = logp;
  }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}
__device__ __forceinline__ void st_split4(float* hi_plane, long long plane, size_t idx, const float4& v) {
  float4 hi, lo;
  tc::split_tf32(v.x, hi.x, lo.x); tc::split_tf32(v.y, hi.y, lo.y); tc::split_tf32(v.z, hi.z, lo.z); tc::split_tf32(v.w, hi.w, lo.w);
  *reinterpret_cast<float4*>(hi_plane + idx) = hi;
  *reinterpret_cast<float4*>(hi_plane + plane + idx) = lo;
}

// Q heads of all five Q passes, Bellman target, the three losses (actor_learner.py:58-69), the
// output-layer gradients dq, and dZ2 = dq (x) w3^T masked by relu'(H2) for the three differentiated
// passes.  Loss sums: per-CTA partials in double, combined in CTA order by the last CTA to finish
// (deterministic).  One warp per row; when h2 % 4 == 0 every access is a 128-bit load / store.
__global__ void __launch_bounds__(ROW_WARPS * 32) k_qheads_losses(
    StepState* st, int B, int h2, float gamma, const float* __restrict__ H2d, const float* __restrict__ H2e,
    const float* __restrict__ H2f, const float* __restrict__ H2g, const float* __restrict__ H2h,
    const float* __restrict__ W3q1, const float* __restrict__ W3q2, const float* __restrict__ W3q1t,
    const float* __restrict__ W3q2t, const float* __restrict__ R, const float* __restrict__ DN,
    const float* __restrict__ LOGP1, float* LOGP2, float* dQd, float* dQe, float* dZ2d, float* dZ2e,
    float* dZ2f, double* partials, unsigned int* ticket, float* SCAL, int ldz, long long zlo,
    // fold_b: the narrow policy head of pass b (main pi(x2) -> logp2) is evaluated here, by the warp that owns the row,
    // instead of by a launch of its own between the second forward stage and this kernel
    int fold_b, const float* __restrict__ H2b, const float* __restrict__ Whead, const float* __restrict__ NOISE, int A, int ldh,
    float act_scale) {
  __shared__ double s_part[ROW_WARPS][4];
  __shared__ double s_red[ROW_WARPS * 32][4];
  __shared__ float s_hout[ROW_WARPS][16];
  __shared__ bool s_last;
  pdl_trigger();
  pdl_wait();
  if (ldz == 0) ldz = h2;   // zlo != 0: hi/lo planes for the tensor-core GEMMs (lo plane zlo floats after hi)
  const int vb = blockIdx.x, vgrid = gridDim.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = vb * ROW_WARPS + w;
  const float alpha = st->alpha_cur;
  const StepDyn& d = st->dyn;
  const bool vec = (h2 & 3) == 0 && (ldz & 3) == 0;
  double t_pi = 0.0, t_q1 = 0.0, t_q2 = 0.0, t_lp = 0.0;
  if (row < B) {
    const float *hd = H2d + (size_t)row * h2, *he = H2e + (size_t)row * h2, *hf = H2f + (size_t)row * h2,
                *hg = H2g + (size_t)row * h2, *hh = H2h + (size_t)row * h2;
    float qd = 0.f, qe = 0.f, qf = 0.f, qg = 0.f, qh = 0.f;
    // h2 <= 256 and 128-bit rows: every lane keeps its (<= 2 x 4) features of the five rows and the four weight vectors
    // in registers — all loads of the kernel are issued before anything is consumed (one L2 round trip), and the dZ2
    // pass below needs no second read
    constexpr int NV = 2;
    const bool fast = vec && h2 <= 128 * NV;
    float4 xd[NV], xe[NV], xf[NV], w1[NV], w2[NV];
    if (fast) {
      float4 xg[NV], xh[NV], w1t[NV], w2t[NV];
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int k = lane * 4 + 128 * u;
        const bool ok = k < h2;
        xd[u] = ok ? ld4(hd + k) : z4; xe[u] = ok ? ld4(he + k) : z4; xf[u] = ok ? ld4(hf + k) : z4;
        xg[u] = ok ? ld4(hg + k) : z4; xh[u] = ok ? ld4(hh + k) : z4;
        w1[u] = ok ? ld4(W3q1 + k) : z4; w2[u] = ok ? ld4(W3q2 + k) : z4;
        w1t[u] = ok ? ld4(W3q1t + k) : z4; w2t[u] = ok ? ld4(W3q2t + k) : z4;
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        qd = dot4(xd[u], w1[u], qd); qe = dot4(xe[u], w2[u], qe); qf = dot4(xf[u], w1[u], qf);
        qg = dot4(xg[u], w1t[u], qg); qh = dot4(xh[u], w2t[u], qh);
      }
    } else if (vec) {
      for (int k = lane * 4; k < h2; k += 128) {
        const float4 v1 = ld4(W3q1 + k), v2 = ld4(W3q2 + k), v1t = ld4(W3q1t + k), v2t = ld4(W3q2t + k);
        qd = dot4(ld4(hd + k), v1, qd);
        qe = dot4(ld4(he + k), v2, qe);
        qf = dot4(ld4(hf + k), v1, qf);
        qg = dot4(ld4(hg + k), v1t, qg);
        qh = dot4(ld4(hh + k), v2t, qh);
      }
    } else {
      for (int k = lane; k < h2; k += 32) {
        const float v1 = W3q1[k], v2 = W3q2[k];
        qd = fmaf(hd[k], v1, qd);
        qe = fmaf(he[k], v2, qe);
        qf = fmaf(hf[k], v1, qf);
        qg = fmaf(hg[k], W3q1t[k], qg);
        qh = fmaf(hh[k], W3q2t[k], qh);
      }
    }
    const float invB = 1.0f / (float)B;
    const float lp1 = LOGP1[row], rew = R[row], dn = DN[row];
    float lp2;
    if (fold_b) {
      // (its loads are issued while the rows above are still in flight; the reductions below come after)
      float unused;
      lp2 = heads_row_eval(H2b + (size_t)row * h2, h2, ldh, A, Whead, NOISE + ((size_t)B + row) * A, act_scale, s_hout[w], lane,
                           &unused);
      if (lane == 0) LOGP2[row] = lp2;
    } else {
      lp2 = LOGP2[row];
    }
    qd = warp_sum(qd) + W3q1[h2];
    qe = warp_sum(qe) + W3q2[h2];
    qf = warp_sum(qf) + W3q1[h2];
    qg = warp_sum(qg) + W3q1t[h2];
    qh = warp_sum(qh) + W3q2t[h2];
    const float min_q = fminf(qg, qh);
    const float v_backup = __fsub_rn(min_q, __fmul_rn(alpha, lp2));
    const float q_backup = __fadd_rn(rew, __fmul_rn(__fmul_rn(gamma, __fsub_rn(1.0f, dn)), v_backup));
    const float e1 = __fsub_rn(q_backup, qd), e2 = __fsub_rn(q_backup, qe);
    const float dqd = -e1 * invB, dqe = -e2 * invB, dqf = -invB;
    auto dz4 = [](const float4& x, const float4& wv, float dq) {
      return make_float4(x.x > 0.f ? dq * wv.x : 0.f, x.y > 0.f ? dq * wv.y : 0.f, x.z > 0.f ? dq * wv.z : 0.f, x.w > 0.f ? dq * wv.w : 0.f);
    };
    auto put3 = [&](size_t o, const float4& zd, const float4& ze, const float4& zf) {
      if (zlo) { st_split4(dZ2d, zlo, o, zd); st_split4(dZ2e, zlo, o, ze); st_split4(dZ2f, zlo, o, zf); }
      else {
        *reinterpret_cast<float4*>(dZ2d + o) = zd;
        *reinterpret_cast<float4*>(dZ2e + o) = ze;
        *reinterpret_cast<float4*>(dZ2f + o) = zf;
      }
    };
    if (fast) {
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int k = lane * 4 + 128 * u;
        if (k < h2) put3((size_t)row * ldz + k, dz4(xd[u], w1[u], dqd), dz4(xe[u], w2[u], dqe), dz4(xf[u], w1[u], dqf));
      }
    } else if (vec) {
      for (int k = lane * 4; k < h2; k += 128) {
        const float4 v1 = ld4(W3q1 + k), v2 = ld4(W3q2 + k);
        put3((size_t)row * ldz + k, dz4(ld4(hd + k), v1, dqd), dz4(ld4(he + k), v2, dqe), dz4(ld4(hf + k), v1, dqf));
      }
    } else {
      for (int k = lane; k < h2; k += 32) {
        const float v1 = W3q1[k], v2 = W3q2[k];
        const float zd = hd[k] > 0.0f ? dqd * v1 : 0.0f, ze = he[k] > 0.0f ? dqe * v2 : 0.0f, zf = hf[k] > 0.0f ? dqf * v1 : 0.0f;
        const size_t o = (size_t)row * ldz + k;
        if (zlo) { put_split(dZ2d, zlo, o, zd); put_split(dZ2e, zlo, o, ze); put_split(dZ2f, zlo, o, zf); }
        else { dZ2d[o] = zd; dZ2e[o] = ze; dZ2f[o] = zf; }
      }
    }
    if (lane == 0) {
      dQd[row] = dqd; dQe[row] = dqe;
      if (d.out_q1) d.out_q1[row] = qd;
      if (d.out_q2) d.out_q2[row] = qe;
      if (d.out_logp) d.out_logp[row] = lp1;
      t_pi = (double)__fsub_rn(__fmul_rn(alpha, lp1), qf);
      t_q1 = (double)__fmul_rn(e1, e1);
      t_q2 = (double)__fmul_rn(e2, e2);
      t_lp = (double)lp1;
    }
  }
  if (lane == 0) { s_part[w][0] = t_pi; s_part[w][1] = t_q1; s_part[w][2] = t_q2; s_part[w][3] = t_lp; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double acc = 0.0;
    for (int i = 0; i < ROW_WARPS; ++i) acc += s_part[i][threadIdx.x];
    partials[(size_t)vb * 4 + threadIdx.x] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned int)vgrid - 1);
  __syncthreads();
  if (s_last) {
    // the last CTA combines the per-CTA partials: thread t sums CTAs t, t + 256, ... (independent loads), then a
    // fixed-order tree in shared memory — deterministic whatever the order the CTAs finished in
    __threadfence();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (unsigned int i = threadIdx.x; i < (unsigned int)vgrid; i += ROW_WARPS * 32)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] += partials[(size_t)i * 4 + c];
#pragma unroll
    for (int c = 0; c < 4; ++c) s_red[threadIdx.x][c] = acc[c];
    __syncthreads();
    for (int stride = ROW_WARPS * 16; stride > 0; stride >>= 1) {
      if ((int)threadIdx.x < stride)
#pragma unroll
        for (int c = 0; c < 4; ++c) s_red[threadIdx.x][c] += s_red[threadIdx.x + stride][c];
      __syncthreads();
    }
    if (threadIdx.x < 4) {
      const float v = (float)((threadIdx.x == 1 || threadIdx.x == 2 ? 0.5 : 1.0) * s_red[0][threadIdx.x] / B);
      if (threadIdx.x < 3) { SCAL[threadIdx.x] = v; if (d.out_scalars) d.out_scalars[threadIdx.x] = v; }
      else SCAL[4] = v;   // mean logp1 (entropy-alpha gradient; all-reduced across ranks by the host)
      if (threadIdx.x == 0) { SCAL[3] = alpha; if (d.out_scalars) d.out_scalars[3] = alpha; *ticket = 0u; }
    }
  }
}

// gradient of pi_loss = mean(alpha*logp1 - q1_pi) wrt the policy head pre-activations (chain rule of
// the reference op graph), fused with its two skinny neighbours:
//   dA1  = dZ1(Q1(x,pi)) . W1q1[D:D+A, :]^T                 (input gradient of Q1 wrt the action)
//   dZ2a = [dmu | dls] . Whead[0:h2, :]^T  masked by relu'(H2a)
// One warp per row; 128-bit loads when the widths allow (h1 % 4 == 0; head block rows are ldh = 4-float padded).
__global__ void __launch_bounds__(ROW_WARPS * 32) k_policy_bwd_rows(
    const StepState* __restrict__ st, int B, int A, int h1, int h2, int ldh, float act_scale, const float* __restrict__ HDa,
    const float* __restrict__ NOISE, const float* __restrict__ dZ1f, const float* __restrict__ W1q1_act,
    const float* __restrict__ Whead, const float* __restrict__ H2a, float* dHD, float* dZ2a, int ld1, long long lo1, int ldz,
    long long zlo) {
  __shared__ __align__(16) float s_da[ROW_WARPS][MAX_HEAD];
  __shared__ __align__(16) float s_dhd[ROW_WARPS][MAX_HEAD + 4];
  pdl_trigger();
  pdl_wait();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * ROW_WARPS + w;
  if (row >= B) return;
  if (ld1 == 0) ld1 = h1;
  if (ldz == 0) ldz = h2;
  const float* zhi = dZ1f + (size_t)row * ld1;
  const float* zlo1 = lo1 ? zhi + lo1 : nullptr;
  if ((h1 & 3) == 0 && (ld1 & 3) == 0) {
    for (int j0 = 0; j0 < A; j0 += 8) {
      float acc[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0f;
      for (int k = lane * 4; k < h1; k += 128) {
        float4 x = ld4(zhi + k);
        if (zlo1) { const float4 l = ld4(zlo1 + k); x.x += l.x; x.y += l.y; x.z += l.z; x.w += l.w; }   // hi + lo is exact
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (j0 + jj < A) acc[jj] = dot4(x, ld4(W1q1_act + (size_t)(j0 + jj) * h1 + k), acc[jj]);
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float v = warp_sum(acc[jj]);
        if (j0 + jj < A && lane == 0) s_da[w][j0 + jj] = v;
      }
    }
    __syncwarp();
  } else {
    warp_dots(zhi, h1, W1q1_act, h1, A, true, false, s_da[w], lane, zlo1);
  }
  const float dlogp = st->alpha_cur / (float)B;
  const float* hd = HDa + (size_t)row * 2 * A;
  const float* eps = NOISE + (size_t)row * A;
  for (int j = lane; j < ldh; j += 32) s_dhd[w][j] = 0.0f;   // padding columns of the head block multiply zeros
  __syncwarp();
  for (int j = lane; j < A; j += 32) {
    const float mu = hd[j];
    const PolEl e = policy_elem(mu, hd[A + j], eps[j]);
    // logp -= log(clip_pass(1 - pi^2) + 1e-6):  d/dpi = +2 pi / (clipped + 1e-6) * dlogp
    const float dpi = act_scale * s_da[w][j] + dlogp * (2.0f * e.pi) / (e.clipped + 1e-6f);
    float du = dpi * e.omp;                          // tanh'(u) = 1 - pi^2 (unclipped, as TF's TanhGrad)
    const float dz = -dlogp * e.z;                   // gaussian term: pre = -0.5 (z^2 + 2 log_std + c)
    const float ddiff = dz / e.den;                  // d(pi_raw - mu)
    du += ddiff;
    const float dstd_den = -dz * e.z / e.den;        // through the denominator exp(log_std)+EPS
    const float dmu = du - ddiff;                    // pi_raw path + (x - mu) path
    const float dstd = du * eps[j] + dstd_den;       // pi_raw = mu + eps*std
    const float dlog_std = dstd * e.std - dlogp;     // exp' and the direct 2*log_std term
    const float dls = 11.0f * dlog_std * (1.0f - e.ls_t * e.ls_t);
    s_dhd[w][j] = dmu; s_dhd[w][A + j] = dls;
    dHD[(size_t)row * 2 * A + j] = dmu;
    dHD[(size_t)row * 2 * A + A + j] = dls;
  }
  __syncwarp();
  for (int n = lane; n < h2; n += 32) {
    const float* wr = Whead + (size_t)n * ldh;
    float a0 = 0.0f, a1 = 0.0f;
    for (int j = 0; j < ldh; j += 8) {               // ldh is a multiple of 4; rows are 16-byte aligned
      a0 = dot4(ld4(s_dhd[w] + j), ld4(wr + j), a0);
      if (j + 4 < ldh) a1 = dot4(ld4(s_dhd[w] + j + 4), ld4(wr + j + 4), a1);
    }
    const float z = H2a[(size_t)row * h2 + n] > 0.0f ? a0 + a1 : 0.0f;
    if (zlo) put_split(dZ2a, zlo, (size_t)row * ldz + n, z);
    else dZ2a[(size_t)row * ldz + n] = z;
  }
}

// ------------------------------------------------------------------------------------------------
// Actor inference (Actor.get_action, algos/sac1/actor_learner.py:195-197): mu / pi of the MAIN policy
// for n observations — one warp per observation, the whole 2-layer MLP + heads in one kernel (n is 1
// per env step in the reference; vectorised rollouts pass n = number of envs).
// ------------------------------------------------------------------------------------------------
constexpr int ACT_MAX_H = 512;
__global__ void __launch_bounds__(128) k_actor_forward(int n, int D, int A, int h1, int h2, int ldh, float act_scale, int deterministic,
                                                       const float* __restrict__ OBS, const float* __restrict__ W1,
                                                       const float* __restrict__ W2, const float* __restrict__ Whead,
                                                       const float* __restrict__ noise, unsigned long long seed,
                                                       unsigned long long counter, float* OUT) {
  __shared__ float s_h1[4][ACT_MAX_H];
  __shared__ float s_h2[4][ACT_MAX_H];
  __shared__ float s_out[4][MAX_HEAD];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + w;
  if (row >= n) return;
  const float* x = OBS + (size_t)row * D;
  for (int j = lane; j < h1; j += 32) {          // W1: [D+1, h1], bias = last row
    float acc = W1[(size_t)D * h1 + j];
    for (int k = 0; k < D; ++k) acc = fmaf(x[k], W1[(size_t)k * h1 + j], acc);
    s_h1[w][j] = fmaxf(acc, 0.0f);
  }
  __syncwarp();
  for (int j = lane; j < h2; j += 32) {
    float acc = W2[(size_t)h1 * h2 + j];
    for (int k = 0; k < h1; ++k) acc = fmaf(s_h1[w][k], W2[(size_t)k * h2 + j], acc);
    s_h2[w][j] = fmaxf(acc, 0.0f);
  }
  __syncwarp();
  warp_dots(s_h2[w], h2, Whead, ldh, 2 * A, false, true, s_out[w], lane);
  for (int j = lane; j < A; j += 32) {
    float eps = 0.0f;
    if (!deterministic) {
      if (noise) eps = noise[(size_t)row * A + j];
      else {
        const unsigned long long e = (unsigned long long)row * A + j;
        const Philox4 p = philox4x32_10((uint32_t)(e >> 2), (uint32_t)counter, (uint32_t)(counter >> 32), 0xAC70u,
                                        (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t ua = (e & 2) ? p.z : p.x, ub = (e & 2) ? p.w : p.y;
        const float u0 = ((float)ua + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)ub + 0.5f) * 2.3283064365386963e-10f;
        const float r = sqrtf(-2.0f * logf(fmaxf(u0, 1e-30f)));
        float sn, cs;
        sincospif(2.0f * u1, &sn, &cs);
        eps = (e & 1) ? r * sn : r * cs;
      }
    }
    const float mu = s_out[w][j];
    const PolEl e = policy_elem(mu, s_out[w][A + j], eps);
    OUT[(size_t)row * A + j] = __fmul_rn(deterministic ? tanhf(mu) : e.pi, act_scale);
  }
}

// ------------------------------------------------------------------------------------------------
// optimiser: TF1 Adam (epsilon-hat form) for pi and q parameter ranges, then polyak with the new
// weights (actor_learner.py:73-87); gradients arrive as S split-K partials (S = 1 after all-reduce).
// ------------------------------------------------------------------------------------------------
// hi/lo planes of the weight blocks the tensor-core GEMMs read (kernel rows only; the bias row is added in
// the GEMM epilogue from the fp32 master copy): block b = rows x N[b] at sp_off[b], row pitch pitch[b]
struct SplitMap {
  int nblk;
  int N[6], pitch[6];
  long long int_off[6], size[6], sp_off[6];
  long long plane;
  int vec4;        // every N[b] is a multiple of 4: the optimiser may treat 4 consecutive parameters as one row segment
};
__device__ __forceinline__ void write_split(const SplitMap& mp, int64_t i, float w, float* __restrict__ Wsp) {
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    if (b < mp.nblk && i >= mp.int_off[b] && i < mp.int_off[b] + mp.size[b]) {
      const int64_t e = i - mp.int_off[b];
      const int64_t r = e / mp.N[b], c = e - r * mp.N[b];
      put_split(Wsp, mp.plane, (size_t)(mp.sp_off[b] + r * mp.pitch[b] + c), w);
    }
  }
}
// four consecutive parameters starting at a multiple of 4 (mp.vec4: every block width is a multiple of 4, so they sit in
// one row of one block): one 128-bit store per plane
__device__ __forceinline__ void write_split4(const SplitMap& mp, int64_t i, const float4& w, float* __restrict__ Wsp) {
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    if (b < mp.nblk && i >= mp.int_off[b] && i < mp.int_off[b] + mp.size[b]) {
      const int e = (int)(i - mp.int_off[b]);
      const int r = e / mp.N[b], c = e - r * mp.N[b];
      st_split4(Wsp, mp.plane, (size_t)(mp.sp_off[b] + (long long)r * mp.pitch[b] + c), w);
    }
  }
}
__global__ void __launch_bounds__(256) k_split_weights(SplitMap mp, int64_t P, const float* __restrict__ W,
                                                       const float* __restrict__ Wt, float* Wsp, float* Wtsp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
    write_split(mp, i, W[i], Wsp);
    write_split(mp, i, Wt[i], Wtsp);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of a NARROW first layer (policy W1: K = D <= 32 input features): d[W;b][k, n] = sum_r [x|1][r, k] dZ[r, n].
// As a tensor-core stage this was one more dependent launch of ~10 us for 13 MFLOP at the very end of the backward
// chain; here it is FFMA: grid = (32-column chunks of n) x (64-row slices of the batch).  lane = column, warp w owns
// rows w, w + 8, ... of the slice and ALL K + 1 outputs of its column (K + 1 accumulators per thread, 2 loads per
// 34 FMAs: a load -> few-FMA loop is serialised on the L2 latency by the compiler, measured 47 us); the eight warps'
// partial sums are combined in shared memory in warp order (deterministic).  The slice's [x|1] rows are staged in
// shared memory (hi + lo is exact).  Each slice writes its own partial block; the optimiser sums the slices in index
// order, see NarrowGrad.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_nc_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
constexpr int NW_ROWS = 64;       // batch rows per slice
constexpr int NW_MAXK = 33;       // K + 1 (bias row) <= 33
__global__ void __launch_bounds__(256) k_wgrad_narrow(int B, int K, int N, const float* __restrict__ X, int ldx, long long xlo,
                                                      const float* __restrict__ Z, int ldz, long long zlo, float* Gn) {
  __shared__ float s_x[NW_ROWS][NW_MAXK + 1];
  __shared__ float s_acc[8][NW_MAXK][33];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * NW_ROWS, nr = min(NW_ROWS, B - r0);
  constexpr int RW = NW_ROWS / 8;             // rows per warp
  // this thread's dZ values first (hi + lo planes), so that they are in flight while [x|1] is staged
  float zh[RW], zl[RW];
  {
    const float* zr = Z + (size_t)r0 * ldz + min(col, N - 1);
#pragma unroll
    for (int u = 0; u < RW; ++u) {
      const int ru = min(w + 8 * u, nr - 1);  // rows beyond the slice re-read its last row and are dropped below
      zh[u] = ld_nc_f32(zr + (size_t)ru * ldz);
      zl[u] = ld_nc_f32(zr + zlo + (size_t)ru * ldz);
    }
  }
  for (int i = threadIdx.x; i < NW_ROWS * (NW_MAXK + 1); i += 256) {
    const int r = i / (NW_MAXK + 1), k = i - r * (NW_MAXK + 1);
    float v = 0.0f;
    if (r < nr && k <= K) v = k < K ? X[(size_t)(r0 + r) * ldx + k] + X[xlo + (size_t)(r0 + r) * ldx + k] : 1.0f;
    s_x[r][k] = v;
  }
  __syncthreads();
  float acc[NW_MAXK];
#pragma unroll
  for (int k = 0; k < NW_MAXK; ++k) acc[k] = 0.0f;
#pragma unroll
  for (int u = 0; u < RW; ++u) {
    const int r = w + 8 * u;
    const float z = r < nr ? zh[u] + zl[u] : 0.0f;
#pragma unroll
    for (int k = 0; k < NW_MAXK; ++k) acc[k] = fmaf(s_x[r][k], z, acc[k]);     // columns k > K of s_x are zero
  }
#pragma unroll
  for (int k = 0; k < NW_MAXK; ++k) s_acc[w][k][lane] = acc[k];
  __syncthreads();
  float* out = Gn + (size_t)blockIdx.y * (size_t)(K + 1) * N;
  for (int i = threadIdx.x; i < (K + 1) * 32; i += 256) {
    const int k = i >> 5, l = i & 31;
    float t = s_acc[0][k][l];
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) t += s_acc[ww][k][l];
    if (blockIdx.x * 32 + l < N) out[(size_t)k * N + blockIdx.x * 32 + l] = t;
  }
}
// the parameter range [off, off + size) whose gradient arrives as SN slices from k_wgrad_narrow instead of split-K partials
struct NarrowGrad {
  const float* Gn;
  long long off, size;
  int SN;
};
__device__ __forceinline__ float4 grad4(int64_t i, int64_t P, int S, const float* __restrict__ Gp, const NarrowGrad& ng) {
  if (ng.SN > 0 && i >= ng.off && i < ng.off + ng.size) {
    // slices summed in index order; eight loads are in flight at a time (a load -> add loop is one L2 latency per slice)
    const float* p = ng.Gn + (i - ng.off);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < ng.SN; s0 += 8) {
      float4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        q[u] = s0 + u < ng.SN ? *reinterpret_cast<const float4*>(p + (size_t)(s0 + u) * ng.size) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { g.x += q[u].x; g.y += q[u].y; g.z += q[u].z; g.w += q[u].w; }
    }
    return g;
  }
  float4 g = *reinterpret_cast<const float4*>(Gp + i);
  for (int s = 1; s < S; ++s) {
    const float4 q = *reinterpret_cast<const float4*>(Gp + (size_t)s * P + i);
    g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
  }
  return g;
}
__global__ void __launch_bounds__(256) k_grad_reduce(int64_t P, int S, const float* __restrict__ Gp, float* G, NarrowGrad ng) {
  pdl_trigger();
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < P; i += stride)
    *reinterpret_cast<float4*>(G + i) = grad4(i, P, S, Gp, ng);
}

// one Adam + polyak update of 4 consecutive parameters (i % 4 == 0; P_pi % 4 == 0 so one learning rate applies)
__device__ __forceinline__ void adam4(int64_t i, const float4& g4, float gs, float lr_i, float polyak, float* W, float* Wt, float* Mo,
                                      float* Vo, const SplitMap* mp, float* Wsp, float* Wtsp) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float4 m4 = *reinterpret_cast<const float4*>(Mo + i), v4 = *reinterpret_cast<const float4*>(Vo + i);
  const float4 w4 = *reinterpret_cast<const float4*>(W + i), t4 = *reinterpret_cast<const float4*>(Wt + i);
  const float gv[4] = {g4.x * gs, g4.y * gs, g4.z * gs, g4.w * gs};
  const float mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w},
              tv[4] = {t4.x, t4.y, t4.z, t4.w};
  float mo[4], vo[4], wo[4], to[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    mo[e] = b1 * mv[e] + (1.0f - b1) * gv[e];
    vo[e] = b2 * vv[e] + (1.0f - b2) * gv[e] * gv[e];
    wo[e] = wv[e] - lr_i * mo[e] / (sqrtf(vo[e]) + eps);
    to[e] = polyak * tv[e] + (1.0f - polyak) * wo[e];
  }
  const float4 wn = make_float4(wo[0], wo[1], wo[2], wo[3]), tn = make_float4(to[0], to[1], to[2], to[3]);
  *reinterpret_cast<float4*>(Mo + i) = make_float4(mo[0], mo[1], mo[2], mo[3]);
  *reinterpret_cast<float4*>(Vo + i) = make_float4(vo[0], vo[1], vo[2], vo[3]);
  *reinterpret_cast<float4*>(W + i) = wn;
  *reinterpret_cast<float4*>(Wt + i) = tn;
  if (mp) {
    if (mp->vec4) { write_split4(*mp, i, wn, Wsp); write_split4(*mp, i, tn, Wtsp); }
    else {
#pragma unroll
      for (int e = 0; e < 4; ++e) { write_split(*mp, i + e, wo[e], Wsp); write_split(*mp, i + e, to[e], Wtsp); }
    }
  }
}
// entropy-alpha (reference-intended semantics): one scalar Adam step, unordered wrt the rest; uses the
// pre-update mean(logp1) of this step
__device__ __forceinline__ void alpha_step(StepState* st, float lr, float mean_logp, float target_entropy) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  st->t_alpha += 1;
  const double ta = (double)st->t_alpha;
  const float lr_a = (float)((double)lr * sqrt(1.0 - pow((double)b2, ta)) / (1.0 - pow((double)b1, ta)));
  const float g = -(mean_logp + target_entropy);
  st->alpha_m = b1 * st->alpha_m + (1.0f - b1) * g;
  st->alpha_v = b2 * st->alpha_v + (1.0f - b2) * g * g;
  st->log_alpha -= lr_a * st->alpha_m / (sqrtf(st->alpha_v) + eps);
}
__device__ __forceinline__ void d_adam_polyak(int vb, int vgrid, StepState* st, int64_t P, int64_t P_pi, int S,
                                                     const float* __restrict__ Gp, float lr, float polyak,
                                                     float target_entropy, const float* __restrict__ SCAL,
                                                     float* W, float* Wt, float* Mo, float* Vo,
                                                     const NarrowGrad& ng, const SplitMap* mp = nullptr, float* Wsp = nullptr,
                                                     float* Wtsp = nullptr) {
  const float lr_pi = st->lr_pi, lr_q = st->lr_q;
  const float gs = st->dyn.grad_scale;
  const int64_t stride = (int64_t)vgrid * blockDim.x * 4;
  for (int64_t i = ((int64_t)vb * blockDim.x + threadIdx.x) * 4; i < P; i += stride)      // P, P_pi are multiples of 4
    adam4(i, grad4(i, P, S, Gp, ng), gs, i < P_pi ? lr_pi : lr_q, polyak, W, Wt, Mo, Vo, mp, Wsp, Wtsp);
  if (vb == 0 && threadIdx.x == 0 && st->auto_alpha) alpha_step(st, lr, SCAL[4], target_entropy);
}
__global__ void __launch_bounds__(256) k_adam_polyak(StepState* st, int64_t P, int64_t P_pi, int S, const float* __restrict__ Gp,
                                                     float lr, float polyak, float target_entropy, const float* __restrict__ SCAL,
                                                     float* W, float* Wt, float* Mo, float* Vo, NarrowGrad ng) {
  pdl_trigger();
  pdl_wait();
  d_adam_polyak(blockIdx.x, gridDim.x, st, P, P_pi, S, Gp, lr, polyak, target_entropy, SCAL, W, Wt, Mo, Vo, ng);
}
__global__ void __launch_bounds__(256) k_adam_polyak_split(StepState* st, int64_t P, int64_t P_pi, int S, const float* __restrict__ Gp,
                                                           float lr, float polyak, float target_entropy,
                                                           const float* __restrict__ SCAL, float* W, float* Wt, float* Mo, float* Vo,
                                                           const __grid_constant__ SplitMap mp, float* Wsp, float* Wtsp,
                                                           NarrowGrad ng) {
  pdl_trigger();
  pdl_wait();
  d_adam_polyak(blockIdx.x, gridDim.x, st, P, P_pi, S, Gp, lr, polyak, target_entropy, SCAL, W, Wt, Mo, Vo, ng, &mp, Wsp, Wtsp);
}

// ------------------------------------------------------------------------------------------------
// Data-parallel learners without a collective library call: gradient all-reduce FUSED into the optimiser over
// NVLink peer memory.  Every rank owns one communication buffer [2][Pc] floats (+ 8 arrival flags) that all peers
// map with CUDA IPC.  Step t: k_grad_reduce_comm leaves this rank's flat gradient (split-K partials summed, plus
// the entropy statistic) in slot t & 1; k_adam_polyak_peer then
//   1. (the reduce kernel's last CTA has already stored t into flag[rank] of every peer, st.release.sys)
//   2. waits until its own flags show t from every peer (ld.acquire.sys),
//   3. reads all ranks' slot t & 1 directly (128-bit volatile loads over NVLink), sums them in rank order — every
//      rank computes bit-identical sums — scales by 1/N and applies Adam + polyak (+ the weight split planes).
// Slots alternate, so a slot is rewritten at step t + 2 only after every peer has passed the barrier of step
// t + 1, i.e. finished reading it: one flag exchange per step is the only synchronisation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Poll with RELAXED loads and take one acquire fence when the flag is seen: an ld.acquire.sys in the spin loop
// invalidates the SM's L1 on every iteration, which slowed the kernels running next to a waiting exchange kernel
// threefold (measured on the narrow-W1 kernel: 5 -> 16 us).  The data behind the flags is read with volatile loads.
// Wait until every rank has published `epoch`.  ONE CTA polls the peer-written flags at system scope and re-publishes the
// epoch on a local word at GPU scope; every other CTA waits on that word.  Measured (tools/probes/flag_probe.cu, 2 B200,
// 216 CTAs): 6.9 us per exchange round against 18.5 us when every CTA polled the remote-written flags and issued its own
// fence.acq_rel.sys — system-scope fences are ~1.6 us each and serialise when hundreds are in flight.  CTA 0 never waits
// on another CTA, and the launches that call this keep the whole grid resident.
__device__ __forceinline__ void wait_flags(const unsigned int* flags, unsigned int* relay, int world, unsigned int epoch, int* err) {
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < world) {
      const unsigned int* f = flags + threadIdx.x;
      const long long t0 = clock64();
      unsigned int v;
      for (;;) {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 20000000000LL) { *err = 1; break; }   // ~10 s: a peer died; do not hang the GPU
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("fence.acq_rel.sys;" ::: "memory");              // acquire side of the peers' st.release.sys
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(relay), "r"(epoch) : "memory");
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    unsigned int v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(relay) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 20000000000LL) { *err = 1; break; }
    }
  }
  __syncthreads();
}
struct PeerComm {
  float* buf[8];            // rank r's communication buffer as mapped in this process
  unsigned int* flags[8];   // rank r's arrival flags: flags[r][i] = last step whose gradient rank i has published
  int world, rank, nslice;
  long long Pc;             // floats per slot (P + 4, multiple of 4)
  unsigned int* relay;      // local word: the last epoch whose flags CTA 0 has seen (wait_flags)
  const float* mc;          // nullable: NVLS multicast mapping of the same buffers (ddrl_sac_comm_attach_ptrs): one
                            // multimem.ld_reduce returns the sum over all ranks, added inside the NVSwitch
};
// sum over all ranks of the 4 floats at the same offset of every rank's buffer, reduced in the switch
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__global__ void __launch_bounds__(256) k_grad_reduce_comm(const StepState* __restrict__ st, int64_t P, int S,
                                                          const float* __restrict__ Gp, const float* __restrict__ SCAL,
                                                          const __grid_constant__ PeerComm pc, unsigned int* ticket, NarrowGrad ng) {
  pdl_trigger();
  pdl_wait();
  step_stamp(st, 0);
  float* G = pc.buf[pc.rank] + (size_t)(st->t_pi & 1) * pc.Pc;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < P; i += stride)
    *reinterpret_cast<float4*>(G + i) = grad4(i, P, S, Gp, ng);
  if (blockIdx.x == 0 && threadIdx.x == 0) G[P] = SCAL[4];   // mean logp1 of this rank's batch (entropy-alpha gradient)
  // the last CTA to finish publishes "my gradient of step t is complete" to every peer, so the flags travel while
  // the optimiser kernel is being launched
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x < pc.world) {
    if (threadIdx.x == 0) *ticket = 0u;
    // release at system scope: cumulative over every CTA's writes (each fenced before its ticket increment); one fence
    // per publishing thread, no separate __threadfence_system()
    const unsigned int epoch = (unsigned int)st->t_pi;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pc.flags[threadIdx.x] + pc.rank), "r"(epoch) : "memory");
    if (threadIdx.x == 0 && st->trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); st->trace[1] = t; }
  }
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__global__ void __launch_bounds__(256) k_adam_polyak_peer(StepState* st, int64_t P, int64_t P_pi, float lr, float polyak,
                                                          float target_entropy, float* W, float* Wt, float* Mo, float* Vo,
                                                          const __grid_constant__ SplitMap mp, float* Wsp, float* Wtsp,
                                                          const __grid_constant__ PeerComm pc, int* err) {
  pdl_trigger();
  pdl_wait();
  step_stamp(st, 2);
  const unsigned int epoch = (unsigned int)st->t_pi;
  wait_flags(pc.flags[pc.rank], pc.relay, pc.world, epoch, err);     // (this rank's own flag was published by k_grad_reduce_comm)
  step_stamp(st, 3);
  const float lr_pi = st->lr_pi, lr_q = st->lr_q;
  const float gs = st->dyn.grad_scale;
  const size_t slot = (size_t)(epoch & 1u) * pc.Pc;
  const int world = pc.world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < P; i += stride) {
    // all peers' loads are issued before the first add: one NVLink round trip, not `world` of them
    float4 q[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < world) q[r] = ld_volatile_f4(pc.buf[r] + slot + i);
    float4 g = q[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (r < world) { g.x += q[r].x; g.y += q[r].y; g.z += q[r].z; g.w += q[r].w; }
    adam4(i, g, gs, i < P_pi ? lr_pi : lr_q, polyak, W, Wt, Mo, Vo, Wsp ? &mp : nullptr, Wsp, Wtsp);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && st->auto_alpha) {
    float lp = 0.0f;
    for (int r = 0; r < world; ++r) lp += *reinterpret_cast<volatile const float*>(pc.buf[r] + slot + P);
    alpha_step(st, lr, lp * gs, target_entropy);          // mean over the global batch (equal batch per rank)
  }
  step_stamp(st, 4);      // (CTA 0's end: the other CTAs run the same loop length)
}

// ------------------------------------------------------------------------------------------------
// Data-parallel step, default form: the two-kernel form above as ONE kernel (the split-K reduction is folded in, which
// removes a launch from the critical path): every CTA sums its share of the split-K partials (+ the narrow-W1 slices)
// into this rank's exchange slot; the last CTA to finish publishes the flag on every peer; every CTA waits for all
// ranks' flags, reads all ranks' slots (128-bit volatile loads over NVLink, all issued before the first add), sums in
// rank order — every rank computes the same bits, replicas stay bit-identical — and applies Adam + polyak.
// What was tried instead and measured slower on 2 / 4 / 8 B200 (C2; single-GPU step 96 us):
//   * reduce-scatter / all-gather inside the kernel (1/N of the data, but a second flag round): every flag round a
//     kernel has to WAIT for costs ~5 us at 2 ranks and ~10 us at 8 (publish -> NVLink -> poll + the skew between
//     ranks): 130 vs 124 us per step at N = 8, 130 vs 117 at N = 4;
//   * the same exchange in a kernel on the side stream under the end of the backward: a kernel that SPINS next to
//     others slows them (the narrow-W1 kernel went from 5 to 16 us): 133 us at N = 8;
//   * publishing everything but the policy's first layer early from a non-waiting side-stream kernel and handling that
//     late block last: each of the small dependent phases costs 3-5 us of latency: 124 vs 112 us at N = 2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam_dp(StepState* st, int64_t P, int64_t P_pi, int S, const float* __restrict__ Gp,
                                                 NarrowGrad ng, const float* __restrict__ SCAL, float lr, float polyak,
                                                 float target_entropy, float* W, float* Wt, float* Mo, float* Vo,
                                                 const __grid_constant__ SplitMap mp, float* Wsp, float* Wtsp,
                                                 const __grid_constant__ PeerComm pc, unsigned int* ticket, int* err,
                                                 unsigned long long* trace) {
  __shared__ bool s_last;
  pdl_trigger();
  pdl_wait();
  auto stamp = [&](int i) { if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[i] = tc::gtime(); };
  stamp(0);
  const unsigned int epoch = (unsigned int)st->t_pi;
  const int world = pc.world;
  const size_t slot = (size_t)(epoch & 1u) * pc.Pc;   // slots alternate: a slot is rewritten at step t + 2, after every
                                                      // peer's flag of step t + 1, i.e. after it has finished reading step t
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (int64_t)gridDim.x * blockDim.x;
  float* part = pc.buf[pc.rank] + slot;
  for (int64_t i = gtid * 4; i < P; i += gthreads * 4) *reinterpret_cast<float4*>(part + i) = grad4(i, P, S, Gp, ng);
  if (gtid == 0) part[P] = SCAL[4];                 // mean logp1 of this rank's batch (entropy-alpha gradient)
  stamp(1);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && (int)threadIdx.x < world) {
    if (threadIdx.x == 0) *ticket = 0u;
    st_release_sys(pc.flags[threadIdx.x] + pc.rank, epoch);   // release: cumulative over the CTAs' writes above
  }
  wait_flags(pc.flags[pc.rank], pc.relay, world, epoch, err);
  stamp(2);
  const float lr_pi = st->lr_pi, lr_q = st->lr_q;
  const float gs = st->dyn.grad_scale;
  const SplitMap* mpp = Wsp ? &mp : nullptr;
  if (pc.mc) {
    // NVLS: ONE load per 16 bytes; the switch reads every rank's slot and adds (traffic per rank: P instead of (N-1) P)
    for (int64_t i = gtid * 4; i < P; i += gthreads * 4)
      adam4(i, mc_ld_reduce_f4(pc.mc + slot + i), gs, i < P_pi ? lr_pi : lr_q, polyak, W, Wt, Mo, Vo, mpp, Wsp, Wtsp);
  } else {
    for (int64_t i = gtid * 4; i < P; i += gthreads * 4) {
      float4 q[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < world) q[r] = ld_volatile_f4(pc.buf[r] + slot + i);
      float4 g = q[0];
#pragma unroll
      for (int r = 1; r < 8; ++r)
        if (r < world) { g.x += q[r].x; g.y += q[r].y; g.z += q[r].z; g.w += q[r].w; }
      adam4(i, g, gs, i < P_pi ? lr_pi : lr_q, polyak, W, Wt, Mo, Vo, mpp, Wsp, Wtsp);
    }
  }
  if (gtid == 0 && st->auto_alpha) {
    float lp = 0.0f;
    for (int r = 0; r < world; ++r) lp += *reinterpret_cast<volatile const float*>(pc.buf[r] + slot + P);
    alpha_step(st, lr, lp * gs, target_entropy);      // mean over the global batch (equal batch per rank)
  }
  stamp(3);
}

// external (TF variable order: kernel, bias per dense layer; mu head then log_std head) <-> internal
// flat layout.  Internal blocks start on 16-byte boundaries (vector loads in the GEMM operand fetch)
// and the policy head block is fused: internal [h2+1, ldh] = [Wmu|Wls|0.. ; bmu|bls|0..], ldh = 2A rounded up to 4.
struct LayoutMap {
  int nblk, head_idx, h2, A, ldh;   // ldh: row pitch of the fused head block (2A rounded up to 4 floats)
  long long ext_off[9], int_off[9], size[9];
};
__global__ void __launch_bounds__(256) k_convert_layout(LayoutMap mp, int64_t Pext, int to_internal,
                                                        const float* __restrict__ src, float* dst, float* dst2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Pext; i += stride) {
    int b = 0;
    while (b + 1 < mp.nblk && i >= mp.ext_off[b + 1]) ++b;
    int64_t e = i - mp.ext_off[b];
    if (b == mp.head_idx) {   // external: Wmu [h2,A], bmu [A], Wls [h2,A], bls [A]
      const int64_t half = (int64_t)mp.h2 * mp.A + mp.A;
      const int which = e >= half;
      if (which) e -= half;
      const int64_t r = e / mp.A, c = e % mp.A;  // r == h2 is the bias row
      e = r * mp.ldh + which * mp.A + c;
    }
    const int64_t j = mp.int_off[b] + e;
    if (to_internal) { const float v = src[i]; dst[j] = v; if (dst2) dst2[j] = v; }
    else dst[i] = src[j];
  }
}

// ------------------------------------------------------------------------------------------------
// Bias gradients in tensor-core mode: d b = column sums of dZ (hi + lo planes) over the batch rows of each
// split-K slice.  One CTA per (32-column chunk, split): lane = column (128-byte coalesced rows), the 8 warps
// stride the rows, fixed-order combine in shared memory (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int COLSUM_MAX = 6;
struct ColsumGroup {
  int nprob, B, kps;
  long long split_stride;
  struct {
    const float* z;
    long long lo;
    float* out;
    int ld, N, chunk_begin;
  } p[COLSUM_MAX];
};
__global__ void __launch_bounds__(256) k_colsum(const __grid_constant__ ColsumGroup g) {
  __shared__ float s_acc[8][32];
  int pi = 0;
  while (pi + 1 < g.nprob && (int)blockIdx.x >= g.p[pi + 1].chunk_begin) ++pi;
  const float* z = g.p[pi].z;
  const long long lo = g.p[pi].lo;
  const int ld = g.p[pi].ld, N = g.p[pi].N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = (blockIdx.x - g.p[pi].chunk_begin) * 32 + lane;
  const int r0 = blockIdx.y * g.kps, r1 = min(g.B, r0 + g.kps);
  float acc = 0.0f;
  if (col < N)
    for (int r = r0 + w; r < r1; r += 8) acc += z[(size_t)r * ld + col] + z[lo + (size_t)r * ld + col];
  s_acc[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < N) {
    float t = s_acc[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += s_acc[i][lane];
    g.p[pi].out[(size_t)blockIdx.y * g.split_stride + col] = t;
  }
}

// Skinny weight gradients in tensor-core mode (Q heads: N = 1, policy heads: N = 2A): d[W;b][f, j] =
// sum_b [act|1][b, f] * z[b, j] over the batch rows of each split-K slice.  One CTA per (32-feature chunk, split):
// lane = feature (128-byte coalesced rows of act), z[b, :] is a broadcast load, the 8 warps stride the rows and
// combine in shared memory in fixed order (deterministic).
constexpr int SKINNY_MAX = 4;
struct SkinnyGroup {
  int nprob, B, kps;
  long long split_stride;
  struct {
    const float* act;   // [B, K] plain fp32, pitch ld_act; feature K is the constant-one column (bias row)
    const float* z;     // [B, n], pitch ldz
    float* out;         // [K + 1, ldc]
    int ld_act, K, ldz, n, ldc, chunk_begin;
  } p[SKINNY_MAX];
};
template <int NJ>   // NJ >= n: outputs accumulated per thread in one pass over the rows
__device__ __forceinline__ void skinny_body(const SkinnyGroup& g, int pi, float (*s_acc)[8][33]) {
  const float* act = g.p[pi].act;
  const float* z = g.p[pi].z;
  const int K = g.p[pi].K, ld_act = g.p[pi].ld_act, ldz = g.p[pi].ldz, n = g.p[pi].n, ldc = g.p[pi].ldc;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int f = (blockIdx.x - g.p[pi].chunk_begin) * 32 + lane;
  const int r0 = blockIdx.y * g.kps, r1 = min(g.B, r0 + g.kps);
  float* out = g.p[pi].out + (size_t)blockIdx.y * g.split_stride;
  float acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = 0.0f;
  if (f <= K) {
    for (int r = r0 + w; r < r1; r += 8) {
      const float a = f < K ? act[(size_t)r * ld_act + f] : 1.0f;
      const float* zr = z + (size_t)r * ldz;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j < n) acc[j] = fmaf(a, zr[j], acc[j]);
    }
  }
#pragma unroll
  for (int j0 = 0; j0 < NJ; j0 += 8) {
    if (j0 >= n) break;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) s_acc[w][jj][lane] = acc[j0 + jj];
    __syncthreads();
    if (j0 + w < n && f <= K) {       // warp w sums output j0 + w over the 8 warps, in fixed order
      float t = s_acc[0][w][lane];
#pragma unroll
      for (int i = 1; i < 8; ++i) t += s_acc[i][w][lane];
      out[(size_t)f * ldc + j0 + w] = t;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_skinny_wgrad(const __grid_constant__ SkinnyGroup g) {
  __shared__ float s_acc[8][8][33];
  int pi = 0;
  while (pi + 1 < g.nprob && (int)blockIdx.x >= g.p[pi + 1].chunk_begin) ++pi;
  const int n = g.p[pi].n;   // block-uniform
  if (n <= 8) skinny_body<8>(g, pi, s_acc);
  else if (n <= 24) skinny_body<24>(g, pi, s_acc);
  else skinny_body<MAX_HEAD>(g, pi, s_acc);
}

}  // namespace ddrl

// =================================================================================================
// host side
// =================================================================================================
using namespace ddrl;

namespace {

struct Group {
  std::vector<GemmProb> probs;          // FFMA tiles (cfg 0/1)
  std::vector<tc::TcProb> probs_tc;     // tcgen05 tiles
  std::vector<tc::FusedProb> probs_fz;  // tcgen05 fused first + second layer tiles
  GemmGroup grp{};                      // the same, packed as kernel parameters
  tc::TcGroup grp_tc{};
  tc::FusedGroup grp_fz{};
  int tiles = 0, tiles_tc = 0, tiles_fz = 0;
  int bn = 128;                         // tile width of the tcgen05 launch (64 when 128-wide tiles cannot fill the chip)
};

struct Plan {
  int B = 0, S = 1;
  std::vector<Group> fwd1, fwd2, bwd;  // see build_plan
  std::vector<std::vector<Group>> stages;  // ordered stages; element-wise kernels sit between them
  cudaGraphExec_t exec_full = nullptr, exec_grads = nullptr, exec_apply = nullptr, exec_dp = nullptr;
  int64_t kernels[4] = {0, 0, 0, 0};  // kernels inside each captured graph (for the launch counter)
  // The prologue is the graph's FIRST node: its per-step values travel as kernel parameters, patched into the
  // instantiated graph before every launch (cudaGraphExecKernelNodeSetParams).  Launched on its own in front of the graph
  // it cost two stream-order hand-overs per step instead of one (measured with tools/probes/tick_probe.cu: ~1.6 us and a
  // 2.05 us completion granularity per stream-order dependent launch, against 0.75 us between the nodes of a graph).
  cudaGraph_t graph[4] = {nullptr, nullptr, nullptr, nullptr};       // kept alive: the node handles below belong to them
  cudaGraphNode_t prologue_node[4] = {nullptr, nullptr, nullptr, nullptr};
  bool fused = false;              // first + second layer of every forward pass in ONE tcgen05 launch (fwd_fused_tc)
  bool fold_b = false;             // pass b's policy head is evaluated inside k_qheads_losses (narrow heads)
  ColsumGroup colsum[8] = {};      // tensor-core mode: bias-gradient column sums per stage (side stream)
  int colsum_chunks[8] = {};
  SkinnyGroup skinny[8] = {};      // tensor-core mode: skinny weight gradients per stage (side stream)
  int skinny_chunks[8] = {};
};

}  // namespace

struct ddrl_sac {
  int device = 0, D = 0, A = 0, h1 = 0, h2 = 0, maxB = 0, sms = 148;
  float gamma = 0.99f, polyak = 0.995f, lr = 1e-3f, alpha = 0.2f, act_scale = 1.0f;
  int auto_alpha = 0;
  int64_t P = 0, P_pi = 0, Pext = 0;   // P: internal (padded) float count; Pext: the reference's parameter count
  LayoutMap map{};
  // offsets of the [K+1,N] blocks in the flat buffers
  int64_t o_pi1 = 0, o_pi2 = 0, o_pih = 0, o_q1[3] = {0, 0, 0}, o_q2[3] = {0, 0, 0};
  int Smax = 1;
  float *W = nullptr, *Wt = nullptr, *Mo = nullptr, *Vo = nullptr, *Gp = nullptr, *G = nullptr;
  StepState* st = nullptr;
  float* SCAL = nullptr;
  double* partials = nullptr;       // per-CTA loss partial sums of k_qheads_losses
  unsigned int* ticket = nullptr;   // its last-CTA-done counter
  // batch + activations
  float *X = nullptr, *X2 = nullptr, *ACT = nullptr, *R = nullptr, *DN = nullptr, *NOISE = nullptr;
  float *H1[8] = {}, *H2[8] = {}, *HD[3] = {}, *Q[5] = {};  // passes a..h ; heads a..c ; q d..h
  float *A1 = nullptr, *A3 = nullptr, *LOGP1 = nullptr, *LOGP2 = nullptr;
  float *dQ[3] = {}, *dZ2[3] = {}, *dZ1[3] = {}, *dA1 = nullptr, *dHD = nullptr, *dZ2a = nullptr, *dZ1a = nullptr;
  // tensor-core mode: H1 / dZ1 / dZ2 / dZ2a / dZ1a are hi/lo plane pairs [2][maxB][ld], ld rounded up to 4 floats
  // (16-byte TMA strides); lo* = floats from the hi plane to the lo plane (0 in FFMA mode, where ld = width)
  int ld1 = 0, ld2 = 0, ldx = 0, ldh = 0;   // ldh: head block row pitch
  long long lo1 = 0, lo2 = 0, lox = 0;
  float* XA[3] = {};                    // [x|a], [x|a1], [x2|a3] split planes
  bool fuse_fwd = false;                // narrow inputs: first + second layer of a forward pass in one tcgen05 launch
                                        // (fwd_fused_tc); DDRL_FUSE_L1=0 keeps the two-launch form
  int force_bn = 0;                     // DDRL_TC_BN=64|128 overrides the per-stage tile width choice
  unsigned long long* dp_trace = nullptr;   // DDRL_DP_TRACE=1: 8 phase time stamps of k_adam_dp's CTA 0 (ddrl_sac_dp_trace)
  bool dp_v1 = false;                   // DDRL_DP_V1=1: two-kernel form of the fused data-parallel step (reduce kernel, then full peer read + optimiser)
  bool narrow_w1 = false;               // policy W1 gradient (K = D <= 32) by k_wgrad_narrow instead of a tensor-core stage
  cudaGraphExec_t dbg_exec = nullptr;   // ddrl_sac_debug_stage(reps < 0): the last measurement graph
  int dbg_key[3] = {0, 0, 0};           // its (batch, stage, reps)
  float* host_stage = nullptr;          // device copy of a host batch block (ddrl_sac_step_host), maxB * (2D + A + 2) floats
  float* host_scal = nullptr;           // its 4 output scalars before the D2H copy
  float* Gn = nullptr;                  // its per-slice partial blocks [ceil(maxB / 64)][(D + 1) * h1]
  uint32_t* H1bits[8] = {};             // relu'(H1) of each pass as bit masks [maxB][ldbits] (written by the L1 epilogue)
  int ldbits = 0;
  float *Wsp = nullptr, *Wtsp = nullptr;  // split planes of the main / target weight blocks (SplitMap)
  SplitMap smap{};
  std::vector<void*> allocs;
  std::map<int, Plan> plans;
  bool use_graph = true;
  bool use_tc = true;    // tcgen05 3xTF32 GEMMs from pre-split planes (default); DDRL_GEMM=ffma: fp32 FFMA tiles
  PeerComm pc{};                      // fused data-parallel mode (ddrl_sac_comm_attach): peers' communication buffers
  float* comm = nullptr;              // this rank's buffer: [2][Pc] floats, then 8 arrival flags
  bool peer_opened[8] = {};
  int* d_err = nullptr;
  int t_host = 0;                     // number of updates enqueued so far (Adam step count, noise counter)
  StepDyn last_dyn{};
  cudaStream_t side_stream = nullptr; // tensor-core mode: skinny / bias gradients run beside the main chain (forked with events)
  cudaEvent_t ev[4] = {};
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy
                                      // default stream, which cannot be captured); replay on the caller's
};

namespace {

int dalloc(ddrl_sac* h, float** p, size_t nfloats) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, nfloats * sizeof(float) + 16);
  if (e != cudaSuccess) return fail(DDRL_ENOMEM, "cudaMalloc(%zu floats) failed: %s", nfloats, cudaGetErrorString(e));
  cudaMemset(q, 0, nfloats * sizeof(float) + 16);
  h->allocs.push_back(q);
  *p = (float*)q;
  return 0;
}

GemmProb mk(Seg a0, Seg a1, int ones, int a_trans, const float* Bp, int ldb, int b_trans, float* C, int ldc, int M,
            int N, int K, int epi = EPI_NONE, const float* mask = nullptr, int ldmask = 0) {
  GemmProb p{};
  p.a0 = a0; p.a1 = a1; p.a_ones = ones; p.a_trans = a_trans;
  p.B = Bp; p.ldb = ldb; p.b_trans = b_trans;
  p.C = C; p.ldc = ldc; p.c_split_stride = 0;
  p.M = M; p.N = N; p.K = K;
  p.epi = epi; p.mask = mask; p.ldmask = ldmask;
  p.splits = 1; p.k_per_split = K;
  return p;
}
Seg seg(const float* p, int ld, int w) { return Seg{p, ld, w}; }
Seg none() { return Seg{nullptr, 0, 0}; }

int set_out_map(tc::TcProb* p);
int pack_tc(std::vector<tc::TcProb>& v, tc::TcGroup* g, int* tiles, int bn) {
  if ((int)v.size() > tc::MAX_PROBS) return fail(DDRL_EINVAL, "too many tensor-core problems in one stage (%d)", (int)v.size());
  int t = 0;
  g->nprob = (int)v.size();
  for (size_t i = 0; i < v.size(); ++i) {
    tc::TcProb& p = v[i];
    if (int rc = set_out_map(&p)) return rc;
    p.tiles_m = (p.M + tc::BM - 1) / tc::BM;
    p.tiles_n = (p.N + bn - 1) / bn;
    p.tile_begin = t;
    t += p.tiles_m * p.tiles_n * p.splits;
    g->p[i] = p;
  }
  *tiles = t;
  return 0;
}

int pack(std::vector<GemmProb>& v, GemmGroup* g, int* tiles) {
  if ((int)v.size() > GEMM_MAX_PROBS) return fail(DDRL_EINVAL, "too many GEMM problems in one stage (%d)", (int)v.size());
  int t = 0;
  g->nprob = (int)v.size();
  for (size_t i = 0; i < v.size(); ++i) {
    GemmProb& p = v[i];
    const int BM = p.cfg == 0 ? 64 : 128, BN = p.cfg == 0 ? 64 : 16;
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    p.tile_begin = t;
    t += p.tiles_m * p.tiles_n * p.splits;
    g->p[i] = p;
  }
  *tiles = t;
  return 0;
}

int pack_fz(std::vector<tc::FusedProb>& v, tc::FusedGroup* g, int* tiles) {
  if ((int)v.size() > tc::FZ_MAX_PROBS) return fail(DDRL_EINVAL, "too many fused forward problems in one stage (%d)", (int)v.size());
  int t = 0;
  g->nprob = (int)v.size();
  for (size_t i = 0; i < v.size(); ++i) {
    tc::FusedProb& p = v[i];
    p.tiles_m = (p.M + tc::BM - 1) / tc::BM;
    p.tiles_n = (p.h2 + tc::FZ_BN - 1) / tc::FZ_BN;
    p.tile_begin = t;
    t += p.tiles_m * p.tiles_n;
    g->p[i] = p;
  }
  *tiles = t;
  return 0;
}

int finalize_group(Group& g) {
  int rc = pack(g.probs, &g.grp, &g.tiles);
  if (rc) return rc;
  if ((rc = pack_fz(g.probs_fz, &g.grp_fz, &g.tiles_fz))) return rc;
  return pack_tc(g.probs_tc, &g.grp_tc, &g.tiles_tc, g.bn);
}

int launch_tc(const Group& g, cudaStream_t s) {
  if (g.tiles_fz > 0) {
    DDRL_CUDA(launch_pdl(tc::fwd_fused_tc, dim3(g.tiles_fz), dim3(tc::FZ_THREADS), tc::FZ_SMEM_BYTES, s, g.grp_fz));
    DDRL_LAUNCH_CHECK();
  }
  if (g.tiles_tc > 0) {
    if (g.bn == 64)
      DDRL_CUDA(launch_pdl(tc::gemm_grouped_tc<64>, dim3(g.tiles_tc), dim3(256), tc::Cfg<64>::SMEM_BYTES, s, g.grp_tc));
    else
      DDRL_CUDA(launch_pdl(tc::gemm_grouped_tc<128>, dim3(g.tiles_tc), dim3(256), tc::Cfg<128>::SMEM_BYTES, s, g.grp_tc));
    DDRL_LAUNCH_CHECK();
  }
  return 0;
}
int launch_f32(const Group& g, cudaStream_t s) {
  if (g.tiles > 0) {
    DDRL_CUDA(launch_pdl(gemm_grouped_f32, dim3(g.tiles), dim3(256), 0, s, g.grp));
    DDRL_LAUNCH_CHECK();
  }
  return 0;
}
int launch_group(const Group& g, cudaStream_t s) {
  int rc = launch_tc(g, s);
  return rc ? rc : launch_f32(g, s);
}

// GEMM stages; the row-wise kernels (policy heads, Q heads + losses, policy backward) sit between them
enum { ST_L1 = 0, ST_L2 /*-> k_policy_heads_fwd*/, ST_QL1, ST_QL2 /*-> k_qheads_losses*/, ST_BQ
       /*-> k_policy_bwd_rows*/, ST_BP, ST_BP3, ST_COUNT };

void add(std::vector<Group>& stage, GemmProb p) {
  p.cfg = p.N <= 16 ? 1 : 0;
  if (stage.empty()) stage.emplace_back();
  stage[0].probs.push_back(p);
}
void add_tc(std::vector<Group>& stage, const tc::TcProb& p) {
  if (stage.empty()) stage.emplace_back();
  stage[0].probs_tc.push_back(p);
}

// ---- TMA tensor maps over the pre-split operands ([2][rows][pitch] floats, plane 0 = hi, 1 = lo) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
struct View {          // a pre-split 2-D tensor: `rows` x `cols` valid elements, row pitch `ld`, lo plane `lo` floats after hi
  const float* p;
  int ld;
  long long lo;
  int rows, cols;
};
// operand whose contraction index runs along the COLUMNS of the stored tensor (K-major): box = box_rows x 32 cols
//   (box_rows = 128 for the A operand, the tile width BN for the B operand)
// operand whose contraction index runs along the ROWS (MN-major): box = 32 rows (k) x 32 cols (mn)
int make_map(CUtensorMap* m, const View& v, bool mn_major, int box_rows = 128) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(DDRL_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((v.ld & 3) || (v.lo & 3) || (reinterpret_cast<uintptr_t>(v.p) & 15))
    return fail(DDRL_EINVAL, "tensor map operand is not 16-byte aligned (ld=%d)", v.ld);
  cuuint64_t dims[3] = {(cuuint64_t)v.cols, (cuuint64_t)v.rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)v.ld * 4, (cuuint64_t)v.lo * 4};
  cuuint32_t box[3] = {32, mn_major ? 32u : (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(v.p), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DDRL_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, v.rows, v.cols, v.ld);
  return 0;
}
// C[M,N] = epi( opA . opB + bias ):  a / b are the stored tensors, a_mn / b_mn say which index is contracted
int mk_tc(tc::TcProb* out, int bn, const View& a, bool a_mn, const View& b, bool b_mn, float* C, float* C_lo, int ldc, int M, int N,
          int K, int epi = EPI_NONE, const float* bias = nullptr, const float* mask = nullptr, const float* mask_lo = nullptr,
          int ldmask = 0) {
  tc::TcProb p{};
  int rc = make_map(&p.ta, a, a_mn);
  if (rc) return rc;
  if ((rc = make_map(&p.tb, b, b_mn, bn))) return rc;
  p.C = C; p.C_lo = C_lo; p.ldc = ldc; p.c_split_stride = 0;
  p.mask = mask; p.mask_lo = mask_lo; p.ldmask = ldmask; p.bias = bias;
  p.M = M; p.N = N; p.K = K; p.epi = epi; p.a_mn = a_mn; p.b_mn = b_mn;
  p.splits = 1; p.k_per_split = (K + tc::BK - 1) / tc::BK * tc::BK;
  *out = p;
  return 0;
}
// output tensor map (N, M, plane), box 32 x 32, SWIZZLE_128B (the epilogues stage 32 x 32 blocks in shared memory)
int make_out_map(CUtensorMap* m, float* C, int N, int M, int ldc, int planes, long long plane) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(DDRL_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)(planes > 1 ? plane : (long long)ldc * M) * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, C, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DDRL_ECUDA, "cuTensorMapEncodeTiled(output) failed (%d) M=%d N=%d ldc=%d", (int)r, M, N, ldc);
  return 0;
}
// plane = hi/lo (C_lo set) or the split-K partial index.  Falls back to direct stores (c_tma = 0) when the output is
// not 16-byte aligned / pitched (e.g. N = 33 weight-gradient blocks).
int set_out_map(tc::TcProb* p) {
  const long long plane = p->C_lo ? (long long)(p->C_lo - p->C) : p->c_split_stride;
  const int planes = p->C_lo ? 2 : p->splits;
  p->c_tma = 0;
  if ((p->ldc & 3) || (reinterpret_cast<uintptr_t>(p->C) & 15) || (planes > 1 && (plane <= 0 || (plane & 3)))) return 0;
  if (int rc = make_out_map(&p->tc, p->C, p->N, p->M, p->ldc, planes, plane)) return rc;
  p->c_tma = 1;
  return 0;
}

int build_plan_tc(ddrl_sac* h, int B, Plan& pl);

int build_plan(ddrl_sac* h, int B, Plan& pl) {
  const int D = h->D, A = h->A, h1 = h->h1, h2 = h->h2;
  pl.B = B;
  const int kps = 256;
  pl.S = (B + kps - 1) / kps;
  if (pl.S > h->Smax) return fail(DDRL_EINVAL, "batch %d exceeds max_batch %d", B, h->maxB);
  pl.stages.assign(ST_COUNT, {});
  if (h->use_tc) return build_plan_tc(h, B, pl);
  float *W = h->W, *Wt = h->Wt, *Gp = h->Gp;
  auto wg = [&](GemmProb p) { p.splits = pl.S; p.k_per_split = kps; p.c_split_stride = h->P; return p; };
  enum { a = 0, b, c, d, e, f, g, hh };
  // ---- forward.  First stage: policies a (main@x), c (target@x2) and the data-action Q passes d, e; second stage:
  // policy b (main@x2, only its log-likelihood is used, by the losses) and the policy-action Q passes f = Q1(x,a1),
  // g = Q1_targ(x2,a3), hh = Q2_targ(x2,a3) — four passes per stage
  const float* xin[8] = {h->X, h->X2, h->X2, h->X, h->X, h->X, h->X2, h->X2};
  const float* ain[8] = {nullptr, nullptr, nullptr, h->ACT, h->ACT, h->A1, h->A3, h->A3};
  const float* wsrc[8] = {W + h->o_pi1, W + h->o_pi1, Wt + h->o_pi1, W + h->o_q1[0], W + h->o_q2[0],
                          W + h->o_q1[0], Wt + h->o_q1[0], Wt + h->o_q2[0]};
  const float* wsrc2[8] = {W + h->o_pi2, W + h->o_pi2, Wt + h->o_pi2, W + h->o_q1[1], W + h->o_q2[1],
                           W + h->o_q1[1], Wt + h->o_q1[1], Wt + h->o_q2[1]};
  for (int p = 0; p < 8; ++p) {
    const bool first = p == a || p == c || p == d || p == e;
    const bool isq = ain[p] != nullptr;
    add(pl.stages[first ? ST_L1 : ST_QL1], mk(seg(xin[p], D, D), isq ? seg(ain[p], A, A) : none(), 1, 0, wsrc[p], h1, 0, h->H1[p], h1,
                                              B, h1, D + (isq ? A : 0) + 1, EPI_RELU));
    add(pl.stages[first ? ST_L2 : ST_QL2], mk(seg(h->H1[p], h1, h1), none(), 1, 0, wsrc2[p], h2, 0, h->H2[p], h2, B, h2, h1 + 1, EPI_RELU));
  }
  // ---- backward of the three differentiated Q passes: 0 = d (Q1 data), 1 = e (Q2 data), 2 = f (Q1 pi-path)
  const int pass[3] = {d, e, f};
  const int64_t* oq[3] = {h->o_q1, h->o_q2, h->o_q1};
  for (int i = 0; i < 3; ++i) {
    // dZ1 = dZ2 . W2^T  masked by relu'(H1)        (dZ2 comes from k_qheads_losses)
    add(pl.stages[ST_BQ], mk(seg(h->dZ2[i], h2, h2), none(), 0, 0, W + oq[i][1], h2, 1, h->dZ1[i], h1, B, h1, h2, EPI_MASK,
                             h->H1[pass[i]], h1));
    if (i < 2) {
      // d[W3;b3] = [H2|1]^T dq ; d[W2;b2] = [H1|1]^T dZ2 ; d[W1;b1] = [x|a|1]^T dZ1
      add(pl.stages[ST_BQ], wg(mk(seg(h->H2[pass[i]], h2, h2), none(), 1, 1, h->dQ[i], 1, 0, Gp + oq[i][2], 1, h2 + 1, 1, B)));
      add(pl.stages[ST_BQ], wg(mk(seg(h->H1[pass[i]], h1, h1), none(), 1, 1, h->dZ2[i], h2, 0, Gp + oq[i][1], h2, h1 + 1, h2, B)));
      add(pl.stages[ST_BP], wg(mk(seg(h->X, D, D), seg(h->ACT, A, A), 1, 1, h->dZ1[i], h1, 0, Gp + oq[i][0], h1, D + A + 1, h1, B)));
    }
  }
  // ---- policy backward (pass a); dHD and dZ2a come from k_policy_bwd_rows
  add(pl.stages[ST_BP], wg(mk(seg(h->H2[a], h2, h2), none(), 1, 1, h->dHD, 2 * A, 0, Gp + h->o_pih, h->ldh, h2 + 1, 2 * A, B)));
  add(pl.stages[ST_BP], mk(seg(h->dZ2a, h2, h2), none(), 0, 0, W + h->o_pi2, h2, 1, h->dZ1a, h1, B, h1, h2, EPI_MASK, h->H1[a], h1));
  add(pl.stages[ST_BP], wg(mk(seg(h->H1[a], h1, h1), none(), 1, 1, h->dZ2a, h2, 0, Gp + h->o_pi2, h2, h1 + 1, h2, B)));
  add(pl.stages[ST_BP3], wg(mk(seg(h->X, D, D), none(), 1, 1, h->dZ1a, h1, 0, Gp + h->o_pi1, h1, D + 1, h1, B)));
  for (auto& st : pl.stages)
    for (auto& g2 : st) {
      int rc = finalize_group(g2);
      if (rc) return rc;
    }
  return 0;
}

// Tensor-core plan: every GEMM with N > 16 runs on tcgen05 from the pre-split planes; the skinny weight
// gradients (Q heads, policy heads) and the bias gradients (column sums of dZ) stay on FFMA tiles.
// Stage layout (B = 1024, 256 x 256: every stage is <= 128 tiles of 128 x 64, one per SM, in ONE wave):
//   ST_L1 / ST_L2     passes a (pi main@x), c (pi target@x2), d, e (Q1, Q2 @ (x, a))      -> heads of a, c
//   ST_QL1 / ST_QL2   passes b (pi main@x2), f (Q1 @ (x, a1)), g, h (Q targets @ (x2, a3)) -> Q heads + losses (+ head of b)
//   ST_BQ             dgrad d, e, f; wgrad W2(q1)                                          -> policy backward rows
//   ST_BP             dgrad a; wgrad W2(pi), W2(q2), W1(q1), W1(q2)
//   ST_BP3            wgrad W1(pi)
// With narrow inputs (D + A <= 32, h1 <= 256) the first and second layer of a forward stage are ONE launch of
// fwd_fused_tc (ST_L1 / ST_QL1 stay empty).
int build_plan_tc(ddrl_sac* h, int B, Plan& pl) {
  const int D = h->D, A = h->A, h1 = h->h1, h2 = h->h2;
  float *W = h->W, *Gp = h->Gp;
  const SplitMap& sm = h->smap;
  int rc = 0;
  const int kps = 256;
  auto wg_tc = [&](tc::TcProb p) { p.splits = pl.S; p.k_per_split = kps; p.c_split_stride = h->P; return p; };
  enum { a = 0, b, c, d, e, f, g, hh };
  enum { PI1 = 0, PI2, Q1_0, Q1_1, Q2_0, Q2_1 };   // SplitMap block indices
  const int64_t ioff[6] = {h->o_pi1, h->o_pi2, h->o_q1[0], h->o_q1[1], h->o_q2[0], h->o_q2[1]};
  const int Kb[6] = {D, h1, D + A, h1, D + A, h1};
  const int Nb[6] = {h1, h2, h1, h2, h1, h2};
  auto wview = [&](int blk, bool target) {
    return View{(target ? h->Wtsp : h->Wsp) + sm.sp_off[blk], sm.pitch[blk], sm.plane, Kb[blk], Nb[blk]};
  };
  auto bias_of = [&](int blk, bool target) { return (target ? h->Wt : W) + ioff[blk] + (int64_t)Kb[blk] * Nb[blk]; };
  auto xa = [&](int which, int cols) { return View{h->XA[which], h->ldx, h->lox, B, cols}; };
  auto h1v = [&](int p) { return View{h->H1[p], h->ld1, h->lo1, B, h1}; };
  auto dz2v = [&](const float* p) { return View{p, h->ld2, h->lo2, B, h2}; };
  auto dz1v = [&](const float* p) { return View{p, h->ld1, h->lo1, B, h1}; };
  // The stage definitions below run twice: a dry pass that only counts 128 x 128 tiles per stage (to pick the tile
  // width: 64 when 128-wide tiles would leave SMs idle), then the pass that builds the tensor maps.
  bool dry = true;
  int tiles128[ST_COUNT] = {}, bn_of[ST_COUNT] = {};
  auto count = [&](int stage, int M, int N, int splits) { tiles128[stage] += ((M + 127) / 128) * ((N + 127) / 128) * splits; };
  auto push = [&](int stage, const tc::TcProb& p) {
    if (pl.stages[stage].empty()) pl.stages[stage].emplace_back();
    pl.stages[stage][0].bn = bn_of[stage];
    pl.stages[stage][0].probs_tc.push_back(p);
  };
  auto fwd = [&](int stage, const View& act, int blk, bool target, float* C, float* C_lo, int ldc, uint32_t* bits = nullptr) {
    if (dry) return count(stage, B, Nb[blk], 1);
    tc::TcProb p;
    if (!rc && !(rc = mk_tc(&p, bn_of[stage], act, false, wview(blk, target), true, C, C_lo, ldc, B, Nb[blk], Kb[blk], EPI_RELU,
                            bias_of(blk, target)))) {
      p.relu_bits = bits; p.ldbits = h->ldbits;
      push(stage, p);
    }
  };
  auto fused = [&](int stage, const View& x, int blk1, int blk2, bool target, int pass, bool keep_h1, bool keep_bits) {
    if (dry || rc) return;
    tc::FusedProb p{};
    if ((rc = make_map(&p.tx, x, false))) return;
    if ((rc = make_map(&p.tw1, wview(blk1, target), true))) return;
    if ((rc = make_map(&p.tw2, wview(blk2, target), true))) return;
    if (keep_h1 && (rc = make_out_map(&p.th1, h->H1[pass], h1, B, h->ld1, 2, h->lo1))) return;
    if ((rc = make_out_map(&p.tc, h->H2[pass], h2, B, h2, 1, 0))) return;
    p.bias1 = bias_of(blk1, target); p.bias2 = bias_of(blk2, target);
    p.bits = keep_bits ? h->H1bits[pass] : nullptr; p.ldbits = h->ldbits; p.store_h1 = keep_h1 ? 1 : 0;
    p.M = B; p.K1 = Kb[blk1]; p.h1 = h1; p.h2 = h2;
    if (pl.stages[stage].empty()) pl.stages[stage].emplace_back();
    pl.stages[stage][0].probs_fz.push_back(p);
  };
  auto dgrad = [&](int stage, const float* dz2, int blk, float* dz1, int mask_pass) {
    if (dry) return count(stage, B, Kb[blk], 1);
    tc::TcProb p;   // dZ1 = dZ2 . W2^T masked by relu'(H1)
    if (!rc && !(rc = mk_tc(&p, bn_of[stage], dz2v(dz2), false, wview(blk, false), false, dz1, dz1 + h->lo1, h->ld1, B, Kb[blk],
                            Nb[blk], EPI_MASK, nullptr, h->H1[mask_pass], h->H1[mask_pass] + h->lo1, h->ld1))) {
      p.mask_bits = h->H1bits[mask_pass]; p.ldbits = h->ldbits;
      push(stage, p);
    }
  };
  auto wgrad = [&](int stage, const View& act, const View& dz, int blk) {
    if (dry) return count(stage, Kb[blk], Nb[blk], pl.S);
    tc::TcProb p;   // d[W] = act^T . dZ (kernel rows); the bias row is the column sum of dZ, on FFMA tiles
    if (!rc && !(rc = mk_tc(&p, bn_of[stage], act, true, dz, true, Gp + ioff[blk], nullptr, Nb[blk], Kb[blk], Nb[blk], B)))
      push(stage, wg_tc(p));
    ColsumGroup& cg = pl.colsum[stage];
    if (!rc && cg.nprob >= COLSUM_MAX) rc = fail(DDRL_EINVAL, "too many bias-gradient problems in one stage");
    if (!rc) {
      auto& q = cg.p[cg.nprob++];
      q.z = dz.p; q.lo = dz.lo; q.ld = dz.ld; q.N = Nb[blk];
      q.out = Gp + ioff[blk] + (int64_t)Kb[blk] * Nb[blk];
      q.chunk_begin = pl.colsum_chunks[stage];
      pl.colsum_chunks[stage] += (Nb[blk] + 31) / 32;
      cg.B = B; cg.kps = kps; cg.split_stride = h->P;
    }
  };
  auto skinny = [&](int stage, const float* act, int ld_act, int K, const float* z, int ldz, int n, float* out, int ldc) {
    if (dry) return;
    SkinnyGroup& sg = pl.skinny[stage];
    if (!rc && sg.nprob >= SKINNY_MAX) rc = fail(DDRL_EINVAL, "too many skinny weight-gradient problems in one stage");
    if (rc) return;
    auto& q = sg.p[sg.nprob++];
    q.act = act; q.ld_act = ld_act; q.K = K; q.z = z; q.ldz = ldz; q.n = n; q.out = out; q.ldc = ldc;
    q.chunk_begin = pl.skinny_chunks[stage];
    pl.skinny_chunks[stage] += (K + 1 + 31) / 32;
    sg.B = B; sg.kps = kps; sg.split_stride = h->P;
  };
  pl.fused = h->fuse_fwd;
  // per forward pass: input planes, first / second layer weight block, target net?, is the H1 plane pair / relu mask
  // needed by the backward (wgrad operand: a, d, e; dgrad mask: a, d, e, f)
  struct FwdPass { int xa_idx, cols, blk1, blk2; bool target, keep_h1, keep_bits, first; };
  const FwdPass fp[8] = {
      {0, D, PI1, PI2, false, true, true, true},            // a  pi main   @ x
      {2, D, PI1, PI2, false, false, false, false},         // b  pi main   @ x2
      {2, D, PI1, PI2, true, false, false, true},           // c  pi target @ x2
      {0, D + A, Q1_0, Q1_1, false, true, true, true},      // d  Q1 main   @ (x, a)
      {0, D + A, Q2_0, Q2_1, false, true, true, true},      // e  Q2 main   @ (x, a)
      {1, D + A, Q1_0, Q1_1, false, false, true, false},    // f  Q1 main   @ (x, a1)
      {2, D + A, Q1_0, Q1_1, true, false, false, false},    // g  Q1 target @ (x2, a3)
      {2, D + A, Q2_0, Q2_1, true, false, false, false},    // h  Q2 target @ (x2, a3)
  };
  auto define_stages = [&]() {
    for (int p = 0; p < 8; ++p) {
      const FwdPass& q = fp[p];
      if (pl.fused) {
        fused(q.first ? ST_L2 : ST_QL2, xa(q.xa_idx, q.cols), q.blk1, q.blk2, q.target, p, q.keep_h1, q.keep_bits);
      } else {
        fwd(q.first ? ST_L1 : ST_QL1, xa(q.xa_idx, q.cols), q.blk1, q.target, h->H1[p], h->H1[p] + h->lo1, h->ld1,
            q.keep_bits ? h->H1bits[p] : nullptr);
        fwd(q.first ? ST_L2 : ST_QL2, h1v(p), q.blk2, q.target, h->H2[p], nullptr, h2);
      }
    }
    // ---- backward of the three differentiated Q passes: 0 = d (Q1 data), 1 = e (Q2 data), 2 = f (Q1 pi-path)
    const int pass[3] = {d, e, f};
    const int w2blk[3] = {Q1_1, Q2_1, Q1_1}, w1blk[3] = {Q1_0, Q2_0, Q1_0};
    const int64_t* oq[3] = {h->o_q1, h->o_q2, h->o_q1};
    for (int i = 0; i < 3; ++i) {
      dgrad(ST_BQ, h->dZ2[i], w2blk[i], h->dZ1[i], pass[i]);
      if (i < 2) {
        skinny(ST_BQ, h->H2[pass[i]], h2, h2, h->dQ[i], 1, 1, Gp + oq[i][2], 1);      // d[W3;b3] = [H2|1]^T dq
        wgrad(i == 0 ? ST_BQ : ST_BP, h1v(pass[i]), dz2v(h->dZ2[i]), w2blk[i]);       // W2(q2) rides with the policy stage
        wgrad(ST_BP, xa(0, D + A), dz1v(h->dZ1[i]), w1blk[i]);
      }
    }
    // ---- policy backward (pass a); dHD and dZ2a come from k_policy_bwd_rows
    skinny(ST_BP, h->H2[a], h2, h2, h->dHD, 2 * A, 2 * A, Gp + h->o_pih, h->ldh);      // d[Whead;bhead] = [H2a|1]^T dHD
    dgrad(ST_BP, h->dZ2a, PI2, h->dZ1a, a);
    wgrad(ST_BP, h1v(a), dz2v(h->dZ2a), PI2);
    if (!h->narrow_w1) wgrad(ST_BP3, xa(0, D), dz1v(h->dZ1a), PI1);      // else k_wgrad_narrow (enqueue_grads)
  };
  define_stages();
  for (int st = 0; st < ST_COUNT; ++st) {
    bn_of[st] = tiles128[st] < h->sms ? 64 : 128;
    if (h->force_bn) bn_of[st] = h->force_bn;
  }
  dry = false;
  define_stages();
  if (rc) return rc;
  for (auto& st : pl.stages)
    for (auto& g2 : st)
      if ((rc = finalize_group(g2))) return rc;
  return 0;
}

int run_stage(const Plan& pl, int s, cudaStream_t st, bool tc_only = false) {
  for (const auto& g : pl.stages[s]) {
    int rc = tc_only ? launch_tc(g, st) : launch_group(g, st);
    if (rc) return rc;
  }
  return 0;
}
// tensor-core mode: the FFMA tiles (skinny weight gradients) and bias column sums of stage `st`, on the side stream
int run_side(const Plan& pl, int st, int S, cudaStream_t side) {
  for (const auto& g : pl.stages[st]) {
    int rc = launch_f32(g, side);
    if (rc) return rc;
  }
  if (pl.skinny[st].nprob > 0) {
    k_skinny_wgrad<<<dim3(pl.skinny_chunks[st], S), 256, 0, side>>>(pl.skinny[st]);
    DDRL_LAUNCH_CHECK();
  }
  if (pl.colsum[st].nprob > 0) {
    k_colsum<<<dim3(pl.colsum_chunks[st], S), 256, 0, side>>>(pl.colsum[st]);
    DDRL_LAUNCH_CHECK();
  }
  return 0;
}

XaOut xa_out(const ddrl_sac* h) {
  return h->use_tc ? XaOut{h->XA[0], h->XA[1], h->XA[2], h->ldx, h->lox} : XaOut{nullptr, nullptr, nullptr, 0, 0};
}

int prologue_blocks(const ddrl_sac* h, int B) {
  const int64_t work = std::max<int64_t>((int64_t)B * h->D, 3LL * B * h->A);
  return (int)std::min<int64_t>((work + 255) / 256, h->sms * 4);
}
int launch_prologue(ddrl_sac* h, const Plan& pl, const StepDyn& dyn, cudaStream_t s) {
  const int B = pl.B, D = h->D, A = h->A;
  DDRL_CUDA(launch_pdl(k_prologue, dim3(prologue_blocks(h, B)), dim3(256), 0, s, h->st, dyn, B, D, A, h->X, h->X2, h->ACT, h->R,
                       h->DN, h->NOISE, xa_out(h)));
  DDRL_LAUNCH_CHECK();
  return 0;
}
// this step's values into the prologue node of an instantiated graph
int patch_prologue(ddrl_sac* h, const Plan& pl, cudaGraphExec_t exec, cudaGraphNode_t node, const StepDyn& dyn) {
  int B = pl.B, D = h->D, A = h->A;
  StepDyn d = dyn;
  XaOut xa = xa_out(h);
  void* args[] = {&h->st, &d, &B, &D, &A, &h->X, &h->X2, &h->ACT, &h->R, &h->DN, &h->NOISE, &xa};
  cudaKernelNodeParams np{};
  np.func = (void*)k_prologue;
  np.gridDim = dim3(prologue_blocks(h, B)); np.blockDim = dim3(256);
  np.sharedMemBytes = 0; np.kernelParams = args; np.extra = nullptr;
  DDRL_CUDA(cudaGraphExecKernelNodeSetParams(exec, node, &np));
  return 0;
}
bool narrow_heads(const ddrl_sac* h) { return 2 * h->A <= 16 && (h->h2 & 3) == 0; }
// policy heads of the passes in `pm` (0: main pi(x), 1: main pi(x2), 2: target pi(x2))
int launch_heads(ddrl_sac* h, const Plan& pl, cudaStream_t s, PassMap pm) {
  const int B = pl.B, D = h->D, A = h->A, h2 = h->h2;
  const int rows = pm.n * B;
  if (narrow_heads(h)) {     // warp per row
    DDRL_CUDA(launch_pdl(k_policy_heads_rows, dim3((rows + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, s,
        B, A, h2, h->ldh, h->act_scale, h->H2[0], h->H2[1], h->H2[2], h->W + h->o_pih, h->Wt + h->o_pih, h->NOISE, h->HD[0],
        h->A1, h->A3, h->LOGP1, h->LOGP2, xa_out(h), D, pm));
    DDRL_LAUNCH_CHECK();
    return 0;
  }
  const int groups = std::min(32, HEADS_THREADS / (2 * A));    // row groups per CTA (2A <= 64 -> >= 4)
  auto smem_of = [&](int r) { return ((size_t)r * ((h2 + 3) / 4 * 4 + 4) + (size_t)r * 4 * A) * sizeof(float); };
  // 4 rows per thread when there is enough work to still fill the GPU and a 4-row group never straddles a pass
  // boundary (B % 4 == 0); otherwise one row per thread
  const bool rb4 = (B % 4 == 0) && ((int64_t)rows / (4 * groups) >= 2LL * h->sms) && smem_of(4 * groups) <= HEADS_SMEM_MAX;
  int R = rb4 ? 4 * groups : groups;
  if (!rb4) while (R > 1 && smem_of(R) > HEADS_SMEM_MAX) R /= 2;
  const size_t smem = smem_of(R);
  if (smem > HEADS_SMEM_MAX) return fail(DDRL_EINVAL, "hidden size %d is too wide for the policy-head kernel", h2);
  DDRL_CUDA(launch_pdl(rb4 ? k_policy_heads_fwd<4> : k_policy_heads_fwd<1>, dim3((rows + R - 1) / R), dim3(HEADS_THREADS), smem, s,
      B, A, h2, h->ldh, R, h->act_scale, h->H2[0], h->H2[1], h->H2[2], h->W + h->o_pih, h->Wt + h->o_pih, h->NOISE, h->HD[0], h->A1,
      h->A3, h->LOGP1, h->LOGP2, xa_out(h), D, pm));
  DDRL_LAUNCH_CHECK();
  return 0;
}
int launch_qheads(ddrl_sac* h, const Plan& pl, cudaStream_t s) {
  const int B = pl.B, h2 = h->h2;
  DDRL_CUDA(launch_pdl(k_qheads_losses, dim3((B + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, s,
      h->st, B, h2, h->gamma, h->H2[3], h->H2[4], h->H2[5], h->H2[6], h->H2[7], h->W + h->o_q1[2], h->W + h->o_q2[2],
      h->Wt + h->o_q1[2], h->Wt + h->o_q2[2], h->R, h->DN, h->LOGP1, h->LOGP2, h->dQ[0], h->dQ[1], h->dZ2[0], h->dZ2[1],
      h->dZ2[2], h->partials, h->ticket, h->SCAL, h->ld2, h->lo2,
      narrow_heads(h) ? 1 : 0, h->H2[1], h->W + h->o_pih, h->NOISE, h->A, h->ldh, h->act_scale));
  DDRL_LAUNCH_CHECK();
  return 0;
}
int launch_pbwd(ddrl_sac* h, const Plan& pl, cudaStream_t s) {
  const int B = pl.B, D = h->D, A = h->A, h1 = h->h1, h2 = h->h2;
  DDRL_CUDA(launch_pdl(k_policy_bwd_rows, dim3((B + ROW_WARPS - 1) / ROW_WARPS), dim3(ROW_WARPS * 32), 0, s,
      h->st, B, A, h1, h2, h->ldh, h->act_scale, h->HD[0], h->NOISE, h->dZ1[2], h->W + h->o_q1[0] + (int64_t)D * h1,
      h->W + h->o_pih, h->H2[0], h->dHD, h->dZ2a, h->ld1, h->lo1, h->ld2, h->lo2));
  DDRL_LAUNCH_CHECK();
  return 0;
}

NarrowGrad narrow_grad(const ddrl_sac* h, int B) {
  NarrowGrad ng{};
  if (h->narrow_w1) {
    ng.Gn = h->Gn; ng.off = h->o_pi1; ng.size = (long long)(h->D + 1) * h->h1; ng.SN = (B + NW_ROWS - 1) / NW_ROWS;
  }
  return ng;
}

int enqueue_grads(ddrl_sac* h, const Plan& pl, cudaStream_t s) {
  int rc;
  if ((rc = run_stage(pl, ST_L1, s))) return rc;
  if ((rc = run_stage(pl, ST_L2, s))) return rc;
  if ((rc = launch_heads(h, pl, s, PassMap{2, {0, 2, 0}}))) return rc;          // a -> a1, logp1; c -> a3
  if ((rc = run_stage(pl, ST_QL1, s))) return rc;
  if ((rc = run_stage(pl, ST_QL2, s))) return rc;
  if (!narrow_heads(h) && (rc = launch_heads(h, pl, s, PassMap{1, {1, 0, 0}}))) return rc;   // b -> logp2 (else inside the next kernel)
  if ((rc = launch_qheads(h, pl, s))) return rc;
  const bool tcm = h->use_tc;
  cudaStream_t side = h->side_stream;
  if (tcm) {
    DDRL_CUDA(cudaEventRecord(h->ev[0], s));
    DDRL_CUDA(cudaStreamWaitEvent(side, h->ev[0], 0));
    if ((rc = run_side(pl, ST_BQ, pl.S, side))) return rc;
  }
  if ((rc = run_stage(pl, ST_BQ, s, tcm))) return rc;
  if ((rc = launch_pbwd(h, pl, s))) return rc;
  if (tcm) {
    DDRL_CUDA(cudaEventRecord(h->ev[1], s));
    DDRL_CUDA(cudaStreamWaitEvent(side, h->ev[1], 0));
    if ((rc = run_side(pl, ST_BP, pl.S, side))) return rc;
  }
  if ((rc = run_stage(pl, ST_BP, s, tcm))) return rc;
  if (tcm && h->narrow_w1) {
    // d[W1;b1](pi) = [x|1]^T dZ1a on FFMA, per 64-row slice; the side stream only has to join
    DDRL_CUDA(cudaEventRecord(h->ev[3], side));
    DDRL_CUDA(launch_pdl(k_wgrad_narrow, dim3((h->h1 + 31) / 32, (pl.B + NW_ROWS - 1) / NW_ROWS), dim3(256), 0, s, pl.B, h->D, h->h1,
                         (const float*)h->XA[0], h->ldx, h->lox, (const float*)h->dZ1a, h->ld1, h->lo1, h->Gn));
    DDRL_LAUNCH_CHECK();
    DDRL_CUDA(cudaStreamWaitEvent(s, h->ev[3], 0));
    return 0;
  }
  if (tcm) {
    DDRL_CUDA(cudaEventRecord(h->ev[2], s));
    DDRL_CUDA(cudaStreamWaitEvent(side, h->ev[2], 0));
    if ((rc = run_side(pl, ST_BP3, pl.S, side))) return rc;
    DDRL_CUDA(cudaEventRecord(h->ev[3], side));
  }
  if ((rc = run_stage(pl, ST_BP3, s, tcm))) return rc;
  if (tcm) DDRL_CUDA(cudaStreamWaitEvent(s, h->ev[3], 0));
  return 0;
}

int enqueue_reduce(ddrl_sac* h, const Plan& pl, cudaStream_t s) {
  const int blocks = (int)std::min<int64_t>((h->P / 4 + 255) / 256, h->sms * 8);
  if (h->pc.world > 1)
    DDRL_CUDA(launch_pdl(k_grad_reduce_comm, dim3(blocks), dim3(256), 0, s, (const StepState*)h->st, h->P, pl.S,
                         (const float*)h->Gp, (const float*)h->SCAL, h->pc, h->ticket + 1, narrow_grad(h, pl.B)));
  else
    DDRL_CUDA(launch_pdl(k_grad_reduce, dim3(blocks), dim3(256), 0, s, h->P, pl.S, (const float*)h->Gp, h->G, narrow_grad(h, pl.B)));
  DDRL_LAUNCH_CHECK();
  return 0;
}

int enqueue_apply(ddrl_sac* h, int S, const float* grads, cudaStream_t s, int B = 0) {
  // B > 0: gradients straight from the backward (split-K partials + the narrow-W1 slices); 0: an already reduced buffer
  const NarrowGrad ng = B > 0 ? narrow_grad(h, B) : NarrowGrad{};
  const int blocks = (int)std::min<int64_t>((h->P / 4 + 255) / 256, h->sms * 8);      // four parameters per thread
  if (h->pc.world > 1 && grads == h->G) {     // data-parallel apply: all-reduce fused into the optimiser over peer memory
    DDRL_CUDA(launch_pdl(k_adam_polyak_peer, dim3(blocks), dim3(256), 0, s, h->st, h->P, h->P_pi, h->lr, h->polyak,
                         -(float)h->A, h->W, h->Wt, h->Mo, h->Vo, h->smap, h->use_tc ? h->Wsp : nullptr,
                         h->use_tc ? h->Wtsp : nullptr, h->pc, h->d_err));
    DDRL_LAUNCH_CHECK();
    return 0;
  }
  if (h->use_tc)
    DDRL_CUDA(launch_pdl(k_adam_polyak_split, dim3(blocks), dim3(256), 0, s, h->st, h->P, h->P_pi, S, grads, h->lr, h->polyak,
                         -(float)h->A, h->SCAL, h->W, h->Wt, h->Mo, h->Vo, h->smap, h->Wsp, h->Wtsp, ng));
  else
    DDRL_CUDA(launch_pdl(k_adam_polyak, dim3(blocks), dim3(256), 0, s, h->st, h->P, h->P_pi, S, grads, h->lr, h->polyak,
                         -(float)h->A, h->SCAL, h->W, h->Wt, h->Mo, h->Vo, ng));
  DDRL_LAUNCH_CHECK();
  return 0;
}

enum { MODE_FULL = 0, MODE_GRADS = 1, MODE_APPLY = 2, MODE_DP = 3 };

int enqueue_mode(ddrl_sac* h, const Plan& pl, int mode, cudaStream_t s) {
  int rc = 0;
  if (mode == MODE_FULL) {
    if ((rc = enqueue_grads(h, pl, s))) return rc;
    return enqueue_apply(h, pl.S, h->Gp, s, pl.B);
  }
  if (mode == MODE_GRADS) {
    if ((rc = enqueue_grads(h, pl, s))) return rc;
    return enqueue_reduce(h, pl, s);
  }
  if (mode == MODE_DP && h->dp_v1) {   // first form: gradients -> exchange buffer -> full peer read + optimiser
    if ((rc = enqueue_grads(h, pl, s))) return rc;
    if ((rc = enqueue_reduce(h, pl, s))) return rc;
    return enqueue_apply(h, 1, h->G, s);
  }
  if (mode == MODE_DP) {   // one kernel: split-K reduce + publish + all-peer read over NVLink + optimiser
    if ((rc = enqueue_grads(h, pl, s))) return rc;
    const int blocks = (int)std::min<int64_t>((h->P / 4 + 255) / 256, h->sms * 2);      // fully resident: CTAs wait for flags
    DDRL_CUDA(launch_pdl(k_adam_dp, dim3(blocks), dim3(256), 0, s, h->st, h->P, h->P_pi, pl.S, (const float*)h->Gp,
                         narrow_grad(h, pl.B), (const float*)h->SCAL, h->lr, h->polyak, -(float)h->A, h->W, h->Wt, h->Mo, h->Vo,
                         h->smap, h->use_tc ? h->Wsp : nullptr, h->use_tc ? h->Wtsp : nullptr, h->pc, h->ticket + 2, h->d_err,
                         h->dp_trace));
    DDRL_LAUNCH_CHECK();
    return 0;
  }
  return enqueue_apply(h, 1, h->G, s);
}

int run_mode(ddrl_sac* h, Plan& pl, int mode, cudaStream_t s, const StepDyn* dyn) {
  cudaGraphExec_t* slot = mode == MODE_FULL ? &pl.exec_full : mode == MODE_GRADS ? &pl.exec_grads
                          : mode == MODE_APPLY ? &pl.exec_apply : &pl.exec_dp;
  const int gi = mode == MODE_FULL ? 0 : mode == MODE_GRADS ? 1 : mode == MODE_APPLY ? 2 : 3;
  if (!h->use_graph) {
    if (dyn) { int rc = launch_prologue(h, pl, *dyn, s); if (rc) return rc; }
    return enqueue_mode(h, pl, mode, s);
  }
  if (!*slot) {
    cudaGraph_t graph = nullptr;
    if (!h->cap_stream) DDRL_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const int64_t before = g_launches.load();
    DDRL_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = dyn ? launch_prologue(h, pl, *dyn, h->cap_stream) : 0;
    if (!rc) rc = enqueue_mode(h, pl, mode, h->cap_stream);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    pl.kernels[mode] = g_launches.load() - before;
    g_launches.store(before);  // captured, not executed
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(DDRL_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    if (dyn) {      // the prologue's node: its parameters are replaced before every launch
      size_t n = 0;
      DDRL_CUDA(cudaGraphGetNodes(graph, nullptr, &n));
      std::vector<cudaGraphNode_t> nodes(n);
      DDRL_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n));
      for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp{};
        if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && kp.func == (void*)k_prologue) { pl.prologue_node[gi] = nd; break; }
      }
      if (!pl.prologue_node[gi]) { cudaGraphDestroy(graph); return fail(DDRL_ECUDA, "captured step has no prologue node"); }
    }
    e = cudaGraphInstantiate(slot, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return fail(DDRL_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    pl.graph[gi] = graph;
  } else if (dyn) {
    int rc = patch_prologue(h, pl, *slot, pl.prologue_node[gi], *dyn);
    if (rc) return rc;
  }
  DDRL_CUDA(cudaGraphLaunch(*slot, s));
  g_launches.fetch_add(pl.kernels[mode], std::memory_order_relaxed);
  return 0;
}

int get_plan(ddrl_sac* h, int B, Plan** out) {
  auto it = h->plans.find(B);
  if (it == h->plans.end()) {
    Plan pl;
    int rc = build_plan(h, B, pl);
    if (rc) return rc;
    it = h->plans.emplace(B, std::move(pl)).first;
  }
  *out = &it->second;
  return 0;
}

}  // namespace

extern "C" {

int ddrl_sac_create(int device, int obs_dim, int act_dim, int h1, int h2, int max_batch, float gamma, float polyak,
                    float lr, float alpha, float act_scale, int gemm_mode, ddrl_sac_t* out) {
  if (!out) return fail(DDRL_EINVAL, "ddrl_sac_create: out is NULL");
  *out = nullptr;
  if (obs_dim < 1 || act_dim < 1 || h1 < 1 || h2 < 1 || max_batch < 1)
    return fail(DDRL_EINVAL, "ddrl_sac_create: obs_dim, act_dim, h1, h2, max_batch must be >= 1");
  if (2 * act_dim > MAX_HEAD)
    return fail(DDRL_EINVAL, "ddrl_sac_create: act_dim=%d > %d is not supported", act_dim, MAX_HEAD / 2);
  if (gemm_mode < DDRL_GEMM_AUTO || gemm_mode > DDRL_GEMM_FFMA)
    return fail(DDRL_EINVAL, "ddrl_sac_create: gemm_mode=%d is not DDRL_GEMM_AUTO / _TC / _FFMA", gemm_mode);
  int ndev = 0;
  DDRL_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DDRL_EINVAL, "ddrl_sac_create: device %d out of range", device);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_sac_create: cannot select device %d", device);
  ddrl_sac* h = new ddrl_sac();
  h->device = device; h->D = obs_dim; h->A = act_dim; h->h1 = h1; h->h2 = h2; h->maxB = max_batch;
  h->gamma = gamma; h->polyak = polyak; h->lr = lr; h->act_scale = act_scale;
  h->auto_alpha = alpha < 0.0f; h->alpha = alpha;
  h->sms = sm_count(device);
  const char* ng = getenv("DDRL_NO_GRAPH");
  h->use_graph = !(ng && ng[0] == '1');
  if (gemm_mode == DDRL_GEMM_TC) h->use_tc = true;
  else if (gemm_mode == DDRL_GEMM_FFMA) h->use_tc = false;
  else if (const char* gm = getenv("DDRL_GEMM")) h->use_tc = (gm[0] != 'f');   // "ffma": plain fp32 FFMA tiles
  {
    cudaError_t eh = cudaFuncSetAttribute(k_policy_heads_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEADS_SMEM_MAX);
    if (eh == cudaSuccess) eh = cudaFuncSetAttribute(k_policy_heads_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEADS_SMEM_MAX);
    if (eh != cudaSuccess) { delete h; return fail(DDRL_ECUDA, "cudaFuncSetAttribute(heads smem): %s", cudaGetErrorString(eh)); }
  }
  if (h->use_tc) {
    cudaError_t es = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && es == cudaSuccess; ++i) es = cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming);
    if (es != cudaSuccess) { delete h; return fail(DDRL_ECUDA, "side stream / events: %s", cudaGetErrorString(es)); }
    cudaError_t ea = cudaFuncSetAttribute(tc::gemm_grouped_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<128>::SMEM_BYTES);
    if (ea == cudaSuccess) ea = cudaFuncSetAttribute(tc::gemm_grouped_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<64>::SMEM_BYTES);
    if (ea == cudaSuccess) ea = cudaFuncSetAttribute(tc::fwd_fused_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FZ_SMEM_BYTES);
    if (ea != cudaSuccess) { delete h; return fail(DDRL_ECUDA, "cudaFuncSetAttribute(tc smem): %s", cudaGetErrorString(ea)); }
  }
  const int D = obs_dim, A = act_dim;
  {
    // blocks in the reference's variable order; internal starts rounded up to 4 floats
    h->ldh = (2 * A + 3) / 4 * 4;
    const int64_t sizes[9] = {(int64_t)(D + 1) * h1, (int64_t)(h1 + 1) * h2, (int64_t)(h2 + 1) * h->ldh,
                              (int64_t)(D + A + 1) * h1, (int64_t)(h1 + 1) * h2, (int64_t)(h2 + 1),
                              (int64_t)(D + A + 1) * h1, (int64_t)(h1 + 1) * h2, (int64_t)(h2 + 1)};
    int64_t* slots[9] = {&h->o_pi1, &h->o_pi2, &h->o_pih, &h->o_q1[0], &h->o_q1[1], &h->o_q1[2],
                         &h->o_q2[0], &h->o_q2[1], &h->o_q2[2]};
    int64_t o = 0, e = 0;
    h->map.nblk = 9; h->map.head_idx = 2; h->map.h2 = h2; h->map.A = A; h->map.ldh = h->ldh;
    for (int b = 0; b < 9; ++b) {
      o = (o + 3) / 4 * 4;
      if (b == 3) h->P_pi = o;
      *slots[b] = o;
      h->map.ext_off[b] = e; h->map.int_off[b] = o; h->map.size[b] = sizes[b];
      o += sizes[b];
      e += b == 2 ? (int64_t)(h2 + 1) * 2 * A : sizes[b];   // the reference's head tensors are unpadded
    }
    h->P = (o + 3) / 4 * 4;
    h->Pext = e;
  }
  h->Smax = (max_batch + 255) / 256;
  {
    const char* fz = getenv("DDRL_FUSE_L1");
    h->fuse_fwd = h->use_tc && D + A <= tc::BK && h1 % 32 == 0 && h1 <= tc::FZ_MAX_H1 && h2 % 4 == 0 && !(fz && fz[0] == '0');
    if (const char* bz = getenv("DDRL_TC_BN")) { const int v = atoi(bz); if (v == 64 || v == 128) h->force_bn = v; }
    if (const char* dt = getenv("DDRL_DP_TRACE")) {
      if (dt[0] == '1') { float* t = nullptr; dalloc(h, &t, 16); h->dp_trace = reinterpret_cast<unsigned long long*>(t); }
    }
    // one-kernel exchange (split-K reduce + publish + wait + all-peer read + optimiser, k_adam_dp) is the default; DDRL_DP_V1=1
    // selects the two-kernel form (reduce + publish, then wait + all-peer read + optimiser).  Measured on 2 / 4 / 8 B200 (C2,
    // 1000 steps, single-GPU step 92.8 us), with the single-poller wait_flags: one kernel 105.7 / 110.7 / 118.0 us per step,
    // two kernels 108.0 / 113.2 / 120.9 us, NCCL all-reduce 121.0 us at 2 GPUs.  (Before wait_flags polled from one CTA only
    // the order was the opposite: every CTA's own system-scope fence cost more inside the waiting kernel.)
    const char* d1 = getenv("DDRL_DP_V1");
    h->dp_v1 = d1 && d1[0] == '1';

    const char* nz = getenv("DDRL_NARROW_W1");
    h->narrow_w1 = h->use_tc && D + 1 <= NW_MAXK && h1 % 4 == 0 && !(nz && nz[0] == '0');
  }
  int rc = 0;
  const size_t P = (size_t)h->P, M = (size_t)max_batch;
  auto r4 = [](int v) { return (v + 3) / 4 * 4; };
  h->ld1 = h->use_tc ? r4(h1) : h1;
  h->ld2 = h->use_tc ? r4(h2) : h2;
  h->ldx = r4(D + A);
  h->lo1 = h->use_tc ? (long long)M * h->ld1 : 0;
  h->lo2 = h->use_tc ? (long long)M * h->ld2 : 0;
  h->lox = (long long)M * h->ldx;
  const size_t planes = h->use_tc ? 2 : 1;
  auto A_ = [&](float** p, size_t n) { if (!rc) rc = dalloc(h, p, n); };
  A_(&h->W, P); A_(&h->Wt, P); A_(&h->Mo, P); A_(&h->Vo, P); A_(&h->G, P); A_(&h->Gp, P * h->Smax);
  A_(&h->SCAL, 8);
  {
    float* tmp = nullptr;
    A_(&tmp, 2 * 4 * ((size_t)(max_batch + ROW_WARPS - 1) / ROW_WARPS) + 4);
    h->partials = reinterpret_cast<double*>(tmp);
    tmp = nullptr;
    A_(&tmp, 8);
    h->ticket = reinterpret_cast<unsigned int*>(tmp);   // [0] Q-heads kernel, [1] reduce kernel (first form), [2] k_adam_dp, [3] wait_flags relay
  }
  A_(&h->X, M * D); A_(&h->X2, M * D); A_(&h->ACT, M * A); A_(&h->R, M); A_(&h->DN, M); A_(&h->NOISE, 3 * M * A + 4);
  for (int p = 0; p < 8; ++p) { A_(&h->H1[p], planes * M * h->ld1); A_(&h->H2[p], M * h2); }
  for (int p = 0; p < 3; ++p) A_(&h->HD[p], M * 2 * A);
  for (int p = 0; p < 5; ++p) A_(&h->Q[p], M);
  A_(&h->A1, M * A); A_(&h->A3, M * A); A_(&h->LOGP1, M); A_(&h->LOGP2, M);
  for (int p = 0; p < 3; ++p) { A_(&h->dQ[p], M); A_(&h->dZ2[p], planes * M * h->ld2); A_(&h->dZ1[p], planes * M * h->ld1); }
  A_(&h->dA1, M * A); A_(&h->dHD, M * 2 * A); A_(&h->dZ2a, planes * M * h->ld2); A_(&h->dZ1a, planes * M * h->ld1);
  if (h->narrow_w1) A_(&h->Gn, (size_t)((max_batch + NW_ROWS - 1) / NW_ROWS) * (size_t)(D + 1) * h1);
  if (h->use_tc) {
    for (int i = 0; i < 3; ++i) A_(&h->XA[i], 2 * M * h->ldx);
    h->ldbits = (h1 + 31) / 32;
    for (int p = 0; p < 8; ++p) { float* t = nullptr; A_(&t, M * h->ldbits); h->H1bits[p] = reinterpret_cast<uint32_t*>(t); }
    SplitMap& sm = h->smap;
    const int64_t ioff[6] = {h->o_pi1, h->o_pi2, h->o_q1[0], h->o_q1[1], h->o_q2[0], h->o_q2[1]};
    const int Kb[6] = {D, h1, D + A, h1, D + A, h1};
    const int Nb[6] = {h1, h2, h1, h2, h1, h2};
    long long o = 0;
    sm.nblk = 6;
    for (int b = 0; b < 6; ++b) {
      sm.N[b] = Nb[b]; sm.pitch[b] = r4(Nb[b]);
      sm.int_off[b] = ioff[b]; sm.size[b] = (long long)Kb[b] * Nb[b];   // kernel rows only
      sm.sp_off[b] = o;
      o += (long long)Kb[b] * sm.pitch[b];
    }
    sm.plane = o;
    sm.vec4 = (h1 % 4 == 0 && h2 % 4 == 0) ? 1 : 0;
    A_(&h->Wsp, 2 * (size_t)o); A_(&h->Wtsp, 2 * (size_t)o);
  }
  float* stp = nullptr;
  A_(&stp, (sizeof(StepState) + 3) / 4);
  if (rc) { ddrl_sac_destroy(h); return rc; }
  if (cudaMalloc((void**)&h->d_err, sizeof(int)) != cudaSuccess || cudaMemset(h->d_err, 0, sizeof(int)) != cudaSuccess) {
    ddrl_sac_destroy(h);
    return fail(DDRL_ENOMEM, "cudaMalloc(error flag) failed");
  }
  h->st = reinterpret_cast<StepState*>(stp);
  StepState init{};
  init.auto_alpha = h->auto_alpha;
  init.alpha_const = alpha;
  init.alpha_cur = h->auto_alpha ? 1.0f : alpha;
  init.lr = lr;
  init.trace = h->dp_trace;
  cudaError_t e = cudaMemcpy(h->st, &init, sizeof(init), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { ddrl_sac_destroy(h); return fail(DDRL_ECUDA, "init state: %s", cudaGetErrorString(e)); }
  *out = h;
  return 0;
}

int ddrl_sac_destroy(ddrl_sac_t h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->plans) {
    Plan& pl = kv.second;
    for (auto ex : {pl.exec_full, pl.exec_grads, pl.exec_apply, pl.exec_dp}) if (ex) cudaGraphExecDestroy(ex);
    for (auto g : pl.graph) if (g) cudaGraphDestroy(g);
  }
  if (h->dbg_exec) cudaGraphExecDestroy(h->dbg_exec);
  for (int r = 0; r < 8; ++r) if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->pc.buf[r]);
  if (h->comm) cudaFree(h->comm);
  if (h->d_err) cudaFree(h->d_err);
  for (void* p : h->allocs) cudaFree(p);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (auto e : h->ev) if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

int64_t ddrl_sac_param_count(ddrl_sac_t h) { return h ? h->Pext : -1; }

int ddrl_sac_set_weights(ddrl_sac_t h, const float* d_flat, int also_target, void* stream) {
  if (!h || !d_flat) return fail(DDRL_EINVAL, "ddrl_sac_set_weights: NULL argument");
  DeviceGuard guard(h->device);
  int blocks = (int)std::min<int64_t>((h->Pext + 255) / 256, h->sms * 8);
  k_convert_layout<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->map, h->Pext, 1, d_flat, h->W, also_target ? h->Wt : nullptr);
  DDRL_LAUNCH_CHECK();
  if (h->use_tc) {
    k_split_weights<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->smap, h->P, h->W, h->Wt, h->Wsp, h->Wtsp);
    DDRL_LAUNCH_CHECK();
  }
  return 0;
}

int ddrl_sac_get_weights(ddrl_sac_t h, float* d_flat, int which, void* stream) {
  if (!h || !d_flat) return fail(DDRL_EINVAL, "ddrl_sac_get_weights: NULL argument");
  const float* src = which == 0 ? h->W : which == 1 ? h->Wt : which == 2 ? h->Mo : which == 3 ? h->Vo : which == 4 ? h->G : nullptr;
  if (!src) return fail(DDRL_EINVAL, "ddrl_sac_get_weights: which=%d not in 0..4", which);
  DeviceGuard guard(h->device);
  int blocks = (int)std::min<int64_t>((h->Pext + 255) / 256, h->sms * 8);
  k_convert_layout<<<blocks, 256, 0, (cudaStream_t)stream>>>(h->map, h->Pext, 0, src, d_flat, nullptr);
  DDRL_LAUNCH_CHECK();
  return 0;
}

struct RingSrc {   // fused sample -> update: where the prologue gathers the batch from
  const float* ring = nullptr;
  int row_f = 0;
  uint32_t stream = 0;
  uint64_t seed = 0, counter = 0, size = 0;
};
static int step_common(ddrl_sac_t h, int mode, const float* d_obs1, const float* d_obs2, const float* d_acts,
                       const float* d_rews, const float* d_done, int batch, const float* d_noise, uint64_t seed,
                       float grad_scale, float* d_out_scalars, float* d_out_q1, float* d_out_q2, float* d_out_logp,
                       void* stream, const char* who, const RingSrc* rs = nullptr) {
  if (!h) return fail(DDRL_EINVAL, "%s: NULL handle", who);
  if (batch < 1 || batch > h->maxB) return fail(DDRL_EINVAL, "%s: batch=%d not in [1, %d]", who, batch, h->maxB);
  if (mode != MODE_APPLY && !rs && (!d_obs1 || !d_obs2 || !d_acts || !d_rews || !d_done))
    return fail(DDRL_EINVAL, "%s: NULL batch array", who);
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  Plan* pl = nullptr;
  int rc = get_plan(h, batch, &pl);
  if (rc) return rc;
  if (mode != MODE_APPLY) {
    StepDyn dyn{};
    dyn.obs1 = d_obs1; dyn.obs2 = d_obs2; dyn.acts = d_acts; dyn.rews = d_rews; dyn.done = d_done;
    dyn.noise = d_noise;
    dyn.out_scalars = d_out_scalars; dyn.out_q1 = d_out_q1; dyn.out_q2 = d_out_q2; dyn.out_logp = d_out_logp;
    dyn.seed = seed; dyn.grad_scale = grad_scale;
    if (rs) {
      dyn.ring = rs->ring; dyn.ring_row_f = rs->row_f; dyn.ring_stream = rs->stream;
      dyn.ring_seed = rs->seed; dyn.ring_counter = rs->counter; dyn.ring_size = rs->size;
    }
    h->t_host += 1;
    const double t = (double)h->t_host;
    dyn.t_pi = dyn.t_q = h->t_host;
    dyn.lr_pi = dyn.lr_q = (float)((double)h->lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
    dyn.noise_counter = (unsigned long long)h->t_host;
    h->last_dyn = dyn;
    // the prologue carries the per-step values as kernel parameters of the graph's first node
    return run_mode(h, *pl, mode, s, &dyn);
  }
  return run_mode(h, *pl, mode, s, nullptr);
}

int ddrl_sac_step(ddrl_sac_t h, const float* d_obs1, const float* d_obs2, const float* d_acts, const float* d_rews,
                  const float* d_done, int batch, const float* d_noise, uint64_t seed, float* d_out_scalars,
                  float* d_out_q1, float* d_out_q2, float* d_out_logp, void* stream) {
  return step_common(h, MODE_FULL, d_obs1, d_obs2, d_acts, d_rews, d_done, batch, d_noise, seed, 1.0f, d_out_scalars,
                     d_out_q1, d_out_q2, d_out_logp, stream, "ddrl_sac_step");
}

int ddrl_sac_compute_grads(ddrl_sac_t h, const float* d_obs1, const float* d_obs2, const float* d_acts,
                           const float* d_rews, const float* d_done, int batch, const float* d_noise, uint64_t seed,
                           float grad_scale, float* d_out_scalars, float* d_out_q1, float* d_out_q2, float* d_out_logp,
                           void* stream) {
  return step_common(h, MODE_GRADS, d_obs1, d_obs2, d_acts, d_rews, d_done, batch, d_noise, seed, grad_scale,
                     d_out_scalars, d_out_q1, d_out_q2, d_out_logp, stream, "ddrl_sac_compute_grads");
}

int ddrl_sac_step_dp(ddrl_sac_t h, const float* d_obs1, const float* d_obs2, const float* d_acts, const float* d_rews,
                     const float* d_done, int batch, const float* d_noise, uint64_t seed, float grad_scale,
                     float* d_out_scalars, float* d_out_q1, float* d_out_q2, float* d_out_logp, void* stream) {
  if (h && h->pc.world < 2) return fail(DDRL_ESTATE, "ddrl_sac_step_dp: no peers attached (ddrl_sac_comm_attach)");
  return step_common(h, MODE_DP, d_obs1, d_obs2, d_acts, d_rews, d_done, batch, d_noise, seed, grad_scale, d_out_scalars,
                     d_out_q1, d_out_q2, d_out_logp, stream, "ddrl_sac_step_dp");
}

int ddrl_sac_step_from_buffer(ddrl_sac_t h, ddrl_rb_t rb, int batch, uint64_t rb_seed, uint64_t rb_counter,
                              uint32_t rb_stream, const float* d_noise, uint64_t seed, float* d_out_scalars, float* d_out_q1,
                              float* d_out_q2, float* d_out_logp, void* stream) {
  if (!h || !rb) return fail(DDRL_EINVAL, "ddrl_sac_step_from_buffer: NULL handle");
  int D = 0, A = 0, row_f = 0;
  void* ring = nullptr;
  int rc = ddrl_rb_layout(rb, &D, &A, &row_f, &ring);
  if (rc) return rc;
  if (D != h->D || A != h->A)
    return fail(DDRL_EINVAL, "ddrl_sac_step_from_buffer: buffer rows are (obs %d, act %d), learner expects (%d, %d)", D, A, h->D, h->A);
  // the ring is read by this step's first kernel: order it after stores issued on other streams (read_begin holds the
  // buffer's lock until read_end, so no store can slip between the size snapshot and the launch)
  int64_t size = 0;
  if ((rc = ddrl_rb_read_begin(rb, stream, &size))) return rc;
  if (size == 0) {
    ddrl_rb_read_end(rb, stream, 0);
    return fail(DDRL_EEMPTY, "ddrl_sac_step_from_buffer: ring is empty (the reference raises ValueError: high <= 0)");
  }
  RingSrc rs;
  rs.ring = (const float*)ring; rs.row_f = row_f; rs.stream = rb_stream; rs.seed = rb_seed; rs.counter = rb_counter;
  rs.size = (uint64_t)size;
  const int mode = h->pc.world > 1 ? MODE_DP : MODE_FULL;
  const float gs = h->pc.world > 1 ? 1.0f / (float)h->pc.world : 1.0f;
  rc = step_common(h, mode, nullptr, nullptr, nullptr, nullptr, nullptr, batch, d_noise, seed, gs, d_out_scalars, d_out_q1,
                   d_out_q2, d_out_logp, stream, "ddrl_sac_step_from_buffer", &rs);
  const int rc2 = ddrl_rb_read_end(rb, stream, rc ? 0 : 1);
  return rc ? rc : rc2;
}

int ddrl_sac_step_host(ddrl_sac_t h, const void* h_block, int batch, uint64_t seed, float* h_out_scalars, float* d_out_q1,
                       float* d_out_q2, float* d_out_logp, void* stream) {
  if (!h || !h_block) return fail(DDRL_EINVAL, "ddrl_sac_step_host: NULL argument");
  if (batch < 1 || batch > h->maxB) return fail(DDRL_EINVAL, "ddrl_sac_step_host: batch=%d not in [1, %d]", batch, h->maxB);
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t row = (size_t)(2 * h->D + h->A + 2);
  if (!h->host_stage) {
    int rc = dalloc(h, &h->host_stage, (size_t)h->maxB * row);
    if (!rc) rc = dalloc(h, &h->host_scal, 4);
    if (rc) return rc;
  }
  // Small blocks in PINNED host memory are read by the step's first kernel straight through the block's device mapping
  // (the prologue is a copy kernel anyway), and the four scalars are written straight into the caller's pinned array:
  // no H2D / D2H DMA launches (~10 us of latency each) on the critical path.  Otherwise: one H2D copy of the block
  // (stream order protects the staging: the previous step's prologue has consumed it), the step, one D2H copy of the
  // scalars.  Nothing here waits for the GPU.
  static const bool zero_copy = [] { const char* e = getenv("DDRL_ZERO_COPY"); return !(e && e[0] == '0'); }();
  const size_t B = (size_t)batch;
  float* x = zero_copy && B * row * sizeof(float) <= (2u << 20) ? (float*)host_device_pointer(h_block) : nullptr;
  float* scal = x && h_out_scalars ? (float*)host_device_pointer(h_out_scalars) : nullptr;
  if (!x) {
    DDRL_CUDA(cudaMemcpyAsync(h->host_stage, h_block, B * row * sizeof(float), cudaMemcpyHostToDevice, s));
    x = h->host_stage;
  }
  float* x2 = x + B * h->D;
  float* a = x2 + B * h->D;
  float* r = a + B * h->A;
  float* d = r + B;
  const bool dp = h->pc.world > 1;
  int rc = step_common(h, dp ? MODE_DP : MODE_FULL, x, x2, a, r, d, batch, nullptr, seed, dp ? 1.0f / (float)h->pc.world : 1.0f,
                       scal ? scal : h->host_scal, d_out_q1, d_out_q2, d_out_logp, stream, "ddrl_sac_step_host");
  if (rc) return rc;
  if (h_out_scalars && !scal) DDRL_CUDA(cudaMemcpyAsync(h_out_scalars, h->host_scal, 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  return 0;
}

int ddrl_sac_grad_buffer(ddrl_sac_t h, float** d_grads, int64_t* count, float** d_alpha_stat) {
  if (!h) return fail(DDRL_EINVAL, "ddrl_sac_grad_buffer: NULL handle");
  if (d_grads) *d_grads = h->G;
  if (count) *count = h->P;
  if (d_alpha_stat) *d_alpha_stat = h->SCAL + 4;
  return 0;
}

int ddrl_sac_comm_export(ddrl_sac_t h, void* h_handle64) {
  if (!h || !h_handle64) return fail(DDRL_EINVAL, "ddrl_sac_comm_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  DeviceGuard guard(h->device);
  if (!h->comm) {
    const long long Pc = (h->P + 4 + 3) / 4 * 4;
    h->pc.nslice = 1;     // flags: [0, 8) "gradient ready" per source rank
    const size_t bytes = (size_t)2 * Pc * sizeof(float) + (8 + 8 * (size_t)h->pc.nslice) * sizeof(unsigned int);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);   // its own allocation: CUDA IPC shares whole allocations
    if (e != cudaSuccess) return fail(DDRL_ENOMEM, "cudaMalloc(comm buffer) failed: %s", cudaGetErrorString(e));
    DDRL_CUDA(cudaMemset(p, 0, bytes));
    h->comm = (float*)p;
    h->pc.Pc = Pc;
    if (!h->d_err) {
      DDRL_CUDA(cudaMalloc((void**)&h->d_err, sizeof(int)));
      DDRL_CUDA(cudaMemset(h->d_err, 0, sizeof(int)));
    }
  }
  cudaIpcMemHandle_t hd;
  DDRL_CUDA(cudaIpcGetMemHandle(&hd, h->comm));
  memcpy(h_handle64, &hd, 64);
  return 0;
}

static long long comm_pc(const ddrl_sac* h) { return (h->P + 4 + 3) / 4 * 4; }

int64_t ddrl_sac_comm_bytes(ddrl_sac_t h) {
  if (!h) return -1;
  return (int64_t)((size_t)2 * comm_pc(h) * sizeof(float) + (8 + 8) * sizeof(unsigned int));
}

int ddrl_sac_comm_attach_ptrs(ddrl_sac_t h, int world, int rank, void* const* d_bufs, const void* d_multicast) {
  if (!h || !d_bufs) return fail(DDRL_EINVAL, "ddrl_sac_comm_attach_ptrs: NULL argument");
  if (world < 2 || world > 8 || rank < 0 || rank >= world)
    return fail(DDRL_EINVAL, "ddrl_sac_comm_attach_ptrs: rank %d of %d (2..8 ranks of one node)", rank, world);
  if (h->t_host != 0) return fail(DDRL_ESTATE, "ddrl_sac_comm_attach_ptrs: attach before the first update (ranks step in lockstep)");
  for (int r = 0; r < world; ++r)
    if (!d_bufs[r] || ((uintptr_t)d_bufs[r] & 15)) return fail(DDRL_EINVAL, "ddrl_sac_comm_attach_ptrs: buffer %d is NULL or not 16-byte aligned", r);
  if (((uintptr_t)d_multicast & 15)) return fail(DDRL_EINVAL, "ddrl_sac_comm_attach_ptrs: multicast pointer not 16-byte aligned");
  DeviceGuard guard(h->device);
  h->pc.nslice = 1;
  h->pc.Pc = comm_pc(h);
  if (!h->d_err) {
    DDRL_CUDA(cudaMalloc((void**)&h->d_err, sizeof(int)));
    DDRL_CUDA(cudaMemset(h->d_err, 0, sizeof(int)));
  }
  for (int r = 0; r < world; ++r) {
    h->pc.buf[r] = (float*)d_bufs[r];
    h->pc.flags[r] = reinterpret_cast<unsigned int*>((float*)d_bufs[r] + 2 * h->pc.Pc);
  }
  h->pc.mc = (const float*)d_multicast;
  h->pc.world = world; h->pc.rank = rank;
  h->pc.relay = h->ticket + 3;
  for (auto& kv : h->plans)
    for (cudaGraphExec_t* ex : {&kv.second.exec_full, &kv.second.exec_grads, &kv.second.exec_apply, &kv.second.exec_dp})
      if (*ex) { cudaGraphExecDestroy(*ex); *ex = nullptr; }
  return 0;
}

int ddrl_sac_comm_attach(ddrl_sac_t h, int world, int rank, const void* h_handles) {
  if (!h || !h_handles) return fail(DDRL_EINVAL, "ddrl_sac_comm_attach: NULL argument");
  if (world < 2 || world > 8 || rank < 0 || rank >= world)
    return fail(DDRL_EINVAL, "ddrl_sac_comm_attach: rank %d of %d (2..8 ranks of one node)", rank, world);
  if (!h->comm) return fail(DDRL_ESTATE, "ddrl_sac_comm_attach: call ddrl_sac_comm_export first");
  if (h->t_host != 0) return fail(DDRL_ESTATE, "ddrl_sac_comm_attach: attach before the first update (ranks step in lockstep)");
  DeviceGuard guard(h->device);
  for (int r = 0; r < world; ++r) {
    void* p = h->comm;
    if (r != rank) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)h_handles + 64 * r, 64);
      p = nullptr;
      DDRL_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      h->peer_opened[r] = true;
    }
    h->pc.buf[r] = (float*)p;
    h->pc.flags[r] = reinterpret_cast<unsigned int*>((float*)p + 2 * h->pc.Pc);
  }
  h->pc.world = world; h->pc.rank = rank;
  h->pc.relay = h->ticket + 3;
  // graphs captured before the attach hold the single-GPU kernels
  for (auto& kv : h->plans)
    for (cudaGraphExec_t* ex : {&kv.second.exec_full, &kv.second.exec_grads, &kv.second.exec_apply, &kv.second.exec_dp})
      if (*ex) { cudaGraphExecDestroy(*ex); *ex = nullptr; }
  return 0;
}

int ddrl_sac_dp_trace(ddrl_sac_t h, unsigned long long* h_out8) {
  if (!h || !h_out8) return fail(DDRL_EINVAL, "ddrl_sac_dp_trace: NULL argument");
  if (!h->dp_trace) return fail(DDRL_ESTATE, "ddrl_sac_dp_trace: create the handle with DDRL_DP_TRACE=1 in the environment");
  DeviceGuard guard(h->device);
  DDRL_CUDA(cudaDeviceSynchronize());
  DDRL_CUDA(cudaMemcpy(h_out8, h->dp_trace, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int ddrl_sac_comm_error(ddrl_sac_t h, int* out) {
  if (!h || !out) return fail(DDRL_EINVAL, "ddrl_sac_comm_error: NULL argument");
  *out = 0;
  if (!h->d_err) return 0;
  DeviceGuard guard(h->device);
  DDRL_CUDA(cudaMemcpy(out, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int ddrl_sac_apply_grads(ddrl_sac_t h, int batch, void* stream) {
  return step_common(h, MODE_APPLY, nullptr, nullptr, nullptr, nullptr, nullptr, batch, nullptr, 0, 1.0f, nullptr,
                     nullptr, nullptr, nullptr, stream, "ddrl_sac_apply_grads");
}

int ddrl_sac_act(ddrl_sac_t h, const float* d_obs, int n, int deterministic, const float* d_noise, uint64_t seed,
                 uint64_t counter, float* d_out_act, void* stream) {
  if (!h || !d_obs || !d_out_act) return fail(DDRL_EINVAL, "ddrl_sac_act: NULL argument");
  if (n < 1) return fail(DDRL_EINVAL, "ddrl_sac_act: n=%d < 1", n);
  if (h->h1 > ACT_MAX_H || h->h2 > ACT_MAX_H)
    return fail(DDRL_EINVAL, "ddrl_sac_act: hidden sizes above %d are not supported", ACT_MAX_H);
  DeviceGuard guard(h->device);
  k_actor_forward<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(n, h->D, h->A, h->h1, h->h2, h->ldh, h->act_scale, deterministic, d_obs,
                                                                 h->W + h->o_pi1, h->W + h->o_pi2, h->W + h->o_pih, d_noise, seed,
                                                                 counter, d_out_act);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_sac_debug_stage(ddrl_sac_t h, int batch, int stage, int reps, void* stream) {
  if (!h) return fail(DDRL_EINVAL, "ddrl_sac_debug_stage: NULL handle");
  DeviceGuard guard(h->device);
  Plan* pl = nullptr;
  int rc = get_plan(h, batch, &pl);
  if (rc) return rc;
  // 0..6: GEMM stages (tensor-core tiles only in tc mode); 7 prologue, 8 policy heads, 9 Q heads + losses, 10 policy
  // backward rows, 11 optimiser (state advances!), 12..14: side-stream work of BQ / BP / BP3
  if (stage < 0 || stage > 14) return fail(DDRL_EINVAL, "ddrl_sac_debug_stage: stage %d not in [0,14]", stage);
  cudaStream_t s = (cudaStream_t)stream;
  if (reps < 0) {
    // -reps launches as the nodes of ONE graph: what a kernel costs inside the step's graph.  (Back-to-back launches on
    // a stream complete on a ~2.05 us grid — tools/probes/tick_probe.cu — which rounds every kernel up to the next tick.)
    // the graph of the last (batch, stage, reps) is kept: call once to build + warm up, then time a second call
    if (h->dbg_exec && h->dbg_key[0] == batch && h->dbg_key[1] == stage && h->dbg_key[2] == reps) {
      DDRL_CUDA(cudaGraphLaunch(h->dbg_exec, s));
      g_launches.fetch_add(-reps, std::memory_order_relaxed);
      return 0;
    }
    if (h->dbg_exec) { cudaGraphExecDestroy(h->dbg_exec); h->dbg_exec = nullptr; }
    h->dbg_key[0] = batch; h->dbg_key[1] = stage; h->dbg_key[2] = reps;
    if (!h->cap_stream) DDRL_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    const int64_t before = g_launches.load();
    DDRL_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    rc = ddrl_sac_debug_stage(h, batch, stage, -reps, h->cap_stream);
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    g_launches.store(before + (rc ? 0 : g_launches.load() - before));
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(DDRL_ECUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->dbg_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(DDRL_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    DDRL_CUDA(cudaGraphLaunch(h->dbg_exec, s));
    return 0;
  }
  for (int i = 0; i < reps; ++i) {
    if (stage < ST_COUNT) rc = run_stage(*pl, stage, s, h->use_tc);
    else if (stage == 7) rc = launch_prologue(h, *pl, h->last_dyn, s);
    else if (stage == 8) rc = launch_heads(h, *pl, s, PassMap{2, {0, 2, 0}});
    else if (stage == 9) rc = launch_qheads(h, *pl, s);
    else if (stage == 10) rc = launch_pbwd(h, *pl, s);
    else if (stage == 11) rc = enqueue_apply(h, pl->S, h->Gp, s, pl->B);
    else rc = h->use_tc ? run_side(*pl, ST_BQ + (stage - 12), pl->S, s) : 0;
    if (rc) return rc;
  }
  return 0;
}

// Test entry for the tensor-core GEMM alone: C[M,N] = opA . opB from plain fp32 device matrices.  The stored
// tensors are A [a_rows, a_cols] and B [b_rows, b_cols] (row-major, dense); a_mn / b_mn = 1 when the contraction
// index is the ROW of the stored tensor.  Splits into hi/lo planes, builds the TMA maps, runs one launch.
__global__ void k_split_planes(const float* __restrict__ src, int rows, int cols, int ld, long long plane, float* dst) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    put_split(dst, plane, (size_t)r * ld + c, src[i]);
  }
}
int ddrl_debug_tc_gemm(int device, const float* dA, int a_rows, int a_cols, int a_mn, const float* dB, int b_rows, int b_cols,
                       int b_mn, float* dC, int M, int N, int K, int splits, int bn, void* stream) {
  if (!dA || !dB || !dC) return fail(DDRL_EINVAL, "ddrl_debug_tc_gemm: NULL argument");
  if (bn != 64 && bn != 128) return fail(DDRL_EINVAL, "ddrl_debug_tc_gemm: bn must be 64 or 128");
  DeviceGuard guard(device);
  cudaStream_t s = (cudaStream_t)stream;
  DDRL_CUDA(cudaFuncSetAttribute(tc::gemm_grouped_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<128>::SMEM_BYTES));
  DDRL_CUDA(cudaFuncSetAttribute(tc::gemm_grouped_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<64>::SMEM_BYTES));
  auto r4 = [](int v) { return (v + 3) / 4 * 4; };
  const int lda = r4(a_cols), ldb = r4(b_cols);
  const long long pa = (long long)a_rows * lda, pb = (long long)b_rows * ldb;
  float *sa = nullptr, *sb = nullptr;
  DDRL_CUDA(cudaMalloc(&sa, 2 * pa * sizeof(float)));
  DDRL_CUDA(cudaMalloc(&sb, 2 * pb * sizeof(float)));
  cudaMemsetAsync(sa, 0, 2 * pa * sizeof(float), s);
  cudaMemsetAsync(sb, 0, 2 * pb * sizeof(float), s);
  k_split_planes<<<256, 256, 0, s>>>(dA, a_rows, a_cols, lda, pa, sa);
  k_split_planes<<<256, 256, 0, s>>>(dB, b_rows, b_cols, ldb, pb, sb);
  Group g;
  tc::TcProb p;
  g.bn = bn;
  int rc = mk_tc(&p, bn, View{sa, lda, pa, a_rows, a_cols}, a_mn != 0, View{sb, ldb, pb, b_rows, b_cols}, b_mn != 0, dC, nullptr, N, M,
                 N, K);
  if (!rc) {
    if (splits > 1) { p.splits = splits; p.k_per_split = ((K + splits - 1) / splits + tc::BK - 1) / tc::BK * tc::BK; p.c_split_stride = (long long)M * N; }
    g.probs_tc.push_back(p);
    rc = finalize_group(g);
  }
  if (!rc) rc = launch_group(g, s);
  cudaStreamSynchronize(s);
  cudaFree(sa);
  cudaFree(sb);
  return rc;
}

// profiling aid: run GEMM stage `stage` once with per-CTA phase time stamps (ns, %globaltimer): d_trace [tiles, 8]
int ddrl_sac_trace_stage(ddrl_sac_t h, int batch, int stage, unsigned long long* d_trace, int max_tiles, int* tiles, void* stream) {
  if (!h || !d_trace || !tiles) return fail(DDRL_EINVAL, "ddrl_sac_trace_stage: NULL argument");
  if (stage < 0 || stage >= ST_COUNT) return fail(DDRL_EINVAL, "ddrl_sac_trace_stage: stage %d", stage);
  DeviceGuard guard(h->device);
  Plan* pl = nullptr;
  int rc = get_plan(h, batch, &pl);
  if (rc) return rc;
  *tiles = 0;
  for (auto& g : pl->stages[stage]) {
    if (g.tiles_fz > 0) {
      if (g.tiles_fz > max_tiles) return fail(DDRL_EINVAL, "ddrl_sac_trace_stage: %d tiles > max_tiles", g.tiles_fz);
      tc::FusedGroup grp = g.grp_fz;
      grp.trace = d_trace;
      DDRL_CUDA(launch_pdl(tc::fwd_fused_tc, dim3(g.tiles_fz), dim3(tc::FZ_THREADS), tc::FZ_SMEM_BYTES, (cudaStream_t)stream, grp));
      DDRL_LAUNCH_CHECK();
      *tiles = g.tiles_fz;
      continue;
    }
    if (g.tiles_tc <= 0) continue;
    if (g.tiles_tc > max_tiles) return fail(DDRL_EINVAL, "ddrl_sac_trace_stage: %d tiles > max_tiles", g.tiles_tc);
    tc::TcGroup grp = g.grp_tc;
    grp.trace = d_trace;
    if (g.bn == 64)
      DDRL_CUDA(launch_pdl(tc::gemm_grouped_tc<64>, dim3(g.tiles_tc), dim3(256), tc::Cfg<64>::SMEM_BYTES, (cudaStream_t)stream, grp));
    else
      DDRL_CUDA(launch_pdl(tc::gemm_grouped_tc<128>, dim3(g.tiles_tc), dim3(256), tc::Cfg<128>::SMEM_BYTES, (cudaStream_t)stream, grp));
    DDRL_LAUNCH_CHECK();
    *tiles = g.tiles_tc;
  }
  return 0;
}

int ddrl_sac_state(ddrl_sac_t h, int* t_pi, int* t_q, int* t_alpha, float* log_alpha, void* stream) {
  if (!h) return fail(DDRL_EINVAL, "ddrl_sac_state: NULL handle");
  DeviceGuard guard(h->device);
  StepState s{};
  DDRL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  DDRL_CUDA(cudaMemcpy(&s, h->st, sizeof(s), cudaMemcpyDeviceToHost));
  if (t_pi) *t_pi = s.t_pi;
  if (t_q) *t_q = s.t_q;
  if (t_alpha) *t_alpha = s.t_alpha;
  if (log_alpha) *log_alpha = s.log_alpha;
  return 0;
}

}  // extern "C"
