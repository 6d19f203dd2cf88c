// sac_gemm.cuh — grouped fp32 GEMM used by the SAC1 learner step (parity mode: plain FFMA, no
// TF32/bf16 rounding, so losses and updated weights stay within 1e-5 of the float64 oracle).
//
// One launch executes a GROUP of independent problems  C = epi( opA · opB )  (the 3 policy passes,
// 5 Q passes, or the dgrad/wgrad set of one backward level), each described by a GemmProb.
//   - A is always stored [batch, features] row-major and may be a VIRTUAL CONCAT of two tensors
//     plus a constant-one column:  A = [ a0 | a1 | 1 ].  The one column turns the bias into the last
//     row of the weight block, so every layer's parameters are one contiguous [K+1, N] block
//     (kernel rows, then the bias row) and  dense(x) = [x|1] · [W;b],  d[W;b] = [x|1]^T · dZ.
//     a_trans = 0 : opA = A        (forward / dgrad:  M = batch,     K = features)
//     a_trans = 1 : opA = A^T      (wgrad:            M = features,  K = batch; optional split-K)
//   - B is row-major [K,N] (b_trans = 0) or stored [N,K] (b_trans = 1: dgrad against W^T).
//   - epilogue: none | relu | multiply by (mask > 0)   (relu backward).
#pragma once
#include "common.cuh"

namespace ddrl {

struct Seg {
  const float* p;
  int ld;
  int w;
};

enum Epi : int { EPI_NONE = 0, EPI_RELU = 1, EPI_MASK = 2 };

struct GemmProb {
  Seg a0, a1;
  int a_ones;
  int a_trans;
  const float* B;
  int ldb;
  int b_trans;
  float* C;
  int ldc;
  long long c_split_stride;  // floats between split-K partial outputs
  int M, N, K;
  int epi;
  const float* mask;
  int ldmask;
  int splits, k_per_split;
  int tiles_m, tiles_n, tile_begin;
  int cfg;  // tile shape: 0 = 64x64 (4x4 per thread), 1 = 128x16 (4x2 per thread, skinny N)
};

// predicated (branch-free) global load: all operand loads of a tile are issued back to back; a
// branchy formulation serialises them on the L2 latency (measured: 8000+ cycles per k-block)
__device__ __forceinline__ float ld_pred(const float* p, bool pred) {
  float v;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.f32 %0, [%1];\n\t}"
      : "=f"(v) : "l"(p), "r"((int)pred));
  return v;
}

// element (row, feat) of the virtual concat A = [a0 | a1 | 1]; `ok` = inside the tile's bounds
__device__ __forceinline__ float fetch_a(const GemmProb& P, int row, int feat, bool ok = true) {
  const bool s0 = feat < P.a0.w;
  const int f1 = feat - P.a0.w;
  const bool s1 = !s0 && f1 < P.a1.w;
  const float* p = s0 ? P.a0.p + (size_t)row * P.a0.ld + feat : P.a1.p + (size_t)row * P.a1.ld + f1;
  const float v = ld_pred(p, ok && (s0 || s1));
  return (ok && !s0 && !s1 && P.a_ones && f1 == P.a1.w) ? 1.0f : v;
}

// global -> register fetch of one (BM x BK) A tile slice and one (BK x BN) B tile slice
template <int BM, int BN, int BK>
struct TileRegs {
  float a[BM * BK / 256];
  float b[BN * BK / 256];
};

template <int BM, int BN, int BK>
__device__ __forceinline__ void fetch_tiles(const GemmProb& P, int tid, int m0, int n0, int k0, int kend,
                                            TileRegs<BM, BN, BK>& r) {
  if (!P.a_trans) {
#pragma unroll
    for (int e = 0; e < BM * BK / 256; ++e) {
      const int i = tid + e * 256;
      const int kl = i % BK, ml = i / BK;
      const int m = m0 + ml, k = k0 + kl;
      r.a[e] = fetch_a(P, m, k, m < P.M && k < kend);
    }
  } else {
#pragma unroll
    for (int e = 0; e < BM * BK / 256; ++e) {
      const int i = tid + e * 256;
      const int ml = i % BM, kl = i / BM;
      const int m = m0 + ml, k = k0 + kl;
      r.a[e] = fetch_a(P, k, m, m < P.M && k < kend);
    }
  }
  if (!P.b_trans) {
#pragma unroll
    for (int e = 0; e < BN * BK / 256; ++e) {
      const int i = tid + e * 256;
      const int nl = i % BN, kl = i / BN;
      const int n = n0 + nl, k = k0 + kl;
      r.b[e] = ld_pred(P.B + (size_t)k * P.ldb + n, n < P.N && k < kend);
    }
  } else {
#pragma unroll
    for (int e = 0; e < BN * BK / 256; ++e) {
      const int i = tid + e * 256;
      const int kl = i % BK, nl = i / BK;
      const int n = n0 + nl, k = k0 + kl;
      r.b[e] = ld_pred(P.B + (size_t)n * P.ldb + k, n < P.N && k < kend);
    }
  }
}

constexpr int GEMM_BK = 32;
constexpr int GEMM_SMEM_FLOATS = GEMM_BK * (128 + 4) + GEMM_BK * (64 + 4);

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_tile(const GemmProb& P, float* smem_raw, int tile_in_prob) {
  constexpr int BK = GEMM_BK;
  constexpr int TX = BN / TN;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads per tile");
  float (*As)[BM + 4] = reinterpret_cast<float (*)[BM + 4]>(smem_raw);
  float (*Bs)[BN + 4] = reinterpret_cast<float (*)[BN + 4]>(smem_raw + BK * (BM + 4));
  const int tid = threadIdx.x;

  int t = tile_in_prob;
  const int per_split = P.tiles_m * P.tiles_n;
  const int split = t / per_split;
  t -= split * per_split;
  const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
  const int kbeg = split * P.k_per_split;
  const int kend = min(P.K, kbeg + P.k_per_split);
  const int tx = tid % TX, ty = tid / TX;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  TileRegs<BM, BN, BK> r;
  fetch_tiles<BM, BN, BK>(P, tid, m0, n0, kbeg, kend, r);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // registers -> shared (same index maps as fetch_tiles)
    if (!P.a_trans) {
#pragma unroll
      for (int e = 0; e < BM * BK / 256; ++e) { const int i = tid + e * 256; As[i % BK][i / BK] = r.a[e]; }
    } else {
#pragma unroll
      for (int e = 0; e < BM * BK / 256; ++e) { const int i = tid + e * 256; As[i / BM][i % BM] = r.a[e]; }
    }
    if (!P.b_trans) {
#pragma unroll
      for (int e = 0; e < BN * BK / 256; ++e) { const int i = tid + e * 256; Bs[i / BN][i % BN] = r.b[e]; }
    } else {
#pragma unroll
      for (int e = 0; e < BN * BK / 256; ++e) { const int i = tid + e * 256; Bs[i % BK][i / BK] = r.b[e]; }
    }
    __syncthreads();
    if (k0 + BK < kend) fetch_tiles<BM, BN, BK>(P, tid, m0, n0, k0 + BK, kend, r);  // in flight during the FMAs
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* C = P.C + (size_t)split * P.c_split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= P.N) continue;
      float v = acc[i][j];
      if (P.epi == EPI_RELU) v = fmaxf(v, 0.0f);
      else if (P.epi == EPI_MASK) v = (P.mask[(size_t)m * P.ldmask + n] > 0.0f) ? v : 0.0f;
      C[(size_t)m * P.ldc + n] = v;
    }
  }
}

// The problem descriptors travel as kernel parameters (constant bank): a CTA finds its problem
// without any dependent global load (a descriptor array in global memory cost ~4 us per launch).
constexpr int GEMM_MAX_PROBS = 12;
struct GemmGroup {
  int nprob;
  GemmProb p[GEMM_MAX_PROBS];
};

// internal linkage: the header is compiled into sac.cu (SAC1 step) and qlearn.cu (DDQN / SQN steps)
static __global__ void __launch_bounds__(256, 3) gemm_grouped_f32(const __grid_constant__ GemmGroup grp) {
  __shared__ __align__(16) float smem_raw[GEMM_SMEM_FLOATS];
  pdl_trigger();
  pdl_wait();
  int pi = 0;
  while (pi + 1 < grp.nprob && (int)blockIdx.x >= grp.p[pi + 1].tile_begin) ++pi;
  const GemmProb P = grp.p[pi];   // into registers (indexed constant-bank reads in the inner loops are slow)
  if (P.cfg == 0) gemm_tile<64, 64, 4, 4>(P, smem_raw, blockIdx.x - P.tile_begin);
  else gemm_tile<128, 16, 4, 2>(P, smem_raw, blockIdx.x - P.tile_begin);
}

}  // namespace ddrl
