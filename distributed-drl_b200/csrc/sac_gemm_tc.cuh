// sac_gemm_tc.cuh — tcgen05 (5th-gen tensor core) version of the grouped GEMM of sac_gemm.cuh, with
// fp32-class accuracy through the 3xTF32 split:
//     a = a_hi + a_lo,  a_hi = tf32(a) (top 19 bits), a_lo = a - a_hi (exact in fp32)
//     a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi          (three kind::tf32 MMAs, fp32 accumulation in TMEM)
// which keeps the SAC1 step inside the 1e-5 parity bar (plain TF32 would not: 10-bit mantissa).
//
// One CTA per 128 x 128 output tile (M = 128 TMEM lanes, N = 128 fp32 TMEM columns):
//   all 256 threads   fetch the fp32 operands (same virtual-concat / transpose addressing as the FFMA
//                     kernel), split them and write four K-major SWIZZLE_128B tiles (A_hi, A_lo, B_hi,
//                     B_lo; 128 rows x 32 tf32 = 128 B per row) into one of two 64 KB stages; the next
//                     k-block's global loads are in flight while the current one is split and stored
//   thread 0          issues 12 tcgen05.mma (4 k-steps of 8 x 3 products) per k-block and commits them
//                     to the stage's mbarrier, which is what frees the stage for refilling
//   epilogue          8 warps read their TMEM lane quarter with tcgen05.ld (32 lanes x 32 columns per
//                     load), apply relu / relu-mask and write full 128-byte row segments.
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once
#include "sac_gemm.cuh"

namespace ddrl {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;            // BK tf32 = 128 bytes = one swizzle row
constexpr int TILE_BYTES = 128 * 128;                 // one operand tile (128 rows x 128 B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 64 /*barriers, tmem ptr*/;
constexpr int TMEM_COLS = 128;
__device__ int g_tc_debug = 0;   // experiment switch: 1 skip operand fetch, 2 skip split+store, 4 skip MMA, 8 skip fence.proxy

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B operand tile: 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_byte_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_byte_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TC_DONE;\n\t"
      "bra TC_WAIT;\n\t"
      "TC_DONE:\n\t}" ::"r"(s_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// operand element accessors in (row-of-operand, k) coordinates
__device__ __forceinline__ float op_a(const GemmProb& P, int m, int k, int kend) {
  const bool ok = m < P.M && k < kend;
  return P.a_trans ? fetch_a(P, k, m, ok) : fetch_a(P, m, k, ok);
}
__device__ __forceinline__ float op_b(const GemmProb& P, int n, int k, int kend) {
  const bool ok = n < P.N && k < kend;
  return ld_pred(P.b_trans ? P.B + (size_t)n * P.ldb + k : P.B + (size_t)k * P.ldb + n, ok);
}

struct Frag {          // raw fp32 operands of one k-block held by one thread: 4 (row, 4-k chunk) items per operand
  float a[4][4];
  float b[4][4];
};

// item j of thread tid: r_lo = id % 8, kc = (id / 8) % 8, r_hi = id / 64  (id = tid + 256 j): within a quarter warp the
// eight rows of one swizzle group with one k-chunk -> conflict-free 16-byte shared stores after the XOR swizzle
__device__ __forceinline__ void item_coords(int tid, int j, int& r, int& kc) {
  const int id = tid + 256 * j;
  r = (id >> 6) * 8 + (id & 7);
  kc = (id >> 3) & 7;
}

__device__ __forceinline__ void fetch_frag(const GemmProb& P, int tid, int m0, int n0, int k0, int kend, Frag& f) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int r, kc;
    item_coords(tid, j, r, kc);
    const int k = k0 + 4 * kc;
    bool done = false;
    if (!P.a_trans && k + 3 < kend && m0 + r < P.M && k + 3 < P.a0.w && (P.a0.ld & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(P.a0.p) & 15) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(P.a0.p + (size_t)(m0 + r) * P.a0.ld + k);
      f.a[j][0] = v.x; f.a[j][1] = v.y; f.a[j][2] = v.z; f.a[j][3] = v.w;
      done = true;
    }
    if (!done) {
#pragma unroll
      for (int t = 0; t < 4; ++t) f.a[j][t] = op_a(P, m0 + r, k + t, kend);
    }
    done = false;
    if (P.b_trans && k + 3 < kend && n0 + r < P.N && (P.ldb & 3) == 0 && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(P.B + (size_t)(n0 + r) * P.ldb + k);
      f.b[j][0] = v.x; f.b[j][1] = v.y; f.b[j][2] = v.z; f.b[j][3] = v.w;
      done = true;
    }
    if (!done) {
#pragma unroll
      for (int t = 0; t < 4; ++t) f.b[j][t] = op_b(P, n0 + r, k + t, kend);
    }
  }
}

__device__ __forceinline__ void split_store(unsigned char* tile_hi, unsigned char* tile_lo, int r, int kc, const float* x) {
  float hi[4], lo[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    hi[t] = __uint_as_float(__float_as_uint(x[t]) & 0xFFFFE000u);   // tf32: sign, 8 exponent, 10 mantissa bits
    lo[t] = x[t] - hi[t];                                            // exact; the MMA reads its top 19 bits
  }
  const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((kc ^ (r & 7)) << 4);
  *reinterpret_cast<float4*>(tile_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(tile_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(256, 1) gemm_grouped_tc(const __grid_constant__ GemmGroup grp) {
  extern __shared__ unsigned char smem_dyn[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES);
  int pi = 0;
  while (pi + 1 < grp.nprob && (int)blockIdx.x >= grp.p[pi + 1].tile_begin) ++pi;
  const GemmProb P = grp.p[pi];   // into registers (indexed constant-bank reads in the inner loops are slow)
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) bar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  int t = blockIdx.x - P.tile_begin;
  const int per_split = P.tiles_m * P.tiles_n;
  const int split = t / per_split;
  t -= split * per_split;
  const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
  const int kbeg = split * P.k_per_split;
  const int kend = min(P.K, kbeg + P.k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;
  const uint32_t idesc = make_idesc();

  // Three register sets: the operands of k-blocks kb+1 and kb+2 are in flight while kb is split and
  // stored, so the L2 round trip of the operand fetch is hidden behind two k-blocks of work.
  uint32_t phase_bits = 0;   // bit s = parity the next wait on stage s expects
  const int dbg = g_tc_debug;
  auto step = [&](int kb, const Frag& cur, Frag& pre) {
    const int s = kb & 1;
    if (kb + 2 < nkb && !(dbg & 1)) fetch_frag(P, tid, m0, n0, kbeg + (kb + 2) * BK, kend, pre);
    if (kb >= STAGES && !(dbg & 4)) {   // the MMAs that read this stage two k-blocks ago must have retired
      bar_wait(&bars[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    unsigned char* st = base + s * STAGE_BYTES;
    if (!(dbg & 2)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r, kc;
        item_coords(tid, j, r, kc);
        split_store(st, st + TILE_BYTES, r, kc, cur.a[j]);
        split_store(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, r, kc, cur.b[j]);
      }
    }
    if (!(dbg & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
    __syncthreads();
    if (tid == 0 && !(dbg & 4)) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = s_addr(st), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint32_t ko = ks * 32;   // 8 tf32 = 32 bytes along K inside the 128-byte swizzle row
        mma_tf32(tmem_d, make_sdesc(a_lo + ko), make_sdesc(b_hi + ko), idesc, (kb | ks) ? 1u : 0u);
        mma_tf32(tmem_d, make_sdesc(a_hi + ko), make_sdesc(b_lo + ko), idesc, 1u);
        mma_tf32(tmem_d, make_sdesc(a_hi + ko), make_sdesc(b_hi + ko), idesc, 1u);
      }
      mma_commit(&bars[s]);
    }
  };
  Frag f0, f1, f2;
  fetch_frag(P, tid, m0, n0, kbeg, kend, f0);
  if (nkb > 1) fetch_frag(P, tid, m0, n0, kbeg + BK, kend, f1);
  for (int kb = 0; kb < nkb; kb += 3) {
    step(kb, f0, f2);
    if (kb + 1 < nkb) step(kb + 1, f1, f0);
    if (kb + 2 < nkb) step(kb + 2, f2, f1);
  }
  // the last commit covers every MMA issued before it
  if (!(dbg & 4)) {
    const int s = (nkb - 1) & 1;
    bar_wait(&bars[s], (phase_bits >> s) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*(w/4)..+64
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + 32 * q + lane;
    float* C = P.C + (size_t)split * P.c_split_stride;
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      const int c0 = half * 64 + cb * 32;
      float v[32];
      tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
      if (m < P.M) {
        const int n_base = n0 + c0;
        if (P.epi == EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        } else if (P.epi == EPI_MASK) {
          const float* mk = P.mask + (size_t)m * P.ldmask + n_base;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n_base + i < P.N) v[i] = mk[i] > 0.0f ? v[i] : 0.0f;
        }
        float* dst = C + (size_t)m * P.ldc + n_base;
        if (n_base + 31 < P.N && (P.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n_base + i < P.N) dst[i] = v[i];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace ddrl
