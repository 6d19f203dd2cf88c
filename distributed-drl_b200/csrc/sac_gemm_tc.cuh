// sac_gemm_tc.cuh — tcgen05 (5th-gen tensor core) grouped GEMM of the SAC1 learner step with fp32-class
// accuracy through the 3xTF32 split
//     x = x_hi + x_lo,  x_hi = x rounded to tf32 (cvt.rna),  x_lo = x - x_hi (exact in fp32)
//     a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi       (three kind::tf32 MMAs, fp32 accumulation in TMEM)
// which keeps the step inside the 1e-5 parity bar (plain TF32 would not: 10-bit mantissa).
//
// Every operand already lives in HBM/L2 as a PRE-SPLIT pair of planes [2][rows][pitch] (plane 0 = hi,
// plane 1 = lo): producers (the epilogue below, the row-wise kernels, the prologue, the optimiser) write
// both planes, so this kernel never passes an operand through registers.  All three GEMM kinds of the
// step read the tensors in their NATURAL row-major layout — no transposed copies exist anywhere:
//     forward   H  = act . W      A = act [B,K]   K-major      B = W  [K,N]   MN-major
//     dgrad     dX = dZ  . W^T    A = dZ  [B,N]   K-major      B = W  [K,N]   K-major (contraction = N)
//     wgrad     dW = act^T . dZ   A = act [B,K]   MN-major     B = dZ [B,N]   MN-major (contraction = B)
// (the major-ness is a bit of the instruction descriptor plus the shared-memory descriptor's strides).
//
// One CTA per 128 x 128 output tile, warp-specialised, 3-stage TMA -> tcgen05 pipeline:
//   warps 0, 2 / lane 0   TMA producers (A planes, B planes): per k-block (32 tf32 = one 128-byte swizzle row)
//                     load A_hi, A_lo / B_hi, B_lo with cp.async.bulk.tensor into a 64 KB stage; out-of-range
//                     rows / columns / k are zero-filled by the TMA unit (ragged M, N, K need no code)
//   warp 1 / lane 0   MMA issuer: 4 k-steps x 3 products of tcgen05.mma.kind::tf32 per stage into one of two TMEM
//                     accumulators (alternating k-blocks: halves the accumulation truncation bias), then
//                     tcgen05.commit to the stage's "empty" barrier; the last commit signals the epilogue
//   all 8 warps       epilogue: tcgen05.ld (32 lanes x 32 columns) of both accumulators, summed with RN adds, bias
//                     (prefetched to shared memory during the main loop) / relu (+ bit mask) / relu-mask, then the
//                     plain fp32 tile or the hi/lo pair the next GEMM consumes goes out through swizzled shared
//                     memory and TMA stores (direct stores when the output is not 16-byte pitched).
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) and
// cute/atom/mma_traits_sm100.hpp (canonical K-major / MN-major SWIZZLE_128B layouts).
#pragma once
#include <cuda.h>

#include "sac_gemm.cuh"

namespace ddrl {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;            // BK tf32 = 128 bytes = one swizzle row
constexpr int TILE_BYTES = 128 * 128;                 // one operand plane tile (128 x 32 tf32)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 3;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers, tmem ptr*/ + 512 /*bias of the tile*/;
constexpr int NACC = 2;                               // TMEM accumulators: k-block kb adds into accumulator kb % NACC
constexpr int TMEM_COLS = NACC * 128;
constexpr int MAX_PROBS = 10;

struct alignas(64) TcProb {
  CUtensorMap ta, tb;          // 3-D (inner, rows, plane) maps over the pre-split operands
  CUtensorMap tc;              // output map (N, M, plane | split), box 32 x 32, used when c_tma != 0
  float* C;                    // plain output, or the hi plane when C_lo != nullptr
  float* C_lo;
  const float* mask;           // relu-mask source (hi plane) and its lo plane
  const float* mask_lo;
  const float* bias;           // nullable: added per output column before the activation
  uint32_t* relu_bits;         // nullable (EPI_RELU): bit c of word [row * ldbits + n/32] = output (row, n + c) > 0
  const uint32_t* mask_bits;   // EPI_MASK with bits instead of the mask planes (the producer's relu_bits)
  int ldbits;
  // Dependent tiles inside ONE launch (two stages of the step merged into one kernel): a producer problem bumps
  // sig_ctr[sig_per_mtile ? m-tile : 0] when its tile is globally visible; a consumer problem's TMA producer warp for
  // the dependent operand (dep_a / dep_b) waits until dep_ctr[dep_per_mtile ? m-tile : 0] >= dep_need.  The step's
  // prologue kernel zeroes the counters.
  unsigned int* sig_ctr;
  const unsigned int* dep_ctr;
  int* err;
  int sig_per_mtile, dep_per_mtile, dep_need, dep_a, dep_b;
  long long c_split_stride;    // floats between split-K partial outputs
  int ldc, ldmask;
  int M, N, K;
  int epi;
  int c_tma;                   // 1: epilogue stages 32 x 32 blocks in shared memory and stores them with TMA
  int a_mn, b_mn;              // 1: operand is MN-major (its contraction index is the ROW of the stored tensor)
  int splits, k_per_split;
  int tiles_m, tiles_n, tile_begin;
};
struct TcGroup {
  int nprob;
  unsigned long long* trace;   // nullable (tools/tc_trace.py): per CTA 8 globaltimer stamps of the kernel's phases
  TcProb p[MAX_PROBS];
};
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i) do { if (grp.trace && lane == 0) grp.trace[(size_t)blockIdx.x * 8 + (i)] = gtime(); } while (0)

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));       // round to nearest tf32 (sign, 8 exponent, 10 mantissa bits)
  hi = __uint_as_float(u);
  lo = x - hi;                                               // exact, |lo| <= 2^-12 |x|; the MMA reads its top 19 bits
}

// SWIZZLE_128B shared-memory matrix descriptor (version 1).
//   K-major : rows of 128 B (32 tf32 of K), 8-row groups SBO = 1024 B apart, LBO unused
//   MN-major: 32-bit operands have ONE legal MN-major layout, SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:
//             "for mn-major tf32 operands, SW128_32B is the only available smem layout"): 32 consecutive MN
//             elements per 128-byte row with its 32-byte chunks XOR-swizzled by (row & 3)
//             (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), 4 k-rows per 512-byte atom, k-atoms SBO = 512 B
//             apart, the next 32 MN elements LBO = 4096 B further (one TMA box of 32 k x 32 mn)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_byte_addr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_byte_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;
  d |= (uint64_t)((mn_major ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(mn_major ? 1 : 2) << 61;
  return d;
}
// kind::tf32, fp32 accumulate, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// this lane's 32 consecutive floats -> row `lane` of a 32 x 32 SWIZZLE_128B block (conflict-free 16-byte stores)
__device__ __forceinline__ void stage_row(unsigned char* blk, int lane, const float* v) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(blk + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {   // caller issues tcgen05.wait::ld before use
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
        "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
        "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
        "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr));
}
__global__ void __launch_bounds__(256, 1) gemm_grouped_tc(const __grid_constant__ TcGroup grp) {
  extern __shared__ unsigned char smem_dyn[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_ready = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
  float* s_bias = reinterpret_cast<float*>(base + STAGES * STAGE_BYTES + 128);   // [BN]

  if (tid == 0) TC_STAMP(0);          // CTA start
  int pi = 0;
  while (pi + 1 < grp.nprob && (int)blockIdx.x >= grp.p[pi + 1].tile_begin) ++pi;
  const TcProb& P = grp.p[pi];
  const int M = P.M, N = P.N, K = P.K, a_mn = P.a_mn, b_mn = P.b_mn;
  int t = blockIdx.x - P.tile_begin;
  const int per_split = P.tiles_m * P.tiles_n;
  const int split = t / per_split;
  t -= split * per_split;
  const int m0 = (t / P.tiles_n) * BM, n0 = (t % P.tiles_n) * BN;
  const int kbeg = split * P.k_per_split;
  const int kend = min(K, kbeg + P.k_per_split);
  const int nkb = (kend - kbeg + BK - 1) / BK;
  // one k-block of one operand (both planes) into stage `s`
  auto load_operand = [&](bool is_b, int kb, int s) {
    const CUtensorMap* tm = is_b ? &P.tb : &P.ta;
    const int mn = is_b ? b_mn : a_mn, r0 = is_b ? n0 : m0;
    bar_expect_tx(&full[s], 2 * TILE_BYTES);
    const uint32_t st = s_addr(base + s * STAGE_BYTES) + (is_b ? 2 * TILE_BYTES : 0);
    const int k0 = kbeg + kb * BK;
#pragma unroll
    for (int hl = 0; hl < 2; ++hl) {
      const uint32_t dst = st + hl * TILE_BYTES;
      if (!mn) tma_load_3d(dst, tm, k0, r0, hl, &full[s]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * 4096, tm, r0 + 32 * j, k0, hl, &full[s]);
      }
    }
  };
  // tid 0 may start the first k-block before the CTA-wide setup barrier (it initialised the barriers itself)
  // unless the operand is produced by other tiles of this launch (dependency wait happens in the producer loops)
  const bool early0 = !P.dep_ctr;

  if (tid == 0) {
    prefetch_tmap(&P.ta);
    prefetch_tmap(&P.tb);
    for (int s = 0; s < STAGES; ++s) { bar_init(&full[s], 2); bar_init(&empty[s], 1); }   // full: A and B producers
    bar_init(acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (early0) {
      pdl_wait();
      load_operand(false, 0, 0);
      load_operand(true, 0, 0);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();     // the next kernel of the chain may set up (barriers, TMEM, descriptor prefetch) under this one
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();        // everything above touched no global data; operands are complete and visible from here on
  if (tid == 0) TC_STAMP(1);          // setup done (barriers, TMEM)
  // epilogue operands fetched now, while the main loop runs: the tile's bias row into shared memory (idle warps 4..7),
  // this thread's relu-mask words into registers
  if (tid >= 128 && P.bias) s_bias[tid - 128] = (n0 + tid - 128 < N) ? __ldg(P.bias + n0 + tid - 128) : 0.0f;
  uint32_t mbits[2] = {0u, 0u};
  if (P.epi == EPI_MASK && P.mask_bits) {
    const int mrow = m0 + 32 * (warp & 3) + lane;
    const int w0 = (n0 + (warp >> 2) * 64) >> 5;
    if (mrow < M) {
      if (32 * w0 < N) mbits[0] = P.mask_bits[(size_t)mrow * P.ldbits + w0];
      if (32 * (w0 + 1) < N) mbits[1] = P.mask_bits[(size_t)mrow * P.ldbits + w0 + 1];
    }
  }

  if (warp == 0 || warp == 2) {
    if (lane == 0) {
      // ---- TMA producers: warp 0 loads the A planes, warp 2 the B planes (an MN-major operand is four
      //      32 x 32 boxes per plane, so one thread issuing all 16 copies of a k-block would be the bottleneck) ----
      const bool is_b = warp == 2;
      if (!is_b && P.c_tma) prefetch_tmap(&P.tc);
      if (P.dep_ctr && (is_b ? P.dep_b : P.dep_a)) {
        // this operand is written by other tiles of the same launch
        const unsigned int target = (unsigned int)P.dep_need;
        const unsigned int* ctr = P.dep_ctr + (P.dep_per_mtile ? m0 / BM : 0);
        const long long t0 = clock64();
        unsigned int v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
          if (v < target && clock64() - t0 > 4000000000LL) { *P.err = 2; break; }   // ~2 s: never hang the GPU
        } while (v < target);
        asm volatile("fence.proxy.async;" ::: "memory");   // the acquire orders generic accesses; TMA reads are async-proxy
      }
      for (int kb = early0 ? 1 : 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) bar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
        load_operand(is_b, kb, s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      const uint32_t idesc = make_idesc(a_mn, b_mn);
      const uint32_t a_step = a_mn ? 1024u : 32u, b_step = b_mn ? 1024u : 32u;   // 8 tf32 of K
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        bar_wait(&full[s], (kb / STAGES) & 1);
        if (kb == 0) TC_STAMP(2);     // first stage landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = s_addr(base + s * STAGE_BYTES), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES,
                       b_lo = a_hi + 3 * TILE_BYTES;
        // The tensor core adds each product into its fp32 accumulator with truncation, a bias that grows with the
        // number of additions (measured: 2e-5 gradient error at K ~ 400 against 1e-7 for FFMA).  Alternate k-blocks go
        // to NACC separate accumulators that the epilogue sums with round-to-nearest adds: the bias shrinks by NACC.
        const uint32_t acc = tmem_d + (uint32_t)(kb % NACC) * 128u;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t ao = ks * a_step, bo = ks * b_step;
          mma_tf32(acc, make_sdesc(a_lo + ao, a_mn), make_sdesc(b_hi + bo, b_mn), idesc, (kb >= NACC || ks) ? 1u : 0u);
          mma_tf32(acc, make_sdesc(a_hi + ao, a_mn), make_sdesc(b_lo + bo, b_mn), idesc, 1u);
          mma_tf32(acc, make_sdesc(a_hi + ao, a_mn), make_sdesc(b_hi + bo, b_mn), idesc, 1u);
        }
        mma_commit(&empty[s]);          // frees the stage when these MMAs have read it
      }
      mma_commit(acc_ready);            // covers every MMA issued before it
      TC_STAMP(3);                      // all MMAs issued
    }
    __syncwarp();
  }

  __syncthreads();                    // s_bias is visible; the issuing lanes have left their loops
  bar_wait(acc_ready, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) TC_STAMP(4);          // accumulator complete

  // epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*(w/4)..+64, as two 32-column chunks whose TMEM loads
  // are both in flight before the first is consumed
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + 32 * q + lane;
    const int epi = P.epi, ldc = P.ldc, c_tma = P.c_tma;
    float* C = P.C + (size_t)split * P.c_split_stride;
    float* C_lo = P.C_lo;
    const bool has_bias = P.bias != nullptr;
    float vv[2][32], v2[2][32];
    const bool live0 = n0 + half * 64 < N, live1 = n0 + half * 64 + 32 < N;     // warp-uniform
    const bool two = nkb > 1;                                                    // second accumulator in use
    const uint32_t trow = tmem_d + ((uint32_t)(32 * q) << 16);
    if (live0) tmem_ld32_nowait(trow + (uint32_t)(half * 64), vv[0]);
    if (live1) tmem_ld32_nowait(trow + (uint32_t)(half * 64 + 32), vv[1]);
    if (two && live0) tmem_ld32_nowait(trow + 128u + (uint32_t)(half * 64), v2[0]);
    if (two && live1) tmem_ld32_nowait(trow + 128u + (uint32_t)(half * 64 + 32), v2[1]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (two) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (live0) vv[0][i] += v2[0][i];
        if (live1) vv[1][i] += v2[1][i];
      }
    }
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      const int c0 = half * 64 + cb * 32;
      const int n_base = n0 + c0;
      if (n_base >= N) continue;                     // warp-uniform
      float* v = vv[cb];
      if (m < M) {
        const bool full_n = n_base + 31 < N;
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + i);   // zero beyond N
            v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
          }
        }
        if (epi == EPI_RELU) {
          uint32_t bits = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            bits |= (v[i] > 0.0f ? 1u : 0u) << i;
            v[i] = fmaxf(v[i], 0.0f);
          }
          if (P.relu_bits) P.relu_bits[(size_t)m * P.ldbits + (n_base >> 5)] = bits;
        } else if (epi == EPI_MASK && P.mask_bits) {
          const uint32_t bits = mbits[cb];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ((bits >> i) & 1u) ? v[i] : 0.0f;
        } else if (epi == EPI_MASK) {
          const float* mk = P.mask + (size_t)m * P.ldmask + n_base;
          const float* ml = P.mask_lo ? P.mask_lo + (size_t)m * P.ldmask + n_base : nullptr;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n_base + i < N) v[i] = (mk[i] > 0.0f || (ml && ml[i] > 0.0f)) ? v[i] : 0.0f;
        }
        if (!c_tma) {
          float* dst = C + (size_t)m * ldc + n_base;
          const bool vec = full_n && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
          if (C_lo) {
            float* dlo = C_lo + (size_t)m * ldc + n_base;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float hi, lo;
              split_tf32(v[i], hi, lo);
              if (full_n || n_base + i < N) { dst[i] = hi; dlo[i] = lo; }
            }
          } else if (vec) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n_base + i < N) dst[i] = v[i];
          }
        }
      }
      if (c_tma) {
        // 32 x 32 block per plane -> swizzled shared memory -> one TMA store each (full 128-byte lines; rows / columns
        // beyond M / N are clipped by the TMA unit).  The pipeline stages are free: every load has been consumed.
        unsigned char* blk = base + (warp * 2 + cb) * 8192;
        if (C_lo) {
          float lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) split_tf32(v[i], v[i], lo[i]);
          stage_row(blk + 4096, lane, lo);
        }
        stage_row(blk, lane, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&P.tc, s_addr(blk), n_base, m0 + 32 * q, C_lo ? 0 : split);
          if (C_lo) tma_store_3d(&P.tc, s_addr(blk + 4096), n_base, m0 + 32 * q, 1);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (c_tma && lane == 0) {
      if (P.sig_ctr) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // writes complete, not just smem read
      else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (P.sig_ctr) {
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
    }
  }
  if (tid == 0) TC_STAMP(5);          // this warp's epilogue done (stores issued and drained)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) TC_STAMP(6);          // all warps done
  if (tid == 0 && P.sig_ctr)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.sig_ctr + (P.sig_per_mtile ? m0 / BM : 0)) : "memory");
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace ddrl
