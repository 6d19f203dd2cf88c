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

constexpr int BM = 128, BK = 32;                      // BK tf32 = 128 bytes = one swizzle row
constexpr int TILE_BYTES = 128 * 128;                 // one A plane tile (128 rows x 32 tf32)
// TMEM accumulators per tile: NACC = 4 blocks of BN columns = two sets of (hi . hi | cross terms); the k-steps of a
// k-block alternate between the sets and the epilogue sums the four blocks with round-to-nearest adds.  The tensor core
// adds every product into its fp32 accumulator with truncation, a bias that grows with the number of additions
// (measured: 2e-5 gradient error at K ~ 400 with ONE accumulator against 1e-7 for FFMA); splitting the chain divides
// it.  (The number of accumulators does not change the speed of the main loop, which is bound by shared-memory
// bandwidth: the products re-read the A and B tiles from shared memory while TMA writes the next stage — hence the
// [B_hi | B_lo] descriptor of the MMA issuer below, which reads A twice per k-step instead of three times.)
constexpr int NACC = 4;
constexpr int MAX_PROBS = 10;
// Tile width BN is a template parameter of the kernel: 128 x 128 tiles when they fill the chip on their own, 128 x 64
// tiles for the stages of the small-batch configurations (C1 / C2), where 128-wide tiles leave 68 of the 148 SMs idle
// and the per-CTA main loop (12 MMAs per k-block) is the critical path of a launch.
template <int BN>
struct Cfg {
  static constexpr int B_TILE = BN * 128;                       // one B plane tile (BN x 32 tf32)
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * B_TILE;   // A_hi, A_lo, B_hi, B_lo
  static constexpr int STAGES = BN == 64 ? 4 : 3;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers, tmem ptr*/ + 512 /*bias of the tile*/;
  static constexpr int TMEM_COLS = NACC * BN;
  static constexpr int NCB = BN / 64;                           // 32-column chunks per epilogue warp
};

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
constexpr int TRACE_SLOTS = 32;   // per CTA: 0..7 phase stamps; fused kernel: 8.. MMA issuer per k-block, 16.. / 24.. converter groups
#define TC_STAMP(i) do { if (grp.trace && lane == 0) grp.trace[(size_t)blockIdx.x * TRACE_SLOTS + (i)] = gtime(); } while (0)

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  The tcgen05 / TMA instructions take their operands from uniform registers: issued
// from a divergent `if (lane == 0)` region the compiler wraps every single one in an ELECT / R2UR.BROADCAST /
// BRA.U.ANY uniformisation loop (measured: ~100 cycles per tcgen05.mma, the whole main loop), so the issuing warps
// run their loops convergently and only the instruction itself sits under elect.sync.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  // hi = x rounded to tf32 (sign, 8 exponent, 10 mantissa bits), nearest with ties away from zero — what
  // cvt.rna.tf32.f32 computes, written as two full-rate integer operations (the conversion instruction issues at a
  // quarter of that rate and was the bottleneck of the converter warps: 512 ns per k-block)
  const uint32_t u = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  hi = __uint_as_float(u);
  lo = x - hi;                                               // exact, |lo| <= 2^-11 |x|; the MMA reads its top 19 bits
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
// kind::tf32, fp32 accumulate, M = 128, N = n (multiple of 16, <= 256)
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
// Spin on try_wait (the hardware suspends the thread for a bounded time per attempt).  A barrier that never completes
// is a protocol bug: trap (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = s_addr(bar);
  uint32_t done;
  int spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && ++spins > (1 << 22)) __trap();
  } while (!done);
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
template <int BN>
__global__ void __launch_bounds__(256, 1) gemm_grouped_tc(const __grid_constant__ TcGroup grp) {
  constexpr int STAGES = Cfg<BN>::STAGES, STAGE_BYTES = Cfg<BN>::STAGE_BYTES, B_TILE = Cfg<BN>::B_TILE, NCB = Cfg<BN>::NCB;
  constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
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
    const int plane_bytes = is_b ? B_TILE : TILE_BYTES;
    bar_expect_tx(&full[s], 2 * plane_bytes);
    const uint32_t st = s_addr(base + s * STAGE_BYTES) + (is_b ? 2 * TILE_BYTES : 0);
    const int k0 = kbeg + kb * BK;
#pragma unroll
    for (int hl = 0; hl < 2; ++hl) {
      const uint32_t dst = st + hl * plane_bytes;
      if (!mn) tma_load_3d(dst, tm, k0, r0, hl, &full[s]);          // the map's box holds BM (A) / BN (B) rows
      else {
        const int nbox = is_b ? BN / 32 : 4;
        for (int j = 0; j < nbox; ++j) tma_load_3d(dst + j * 4096, tm, r0 + 32 * j, k0, hl, &full[s]);
      }
    }
  };
  // tid 0 starts the first k-block before the CTA-wide setup barrier (it initialised the barriers itself)
  const bool early0 = true;

  if (warp == 0) {
    if (elect_one()) {
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
    __syncwarp();
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
  if (tid >= 128 && tid - 128 < BN && P.bias) s_bias[tid - 128] = (n0 + tid - 128 < N) ? __ldg(P.bias + n0 + tid - 128) : 0.0f;
  uint32_t mbits[2] = {0u, 0u};
  if (P.epi == EPI_MASK && P.mask_bits) {
    const int mrow = m0 + 32 * (warp & 3) + lane;
    const int w0 = (n0 + (warp >> 2) * 32 * NCB) >> 5;
    if (mrow < M) {
      if (32 * w0 < N) mbits[0] = P.mask_bits[(size_t)mrow * P.ldbits + w0];
      if (NCB > 1 && 32 * (w0 + 1) < N) mbits[1] = P.mask_bits[(size_t)mrow * P.ldbits + w0 + 1];
    }
  }

  if (warp == 0 || warp == 2) {
    // ---- TMA producers: warp 0 loads the A planes, warp 2 the B planes (an MN-major operand is four
    //      32 x 32 boxes per plane, so one thread issuing all 16 copies of a k-block would be the bottleneck) ----
    const bool is_b = warp == 2;
    if (!is_b && P.c_tma && elect_one()) prefetch_tmap(&P.tc);
    for (int kb = early0 ? 1 : 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      if (kb >= STAGES) bar_wait(&empty[s], ((kb / STAGES) - 1) & 1);
      if (elect_one()) load_operand(is_b, kb, s);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer (the whole warp walks the loop; one elected lane issues) ----
    // Two MMAs per k-step instead of three: the B_lo plane tile follows the B_hi tile in the stage at the descriptor's own
    // block stride (K-major: 8-row groups 1024 B apart; MN-major: 32-column blocks 4096 B apart), so ONE descriptor of
    // N = 2 BN addresses [B_hi | B_lo]:
    //     acc[2s]   += A_hi . B_hi      acc[2s+1] += A_hi . B_lo        (one MMA, N = 2 BN, into two adjacent accumulators)
    //     acc[2s+1] += A_lo . B_hi                                      (one MMA, N = BN)
    // with s alternating per k-step.  A is read from shared memory twice per k-step instead of three times (the loop is
    // bound by shared-memory bandwidth), and the cross terms never meet the truncating accumulation of the hi . hi sums.
    const uint32_t idesc = make_idesc(a_mn, b_mn, BN), idesc_w = make_idesc(a_mn, b_mn, 2 * BN);
    static_assert(NACC == 4 && 2 * BN <= 256, "two sets of (hi.hi | cross) accumulators");
    const int a_step = a_mn ? 1024 : 32, b_step = b_mn ? 1024 : 32;   // bytes per 8 tf32 of K
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      bar_wait(&full[s], (kb / STAGES) & 1);
      if (kb == 0) TC_STAMP(2);     // first stage landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = s_addr(base + s * STAGE_BYTES), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES;
      // 8 MMAs per k-block: k-step ks into accumulator set ks & 1 (blocks 2s = hi . hi, 2s + 1 = cross terms)
      const uint64_t da_lo = make_sdesc(a_lo, a_mn), da_hi = make_sdesc(a_hi, a_mn), db_hi = make_sdesc(b_hi, b_mn);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t ao = (uint64_t)((ks * a_step) >> 4), bo = (uint64_t)((ks * b_step) >> 4);   // descriptor address field
          const uint32_t d0 = tmem_d + (uint32_t)(2 * (ks & 1) * BN);
          mma_tf32(d0, da_hi + ao, db_hi + bo, idesc_w, (kb > 0 || ks >= 2) ? 1u : 0u);
          mma_tf32(d0 + (uint32_t)BN, da_lo + ao, db_hi + bo, idesc, 1u);
        }
        mma_commit(&empty[s]);          // frees the stage when these MMAs have read it
        if (kb == nkb - 1) mma_commit(acc_ready);   // covers every MMA issued before it
      }
      __syncwarp();
    }
    TC_STAMP(3);                      // all MMAs issued
  }

  __syncthreads();                    // s_bias is visible; the issuing lanes have left their loops
  bar_wait(acc_ready, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) TC_STAMP(4);          // accumulator complete

  // epilogue: warp w owns TMEM lanes 32*(w%4).. and columns (BN/2)*(w/4)..+BN/2, as NCB 32-column chunks whose TMEM
  // loads are all in flight before the first is consumed
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + 32 * q + lane;
    const int epi = P.epi, ldc = P.ldc, c_tma = P.c_tma;
    float* C = P.C + (size_t)split * P.c_split_stride;
    float* C_lo = P.C_lo;
    const bool has_bias = P.bias != nullptr;
    const int cbase = half * 32 * NCB;                                            // first column of this warp in the tile
    const uint32_t trow = tmem_d + ((uint32_t)(32 * q) << 16);
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) {
      const int c0 = cbase + cb * 32;
      const int n_base = n0 + c0;
      if (n_base >= N) continue;                     // warp-uniform
      // sum of the NACC accumulators in index order (round-to-nearest adds); the next accumulator's TMEM load is in
      // flight while the previous one is added
      float v[32], v2[2][32];
      tmem_ld32_nowait(trow + (uint32_t)c0, v);
      tmem_ld32_nowait(trow + (uint32_t)(BN + c0), v2[0]);
#pragma unroll
      for (int acc = 1; acc < NACC; ++acc) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (acc + 1 < NACC) tmem_ld32_nowait(trow + (uint32_t)((acc + 1) * BN + c0), v2[acc & 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[(acc - 1) & 1][i];
      }
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
        unsigned char* blk = base + (warp * NCB + cb) * 8192;
        if (C_lo) {
          float lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) split_tf32(v[i], v[i], lo[i]);
          stage_row(blk + 4096, lane, lo);
        }
        stage_row(blk, lane, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
          tma_store_3d(&P.tc, s_addr(blk), n_base, m0 + 32 * q, C_lo ? 0 : split);
          if (C_lo) tma_store_3d(&P.tc, s_addr(blk + 4096), n_base, m0 + 32 * q, 1);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (c_tma) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // bulk groups are per thread: a no-op for the others
  }
  if (tid == 0) TC_STAMP(5);          // this warp's epilogue done (stores issued and drained)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) TC_STAMP(6);          // all warps done
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

// =====================================================================================================================
// Fused first + second layer of a forward pass (narrow inputs: K1 = D or D + A <= 32, h1 <= 256, h1 % 32 == 0):
//     H1 = relu([x|a] . W1 + b1)      one k-block, M = 128, N = h1: accumulates in TMEM columns [0, h1)
//     H2 = relu(H1 . W2 + b2)         128 x 64 output tile per CTA, K = h1
// One CTA = (pass, 128-row block, 64-column block of H2).  H1 never leaves the SM as an operand: the layer-1 accumulator
// (TMEM lane = row, column = H1 feature) already HAS the layout of a K-major A operand in tensor memory, so the
// converter warps read it with tcgen05.ld, add the bias, apply relu, split into tf32 hi / lo and write hi back IN
// PLACE and lo into a 4-k-block ring of TMEM columns with tcgen05.st; the second layer then runs
// tcgen05.mma with A FROM TENSOR MEMORY and only W2 (streamed by TMA) from shared memory.  Measured reason: with A in
// shared memory the 3xTF32 products re-read it three times per k-step and the launch was bound by shared-memory
// bandwidth (MMA operand reads + converter stores + TMA writes: ~1 us per k-block); from TMEM the second layer is
// bound by the tensor pipe.  The four CTAs of a row block recompute the (cheap, one k-block) first layer; each of
// them writes a quarter of H1's k-blocks (hi / lo planes by TMA store, relu bit masks) for the passes the backward needs.
//   warp 0 / lane 0    TMA producer: [x|a] and W1 (once), then the W2 k-block ring (4 stages of 16 KB)
//   warp 1 / lane 0    MMA issuer: first layer (<= 4 k-steps x 3 products, N = h1), then per k-block 8 MMAs
//                      (4 k-steps x [N = 128 against W2_hi|W2_lo, N = 64 against W2_hi])
//   warps 2..9         converters (two groups of four warps = the four TMEM lane quadrants; even / odd k-blocks),
//                      then the H2 epilogue (bias, relu, TMA store)
// TMEM columns: [0, 256) layer-1 accumulator -> hi plane; [256, 384) lo ring (4 x 32); [384, 448) layer-2 accumulator of
// the hi . hi products, [448, 512) of the two cross products (hi . lo + lo . hi), summed by the epilogue with a
// round-to-nearest add (the small terms never meet the truncating accumulation of the large ones).
// =====================================================================================================================
constexpr int FZ_THREADS = 320;
constexpr int FZ_BN = 64;
constexpr int FZ_SA = 4, FZ_SB = 4;                        // lo ring (TMEM) / W2 ring (shared memory) depth; FZ_SA is even
                                                           // so that a converter group (even / odd k-blocks) owns its stages
constexpr int FZ_B_STAGE = 2 * FZ_BN * 128;                // hi + lo, 64 x 32 tf32 each
constexpr int FZ_MAX_H1 = 256;
constexpr int FZ_IN_BYTES = 2 * TILE_BYTES + 2 * (FZ_MAX_H1 / 32) * 4096;   // [x|a] planes + W1 planes; later: H1 store staging
constexpr int FZ_SMEM_BYTES = FZ_IN_BYTES + FZ_SB * FZ_B_STAGE + 1024 /*align*/ + 256 /*barriers, tmem ptr*/ +
                              4 * FZ_MAX_H1 + 4 * FZ_BN;
constexpr int FZ_TMEM_COLS = 512;
constexpr int FZ_LO_COL = FZ_MAX_H1, FZ_ACC_COL = FZ_MAX_H1 + FZ_SA * BK;
constexpr int FZ_MAX_PROBS = 4;
struct alignas(64) FusedProb {
  CUtensorMap tx;              // [x|a] planes, K-major, box 32 x 128
  CUtensorMap tw1, tw2;        // W1 [K1, h1] / W2 [h1, h2] planes, MN-major, box 32 x 32
  CUtensorMap th1;             // H1 planes (h1, M, 2), box 32 x 32  (store_h1)
  CUtensorMap tc;              // H2 (h2, M, 1), box 32 x 32
  const float *bias1, *bias2;
  uint32_t* bits;              // nullable: relu'(H1) bit masks [M][ldbits]
  int ldbits, store_h1;
  int M, K1, h1, h2;
  int tiles_m, tiles_n, tile_begin;
};
struct FusedGroup {
  int nprob;
  unsigned long long* trace;   // nullable (tools/tc_trace.py): per CTA TRACE_SLOTS globaltimer stamps
  FusedProb p[FZ_MAX_PROBS];
};
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_addr(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand is read from tensor memory (lane = row, one column per tf32 of K)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {   // caller issues tcgen05.wait::st
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
        "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
        "f"(v[30]), "f"(v[31]) : "memory");
}
__global__ void __launch_bounds__(FZ_THREADS, 1) fwd_fused_tc(const __grid_constant__ FusedGroup grp) {
  extern __shared__ unsigned char smem_dyn[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* in1 = base;                                    // [x|a] planes (32 KB) + W1 planes; then H1 store staging
  unsigned char* bring = base + FZ_IN_BYTES;                    // W2 ring; then H2 staging blocks in the epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(bring + FZ_SB * FZ_B_STAGE);
  uint64_t* fullA = bars;                  // [FZ_SA]  4 converter warps have written hi (in place) + lo ring stage
  uint64_t* emptyA = fullA + FZ_SA;        // [FZ_SA]  the MMAs that read the lo ring stage have completed
  uint64_t* fullB = emptyA + FZ_SA;        // [FZ_SB]
  uint64_t* emptyB = fullB + FZ_SB;        // [FZ_SB]
  uint64_t* bar_in1 = emptyB + FZ_SB;      // layer-1 operands landed
  uint64_t* bar_h1 = bar_in1 + 1;          // layer-1 accumulator complete
  uint64_t* acc_ready = bar_h1 + 1;        // layer-2 accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
  float* s_bias1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);   // [h1]
  float* s_bias2 = s_bias1 + FZ_MAX_H1;                                                      // [64]

  if (tid == 0) TC_STAMP(0);          // CTA start
  int pi = 0;
  while (pi + 1 < grp.nprob && (int)blockIdx.x >= grp.p[pi + 1].tile_begin) ++pi;
  const FusedProb& P = grp.p[pi];
  const int M = P.M, h1 = P.h1, h2 = P.h2;
  const int t = blockIdx.x - P.tile_begin;
  const int nt = t % P.tiles_n;
  const int m0 = (t / P.tiles_n) * BM, n0 = nt * FZ_BN;
  const int nkb = h1 / BK;                                      // k-blocks of the second layer
  const int w1_plane = (h1 / 32) * 4096;                        // bytes of one W1 plane in shared memory (32 k x h1)

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&P.tx); prefetch_tmap(&P.tw1); prefetch_tmap(&P.tw2);
      for (int s = 0; s < FZ_SA; ++s) { bar_init(&fullA[s], 4); bar_init(&emptyA[s], 1); }
      for (int s = 0; s < FZ_SB; ++s) { bar_init(&fullB[s], 1); bar_init(&emptyB[s], 1); }
      bar_init(bar_in1, 1); bar_init(bar_h1, 1); bar_init(acc_ready, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // the layer-1 operands are requested right away, under the rest of the CTA's setup (TMEM allocation, barrier)
      pdl_wait();
      bar_expect_tx(bar_in1, 2 * TILE_BYTES + 2 * w1_plane);
      const uint32_t xa = s_addr(in1), w1s = xa + 2 * TILE_BYTES;
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) tma_load_3d(xa + hl * TILE_BYTES, &P.tx, 0, m0, hl, bar_in1);
      for (int hl = 0; hl < 2; ++hl)
        for (int j = 0; j < h1 / 32; ++j) tma_load_3d(w1s + hl * w1_plane + j * 4096, &P.tw1, 32 * j, 0, hl, bar_in1);
    }
    __syncwarp();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "n"(FZ_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl_wait();
  if (tid == 0) TC_STAMP(1);          // setup done (barriers, TMEM)

  if (warp == 0) {
    // ---- TMA producer (the whole warp walks the loop; one elected lane issues) ----
    if (elect_one()) {
      if (P.store_h1) prefetch_tmap(&P.th1);
      prefetch_tmap(&P.tc);
    }
    __syncwarp();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % FZ_SB;
      if (kb >= FZ_SB) bar_wait(&emptyB[s], ((kb / FZ_SB) - 1) & 1);
      if (elect_one()) {
        bar_expect_tx(&fullB[s], FZ_B_STAGE);
        const uint32_t dst = s_addr(bring + s * FZ_B_STAGE);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int j = 0; j < FZ_BN / 32; ++j)
            tma_load_3d(dst + hl * (FZ_B_STAGE / 2) + j * 4096, &P.tw2, n0 + 32 * j, kb * BK, hl, &fullB[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer (the whole warp walks the loop; one elected lane issues) ----
    bar_wait(bar_in1, 0);
    TC_STAMP(2);                    // layer-1 operands landed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const uint32_t idesc1 = make_idesc(0, 1, h1);
      const uint32_t a_hi = s_addr(in1), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = b_hi + w1_plane;
      const int nks = (P.K1 + 7) / 8;
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t ao = ks * 32u, bo = ks * 1024u;
          mma_tf32(tmem_d, make_sdesc(a_lo + ao, false), make_sdesc(b_hi + bo, true), idesc1, ks ? 1u : 0u);
          mma_tf32(tmem_d, make_sdesc(a_hi + ao, false), make_sdesc(b_lo + bo, true), idesc1, 1u);
          mma_tf32(tmem_d, make_sdesc(a_hi + ao, false), make_sdesc(b_hi + bo, true), idesc1, 1u);
        }
        mma_commit(bar_h1);
      }
      __syncwarp();
    }
    // Two MMAs per k-step instead of three: the W2 stage holds the hi plane's two 32-column blocks followed by the lo
    // plane's (uniform 4096-byte LBO), so ONE descriptor of N = 128 addresses [W2_hi | W2_lo] and
    //     acc[0, 64)   += H1_hi . W2_hi          acc[64, 128) += H1_hi . W2_lo      (one MMA, N = 128)
    //     acc[64, 128) += H1_lo . W2_hi                                             (one MMA, N = 64)
    // The A operand (tensor memory) is read twice per k-step instead of three times — the second layer was paced by
    // those reads plus the converters' tcgen05.ld on the same port (512 ns per k-block for 384 cycles of tensor math).
    const uint32_t idesc2 = make_idesc(0, 1, FZ_BN), idesc2w = make_idesc(0, 1, 2 * FZ_BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int sa = kb % FZ_SA, sb = kb % FZ_SB;
      bar_wait(&fullB[sb], (kb / FZ_SB) & 1);
      bar_wait(&fullA[sa], (kb / FZ_SA) & 1);
      TC_STAMP(8 + kb);             // operands of k-block kb ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = tmem_d + (uint32_t)(kb * BK), a_lo = tmem_d + (uint32_t)(FZ_LO_COL + sa * BK);
      const uint32_t b_hi = s_addr(bring + sb * FZ_B_STAGE);      // the lo plane follows at + FZ_B_STAGE / 2
      const uint32_t acc = tmem_d + (uint32_t)FZ_ACC_COL;
      const uint64_t db_hi = make_sdesc(b_hi, true);
      static_assert(FZ_B_STAGE / 2 == (FZ_BN / 32) * 4096, "the lo plane's blocks continue the hi plane's at the descriptor's LBO");
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t bo = (uint64_t)((ks * 1024) >> 4);
          const uint32_t ac = (uint32_t)(ks * 8);                 // 8 tf32 of K = 8 TMEM columns
          mma_tf32_ts(acc, a_hi + ac, db_hi + bo, idesc2w, (kb || ks) ? 1u : 0u);
          mma_tf32_ts(acc + FZ_BN, a_lo + ac, db_hi + bo, idesc2, 1u);
        }
        mma_commit(&emptyA[sa]);
        mma_commit(&emptyB[sb]);
        if (kb == nkb - 1) mma_commit(acc_ready);
      }
      __syncwarp();
    }
    TC_STAMP(4);                    // all MMAs issued
  } else {
    // ---- converters: layer-1 accumulator -> relu(+ b1) -> hi (in place) / lo (ring) A operand of layer 2 in TMEM ----
    const int cw = warp - 2;                     // 0..7
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int cg = cw >> 2;                      // converter group: even / odd k-blocks; epilogue column chunk
    const int ctid = tid - 64;                   // 0..255
    for (int i = ctid; i < h1; i += 256) s_bias1[i] = __ldg(P.bias1 + i);
    if (ctid < FZ_BN) s_bias2[ctid] = (n0 + ctid < h2) ? __ldg(P.bias2 + n0 + ctid) : 0.0f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int row = 32 * q + lane;               // row of the tile = TMEM lane
    const int m = m0 + row;
    const uint32_t trow = tmem_d + ((uint32_t)(32 * q) << 16);
    unsigned char* stg = in1 + cw * 8192;        // this warp's H1 store staging (hi 4 KB + lo 4 KB): free after layer 1
    bar_wait(bar_h1, 0);
    if (warp == 2) TC_STAMP(3);       // layer-1 accumulator complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kb = cg; kb < nkb; kb += 2) {
      const int s = kb % FZ_SA;
      float v[32], lo[32];
      const bool probe = q == 0 && kb == cg + 2;              // trace: sub-steps of this group's second k-block
      if (probe) TC_STAMP(24 + 4 * cg);
      tmem_ld32_nowait(trow + (uint32_t)(kb * BK), v);
      if (kb >= FZ_SA) {
        bar_wait(&emptyA[s], ((kb / FZ_SA) - 1) & 1);          // the MMAs that read lo ring stage s have completed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (probe) TC_STAMP(25 + 4 * cg);
      const bool mine = (kb % P.tiles_n) == nt;                // this CTA's share of H1 for the backward
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias1 + kb * BK + i);
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
      }
      if (P.bits && mine) {                                    // warp-uniform
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) bits |= (v[i] > 0.0f ? 1u : 0u) << i;
        if (m < M) P.bits[(size_t)m * P.ldbits + kb] = bits;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) split_tf32(fmaxf(v[i], 0.0f), v[i], lo[i]);
      tmem_st32(trow + (uint32_t)(kb * BK), v);
      tmem_st32(trow + (uint32_t)(FZ_LO_COL + s * BK), lo);
      if (probe) TC_STAMP(26 + 4 * cg);
      if (P.store_h1 && mine) {
        // this warp's previous TMA store has finished reading the staging block (bulk groups are per thread: only
        // the lane that issued it waits)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        stage_row(stg, lane, v);
        stage_row(stg + 4096, lane, lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (probe) TC_STAMP(27 + 4 * cg);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (q == 0) TC_STAMP(16 + kb);  // this warp has converted its rows of k-block kb
      if (elect_one()) {
        bar_arrive(&fullA[s]);
        if (P.store_h1 && mine) {
          tma_store_3d(&P.th1, s_addr(stg), kb * BK, m0 + 32 * q, 0);
          tma_store_3d(&P.th1, s_addr(stg + 4096), kb * BK, m0 + 32 * q, 1);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    // ---- H2 epilogue: this warp owns rows 32q.. and columns 32 cg..+32 of the 128 x 64 tile ----
    bar_wait(acc_ready, 0);
    if (warp == 2) TC_STAMP(5);       // layer-2 accumulators complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int c0 = 32 * cg, n_base = n0 + c0;
    if (n_base < h2) {                           // warp-uniform
      float v[32], v2[32];
      tmem_ld32_nowait(trow + (uint32_t)(FZ_ACC_COL + c0), v);                 // hi . hi
      tmem_ld32_nowait(trow + (uint32_t)(FZ_ACC_COL + FZ_BN + c0), v2);        // hi . lo + lo . hi
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf((v[i] + v2[i]) + s_bias2[c0 + i], 0.0f);
      // staging in the W2 ring: every W2 load has been consumed
      unsigned char* blk = bring + cw * 4096;
      stage_row(blk, lane, v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(&P.tc, s_addr(blk), n_base, m0 + 32 * q, 0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) TC_STAMP(6);          // all warps done
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(FZ_TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace ddrl
