// replay.cu — HBM-resident replay ring: batched store, Philox index generation + row gather.
//
// Replaces the numpy ring of the reference's ReplayBuffer (example/dsac.py:14-48,
// algos/sac1/sac1.py:28-63).  The reference keeps five arrays (obs1/obs2/acts/rews/done); here a
// transition is ONE packed, 16-byte aligned row
//     [ obs1 (D) | obs2 (D) | acts (A) | rew | done | 0-pad ]      row_f = round_up(2D+A+2, 4) floats
// so a sampled transition is a single contiguous read (2..7 DRAM sectors instead of >= 5 scattered
// ones) and a batched store is a single streaming write.  The API-visible arrays (the dict
// sample_batch returns) are produced by the gather kernels in the reference's five-array form.
//
// Kernels (all HBM-bandwidth bound; algorithmic bytes per transition = 2 * 4 * (2D+A+2)):
//   rb_gather_narrow<LANES>  rows of <= 32 float4: LANES lanes per row, one 128-bit load per lane,
//                            4 rows in flight per lane group; indices drawn 32 at a time per warp
//   rb_gather_wide           rows of  > 32 float4: one warp per row, 4 x 128-bit loads in flight
//   rb_store_rows<T>         SoA inputs -> packed rows, flat (row, chunk) mapping, 128-bit stores
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"

namespace ddrl {

enum IdxMode : int { IDX_INJECT = 0, IDX_PHILOX = 1, IDX_IDENTITY = 2 };

struct GatherArgs {
  const float4* ring;
  int D, A, row_f4;    // row_f4: row stride in float4 (rows may be padded to a 64/128-byte multiple)
  int used_f4;         // float4 chunks of a row that carry data: ceil((2D+A+2)/4)
  uint64_t size;       // sampling range [0,size)
  int64_t total;       // rows to produce
  const int64_t* idx_in;
  int idx_mode;
  uint64_t seed, counter;
  uint32_t rng_stream;
  float *o1, *o2, *oa, *orw, *od;
  int64_t* oidx;
  // global-uniform mode over a replay sharded across GPUs: index g addresses row g - cum[s] of shard s,
  // whose ring is a peer mapping (cudaIpc) read directly over NVLink.  nshards == 0: local ring only.
  int nshards;
  const float4* rings[8];
  int64_t cum[9];
};

__device__ __forceinline__ const float4* row_ptr(const GatherArgs& a, int64_t g) {
  if (a.nshards == 0) return a.ring + g * a.row_f4;
  int s = 0;
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (i < a.nshards && g >= a.cum[i]) s = i;
  return a.rings[s] + (g - a.cum[s]) * a.row_f4;
}

__device__ __forceinline__ void route_scalar(const GatherArgs& a, int64_t b, int f, float x) {
  const int D = a.D, A = a.A;
  if (f < D) a.o1[b * D + f] = x;
  else if (f < 2 * D) a.o2[b * D + (f - D)] = x;
  else if (f < 2 * D + A) a.oa[b * A + (f - 2 * D)] = x;
  else if (f == 2 * D + A) a.orw[b] = x;
  else if (f == 2 * D + A + 1) a.od[b] = x;
}

// chunk c (float4) of packed row -> the five output arrays
__device__ __forceinline__ void route_chunk(const GatherArgs& a, bool aligned, int64_t b, int c,
                                            const float4& v) {
  const int D4 = a.D >> 2;
  if (aligned && c < D4) {
    st_f4(reinterpret_cast<float4*>(a.o1 + b * a.D) + c, v);
  } else if (aligned && c < 2 * D4) {
    st_f4(reinterpret_cast<float4*>(a.o2 + b * a.D) + (c - D4), v);
  } else {
    const int f = c * 4;
    route_scalar(a, b, f + 0, v.x);
    route_scalar(a, b, f + 1, v.y);
    route_scalar(a, b, f + 2, v.z);
    route_scalar(a, b, f + 3, v.w);
  }
}

__device__ __forceinline__ int64_t draw_index(const GatherArgs& a, int64_t ordinal) {
  if (a.idx_mode == IDX_INJECT) return a.idx_in[ordinal];
  if (a.idx_mode == IDX_PHILOX)
    return philox_index((uint64_t)ordinal, a.seed, a.counter, a.rng_stream, a.size);
  return ordinal;
}

__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
  int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffff), src);
  int hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
  return ((int64_t)hi << 32) | (uint32_t)lo;
}

// ---- rows that fit in LANES float4 (LANES in {2,4,8,16,32}) ---------------------------------
// A lane always handles the same chunk l of a row, so where its four floats go is loop-invariant:
// it is resolved once into (base pointer, per-row stride) pairs; the row loop is then
// shuffle -> 128-bit load -> 128-bit store (or up to four 32-bit stores for the acts/rew/done tail).
struct LaneRoute {
  float* vbase;      // non-null: the whole chunk goes to one 16-byte aligned destination
  int vstride;       // floats between consecutive output rows of that destination
  float* sbase[4];   // otherwise: per-float destinations (null = padding, dropped)
  int sstride[4];
};

__device__ __forceinline__ LaneRoute make_route(const GatherArgs& a, int c) {
  LaneRoute r;
  r.vbase = nullptr; r.vstride = 0;
  const int D = a.D, A = a.A, D4 = a.D >> 2;
  const bool aligned = (D & 3) == 0;
  if (aligned && c < D4) { r.vbase = a.o1 + 4 * c; r.vstride = D; }
  else if (aligned && c < 2 * D4) { r.vbase = a.o2 + 4 * (c - D4); r.vstride = D; }
  else if (aligned && (A & 3) == 0 && c < 2 * D4 + (A >> 2)) { r.vbase = a.oa + 4 * (c - 2 * D4); r.vstride = A; }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = 4 * c + j;
    r.sbase[j] = nullptr; r.sstride[j] = 0;
    if (f < D) { r.sbase[j] = a.o1 + f; r.sstride[j] = D; }
    else if (f < 2 * D) { r.sbase[j] = a.o2 + (f - D); r.sstride[j] = D; }
    else if (f < 2 * D + A) { r.sbase[j] = a.oa + (f - 2 * D); r.sstride[j] = A; }
    else if (f == 2 * D + A) { r.sbase[j] = a.orw; r.sstride[j] = 1; }
    else if (f == 2 * D + A + 1) { r.sbase[j] = a.od; r.sstride[j] = 1; }
  }
  return r;
}

__device__ __forceinline__ void route_store(const LaneRoute& r, int64_t b, const float4& v) {
  if (r.vbase) {
    st_f4(reinterpret_cast<float4*>(r.vbase + b * r.vstride), v);
  } else {
    if (r.sbase[0]) r.sbase[0][b * r.sstride[0]] = v.x;
    if (r.sbase[1]) r.sbase[1][b * r.sstride[1]] = v.y;
    if (r.sbase[2]) r.sbase[2][b * r.sstride[2]] = v.z;
    if (r.sbase[3]) r.sbase[3][b * r.sstride[3]] = v.w;
  }
}

template <int LANES, int U>
__global__ void __launch_bounds__(256, (U >= 8 || LANES <= 4) ? 2 : ((LANES == 8 || LANES == 16) ? 4 : 3)) rb_gather_narrow(const GatherArgs a) {
  constexpr int RPP = 32 / LANES;  // rows per pass of a warp
  static_assert(U <= LANES && LANES % U == 0, "U rows in flight per lane group");
  const int lane = threadIdx.x & 31;
  const int g = lane / LANES, l = lane % LANES;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool lane_live = l < a.used_f4;
  const LaneRoute route = make_route(a, l);

  for (int64_t base = warp * 32; base < a.total; base += nwarps * 32) {
    const int64_t mine = base + lane;
    int64_t myidx = 0;
    if (mine < a.total) {
      myidx = draw_index(a, mine);
      if (a.oidx) a.oidx[mine] = myidx;
    }
#pragma unroll 1
    for (int p0 = 0; p0 < LANES; p0 += U) {
      float4 v[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = (p0 + u) * RPP + g;
        const int64_t src = shfl_i64(myidx, r);
        ok[u] = lane_live && (base + r) < a.total;
        if (ok[u]) v[u] = ld_nc_f4(row_ptr(a, src) + l);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) route_store(route, base + (p0 + u) * RPP + g, v[u]);
    }
  }
}

// ---- narrow rows whose float4 count is far from a power of two (C1: 5 chunks of 16 B -> 3 of 8 lanes idle above) ------
// Same warp-level scheme (32 rows per pass, one index per lane), but the 32 * used_f4 chunks of the pass are dealt to the
// lanes FLAT: chunk e = 32 * it + lane belongs to row e / used_f4, so every lane moves data in every iteration.
__global__ void __launch_bounds__(256, 6) rb_gather_flat(const GatherArgs a, uint32_t magic) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool aligned = (a.D & 3) == 0;
  const int used = a.used_f4;
  for (int64_t base = warp * 32; base < a.total; base += nwarps * 32) {
    const int64_t mine = base + lane;
    int64_t myidx = 0;
    if (mine < a.total) {
      myidx = draw_index(a, mine);
      if (a.oidx) a.oidx[mine] = myidx;
    }
    const int rows_here = (int)min((int64_t)32, a.total - base);
    const int chunks = rows_here * used;
#pragma unroll 1
    for (int e0 = lane; e0 < 32 * used; e0 += 32 * 4) {      // all lanes walk the loop (shuffles); `chunks` masks the tail
      float4 v[4];
      int r[4], c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + 32 * u;
        r[u] = (int)(((uint64_t)e * magic) >> 32);            // e / used (exact: e < 1024, magic = ceil(2^32 / used))
        c[u] = e - r[u] * used;
        const int64_t src = shfl_i64(myidx, r[u] & 31);
        if (e < chunks) v[u] = ld_nc_f4(row_ptr(a, src) + c[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + 32 * u < chunks) route_chunk(a, aligned, base + r[u], c[u], v[u]);
    }
  }
}

// ---- wide rows: one warp per row ------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) rb_gather_wide(const GatherArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool aligned = (a.D & 3) == 0;
  // a warp owns 8 consecutive rows per step (enough to amortise the index draw, small enough to
  // spread a 4096-row batch over the whole chip)
  constexpr int ROWS = 8;
  for (int64_t base = warp * ROWS; base < a.total; base += nwarps * ROWS) {
    const int64_t mine = base + lane;
    int64_t myidx = 0;
    if (lane < ROWS && mine < a.total) {
      myidx = draw_index(a, mine);
      if (a.oidx) a.oidx[mine] = myidx;
    }
    for (int r = 0; r < ROWS; ++r) {
      const int64_t b = base + r;
      const int64_t src = shfl_i64(myidx, r);
      if (b >= a.total) break;  // warp-uniform
      const float4* row = row_ptr(a, src);
      for (int c0 = lane; c0 < a.used_f4; c0 += 32 * 4) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + 32 * k;
          if (c < a.used_f4) v[k] = ld_nc_f4(row + c);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + 32 * k;
          if (c < a.used_f4) route_chunk(a, aligned, b, c, v[k]);
        }
      }
    }
  }
}

// ---- very wide rows (frame observations: 113 KB per row): a row is cut into 16 KB segments and every (row,
// segment) pair is one warp's work, so a 512-row batch still spreads over the whole chip -----------------------
constexpr int XW_SEG = 1024;     // float4 per segment
__global__ void __launch_bounds__(256, 4) rb_gather_xwide(const GatherArgs a, int nseg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool aligned = (a.D & 3) == 0;
  for (int64_t item = warp; item < a.total * nseg; item += nwarps) {
    const int64_t b = item / nseg;
    const int sg = (int)(item - b * nseg);
    int64_t src = 0;
    if (lane == 0) {
      src = draw_index(a, b);
      if (a.oidx && sg == 0) a.oidx[b] = src;
    }
    src = shfl_i64(src, 0);
    const float4* row = row_ptr(a, src);
    const int cend = min(a.used_f4, (sg + 1) * XW_SEG);
    for (int c0 = sg * XW_SEG + lane; c0 < cend; c0 += 32 * 4) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 32 * k;
        if (c < cend) v[k] = ld_nc_f4(row + c);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 32 * k;
        if (c < cend) route_chunk(a, aligned, b, c, v[k]);
      }
    }
  }
}

// ---- bulk-async staged gather (TMA engine, no register staging) -------------------------------
// The packed row is ONE contiguous 16-byte-aligned run, so a sampled row is ONE `cp.async.bulk`
// (global -> shared, completion counted in bytes on an mbarrier).  A CTA keeps STAGES tiles of R rows
// in flight (tens of KB per CTA, ~150+ KB per SM) — the memory-level parallelism a random 224-byte
// gather needs to approach the HBM roofline, which per-thread register loads cannot hold — and
// drains each landed tile to the five output arrays with 128-bit stores.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds_f4(const void* p) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_u32(p)));
  return r;
}

template <int STAGES, bool NARROW>
__global__ void __launch_bounds__(256) rb_gather_bulk(const GatherArgs a, int R, int lanes) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int row_bytes = a.used_f4 * 16;
  const int stage_bytes = (R * row_bytes + 127) / 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);
  const int tid = threadIdx.x;
  const int64_t ntiles = (a.total + R - 1) / R;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int stage) {
    const int64_t row0 = tile * R;
    const int rows = (int)min((int64_t)R, a.total - row0);
    if (tid == 0) mbar_arrive_expect_tx(&bars[stage], (uint32_t)(rows * row_bytes));
    if (tid < rows) {
      const int64_t ord = row0 + tid;
      const int64_t idx = draw_index(a, ord);
      if (a.oidx) a.oidx[ord] = idx;
      bulk_g2s(smem + (size_t)stage * stage_bytes + (size_t)tid * row_bytes, row_ptr(a, idx), (uint32_t)row_bytes,
               &bars[stage]);
    }
  };

  // prologue: fill the pipeline
  int64_t next = blockIdx.x;
  for (int s = 0; s < STAGES; ++s, next += gridDim.x)
    if (next < ntiles) issue(next, s);

  const bool aligned = (a.D & 3) == 0;
  LaneRoute route;
  int l = 0, rsub = 0, rows_per_pass = 1;
  if (NARROW) {
    l = tid % lanes; rsub = tid / lanes; rows_per_pass = 256 / lanes;
    route = make_route(a, l);
  }
  const int q = 256 / a.used_f4, rem = 256 % a.used_f4;

  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * R;
    const int rows = (int)min((int64_t)R, a.total - row0);
    mbar_wait(&bars[stage], parity);
    const unsigned char* sbase = smem + (size_t)stage * stage_bytes;
    if (NARROW) {
      if (l < a.used_f4) {
#pragma unroll 4
        for (int r = rsub; r < rows; r += rows_per_pass)
          route_store(route, row0 + r, lds_f4(sbase + (size_t)r * row_bytes + l * 16));
      }
    } else {
      int r = tid / a.used_f4, c = tid - r * a.used_f4;
      while (r < rows) {
        route_chunk(a, aligned, row0 + r, c, lds_f4(sbase + (size_t)r * row_bytes + c * 16));
        r += q; c += rem;
        if (c >= a.used_f4) { c -= a.used_f4; ++r; }
      }
    }
    __syncthreads();   // every thread is done reading this stage before it is refilled
    if (next < ntiles) issue(next, stage);
    next += gridDim.x;
    if (++stage == STAGES) { stage = 0; parity ^= 1; }
  }
}

// ---- TMA-only gather: no byte of an observation passes through a register -------------------------------------
// One source run (a packed row, or one frame of the frame ring) is cut into parts of <= 8 KB; a (sampled row, part)
// unit is ONE bulk copy global -> shared (mbarrier, byte-counted) followed by ONE OR TWO bulk copies shared -> global
// into the output arrays (the obs1 / obs2 slices of a packed row are contiguous 16-byte aligned pieces of the run; a
// frame of the deduplicated ring lands in slot f of obs1 and slot f-1 of obs2).  A CTA is one warp that keeps STAGES
// units in flight; lane 0 issues the bulk copies, the other lanes carry the (A + 2)-float tail / the per-transition
// scalars.  Tens of CTAs per SM (shared memory is the limit) give ~200 KB in flight per SM without any register
// staging, which is what a 512-row batch of 56 KB rows needs to cover the chip.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

struct TmaGatherArgs {
  const char* src;          // ring base
  int64_t src_stride;       // bytes between ring rows / frames
  int run_bytes;            // bytes of one source run (packed row: used_f4 * 16; frame ring: frame bytes)
  int part_bytes, nparts;   // the run is cut into nparts pieces of part_bytes (the last one shorter), multiples of 16
  int obs_bytes;            // packed row: D * 4;  frame ring: frame bytes
  int stack;                // frame ring: frames per stacked observation
  int A;                    // packed row: action floats (tail = A + 2 floats at byte 2 * obs_bytes)
  int64_t cap, size, base;  // frame ring: capacity, valid frames, ring position of the oldest frame
  int64_t total;            // transitions to produce
  const int64_t* idx_in;
  int idx_mode;
  uint64_t seed, counter;
  uint32_t rng_stream;
  char *o1, *o2;
  float *oa, *orw, *od;
  int64_t* oidx;
  const float *act, *rew, *done;   // frame ring: per-transition scalars [cap]
};

constexpr int TG_STAGES = 4;

template <bool FRAMES>
__global__ void __launch_bounds__(32) rb_gather_tma(const TmaGatherArgs a, int slot_stride) {
  extern __shared__ __align__(128) unsigned char tg_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(tg_smem + (size_t)TG_STAGES * slot_stride);
  const int lane = threadIdx.x;
  const int per_b = FRAMES ? (a.stack + 1) * a.nparts : a.nparts;
  const int64_t nunits = a.total * per_b;
  const int64_t G = gridDim.x;
  if (lane == 0) {
    for (int s = 0; s < TG_STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // unit -> (transition b, frame f, byte offset p0 of the part inside the run, bytes, source ring index)
  struct Unit { int64_t b, src; int f, p0, bytes; };
  auto decode = [&](int64_t u) {
    Unit d;
    d.b = u / per_b;
    int r = (int)(u - d.b * per_b);
    d.f = FRAMES ? r / a.nparts : 0;
    const int part = r - d.f * a.nparts;
    d.p0 = part * a.part_bytes;
    d.bytes = min(a.part_bytes, a.run_bytes - d.p0);
    if (FRAMES) {
      // transition i: obs1 = frames i-S+1 .. i, obs2 = i-S+2 .. i+1; Philox draws an AGE u in [0, size - S) counted from the
      // oldest frame, so no window crosses the write head (ring position = base + S-1 + u)
      if (a.idx_mode == IDX_INJECT) d.src = a.idx_in[d.b];
      else {
        d.src = a.base + (a.stack - 1) + (int64_t)philox_index((uint64_t)d.b, a.seed, a.counter, a.rng_stream, (uint64_t)(a.size - a.stack));
        if (d.src >= a.cap) d.src -= a.cap;
      }
    } else {
      d.src = a.idx_mode == IDX_INJECT ? a.idx_in[d.b]
              : (a.idx_mode == IDX_PHILOX ? (int64_t)philox_index((uint64_t)d.b, a.seed, a.counter, a.rng_stream, (uint64_t)a.size) : d.b);
    }
    return d;
  };
  auto issue = [&](int64_t u, int slot) {   // lane 0 only
    const Unit d = decode(u);
    int64_t row = d.src;
    if (FRAMES) { row = d.src - (a.stack - 1) + d.f; row %= a.cap; if (row < 0) row += a.cap; }
    mbar_arrive_expect_tx(&bar[slot], (uint32_t)d.bytes);
    bulk_g2s(tg_smem + (size_t)slot * slot_stride, a.src + row * a.src_stride + d.p0, (uint32_t)d.bytes, &bar[slot]);
  };

  int64_t u_load = blockIdx.x;
  for (int s = 0; s < TG_STAGES; ++s, u_load += G)
    if (u_load < nunits && lane == 0) issue(u_load, s);
  int k = 0;
  for (int64_t u = blockIdx.x; u < nunits; u += G, ++k) {
    const int slot = k % TG_STAGES;
    mbar_wait(&bar[slot], (uint32_t)((k / TG_STAGES) & 1));
    const Unit d = decode(u);
    unsigned char* sm = tg_smem + (size_t)slot * slot_stride;
    if (FRAMES) {
      if (d.f == 0 && d.p0 == 0) {
        if (lane == 0) { a.oa[d.b] = a.act[d.src]; if (a.oidx) a.oidx[d.b] = d.src; }
        if (lane == 1) a.orw[d.b] = a.rew[d.src];
        if (lane == 2) a.od[d.b] = a.done[d.src];
      }
      if (lane == 0) {
        const int64_t ob = (int64_t)a.stack * a.obs_bytes;
        if (d.f < a.stack) bulk_s2g(a.o1 + d.b * ob + (int64_t)d.f * a.obs_bytes + d.p0, sm, (uint32_t)d.bytes);
        if (d.f >= 1) bulk_s2g(a.o2 + d.b * ob + (int64_t)(d.f - 1) * a.obs_bytes + d.p0, sm, (uint32_t)d.bytes);
        bulk_commit();
      }
    } else {
      const int OB = a.obs_bytes, p1 = d.p0 + d.bytes;
      // tail: acts (A), rew, done — floats at byte 2 * OB of the run
      for (int t = lane; t < a.A + 2; t += 32) {
        const int pos = 2 * OB + 4 * t;
        if (pos >= d.p0 && pos < p1) {
          const float v = *reinterpret_cast<const float*>(sm + (pos - d.p0));
          if (t < a.A) a.oa[d.b * a.A + t] = v;
          else if (t == a.A) a.orw[d.b] = v;
          else a.od[d.b] = v;
        }
      }
      if (lane == 0) {
        if (a.oidx && d.p0 == 0) a.oidx[d.b] = d.src;
        int lo = d.p0, hi = min(p1, OB);
        if (lo < hi) bulk_s2g(a.o1 + d.b * OB + lo, sm + (lo - d.p0), (uint32_t)(hi - lo));
        lo = max(d.p0, OB); hi = min(p1, 2 * OB);
        if (lo < hi) bulk_s2g(a.o2 + d.b * OB + (lo - OB), sm + (lo - d.p0), (uint32_t)(hi - lo));
        bulk_commit();
      }
    }
    __syncwarp();
    // the slot consumed in the PREVIOUS iteration is free once its stores have read shared memory
    if (k >= 1 && u_load < nunits) {
      if (lane == 0) { bulk_wait_read<1>(); issue(u_load, (k - 1) % TG_STAGES); }
      u_load += G;
    }
  }
  if (lane == 0) bulk_wait_read<0>();
}

// ---- frame-deduplicated ring (Atari-shaped replay, BASELINE config C4) ---------------------------------------
// The frame ring holds ONE frame per env step; transition i is obs1 = frames[i-S+1 .. i], obs2 = frames[i-S+2 .. i+1]
// (S = stack depth), i.e. S+1 consecutive frames instead of the 2S a naive obs1/obs2 row stores.  Sampling is
// rb_gather_tma<true> above: each of the S+1 frames is read ONCE and lands in obs1 and / or obs2.
// ---- stores of the frame ring and of the N-step sequence ring -------------------------------------------------
// store_frames: frame i of the input lands in ring slot (ptr0 + i) % cap (one CTA per frame, 128-bit copies, all of a
// frame's loads of a thread in flight before its stores); the transition scalars of the step that produced frame i
// belong to the PREVIOUS slot (transition j = stack ending at frame j -> stack ending at j + 1).
struct FrameStoreArgs {
  float4* ring; const float4* in;
  int frame_f4;
  int64_t cap, ptr0, first, n;      // frames [first, n) of the input are written (first > 0 iff n > cap)
  const float *act, *rew, *done;    // [n]
  float *ract, *rrew, *rdone;       // [cap]
};
constexpr int FS_GROUP = 4;   // frames per CTA pass: 4 x 441 float4 (84 x 84) = 7 x 256 -> every load of a thread in flight at once
__global__ void __launch_bounds__(256) fb_store_frames(const FrameStoreArgs a) {
  const int64_t ngroups = (a.n - a.first + FS_GROUP - 1) / FS_GROUP;
  for (int64_t gidx = blockIdx.x; gidx < ngroups; gidx += gridDim.x) {
    const int64_t i0 = a.first + gidx * FS_GROUP;
    const int nf = (int)min((int64_t)FS_GROUP, a.n - i0);
    const float4* src = a.in + i0 * a.frame_f4;                    // the group's frames are contiguous in the input
    const int total = nf * a.frame_f4;
    for (int e0 = threadIdx.x; e0 < total; e0 += 256 * 8) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) if (e0 + 256 * k < total) v[k] = ld_nc_f4(src + e0 + 256 * k);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = e0 + 256 * k;
        if (e < total) {
          const int f = e / a.frame_f4;                            // ring slots wrap frame by frame
          st_f4(a.ring + ((a.ptr0 + i0 + f) % a.cap) * a.frame_f4 + (e - f * a.frame_f4), v[k]);
        }
      }
    }
    if (threadIdx.x < nf) {
      const int64_t i = i0 + threadIdx.x;
      const int64_t pos = (a.ptr0 + i) % a.cap;
      const int64_t prev = pos == 0 ? a.cap - 1 : pos - 1;
      a.ract[prev] = a.act[i]; a.rrew[prev] = a.rew[i]; a.rdone[prev] = a.done[i];
    }
  }
}

// seg_store: the inverse of rb_gather_segments — nseg dense inputs [n, w_s] -> packed rows [obs | acts | rews | done | 0-pad]
// at ring slots (ptr0 + i) % cap; one warp per row, every lane assembles whole 16-byte chunks of the packed row.
constexpr int SEG_TBL = 4096;   // chunks of a packed row covered by the per-CTA lookup tables
struct SegStoreArgs {
  float4* ring;
  int row_f4, nseg;
  int off[8], w[8];
  const float* in[8];
  int64_t cap, ptr0, first, n;
};
// per 16-byte chunk of the packed row: the segment it lies in when it can arrive as ONE 128-bit load (inside one segment,
// 16-byte aligned on both sides), 254 for pure padding, else 255 (assembled float by float); built once per CTA
__device__ __forceinline__ int seg_store_chunk(const SegStoreArgs& a, int c) {
  const int f = 4 * c;
  int hit = 254;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    if (s >= a.nseg) continue;
    const int lo = a.off[s], hi = a.off[s] + a.w[s];
    if (f >= lo && f + 3 < hi && ((a.w[s] | (f - lo)) & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.in[s]) & 15) == 0)) return s;
    if (f + 3 >= lo && f < hi) hit = 255;          // overlaps the segment without being a whole aligned chunk of it
  }
  return hit;
}
__global__ void __launch_bounds__(256) seg_store_rows(const SegStoreArgs a) {
  __shared__ unsigned char s_seg[SEG_TBL];
  for (int c = threadIdx.x; c < min(a.row_f4, SEG_TBL); c += blockDim.x) s_seg[c] = (unsigned char)seg_store_chunk(a, c);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = a.first + warp; i < a.n; i += nwarps) {
    float4* dst = a.ring + ((a.ptr0 + i) % a.cap) * a.row_f4;
    for (int c0 = lane; c0 < a.row_f4; c0 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        if (c >= a.row_f4) continue;
        const int sg = c < SEG_TBL ? (int)s_seg[c] : seg_store_chunk(a, c);
        if (sg < 8) {
          v[u] = ld_nc_f4(reinterpret_cast<const float4*>(a.in[sg] + i * a.w[sg] + (4 * c - a.off[sg])));
        } else {
          float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};               // row padding / gaps between segments stay zero
          if (sg == 255) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int f = 4 * c + j;
#pragma unroll
              for (int s = 0; s < 8; ++s)
                if (s < a.nseg && f >= a.off[s] && f < a.off[s] + a.w[s]) x[j] = __ldg(a.in[s] + i * a.w[s] + (f - a.off[s]));
            }
          }
          v[u] = make_float4(x[0], x[1], x[2], x[3]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + 32 * u < a.row_f4) st_f4(dst + c0 + 32 * u, v[u]);
    }
  }
}

// ---- batched store --------------------------------------------------------------------------
template <typename T>
struct StoreArgs {
  float4* ring;
  int D, A, row_f4, used_f4;
  int64_t cap, ptr0;
  int64_t first, n;   // rows [first, n) of the inputs are written (first > 0 iff n > cap)
  const T *obs, *act, *rew, *nxt, *done;
  int vec_ok;         // inputs are float, 16 B aligned and D % 4 == 0
};

template <typename T>
__device__ __forceinline__ float fetch_field(const StoreArgs<T>& s, int64_t i, int f) {
  const int D = s.D, A = s.A;
  if (f < D) return (float)s.obs[i * D + f];
  if (f < 2 * D) return (float)s.nxt[i * D + (f - D)];
  if (f < 2 * D + A) return (float)s.act[i * A + (f - 2 * D)];
  if (f == 2 * D + A) return (float)s.rew[i];
  if (f == 2 * D + A + 1) return (float)s.done[i];
  return 0.0f;
}

// Flat (row, chunk) map over the CTA's block of rows: consecutive threads write consecutive 16-byte
// chunks of the ring (full 128-byte lines, whole padded rows), row/chunk recovered with a 32-bit
// multiply-shift instead of a division, four independent chunks in flight per thread.
template <typename T>
__device__ __forceinline__ float4 load_chunk(const StoreArgs<T>& s, int64_t i, int c) {
  const int D4 = s.D >> 2;
  if constexpr (sizeof(T) == 4) {
    if (s.vec_ok && c < 2 * D4) {
      const float* src = (c < D4) ? (const float*)s.obs + i * s.D + 4 * c : (const float*)s.nxt + i * s.D + 4 * (c - D4);
      return ld_nc_f4(reinterpret_cast<const float4*>(src));
    }
    if (s.vec_ok && (s.A & 3) == 0 && c < 2 * D4 + (s.A >> 2) && (((uintptr_t)s.act) & 15) == 0)
      return ld_nc_f4(reinterpret_cast<const float4*>((const float*)s.act + i * s.A + 4 * (c - 2 * D4)));
  }
  if (c >= s.used_f4) return make_float4(0.f, 0.f, 0.f, 0.f);   // row padding
  float4 v;
  const int f = 4 * c;
  v.x = fetch_field(s, i, f + 0);
  v.y = fetch_field(s, i, f + 1);
  v.z = fetch_field(s, i, f + 2);
  v.w = fetch_field(s, i, f + 3);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) rb_store_rows(const StoreArgs<T> s, int rows_per_block, uint32_t magic) {
  const int64_t nrows = s.n - s.first;
  const uint32_t rf4 = (uint32_t)s.row_f4;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block; r0 < nrows; r0 += (int64_t)gridDim.x * rows_per_block) {
    const uint32_t rows = (uint32_t)min((int64_t)rows_per_block, nrows - r0);
    const uint32_t nchunks = rows * rf4;
    for (uint32_t e0 = threadIdx.x; e0 < nchunks; e0 += 4 * 256) {
      float4 v[4];
      int64_t dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t e = e0 + u * 256;
        dst[u] = -1;
        if (e < nchunks) {
          const uint32_t lr = __umulhi(e, magic);          // e / row_f4 (exact for e < 2^22, see host)
          const uint32_t c = e - lr * rf4;
          const int64_t i = s.first + r0 + lr;
          int64_t pos = s.ptr0 + i;                        // ptr0 < cap and i - first < cap
          while (pos >= s.cap) pos -= s.cap;
          v[u] = load_chunk(s, i, (int)c);
          dst[u] = pos * s.row_f4 + c;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) st_f4(s.ring + dst[u], v[u]);
    }
  }
}

// ---- segmented rows (N-step sequence replay, algos/sac1/sac_ray.py:34-83) --------------------------------
// A row is the concatenation of up to 8 float segments (obs [Ln+1, D] | acts [Ln, A] | rews [Ln] | done [Ln]),
// padded to a multiple of 4 floats.  One warp per sampled row, 128-bit loads, four in flight; a chunk that lies
// inside one segment at a 16-byte aligned offset leaves as one 128-bit store, others float by float.
struct SegArgs {
  const float4* ring;
  int row_f4, nseg;
  int off[8], w[8];      // segment start (floats) inside the row, width (floats)
  float* out[8];         // dense [batch, w[s]] outputs
  uint64_t size;
  int64_t total;
  const int64_t* idx_in;
  int idx_mode;
  uint64_t seed, counter;
  uint32_t rng_stream;
  int64_t* oidx;
};
__device__ __forceinline__ void seg_route(const SegArgs& a, int64_t b, int f, float x) {
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < a.nseg && f >= a.off[s] && f < a.off[s] + a.w[s]) a.out[s][b * a.w[s] + (f - a.off[s])] = x;
}
// which segment a 16-byte chunk belongs to when it can leave as ONE 128-bit store (inside one segment, 16-byte aligned
// on both sides), else 255: looked up per chunk from a table built once per CTA instead of searched per chunk
__device__ __forceinline__ int seg_of_chunk(const SegArgs& a, int c) {
  const int f = 4 * c;
  int hit = 255;
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < a.nseg && f >= a.off[s] && f + 3 < a.off[s] + a.w[s] && ((a.w[s] | (f - a.off[s])) & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(a.out[s]) & 15) == 0))
      hit = s;
  return hit;
}
__global__ void __launch_bounds__(256) rb_gather_segments(const SegArgs a) {
  __shared__ unsigned char s_seg[SEG_TBL];
  for (int c = threadIdx.x; c < min(a.row_f4, SEG_TBL); c += blockDim.x) s_seg[c] = (unsigned char)seg_of_chunk(a, c);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < a.total; b += nwarps) {
    int64_t idx = 0;
    if (lane == 0) {
      idx = a.idx_mode == IDX_INJECT ? a.idx_in[b] : philox_index((uint64_t)b, a.seed, a.counter, a.rng_stream, a.size);
      if (a.oidx) a.oidx[b] = idx;
    }
    idx = shfl_i64(idx, 0);
    const float4* src = a.ring + idx * a.row_f4;
    for (int c0 = lane; c0 < a.row_f4; c0 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c0 + 32 * u < a.row_f4) v[u] = ld_nc_f4(src + c0 + 32 * u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        if (c >= a.row_f4) continue;
        const int sg = c < SEG_TBL ? (int)s_seg[c] : seg_of_chunk(a, c);
        if (sg != 255) {
          st_f4(reinterpret_cast<float4*>(a.out[sg] + b * a.w[sg] + (4 * c - a.off[sg])), v[u]);
        } else {
          const int f = 4 * c;
          seg_route(a, b, f + 0, v[u].x);
          seg_route(a, b, f + 1, v[u].y);
          seg_route(a, b, f + 2, v[u].z);
          seg_route(a, b, f + 3, v[u].w);
        }
      }
    }
  }
}

// Staged variant for float32 inputs with D % 4 == 0 (the common case): per pass a CTA takes R consecutive
// rows.  Their obs / next_obs slices are CONTIGUOUS blocks of the SoA inputs (R*D floats each), read with fully
// coalesced 128-bit loads and scattered into packed-row order in shared memory; one thread per row adds the
// acts | rew | done | pad tail; then the R packed rows — contiguous in the ring — leave as coalesced 128-bit
// stores.  A handful of instructions per 16 bytes instead of the per-chunk field routing of rb_store_rows.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// Two shared-memory stages per CTA: the cp.async (LDGSTS) loads of the next block of rows are in flight while the
// current block is written to the ring, so neither the load latency nor the store issue is exposed.
__global__ void __launch_bounds__(256) rb_store_staged(const StoreArgs<float> s, int R, uint32_t magic_d4, uint32_t magic_rf4) {
  extern __shared__ float4 st_sm[];
  const int64_t nrows = s.n - s.first;
  const uint32_t D4 = (uint32_t)s.D >> 2, rf4 = (uint32_t)s.row_f4;
  const int D = s.D, A = s.A, row_f = s.row_f4 * 4;
  const size_t stage_f4 = (size_t)R * rf4;
  auto issue = [&](int64_t r0, int stage) {
    float4* sm = st_sm + stage * stage_f4;
    float* smf = reinterpret_cast<float*>(sm);
    const uint32_t rows = (uint32_t)min((int64_t)R, nrows - r0);
    const int64_t i0 = s.first + r0;
    const float4* o1 = reinterpret_cast<const float4*>(s.obs + i0 * D);
    const float4* o2 = reinterpret_cast<const float4*>(s.nxt + i0 * D);
    const uint32_t nobs = rows * D4;
    for (uint32_t e = threadIdx.x; e < nobs; e += 256) {
      const uint32_t r = magic_d4 ? __umulhi(e, magic_d4) : e, c = e - r * D4;   // magic 0: D4 == 1
      cp_async16(sm + r * rf4 + c, o1 + e);
      cp_async16(sm + r * rf4 + D4 + c, o2 + e);
    }
    // the acts | rew | done tail of every row is asynchronous too (16-byte copies when the action row allows, 4-byte
    // copies otherwise): nothing in this function waits for a load
    const bool act16 = (A & 3) == 0 && ((reinterpret_cast<uintptr_t>(s.act) & 15) == 0);
    for (uint32_t r = threadIdx.x; r < rows; r += 256) {
      float* row = smf + (size_t)r * row_f + 2 * D;
      const float* act = s.act + (i0 + r) * A;
      if (act16) {
        for (int j = 0; j < A; j += 4) cp_async16(row + j, act + j);
      } else {
        for (int j = 0; j < A; ++j) cp_async4(row + j, act + j);
      }
      cp_async4(row + A, s.rew + i0 + r);
      cp_async4(row + A + 1, s.done + i0 + r);
      for (int j = 2 * D + A + 2; j < row_f; ++j) smf[(size_t)r * row_f + j] = 0.0f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  constexpr int NST = 2;        // stages: the next block of rows is in flight while one is written out (measured: 2 x 16 KB
                                // beats 4 x 8 KB by 1.7x — the two CTA barriers per block want large blocks)
  const int64_t stride = (int64_t)gridDim.x * R;
  int64_t r0 = (int64_t)blockIdx.x * R;
  int stage = 0;
#pragma unroll
  for (int p = 0; p < NST - 1; ++p) {
    if (r0 + p * stride < nrows) issue(r0 + p * stride, p);
    else asm volatile("cp.async.commit_group;" ::: "memory");     // keep the group count uniform
  }
  for (; r0 < nrows; r0 += stride, stage = (stage + 1) % NST) {
    const int64_t nxt = r0 + (NST - 1) * stride;
    if (nxt < nrows) issue(nxt, (stage + NST - 1) % NST);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
    __syncthreads();
    const float4* sm = st_sm + stage * stage_f4;
    const uint32_t rows = (uint32_t)min((int64_t)R, nrows - r0);
    int64_t pos0 = s.ptr0 + s.first + r0;
    while (pos0 >= s.cap) pos0 -= s.cap;
    const uint32_t nch = rows * rf4;
    for (uint32_t e = threadIdx.x; e < nch; e += 256) {
      const uint32_t r = magic_rf4 ? __umulhi(e, magic_rf4) : e;
      int64_t pos = pos0 + r;
      if (pos >= s.cap) pos -= s.cap;
      st_f4(s.ring + pos * rf4 + (e - r * rf4), sm[e]);
    }
    __syncthreads();     // the stage is refilled by the next iteration's issue
  }
}

// =============================================================================================
// host side
// =============================================================================================
struct Staging {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    size_t want = need + need / 2;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(DDRL_ENOMEM, "cudaMalloc(%zu) for staging failed: %s", want,
                                      cudaGetErrorString(e));
    bytes = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace ddrl

// Cross-stream ordering of ring accesses.  Producers may store from their own threads / streams while the learner samples
// on another (BASELINE config 5; the reference's fire-and-forget `replay_buffer.store.remote`, algos/sac1/sac1.py:195).
// The handle serialises the RESERVATION (ptr / size advance under a mutex, so the ring always equals some serial order
// of the calls, as a Ray actor gives the reference) and orders the KERNELS with events: a sample waits for every earlier
// store issued on another stream, a store waits for every earlier sample and store issued on another stream.  While
// every call arrives on one stream (the common case) no event is recorded at all.
struct StreamDep { cudaStream_t s = nullptr; cudaEvent_t ev = nullptr; bool valid = false; };
struct StagingSet { ddrl::Staging in, out, idx; };

struct ddrl_rb {
  int device = 0, D = 0, A = 0, row_f = 0, row_f4 = 0, used_f4 = 0, sms = 148, gather_u = 8;
  int zero_copy = 1;              // DDRL_ZERO_COPY=0: always stage host results in device memory + cudaMemcpyAsync
  int gather_flat = 1;            // DDRL_GATHER_FLAT=0: keep the power-of-two lane groups for every narrow row
  int gather_mode = 0;            // 0 auto, 1 always bulk-async + register drain, 2 register kernels, 3 TMA-only for wide rows
  int64_t bulk_min_bytes = 4 << 20;
  int64_t cap = 0, ptr = 0, size = 0, steps = 0, sample_times = 0;
  float* ring = nullptr;
  std::recursive_mutex mu;
  bool have_home = false, multi = false;
  cudaStream_t home = nullptr;
  static constexpr int NDEP = 8;
  StreamDep writers[NDEP], readers[NDEP];
  std::map<cudaStream_t, StagingSet> staging;     // device staging of the host-array entry points, per calling stream
  // by-value capture of host stores (ddrl_rb_store_batch_host_copy): two pinned blocks, alternated; a block is reused only
  // after the H2D copy that read it has completed (its event)
  struct PinnedStage { void* p = nullptr; size_t bytes = 0; cudaEvent_t ev = nullptr; bool busy = false; } pinned[2];
  int pinned_cur = 0;
  int nshards = 0, my_shard = 0;
  const float4* peer[8] = {};
  bool peer_opened[8] = {};
};

namespace ddrl {
using RbLock = std::unique_lock<std::recursive_mutex>;

static int dep_record(StreamDep* tab, cudaStream_t s) {
  StreamDep* slot = nullptr;
  for (int i = 0; i < ddrl_rb::NDEP && !slot; ++i) if (tab[i].ev && tab[i].s == s) slot = &tab[i];
  for (int i = 0; i < ddrl_rb::NDEP && !slot; ++i) if (!tab[i].valid) slot = &tab[i];
  if (!slot) {                       // more concurrent streams than slots: retire the first one on the host
    slot = &tab[0];
    DDRL_CUDA(cudaEventSynchronize(slot->ev));
  }
  if (!slot->ev) DDRL_CUDA(cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming));
  slot->s = s; slot->valid = true;
  DDRL_CUDA(cudaEventRecord(slot->ev, s));
  return 0;
}
// before launching a ring access on stream s (caller holds rb->mu)
static int order_before(ddrl_rb* rb, cudaStream_t s, bool is_write) {
  if (!rb->have_home) { rb->have_home = true; rb->home = s; return 0; }
  if (!rb->multi) {
    if (s == rb->home) return 0;
    rb->multi = true;                // first call from a second stream: everything issued so far on the home stream
    int rc = dep_record(rb->writers, rb->home);
    if (rc) return rc;
  }
  for (auto& w : rb->writers)
    if (w.valid) {
      if (w.s != s) DDRL_CUDA(cudaStreamWaitEvent(s, w.ev, 0));
      if (is_write) w.valid = false;     // superseded: later accesses order against THIS store
    }
  if (is_write)
    for (auto& r : rb->readers)
      if (r.valid) {
        if (r.s != s) DDRL_CUDA(cudaStreamWaitEvent(s, r.ev, 0));
        r.valid = false;
      }
  return 0;
}
static int order_after(ddrl_rb* rb, cudaStream_t s, bool is_write) {
  if (!rb->multi) return 0;
  return dep_record(is_write ? rb->writers : rb->readers, s);
}
}  // namespace ddrl

using namespace ddrl;

static int launch_gather_bulk(ddrl_rb* rb, const GatherArgs& a, cudaStream_t st) {
  constexpr int STAGES = 4;
  const int row_bytes = rb->used_f4 * 16;
  // ~16 KB of rows per stage, at most one row per thread, at least 4 rows
  int R = 16384 / row_bytes;
  if (R > 256) R = 256;
  if (R < 4) R = 4;
  const int stage_bytes = (R * row_bytes + 127) / 128 * 128;
  const size_t smem = (size_t)STAGES * stage_bytes + STAGES * sizeof(uint64_t);
  const bool narrow = rb->used_f4 <= 32;
  int lanes = 2;
  while (lanes < rb->used_f4) lanes <<= 1;
  {
    // the attribute is per device / context and the call is cheap: set it on every launch (a process may hold buffers
    // on several GPUs, from several threads)
    cudaError_t e = narrow ? cudaFuncSetAttribute(rb_gather_bulk<STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                           : cudaFuncSetAttribute(rb_gather_bulk<STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(DDRL_ECUDA, "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
  }
  const int64_t ntiles = (a.total + R - 1) / R;
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = std::min<int64_t>(ntiles, (int64_t)rb->sms * per_sm);
  if (narrow) rb_gather_bulk<STAGES, true><<<(int)blocks, 256, smem, st>>>(a, R, lanes);
  else rb_gather_bulk<STAGES, false><<<(int)blocks, 256, smem, st>>>(a, R, lanes);
  DDRL_LAUNCH_CHECK();
  return 0;
}

// parts of <= 8 KB (multiples of 16 bytes), slot stride a multiple of 128 bytes, grid = as many one-warp CTAs as the
// shared memory of the chip holds (each with TG_STAGES slots), capped by the number of units
static void tma_geometry(int run_bytes, int* part_bytes, int* nparts, int* slot_stride, size_t* smem) {
  const int np = (run_bytes + 8191) / 8192;
  const int pb = ((run_bytes + np - 1) / np + 15) / 16 * 16;
  *nparts = (run_bytes + pb - 1) / pb; *part_bytes = pb;
  *slot_stride = (pb + 127) / 128 * 128;
  *smem = (size_t)TG_STAGES * *slot_stride + TG_STAGES * sizeof(uint64_t);
}
static int tma_grid(int sms, size_t smem, int64_t units) {
  int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
  if (per_sm > 24) per_sm = 24;
  if (per_sm < 1) per_sm = 1;
  return (int)std::min<int64_t>(units, (int64_t)sms * per_sm);
}
static bool tma_gather_ok(const ddrl_rb* rb, const GatherArgs& a) {
  return a.nshards == 0 && (rb->D & 3) == 0 && ((((uintptr_t)a.o1) | ((uintptr_t)a.o2)) & 15) == 0;
}
static int launch_gather_tma(ddrl_rb* rb, const GatherArgs& a, cudaStream_t st) {
  TmaGatherArgs t{};
  t.src = reinterpret_cast<const char*>(a.ring);
  t.src_stride = (int64_t)rb->row_f4 * 16;
  t.run_bytes = rb->used_f4 * 16;
  t.obs_bytes = rb->D * 4; t.stack = 1; t.A = rb->A;
  t.cap = rb->cap; t.size = (int64_t)a.size; t.base = 0; t.total = a.total;
  t.idx_in = a.idx_in; t.idx_mode = a.idx_mode; t.seed = a.seed; t.counter = a.counter; t.rng_stream = a.rng_stream;
  t.o1 = reinterpret_cast<char*>(a.o1); t.o2 = reinterpret_cast<char*>(a.o2);
  t.oa = a.oa; t.orw = a.orw; t.od = a.od; t.oidx = a.oidx;
  int slot = 0; size_t smem = 0;
  tma_geometry(t.run_bytes, &t.part_bytes, &t.nparts, &slot, &smem);
  DDRL_CUDA(cudaFuncSetAttribute(rb_gather_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  rb_gather_tma<false><<<tma_grid(rb->sms, smem, a.total * t.nparts), 32, smem, st>>>(t, slot);
  DDRL_LAUNCH_CHECK();
  return 0;
}

static int launch_gather(ddrl_rb* rb, const GatherArgs& a, cudaStream_t st) {
  if (a.total <= 0) return 0;
  // rows wider than 512 B (C3: 3088 B, frame rows: 56 KB): TMA only — one bulk copy in, two out, no register pass.
  // Measured on B200 (profiles/r02_replay_curves.json): C3 one batch of 4096 rows 8.7 us against 22.2 us for the
  // one-warp-per-row register kernel, 0.79 against 0.75 of the HBM peak at 32 k rows, 0.83 against 0.84 at 262 k rows.
  // DDRL_GATHER_MODE=2 keeps the register kernels (also used for D % 4 != 0, unaligned outputs and peer rings).
  if (tma_gather_ok(rb, a) && rb->used_f4 > 32 && (rb->gather_mode == 0 || rb->gather_mode == 3))
    return launch_gather_tma(rb, a, st);
  // bulk-async pipeline for launches big enough to fill the chip when a row is at least 128 B (measured on B200, C2 rows:
  // 5.5 TB/s bulk vs 4.5 TB/s registers); shorter rows (C1: 80 B) are bound by the rate of bulk-copy operations — one
  // 80-byte copy per row — and stay on the register kernels (0.48 vs 0.41 of the HBM peak at 524 k rows per launch)
  // peer (NVLink) rows go through the register kernels: plain ld.global on the mapped peer pointer
  if (a.nshards == 0 &&
      (rb->gather_mode == 1 ||
       (rb->gather_mode == 0 && rb->used_f4 <= 32 && rb->used_f4 >= 8 && a.total * rb->used_f4 * 16 >= (int64_t)rb->bulk_min_bytes)))
    return launch_gather_bulk(rb, a, st);
  const int threads = 256;
  const int max_blocks = rb->sms * 8;
  if (rb->used_f4 <= 32) {
    int lanes = 2;
    while (lanes < rb->used_f4) lanes <<= 1;
    const int64_t rows_per_block = (threads / 32) * 32;
    int64_t blocks = (a.total + rows_per_block - 1) / rows_per_block;
    if (blocks > max_blocks) blocks = max_blocks;
    const bool u8 = rb->gather_u >= 8;
    // a quarter or more of the lanes of the power-of-two kernel would idle (5, 6 of 8; 9..12 of 16; 17..24 of 32): flat map
    if (rb->used_f4 * 4 <= lanes * 3 && rb->gather_flat) {
      const uint32_t magic = (uint32_t)((0x100000000ull + rb->used_f4 - 1) / rb->used_f4);
      rb_gather_flat<<<(int)blocks, threads, 0, st>>>(a, magic);
      DDRL_LAUNCH_CHECK();
      return 0;
    }
    switch (lanes) {
      case 2: rb_gather_narrow<2, 2><<<(int)blocks, threads, 0, st>>>(a); break;
      case 4: rb_gather_narrow<4, 4><<<(int)blocks, threads, 0, st>>>(a); break;
      case 8: rb_gather_narrow<8, 4><<<(int)blocks, threads, 0, st>>>(a); break;
      case 16:
        if (u8) rb_gather_narrow<16, 8><<<(int)blocks, threads, 0, st>>>(a);
        else rb_gather_narrow<16, 4><<<(int)blocks, threads, 0, st>>>(a);
        break;
      default:
        if (u8) rb_gather_narrow<32, 8><<<(int)blocks, threads, 0, st>>>(a);
        else rb_gather_narrow<32, 4><<<(int)blocks, threads, 0, st>>>(a);
        break;
    }
  } else if (rb->used_f4 >= 2 * XW_SEG) {
    const int nseg = (rb->used_f4 + XW_SEG - 1) / XW_SEG;
    int64_t blocks = (a.total * nseg + threads / 32 - 1) / (threads / 32);
    if (blocks > max_blocks) blocks = max_blocks;
    rb_gather_xwide<<<(int)blocks, threads, 0, st>>>(a, nseg);
  } else {
    const int64_t rows_per_block = (threads / 32) * 8;
    int64_t blocks = (a.total + rows_per_block - 1) / rows_per_block;
    if (blocks > max_blocks) blocks = max_blocks;
    rb_gather_wide<<<(int)blocks, threads, 0, st>>>(a);
  }
  DDRL_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_store(ddrl_rb* rb, const void* obs, const void* act, const void* rew,
                        const void* nxt, const void* done, int64_t ptr0, int64_t n,
                        cudaStream_t st) {
  StoreArgs<T> s;
  s.ring = reinterpret_cast<float4*>(rb->ring);
  s.D = rb->D; s.A = rb->A; s.row_f4 = rb->row_f4; s.used_f4 = rb->used_f4;
  s.cap = rb->cap; s.ptr0 = ptr0;
  s.first = n > rb->cap ? n - rb->cap : 0;
  s.n = n;
  s.obs = (const T*)obs; s.act = (const T*)act; s.rew = (const T*)rew;
  s.nxt = (const T*)nxt; s.done = (const T*)done;
  s.vec_ok = sizeof(T) == 4 && (rb->D % 4 == 0) && (((uintptr_t)obs | (uintptr_t)nxt) % 16 == 0);
  if constexpr (sizeof(T) == 4) {
    if (s.vec_ok && rb->row_f4 * 16 <= 32768) {
      // staged kernel: R rows (~32 KB of packed rows) per CTA pass; magic multipliers for e / D4 and e / row_f4
      // (exact while e < 2^16 * d, and e < R * row_f4 <= 2048 + row_f4 here)
      int R = (rb->row_f4 * 16 >= 1024 ? 32768 : 16384) / (rb->row_f4 * 16);   // two stages of ~32 KB (wide rows) / ~16 KB per CTA
      if (R > 128) R = 128;
      if (R < 1) R = 1;
      const uint32_t D4 = (uint32_t)rb->D / 4;
      const uint32_t m1 = D4 > 1 ? (uint32_t)((0x100000000ull + D4 - 1) / D4) : 0u;   // ceil(2^32 / 1) does not fit: 0 = "d is 1"
      const uint32_t m2 = rb->row_f4 > 1 ? (uint32_t)((0x100000000ull + rb->row_f4 - 1) / rb->row_f4) : 0u;
      int64_t nb = ((s.n - s.first) + R - 1) / R;
      if (nb < 1) nb = 1;
      if (nb > rb->sms * 6) nb = rb->sms * 6;
      // per device / context, cheap: set on every launch
      DDRL_CUDA(cudaFuncSetAttribute(rb_store_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768 + 1024));
      rb_store_staged<<<(int)nb, 256, (size_t)2 * R * rb->row_f4 * 16, st>>>(s, R, m1, m2);
      DDRL_LAUNCH_CHECK();
      return 0;
    }
  }
  // rows per CTA pass: ~4 chunks per thread, at least one row
  int rows_per_block = (4 * 256 + rb->row_f4 - 1) / rb->row_f4;
  if (rows_per_block < 1) rows_per_block = 1;
  // magic for e / row_f4 by multiply-high: ceil(2^32 / d) is exact while e * d < 2^32 / ... ; e stays
  // below rows_per_block * row_f4 <= 1024 + row_f4, so the error term e * (d - 2^32 mod d) / 2^32 < 1
  const uint32_t magic = (uint32_t)((0x100000000ull + rb->row_f4 - 1) / rb->row_f4);
  const int threads = 256;
  int64_t blocks = ((s.n - s.first) + rows_per_block - 1) / rows_per_block;
  if (blocks < 1) blocks = 1;
  if (blocks > rb->sms * 8) blocks = rb->sms * 8;
  rb_store_rows<T><<<(int)blocks, threads, 0, st>>>(s, rows_per_block, magic);
  DDRL_LAUNCH_CHECK();
  return 0;
}

extern "C" {

int ddrl_rb_create(int device, int obs_dim, int act_dim, int64_t capacity, ddrl_rb_t* out) {
  if (!out) return fail(DDRL_EINVAL, "ddrl_rb_create: out is NULL");
  *out = nullptr;
  if (obs_dim < 1 || act_dim < 1 || capacity < 1)
    return fail(DDRL_EINVAL, "ddrl_rb_create: obs_dim=%d act_dim=%d capacity=%lld must be >= 1",
                obs_dim, act_dim, (long long)capacity);
  int ndev = 0;
  DDRL_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    return fail(DDRL_EINVAL, "ddrl_rb_create: device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_rb_create: cannot select device %d", device);
  ddrl_rb* rb = new ddrl_rb();
  rb->device = device;
  rb->D = obs_dim; rb->A = act_dim;
  rb->used_f4 = ((2 * obs_dim + act_dim + 2) + 3) / 4;
  {
    // Row stride.  DRAM is fetched in 64..128-byte granules: a row that straddles granules drags in
    // bytes of its neighbours (measured: 302 B read per 224 B row at C2).  Rows longer than 128 B are
    // therefore padded to a 128-byte multiple (C2 224 -> 256 B, C3 3088 -> 3200 B); shorter rows stay
    // 16-byte packed (C1: 80 B, L2-resident at 1e6 rows).  DDRL_ROW_ALIGN=<bytes> overrides.
    int align = (rb->used_f4 * 16 > 128) ? 128 : 16;
    if (const char* e = getenv("DDRL_ROW_ALIGN")) { int v = atoi(e); if (v >= 16 && v % 16 == 0) align = v; }
    const int bytes = (rb->used_f4 * 16 + align - 1) / align * align;
    rb->row_f4 = bytes / 16;
    rb->row_f = rb->row_f4 * 4;
    if (const char* e = getenv("DDRL_GATHER_U")) rb->gather_u = atoi(e);
    if (const char* e = getenv("DDRL_GATHER_MODE")) rb->gather_mode = atoi(e);
    if (const char* e = getenv("DDRL_GATHER_FLAT")) rb->gather_flat = atoi(e);
    if (const char* e = getenv("DDRL_ZERO_COPY")) rb->zero_copy = atoi(e);
    if (const char* e = getenv("DDRL_BULK_MIN_BYTES")) rb->bulk_min_bytes = atoll(e);
  }
  rb->cap = capacity;
  rb->sms = sm_count(device);
  const size_t bytes = (size_t)capacity * rb->row_f * sizeof(float);
  cudaError_t e = cudaMalloc(&rb->ring, bytes);
  if (e != cudaSuccess) {
    delete rb;
    return fail(DDRL_ENOMEM, "ddrl_rb_create: cudaMalloc(%zu bytes) failed: %s", bytes,
                cudaGetErrorString(e));
  }
  e = cudaMemset(rb->ring, 0, bytes);
  if (e != cudaSuccess) {
    cudaFree(rb->ring);
    delete rb;
    return fail(DDRL_ECUDA, "ddrl_rb_create: cudaMemset failed: %s", cudaGetErrorString(e));
  }
  *out = rb;
  return 0;
}

int ddrl_rb_destroy(ddrl_rb_t rb) {
  if (!rb) return 0;
  DeviceGuard guard(rb->device);
  cudaDeviceSynchronize();
  for (int s = 0; s < 8; ++s) if (rb->peer_opened[s]) cudaIpcCloseMemHandle(const_cast<float4*>(rb->peer[s]));
  if (rb->ring) cudaFree(rb->ring);
  for (auto& kv : rb->staging) { kv.second.in.release(); kv.second.out.release(); kv.second.idx.release(); }
  for (auto& ps : rb->pinned) { if (ps.p) cudaFreeHost(ps.p); if (ps.ev) cudaEventDestroy(ps.ev); }
  for (auto* tab : {rb->writers, rb->readers})
    for (int i = 0; i < ddrl_rb::NDEP; ++i) if (tab[i].ev) cudaEventDestroy(tab[i].ev);
  delete rb;
  return 0;
}

static int store_common(ddrl_rb_t rb, const void* obs, const void* act, const void* rew,
                        const void* nxt, const void* done, int64_t n, int in_dtype,
                        cudaStream_t st) {
  int rc = order_before(rb, st, true);
  if (rc) return rc;
  if (in_dtype == DDRL_F32) rc = launch_store<float>(rb, obs, act, rew, nxt, done, rb->ptr, n, st);
  else rc = launch_store<double>(rb, obs, act, rew, nxt, done, rb->ptr, n, st);
  if (rc) return rc;
  rb->ptr = (rb->ptr + n) % rb->cap;
  rb->size = (rb->size + n < rb->cap) ? rb->size + n : rb->cap;
  rb->steps += n;
  return order_after(rb, st, true);
}

int ddrl_rb_store_batch(ddrl_rb_t rb, const void* d_obs, const void* d_act, const void* d_rew,
                        const void* d_next_obs, const void* d_done, int64_t n, int in_dtype,
                        void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_store_batch: NULL handle");
  if (n < 0) return fail(DDRL_EINVAL, "ddrl_rb_store_batch: n=%lld < 0", (long long)n);
  if (in_dtype != DDRL_F32 && in_dtype != DDRL_F64)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch: in_dtype must be DDRL_F32 or DDRL_F64");
  if (n == 0) return 0;
  if (!d_obs || !d_act || !d_rew || !d_next_obs || !d_done)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch: NULL input array");
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  return store_common(rb, d_obs, d_act, d_rew, d_next_obs, d_done, n, in_dtype, (cudaStream_t)stream);
}

// device staging of one host store: five 256-byte aligned sub-arrays obs | next_obs | acts | rews | done
struct HostStoreLayout { size_t b_obs, b_act, b_s, total; };
static HostStoreLayout host_store_layout(const ddrl_rb* rb, int64_t n, size_t es) {
  auto up256 = [](size_t x) { return (x + 255) / 256 * 256; };
  HostStoreLayout L;
  L.b_obs = up256((size_t)n * rb->D * es); L.b_act = up256((size_t)n * rb->A * es); L.b_s = up256((size_t)n * es);
  L.total = 2 * L.b_obs + L.b_act + 2 * L.b_s;
  return L;
}

int ddrl_rb_store_batch_host(ddrl_rb_t rb, const void* h_obs, const void* h_act, const void* h_rew,
                             const void* h_next_obs, const void* h_done, int64_t n, int in_dtype,
                             void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host: NULL handle");
  if (n < 0) return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host: n=%lld < 0", (long long)n);
  if (in_dtype != DDRL_F32 && in_dtype != DDRL_F64)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host: in_dtype must be DDRL_F32 or DDRL_F64");
  if (n == 0) return 0;
  if (!h_obs || !h_act || !h_rew || !h_next_obs || !h_done)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host: NULL input array");
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = in_dtype == DDRL_F32 ? 4 : 8;
  const HostStoreLayout L = host_store_layout(rb, n, es);
  // the staging block may still be read by a previous store on this stream: stream order protects
  // re-use, growth (cudaFree) synchronises the device.
  Staging& st_in = rb->staging[st].in;
  int rc = st_in.ensure(L.total);
  if (rc) return rc;
  char* d_obs = (char*)st_in.p;
  char* d_nxt = d_obs + L.b_obs;
  char* d_act = d_nxt + L.b_obs;
  char* d_rew = d_act + L.b_act;
  char* d_done = d_rew + L.b_s;
  DDRL_CUDA(cudaMemcpyAsync(d_obs, h_obs, (size_t)n * rb->D * es, cudaMemcpyHostToDevice, st));
  DDRL_CUDA(cudaMemcpyAsync(d_nxt, h_next_obs, (size_t)n * rb->D * es, cudaMemcpyHostToDevice, st));
  DDRL_CUDA(cudaMemcpyAsync(d_act, h_act, (size_t)n * rb->A * es, cudaMemcpyHostToDevice, st));
  DDRL_CUDA(cudaMemcpyAsync(d_rew, h_rew, (size_t)n * es, cudaMemcpyHostToDevice, st));
  DDRL_CUDA(cudaMemcpyAsync(d_done, h_done, (size_t)n * es, cudaMemcpyHostToDevice, st));
  return store_common(rb, d_obs, d_act, d_rew, d_nxt, d_done, n, in_dtype, st);
}

int ddrl_rb_store_batch_host_copy(ddrl_rb_t rb, const void* h_obs, const void* h_act, const void* h_rew,
                                  const void* h_next_obs, const void* h_done, int64_t n, int in_dtype, void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host_copy: NULL handle");
  if (n < 0) return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host_copy: n=%lld < 0", (long long)n);
  if (in_dtype != DDRL_F32 && in_dtype != DDRL_F64)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host_copy: in_dtype must be DDRL_F32 or DDRL_F64");
  if (n == 0) return 0;
  if (!h_obs || !h_act || !h_rew || !h_next_obs || !h_done)
    return fail(DDRL_EINVAL, "ddrl_rb_store_batch_host_copy: NULL input array");
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  const size_t es = in_dtype == DDRL_F32 ? 4 : 8;
  const HostStoreLayout L = host_store_layout(rb, n, es);
  auto& ps = rb->pinned[rb->pinned_cur];
  rb->pinned_cur ^= 1;
  if (ps.busy) { DDRL_CUDA(cudaEventSynchronize(ps.ev)); ps.busy = false; }
  if (ps.bytes < L.total) {
    if (ps.p) { cudaFreeHost(ps.p); ps.p = nullptr; ps.bytes = 0; }
    const size_t want = L.total + L.total / 2;
    cudaError_t e = cudaHostAlloc(&ps.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(DDRL_ENOMEM, "cudaHostAlloc(%zu) for store staging failed: %s", want, cudaGetErrorString(e));
    ps.bytes = want;
  }
  if (!ps.ev) DDRL_CUDA(cudaEventCreateWithFlags(&ps.ev, cudaEventDisableTiming));
  // the caller's arrays are captured NOW (the reference's `store.remote` pickles its arguments at call time)
  char* b = (char*)ps.p;
  memcpy(b, h_obs, (size_t)n * rb->D * es);
  memcpy(b + L.b_obs, h_next_obs, (size_t)n * rb->D * es);
  memcpy(b + 2 * L.b_obs, h_act, (size_t)n * rb->A * es);
  memcpy(b + 2 * L.b_obs + L.b_act, h_rew, (size_t)n * es);
  memcpy(b + 2 * L.b_obs + L.b_act + L.b_s, h_done, (size_t)n * es);
  int rc = ddrl_rb_store_block_host(rb, b, n, in_dtype, stream);
  if (rc) return rc;
  DDRL_CUDA(cudaEventRecord(ps.ev, (cudaStream_t)stream));
  ps.busy = true;
  return 0;
}

int64_t ddrl_rb_store_block_bytes(ddrl_rb_t rb, int64_t n, int in_dtype) {
  if (!rb || n < 0 || (in_dtype != DDRL_F32 && in_dtype != DDRL_F64)) return -1;
  return (int64_t)host_store_layout(rb, n, in_dtype == DDRL_F32 ? 4 : 8).total;
}

int ddrl_rb_store_block_host(ddrl_rb_t rb, const void* h_block, int64_t n, int in_dtype, void* stream) {
  if (!rb || !h_block) return fail(DDRL_EINVAL, "ddrl_rb_store_block_host: NULL argument");
  if (n < 0) return fail(DDRL_EINVAL, "ddrl_rb_store_block_host: n=%lld < 0", (long long)n);
  if (in_dtype != DDRL_F32 && in_dtype != DDRL_F64)
    return fail(DDRL_EINVAL, "ddrl_rb_store_block_host: in_dtype must be DDRL_F32 or DDRL_F64");
  if (n == 0) return 0;
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = in_dtype == DDRL_F32 ? 4 : 8;
  const HostStoreLayout L = host_store_layout(rb, n, es);
  Staging& st_in = rb->staging[st].in;
  int rc = st_in.ensure(L.total);
  if (rc) return rc;
  char* d_obs = (char*)st_in.p;
  // the caller's block has the layout of the device staging: ONE copy (up to the last used byte)
  DDRL_CUDA(cudaMemcpyAsync(d_obs, h_block, 2 * L.b_obs + L.b_act + L.b_s + (size_t)n * es, cudaMemcpyHostToDevice, st));
  return store_common(rb, d_obs, d_obs + 2 * L.b_obs, d_obs + 2 * L.b_obs + L.b_act, d_obs + L.b_obs,
                      d_obs + 2 * L.b_obs + L.b_act + L.b_s, n, in_dtype, st);
}

static int sample_check(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const char* who) {
  if (!rb) return fail(DDRL_EINVAL, "%s: NULL handle", who);
  if (batch < 0 || n_batches < 0)
    return fail(DDRL_EINVAL, "%s: batch=%lld n_batches=%lld must be >= 0", who, (long long)batch,
                (long long)n_batches);
  if ((double)batch * (double)n_batches >= 4294967296.0)
    return fail(DDRL_EINVAL, "%s: batch*n_batches must be < 2^32", who);
  if (rb->size == 0 && batch * n_batches > 0)
    return fail(DDRL_EEMPTY, "%s: ring is empty (the reference raises ValueError: high <= 0)", who);
  return 0;
}

int ddrl_rb_sample(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* d_idx_in,
                   uint64_t seed, uint64_t counter, uint32_t rng_stream, float* d_out_obs1,
                   float* d_out_obs2, float* d_out_acts, float* d_out_rews, float* d_out_done,
                   int64_t* d_out_idx, void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_sample: NULL handle");
  RbLock lock(rb->mu);
  int rc = sample_check(rb, batch, n_batches, "ddrl_rb_sample");
  if (rc) return rc;
  const int64_t total = batch * n_batches;
  if (total > 0 && (!d_out_obs1 || !d_out_obs2 || !d_out_acts || !d_out_rews || !d_out_done))
    return fail(DDRL_EINVAL, "ddrl_rb_sample: NULL output array");
  DeviceGuard guard(rb->device);
  GatherArgs a;
  a.ring = reinterpret_cast<const float4*>(rb->ring);
  a.D = rb->D; a.A = rb->A; a.row_f4 = rb->row_f4; a.used_f4 = rb->used_f4;
  a.size = (uint64_t)rb->size; a.total = total;
  a.idx_in = d_idx_in; a.idx_mode = d_idx_in ? IDX_INJECT : IDX_PHILOX;
  a.seed = seed; a.counter = counter; a.rng_stream = rng_stream;
  a.o1 = d_out_obs1; a.o2 = d_out_obs2; a.oa = d_out_acts; a.orw = d_out_rews; a.od = d_out_done;
  a.oidx = d_out_idx;
  a.nshards = 0;
  if ((rc = order_before(rb, (cudaStream_t)stream, false))) return rc;
  rc = launch_gather(rb, a, (cudaStream_t)stream);
  if (rc) return rc;
  rb->sample_times += n_batches;
  return order_after(rb, (cudaStream_t)stream, false);
}

int64_t ddrl_rb_sample_block_bytes(ddrl_rb_t rb, int64_t n) {
  if (!rb || n < 0) return -1;
  int64_t f = n * (2 * (int64_t)rb->D + rb->A + 2) * 4;
  f = (f + 7) / 8 * 8;
  return f + n * 8;
}

int ddrl_rb_sample_host(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_idx_in,
                        uint64_t seed, uint64_t counter, uint32_t rng_stream, void* h_out_block,
                        int64_t block_bytes, void* stream) {
  int rc = ddrl_rb_sample_host_async(rb, batch, n_batches, h_idx_in, seed, counter, rng_stream, h_out_block, block_bytes, stream);
  if (rc) return rc;
  DeviceGuard guard(rb->device);
  DDRL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int ddrl_rb_sample_host_async(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_idx_in,
                              uint64_t seed, uint64_t counter, uint32_t rng_stream, void* h_out_block,
                              int64_t block_bytes, void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_sample_host: NULL handle");
  RbLock lock(rb->mu);
  int rc = sample_check(rb, batch, n_batches, "ddrl_rb_sample_host");
  if (rc) return rc;
  const int64_t n = batch * n_batches;
  const int64_t need = ddrl_rb_sample_block_bytes(rb, n);
  if (!h_out_block || block_bytes < need)
    return fail(DDRL_EINVAL, "ddrl_rb_sample_host: output block too small (%lld < %lld)",
                (long long)block_bytes, (long long)need);
  if (n == 0) return 0;
  DeviceGuard guard(rb->device);
  cudaStream_t st = (cudaStream_t)stream;
  StagingSet& sg = rb->staging[st];
  const int64_t* d_idx_in = nullptr;
  if (h_idx_in) {
    rc = sg.idx.ensure((size_t)n * 8);
    if (rc) return rc;
    DDRL_CUDA(cudaMemcpyAsync(sg.idx.p, h_idx_in, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    d_idx_in = (const int64_t*)sg.idx.p;
  }
  // Small batches into PINNED host memory: the gather kernel writes the block itself through the block's device
  // mapping (posted PCIe writes of whole row slices) — no device staging, no separate D2H copy and its ~10 us of DMA
  // launch latency.  Larger results and the TMA kernels (rows > 512 B) keep staging + one cudaMemcpyAsync.
  void* mapped = rb->zero_copy && need <= (2 << 20) && rb->used_f4 <= 32 ? host_device_pointer(h_out_block) : nullptr;
  if (!mapped) {
    rc = sg.out.ensure((size_t)need);
    if (rc) return rc;
  }
  void* stage = mapped ? mapped : sg.out.p;
  float* o1 = (float*)stage;
  float* o2 = o1 + n * rb->D;
  float* oa = o2 + n * rb->D;
  float* orw = oa + n * rb->A;
  float* od = orw + n;
  int64_t* oidx = (int64_t*)((char*)stage + (need - n * 8));
  const int64_t fbytes = n * (2 * (int64_t)rb->D + rb->A + 2) * 4;
  if (need - n * 8 > fbytes) {    // the 4 alignment bytes in front of the index array: defined contents
    if (mapped) memset((char*)h_out_block + fbytes, 0, (size_t)(need - n * 8 - fbytes));
    else DDRL_CUDA(cudaMemsetAsync((char*)stage + fbytes, 0, (size_t)(need - n * 8 - fbytes), st));
  }
  rc = ddrl_rb_sample(rb, batch, n_batches, d_idx_in, seed, counter, rng_stream, o1, o2, oa, orw, od,
                      oidx, stream);
  if (rc) return rc;
  // the block is complete when `stream` reaches this point: the caller synchronises (ddrl_rb_sample_host does; a
  // prefetcher records an event and keeps going).  The device staging of this stream is reused by its next call.
  if (!mapped) DDRL_CUDA(cudaMemcpyAsync(h_out_block, stage, (size_t)need, cudaMemcpyDeviceToHost, st));
  return 0;
}

int ddrl_rb_ipc_export(ddrl_rb_t rb, void* h_handle64) {
  if (!rb || !h_handle64) return fail(DDRL_EINVAL, "ddrl_rb_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  DeviceGuard guard(rb->device);
  cudaIpcMemHandle_t hd;
  DDRL_CUDA(cudaIpcGetMemHandle(&hd, rb->ring));
  memcpy(h_handle64, &hd, 64);
  return 0;
}

int ddrl_rb_peer_attach(ddrl_rb_t rb, int n_shards, int shard, const void* h_handle64) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_peer_attach: NULL handle");
  if (n_shards < 1 || n_shards > 8 || shard < 0 || shard >= n_shards)
    return fail(DDRL_EINVAL, "ddrl_rb_peer_attach: shard %d of %d (max 8 shards)", shard, n_shards);
  DeviceGuard guard(rb->device);
  rb->nshards = n_shards;
  if (!h_handle64) {          // this rank's own shard
    rb->my_shard = shard;
    rb->peer[shard] = reinterpret_cast<const float4*>(rb->ring);
    return 0;
  }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, h_handle64, 64);
  void* p = nullptr;
  DDRL_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  rb->peer[shard] = reinterpret_cast<const float4*>(p);
  rb->peer_opened[shard] = true;
  return 0;
}

int ddrl_rb_sample_global(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_shard_sizes,
                          const int64_t* d_idx_in, uint64_t seed, uint64_t counter, uint32_t rng_stream,
                          float* d_out_obs1, float* d_out_obs2, float* d_out_acts, float* d_out_rews,
                          float* d_out_done, int64_t* d_out_idx, void* stream) {
  if (!rb || !h_shard_sizes) return fail(DDRL_EINVAL, "ddrl_rb_sample_global: NULL argument");
  if (rb->nshards < 1) return fail(DDRL_ESTATE, "ddrl_rb_sample_global: no shards attached (ddrl_rb_peer_attach)");
  const int64_t total = batch * n_batches;
  if (batch < 0 || n_batches < 0 || (double)batch * (double)n_batches >= 4294967296.0)
    return fail(DDRL_EINVAL, "ddrl_rb_sample_global: bad batch / n_batches");
  if (total > 0 && (!d_out_obs1 || !d_out_obs2 || !d_out_acts || !d_out_rews || !d_out_done))
    return fail(DDRL_EINVAL, "ddrl_rb_sample_global: NULL output array");
  GatherArgs a;
  a.ring = reinterpret_cast<const float4*>(rb->ring);
  a.D = rb->D; a.A = rb->A; a.row_f4 = rb->row_f4; a.used_f4 = rb->used_f4;
  a.nshards = rb->nshards;
  a.cum[0] = 0;
  for (int s = 0; s < 8; ++s) {
    a.rings[s] = s < rb->nshards ? rb->peer[s] : nullptr;
    if (s < rb->nshards && !rb->peer[s]) return fail(DDRL_ESTATE, "ddrl_rb_sample_global: shard %d not attached", s);
    a.cum[s + 1] = a.cum[s] + (s < rb->nshards ? h_shard_sizes[s] : 0);
  }
  if (a.cum[rb->nshards] == 0 && total > 0)
    return fail(DDRL_EEMPTY, "ddrl_rb_sample_global: all shards are empty (the reference raises ValueError: high <= 0)");
  a.size = (uint64_t)a.cum[rb->nshards]; a.total = total;
  a.idx_in = d_idx_in; a.idx_mode = d_idx_in ? IDX_INJECT : IDX_PHILOX;
  a.seed = seed; a.counter = counter; a.rng_stream = rng_stream;
  a.o1 = d_out_obs1; a.o2 = d_out_obs2; a.oa = d_out_acts; a.orw = d_out_rews; a.od = d_out_done; a.oidx = d_out_idx;
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  int rc = order_before(rb, (cudaStream_t)stream, false);
  if (rc) return rc;
  rc = launch_gather(rb, a, (cudaStream_t)stream);
  if (rc) return rc;
  rb->sample_times += n_batches;
  return order_after(rb, (cudaStream_t)stream, false);
}

int ddrl_fb_sample_stack(int device, const void* d_frames, int64_t frame_bytes, int stack, int64_t capacity, int64_t size,
                         int64_t oldest, const float* d_act, const float* d_rew, const float* d_done, int64_t batch,
                         const int64_t* d_idx_in, uint64_t seed, uint64_t counter, uint32_t rng_stream,
                         void* d_out_obs1, void* d_out_obs2, float* d_out_acts, float* d_out_rews, float* d_out_done,
                         int64_t* d_out_idx, void* stream) {
  if (!d_frames || !d_act || !d_rew || !d_done || !d_out_obs1 || !d_out_obs2 || !d_out_acts || !d_out_rews || !d_out_done)
    return fail(DDRL_EINVAL, "ddrl_fb_sample_stack: NULL array");
  if (frame_bytes < 16 || frame_bytes % 16 != 0) return fail(DDRL_EINVAL, "ddrl_fb_sample_stack: frame_bytes must be a multiple of 16");
  if (stack < 1 || capacity < stack + 1 || size > capacity || batch < 0 || oldest < 0 || oldest >= capacity)
    return fail(DDRL_EINVAL, "ddrl_fb_sample_stack: bad stack / capacity / size / oldest / batch");
  if (size < stack + 1 && batch > 0)
    return fail(DDRL_EEMPTY, "ddrl_fb_sample_stack: fewer than stack+1 frames stored");
  if (((((uintptr_t)d_frames) | ((uintptr_t)d_out_obs1) | ((uintptr_t)d_out_obs2)) & 15) != 0)
    return fail(DDRL_EINVAL, "ddrl_fb_sample_stack: frame ring and stacked outputs must be 16-byte aligned");
  if (batch == 0) return 0;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_fb_sample_stack: cannot select device %d", device);
  TmaGatherArgs t{};
  t.src = reinterpret_cast<const char*>(d_frames);
  t.src_stride = frame_bytes; t.run_bytes = (int)frame_bytes; t.obs_bytes = (int)frame_bytes; t.stack = stack; t.A = 0;
  t.cap = capacity; t.size = size; t.base = oldest; t.total = batch;
  t.idx_in = d_idx_in; t.idx_mode = d_idx_in ? IDX_INJECT : IDX_PHILOX;
  t.seed = seed; t.counter = counter; t.rng_stream = rng_stream;
  t.o1 = reinterpret_cast<char*>(d_out_obs1); t.o2 = reinterpret_cast<char*>(d_out_obs2);
  t.oa = d_out_acts; t.orw = d_out_rews; t.od = d_out_done; t.oidx = d_out_idx;
  t.act = d_act; t.rew = d_rew; t.done = d_done;
  int slot = 0; size_t smem = 0;
  tma_geometry(t.run_bytes, &t.part_bytes, &t.nparts, &slot, &smem);
  DDRL_CUDA(cudaFuncSetAttribute(rb_gather_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  rb_gather_tma<true><<<tma_grid(sm_count(device), smem, batch * (stack + 1) * t.nparts), 32, smem, (cudaStream_t)stream>>>(t, slot);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_fb_store_frames(int device, void* d_frames, int64_t frame_bytes, int64_t capacity, int64_t ptr, float* d_act,
                         float* d_rew, float* d_done, const void* d_in_frames, const float* d_in_act, const float* d_in_rew,
                         const float* d_in_done, int64_t n, void* stream) {
  if (!d_frames || !d_act || !d_rew || !d_done || !d_in_frames || !d_in_act || !d_in_rew || !d_in_done)
    return fail(DDRL_EINVAL, "ddrl_fb_store_frames: NULL array");
  if (frame_bytes < 16 || frame_bytes % 16 != 0 || ((((uintptr_t)d_frames) | ((uintptr_t)d_in_frames)) & 15) != 0)
    return fail(DDRL_EINVAL, "ddrl_fb_store_frames: frames must be 16-byte aligned multiples of 16 bytes");
  if (capacity < 1 || ptr < 0 || ptr >= capacity || n < 0) return fail(DDRL_EINVAL, "ddrl_fb_store_frames: bad capacity / ptr / n");
  if (n == 0) return 0;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_fb_store_frames: cannot select device %d", device);
  FrameStoreArgs a;
  a.ring = reinterpret_cast<float4*>(d_frames); a.in = reinterpret_cast<const float4*>(d_in_frames);
  a.frame_f4 = (int)(frame_bytes / 16); a.cap = capacity; a.ptr0 = ptr;
  a.first = n > capacity ? n - capacity : 0; a.n = n;
  a.act = d_in_act; a.rew = d_in_rew; a.done = d_in_done; a.ract = d_act; a.rrew = d_rew; a.rdone = d_done;
  const int64_t blocks = std::min<int64_t>((n - a.first + FS_GROUP - 1) / FS_GROUP, (int64_t)sm_count(device) * 8);
  fb_store_frames<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_seg_store(int device, float* d_ring, int row_floats, int64_t capacity, int64_t ptr, int nseg, const int* h_seg_off,
                   const int* h_seg_w, const float* const* h_d_in, int64_t n, void* stream) {
  if (!d_ring || !h_seg_off || !h_seg_w || !h_d_in) return fail(DDRL_EINVAL, "ddrl_seg_store: NULL argument");
  if (row_floats < 4 || row_floats % 4 != 0) return fail(DDRL_EINVAL, "ddrl_seg_store: row_floats must be a positive multiple of 4");
  if (nseg < 1 || nseg > 8) return fail(DDRL_EINVAL, "ddrl_seg_store: 1..8 segments");
  if (capacity < 1 || ptr < 0 || ptr >= capacity || n < 0) return fail(DDRL_EINVAL, "ddrl_seg_store: bad capacity / ptr / n");
  if (n == 0) return 0;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_seg_store: cannot select device %d", device);
  SegStoreArgs a{};
  a.ring = reinterpret_cast<float4*>(d_ring);
  a.row_f4 = row_floats / 4; a.nseg = nseg;
  int prev_end = 0;
  for (int s = 0; s < nseg; ++s) {
    if (h_seg_off[s] < prev_end || h_seg_w[s] < 1 || h_seg_off[s] + h_seg_w[s] > row_floats || !h_d_in[s])
      return fail(DDRL_EINVAL, "ddrl_seg_store: segment %d is out of order, outside the row or has no input", s);
    a.off[s] = h_seg_off[s]; a.w[s] = h_seg_w[s]; a.in[s] = h_d_in[s];
    prev_end = h_seg_off[s] + h_seg_w[s];
  }
  a.cap = capacity; a.ptr0 = ptr; a.first = n > capacity ? n - capacity : 0; a.n = n;
  const int64_t blocks = std::min<int64_t>((n - a.first + 7) / 8, (int64_t)sm_count(device) * 8);
  seg_store_rows<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_seg_sample(int device, const float* d_ring, int row_floats, int64_t size, int nseg, const int* h_seg_off,
                    const int* h_seg_w, float* const* h_d_out, int64_t batch, const int64_t* d_idx_in, uint64_t seed,
                    uint64_t counter, uint32_t rng_stream, int64_t* d_out_idx, void* stream) {
  if (!d_ring || !h_seg_off || !h_seg_w || !h_d_out) return fail(DDRL_EINVAL, "ddrl_seg_sample: NULL argument");
  if (row_floats < 4 || row_floats % 4 != 0) return fail(DDRL_EINVAL, "ddrl_seg_sample: row_floats must be a positive multiple of 4");
  if (nseg < 1 || nseg > 8) return fail(DDRL_EINVAL, "ddrl_seg_sample: 1..8 segments");
  if (batch < 0) return fail(DDRL_EINVAL, "ddrl_seg_sample: batch < 0");
  if (size <= 0 && batch > 0) return fail(DDRL_EEMPTY, "ddrl_seg_sample: ring is empty (the reference raises ValueError: high <= 0)");
  if (batch == 0) return 0;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(DDRL_ECUDA, "ddrl_seg_sample: cannot select device %d", device);
  SegArgs a{};
  a.ring = reinterpret_cast<const float4*>(d_ring);
  a.row_f4 = row_floats / 4; a.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    if (h_seg_off[s] < 0 || h_seg_w[s] < 1 || h_seg_off[s] + h_seg_w[s] > row_floats || !h_d_out[s])
      return fail(DDRL_EINVAL, "ddrl_seg_sample: segment %d is outside the row or has no output", s);
    a.off[s] = h_seg_off[s]; a.w[s] = h_seg_w[s]; a.out[s] = h_d_out[s];
  }
  a.size = (uint64_t)size; a.total = batch;
  a.idx_in = d_idx_in; a.idx_mode = d_idx_in ? IDX_INJECT : IDX_PHILOX;
  a.seed = seed; a.counter = counter; a.rng_stream = rng_stream; a.oidx = d_out_idx;
  const int sms = sm_count(device);
  int64_t blocks = (batch + 7) / 8;
  if (blocks > sms * 8) blocks = sms * 8;
  rb_gather_segments<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DDRL_LAUNCH_CHECK();
  return 0;
}

int ddrl_rb_counts(ddrl_rb_t rb, int64_t* ptr, int64_t* size, int64_t* capacity, int64_t* steps,
                   int64_t* sample_times) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_counts: NULL handle");
  RbLock lock(rb->mu);
  if (ptr) *ptr = rb->ptr;
  if (size) *size = rb->size;
  if (capacity) *capacity = rb->cap;
  if (steps) *steps = rb->steps;
  if (sample_times) *sample_times = rb->sample_times;
  return 0;
}

int ddrl_rb_note_samples(ddrl_rb_t rb, int64_t n_batches) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_note_samples: NULL handle");
  RbLock lock(rb->mu);
  rb->sample_times += n_batches;
  return 0;
}

/* A consumer that reads the ring from its OWN kernel (the fused sample -> update step, ddrl_sac_step_from_buffer) brackets
 * the launch of that kernel: read_begin takes the handle's lock, makes `stream` wait for every earlier store issued on
 * another stream and reports the sampling range; read_end counts the samples, records the read for later stores and
 * releases the lock. */
int ddrl_rb_read_begin(ddrl_rb_t rb, void* stream, int64_t* size) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_read_begin: NULL handle");
  rb->mu.lock();
  int rc = order_before(rb, (cudaStream_t)stream, false);
  if (rc) { rb->mu.unlock(); return rc; }
  if (size) *size = rb->size;
  return 0;
}

int ddrl_rb_read_end(ddrl_rb_t rb, void* stream, int64_t n_batches) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_read_end: NULL handle");
  rb->sample_times += n_batches;
  int rc = order_after(rb, (cudaStream_t)stream, false);
  rb->mu.unlock();
  return rc;
}

int ddrl_rb_layout(ddrl_rb_t rb, int* obs_dim, int* act_dim, int* row_floats, void** d_ring) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_layout: NULL handle");
  if (obs_dim) *obs_dim = rb->D;
  if (act_dim) *act_dim = rb->A;
  if (row_floats) *row_floats = rb->row_f;
  if (d_ring) *d_ring = rb->ring;
  return 0;
}

int ddrl_rb_export(ddrl_rb_t rb, float* d_obs1, float* d_obs2, float* d_acts, float* d_rews,
                   float* d_done, void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_export: NULL handle");
  if (!d_obs1 || !d_obs2 || !d_acts || !d_rews || !d_done)
    return fail(DDRL_EINVAL, "ddrl_rb_export: NULL output array");
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  int rc = order_before(rb, (cudaStream_t)stream, false);
  if (rc) return rc;
  GatherArgs a;
  a.ring = reinterpret_cast<const float4*>(rb->ring);
  a.D = rb->D; a.A = rb->A; a.row_f4 = rb->row_f4; a.used_f4 = rb->used_f4;
  a.size = (uint64_t)rb->cap; a.total = rb->cap;
  a.idx_in = nullptr; a.idx_mode = IDX_IDENTITY;
  a.seed = a.counter = 0; a.rng_stream = 0;
  a.o1 = d_obs1; a.o2 = d_obs2; a.oa = d_acts; a.orw = d_rews; a.od = d_done; a.oidx = nullptr;
  a.nshards = 0;
  if ((rc = launch_gather(rb, a, (cudaStream_t)stream))) return rc;
  return order_after(rb, (cudaStream_t)stream, false);
}

int ddrl_rb_import(ddrl_rb_t rb, const float* d_obs1, const float* d_obs2, const float* d_acts,
                   const float* d_rews, const float* d_done, int64_t ptr, int64_t size,
                   int64_t steps, int64_t sample_times, void* stream) {
  if (!rb) return fail(DDRL_EINVAL, "ddrl_rb_import: NULL handle");
  if (!d_obs1 || !d_obs2 || !d_acts || !d_rews || !d_done)
    return fail(DDRL_EINVAL, "ddrl_rb_import: NULL input array");
  if (ptr < 0 || ptr >= rb->cap || size < 0 || size > rb->cap || steps < 0 || sample_times < 0)
    return fail(DDRL_EINVAL, "ddrl_rb_import: inconsistent counters ptr=%lld size=%lld cap=%lld",
                (long long)ptr, (long long)size, (long long)rb->cap);
  DeviceGuard guard(rb->device);
  RbLock lock(rb->mu);
  int rc = order_before(rb, (cudaStream_t)stream, true);
  if (rc) return rc;
  rc = launch_store<float>(rb, d_obs1, d_acts, d_rews, d_obs2, d_done, 0, rb->cap, (cudaStream_t)stream);
  if (rc) return rc;
  rb->ptr = ptr; rb->size = size; rb->steps = steps; rb->sample_times = sample_times;
  return order_after(rb, (cudaStream_t)stream, true);
}

}  // extern "C"
