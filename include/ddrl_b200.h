/*
 * ddrl_b200.h — C ABI of libddrl_b200.so: the B200-native (sm_100a) learner-side data path of
 * createamind/Distributed-DRL.
 *
 * The reference has NO native / FFI interface: its ReplayBuffer, ParameterServer and Learner are
 * Python classes defined inline in each driver script.  The entry points below are therefore what a
 * ctypes binding *inside those reference classes* would call (INTEGRATION.md shows that binding);
 * each one cites the reference method it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C symbols, opaque handles, no torch / CUDA types in signatures (streams are `void*`
 *     holding a cudaStream_t; NULL = the legacy default stream);
 *   - every function returns 0 on success or a negative DDRL_E* code; ddrl_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - pointers named `d_*` are DEVICE pointers owned by the caller (e.g. torch tensors), pointers
 *     named `h_*` are HOST pointers; all device work is enqueued on `stream` and is asynchronous
 *     unless the function name ends in `_host` (those synchronise `stream` before returning,
 *     because their results land in host memory);
 *   - a handle is bound to one device and must be driven by one host thread at a time, with all
 *     calls on one stream (or externally ordered) — the same one-call-at-a-time discipline a Ray
 *     actor gives the reference classes;
 *   - there is no CPU fallback: on a machine without a CUDA device every compute entry point
 *     fails with DDRL_ECUDA.
 */
#ifndef DDRL_B200_H_
#define DDRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDRL_ABI_VERSION 1

/* error codes */
#define DDRL_OK        0
#define DDRL_EINVAL   -1  /* bad argument (message says which)                                  */
#define DDRL_ECUDA    -2  /* CUDA runtime error (message carries cudaGetErrorString)            */
#define DDRL_EEMPTY   -3  /* sample from an empty ring: the reference raises ValueError there   */
#define DDRL_ENOMEM   -4
#define DDRL_ESTATE   -5  /* call not valid in the handle's current state                       */

/* element types of store inputs / ring observations */
#define DDRL_F32 0
#define DDRL_F64 1   /* store inputs only: cast to f32 round-to-nearest-even, as numpy assignment does */
#define DDRL_U8  2   /* frame observations (Atari-shaped rows)                                  */

int         ddrl_abi_version(void);
const char* ddrl_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
int64_t     ddrl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Replay ring  — replaces ReplayBuffer (example/dsac.py:14-48, algos/sac1/sac1.py:28-63,
 *                scalar-action flavour algos/dqn/train.py:37-80)
 *
 * HBM layout (internal, not the reference's five arrays): one packed row per transition,
 *     [ obs1 (D) | obs2 (D) | acts (A) | rew | done | zero pad ]  padded to a multiple of 4 floats,
 * so that a sampled transition is ONE contiguous, 16-byte aligned read instead of five scattered
 * ones.  ddrl_rb_export / ddrl_rb_import convert to / from the reference's five-array form.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ddrl_rb* ddrl_rb_t;

/* ReplayBuffer.__init__(obs_dim, act_dim, size)  (example/dsac.py:20-27).  Allocates and zero-fills
 * the ring on `device`.  act_dim >= 1 (the dqn flavour's scalar action is act_dim = 1). */
int ddrl_rb_create(int device, int obs_dim, int act_dim, int64_t capacity, ddrl_rb_t* out);
int ddrl_rb_destroy(ddrl_rb_t rb);

/* n x ReplayBuffer.store(obs, act, rew, next_obs, done)  (example/dsac.py:29-37), batched: the
 * result equals n sequential store() calls in row order — rows land at (ptr+i) % capacity, ptr and
 * size advance once, steps += n; if n > capacity only the last `capacity` rows survive.
 * Inputs are row-major [n,D], [n,A], [n], [n,D], [n] arrays of `in_dtype` (DDRL_F32 | DDRL_F64). */
int ddrl_rb_store_batch(ddrl_rb_t rb, const void* d_obs, const void* d_act, const void* d_rew,
                        const void* d_next_obs, const void* d_done, int64_t n, int in_dtype,
                        void* stream);
/* same, inputs in HOST memory (pinned memory makes the copies asynchronous); copies are staged
 * through device buffers owned by the handle.  Does not synchronise: the host arrays must stay
 * valid until `stream` has passed this call (the Python layer owns pinned staging for that). */
int ddrl_rb_store_batch_host(ddrl_rb_t rb, const void* h_obs, const void* h_act, const void* h_rew,
                             const void* h_next_obs, const void* h_done, int64_t n, int in_dtype,
                             void* stream);

/* n_batches x ReplayBuffer.sample_batch(batch)  (example/dsac.py:39-45; algos/sac1/sac1.py:53-60)
 * in one launch.  Index stream, uniform on [0,size) with replacement like np.random.randint:
 *     d_idx_in != NULL : injected int64 [n_batches*batch] (parity mode: bit-exact vs the reference
 *                        numpy buffer driven by the same indices); out-of-range -> DDRL_EINVAL is
 *                        NOT checked on the device, the caller guarantees 0 <= idx < size;
 *     d_idx_in == NULL : Philox4x32-10, key = seed, counter = (ordinal, counter, rng_stream),
 *                        idx = mulhi64(u64, size).
 * Outputs are the reference's dict arrays, row-major f32: obs1,obs2 [n_batches*batch, D],
 * acts [.., A], rews, done [..]; d_out_idx (nullable) receives the indices used.
 * sample_times += n_batches.  Returns DDRL_EEMPTY when size == 0. */
int ddrl_rb_sample(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* d_idx_in,
                   uint64_t seed, uint64_t counter, uint32_t rng_stream,
                   float* d_out_obs1, float* d_out_obs2, float* d_out_acts, float* d_out_rews,
                   float* d_out_done, int64_t* d_out_idx, void* stream);
/* same, result delivered to ONE host block (pinned recommended) laid out as
 *     [obs1 | obs2 | acts | rews | done] f32, each segment n*width floats, then [idx] int64 at the
 * next 8-byte boundary;  h_idx_in (nullable) is a HOST index stream.  Synchronises `stream`. */
int ddrl_rb_sample_host(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_idx_in,
                        uint64_t seed, uint64_t counter, uint32_t rng_stream,
                        void* h_out_block, int64_t block_bytes, void* stream);
/* bytes ddrl_rb_sample_host writes for n = batch*n_batches rows */
int64_t ddrl_rb_sample_block_bytes(ddrl_rb_t rb, int64_t n);

/* ReplayBuffer.get_counts()  (algos/sac1/sac1.py:62-63) plus ptr / capacity.  Any out may be NULL. */
int ddrl_rb_counts(ddrl_rb_t rb, int64_t* ptr, int64_t* size, int64_t* capacity, int64_t* steps,
                   int64_t* sample_times);
/* geometry of the packed row: floats per padded row, device pointer of the ring (for zero-copy
 * consumers such as the fused sample->update step).  Any out may be NULL. */
int ddrl_rb_layout(ddrl_rb_t rb, int* obs_dim, int* act_dim, int* row_floats, void** d_ring);

/* ring <-> the reference's five arrays ([capacity,D],[capacity,D],[capacity,A],[capacity],[capacity]
 * f32, device memory): the on-disk format of algos/dqn/train.py:82-108 (save/load) goes through
 * these.  import also restores the counters (buffer_infos). */
int ddrl_rb_export(ddrl_rb_t rb, float* d_obs1, float* d_obs2, float* d_acts, float* d_rews,
                   float* d_done, void* stream);
int ddrl_rb_import(ddrl_rb_t rb, const float* d_obs1, const float* d_obs2, const float* d_acts,
                   const float* d_rews, const float* d_done, int64_t ptr, int64_t size,
                   int64_t steps, int64_t sample_times, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDRL_B200_H_ */
