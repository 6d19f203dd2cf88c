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
/* same from ONE host block laid out like the handle's device staging — five sub-arrays obs | next_obs | acts | rews |
 * done, each starting at a 256-byte multiple (sizes n*D, n*D, n*A, n, n elements of in_dtype rounded up to 256 bytes;
 * ddrl_rb_store_block_bytes gives the block size): a single H2D copy.  Vectorised producers stage their rows in such a
 * pinned block (the by-value capture of `replay_buffer.store.remote`, algos/sac1/sac1.py:195). */
int64_t ddrl_rb_store_block_bytes(ddrl_rb_t rb, int64_t n, int in_dtype);
/* ddrl_rb_store_batch_host with BY-VALUE capture: the five host arrays (any host memory) are copied at call time into a
 * pinned block owned by the handle (two blocks, alternated; a block waits for the H2D copy that last read it), then
 * stored like ddrl_rb_store_block_host.  The caller may overwrite its arrays as soon as the call returns — the semantics
 * of `replay_buffer.store.remote(o, a, r, o2, d)` (algos/sac1/sac1.py:195), whose arguments are pickled at call time. */
int ddrl_rb_store_batch_host_copy(ddrl_rb_t rb, const void* h_obs, const void* h_act, const void* h_rew,
                                  const void* h_next_obs, const void* h_done, int64_t n, int in_dtype, void* stream);
int ddrl_rb_store_block_host(ddrl_rb_t rb, const void* h_block, int64_t n, int in_dtype, void* stream);

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
 * next 8-byte boundary;  h_idx_in (nullable) is a HOST index stream.  Synchronises `stream`.
 * When the block is pinned host memory and the result is <= 2 MB of rows <= 512 B, the gather kernel writes it directly
 * through the block's device mapping (no staging, no DMA copy; DDRL_ZERO_COPY=0 disables). */
int ddrl_rb_sample_host(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_idx_in,
                        uint64_t seed, uint64_t counter, uint32_t rng_stream,
                        void* h_out_block, int64_t block_bytes, void* stream);
/* same without the final synchronisation: the block is complete when `stream` has executed the call (record an event
 * after it).  This is what a prefetcher uses — the reference's Cache process samples ahead of the learner into a
 * Queue(10) (algos/sac1/sac1.py:103-130); here the gather and the D2H copy of batch k+1 run on the prefetcher's stream
 * while the learner's stream runs update k. */
int ddrl_rb_sample_host_async(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_idx_in,
                              uint64_t seed, uint64_t counter, uint32_t rng_stream,
                              void* h_out_block, int64_t block_bytes, void* stream);
/* bytes ddrl_rb_sample_host writes for n = batch*n_batches rows */
int64_t ddrl_rb_sample_block_bytes(ddrl_rb_t rb, int64_t n);

/* Replay sharded over the GPUs of one node (the reference shards replay into several buffer actors and
 * picks one shard per call, algos/sac1/sac_ray.py:137-141,246).  Each rank owns one ring; for the
 * GLOBAL-UNIFORM mode every rank maps its peers' rings (CUDA IPC) and the gather kernel reads remote
 * rows directly over NVLink — no collective on the data path.
 *   ddrl_rb_ipc_export   64-byte handle of this ring (exchange it with torch.distributed all_gather)
 *   ddrl_rb_peer_attach  register shard `shard` of `n_shards` (<= 8): h_handle64 == NULL marks this
 *                        rank's own ring, otherwise the peer's exported handle is opened
 *   ddrl_rb_sample_global  like ddrl_rb_sample, but indices address the concatenation of all shards:
 *                        g in [0, sum(h_shard_sizes)) -> shard s, row g - sum(sizes[:s]); h_shard_sizes
 *                        are the current fill counts of the shards (host array, n_shards entries).
 *                        The caller orders remote stores against this call (e.g. a barrier between
 *                        the store phase and the sample phase of a synchronous step). */
int ddrl_rb_ipc_export(ddrl_rb_t rb, void* h_handle64);
int ddrl_rb_peer_attach(ddrl_rb_t rb, int n_shards, int shard, const void* h_handle64);
int ddrl_rb_sample_global(ddrl_rb_t rb, int64_t batch, int64_t n_batches, const int64_t* h_shard_sizes,
                          const int64_t* d_idx_in, uint64_t seed, uint64_t counter, uint32_t rng_stream,
                          float* d_out_obs1, float* d_out_obs2, float* d_out_acts, float* d_out_rews,
                          float* d_out_done, int64_t* d_out_idx, void* stream);

/* Frame-stack replay (BASELINE.json config 4; a synthetic extension — the reference has no uint8 frame
 * buffer, its nearest relatives are the dqn-family ring algos/dqn/train.py:37-80 and the env-side
 * FrameStack algos/trading_env.py:289-325).  Frames live in a caller-owned ring d_frames[capacity]
 * of frame_bytes each (one frame per env step) with per-transition scalars d_act/d_rew/d_done[capacity];
 * transition i is obs1 = frames[i-stack+1..i], obs2 = frames[i-stack+2..i+1] (ring positions, no
 * episode-boundary handling).  Gathers `batch` transitions: injected ring positions (caller's responsibility that the
 * window does not cross the write head) or Philox-drawn: an age u uniform over [0, size - stack) counted from the OLDEST
 * frame, i = (oldest + stack-1 + u) % capacity — `oldest` is 0 until the ring wraps and the write pointer afterwards, so
 * no drawn window straddles the write head.  Outputs: obs1, obs2 [batch, stack*frame_bytes] bytes; acts, rews, done
 * [batch] f32.  One bulk copy global -> shared per frame and one or two shared -> global (TMA engine, no register pass);
 * the ring and the stacked outputs must be 16-byte aligned. */
int ddrl_fb_sample_stack(int device, const void* d_frames, int64_t frame_bytes, int stack, int64_t capacity,
                         int64_t size, int64_t oldest, const float* d_act, const float* d_rew, const float* d_done,
                         int64_t batch, const int64_t* d_idx_in, uint64_t seed, uint64_t counter,
                         uint32_t rng_stream, void* d_out_obs1, void* d_out_obs2, float* d_out_acts,
                         float* d_out_rews, float* d_out_done, int64_t* d_out_idx, void* stream);
/* store side of the same ring: frame j of d_in_frames [n, frame_bytes] goes to slot (ptr + j) % capacity, the scalars of
 * the env step that produced it (d_in_act/rew/done [n]) to the slot before it; n > capacity keeps the last `capacity`
 * frames, like n sequential stores of the dqn-family ring (algos/dqn/train.py:60-67).  The caller advances ptr / size. */
int ddrl_fb_store_frames(int device, void* d_frames, int64_t frame_bytes, int64_t capacity, int64_t ptr, float* d_act,
                         float* d_rew, float* d_done, const void* d_in_frames, const float* d_in_act,
                         const float* d_in_rew, const float* d_in_done, int64_t n, void* stream);

/* N-step sequence replay — sample_batch of the SQN_N_STEP ring (algos/sac1/sac_ray.py:34-83: rows of obs [Ln+1, D],
 * acts [Ln, A], rews [Ln], done [Ln]).  The ring is a caller-owned device array [capacity, row_floats] whose rows are
 * the concatenation of `nseg` float segments (start h_seg_off[s], width h_seg_w[s], row_floats a multiple of 4);
 * batch rows drawn like ddrl_rb_sample (injected indices or Philox) are gathered into one dense output
 * [batch, h_seg_w[s]] per segment (h_d_out[s], device pointers in a host array). */
int ddrl_seg_sample(int device, const float* d_ring, int row_floats, int64_t size, int nseg, const int* h_seg_off,
                    const int* h_seg_w, float* const* h_d_out, int64_t batch, const int64_t* d_idx_in, uint64_t seed,
                    uint64_t counter, uint32_t rng_stream, int64_t* d_out_idx, void* stream);

/* store side of the N-step ring (algos/sac1/sac_ray.py:52-68: `buffer[ptr] = ...` per field): row j of the dense inputs
 * h_d_in[s] [n, h_seg_w[s]] is packed into ring slot (ptr + j) % capacity (gaps and row padding are zero-filled; segments
 * in ascending offset order); n > capacity keeps the last `capacity` rows.  The caller advances ptr / size. */
int ddrl_seg_store(int device, float* d_ring, int row_floats, int64_t capacity, int64_t ptr, int nseg, const int* h_seg_off,
                   const int* h_seg_w, const float* const* h_d_in, int64_t n, void* stream);

/* ReplayBuffer.get_counts()  (algos/sac1/sac1.py:62-63) plus ptr / capacity.  Any out may be NULL. */
int ddrl_rb_counts(ddrl_rb_t rb, int64_t* ptr, int64_t* size, int64_t* capacity, int64_t* steps,
                   int64_t* sample_times);
/* geometry of the packed row: floats per padded row, device pointer of the ring (for zero-copy
 * consumers such as the fused sample->update step).  Any out may be NULL. */
int ddrl_rb_layout(ddrl_rb_t rb, int* obs_dim, int* act_dim, int* row_floats, void** d_ring);
/* sample_times += n_batches for a consumer that gathered from the ring itself (the fused sample->update step) */
int ddrl_rb_note_samples(ddrl_rb_t rb, int64_t n_batches);
/* Concurrency contract of a ddrl_rb_t (the Ray-actor semantics of the reference's ReplayBuffer, example/dsac.py:13,
 * algos/sac1/sac1.py:27, 195): every entry point may be called from any host thread on any stream.  Reservations are
 * serialised by the handle, so the ring always equals SOME serial order of the calls; kernels issued on different streams
 * are ordered with events (a sample waits for earlier stores, a store for earlier samples and stores), so producers'
 * copies and store kernels overlap the learner's stream.  A consumer that reads the ring from its own kernel brackets the
 * launch with read_begin (takes the handle's lock, orders `stream`, reports the sampling range) / read_end (counts
 * n_batches samples, records the read, releases the lock). */
int ddrl_rb_read_begin(ddrl_rb_t rb, void* stream, int64_t* size);
int ddrl_rb_read_end(ddrl_rb_t rb, void* stream, int64_t n_batches);

/* ring <-> the reference's five arrays ([capacity,D],[capacity,D],[capacity,A],[capacity],[capacity]
 * f32, device memory): the on-disk format of algos/dqn/train.py:82-108 (save/load) goes through
 * these.  import also restores the counters (buffer_infos). */
int ddrl_rb_export(ddrl_rb_t rb, float* d_obs1, float* d_obs2, float* d_acts, float* d_rews,
                   float* d_done, void* stream);
int ddrl_rb_import(ddrl_rb_t rb, const float* d_obs1, const float* d_obs2, const float* d_acts,
                   const float* d_rews, const float* d_done, int64_t ptr, int64_t size,
                   int64_t steps, int64_t sample_times, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SAC1 learner — replaces Learner (algos/sac1/actor_learner.py:19-148; nets algos/sac1/core.py:15-121)
 *
 * All state (main / target weights, Adam moments, gradients, activations) is device-resident in
 * flat fp32 buffers.  "flat" weight vectors exchanged through this API are in the reference's
 * variable order  main/pi/dense{,_1,_2,_3}/{kernel,bias}, main/q1/dense{,_1,_2}/{kernel,bias},
 * main/q2/...  (kernel [in,out] row-major, then bias), P = ddrl_sac_param_count() floats.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ddrl_sac* ddrl_sac_t;

/* Learner.__init__ (actor_learner.py:20-123).  hidden sizes (h1,h2) are parameters (the reference
 * hard-wires core.py:91's (400,300) default); alpha < 0 selects the entropy-alpha ('auto') branch
 * (actor_learner.py:46-55, reference-intended semantics); act_scale = action_space.high[0]
 * (core.py:104-106).  Buffers are sized for batches up to max_batch.  Weights start at zero: call
 * ddrl_sac_set_weights.  gemm_mode: DDRL_GEMM_TC (tcgen05 tensor cores, 3xTF32 split, fp32-class accuracy),
 * DDRL_GEMM_FFMA (plain fp32 FFMA tiles), DDRL_GEMM_AUTO (tensor cores unless the environment says DDRL_GEMM=ffma). */
#define DDRL_GEMM_AUTO 0
#define DDRL_GEMM_TC 1
#define DDRL_GEMM_FFMA 2
int ddrl_sac_create(int device, int obs_dim, int act_dim, int h1, int h2, int max_batch, float gamma,
                    float polyak, float lr, float alpha, float act_scale, int gemm_mode, ddrl_sac_t* out);
int ddrl_sac_destroy(ddrl_sac_t sac);
int64_t ddrl_sac_param_count(ddrl_sac_t sac);

/* Learner.set_weights (actor_learner.py:125-127): assign main weights; also_target != 0 runs
 * target_init (target <- main), which the reference always does for the Learner. */
int ddrl_sac_set_weights(ddrl_sac_t sac, const float* d_flat, int also_target, void* stream);
/* Learner.get_weights (actor_learner.py:129-133) for which = 0 (main); 1 = target, 2 = Adam m,
 * 3 = Adam v, 4 = last reduced gradient (diagnostics / parity tests). */
int ddrl_sac_get_weights(ddrl_sac_t sac, float* d_flat, int which, void* stream);

/* Learner.train(batch) (actor_learner.py:135-142) = one sess.run(step_ops): forward, both losses,
 * backward, Adam(pi), Adam(q), polyak (and the alpha step in 'auto' mode), replayed as one CUDA
 * graph.  Batch arrays are the ReplayBuffer.sample_batch dict ([B,D],[B,D],[B,A],[B],[B] f32).
 * d_noise: the three N(0,1) draws of the step, [3,B,A] (eps for pi(x), pi(x2), target pi(x2)), or
 * NULL to draw them on the device (Philox4x32-10 + Box-Muller keyed by `seed` and the step count).
 * Fetches (all nullable, pre-update values like the reference's, actor_learner.py:97-101):
 * d_out_scalars[4] = pi_loss, q1_loss, q2_loss, alpha;  d_out_q1, d_out_q2, d_out_logp: [B]. */
int ddrl_sac_step(ddrl_sac_t sac, const float* d_obs1, const float* d_obs2, const float* d_acts,
                  const float* d_rews, const float* d_done, int batch, const float* d_noise,
                  uint64_t seed, float* d_out_scalars, float* d_out_q1, float* d_out_q2,
                  float* d_out_logp, void* stream);

/* Data-parallel split of the same step (the reference's compute_gradients / apply_gradients stubs,
 * actor_learner.py:144-148, are empty): compute_grads leaves the flat gradient in the buffer
 * returned by ddrl_sac_grad_buffer (all-reduce it across ranks, e.g. NCCL SUM), apply_grads
 * multiplies it by the grad_scale given to compute_grads (1/world_size) and runs Adam + polyak.
 * d_alpha_stat points at mean(logp_pi) of the local batch (average it too in 'auto' mode). */
int ddrl_sac_compute_grads(ddrl_sac_t sac, const float* d_obs1, const float* d_obs2,
                           const float* d_acts, const float* d_rews, const float* d_done, int batch,
                           const float* d_noise, uint64_t seed, float grad_scale,
                           float* d_out_scalars, float* d_out_q1, float* d_out_q2,
                           float* d_out_logp, void* stream);
int ddrl_sac_grad_buffer(ddrl_sac_t sac, float** d_grads, int64_t* count, float** d_alpha_stat);
/* Fused `batch = replay_buffer.sample_batch(B); agent.train(batch)` (algos/sac1/sac1.py:146-148; example/model.py:92-101
 * Model.train(replay_buffer, args)): the step's first kernel gathers the batch straight from the ring, drawing row i as
 * philox_index(i, rb_seed, rb_counter, rb_stream, size) — the same rows ddrl_rb_sample(rb, batch, 1, NULL, rb_seed,
 * rb_counter, rb_stream, ...) would return — so the sampled batch never makes a round trip through HBM arrays.
 * Uses the fused data-parallel exchange when peers are attached.  Counts as one sample_batch call. */
int ddrl_sac_step_from_buffer(ddrl_sac_t sac, ddrl_rb_t rb, int batch, uint64_t rb_seed, uint64_t rb_counter,
                              uint32_t rb_stream, const float* d_noise, uint64_t seed, float* d_out_scalars,
                              float* d_out_q1, float* d_out_q2, float* d_out_logp, void* stream);
/* `agent.train(batch)` with the batch in HOST memory, as the reference feeds it (feed_dict of numpy arrays,
 * algos/sac1/actor_learner.py:135-142): h_block is ONE block [obs1 | obs2 | acts | rews | done] f32 for `batch` rows —
 * the layout ddrl_rb_sample_host delivers (pinned memory recommended).  One H2D copy, the update, one D2H copy of
 * (pi_loss, q1_loss, q2_loss, alpha) into h_out_scalars (nullable; pinned); nothing waits for the GPU — the scalars are
 * valid once `stream` has executed the call.  Policy noise is drawn on the device.  Uses the fused data-parallel
 * exchange when peers are attached.  A pinned block of <= 2 MB is read by the step's first kernel through its device
 * mapping and the scalars are written straight to the pinned h_out_scalars (no DMA copies; DDRL_ZERO_COPY=0 disables). */
int ddrl_sac_step_host(ddrl_sac_t sac, const void* h_block, int batch, uint64_t seed, float* h_out_scalars,
                       float* d_out_q1, float* d_out_q2, float* d_out_logp, void* stream);
int ddrl_sac_apply_grads(ddrl_sac_t sac, int batch, void* stream);
/* Fused data-parallel mode (one process per GPU of one node, 2..8 ranks; replaces the NCCL all-reduce between
 * compute_grads and apply_grads, i.e. what `north_star` asks of algos/sac1's multi-learner setup, sac1.py:273-276):
 *   ddrl_sac_comm_export  -> 64-byte CUDA IPC handle of this rank's gradient exchange buffer
 *   ddrl_sac_comm_attach  -> h_handles = world x 64 bytes (rank order, as exported); maps every peer's buffer.
 * Afterwards compute_grads leaves the flat gradient in the exchange buffer and apply_grads runs ONE kernel that
 * exchanges arrival flags with all peers, reads their gradients over NVLink, sums them in rank order (replicas stay
 * bit-identical), scales by grad_scale and applies Adam + polyak.  All ranks must call the step functions in
 * lockstep.  ddrl_sac_comm_error reports a peer time-out (a rank died) observed by the kernel. */
int ddrl_sac_comm_export(ddrl_sac_t sac, void* h_handle64);
/* compute_grads + apply_grads of the fused data-parallel mode as ONE captured graph (same arguments as compute_grads) */
int ddrl_sac_step_dp(ddrl_sac_t sac, const float* d_obs1, const float* d_obs2, const float* d_acts,
                     const float* d_rews, const float* d_done, int batch, const float* d_noise, uint64_t seed,
                     float grad_scale, float* d_out_scalars, float* d_out_q1, float* d_out_q2, float* d_out_logp,
                     void* stream);
int ddrl_sac_comm_attach(ddrl_sac_t sac, int world, int rank, const void* h_handles);
/* The same attachment from pointers the caller already holds: d_bufs[r] = rank r's exchange buffer as mapped in THIS
 * process (ddrl_sac_comm_bytes bytes each, zero-filled, 16-byte aligned; e.g. a symmetric-memory allocation), and
 * optionally d_multicast = the NVLS multicast mapping of those buffers: the optimiser kernel then fetches the sum of all
 * ranks' gradients with ONE multimem.ld_reduce per 16 bytes — the NVSwitch adds — instead of reading every peer
 * ((N-1) P bytes per rank become P).  The one-kernel exchange only (DDRL_DP_V1 unset). */
int64_t ddrl_sac_comm_bytes(ddrl_sac_t sac);
int ddrl_sac_comm_attach_ptrs(ddrl_sac_t sac, int world, int rank, void* const* d_bufs, const void* d_multicast);
int ddrl_sac_comm_error(ddrl_sac_t sac, int* out_error);
/* profiling aid (handle created with DDRL_DP_TRACE=1 in the environment): %globaltimer stamps (ns) of the last fused
 * data-parallel optimiser launch, CTA 0: start, own gradient written, all gradients published, slice scattered, all
 * slices written, end, 0, 0.  Synchronises the device. */
int ddrl_sac_dp_trace(ddrl_sac_t sac, unsigned long long* h_out8);
/* Actor.get_action(o, deterministic) (algos/sac1/actor_learner.py:195-197) for n observations at once:
 * d_out_act[n, A] = act_scale * tanh(mu) (deterministic) or act_scale * tanh(mu + eps * std); eps from
 * d_noise [n, A] or, when NULL, Philox keyed by (seed, counter).  Uses the handle's main policy weights. */
int ddrl_sac_act(ddrl_sac_t sac, const float* d_obs, int n, int deterministic, const float* d_noise,
                 uint64_t seed, uint64_t counter, float* d_out_act, void* stream);
/* profiling aid: enqueue one phase of the step `reps` times (0..6: GEMM stages L1, L2, QL1, QL2, BQ, BP, BP3 — with
 * narrow inputs L1 / QL1 are empty and L2 / QL2 are the fused first + second layer launches;
 * 7 prologue, 8 policy heads, 9 Q heads + losses, 10 policy backward rows, 11 optimiser, 12..14 side-stream work).
 * reps < 0: the -reps launches run as the nodes of ONE CUDA graph (what a kernel costs inside the step's graph; the graph
 * of the last (batch, stage, reps) is cached, so call once to build it and time the second call). */
int ddrl_sac_debug_stage(ddrl_sac_t sac, int batch, int stage, int reps, void* stream);
/* Test entry for the tcgen05 3xTF32 GEMM alone: C[M,N] (splits > 1: `splits` partial outputs M*N floats apart)
 * = opA . opB from dense row-major fp32 device matrices A [a_rows,a_cols], B [b_rows,b_cols]; a_mn / b_mn = 1
 * when the contraction index is the ROW of the stored tensor (MN-major operand), 0 when it is the column.
 * bn = 64 | 128: output tile width of the launch (128 x bn tiles). */
int ddrl_debug_tc_gemm(int device, const float* d_a, int a_rows, int a_cols, int a_mn, const float* d_b, int b_rows,
                       int b_cols, int b_mn, float* d_c, int m, int n, int k, int splits, int bn, void* stream);
/* profiling aid: run the tensor-core launch of GEMM stage `stage` once with per-CTA phase time stamps
 * (d_trace [tiles, 8] of %globaltimer ns: start, setup done, first stage landed, MMAs issued, accumulator complete,
 * warp-0 epilogue done, all warps done) */
int ddrl_sac_trace_stage(ddrl_sac_t sac, int batch, int stage, unsigned long long* d_trace, int max_tiles, int* tiles,
                         void* stream);
/* optimiser step counters and log_alpha (synchronises `stream`).  Any out may be NULL. */
int ddrl_sac_state(ddrl_sac_t sac, int* t_pi, int* t_q, int* t_alpha, float* log_alpha, void* stream);

/* ---- discrete-action learners: DDQN (algos/dqn/actor_learner.py:20-130) and SQN (algos/sqn/actor_learner.py:20-125) ----
 * n_nets = 1: DDQN — q = mlp(x) -> [B, n_actions] (hidden h1, h2, relu; algos/dqn/core.py:53-63); loss
 *             0.5 * mean((r + gamma (1-d) q_target(x2)[argmax_a q_main(x2)] - q(x)[a])^2)   (actor_learner.py:41-55)
 * n_nets = 2: SQN  — q1, q2; backup r + gamma (1-d) (min(max_a q1_target(x2), max_a q2_target(x2)) - alpha * sum_a p log p),
 *             p = softmax(q1_main(x2) / alpha) (algos/sqn/core.py:30-45, actor_learner.py:43-58); loss = q1_loss + q2_loss.
 * Both: tf.train.AdamOptimizer(lr) on the main networks, then polyak averaging of every main variable into its target
 * (actor_learner.py:59-67).  Flat parameter vector = the reference's variable order, per network dense/kernel [D,h1],
 * dense/bias [h1], dense_1/kernel [h1,h2], dense_1/bias, dense_2/kernel [h2,n_actions], dense_2/bias. */
typedef struct ddrl_ql* ddrl_ql_t;
int ddrl_ql_create(int device, int obs_dim, int n_actions, int h1, int h2, int max_batch, int n_nets, float gamma,
                   float polyak, float lr, float alpha, ddrl_ql_t* out);
int ddrl_ql_destroy(ddrl_ql_t ql);
int64_t ddrl_ql_param_count(ddrl_ql_t ql);
/* Learner.set_weights (assign main; also_target = the reference's target_init) / get_weights; which: 0 main, 1 target,
 * 2 Adam m, 3 Adam v, 4 the last step's gradient */
int ddrl_ql_set_weights(ddrl_ql_t ql, const float* d_flat, int also_target, void* stream);
int ddrl_ql_get_weights(ddrl_ql_t ql, float* d_flat, int which, void* stream);
/* Learner.train(batch, cnt): one update on device arrays obs1, obs2 [batch, obs_dim], acts (action indices stored as
 * floats, as the dqn-family ring keeps them), rews, done [batch].  d_out_loss (nullable): n_nets per-network losses then
 * their sum; d_out_q (nullable): q(x) of each network, [n_nets, batch, n_actions] — the reference's fetch list. */
int ddrl_ql_step(ddrl_ql_t ql, const float* d_obs1, const float* d_obs2, const float* d_acts, const float* d_rews,
                 const float* d_done, int batch, float* d_out_loss, float* d_out_q, void* stream);
/* Actor side (algos/dqn/actor_learner.py:190-200, algos/sqn/core.py:30-45): main network `net`'s Q values for n
 * observations; the caller takes the argmax / epsilon-greedy / softmax(q / alpha) sample. */
int ddrl_ql_forward(ddrl_ql_t ql, const float* d_obs, int n, int net, float* d_out_q, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDRL_B200_H_ */
