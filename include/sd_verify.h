/*
 * sd_verify.h -- C ABI of the B200-native speculative-decoding verify path.
 *
 * One shared library (libsd_verify_b200.so, built from llm-inference-lab_b200/csrc for sm_100a)
 * exports exactly the entry points below.  They are what a foreign-function binding of the
 * reference (GogoRit/llm-inference-lab) would bind for this path; each entry cites the reference
 * interface it replaces (paths relative to the reference root).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  The compute entry points
 *     (sd_verify_*, sd_kv_append*, sd_verify_attention, sd_sample_bonus, sd_exchange_wait) neither
 *     allocate nor synchronise: every call only enqueues work on `stream`, so calls are CUDA-graph
 *     capturable and callers keep ownership of every buffer.  Exceptions, none of them on the step:
 *     sd_exchange_alloc / _open / _close / _free manage the peer-mapped exchange buffer (cudaMalloc +
 *     CUDA IPC), and the sd_*_set_* hooks store process-wide tuning choices for evidence runs.
 *   - sd_step_args, sd_exchange and sd_kv_layer are plain C structs (sd_exchange and the sd_kv_layer
 *     table live in DEVICE memory, sd_step_args in host memory).
 *   - Return value: 0 = success; > 0 = a cudaError_t from the launch; < 0 = an sd_status
 *     argument error (nothing was enqueued).  sd_error_string() names either.
 *   - Strides are in ELEMENTS for logits / ids and in BYTES for KV tensors (stated per argument).
 *   - No CPU fallback exists: on a machine without an sm_100 device the launches fail loudly.
 */
#ifndef SD_VERIFY_H_
#define SD_VERIFY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 3

/* element type of logits / KV / attention tensors */
enum sd_dtype { SD_F32 = 0, SD_F16 = 1, SD_BF16 = 2 };
/* element type of token ids */
enum sd_id_dtype { SD_I32 = 0, SD_I64 = 1 };

enum sd_status {
    SD_OK = 0,
    SD_ERR_INVALID_ARG = -1, /* null pointer, negative size, unknown dtype */
    SD_ERR_SHAPE = -2,       /* reference raises AssertionError: B / K / H / D mismatch */
    SD_ERR_WORKSPACE = -3,   /* workspace too small or missing */
    SD_ERR_ALIGNMENT = -4,   /* pointer / stride not aligned to the element size */
    SD_ERR_UNSUPPORTED = -5, /* e.g. head_dim not supported by the attention kernel */
    SD_ERR_CAPACITY = -6     /* KV arena capacity would be exceeded (checked on host-known sizes) */
};

/* acceptance rule evaluated by sd_verify_step */
enum sd_accept_mode {
    SD_ACCEPT_GREEDY = 0,     /* argmax == draft           reference: src/kernels/reference.py:37-54 */
    SD_ACCEPT_TYPICAL = 1,    /* softmax(target)[draft] >= threshold
                                 reference: src/specdec/policies/policies.py:358-373 */
    SD_ACCEPT_CONFIDENCE = 2, /* max softmax(logits) >= threshold
                                 reference: src/specdec/policies/policies.py:240-250 */
    SD_ACCEPT_TOPK = 3        /* draft token among the target's top-k: fewer than `threshold` (= k) logits
                                 strictly above it; token_prob then returns that count, row_sumexp is
                                 meaningless.  reference: src/specdec/policies/policies.py:300-314 */
};

int sd_abi_version(void);
/* sizeof / offsetof of the public structs, for bindings to check their mirror of the layout:
 * 0 sizeof(sd_kv_layer), 1 sizeof(sd_exchange), 2 offsetof(sd_exchange, peer), 3 sizeof(sd_step_args),
 * 4..7 offsetof(sd_step_args, layers | row_bytes | seq_lens | step_state); other values -> 0. */
size_t sd_struct_layout(int which);
const char* sd_error_string(int code);
/* Programmatic dependent launch for every kernel of the library (default on; env SD_B200_PDL=0 turns
 * it off): a kernel's CTAs may become resident while the previous kernel on the stream drains, and
 * block in griddepcontrol.wait until it has completed.  Capturable in CUDA graphs. */
int sd_set_pdl(int enabled);
/* Fills SM count and compute capability of the current device.  Returns 0 or a cudaError_t. */
int sd_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Device-side waits are BOUNDED.  The kernels of this library wait across CTAs (a verify finalizer for the partial
 * results of its rows, an append for a sequence's completion word) and across GPUs (the exchange gather for the
 * peers' records, the start gate).  Every such poll gives up after 2 s of %globaltimer, records which kind of
 * wait expired and lets the launch run to its end (its outputs are then garbage; accesses stay in bounds), so a dead
 * peer, a stale workspace or a violated dispatch-order assumption (see sd_verify_append) surfaces here as a code
 * instead of a hung GPU.  sites_out (nullable) receives the OR of: 1 partial slots, 2 sequence word / flag,
 * 4 all-sequence wait, 8 peer words, 16 start gate; 0 = no wait has expired since the last clear.  Reads device
 * memory with a synchronising copy: call it at step boundaries / in error paths, not per launch.
 * Returns 0 or a cudaError_t. */
int sd_device_error(int clear, unsigned* sites_out);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU result exchange over peer memory (batch-sharded step, one process per GPU).
 *
 * No reference counterpart: the reference is single-device.  Sequences are
 * independent, so ranks only exchange the packed [B_local, K+2] verify records.  Instead of a NCCL
 * all-gather launch per step, the verify kernel's sequence finalizer stores each record straight into
 * EVERY rank's exchange buffer (st.global on CUDA-IPC mapped peer pointers, NVLink).  Every value
 * travels in ONE 8-byte store together with an epoch tag -- {epoch + 1, value} -- so the data path needs
 * no fence, atomic or arrival counter: the consumer (the last CTA of sd_kv_append, or
 * sd_exchange_wait) polls the words of the current epoch until their tags match, copies the values
 * out and advances the device-side epoch.  Words live in 4 slots indexed by epoch & 3, which covers
 * lag <= 1 (a producer is never more than two epochs ahead of a slot still being read).
 *
 * Exchange buffer (identical on every rank, sd_exchange_bytes): 256-byte header {u32 epoch} +
 * u64 words[4][world * rows_per_rank][width].  The descriptor lives in DEVICE memory; peer[rank] is the
 * rank's own buffer.  Rank r holds total_rows / world (+1 for the first total_rows % world ranks) rows.
 */
#define SD_MAX_PEERS 16
typedef struct sd_exchange {
    int32_t world, rank;
    int32_t rows_per_rank; /* record rows reserved per rank (max local batch) */
    int32_t width;         /* K + 2 */
    int32_t total_rows;    /* sum of the local batches of all ranks = records arriving per step */
    int32_t lag;           /* 0: a step gathers its own records (ranks in lockstep); 1: the previous step's
                              (a rank may run one step ahead of the slowest; the table lags one step) */
    int32_t reserved[2];
    void* peer[SD_MAX_PEERS];
} sd_exchange;

size_t sd_exchange_bytes(int world, int rows_per_rank, int width);
/* cudaMalloc + zero an exchange buffer on the current device; handle_out receives the 64-byte
 * cudaIpcMemHandle_t to send to the other ranks (nullable). */
int sd_exchange_alloc(size_t bytes, void** ptr_out, void* handle_out);
/* Map a peer's buffer from its 64-byte handle (cudaIpcOpenMemHandle, lazy peer access). */
int sd_exchange_open(const void* handle, void** peer_ptr_out);
int sd_exchange_close(void* peer_ptr);
int sd_exchange_free(void* ptr);
/* Wait until every rank's records of the current epoch have arrived in this rank's buffer, copy
 * them to out [world * rows_per_rank, width] int32 (nullable) and advance the epoch.  One small
 * kernel on `stream`; call it once per verify launch that carried `exchange`. */
int sd_exchange_wait(const sd_exchange* exchange, int32_t* out, void* stream);
/* Start gate: one tiny kernel on `stream` behind which the work enqueued after it starts (a) only once the host has
 * stored go_value into *host_go (nullable; a uint32 in pinned, device-accessible host memory -- set it after the
 * work behind the gate has been enqueued, so the GPU never outruns its launch queue), and (b) only once every rank
 * of `exchange` (nullable = single GPU) has reached the same gate: each rank stores an arrival word into every
 * peer's header and polls its own.  Ranks then leave the gate within ~2 us of each other whatever the skew of
 * their host threads was -- a lock-step multi-GPU loop (lag 0 exchange) otherwise absorbs that skew in its first
 * gather.  Collective over the ranks of the exchange; waits are bounded (sd_device_error, site 16).  No reference
 * counterpart (single GPU). */
int sd_start_gate(const sd_exchange* exchange, const void* host_go, unsigned go_value, void* stream);
/* Host-side self-test of the start gate's bounded-wait rule (no GPU): feeds `readings` (successive %globaltimer values,
 * ns, as the gate sees them every 128 polls) through the same inline function the kernel uses and returns the index of
 * the reading at which the gate gives up, -1 if it keeps waiting through all of them, -2 on a bad argument.  The rule:
 * the window restarts on a backwards clock and the budget must be exceeded twice, each time from a fresh reading, so
 * one step of the clock (seen on the first launch after seconds of GPU idle) opens nothing. */
int sd_selftest_gate_clock(const unsigned long long* readings, int n, unsigned long long budget_ns);

/* ------------------------------------------------------------------------------------------
 * Kernel 1: fused verify.
 *
 * Replaces  verify_prefix(logits[B,K,V], draft_ids[B,K]) -> (accept_len[B] i32, mask[B,K] u8)
 *           reference: src/kernels/reference.py:13-56 (oracle), src/kernels/triton/verify.py:101-169
 *           (on-GPU incumbent), src/kernels/cuda/verify.cu:154-217 (never registered);
 *           call site src/specdec/policies/policies.py:128-134.
 * and, when rows == K + 1, additionally the per-sequence bonus-token selection of
 *           src/specdec/core/pipeline.py:3140-3163 / :3174-3206 / :3341-3362 (greedy branch
 *           :135-137): bonus[b] = argmax(logits[b, accept_len[b], :]).
 *
 * logits      [B, rows, V], element strides (stride_b, stride_r, 1); rows == K or K + 1
 * draft_ids   [B, K], element strides (id_stride_b, 1)
 * accept_len  [B] int32                 (required)
 * accepted_mask [B, K] uint8            (nullable)
 * target_ids  [B, rows] int32           (nullable)  argmax of every row, torch.argmax semantics:
 *                                        first maximal index, NaN maximal (first NaN), -0 == +0
 * bonus       [B] int32                 (nullable)  target_ids[b, accept_len[b]] (rows == K+1) else -1
 * packed      [B, K + 2] int32          (nullable)  {accept_len, target_ids[0..K]} -- the record the
 *                                        multi-GPU allgather exchanges (target_ids[K] = -1 if rows == K)
 * exchange    (nullable, device pointer) when set, the finalizer of sequence b also stores the packed
 *             record into row rank * rows_per_rank + b of every peer's exchange buffer (see above);
 *             requires width == K + 2 and B <= rows_per_rank
 * seq_ready   (nullable) [B] int32, zero before the launch: flag b is raised (release) once every output
 *             of sequence b is final.  Pass the same array to the sd_kv_append call that follows on the
 *             stream: its CTAs then start on sequence b as soon as b is verified instead of waiting
 *             for the whole verify grid, and its last CTA re-zeroes the flags.
 * workspace   sd_verify_workspace_bytes(B, rows) bytes, 8-byte aligned, zeroed once with sd_verify_workspace_init;
 *             the kernel leaves it zeroed again, so one init serves any number of launches
 *             (including graph replays).  Must not be shared by launches that may overlap.
 *
 * One kernel launch.  Algorithmic traffic: B * rows * V * sizeof(element) bytes read once.
 * ------------------------------------------------------------------------------------------ */
size_t sd_verify_workspace_bytes(int B, int rows);
int sd_verify_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

int sd_verify_prefix(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                     const void* draft_ids, int id_dtype, int64_t id_stride_b,
                     int32_t* accept_len, uint8_t* accepted_mask, int32_t* target_ids,
                     int32_t* bonus, int32_t* packed, void* workspace, size_t workspace_bytes,
                     int B, int rows, int K, int V, void* stream, const sd_exchange* exchange,
                     int32_t* seq_ready);

/*
 * Same pass with softmax statistics ("sampling mode"): one read of the logits yields, per row,
 * the argmax, the row maximum and sum(exp(x - max)); the acceptance rule `mode` is evaluated on
 * device.  Replaces the per-sequence softmax policies of src/specdec/policies/policies.py:213-396
 * (TypicalAcceptancePolicy :358-373, ConfidenceThresholdPolicy :240-250) for a whole batch.
 *
 * inv_temperature multiplies the logits before the softmax (1.0f = none;
 *                 reference: src/specdec/core/pipeline.py:91-92 divides by temperature).
 * token_prob  [B, K] float (nullable)   softmax(logits[b,k])[draft_ids[b,k]]
 * row_max     [B, rows] float (nullable) max_v logits * inv_temperature
 * row_sumexp  [B, rows] float (nullable) sum_v exp(logits * inv_temperature - row_max)
 *                                        (max probability of the row = 1 / row_sumexp)
 * Other arguments as sd_verify_prefix.  accept_len is the leading run of positions passing `mode`.
 */
int sd_verify_step(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                   const void* draft_ids, int id_dtype, int64_t id_stride_b, int mode,
                   float threshold, float inv_temperature, int32_t* accept_len,
                   uint8_t* accepted_mask, int32_t* target_ids, int32_t* bonus, int32_t* packed,
                   float* token_prob, float* row_max, float* row_sumexp, void* workspace,
                   size_t workspace_bytes, int B, int rows, int K, int V, void* stream,
                   const sd_exchange* exchange, int32_t* seq_ready);

/* Tuning hook used by tools/ and bench.py sweeps: threads per CTA (128..1024, multiple of 32),
 * 16-byte vectors in flight per thread (1..16), 0 restores the built-in heuristic. */
int sd_verify_set_tuning(int threads, int vectors_per_thread);
/* Kernel variant: 0 = one CTA per (row, chunk) with register-staged 128-bit loads; 1 = persistent grid
 * with equal contiguous unit ranges, register-staged; 2 = persistent grid, cp.async.bulk (TMA 1-D)
 * copies into a shared-memory mbarrier ring consumed by 256 threads; 3 = one CTA per (row, chunk) whose
 * whole chunk (vectors_per_thread x 4 KB, <= 48 KB) is fetched by a single cp.async.bulk.  ctas_per_sm caps the persistent
 * grids (0 = occupancy limit for 1, two CTAs per SM for 2); for variant 2 sd_verify_set_tuning's
 * vectors_per_thread (2 or 4) selects the ring slot size (8 or 16 KB). */
int sd_verify_set_variant(int variant, int ctas_per_sm);
/* Evidence runs: per-CTA timestamps of the variant-0 kernel.  device_buffer holds capacity_ctas x 4
 * uint64 {globaltimer ns at entry, after the streaming scan, at exit, SM id}; NULL switches it off
 * (the default; the kernel then only tests one pointer). */
int sd_verify_set_trace(void* device_buffer, size_t capacity_ctas);

/* ------------------------------------------------------------------------------------------
 * Kernel 2: kv_append into a preallocated arena, with rollback of rejected slots.
 *
 * Replaces  kv_append(base_k, base_v, new_k, new_v) -> cat on dim 2
 *           reference: src/kernels/reference.py:59-93; call site
 *           src/specdec/models/hf_wrappers.py:985-1030 (one call per layer + device sync)
 * and       kv_append_with_mask(base_k, base_v, draft_k, draft_v, mask, accept_len, offset)
 *           reference: src/kernels/reference.py:96-159; src/kernels/cuda/kv_cache.cu:14-95
 * and the rollback of src/specdec/cache/kv_cache_manager.py:353-479 (realign_kv_cache), which in
 * the arena layout is a length update plus zeroing of the rejected slots.
 *
 * One launch serves every layer.  `layers` is a DEVICE array of n_layers records:
 */
typedef struct sd_kv_layer {
    void* arena_k;      /* [B, H, C, D] rows of row_bytes; byte strides arena_stride_{b,h}, row stride row_bytes */
    void* arena_v;
    const void* new_k;  /* [B, H, R, D] offered rows; byte strides new_stride_{b,h,r} */
    const void* new_v;
} sd_kv_layer;

/*
 * For every layer, K and V, sequence b and head h:
 *   n   = n_keep ? clamp(n_keep[b] + keep_bias, 0, R) : R
 *   dst = arena[b, h, seq_lens[b] + j]   for j in [0, R)
 *   accepted_mask == NULL : dst_j = new[b, h, j]          for j <  n
 *   accepted_mask != NULL : dst_j = new[b, h, idx_j(b)]   for j <  n, idx_j = position of the j-th set
 *                           bit of accepted_mask[b, 0..R)  (reference.py:149-157 compaction)
 *   zero_rejected         : dst_j = 0                     for n <= j < R      (reference.py:133-134)
 * and, after every copy has been issued, seq_lens_out[b] = seq_lens[b] + n (seq_lens_out may alias
 * seq_lens: the update is performed by the last CTA to retire).  n_keep is typically the
 * accept_len written by sd_verify_prefix on the same stream with keep_bias = 1 (accepted drafts +
 * the token they were conditioned on), so no host round trip separates verify from append.
 *
 * seq_lens    [B] int32  rows already valid per sequence (NULL = all zero)
 * capacity    C; the host-checkable bound max_seq_len_host + R <= C is verified when
 *             max_seq_len_host >= 0; on device, rows that would fall beyond C are dropped and
 *             *overflow_flag (nullable, int32) is set non-zero.
 * ticket      one int32 in device memory, zero before the first launch (the kernel re-zeroes it).
 * exchange / exchange_out  (nullable) multi-GPU step: the records that the verify launch of this step
 *             pushed into every rank's exchange buffer are gathered HERE, by the first CTA once its own
 *             rows are copied (poll the tagged words of the epoch -> copy the
 *             [world * rows_per_rank, width] table to exchange_out -> advance the epoch), so the
 *             exchange hides behind the append and needs no launch of its own.  Use this INSTEAD of
 *             sd_exchange_wait for that step.  With seq_ready the gathering CTA first waits for the flags
 *             of ALL sequences: the epoch must not advance while a finalizer has yet to read it.
 * exchange_packed  (nullable) [B, K+2] int32 packed records of THIS rank, complete before this launch
 *             (written by the verify launch): when set, the first CTA of this launch pushes them to
 *             every peer at its start -- the NVLink round trip then hides behind the append -- and
 *             the verify launch must be issued WITHOUT `exchange`.
 * seq_ready   (nullable) the flag array given to the sd_verify_* launch just before on the stream (n_keep
 *             must be that launch's accept_len; column kernel only, i.e. unmasked 16-byte aligned rows,
 *             else SD_ERR_UNSUPPORTED).  The launch then does NOT wait for the whole verify grid: every
 *             thread waits for its own sequence's flag, so the append of early sequences overlaps the
 *             verification of late ones; the last CTA waits for the verify grid, re-zeroes the flags
 *             and commits the lengths.
 * rows_ready_early  non-zero = the caller guarantees that the layer table and the offered rows were
 *             written at least two kernels before this call on the stream (in the verify step they
 *             come from the target forward; the kernel just before this one is the fused verify,
 *             which writes only accept_len).  The kernel then fetches the first rows of K and V
 *             before griddepcontrol.wait, overlapping its load phase with the tail of the previous
 *             kernel; all R <= 6 offered rows are read (rejected ones included).  0 = always safe.
 *
 * Algorithmic traffic per launch: n_layers * 2 * B * H * row_bytes * (mean n read + R written
 * when zero_rejected, else mean n written).
 */
int sd_kv_append(const sd_kv_layer* layers, int n_layers, int B, int H, int R, int64_t row_bytes,
                 int64_t capacity, int64_t arena_stride_b, int64_t arena_stride_h,
                 int64_t new_stride_b, int64_t new_stride_h, int64_t new_stride_r,
                 const int32_t* seq_lens, int32_t* seq_lens_out, const int32_t* n_keep,
                 int keep_bias, const uint8_t* accepted_mask, int zero_rejected,
                 int64_t max_seq_len_host, int32_t* overflow_flag, int32_t* ticket, int rows_ready_early,
                 void* stream, const sd_exchange* exchange, int32_t* exchange_out,
                 const int32_t* exchange_packed, int32_t* seq_ready);

/*
 * Single-layer form with the layer record passed by value and a uniform current length
 * `base_len` (every sequence holds base_len rows): the building block of the functional drop-ins
 *   kv_append(base_k, base_v, new_k, new_v)                       reference.py:59-93
 *   kv_append_with_mask(base_k, ..., accepted_mask, accept_len)   reference.py:96-159
 * whose outputs are views [B,H,base_len+R,D] of the arena.  Same per-row semantics as sd_kv_append.
 */
int sd_kv_append_layer(void* arena_k, void* arena_v, const void* new_k, const void* new_v, int B, int H,
                       int R, int64_t row_bytes, int64_t capacity, int64_t arena_stride_b,
                       int64_t arena_stride_h, int64_t new_stride_b, int64_t new_stride_h,
                       int64_t new_stride_r, int64_t base_len, const int32_t* n_keep, int keep_bias,
                       const uint8_t* accepted_mask, int zero_rejected, int32_t* ticket, void* stream);

/* ------------------------------------------------------------------------------------------
 * The whole verify step in ONE launch: fused verify (greedy) + kv_append of all layers.
 *
 * Same results as sd_verify_prefix(rows = K + 1) followed by sd_kv_append(n_keep = accept_len, keep_bias = 1,
 * seq_lens in place), i.e. the reference's per-sequence accept / bonus / append block
 * (src/specdec/core/pipeline.py:2958-3620 for the arithmetic part; rollback semantics
 * src/kernels/reference.py:133-157: kept rows compacted after the valid ones, rejected slots zero).
 *
 * Default layout ("lagged append"): every verify CTA reduces one chunk of one logits row AND carries one piece of
 * the append of the sequence `lag` places earlier in the batch -- requested while its own chunk is in flight,
 * stored after its partial result is published -- so the append traffic streams under the verify phase instead
 * of behind it; a short tail of CTAs carries the pieces of the last `lag` sequences.  The finalizer of a
 * sequence publishes ONE 8-byte word {launch tag, accept_len + 1, old length} in step_state and stores the new
 * length; nothing is ever re-armed (words of earlier launches carry another tag).
 * DISPATCH-ORDER REQUIREMENT: a CTA only waits for results of CTAs with a lower block index; the kernel relies on
 * the hardware dispatching the blocks of a 1-D grid in index order (true of every CUDA GPU to date, not promised
 * by the programming model).  All waits are bounded: see sd_device_error.
 * Requirements: greedy acceptance, 16-byte aligned unmasked KV rows, the bulk-chunk verify variant (the default);
 * otherwise SD_ERR_UNSUPPORTED: issue the two launches.
 * With `exchange`, the records are pushed by the verify finalizers, each as soon as its sequence is verified
 * (exchange_push_in_append == 0, the faster form), or by the first tail CTA once all of them are final; in both
 * forms that CTA gathers every rank's records into exchange_out and advances the exchange epoch -- only after every
 * finalizer of the launch has read it.
 */
typedef struct sd_step_args {
    /* fused verify: see sd_verify_prefix */
    const void* logits;
    int32_t logit_dtype, id_dtype;
    int64_t stride_b, stride_r, id_stride_b;
    const void* draft_ids;
    int32_t* accept_len;
    uint8_t* accepted_mask; /* nullable */
    int32_t* target_ids;    /* nullable */
    int32_t* bonus;         /* nullable */
    int32_t* packed;        /* nullable (required with exchange) */
    void* workspace;
    size_t workspace_bytes;
    int32_t B, rows, K, V;
    /* kv_append of the rows == K + 1 offered rows: see sd_kv_append */
    const sd_kv_layer* layers;
    int32_t n_layers, H, zero_rejected, exchange_push_in_append;
    int64_t row_bytes, capacity, arena_stride_b, arena_stride_h, new_stride_b, new_stride_h, new_stride_r;
    int64_t max_seq_len_host;
    int32_t* seq_lens;      /* updated in place */
    int32_t* overflow_flag; /* nullable */
    int32_t* ticket;        /* one int32, zero before the first launch (the kernel re-zeroes it) */
    int32_t* seq_ready;     /* mode 1 only: [B] int32, zero before the launch, re-zeroed by it */
    /* multi-GPU exchange (nullable) */
    const sd_exchange* exchange;
    int32_t* exchange_out;
    /* mode 0 (default): sd_step_state_bytes(B) bytes, 8-byte aligned, zero before the FIRST launch and never touched
     * by the caller afterwards: u64 sequence words[B], u32 launch epoch, u32 reserved.  One state per stream of
     * steps (launches that may overlap need their own). */
    int32_t* step_state;
    /* Optional (mode 0): uniform layers.  When either stride is non-zero, layer l's pointers are
     * layer0.{arena_k, arena_v} + l * layer_stride_arena and layer0.{new_k, new_v} + l * layer_stride_new (bytes,
     * multiples of 16) -- the layout of a KV cache allocated as ONE [layers, B, H, C, D] tensor -- and the device
     * table `layers` is not read by the append part: a row request then depends on nothing but the thread index
     * (no table lookup in front of it).  `layers` must still describe the same pointers. */
    sd_kv_layer layer0;
    int64_t layer_stride_arena, layer_stride_new;
} sd_step_args;

size_t sd_step_state_bytes(int B);
int sd_verify_append(const sd_step_args* args, void* stream);
/* Layout of sd_verify_append: mode 0 (default) = lagged append, `lag` sequences (0 = heuristic: about 1.5 waves
 * of resident CTAs); mode 1 = round-1 layout kept for A/B evidence: all append CTAs behind all verify CTAs, each
 * thread waiting for its sequence's seq_ready flag -- the two phases serialise (34.6 us vs the sum of its parts
 * 18.9 + 14.9 us at B = 64, K = 4, V = 128256, 28 layers). */
int sd_verify_append_set_mode(int mode, int lag);

/* Variant selector for sd_kv_append (evidence runs): 0 = heuristic, 1 = 128-bit LDG/STG kernel,
 * 2 = cp.async.bulk (TMA 1-D) shared-memory staged kernel, 3 = column kernel (one thread per 16-byte
 * column, all R rows of K and V in flight). */
int sd_kv_set_variant(int variant);

/* ------------------------------------------------------------------------------------------
 * Kernel 3: (K+1)-query causal decode attention for the verify pass (GQA aware, split-KV).
 *
 * Replaces the target model's attention inner loop during verification, which the reference runs as
 * K sequential single-query SDPA calls inside `transformers`
 * (src/specdec/models/hf_wrappers.py:98 attn_implementation="sdpa";
 *  src/scheduler/speculative_scheduler.py:192-199 K greedy steps), with ONE pass over the KV cache.
 *
 * q        [B, Hq, R, D]   contiguous, dtype f16/bf16, D = 64 or 128; R = K+1 new positions
 *                          (R <= 32 for D = 128, R <= 64 for D = 64; query heads sharing a KV head are
 *                          processed together while (Hq/Hkv) * R fits the accumulator budget, else in
 *                          head groups that re-read the KV tile from L2)
 * k_arena  [B, Hkv, C, D]  byte strides kv_stride_b, kv_stride_h, row stride D*elt; the R new rows
 *                          are already appended: rows [0, seq_lens[b] + R) are valid
 * v_arena  same layout
 * seq_lens [B] int32       rows valid BEFORE the R new ones (query i sees rows [0, seq_lens[b]+i])
 * out      [B, Hq, R, D]   same dtype as q
 * workspace sd_verify_attention_workspace_bytes(...) bytes (split-KV partials), no init needed
 * scale    softmax scale (1/sqrt(D) when <= 0)
 * n_splits 0 = heuristic (wave-quantised cost model on max_seq_len_host, bounded by the workspace);
 *          > 1 needs sd_verify_attention_workspace_bytes(..., n_splits) bytes;
 *          SD_ATTN_BALANCED = the persistent form described below (its own, initialised workspace)
 * max_seq_len_host  upper bound of seq_lens known on the host (-1 = derive from kv_stride_h); the split heuristic
 *          uses it -- the kernel reads the true lengths from seq_lens -- and max_seq_len_host + R > capacity is
 *          refused (SD_ERR_CAPACITY)
 * capacity  rows per (b, kv head) slab of the arena, C (<= 0: kv_stride_h / (D * elt)).  On the device
 *          seq_lens[b] + R is clamped to it (a sequence can sit AT capacity after an append that overflowed): rows of
 *          the next head are never read.  overflow_flag (nullable, int32) is set when a length was clamped or negative;
 *          that sequence's output is then not meaningful.
 * Returns SD_ERR_UNSUPPORTED for dtype f32, D not in {64,128} or R beyond the limits above.
 * Algorithmic traffic: 2 * B * Hkv * (mean len + R) * D * elt bytes of KV read once per launch.
 * One launch (+ one small merge launch when n_splits > 1).
 */
size_t sd_verify_attention_workspace_bytes(int B, int Hq, int R, int D, int max_splits);

/* Balanced (persistent) form of kernel 3: n_splits = SD_ATTN_BALANCED.  The grid is exactly the resident CTAs and the
 * key tiles of all (sequence, KV head) items are cut into equal spans, one per CTA; an item cut by a span boundary is
 * finished by whichever of its CTAs arrives last (per-item counter; nothing waits).  Same results as the standard
 * form up to fp32 summation order.  Worth it when B * Hkv is not close to a multiple of the resident CTAs (512 items on
 * 296 slots).  B <= 512.  The workspace (sd_verify_attention_balanced_workspace_bytes) holds the counters and the
 * partial slots and must be zeroed ONCE with sd_verify_attention_balanced_workspace_init; every launch leaves the
 * counters zero again, so it is graph-replayable.  One workspace serves one launch at a time (launches on one stream,
 * PDL included, are fine); re-initialise it after an aborted launch. */
#define SD_ATTN_BALANCED (-1)
size_t sd_verify_attention_balanced_workspace_bytes(int B, int Hq, int Hkv, int R, int D);
int sd_verify_attention_balanced_workspace_init(void* workspace, size_t workspace_bytes, void* stream);
int sd_verify_attention(const void* q, const void* k_arena, const void* v_arena,
                        const int32_t* seq_lens, void* out, void* workspace,
                        size_t workspace_bytes, int dtype, int B, int Hq, int Hkv, int R, int D,
                        int64_t kv_stride_b, int64_t kv_stride_h, int64_t max_seq_len_host,
                        int64_t capacity, int32_t* overflow_flag, float scale, int n_splits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Kernel 4: batched bonus-token sampler.
 *
 * Replaces  sample_bonus_token_from_logits(logits[V], temperature, do_sample, top_p, top_k)
 *           reference: src/specdec/core/pipeline.py:48-147, called once per sequence at
 *           :3140-3163 / :3174-3206 / :3341-3362 on logits row accept_len[b].
 * for the whole batch in one launch (one CTA per sequence, no sort of the vocabulary: with 0 < top_k <= 1024 two
 * passes over the row collect a superset of the top-k, which is finished in shared memory; otherwise radix select by
 * count for top-k and by probability mass for top-p; fixed-point masses -> deterministic).
 *
 * logits     [B, rows, V], element strides (stride_b, stride_r, 1)
 * row_index  [B] int32 (nullable): row = clamp(row_index[b] + row_bias, 0, rows-1); NULL -> row_const.
 *            Typically the accept_len written by sd_verify_* on the same stream (row_bias = 0).
 * uniform    [B] float in [0,1) (nullable).  NULL = greedy branch (pipeline.py:135-137): argmax.
 *            Otherwise token = first index t (index order) whose cumulative probability over the
 *            filtered distribution exceeds uniform[b]: the distribution is the reference's
 *            (temperature :91-92, top-k :95-102, top-p :105-121 with "keep at least one"); its
 *            torch.multinomial draw depends on torch's RNG stream and cannot be reproduced.
 * top_k      <= 0 = off;  top_p >= 1 (or <= 0) = off;  temperature > 0 when sampling
 * kept_count [B] int32 (nullable): size of the final candidate set
 */
int sd_sample_bonus(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                    const int32_t* row_index, int row_bias, int row_const, const float* uniform,
                    float temperature, int top_k, float top_p, int32_t* out_tokens,
                    int32_t* kept_count, int B, int rows, int V, void* stream);

/* Bonus sampler: 1 (default) = requests with 0 < top_k <= 1024 take the two-pass candidate path (the k-th largest
 * per-thread maximum bounds the threshold, one more pass collects the superset, sorts in shared memory finish the
 * job) and fall back to the general radix-select path only when the superset does not fit; 0 = always the general
 * path.  Both give the same token (tests/test_sample_gpu.py compares them). */
int sd_sample_set_fast_path(int enabled);

/* Tile configuration of the attention kernel (evidence runs): 0 = heuristic, 1 = 16-key tiles x 3
 * stages (two CTAs per SM at D = 128), 2 = 32-key tiles x 3 (4 at D = 64) stages, 3 = 32-key x 2; D = 64 only:
 * 4 / 5 / 6 = 16-key tiles x 3 / 4 / 6 stages (measured slower than the default, csrc/sd_attn.cu). */
int sd_verify_attention_set_tuning(int config);

#ifdef __cplusplus
}
#endif
#endif /* SD_VERIFY_H_ */
