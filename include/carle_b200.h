/*
 * carle_b200.h — C ABI of the B200-native CARLE environment step.
 *
 * The reference (riveSunder/carle) has no FFI: its hot path is the Python class
 * `CARLE` in carle/env.py, running torch ops on a float32 [N,1,H,W] tensor.  This
 * library sits UNDER a drop-in `CARLE` class (carle_b200/env.py); each entry point
 * below names the reference lines whose work it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the
 *     handle's device unless the name says `host`; `stream` is a cudaStream_t
 *     passed as void* (0 = legacy default stream).
 *   - every function returns 0 on success, a negative CARLE_E* code on failure;
 *     carle_last_error() returns a thread-local message.  Nothing throws.
 *   - no allocation on the step path: all buffers are caller-owned; the handle
 *     owns only a small scratch allocated in carle_create().
 *   - a handle is bound to one device; calls on one handle are not thread-safe,
 *     distinct handles are independent.
 *
 * Packed layouts (all little-endian uint32 words)
 *   state : [N][H][WPR], WPR = ceil(W/32); bit b of word w of a row = column 32*w+b;
 *           bits at columns >= W are always 0.
 *   action: [K][B][AW][AWPR], B = 1 (broadcast) or N.  Element [r][c] of the window
 *           toggles universe cell [row0 + r][col0 + c] (carle/env.py:172-182; the
 *           reference's (aw, ah) axis order is kept).  Packed rows are ALIGNED TO THE
 *           UNIVERSE'S WORD GRID so a toggle is one XOR: AW0 = col0/32 is the first
 *           universe word the window touches, AWPR = (col0+AH-1)/32 - AW0 + 1, and bit b
 *           of word j of row r is the toggle of universe column 32*(AW0+j)+b, i.e.
 *           window column 32*(AW0+j)+b-col0 (bits outside the window are 0).
 */
#ifndef CARLE_B200_H
#define CARLE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CARLE_API __attribute__((visibility("default")))
#else
#define CARLE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct carle_ctx* carle_handle_t;

enum {
    CARLE_OK = 0,
    CARLE_EINVAL = -1,      /* bad argument / geometry the reference rejects too   */
    CARLE_ECUDA = -2,       /* a CUDA runtime call failed (message has the detail) */
    CARLE_ENODEV = -3,      /* no usable CUDA device                               */
    CARLE_ERULE = -4        /* empty birth or survive set (reference: TypeError)   */
};

/* element types of unpacked cell / action buffers */
enum {
    CARLE_F32 = 0,          /* float32, any non-zero value is "on" (env.py:182)    */
    CARLE_U8 = 1,           /* uint8 / bool                                        */
    CARLE_PACKED = 2        /* already packed uint32 words (actions only)          */
};

/* indices into the per-step flags pair written by carle_pack_action */
enum { CARLE_FLAG_NOT_ALL_ONES = 0, CARLE_FLAG_ANY_TOGGLE = 1 };
/* indices into the device counters block (int64[8]) updated by carle_step* */
enum { CARLE_CNT_STEP_NUMBER = 0, CARLE_CNT_STEPS_SINCE_ACTION = 1,
       CARLE_CNT_RESETS = 2, CARLE_CNT_GENERATIONS = 3,
       CARLE_CNT_LAST_NOT_ALL_ONES = 4,    /* 0 <=> the last step cleared the universe          */
       CARLE_CNT_LAST_ANY_TOGGLE = 5,
       CARLE_CNT_LAST_RESET_COND = 6 };    /* 1 <=> the last step's reset condition held (also
                                              when the clear was deferred, carle_step_args)     */
/* columns of the reductions block (int64[N][4]) */
enum { CARLE_RED_LIVE = 0, CARLE_RED_SH = 1, CARLE_RED_SW = 2, CARLE_RED_WINDOW_LIVE = 3 };

CARLE_API int carle_version(void);
CARLE_API const char* carle_last_error(void);

/* Replaces CARLE.__init__ geometry + set_action_padding (carle/env.py:17-59,
 * 119-132).  Same arithmetic, same rejections: the padded action must come out
 * exactly H x W, which fails for odd or non-square grids.  action_height /
 * action_width are the ctor kwargs; the adjusted values are readable through
 * carle_geometry(). */
CARLE_API int carle_create(carle_handle_t* out, int device, int64_t instances, int height,
                 int width, int action_height, int action_width);
CARLE_API int carle_destroy(carle_handle_t h);

/* geo[0..7] = row0, col0, aw (window rows), ah (window cols), WPR, AWPR,
 *             kernel family (0 generic, 1 warp-resident, 2 tiled), AW0 */
CARLE_API int carle_geometry(carle_handle_t h, int32_t geo[8]);

/* Replaces the per-step `elem == my_neighborhood` rule evaluation set-up
 * (carle/env.py:221-224): bit k of each mask <=> neighbour count k is in the rule,
 * k = 0..8.  An empty mask returns CARLE_ERULE (the reference raises TypeError from
 * reduce() on an empty list).  Callers re-derive the masks from env.birth /
 * env.survive before every step because the reference lets callers assign those
 * lists directly (carle/train_mcl.py:56-57). */
CARLE_API int carle_set_rule(carle_handle_t h, uint32_t birth_mask, uint32_t survive_mask);

/* float32/uint8 cells [N][H][W]  <->  packed state.  (The reference keeps the
 * float tensor as its state, carle/env.py:136; these are the boundary converters
 * behind the `universe` property and the float32 observation.) */
CARLE_API int carle_pack_state(carle_handle_t h, const void* cells, int dtype,
                     uint32_t* packed, void* stream);
CARLE_API int carle_unpack_state(carle_handle_t h, const uint32_t* packed, void* cells,
                       int dtype, void* stream);

/* Replaces torch.sum(action) / torch.mean(action) == 1.0 (carle/env.py:191, 208)
 * and the ZeroPad2d + logical_xor operand preparation (carle/env.py:179-182).
 * action: [steps][batch][AW][AH] of dtype; batch is 1 or N.
 * packed_action: [steps][batch][AW][AWPR] out.
 * flags: int32 [steps][2] in/out, MUST BE ZERO ON ENTRY (the kernel only ever stores
 * 1s); [.][0] != 0 <=> some element != 1.0 (no master reset), [.][1] != 0 <=> some
 * element != 0 (an action was taken).  carle_step / carle_step_many consume the flags
 * and re-zero them (the last block to retire does it), so a buffer zeroed once at
 * allocation can be reused every step with no memset in between. */
CARLE_API int carle_pack_action(carle_handle_t h, const void* action, int dtype,
                      int64_t batch, int64_t steps, uint32_t* packed_action,
                      int32_t* flags, void* stream);

/* The same packing ON THE HOST, for actions that live in host memory (the reference's agents
 * produce them there, carle/agents.py:35-42, and env.py:158-160 ships 4 bytes per toggle to the
 * device): action_host [batch][AW][AH] of dtype (CARLE_F32 / CARLE_U8) -> packed_host
 * [batch][AW][AWPR], both HOST pointers (pinned memory for an asynchronous copy afterwards),
 * packed by `threads` host threads (a persistent pool inside the library).  No device is involved:
 * the geometry is passed explicitly -- aw, ah, awpr = geo[2], geo[3], geo[5] of carle_geometry and
 * bit0 = col0 - 32 * AW0 = geo[1] - 32 * geo[7], the bit of a row's first word that holds window
 * column 0.
 * flags[0] != 0 <=> some element != 1.0, flags[1] != 0 <=> some element != 0,
 * flags[2] != 0 <=> some element is neither 0 nor 1: the packed words then cannot carry the
 * reference's mean / sum predicates (uint8: "every element == 1") and the caller must ship the
 * unpacked action instead. */
CARLE_API int carle_pack_action_host(int32_t aw, int32_t ah, int32_t awpr, int32_t bit0,
                                     const void* action_host, int dtype, int64_t batch,
                                     uint32_t* packed_host, int32_t* flags, int32_t threads);

/* carle_pack_action_host that also ships what it packs (env.py:158-160, the copy of the action to the
 * device): the packed words are copied from packed_host (pinned) to packed_device ([batch][AW][AWPR] on
 * CUDA device `device`) with cudaMemcpyAsync on `stream`, in up to eight slices, each enqueued by the host
 * thread that finishes packing it -- when the packing ends only the last slice's copy is still to run.
 * All copies have been enqueued when the call returns (also when flags[2] says the words are of no use).
 * CARLE_ENODEV without a device, CARLE_ECUDA if a copy could not be enqueued. */
CARLE_API int carle_pack_action_host_copy(int32_t aw, int32_t ah, int32_t awpr, int32_t bit0,
                                          const void* action_host, int dtype, int64_t batch,
                                          uint32_t* packed_host, int32_t* flags, int32_t threads,
                                          uint32_t* packed_device, int32_t device, void* stream);

/* Replaces CARLE.step (carle/env.py:188-242): action XOR -> master reset if the
 * whole action tensor was ones -> one Life-like generation with toroidal wrap.
 * state_in may equal state_out only for the warp-resident family (geo[6] == 1).
 * packed_action / flags as written by carle_pack_action (NULL action = no
 * toggles; NULL flags = "no reset, no action").  counters: int64[8] device block
 * or NULL.  reductions: int64 [N][4] or NULL (live, sum i*m*u, sum j*m*u, live
 * inside window) of the NEW state — the SpeedDetector sums of carle/mcl.py:773-779,
 * fused. */
CARLE_API int carle_step(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
               const uint32_t* packed_action, int64_t action_batch,
               int32_t* flags, int64_t* counters, int64_t* reductions,
               void* stream);

/* The whole of CARLE.step (carle/env.py:188-242) from the caller's UNPACKED action in one
 * call.  action: [batch][AW][AH] float32 or uint8, batch = 1 or N.  For the batched
 * shapes of BASELINE.json (64x64 and 128x128 grids with a 32x32 window, 256x256 with
 * 64x64) this is ONE kernel: each warp ballots its instance's action while its state loads are
 * in flight, and the batch-wide master-reset test is resolved by the last block to
 * retire (it clears the freshly written state in the rare all-ones case).  Otherwise it
 * is carle_pack_action + carle_step on a scratch buffer owned by the handle (allocated
 * on first use).  Same outputs as carle_step. */
CARLE_API int carle_step_action(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                                const void* action, int dtype, int64_t action_batch,
                                int64_t* counters, int64_t* reductions, void* stream);

/* carle_step_action with the per-step outputs a drop-in `CARLE.step` has to produce FUSED INTO
 * THE STEP KERNEL (carle/env.py:236-240), and the switch an instance-sharded batch needs:
 *   reward_zero   float32 [N] or NULL: filled with 0.0f by the step kernel -- the reference
 *                 returns a fresh all-zero reward tensor every step (env.py:238); a caller can hand
 *                 in uninitialised memory instead of paying a fill launch.
 *   obs           [N][H][W] float32 (CARLE_F32) / uint8 (CARLE_U8) or NULL: the NEW state unpacked,
 *                 written by the step kernel from the registers that hold the new rows (the
 *                 reference's observation is its float32 state, env.py:184-186, 236) -- no second
 *                 launch and no re-read of the packed state.
 *   defer_reset   non-zero: the batch is ONE SHARD of a larger batch; the master reset test of the
 *                 reference spans the whole batch (env.py:208), so this call never clears: it
 *                 reports whether the shard's own condition held in counters[CARLE_CNT_LAST_RESET_COND]
 *                 and the caller combines the shards and calls carle_apply_reset.
 * action may be CARLE_F32 / CARLE_U8 ([batch][AW][AH]), CARLE_PACKED ([batch][AW][AWPR]) or NULL.
 * Geometries / dtypes without a one-launch kernel get the same results from extra launches.
 * Set struct_size = sizeof(carle_step_args); fields beyond it are taken as zero. */
typedef struct carle_step_args {
    uint32_t struct_size;
    int32_t action_dtype;
    const uint32_t* state_in;
    uint32_t* state_out;
    const void* action;
    int64_t action_batch;
    int64_t* counters;
    int64_t* reductions;
    float* reward_zero;
    void* obs;
    int32_t obs_dtype;
    int32_t defer_reset;
    /* SpeedDetector tail (carle/mcl.py:777-795, see carle_speed_tail) as part of the step: needs
     * `reductions`; the step then also produces, from the sums of the NEW state,
     *   speed_com_next[2][N]  this step's centres of mass (speed_com_prev: the previous step's,
     *                         read only -- two distinct buffers, swapped by the caller),
     *   speed_velocity[2][N]  (optional), speed_out[1], speed_sumsq[1] (float64, optional),
     *   reward_zero[i] = 0 + speed instead of 0,
     * with the wrapper's "first step records the centre of mass only" decided by the device flag
     * speed_primed (int32, read and then set).  For the batched shapes all of it happens inside
     * the one step kernel; otherwise carle_speed_tail is launched behind the step. */
    const float* speed_com_prev;
    float* speed_com_next;
    float* speed_velocity;
    float* speed_out;
    double* speed_sumsq;
    int32_t* speed_primed;
} carle_step_args;
CARLE_API int carle_step_ex(carle_handle_t h, const carle_step_args* args, void* stream);

/* The clear of a master reset decided outside the step (instance-sharded batches: env.py:208
 * evaluated over all shards).  *decision (device int32) != 0: zero `state`, `obs` (optional,
 * dtype as above), `reductions` (optional) and set the counters as reset() does (env.py:142-145);
 * *decision == 0: nothing happens.  One launch, no host synchronisation. */
CARLE_API int carle_apply_reset(carle_handle_t h, const int32_t* decision, uint32_t* state, void* obs,
                                int obs_dtype, int64_t* reductions, int64_t* counters, void* stream);

/* K generations in one call == K x carle_step with actions[k], flags[k]; the
 * warp-resident family keeps the state in registers across all K generations
 * (one HBM round trip).  reductions: int64 [K][N][4] or NULL.  scratch: a second
 * state-sized buffer, required by the generic family when K > 1 (may be NULL for
 * the warp-resident family). */
CARLE_API int carle_step_many(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                    uint32_t* scratch, const uint32_t* packed_actions,
                    int64_t action_batch, int64_t steps, int32_t* flags,
                    int64_t* counters, int64_t* reductions, void* stream);

/* ---- one giant grid split into row bands over several GPUs (BASELINE config 5) ----------
 * The reference cannot do this at all (one dense tensor on one device, carle/env.py:136);
 * the semantics are those of CARLE.step on the whole H x W torus.  Each rank owns rows
 * [band_row0, band_row0 + band_rows) and keeps a local packed buffer of
 * (band_rows + 2*halo) rows x W/32 words: [halo rows above | band | halo rows below].
 * carle_band_step advances the band `generations` <= halo generations in one launch
 * (temporal blocking inside 256x256 register tiles) reading `in` and writing the band rows of
 * `out`; the rows that are the neighbours' halos are ALSO stored directly into the
 * neighbouring ranks' `out` buffers (peer_up_out / peer_dn_out: device pointers of those
 * buffers mapped into this process with carle_ipc_open; NVLink P2P stores from inside the
 * compute kernel).  Consecutive calls must be separated by a barrier with the two neighbours:
 * either the caller's (sync_* NULL), or -- no host, no NCCL -- the kernel's own: sync_local is this
 * rank's int32[4] (cudaMalloc'ed, zeroed, mapped into the neighbours), sync_up / sync_dn the
 * neighbours' int32[4] mapped here.  A launch then spins (ld.acquire.sys) until both neighbours
 * have finished as many temporal blocks as this rank, and its last CTA publishes the new count in
 * the neighbours' words (st.release.sys over NVLink) behind system-scope fences of every CTA.
 * packed_actions: [generations][AW][AWPR] for the whole-grid window (every rank passes the same
 * action) or NULL; flags as in carle_step. */
CARLE_API int carle_band_create(carle_handle_t* out, int device, int height, int width,
                                int action_height, int action_width, int band_row0,
                                int band_rows, int halo);
CARLE_API int carle_band_step(carle_handle_t h, const uint32_t* in, uint32_t* out,
                              uint32_t* peer_up_out, uint32_t* peer_dn_out, int generations,
                              const uint32_t* packed_actions, int32_t* flags,
                              int64_t* counters, int32_t* sync_local, int32_t* sync_up,
                              int32_t* sync_dn, void* stream);
/* Copy this band's edge rows of `buf` into the neighbours' halo rows (initial state / after
 * the caller rewrote the band); peers may be NULL. */
CARLE_API int carle_band_push_halos(carle_handle_t h, const uint32_t* buf, uint32_t* peer_up_buf,
                                    uint32_t* peer_dn_buf, void* stream);
/* cudaMalloc / cudaFree for the band buffers: CUDA IPC exports whole allocations, so the
 * buffers must not be sub-allocated by a caching allocator. */
CARLE_API int carle_dev_alloc(int device, uint64_t bytes, void** out);
CARLE_API int carle_dev_free(int device, void* ptr);
/* CUDA IPC plumbing for the peer buffers (64-byte handles travel over torch.distributed). */
CARLE_API int carle_ipc_export(const void* dev_ptr, unsigned char handle_out[64]);
CARLE_API int carle_ipc_open(const unsigned char handle[64], void** dev_ptr_out);
CARLE_API int carle_ipc_close(void* dev_ptr);

/* Replaces RandomAgent.forward (carle/agents.py:35-42: 1.0 * (rand <= toggle_rate)) on the
 * device and without the float tensor: Bernoulli(toggle_rate) toggles (16-bit resolution) for
 * every window cell of `batch` entries, written in the packed action layout.  Stateless
 * Philox4x32-10 keyed by (seed; entry, row, chunk, step): the same arguments always give the
 * same action.  The distribution matches the reference's agent; the bit stream is of course
 * not torch's CPU generator. */
CARLE_API int carle_random_action(carle_handle_t h, uint64_t seed, uint32_t step, double toggle_rate,
                                  int64_t batch, uint32_t* packed_action, void* stream);
/* CARLE.step with the random agent fused in: state' = step(state, random_action(seed, step,
 * toggle_rate)) in ONE kernel for the batched shapes (the toggles are drawn inside the step
 * kernel from the same Philox streams carle_random_action uses, so both paths agree bit for
 * bit); other geometries generate into `packed_scratch` ([batch][AW][AWPR], required then) and
 * run flags + step. */
CARLE_API int carle_step_random(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                                uint64_t seed, uint32_t step, double toggle_rate,
                                int64_t action_batch, uint32_t* packed_scratch,
                                int64_t* counters, int64_t* reductions, void* stream);
/* packed action -> float32 [batch][AW][AH] (to hand a device-generated action to code that
 * expects the reference's format). */
CARLE_API int carle_unpack_action(carle_handle_t h, const uint32_t* packed_action, int64_t batch,
                                  float* action, void* stream);

/* Replaces CARLE.apply_action used on its own (carle/env.py:150-182): toggle the
 * window cells in place, no generation. */
CARLE_API int carle_apply_action(carle_handle_t h, uint32_t* state,
                                 const uint32_t* packed_action, int64_t action_batch,
                                 void* stream);

/* Standalone grid reductions on a packed state (carle/mcl.py:773-779 SpeedDetector
 * numerators; :832 PufferDetector live count is the column-0 sum).
 * out: int64 [N][4] as above. */
CARLE_API int carle_reduce(carle_handle_t h, const uint32_t* state, int64_t* out, void* stream);

/* Fixed-mask weighted popcount (carle/mcl.py:222-223 CornerBonus):
 * out[n] = popcount(state[n] & plus_mask) - popcount(state[n] & minus_mask);
 * masks are packed [H][WPR], either may be NULL.  out: int64 [N]. */
CARLE_API int carle_masked_count(carle_handle_t h, const uint32_t* state,
                       const uint32_t* plus_mask, const uint32_t* minus_mask,
                       int64_t* out, void* stream);

/* Per-entry action popcount (carle/mcl.py:102-103 ParsimonyBonus denominator,
 * exact for 0/1 actions).  out: int64 [batch]. */
CARLE_API int carle_action_count(carle_handle_t h, const uint32_t* packed_action,
                       int64_t batch, int64_t* out, void* stream);

/* SpeedDetector tail (carle/mcl.py:777-795) in one launch, from the fused sums of a step:
 *   com = (Sh, Sw) / (live + 1e-7)         float32 [2][N], IN: previous step's, OUT: this step's
 *   velocity = com_previous - com          float32 [2][N] (optional)
 *   speed = || velocity ||_2 over the whole batch -> speed_out[0]
 *   reward[i] += speed                     float32 [N] (optional)
 *   sumsq_out[0] = sum of velocity^2       float64 (optional): the part of the norm one shard of
 *                                          an instance-sharded batch contributes
 * have_previous = 0 on the wrapper's first step (mcl.py:784: only the centre of mass is
 * recorded; velocity, speed, reward and sumsq are left untouched).  primed (device int32,
 * optional) moves that decision onto the device: when given, *primed != 0 is used instead of
 * have_previous and the launch sets *primed = 1 -- the same call then serves the first step and
 * every later one, which is what a CUDA-graph replay of a rollout needs. */
CARLE_API int carle_speed_tail(carle_handle_t h, const int64_t* reductions, float* center_of_mass,
                               int have_previous, float* velocity_out, float* speed_out,
                               float* reward, double* sumsq_out, int32_t* primed, void* stream);

/* PufferDetector tail (carle/mcl.py:828-850) without the reference's per-step host round trip:
 * total live cells of the step = sum of reductions[.][CARLE_RED_LIVE] (exact int64; the reference
 * sums float32), then the wrapper's sliding window ON THE DEVICE:
 *   no toggle this step (counters[CARLE_CNT_LAST_ANY_TOGGLE] == 0; mcl.py:833 tests sum(action)):
 *       append the total; once the window holds more than growth_threshold entries,
 *       slope = newest - oldest, drop the oldest, and if slope > 0: reward[i] += 1 for every i
 *   else: the window is emptied.
 * ring : int64 [growth_threshold + 1]; state: int64 [8], ZERO AT ALLOCATION, kept by the kernel:
 *   [0] entries in the window, [1] index of the oldest, [2] / [3] scratch (zero between calls),
 *   [4] live total of the last step, [5] 1 when the last step paid the bonus, [6] bonuses paid.
 * reward: float32 [N] (optional). */
CARLE_API int carle_puffer_tail(carle_handle_t h, const int64_t* reductions, const int64_t* counters,
                                int64_t* ring, int64_t* state, int32_t growth_threshold, float* reward,
                                void* stream);

/* MorphoBonus's template match (carle/mcl.py:176-185: F.conv2d(|universe - action|, patterns),
 * no padding, then the maximum and the minimum over patterns and positions per instance) on the
 * packed state.  A pattern is 8x8 with weight +weights[p] on its live cells and -1 on the others
 * (mcl.py:153-159), so a window X scores weights[p] * popcount(X & P) - popcount(X & ~P).
 *   patterns[p]: bit 8*r + c = pattern cell [r][c] (the conv2d weight at [r][c]).
 *   toggles    : optional packed plane(s) [toggle_batch][H][WPR] XOR-ed onto the state as it is
 *                read (the action of the step about to be taken; toggle_batch 1 or N).
 *   out_max / out_min: float32 [N]; exact whenever the weights are integers (15 / 5 for the
 *                reference's gliders), else within float32 rounding of the conv2d sum. */
CARLE_API int carle_morpho_match(carle_handle_t h, const uint32_t* state, const uint32_t* toggles,
                                 int64_t toggle_batch, const uint64_t* patterns, const float* weights,
                                 int32_t n_patterns, float* out_max, float* out_min, void* stream);

/* RLE codec on packed words in HOST memory (carle/env.py:408-464 get_rle's run tokens and line
 * breaks; env.py:260-328 rle_to_grid).  The header lines stay with the caller.
 * encode: rows [height][ceil(width/32)] -> "<count><b|o>...$" tokens, a line break whenever a line
 *   exceeds 69 characters, "!" at the end; byte-identical to what the reference emits with
 *   flags = 0 (it drops the last partial line, env.py:453-455); CARLE_RLE_KEEP_TAIL keeps it.
 *   Returns the length of the text (writes only if it fits `capacity`; call with out = NULL to
 *   size the buffer) or a negative error code.
 * decode: tokens with optional counts, upper or lower case, counts may span line breaks; cells
 *   outside height x width are dropped (the reference raises IndexError for rows). */
enum { CARLE_RLE_KEEP_TAIL = 1 };
CARLE_API int64_t carle_rle_encode_host(const uint32_t* packed_host, int32_t height, int32_t width,
                                        int32_t flags, char* out, int64_t capacity);
CARLE_API int carle_rle_decode_host(const char* text, int64_t length, int32_t height, int32_t width,
                                    uint32_t* packed_host_out);

/* Run-time rule specialisation (no reference equivalent: carle/env.py:221-229 evaluates any
 * rule list with the same torch ops).  Rules other than the four built-in ones are compiled
 * with NVRTC into StaticRule kernels the first time they are stepped on a device (the library
 * dlopens libnvrtc; without it, or with CARLE_JIT=0, the slower run-time-rule kernels run --
 * still on the GPU).  carle_jit_probe() compiles, without loading, the specialised one-launch
 * step kernel of `shape` (1: 64x64 / 32x32 window, 2: 128x128 / 32x32, 3: 256x256 / 64x64, float32
 * actions; 4: the multi-generation 256x256 kernel of carle_step_many; 5: the any-shape kernel;
 * 6: the tiled large-grid / row-band kernel (256-row register tiles; 8: its 128-row variant);
 * 7: the 128x128 step with the device random agent; 9 / 10: the 256x256 strip and the 128x128
 * stream kernel fed packed actions)
 * and reports the CUBIN size: a build-time check that
 * needs no GPU.  CARLE_ECUDA
 * with the NVRTC log in carle_last_error() when the compilation fails. */
CARLE_API int carle_jit_probe(int shape, uint32_t birth_mask, uint32_t survive_mask,
                              int64_t* cubin_bytes);

/* Number of NVRTC-specialised kernels loaded by this process so far (all devices); lets a
 * caller / test see whether a rule runs specialised or on the run-time-rule kernels. */
CARLE_API int carle_jit_loaded(void);

#ifdef __cplusplus
}
#endif
#endif /* CARLE_B200_H */
