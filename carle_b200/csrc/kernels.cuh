// kernels.cuh — sm_100a kernels for the CARLE environment step.
//
//  step_stream_kernel<...>       THE per-step kernel for 64x64 / 128x128: persistent warps, each
//                                owning whole instances; packed state + unpacked float32 / uint8
//                                action staged by TMA bulk copies (mbarrier completion, DEPTH slots
//                                per warp), action ballotted in the kernel, one generation,
//                                optional SpeedDetector sums, fence-free retirement, programmatic
//                                dependent launch.  With T = DeviceRandom the toggles are drawn in
//                                registers (Philox) instead of fetched.  (256x256: strip.cuh.)
//  step_warp_kernel<WPR, Rule>   one warp owns one whole instance (H = W = 32*WPR <= 256)
//                                in registers: lane L holds rows [L*WPR, (L+1)*WPR), each
//                                WPR words.  Horizontal neighbours = funnel shifts inside
//                                the lane (toroidal: word WPR-1 wraps to word 0), vertical
//                                neighbours across lanes = warp shuffles of the row-triple
//                                planes (lane 31 wraps to lane 0).  K generations run
//                                without leaving registers (temporal blocking for the
//                                batched configs) from PRE-PACKED actions; master reset and the
//                                SpeedDetector sums are fused.
//  step_fused_kernel,            earlier one-launch variants (one warp per instance, plain loads;
//  step_random_kernel            device random agent at 256x256), kept for A/B runs and unaligned
//                                action pointers.
//  step_generic_kernel<Rule>     any even square shape (W not a multiple of 32, W > 256):
//                                one thread per word, one generation per launch.
//  pack / unpack / reduce /      boundary converters, standalone reductions, the SpeedDetector
//  speed_tail                    tail.
// The same text is compiled at run time by NVRTC (jit.cu) to specialise any rule: keep it free
// of host-only constructs outside `#if !defined(__CUDACC_RTC__)`.
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "ca_core.cuh"

// A/B switches of the optional parts of the one-launch step kernels (build an alternative library
// with CARLE_NVCC_EXTRA="-DCARLE_FEAT_...=0" and select it with CARLE_B200_LIB; tools/ab_features.sh)
#ifndef CARLE_FEAT_SD
#define CARLE_FEAT_SD 1        // SpeedDetector tail inside the kernel (sd_com)
#endif
#ifndef CARLE_FEAT_OBS
#define CARLE_FEAT_OBS 1       // unpacked observation written by the kernel (obs)
#endif
#ifndef CARLE_FEAT_NONBIN
#define CARLE_FEAT_NONBIN 1    // "some action element is neither 0 nor 1" (reference mean / sum predicates)
#endif

namespace carle {

struct StepParams {
    const uint32_t* in;
    uint32_t* out;
    const uint32_t* act;        // packed actions [K][B][AW][AWPR] or nullptr
    long long act_step_stride;  // words between consecutive steps
    long long act_inst_stride;  // words between instances (0: batch-1 broadcast)
    const void* raw;            // FUSED mode: unpacked action [B][AW][AH] (float32 / uint8), K = 1;
                                // the kernel ballots it itself and detects the master reset
    long long raw_inst_stride;  // elements between instances (0: batch-1 broadcast)
    int raw_u8;                 // element type of `raw`: 0 float32, 1 uint8, 2 packed words
                                // ([B][AW][AWPR], persistent kernels only)
    int* flags;                 // [K][2] or nullptr; consumed and re-zeroed by the step
    long long* counters;        // int64[8] or nullptr
    unsigned int* retire;       // handle-owned block-retirement counter (last-block pattern)
    unsigned long long* retire64;   // same, for the fence-free protocol of the persistent kernels
    long long* red;             // int64 [K][N][4] or nullptr
    uint32_t rand_key0, rand_key1, rand_step, rand_threshold;   // device random agent (Philox key,
                                // environment step, Bernoulli threshold * 65536)
    uint32_t zero;              // always 0, but opaque to the compiler: the TMA kernels fold
                                // (loaded registers & zero) into the refill's byte count so the
                                // bulk copy cannot issue before the slot's LDS reads returned
    int rank_blocked;           // persistent kernels: 1 = warp rank block*W + warp (neighbouring
                                // warps stream neighbouring instances: best for many trips),
                                // 0 = warp*gridDim + block (a partial last trip is spread evenly
                                // over the SMs: best for one or two trips)
    unsigned int* strip_part;   // strip kernel: handle-owned uint64 [N][2] sum accumulators
                                // (zero between launches)
    int reverse;                // persistent one-launch kernels: walk the units from the last to the first.
                                // A rollout ping-pongs two state buffers; walking them in alternate
                                // directions lets a step start on the rows the previous step wrote LAST,
                                // which are the ones still in L2 (the launchers set it from the buffer order)
    int act_evict_first;        // bulk copies of the unpacked action carry an L2 evict-first policy: the
                                // action is read once and must not push the state out of L2
    float* reward_zero;         // float32 [N] or nullptr: zero-filled by the step kernel (the fresh
                                // all-zero reward tensor of carle/env.py:238, without a fill launch)
    void* obs;                  // [N][H][W] float32 / uint8 or nullptr: the NEW state unpacked by the
                                // step kernel itself (strict drop-in observation, no second launch)
    int obs_u8;                 // element type of `obs`: 0 float32, 1 uint8
    // SpeedDetector tail (carle/mcl.py:777-795) fused into the step kernel (sd_com != nullptr; needs
    // red != nullptr and !defer_reset): whoever completes an instance's sums also turns them into the
    // centre of mass / velocity, the squared velocities meet in sd_acc, and the grid's last warp
    // writes speed and the reward column (reward_zero[i] = 0 + speed)
    const float* sd_com_prev;   // float32 [2][N]: the previous step's centres of mass (left intact:
                                // a master reset is only known at the end of the launch, and its
                                // velocities are relative to these)
    float* sd_com;              // float32 [2][N] out: this step's (a different buffer)
    float* sd_vel;              // float32 [2][N] or nullptr
    float* sd_speed;            // float32 [1]
    int* sd_primed;             // device flag: a previous centre of mass exists (set by this launch)
    double* sd_sumsq;           // float64 [1] or nullptr: sum of v^2
    double* sd_acc;             // handle-owned accumulator (zero between launches)
    int defer_reset;            // instance-sharded batches: never clear in this launch; whether THIS
                                // shard's reset condition held goes to counters[6] and the caller
                                // combines the shards (carle_apply_reset)
    long long n;                // instances
    int k;                      // generations in this launch
    int h, w, wpr;              // grid
    int row0, col0, aw, ah;     // window: rows [row0,row0+aw), cols [col0,col0+ah)
    int aw0, awpr;              // first universe word the window touches, words it spans
    uint32_t birth, survive;    // 9-bit rule masks
    ca::RuleMasks masks;        // expanded for the run-time rule path
};

// ---- rule functors --------------------------------------------------------------------
template <uint32_t B, uint32_t S>
struct StaticRule {
    __device__ __forceinline__ explicit StaticRule(const StepParams&) {}
    __device__ __forceinline__ uint32_t operator()(uint32_t x, ca::Sum9 s) const {
        return ca::next_static<B, S>(x, s);
    }
    // next state from the row triples above / of / below the cell (Life: 7 LOP3 instead of 8)
    __device__ __forceinline__ uint32_t from_triples(uint32_t x, ca::Triple a, ca::Triple c,
                                                     ca::Triple b) const {
        return ca::next_static_triples<B, S>(x, a, c, b);
    }
};
struct DynamicRule {
    const ca::RuleMasks& m;      // stays in the kernel-parameter constant bank
    __device__ __forceinline__ explicit DynamicRule(const StepParams& p) : m(p.masks) {}
    __device__ __forceinline__ uint32_t operator()(uint32_t x, ca::Sum9 s) const {
        return ca::next_dynamic(x, s, m);
    }
    __device__ __forceinline__ uint32_t from_triples(uint32_t x, ca::Triple a, ca::Triple c,
                                                     ca::Triple b) const {
        return ca::next_dynamic(x, ca::add3(a, c, b), m);
    }
};

// Packed action rows are stored ALIGNED TO THE UNIVERSE'S WORD GRID: word j of action row
// r holds the toggles of universe columns [32*(aw0+j), 32*(aw0+j)+32) of universe row
// row0+r, so applying an action is one load + one XOR per touched word, no shifting.
__device__ __forceinline__ uint32_t action_bits(const StepParams& p, const uint32_t* act_inst,
                                                int row, int w) {
    const int ar = row - p.row0, j = w - p.aw0;
    if (ar < 0 || ar >= p.aw || j < 0 || j >= p.awpr) return 0u;
    return act_inst[(long long)ar * p.awpr + j];
}

// mask of the window columns inside word w
__device__ __forceinline__ uint32_t window_col_mask(const StepParams& p, int w) {
    int lo = max(p.col0 - 32 * w, 0);
    int hi = min(p.col0 + p.ah - 32 * w, 32);
    if (hi <= lo) return 0u;
    uint32_t m = (hi - lo == 32) ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u);
    return m << lo;
}

// Host bookkeeping (carle/env.py:142-145, 200, 230), read lazily by the host.  Runs in the
// LAST block to retire, i.e. after every block has consumed the step's flags, and then
// re-zeroes those flags so the caller's buffer is ready for the next carle_pack_action
// without a memset node in between.
// counters[6] = 1 when the reset condition of the last generation held (also when the clear itself
// was deferred to the caller, StepParams::defer_reset).
__device__ __forceinline__ bool finish_step(const StepParams& p) {
    bool reset = false, any = false, cond = false;
    {
        long long step_number = 0, since = 0, resets = 0;
        if (p.counters) { step_number = p.counters[0]; since = p.counters[1]; resets = p.counters[2]; }
        for (int g = 0; g < p.k; ++g) {
            cond = p.flags && p.flags[2 * g] == 0;
            reset = cond && !p.defer_reset;
            any = p.flags && p.flags[2 * g + 1] != 0;
            if (!any) since += 1;
            if (reset) { step_number = 0; since = 0; resets += 1; }
            else step_number += 1;
        }
        if (p.counters) {
            p.counters[0] = step_number;
            p.counters[1] = since;
            p.counters[2] = resets;
            p.counters[3] += p.k;
            p.counters[4] = reset ? 0 : 1;      // flags of the last generation, for the host
            p.counters[5] = any ? 1 : 0;
            p.counters[6] = cond ? 1 : 0;
        }
    }
    if (p.flags)
        for (int g = 0; g < 2 * p.k; ++g) p.flags[g] = 0;
    if (p.retire) p.retire[7] = reset ? 1u : 0u;      // "the last step cleared the universe"
    return reset;
}

// Call at the end of a (non-fused) step kernel, by every thread of every block: the last
// block to retire does the bookkeeping and re-zeroes the consumed flags.
__device__ __forceinline__ void retire_block(const StepParams& p) {
    if (!p.retire) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.retire, 1u) == gridDim.x - 1) {
            __threadfence();
            finish_step(p);
            *p.retire = 0u;
        }
    }
}

// Host bookkeeping of ONE fused step whose batch-wide flags are already known (fence-free
// retirement below): same arithmetic as finish_step for k = 1.  `cond`: the reset condition
// held; `reset`: the universe is actually cleared by this launch (cond && !defer_reset).
__device__ __forceinline__ void finish_fused_step(const StepParams& p, bool cond, bool reset, bool any) {
    if (p.retire) p.retire[7] = reset ? 1u : 0u;
    if (!p.counters) return;
    long long step_number = p.counters[0], since = p.counters[1], resets = p.counters[2];
    if (!any) since += 1;
    if (reset) { step_number = 0; since = 0; resets += 1; }
    else step_number += 1;
    p.counters[0] = step_number;
    p.counters[1] = since;
    p.counters[2] = resets;
    p.counters[3] += 1;
    p.counters[4] = reset ? 0 : 1;
    p.counters[5] = any ? 1 : 0;
    p.counters[6] = cond ? 1 : 0;
}

// helpers of the fused kernels' action ingestion
template <typename T> struct IsFloatAction;          // (defined with the action element types below)
__device__ __forceinline__ uint32_t bits_of(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t bits_of(uint8_t v) { return v; }
template <typename T> struct OneBits;
template <> struct OneBits<float> { static constexpr uint32_t value = 0x3F800000u; };
template <> struct OneBits<uint8_t> { static constexpr uint32_t value = 1u; };

// "Some element is neither 0.0 nor 1.0".  The reference tests `torch.sum(action)` and
// `torch.mean(action) == 1.0` (carle/env.py:191, 208); for 0/1-valued actions these equal "some
// toggle is set" / "every toggle is set", which the kernels read off the ballot masks.  They differ
// only when an element is neither 0 nor 1 (a 0/2 checkerboard has mean 1.0; +1 and -1 cancel), so
// the kernels also carry a flag "some element could make the predicates differ" (NonBinary below)
// and the grid's last warp then evaluates the reference's predicates on the action tensor itself
// (resolve_action_mean).  uint8
// actions are this library's extension: "all elements == 1" / "some element != 0".
struct NonBinary {
#if defined(CARLE_NB_LEGACY)
    float acc = 0.f;
    __device__ __forceinline__ void see(float v) { if (CARLE_FEAT_NONBIN) acc += fabsf(fmaf(v, v, -v)); }
    __device__ __forceinline__ void see(float a, float b) { see(a); see(b); }
    __device__ __forceinline__ bool any_lane() const { return acc != 0.f; }    // (NaN != 0 is true)
#else
    // A FILTER for the rare exact path (resolve_action_mean), not the predicate itself: the raw bit
    // patterns are OR-ed -- one three-input LOP3 per TWO values -- and any bit outside 1.0's pattern
    // (0x3F800000) flags the instance: every negative value, every value above 1 (2.0 is
    // 0x40000000), every value with a mantissa, NaN and inf.  What passes unflagged besides 0 and 1.0
    // are the powers of two 2^-1 .. 2^-126, all inside (0, 1): they cannot cancel a sum to zero, and
    // they can make mean == 1.0 without every element being 1.0 only by float32 rounding, which
    // needs 2^24 action elements per step or more -- where the reference's own float32 mean of a
    // binary action stops being exact as well ("reset predicate").
    uint32_t acc = 0u;
    __device__ __forceinline__ void see(float v) {
        if (CARLE_FEAT_NONBIN) acc |= __float_as_uint(v);
    }
    __device__ __forceinline__ void see(float a, float b) {
        if (CARLE_FEAT_NONBIN) acc |= __float_as_uint(a) | __float_as_uint(b);
    }
    __device__ __forceinline__ bool any_lane() const { return (acc & ~0x3F800000u) != 0u; }
#endif
    __device__ __forceinline__ void see(uint8_t) {}
    __device__ __forceinline__ void see(uint8_t, uint8_t) {}
    // N values at once, in pairs
    template <typename T, int N>
    __device__ __forceinline__ void see_all(const T (&v)[N]) {
#pragma unroll
        for (int k = 0; k + 1 < N; k += 2) see(v[k], v[k + 1]);
        if (N & 1) see(v[N - 1]);
    }
};

// Rare path: the reference's own predicates, evaluated by ONE warp over the whole float32 action
// tensor ([B][AW][AH], B = 1 or N as the caller passed it): sum in float64 (exact for any realistic
// input, hence independent of the summation order the reference's float32 sum depends on),
// reset = (float32(sum / count) == 1.0), any = (sum != 0).
static __device__ __noinline__ void resolve_action_mean(const StepParams& p, int lane, bool& reset, bool& any) {
    const long long count = (p.raw_inst_stride ? p.n : 1) * (long long)p.aw * p.ah;
    const float* a = static_cast<const float*>(p.raw);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long i = lane;
    for (; i + 96 < count; i += 128) {
        const float v0 = a[i], v1 = a[i + 32], v2 = a[i + 64], v3 = a[i + 96];
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; i < count; i += 32) s0 += a[i];
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    reset = count > 0 && (float)(s / (double)count) == 1.0f;
    any = s != 0.0;
    if (reset) {
        // The clear that follows overwrites stores of warps that did not fence (only all-ones and
        // non-binary instances fence, see fence_if_all_ones); this path has already spent far longer
        // reading the action than any store stays in flight, and waits once more on top.
        for (int k = 0; k < 64; ++k) __nanosleep(1000);
        __threadfence();
    }
}

// Fence-free retirement of the persistent fused-step kernels (<= 65535 blocks, <= 255 warps).
// The batch-wide flags travel INSIDE the atomics: every warp adds
// (1 | not_one << 8 | any << 16 | non_binary << 24) to the block's shared word, the block's last
// warp adds (1 | not_one << 16 | any << 32 | non_binary << 48) to the handle's 64-bit word, and the
// grid's last block reads the totals off its own atomic's return value -- no flag stores, no
// __threadfence on the common path.  Ordering is only needed for the master reset
// (carle/env.py:208-216), where the last block overwrites every block's output with zeros: a warp
// whose instance saw ONLY 1.0 toggles (or a non-binary value) fences its stores itself
// (fence_if_all_ones), and an all-ones reset happens only if every warp did.
// Returns 0, or (in every lane of the grid's last warp) 1 = no clear, 2 = clear the universe.
template <typename T>
__device__ __forceinline__ int retire_fused(const StepParams& p, unsigned int* s_word, int lane,
                                            int warps_per_block, bool warp_not_one, bool warp_any,
                                            bool warp_nonbin, double* s_sd = nullptr) {
    __syncwarp();
    int last_of_grid = 0;
    unsigned int hi = 0u;                       // bits 16.. of the grid total
    if (lane == 0) {
        const unsigned int mine = 1u | (warp_not_one ? 1u << 8 : 0u) | (warp_any ? 1u << 16 : 0u) |
                                  (warp_nonbin ? 1u << 24 : 0u);
        const unsigned int tot = atomicAdd(s_word, mine) + mine;
        if ((tot & 0xFFu) == (unsigned)warps_per_block) {
            unsigned long long blk = 1ull | (((tot >> 8) & 0xFFu) ? 1ull << 16 : 0ull) |
                                     (((tot >> 16) & 0xFFu) ? 1ull << 32 : 0ull) |
                                     ((tot >> 24) ? 1ull << 48 : 0ull);
            if (s_sd) {
                // fused SpeedDetector tail: the CTA's sum of squared velocities (see speed_warp_done)
                const double cta = *reinterpret_cast<volatile double*>(s_sd);
                if (cta != 0.0) {
                    const double old = atomicAdd(p.sd_acc, cta);
                    blk += (unsigned long long)((unsigned)__double2hiint(old) & p.zero);   // == 0: ordering only
                }
            }
            const unsigned long long g = atomicAdd(p.retire64, blk) + blk;
            if ((g & 0xFFFFull) == gridDim.x) {
                *p.retire64 = 0ull;
                last_of_grid = 1;
                // not_one | any | non_binary totals, folded to three bits
                hi = (((g >> 16) & 0xFFFFull) ? 1u : 0u) | (((g >> 32) & 0xFFFFull) ? 2u : 0u) |
                     ((g >> 48) ? 4u : 0u);
            }
        }
    }
    last_of_grid = __shfl_sync(0xFFFFFFFFu, last_of_grid, 0);
    if (!last_of_grid) return 0;
    hi = __shfl_sync(0xFFFFFFFFu, hi, 0);
    bool cond = (hi & 1u) == 0u, any = (hi & 2u) != 0u;
    if constexpr (IsFloatAction<T>::value) {
        if (hi & 4u) resolve_action_mean(p, lane, cond, any);
    }
    const bool reset = cond && !p.defer_reset;
    if (lane == 0) finish_fused_step(p, cond, reset, any);
    return reset ? 2 : 1;
}

// see retire_fused: called after an instance's results are stored
__device__ __forceinline__ void fence_if_all_ones(bool instance_not_one) {
    if (!instance_not_one) __threadfence();
}

// Retirement of the non-persistent one-launch kernels (step_fused_kernel, step_quad_kernel,
// step_random_kernel: A/B variants and unaligned action pointers): block flags in shared memory,
// grid flags in the handle's scratch (p.flags = retire[2..3], non-binary: retire[6]), fences on
// both levels, the last warp of the grid does the bookkeeping.  Returns 0 / 1 / 2 like retire_fused.
template <typename T>
__device__ __forceinline__ int retire_legacy(const StepParams& p, unsigned int* s_done, int* s_flag,
                                             int lane, int warps_per_block, bool not_one, bool any,
                                             bool nonbin) {
    __syncwarp();
    int last_of_grid = 0;
    if (lane == 0) {
        if (not_one) s_flag[0] = 1;
        if (any) s_flag[1] = 1;
        if (nonbin) s_flag[2] = 1;
        __threadfence_block();
        if (atomicAdd(s_done, 1u) == (unsigned)warps_per_block - 1u) {
            __threadfence_block();
            if (s_flag[0]) p.flags[0] = 1;
            if (s_flag[1]) p.flags[1] = 1;
            if (s_flag[2]) p.retire[6] = 1u;
            __threadfence();
            if (atomicAdd(p.retire, 1u) == gridDim.x - 1) {
                __threadfence();
                last_of_grid = 1;
            }
        }
    }
    last_of_grid = __shfl_sync(0xFFFFFFFFu, last_of_grid, 0);
    if (!last_of_grid) return 0;
    if constexpr (IsFloatAction<T>::value) {
        if (*reinterpret_cast<volatile unsigned int*>(p.retire + 6) != 0u) {
            bool cond = false, some = false;
            resolve_action_mean(p, lane, cond, some);
            __syncwarp();
            if (lane == 0) { p.flags[0] = cond ? 0 : 1; p.flags[1] = some ? 1 : 0; p.retire[6] = 0u; }
            __syncwarp();
        }
    }
    int code = 0;
    if (lane == 0) {
        code = finish_step(p) ? 2 : 1;            // batch-wide master reset known here
        *p.retire = 0u;
    }
    return __shfl_sync(0xFFFFFFFFu, code, 0);
}

// ---- SpeedDetector tail fused into the persistent step kernels ------------------------------------
// one instance, by one lane, exactly once per step: mcl.py:777-787 in float32 exactly as the
// reference computes it; returns this instance's v^2 (0 on the wrapper's first step)
__device__ __forceinline__ double speed_instance(const StepParams& p, long long inst, bool primed,
                                                 float prev_h, float prev_w, uint32_t live,
                                                 unsigned long long sh, unsigned long long sw) {
    const float denom = (float)live + 1e-7f;                                       // mcl.py:777
    const float ch = (float)(long long)sh / denom, cw = (float)(long long)sw / denom;
    double v2 = 0.0;
    if (primed) {
        const float vh = prev_h - ch, vw = prev_w - cw;                            // mcl.py:787
        if (p.sd_vel) { p.sd_vel[inst] = vh; p.sd_vel[p.n + inst] = vw; }
        v2 = (double)vh * vh + (double)vw * vw;
    }
    p.sd_com[inst] = ch;
    p.sd_com[p.n + inst] = cw;
    return v2;
}

// every warp, before it retires: its share of sum v^2 (made visible before the retirement atomics)
// The warp's sum of squared velocities joins the CTA's shared accumulator; the CTA's last warp adds
// the CTA total to the handle's accumulator INSIDE retire_fused, ahead of the retirement atomic and
// ordered before it by a data dependency on the atomic's return value (both are performed at L2).
// No __threadfence: a fence here made every warp wait for all of its state stores to be acknowledged
// before it could retire -- the only fence on the common path of the otherwise fence-free kernels.
__device__ __forceinline__ void speed_warp_done(double* s_sd, int lane, double local) {
    if (lane == 0) {
        if (local != 0.0) atomicAdd(s_sd, local);
        __threadfence_block();                   // (ahead of this warp's arrival on the CTA's counter)
    }
    __syncwarp();
}
__device__ __forceinline__ void speed_grid_done(const StepParams& p, int lane, bool primed, bool cleared) {
    double total = atomicAdd(p.sd_acc, 0.0);      // (read at L2, behind every CTA's contribution)
    if (cleared) {
        double local = 0.0;
        for (long long i = lane; i < p.n; i += 32)
            local += speed_instance(p, i, primed, p.sd_com_prev[i], p.sd_com_prev[p.n + i], 0u, 0ull, 0ull);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, off);
        total = local;
    }
    const float speed = primed ? sqrtf((float)total) : 0.f;                        // mcl.py:789
    if (lane == 0) {
        if (primed) {
            *p.sd_speed = speed;
            if (p.sd_sumsq) *p.sd_sumsq = total;
        }
        *p.sd_acc = 0.0;
        *p.sd_primed = 1;
    }
    if (p.reward_zero) {
        long long i0 = 0;
        if ((reinterpret_cast<unsigned long long>(p.reward_zero) & 15ull) == 0ull) {
            float4* r4 = reinterpret_cast<float4*>(p.reward_zero);
            const long long n4 = p.n >> 2;
#pragma unroll 4
            for (long long i = lane; i < n4; i += 32) r4[i] = make_float4(speed, speed, speed, speed);
            i0 = n4 << 2;
        }
        for (long long i = i0 + lane; i < p.n; i += 32) p.reward_zero[i] = speed;
    }
}

// the rare clear after a master reset, by the grid's last warp
__device__ __forceinline__ void clear_after_reset(const StepParams& p, int lane) {
    const long long words = p.n * (long long)p.h * p.wpr;
    for (long long i = lane; i < words; i += 32) p.out[i] = 0u;
    if (p.red)
        for (long long i = lane; i < p.n * 4; i += 32) p.red[i] = 0;
    __syncwarp();
}

// =========================================================================================
// warp-resident family
// =========================================================================================
// resident CTAs per SM asked of ptxas (register budget 65536 / (128 * CTAs))
constexpr int warp_kernel_min_ctas(int wpr) {
    return wpr <= 4 ? 7 : (wpr <= 6 ? 3 : 2);
}

// ---- building blocks shared by the two warp-resident kernels ---------------------------------
template <int WPR>
__device__ __forceinline__ void load_state(uint32_t (&x)[WPR][WPR], const uint32_t* src) {
    constexpr int WORDS = WPR * WPR;
    if constexpr (WORDS % 4 == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
        for (int i = 0; i < WORDS / 4; ++i) {
            const uint4 v = s4[i];
            (&x[0][0])[4 * i + 0] = v.x; (&x[0][0])[4 * i + 1] = v.y;
            (&x[0][0])[4 * i + 2] = v.z; (&x[0][0])[4 * i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < WORDS; ++i) (&x[0][0])[i] = src[i];
    }
}

template <int WPR>
__device__ __forceinline__ void store_state(const uint32_t (&x)[WPR][WPR], uint32_t* dst) {
    constexpr int WORDS = WPR * WPR;
    if constexpr (WORDS % 4 == 0) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < WORDS / 4; ++i)
            d4[i] = make_uint4((&x[0][0])[4 * i + 0], (&x[0][0])[4 * i + 1],
                               (&x[0][0])[4 * i + 2], (&x[0][0])[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < WORDS; ++i) dst[i] = (&x[0][0])[i];
    }
}

// one generation (carle/env.py:219-229) of the R x W words per lane held by this warp (lane L:
// rows [L*R, (L+1)*R) of a torus 32*R rows high and 32*W columns wide)
template <int R, int W, class Rule>
__device__ __forceinline__ void generation_rw(uint32_t (&x)[R][W], const Rule& rule,
                                              int up_lane, int dn_lane) {
    // row triples of the first and last row feed the neighbouring lanes
    constexpr bool KEEP_LAST = (R > 1 && R * W <= 16);   // larger tiles: recomputing the last row's
                                                         // triple is cheaper than 2*W live registers
    ca::Triple prev[W], cur[W], dn[W], last[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int wl = (w + W - 1) % W, wr = (w + 1) % W;
        cur[w] = ca::row_triple(ca::west(x[0][wl], x[0][w]), x[0][w],
                                ca::east(x[0][w], x[0][wr]));
        last[w] = cur[w];
        if constexpr (R > 1)
            last[w] = ca::row_triple(ca::west(x[R - 1][wl], x[R - 1][w]), x[R - 1][w],
                                     ca::east(x[R - 1][w], x[R - 1][wr]));
        prev[w].lo = __shfl_sync(0xFFFFFFFFu, last[w].lo, up_lane);
        prev[w].hi = __shfl_sync(0xFFFFFFFFu, last[w].hi, up_lane);
        dn[w].lo = __shfl_sync(0xFFFFFFFFu, cur[w].lo, dn_lane);
        dn[w].hi = __shfl_sync(0xFFFFFFFFu, cur[w].hi, dn_lane);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ca::Triple nxt[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if (r + 1 < R && !(KEEP_LAST && r + 2 == R)) {
                const int wl = (w + W - 1) % W, wr = (w + 1) % W;
                nxt[w] = ca::row_triple(ca::west(x[r + 1][wl], x[r + 1][w]), x[r + 1][w],
                                        ca::east(x[r + 1][w], x[r + 1][wr]));
            } else if (r + 1 < R) {
                nxt[w] = last[w];
            } else {
                nxt[w] = dn[w];
            }
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            x[r][w] = rule.from_triples(x[r][w], prev[w], cur[w], nxt[w]);
            prev[w] = cur[w];
            cur[w] = nxt[w];
        }
    }
}

// the square case: one warp holds a whole 32*WPR x 32*WPR instance
template <int WPR, class Rule>
__device__ __forceinline__ void generation(uint32_t (&x)[WPR][WPR], const Rule& rule,
                                           int up_lane, int dn_lane) {
    generation_rw<WPR, WPR>(x, rule, up_lane, dn_lane);
}

// fused SpeedDetector sums (carle/mcl.py:773-779) of the instance held by this warp
template <int WPR>
__device__ __forceinline__ void instance_sums(const StepParams& p, const uint32_t (&x)[WPR][WPR],
                                              int lane, long long* out) {
    uint32_t live = 0, sh = 0, sw = 0, wl = 0;
#pragma unroll
    for (int r = 0; r < WPR; ++r) {
        const int row = lane * WPR + r;
        const bool in_rows = (row >= p.row0) && (row < p.row0 + p.aw);
        uint32_t rowcnt = 0;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
            const uint32_t v = x[r][w];
            const uint32_t inside = in_rows ? (v & window_col_mask(p, w)) : 0u;
            const uint32_t outside = v ^ inside;
            const uint32_t c = ca::popc32(outside);
            live += ca::popc32(v);
            wl += ca::popc32(inside);
            rowcnt += c;
            sw += 32u * w * c + ca::bit_index_sum(outside);
        }
        sh += (uint32_t)row * rowcnt;
    }
    live = __reduce_add_sync(0xFFFFFFFFu, live);
    sh = __reduce_add_sync(0xFFFFFFFFu, sh);
    sw = __reduce_add_sync(0xFFFFFFFFu, sw);
    wl = __reduce_add_sync(0xFFFFFFFFu, wl);
    if (lane == 0) {
        longlong2* o = reinterpret_cast<longlong2*>(out);
        o[0] = make_longlong2(live, sh);
        o[1] = make_longlong2(sw, wl);
    }
}

// ---- observation materialised by the step kernel itself -------------------------------------------
// The reference's observation IS its float32 state (carle/env.py:184-186, 236).  The lane layout of
// the register-resident kernels (lane L holds R consecutive rows of W words) makes a lane's R*W
// words -- and therefore their 32*R*W unpacked cells -- contiguous in memory.  One warp store
// instruction writes FOUR whole words (4 x 128 bytes of float32: lanes 8g..8g+7 expand word i of
// source lane 4t+g, one nibble each), fetched with ONE shuffle whose source lane differs per lane
// group.  No shared memory, no second launch, no re-read of the packed state.
// `unit`: first unpacked cell of the warp's unit (instance or strip), out[unit + ((L*R*W + i)*32 + b)]
// = bit b of word i of lane L.
template <typename O> struct ObsVec;
template <> struct ObsVec<float> {
    using type = float4;
    static __device__ __forceinline__ float4 expand(uint32_t nib) {
        return make_float4((nib & 1u) ? 1.f : 0.f, (nib & 2u) ? 1.f : 0.f, (nib & 4u) ? 1.f : 0.f,
                           (nib & 8u) ? 1.f : 0.f);
    }
};
template <> struct ObsVec<uint8_t> {
    using type = uint32_t;
    static __device__ __forceinline__ uint32_t expand(uint32_t nib) {
        // bit k -> byte k: spread the nibble with a multiply (0x00204081 places bit k at 8k)
        return ((nib & 0xFu) * 0x00204081u) & 0x01010101u;
    }
};

#ifndef CARLE_OBS_CONTIGUOUS
#define CARLE_OBS_CONTIGUOUS 1   // one warp store = 512 contiguous bytes (four words of ONE source lane)
#endif
template <typename O, int WORDS>
__device__ __forceinline__ void emit_obs(const uint32_t* x, O* out, long long unit, int lane) {
    using V = typename ObsVec<O>::type;
    const int g = lane >> 3, sh = (lane & 7) * 4;
#if CARLE_OBS_CONTIGUOUS
    // Every store instruction covers four consecutive words of one source lane: 32 lanes x 16 bytes
    // (float32) = 512 contiguous bytes, walking the unit's unpacked cells front to back.  Lane group g
    // needs word 4k + g of the source lane, a different REGISTER per group: four shuffles and a select.
    static_assert(WORDS % 4 == 0, "words per lane");
    V* dst = reinterpret_cast<V*>(out + unit) + lane;
#pragma unroll 1
    for (int src = 0; src < 32; ++src) {
#pragma unroll
        for (int k = 0; k < WORDS / 4; ++k) {
            const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, x[4 * k + 0], src);
            const uint32_t w1 = __shfl_sync(0xFFFFFFFFu, x[4 * k + 1], src);
            const uint32_t w2 = __shfl_sync(0xFFFFFFFFu, x[4 * k + 2], src);
            const uint32_t w3 = __shfl_sync(0xFFFFFFFFu, x[4 * k + 3], src);
            const uint32_t word = (g & 2) ? ((g & 1) ? w3 : w2) : ((g & 1) ? w1 : w0);
            __stcs(dst + ((long long)src * WORDS + 4 * k) * 8, ObsVec<O>::expand(word >> sh));
        }
    }
#else
    V* dst = reinterpret_cast<V*>(out + unit) + (lane & 7);
#pragma unroll 2
    for (int t = 0; t < 8; ++t) {
        const int src = 4 * t + g;
#pragma unroll
        for (int i = 0; i < WORDS; ++i) {
            const uint32_t word = __shfl_sync(0xFFFFFFFFu, x[i], src);
            // (streaming stores; plain write-back stores measured the same or slower, 55.2 vs 56.4 us
            //  at 4096 x 128x128 -- and a run-time switch between the two cost 240 bytes of spills)
            __stcs(dst + ((long long)src * WORDS + i) * 8, ObsVec<O>::expand(word >> sh));
        }
    }
#endif
}

template <int WORDS>
__device__ __forceinline__ void emit_obs_any(const StepParams& p, const uint32_t* x, long long unit,
                                             int lane) {
    if (p.obs_u8) emit_obs<uint8_t, WORDS>(x, static_cast<uint8_t*>(p.obs), unit, lane);
    else emit_obs<float, WORDS>(x, static_cast<float*>(p.obs), unit, lane);
}

// ---- K generations, pre-packed actions -----------------------------------------------------
template <int WPR, class Rule>
__global__ void __launch_bounds__(128, warp_kernel_min_ctas(WPR))
step_warp_kernel(const __grid_constant__ StepParams p) {
    constexpr int WORDS = WPR * WPR;            // words per lane
    const int lane = threadIdx.x & 31;
    const long long warps_per_block = blockDim.x >> 5;
    // (warp index through a shuffle: the compiler then knows the instance loop is warp-uniform and
    //  drops the WARPSYNC / ENDCOLLECTIVE pair it otherwise wraps around every neighbour shuffle)
    const long long warp0 = (long long)blockIdx.x * warps_per_block +
                            __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const Rule rule(p);
    const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;

    for (long long inst = warp0; inst < p.n; inst += nwarps) {
        uint32_t x[WPR][WPR];
        load_state<WPR>(x, p.in + inst * (32LL * WORDS) + (long long)lane * WORDS);
        for (int g = 0; g < p.k; ++g) {
            // ---- action XOR (carle/env.py:179-182) ----
            if (p.act) {
                const uint32_t* act_inst = p.act + (long long)g * p.act_step_stride +
                                           inst * p.act_inst_stride;
#pragma unroll
                for (int r = 0; r < WPR; ++r) {
                    const int ar = lane * WPR + r - p.row0;
                    if (ar >= 0 && ar < p.aw) {
                        const uint32_t* arow = act_inst + (long long)ar * p.awpr - p.aw0;
#pragma unroll
                        for (int w = 0; w < WPR; ++w)
                            if (w >= p.aw0 && w < p.aw0 + p.awpr) x[r][w] ^= arow[w];
                    }
                }
            }
            const bool reset = p.flags && p.flags[2 * g] == 0 && !p.defer_reset;   // warp-uniform
            if (reset) {
                // master reset (carle/env.py:208-216): every toggle was 1.0
#pragma unroll
                for (int i = 0; i < WORDS; ++i) (&x[0][0])[i] = 0u;
            } else {
                generation<WPR>(x, rule, up_lane, dn_lane);
            }
            if (p.red) instance_sums<WPR>(p, x, lane, p.red + ((long long)g * p.n + inst) * 4);
        }
        store_state<WPR>(x, p.out + inst * (32LL * WORDS) + (long long)lane * WORDS);
    }
    retire_block(p);
}

// ---- one env step from the caller's unpacked action, ONE kernel ---------------------------------
// One warp per instance, exactly one instance per warp (no loop).  G = AW / WPR window-row
// groups and C = AH / 32 chunks are compile-time so the 32 (64) action loads use immediate
// offsets from one base pointer.  Retirement is counted per warp through shared memory (no
// __syncthreads at the tail: warps of a block finish at different times).
#ifndef CARLE_FUSED_W8_CTAS
#define CARLE_FUSED_W8_CTAS 2
#endif
constexpr int fused_min_ctas(int wpr) { return wpr <= 4 ? 7 : CARLE_FUSED_W8_CTAS; }

template <int WPR, class Rule, typename T, int C, int G>
__global__ void __launch_bounds__(128, fused_min_ctas(WPR))
step_fused_kernel(const __grid_constant__ StepParams p) {
    constexpr int WORDS = WPR * WPR;
    __shared__ unsigned int s_done;
    __shared__ int s_flag[3];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long inst = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (threadIdx.x == 0) { s_done = 0u; s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0; }
    __syncthreads();
    bool not_one = false, any = false, nonbin = false;
    if (inst < p.n) {
        const Rule rule(p);
        uint32_t x[WPR][WPR];
        load_state<WPR>(x, p.in + inst * (32LL * WORDS) + (long long)lane * WORDS);
        // ---- action ingestion (carle/env.py:179-182, 191, 208) ----
        const T* a = static_cast<const T*>(p.raw) + inst * p.raw_inst_stride + lane;
        const int my_group = lane - p.row0 / WPR;       // window row group this lane owns
        uint32_t mine[WPR][C];
#pragma unroll
        for (int s = 0; s < WPR; ++s)
#pragma unroll
            for (int c = 0; c < C; ++c) mine[s][c] = 0u;
        uint32_t differs = 0u, seen = 0u;
        NonBinary nb;
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int s = 0; s < WPR; ++s)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const T v = a[((g * WPR + s) * C + c) * 32];
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, v != T(0));
                    differs |= bits_of(v) ^ OneBits<T>::value;
                    nb.see(v);
                    seen |= m;
                    if (my_group == g) mine[s][c] = m;
                }
        }
        not_one = __any_sync(0xFFFFFFFFu, differs != 0u);
        nonbin = __any_sync(0xFFFFFFFFu, nb.any_lane());
        any = seen != 0u;
        const int bit0 = p.col0 - 32 * p.aw0;
#pragma unroll
        for (int s = 0; s < WPR; ++s) {
            // the row's toggles as a bit string shifted left by bit0, split over <= C+1 words
            uint32_t word[C + 1];
#pragma unroll
            for (int c = 0; c <= C; ++c) {
                const uint32_t cur = (c < C) ? mine[s][c] : 0u;
                const uint32_t prv = (c > 0) ? mine[s][c - 1] : 0u;
                word[c] = bit0 ? ((cur << bit0) | (prv >> (32 - bit0))) : cur;
            }
#pragma unroll
            for (int w = 0; w < WPR; ++w)
#pragma unroll
                for (int c = 0; c <= C; ++c)
                    if (w == p.aw0 + c) x[s][w] ^= word[c];
        }
        generation<WPR>(x, rule, (lane + 31) & 31, (lane + 1) & 31);
        if (p.red) instance_sums<WPR>(p, x, lane, p.red + inst * 4);
        store_state<WPR>(x, p.out + inst * (32LL * WORDS) + (long long)lane * WORDS);
        if (p.reward_zero && lane == 0) p.reward_zero[inst] = 0.f;
        if (p.obs) emit_obs_any<WORDS>(p, &x[0][0], inst * (1024LL * WORDS), lane);
    }
    // ---- retirement: warp -> block (shared memory) -> grid (global), no tail barrier ----
    if (retire_legacy<T>(p, &s_done, s_flag, lane, warps_per_block, not_one, any, nonbin) == 2)
        clear_after_reset(p, lane);           // rare: the whole batch asked for the master reset
}

// =========================================================================================
// device-side random agent: generator (carle/agents.py:35-42: Bernoulli(toggle_rate) per toggle)
// =========================================================================================
// Philox4x32-10 counter-based generator (Salmon et al.): stateless, so every (entry, row, chunk,
// step) draws its own stream and the result does not depend on the launch configuration.
struct Philox {
    static __device__ __forceinline__ uint4 rounds(uint4 c, uint2 k) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
        }
        return c;
    }
};

// 32 Bernoulli(threshold / 65536) toggles: window columns [32j, 32j+32) of window row `row` of
// action entry `entry` at environment step `step`.  Four Philox calls give 16 words = 16 BIT
// PLANES of 32 independent 16-bit uniforms (plane k bit i = bit k of cell i's number); the
// comparison "number < threshold" runs bit-sliced from the LSB up, one LOP3 per plane:
//   lt' = T_k ? (~r_k | lt) : (~r_k & lt)   with T_k the k-th threshold bit.
__device__ __forceinline__ uint32_t random_chunk(long long entry, uint32_t row, int j, uint32_t step,
                                                 uint2 key, uint32_t threshold) {
    uint32_t lt = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 r = Philox::rounds(
            make_uint4((uint32_t)entry, (uint32_t)(entry >> 32), row * 64u + j * 4u + q, step), key);
        const uint32_t plane[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t tk = 0u - ((threshold >> (4 * q + i)) & 1u);
            lt = ca::lop3<0x8E>(plane[i], lt, tk);      // (~r & (lt | tk)) | (lt & tk)
        }
    }
    return threshold > 0xFFFFu ? 0xFFFFFFFFu : lt;
}

// action element type of the step kernels whose action IS the device-side random agent
struct DeviceRandom { unsigned char unused; };
template <typename T> struct IsDeviceRandom { static constexpr bool value = false; };
template <> struct IsDeviceRandom<DeviceRandom> { static constexpr bool value = true; };
// action "element" type of the step kernels fed ALREADY PACKED actions ([B][AW][AWPR] words aligned
// to the universe's word grid, include/carle_b200.h): 1 bit per toggle instead of 32 -- the
// information minimum of SURVEY.md section 8(d); a toggle is one shared-memory load and one XOR
struct PackedWords { uint32_t w; };
template <typename T> struct IsPackedWords { static constexpr bool value = false; };
template <> struct IsPackedWords<PackedWords> { static constexpr bool value = true; };
template <typename T> struct IsFloatAction { static constexpr bool value = false; };
template <> struct IsFloatAction<float> { static constexpr bool value = true; };
// the valid toggle bits of packed word j of a window row (window of AH columns from COL0)
template <int COL0, int AH>
__device__ __forceinline__ constexpr uint32_t packed_valid_mask(int j) {
    return ca::window_col_mask_of(COL0 / 32 + j, COL0, AH);
}

// XOR one action row, given as C ballot masks, into the words of a universe row
template <int WPL, int AW0, int BIT0, int C>
__device__ __forceinline__ void xor_action_row(uint32_t (&row)[WPL], const uint32_t (&m)[C]) {
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        if (c == C && BIT0 == 0) break;
        const uint32_t cur = (c < C) ? m[c < C ? c : 0] : 0u;
        const uint32_t prv = (c > 0) ? m[c > 0 ? c - 1 : 0] : 0u;
        row[AW0 + c] ^= BIT0 ? ((cur << BIT0) | (prv >> ((32 - BIT0) & 31))) : cur;
    }
}

// ---- the same step as a PERSISTENT, TMA-staged pipeline -------------------------------------
// One CTA per SM slot, every warp walks instances warp_id, warp_id + W, ...  Each warp owns a
// two-slot ring in shared memory; one elected lane issues cp.async.bulk (the TMA engine's 1-D
// bulk copy, UBLKCP in SASS) for the NEXT instance's packed state and unpacked action, completion
// is signalled on an mbarrier, and the warp meanwhile ballots / advances the CURRENT instance
// out of shared memory.  Global-memory latency is therefore hidden by the pipeline, not by
// occupancy (which the register-heavy WPR = 8 variant does not have), and the next state leaves
// through a bulk shared->global store.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy(bool evict_first) {
    unsigned long long pol;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// true in exactly one (the lowest active) lane of the converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
}  // namespace tma

// CTA shape of the persistent stream kernel (measured, profiles/r1c_ab_ctas.txt,
// r1d_ab_cta_shape.txt): ONE 16-warp CTA per SM with two TMA slots per warp wins for batches of
// one or two instances per warp (4096 x 128x128: 7.17 us vs 7.42 us as two 8-warp CTAs: fewer
// retirement atomics, one prologue per SM) and for 64x64; long 128x128 batches (`big`) run best
// as three 8-warp CTAs with one slot (32768 instances: 45.6 us vs 49.5 / 50.4 us).  256x256
// keeps 8 warps (255 registers per thread).
constexpr int stream_warps(int wpr, bool big) { return wpr <= 2 ? 16 : (wpr <= 4 ? (big ? 8 : 16) : 8); }
constexpr int stream_min_ctas(int wpr, bool big) { return (wpr > 2 && wpr <= 4 && big) ? 3 : 1; }

template <int WPR, typename T, int C, int G>
struct StreamLayout {
    static constexpr int STATE_BYTES = 32 * WPR * WPR * 4;
    static constexpr bool RANDOM = IsDeviceRandom<T>::value;   // toggles drawn in the kernel
    static constexpr bool PACKED = IsPackedWords<T>::value;    // toggles arrive as grid-aligned words
    static constexpr int COL0 = (32 * WPR - 32 * C) / 2;
    static constexpr int AWPR = C + (COL0 % 32 ? 1 : 0);       // packed words per window row
    static constexpr int ACT_BYTES = RANDOM ? 0 : PACKED ? G * WPR * AWPR * 4
                                                         : G * WPR * C * 32 * (int)sizeof(T);
    static_assert(ACT_BYTES % 16 == 0, "bulk copy size");
    static constexpr int SLOT_BYTES = STATE_BYTES + ACT_BYTES;
    static constexpr int MASK_BYTES = (RANDOM || PACKED) ? 0 : G * WPR * C * 4;   // one ballot mask per 32 toggles
    static constexpr int warp_bytes(int depth) {             // slots + masks + mbarriers
        return depth * SLOT_BYTES + MASK_BYTES + 16;
    }
};

// programmatic dependent launch (no-ops when the launch carries no programmatic dependency)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// A slot is drained into registers (state words + ballotted action masks) as soon as it lands
// and is refilled at once, so the bulk copy of a later instance flies while this instance's
// generation is computed.  DEPTH slots per warp: with DEPTH = 2 the first TWO instances of every
// warp are requested at kernel start, which matters for batches of only one or two instances
// per warp (4096 x 128 x 128: the whole step is a single memory round trip).
// For short batches warps are ranked block-interleaved (rank = warp_in_block * gridDim + block),
// so a partial last trip is spread evenly over the CTAs / SMs instead of filling the first blocks.
// Programmatic dependent launch: launch_dependents is signalled at once and the previous grid is
// awaited (griddepcontrol.wait) before global memory is touched, so back-to-back steps overlap
// the launch latency and this prologue with the previous step's tail.
template <int WPR, class Rule, typename T, int C, int G, int DEPTH, bool BIG>
__global__ void __launch_bounds__(32 * stream_warps(WPR, BIG), stream_min_ctas(WPR, BIG))
step_stream_kernel(const __grid_constant__ StepParams p) {
    using L = StreamLayout<WPR, T, C, G>;
    constexpr int WORDS = WPR * WPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned int s_done;
    __shared__ double s_sd;
    const int lane = threadIdx.x & 31;
    // (through a shuffle so the compiler knows it is warp-uniform: the bulk-copy operands then
    //  live in uniform registers)
    const int wib = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
    const int warps_per_block = blockDim.x >> 5;
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const long long warp = p.rank_blocked ? (long long)blockIdx.x * warps_per_block + wib
                                          : (long long)wib * gridDim.x + blockIdx.x;
    unsigned char* wbase = smem_raw + (size_t)wib * L::warp_bytes(DEPTH);
    uint32_t* amask = reinterpret_cast<uint32_t*>(wbase + DEPTH * L::SLOT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + DEPTH * L::SLOT_BYTES + L::MASK_BYTES);

    pdl_launch_dependents();
    if (threadIdx.x == 0) { s_done = 0u; s_sd = 0.0; }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) tma::mbar_init(bars + s, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();                                     // the previous step's state / flags are final

    const Rule rule(p);
    const char* in_bytes = reinterpret_cast<const char*>(p.in);
    const char* act_bytes = static_cast<const char*>(p.raw);
    const long long act_stride = p.raw_inst_stride * (long long)sizeof(T);
    const unsigned long long act_policy = tma::l2_policy(p.act_evict_first != 0);
    // called by the whole (converged) warp; one elected lane issues.  dep == 0 (StepParams::zero)
    auto issue = [&](int s, long long inst, uint32_t dep) {
        const uint32_t slot = tma::smem_u32(wbase + s * L::SLOT_BYTES);
        const uint32_t bar = tma::smem_u32(bars + s);
        if (tma::elect_one()) {
            tma::mbar_expect_tx_u32(bar, L::SLOT_BYTES + dep);
            tma::bulk_g2s_u32(slot, in_bytes + inst * L::STATE_BYTES, L::STATE_BYTES, bar);
            if constexpr (!L::RANDOM)
                tma::bulk_g2s_hint(slot + L::STATE_BYTES, act_bytes + inst * act_stride, L::ACT_BYTES, bar,
                                   act_policy);
        }
        __syncwarp();
    };
    // unit `it` of this warp's walk -> instance (see StepParams::reverse)
    auto unit_inst = [&](long long it) { return p.reverse ? p.n - 1 - it : it; };

    bool warp_not_one = false, warp_any = false, warp_nonbin = false;
    const bool sd_on = CARLE_FEAT_SD && p.sd_com != nullptr;
    // centred window (carle/env.py:119-132): the geometry follows from the template shape
    constexpr int ROW0 = (32 * WPR - G * WPR) / 2, COL0 = (32 * WPR - 32 * C) / 2;
    static_assert(ROW0 % WPR == 0, "window rows start on a lane boundary");
    const int my_group = lane - ROW0 / WPR;             // window row group this lane owns
#pragma unroll
    for (int s = 0; s < DEPTH; ++s)
        if (warp + s * nwarps < p.n) issue(s, unit_inst(warp + s * nwarps), 0u);
    int trip = 0;
    for (long long it = warp; it < p.n; it += nwarps, ++trip) {
        const long long inst = unit_inst(it);
        const int sl = trip % DEPTH;
        const unsigned char* slot = wbase + sl * L::SLOT_BYTES;
        tma::mbar_wait(bars + sl, (uint32_t)((trip / DEPTH) & 1));
        uint32_t x[WPR][WPR];
        load_state<WPR>(x, reinterpret_cast<const uint32_t*>(slot) + lane * WORDS);
        uint32_t mine[WPR][C];
        uint32_t pw[WPR][L::AWPR];                       // packed actions: this lane's grid-aligned words
        bool inst_not_one, inst_any, inst_nonbin = false;
        if constexpr (L::PACKED) {
            // ---- the action arrives packed: lanes that own window rows fetch their words; the
            //      batch-wide flags come from the words (all valid bits set / some bit set) ----
            const bool in = (unsigned)my_group < (unsigned)G;
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(slot + L::STATE_BYTES) +
                                 (in ? my_group : 0) * (WPR * L::AWPR);
            uint32_t seen = 0u, all_set = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < WPR; ++r)
#pragma unroll
                for (int j = 0; j < L::AWPR; ++j) {
                    const uint32_t w = in ? aw[r * L::AWPR + j] : 0u;
                    constexpr int CC0 = L::COL0;
                    const uint32_t valid = packed_valid_mask<CC0, 32 * C>(j);
                    pw[r][j] = w & valid;
                    seen |= w & valid;
                    all_set &= in ? (w | ~valid) : 0xFFFFFFFFu;
                }
            inst_any = __any_sync(0xFFFFFFFFu, seen != 0u);
            inst_not_one = __any_sync(0xFFFFFFFFu, all_set != 0xFFFFFFFFu);
        } else if constexpr (L::RANDOM) {
            // ---- the action IS the random agent (carle/agents.py:35-42): lane l draws window rows
            //      l, l+32, ..; the lanes that own those universe rows fetch them by shuffle ----
            constexpr int AWR = G * WPR, SLOTS = (AWR + 31) / 32;
            const long long entry = p.raw_inst_stride ? inst : 0;      // batch-1: one shared action
            const uint2 key = make_uint2(p.rand_key0, p.rand_key1);
            uint32_t drawn[SLOTS][C], all = 0xFFFFFFFFu, some = 0u;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int r = sl * 32 + lane;
                    drawn[sl][c] = (r < AWR) ? random_chunk(entry, (uint32_t)r, c, p.rand_step, key,
                                                            p.rand_threshold) : 0u;
                    if (r < AWR) { all &= drawn[sl][c]; some |= drawn[sl][c]; }
                }
            inst_not_one = __any_sync(0xFFFFFFFFu, all != 0xFFFFFFFFu);
            inst_any = __any_sync(0xFFFFFFFFu, some != 0u);
            const bool in = (unsigned)my_group < (unsigned)G;
#pragma unroll
            for (int r = 0; r < WPR; ++r) {
                const int row = (in ? my_group : 0) * WPR + r;         // window row of x[r][*]
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    uint32_t m = 0u;
#pragma unroll
                    for (int sl = 0; sl < SLOTS; ++sl) {
                        const uint32_t got = __shfl_sync(0xFFFFFFFFu, drawn[sl][c], row & 31);
                        if ((row >> 5) == sl) m = got;
                    }
                    mine[r][c] = in ? m : 0u;
                }
            }
        } else {
        // ---- action ingestion out of shared memory (carle/env.py:179-182, 191, 208): one ballot
        //      per 32 toggles, the masks parked in the warp's mask area ----
        const T* a = reinterpret_cast<const T*>(slot + L::STATE_BYTES) + lane;
        NonBinary nb;
#pragma unroll
        for (int j = 0; j < G * WPR; j += 4) {
            T v[4][C];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c) v[i][c] = a[((j + i) * C + c) * 32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, v[i][c] != T(0));
                    if (lane == 0) amask[(j + i) * C + c] = m;
                }
            nb.see_all(reinterpret_cast<const T(&)[4 * C]>(v));
        }
        inst_nonbin = __any_sync(0xFFFFFFFFu, nb.any_lane());
        __syncwarp();
        {
            const bool in = (unsigned)my_group < (unsigned)G;      // this lane's rows are window rows
            const uint32_t* mrow = amask + (in ? my_group : 0) * (WPR * C);
#pragma unroll
            for (int r = 0; r < WPR; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c) mine[r][c] = in ? mrow[r * C + c] : 0u;
        }
        // batch-wide flags: some toggle != 0, and some toggle != 1.0 (master reset, env.py:208).  A
        // zero toggle settles the second, and the masks already say whether there is one; only if
        // EVERY toggle of the instance is non-zero are the values themselves compared with 1.0
        // (the slot is not refilled yet).
        uint32_t seen = 0u, all_set = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < (G * WPR * C + 31) / 32; ++k)
            if (k * 32 + lane < G * WPR * C) {
                const uint32_t m = amask[k * 32 + lane];
                seen |= m;
                all_set &= m;
            }
        inst_any = __any_sync(0xFFFFFFFFu, seen != 0u);
        inst_not_one = __any_sync(0xFFFFFFFFu, all_set != 0xFFFFFFFFu);
        if (!inst_not_one) {
            uint32_t differs = 0u;
            for (int j = 0; j < G * WPR * C; ++j) differs |= bits_of(a[j * 32]) ^ OneBits<T>::value;
            inst_not_one = __any_sync(0xFFFFFFFFu, differs != 0u);
        }
        }
        // The refill overwrites the slot through the async proxy, and a bank-conflicted LDS can
        // still be queued in the LSU when later instructions issue: make the refill's operands
        // depend on one register of every state load (the ballots consumed the action values).
        uint32_t dep = 0u;
#pragma unroll
        for (int i = 0; i < (WORDS + 3) / 4; ++i) dep ^= (&x[0][0])[(4 * i < WORDS) ? 4 * i : 0];
        if constexpr (L::PACKED) {
#pragma unroll
            for (int r = 0; r < WPR; ++r) dep ^= pw[r][0] ^ pw[r][L::AWPR - 1];
        }
        dep &= p.zero;
        __syncwarp();                                   // the slot is drained: refill it
        const long long next = it + DEPTH * nwarps;
        if (next < p.n) issue(sl, unit_inst(next), dep);
        warp_not_one |= inst_not_one;
        warp_any |= inst_any;
        warp_nonbin |= inst_nonbin;
        if constexpr (L::PACKED) {
#pragma unroll
            for (int r = 0; r < WPR; ++r)
#pragma unroll
                for (int j = 0; j < L::AWPR; ++j) x[r][COL0 / 32 + j] ^= pw[r][j];
        } else {
#pragma unroll
            for (int r = 0; r < WPR; ++r) xor_action_row<WPR, COL0 / 32, COL0 % 32, C>(x[r], mine[r]);
        }
        generation<WPR>(x, rule, (lane + 31) & 31, (lane + 1) & 31);
        if (p.red) {                                    // fused SpeedDetector sums (carry-save form)
            uint32_t live = 0, sh = 0, sw = 0, wl = 0;
            ca::strip_lane_sums<WPR, WPR, 32 * C>(x, lane * WPR, live, sh, sw, wl);
            live = __reduce_add_sync(0xFFFFFFFFu, live);
            sh = __reduce_add_sync(0xFFFFFFFFu, sh);
            sw = __reduce_add_sync(0xFFFFFFFFu, sw);
            wl = __reduce_add_sync(0xFFFFFFFFu, wl);
            if (lane == 0) {
                longlong2* o = reinterpret_cast<longlong2*>(p.red + inst * 4);
                o[0] = make_longlong2(live, sh);
                o[1] = make_longlong2(sw, wl);
            }
        }
        store_state<WPR>(x, p.out + inst * (32LL * WORDS) + (long long)lane * WORDS);
        if (p.reward_zero && !sd_on && lane == 0) p.reward_zero[inst] = 0.f;
        if (CARLE_FEAT_OBS && p.obs) emit_obs_any<WORDS>(p, &x[0][0], inst * (1024LL * WORDS), lane);
        fence_if_all_ones(inst_not_one && !inst_nonbin);
    }
    // ---- retirement: warp -> block (shared memory) -> grid (global), flags inside the atomics ----
    bool sd_primed = false;
    if (sd_on) {
        // SpeedDetector tail (carle/mcl.py:777-795) as a pass BEHIND the instance loop, one of this
        // warp's instances per lane: the sums are read back from where lane 0 stored them above
        // (same warp: ordered by the __syncwarp), so the loop itself carries no state of the tail
        // and the float work runs once per 32 instances instead of once per instance on one lane
        sd_primed = *p.sd_primed != 0;                  // (the previous step set it)
        double sd_local = 0.0;
        __syncwarp();
        for (long long k = lane; warp + k * nwarps < p.n; k += 32) {
            const long long inst = unit_inst(warp + k * nwarps);
            const longlong2* o = reinterpret_cast<const longlong2*>(p.red + inst * 4);
            const longlong2 a = o[0], b = o[1];
            sd_local += speed_instance(p, inst, sd_primed, p.sd_com_prev[inst], p.sd_com_prev[p.n + inst],
                                       (uint32_t)a.x, (unsigned long long)a.y, (unsigned long long)b.x);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sd_local += __shfl_xor_sync(0xFFFFFFFFu, sd_local, off);
        speed_warp_done(&s_sd, lane, sd_local);
    }
    const int last_of_grid = retire_fused<T>(p, &s_done, lane, warps_per_block, warp_not_one,
                                             warp_any, warp_nonbin, sd_on ? &s_sd : nullptr);
    if (last_of_grid == 2) clear_after_reset(p, lane);
    if (last_of_grid && sd_on) speed_grid_done(p, lane, sd_primed, last_of_grid == 2);
}

// =========================================================================================
// generic family: any (even, square) shape, one generation per launch
// =========================================================================================
template <class Rule>
__global__ void __launch_bounds__(256)
step_generic_kernel(const __grid_constant__ StepParams p) {
    const Rule rule(p);
    const long long words_per_inst = (long long)p.h * p.wpr;
    const long long total = p.n * words_per_inst;
    const int tail = p.w & 31;
    const uint32_t tailmask = tail ? ((1u << tail) - 1u) : 0xFFFFFFFFu;
    const bool reset = p.flags && p.flags[0] == 0 && !p.defer_reset;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (reset) { p.out[idx] = 0u; continue; }
        const long long inst = idx / words_per_inst;
        const int rem = (int)(idx - inst * words_per_inst);
        const int r = rem / p.wpr, w = rem - r * p.wpr;
        const uint32_t* base = p.in + inst * words_per_inst;
        const uint32_t* act_inst = p.act ? p.act + inst * p.act_inst_stride : nullptr;
        const int wl = (w == 0) ? p.wpr - 1 : w - 1;
        const int wr = (w == p.wpr - 1) ? 0 : w + 1;
        ca::Triple t[3];
        uint32_t centre = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int rr = r + d - 1;
            rr = (rr < 0) ? p.h - 1 : (rr >= p.h ? 0 : rr);
            const uint32_t* row = base + (long long)rr * p.wpr;
            uint32_t xl = row[wl], xc = row[w], xr = row[wr];
            if (act_inst) {
                xl ^= action_bits(p, act_inst, rr, wl);
                xc ^= action_bits(p, act_inst, rr, w);
                xr ^= action_bits(p, act_inst, rr, wr);
            }
            // seam: the row's last word holds only `tail` cells when W % 32 != 0
            uint32_t prev = (w == 0 && tail) ? (xl << (32 - tail)) : xl;
            uint32_t west = ca::west(prev, xc);
            uint32_t east = (w == p.wpr - 1 && tail)
                                ? ((xc >> 1) | ((xr & 1u) << (tail - 1)))
                                : ca::east(xc, xr);
            t[d] = ca::row_triple(west, xc, east);
            if (d == 1) centre = xc;
        }
        uint32_t nx = rule.from_triples(centre, t[0], t[1], t[2]);
        if (w == p.wpr - 1) nx &= tailmask;
        p.out[idx] = nx;
    }
    retire_block(p);
}

// =========================================================================================
// boundary converters
// =========================================================================================
template <typename T> __device__ __forceinline__ bool is_on(T v) { return v != T(0); }

// cells [N*H][W] (T) -> packed words [N*H][WPR]; one warp packs 32 words per trip
template <typename T>
__global__ void __launch_bounds__(256)
pack_state_kernel(const T* __restrict__ cells, uint32_t* __restrict__ packed,
                  long long rows, int w, int wpr) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = rows * wpr;
    for (long long base = warp * 32; base < total; base += nwarps * 32) {
        uint32_t mine = 0;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            long long j = base + i;
            bool on = false;
            if (j < total) {
                long long row = j / wpr;
                int wi = (int)(j - row * wpr);
                int col = 32 * wi + lane;
                if (col < w) on = is_on(cells[row * w + col]);
            }
            uint32_t word = __ballot_sync(0xFFFFFFFFu, on);
            if (i == lane) mine = word;
        }
        if (base + lane < total) packed[base + lane] = mine;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_state_kernel(const uint32_t* __restrict__ packed, T* __restrict__ cells,
                    long long rows, int w, int wpr) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = rows * wpr;
    for (long long base = warp * 32; base < total; base += nwarps * 32) {
        uint32_t mine = (base + lane < total) ? packed[base + lane] : 0u;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            long long j = base + i;
            uint32_t word = __shfl_sync(0xFFFFFFFFu, mine, i);
            if (j < total) {
                long long row = j / wpr;
                int wi = (int)(j - row * wpr);
                int col = 32 * wi + lane;
                if (col < w) cells[row * w + col] = T((word >> lane) & 1u);
            }
        }
    }
}

// fast path for W % 32 == 0 and float32: 8 lanes write one word as 8 float4 (512 B per
// warp store instruction, fully coalesced)
static __global__ void __launch_bounds__(256)
unpack_state_f32_kernel(const uint32_t* __restrict__ packed, float4* __restrict__ cells,
                        long long total_words) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp * 32; base < total_words; base += nwarps * 32) {
        uint32_t mine = (base + lane < total_words) ? __ldg(packed + base + lane) : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int j = i * 4 + (lane >> 3);
            uint32_t word = __shfl_sync(0xFFFFFFFFu, mine, j);
            uint32_t nib = word >> ((lane & 7) * 4);
            float4 v = make_float4((nib & 1u) ? 1.f : 0.f, (nib & 2u) ? 1.f : 0.f,
                                   (nib & 4u) ? 1.f : 0.f, (nib & 8u) ? 1.f : 0.f);
            if (base + j < total_words) __stcs(cells + (base + j) * 8 + (lane & 7), v);
        }
    }
}

// action [K][B*AW rows][AH] (T) -> grid-aligned packed [K][B*AW][AWPR] + per-step flags.
// Bit b of output word j of a row is window column c = 32*(aw0+j) + b - col0 (0 if outside),
// i.e. the row's toggles as a bit string shifted left by bit0 = col0 - 32*aw0.
// grid = (blocks, K): blockIdx.y is the step, so no division is needed for the flags.
// One warp packs R rows per trip: R independent coalesced loads per lane, then R ballots.
template <typename T, int R>
__global__ void __launch_bounds__(256)
pack_action_kernel(const T* __restrict__ action, uint32_t* __restrict__ packed,
                   int* __restrict__ flags, long long rows_per_step, int ah, int awpr,
                   int bit0, unsigned int* __restrict__ nonbin_flag) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long step = blockIdx.y;
    action += step * rows_per_step * ah;
    packed += step * rows_per_step * awpr;
    const int nchunks = (ah + 31) >> 5;
    bool not_one = false, any = false;
    NonBinary nb;
    for (long long r0 = warp * R; r0 < rows_per_step; r0 += nwarps * R) {
        uint32_t carry[R];
#pragma unroll
        for (int i = 0; i < R; ++i) carry[i] = 0u;
        for (int j = 0; j < awpr; ++j) {
            uint32_t m[R];
            const int col = 32 * j + lane;
            const bool have = (j < nchunks) && (col < ah);
            T v[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const long long r = r0 + i;
                v[i] = (have && r < rows_per_step) ? action[r * ah + col] : T(0);
            }
            bool n1 = false;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                m[i] = __ballot_sync(0xFFFFFFFFu, v[i] != T(0));
                n1 |= have && (r0 + i < rows_per_step) && (v[i] != T(1));
                nb.see(v[i]);
            }
            not_one |= n1;
            uint32_t mine = 0u;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const uint32_t word = bit0 ? ((m[i] << bit0) | (carry[i] >> (32 - bit0))) : m[i];
                carry[i] = m[i];
                any |= (m[i] != 0u);
                if (lane == i) mine = word;
            }
            if (lane < R && r0 + lane < rows_per_step) packed[(r0 + lane) * awpr + j] = mine;
        }
    }
    not_one = __any_sync(0xFFFFFFFFu, not_one);
    const bool nonbin = __any_sync(0xFFFFFFFFu, nb.any_lane());
    if (lane == 0) {
        if (not_one) flags[2 * step] = 1;
        if (any) flags[2 * step + 1] = 1;
        if (nonbin && nonbin_flag) *nonbin_flag = 1u;      // -> resolve_packed_flags_kernel
    }
}

// Fast path of the above for AH % 32 == 0 with C = AH/32 a power of two (32, 64, 128, 256
// wide windows): the action tensor is then a flat stream of 32-float chunks, one warp-wide
// coalesced load + one ballot each.  A warp takes 32 consecutive chunks per trip (32 loads
// in flight per lane), lane i keeps chunk i's mask, neighbouring chunk masks of the same row
// come from lane-1 by shuffle, and the grid-aligned words are written with coalesced stores.
template <typename T>
__global__ void __launch_bounds__(256, 4)
pack_action_stream_kernel(const T* __restrict__ action, uint32_t* __restrict__ packed,
                          int* __restrict__ flags, long long chunks_per_step, int cshift,
                          int awpr, int bit0, unsigned int* __restrict__ nonbin_flag) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long step = blockIdx.y;
    const int C = 1 << cshift;                       // chunks per row
    action += step * chunks_per_step * 32;
    packed += step * (chunks_per_step >> cshift) * awpr;
    bool not_one = false, any = false;
    NonBinary nb;
    for (long long q0 = warp * 32; q0 < chunks_per_step; q0 += nwarps * 32) {
        T v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
            v[i] = (q0 + i < chunks_per_step) ? action[(q0 + i) * 32 + lane] : T(1);
        uint32_t mine = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, v[i] != T(0));
            not_one |= (v[i] != T(1));
            nb.see(v[i]);
            if (i == lane) mine = m;
        }
        const long long q = q0 + lane;               // this lane's chunk
        if (q >= chunks_per_step) mine = 0u;
        any |= (mine != 0u);
        const int c = (int)(q & (C - 1));            // chunk index inside its row
        const long long row = q >> cshift;
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, mine, 1);
        if (c == 0) prev = 0u;                       // rows never straddle a trip (32 % C == 0)
        if (q < chunks_per_step) {
            uint32_t* dst = packed + row * awpr;
            dst[c] = bit0 ? ((mine << bit0) | (prev >> (32 - bit0))) : mine;
            if (c == C - 1 && awpr > C) dst[C] = mine >> (32 - bit0);
        }
    }
    not_one = __any_sync(0xFFFFFFFFu, not_one);
    any = __any_sync(0xFFFFFFFFu, any);
    const bool nonbin = __any_sync(0xFFFFFFFFu, nb.any_lane());
    if (lane == 0) {
        if (not_one) flags[2 * step] = 1;
        if (any) flags[2 * step + 1] = 1;
        if (nonbin && nonbin_flag) *nonbin_flag = 1u;      // -> resolve_packed_flags_kernel
    }
}

// Rare path behind the two kernels above (one block, launched after them for float32 actions; exits
// at once unless some element of some step was neither 0.0 nor 1.0): re-derive the flags of every
// step that holds such an element from the reference's own predicates (carle/env.py:191, 208:
// `torch.sum(action)`, `torch.mean(action) == 1.0`; sum in float64, see resolve_action_mean).
static __global__ void __launch_bounds__(256)
resolve_packed_flags_kernel(const float* __restrict__ action, int* __restrict__ flags,
                            long long steps, long long elems_per_step,
                            unsigned int* __restrict__ nonbin_flag) {
    if (*reinterpret_cast<volatile unsigned int*>(nonbin_flag) == 0u) return;
    __shared__ double s_sum[8];
    __shared__ int s_nb[8];
    for (long long st = 0; st < steps; ++st) {
        const float* a = action + st * elems_per_step;
        double sum = 0.0;
        NonBinary nb;
        for (long long i = threadIdx.x; i < elems_per_step; i += blockDim.x) {
            const float v = a[i];
            sum += v;
            nb.see(v);
        }
        int mine_nb = nb.any_lane() ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
            mine_nb |= __shfl_xor_sync(0xFFFFFFFFu, mine_nb, off);
        }
        if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_nb[threadIdx.x >> 5] = mine_nb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            int any_nb = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tot += s_sum[w]; any_nb |= s_nb[w]; }
            if (any_nb) {
                flags[2 * st] = ((float)(tot / (double)elems_per_step) == 1.0f) ? 0 : 1;
                flags[2 * st + 1] = (tot != 0.0) ? 1 : 0;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *nonbin_flag = 0u;
}

// universe / observation / sums cleared by a whole grid: after a reset decided OUTSIDE the step
// kernel (instance-sharded batches, carle_apply_reset: `decision` non-zero) or, with
// decision == nullptr, after a step whose own last block reported a clear in `fired` (the float
// observation a fused step wrote is far too large for that one warp).  Exits at once otherwise.
static __global__ void __launch_bounds__(256)
clear_if_kernel(const int* __restrict__ decision, const unsigned int* __restrict__ fired,
                uint32_t* __restrict__ state, long long state_words, uint32_t* __restrict__ obs,
                long long obs_words, long long* __restrict__ red, long long red_n,
                long long* __restrict__ counters) {
    const bool go = decision ? (*decision != 0) : (*reinterpret_cast<const volatile unsigned int*>(fired) != 0u);
    if (!go) return;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    auto zero = [&](uint32_t* ptr, long long words) {           // 16-byte aligned base
        if (!ptr) return;
        uint4* v = reinterpret_cast<uint4*>(ptr);
        const long long nv = words >> 2;
        for (long long i = tid; i < nv; i += nth) v[i] = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = (nv << 2) + tid; i < words; i += nth) ptr[i] = 0u;
    };
    zero(state, state_words);
    zero(obs, obs_words);
    if (red) for (long long i = tid; i < red_n; i += nth) red[i] = 0;
    if (decision && counters && tid == 0) {
        // the deferred reset's bookkeeping (carle/env.py:142-145): the step itself counted a
        // generation; reset() zeroes both counters
        counters[0] = 0;
        counters[1] = 0;
        counters[2] += 1;
        counters[4] = 0;
    }
}

// flags for already-packed (grid-aligned) actions: "all ones" <=> every valid bit set
static __global__ void __launch_bounds__(256)
packed_action_flags_kernel(const uint32_t* __restrict__ packed, int* __restrict__ flags,
                           long long rows, long long rows_per_step, int ah, int awpr,
                           int bit0) {
    const long long total = rows * awpr;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (long long)gridDim.x * blockDim.x) {
        long long row = j / awpr;
        int wi = (int)(j - row * awpr);
        int lo = max(bit0 - 32 * wi, 0), hi = min(bit0 + ah - 32 * wi, 32);
        uint32_t full = 0u;
        if (hi > lo) full = ((hi - lo == 32) ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u)) << lo;
        uint32_t v = packed[j] & full;
        long long step = row / rows_per_step;
        if (v != full) flags[2 * step] = 1;
        if (v != 0u) flags[2 * step + 1] = 1;
    }
}

// state[n][row0+r][aw0+j] ^= act[n or 0][r][j]   (apply_action without a generation)
static __global__ void __launch_bounds__(256)
apply_action_kernel(const StepParams p, uint32_t* __restrict__ state) {
    const long long per_inst = (long long)p.aw * p.awpr;
    const long long total = p.n * per_inst;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long inst = i / per_inst;
        const int rem = (int)(i - inst * per_inst);
        const int r = rem / p.awpr, j = rem - r * p.awpr;
        const uint32_t a = p.act[inst * p.act_inst_stride + rem];
        if (a) state[(inst * p.h + p.row0 + r) * p.wpr + p.aw0 + j] ^= a;
    }
}

// =========================================================================================
// device-side random agent (carle/agents.py:35-42: Bernoulli(toggle_rate) per toggle)
// =========================================================================================
// packed[b][r][0..awpr) <- Bernoulli(threshold / 65536) toggles for every window cell, written
// straight in the grid-aligned packed layout the step kernels consume (no float tensor at all).
// One thread per (entry, window row); 16 random bits per cell.
static __global__ void __launch_bounds__(256)
random_action_kernel(uint32_t* __restrict__ packed, long long rows_total, int aw, int ah, int awpr,
                     int bit0, uint32_t threshold, uint2 key, uint32_t step) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_total) return;
    const long long entry = t / aw;
    const uint32_t row = (uint32_t)(t - entry * aw);
    uint32_t* out = packed + t * awpr;
    uint32_t carry = 0u;
    const int chunks = (ah + 31) >> 5;
    for (int j = 0; j < awpr; ++j) {
        uint32_t m = 0u;
        if (j < chunks) {
            m = random_chunk(entry, row, j, step, key, threshold);
            const int valid = ah - 32 * j;            // columns of this chunk inside the window
            if (valid < 32) m &= (1u << valid) - 1u;
        }
        out[j] = bit0 ? ((m << bit0) | (carry >> (32 - bit0))) : m;
        carry = m;
    }
}

// One env step whose action IS the device-side random agent: the toggles are generated inside
// the step kernel (same Philox streams as random_action_kernel, so both paths agree bit for
// bit), so a random-agent rollout has no action tensor and no action traffic at all.  Lane l
// draws window rows l, l+32, ...; the lanes that own those universe rows fetch them by shuffle.
// One warp per instance; batch-wide flags / master reset resolved by the last warp to retire.
template <int WPR, class Rule, int C, int G>
__global__ void __launch_bounds__(128, fused_min_ctas(WPR))
step_random_kernel(const __grid_constant__ StepParams p, uint2 key, uint32_t step,
                   uint32_t threshold) {
    constexpr int WORDS = WPR * WPR;
    constexpr int AW = G * WPR;                     // window rows
    constexpr int SLOTS = (AW + 31) / 32;           // window rows drawn per lane
    __shared__ unsigned int s_done;
    __shared__ int s_flag[3];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long inst = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (threadIdx.x == 0) { s_done = 0u; s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0; }
    __syncthreads();
    bool not_one = false, any = false;
    if (inst < p.n) {
        const Rule rule(p);
        uint32_t x[WPR][WPR];
        load_state<WPR>(x, p.in + inst * (32LL * WORDS) + (long long)lane * WORDS);
        const long long entry = p.raw_inst_stride ? inst : 0;      // batch-1: one shared action
        uint32_t drawn[SLOTS][C];
        uint32_t all = 0xFFFFFFFFu, seen = 0u;
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int r = sl * 32 + lane;
                drawn[sl][c] = (r < AW) ? random_chunk(entry, (uint32_t)r, c, step, key, threshold) : 0u;
                if (r < AW) { all &= drawn[sl][c]; seen |= drawn[sl][c]; }
            }
        not_one = __any_sync(0xFFFFFFFFu, all != 0xFFFFFFFFu);
        any = __any_sync(0xFFFFFFFFu, seen != 0u);
        const int my_group = lane - p.row0 / WPR;
        const bool owner = my_group >= 0 && my_group < G;
        const int bit0 = p.col0 - 32 * p.aw0;
#pragma unroll
        for (int s = 0; s < WPR; ++s) {
            const int r = (owner ? my_group : 0) * WPR + s;        // window row of x[s][*]
            uint32_t mine[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t m = 0u;
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {
                    const uint32_t got = __shfl_sync(0xFFFFFFFFu, drawn[sl][c], r & 31);
                    if ((r >> 5) == sl) m = got;
                }
                mine[c] = owner ? m : 0u;
            }
            uint32_t word[C + 1];
#pragma unroll
            for (int c = 0; c <= C; ++c) {
                const uint32_t cur = (c < C) ? mine[c] : 0u;
                const uint32_t prv = (c > 0) ? mine[c - 1] : 0u;
                word[c] = bit0 ? ((cur << bit0) | (prv >> (32 - bit0))) : cur;
            }
#pragma unroll
            for (int w = 0; w < WPR; ++w)
#pragma unroll
                for (int c = 0; c <= C; ++c)
                    if (w == p.aw0 + c) x[s][w] ^= word[c];
        }
        generation<WPR>(x, rule, (lane + 31) & 31, (lane + 1) & 31);
        if (p.red) instance_sums<WPR>(p, x, lane, p.red + inst * 4);
        store_state<WPR>(x, p.out + inst * (32LL * WORDS) + (long long)lane * WORDS);
        if (p.reward_zero && lane == 0) p.reward_zero[inst] = 0.f;
        if (p.obs) emit_obs_any<WORDS>(p, &x[0][0], inst * (1024LL * WORDS), lane);
    }
    if (retire_legacy<uint8_t>(p, &s_done, s_flag, lane, warps_per_block, not_one, any, false) == 2)
        clear_after_reset(p, lane);
}

// grid-aligned packed action -> float32 [B][AW][AH] (the reference's action format)
static __global__ void __launch_bounds__(256)
unpack_action_kernel(const uint32_t* __restrict__ packed, float* __restrict__ action,
                     long long total, int aw, int ah, int awpr, int bit0) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / ah;                      // entry * aw + r
        const int c = (int)(i - row * ah), bit = c + bit0;
        const uint32_t word = packed[row * awpr + (bit >> 5)];
        action[i] = (float)((word >> (bit & 31)) & 1u);
    }
}

// =========================================================================================
// standalone reductions
// =========================================================================================
// grid = (blocks_per_instance, N); out pre-zeroed; int64 [N][4]
static __global__ void __launch_bounds__(256)
reduce_kernel(const StepParams p, const uint32_t* __restrict__ state,
              unsigned long long* __restrict__ out) {
    const long long inst = blockIdx.y;
    const long long words_per_inst = (long long)p.h * p.wpr;
    const uint32_t* base = state + inst * words_per_inst;
    unsigned long long live = 0, sh = 0, sw = 0, wl = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_inst;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / p.wpr), w = (int)(i - (long long)r * p.wpr);
        const uint32_t v = base[i];
        const bool in_rows = (r >= p.row0) && (r < p.row0 + p.aw);
        const uint32_t inside = in_rows ? (v & window_col_mask(p, w)) : 0u;
        const uint32_t outside = v ^ inside;
        const uint32_t c = ca::popc32(outside);
        live += ca::popc32(v);
        wl += ca::popc32(inside);
        sh += (unsigned long long)r * c;
        sw += 32ull * w * c + ca::bit_index_sum(outside);
    }
    __shared__ unsigned long long acc[4];
    if (threadIdx.x < 4) acc[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        live += __shfl_xor_sync(0xFFFFFFFFu, live, off);
        sh += __shfl_xor_sync(0xFFFFFFFFu, sh, off);
        sw += __shfl_xor_sync(0xFFFFFFFFu, sw, off);
        wl += __shfl_xor_sync(0xFFFFFFFFu, wl, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], live); atomicAdd(&acc[1], sh);
        atomicAdd(&acc[2], sw);   atomicAdd(&acc[3], wl);
    }
    __syncthreads();
    if (threadIdx.x < 4 && acc[threadIdx.x])
        atomicAdd(out + inst * 4 + threadIdx.x, acc[threadIdx.x]);
}

// out[n] = popc(state & plus) - popc(state & minus); grid = (blocks_per_instance, N)
static __global__ void __launch_bounds__(256)
masked_count_kernel(const uint32_t* __restrict__ state, const uint32_t* __restrict__ plus,
                    const uint32_t* __restrict__ minus, long long words_per_inst,
                    long long* __restrict__ out) {
    const long long inst = blockIdx.y;
    const uint32_t* base = state + inst * words_per_inst;
    long long acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_inst;
         i += (long long)gridDim.x * blockDim.x) {
        const uint32_t v = base[i];
        if (plus) acc += ca::popc32(v & plus[i]);
        if (minus) acc -= ca::popc32(v & minus[i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, off);
    __shared__ long long blk;
    if (threadIdx.x == 0) blk = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && acc)
        atomicAdd(reinterpret_cast<unsigned long long*>(&blk), (unsigned long long)acc);
    __syncthreads();
    if (threadIdx.x == 0 && blk)
        atomicAdd(reinterpret_cast<unsigned long long*>(out + inst), (unsigned long long)blk);
}

// SpeedDetector tail (carle/mcl.py:777-795) in ONE launch: per instance the centre of mass
// (Sh, Sw) / (live + 1e-7) in float32 exactly as the reference computes it, the batch-wide
// speed = || com_prev - com ||_2 (squares accumulated in double, so the result does not depend
// on the reduction order), and reward += speed by the last block to retire.
// scratch: handle-owned {double acc; unsigned retired;} kept zero between launches.
static __global__ void __launch_bounds__(256)
speed_tail_kernel(const long long* __restrict__ red, float* __restrict__ com, long long n,
                  int have_prev, float* __restrict__ velocity, float* __restrict__ speed_out,
                  float* __restrict__ reward, double* __restrict__ sumsq_out, int* primed,
                  double* acc, unsigned int* retired) {
    pdl_launch_dependents();
    pdl_wait();                                     // the step kernel's sums are final
    // `primed` (device flag, optional): "a previous centre of mass exists" decided on the device,
    // so the SAME launch serves the wrapper's first step and all later ones (CUDA-graph replays);
    // set by the last block, i.e. after every block has read it
    if (primed) have_prev = *reinterpret_cast<volatile int*>(primed) != 0;
    double local = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const longlong2 a = *reinterpret_cast<const longlong2*>(red + 4 * i);      // live, sh
        const float denom = (float)a.x + 1e-7f;                                    // mcl.py:777
        const float ch = (float)a.y / denom, cw = (float)red[4 * i + 2] / denom;
        if (have_prev) {
            const float vh = com[i] - ch, vw = com[n + i] - cw;                    // mcl.py:787
            if (velocity) { velocity[i] = vh; velocity[n + i] = vw; }
            local += (double)vh * vh + (double)vw * vw;
        }
        com[i] = ch;
        com[n + i] = cw;
    }
    if (!have_prev && !primed) return;
    __shared__ double part[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    __shared__ float s_speed;
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        double blk = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) blk += part[w];
        if (have_prev) atomicAdd(acc, blk);
        __threadfence();
        s_last = (atomicAdd(retired, 1u) == gridDim.x - 1) ? 1 : 0;
        if (s_last) {
            __threadfence();
            if (have_prev) {
                const double total = *reinterpret_cast<volatile double*>(acc);
                s_speed = sqrtf((float)total);                                     // mcl.py:789
                *speed_out = s_speed;
                if (sumsq_out) *sumsq_out = total;
            }
            *acc = 0.0;
            *retired = 0u;
            if (primed) *primed = 1;
        }
    }
    __syncthreads();
    if (s_last && reward && have_prev) {                                               // mcl.py:795
        // one block updates the whole reward column: 16-byte accesses, four in flight per thread
        const float sp = s_speed;
        long long i0 = 0;
        if ((reinterpret_cast<unsigned long long>(reward) & 15ull) == 0ull) {
            float4* r4 = reinterpret_cast<float4*>(reward);
            const long long n4 = n >> 2;
#pragma unroll 4
            for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
                float4 v = r4[i];
                v.x += sp; v.y += sp; v.z += sp; v.w += sp;
                r4[i] = v;
            }
            i0 = n4 << 2;
        }
        for (long long i = i0 + threadIdx.x; i < n; i += blockDim.x) reward[i] += sp;
    }
}

// out[b] = popcount of action entry b; one warp per entry
static __global__ void __launch_bounds__(256)
action_count_kernel(const uint32_t* __restrict__ packed, long long batch,
                    long long words_per_entry, long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < batch; b += nwarps) {
        uint32_t c = 0;
        for (long long i = lane; i < words_per_entry; i += 32)
            c += ca::popc32(packed[b * words_per_entry + i]);
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) out[b] = c;
    }
}

}  // namespace carle
