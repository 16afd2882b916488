// abi_internal.h — shared between the translation units of libcarle_b200.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "kernels.cuh"

struct carle_ctx {
    int device;
    int64_t n;
    int h, w, wpr;
    int row0, col0, aw, ah;
    int aw0, awpr;            // grid-aligned packed action: first word, words per row
    int family;               // 0 generic, 1 warp-resident, 2 tiled (W % 32 == 0, beyond 256)
    int band_row0, band_rows, halo;   // row band of a giant grid (halo == 0: whole grid)
    int grid_h;               // height of the WHOLE grid (== h unless this is a band)
    uint32_t birth, survive;
    int rule_id;
    int sm_count;
    unsigned int* retire;     // device scratch (16 words): [8..9] double accumulator and [10] block
                              // counter of carle_speed_tail, [4..5] 64-bit retirement word of the
                              // persistent fused kernels, [0] block-retirement counter of the others,
                              // [2..3] batch-wide flags of the fused step (kept zero between calls),
                              // [6] / [11] "some action element is neither 0 nor 1" of the
                              // non-persistent fused kernels / of carle_pack_action (zero between
                              // calls), [7] "the last step cleared the universe"
    uint32_t* act_scratch;    // packed action for the unfused fallback of carle_step_action
    size_t act_scratch_words;
    unsigned int* strip_scratch;   // strip kernel: uint64 [N][2] sum accumulators (or NULL)
    int strip_u;
    int defer_reset;          // set by carle_step_ex around its unfused launches
};

namespace carle {

// carle_abi.cu: records the thread-local message behind carle_last_error() and returns `code`
int abi_fail(int code, const std::string& msg);

constexpr uint32_t kLifeB = 0x008, kLifeS = 0x00C;          // B3/S23
constexpr uint32_t kMorleyB = 0x148, kMorleyS = 0x034;      // B368/S245
constexpr uint32_t kHighB = 0x048, kHighS = 0x00C;          // B36/S23
constexpr uint32_t kDayNightB = 0x1C8, kDayNightS = 0x1D8;  // B3678/S34678

constexpr int kMaxDevices = 64;     // per-device caches of launch configurations

enum RuleId { RULE_DYNAMIC = 0, RULE_LIFE, RULE_MORLEY, RULE_HIGHLIFE, RULE_DAYNIGHT };

// environment switches for A/B measurements (re-read on every call: a getenv is noise next to a
// launch, and one test process can exercise every variant)
inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
inline bool pdl_enabled() { return env_int("CARLE_PDL", 1) != 0; }

// warp ranking of the persistent kernels (StepParams::rank_blocked): blocked when every warp makes
// many trips, interleaved otherwise; CARLE_RANK=blocked|interleaved forces one (A/B runs)
inline int rank_blocked_for(long long units, long long nwarps) {
    const char* e = getenv("CARLE_RANK");
    if (e && e[0] == 'b') return 1;
    if (e && e[0] == 'i') return 0;
    return units >= 8 * nwarps ? 1 : 0;
}

// StepParams::reverse / act_evict_first of a persistent one-launch step from p.in to p.out: a rollout
// ping-pongs two buffers, so "walk backwards when stepping from the higher buffer to the lower one"
// alternates the direction from step to step without any state (CARLE_REVERSE=0 switches it off for A/B
// runs: 0.7 - 1.3 us of ~100 at 16384 x 256x256 and 131072 x 64x64).  The evict-first policy on the action
// copies measured SLOWER (+2 to +3 us on the same shapes, profiles/r2c_ab_walk_and_features.txt) and is off
// unless CARLE_ACT_EVICT_FIRST=1.
inline void walk_policy(StepParams& q) {
    q.reverse = (env_int("CARLE_REVERSE", 1) != 0 &&
                 reinterpret_cast<uintptr_t>(q.in) > reinterpret_cast<uintptr_t>(q.out)) ? 1 : 0;
    q.act_evict_first = env_int("CARLE_ACT_EVICT_FIRST", 0) != 0 ? 1 : 0;
}

// strip_abi.cu: one env step with the strip kernel (strip.cuh).  `shape` as fused_shape():
// 2 = 128x128 / 32x32 window, 3 = 256x256 / 64x64 window; rows_per_lane in {2, 4} (shape 3).
// Returns cudaErrorInvalidValue for an unsupported combination.
cudaError_t launch_strip(int device, int rule_id, int shape, int rows_per_lane, int sm_count,
                         bool pdl, const StepParams& p, cudaStream_t s);

// stream_abi.cu: the persistent TMA-staged one-launch step (float32 / uint8 actions); shape as
// fused_shape(): 1 = 64x64 / 32x32 window, 2 = 128x128 / 32x32, 3 = 256x256 / 64x64.
cudaError_t launch_stream(int device, int rule_id, int shape, int sm_count, bool pdl,
                          const StepParams& p, cudaStream_t s);

// fused_abi.cu: the non-persistent one-launch variants (A/B runs, unaligned action pointers)
cudaError_t launch_fused(int rule_id, int shape, const StepParams& p, cudaStream_t s);
cudaError_t launch_quad(int rule_id, int sm_count, const StepParams& p, cudaStream_t s);
cudaError_t launch_random_direct(int rule_id, int shape, const StepParams& p, uint2 key, uint32_t step,
                                 uint32_t thr, cudaStream_t s);

// random_abi.cu: one env step whose action is the device-side random agent, through the persistent
// stream kernel (shape 1: 64x64 / 32x32 window, 2: 128x128 / 32x32).  p.rand_* must be set.
cudaError_t launch_stream_random(int device, int rule_id, int shape, int sm_count, bool pdl,
                                 const StepParams& p, cudaStream_t s);

// jit.cu: NVRTC specialisation of the step kernels for arbitrary rules.
//  jit_kernel   -> driver function handle for `instantiation` on `device` (compiled once, cached),
//                  or nullptr when the JIT is disabled / unavailable / failed.
//  jit_launch_grid -> plain launch of `blocks` CTAs; `params` (and `params2`, for kernels with a
//                  second one) point at the kernel's __grid_constant__ parameter structs.
//  jit_launch   -> persistent launch: grid = min(SMs * occupancy, max_blocks) rounded down to a
//                  multiple of block_multiple.
//  jit_compile  -> compile only (no GPU needed); 0 on success.
bool jit_enabled();
int jit_loaded();
void* jit_kernel(int device, const std::string& instantiation);
cudaError_t jit_launch(void* function, int sm_count, int threads, size_t smem, long long max_blocks,
                       int block_multiple, bool pdl, StepParams p, long long units, cudaStream_t s,
                       const void* params2 = nullptr);
cudaError_t jit_launch_grid(void* function, long long blocks, int threads, size_t smem, bool pdl,
                            const void* params, cudaStream_t s, const void* params2 = nullptr);
int jit_compile(const char* instantiation, std::vector<char>* cubin, std::string* lowered,
                std::string* log);

}  // namespace carle
