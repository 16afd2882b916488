// carle_abi.cu — extern "C" boundary of libcarle_b200.so (see include/carle_b200.h).
// Host-side only: argument validation mirroring the reference's geometry arithmetic
// (carle/env.py:119-132), kernel selection and launches.  No torch types, no allocation
// on the step path, no CPU implementation of the update: without a CUDA device every
// entry point fails with CARLE_ENODEV / CARLE_ECUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <type_traits>

#include "../../include/carle_b200.h"
#include "abi_internal.h"
#include "stream_launch.h"
#include "tiled.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess)                                                      \
            return fail(CARLE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// switch to the handle's device for the duration of a call, then restore the caller's
// (torch tracks the current device through the runtime; do not leave it changed)
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define DEVICE_GUARD(h)                                                        \
    DeviceGuard _guard((h)->device);                                           \
    if (_guard.err != cudaSuccess)                                             \
        return fail(CARLE_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_guard.err))

using carle::kLifeB; using carle::kLifeS; using carle::kMorleyB; using carle::kMorleyS;
using carle::kHighB; using carle::kHighS; using carle::kDayNightB; using carle::kDayNightS;
using carle::RULE_DYNAMIC; using carle::RULE_LIFE; using carle::RULE_MORLEY;
using carle::RULE_HIGHLIFE; using carle::RULE_DAYNIGHT;

using carle::env_int;
using carle::pdl_enabled;

}  // namespace

namespace carle {
int abi_fail(int code, const std::string& msg) { return fail(code, msg); }
}

// (struct carle_ctx: abi_internal.h)

namespace {

carle::StepParams base_params(const carle_ctx* c) {
    carle::StepParams p;
    memset(&p, 0, sizeof(p));
    p.n = c->n;
    p.k = 1;
    p.h = c->h; p.w = c->w; p.wpr = c->wpr;
    p.row0 = c->row0; p.col0 = c->col0; p.aw = c->aw; p.ah = c->ah;
    p.aw0 = c->aw0; p.awpr = c->awpr;
    p.birth = c->birth; p.survive = c->survive;
    p.masks = ca::expand_rule(c->birth, c->survive);
    p.retire = c->retire;
    p.retire64 = reinterpret_cast<unsigned long long*>(c->retire + 4);
    p.strip_part = c->strip_scratch;
    p.defer_reset = c->defer_reset;
    return p;
}

// device of the handle a launch belongs to (for the per-device cache of NVRTC kernels); set by
// the entry points right before they dispatch
thread_local int t_device = 0;

template <int WPR, class Rule>
cudaError_t launch_warp(const carle::StepParams& p, cudaStream_t s) {
    const int warps_per_block = 4;
    long long blocks = (p.n + warps_per_block - 1) / warps_per_block;
    if (blocks > (1LL << 30)) blocks = 1LL << 30;
    if constexpr (std::is_same<Rule, carle::DynamicRule>::value) {
        char inst[128];
        snprintf(inst, sizeof inst, "carle::step_warp_kernel<%d, carle::StaticRule<%uu, %uu>>", WPR,
                 p.birth, p.survive);
        if (void* fn = carle::jit_kernel(t_device, inst))
            return carle::jit_launch_grid(fn, blocks, warps_per_block * 32, 0, false, &p, s);
    }
    carle::step_warp_kernel<WPR, Rule><<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(p);
    return cudaGetLastError();
}

template <class Rule>
cudaError_t launch_warp_wpr(int wpr, const carle::StepParams& p, cudaStream_t s) {
    switch (wpr) {
        case 1: return launch_warp<1, Rule>(p, s);
        case 2: return launch_warp<2, Rule>(p, s);
        case 3: return launch_warp<3, Rule>(p, s);
        case 4: return launch_warp<4, Rule>(p, s);
        case 5: return launch_warp<5, Rule>(p, s);
        case 6: return launch_warp<6, Rule>(p, s);
        case 7: return launch_warp<7, Rule>(p, s);
        case 8: return launch_warp<8, Rule>(p, s);
    }
    return cudaErrorInvalidValue;
}

// the supported fused shapes: 64x64/32 (cfg 1, 4), 128x128/32 (cfg 2), 256x256/64 (cfg 3,
// the reference's defaults)
inline int fused_shape(int wpr, int aw, int ah) {
    if (wpr == 2 && aw == 32 && ah == 32) return 1;
    if (wpr == 4 && aw == 32 && ah == 32) return 2;
    if (wpr == 8 && aw == 64 && ah == 64) return 3;
    return 0;
}

template <class Rule>
cudaError_t launch_generic(const carle_ctx* c, const carle::StepParams& p, cudaStream_t s) {
    long long total = p.n * (long long)p.h * p.wpr;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)c->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if constexpr (std::is_same<Rule, carle::DynamicRule>::value) {
        char inst[128];
        snprintf(inst, sizeof inst, "carle::step_generic_kernel<carle::StaticRule<%uu, %uu>>", p.birth,
                 p.survive);
        if (void* fn = carle::jit_kernel(c->device, inst))
            return carle::jit_launch_grid(fn, blocks, 256, 0, false, &p, s);
    }
    carle::step_generic_kernel<Rule><<<(unsigned)blocks, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_step(const carle_ctx* c, const carle::StepParams& p, cudaStream_t s) {
    using namespace carle;
    t_device = c->device;
    if (c->family == 1) {
        switch (c->rule_id) {
            case RULE_LIFE: return launch_warp_wpr<StaticRule<kLifeB, kLifeS>>(c->wpr, p, s);
            case RULE_MORLEY: return launch_warp_wpr<StaticRule<kMorleyB, kMorleyS>>(c->wpr, p, s);
            case RULE_HIGHLIFE: return launch_warp_wpr<StaticRule<kHighB, kHighS>>(c->wpr, p, s);
            case RULE_DAYNIGHT: return launch_warp_wpr<StaticRule<kDayNightB, kDayNightS>>(c->wpr, p, s);
            default: return launch_warp_wpr<DynamicRule>(c->wpr, p, s);
        }
    }
    switch (c->rule_id) {
        case RULE_LIFE: return launch_generic<StaticRule<kLifeB, kLifeS>>(c, p, s);
        case RULE_MORLEY: return launch_generic<StaticRule<kMorleyB, kMorleyS>>(c, p, s);
        case RULE_HIGHLIFE: return launch_generic<StaticRule<kHighB, kHighS>>(c, p, s);
        case RULE_DAYNIGHT: return launch_generic<StaticRule<kDayNightB, kDayNightS>>(c, p, s);
        default: return launch_generic<DynamicRule>(c, p, s);
    }
}

template <class Rule, int R>
cudaError_t launch_tiled_rule(const carle_ctx* c, const carle::TiledParams& tp, cudaStream_t s) {
    const long long tiles = tp.s.n * (long long)tp.tiles_y * tp.tiles_x;
    long long blocks = (tiles + 3) / 4;
    // persistent: as many 4-warp CTAs per SM as the register file holds
    const long long cap = (long long)c->sm_count * carle::tiled_min_ctas(R);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if constexpr (std::is_same<Rule, carle::DynamicRule>::value) {
        char inst[128];
        snprintf(inst, sizeof inst, "carle::step_tiled_kernel<carle::StaticRule<%uu, %uu>, %d>", tp.s.birth,
                 tp.s.survive, R);
        if (void* fn = carle::jit_kernel(c->device, inst))
            return carle::jit_launch_grid(fn, blocks, 128, carle::tiled_smem_bytes(R), false, &tp, s);
    }
    auto kernel = carle::step_tiled_kernel<Rule, R>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         carle::tiled_smem_bytes(R));
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)blocks, 128, carle::tiled_smem_bytes(R), s>>>(tp);
    return cudaGetLastError();
}

template <class Rule>
cudaError_t launch_tiled_r(const carle_ctx* c, const carle::TiledParams& tp, int r, cudaStream_t s) {
    return r == 4 ? launch_tiled_rule<Rule, 4>(c, tp, s) : launch_tiled_rule<Rule, 8>(c, tp, s);
}

// rows per lane of the tiled family's register tiles: 8 (256-row tiles, 8 warps per SM) or 4
// (128-row tiles, 16 warps per SM); CARLE_TILE_R overrides the default
int tile_rows_per_lane() {
    static const int r = [] {
        const char* e = getenv("CARLE_TILE_R");
        const int v = e ? atoi(e) : 8;
        return v == 4 ? 4 : 8;
    }();
    return r;
}

// fills the tile grid of `tp` (tv, out_rows and s.k set) for tiles of 32*r rows
void set_tile_grid(carle::TiledParams& tp, int r, int wpr, int generations) {
    const int interior = 32 * r - 2 * tp.tv;
    tp.tiles_y = (tp.out_rows + interior - 1) / interior;
    tp.xstride = (generations <= 16 && wpr >= 8) ? 7 : 6;
    tp.tiles_x = (wpr + tp.xstride - 1) / tp.xstride;
}

cudaError_t launch_tiled(const carle_ctx* c, carle::TiledParams& tp, cudaStream_t s) {
    using namespace carle;
    // 128-row tiles need a halo that leaves an interior: tv < 64
    const int r = (tile_rows_per_lane() == 4 && tp.tv <= 32) ? 4 : 8;
    set_tile_grid(tp, r, c->wpr, tp.s.k);
    switch (c->rule_id) {
        case RULE_LIFE: return launch_tiled_r<StaticRule<kLifeB, kLifeS>>(c, tp, r, s);
        case RULE_MORLEY: return launch_tiled_r<StaticRule<kMorleyB, kMorleyS>>(c, tp, r, s);
        case RULE_HIGHLIFE: return launch_tiled_r<StaticRule<kHighB, kHighS>>(c, tp, r, s);
        case RULE_DAYNIGHT: return launch_tiled_r<StaticRule<kDayNightB, kDayNightS>>(c, tp, r, s);
        default: return launch_tiled_r<DynamicRule>(c, tp, r, s);
    }
}

// generations per temporal block of the tiled family (halo rows = this rounded up to 8)
int tile_block_generations() {
    static const int t = [] {
        const char* e = getenv("CARLE_TILE_T");
        int v = e ? atoi(e) : 16;
        return v < 1 ? 1 : (v > 32 ? 32 : v);
    }();
    return t;
}

int grid_for(long long work_items, int per_block, int sm_count, int waves = 16) {
    long long blocks = (work_items + per_block - 1) / per_block;
    long long cap = (long long)sm_count * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_reduce(const carle_ctx* c, const uint32_t* state, int64_t* out, cudaStream_t s) {
    carle::StepParams p = base_params(c);
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int64_t) * 4 * c->n, s));
    long long words = (long long)c->h * c->wpr;
    int bx = (int)((words + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 4096) bx = 4096;
    long long done = 0;
    while (done < c->n) {                       // gridDim.y <= 65535
        long long chunk = c->n - done;
        if (chunk > 65535) chunk = 65535;
        dim3 grid(bx, (unsigned)chunk);
        carle::reduce_kernel<<<grid, 256, 0, s>>>(
            p, state + done * words, reinterpret_cast<unsigned long long*>(out) + done * 4);
        CUDA_TRY(cudaGetLastError());
        done += chunk;
    }
    return CARLE_OK;
}

}  // namespace

extern "C" {

CARLE_API int carle_version(void) { return 100; }

CARLE_API const char* carle_last_error(void) { return g_err.c_str(); }

CARLE_API int carle_create(carle_handle_t* out, int device, int64_t instances, int height, int width,
                 int action_height, int action_width) {
    if (!out) return fail(CARLE_EINVAL, "carle_create: out is NULL");
    *out = nullptr;
    if (instances < 1 || height < 1 || width < 1 || action_height < 0 || action_width < 0)
        return fail(CARLE_EINVAL, "carle_create: non-positive size");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(CARLE_ENODEV, "carle_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= count)
        return fail(CARLE_EINVAL, "carle_create: device index out of range");
    // carle/env.py:119-132 — same arithmetic, including the axis swap in ZeroPad2d
    const int asym_w = (width - action_width) % 2, asym_h = (height - action_height) % 2;
    const int aw = action_width - (width % 2), ah = action_height - (height % 2);
    const int wp = (width - aw) / 2, hp = (height - ah) / 2;
    if (aw < 0 || ah < 0 || wp < 0 || hp < 0 || asym_w < 0 || asym_h < 0)
        return fail(CARLE_EINVAL, "carle_create: action window larger than the grid");
    const int rows = aw + 2 * wp + asym_w, cols = ah + 2 * hp + asym_h;
    if (rows != height || cols != width) {
        char buf[160];
        snprintf(buf, sizeof buf,
                 "carle_create: padded action is %dx%d but the universe is %dx%d "
                 "(the reference only works for even, square grids)", rows, cols, height, width);
        return fail(CARLE_EINVAL, buf);
    }
    carle_ctx* c = new carle_ctx();
    c->device = device;
    c->n = instances;
    c->h = height; c->w = width; c->wpr = (width + 31) / 32;
    c->row0 = wp; c->col0 = hp; c->aw = aw; c->ah = ah;
    c->aw0 = hp / 32;
    c->awpr = (ah > 0) ? ((hp + ah - 1) / 32 - c->aw0 + 1) : 1;
    c->family = (width % 32 == 0 && height == width && width <= 256) ? 1 : (width % 32 == 0 ? 2 : 0);
    c->band_row0 = 0; c->band_rows = height; c->halo = 0; c->grid_h = height;
    c->birth = kLifeB; c->survive = kLifeS; c->rule_id = RULE_LIFE;   // carle/env.py:58-59
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return fail(CARLE_ECUDA, "carle_create: cudaGetDeviceProperties failed");
    }
    c->sm_count = prop.multiProcessorCount;
    c->act_scratch = nullptr; c->act_scratch_words = 0;
    c->strip_scratch = nullptr; c->strip_u = 0;
    c->defer_reset = 0;
    {
        DeviceGuard guard(device);
        if (guard.err != cudaSuccess || cudaMalloc(&c->retire, 16 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(c->retire, 0, 16 * sizeof(unsigned int)) != cudaSuccess) {
            delete c;
            return fail(CARLE_ECUDA, "carle_create: cannot allocate the handle's device scratch");
        }
        // strip kernel (strip.cuh): per-instance sum accumulators, allocated here so that no step
        // call ever allocates
        const int shape = (c->family == 1) ? fused_shape(c->wpr, c->aw, c->ah) : 0;
        if (shape == 2 || shape == 3) {
            c->strip_u = c->wpr / 2;
            const size_t bytes = (size_t)c->n * 16;     // two 64-bit accumulators per instance
            if (cudaMalloc(&c->strip_scratch, bytes) != cudaSuccess ||
                cudaMemset(c->strip_scratch, 0, bytes) != cudaSuccess) {
                cudaFree(c->retire);
                delete c;
                return fail(CARLE_ECUDA, "carle_create: cannot allocate the strip scratch");
            }
        }
    }
    *out = c;
    return CARLE_OK;
}

CARLE_API int carle_destroy(carle_handle_t h) {
    if (h && h->retire) {
        DeviceGuard guard(h->device);
        cudaFree(h->retire);
        if (h->act_scratch) cudaFree(h->act_scratch);
        if (h->strip_scratch) cudaFree(h->strip_scratch);
    }
    delete h;
    return CARLE_OK;
}

CARLE_API int carle_geometry(carle_handle_t h, int32_t geo[8]) {
    if (!h || !geo) return fail(CARLE_EINVAL, "carle_geometry: NULL argument");
    geo[0] = h->row0; geo[1] = h->col0; geo[2] = h->aw; geo[3] = h->ah;
    geo[4] = h->wpr; geo[5] = h->awpr; geo[6] = h->family; geo[7] = h->aw0;
    return CARLE_OK;
}

CARLE_API int carle_set_rule(carle_handle_t h, uint32_t birth_mask, uint32_t survive_mask) {
    if (!h) return fail(CARLE_EINVAL, "carle_set_rule: NULL handle");
    birth_mask &= 0x1FFu; survive_mask &= 0x1FFu;
    if (birth_mask == 0 || survive_mask == 0)
        return fail(CARLE_ERULE, "carle_set_rule: empty birth or survive set "
                                 "(reduce() of empty sequence with no initial value)");
    h->birth = birth_mask; h->survive = survive_mask;
    if (birth_mask == kLifeB && survive_mask == kLifeS) h->rule_id = RULE_LIFE;
    else if (birth_mask == kMorleyB && survive_mask == kMorleyS) h->rule_id = RULE_MORLEY;
    else if (birth_mask == kHighB && survive_mask == kHighS) h->rule_id = RULE_HIGHLIFE;
    else if (birth_mask == kDayNightB && survive_mask == kDayNightS) h->rule_id = RULE_DAYNIGHT;
    else h->rule_id = RULE_DYNAMIC;
    return CARLE_OK;
}

CARLE_API int carle_pack_state(carle_handle_t h, const void* cells, int dtype, uint32_t* packed,
                     void* stream) {
    if (!h || !cells || !packed) return fail(CARLE_EINVAL, "carle_pack_state: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long rows = h->n * h->h, total = rows * h->wpr;
    const int grid = grid_for(total, 8 * 32, h->sm_count);
    if (dtype == CARLE_F32)
        carle::pack_state_kernel<float><<<grid, 256, 0, s>>>(
            static_cast<const float*>(cells), packed, rows, h->w, h->wpr);
    else if (dtype == CARLE_U8)
        carle::pack_state_kernel<uint8_t><<<grid, 256, 0, s>>>(
            static_cast<const uint8_t*>(cells), packed, rows, h->w, h->wpr);
    else
        return fail(CARLE_EINVAL, "carle_pack_state: dtype must be CARLE_F32 or CARLE_U8");
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_unpack_state(carle_handle_t h, const uint32_t* packed, void* cells, int dtype,
                       void* stream) {
    if (!h || !cells || !packed) return fail(CARLE_EINVAL, "carle_unpack_state: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long rows = h->n * h->h, total = rows * h->wpr;
    const int grid = grid_for(total, 8 * 32, h->sm_count);
    if (dtype == CARLE_F32 && h->w % 32 == 0 &&
        (reinterpret_cast<uintptr_t>(cells) & 15u) == 0)
        carle::unpack_state_f32_kernel<<<grid, 256, 0, s>>>(
            packed, static_cast<float4*>(cells), total);
    else if (dtype == CARLE_F32)
        carle::unpack_state_kernel<float><<<grid, 256, 0, s>>>(
            packed, static_cast<float*>(cells), rows, h->w, h->wpr);
    else if (dtype == CARLE_U8)
        carle::unpack_state_kernel<uint8_t><<<grid, 256, 0, s>>>(
            packed, static_cast<uint8_t*>(cells), rows, h->w, h->wpr);
    else
        return fail(CARLE_EINVAL, "carle_unpack_state: dtype must be CARLE_F32 or CARLE_U8");
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_pack_action(carle_handle_t h, const void* action, int dtype, int64_t batch,
                      int64_t steps, uint32_t* packed_action, int32_t* flags, void* stream) {
    if (!h || !action || !flags)
        return fail(CARLE_EINVAL, "carle_pack_action: NULL argument");
    if (batch != 1 && batch != h->n)
        return fail(CARLE_EINVAL, "carle_pack_action: action batch must be 1 or N");
    if (steps < 1) return fail(CARLE_EINVAL, "carle_pack_action: steps < 1");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long rows_per_step = batch * h->aw;
    const long long rows = steps * rows_per_step;
    // zero-sized window: nothing to toggle.  The flags are left untouched, and zeroed flags read
    // as "every toggle is 1.0": callers pass NULL flags to the step then (the mean of an empty
    // tensor is NaN upstream: no reset)
    if (rows == 0 || h->ah == 0) return CARLE_OK;
    const long long total = rows * h->awpr;
    if (dtype == CARLE_PACKED) {
        carle::packed_action_flags_kernel<<<grid_for(total, 256, h->sm_count), 256, 0, s>>>(
            static_cast<const uint32_t*>(action), flags, rows, rows_per_step, h->ah, h->awpr,
            h->col0 - 32 * h->aw0);
    } else {
        if (!packed_action) return fail(CARLE_EINVAL, "carle_pack_action: packed_action is NULL");
        // streaming fast path: window width a power-of-two multiple of 32
        int cshift = -1;
        for (int k = 0; k < 6; ++k) if (h->ah == (32 << k)) cshift = k;
        if (cshift >= 0 && (dtype == CARLE_F32 || dtype == CARLE_U8)) {
            const long long chunks_per_step = rows_per_step << cshift;
            long long bx = (chunks_per_step + 8 * 32 - 1) / (8 * 32);   // 8 warps x 32 chunks
            long long cap = ((long long)h->sm_count * 8 + steps - 1) / steps;
            if (bx > cap) bx = cap;
            if (bx < 1) bx = 1;
            long long done = 0;
            while (done < steps) {
                long long chunk = steps - done;
                if (chunk > 65535) chunk = 65535;
                dim3 g((unsigned)bx, (unsigned)chunk);
                const long long in_off = done * rows_per_step * h->ah;
                const long long out_off = done * rows_per_step * h->awpr;
                if (dtype == CARLE_F32)
                    carle::pack_action_stream_kernel<float><<<g, 256, 0, s>>>(
                        static_cast<const float*>(action) + in_off, packed_action + out_off,
                        flags + 2 * done, chunks_per_step, cshift, h->awpr,
                        h->col0 - 32 * h->aw0, h->retire + 11);
                else
                    carle::pack_action_stream_kernel<uint8_t><<<g, 256, 0, s>>>(
                        static_cast<const uint8_t*>(action) + in_off, packed_action + out_off,
                        flags + 2 * done, chunks_per_step, cshift, h->awpr,
                        h->col0 - 32 * h->aw0, nullptr);
                done += chunk;
            }
            CUDA_TRY(cudaGetLastError());
            if (dtype == CARLE_F32) {
                carle::resolve_packed_flags_kernel<<<1, 256, 0, s>>>(
                    static_cast<const float*>(action), flags, steps, rows_per_step * h->ah, h->retire + 11);
                CUDA_TRY(cudaGetLastError());
            }
            return CARLE_OK;
        }
        constexpr int R = 8;
        long long bx = (rows_per_step + 8 * R - 1) / (8 * R);        // 8 warps per block
        long long cap = ((long long)h->sm_count * 16 + steps - 1) / steps;
        if (bx > cap) bx = cap;
        if (bx < 1) bx = 1;
        long long done = 0;
        while (done < steps) {                                      // gridDim.y <= 65535
            long long chunk = steps - done;
            if (chunk > 65535) chunk = 65535;
            dim3 g((unsigned)bx, (unsigned)chunk);
            const long long in_off = done * rows_per_step * h->ah;
            const long long out_off = done * rows_per_step * h->awpr;
            if (dtype == CARLE_F32)
                carle::pack_action_kernel<float, R><<<g, 256, 0, s>>>(
                    static_cast<const float*>(action) + in_off, packed_action + out_off,
                    flags + 2 * done, rows_per_step, h->ah, h->awpr, h->col0 - 32 * h->aw0,
                    h->retire + 11);
            else if (dtype == CARLE_U8)
                carle::pack_action_kernel<uint8_t, R><<<g, 256, 0, s>>>(
                    static_cast<const uint8_t*>(action) + in_off, packed_action + out_off,
                    flags + 2 * done, rows_per_step, h->ah, h->awpr, h->col0 - 32 * h->aw0, nullptr);
            else
                return fail(CARLE_EINVAL, "carle_pack_action: bad dtype");
            done += chunk;
        }
        if (dtype == CARLE_F32) {
            CUDA_TRY(cudaGetLastError());
            carle::resolve_packed_flags_kernel<<<1, 256, 0, s>>>(
                static_cast<const float*>(action), flags, steps, rows_per_step * h->ah, h->retire + 11);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_step_many(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                    uint32_t* scratch, const uint32_t* packed_actions, int64_t action_batch,
                    int64_t steps, int32_t* flags, int64_t* counters,
                    int64_t* reductions, void* stream) {
    if (!h || !state_in || !state_out)
        return fail(CARLE_EINVAL, "carle_step: NULL state pointer");
    if (steps < 1) return fail(CARLE_EINVAL, "carle_step: steps < 1");
    if (packed_actions && action_batch != 1 && action_batch != h->n)
        return fail(CARLE_EINVAL, "carle_step: action batch must be 1 or N");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    carle::StepParams p = base_params(h);
    const long long entry_words = (long long)h->aw * h->awpr;
    if (h->aw == 0 || h->ah == 0) { packed_actions = nullptr; flags = nullptr; }   // empty window
    p.act_inst_stride = (action_batch == 1) ? 0 : entry_words;
    p.act_step_stride = action_batch * entry_words;
    p.counters = reinterpret_cast<long long*>(counters);
    if (h->family == 1) {
        p.in = state_in; p.out = state_out;
        p.act = packed_actions; p.flags = flags;
        p.red = reinterpret_cast<long long*>(reductions);
        p.k = (int)steps;
        CUDA_TRY(launch_step(h, p, s));
        return CARLE_OK;
    }
    if (h->family == 2 && h->halo == 0) {
        // tiled family: blocks of T generations per launch (temporal blocking), ping-pong so
        // the last block lands in state_out.  Per-generation sums force T = 1.
        if (state_in == state_out)
            return fail(CARLE_EINVAL, "carle_step: in-place update needs the warp-resident family");
        const int tmax = reductions ? 1 : tile_block_generations();
        const int64_t nblocks = (steps + tmax - 1) / tmax;
        if (nblocks > 1 && !scratch)
            return fail(CARLE_EINVAL, "carle_step_many: tiled family needs a scratch buffer "
                                      "when the generations do not fit one temporal block");
        const uint32_t* src = state_in;
        int64_t g = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            const int t = (int)((steps - g) < tmax ? (steps - g) : tmax);
            uint32_t* dst = ((nblocks - 1 - b) % 2 == 0) ? state_out : scratch;
            carle::TiledParams tp;
            memset(&tp, 0, sizeof(tp));
            tp.s = p;
            tp.s.in = src; tp.s.out = dst;
            tp.s.act = packed_actions ? packed_actions + g * p.act_step_stride : nullptr;
            tp.s.flags = flags ? flags + 2 * g : nullptr;
            tp.s.k = t;
            tp.tv = (t + 7) / 8 * 8;
            tp.out_row0 = 0; tp.out_rows = h->h; tp.vwrap = 1; tp.act_row_shift = 0;
            tp.grid_h = h->h;
            CUDA_TRY(launch_tiled(h, tp, s));                 // (sets the tile grid)
            if (reductions) {
                int rc = launch_reduce(h, dst, reductions + g * h->n * 4, s);
                if (rc) return rc;
            }
            src = dst;
            g += t;
        }
        return CARLE_OK;
    }
    if (h->halo != 0)
        return fail(CARLE_EINVAL, "carle_step: this handle is a row band; use carle_band_step");
    // generic family: one launch per generation, ping-pong so the last lands in state_out
    if (state_in == state_out)
        return fail(CARLE_EINVAL, "carle_step: in-place update needs the warp-resident family");
    if (steps > 1 && !scratch)
        return fail(CARLE_EINVAL, "carle_step_many: generic family needs a scratch buffer for K > 1");
    const uint32_t* src = state_in;
    for (int64_t g = 0; g < steps; ++g) {
        uint32_t* dst = ((steps - 1 - g) % 2 == 0) ? state_out : scratch;
        p.in = src; p.out = dst;
        p.act = packed_actions ? packed_actions + g * p.act_step_stride : nullptr;
        p.flags = flags ? flags + 2 * g : nullptr;
        p.k = 1;
        CUDA_TRY(launch_step(h, p, s));
        if (reductions) {
            int rc = launch_reduce(h, dst, reductions + g * h->n * 4, s);
            if (rc) return rc;
        }
        src = dst;
    }
    return CARLE_OK;
}

CARLE_API int carle_step(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
               const uint32_t* packed_action, int64_t action_batch, int32_t* flags,
               int64_t* counters, int64_t* reductions, void* stream) {
    return carle_step_many(h, state_in, state_out, nullptr, packed_action, action_batch, 1,
                           flags, counters, reductions, stream);
}

// cells of the whole batch as 32-bit words of an unpacked observation
static long long obs_words_of(const carle_ctx* h, int obs_dtype) {
    const long long cells = h->n * (long long)h->h * h->w;
    return obs_dtype == CARLE_U8 ? cells / 4 : cells;
}

CARLE_API int carle_step_ex(carle_handle_t h, const carle_step_args* args, void* stream) {
    if (!h || !args) return fail(CARLE_EINVAL, "carle_step_ex: NULL argument");
    carle_step_args a;
    memset(&a, 0, sizeof a);
    memcpy(&a, args, args->struct_size < sizeof a ? args->struct_size : sizeof a);
    if (!a.state_in || !a.state_out) return fail(CARLE_EINVAL, "carle_step_ex: NULL state pointer");
    if (a.action && a.action_batch != 1 && a.action_batch != h->n)
        return fail(CARLE_EINVAL, "carle_step_ex: action batch must be 1 or N");
    if (a.action && a.action_dtype != CARLE_F32 && a.action_dtype != CARLE_U8 && a.action_dtype != CARLE_PACKED)
        return fail(CARLE_EINVAL, "carle_step_ex: action dtype must be CARLE_F32, CARLE_U8 or CARLE_PACKED");
    if (a.obs && a.obs_dtype != CARLE_F32 && a.obs_dtype != CARLE_U8)
        return fail(CARLE_EINVAL, "carle_step_ex: obs dtype must be CARLE_F32 or CARLE_U8");
    if (h->halo != 0) return fail(CARLE_EINVAL, "carle_step_ex: this handle is a row band; use carle_band_step");
    const bool sd = a.speed_com_next != nullptr;
    if (sd && (!a.reductions || !a.speed_com_prev || !a.speed_out || !a.speed_primed ||
               a.speed_com_prev == a.speed_com_next))
        return fail(CARLE_EINVAL, "carle_step_ex: the fused SpeedDetector tail needs reductions, two "
                                  "distinct centre-of-mass buffers, speed_out and speed_primed");
    if (sd && a.defer_reset)
        return fail(CARLE_EINVAL, "carle_step_ex: with defer_reset the tail follows carle_apply_reset "
                                  "(carle_speed_tail), it cannot be part of the step");
    // behind a step whose kernel cannot carry the tail: copy the previous centres and run the
    // stand-alone tail kernel in place on them
    auto tail_behind = [&]() -> int {
        DEVICE_GUARD(h);
        CUDA_TRY(cudaMemcpyAsync(a.speed_com_next, a.speed_com_prev, sizeof(float) * 2 * (size_t)h->n,
                                 cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
        return carle_speed_tail(h, a.reductions, a.speed_com_next, 0, a.speed_velocity, a.speed_out,
                                a.reward_zero, a.speed_sumsq, a.speed_primed, stream);
    };
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t* gflags = reinterpret_cast<int32_t*>(h->retire + 2);
    const bool raw = a.action && (a.action_dtype == CARLE_F32 || a.action_dtype == CARLE_U8);
    const bool packed = a.action && a.action_dtype == CARLE_PACKED;
    const int shape = (a.action && h->family == 1 && h->row0 % h->wpr == 0 && h->n < (1LL << 32))
                          ? fused_shape(h->wpr, h->aw, h->ah) : 0;
    const bool obs_in_kernel_ok = !a.obs || (reinterpret_cast<uintptr_t>(a.obs) & 15u) == 0;
    // packed actions ride in the persistent kernels only (bulk copies: 16-byte aligned pointer)
    const bool packed_ok = !packed || ((reinterpret_cast<uintptr_t>(a.action) & 15u) == 0 &&
                                       (shape != 3 || h->strip_scratch));
    if (shape && obs_in_kernel_ok && packed_ok) {
        DEVICE_GUARD(h);
        carle::StepParams p = base_params(h);
        p.in = a.state_in; p.out = a.state_out;
        p.raw = a.action;
        p.raw_u8 = packed ? 2 : (a.action_dtype == CARLE_U8) ? 1 : 0;
        p.raw_inst_stride = (a.action_batch == 1) ? 0
                          : packed ? (long long)h->aw * h->awpr : (long long)h->aw * h->ah;
        p.flags = gflags;
        p.counters = reinterpret_cast<long long*>(a.counters);
        p.red = reinterpret_cast<long long*>(a.reductions);
        p.reward_zero = a.reward_zero;
        p.obs = a.obs;
        p.obs_u8 = (a.obs_dtype == CARLE_U8) ? 1 : 0;
        p.defer_reset = a.defer_reset ? 1 : 0;
        p.k = 1;
        // Kernel choice (A/B switches: CARLE_FUSED_IMPL=direct|tma|quad|strip, CARLE_STRIP_R=2|4,
        // CARLE_STRIP128=1, CARLE_PDL=0).  Measured on B200 (profiles/): the persistent TMA
        // pipeline (step_stream_kernel) reaches the HBM roofline for 64x64 and wins at 128x128;
        // 256x256 runs as independent strips (step_strip_kernel; 128-row strips loaded through a
        // swizzled tensor-map copy, 64-row strips if the driver refuses the tensor map): the
        // one-warp kernels need 255 registers there and the four-warp kernel (quad.cuh) pays three
        // barriers per instance.
        // (the switches are re-read on every call -- a getenv is noise next to a launch -- so
        //  one test process can exercise every variant)
        const int forced = [] {
            const char* e = getenv("CARLE_FUSED_IMPL");
            if (!e) return 0;
            return strcmp(e, "direct") == 0 ? 1 : strcmp(e, "tma") == 0 ? 2
                 : strcmp(e, "quad") == 0 ? 3 : strcmp(e, "strip") == 0 ? 4 : 0;
        }();
        const int strip_r = env_int("CARLE_STRIP_R", 4);
        const bool strip128 = env_int("CARLE_STRIP128", 0) != 0;
        const bool aligned16 = (reinterpret_cast<uintptr_t>(a.action) & 15u) == 0;   // bulk copies
        const bool strip_ok = aligned16 && h->strip_scratch &&
                              ((shape == 3 && (strip_r == 2 || strip_r == 4)) || shape == 2);
        const bool want_strip = forced == 4 || (forced == 0 && (shape == 3 || (shape == 2 && strip128)));
        const bool extras = a.reward_zero || a.obs;           // (the four-warp kernel has none)
        const bool use_strip = strip_ok && (want_strip || (packed && shape == 3 && forced != 2));
        const bool use_quad = !use_strip && !packed && shape == 3 && forced == 3 && aligned16 && !extras;
        const bool use_direct = !use_strip && !use_quad && !packed &&
                                (forced == 1 || !aligned16 || (forced != 2 && h->wpr >= 8));
        // the persistent kernels (strips, TMA stream) can carry the SpeedDetector tail themselves
        const bool fuse_sd = sd && !use_quad && !use_direct;
        if (fuse_sd) {
            p.sd_com_prev = a.speed_com_prev; p.sd_com = a.speed_com_next;
            p.sd_vel = a.speed_velocity; p.sd_speed = a.speed_out; p.sd_sumsq = a.speed_sumsq;
            p.sd_primed = a.speed_primed;
            p.sd_acc = reinterpret_cast<double*>(h->retire + 8);
        }
        cudaError_t e;
        if (use_strip) {
            e = carle::launch_strip(h->device, h->rule_id, shape, shape == 3 ? strip_r : 2,
                                    h->sm_count, pdl_enabled(), p, s);
            if (e == cudaErrorNotSupported && shape == 3 && strip_r == 4)    // no tensor map: 64-row strips
                e = carle::launch_strip(h->device, h->rule_id, shape, 2, h->sm_count, pdl_enabled(), p, s);
        } else if (use_quad) {
            e = carle::launch_quad(h->rule_id, h->sm_count, p, s);
        } else if (use_direct) {
            e = carle::launch_fused(h->rule_id, shape, p, s);
        } else {
            e = carle::launch_stream(h->device, h->rule_id, shape, h->sm_count, pdl_enabled(), p, s);
        }
        CUDA_TRY(e);
        if (a.obs) {
            // rare master reset: the step's last warp cleared the packed state; the unpacked
            // observation is cleared by a whole grid that exits at once in the common case
            carle::clear_if_kernel<<<h->sm_count * 4, 256, 0, s>>>(
                nullptr, h->retire + 7, nullptr, 0, static_cast<uint32_t*>(a.obs),
                obs_words_of(h, a.obs_dtype), nullptr, 0, nullptr);
            CUDA_TRY(cudaGetLastError());
        }
        if (sd && !fuse_sd) return tail_behind();
        return CARLE_OK;
    }
    // ---- no one-launch kernel for this geometry / action format: pack (or flags) + step, and
    //      the extra outputs from their own launches ----
    int rc = CARLE_OK;
    const bool window = h->aw > 0 && h->ah > 0;
    h->defer_reset = a.defer_reset ? 1 : 0;
    if (a.action && window && raw) {
        const size_t need = (size_t)a.action_batch * h->aw * h->awpr;
        if (need > h->act_scratch_words) {             // (allocated on first use)
            DEVICE_GUARD(h);
            if (h->act_scratch) cudaFree(h->act_scratch);
            h->act_scratch = nullptr; h->act_scratch_words = 0;
            CUDA_TRY(cudaMalloc(&h->act_scratch, (size_t)h->n * h->aw * h->awpr * sizeof(uint32_t)));
            h->act_scratch_words = (size_t)h->n * h->aw * h->awpr;
        }
        rc = carle_pack_action(h, a.action, a.action_dtype, a.action_batch, 1, h->act_scratch, gflags, stream);
        if (!rc) rc = carle_step(h, a.state_in, a.state_out, h->act_scratch, a.action_batch, gflags,
                                 a.counters, a.reductions, stream);
    } else if (a.action && window) {                   // already packed: flags from the words
        rc = carle_pack_action(h, a.action, CARLE_PACKED, a.action_batch, 1, nullptr, gflags, stream);
        if (!rc) rc = carle_step(h, a.state_in, a.state_out, static_cast<const uint32_t*>(a.action),
                                 a.action_batch, gflags, a.counters, a.reductions, stream);
    } else {
        rc = carle_step(h, a.state_in, a.state_out, nullptr, 1, nullptr, a.counters, a.reductions, stream);
    }
    h->defer_reset = 0;
    if (rc) return rc;
    if (a.reward_zero) {
        DEVICE_GUARD(h);
        CUDA_TRY(cudaMemsetAsync(a.reward_zero, 0, sizeof(float) * (size_t)h->n, s));
    }
    if (a.obs) {
        rc = carle_unpack_state(h, a.state_out, a.obs, a.obs_dtype, stream);
        if (rc) return rc;
    }
    if (sd) return tail_behind();
    return CARLE_OK;
}

CARLE_API int carle_step_action(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                                const void* action, int dtype, int64_t action_batch,
                                int64_t* counters, int64_t* reductions, void* stream) {
    if (!h || !state_in || !state_out || !action)
        return fail(CARLE_EINVAL, "carle_step_action: NULL argument");
    if (dtype != CARLE_F32 && dtype != CARLE_U8)
        return fail(CARLE_EINVAL, "carle_step_action: dtype must be CARLE_F32 or CARLE_U8");
    carle_step_args a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a;
    a.state_in = state_in; a.state_out = state_out;
    a.action = action; a.action_dtype = dtype; a.action_batch = action_batch;
    a.counters = counters; a.reductions = reductions;
    return carle_step_ex(h, &a, stream);
}

CARLE_API int carle_apply_reset(carle_handle_t h, const int32_t* decision, uint32_t* state, void* obs,
                                int obs_dtype, int64_t* reductions, int64_t* counters, void* stream) {
    if (!h || !decision) return fail(CARLE_EINVAL, "carle_apply_reset: NULL argument");
    if (obs && obs_dtype != CARLE_F32 && obs_dtype != CARLE_U8)
        return fail(CARLE_EINVAL, "carle_apply_reset: obs dtype must be CARLE_F32 or CARLE_U8");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    carle::clear_if_kernel<<<h->sm_count * 4, 256, 0, s>>>(
        decision, nullptr, state, state ? h->n * (long long)h->h * h->wpr : 0,
        static_cast<uint32_t*>(obs), obs ? obs_words_of(h, obs_dtype) : 0,
        reinterpret_cast<long long*>(reductions), reductions ? h->n * 4 : 0,
        reinterpret_cast<long long*>(counters));
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_band_create(carle_handle_t* out, int device, int height, int width,
                                int action_height, int action_width, int band_row0,
                                int band_rows, int halo) {
    if (width % 32 != 0)
        return fail(CARLE_EINVAL, "carle_band_create: width must be a multiple of 32");
    if (halo < 8 || halo > 32 || halo % 8 != 0)
        return fail(CARLE_EINVAL, "carle_band_create: halo must be 8, 16, 24 or 32 rows");
    if (band_rows < halo || band_rows % 8 != 0 || band_row0 < 0 || band_row0 + band_rows > height)
        return fail(CARLE_EINVAL, "carle_band_create: band must lie inside the grid, be a "
                                  "multiple of 8 rows and at least one halo tall");
    int rc = carle_create(out, device, 1, height, width, action_height, action_width);
    if (rc) return rc;
    carle_ctx* c = *out;
    c->family = 2;
    c->band_row0 = band_row0; c->band_rows = band_rows; c->halo = halo;
    c->grid_h = height;
    c->h = band_rows + 2 * halo;              // rows of the local buffer
    return CARLE_OK;
}

CARLE_API int carle_band_step(carle_handle_t h, const uint32_t* in, uint32_t* out,
                              uint32_t* peer_up_out, uint32_t* peer_dn_out, int generations,
                              const uint32_t* packed_actions, int32_t* flags, int64_t* counters,
                              int32_t* sync_local, int32_t* sync_up, int32_t* sync_dn, void* stream) {
    if (!h || !in || !out) return fail(CARLE_EINVAL, "carle_band_step: NULL buffer");
    if (h->halo == 0) return fail(CARLE_EINVAL, "carle_band_step: handle is not a band");
    if (generations < 1 || generations > h->halo)
        return fail(CARLE_EINVAL, "carle_band_step: 1 <= generations <= halo");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    carle::TiledParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.s = base_params(h);
    const long long entry_words = (long long)h->aw * h->awpr;
    tp.s.in = in; tp.s.out = out;
    tp.s.act = entry_words ? packed_actions : nullptr;
    tp.s.act_inst_stride = 0;
    tp.s.act_step_stride = entry_words;
    tp.s.flags = flags;
    tp.s.counters = reinterpret_cast<long long*>(counters);
    tp.s.k = generations;
    tp.tv = h->halo;
    tp.out_row0 = h->halo; tp.out_rows = h->band_rows; tp.vwrap = 0;
    tp.act_row_shift = h->band_row0 - h->halo;
    tp.grid_h = h->grid_h;
    tp.peer_up = peer_up_out; tp.peer_dn = peer_dn_out;
    if (sync_local) {
        if (!sync_up || !sync_dn)
            return fail(CARLE_EINVAL, "carle_band_step: sync_local needs both neighbours' sync words");
        tp.sync_local = sync_local;
        tp.sync_up_word = sync_up + 1;          // I am my upper neighbour's LOWER neighbour
        tp.sync_dn_word = sync_dn + 0;          // ... and my lower neighbour's UPPER one
    }
    CUDA_TRY(launch_tiled(h, tp, s));
    return CARLE_OK;
}

CARLE_API int carle_band_push_halos(carle_handle_t h, const uint32_t* buf, uint32_t* peer_up_buf,
                                    uint32_t* peer_dn_buf, void* stream) {
    if (!h || !buf) return fail(CARLE_EINVAL, "carle_band_push_halos: NULL buffer");
    if (h->halo == 0) return fail(CARLE_EINVAL, "carle_band_push_halos: handle is not a band");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const size_t row = (size_t)h->wpr * sizeof(uint32_t), halo_bytes = row * h->halo;
    // my first `halo` band rows are the upper neighbour's bottom halo, and vice versa
    if (peer_up_buf)
        CUDA_TRY(cudaMemcpyAsync(peer_up_buf + (size_t)(h->halo + h->band_rows) * h->wpr,
                                 buf + (size_t)h->halo * h->wpr, halo_bytes, cudaMemcpyDefault, s));
    if (peer_dn_buf)
        CUDA_TRY(cudaMemcpyAsync(peer_dn_buf, buf + (size_t)h->band_rows * h->wpr, halo_bytes,
                                 cudaMemcpyDefault, s));
    return CARLE_OK;
}

CARLE_API int carle_dev_alloc(int device, uint64_t bytes, void** out) {
    if (!out) return fail(CARLE_EINVAL, "carle_dev_alloc: NULL argument");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(CARLE_ECUDA, "carle_dev_alloc: cudaSetDevice failed");
    CUDA_TRY(cudaMalloc(out, bytes));
    CUDA_TRY(cudaMemset(*out, 0, bytes));
    return CARLE_OK;
}

CARLE_API int carle_dev_free(int device, void* ptr) {
    if (!ptr) return CARLE_OK;
    DeviceGuard guard(device);
    CUDA_TRY(cudaFree(ptr));
    return CARLE_OK;
}

CARLE_API int carle_ipc_export(const void* dev_ptr, unsigned char handle_out[64]) {
    if (!dev_ptr || !handle_out) return fail(CARLE_EINVAL, "carle_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    CUDA_TRY(cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr)));
    memcpy(handle_out, &hd, 64);
    return CARLE_OK;
}

CARLE_API int carle_ipc_open(const unsigned char handle[64], void** dev_ptr_out) {
    if (!handle || !dev_ptr_out) return fail(CARLE_EINVAL, "carle_ipc_open: NULL argument");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
    return CARLE_OK;
}

CARLE_API int carle_ipc_close(void* dev_ptr) {
    if (!dev_ptr) return CARLE_OK;
    CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return CARLE_OK;
}

CARLE_API int carle_random_action(carle_handle_t h, uint64_t seed, uint32_t step, double toggle_rate,
                                  int64_t batch, uint32_t* packed_action, void* stream) {
    if (!h || !packed_action) return fail(CARLE_EINVAL, "carle_random_action: NULL argument");
    if (batch != 1 && batch != h->n)
        return fail(CARLE_EINVAL, "carle_random_action: action batch must be 1 or N");
    if (!(toggle_rate >= 0.0 && toggle_rate <= 1.0))
        return fail(CARLE_EINVAL, "carle_random_action: toggle_rate must be in [0, 1]");
    if (h->aw == 0 || h->ah == 0) return CARLE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long rows = batch * h->aw;
    const uint32_t threshold = (uint32_t)(toggle_rate * 65536.0 + 0.5);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    carle::random_action_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
        packed_action, rows, h->aw, h->ah, h->awpr, h->col0 - 32 * h->aw0, threshold, key, step);
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_step_random(carle_handle_t h, const uint32_t* state_in, uint32_t* state_out,
                                uint64_t seed, uint32_t step, double toggle_rate,
                                int64_t action_batch, uint32_t* packed_scratch,
                                int64_t* counters, int64_t* reductions, void* stream) {
    if (!h || !state_in || !state_out)
        return fail(CARLE_EINVAL, "carle_step_random: NULL state pointer");
    if (action_batch != 1 && action_batch != h->n)
        return fail(CARLE_EINVAL, "carle_step_random: action batch must be 1 or N");
    if (!(toggle_rate >= 0.0 && toggle_rate <= 1.0))
        return fail(CARLE_EINVAL, "carle_step_random: toggle_rate must be in [0, 1]");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int32_t* gflags = reinterpret_cast<int32_t*>(h->retire + 2);
    const int shape = (h->family == 1 && h->row0 % h->wpr == 0 && h->n < (1LL << 32))
                          ? fused_shape(h->wpr, h->aw, h->ah) : 0;
    if (shape) {
        DEVICE_GUARD(h);
        carle::StepParams p = base_params(h);
        p.in = state_in; p.out = state_out;
        p.raw_inst_stride = (action_batch == 1) ? 0 : 1;
        p.flags = gflags;
        p.counters = reinterpret_cast<long long*>(counters);
        p.red = reinterpret_cast<long long*>(reductions);
        p.k = 1;
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t threshold = (uint32_t)(toggle_rate * 65536.0 + 0.5);
        // 64x64 / 128x128: the persistent stream kernel with the toggles drawn in registers (PDL,
        // fence-free retirement); CARLE_RANDOM_IMPL=direct keeps the one-warp-per-instance kernel
        const char* impl = getenv("CARLE_RANDOM_IMPL");
        if ((shape == 1 || shape == 2) && !(impl && strcmp(impl, "direct") == 0)) {
            p.rand_key0 = key.x; p.rand_key1 = key.y; p.rand_step = step; p.rand_threshold = threshold;
            CUDA_TRY(carle::launch_stream_random(h->device, h->rule_id, shape, h->sm_count,
                                                 pdl_enabled(), p, s));
            return CARLE_OK;
        }
        // 256x256: the toggles are drawn into the caller's scratch as packed words (8 MiB at 16384
        // instances) and the strip kernel ingests them -- two launches, ~3x faster than the one-warp
        // kernel (255 registers, 8 warps per SM) that draws them in registers
        if (shape == 3 && packed_scratch && h->strip_scratch && !(impl && strcmp(impl, "direct") == 0) &&
            (reinterpret_cast<uintptr_t>(packed_scratch) & 15u) == 0) {
            int rc = carle_random_action(h, seed, step, toggle_rate, action_batch, packed_scratch, stream);
            if (rc) return rc;
            carle_step_args a;
            memset(&a, 0, sizeof a);
            a.struct_size = sizeof a;
            a.action_dtype = CARLE_PACKED;
            a.state_in = state_in; a.state_out = state_out;
            a.action = packed_scratch; a.action_batch = action_batch;
            a.counters = counters; a.reductions = reductions;
            return carle_step_ex(h, &a, stream);
        }
        CUDA_TRY(carle::launch_random_direct(h->rule_id, shape, p, key, step, threshold, s));
        return CARLE_OK;
    }
    // other geometries: generate into the caller's scratch, then flags + step (3 launches)
    if (!packed_scratch)
        return fail(CARLE_EINVAL, "carle_step_random: this geometry needs packed_scratch");
    int rc = carle_random_action(h, seed, step, toggle_rate, action_batch, packed_scratch, stream);
    if (rc) return rc;
    if (h->aw > 0 && h->ah > 0) {
        rc = carle_pack_action(h, packed_scratch, CARLE_PACKED, action_batch, 1, nullptr, gflags, stream);
        if (rc) return rc;
        return carle_step(h, state_in, state_out, packed_scratch, action_batch, gflags, counters,
                          reductions, stream);
    }
    return carle_step(h, state_in, state_out, nullptr, 1, nullptr, counters, reductions, stream);
}

CARLE_API int carle_unpack_action(carle_handle_t h, const uint32_t* packed_action, int64_t batch,
                                  float* action, void* stream) {
    if (!h || !packed_action || !action)
        return fail(CARLE_EINVAL, "carle_unpack_action: NULL argument");
    if (h->aw == 0 || h->ah == 0) return CARLE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long total = batch * (long long)h->aw * h->ah;
    carle::unpack_action_kernel<<<grid_for(total, 256, h->sm_count), 256, 0, s>>>(
        packed_action, action, total, h->aw, h->ah, h->awpr, h->col0 - 32 * h->aw0);
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_apply_action(carle_handle_t h, uint32_t* state, const uint32_t* packed_action,
                                 int64_t action_batch, void* stream) {
    if (!h || !state || !packed_action)
        return fail(CARLE_EINVAL, "carle_apply_action: NULL argument");
    if (action_batch != 1 && action_batch != h->n)
        return fail(CARLE_EINVAL, "carle_apply_action: action batch must be 1 or N");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    carle::StepParams p = base_params(h);
    const long long entry_words = (long long)h->aw * h->awpr;
    if (entry_words == 0) return CARLE_OK;
    p.act = packed_action;
    p.act_inst_stride = (action_batch == 1) ? 0 : entry_words;
    carle::apply_action_kernel<<<grid_for(h->n * entry_words, 256, h->sm_count), 256, 0, s>>>(
        p, state);
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_reduce(carle_handle_t h, const uint32_t* state, int64_t* out, void* stream) {
    if (!h || !state || !out) return fail(CARLE_EINVAL, "carle_reduce: NULL argument");
    DEVICE_GUARD(h);
    return launch_reduce(h, state, out, static_cast<cudaStream_t>(stream));
}

CARLE_API int carle_masked_count(carle_handle_t h, const uint32_t* state, const uint32_t* plus_mask,
                       const uint32_t* minus_mask, int64_t* out, void* stream) {
    if (!h || !state || !out) return fail(CARLE_EINVAL, "carle_masked_count: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int64_t) * h->n, s));
    const long long words = (long long)h->h * h->wpr;
    int bx = (int)((words + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 4096) bx = 4096;
    long long done = 0;
    while (done < h->n) {
        long long chunk = h->n - done;
        if (chunk > 65535) chunk = 65535;
        dim3 grid(bx, (unsigned)chunk);
        carle::masked_count_kernel<<<grid, 256, 0, s>>>(
            state + done * words, plus_mask, minus_mask, words,
            reinterpret_cast<long long*>(out) + done);
        CUDA_TRY(cudaGetLastError());
        done += chunk;
    }
    return CARLE_OK;
}

CARLE_API int carle_action_count(carle_handle_t h, const uint32_t* packed_action, int64_t batch,
                       int64_t* out, void* stream) {
    if (!h || !packed_action || !out)
        return fail(CARLE_EINVAL, "carle_action_count: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    const long long words = (long long)h->aw * h->awpr;
    carle::action_count_kernel<<<grid_for(batch, 8, h->sm_count), 256, 0, s>>>(
        packed_action, batch, words, reinterpret_cast<long long*>(out));
    CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_speed_tail(carle_handle_t h, const int64_t* reductions, float* center_of_mass,
                               int have_previous, float* velocity_out, float* speed_out,
                               float* reward, double* sumsq_out, int32_t* primed, void* stream) {
    if (!h || !reductions || !center_of_mass || !speed_out)
        return fail(CARLE_EINVAL, "carle_speed_tail: NULL argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DEVICE_GUARD(h);
    long long blocks = (h->n + 255) / 256;
    if (blocks > (long long)h->sm_count * 4) blocks = (long long)h->sm_count * 4;
    // launched as a programmatic dependent of the step kernel in front of it (its launch latency
    // and prologue overlap that kernel's tail; it waits for the sums with griddepcontrol.wait)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, carle::speed_tail_kernel,
                                reinterpret_cast<const long long*>(reductions), center_of_mass,
                                (long long)h->n, have_previous ? 1 : 0, velocity_out, speed_out, reward,
                                sumsq_out, primed, reinterpret_cast<double*>(h->retire + 8), h->retire + 10));
    return CARLE_OK;
}

CARLE_API int carle_jit_probe(int shape, uint32_t birth_mask, uint32_t survive_mask,
                              int64_t* cubin_bytes) {
    birth_mask &= 0x1FFu; survive_mask &= 0x1FFu;
    if (birth_mask == 0 || survive_mask == 0)
        return fail(CARLE_ERULE, "carle_jit_probe: empty birth or survive set");
    char inst[192];
    if (shape == 1)
        carle::stream_instantiation<2, float, 1, 16, false>(inst, sizeof inst, birth_mask, survive_mask);
    else if (shape == 2)
        carle::stream_instantiation<4, float, 1, 8, false>(inst, sizeof inst, birth_mask, survive_mask);
    else if (shape == 7)
        carle::stream_instantiation<4, carle::DeviceRandom, 1, 8, true>(inst, sizeof inst, birth_mask,
                                                                        survive_mask);
    else if (shape == 3)
        snprintf(inst, sizeof inst,
                 "carle::step_strip_kernel<8, 2, 64, carle::StaticRule<%uu, %uu>, float, 1>",
                 birth_mask, survive_mask);
    else if (shape == 4)
        snprintf(inst, sizeof inst, "carle::step_warp_kernel<8, carle::StaticRule<%uu, %uu>>", birth_mask,
                 survive_mask);
    else if (shape == 5)
        snprintf(inst, sizeof inst, "carle::step_generic_kernel<carle::StaticRule<%uu, %uu>>", birth_mask,
                 survive_mask);
    else if (shape == 6)
        snprintf(inst, sizeof inst, "carle::step_tiled_kernel<carle::StaticRule<%uu, %uu>, 8>", birth_mask,
                 survive_mask);
    else if (shape == 8)
        snprintf(inst, sizeof inst, "carle::step_tiled_kernel<carle::StaticRule<%uu, %uu>, 4>", birth_mask,
                 survive_mask);
    else if (shape == 9)
        snprintf(inst, sizeof inst,
                 "carle::step_strip_kernel<8, 4, 64, carle::StaticRule<%uu, %uu>, carle::PackedWords, 1>",
                 birth_mask, survive_mask);
    else if (shape == 10)
        carle::stream_instantiation<4, carle::PackedWords, 1, 8, false>(inst, sizeof inst, birth_mask, survive_mask);
    else
        return fail(CARLE_EINVAL, "carle_jit_probe: shape must be 1..10");
    std::vector<char> cubin;
    std::string lowered, log;
    if (carle::jit_compile(inst, &cubin, &lowered, &log) != 0)
        return fail(CARLE_ECUDA, std::string("carle_jit_probe: ") + inst + ": " + log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return CARLE_OK;
}

CARLE_API int carle_jit_loaded(void) { return carle::jit_loaded(); }

}  // extern "C"
