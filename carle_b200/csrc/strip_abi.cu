// strip_abi.cu — instantiations and launcher of step_strip_kernel (second translation unit of
// libcarle_b200.so, compiled in parallel with carle_abi.cu).
#include <cuda.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include "abi_internal.h"
#include "strip.cuh"

namespace carle {
namespace {

// Tensor map over a packed state buffer seen as [lines][32 words] (128-byte lines), box =
// box_lines x 128 bytes, 128-byte swizzle.  Encoded once per (buffer, size) and cached: a rollout
// alternates between two buffers.
bool state_tensor_map(const void* base, long long lines, int box_lines, TensorMap* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static const EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode || lines >= (1LL << 31)) return false;
    static std::mutex mu;
    static std::map<std::tuple<const void*, long long, int>, TensorMap> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(base, lines, box_lines);
    auto it = cache.find(key);
    if (it == cache.end()) {
        static_assert(sizeof(CUtensorMap) == sizeof(TensorMap), "tensor map size");
        CUtensorMap m;
        const cuuint64_t dims[2] = {32, (cuuint64_t)lines};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {32, (cuuint32_t)box_lines};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        TensorMap t;
        memcpy(&t, &m, sizeof t);
        if (cache.size() > 64) cache.clear();
        it = cache.emplace(key, t).first;
    }
    *out = it->second;
    return true;
}

template <int WPL, int R, int AWIN, class Rule, typename T, int DEPTH>
cudaError_t launch_strip_d(int device, int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    using L = StripLayout<WPL, R, AWIN, T>;
    constexpr int warps = CARLE_STRIP_WARPS;
    const size_t smem = (size_t)warps * L::warp_bytes(DEPTH);
    if (p.n * L::U >= (1LL << 31)) return cudaErrorNotSupported;     // (32-bit unit indices in the kernel)
    TensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if constexpr (L::SWZ) {
        if (!state_tensor_map(p.in, p.n * L::H * (long long)L::ROW_BYTES / 128, L::BODY_LINES, &tmap))
            return cudaErrorNotSupported;               // the caller falls back to 64-row strips
    }
    if constexpr (std::is_same<Rule, DynamicRule>::value) {
        // any rule without a built-in instantiation: NVRTC-specialised StaticRule kernel (jit.cu)
        char inst[192];
        snprintf(inst, sizeof inst,
                 "carle::step_strip_kernel<%d, %d, %d, carle::StaticRule<%uu, %uu>, %s, %d>", WPL, R, AWIN,
                 p.birth, p.survive,
                 IsPackedWords<T>::value ? "carle::PackedWords" : (sizeof(T) == 1 ? "unsigned char" : "float"),
                 DEPTH);
        if (void* fn = jit_kernel(device, inst))
            return jit_launch(fn, sm_count, warps * 32, smem, (p.n * L::U + warps - 1) / warps, L::U,
                              pdl, p, p.n * L::U, s, &tmap);
    }
    auto kernel = step_strip_kernel<WPL, R, AWIN, Rule, T, DEPTH>;
    // (queried once per device for this instantiation, see stream_launch.h)
    static int cached_ctas[kMaxDevices] = {0};
    int ctas_per_sm = (device >= 0 && device < kMaxDevices) ? cached_ctas[device] : 0;
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, warps * 32, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        if (device >= 0 && device < kMaxDevices) cached_ctas[device] = ctas_per_sm;
    }
    long long blocks = (long long)sm_count * ctas_per_sm;
    const long long need = (p.n * L::U + warps - 1) / warps;
    if (blocks > need) blocks = need;
    blocks -= blocks % L::U;                      // the strips of an instance share a trip
    if (blocks < L::U) blocks = L::U;
    StepParams q = p;
    q.rank_blocked = rank_blocked_for(p.n * L::U, blocks * warps);
    walk_policy(q);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(warps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, q, tmap);
}

template <int WPL, int R, int AWIN, class Rule, int DEPTH>
cudaError_t launch_strip_t(int device, int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    if (p.raw_u8 == 2) return launch_strip_d<WPL, R, AWIN, Rule, PackedWords, DEPTH>(device, sm_count, pdl, p, s);
    if (p.raw_u8) return launch_strip_d<WPL, R, AWIN, Rule, uint8_t, DEPTH>(device, sm_count, pdl, p, s);
    return launch_strip_d<WPL, R, AWIN, Rule, float, DEPTH>(device, sm_count, pdl, p, s);
}

template <class Rule>
cudaError_t launch_strip_rule(int device, int shape, int r, int sm_count, bool pdl,
                              const StepParams& p, cudaStream_t s) {
    if (shape == 3 && r == 2) return launch_strip_t<8, 2, 64, Rule, 1>(device, sm_count, pdl, p, s);
    if (shape == 3 && r == 4) return launch_strip_t<8, 4, 64, Rule, 1>(device, sm_count, pdl, p, s);
    if (shape == 2 && r == 2) return launch_strip_t<4, 2, 32, Rule, 2>(device, sm_count, pdl, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_strip(int device, int rule_id, int shape, int rows_per_lane, int sm_count,
                         bool pdl, const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_strip_rule<StaticRule<kLifeB, kLifeS>>(device, shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_MORLEY: return launch_strip_rule<StaticRule<kMorleyB, kMorleyS>>(device, shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_HIGHLIFE: return launch_strip_rule<StaticRule<kHighB, kHighS>>(device, shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_DAYNIGHT: return launch_strip_rule<StaticRule<kDayNightB, kDayNightS>>(device, shape, rows_per_lane, sm_count, pdl, p, s);
        default: return launch_strip_rule<DynamicRule>(device, shape, rows_per_lane, sm_count, pdl, p, s);
    }
}

}  // namespace carle
