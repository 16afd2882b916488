// stream_launch.h — host-side launcher of step_stream_kernel, shared by the translation units that
// instantiate it (carle_abi.cu: float32 / uint8 actions, random_abi.cu: device random agent).
#pragma once
#include <stdio.h>
#include <string.h>
#include <type_traits>
#include "abi_internal.h"

namespace carle {

template <typename T> inline const char* stream_type_name() {
    return IsDeviceRandom<T>::value ? "carle::DeviceRandom"
         : IsPackedWords<T>::value ? "carle::PackedWords" : (sizeof(T) == 1 ? "unsigned char" : "float");
}

// slots per warp: two whenever the CTAs asked of ptxas still fit an SM with them
template <int WPR, typename T, int C, int G, bool BIG>
constexpr int stream_depth() {
    using L = StreamLayout<WPR, T, C, G>;
    return (stream_min_ctas(WPR, BIG) * (stream_warps(WPR, BIG) * L::warp_bytes(2) + 1024) <= 227 * 1024) ? 2 : 1;
}

// the name expression NVRTC instantiates for a run-time rule (jit.cu); also used by
// carle_jit_probe, so the CPU test-suite compiles exactly what the launcher asks for
template <int WPR, typename T, int C, int G, bool BIG>
void stream_instantiation(char* buf, size_t size, uint32_t birth, uint32_t survive) {
    snprintf(buf, size,
             "carle::step_stream_kernel<%d, carle::StaticRule<%uu, %uu>, %s, %d, %d, %d, %s>", WPR, birth,
             survive, stream_type_name<T>(), C, G, stream_depth<WPR, T, C, G, BIG>(),
             BIG ? "true" : "false");
}

template <int WPR, class Rule, typename T, int C, int G, bool BIG>
cudaError_t launch_stream_b(int device, int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    using L = StreamLayout<WPR, T, C, G>;
    const int warps = stream_warps(WPR, BIG);
    constexpr int DEPTH = stream_depth<WPR, T, C, G, BIG>();
    const size_t smem = (size_t)warps * L::warp_bytes(DEPTH);
    if constexpr (std::is_same<Rule, DynamicRule>::value) {
        // any rule without a built-in instantiation: NVRTC-specialised StaticRule kernel (jit.cu)
        char inst[192];
        stream_instantiation<WPR, T, C, G, BIG>(inst, sizeof inst, p.birth, p.survive);
        if (void* fn = jit_kernel(device, inst))
            return jit_launch(fn, sm_count, warps * 32, smem, (p.n + warps - 1) / warps, 1,
                              pdl, p, p.n, s);
    }
    auto kernel = step_stream_kernel<WPR, Rule, T, C, G, DEPTH, BIG>;
    // shared-memory opt-in and occupancy: queried once per device for this instantiation (two
    // runtime calls that cost more host time than the launch itself)
    static int cached_ctas[kMaxDevices] = {0};
    int ctas_per_sm = (device >= 0 && device < kMaxDevices) ? cached_ctas[device] : 0;
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, warps * 32, smem);
        if (e != cudaSuccess) return e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        if (device >= 0 && device < kMaxDevices) cached_ctas[device] = ctas_per_sm;
    }
    long long blocks = (long long)sm_count * ctas_per_sm;
    const long long need = (p.n + warps - 1) / warps;
    if (blocks > need) blocks = need;
    StepParams q = p;
    q.rank_blocked = rank_blocked_for(p.n, blocks * warps);
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
    return cudaLaunchKernelEx(&cfg, kernel, q);
}


}  // namespace carle
