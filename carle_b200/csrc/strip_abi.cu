// strip_abi.cu — instantiations and launcher of step_strip_kernel (second translation unit of
// libcarle_b200.so, compiled in parallel with carle_abi.cu).
#include <string.h>
#include "abi_internal.h"
#include "strip.cuh"

namespace carle {
namespace {

template <int WPL, int R, int AWIN, class Rule, typename T, int DEPTH>
cudaError_t launch_strip_d(int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    using L = StripLayout<WPL, R, AWIN, T>;
    constexpr int warps = 4;
    const size_t smem = (size_t)warps * L::warp_bytes(DEPTH);
    auto kernel = step_strip_kernel<WPL, R, AWIN, Rule, T, DEPTH>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas_per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    long long blocks = (long long)sm_count * ctas_per_sm;
    const long long need = (p.n * L::U + warps - 1) / warps;
    if (blocks > need) blocks = need;
    blocks -= blocks % L::U;                      // the strips of an instance share a trip
    if (blocks < L::U) blocks = L::U;
    StepParams q = p;
    q.rank_blocked = rank_blocked_for(p.n * L::U, blocks * warps);
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

template <int WPL, int R, int AWIN, class Rule, int DEPTH>
cudaError_t launch_strip_t(int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    if (p.raw_u8) return launch_strip_d<WPL, R, AWIN, Rule, uint8_t, DEPTH>(sm_count, pdl, p, s);
    return launch_strip_d<WPL, R, AWIN, Rule, float, DEPTH>(sm_count, pdl, p, s);
}

template <class Rule>
cudaError_t launch_strip_rule(int shape, int r, int sm_count, bool pdl, const StepParams& p,
                              cudaStream_t s) {
    if (shape == 3 && r == 2) return launch_strip_t<8, 2, 64, Rule, 1>(sm_count, pdl, p, s);
    if (shape == 3 && r == 4) return launch_strip_t<8, 4, 64, Rule, 1>(sm_count, pdl, p, s);
    if (shape == 2 && r == 2) return launch_strip_t<4, 2, 32, Rule, 2>(sm_count, pdl, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_strip(int rule_id, int shape, int rows_per_lane, int sm_count, bool pdl,
                         const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_strip_rule<StaticRule<kLifeB, kLifeS>>(shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_MORLEY: return launch_strip_rule<StaticRule<kMorleyB, kMorleyS>>(shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_HIGHLIFE: return launch_strip_rule<StaticRule<kHighB, kHighS>>(shape, rows_per_lane, sm_count, pdl, p, s);
        case RULE_DAYNIGHT: return launch_strip_rule<StaticRule<kDayNightB, kDayNightS>>(shape, rows_per_lane, sm_count, pdl, p, s);
        default: return launch_strip_rule<DynamicRule>(shape, rows_per_lane, sm_count, pdl, p, s);
    }
}

}  // namespace carle
