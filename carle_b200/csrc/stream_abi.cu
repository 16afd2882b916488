// stream_abi.cu — instantiations of step_stream_kernel for float32 / uint8 actions (compiled in
// parallel with the other translation units of libcarle_b200.so).
#include "stream_launch.h"

namespace carle {
namespace {

template <int WPR, class Rule, typename T, int C, int G>
cudaError_t launch_stream_tt(int device, int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    if constexpr (WPR == 4) {
        // long 128 x 128 batches: three 8-warp CTAs per SM (see stream_warps)
        if (p.n >= 8LL * sm_count * 24)
            return launch_stream_b<WPR, Rule, T, C, G, true>(device, sm_count, pdl, p, s);
    }
    return launch_stream_b<WPR, Rule, T, C, G, false>(device, sm_count, pdl, p, s);
}

template <int WPR, class Rule, int C, int G>
cudaError_t launch_stream_t(int device, int sm_count, bool pdl, const StepParams& p, cudaStream_t s) {
    if (p.raw_u8 == 2) return launch_stream_tt<WPR, Rule, PackedWords, C, G>(device, sm_count, pdl, p, s);
    if (p.raw_u8) return launch_stream_tt<WPR, Rule, uint8_t, C, G>(device, sm_count, pdl, p, s);
    return launch_stream_tt<WPR, Rule, float, C, G>(device, sm_count, pdl, p, s);
}

template <class Rule>
cudaError_t launch_stream_rule(int device, int shape, int sm_count, bool pdl, const StepParams& p,
                               cudaStream_t s) {
    switch (shape) {
        case 1: return launch_stream_t<2, Rule, 1, 16>(device, sm_count, pdl, p, s);
        case 2: return launch_stream_t<4, Rule, 1, 8>(device, sm_count, pdl, p, s);
        case 3: return launch_stream_t<8, Rule, 2, 8>(device, sm_count, pdl, p, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_stream(int device, int rule_id, int shape, int sm_count, bool pdl,
                          const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_stream_rule<StaticRule<kLifeB, kLifeS>>(device, shape, sm_count, pdl, p, s);
        case RULE_MORLEY: return launch_stream_rule<StaticRule<kMorleyB, kMorleyS>>(device, shape, sm_count, pdl, p, s);
        case RULE_HIGHLIFE: return launch_stream_rule<StaticRule<kHighB, kHighS>>(device, shape, sm_count, pdl, p, s);
        case RULE_DAYNIGHT: return launch_stream_rule<StaticRule<kDayNightB, kDayNightS>>(device, shape, sm_count, pdl, p, s);
        default: return launch_stream_rule<DynamicRule>(device, shape, sm_count, pdl, p, s);
    }
}

}  // namespace carle
