// random_abi.cu — instantiations of step_stream_kernel whose action is the device-side random agent
// (third translation unit of libcarle_b200.so, compiled in parallel with the others).
#include "stream_launch.h"

namespace carle {
namespace {

template <int WPR, class Rule, int C, int G>
cudaError_t launch_random_shape(int device, int sm_count, bool pdl, const StepParams& p,
                                cudaStream_t s) {
    if constexpr (WPR == 4) {
        if (p.n >= 8LL * sm_count * 24)
            return launch_stream_b<WPR, Rule, DeviceRandom, C, G, true>(device, sm_count, pdl, p, s);
    }
    return launch_stream_b<WPR, Rule, DeviceRandom, C, G, false>(device, sm_count, pdl, p, s);
}

template <class Rule>
cudaError_t launch_random_rule(int device, int shape, int sm_count, bool pdl, const StepParams& p,
                               cudaStream_t s) {
    if (shape == 1) return launch_random_shape<2, Rule, 1, 16>(device, sm_count, pdl, p, s);
    if (shape == 2) return launch_random_shape<4, Rule, 1, 8>(device, sm_count, pdl, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_stream_random(int device, int rule_id, int shape, int sm_count, bool pdl,
                                 const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_random_rule<StaticRule<kLifeB, kLifeS>>(device, shape, sm_count, pdl, p, s);
        case RULE_MORLEY: return launch_random_rule<StaticRule<kMorleyB, kMorleyS>>(device, shape, sm_count, pdl, p, s);
        case RULE_HIGHLIFE: return launch_random_rule<StaticRule<kHighB, kHighS>>(device, shape, sm_count, pdl, p, s);
        case RULE_DAYNIGHT: return launch_random_rule<StaticRule<kDayNightB, kDayNightS>>(device, shape, sm_count, pdl, p, s);
        default: return launch_random_rule<DynamicRule>(device, shape, sm_count, pdl, p, s);
    }
}

}  // namespace carle
