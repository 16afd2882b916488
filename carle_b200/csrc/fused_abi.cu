// fused_abi.cu — instantiations of the non-persistent one-launch step kernels kept for A/B runs and
// for action pointers that are not 16-byte aligned: step_fused_kernel (one warp per instance, plain
// loads), step_quad_kernel (four warps per 256 x 256 instance, quad.cuh) and step_random_kernel
// (device random agent, 256 x 256).  Compiled in parallel with the other translation units.
#include "abi_internal.h"
#include "quad.cuh"

namespace carle {
namespace {

// (WPR, window) combinations with compile-time group / chunk counts
template <int WPR, class Rule, int C, int G>
cudaError_t launch_fused_t(const StepParams& p, cudaStream_t s) {
    const int warps_per_block = 4;
    const long long blocks = (p.n + warps_per_block - 1) / warps_per_block;
    if (p.raw_u8)
        step_fused_kernel<WPR, Rule, uint8_t, C, G><<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(p);
    else
        step_fused_kernel<WPR, Rule, float, C, G><<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(p);
    return cudaGetLastError();
}

template <class Rule>
cudaError_t launch_fused_rule(int shape, const StepParams& p, cudaStream_t s) {
    switch (shape) {
        case 1: return launch_fused_t<2, Rule, 1, 16>(p, s);
        case 2: return launch_fused_t<4, Rule, 1, 8>(p, s);
        case 3: return launch_fused_t<8, Rule, 2, 8>(p, s);
    }
    return cudaErrorInvalidValue;
}

template <class Rule>
cudaError_t launch_quad_rule(int sm_count, const StepParams& p, cudaStream_t s) {
    const size_t smem = 2 * (p.raw_u8 ? sizeof(QuadGroupSmem<uint8_t>) : sizeof(QuadGroupSmem<float>));
    long long blocks = (long long)sm_count * CARLE_QUAD_CTAS;
    const long long need = (p.n + 1) / 2;
    if (blocks > need) blocks = need;
    if (p.raw_u8) {
        auto k = step_quad_kernel<Rule, uint8_t>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<(unsigned)blocks, 256, smem, s>>>(p);
    } else {
        auto k = step_quad_kernel<Rule, float>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<(unsigned)blocks, 256, smem, s>>>(p);
    }
    return cudaGetLastError();
}

template <int WPR, class Rule, int C, int G>
cudaError_t launch_random_t(const StepParams& p, uint2 key, uint32_t step, uint32_t thr, cudaStream_t s) {
    const int warps_per_block = 4;
    const long long blocks = (p.n + warps_per_block - 1) / warps_per_block;
    step_random_kernel<WPR, Rule, C, G><<<(unsigned)blocks, warps_per_block * 32, 0, s>>>(p, key, step, thr);
    return cudaGetLastError();
}

template <class Rule>
cudaError_t launch_random_rule(int shape, const StepParams& p, uint2 key, uint32_t step, uint32_t thr,
                               cudaStream_t s) {
    switch (shape) {
        case 1: return launch_random_t<2, Rule, 1, 16>(p, key, step, thr, s);
        case 2: return launch_random_t<4, Rule, 1, 8>(p, key, step, thr, s);
        case 3: return launch_random_t<8, Rule, 2, 8>(p, key, step, thr, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_fused(int rule_id, int shape, const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_fused_rule<StaticRule<kLifeB, kLifeS>>(shape, p, s);
        case RULE_MORLEY: return launch_fused_rule<StaticRule<kMorleyB, kMorleyS>>(shape, p, s);
        case RULE_HIGHLIFE: return launch_fused_rule<StaticRule<kHighB, kHighS>>(shape, p, s);
        case RULE_DAYNIGHT: return launch_fused_rule<StaticRule<kDayNightB, kDayNightS>>(shape, p, s);
        default: return launch_fused_rule<DynamicRule>(shape, p, s);
    }
}

cudaError_t launch_quad(int rule_id, int sm_count, const StepParams& p, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_quad_rule<StaticRule<kLifeB, kLifeS>>(sm_count, p, s);
        case RULE_MORLEY: return launch_quad_rule<StaticRule<kMorleyB, kMorleyS>>(sm_count, p, s);
        case RULE_HIGHLIFE: return launch_quad_rule<StaticRule<kHighB, kHighS>>(sm_count, p, s);
        case RULE_DAYNIGHT: return launch_quad_rule<StaticRule<kDayNightB, kDayNightS>>(sm_count, p, s);
        default: return launch_quad_rule<DynamicRule>(sm_count, p, s);
    }
}

cudaError_t launch_random_direct(int rule_id, int shape, const StepParams& p, uint2 key, uint32_t step,
                                 uint32_t thr, cudaStream_t s) {
    switch (rule_id) {
        case RULE_LIFE: return launch_random_rule<StaticRule<kLifeB, kLifeS>>(shape, p, key, step, thr, s);
        case RULE_MORLEY: return launch_random_rule<StaticRule<kMorleyB, kMorleyS>>(shape, p, key, step, thr, s);
        case RULE_HIGHLIFE: return launch_random_rule<StaticRule<kHighB, kHighS>>(shape, p, key, step, thr, s);
        case RULE_DAYNIGHT: return launch_random_rule<StaticRule<kDayNightB, kDayNightS>>(shape, p, key, step, thr, s);
        default: return launch_random_rule<DynamicRule>(shape, p, key, step, thr, s);
    }
}

}  // namespace carle
