// jit.cu — run-time specialisation of the per-step kernels for ARBITRARY Life-like rules.
//
// The library ships compile-time instantiations for four named rules (B3/S23, B368/S245, B36/S23,
// B3678/S34678: 4-7 LOP3 per word for the rule function) and a run-time-rule path for the other
// 2^18 - 4 (a branch-free AND-OR over 18 mask words, ~29 LOP3 per word: 25-40 % slower per step,
// profiles/r1c_ab_dynamic_rule.txt).  Here the same kernel templates are instantiated for
// StaticRule<birth, survive> with NVRTC the first time a rule is stepped on a device, loaded as a
// CUBIN through the driver API and cached; every later step of that rule launches the
// specialised kernel (also under stream capture).  libnvrtc / libcuda are dlopen'ed: if either is
// missing, or CARLE_JIT=0, the run-time-rule kernels are used -- still on the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "abi_internal.h"

namespace carle {
namespace {

// the kernel sources, embedded by carle_b200/build.py
#include "embedded_sources.inc"

struct Nvrtc {
    void* lib = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    bool ok = false;
};

const Nvrtc& nvrtc() {
    static const Nvrtc n = [] {
        Nvrtc r;
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* name : names) {
            r.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (r.lib) break;
        }
        if (!r.lib) return r;
#define CARLE_SYM(field, sym) r.field = reinterpret_cast<decltype(r.field)>(dlsym(r.lib, sym))
        CARLE_SYM(CreateProgram, "nvrtcCreateProgram");
        CARLE_SYM(DestroyProgram, "nvrtcDestroyProgram");
        CARLE_SYM(AddNameExpression, "nvrtcAddNameExpression");
        CARLE_SYM(CompileProgram, "nvrtcCompileProgram");
        CARLE_SYM(GetLoweredName, "nvrtcGetLoweredName");
        CARLE_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        CARLE_SYM(GetCUBIN, "nvrtcGetCUBIN");
        CARLE_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        CARLE_SYM(GetProgramLog, "nvrtcGetProgramLog");
#undef CARLE_SYM
        r.ok = r.CreateProgram && r.DestroyProgram && r.AddNameExpression && r.CompileProgram &&
               r.GetLoweredName && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize &&
               r.GetProgramLog;
        return r;
    }();
    return n;
}

struct Driver {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    bool ok = false;
};

const Driver& driver() {
    static const Driver d = [] {
        Driver r;
        auto get = [](const char* name) -> void* {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult st;
            if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess ||
                st != cudaDriverEntryPointSuccess)
                return nullptr;
            return fn;
        };
        r.ModuleLoadData = reinterpret_cast<decltype(r.ModuleLoadData)>(get("cuModuleLoadData"));
        r.ModuleGetFunction = reinterpret_cast<decltype(r.ModuleGetFunction)>(get("cuModuleGetFunction"));
        r.FuncSetAttribute = reinterpret_cast<decltype(r.FuncSetAttribute)>(get("cuFuncSetAttribute"));
        r.LaunchKernelEx = reinterpret_cast<decltype(r.LaunchKernelEx)>(get("cuLaunchKernelEx"));
        r.OccupancyMaxActiveBlocksPerMultiprocessor =
            reinterpret_cast<decltype(r.OccupancyMaxActiveBlocksPerMultiprocessor)>(
                get("cuOccupancyMaxActiveBlocksPerMultiprocessor"));
        r.ok = r.ModuleLoadData && r.ModuleGetFunction && r.FuncSetAttribute && r.LaunchKernelEx &&
               r.OccupancyMaxActiveBlocksPerMultiprocessor;
        return r;
    }();
    return d;
}

struct Cubin { std::vector<char> image; std::string lowered; };
std::mutex g_mu;
std::map<std::string, CUfunction> g_cache;      // "<device>|<instantiation>" -> function (or null)
std::map<std::string, Cubin> g_cubins;          // instantiation -> compiled image (empty: failed)

}  // namespace

bool jit_enabled() {
    const char* e = getenv("CARLE_JIT");
    return !(e && e[0] == '0');
}

int jit_compile(const char* instantiation, std::vector<char>* cubin, std::string* lowered,
                std::string* log) {
    const Nvrtc& n = nvrtc();
    if (!n.ok) {
        if (log) *log = "libnvrtc is not available";
        return -1;
    }
    static const char* kSource = "#include \"strip.cuh\"\n#include \"tiled.cuh\"\n";
    const char* headers[] = {kSrcCaCore, kSrcKernels, kSrcStrip, kSrcTiled};
    const char* names[] = {"ca_core.cuh", "kernels.cuh", "strip.cuh", "tiled.cuh"};
    nvrtcProgram prog;
    if (n.CreateProgram(&prog, kSource, "carle_jit.cu", 4, headers, names) != NVRTC_SUCCESS) {
        if (log) *log = "nvrtcCreateProgram failed";
        return -1;
    }
    int rc = -1;
    do {
        if (n.AddNameExpression(prog, instantiation) != NVRTC_SUCCESS) {
            if (log) *log = "nvrtcAddNameExpression failed";
            break;
        }
        // (-default-device: the headers' plain constexpr helpers become __device__ functions)
        const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-DNDEBUG"};
        const nvrtcResult cr = n.CompileProgram(prog, 4, opts);
        if (cr != NVRTC_SUCCESS) {
            size_t sz = 0;
            n.GetProgramLogSize(prog, &sz);
            std::string text(sz, '\0');
            if (sz) n.GetProgramLog(prog, &text[0]);
            if (log) *log = "NVRTC: " + text;
            break;
        }
        const char* low = nullptr;
        if (n.GetLoweredName(prog, instantiation, &low) != NVRTC_SUCCESS || !low) {
            if (log) *log = "nvrtcGetLoweredName failed";
            break;
        }
        if (lowered) *lowered = low;
        size_t sz = 0;
        if (n.GetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
            if (log) *log = "nvrtcGetCUBINSize failed";
            break;
        }
        cubin->resize(sz);
        if (n.GetCUBIN(prog, cubin->data()) != NVRTC_SUCCESS) {
            if (log) *log = "nvrtcGetCUBIN failed";
            break;
        }
        rc = 0;
    } while (0);
    n.DestroyProgram(&prog);
    return rc;
}

int jit_loaded() {
    std::lock_guard<std::mutex> lock(g_mu);
    int n = 0;
    for (const auto& kv : g_cache) n += kv.second != nullptr;
    return n;
}

void* jit_kernel(int device, const std::string& instantiation) {
    if (!jit_enabled()) return nullptr;
    const std::string key = std::to_string(device) + "|" + instantiation;
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    const Driver& d = driver();
    if (!d.ok) return g_cache[key] = nullptr;
    // compile once per instantiation (all devices share the CUBIN); a failed compilation is
    // remembered, a failed LOAD is not: loading is refused while the calling thread captures a
    // CUDA graph, and must succeed on the next ordinary call
    auto cb = g_cubins.find(instantiation);
    if (cb == g_cubins.end()) {
        Cubin c;
        std::string log;
        if (jit_compile(instantiation.c_str(), &c.image, &c.lowered, &log) != 0) {
            c.image.clear();
            if (getenv("CARLE_JIT_VERBOSE"))
                fprintf(stderr, "carle_b200: JIT of %s failed: %s\n", instantiation.c_str(), log.c_str());
        }
        cb = g_cubins.emplace(instantiation, std::move(c)).first;
    }
    if (cb->second.image.empty()) return g_cache[key] = nullptr;
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    if (d.ModuleLoadData(&mod, cb->second.image.data()) != CUDA_SUCCESS ||
        d.ModuleGetFunction(&fn, mod, cb->second.lowered.c_str()) != CUDA_SUCCESS) {
        if (getenv("CARLE_JIT_VERBOSE"))
            fprintf(stderr, "carle_b200: loading the JIT image of %s failed (will retry)\n",
                    instantiation.c_str());
        return nullptr;
    }
    return g_cache[key] = fn;
}

cudaError_t jit_launch_grid(void* function, long long blocks, int threads, size_t smem, bool pdl,
                            const void* params, cudaStream_t s, const void* params2) {
    const Driver& d = driver();
    CUfunction fn = static_cast<CUfunction>(function);
    if (smem > 48 * 1024 &&
        d.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDimX = (unsigned)blocks; cfg.gridDimY = 1; cfg.gridDimZ = 1;
    cfg.blockDimX = (unsigned)threads; cfg.blockDimY = 1; cfg.blockDimZ = 1;
    cfg.sharedMemBytes = (unsigned)smem;
    cfg.hStream = reinterpret_cast<CUstream>(s);
    CUlaunchAttribute attr[1];
    memset(attr, 0, sizeof(attr));
    attr[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr[0].value.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    void* args[] = {const_cast<void*>(params), const_cast<void*>(params2)};
    return d.LaunchKernelEx(&cfg, fn, args, nullptr) == CUDA_SUCCESS ? cudaSuccess
                                                                     : cudaErrorLaunchFailure;
}

cudaError_t jit_launch(void* function, int sm_count, int threads, size_t smem, long long max_blocks,
                       int block_multiple, bool pdl, StepParams p, long long units,
                       cudaStream_t s, const void* params2) {
    const Driver& d = driver();
    CUfunction fn = static_cast<CUfunction>(function);
    if (d.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    int ctas = 0;
    if (d.OccupancyMaxActiveBlocksPerMultiprocessor(&ctas, fn, threads, smem) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    if (ctas < 1) ctas = 1;
    long long blocks = (long long)sm_count * ctas;
    if (blocks > max_blocks) blocks = max_blocks;
    blocks -= blocks % block_multiple;
    if (blocks < block_multiple) blocks = block_multiple;
    p.rank_blocked = rank_blocked_for(units, blocks * (threads / 32));
    walk_policy(p);
    return jit_launch_grid(function, blocks, threads, smem, pdl, &p, s, params2);
}

}  // namespace carle
