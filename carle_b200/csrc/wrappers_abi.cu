// wrappers_abi.cu — the SURVEY §8(f) rows beside the step: PufferDetector's sliding window kept on the
// device (carle/mcl.py:828-850), MorphoBonus as a bit-parallel template match on packed rows
// (carle/mcl.py:107-195) and the RLE codec on packed words (carle/env.py:260-328, 408-464; host side,
// it is the on-disk format).  Entry points declared in include/carle_b200.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>

#include "../../include/carle_b200.h"
#include "abi_internal.h"

namespace {

#define W_CUDA_TRY(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return carle::abi_fail(CARLE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------
// PufferDetector tail.  state[0] = entries in the window, [1] = index of the oldest, [2] = running
// total of the step's live cells (zero between calls), [3] = blocks retired (zero between calls),
// [4] = live total of the last step, [5] = 1 when the last step paid the bonus, [6] = bonuses paid.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
puffer_tail_kernel(const long long* __restrict__ red, long long n, const long long* __restrict__ counters,
                   long long* __restrict__ ring, long long* __restrict__ state, int threshold,
                   float* __restrict__ reward) {
    long long part = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        part += red[4 * i + CARLE_RED_LIVE];
    for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    __shared__ long long warp_part[8];
    __shared__ int last, fire;
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += warp_part[w];
        atomicAdd(reinterpret_cast<unsigned long long*>(state + 2), (unsigned long long)total);
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned long long*>(state + 3), 1ull) == gridDim.x - 1;
        fire = 0;
        if (last) {
            __threadfence();
            const long long live = *reinterpret_cast<volatile long long*>(state + 2);
            state[2] = 0;
            state[3] = 0;
            state[4] = live;
            const int cap = threshold + 1;
            long long count = state[0], head = state[1];
            if (counters[CARLE_CNT_LAST_ANY_TOGGLE] == 0) {        // `if not(torch.sum(action))`, mcl.py:833
                ring[(head + count) % cap] = live;
                ++count;
                if (count > threshold) {                           // mcl.py:837-842
                    const long long slope = live - ring[head];
                    head = (head + 1) % cap;
                    --count;
                    fire = slope > 0;                              // integers: slope > 0.01 <=> slope >= 1
                }
            } else {                                               // mcl.py:845-847
                count = 0;
                head = 0;
            }
            state[0] = count;
            state[1] = head;
            state[5] = fire;
            state[6] += fire;
        }
    }
    __syncthreads();
    if (last && fire && reward)
        for (long long i = threadIdx.x; i < n; i += blockDim.x) reward[i] += 1.0f;
}

// ---------------------------------------------------------------------------------------------
// MorphoBonus: F.conv2d(grid, patterns) without padding, then max and min over patterns and
// positions per instance (mcl.py:176-185).  A pattern is 8x8 with weight +w on its live cells and
// -1 on the others, so a window X scores  w * |X & P| - |X & ~P|:  two population counts of 64-bit
// words.  One thread owns 32 horizontally adjacent window positions of one row.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPatterns = 64;

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void morpho_init_kernel(float* mx, float* mn, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        mx[i] = __int_as_float(0xff800000);      // -inf
        mn[i] = __int_as_float(0x7f800000);      // +inf
    }
}

__global__ void __launch_bounds__(256)
morpho_match_kernel(const uint32_t* __restrict__ state, const uint32_t* __restrict__ toggles,
                    long long toggle_stride, int h, int w, int wpr,
                    const unsigned long long* __restrict__ patterns, const float* __restrict__ weights,
                    int npat, float* __restrict__ out_max, float* __restrict__ out_min) {
    __shared__ unsigned long long s_pat[kMaxPatterns];
    __shared__ float s_w[kMaxPatterns];
    __shared__ float s_max[8], s_min[8];
    for (int p = threadIdx.x; p < npat; p += blockDim.x) {
        s_pat[p] = patterns[p];
        s_w[p] = weights[p];
    }
    __syncthreads();
    const long long inst = blockIdx.y;
    const uint32_t* u = state + inst * (long long)h * wpr;
    const uint32_t* t = toggles ? toggles + inst * toggle_stride : nullptr;
    const int rows_out = h - 7, cols_out = w - 7;
    float best = __int_as_float(0xff800000), worst = __int_as_float(0x7f800000);
    for (long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x; item < (long long)rows_out * wpr;
         item += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(item / wpr), wd = (int)(item % wpr);
        const int jn = min(32, cols_out - 32 * wd);
        if (jn <= 0) continue;
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long at = (long long)(i + r) * wpr + wd;
            lo[r] = u[at];
            hi[r] = wd + 1 < wpr ? u[at + 1] : 0u;
            if (t) {
                lo[r] ^= t[at];
                if (wd + 1 < wpr) hi[r] ^= t[at + 1];
            }
        }
        for (int j = 0; j < jn; ++j) {
            // byte r of the window = columns j .. j+7 of row i + r
            uint32_t b[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) b[r] = __funnelshift_r(lo[r], hi[r], j);
            const uint32_t x_lo = __byte_perm(__byte_perm(b[0], b[1], 0x0040), __byte_perm(b[2], b[3], 0x0040), 0x5410);
            const uint32_t x_hi = __byte_perm(__byte_perm(b[4], b[5], 0x0040), __byte_perm(b[6], b[7], 0x0040), 0x5410);
            const unsigned long long x = ((unsigned long long)x_hi << 32) | x_lo;
            const int total = __popcll(x);
            for (int p = 0; p < npat; ++p) {
                const int a = __popcll(x & s_pat[p]);
                const float score = s_w[p] * (float)a - (float)(total - a);
                best = fmaxf(best, score);
                worst = fminf(worst, score);
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        best = fmaxf(best, __shfl_down_sync(0xffffffffu, best, o));
        worst = fminf(worst, __shfl_down_sync(0xffffffffu, worst, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_max[threadIdx.x >> 5] = best;
        s_min[threadIdx.x >> 5] = worst;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            best = fmaxf(best, s_max[k]);
            worst = fminf(worst, s_min[k]);
        }
        if (best != __int_as_float(0xff800000)) {
            atomic_max_float(out_max + inst, best);
            atomic_min_float(out_min + inst, worst);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RLE on packed words (host).
// ---------------------------------------------------------------------------------------------
struct Sink {
    char* out;
    int64_t cap, len = 0;
    void put(const char* s, int64_t n) {
        if (out && len + n <= cap) memcpy(out + len, s, (size_t)n);
        len += n;
    }
};

inline int bit_at(const uint32_t* row, int c) { return (row[c >> 5] >> (c & 31)) & 1; }

// first column >= c whose cell differs from `v` (or w)
inline int run_end(const uint32_t* row, int c, int w, int v) {
    const int wpr = (w + 31) / 32;
    int word = c >> 5;
    uint32_t x = (v ? ~row[word] : row[word]) & (~0u << (c & 31));
    while (true) {
        if (x) {
            const int at = 32 * word + __builtin_ctz(x);
            return at < w ? at : w;
        }
        if (++word >= wpr) return w;
        x = v ? ~row[word] : row[word];
    }
}

}  // namespace

extern "C" {

CARLE_API int carle_puffer_tail(carle_handle_t h, const int64_t* reductions, const int64_t* counters,
                                int64_t* ring, int64_t* state, int32_t growth_threshold, float* reward,
                                void* stream) {
    if (!h || !reductions || !counters || !ring || !state)
        return carle::abi_fail(CARLE_EINVAL, "carle_puffer_tail: NULL argument");
    if (growth_threshold < 1) return carle::abi_fail(CARLE_EINVAL, "carle_puffer_tail: growth_threshold < 1");
    DevGuard guard(h->device);
    long long blocks = (h->n + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 2LL * h->sm_count) blocks = 2LL * h->sm_count;
    puffer_tail_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(reductions), h->n, reinterpret_cast<const long long*>(counters),
        reinterpret_cast<long long*>(ring), reinterpret_cast<long long*>(state), growth_threshold, reward);
    W_CUDA_TRY(cudaGetLastError());
    return CARLE_OK;
}

CARLE_API int carle_morpho_match(carle_handle_t h, const uint32_t* state, const uint32_t* toggles,
                                 int64_t toggle_batch, const uint64_t* patterns, const float* weights,
                                 int32_t n_patterns, float* out_max, float* out_min, void* stream) {
    if (!h || !state || !patterns || !weights || !out_max || !out_min)
        return carle::abi_fail(CARLE_EINVAL, "carle_morpho_match: NULL argument");
    if (n_patterns < 1 || n_patterns > kMaxPatterns)
        return carle::abi_fail(CARLE_EINVAL, "carle_morpho_match: 1..64 patterns");
    if (h->h < 8 || h->w < 8)
        return carle::abi_fail(CARLE_EINVAL, "carle_morpho_match: the grid is smaller than the 8x8 patterns");
    if (toggles && toggle_batch != 1 && toggle_batch != h->n)
        return carle::abi_fail(CARLE_EINVAL, "carle_morpho_match: toggle_batch must be 1 or N");
    DevGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long words = (long long)h->h * h->wpr;
    morpho_init_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, s>>>(out_max, out_min, h->n);
    W_CUDA_TRY(cudaGetLastError());
    const long long items = (long long)(h->h - 7) * h->wpr;
    long long bx = (items + 255) / 256;
    if (bx > 1024) bx = 1024;
    long long done = 0;
    while (done < h->n) {
        long long chunk = h->n - done;
        if (chunk > 65535) chunk = 65535;
        dim3 grid((unsigned)bx, (unsigned)chunk);
        morpho_match_kernel<<<grid, 256, 0, s>>>(
            state + done * words, toggles ? toggles + (toggle_batch == 1 ? 0 : done * words) : nullptr,
            toggle_batch == 1 ? 0 : words, h->h, h->w, h->wpr,
            reinterpret_cast<const unsigned long long*>(patterns), weights, n_patterns, out_max + done,
            out_min + done);
        W_CUDA_TRY(cudaGetLastError());
        done += chunk;
    }
    return CARLE_OK;
}

CARLE_API int64_t carle_rle_encode_host(const uint32_t* packed_host, int32_t height, int32_t width,
                                        int32_t flags, char* out, int64_t capacity) {
    if (!packed_host || height < 0 || width < 1 || capacity < 0)
        return carle::abi_fail(CARLE_EINVAL, "carle_rle_encode_host: bad argument");
    const int wpr = (width + 31) / 32;
    Sink sink{out, capacity};
    char line[160];
    int fill = 0;
    auto token = [&](int count, char tag, bool row_end) {
        fill += snprintf(line + fill, sizeof(line) - (size_t)fill, "%d%c", count, tag);
        if (row_end) line[fill++] = '$';
        if (fill > 69) {                                   // env.py:439-441, 449-451
            line[fill++] = '\n';
            sink.put(line, fill);
            fill = 0;
        }
    };
    for (int r = 0; r < height; ++r) {
        const uint32_t* row = packed_host + (size_t)r * wpr;
        int c = 0;
        while (c < width) {
            const int v = bit_at(row, c);
            const int e = run_end(row, c, width, v);
            token(e - c, v ? 'o' : 'b', e == width);
            c = e;
        }
    }
    if (flags & CARLE_RLE_KEEP_TAIL) sink.put(line, fill);   // upstream drops this partial line (env.py:453-455)
    sink.put("!", 1);
    return sink.len;
}

CARLE_API int carle_rle_decode_host(const char* text, int64_t length, int32_t height, int32_t width,
                                    uint32_t* packed_host_out) {
    if (!text || !packed_host_out || height < 0 || width < 1 || length < 0)
        return carle::abi_fail(CARLE_EINVAL, "carle_rle_decode_host: bad argument");
    const int wpr = (width + 31) / 32;
    memset(packed_host_out, 0, sizeof(uint32_t) * (size_t)height * wpr);
    long long row = 0, col = 0, count = 0;
    bool have = false;
    for (int64_t i = 0; i < length; ++i) {
        const char ch = text[i];
        if (ch >= '0' && ch <= '9') {
            count = count < (1LL << 40) ? count * 10 + (ch - '0') : count;
            have = true;
            continue;
        }
        if (ch == '\n' || ch == '\r') continue;            // a count may continue on the next line (env.py:277-278)
        const long long run = have ? count : 1;
        count = 0;
        have = false;
        if (ch == 'b' || ch == 'B') {
            col += run;
        } else if (ch == 'o' || ch == 'O') {
            if (row < height) {
                long long e = col + run < width ? col + run : width;
                for (long long c = col; c < e; ++c)
                    packed_host_out[row * wpr + (c >> 5)] |= 1u << (c & 31);
            }
            col += run;
        } else if (ch == '$') {
            row += run;
            col = 0;
        } else if (ch == '!') {
            break;
        }
    }
    return CARLE_OK;
}

}  // extern "C"
