// host_pack.cu (host code only) — bit-packing of an unpacked action ON THE HOST, before it crosses the bus.
//
// The reference's agents produce float32 actions in host memory (carle/agents.py:35-42:
// `1.0 * (torch.rand(...) <= 0.1)` on the CPU) and CARLE.apply_action moves them to the device
// (carle/env.py:158-160): 4 bytes per toggle over PCIe, 256 MiB per step at 16384 x 64x64 windows --
// 5 ms at 52 GB/s, fifty times the step kernel.  carle_pack_action_host turns the same host tensor
// into the library's grid-aligned packed words (1 bit per toggle, include/carle_b200.h) with a small
// persistent pool of host threads (AVX-512 / AVX2 / SSE2 compare + mask, memory bound), so 8 MiB cross the bus
// instead.  It also reports what the reference's predicates need: some element != 0, some element
// != 1.0, and "some element is neither 0 nor 1" -- in which case the caller falls back to shipping
// the floats, because only the device path evaluates mean(action) == 1.0 on non-binary values.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/carle_b200.h"
#include "abi_internal.h"

namespace {

struct Job {
    const void* action;
    int u8;
    int64_t batch;
    int aw, ah, awpr, bit0;
    int avx2;
    int flat;                 // 0: entry by entry; 1 / 2: the flat path with AVX2 / AVX-512 (the flat path below)
    int streams, prefetch;    // flat path: interleaved streams per thread, prefetch distance in bytes
    uint32_t* out;
    std::atomic<int64_t> next{0};
    std::atomic<uint32_t> any{0}, not_one{0}, nonbin{0};
    // copy-as-you-pack (carle_pack_action_host_copy): the packed words go to `dev` in `nslices` pieces, each
    // enqueued on `stream` by the thread that packs the piece's last grab
    int64_t units = 0, grab = 1;          // flat: elements, entry by entry: entries; per fetch of `next`
    double words_per_unit = 0;            // 1/32 (flat) or aw * awpr
    uint32_t* dev = nullptr;
    int device = -1;
    cudaStream_t stream = nullptr;
    int64_t grabs_per_slice = 0;
    int nslices = 0;
    std::atomic<int64_t>* slice_left = nullptr;
    std::atomic<int> cuda_err{0};
    std::thread::id caller = std::this_thread::get_id();
};

// grab `g` of the job is packed: the thread that completes a slice sends it
void grab_done(Job& j, int64_t g) {
    if (!j.dev) return;
    const int64_t s = g / j.grabs_per_slice;
    // (release: this thread's words; acquire: the last one sees everybody's before it starts the DMA)
    if (j.slice_left[s].fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    const int64_t u0 = s * j.grabs_per_slice * j.grab;
    int64_t u1 = (s + 1) * j.grabs_per_slice * j.grab;
    if (u1 > j.units) u1 = j.units;
    const int64_t w0 = (int64_t)(u0 * j.words_per_unit), w1 = (int64_t)(u1 * j.words_per_unit);
    // the current device is per-thread state: the caller's thread gets its own back, a pool thread keeps this one
    int prev = -1;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess && prev != j.device) e = cudaSetDevice(j.device);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(j.dev + w0, j.out + w0, (size_t)(w1 - w0) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            j.stream);
    if (prev >= 0 && prev != j.device && j.caller == std::this_thread::get_id()) cudaSetDevice(prev);
    if (e != cudaSuccess) j.cuda_err.store((int)e, std::memory_order_relaxed);
}

// one action entry [aw][ah] -> [aw][awpr] words; bit (bit0 + c) of the row's word string = element c != 0
template <typename T>
void pack_entry(const T* a, const Job& j, uint32_t* out, uint32_t& any, uint32_t& not_one, uint32_t& nonbin);

template <>
void pack_entry<float>(const float* a, const Job& j, uint32_t* out, uint32_t& any, uint32_t& not_one,
                       uint32_t& nonbin) {
    const __m128 zero = _mm_setzero_ps(), one = _mm_set1_ps(1.0f);
    for (int r = 0; r < j.aw; ++r, a += j.ah, out += j.awpr) {
        uint64_t acc = 0;                       // bits waiting to be written, `fill` of them valid
        int fill = j.bit0, word = 0;
        int c = 0;
        for (; c + 4 <= j.ah; c += 4) {
            const __m128 v = _mm_loadu_ps(a + c);
            const uint32_t nz = (uint32_t)_mm_movemask_ps(_mm_cmpneq_ps(v, zero));
            const uint32_t eq1 = (uint32_t)_mm_movemask_ps(_mm_cmpeq_ps(v, one));
            any |= nz;
            not_one |= eq1 ^ 0xFu;
            nonbin |= nz ^ eq1;                 // non-zero and not 1.0 (NaN: nz = 1, eq1 = 0)
            acc |= (uint64_t)nz << fill;
            fill += 4;
            if (fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        for (; c < j.ah; ++c) {
            const float v = a[c];
            const uint32_t nz = v != 0.0f, eq1 = v == 1.0f;
            any |= nz; not_one |= eq1 ^ 1u; nonbin |= nz ^ eq1;
            acc |= (uint64_t)nz << fill;
            if (++fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        if (word < j.awpr) out[word++] = (uint32_t)acc;
        for (; word < j.awpr; ++word) out[word] = 0u;
    }
}

// the same with 8 floats per instruction and the three flags kept in vector registers until the end of
// the entry (chosen at run time when the CPU has AVX2)
__attribute__((target("avx2")))
void pack_entry_avx2(const float* a, const Job& j, uint32_t* out, uint32_t& any, uint32_t& not_one,
                     uint32_t& nonbin) {
    const __m256 zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.0f);
    __m256 v_any = zero, v_all = _mm256_castsi256_ps(_mm256_set1_epi32(-1)), v_nb = zero;
    for (int r = 0; r < j.aw; ++r, a += j.ah, out += j.awpr) {
        uint64_t acc = 0;
        int fill = j.bit0, word = 0;
        int c = 0;
        if (j.bit0 == 0) {
            // word-aligned window (64x64 on 256x256, 32x32 on 128x128): 32 floats -> one output word
            for (; c + 32 <= j.ah; c += 32) {
                uint32_t w = 0;
                for (int q = 0; q < 4; ++q) {
                    const __m256 v = _mm256_loadu_ps(a + c + 8 * q);
                    const __m256 nz = _mm256_cmp_ps(v, zero, _CMP_NEQ_UQ), eq1 = _mm256_cmp_ps(v, one, _CMP_EQ_OQ);
                    v_any = _mm256_or_ps(v_any, nz);
                    v_all = _mm256_and_ps(v_all, eq1);
                    v_nb = _mm256_or_ps(v_nb, _mm256_xor_ps(nz, eq1));
                    w |= (uint32_t)_mm256_movemask_ps(nz) << (8 * q);
                }
                out[word++] = w;
            }
        }
        for (; c + 8 <= j.ah; c += 8) {
            const __m256 v = _mm256_loadu_ps(a + c);
            const __m256 nz = _mm256_cmp_ps(v, zero, _CMP_NEQ_UQ), eq1 = _mm256_cmp_ps(v, one, _CMP_EQ_OQ);
            v_any = _mm256_or_ps(v_any, nz);
            v_all = _mm256_and_ps(v_all, eq1);
            v_nb = _mm256_or_ps(v_nb, _mm256_xor_ps(nz, eq1));
            acc |= (uint64_t)(uint32_t)_mm256_movemask_ps(nz) << fill;
            fill += 8;
            if (fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        for (; c < j.ah; ++c) {
            const float v = a[c];
            const uint32_t nz = v != 0.0f, eq1 = v == 1.0f;
            any |= nz; not_one |= eq1 ^ 1u; nonbin |= nz ^ eq1;
            acc |= (uint64_t)nz << fill;
            if (++fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        if (word < j.awpr) out[word++] = (uint32_t)acc;
        for (; word < j.awpr; ++word) out[word] = 0u;
    }
    if (j.ah >= 8) {
        any |= _mm256_movemask_ps(v_any) != 0;
        not_one |= _mm256_movemask_ps(v_all) != 0xFF;
        nonbin |= _mm256_movemask_ps(v_nb) != 0;
    }
}

template <>
void pack_entry<uint8_t>(const uint8_t* a, const Job& j, uint32_t* out, uint32_t& any, uint32_t& not_one,
                         uint32_t& nonbin) {
    // (uint8 actions: "all elements == 1" / "some element != 0" on the device; a byte above 1 toggles but is
    //  not 1, which packed words cannot say: flagged like a non-binary float)
    const __m128i zero = _mm_setzero_si128(), one = _mm_set1_epi8(1);
    for (int r = 0; r < j.aw; ++r, a += j.ah, out += j.awpr) {
        uint64_t acc = 0;
        int fill = j.bit0, word = 0;
        int c = 0;
        for (; c + 16 <= j.ah; c += 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(a + c));
            const uint32_t nz = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, zero)) ^ 0xFFFFu;
            const uint32_t eq1 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, one));
            any |= nz;
            not_one |= eq1 ^ 0xFFFFu;
            nonbin |= nz ^ eq1;
            acc |= (uint64_t)nz << fill;
            fill += 16;
            if (fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        for (; c < j.ah; ++c) {
            const uint32_t nz = a[c] != 0;
            any |= nz; not_one |= (a[c] != 1); nonbin |= (a[c] > 1);
            acc |= (uint64_t)nz << fill;
            if (++fill >= 32) { out[word++] = (uint32_t)acc; acc >>= 32; fill -= 32; }
        }
        if (word < j.awpr) out[word++] = (uint32_t)acc;
        for (; word < j.awpr; ++word) out[word] = 0u;
    }
}

// ---- the flat path -------------------------------------------------------------------------------------
// A word-aligned window whose rows are whole words (bit0 == 0, ah % 32 == 0: 64x64 on 256x256, 32x32 on
// 128x128) makes the whole action ONE string of elements that maps onto ONE string of words, 32 elements
// per word, whatever the entry and row structure.  A thread that walks such a string front to back is bound
// by the latency of its one stream of cache misses (~7 GB/s per core measured: the hardware prefetcher
// stops at every 4 KiB page).  pack_flat therefore cuts each piece of work into `streams` equal runs
// and advances them together, 256 bytes of each per trip, with a software prefetch `prefetch` bytes ahead in
// every run: several pages in flight per core, 1.5x per thread on the same memory.
struct FlatFlags {
    uint32_t any = 0, all = 0xFFFFFFFFu, nb = 0;      // OR of the toggle words, AND of the ==1 words, OR of their XOR
};

// 64 floats -> two words
__attribute__((target("avx512f,avx512bw")))
inline void f32_block_avx512(const float* p, uint32_t* o, FlatFlags& f) {
    const __m512i absmask = _mm512_set1_epi32(0x7fffffff), one = _mm512_set1_epi32(0x3f800000);
    const __m512i v0 = _mm512_loadu_si512(p), v1 = _mm512_loadu_si512(p + 16);
    const __m512i v2 = _mm512_loadu_si512(p + 32), v3 = _mm512_loadu_si512(p + 48);
    // != 0 (either zero; a NaN toggles, as `v != 0` says) and == 1.0f (which has one encoding) on the raw bits
    const uint32_t w0 = (uint32_t)_mm512_test_epi32_mask(v0, absmask) | ((uint32_t)_mm512_test_epi32_mask(v1, absmask) << 16);
    const uint32_t w1 = (uint32_t)_mm512_test_epi32_mask(v2, absmask) | ((uint32_t)_mm512_test_epi32_mask(v3, absmask) << 16);
    const uint32_t q0 = (uint32_t)_mm512_cmpeq_epi32_mask(v0, one) | ((uint32_t)_mm512_cmpeq_epi32_mask(v1, one) << 16);
    const uint32_t q1 = (uint32_t)_mm512_cmpeq_epi32_mask(v2, one) | ((uint32_t)_mm512_cmpeq_epi32_mask(v3, one) << 16);
    f.any |= w0 | w1; f.all &= q0 & q1; f.nb |= (w0 ^ q0) | (w1 ^ q1);
    o[0] = w0; o[1] = w1;
}

__attribute__((target("avx2")))
inline uint32_t f32_word_avx2(const float* p, uint32_t& q) {
    const __m256i absmask = _mm256_set1_epi32(0x7fffffff), one = _mm256_set1_epi32(0x3f800000);
    const __m256i zero = _mm256_setzero_si256();
    uint32_t z = 0;
    q = 0;
    for (int k = 0; k < 4; ++k) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8 * k));
        z |= (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(_mm256_and_si256(v, absmask), zero))) << (8 * k);
        q |= (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, one))) << (8 * k);
    }
    return ~z;
}

__attribute__((target("avx2")))
inline void f32_block_avx2(const float* p, uint32_t* o, FlatFlags& f) {
    uint32_t q0, q1;
    const uint32_t w0 = f32_word_avx2(p, q0), w1 = f32_word_avx2(p + 32, q1);
    f.any |= w0 | w1; f.all &= q0 & q1; f.nb |= (w0 ^ q0) | (w1 ^ q1);
    o[0] = w0; o[1] = w1;
}

// 256 bytes -> eight words
__attribute__((target("avx512f,avx512bw")))
inline void u8_block_avx512(const uint8_t* p, uint32_t* o, FlatFlags& f) {
    const __m512i one = _mm512_set1_epi8(1);
    for (int k = 0; k < 4; ++k) {
        const __m512i v = _mm512_loadu_si512(p + 64 * k);
        const uint64_t w = _mm512_test_epi8_mask(v, v), q = _mm512_cmpeq_epi8_mask(v, one);
        const uint64_t x = w ^ q;
        f.any |= (uint32_t)(w | (w >> 32)); f.all &= (uint32_t)(q & (q >> 32)); f.nb |= (uint32_t)(x | (x >> 32));
        o[2 * k] = (uint32_t)w; o[2 * k + 1] = (uint32_t)(w >> 32);
    }
}

__attribute__((target("avx2")))
inline void u8_block_avx2(const uint8_t* p, uint32_t* o, FlatFlags& f) {
    const __m256i zero = _mm256_setzero_si256(), one = _mm256_set1_epi8(1);
    for (int k = 0; k < 8; ++k) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 32 * k));
        const uint32_t w = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero));
        const uint32_t q = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, one));
        f.any |= w; f.all &= q; f.nb |= w ^ q;
        o[k] = w;
    }
}

// elements [0, m) of a piece of the flat string, m a multiple of 32; 256 bytes per block (float: 64, uint8: 256
// elements).  (A macro, not a template: the loop has to be compiled inside a function that carries the block's
// target attribute, or the compiler will not inline the block into it.)
#define CARLE_FLAT_PIECE(T, BLOCK)                                                                              \
    constexpr int kBlock = 256 / (int)sizeof(T);                                                                \
    const int64_t run = m / ((int64_t)streams * kBlock) * kBlock;         /* elements per stream */             \
    for (int64_t c = 0; c < run; c += kBlock)                                                                   \
        for (int s = 0; s < streams; ++s) {                                                                     \
            const T* p = a + s * run + c;                                                                       \
            if (prefetch)                                                                                       \
                for (int k = 0; k < 4; ++k)                                                                     \
                    _mm_prefetch(reinterpret_cast<const char*>(p) + prefetch + 64 * k, _MM_HINT_T0);            \
            BLOCK(p, out + ((s * run + c) >> 5), f);                                                            \
        }                                                                                                       \
    for (int64_t c = (int64_t)streams * run; c < m; c += 32) {           /* the streams' left-over, by word */  \
        uint32_t w = 0, q = 0;                                                                                  \
        for (int k = 0; k < 32; ++k) {                                                                          \
            w |= (uint32_t)(a[c + k] != (T)0) << k;                                                             \
            q |= (uint32_t)(a[c + k] == (T)1) << k;                                                             \
        }                                                                                                       \
        f.any |= w; f.all &= q; f.nb |= w ^ q;                                                                  \
        out[c >> 5] = w;                                                                                        \
    }

__attribute__((target("avx512f,avx512bw")))
void flat_f32_avx512(const float* a, uint32_t* out, int64_t m, int streams, int prefetch, FlatFlags& f) {
    CARLE_FLAT_PIECE(float, f32_block_avx512)
}
__attribute__((target("avx2")))
void flat_f32_avx2(const float* a, uint32_t* out, int64_t m, int streams, int prefetch, FlatFlags& f) {
    CARLE_FLAT_PIECE(float, f32_block_avx2)
}
__attribute__((target("avx512f,avx512bw")))
void flat_u8_avx512(const uint8_t* a, uint32_t* out, int64_t m, int streams, int prefetch, FlatFlags& f) {
    CARLE_FLAT_PIECE(uint8_t, u8_block_avx512)
}
__attribute__((target("avx2")))
void flat_u8_avx2(const uint8_t* a, uint32_t* out, int64_t m, int streams, int prefetch, FlatFlags& f) {
    CARLE_FLAT_PIECE(uint8_t, u8_block_avx2)
}
#undef CARLE_FLAT_PIECE

void run_job_flat(Job& j) {
    const int64_t total = j.units, grab = j.grab;                          // elements; multiples of 32
    FlatFlags f;
    for (;;) {
        const int64_t e0 = j.next.fetch_add(grab, std::memory_order_relaxed);
        if (e0 >= total) break;
        const int64_t m = e0 + grab < total ? grab : total - e0;
        uint32_t* out = j.out + (e0 >> 5);
        if (j.u8) {
            const uint8_t* a = static_cast<const uint8_t*>(j.action) + e0;
            if (j.flat == 2) flat_u8_avx512(a, out, m, j.streams, j.prefetch, f);
            else flat_u8_avx2(a, out, m, j.streams, j.prefetch, f);
        } else {
            const float* a = static_cast<const float*>(j.action) + e0;
            if (j.flat == 2) flat_f32_avx512(a, out, m, j.streams, j.prefetch, f);
            else flat_f32_avx2(a, out, m, j.streams, j.prefetch, f);
        }
        grab_done(j, e0 / grab);
    }
    if (f.any) j.any.store(1, std::memory_order_relaxed);
    if (f.all != 0xFFFFFFFFu) j.not_one.store(1, std::memory_order_relaxed);
    if (f.nb) j.nonbin.store(1, std::memory_order_relaxed);
}

void run_job(Job& j) {
    if (j.flat) { run_job_flat(j); return; }
    const int64_t chunk = j.grab;               // entries per grab
    const size_t entry = (size_t)j.aw * j.ah, entry_out = (size_t)j.aw * j.awpr;
    uint32_t any = 0, not_one = 0, nonbin = 0;
    for (;;) {
        const int64_t b0 = j.next.fetch_add(chunk, std::memory_order_relaxed);
        if (b0 >= j.batch) break;
        const int64_t b1 = b0 + chunk < j.batch ? b0 + chunk : j.batch;
        for (int64_t b = b0; b < b1; ++b) {
            if (j.u8)
                pack_entry<uint8_t>(static_cast<const uint8_t*>(j.action) + b * entry, j, j.out + b * entry_out,
                                    any, not_one, nonbin);
            else if (j.avx2)
                pack_entry_avx2(static_cast<const float*>(j.action) + b * entry, j, j.out + b * entry_out,
                                any, not_one, nonbin);
            else
                pack_entry<float>(static_cast<const float*>(j.action) + b * entry, j, j.out + b * entry_out,
                                  any, not_one, nonbin);
        }
        grab_done(j, b0 / chunk);
    }
    if (any) j.any.store(1, std::memory_order_relaxed);
    if (not_one) j.not_one.store(1, std::memory_order_relaxed);
    if (nonbin) j.nonbin.store(1, std::memory_order_relaxed);
}

// persistent workers: woken per call, parked on a condition variable in between
class Pool {
  public:
    void run(Job& job, int threads) {
        std::unique_lock<std::mutex> call(call_mu_);          // one packing call at a time
        if (threads < 1) threads = 1;
        while ((int)workers_.size() < threads - 1) workers_.emplace_back([this] { loop(); });
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &job;
            wanted_ = threads - 1;
            started_ = 0;
            finished_ = 0;
            ++epoch_;
        }
        cv_.notify_all();
        run_job(job);
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return finished_ == wanted_; });
        job_ = nullptr;
    }

  private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            Job* job;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return epoch_ != seen && started_ < wanted_; });
                seen = epoch_;
                ++started_;
                job = job_;
            }
            run_job(*job);
            {
                std::lock_guard<std::mutex> g(mu_);
                ++finished_;
            }
            done_cv_.notify_one();
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> workers_;
    Job* job_ = nullptr;
    int wanted_ = 0, started_ = 0, finished_ = 0;
    uint64_t epoch_ = 0;
};

Pool& pool() {
    static Pool* p = new Pool();                // (never destroyed: the workers are detached in spirit)
    return *p;
}

}  // namespace

namespace {

int pack_host(int32_t aw, int32_t ah, int32_t awpr, int32_t bit0, const void* action_host, int dtype, int64_t batch,
              uint32_t* packed_host, int32_t* flags, int32_t threads, uint32_t* packed_device, int device,
              void* stream, const char* who) {
    const std::string name(who);
    if (!action_host || !packed_host || !flags) return carle::abi_fail(CARLE_EINVAL, name + ": NULL argument");
    if (dtype != CARLE_F32 && dtype != CARLE_U8)
        return carle::abi_fail(CARLE_EINVAL, name + ": dtype must be CARLE_F32 or CARLE_U8");
    if (batch < 1) return carle::abi_fail(CARLE_EINVAL, name + ": empty batch");
    if (aw < 0 || ah < 0 || bit0 < 0 || bit0 > 31 || awpr != (ah > 0 ? (bit0 + ah + 31) / 32 : awpr))
        return carle::abi_fail(CARLE_EINVAL, name + ": geometry (see carle_geometry)");
    Job job;
    job.action = action_host;
    job.u8 = dtype == CARLE_U8;
    job.batch = batch;
    job.aw = aw; job.ah = ah; job.awpr = awpr;
    job.bit0 = bit0;
    job.out = packed_host;
    // CARLE_HOST_PACK_ISA = sse2 | avx2 | avx512 caps the instruction set (tests, A/B runs), CARLE_HOST_PACK_FLAT=0
    // keeps the entry-by-entry walk, CARLE_HOST_PACK_STREAMS / CARLE_HOST_PACK_PREFETCH tune the flat path
    const char* isa = getenv("CARLE_HOST_PACK_ISA");
    const int cap = !isa ? 3 : (isa[0] == 's' ? 0 : (strcmp(isa, "avx2") == 0 ? 1 : 3));
    job.avx2 = (cap >= 1 && __builtin_cpu_supports("avx2")) ? 1 : 0;
    const bool avx512 = cap >= 2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    job.flat = 0;
    if (job.avx2 && bit0 == 0 && ah > 0 && ah % 32 == 0 && carle::env_int("CARLE_HOST_PACK_FLAT", 1))
        job.flat = avx512 ? 2 : 1;
    job.streams = carle::env_int("CARLE_HOST_PACK_STREAMS", 4);
    if (job.streams < 1 || job.streams > 64) job.streams = 4;
    job.prefetch = carle::env_int("CARLE_HOST_PACK_PREFETCH", 4096);
    if (job.prefetch < 0 || job.prefetch > (1 << 20)) job.prefetch = 4096;
    if (job.flat) {
        job.units = batch * (int64_t)aw * ah;
        job.grab = job.u8 ? (int64_t)1 << 19 : (int64_t)1 << 17;         // 512 KiB of input per grab
        job.words_per_unit = 1.0 / 32.0;
    } else {
        job.units = batch;
        job.grab = 8;
        job.words_per_unit = (double)aw * awpr;
    }
    std::unique_ptr<std::atomic<int64_t>[]> slice_left;
    if (job.aw > 0 && job.ah > 0) {
        if (packed_device) {
            // one slice per MiB of packed words, eight at most: the last slice's copy is all that is left
            // to do when the packing ends (CARLE_HOST_PACK_SLICES forces a count: tests)
            const int64_t grabs = (job.units + job.grab - 1) / job.grab;
            const int64_t words = batch * (int64_t)aw * awpr;
            int64_t want = carle::env_int("CARLE_HOST_PACK_SLICES", 0);
            if (want < 1) want = words * 4 / (1 << 20);
            want = want < 1 ? 1 : (want > 8 ? 8 : want);
            job.grabs_per_slice = (grabs + want - 1) / want;
            job.nslices = (int)((grabs + job.grabs_per_slice - 1) / job.grabs_per_slice);
            slice_left.reset(new std::atomic<int64_t>[job.nslices]);
            for (int s = 0; s < job.nslices; ++s) {
                const int64_t g1 = (s + 1) * job.grabs_per_slice < grabs ? (s + 1) * job.grabs_per_slice : grabs;
                slice_left[s].store(g1 - s * job.grabs_per_slice, std::memory_order_relaxed);
            }
            job.slice_left = slice_left.get();
            job.dev = packed_device;
            job.device = device;
            job.stream = static_cast<cudaStream_t>(stream);
        }
        if (batch * (int64_t)job.aw * job.ah < (1 << 16)) threads = 1;       // small: not worth a wake-up
        pool().run(job, threads);
    }
    flags[0] = (int32_t)job.not_one.load();
    flags[1] = (int32_t)job.any.load();
    flags[2] = (int32_t)job.nonbin.load();
    if (job.cuda_err.load())
        return carle::abi_fail(CARLE_ECUDA, name + ": " + cudaGetErrorString((cudaError_t)job.cuda_err.load()));
    return CARLE_OK;
}

}  // namespace

extern "C" CARLE_API int carle_pack_action_host(int32_t aw, int32_t ah, int32_t awpr, int32_t bit0,
                                                const void* action_host, int dtype, int64_t batch,
                                                uint32_t* packed_host, int32_t* flags, int32_t threads) {
    return pack_host(aw, ah, awpr, bit0, action_host, dtype, batch, packed_host, flags, threads, nullptr, -1, nullptr,
                     "carle_pack_action_host");
}

extern "C" CARLE_API int carle_pack_action_host_copy(int32_t aw, int32_t ah, int32_t awpr, int32_t bit0,
                                                     const void* action_host, int dtype, int64_t batch,
                                                     uint32_t* packed_host, int32_t* flags, int32_t threads,
                                                     uint32_t* packed_device, int32_t device, void* stream) {
    if (!packed_device) return carle::abi_fail(CARLE_EINVAL, "carle_pack_action_host_copy: NULL device buffer");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return carle::abi_fail(CARLE_ENODEV, "carle_pack_action_host_copy: no such CUDA device");
    return pack_host(aw, ah, awpr, bit0, action_host, dtype, batch, packed_host, flags, threads, packed_device, device,
                     stream, "carle_pack_action_host_copy");
}
