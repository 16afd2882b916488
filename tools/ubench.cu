// Micro-benchmarks that calibrate the rooflines used in DESIGN.md / bench.py (run on the B200):
//   1. launch floor: empty kernels of the step kernel's grid shape, replayed from a CUDA graph
//   2. LOP3 issue peak (the integer-pipe roofline of the bit-sliced update)
//   3. streaming read bandwidth with the pack/ingest access pattern (32-float chunks)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 12345) *p = 1; }

__global__ void touch_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { uint4 v = in[i]; v.x ^= 1u; out[i] = v; }
}

template <int ITERS>
__global__ void lop3_kernel(uint32_t* out, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = a * 3u, c = a * 5u, d = a * 7u;
    uint32_t e = a ^ 11u, f = b ^ 13u, g = c ^ 17u, h = d ^ 19u;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b) : "r"(c), "r"(d));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c) : "r"(d), "r"(e));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(d) : "r"(e), "r"(f));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(f), "r"(g));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(f) : "r"(g), "r"(h));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(a));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(h) : "r"(a), "r"(b));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}


// ---- 2b. issue rate of the other instructions the step kernels lean on -----------------------
// 8 independent dependency chains per thread, asm volatile so nothing is folded away.
enum { OP_POPC, OP_SHFL, OP_VOTE, OP_REDUX, OP_IMAD, OP_SHF, OP_SEL, OP_FSETP_VOTE, OP_LDS, OP_IADD3 };
template <int OP, int ITERS>
__global__ void op_kernel(uint32_t* out, uint32_t seed) {
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = seed * (k + 3) + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (OP == OP_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(r[k]));
                if (OP == OP_SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(r[k]) : "r"((threadIdx.x + 1) & 31));
                if (OP == OP_VOTE) asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(r[k]));
                if (OP == OP_REDUX) asm volatile("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(r[k]));
                if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[k]) : "r"(r[(k + 1) & 7]), "r"(seed));
                if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(r[k]) : "r"(r[(k + 1) & 7]));
                if (OP == OP_SEL) asm volatile("{ .reg .pred p; setp.eq.u32 p, %1, 5; selp.u32 %0, %0, %2, p; }" : "+r"(r[k]) : "r"(threadIdx.x & 31), "r"(r[(k + 1) & 7]));
                if (OP == OP_FSETP_VOTE) asm volatile("{ .reg .pred p; .reg .f32 f; mov.b32 f, %0; setp.neu.f32 p, f, 0f00000000; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(r[k]));
                if (OP == OP_LDS) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[k]) : "r"((uint32_t)__cvta_generic_to_shared(sm + ((r[k] + threadIdx.x) & 1023))));
                if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[k]) : "r"(r[(k + 1) & 7]));
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= r[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
static void time_op(const char* name, cudaStream_t s, const cudaDeviceProp& prop, uint32_t* out) {
    constexpr int ITERS = 2048;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    op_kernel<OP, ITERS><<<blocks, threads, 0, s>>>(out, 1); cudaStreamSynchronize(s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a, s); op_kernel<OP, ITERS><<<blocks, threads, 0, s>>>(out, r + 2);
        cudaEventRecord(b, s); cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * 32;
    printf("%-22s %.3e thread-ops/s  (%.1f per clk per SM)\n", name, ops / (best * 1e-3),
           ops / (best * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3));
}

// every warp reads 32 consecutive 128-byte chunks (one LDG.32 per lane per chunk)
__global__ void stream_read_kernel(const float* __restrict__ in, uint32_t* out, long long chunks) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long q0 = warp * 32;
    if (q0 >= chunks) return;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = in[(q0 + i) * 32 + lane];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc |= __ballot_sync(0xFFFFFFFFu, v[i] != 0.f);
    if (lane == 0) out[warp] = acc;
}

static float time_graph(cudaStream_t s, cudaGraphExec_t g, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaGraphLaunch(g, s); cudaStreamSynchronize(s);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a, s); cudaGraphLaunch(g, s); cudaEventRecord(b, s);
        cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int L = 200;
    // ---- 1. launch floor ----
    int shapes[][2] = {{1, 32}, {148, 128}, {1024, 128}, {4096, 32}, {512, 256}, {2048, 64}};
    for (auto& sh : shapes) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < L; ++i) empty_kernel<<<sh[0], sh[1], 0, s>>>(nullptr);
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        printf("empty kernel <<<%d,%d>>> in a %d-node graph: %.3f us / launch\n", sh[0], sh[1], L,
               1e3f * time_graph(s, ge, 20) / L);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
    // copy-like kernel over an 8 MiB L2-resident buffer (the cfg-2 state)
    {
        const long long n = 8ll << 20 >> 4;   // uint4 elements
        uint4 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16));
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < L; ++i) {
            touch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(i & 1 ? b : a, i & 1 ? a : b, n);
        }
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        printf("8 MiB read + 8 MiB write (L2-resident ping-pong), uint4 per thread: %.3f us / launch\n",
               1e3f * time_graph(s, ge, 20) / L);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        cudaFree(a); cudaFree(b);
    }
    // ---- 2. LOP3 peak ----
    {
        uint32_t* out; const int blocks = prop.multiProcessorCount * 8, threads = 256;
        CK(cudaMalloc(&out, blocks * threads * 4));
        constexpr int ITERS = 4096;
        lop3_kernel<ITERS><<<blocks, threads, 0, s>>>(out, 1); CK(cudaStreamSynchronize(s));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a, s); lop3_kernel<ITERS><<<blocks, threads, 0, s>>>(out, r);
            cudaEventRecord(b, s); cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        double ops = (double)blocks * threads * ITERS * 16 * 8;
        printf("LOP3 peak: %.3e thread-LOP3/s  (%.1f per clk per SM at %d kHz nominal)\n",
               ops / (best * 1e-3), ops / (best * 1e-3) / prop.multiProcessorCount / (prop.clockRate * 1e3),
               prop.clockRate);
        cudaFree(out);
    }
    {
        uint32_t* out; CK(cudaMalloc(&out, prop.multiProcessorCount * 8 * 256 * 4));
        time_op<OP_POPC>("POPC", s, prop, out);
        time_op<OP_SHFL>("SHFL.IDX", s, prop, out);
        time_op<OP_VOTE>("ISETP+VOTE.ballot", s, prop, out);
        time_op<OP_FSETP_VOTE>("FSETP+VOTE.ballot", s, prop, out);
        time_op<OP_REDUX>("REDUX.add", s, prop, out);
        time_op<OP_IMAD>("IMAD", s, prop, out);
        time_op<OP_SHF>("SHF (funnel shift)", s, prop, out);
        time_op<OP_SEL>("ISETP+SEL", s, prop, out);
        time_op<OP_IADD3>("IADD", s, prop, out);
        time_op<OP_LDS>("LDS.32 (random banks)", s, prop, out);
        cudaFree(out);
    }
    // ---- 3. streaming read with the ingest pattern ----
    {
        const int POOL = 24; const long long chunks = 131072;      // 16 MiB per batch
        std::vector<float*> pool(POOL);
        for (auto& p : pool) { CK(cudaMalloc(&p, chunks * 128)); CK(cudaMemset(p, 0, chunks * 128)); }
        uint32_t* out; CK(cudaMalloc(&out, 4096 * 4));
        for (int bs : {128, 256}) {
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
            for (int i = 0; i < L; ++i)
                stream_read_kernel<<<(unsigned)(chunks / 32 / (bs / 32)), bs, 0, s>>>(pool[i % POOL], out, chunks);
            CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
            float us = 1e3f * time_graph(s, ge, 20) / L;
            printf("stream-read 16 MiB (32 LDG.32 per lane, block %d, rotating 384 MiB pool): %.3f us / launch = %.0f GB/s\n",
                   bs, us, 16.777216e6 / us / 1e3);
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
    }
    return 0;
}
