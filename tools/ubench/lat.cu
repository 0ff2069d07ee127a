// micro-benchmarks that size the window engine's critical path on B200 (latencies in SM cycles)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_l1(unsigned* buf, long long* out) {
    // buf: private array; measure cold load, repeat load, load after own store (same line), ld.cg
    unsigned* p = buf + 1024 * (blockIdx.x + 1);
    long long t0, t1; unsigned v, acc = 0;
    t0 = clock64(); v = p[0]; acc += v; asm volatile("" :: "r"(acc)); __syncwarp(); t1 = clock64(); out[0] = t1 - t0;       // cold
    t0 = clock64(); v = p[(acc & 1) + 1]; acc += v; asm volatile("" :: "r"(acc)); t1 = clock64(); out[1] = t1 - t0;          // same line again
    p[2] = acc + 7;                                                                                                     // store to the line
    t0 = clock64(); v = p[(acc & 1) + 3]; acc += v; asm volatile("" :: "r"(acc)); t1 = clock64(); out[2] = t1 - t0;          // load after store (same line)
    t0 = clock64(); v = p[2]; acc += v; asm volatile("" :: "r"(acc)); t1 = clock64(); out[3] = t1 - t0;                      // load the stored word
    t0 = clock64(); v = __ldcg(p + 4 + (acc & 1)); acc += v; asm volatile("" :: "r"(acc)); t1 = clock64(); out[4] = t1 - t0; // ld.cg (L2)
    t0 = clock64(); v = atomicAdd(p + 64, 1u); acc += v; asm volatile("" :: "r"(acc)); t1 = clock64(); out[5] = t1 - t0;     // atomic round trip
    // dependent chain of 16 loads on one private line after stores
    t0 = clock64();
    unsigned idx = acc & 3;
    for (int i = 0; i < 16; i++) { idx = p[8 + (idx & 15)] & 15; p[32 + i] = idx; }
    t1 = clock64(); out[6] = (t1 - t0) / 16;
    // fp64 division latency
    double d = 1.0 + acc; t0 = clock64();
    for (int i = 0; i < 16; i++) d = 4096.0 / (d + 1.5);
    t1 = clock64(); out[7] = (t1 - t0) / 16;
    // integer division by a runtime divisor
    unsigned q = acc + 12345, dv = (acc & 7) + 3; t0 = clock64();
    for (int i = 0; i < 16; i++) q = q / dv + 77777;
    t1 = clock64(); out[8] = (t1 - t0) / 16;
    buf[0] = acc + (unsigned)d + q + idx;
}
__global__ void k_bar(unsigned long long* ctr, long long* out, int iters) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < (unsigned long long)i * gridDim.x);
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
int main() {
    unsigned* buf; long long* out; unsigned long long* ctr;
    cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 0, 64 << 20);
    cudaMallocManaged(&out, 64 * sizeof(long long)); cudaMalloc(&ctr, 8);
    k_l1<<<1, 32>>>(buf, out); cudaDeviceSynchronize();
    printf("cold %lld  again %lld  after-store-same-line %lld  stored-word %lld  ldcg %lld  atomic %lld  dep-chain/ld %lld  ddiv %lld  udiv %lld\n",
           out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8]);
    int grids[] = {1, 8, 16, 32, 74, 148, 296, 592};
    for (int g : grids) {
        cudaMemset(ctr, 0, 8);
        void* args[] = {&ctr, &out, nullptr}; int iters = 2000; args[2] = &iters;
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_bar, dim3(g), dim3(128), args, 0, 0);
        cudaDeviceSynchronize();
        printf("barrier grid=%d: %lld cycles (%s)\n", g, out[0], cudaGetErrorString(e));
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
    return 0;
}
