// does an L1-resident line survive (a) red.release.gpu, (b) ld.relaxed.gpu polls, (c) st.cg to another line ?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long timed_load(const unsigned* p, unsigned* acc) {
    long long t0 = clock64();
    unsigned v = *(volatile const unsigned*)0 == 0 ? 0 : 0; (void)v;
    return t0;
}
__global__ void k(unsigned* buf, unsigned long long* ctr, long long* out) {
    unsigned* p = buf + 4096;
    unsigned acc = 0; long long t0, t1;
#define TL(slot, expr) t0 = clock64(); acc += (expr); asm volatile("" :: "r"(acc) : "memory"); t1 = clock64(); if (acc == 0x7fffffff) out[63] = 1; out[slot] = t1 - t0;
    TL(0, p[0])              // cold
    TL(1, p[1])              // L1 hit
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
    TL(2, p[2])              // after release-red
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); acc += (unsigned)v;
    TL(3, p[3])              // after relaxed poll
    __stcg(buf + 8192, acc);
    TL(4, p[4])              // after st.cg elsewhere
    p[5] = acc;              // plain store into the line
    TL(5, p[6])              // after own store
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    TL(6, p[7])              // after acq_rel fence (expected miss)
    TL(7, p[8])
    TL(8, __ldcg(p + 64))    // L2 latency, fresh line (DRAM)
    TL(9, __ldcg(p + 65))    // L2 hit
    t0 = clock64(); unsigned old = atomicAdd(buf + 16384, 1u); acc += old; asm volatile("" :: "r"(acc) : "memory"); t1 = clock64(); out[10] = t1 - t0;
    t0 = clock64(); old = atomicAdd(buf + 16384, 1u); acc += old; asm volatile("" :: "r"(acc) : "memory"); t1 = clock64(); out[11] = t1 - t0;
    buf[0] = acc;
}
int main() {
    unsigned* buf; unsigned long long* ctr; long long* out;
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20); cudaMalloc(&ctr, 8); cudaMemset(ctr, 0, 8);
    cudaMallocManaged(&out, 64 * 8);
    for (int rep = 0; rep < 2; rep++) {
        k<<<1, 1>>>(buf, ctr, out); cudaDeviceSynchronize();
        printf("cold %lld | L1 hit %lld | after red.release %lld | after relaxed poll %lld | after st.cg %lld | after own store %lld | after acq_rel fence %lld, again %lld | ldcg fresh %lld | ldcg L2 hit %lld | atomic first %lld second %lld\n",
               out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[9], out[10], out[11]);
    }
    return 0;
}
