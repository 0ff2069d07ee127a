// which "fresh from L2" load flavours leave other L1-resident lines alone?
#include <cstdio>
#include <cuda_runtime.h>
#define TL(slot, expr) t0 = clock64(); acc += (expr); asm volatile("" :: "r"(acc) : "memory"); t1 = clock64(); out[slot] = t1 - t0;
__device__ __forceinline__ unsigned ld_noalloc(const unsigned* p) { unsigned v; asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_weak_cg(const unsigned* p) { unsigned v; asm volatile("ld.weak.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed_cta(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.cta.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_nc(const unsigned* p) { unsigned v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_cv(const unsigned* p) { unsigned v; asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__global__ void k(unsigned* buf, long long* out) {
    unsigned* A = buf + 4096; unsigned* B = buf + 8192; unsigned* C = buf + 65536;
    unsigned acc = 0; long long t0, t1;
    TL(0, A[0]) TL(1, B[0])                          // warm A, B
    TL(2, ld_noalloc(C)) TL(3, A[1]) TL(4, B[1])     // no_allocate load of C, then A/B
    TL(5, ld_weak_cg(C + 64)) TL(6, A[2]) TL(7, B[2])
    TL(8, ld_relaxed_cta(C + 128)) TL(9, A[3]) TL(10, B[3])
    TL(11, ld_nc(C + 192)) TL(12, A[4]) TL(13, B[4])
    TL(14, atomicAdd(C + 256, 0u)) TL(15, A[5]) TL(16, B[5])
    TL(17, atomicOr(C + 320, 0u)) TL(18, A[6]) TL(19, B[6])
    TL(20, ld_cv(C + 384)) TL(21, A[7]) TL(22, B[7])
    TL(23, A[8]) TL(24, B[8])
    TL(25, ld_relaxed_gpu(C + 448)) TL(26, A[9]) TL(27, B[9])
    TL(28, __ldcg(C + 512)) TL(29, A[10]) TL(30, B[10])
    buf[0] = acc;
}
int main() {
    unsigned* buf; long long* out;
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20);
    cudaMallocManaged(&out, 64 * 8);
    for (int rep = 0; rep < 2; rep++) {
        k<<<1, 1>>>(buf, out); cudaDeviceSynchronize();
        const char* n[] = {"L1::no_allocate", "weak.cg", "relaxed.cta", "nc", "atomicAdd0", "atomicOr0", "cv(volatile)", "", "relaxed.gpu", "__ldcg"};
        printf("warm A %lld B %lld\n", out[0], out[1]);
        for (int i = 0; i < 10; i++) { if (i == 7) { printf("  (plain again) A %lld B %lld\n", out[23], out[24]); continue; } int b = i < 7 ? 2 + 3 * i : (i == 8 ? 25 : 28); printf("  %-16s itself %5lld | then A %5lld  B %5lld\n", n[i], out[b], out[b + 1], out[b + 2]); }
    }
    return 0;
}
