// Does a store allocate / keep a line in L1 on sm_100a?  Pointer chases (every address depends on the previous
// value) over 16 lines, single thread; cycles per load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1store l1store.cu && ./l1store
#include <cstdio>
#include <cuda_runtime.h>
#define N 16
__device__ __forceinline__ unsigned chase(unsigned* base, unsigned idx, long long* cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) idx = base[idx * 32];
    long long t1 = clock64();
    *cyc = (t1 - t0) / N + (idx & 0);
    return idx;
}
__global__ void k(unsigned* buf, long long* out) {
    unsigned acc = 0;
    // A: lines never loaded by this SM; store the ring, then chase it twice
    unsigned* A = buf + 32 * 1024;
    for (int i = 0; i < N; i++) A[i * 32] = (i + 1) % N;
    __threadfence_block();
    acc = chase(A, acc & 15, &out[0]);
    acc = chase(A, acc & 15, &out[1]);
    // B: ring already in L1 (just chased): overwrite every line with the same ring, chase again
    for (int i = 0; i < N; i++) A[i * 32] = (i + 1) % N + (acc & 0);
    __threadfence_block();
    acc = chase(A, acc & 15, &out[2]);
    // C: ring set up by the host (memset + kernel arg), first chase = L2/DRAM, second = L1
    unsigned* C = buf + 32 * 4096;
    acc = chase(C, acc & 15, &out[3]);
    acc = chase(C, acc & 15, &out[4]);
    // D: same as B but the store goes to a different 32-byte sector of each line
    for (int i = 0; i < N; i++) C[i * 32 + 8] = acc;
    __threadfence_block();
    acc = chase(C, acc & 15, &out[5]);
    // E: L1-resident ring, wait 50k cycles (other traffic: none), chase
    long long s = clock64();
    while (clock64() - s < 50000) {}
    acc = chase(C, acc & 15, &out[6]);
    buf[0] = acc;
}
int main() {
    unsigned* buf; long long* out;
    cudaMalloc(&buf, 4 << 20); cudaMemset(buf, 0, 4 << 20);
    unsigned ring[N * 32] = {0};
    for (int i = 0; i < N; i++) ring[i * 32] = (i + 1) % N;
    cudaMemcpy(buf + 32 * 4096, ring, sizeof ring, cudaMemcpyHostToDevice);
    cudaMallocManaged(&out, 16 * sizeof(long long));
    k<<<1, 1>>>(buf, out); cudaDeviceSynchronize();
    const char* n[] = {"store to never-loaded lines, then load", "  second chase", "L1-resident lines, store to them, load", "lines never touched by the SM: first chase", "  second chase",
                       "L1-resident, store to another sector, load", "L1-resident, idle 50k cycles, load"};
    for (int i = 0; i < 7; i++) printf("%-48s %lld cycles/load\n", n[i], out[i]);
    return 0;
}
