// What does fence.release.sys cost on B200 when stores to a peer GPU (NVLink) are outstanding, and how does the cost change
// when the fence comes some time after the stores?  One warp per block, lane 0 measures.  Needs 2 GPUs with peer access.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fence_sys fence_sys.cu && ./fence_sys
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k(int4* remote, int4* local, int mode, int nstores, int delay_cycles, int iters, long long* out) {
    // mode 4: fence.gpu right after remote stores; mode 0: fence.gpu after local stores; 1: fence.sys, no stores; 2: fence.sys right after remote stores; 3: fence.sys `delay` cycles after them
    const int lane = threadIdx.x & 31;
    int4* r = remote + (size_t)blockIdx.x * 4096;
    int4* l = local + (size_t)blockIdx.x * 4096;
    long long total = 0, worst = 0;
    for (int it = 0; it < iters; it++) {
        if (lane == 0) {
            int4 v = make_int4(it, it, it, it);
            if (mode == 0)
                for (int i = 0; i < nstores; i++) __stcg(&l[(it * nstores + i) & 4095], v);
            if (mode >= 2)
                for (int i = 0; i < nstores; i++) __stcg(&r[(it * nstores + i) & 4095], v);
            if (mode == 3) {
                long long t = clock64();
                while (clock64() - t < delay_cycles) {}
            }
            long long t0 = clock64();
            if (mode == 0 || mode == 4)
                asm volatile("fence.release.gpu;" ::: "memory");
            else
                asm volatile("fence.release.sys;" ::: "memory");
            long long dt = clock64() - t0;
            total += dt;
            if (dt > worst) worst = dt;
            // some unrelated work between iterations so that iterations do not pile up
            long long t = clock64();
            while (clock64() - t < 20000) {}
        }
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&out[0], (unsigned long long)total);
        atomicMax((unsigned long long*)&out[1], (unsigned long long)worst);
    }
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(1));
    int4* remote;
    CK(cudaMalloc(&remote, 2048 * 4096 * sizeof(int4)));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    int4* local;
    long long* out;
    CK(cudaMalloc(&local, 2048 * 4096 * sizeof(int4)));
    CK(cudaMalloc(&out, 16));
    const char* names[] = {"fence.gpu after local stores", "fence.sys, nothing outstanding", "fence.sys right after remote stores", "fence.sys after remote stores + delay", "fence.gpu right after remote stores"};
    int blocks_list[] = {1, 148, 1036};
    for (int blocks : blocks_list)
        for (int mode = 0; mode < 5; mode++)
            for (int delay : {2000, 6000, 12000}) {
                if (mode != 3 && delay != 2000) continue;
                const int iters = 200, nstores = 12;
                CK(cudaMemset(out, 0, 16));
                k<<<blocks, 32>>>(remote, local, mode, nstores, delay, iters, out);
                CK(cudaDeviceSynchronize());
                long long h[2];
                CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
                printf("blocks %4d  %-40s delay %5d cyc: mean %7.0f cycles, worst %lld\n", blocks, names[mode], mode == 3 ? delay : 0, (double)h[0] / ((double)iters * blocks), h[1]);
            }
    return 0;
}
