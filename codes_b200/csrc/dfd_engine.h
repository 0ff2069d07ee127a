// Host-side handle of the device engine (internal; the public C-ABI is include/dfdally_b200.h).
#pragma once
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "dfd_types.h"

struct DfdHostTopo {  // flattened connection tables (see DfdArrays)
    std::vector<int> loc_off, loc_port, loc_dest, lout_lane, lin_src, gs_port, gs_group, gdest, gout_lane, gin_src, gin_port, routed_off, routed_lid;
    // R_SNAPSHOT schedule (PARAMS router_buffer_snapshots): the times before ts_end sorted by (time, configured index)
    std::vector<double> snap_ts;
    std::vector<unsigned> snap_id;
};

// page-locked host array (device -> host copies at full PCIe rate, no staging); grows, never shrinks
template <typename T> struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void release() {
        if (p) {
            if (pinned)
                cudaFreeHost(p);
            else
                free(p);
        }
        p = nullptr;
        n = cap = 0;
    }
    void resize(size_t count) {
        if (count > cap) {
            release();
            void* q = nullptr;
            if (cudaHostAlloc(&q, std::max<size_t>(1, count) * sizeof(T), cudaHostAllocDefault) == cudaSuccess) {
                pinned = true;
            } else {
                (void)cudaGetLastError();
                q = malloc(std::max<size_t>(1, count) * sizeof(T));
                pinned = false;
            }
            p = (T*)q;
            cap = count;
            memset((void*)p, 0, std::max<size_t>(1, count) * sizeof(T));
        }
        n = count;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

struct DfdHostResults {  // indexed by GLOBAL node / router id
    PinnedBuf<NodeState> node;
    PinnedBuf<RouterHdr> rh;
    PinnedBuf<PortState> port;
    PinnedBuf<int4> snap;  // [NR][nsnap_all][radix] vc_occupancy captured by the R_SNAPSHOT events
    PinnedBuf<unsigned> ack_count;
    PinnedBuf<unsigned long long> ack_first, ack_last;
    unsigned long long counters[64];
    size_t d2h_bytes = 0;
};

struct DfdDevice {
    DfdParams P;
    DfdArrays A;
    DfdSeedConsts seeds;
    DfdPeers X;
    char* xregion = nullptr;
    size_t xregion_bytes = 0;
    size_t x_ack_off = 0;  // start of the ACK arrays inside the exchange region
    size_t x_cnt_off = 0;  // ... of the counters: the statistics of a partition live in the exchange region too, so that
                           // rank 0 can read every partition over its CUDA-IPC mapping when the run is over
    int ntl = 0, nrl = 0;        // nodes / routers owned by this rank
    int ntmax = 0, nrmax = 0;    // largest partition of any rank (exchange-region strides)
    std::vector<void*> ipc_opened;
    std::vector<void*> allocs;
    size_t bytes_allocated = 0, topo_bytes = 0;
    std::vector<size_t> topo_sizes;  // element counts of the uploaded tables (reuse check)
    int num_sms = 0, grid = 0, block = 0, blocks_per_sm = 0;
    int warps = 0;  // resident warps of the event-loop kernel (grid x warps per block)
    size_t smem = 0;
    const void* kernel = nullptr;  // dfd_run_kernel instantiation in use
    char error[512] = {0};
};

void dfd_partition(int G, int a, int p, int rank, int world, DfdPeers* X);
int dfd_device_create(DfdDevice* dev, const DfdParams& P, const DfdHostTopo& T, int rank, int world, int ensemble = 1);
int dfd_device_ipc_export(DfdDevice* dev, void* handle64);
int dfd_device_ipc_import(DfdDevice* dev, const void* handles, int world);
int dfd_device_upload_topo(DfdDevice* dev, const DfdHostTopo& T, cudaStream_t st);
int dfd_device_reset(DfdDevice* dev, cudaStream_t st);
int dfd_device_run(DfdDevice* dev, cudaStream_t st);
// what dfd_device_download fetches: this rank's slice; every rank's slice through the peer mappings (rank 0 of a sharded
// run, once every rank's kernel has finished); or only this rank's counters and control block
enum { DFD_DL_OWN = 0, DFD_DL_ALL = 1, DFD_DL_CTRL = 2 };
int dfd_device_download(DfdDevice* dev, DfdHostResults* out, cudaStream_t st, int mode = DFD_DL_OWN);
void dfd_device_destroy(DfdDevice* dev);
