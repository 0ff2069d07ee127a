// Host-side handle of the device engine (internal; the public C-ABI is include/dfdally_b200.h).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "dfd_types.h"

struct DfdSeedConsts {  // CLCG4 constants: moduli, default seed, a^(2^(31+41)) mod m (stream jump)
    unsigned long long m[4], seed0[4], avw[4];
};

struct DfdHostTopo {  // flattened connection tables (see DfdArrays)
    std::vector<int> loc_off, loc_port, loc_dest, gs_port, gs_group, gdest, routed_off, routed_lid;
};

struct DfdHostResults {
    std::vector<NodeState> node;
    std::vector<RouterHdr> rh;
    std::vector<PortState> port;
    std::vector<unsigned> ack_count;
    std::vector<unsigned long long> ack_first, ack_last;
    unsigned long long counters[64];
    std::vector<unsigned long long> winmax;
    size_t d2h_bytes = 0;
};

struct DfdDevice {
    DfdParams P;
    DfdArrays A;
    DfdSeedConsts seeds;
    DfdPeers X;
    char* xregion = nullptr;
    size_t xregion_bytes = 0;
    size_t x_zero_off = 0, x_zero_bytes = 0;  // lane counts + masks inside the exchange region (sharded runs)
    std::vector<void*> ipc_opened;
    std::vector<void*> allocs;
    size_t bytes_allocated = 0, topo_bytes = 0;
    int num_sms = 0, grid = 0, block = 0, blocks_per_sm = 0;
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
int dfd_device_run(DfdDevice* dev, cudaStream_t st, unsigned long long max_windows);
int dfd_device_download(DfdDevice* dev, DfdHostResults* out, cudaStream_t st);
int dfd_device_read_counters(DfdDevice* dev, unsigned long long* counters, cudaStream_t st);
void dfd_device_destroy(DfdDevice* dev);
