// dfd_engine.cu -- B200 (sm_100a) asynchronous engine for the CODES dragonfly-dally hot path.
//
// One persistent kernel executes the whole simulation with NO grid barrier.  Every warp owns a fixed set of groups
// (one router + its terminals each, dfd_core.h) and activates them round-robin; an activation polls the router's
// channel words, executes what is provably safe, and publishes new bounds to the neighbours.  LP-private state stays
// L1-resident (a group is always processed by the same SM); records and channel words are exchanged through L2 with
// ld.relaxed / st.relaxed behind one release fence per activation.  Termination is detected by warp 0 of the grid
// (created == consumed, DfdCtrl).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (timestamps must be plain IEEE double
// operations in the reference's order, so FMA contraction is disabled for the whole file).

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dfd_engine.h"
#include "dfd_init.h"

#define DFD_POOL_WARPS 16  // warps of one dfd_run_kernel_pool block: 16 x 32 threads x 128 registers = one SM's register file

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Monitor (warp 0 of the grid, while it is idle): termination and watchdog.  Reads every rank's `consumed`, then --
// after a fence -- every rank's `created`; equal sums prove that nothing was pending at a moment in between.
__device__ bool monitor_check(const DfdArrays& A, const DfdPeers& X, unsigned long long* last_consumed, unsigned long long* last_change_ns,
                              unsigned long long timeout_ns, Ctx& c) {
    const int lane = DFD_LANE();
    unsigned long long cons = 0, crea = 0;
    if (lane < X.world) {
        const DfdCtrl* pc = lane == X.rank ? A.ctrl : (const DfdCtrl*)(X.base[lane] + ((const char*)A.ctrl - X.local_base));
        cons = ld_strong64_sys(&pc->consumed);
    }
    __threadfence_system();
    if (lane < X.world) {
        const DfdCtrl* pc = lane == X.rank ? A.ctrl : (const DfdCtrl*)(X.base[lane] + ((const char*)A.ctrl - X.local_base));
        crea = ld_strong64_sys(&pc->created);
    }
    for (int o = 16; o > 0; o >>= 1) {
        cons += __shfl_xor_sync(0xFFFFFFFFu, cons, o);
        crea += __shfl_xor_sync(0xFFFFFFFFu, crea, o);
    }
    if (cons == crea) {
        if (lane == 0) st_strong64(&A.ctrl->done, 1ULL, false);
        return true;
    }
    unsigned long long now = globaltimer_ns();
    if (cons != *last_consumed) {
        *last_consumed = cons;
        *last_change_ns = now;
    } else if (now - *last_change_ns > timeout_ns) {
        if (lane == 0) set_error(c, DFD_ERR_STALL, crea - cons);
    }
    return false;
}

// Publisher of queue qi (sharded runs; PubCtl in dfd_types.h): takes the entries that are ready, executes ONE system-scope
// fence for the batch and stores the channel words into the other ranks' memory.  Within a batch the last entry for a
// target wins (a newer word supersedes an older one: counts and bounds only grow).
__device__ void publisher_loop(const DfdParams& P, const DfdArrays& A, unsigned qi) {
    // chunks of 32 entries per batch.  1 is what was validated on 2, 4 and 8 GPUs; 4 (up to 128 words per fence) passes the
    // ranks-on-one-GPU tests and was not faster there
    constexpr int CH = 1;
    const int lane = DFD_LANE();
    const ulonglong2* q = A.pubq + (size_t)qi * DFD_PUBQ_CAP;
    PubCtl* ctl = &A.pubctl[qi];
    unsigned long long head = 0;
    for (;;) {
        unsigned long long a[CH], b[CH];
        int n = 0;  // entries ready in ticket order
        bool open = true;
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const unsigned long long t = head + (unsigned)(32 * j + lane);
            ld_relaxed_b128(&q[t & (DFD_PUBQ_CAP - 1)], a[j], b[j]);
            const bool ok = (a[j] & 7ULL) == (t / DFD_PUBQ_CAP) % 3ULL + 1ULL;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (open) {
                if (m == 0xFFFFFFFFu)
                    n += 32;
                else {
                    n += __ffs((int)~m) - 1;
                    open = false;
                }
            }
        }
        if (n == 0) {
            if (ld_strong64(&A.ctrl->error) | ld_strong64(&A.ctrl->done)) break;
            __nanosleep(100);
            continue;
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
#pragma unroll
        for (int j = 0; j < CH; j++) {
            // inside a chunk the last entry for a target wins (one instruction: no order among its lanes); a later chunk
            // overwrites an earlier one: __syncwarp orders the stores of two chunks, and both are system-scope stores to one location
            const bool mine = 32 * j + lane < n;
            const unsigned long long tgt = a[j] & ~7ULL;
            const unsigned same = __match_any_sync(0xFFFFFFFFu, mine ? tgt : (unsigned long long)lane);  // real targets are pointers: never < 32
            if (mine && (31 - __clz((int)same)) == lane) st_strong64((unsigned long long*)tgt, b[j], true);
            __syncwarp();
        }
        head += (unsigned)n;
        if (lane == 0) st_strong64(&ctl->head, head, false);
    }
}

// What a warp accumulates over its life and folds into the run counters when it leaves the event loop.
struct WarpTally {
    unsigned long long n_act = 0, n_prod = 0, n_exec = 0, passes = 0;
    long long t_busy = 0, t_idle = 0;
};
__device__ __forceinline__ void tally_commit(const DfdArrays& A, const Ctx& c, const WarpTally& t) {
    if (DFD_LANE() == 0) {
        atomicAdd(&A.counters[1], t.n_prod);
        atomicAdd(&A.counters[20], (unsigned long long)t.t_busy);
        atomicAdd(&A.counters[21], (unsigned long long)t.t_idle);
        atomicAdd(&A.counters[22], t.n_act);
        atomicAdd(&A.counters[23], t.passes);
        atomicMax(&A.counters[24], (unsigned long long)t.t_busy);
        atomicAdd(&A.counters[25], t.n_exec);
        for (int i = 0; i < 5; i++) atomicAdd(&A.counters[30 + i], (unsigned long long)c.cyc[i]);
        atomicAdd(&A.counters[35], c.nrounds);
    }
}
__device__ __forceinline__ void ctx_init(Ctx& c, const DfdParams* P, const DfdArrays* A, const DfdPeers* X, unsigned* ev) {
    c.P = P;
    c.A = A;
    c.X = X;
    c.ev = ev;
    c.port = nullptr;
    c.rl = 0xFFFFFFFFu;
    for (int i = 0; i < 6; i++) c.cyc[i] = 0;
    c.nrounds = 0;
}

// Latency-bound partitions (at most one group per resident warp): one warp per block, every warp owns its groups for
// the whole run.  MIN_BLOCKS trades registers for resident warps: <8> leaves 255 registers per thread for partitions of
// up to 8 groups per SM, <16> caps at 128 registers for larger ones.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(32, MIN_BLOCKS) dfd_run_kernel(DfdParams P, DfdArrays A, DfdPeers X, unsigned long long timeout_ns) {
    __shared__ unsigned s_ev[EV_NTYPES];
    const int lane = DFD_LANE();
    if (lane < EV_NTYPES) s_ev[lane] = 0;
    __syncwarp();
    Ctx c;
    ctx_init(c, &P, &A, &X, s_ev);
    const unsigned ngroups = (unsigned)(X.r1 - X.r0);
    const unsigned w = blockIdx.x, nw = gridDim.x - (unsigned)P.npub_blocks;
    if (w >= nw) {  // the last blocks of the grid are the publishers of a sharded run
        publisher_loop(P, A, w - nw);
        return;
    }
    WarpTally t;
    unsigned long long idle_passes = 0;
    unsigned long long last_consumed = ~0ULL, last_change_ns = globaltimer_ns();
    for (;;) {
        long long c0 = clock64();
        // run control: requested before the pass so that its latency overlaps the channel-word poll
        const unsigned long long err = ld_strong64(&A.ctrl->error), done = ld_strong64(&A.ctrl->done);
        unsigned nexec = 0;
        for (unsigned g = w; g < ngroups; g += nw) {
            ActResult a = group_activate<false>(c, g);
            t.n_act += a.polled_change ? 1 : 0;
            t.n_prod += a.nexec ? 1 : 0;
            nexec += a.nexec;
        }
        long long c1 = clock64();
        if (nexec) {
            t.t_busy += c1 - c0;
            idle_passes = 0;
        } else {
            t.t_idle += c1 - c0;
            idle_passes++;
        }
        t.n_exec += nexec;
        t.passes++;
        if (err | done) {
            if (err && w == 0 && lane == 0 && X.world > 1) propagate_error(A, X);
            break;
        }
        if (w == 0 && nexec == 0 && (idle_passes & 7) == 1) {
            if (monitor_check(A, X, &last_consumed, &last_change_ns, timeout_ns, c)) break;
        }
        if (nexec == 0 && P.idle_sleep_ns) __nanosleep((unsigned)P.idle_sleep_ns);
    }
    __syncwarp();
    if (lane < EV_NTYPES && s_ev[lane]) {
        atomicAdd(&A.counters[4 + lane], (unsigned long long)s_ev[lane]);
        atomicAdd(&A.counters[0], (unsigned long long)s_ev[lane]);
    }
    tally_commit(A, c, t);
}

// Throughput-bound partitions (several groups per resident warp): one block of WARPS warps per SM owns a contiguous
// slice of the groups, and its warps take turns on them -- a ticket counter deals the groups out round-robin, a lock word
// per group keeps two warps off the same group.  A group whose activation runs long (a hot router, a burst of
// arrivals) then holds up one warp, not the other groups a static assignment would have queued behind it; the group's
// state still never leaves the SM, so its L1 lines stay where they are.  Lock and unlock are block-scope
// acquire/release: the warps of a block share one L1, plain loads and stores of LP-private state are coherent there.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) dfd_run_kernel_pool(DfdParams P, DfdArrays A, DfdPeers X, unsigned long long timeout_ns) {
    extern __shared__ unsigned s_lock[];  // [groups of this block]
    __shared__ unsigned s_ev[EV_NTYPES];
    __shared__ unsigned s_ticket;
    const int lane = DFD_LANE();
    const unsigned wib = threadIdx.x >> 5;
    const unsigned ngroups = (unsigned)(X.r1 - X.r0);
    const unsigned b = blockIdx.x, nb = gridDim.x - (unsigned)P.npub_blocks;
    if (b >= nb) {  // publisher block of a sharded run: one queue per warp
        const unsigned qi = (b - nb) * WARPS + wib;
        if (qi < (unsigned)P.npubq) publisher_loop(P, A, qi);
        return;
    }
    const unsigned g0 = (unsigned)(((unsigned long long)ngroups * b) / nb), g1 = (unsigned)(((unsigned long long)ngroups * (b + 1)) / nb);
    const unsigned ng = g1 - g0;
    for (unsigned i = threadIdx.x; i < ng; i += blockDim.x) s_lock[i] = 0;
    if (threadIdx.x < EV_NTYPES) s_ev[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_ticket = 0;
    __syncthreads();
    Ctx c;
    ctx_init(c, &P, &A, &X, s_ev);
    WarpTally t;
    const bool monitor = b == 0 && wib == 0;
    const unsigned per_pass = ng ? (ng + WARPS - 1) / WARPS : 0;
    unsigned long long idle_passes = 0;
    unsigned long long last_consumed = ~0ULL, last_change_ns = globaltimer_ns();
    for (;;) {
        long long c0 = clock64();
        const unsigned long long err = ld_strong64(&A.ctrl->error), done = ld_strong64(&A.ctrl->done);
        unsigned nexec = 0;
        for (unsigned k = 0; k < per_pass; k++) {
            unsigned i = 0;
            if (lane == 0) {
                i = atomicInc(&s_ticket, ng - 1);  // round-robin over the block's groups, wraps at ng
                if (atomicExch_block(&s_lock[i], 1u) != 0u) i = 0xFFFFFFFFu;  // another warp is on it
            }
            i = __shfl_sync(0xFFFFFFFFu, i, 0);
            if (i == 0xFFFFFFFFu) continue;
            __threadfence_block();
            ActResult a = group_activate<true>(c, g0 + i);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicExch_block(&s_lock[i], 0u);
            t.n_act += a.polled_change ? 1 : 0;
            t.n_prod += a.nexec ? 1 : 0;
            nexec += a.nexec;
        }
        long long c1 = clock64();
        if (nexec) {
            t.t_busy += c1 - c0;
            idle_passes = 0;
        } else {
            t.t_idle += c1 - c0;
            idle_passes++;
        }
        t.n_exec += nexec;
        t.passes++;
        if (err | done) {
            if (err && monitor && lane == 0 && X.world > 1) propagate_error(A, X);
            break;
        }
        if (monitor && nexec == 0 && (idle_passes & 7) == 1) {
            if (monitor_check(A, X, &last_consumed, &last_change_ns, timeout_ns, c)) break;
        }
        if (nexec == 0 && P.idle_sleep_ns) __nanosleep((unsigned)P.idle_sleep_ns);
    }
    __syncthreads();
    if (threadIdx.x < EV_NTYPES && s_ev[threadIdx.x]) {
        atomicAdd(&A.counters[4 + threadIdx.x], (unsigned long long)s_ev[threadIdx.x]);
        atomicAdd(&A.counters[0], (unsigned long long)s_ev[threadIdx.x]);
    }
    tally_commit(A, c, t);
}

// ================================================================================================
// state initialisation (device side)
// ================================================================================================
__global__ void dfd_init_kernel(DfdParams P, DfdArrays A, DfdPeers X, DfdSeedConsts K) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ntl = (size_t)(X.n1 - X.n0), nrl = (size_t)(X.r1 - X.r0);
    if (gtid == 0) init_ctrl(P, A, X);
    for (size_t i = gtid; i < (size_t)P.npubq * DFD_PUBQ_CAP; i += nthreads) A.pubq[i] = make_ulonglong2(0ULL, 0ULL);
    for (size_t i = gtid; i < (size_t)P.npubq; i += nthreads) A.pubctl[i].reserve = A.pubctl[i].head = 0ULL;
    for (size_t n = gtid; n < ntl; n += nthreads) init_node(P, A, X, K, (unsigned)n);
    for (size_t r = gtid; r < nrl; r += nthreads) init_router(P, A, X, K, (unsigned)r);
    for (size_t i = gtid; i < nrl * P.radix; i += nthreads) init_lane(P, A, X, i);
    const size_t nqw = nrl * (size_t)P.radix * P.radix / 4;  // radix^2 bytes per router, cleared as words
    for (size_t i = gtid; i < nqw + 1; i += nthreads)
        if (i < nqw)
            ((unsigned*)A.nq)[i] = 0u;
        else
            for (size_t b = nqw * 4; b < nrl * (size_t)P.radix * P.radix; b++) A.nq[b] = 0;
    const size_t nslots = nrl * (size_t)P.rpool_cap;
    for (size_t i = gtid; i < nslots; i += nthreads) {
        unsigned k = (unsigned)(i % P.rpool_cap);
        A.pfree[i] = (unsigned)P.rpool_cap - 1 - k;  // popped from the top: slot 0 first
    }
}

// ================================================================================================
// host side of the device engine
// ================================================================================================
#define CK(x)                                                                                          \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) {                                                                       \
            snprintf(dev->error, sizeof dev->error, "%s failed: %s", #x, cudaGetErrorString(e_));      \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)

template <typename T> static int dalloc(DfdDevice* dev, T** p, size_t n) {
    size_t bytes = n * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    CK(cudaMalloc((void**)p, bytes));
    dev->allocs.push_back((void*)*p);
    dev->bytes_allocated += bytes;
    return 0;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

void dfd_partition(int G, int a, int p, int rank, int world, DfdPeers* X) {
    memset(X, 0, sizeof *X);
    X->rank = rank;
    X->world = world;
    for (int k = 0; k <= world; k++) X->grp_lo[k] = (int)(((long long)k * G) / world);  // contiguous groups
    for (int k = world + 1; k <= DFD_MAX_RANKS; k++) X->grp_lo[k] = G;
    X->r0 = X->grp_lo[rank] * a;
    X->r1 = X->grp_lo[rank + 1] * a;
    X->n0 = X->r0 * p;
    X->n1 = X->r1 * p;
}

// ring geometry by lane index (DfdArrays::lgeo): local links, global links, terminal lanes
static void lane_geometry(const DfdParams& P, std::vector<uint2>* g) {
    g->resize((size_t)P.radix * 2);
    for (int j = 0; j < P.radix; j++) {
        unsigned coff, ccap, koff, kcap;
        if (j < P.intra_radix) {
            coff = koff = (unsigned)j * P.cap_l;
            ccap = kcap = P.cap_l;
        } else if (j < P.ext) {
            coff = koff = (unsigned)P.intra_radix * P.cap_l + (unsigned)(j - P.intra_radix) * P.cap_g;
            ccap = kcap = P.cap_g;
        } else {
            coff = (unsigned)P.intra_radix * P.cap_l + (unsigned)P.global_radix * P.cap_g + (unsigned)(j - P.ext) * P.cap_tc;
            koff = (unsigned)P.intra_radix * P.cap_l + (unsigned)P.global_radix * P.cap_g + (unsigned)(j - P.ext) * P.cap_tk;
            ccap = P.cap_tc;
            kcap = P.cap_tk;
        }
        (*g)[2 * j] = make_uint2(coff, ccap - 1);
        (*g)[2 * j + 1] = make_uint2(koff, kcap - 1);
    }
}

static std::vector<const std::vector<int>*> topo_tables(const DfdHostTopo& T) {
    return {&T.loc_off, &T.loc_port, &T.loc_dest, &T.lout_lane, &T.lin_src, &T.gs_port, &T.gs_group, &T.gdest, &T.gout_lane, &T.gin_src, &T.gin_port, &T.routed_off, &T.routed_lid};
}

int dfd_device_create(DfdDevice* dev, const DfdParams& P, const DfdHostTopo& T, int rank, int world, int ensemble) {
    dev->P = P;
    DfdArrays& A = dev->A;
    memset(&A, 0, sizeof A);
    if (world < 1 || world > DFD_MAX_RANKS || rank < 0 || rank >= world) {
        snprintf(dev->error, sizeof dev->error, "bad partition: rank %d of %d (at most %d ranks)", rank, world, DFD_MAX_RANKS);
        return 1;
    }
    if (world > P.G) {
        snprintf(dev->error, sizeof dev->error, "more ranks (%d) than dragonfly groups (%d)", world, P.G);
        return 1;
    }
    dfd_partition(P.G, P.a, P.p, rank, world, &dev->X);
    int devid = 0;
    CK(cudaGetDevice(&devid));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, devid));
    dev->num_sms = prop.multiProcessorCount;
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, devid));
    if (!coop) {
        snprintf(dev->error, sizeof dev->error, "device does not support cooperative launch");
        return 1;
    }
    // own-slice-only allocation: private arrays hold this rank's nodes and routers, the exchange region is laid out
    // for the largest partition of any rank so that the same offset is valid on every rank
    int gmax = 0;
    for (int k = 0; k < world; k++) gmax = std::max(gmax, dev->X.grp_lo[k + 1] - dev->X.grp_lo[k]);
    dev->nrl = dev->X.r1 - dev->X.r0;
    dev->ntl = dev->X.n1 - dev->X.n0;
    dev->nrmax = gmax * P.a;
    dev->ntmax = dev->nrmax * P.p;
    const size_t NT = (size_t)dev->ntl, NR = (size_t)dev->nrl, NRX = (size_t)dev->nrmax, NTX = (size_t)dev->ntmax, radix = (size_t)P.radix;
    if (dalloc(dev, &A.nev, NT * P.nev_cap)) return 1;
    if (dalloc(dev, &A.sq, NT * P.nsq_cap)) return 1;
    if (dalloc(dev, &A.mq, NT * P.nsq_cap)) return 1;
    if (dalloc(dev, &A.tq, NT * P.ntq_cap)) return 1;
    if (dalloc(dev, &A.ncin, NT * P.ncin_cap)) return 1;
    if (dalloc(dev, &A.nkin, NT * P.nkin_cap)) return 1;
    if (dalloc(dev, &A.mtab, NT * (size_t)P.nreasm_cap)) return 1;
    if (dalloc(dev, &A.ptab, NT * (size_t)P.nreasm_cap)) return 1;
    if (dalloc(dev, &A.ncin_tail, NT)) return 1;
    if (dalloc(dev, &A.nkin_tail, NT)) return 1;
    if (dalloc(dev, &A.portq, NR * radix)) return 1;
    if (dalloc(dev, &A.pollhdr, NR)) return 1;
    if (dalloc(dev, &A.skey, NR * radix)) return 1;
    if (dalloc(dev, &A.ckey, NR * radix)) return 1;
    if (dalloc(dev, &A.kkey, NR * radix)) return 1;
    if (dalloc(dev, &A.lane, NR * radix)) return 1;
    if (dalloc(dev, &A.seen_c, NR * radix)) return 1;
    if (dalloc(dev, &A.seen_k, NR * radix)) return 1;
    if (dalloc(dev, &A.pub_c, NR * radix)) return 1;
    if (dalloc(dev, &A.pub_k, NR * radix)) return 1;
    if (dalloc(dev, &A.tgt_c, NR * radix)) return 1;
    if (dalloc(dev, &A.tgt_k, NR * radix)) return 1;
    if (dalloc(dev, &A.nq, NR * radix * radix + 16)) return 1;
    if (dalloc(dev, &A.qbits, NR * radix * 2)) return 1;
    if (dalloc(dev, &A.iblk, NR * 2 * P.p)) return 1;
    if (dalloc(dev, &A.pool, NR * P.rpool_cap)) return 1;
    if (dalloc(dev, &A.pfree, NR * P.rpool_cap)) return 1;
    // exchange region: everything another rank may write (control block, channel words, lane rings, ACK statistics)
    // in one allocation behind one CUDA-IPC handle
    size_t off = align256(sizeof(DfdCtrl));
    size_t o_cwc = off;
    off = align256(off + NRX * radix * sizeof(unsigned long long));
    size_t o_cwk = off;
    off = align256(off + NRX * radix * sizeof(unsigned long long));
    size_t o_rk = off;
    off = align256(off + NRX * (size_t)P.kslots * sizeof(Key));
    size_t o_rc = off;
    off = align256(off + NRX * (size_t)P.cslots * sizeof(ChunkRec));
    size_t o_ac = off;
    off = align256(off + NTX * sizeof(unsigned));
    size_t o_af = off;
    off = align256(off + NTX * sizeof(unsigned long long));
    size_t o_al = off;
    off = align256(off + NTX * sizeof(unsigned long long));
    // ... and the per-LP statistics (read by rank 0 over the same mapping when the run is over)
    size_t o_node = off;
    off = align256(off + NTX * sizeof(NodeState));
    size_t o_rh = off;
    off = align256(off + NRX * sizeof(RouterHdr));
    size_t o_port = off;
    off = align256(off + NRX * radix * sizeof(PortState));
    size_t o_cnt = off;
    off = align256(off + 64 * sizeof(unsigned long long));
    size_t o_snap = off;
    off = align256(off + NRX * (size_t)P.nsnap_all * radix * sizeof(int4));
    char* xr = nullptr;
    if (dalloc(dev, &xr, off)) return 1;
    dev->xregion = xr;
    dev->xregion_bytes = off;
    dev->x_ack_off = o_ac;
    dev->x_cnt_off = o_cnt;
    A.node = (NodeState*)(xr + o_node);
    A.rh = (RouterHdr*)(xr + o_rh);
    A.port = (PortState*)(xr + o_port);
    A.counters = (unsigned long long*)(xr + o_cnt);
    A.snap = (int4*)(xr + o_snap);
    {
        double* d = nullptr;
        unsigned* u = nullptr;
        if (dalloc(dev, &d, (size_t)P.nsnap_all)) return 1;
        if (dalloc(dev, &u, (size_t)P.nsnap_all)) return 1;
        A.snap_ts = d;
        A.snap_id = u;
    }
    A.ctrl = (DfdCtrl*)xr;
    A.cw_c = (unsigned long long*)(xr + o_cwc);
    A.cw_k = (unsigned long long*)(xr + o_cwk);
    A.ring_k = (Key*)(xr + o_rk);
    A.ring_c = (ChunkRec*)(xr + o_rc);
    A.ack_count = (unsigned*)(xr + o_ac);
    A.ack_first = (unsigned long long*)(xr + o_af);
    A.ack_last = (unsigned long long*)(xr + o_al);
    for (int k = 0; k < DFD_MAX_RANKS; k++) dev->X.base[k] = nullptr;
    dev->X.base[rank] = xr;
    dev->X.local_base = xr;
    dev->X.local_bytes = off;
    {
        std::vector<uint2> geo;
        lane_geometry(P, &geo);
        uint2* d = nullptr;
        if (dalloc(dev, &d, geo.size())) return 1;
        CK(cudaMemcpy(d, geo.data(), geo.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        A.lgeo = d;
    }
    // connection tables
    std::vector<const std::vector<int>*> tabs = topo_tables(T);
    const int** slots[] = {&A.loc_off, &A.loc_port, &A.loc_dest, &A.lout_lane, &A.lin_src, &A.gs_port, &A.gs_group, &A.gdest, &A.gout_lane, &A.gin_src, &A.gin_port, &A.routed_off, &A.routed_lid};
    dev->topo_sizes.clear();
    for (size_t i = 0; i < tabs.size(); i++) {
        int* d = nullptr;
        if (dalloc(dev, &d, tabs[i]->size())) return 1;
        *slots[i] = d;
        dev->topo_sizes.push_back(tabs[i]->size());
    }
    dev->topo_bytes = 0;
    // grid: persistent, every block (= one warp) co-resident: warps wait for one another's channel words.  One group
    // per warp while the device has room, several per warp beyond that.
    dev->block = 32;
    const size_t groups = NR;
    bool wide = groups <= (size_t)dev->num_sms * 8 && ensemble <= 1;
    if (const char* ev = getenv("DFD_WIDE_REGS")) wide = atoi(ev) != 0;
    dev->kernel = wide ? (const void*)dfd_run_kernel<8> : (const void*)dfd_run_kernel<16>;
    if (const char* ev = getenv("DFD_MINBLOCKS")) {
        int mb = atoi(ev);
        dev->kernel = mb == 8 ? (const void*)dfd_run_kernel<8> : mb == 32 ? (const void*)dfd_run_kernel<32> : (const void*)dfd_run_kernel<16>;
    }
    dev->smem = 0;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (void (*)(DfdParams, DfdArrays, DfdPeers, unsigned long long))dev->kernel, dev->block, dev->smem));
    if (per_sm < 1) {
        snprintf(dev->error, sizeof dev->error, "run kernel does not fit on an SM");
        return 1;
    }
    dev->blocks_per_sm = per_sm;
    size_t slots_total = (size_t)per_sm * dev->num_sms;
    if (ensemble > 1) slots_total = std::max<size_t>(1, slots_total / ensemble);  // the GPU is shared by `ensemble` engines
    // Sharded runs: the last blocks of the grid are publishers (PubCtl in dfd_types.h); DFD_NPUB=0 goes back to one system-scope
    // fence per activation
    // One warp per group: an activation that ends in fence.release.sys stalls its group for 5 us at N=8 (a run with gpu-scope
    // fences, timing only, is 24 % faster), so the words for other ranks can go through publisher blocks: as many as fit beside
    // the groups, 4..32 (a publisher moves at most 32 words per fence).  The word still reaches the other rank only after a
    // system-scope fence -- the publisher's -- so what is gained is the stall, not the latency: measured on the bench network,
    // 8 GPUs (1,028 groups per GPU) 980 M events/s with publishers against 940 M without; 4 GPUs (2,056 groups per GPU, twice the
    // words per GPU) 576 M against 619 M; 2 GPUs (pooled kernel, whose other warps hide the fence) 370 M against 425 M.  Hence:
    // publishers only for partitions of at most 8 groups per SM (the 226-register kernel), and for ranks that share a GPU (tests).
    int npub_blocks = 0;
    if (world > 1 && ensemble > 1) npub_blocks = 1;
    if (world > 1 && ensemble <= 1 && wide) npub_blocks = (int)std::max<long long>(4, std::min<long long>(32, (long long)slots_total - (long long)groups));
    if (const char* ev = getenv("DFD_NPUB")) npub_blocks = world > 1 ? std::max(0, std::min(16, atoi(ev))) : 0;
    if (slots_total < (size_t)npub_blocks + 1) npub_blocks = 0;
    dev->grid = (int)std::max<size_t>(1, std::min(groups, slots_total - npub_blocks));
    if (const char* ev = getenv("DFD_GRID")) dev->grid = std::max(1, std::min(dev->grid, atoi(ev)));
    int npubq = npub_blocks;  // one warp per block: one queue per publisher block
    // More groups than resident warps: one 16-warp block per SM whose warps share the block's groups (dfd_run_kernel_pool)
    bool pool = groups > (size_t)dev->grid && ensemble <= 1 && !getenv("DFD_GRID") && !getenv("DFD_MINBLOCKS");
    if (const char* ev = getenv("DFD_POOL")) pool = atoi(ev) != 0 && ensemble <= 1;
    if (pool) {
        constexpr int W = DFD_POOL_WARPS;
        const int pool_pub = npub_blocks && getenv("DFD_POOL_PUB") ? 1 : 0;  // (experiment) one publisher block of W warps: W queues
        const int nb = (int)std::min<size_t>((size_t)dev->num_sms - pool_pub, (groups + W - 1) / W);
        const size_t per_block = (groups + nb - 1) / nb;
        const size_t smem = per_block * sizeof(unsigned);
        const void* k = (const void*)dfd_run_kernel_pool<W>;
        int fits = 0;
        if (smem <= 160 * 1024) {
            if (smem > 40 * 1024) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fits, (void (*)(DfdParams, DfdArrays, DfdPeers, unsigned long long))k, W * 32, smem));
        }
        if (fits >= 1) {
            dev->kernel = k;
            dev->block = W * 32;
            dev->grid = nb;
            dev->smem = smem;
            dev->blocks_per_sm = 1;
            npub_blocks = pool_pub;
            npubq = pool_pub * W;
        }
    }
    dev->warps = dev->grid * (dev->block / 32);  // worker warps
    dev->grid += npub_blocks;
    dev->P.npub_blocks = npub_blocks;
    dev->P.npubq = npubq;
    if (dalloc(dev, &A.pubq, (size_t)npubq * DFD_PUBQ_CAP)) return 1;
    if (dalloc(dev, &A.pubctl, (size_t)npubq)) return 1;
    // Event budget of one activation.  With a warp per group the run is bound by the chain of dependent activations:
    // publishing early lets the neighbours start.  With several groups per warp it is bound by throughput: the fixed
    // cost of an activation (poll, bounds, publish) is spread over more events.
    // With a warp per group the best budget grows with the radix -- the fixed cost of an activation does -- measured on B200
    // (tools/gpu_batch24.sh): radix 15: 2 (13.9 M events/s against 12.7 M with 4), radix 31: 4-6, radix 63: 10-12 (178 M against
    // 154 M with 4 on 1,056 routers, 196 M against 170 M on 2,080).
    if (!getenv("DFD_ACT_BUDGET")) dev->P.act_budget = groups > (size_t)dev->warps ? 16 : std::max(2, std::min(16, P.radix / 6));
    // Idle warps of a throughput-bound run compete with the busy ones for instruction fetch: let them pause.  With a
    // warp per group the poll latency is on the critical path: no pause.
    dev->P.idle_sleep_ns = groups > (size_t)dev->warps ? 8000 : 0;
    dev->P.prefetch = 0;  // measured on B200 at 131,584 and 1,050,624 terminals: no gain (-2 %), kept as a knob
    if (const char* ev = getenv("DFD_PREFETCH")) dev->P.prefetch = atoi(ev);
    if (const char* ev = getenv("DFD_IDLE_SLEEP_NS")) dev->P.idle_sleep_ns = std::max(0, atoi(ev));
    return 0;
}

// CUDA-IPC plumbing of the exchange regions (one process per GPU)
int dfd_device_ipc_export(DfdDevice* dev, void* handle64) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, dev->xregion));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return 0;
}
int dfd_device_ipc_import(DfdDevice* dev, const void* handles, int world) {
    if (world != dev->X.world) {
        snprintf(dev->error, sizeof dev->error, "ipc import: %d handles for a world of %d", world, dev->X.world);
        return 1;
    }
    for (int k = 0; k < world; k++) {
        if (k == dev->X.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * k, 64);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        dev->X.base[k] = (char*)ptr;
        dev->ipc_opened.push_back(ptr);
    }
    return 0;
}

int dfd_device_upload_topo(DfdDevice* dev, const DfdHostTopo& T, cudaStream_t st) {
    DfdArrays& A = dev->A;
    std::vector<const std::vector<int>*> tabs = topo_tables(T);
    const int* dst[] = {A.loc_off, A.loc_port, A.loc_dest, A.lout_lane, A.lin_src, A.gs_port, A.gs_group, A.gdest, A.gout_lane, A.gin_src, A.gin_port, A.routed_off, A.routed_lid};
    size_t total = 0;
    for (size_t i = 0; i < tabs.size(); i++) {
        if (tabs[i]->size() > dev->topo_sizes[i]) {
            snprintf(dev->error, sizeof dev->error, "connection table %zu grew from %zu to %zu entries: the device arrays must be re-created", i, dev->topo_sizes[i], tabs[i]->size());
            return 1;
        }
        if (!tabs[i]->empty()) CK(cudaMemcpyAsync((void*)dst[i], tabs[i]->data(), tabs[i]->size() * sizeof(int), cudaMemcpyHostToDevice, st));
        total += tabs[i]->size() * sizeof(int);
    }
    if ((int)T.snap_ts.size() != dev->P.nsnap || dev->P.nsnap > dev->P.nsnap_all) {
        snprintf(dev->error, sizeof dev->error, "snapshot schedule does not match the device parameters");
        return 1;
    }
    if (dev->P.nsnap) {
        CK(cudaMemcpyAsync((void*)A.snap_ts, T.snap_ts.data(), T.snap_ts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync((void*)A.snap_id, T.snap_id.data(), T.snap_id.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        total += T.snap_ts.size() * (sizeof(double) + sizeof(unsigned));
    }
    dev->topo_bytes = total;
    return 0;
}

int dfd_device_reset(DfdDevice* dev, cudaStream_t st) {
    int blocks = dev->num_sms * 8;
    dfd_init_kernel<<<blocks, 256, 0, st>>>(dev->P, dev->A, dev->X, dev->seeds);
    CK(cudaGetLastError());
    return 0;
}

int dfd_device_run(DfdDevice* dev, cudaStream_t st) {
    for (int k = 0; k < dev->X.world; k++)
        if (!dev->X.base[k]) {
            snprintf(dev->error, sizeof dev->error, "rank %d: the exchange region of rank %d has not been imported", dev->X.rank, k);
            return 1;
        }
    // watchdog: no event executed anywhere for this long while events are pending -> DFD_ERR_STALL instead of a hang
    unsigned long long timeout_ns = 20ULL * 1000000000ULL;
    if (const char* ev = getenv("DFD_STALL_TIMEOUT_S")) timeout_ns = (unsigned long long)(atof(ev) * 1e9);
    void* args[] = {(void*)&dev->P, (void*)&dev->A, (void*)&dev->X, (void*)&timeout_ns};
    CK(cudaLaunchCooperativeKernel(dev->kernel, dim3(dev->grid), dim3(dev->block), args, dev->smem, st));
    return 0;
}

int dfd_device_download(DfdDevice* dev, DfdHostResults* out, cudaStream_t st, int mode) {
    const DfdParams& P = dev->P;
    const DfdArrays& A = dev->A;
    const DfdPeers& X = dev->X;
    out->node.resize(P.NT);
    out->rh.resize(P.NR);
    out->port.resize((size_t)P.NR * P.radix);
    out->snap.resize((size_t)P.NR * P.nsnap_all * P.radix);
    out->ack_count.resize(P.NT);
    out->ack_first.resize(P.NT);
    out->ack_last.resize(P.NT);
    out->d2h_bytes = 0;
    DfdCtrl ctrl[DFD_MAX_RANKS];
    unsigned long long cnt[DFD_MAX_RANKS][64];
    const int k0 = mode == DFD_DL_ALL ? 0 : X.rank, k1 = mode == DFD_DL_ALL ? X.world : X.rank + 1;
    for (int k = k0; k < k1; k++) {
        // every array of the statistics sits at the same offset of every rank's exchange region
        const char* base = X.base[k];
        if (!base) {
            snprintf(dev->error, sizeof dev->error, "download: the exchange region of rank %d is not mapped", k);
            return 1;
        }
        auto at = [&](const void* local) { return base + ((const char*)local - X.local_base); };
        DfdPeers Y;
        dfd_partition(P.G, P.a, P.p, k, X.world, &Y);
        const size_t nt = (size_t)(Y.n1 - Y.n0), nr = (size_t)(Y.r1 - Y.r0);
        if (mode != DFD_DL_CTRL) {
            CK(cudaMemcpyAsync(out->node.data() + Y.n0, at(A.node), sizeof(NodeState) * nt, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(out->rh.data() + Y.r0, at(A.rh), sizeof(RouterHdr) * nr, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(out->port.data() + (size_t)Y.r0 * P.radix, at(A.port), sizeof(PortState) * nr * P.radix, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(out->ack_count.data() + Y.n0, at(A.ack_count), sizeof(unsigned) * nt, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(out->ack_first.data() + Y.n0, at(A.ack_first), sizeof(unsigned long long) * nt, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(out->ack_last.data() + Y.n0, at(A.ack_last), sizeof(unsigned long long) * nt, cudaMemcpyDeviceToHost, st));
            const size_t snap_per_router = (size_t)P.nsnap_all * P.radix;
            if (P.nsnap)
                CK(cudaMemcpyAsync(out->snap.data() + (size_t)Y.r0 * snap_per_router, at(A.snap), sizeof(int4) * nr * snap_per_router, cudaMemcpyDeviceToHost, st));
            out->d2h_bytes += sizeof(NodeState) * nt + sizeof(RouterHdr) * nr + sizeof(PortState) * nr * P.radix + (sizeof(unsigned) + 2 * sizeof(unsigned long long)) * nt +
                              (P.nsnap ? sizeof(int4) * nr * snap_per_router : 0);
        }
        CK(cudaMemcpyAsync(cnt[k], at(A.counters), sizeof cnt[k], cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ctrl[k], at(A.ctrl), sizeof(DfdCtrl), cudaMemcpyDeviceToHost, st));
        out->d2h_bytes += sizeof cnt[k] + sizeof(DfdCtrl);
    }
    CK(cudaStreamSynchronize(st));
    memset(out->counters, 0, sizeof out->counters);
    for (int k = k0; k < k1; k++) {
        for (int i = 0; i < 64; i++)
            if (i == 24)
                out->counters[i] = std::max(out->counters[i], cnt[k][i]);  // busiest warp
            else
                out->counters[i] += cnt[k][i];
    }
    out->counters[2] = out->counters[3] = 0;
    for (int k = k0; k < k1; k++) {
        out->counters[26] += ctrl[k].created;
        out->counters[27] += ctrl[k].consumed;
        // the first error: a rank's own error wins over the "a peer failed" notice it may have received
        if (ctrl[k].error && (!out->counters[2] || (out->counters[2] == DFD_ERR_PEER_ERROR && ctrl[k].error != DFD_ERR_PEER_ERROR))) {
            out->counters[2] = ctrl[k].error;
            out->counters[3] = ctrl[k].error_info;
        }
    }
    return 0;
}

void dfd_device_destroy(DfdDevice* dev) {
    for (void* p : dev->ipc_opened) cudaIpcCloseMemHandle(p);
    dev->ipc_opened.clear();
    for (void* p : dev->allocs) cudaFree(p);
    dev->allocs.clear();
    dev->bytes_allocated = 0;
}
