// dfd_cpusim.cc -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for codes_b200/csrc/dfd_engine.cu.
//
// Compiles codes_b200/csrc/dfd_core.h / dfd_init.h (the asynchronous scheduler and the LP handlers, the very code
// the CUDA kernel runs) as plain C++ and drives the group activations from a seeded random sequential scheduler.
// Linked with dfd_host.cc into tests/cpusim/libdfdally_cpusim.so it lets the `-m "not gpu"` suite check, without a
// GPU, that the channel-word bounds never let an LP run ahead of an event it has yet to receive (DFD_ERR_LATE_EVENT),
// that the scheme cannot deadlock (a sweep that changes nothing while events are pending is an error), and that the
// results are byte-identical to the oracle for any interleaving.  It is never loaded by the product: codes_b200
// binds libdfdally_b200.so only, which has no CPU path.
//
//   DFD_CPUSIM_SEED   scheduler seed (default 1)
//   DFD_CPUSIM_WORLD  emulate a group-sharded run over N ranks inside one address space (default 1)
//   DFD_CPUSIM_WARPS  warps per rank (default: one per group)
//   DFD_CPUSIM_SYNC   1: synchronous rounds -- every group is activated once per round and sees only what was published
//                     in earlier rounds; the number of rounds is the length of the dependency chain the bounds leave
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <vector>

#include "../../codes_b200/csrc/dfd_engine.h"
#include "../../codes_b200/csrc/dfd_init.h"

// ---- the two CUDA runtime calls dfd_host.cc makes ------------------------------------------------------------------
extern "C" cudaError_t cudaGetDeviceCount(int* count) {
    *count = 1;
    return cudaSuccess;
}
extern "C" cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
extern "C" cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) {
    *p = malloc(bytes);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
extern "C" cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
extern "C" cudaError_t cudaGetLastError(void) { return cudaSuccess; }

namespace {
struct Rank {
    DfdArrays A;
    DfdPeers X;
    std::vector<void*> allocs;
    int ntl = 0, nrl = 0;
};
struct Sim {
    std::vector<Rank> ranks;
    int world = 1;
};
Sim* g_sim = nullptr;  // one simulated device per process is all the tests need; re-created by dfd_device_create

template <typename T> T* halloc(Rank& R, size_t n) {
    if (n == 0) n = 1;
    T* p = (T*)calloc(n, sizeof(T));
    R.allocs.push_back(p);
    return p;
}
}  // namespace

void dfd_partition(int G, int a, int p, int rank, int world, DfdPeers* X) {
    memset(X, 0, sizeof *X);
    X->rank = rank;
    X->world = world;
    for (int k = 0; k <= world; k++) X->grp_lo[k] = (int)(((long long)k * G) / world);
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
        (*g)[2 * j] = uint2{coff, ccap - 1};
        (*g)[2 * j + 1] = uint2{koff, kcap - 1};
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int dfd_device_create(DfdDevice* dev, const DfdParams& P, const DfdHostTopo& T, int rank, int world, int) {
    if (rank != 0 || world != 1) {
        snprintf(dev->error, sizeof dev->error, "cpusim: set DFD_CPUSIM_WORLD to emulate a sharded run");
        return 1;
    }
    dev->P = P;
    int W = 1;
    if (const char* ev = getenv("DFD_CPUSIM_WORLD")) W = std::max(1, std::min(DFD_MAX_RANKS, atoi(ev)));
    if (W > P.G) W = P.G;
    delete g_sim;
    g_sim = new Sim();
    g_sim->world = W;
    g_sim->ranks.resize(W);
    int gmax = 0;
    for (int k = 0; k < W; k++) {
        DfdPeers X;
        dfd_partition(P.G, P.a, P.p, k, W, &X);
        gmax = std::max(gmax, X.grp_lo[k + 1] - X.grp_lo[k]);
    }
    const size_t NRX = (size_t)gmax * P.a, NTX = NRX * P.p, radix = (size_t)P.radix;
    // topology tables are shared by the emulated ranks
    std::vector<const std::vector<int>*> tabs = {&T.loc_off, &T.loc_port, &T.loc_dest, &T.lout_lane, &T.lin_src, &T.gs_port, &T.gs_group, &T.gdest, &T.gout_lane, &T.gin_src, &T.gin_port, &T.routed_off, &T.routed_lid};
    std::vector<int*> tcopy;
    for (auto* v : tabs) {
        int* d = (int*)calloc(std::max<size_t>(1, v->size()), sizeof(int));
        memcpy(d, v->data(), v->size() * sizeof(int));
        tcopy.push_back(d);
    }
    for (int k = 0; k < W; k++) {
        Rank& R = g_sim->ranks[k];
        dfd_partition(P.G, P.a, P.p, k, W, &R.X);
        R.nrl = R.X.r1 - R.X.r0;
        R.ntl = R.X.n1 - R.X.n0;
        const size_t NT = (size_t)R.ntl, NR = (size_t)R.nrl;
        DfdArrays& A = R.A;
        memset(&A, 0, sizeof A);
        if (k == 0)
            for (int* d : tcopy) R.allocs.push_back(d);
        A.node = halloc<NodeState>(R, NT);
        A.nev = halloc<NodeEv>(R, NT * P.nev_cap);
        A.sq = halloc<SchedReq>(R, NT * P.nsq_cap);
        A.mq = halloc<MsgReq>(R, NT * P.nsq_cap);
        A.tq = halloc<TermChunk>(R, NT * P.ntq_cap);
        A.ncin = halloc<ChunkRec>(R, NT * P.ncin_cap);
        A.nkin = halloc<Key>(R, NT * P.nkin_cap);
        A.mtab = halloc<uint4>(R, NT * (size_t)P.nreasm_cap);
        A.ptab = halloc<uint4>(R, NT * (size_t)P.nreasm_cap);
        A.ncin_tail = halloc<unsigned>(R, NT);
        A.nkin_tail = halloc<unsigned>(R, NT);
        A.rh = halloc<RouterHdr>(R, NR);
        A.port = halloc<PortState>(R, NR * radix);
        A.portq = halloc<PortQ>(R, NR * radix);
        A.pollhdr = halloc<PollHdr>(R, NR);
        A.skey = halloc<Key>(R, NR * radix);
        A.ckey = halloc<Key>(R, NR * radix);
        A.kkey = halloc<Key>(R, NR * radix);
        A.lane = halloc<LaneCtl>(R, NR * radix);
        A.seen_c = halloc<unsigned long long>(R, NR * radix);
        A.seen_k = halloc<unsigned long long>(R, NR * radix);
        A.pub_c = halloc<unsigned long long>(R, NR * radix);
        A.pub_k = halloc<unsigned long long>(R, NR * radix);
        A.tgt_c = halloc<unsigned long long*>(R, NR * radix);
        A.tgt_k = halloc<unsigned long long*>(R, NR * radix);
        A.nq = halloc<unsigned char>(R, NR * radix * radix + 16);
        A.qbits = halloc<unsigned long long>(R, NR * radix * 2);
        A.iblk = halloc<double>(R, NR * 2 * P.p);
        A.pool = halloc<PoolSlot>(R, NR * P.rpool_cap);
        A.pfree = halloc<unsigned>(R, NR * P.rpool_cap);
        size_t off = align256(sizeof(DfdCtrl));
        size_t o_cwc = off;
        off = align256(off + NRX * radix * 8);
        size_t o_cwk = off;
        off = align256(off + NRX * radix * 8);
        size_t o_rk = off;
        off = align256(off + NRX * (size_t)P.kslots * sizeof(Key));
        size_t o_rc = off;
        off = align256(off + NRX * (size_t)P.cslots * sizeof(ChunkRec));
        size_t o_ac = off;
        off = align256(off + NTX * 4);
        size_t o_af = off;
        off = align256(off + NTX * 8);
        size_t o_al = off;
        off = align256(off + NTX * 8);
        char* xr = halloc<char>(R, off + 256);
        xr = (char*)(((uintptr_t)xr + 255) & ~(uintptr_t)255);
        A.ctrl = (DfdCtrl*)xr;
        A.cw_c = (unsigned long long*)(xr + o_cwc);
        A.cw_k = (unsigned long long*)(xr + o_cwk);
        A.ring_k = (Key*)(xr + o_rk);
        A.ring_c = (ChunkRec*)(xr + o_rc);
        A.ack_count = (unsigned*)(xr + o_ac);
        A.ack_first = (unsigned long long*)(xr + o_af);
        A.ack_last = (unsigned long long*)(xr + o_al);
        A.counters = halloc<unsigned long long>(R, 64);
        A.snap = halloc<int4>(R, NR * (size_t)P.nsnap_all * radix + 1);
        A.snap_ts = halloc<double>(R, (size_t)P.nsnap_all + 1);
        A.snap_id = halloc<unsigned>(R, (size_t)P.nsnap_all + 1);
        {
            std::vector<uint2> geo;
            lane_geometry(P, &geo);
            uint2* d = halloc<uint2>(R, geo.size());
            memcpy(d, geo.data(), geo.size() * sizeof(uint2));
            A.lgeo = d;
        }
        A.loc_off = tcopy[0];
        A.loc_port = tcopy[1];
        A.loc_dest = tcopy[2];
        A.lout_lane = tcopy[3];
        A.lin_src = tcopy[4];
        A.gs_port = tcopy[5];
        A.gs_group = tcopy[6];
        A.gdest = tcopy[7];
        A.gout_lane = tcopy[8];
        A.gin_src = tcopy[9];
        A.gin_port = tcopy[10];
        A.routed_off = tcopy[11];
        A.routed_lid = tcopy[12];
        R.X.local_base = xr;
    }
    for (int k = 0; k < W; k++)
        for (int j = 0; j < W; j++) g_sim->ranks[k].X.base[j] = g_sim->ranks[j].X.local_base;
    dfd_partition(P.G, P.a, P.p, 0, 1, &dev->X);
    dev->A = g_sim->ranks[0].A;
    dev->ntl = P.NT;
    dev->nrl = P.NR;
    dev->grid = P.NR;
    dev->block = 32;
    dev->num_sms = 1;
    dev->bytes_allocated = 0;
    dev->topo_sizes.clear();
    for (auto* v : tabs) dev->topo_sizes.push_back(v->size());
    return 0;
}

int dfd_device_ipc_export(DfdDevice* dev, void*) {
    snprintf(dev->error, sizeof dev->error, "cpusim: no IPC");
    return 1;
}
int dfd_device_ipc_import(DfdDevice* dev, const void*, int) {
    snprintf(dev->error, sizeof dev->error, "cpusim: no IPC");
    return 1;
}
int dfd_device_upload_topo(DfdDevice* dev, const DfdHostTopo& T, cudaStream_t) {
    // tables were copied at create; re-copy (same sizes: checked by the caller's reuse test)
    const std::vector<int>* tabs[] = {&T.loc_off, &T.loc_port, &T.loc_dest, &T.lout_lane, &T.lin_src, &T.gs_port, &T.gs_group, &T.gdest, &T.gout_lane, &T.gin_src, &T.gin_port, &T.routed_off, &T.routed_lid};
    const DfdArrays& A = g_sim->ranks[0].A;
    const int* dst[] = {A.loc_off, A.loc_port, A.loc_dest, A.lout_lane, A.lin_src, A.gs_port, A.gs_group, A.gdest, A.gout_lane, A.gin_src, A.gin_port, A.routed_off, A.routed_lid};
    size_t total = 0;
    for (int i = 0; i < 13; i++) {
        memcpy((void*)dst[i], tabs[i]->data(), tabs[i]->size() * sizeof(int));
        total += tabs[i]->size() * sizeof(int);
    }
    for (Rank& R : g_sim->ranks) {
        memcpy((void*)R.A.snap_ts, T.snap_ts.data(), T.snap_ts.size() * sizeof(double));
        memcpy((void*)R.A.snap_id, T.snap_id.data(), T.snap_id.size() * sizeof(unsigned));
    }
    dev->topo_bytes = total;
    return 0;
}

int dfd_device_reset(DfdDevice* dev, cudaStream_t) {
    const DfdParams& P = dev->P;
    for (Rank& R : g_sim->ranks) {
        init_ctrl(P, R.A, R.X);
        for (int n = 0; n < R.ntl; n++) init_node(P, R.A, R.X, dev->seeds, (unsigned)n);
        for (int r = 0; r < R.nrl; r++) init_router(P, R.A, R.X, dev->seeds, (unsigned)r);
        for (size_t i = 0; i < (size_t)R.nrl * P.radix; i++) init_lane(P, R.A, R.X, i);
        memset(R.A.nq, 0, (size_t)R.nrl * P.radix * P.radix);
        for (size_t i = 0; i < (size_t)R.nrl * P.rpool_cap; i++) R.A.pfree[i] = (unsigned)P.rpool_cap - 1 - (unsigned)(i % P.rpool_cap);
    }
    return 0;
}

int dfd_device_run(DfdDevice* dev, cudaStream_t) {
    const DfdParams& P = dev->P;
    unsigned seed = 1;
    if (const char* ev = getenv("DFD_CPUSIM_SEED")) seed = (unsigned)atoi(ev);
    std::mt19937 rng(seed);
    struct Warp {
        int rank;
        std::vector<unsigned> groups;
    };
    std::vector<Warp> warps;
    for (int k = 0; k < g_sim->world; k++) {
        Rank& R = g_sim->ranks[k];
        int nw = R.nrl;
        if (const char* ev = getenv("DFD_CPUSIM_WARPS")) nw = std::max(1, std::min(R.nrl, atoi(ev)));
        for (int w = 0; w < nw; w++) {
            Warp wp;
            wp.rank = k;
            for (int g = w; g < R.nrl; g += nw) wp.groups.push_back((unsigned)g);
            warps.push_back(wp);
        }
    }
    std::vector<unsigned> evcount(EV_NTYPES * g_sim->world, 0);
    unsigned long long productive = 0, activations = 0, passes = 0;
    // schedule: sweeps over all warps in random order, with random repeats in front.  The run ends when every created
    // event has been consumed (the monitor's test); a sweep in which no poll sees news, or a long series of sweeps that
    // execute nothing, while events are pending is a deadlock.
    unsigned long long barren = 0, barren_limit = 2000000;
    if (const char* ev = getenv("DFD_CPUSIM_BARREN")) barren_limit = strtoull(ev, nullptr, 10);
    const bool sync_rounds = getenv("DFD_CPUSIM_SYNC") && atoi(getenv("DFD_CPUSIM_SYNC"));
    // which instantiation of the activation runs: the pooled kernel's (router horizon in a pass of its own) or the
    // one-warp-per-block kernels' (folded into every pick scan); by default the seed decides, so the suite covers both
    const bool split_horizon = getenv("DFD_CPUSIM_SPLIT") ? atoi(getenv("DFD_CPUSIM_SPLIT")) != 0 : (seed & 1u) != 0;
    std::vector<std::pair<unsigned long long*, unsigned long long>> deferred;
    unsigned long long rounds = 0;
    for (;;) {
        bool any = false;
        unsigned long long executed = 0;
        rounds++;
        if (sync_rounds) dfd_host_deferred_words = &deferred;
        std::vector<int> order(warps.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
        std::shuffle(order.begin(), order.end(), rng);
        // a few random extra activations in front of the sweep: repeated and skewed progress
        size_t extra = warps.size() && !sync_rounds ? rng() % (warps.size() + 1) : 0;
        for (size_t i = 0; i < extra; i++) order.insert(order.begin(), (int)(rng() % warps.size()));
        for (int wi : order) {
            Warp& wp = warps[wi];
            Rank& R = g_sim->ranks[wp.rank];
            Ctx c;
            memset(&c, 0, sizeof c);
            c.P = &P;
            c.A = &R.A;
            c.X = &R.X;
            c.ev = &evcount[(size_t)wp.rank * EV_NTYPES];
            for (unsigned g : wp.groups) {
                ActResult a = split_horizon ? group_activate<true>(c, g) : group_activate<false>(c, g);
                if (a.polled_change) {
                    any = true;
                    activations++;
                }
                if (a.nexec) productive++;
                executed += a.nexec;
            }
            passes++;
            bool err = false;
            for (Rank& Q : g_sim->ranks) err |= Q.A.ctrl->error != 0;
            if (err) goto done;
        }
        dfd_host_deferred_words = nullptr;
        for (auto& pw : deferred) *pw.first = pw.second;
        if (!deferred.empty()) any = true;
        deferred.clear();
        unsigned long long created = 0, consumed = 0;
        for (Rank& R : g_sim->ranks) {
            created += R.A.ctrl->created;
            consumed += R.A.ctrl->consumed;
        }
        if (created == consumed) break;
        barren = executed ? 0 : barren + 1;
        if (!any || barren > barren_limit) {  // nothing can move although events are pending
            g_sim->ranks[0].A.ctrl->error = DFD_ERR_STALL;
            g_sim->ranks[0].A.ctrl->error_info = created - consumed;
            break;
        }
    }
done:
    dfd_host_deferred_words = nullptr;
    if (getenv("DFD_CPUSIM_VERBOSE")) fprintf(stderr, "cpusim: %llu rounds, %llu activations past the poll, %llu productive\n", rounds, activations, productive);
    for (int k = 0; k < g_sim->world; k++) {
        Rank& R = g_sim->ranks[k];
        for (int t = 0; t < EV_NTYPES; t++) {
            R.A.counters[4 + t] = evcount[(size_t)k * EV_NTYPES + t];
            R.A.counters[0] += evcount[(size_t)k * EV_NTYPES + t];
        }
    }
    g_sim->ranks[0].A.counters[1] = productive;
    g_sim->ranks[0].A.counters[22] = activations;
    g_sim->ranks[0].A.counters[23] = passes;
    return 0;
}

int dfd_device_download(DfdDevice* dev, DfdHostResults* out, cudaStream_t, int) {
    const DfdParams& P = dev->P;
    out->node.resize(P.NT);
    out->rh.resize(P.NR);
    out->port.resize((size_t)P.NR * P.radix);
    out->snap.resize((size_t)P.NR * P.nsnap_all * P.radix);
    out->ack_count.resize(P.NT);
    out->ack_first.resize(P.NT);
    out->ack_last.resize(P.NT);
    memset(out->counters, 0, sizeof out->counters);
    for (Rank& R : g_sim->ranks) {
        memcpy(out->node.data() + R.X.n0, R.A.node, sizeof(NodeState) * R.ntl);
        memcpy(out->rh.data() + R.X.r0, R.A.rh, sizeof(RouterHdr) * R.nrl);
        memcpy(out->port.data() + (size_t)R.X.r0 * P.radix, R.A.port, sizeof(PortState) * (size_t)R.nrl * P.radix);
        memcpy(out->snap.data() + (size_t)R.X.r0 * P.nsnap_all * P.radix, R.A.snap, sizeof(int4) * (size_t)R.nrl * P.nsnap_all * P.radix);
        memcpy(out->ack_count.data() + R.X.n0, R.A.ack_count, sizeof(unsigned) * R.ntl);
        memcpy(out->ack_first.data() + R.X.n0, R.A.ack_first, 8 * (size_t)R.ntl);
        memcpy(out->ack_last.data() + R.X.n0, R.A.ack_last, 8 * (size_t)R.ntl);
        for (int i = 0; i < 64; i++) out->counters[i] += R.A.counters[i];
        if (R.A.ctrl->error && !out->counters[2]) {
            out->counters[2] = R.A.ctrl->error;
            out->counters[3] = R.A.ctrl->error_info;
        }
        out->counters[26] += R.A.ctrl->created;
        out->counters[27] += R.A.ctrl->consumed;
    }
    // counters[2] was summed above: restore the first error code
    for (Rank& R : g_sim->ranks)
        if (R.A.ctrl->error) {
            out->counters[2] = R.A.ctrl->error;
            out->counters[3] = R.A.ctrl->error_info;
            break;
        }
    out->d2h_bytes = 0;
    return 0;
}

void dfd_device_destroy(DfdDevice* dev) {
    if (g_sim) {
        for (Rank& R : g_sim->ranks)
            for (void* p : R.allocs) free(p);
        delete g_sim;
        g_sim = nullptr;
    }
    dev->allocs.clear();
}
