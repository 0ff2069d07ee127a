// dfd_core.h -- LP handlers and the group activation of the asynchronous dragonfly-dally engine.
//
// Conservative, barrier-free scheduling.  A "group" (one router + its p compute nodes) is owned by one warp.  An
// activation of a group
//   (1) polls the channel words of the router's external lanes (one coalesced read),
//   (2) runs rounds of { router phase, node phase } until nothing more is safe: an LP executes its earliest pending
//       event only while its timestamp is below every bound of its currently empty input lanes,
//   (3) publishes, per external link, a chunk bound (no chunk will arrive at the neighbour before it) and a credit
//       bound (no credit will), computed from what this router actually holds:
//         chunk bound(q)  = max(next_output_available_time[q], time of the next possible R_SEND on q)
//                           + min serialisation + router_delay + link delay          (dally:7811-7833, 7910-7911)
//         credit bound(s) = credit delay + min( next R_ARRIVE on in-lane s,
//                                               next R_BUFFER on any port that holds overflowed chunks of s )
//       (a credit only leaves at an R_ARRIVE that finds room, dally:7537-7548, or at an R_BUFFER that promotes a
//       queued chunk, dally:8111-8129).
// Each LP still consumes its events in (timestamp, sender gid, sender sequence) order, so results are identical to
// the sequential reference order whatever the interleaving; a violated bound would surface as DFD_ERR_LATE_EVENT.
//
// The file compiles twice: as device code (nvcc, the product) and as plain C++ (tests/cpusim: the same scheduling
// logic driven by a randomised sequential scheduler, test infrastructure only).  Handlers restate the reference's
// forward handlers and cite the lines they follow ("dally" = src/networks/model-net/dragonfly-dally.cxx,
// "synthetic" = src/network-workloads/model-net-synthetic-dragonfly-all.c).
#pragma once
#include <string.h>

#include "dfd_types.h"

#define INF_BITS 0x7FF0000000000000ULL
#define NIL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// portability layer
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define DFD_FN __device__ __forceinline__
#define DFD_FN_NOINLINE __device__
#define DFD_FN_OUTLINE __device__ __noinline__  // shared or cold code: one copy instead of one per call site (instruction cache)
#define DFD_LANE() ((int)(threadIdx.x & 31u))
#define LANE_FOR(i, n) for (int i = DFD_LANE(); i < (int)(n); i += 32)
#define IS_LANE0() (DFD_LANE() == 0)
#define WARP_SYNC() __syncwarp()
DFD_FN double bits2d(unsigned long long b) { return __longlong_as_double((long long)b); }
DFD_FN unsigned long long d2bits(double d) { return (unsigned long long)__double_as_longlong(d); }
DFD_FN unsigned long long umul64hi(unsigned long long a, unsigned long long b) { return __umul64hi(a, b); }
DFD_FN int4 ld_cg16(const void* p) { return __ldcg((const int4*)p); }
DFD_FN void st_cg16(void* p, int4 v) { __stcg((int4*)p, v); }
DFD_FN unsigned long long ld_strong64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DFD_FN unsigned long long ld_strong64_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DFD_FN void st_strong64(unsigned long long* p, unsigned long long v, bool sys) {
    if (sys)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one 16-byte access, single-copy atomic (PTX .b128 scalar type): an entry of a publisher queue
DFD_FN void st_relaxed_b128(void* p, unsigned long long lo, unsigned long long hi) {
    asm volatile("{ .reg .b128 t; mov.b128 t, {%1, %2}; st.relaxed.gpu.global.b128 [%0], t; }" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
DFD_FN void ld_relaxed_b128(const void* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("{ .reg .b128 t; ld.relaxed.gpu.global.b128 t, [%2]; mov.b128 {%0, %1}, t; }" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
// release fence without the L1 invalidation an acq_rel fence carries (SASS: MEMBAR.ALL.{GPU,SYS}, no CCTL.IVALL):
// LP-private state stays L1-resident, everything another LP writes is read with ld.cg / ld.relaxed (L2)
DFD_FN void fence_release(bool sys) {
    if (sys)
        asm volatile("fence.release.sys;" ::: "memory");
    else
        asm volatile("fence.release.gpu;" ::: "memory");
}
DFD_FN void red_add_i64(long long* p, long long v, bool sys) {
    if (sys)
        asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DFD_FN void ack_update(unsigned* cnt, unsigned long long* first, unsigned long long* last, unsigned long long tb, bool sys) {
    if (sys) {
        atomicAdd_system(cnt, 1u);
        atomicMin_system(first, tb);
        atomicMax_system(last, tb);
    } else {
        atomicAdd(cnt, 1u);
        atomicMin(first, tb);
        atomicMax(last, tb);
    }
}
DFD_FN bool cas_u64(unsigned long long* p, unsigned long long expect, unsigned long long v) { return atomicCAS(p, expect, v) == expect; }
DFD_FN void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
DFD_FN void count_ev(unsigned* ev, int t) { atomicAdd(&ev[t], 1u); }
DFD_FN bool warp_any(bool p) { return __any_sync(0xFFFFFFFFu, p) != 0; }
DFD_FN unsigned warp_sum_u(unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
DFD_FN int warp_sum_i(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
#else  // ---- plain C++ (tests/cpusim) ----
#include <math.h>
#define DFD_FN static inline
#define DFD_FN_NOINLINE static
#define DFD_FN_OUTLINE static
#define DFD_LANE() 0
#define LANE_FOR(i, n) for (int i = 0; i < (int)(n); i++)
#define IS_LANE0() true
#define WARP_SYNC() ((void)0)
DFD_FN double bits2d(unsigned long long b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}
DFD_FN unsigned long long d2bits(double d) {
    unsigned long long b;
    memcpy(&b, &d, 8);
    return b;
}
DFD_FN unsigned long long umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
DFD_FN int4 ld_cg16(const void* p) { return *(const int4*)p; }
DFD_FN void st_cg16(void* p, int4 v) { *(int4*)p = v; }
DFD_FN unsigned long long ld_strong64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
DFD_FN unsigned long long ld_strong64_sys(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
// test hook (tests/cpusim, "synchronous rounds" mode): channel words published during a sweep become visible after it
#include <utility>
#include <vector>
static std::vector<std::pair<unsigned long long*, unsigned long long>>* dfd_host_deferred_words = nullptr;
DFD_FN void st_strong64(unsigned long long* p, unsigned long long v, bool) {
    if (dfd_host_deferred_words)
        dfd_host_deferred_words->push_back(std::make_pair(p, v));
    else
        *(volatile unsigned long long*)p = v;
}
DFD_FN void fence_release(bool) {}
DFD_FN void red_add_i64(long long* p, long long v, bool) { *p += v; }
DFD_FN void ack_update(unsigned* cnt, unsigned long long* first, unsigned long long* last, unsigned long long tb, bool) {
    (*cnt)++;
    if (tb < *first) *first = tb;
    if (tb > *last) *last = tb;
}
DFD_FN void prefetch_l1(const void*) {}
DFD_FN bool cas_u64(unsigned long long* p, unsigned long long expect, unsigned long long v) {
    if (*p != expect) return false;
    *p = v;
    return true;
}
DFD_FN void count_ev(unsigned* ev, int t) { ev[t]++; }
DFD_FN bool warp_any(bool p) { return p; }
DFD_FN unsigned warp_sum_u(unsigned v) { return v; }
DFD_FN int warp_sum_i(int v) { return v; }
#endif

#define DINF bits2d(INF_BITS)

DFD_FN double maxd(double a, double b) { return a < b ? b : a; }
DFD_FN double mind(double a, double b) { return b < a ? b : a; }
// a strictly smaller value: absorbs the rounding differences between a bound and the handler arithmetic it bounds
DFD_FN double lower(double x) { return x < DFD_TS_BIG ? x - x * 0x1p-40 : DFD_TS_BIG; }
DFD_FN int ctz64(unsigned long long v) {
#ifdef __CUDACC__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

DFD_FN bool key_less(double ts, unsigned src, unsigned seq, double ts2, unsigned src2, unsigned seq2) {
    if (ts != ts2) return ts < ts2;
    if (src != src2) return src < src2;
    return seq < seq2;
}
DFD_FN bool key_lt(const Key& a, const Key& b) { return key_less(a.ts, a.src, a.seq, b.ts, b.src, b.seq); }
DFD_FN Key mk_key(double ts, unsigned src, unsigned seq) {
    Key k;
    k.ts = ts;
    k.src = src;
    k.seq = seq;
    return k;
}
DFD_FN Key key_from_int4(int4 v) {
    Key k;
    k.ts = bits2d(((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x);
    k.src = (unsigned)v.z;
    k.seq = (unsigned)v.w;
    return k;
}

// warp-wide argmin of (key, idx): every lane contributes its local best, every lane receives the winner.
// Timestamps are non-negative doubles, so their bit patterns order like unsigned integers: two 32-bit redux.min
// steps settle almost every pick; equal timestamps fall through to sender gid, sequence number and index.
DFD_FN void warp_argmin_key(Key& k, int& idx) {
#ifdef __CUDACC__
    const unsigned full = 0xFFFFFFFFu;
    const unsigned long long tb = d2bits(k.ts);
    const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
    const unsigned mhi = __reduce_min_sync(full, hi);
    const bool in1 = hi == mhi;
    const unsigned mlo = __reduce_min_sync(full, in1 ? lo : 0xFFFFFFFFu);
    unsigned m = __ballot_sync(full, in1 && lo == mlo);
    if (m & (m - 1)) {  // several lanes hold the same timestamp
        bool in = (m >> DFD_LANE()) & 1u;
        const unsigned msrc = __reduce_min_sync(full, in ? k.src : 0xFFFFFFFFu);
        in = in && k.src == msrc;
        const unsigned mseq = __reduce_min_sync(full, in ? k.seq : 0xFFFFFFFFu);
        in = in && k.seq == mseq;
        const unsigned midx = __reduce_min_sync(full, in ? (unsigned)idx : 0xFFFFFFFFu);
        m = __ballot_sync(full, in && (unsigned)idx == midx);
    }
    const int win = __ffs((int)m) - 1;
    k.ts = __shfl_sync(full, k.ts, win);
    k.src = __shfl_sync(full, k.src, win);
    k.seq = __shfl_sync(full, k.seq, win);
    idx = __shfl_sync(full, idx, win);
#else
    (void)k;
    (void)idx;
#endif
}
// warp-wide minimum of non-negative doubles: their bit patterns order like unsigned integers, so two redux.min steps do it
DFD_FN double warp_min_d(double v) {
#ifdef __CUDACC__
    const unsigned full = 0xFFFFFFFFu;
    const unsigned long long tb = d2bits(v);
    const unsigned hi = (unsigned)(tb >> 32), lo = (unsigned)tb;
    const unsigned mhi = __reduce_min_sync(full, hi);
    const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xFFFFFFFFu);
    v = bits2d(((unsigned long long)mhi << 32) | mlo);
#endif
    return v;
}

struct Ctx {
    const DfdParams* P;
    const DfdArrays* A;
    const DfdPeers* X;
    unsigned* ev;  // per-block event counters by type
    // --- the group being activated
    unsigned r, rl;  // router: global id, local index
    unsigned gid;
    int lid, grp;
    RouterHdr* h;
    PortState* port;
    PortQ* pq;
    Key* skey;
    Key* ckey;
    Key* kkey;
    LaneCtl* lane;
    unsigned long long* seen_c;
    unsigned long long* seen_k;
    unsigned char* nq;
    unsigned long long* qbits;
    PoolSlot* pool;
    unsigned* pfree;
    ChunkRec* ring_c;
    Key* ring_k;
    double* iblk;   // [2][p]: bounds of the (empty) terminal lanes, set at the start of every router phase
    double rclock;  // lower bound on the router's next event
    bool nothing_pending;  // no event is pending anywhere in the group and no external lane has a finite bound
    double ev_min;         // earliest pending event of the group as of the last router-phase scan (events only, no bounds)
    double node_floor;     // earliest pending event of any terminal of the group (as of the last bounds pass)
    double node_push_min;  // earliest record the router handed to a terminal since then (lane 0)
    // --- per-lane accounting of this activation
    unsigned ncreated;  // events created
    unsigned nexec;     // events executed
    bool wrote_remote;  // pushed a record into another rank's memory
    // --- cycle accounting of the activations of this warp (lane 0): poll, terminal bounds, router phase, node phase, publish
    long long cyc[6];
    unsigned long long nrounds;
};
#ifdef DFD_DEBUG_HOST
#include <stdio.h>
#include <stdlib.h>
#endif
#ifdef __CUDACC__
#define CYC_NOW() clock64()
#else
#define CYC_NOW() 0LL
#endif

// First error wins.  Outlined (cold, dozens of call sites) and given only a global pointer, so that no kernel-parameter
// struct has to be materialised in local memory for it.  Peers learn about the error from this rank's monitor path
// (propagate_error), which every warp runs when it leaves the event loop.
DFD_FN_OUTLINE void set_error_at(DfdCtrl* ctrl, unsigned code, unsigned long long info) {
    if (cas_u64(&ctrl->error, 0ULL, (unsigned long long)code)) ctrl->error_info = info;
}
DFD_FN void set_error(const Ctx& c, unsigned code, unsigned long long info) { set_error_at(c.A->ctrl, code, info); }
DFD_FN void propagate_error(const DfdArrays& A, const DfdPeers& X) {
#ifdef __CUDACC__
    for (int k = 0; k < X.world; k++)
        if (k != X.rank) {
            DfdCtrl* pc = (DfdCtrl*)(X.base[k] + ((char*)A.ctrl - X.local_base));
            atomicCAS_system(&pc->error, 0ULL, (unsigned long long)DFD_ERR_PEER_ERROR);
        }
#else
    (void)A;
    (void)X;
#endif
}

// ------------------------------------------------------------------------------------------------
// CLCG4 (L'Ecuyer-Andres), int32 Schrage form; identical values to ROSS' rng_gen_val
// ------------------------------------------------------------------------------------------------
struct Clcg4Step {
    int32_t s0, s1, s2, s3;
    double u;
};
// one step of the four component generators; by value in and out so that the caller's state stays in registers
DFD_FN_OUTLINE Clcg4Step clcg4_step(int32_t s0, int32_t s1, int32_t s2, int32_t s3) {
    int k, v;
    double u = 0.0;
    v = s0;
    k = v / 46693;
    v = 45991 * (v - k * 46693) - k * 25884;
    if (v < 0) v = v + 2147483647;
    s0 = v;
    u = u + 4.65661287524579692e-10 * v;
    v = s1;
    k = v / 10339;
    v = 207707 * (v - k * 10339) - k * 870;
    if (v < 0) v = v + 2147483543;
    s1 = v;
    u = u - 4.65661310075985993e-10 * v;
    if (u < 0) u = u + 1.0;
    v = s2;
    k = v / 15499;
    v = 138556 * (v - k * 15499) - k * 3979;
    if (v < 0) v = v + 2147483423;
    s2 = v;
    u = u + 4.65661336096842131e-10 * v;
    if (u >= 1.0) u = u - 1.0;
    v = s3;
    k = v / 43218;
    v = 49689 * (v - k * 43218) - k * 24121;
    if (v < 0) v = v + 2147483323;
    s3 = v;
    u = u - 4.65661357780891134e-10 * v;
    if (u < 0) u = u + 1.0;
    Clcg4Step o;
    o.s0 = s0;
    o.s1 = s1;
    o.s2 = s2;
    o.s3 = s3;
    o.u = u;
    return o;
}
DFD_FN double clcg4_unif(int32_t* s) {
    Clcg4Step o = clcg4_step(s[0], s[1], s[2], s[3]);
    s[0] = o.s0;
    s[1] = o.s1;
    s[2] = o.s2;
    s[3] = o.s3;
    return o.u;
}
// tw_rand_integer: low + (long)(unif * (high + 1 - low)); no draw when high < low
DFD_FN long long clcg4_integer(int32_t* s, long long low, long long high, unsigned long long* draws) {
    if (high < low) return 0;
    (*draws)++;
    return low + (long long)(clcg4_unif(s) * (double)(high + 1 - low));
}

DFD_FN double bytes_to_ns(unsigned long long bytes, double GB_p_s) {  // dally:1588-1599
    double time = ((double)bytes) / (1024.0 * 1024.0 * 1024.0);
    time = time / GB_p_s;
    time = time * 1000.0 * 1000.0 * 1000.0;
    return time;
}
// exact n / d for n < 2^32 via the precomputed reciprocal (magic == 0 encodes d == 1)
DFD_FN unsigned fastdiv(unsigned n, unsigned long long magic) { return magic ? (unsigned)umul64hi((unsigned long long)n, magic) : n; }
// bytes_to_ns(bytes, bw) for bw in {cn, local, global}: the two sizes every chunk of this workload uses are precomputed
enum { BW_CN = 0, BW_LOCAL = 1, BW_GLOBAL = 2 };
DFD_FN double inj_delay(const DfdParams& P, unsigned bytes, int which) {
    if (bytes == (unsigned)P.chunk_size) return which == BW_CN ? P.inj_chunk_cn : which == BW_LOCAL ? P.inj_chunk_local : P.inj_chunk_global;
    if (bytes == (unsigned)P.payload_sz % (unsigned)P.chunk_size) return which == BW_CN ? P.inj_pay_cn : which == BW_LOCAL ? P.inj_pay_local : P.inj_pay_global;
    return bytes_to_ns(bytes, which == BW_CN ? P.cn_bw : which == BW_LOCAL ? P.local_bw : P.global_bw);
}
DFD_FN unsigned mod_chunk(unsigned x, unsigned chunk) { return x < chunk ? x : x % chunk; }

// gid arithmetic (codes_mapping.c:149-378 collapsed to closed forms)
DFD_FN unsigned svr_gid(const DfdParams& P, unsigned k) { return (k / P.p) * P.lps_per_rep + P.off_svr + k % P.p; }
DFD_FN unsigned term_gid(const DfdParams& P, unsigned k) { return (k / P.p) * P.lps_per_rep + P.off_term + k % P.p; }
DFD_FN unsigned rtr_gid(const DfdParams& P, unsigned r) { return r * P.lps_per_rep + P.off_rtr; }

// ------------------------------------------------------------------------------------------------
// lanes: geometry, channel words, group sharding
// ------------------------------------------------------------------------------------------------
DFD_FN unsigned lane_c_off(const DfdParams& P, int s) {
    if (s < P.intra_radix) return (unsigned)s * P.cap_l;
    if (s < P.ext) return (unsigned)P.intra_radix * P.cap_l + (unsigned)(s - P.intra_radix) * P.cap_g;
    return (unsigned)P.intra_radix * P.cap_l + (unsigned)P.global_radix * P.cap_g + (unsigned)(s - P.ext) * P.cap_tc;
}
DFD_FN unsigned lane_c_cap(const DfdParams& P, int s) { return s < P.intra_radix ? P.cap_l : s < P.ext ? P.cap_g : P.cap_tc; }
DFD_FN unsigned lane_k_off(const DfdParams& P, int q) {
    if (q < P.intra_radix) return (unsigned)q * P.cap_l;
    if (q < P.ext) return (unsigned)P.intra_radix * P.cap_l + (unsigned)(q - P.intra_radix) * P.cap_g;
    return (unsigned)P.intra_radix * P.cap_l + (unsigned)P.global_radix * P.cap_g + (unsigned)(q - P.ext) * P.cap_tk;
}
DFD_FN unsigned lane_k_cap(const DfdParams& P, int q) { return q < P.intra_radix ? P.cap_l : q < P.ext ? P.cap_g : P.cap_tk; }
// ring slot of record number `count` of chunk in-lane s / credit lane q: offset and capacity mask of every lane index come
// from a small table (the same for all routers) instead of three-way branches at every use
DFD_FN unsigned ring_c_idx(const DfdArrays& A, int s, unsigned count) {
    const uint2 g = A.lgeo[2 * s];
    return g.x + (count & g.y);
}
DFD_FN unsigned ring_k_idx(const DfdArrays& A, int q, unsigned count) {
    const uint2 g = A.lgeo[2 * q + 1];
    return g.x + (count & g.y);
}

DFD_FN unsigned long long cw_pack(double bound, unsigned count) {
    if (!(bound < DFD_TS_BIG)) bound = DFD_TS_BIG;
    return (d2bits(bound) & ~DFD_STAMP_MASK) | ((unsigned long long)count & DFD_STAMP_MASK);
}
DFD_FN double cw_bound(unsigned long long w) { return bits2d(w & ~DFD_STAMP_MASK); }
// record count of a lane from the stamp in its channel word, given a count `head` that is at most 4095 behind
DFD_FN uint16_t cw_tail(unsigned long long w, uint16_t head) { return (uint16_t)(head + (((unsigned)w - head) & (unsigned)DFD_STAMP_MASK)); }

DFD_FN int owner_of_group(const DfdPeers& X, int g) {
    int k = 0;
    while (g >= X.grp_lo[k + 1]) k++;
    return k;
}
template <typename T> DFD_FN T* on_rank(const DfdPeers& X, T* local, int owner) { return (T*)(X.base[owner] + ((char*)local - X.local_base)); }
// owner rank and local index of router d
DFD_FN int router_home(const Ctx& c, unsigned d, unsigned* dl) {
    const DfdPeers& X = *c.X;
    if (X.world == 1) {
        *dl = d;
        return X.rank;
    }
    int owner = owner_of_group(X, (int)fastdiv(d, c.P->magic_a));
    *dl = d - (unsigned)X.grp_lo[owner] * (unsigned)c.P->a;
    return owner;
}

// ------------------------------------------------------------------------------------------------
// event delivery
// ------------------------------------------------------------------------------------------------
// chunk leaving router port q towards router d
DFD_FN void push_chunk_to_router(Ctx& c, int q, unsigned d, const ChunkRec& rec) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (rec.ts >= P.ts_end) return;
    int s = q < P.intra_radix ? A.lout_lane[c.lid * P.intra_radix + q] : A.gout_lane[(size_t)c.r * P.h + (q - P.intra_radix)];
    unsigned dl;
    int owner = router_home(c, d, &dl);
    LaneCtl& L = c.lane[q];
    ChunkRec* slot = A.ring_c + (size_t)dl * P.cslots + ring_c_idx(A, s, L.sent_c);
    L.sent_c++;
    if (owner != c.X->rank) {
        slot = on_rank(*c.X, slot, owner);
        c.wrote_remote = true;
    }
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 0; i < 6; i++) st_cg16((int4*)slot + i, src[i]);
    c.ncreated++;
}
// credit for in-lane s of this router back to router u, whose port `port` fed it
DFD_FN void push_credit_to_router(Ctx& c, int s, unsigned u, unsigned port, double ts, unsigned seq, unsigned vc) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (ts >= P.ts_end) return;
    unsigned ul;
    int owner = router_home(c, u, &ul);
    LaneCtl& L = c.lane[s];
    Key* slot = A.ring_k + (size_t)ul * P.kslots + ring_k_idx(A, (int)port, L.sent_k);
    L.sent_k++;
    if (owner != c.X->rank) {
        slot = on_rank(*c.X, slot, owner);
        c.wrote_remote = true;
    }
    Key rec = mk_key(ts, c.gid | (vc << 30), seq);
    st_cg16(slot, *(const int4*)&rec);
    c.ncreated++;
}
// chunk / credit from the router to one of its terminals (same warp: plain rings with head / tail counters)
DFD_FN void push_chunk_to_node(Ctx& c, unsigned n, const ChunkRec& rec) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (rec.ts >= P.ts_end) return;
    unsigned nl = n - (unsigned)c.X->n0;
    unsigned idx = A.ncin_tail[nl]++;
    int4* dst = (int4*)&A.ncin[(size_t)nl * P.ncin_cap + (idx & (P.ncin_cap - 1))];
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 0; i < 6; i++) dst[i] = src[i];
    c.ncreated++;
    c.node_push_min = mind(c.node_push_min, rec.ts);
}
DFD_FN void push_credit_to_node(Ctx& c, unsigned n, double ts, unsigned src, unsigned seq) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (ts >= P.ts_end) return;
    unsigned nl = n - (unsigned)c.X->n0;
    unsigned idx = A.nkin_tail[nl]++;
    A.nkin[(size_t)nl * P.nkin_cap + (idx & (P.nkin_cap - 1))] = mk_key(ts, src, seq);
    c.ncreated++;
    c.node_push_min = mind(c.node_push_min, ts);
#ifdef DFD_DEBUG_HOST
    if (getenv("DFD_DBG_NODE") && (unsigned)atoi(getenv("DFD_DBG_NODE")) == n) fprintf(stderr, "  push credit to node %u ts=%.6f seq=%u (router clock %.6f)\n", n, ts, seq, c.rclock);
#endif
}
// chunk / credit from terminal t of this group to its router (terminal lane ext + t; same warp)
DFD_FN void node_push_chunk(Ctx& c, int t, const ChunkRec& rec) {
    const DfdParams& P = *c.P;
    if (rec.ts >= P.ts_end) return;
    int s = P.ext + t;
    LaneCtl& L = c.lane[s];
    ChunkRec* slot = c.ring_c + ring_c_idx(*c.A, s, L.ctail);
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 0; i < 6; i++) st_cg16((int4*)slot + i, src[i]);
    if (L.chead == L.ctail) c.ckey[s] = mk_key(rec.ts, rec.src, rec.seq);
    L.ctail++;
    c.ncreated++;
}
DFD_FN void node_push_credit(Ctx& c, int t, double ts, unsigned src, unsigned seq, unsigned vc) {
    const DfdParams& P = *c.P;
    if (ts >= P.ts_end) return;
    int s = P.ext + t;
    LaneCtl& L = c.lane[s];
    Key rec = mk_key(ts, src | (vc << 30), seq);
    st_cg16(c.ring_k + ring_k_idx(*c.A, s, L.ktail), *(const int4*)&rec);
    if (L.khead == L.ktail) c.kkey[s] = mk_key(ts, src, seq);
    L.ktail++;
    c.ncreated++;
}

// ================================================================================================
// node LP
// ================================================================================================
DFD_FN void node_add_event(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double ts, unsigned src, unsigned seq, unsigned type,
                           unsigned a = 0, unsigned b = 0, unsigned cc = 0, double d = 0.0, unsigned e = 0) {
    if (ts >= c.P->ts_end) return;
    if (s->nev >= (unsigned)c.P->nev_cap) {
        set_error(c, DFD_ERR_NEV_OVERFLOW, n);
        return;
    }
    NodeEv* ev = &list[s->nev++];
    ev->ts = ts;
    ev->src = src;
    ev->seq = seq;
    ev->type = type;
    ev->a = a;
    ev->b = b;
    ev->c = cc;
    ev->d = d;
    ev->e = e;
    if (s->nev > s->nev_hwm) s->nev_hwm = s->nev;
    c.ncreated++;
}

DFD_FN double node_cll(const DfdParams& P, NodeState* s, int32_t* rng) {  // codes/codes.h:61-66
    s->rng_draws++;
    return P.lookahead + 0.5 + clcg4_unif(rng) * 0.5;
}

// issue_event synthetic:231-271: next KICKOFF at now + mean_interval
DFD_FN void svr_issue_event(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    node_add_event(c, n, s, list, now + P.mean_interval, s->svr_gid, s->svr_seq++, EV_KICKOFF);
}

// handle_kickoff_event synthetic:329-413 + model_net_event_impl_base model-net.c:288-380
DFD_FN_NOINLINE void svr_kickoff(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    if (s->msg_sent_count >= P.num_msgs) return;
    long long local_dest = -1;
    if (P.traffic == DFD_TR_UNIFORM) {
        local_dest = clcg4_integer(s->rng_svr0, 1, (long long)P.num_nodes - 2, &s->rng_draws);
        local_dest = (long long)n + local_dest;  // (svr_id + d) % num_nodes with d < num_nodes
        if (local_dest >= P.num_nodes) local_dest -= P.num_nodes;
    } else if (P.traffic == DFD_TR_NEAREST_GROUP) {
        local_dest = (long long)n + P.nodes_per_grp;
        if (local_dest >= P.num_nodes) local_dest -= P.num_nodes;
    } else if (P.traffic == DFD_TR_NEAREST_NEIGHBOR) {
        local_dest = (long long)n + 1;
        if (local_dest >= P.num_nodes) local_dest -= P.num_nodes;
    } else if (P.traffic == DFD_TR_RAND_PERM) {
        if (s->dest_id == -1 || s->rperm_data > P.rperm_threshold) {
            s->dest_id = (int)clcg4_integer(s->rng_svr0, 0, (long long)P.num_nodes - 1, &s->rng_draws);
            local_dest = s->dest_id;
            s->rperm_data = P.payload_sz;
        } else {
            local_dest = s->dest_id;
            s->rperm_data += P.payload_sz;
        }
    } else {  // RANDOM_OTHER_GROUP synthetic:375-392
        int my_group_id = (int)s->grp;
        int sel = (int)clcg4_integer(s->rng_svr0, 0, (long long)P.G - 2, &s->rng_draws);
        int rand_group = sel >= my_group_id ? sel + 1 : sel;
        int rand_node = (int)clcg4_integer(s->rng_svr0, 0, (long long)P.nodes_per_grp - 1, &s->rng_draws);
        local_dest = (long long)rand_group * P.nodes_per_grp + rand_node;
    }
    if (now >= P.warm_up_time) s->warm_msg_sent_count++;
    s->msg_sent_count++;
    s->last_send_ts = now;
    unsigned my_svr = s->svr_gid;
    // model_net_event(net_id, "test", global_dest, PAYLOAD_SZ, 0.0, remote, local, lp)
    if ((unsigned)local_dest == n && P.payload_sz < P.node_eager_limit && !P.noop_bypass) {
        // model_net_noop_event model-net.c:252-286 (same terminal on both ends)
        double poffset = 0.0;
        double delay = node_cll(P, s, s->rng_svr2);
        double sendTime = (double)(unsigned long long)P.payload_sz * P.codes_cn_delay;
        poffset += delay;
        node_add_event(c, n, s, list, now + (poffset + 0.0 + sendTime), my_svr, s->svr_seq++, EV_LOCAL);
        poffset += delay;
        node_add_event(c, n, s, list, now + (poffset + 0.0 + sendTime), my_svr, s->svr_seq++, EV_REMOTE, n, 0, 0, now);
    } else {
        double poffset = node_cll(P, s, s->rng_svr2);
        node_add_event(c, n, s, list, now + (poffset + 0.0), my_svr, s->svr_seq++, EV_NEW_MSG | NEV_FLAG_QREQ, (unsigned)local_dest,
                       (unsigned)P.payload_sz, 0, now);
    }
    svr_issue_event(c, n, s, list, now);
}

// handle_remote_event synthetic:468-498; the ACK (handle_ack_event :415-432) only updates commutative
// statistics of the source svr, so it is applied with atomics instead of travelling as an event.
DFD_FN_NOINLINE void svr_remote(Ctx& c, unsigned n, NodeState* s, double now, unsigned src_svr, double msg_start_time) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const DfdPeers& X = *c.X;
    if (now >= P.warm_up_time) {
        s->msg_recvd_count++;
        double packet_latency = now - msg_start_time;
        s->sum_server_latency += packet_latency;
        if (packet_latency > s->max_server_latency) s->max_server_latency = packet_latency;
        s->svr_seq++;  // tw_event_new for the ACK
        if (now < P.ts_end) {
            int owner = X.rank;
            unsigned sl = src_svr;
            if (X.world > 1) {
                owner = owner_of_group(X, (int)fastdiv(fastdiv(src_svr, P.magic_p), P.magic_a));
                sl = src_svr - (unsigned)X.grp_lo[owner] * (unsigned)P.a * (unsigned)P.p;
            } else
                sl = src_svr - (unsigned)X.n0;
            if (owner != X.rank)
                ack_update(on_rank(X, &A.ack_count[sl], owner), on_rank(X, &A.ack_first[sl], owner), on_rank(X, &A.ack_last[sl], owner), d2bits(now), true);
            else
                ack_update(&A.ack_count[sl], &A.ack_first[sl], &A.ack_last[sl], d2bits(now), false);
            count_ev(c.ev, EV_ACK);
        }
    }
    (void)n;
}

// dragonfly_dally_packet_event dally:5122-5183 via fcfs_next model-net-sched-impl.c:187-261
DFD_FN int mn_fcfs_next(Ctx& c, unsigned n, unsigned nl, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (s->sq_count == 0) return -1;
    SchedReq* q = &A.sq[(size_t)nl * P.nsq_cap + (s->sq_head & (P.nsq_cap - 1))];
    int is_last_packet;
    unsigned psize;
    if (P.mn_packet_size >= q->rem) {
        psize = q->rem;
        is_last_packet = 1;
    } else {
        psize = P.mn_packet_size;
        is_last_packet = 0;
    }
    node_add_event(c, n, s, list, now + 0.0, s->term_gid, s->term_seq++, EV_T_GENERATE | (is_last_packet ? NEV_FLAG_LAST : 0), q->dst_svr,
                   q->msg_size, psize, q->msg_start_time, q->msg_id);
    if (is_last_packet) {
        s->sq_head++;
        s->sq_count--;
        return 1;
    }
    q->rem -= psize;
    return 0;
}
DFD_FN void mn_sched_next(Ctx& c, unsigned n, unsigned nl, NodeState* s, NodeEv* list, double now) {  // model-net-lp.c:838-873
    if (mn_fcfs_next(c, n, nl, s, list, now) == -1) s->in_sched_send_loop = 0;
}

// handle_new_msg model-net-lp.c:643-802
DFD_FN_NOINLINE void mn_new_msg(Ctx& c, unsigned n, unsigned nl, NodeState* s, NodeEv* list, double now, unsigned type, unsigned dst, unsigned msg_size,
                                double msg_start_time) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (dst == n) {
        // lp->gid == dest_mn_lp: a message to the sender's own terminal that did not take the no-op shortcut (size >=
        // node_eager_limit, or codes_noop_bypass): the base LP copies it inside the node, model-net-lp.c:672-704 (one
        // node-copy queue); remote event to the destination svr, then the self event to the sender, both from this LP
        double exp_time = (s->node_copy_next_available_time > now) ? s->node_copy_next_available_time : now;
        exp_time += (double)(unsigned long long)msg_size * P.codes_cn_delay;
        exp_time -= now;
        double delay = node_cll(P, s, s->rng_term2);
        s->node_copy_next_available_time = now + exp_time;
        exp_time += delay;
        node_add_event(c, n, s, list, now + exp_time, s->term_gid, s->term_seq++, EV_REMOTE, n, 0, 0, msg_start_time);
        exp_time += delay;
        node_add_event(c, n, s, list, now + exp_time, s->term_gid, s->term_seq++, EV_LOCAL);
        return;
    }
    if (type & NEV_FLAG_QREQ) {
        // NIC sequencing: the message is re-queued at next_available_time, which only grows -> a FIFO of its own
        double exp_time = (s->next_available_time > now) ? s->next_available_time : now;
        exp_time += P.nic_seq_delay + node_cll(P, s, s->rng_term2);
        s->next_available_time = exp_time;
        double ts = now + (exp_time - now);
        unsigned seq = s->term_seq++;
        if (ts >= P.ts_end) return;
        if (s->mq_count >= (unsigned)P.nsq_cap) {
            set_error(c, DFD_ERR_SQ_OVERFLOW, n);
            return;
        }
        MsgReq* m = &A.mq[(size_t)nl * P.nsq_cap + ((s->mq_head + s->mq_count) & (P.nsq_cap - 1))];
        m->ts = ts;
        m->seq = seq;
        m->dst_svr = dst;
        m->msg_size = msg_size;
        m->pad = 0;
        m->msg_start_time = msg_start_time;
        s->mq_count++;
        c.ncreated++;
        return;
    }
    if (s->sq_count >= (unsigned)P.nsq_cap) {
        set_error(c, DFD_ERR_SQ_OVERFLOW, n);
        return;
    }
    SchedReq* q = &A.sq[(size_t)nl * P.nsq_cap + ((s->sq_head + s->sq_count) & (P.nsq_cap - 1))];
    q->dst_svr = dst;
    q->msg_size = msg_size;
    q->rem = msg_size;
    q->msg_id = s->msg_id++;
    q->msg_start_time = msg_start_time;
    s->sq_count++;
    if (s->sq_count > s->sq_hwm) s->sq_hwm = s->sq_count;
    if (s->in_sched_send_loop == 0) {
        s->in_sched_send_loop = 1;
        mn_sched_next(c, n, nl, s, list, now);
    }
}

DFD_FN unsigned num_chunks_for(unsigned sz, unsigned chunk) { return sz <= chunk ? 1u : (sz + chunk - 1) / chunk; }  // dally:103-104

// packet_generate dally:5417-5774
DFD_FN_NOINLINE void term_packet_generate(Ctx& c, unsigned n, unsigned nl, NodeState* s, NodeEv* list, double now, const NodeEv& ev) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    unsigned dst_svr = ev.a, total_size = ev.b, packet_size = ev.c, message_id = ev.e;
    unsigned dst_term = dst_svr;  // one svr per terminal
    s->packet_gen++;
    s->total_gen_size += packet_size;
    unsigned nchunks = num_chunks_for(packet_size, P.chunk_size);
    int my_router = (int)s->router;
    int dest_router_id = (int)fastdiv(dst_term, P.magic_p);
    int dest_grp_id = (int)fastdiv((unsigned)dest_router_id, P.magic_a);
    int src_grp_id = (int)s->grp;
    if (src_grp_id == dest_grp_id) {
        if (dest_router_id == my_router)
            s->local_sr++;
        else
            s->local_sg++;
    } else
        s->remote_pk++;
    unsigned packet_id = s->packet_counter++;
    unsigned flags = (ev.type & NEV_FLAG_LAST) ? 3u : 0u;  // remote + local event ride on the last packet
    for (unsigned i = 0; i < nchunks; i++) {
        if (s->tq_count >= (unsigned)P.ntq_cap) {
            set_error(c, DFD_ERR_TQ_OVERFLOW, n);
            return;
        }
        TermChunk* t = &A.tq[(size_t)nl * P.ntq_cap + ((s->tq_head + s->tq_count) & (P.ntq_cap - 1))];
        t->dst_svr = dst_svr;
        t->packet_id = packet_id;
        t->packet_size = packet_size;
        t->total_size = total_size;
        t->message_id = message_id;
        t->chunk_id = i;
        t->flags = flags | ((unsigned)dest_grp_id << 16);
        t->dst_router = (unsigned)dest_router_id;
        t->msg_start_time = ev.d;
        s->tq_count++;
        s->terminal_length += P.chunk_size;
    }
    if (s->tq_count > s->tq_hwm) s->tq_hwm = s->tq_count;
    double injection_ts = packet_size == (unsigned)P.payload_sz ? P.inj_packet_cn : bytes_to_ns(packet_size, P.cn_bw);
    double nic_ts = injection_ts;
    unsigned tg = s->term_gid;
    if (s->terminal_length < P.cn_vc) {
        node_add_event(c, n, s, list, now + nic_ts, tg, s->term_seq++, EV_SCHED_NEXT);  // model_net_method_idle_event2
    } else {
        s->issueIdle = 1;
        if (s->last_buf_full == 0.0) s->last_buf_full = now;
    }
    if (s->in_send_loop == 0) {
        node_add_event(c, n, s, list, now + 0.0, tg, s->term_seq++, EV_T_SEND);
        s->in_send_loop = 1;
    }
    long long total_event_size = 416 + ((flags & 1) ? 72 : 0) + ((flags & 2) ? 72 : 0);  // model-net.c:500-504, sizeof(svr_msg)=72
    s->stats_used = 1;
    s->send_count++;
    s->send_bytes += packet_size;
    s->send_time += P.inv_cn_bw * packet_size;
    if (s->max_event_size < total_event_size) s->max_event_size = total_event_size;
}

// packet_send dally:5836-6026 (+ get_next_vcg :3440-3449 for one QoS level)
DFD_FN_NOINLINE void term_packet_send(Ctx& c, unsigned n, unsigned nl, int t, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (s->tq_count == 0 || s->vc_occupancy + P.chunk_size > P.cn_vc) {
        s->in_send_loop = 0;
        s->stalled_chunks++;
        if (!s->last_buf_full) s->last_buf_full = now;
        return;
    }
    TermChunk cur = A.tq[(size_t)nl * P.ntq_cap + (s->tq_head & (P.ntq_cap - 1))];
    unsigned nchunks = num_chunks_for(cur.packet_size, P.chunk_size);
    double injection_delay = P.inj_chunk_cn;
    if ((cur.packet_size < (unsigned)P.chunk_size) && (cur.chunk_id == nchunks - 1)) {
        unsigned data_size = mod_chunk(cur.packet_size, P.chunk_size);
        injection_delay = inj_delay(P, data_size, BW_CN);
    }
    double propagation_delay = P.cn_delay;
    s->terminal_available_time = maxd(s->terminal_available_time, now);
    s->terminal_available_time += injection_delay;
    double injection_ts = s->terminal_available_time - now;
    double propagation_ts = injection_ts + propagation_delay;
    unsigned tg = s->term_gid;
    {
        ChunkRec rec;
        rec.ts = now + propagation_ts;
        rec.src = tg;
        rec.seq = s->term_seq++;
        rec.packet_id = cur.packet_id;
        rec.src_term = n;
        rec.dst_term = cur.dst_svr;
        rec.message_id = cur.message_id;
        rec.packet_size = cur.packet_size;
        rec.total_size = cur.total_size;
        rec.chunk_id = (uint16_t)cur.chunk_id;
        rec.dst_grp = (uint16_t)(cur.flags >> 16);
        rec.intm_grp_id = -1;
        rec.org_grp = (uint16_t)s->grp;
        rec.prev_lp = n;
        rec.prev_port = 0;  // vc_index = vcg = 0
        rec.prev_vc = 0;    // output_chan = vcg
        rec.last_hop = DFD_HOP_TERMINAL;
        rec.path_type = -1;
        rec.is_intm_visited = 0;
        rec.n_hop = rec.l_hop = rec.g_hop = rec.hops_cur_group = 0;
        rec.flags = (uint8_t)(cur.flags & 1u);  // local_event_size_bytes = 0 on the wire
        rec.pad0 = 0;
        rec.travel_start_time = now;
        rec.msg_start_time = cur.msg_start_time;
        rec.origin_router = s->router;
        rec.src_svr = n;
        rec.dst_svr = cur.dst_svr;
        rec.dst_router = cur.dst_router;
        node_push_chunk(c, t, rec);
    }
    if (cur.chunk_id == nchunks - 1 && (cur.flags & 2u)) {  // LOCAL completion event to the sender svr at +0
        node_add_event(c, n, s, list, now + 0.0, tg, s->term_seq++, EV_LOCAL);
    }
    s->vc_occupancy += P.chunk_size;
    s->tq_head++;
    s->tq_count--;
    s->terminal_length -= P.chunk_size;
    s->link_traffic += P.chunk_size;
    s->total_chunks++;
    s->injected_chunks++;
    if (s->tq_count != 0 && s->vc_occupancy + P.chunk_size <= P.cn_vc) {
        node_add_event(c, n, s, list, now + injection_ts, tg, s->term_seq++, EV_T_SEND);
    } else {
        s->in_send_loop = 0;
    }
    if (s->issueIdle && s->terminal_length < P.cn_vc) {
        s->issueIdle = 0;
        node_add_event(c, n, s, list, now + injection_ts, tg, s->term_seq++, EV_SCHED_NEXT);
        if (s->last_buf_full > 0.0) {
            s->busy_time += (now - s->last_buf_full);
            s->last_buf_full = 0.0;
        }
    }
}

// terminal_buf_update dally:6762-6789
DFD_FN void term_buf_update(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    s->vc_occupancy -= P.chunk_size;
    if (s->in_send_loop == 0 && s->tq_count != 0) {
        node_add_event(c, n, s, list, now + 0.0, s->term_gid, s->term_seq++, EV_T_SEND);
        s->in_send_loop = 1;
    }
}

// packet_arrive dally:6453-6744.  Packets of one chunk that are a whole message (the BASELINE workloads) never touch
// the reassembly tables: their rank_tbl entry would be created and retired inside this one call.
DFD_FN_NOINLINE void term_packet_arrive(Ctx& c, unsigned n, unsigned nl, int t, NodeState* s, NodeEv* list, double now, const ChunkRec& m) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (m.dst_term != n) {
        set_error(c, DFD_ERR_WRONG_TERMINAL, n);
        return;
    }
    unsigned tg = s->term_gid;
    // credit to the last router (dally:6483-6495): always the router this terminal hangs off, over its terminal port
    node_push_credit(c, t, now + P.cn_credit, tg, s->term_seq++, m.prev_vc);
    s->finished_chunks++;
    s->ejected_chunks++;
    if (m.path_type == DFD_MINIMAL) s->minimal_count++;
    if (m.path_type == DFD_NON_MINIMAL) s->nonmin_count++;
    double ete_latency = now - m.travel_start_time;
    s->total_time += ete_latency;
    s->total_hops += m.n_hop;
    s->stats_used = 1;
    s->recv_time += ete_latency;
    unsigned nchunks = num_chunks_for(m.packet_size, P.chunk_size);
    if (m.chunk_id == nchunks - 1) {
        s->packet_fin++;
        s->recv_count++;
        s->recv_bytes += m.packet_size;
        s->finished_packets++;
    }
    if (s->min_latency > ete_latency) s->min_latency = ete_latency;
    if (s->max_latency < ete_latency) s->max_latency = ete_latency;
    bool message_done;
    bool has_remote = (m.flags & 1u) != 0;
    if (m.packet_size <= (unsigned)P.chunk_size && m.total_size <= (unsigned)P.packet_size) {
        message_done = true;  // one chunk == one packet == the whole message
    } else {
        // remaining_sz_packets (dally:6590-6618): bytes still missing of a multi-chunk packet
        bool packet_done = false;
        uint4* pt = &A.ptab[(size_t)nl * P.nreasm_cap];
        unsigned pi = 0;
        for (; pi < s->pt_n; pi++)
            if (pt[pi].x == m.packet_id && pt[pi].y == m.src_term) break;
        if (pi < s->pt_n) {
            unsigned actual = (unsigned)P.chunk_size < pt[pi].z ? (unsigned)P.chunk_size : pt[pi].z;
            pt[pi].z -= actual;
            if (pt[pi].z == 0) {
                packet_done = true;
                pt[pi] = pt[--s->pt_n];
            }
        } else if ((unsigned)P.chunk_size < m.packet_size) {
            if (s->pt_n >= (unsigned)P.nreasm_cap) {
                set_error(c, DFD_ERR_REASM_OVERFLOW, n);
                return;
            }
            uint4 e4;
            e4.x = m.packet_id;
            e4.y = m.src_term;
            e4.z = m.packet_size - P.chunk_size;
            e4.w = 0;
            pt[s->pt_n++] = e4;
        } else
            packet_done = true;
        // rank_tbl (dally:6620-6720): packets still missing of the message, keyed by (message_id, sender)
        uint4* mt = &A.mtab[(size_t)nl * P.nreasm_cap];
        unsigned mi = 0;
        for (; mi < s->mt_n; mi++)
            if (mt[mi].x == m.message_id && mt[mi].y == m.src_svr) break;
        if (mi == s->mt_n) {
            if (s->mt_n >= (unsigned)P.nreasm_cap) {
                set_error(c, DFD_ERR_REASM_OVERFLOW, n);
                return;
            }
            unsigned total_packets = m.total_size / (unsigned)P.packet_size + (m.total_size % (unsigned)P.packet_size ? 1 : 0);
            if (total_packets == 0) total_packets = 1;
            uint4 e4;
            e4.x = m.message_id;
            e4.y = m.src_svr;
            e4.z = total_packets;
            e4.w = 0;
            mt[s->mt_n++] = e4;
        }
        if (has_remote) mt[mi].w = 1;  // the last packet's chunks carry the remote event
        if (packet_done) mt[mi].z--;
        message_done = mt[mi].z == 0;
        if (message_done) {
            has_remote = mt[mi].w != 0;
            mt[mi] = mt[--s->mt_n];
        }
    }
    if (message_done) {
        s->total_msg_size += m.total_size;
        s->finished_msgs++;
        if (has_remote)  // send_remote_event dally:6124-6151: REMOTE to the destination svr at +0
            node_add_event(c, n, s, list, now + 0.0, tg, s->term_seq++, EV_REMOTE, m.src_svr, 0, 0, m.msg_start_time);
    }
}

// earliest timestamp among the node's self-scheduled events (+inf when there is none)
DFD_FN double node_self_next(const DfdParams& P, const DfdArrays& A, unsigned nl, const NodeState* s) {
    double m = DINF;
    const NodeEv* list = &A.nev[(size_t)nl * P.nev_cap];
    for (unsigned i = 0; i < s->nev; i++) m = mind(m, list[i].ts);
    if (s->mq_count) m = mind(m, A.mq[(size_t)nl * P.nsq_cap + (s->mq_head & (P.nsq_cap - 1))].ts);
    return m;
}

// earliest R_BUFFER that could promote an overflowed chunk of in-lane s: min over the ports holding such chunks of
// that port's next credit (head of its credit lane, or the lane's bound while it is empty)
DFD_FN double late_credit_bound(const Ctx& c, int s) {
    double m = DFD_TS_BIG;
    for (int w = 0; w < 2; w++) {
        unsigned long long bits = c.qbits[2 * s + w];
        while (bits) {
            int b = ctz64(bits);
            bits &= bits - 1;
            m = mind(m, c.kkey[64 * w + b].ts);
        }
    }
    return m;
}

// Bounds on the four lanes between terminal t and its router, from the state both currently hold:
//   rb: no chunk will arrive at the terminal before rb        kb: no credit will arrive at the terminal before kb
//   x : no chunk of the terminal will arrive at the router before x    y: no credit of the terminal will ... before y
struct NodeBounds {
    double rb, kb, x, y;
};
DFD_FN NodeBounds node_bounds(const Ctx& c, int t, unsigned nl, const NodeState* ns) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    NodeBounds b;
    const int s = P.ext + t;
    const PortState& ps = c.port[s];
    // router_packet_send towards the terminal: noat = max(noat, now) + (inj + router_delay); arrival = noat + cn_delay
    double st = ps.in_send_loop ? c.skey[s].ts : c.rclock;
    b.rb = lower((maxd(ps.noat, st) + P.lb_injrd_cn) + P.cn_delay);
    const double S = ns->self_next;
    const LaneCtl& L = c.lane[s];
    // Credits already waiting in the terminal's ring, and the next self event: what can start a T_SEND before any
    // further credit arrives (terminal_buf_update only restarts the send loop while chunks are waiting).
    const unsigned ktail = A.nkin_tail[nl];
    const bool kr_ne = ns->kin_head != ktail;
    const double khead = kr_ne ? A.nkin[(size_t)nl * P.nkin_cap + (ns->kin_head & (P.nkin_cap - 1))].ts : DFD_TS_BIG;
    const double trig0 = ns->tq_count > 0 ? mind(S, khead) : S;
    // next R_ARRIVE of one of this terminal's chunks, ignoring T_SENDs triggered by credits that are not in the ring yet:
    // such a credit cannot come before kb itself ("credit bound of a terminal lane")
    double a_in = L.chead != L.ctail ? c.ckey[s].ts : lower((maxd(ns->terminal_available_time, trig0) + P.lb_inj_term) + P.cn_delay);
    b.kb = P.cn_credit + mind(a_in, late_credit_bound(c, s));
    double knext = kr_ne ? khead : b.kb;
    // a T_SEND needs a self event (T_GENERATE, a re-armed T_SEND) or, with chunks waiting, a credit (terminal_buf_update)
    double trig = ns->tq_count > 0 ? mind(S, knext) : S;
    b.x = lower((maxd(ns->terminal_available_time, trig) + P.lb_inj_term) + P.cn_delay);
    unsigned ctail = A.ncin_tail[nl];
    double tarr = ns->cin_head != ctail ? A.ncin[(size_t)nl * P.ncin_cap + (ns->cin_head & (P.ncin_cap - 1))].ts : b.rb;
    b.y = tarr + P.cn_credit;
    return b;
}

// Causality check.  An LP's timestamps never go back, and among the events that reach it from OTHER LPs (lane
// records) the keys never go back either.  A self event scheduled at +0 may carry a smaller key than the event that
// created it (the sequential engine executes it afterwards all the same), so self events are only checked for time.
// last_src / last_seq hold the key of the last lane record executed at timestamp last_ts (0, 0: none yet).
DFD_FN void check_order(Ctx& c, double* last_ts, uint32_t* last_src, uint32_t* last_seq, double ts, unsigned src, unsigned seq, bool external,
                        unsigned long long who) {
    bool late = ts < *last_ts || (external && ts == *last_ts && (src < *last_src || (src == *last_src && seq < *last_seq)));
    if (late) {
#ifdef DFD_DEBUG_HOST
        fprintf(stderr, "LATE who=%llx ev=(%.6f,%u,%u,%d) last=(%.6f,%u,%u) rclock=%.6f\n", who, ts, src, seq, (int)external, *last_ts, *last_src, *last_seq, c.rclock);
#endif
        set_error(c, DFD_ERR_LATE_EVENT, who);
    }
    if (ts > *last_ts) {
        *last_ts = ts;
        *last_src = 0;
        *last_seq = 0;
    }
    if (external) {
        *last_src = src;
        *last_seq = seq;
    }
}

// executes the events of terminal t (node n) that are safe: earlier than rb while its chunk ring is empty and
// earlier than kb while its credit ring is empty
DFD_FN_NOINLINE unsigned node_process(Ctx& c, int t, unsigned n, unsigned nl, double rb, double kb) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    NodeState* s = &A.node[nl];
    NodeEv* list = &A.nev[(size_t)nl * P.nev_cap];
    const unsigned cmask = P.ncin_cap - 1, kmask = P.nkin_cap - 1;
    const unsigned ctail = A.ncin_tail[nl], ktail = A.nkin_tail[nl];
    unsigned nproc = 0;
    for (;;) {
        int bi = -1;
        double bts = DINF;
        unsigned bsrc = 0, bseq = 0;
        int kind = 0;  // 0 self, 1 chunk ring, 2 credit ring, 3 second-stage NEW_MSG
        for (unsigned i = 0; i < s->nev; i++) {
            const NodeEv& e = list[i];
            if (bi < 0 || key_less(e.ts, e.src, e.seq, bts, bsrc, bseq)) {
                bi = (int)i;
                bts = e.ts;
                bsrc = e.src;
                bseq = e.seq;
            }
        }
        const MsgReq* mq = nullptr;
        if (s->mq_count) {
            mq = &A.mq[(size_t)nl * P.nsq_cap + (s->mq_head & (P.nsq_cap - 1))];
            if (bi < 0 || key_less(mq->ts, s->term_gid, mq->seq, bts, bsrc, bseq)) {
                kind = 3;
                bts = mq->ts;
                bsrc = s->term_gid;
                bseq = mq->seq;
            }
        }
        s->self_next = bts;  // exact whenever the loop is left: nothing is added to the lists outside this function
        double horizon = DFD_TS_BIG;
        const ChunkRec* ch = nullptr;
        if (s->cin_head != ctail) {
            ch = &A.ncin[(size_t)nl * P.ncin_cap + (s->cin_head & cmask)];
            if (d2bits(bts) == INF_BITS || key_less(ch->ts, ch->src, ch->seq, bts, bsrc, bseq)) {
                kind = 1;
                bts = ch->ts;
                bsrc = ch->src;
                bseq = ch->seq;
            }
        } else
            horizon = rb;
        const Key* kr = nullptr;
        if (s->kin_head != ktail) {
            kr = &A.nkin[(size_t)nl * P.nkin_cap + (s->kin_head & kmask)];
            if (d2bits(bts) == INF_BITS || key_less(kr->ts, kr->src, kr->seq, bts, bsrc, bseq)) {
                kind = 2;
                bts = kr->ts;
                bsrc = kr->src;
                bseq = kr->seq;
            }
        } else
            horizon = mind(horizon, kb);
        if (!(bts < horizon)) break;
        double now = bts;
#ifdef DFD_DEBUG_HOST
        if (getenv("DFD_DBG_NODE") && (unsigned)atoi(getenv("DFD_DBG_NODE")) == n)
            fprintf(stderr, "  node %u exec kind=%d key=(%.6f,%u,%u) rb=%.6f kb=%.6f horizon=%.6f rclock=%.6f tq=%u vc=%d\n", n, kind, bts, bsrc, bseq, rb, kb, horizon, c.rclock, s->tq_count, s->vc_occupancy);
#endif
        check_order(c, &s->last_ts, &s->last_src, &s->last_seq, bts, bsrc, bseq, kind == 1 || kind == 2, 0x8000000000000000ULL | n);
        s->events++;
        nproc++;
        if (kind == 1) {
            ChunkRec m = *ch;
            s->cin_head++;
            count_ev(c.ev, EV_T_ARRIVE);
            term_packet_arrive(c, n, nl, t, s, list, now, m);
        } else if (kind == 2) {
            s->kin_head++;
            count_ev(c.ev, EV_T_BUFFER);
            term_buf_update(c, n, s, list, now);
        } else if (kind == 3) {
            MsgReq m = *mq;
            s->mq_head++;
            s->mq_count--;
            count_ev(c.ev, EV_NEW_MSG);
            mn_new_msg(c, n, nl, s, list, now, EV_NEW_MSG, m.dst_svr, m.msg_size, m.msg_start_time);
        } else {
            NodeEv ev = list[bi];
            list[bi] = list[s->nev - 1];
            s->nev--;
            unsigned type = ev.type & 0xFF;
            count_ev(c.ev, (int)type);
            switch (type) {
            case EV_KICKOFF: svr_kickoff(c, n, s, list, now); break;
            case EV_REMOTE: svr_remote(c, n, s, now, ev.a, ev.d); break;
            case EV_LOCAL:  // handle_local_event synthetic:508-516
                if (now >= P.warm_up_time) s->local_recvd_count++;
                break;
            case EV_NEW_MSG: mn_new_msg(c, n, nl, s, list, now, ev.type, ev.a, ev.b, ev.d); break;
            case EV_SCHED_NEXT: mn_sched_next(c, n, nl, s, list, now); break;
            case EV_T_GENERATE: term_packet_generate(c, n, nl, s, list, now, ev); break;
            case EV_T_SEND: term_packet_send(c, n, nl, t, s, list, now); break;
            default: break;
            }
        }
    }
    return nproc;
}

// ================================================================================================
// router LP
// ================================================================================================
// --- connection queries (DragonflyConnectionManager, dragonfly-network-manager.cxx) ----------------
// local ports from my local id to router-local id `lid` (get_connections_to_gid(.., CONN_LOCAL)): file order
DFD_FN void loc_range(const DfdParams& P, const DfdArrays& A, int mylid, int lid, int* lo, int* hi) {
    *lo = A.loc_off[mylid * P.a + lid];
    *hi = A.loc_off[mylid * P.a + lid + 1];
}
// direct global connections to group g (get_connections_to_group): contiguous range of the by-dest-gid list
// (the by-gid list is sorted by destination group, unused slots hold INT_MAX at the end: lower bound by bisection)
DFD_FN_OUTLINE int grp_range_find(const int* gg, int h, int g) {
    int l = 0, n = h;
    while (n > 0) {
        int half = n >> 1;
        if (gg[l + half] < g) {
            l += half + 1;
            n -= half + 1;
        } else
            n = half;
    }
    int u = l;
    while (u < h && gg[u] == g) u++;
    return l | (u << 16);
}
DFD_FN void grp_range(const DfdParams& P, const DfdArrays& A, unsigned r, int g, int* lo, int* hi) {
    int v = grp_range_find(A.gs_group + (size_t)r * P.h, P.h, g);
    *lo = v & 0xFFFF;
    *hi = v >> 16;
}

// a candidate list of next-hop connections, never materialised:
//  kind 0: empty; 1: direct global links [lo,hi) of the by-gid list; 2: "routed" list towards group g
//  (for each global link my group -> g in inter-file order: all my local links to its owner,
//  get_routed_connections_to_group(g, true) network-manager.cxx:772-814); 3: local links to lid lo
struct Cands {
    int kind, lo, hi, n;
};
DFD_FN int routed_count(const DfdParams& P, const DfdArrays& A, const Ctx& R, int g) {
    int o0 = A.routed_off[R.grp * P.G + g], o1 = A.routed_off[R.grp * P.G + g + 1];
    int n = 0;
    for (int i = o0; i < o1; i++) {
        int lid = A.routed_lid[i];
        n += A.loc_off[R.lid * P.a + lid + 1] - A.loc_off[R.lid * P.a + lid];
    }
    return n;
}
DFD_FN int cand_port(const DfdParams& P, const DfdArrays& A, const Ctx& R, const Cands& cd, int idx) {
    if (cd.kind == 1) return A.gs_port[(size_t)R.r * P.h + cd.lo + idx];
    if (cd.kind == 3) return A.loc_port[cd.lo + idx];
    // kind 2
    int o0 = A.routed_off[R.grp * P.G + cd.lo], o1 = A.routed_off[R.grp * P.G + cd.lo + 1];
    for (int i = o0; i < o1; i++) {
        int lid = A.routed_lid[i];
        int l0 = A.loc_off[R.lid * P.a + lid], l1 = A.loc_off[R.lid * P.a + lid + 1];
        if (idx < l1 - l0) return A.loc_port[l0 + idx];
        idx -= l1 - l0;
    }
    return -1;
}
// get_legal_minimal_stops dally:9793-9849
DFD_FN Cands legal_minimal_stops(const DfdParams& P, const DfdArrays& A, const Ctx& R, int fdest_router, int fg) {
    Cands cd;
    if (R.grp != fg) {
        grp_range(P, A, R.r, fg, &cd.lo, &cd.hi);
        if (cd.hi > cd.lo) {
            cd.kind = 1;
            cd.n = cd.hi - cd.lo;
            return cd;
        }
        cd.kind = 2;
        cd.lo = fg;
        cd.hi = 0;
        cd.n = routed_count(P, A, R, fg);
        return cd;
    }
    int lo, hi;
    loc_range(P, A, R.lid, fdest_router - fg * P.a, &lo, &hi);
    cd.kind = 3;
    cd.lo = lo;
    cd.hi = hi;
    cd.n = hi - lo;
    return cd;
}

DFD_FN int score_port(const DfdParams& P, const Ctx& R, int port, bool nonmin) {  // dally:1649-1684
    if (port < 0) return 2147483647;
    const PortState& ps = R.port[port];
    int score = ps.vc_occ[0] + ps.vc_occ[1] + ps.vc_occ[2] + ps.vc_occ[3] + ps.queued_count;
    if (P.scoring == DFD_SCORE_DELTA && nonmin) score = score * 2;
    return score;
}
DFD_FN long long r_integer(Ctx& R, long long lo, long long hi) { return clcg4_integer(R.h->rng, lo, hi, &R.h->rng_draws); }

// get_absolute_best_connection_from_conns dally:1687-1723 over a candidate list
DFD_FN int best_from_cands(const DfdParams& P, const DfdArrays& A, Ctx& R, const Cands& cd) {
    if (cd.n == 0) return -1;
    if (cd.n == 1) return cand_port(P, A, R, cd, 0);
    int best_score = 2147483647, nbest = 0;
    for (int i = 0; i < cd.n; i++) {
        int sc = score_port(P, R, cand_port(P, A, R, cd, i), false);
        if (sc < best_score) {
            best_score = sc;
            nbest = 1;
        } else if (sc == best_score)
            nbest++;
    }
    int pick = (int)r_integer(R, 0, nbest - 1);
    for (int i = 0; i < cd.n; i++) {
        int port = cand_port(P, A, R, cd, i);
        if (score_port(P, R, port, false) == best_score) {
            if (pick == 0) return port;
            pick--;
        }
    }
    return -1;
}
// dfdally_get_best_from_k_connections dally:1796-1859 with k == 2
DFD_FN int best_from_k2(const DfdParams& P, const DfdArrays& A, Ctx& R, const Cands& cd) {
    if (cd.n == 0) return -1;
    if (cd.n == 1) return cand_port(P, A, R, cd, 0);
    long long n = cd.n;
    long long s1 = r_integer(R, 0, n - 1);
    long long off = r_integer(R, 1, n - 1);
    long long s2 = s1 + off;  // (s1 + off) % n with s1, off < n
    if (s2 >= n) s2 -= n;
    int p1 = cand_port(P, A, R, cd, (int)s1), p2 = cand_port(P, A, R, cd, (int)s2);
    int sc1 = score_port(P, R, p1, false), sc2 = score_port(P, R, p2, false);
    // best of the two polled connections; ties broken by one more draw (always drawn: size == 2)
    if (sc1 < sc2) {
        r_integer(R, 0, 0);
        return p1;
    }
    if (sc2 < sc1) {
        r_integer(R, 0, 0);
        return p2;
    }
    return r_integer(R, 0, 1) == 0 ? p1 : p2;
}

// dfdally_get_best_from_k_connections dally:1796-1859 for any k: k distinct indices drawn without replacement
// (dally:1826-1839), polled in ascending index order, best score wins, ties by one more draw.  The reference loops
// forever when k exceeds the number of connections (its guard is commented out, dally:1822-1823); here that is a
// device error instead of a hung GPU.
DFD_FN_NOINLINE int best_from_k(const DfdParams& P, const DfdArrays& A, Ctx& R, const Cands& cd, int k) {
    if (cd.n == 0) return -1;
    if (cd.n == 1) return cand_port(P, A, R, cd, 0);
    if (k > cd.n || cd.n > 64 || k < 1) {
        set_error(R, DFD_ERR_DEAD_END, 0x4000000000000000ULL | R.r);
        return -1;
    }
    unsigned long long sel = 0;
    long long n = cd.n, last = 0;
    for (int i = 0; i < k; i++) {
        long long ri = r_integer(R, 0, (n - 1) - i);
        long long att = (last + ri) % n;
        while ((sel >> att) & 1ULL) att = (att + 1) % n;
        sel |= 1ULL << att;
        last = att;
    }
    if (k == 1) return cand_port(P, A, R, cd, (int)last);  // one polled connection: returned without a draw
    int best_score = 2147483647, nbest = 0;
    for (unsigned long long m = sel; m; m &= m - 1) {
        int sc = score_port(P, R, cand_port(P, A, R, cd, ctz64(m)), false);
        if (sc < best_score) {
            best_score = sc;
            nbest = 1;
        } else if (sc == best_score)
            nbest++;
    }
    int pick = (int)r_integer(R, 0, nbest - 1);
    for (unsigned long long m = sel; m; m &= m - 1) {
        int port = cand_port(P, A, R, cd, ctz64(m));
        if (score_port(P, R, port, false) == best_score) {
            if (pick == 0) return port;
            pick--;
        }
    }
    return -1;
}

// dfdally_select_intermediate_group dally:9738-9790
DFD_FN void select_intermediate_group(const DfdParams& P, const DfdArrays& A, Ctx& R, PoolSlot& m) {
    int fg = m.dst_grp;
    int og = m.org_grp;
    if (m.intm_grp_id == -1) {
        int sel = (int)r_integer(R, 0, (long long)P.G - 3);  // groups other than origin and destination, ascending
        int lo = og < fg ? og : fg, hi = og < fg ? fg : og;
        int g = sel;
        if (g >= lo) g++;
        if (g >= hi) g++;
        m.intm_grp_id = (int16_t)g;
    } else {
        // re-pick among the groups this router links to directly (set<int>: ascending, unique)
        const int* gg = A.gs_group + (size_t)R.r * P.h;
        // (a router may have fewer links than global ports: unused slots of the by-gid list hold INT_MAX, at the end)
        int nvalid = 0, prev = -1;
        for (int i = 0; i < P.h; i++) {
            int g = gg[i];
            if (g == 0x7FFFFFFF) break;
            if (g != prev && g != fg && g != og) nvalid++;
            prev = g;
        }
        int sel = (int)r_integer(R, 0, (long long)nvalid - 1);
        prev = -1;
        for (int i = 0; i < P.h; i++) {
            int g = gg[i];
            if (g == 0x7FFFFFFF) break;
            if (g != prev && g != fg && g != og) {
                if (sel == 0) {
                    m.intm_grp_id = (int16_t)g;
                    break;
                }
                sel--;
            }
            prev = g;
        }
    }
}

// get_legal_nonminimal_stops dally:9853-9910
DFD_FN Cands legal_nonminimal_stops(const DfdParams& P, const DfdArrays& A, Ctx& R, PoolSlot& m) {
    Cands cd;
    cd.kind = 0;
    cd.lo = cd.hi = cd.n = 0;
    int og = m.org_grp;
    if (R.grp == og) {
        grp_range(P, A, R.r, m.intm_grp_id, &cd.lo, &cd.hi);
        if (cd.hi > cd.lo) {
            cd.kind = 1;
            cd.n = cd.hi - cd.lo;
            return cd;
        }
        if (R.r == m.origin_router) {
            cd.kind = 2;
            cd.lo = m.intm_grp_id;
            cd.n = routed_count(P, A, R, m.intm_grp_id);
            return cd;
        }
        select_intermediate_group(P, A, R, m);
        grp_range(P, A, R.r, m.intm_grp_id, &cd.lo, &cd.hi);
        cd.kind = 1;
        cd.n = cd.hi - cd.lo;
        return cd;
    }
    return cd;
}

// do_dfdally_routing dally:7110-7215 -> output port
DFD_FN int do_routing(Ctx& R, PoolSlot& m, int fdest_router) {
    const DfdParams& P = *R.P;
    const DfdArrays& A = *R.A;
    int fg = m.dst_grp;
    if ((int)R.r == fdest_router) {  // destination router: the terminal port (one rail -> no draw)
        return P.intra_radix + P.global_radix + (int)(m.dst_term - (unsigned)fdest_router * P.p);
    } else if (R.grp == fg) {  // destination group: local link(s) to the destination router
        Cands cd;
        loc_range(P, A, R.lid, fdest_router - fg * P.a, &cd.lo, &cd.hi);
        cd.kind = 3;
        cd.n = cd.hi - cd.lo;
        if (cd.n < 1) {
            set_error(R, DFD_ERR_DEAD_END, R.r);
            return -1;
        }
        if (P.routing == DFD_PROG_ADAPTIVE) return best_from_cands(P, A, R, cd);
        return cand_port(P, A, R, cd, (int)r_integer(R, 0, cd.n - 1));
    }
    if (m.last_hop == DFD_HOP_TERMINAL && m.intm_grp_id == -1) select_intermediate_group(P, A, R, m);
    if (P.routing == DFD_MINIMAL) {  // dfdally_minimal_routing dally:9913-9932
        Cands cd = legal_minimal_stops(P, A, R, fdest_router, fg);
        if (cd.n < 1) {
            set_error(R, DFD_ERR_DEAD_END, R.r);
            return -1;
        }
        return cand_port(P, A, R, cd, (int)r_integer(R, 0, cd.n - 1));
    }
    if (P.routing == DFD_NON_MINIMAL) {  // dfdally_nonminimal_routing dally:9939-9992
        m.path_type = DFD_NON_MINIMAL;
        if (R.grp == m.intm_grp_id) m.is_intm_visited = 1;
        int next_group = m.is_intm_visited == 1 ? fg : m.intm_grp_id;
        Cands cd;
        grp_range(P, A, R.r, next_group, &cd.lo, &cd.hi);
        if (cd.hi > cd.lo) {
            cd.kind = 1;
            cd.n = cd.hi - cd.lo;
        } else {
            cd.kind = 2;
            cd.lo = next_group;
            cd.n = routed_count(P, A, R, next_group);
        }
        if (cd.n < 1) {
            set_error(R, DFD_ERR_DEAD_END, R.r);
            return -1;
        }
        return cand_port(P, A, R, cd, (int)r_integer(R, 0, cd.n - 1));
    }
    // dfdally_prog_adaptive_routing dally:9995-10061
    if (R.grp == m.intm_grp_id) m.is_intm_visited = 1;
    Cands mins = legal_minimal_stops(P, A, R, fdest_router, fg);
    Cands nonmins = legal_nonminimal_stops(P, A, R, m);
    int best_min, best_non;
    if (mins.n > 0 && mins.kind == 1)
        best_min = P.global_k_picks == 2 ? best_from_k2(P, A, R, mins) : best_from_k(P, A, R, mins, P.global_k_picks);
    else
        best_min = best_from_cands(P, A, R, mins);
    if (nonmins.n > 0 && nonmins.kind == 1)
        best_non = P.global_k_picks == 2 ? best_from_k2(P, A, R, nonmins) : best_from_k(P, A, R, nonmins, P.global_k_picks);
    else
        best_non = best_from_cands(P, A, R, nonmins);
    int min_score = score_port(P, R, best_min, false);
    int non_score = score_port(P, R, best_non, true);
    if (m.path_type == DFD_NON_MINIMAL && m.is_intm_visited != 1) return best_non;
    if (min_score <= P.adaptive_threshold) return best_min;
    if (min_score <= non_score) return best_min;
    m.path_type = DFD_NON_MINIMAL;
    return best_non;
}

// 4-element arrays held in registers: select / update by index without dynamic addressing (which would push the
// whole struct to local memory)
template <typename T> DFD_FN T get4(const T* a, int i) { return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : a[3]; }
template <typename T> DFD_FN void set4(T* a, int i, T v) {
    if (i == 0) a[0] = v;
    else if (i == 1) a[1] = v;
    else if (i == 2) a[2] = v;
    else a[3] = v;
}
// whole-struct copies as 128-bit accesses (all of these structs are 16-byte aligned multiples of 16 bytes)
template <typename T> DFD_FN void load_vec(T& dst, const T* src) {
    static_assert(sizeof(T) % 16 == 0, "vector copy");
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) ((int4*)&dst)[i] = ((const int4*)src)[i];
}
template <typename T> DFD_FN void store_vec(T* dst, const T& src) {
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) ((int4*)dst)[i] = ((const int4*)&src)[i];
}

// the R_SEND timer of a port (at most one is pending: in_send_loop)
DFD_FN void schedule_send(Ctx& R, int port, double ts, unsigned seq) {
    if (ts >= R.P->ts_end) return;
    R.skey[port] = mk_key(ts, R.gid, seq);
    R.ncreated++;
}

// router_credit_send dally:7269-7326
DFD_FN void router_credit_send(Ctx& R, double now, unsigned last_hop, unsigned prev_lp, unsigned port, unsigned vc, int in_lane) {
    const DfdParams& P = *R.P;
    unsigned seq = R.h->seq++;
    if (last_hop == DFD_HOP_TERMINAL)
        push_credit_to_node(R, prev_lp, now + P.cn_credit, R.gid, seq);
    else if (last_hop == DFD_HOP_GLOBAL)
        push_credit_to_router(R, in_lane, prev_lp, port, now + P.global_credit, seq, vc);
    else
        push_credit_to_router(R, in_lane, prev_lp, port, now + P.local_credit, seq, vc);
}

// router_packet_receive dally:7377-7595; `rec` is the chunk as it came off in-lane `in_lane`
DFD_FN_NOINLINE void router_packet_receive(Ctx& R, double now, const ChunkRec& rec, int in_lane) {
    const DfdParams& P = *R.P;
    const DfdArrays& A = *R.A;
    union {
        ChunkRec rec;
        PoolSlot slot;
    } u;
    u.rec = rec;
    PoolSlot& m = u.slot;  // the two layouts agree from byte 16 on
    int dest_router_id = (int)m.dst_router;
    int in_path_type = m.path_type;  // msg->path_type (as received)
    unsigned in_port = m.prev_port, in_vc = m.prev_vc, in_last_hop = m.last_hop, in_prev = m.prev_lp;
    if (m.last_hop == DFD_HOP_TERMINAL) m.path_type = DFD_MINIMAL;
    int output_port = do_routing(R, m, dest_router_id);
    if (output_port < 0 || output_port >= P.radix) {
#ifdef DFD_DEBUG_HOST
        fprintf(stderr, "dead end r=%u lane=%d now=%f src=%u seq=%u dst_term=%u dst_router=%u dst_grp=%u last_hop=%u port=%d\n", R.r, in_lane, now, rec.src, rec.seq, rec.dst_term, rec.dst_router, rec.dst_grp, rec.last_hop, output_port);
#endif
        set_error(R, DFD_ERR_DEAD_END, R.r);
        return;
    }
    PortState ps;
    PortQ pq;
    load_vec(ps, &R.port[output_port]);
    load_vec(pq, &R.pq[output_port]);
    int conn_local = output_port < P.intra_radix;
    int conn_global = !conn_local && output_port < P.intra_radix + P.global_radix;
    if (conn_local) {
        int dlid = A.loc_dest[R.lid * P.intra_radix + output_port];
        m.next_stop = R.grp * P.a + dlid;
        m.next_is_term = 0;
    } else if (conn_global) {
        m.next_stop = A.gdest[(size_t)R.r * P.h + (output_port - P.intra_radix)];
        m.next_is_term = 0;
    } else {
        m.next_stop = m.dst_term;
        m.next_is_term = 1;
    }
    m.out_port = (uint16_t)output_port;
    m.in_lane = (uint16_t)in_lane;
    m.pad = 0;
    int max_vc_size = P.cn_vc;
    int src_group_id = m.org_grp;
    int dest_group_id = m.dst_grp;
    int output_chan = 0;
    if (R.grp == src_group_id) {
        output_chan = m.l_hop;
        if (in_path_type == DFD_NON_MINIMAL) output_chan = 1;
    } else if (R.grp != dest_group_id) {
        output_chan = 2;
    } else {
        output_chan = 3;
    }
    if (conn_local) {
        max_vc_size = P.local_vc;
        m.l_hop++;
        m.hops_cur_group++;
    }
    if (conn_global) {
        max_vc_size = P.global_vc;
        m.hops_cur_group = 0;
        m.g_hop++;
    }
    if (output_chan >= DFD_NUM_VCS) {
        set_error(R, DFD_ERR_BAD_VC, R.r);
        return;
    }
    m.out_vc = (uint8_t)output_chan;
    m.n_hop++;
    m.next = NIL;
    if (R.h->pool_nfree == 0) {
        set_error(R, DFD_ERR_POOL_OVERFLOW, R.r);
        return;
    }
    unsigned slot = R.pfree[--R.h->pool_nfree];
    int occ = get4(ps.vc_occ, output_chan);
    if (occ + P.chunk_size <= max_vc_size) {
        router_credit_send(R, now, in_last_hop, in_prev, in_port, in_vc, in_lane);
        set4(ps.vc_occ, output_chan, occ + P.chunk_size);
        if (ps.in_send_loop == 0) {
            double ts = maxd(ps.noat, now) - now;
            schedule_send(R, output_port, now + ts, R.h->seq++);
            ps.in_send_loop = 1;
        }
        unsigned tl = get4(pq.pend_tail, output_chan);  // append to pending_msgs[port][vc]
        if (tl == NIL)
            set4(pq.pend_head, output_chan, slot);
        else
            R.pool[tl].next = slot;
        set4(pq.pend_tail, output_chan, slot);
    } else {
        ps.stalled_chunks++;
        // saved_vc / saved_channel = upstream port / vc: already kept in prev_port / prev_vc
        unsigned tl = get4(pq.queue_tail, output_chan);  // append to queued_msgs[port][vc]
        if (tl == NIL)
            set4(pq.queue_head, output_chan, slot);
        else
            R.pool[tl].next = slot;
        set4(pq.queue_tail, output_chan, slot);
        ps.queued_count += P.chunk_size;
        if (ps.last_buf_full == 0.0) ps.last_buf_full = now;
        // the credit for this chunk now waits for an R_BUFFER on output_port: remember which in-lane it is owed to
        unsigned char* cnt = &R.nq[(size_t)in_lane * P.radix + output_port];
        if ((*cnt)++ == 0) R.qbits[2 * in_lane + (output_port >> 6)] |= 1ULL << (output_port & 63);
    }
    store_vec(&R.pool[slot], m);
    store_vec(&R.port[output_port], ps);
    store_vec(&R.pq[output_port], pq);
}

// The part of router_packet_send (dally:7705-8032) that moves one chunk out of `output_port`: serialisation delay, the
// record for the next hop, link statistics.  `m` is the chunk (pool-slot layout); returns the injection time without
// the router delay (when the port can start its next chunk, relative to now).
DFD_FN double router_send_core(Ctx& R, double now, int output_port, PortState& ps, const PoolSlot& m) {
    const DfdParams& P = *R.P;
    if (ps.last_buf_full) {
        ps.busy_time += (now - ps.last_buf_full);
        ps.last_buf_full = 0.0;
    }
    int to_terminal = 1, global = 0, bw = BW_CN;
    double delay = P.cn_delay;
    if (output_port < P.intra_radix) {
        to_terminal = 0;
        delay = P.local_delay;
        bw = BW_LOCAL;
    } else if (output_port < P.intra_radix + P.num_global_channels) {
        to_terminal = 0;
        global = 1;
        delay = P.global_delay;
        bw = BW_GLOBAL;
    }
    unsigned nchunks = num_chunks_for(m.packet_size, P.chunk_size);
    unsigned tail = mod_chunk(m.packet_size, P.chunk_size);
    double propagation_delay = delay;
    double injection_delay = inj_delay(P, P.chunk_size, bw);
    if (m.packet_size == 0) injection_delay = inj_delay(P, P.credit_size, bw);
    if ((m.packet_size < (unsigned)P.chunk_size) && (m.chunk_id == nchunks - 1)) injection_delay = inj_delay(P, tail, bw);
    injection_delay += P.router_delay;
    ps.noat = maxd(ps.noat, now);
    ps.noat += injection_delay;
    double injection_ts = ps.noat - now;
    double propagation_ts = injection_ts + propagation_delay;
    {
        // outgoing record = the pool slot with a new key and the "previous hop" fields replaced
        union {
            ChunkRec rec;
            PoolSlot slot;
        } u;
        u.slot = m;
        u.rec.ts = now + propagation_ts;
        u.rec.src = R.gid;
        u.rec.seq = R.h->seq++;
        u.rec.prev_lp = R.r;  // intm_lp_id
        u.rec.prev_port = m.out_port;
        u.rec.prev_vc = m.out_vc;
        u.rec.last_hop = global ? DFD_HOP_GLOBAL : DFD_HOP_LOCAL;
        if (to_terminal)
            push_chunk_to_node(R, m.next_stop, u.rec);
        else
            push_chunk_to_router(R, output_port, m.next_stop, u.rec);
    }
    if ((tail || m.packet_size == 0) && (m.chunk_id == nchunks - 1))
        ps.link_traffic += tail;
    else
        ps.link_traffic += P.chunk_size;
    ps.total_chunks++;
    ps.noat -= P.router_delay;
    return injection_ts - P.router_delay;
}

// router_packet_send dally:7705-8032 (+ get_next_router_vcg :3499-3557)
DFD_FN_NOINLINE void router_packet_send(Ctx& R, double now, int output_port) {
    PortState ps;
    PortQ pq;
    load_vec(ps, &R.port[output_port]);
    load_vec(pq, &R.pq[output_port]);
    int output_chan = -1;
    {
        int next_rr_vc = (ps.last_qos_lvl + 1) % DFD_NUM_VCS;
#pragma unroll
        for (int i = 0; i < DFD_NUM_VCS; i++) {
            if (output_chan < 0) {
                if (get4(pq.pend_head, next_rr_vc) != NIL) {
                    ps.last_qos_lvl = (uint8_t)next_rr_vc;
                    output_chan = next_rr_vc;
                } else
                    next_rr_vc = (next_rr_vc + 1) % DFD_NUM_VCS;
            }
        }
    }
    if (output_chan < 0) {
        ps.in_send_loop = 0;
        if (ps.queued_count && !ps.last_buf_full) ps.last_buf_full = now;
        store_vec(&R.port[output_port], ps);
        return;
    }
    unsigned slot = get4(pq.pend_head, output_chan);
    PoolSlot m;
    load_vec(m, &R.pool[slot]);
    double injection_ts = router_send_core(R, now, output_port, ps, m);
    set4(pq.pend_head, output_chan, m.next);  // pop pending_msgs[port][vc]
    if (m.next == NIL) set4(pq.pend_tail, output_chan, NIL);
    R.pfree[R.h->pool_nfree++] = slot;  // free the slot
    bool more = (pq.pend_head[0] != NIL) | (pq.pend_head[1] != NIL) | (pq.pend_head[2] != NIL) | (pq.pend_head[3] != NIL);
    if (!more)
        ps.in_send_loop = 0;
    else
        schedule_send(R, output_port, now + injection_ts, R.h->seq++);
    store_vec(&R.port[output_port], ps);
    store_vec(&R.pq[output_port], pq);
}

// router_buf_update dally:8069-8145
DFD_FN_NOINLINE void router_buf_update(Ctx& R, double now, int indx, int output_chan) {
    const DfdParams& P = *R.P;
    PortState ps;
    PortQ pq;
    load_vec(ps, &R.port[indx]);
    load_vec(pq, &R.pq[indx]);
    int occ = get4(ps.vc_occ, output_chan) - P.chunk_size;
    if (ps.last_buf_full > 0.0) {
        ps.busy_time += (now - ps.last_buf_full);
        ps.last_buf_full = 0.0;
    }
    unsigned qh = get4(pq.queue_head, output_chan);
    if (qh != NIL) {  // promote the oldest overflowed chunk: credit its sender, move it to pending_msgs
        PoolSlot* hd = &R.pool[qh];
        unsigned nxt = hd->next, prev_lp = hd->prev_lp, prev_port = hd->prev_port, prev_vc = hd->prev_vc, last_hop = hd->last_hop;
        int in_lane = hd->in_lane;
        set4(pq.queue_head, output_chan, nxt);
        if (nxt == NIL) set4(pq.queue_tail, output_chan, NIL);
        router_credit_send(R, now, last_hop, prev_lp, prev_port, prev_vc, in_lane);
        unsigned char* cnt = &R.nq[(size_t)in_lane * P.radix + indx];
        if (--(*cnt) == 0) R.qbits[2 * in_lane + (indx >> 6)] &= ~(1ULL << (indx & 63));
        hd->next = NIL;
        unsigned tl = get4(pq.pend_tail, output_chan);
        if (tl == NIL)
            set4(pq.pend_head, output_chan, qh);
        else
            R.pool[tl].next = qh;
        set4(pq.pend_tail, output_chan, qh);
        occ += P.chunk_size;
        ps.queued_count -= P.chunk_size;
    }
    set4(ps.vc_occ, output_chan, occ);
    if (ps.in_send_loop == 0 && get4(pq.pend_head, output_chan) != NIL) {
        double ts = maxd(ps.noat, now) - now;
        schedule_send(R, indx, now + ts, R.h->seq++);
        ps.in_send_loop = 1;
    }
    store_vec(&R.port[indx], ps);
    store_vec(&R.pq[indx], pq);
}

// bound of an empty lane: from the channel word of an external lane, from iblk for a terminal lane
// R_SNAPSHOT (PARAMS router_buffer_snapshots): router_send_snapshot_events dally:4363-4417 sends every router one self event
// per configured timestamp at init; here they are a cursor (RouterHdr::snap_next) into the time-sorted list.  The key of
// the pending one is (time, router gid, configured index): the index is the sequence number the event got at init.
// Read from the group's header in memory so that a run without snapshots (P.nsnap == 0, the common case) pays one
// uniform branch and no register.
DFD_FN bool snap_pending(const Ctx& c, Key* k) {
    const unsigned i = c.A->rh[c.rl].snap_next;
    if (i >= (unsigned)c.P->nsnap) return false;
    *k = mk_key(c.A->snap_ts[i], c.gid, c.A->snap_id[i]);
    return true;
}
DFD_FN double snap_next_ts(const Ctx& c) {
    const unsigned i = c.A->rh[c.rl].snap_next;
    return i < (unsigned)c.P->nsnap ? c.A->snap_ts[i] : DFD_TS_BIG;
}
// router_handle_snapshot_event dally:4419-4437: vc_occupancy[port][vc] of every port into snapshot_data[packet_ID]
DFD_FN_NOINLINE void router_snapshot(Ctx& c, RouterHdr* hd) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const unsigned i = hd->snap_next;
    int4* dst = A.snap + ((size_t)c.rl * P.nsnap_all + A.snap_id[i]) * P.radix;
    for (int j = 0; j < P.radix; j++) dst[j] = *(const int4*)c.port[j].vc_occ;
    hd->snap_next = i + 1;
    A.rh[c.rl].snap_next = i + 1;  // what the other lanes (and the next pick) read
}

DFD_FN double empty_bound_c(const Ctx& c, int j) { return j < c.P->ext ? cw_bound(c.seen_c[j]) : c.iblk[j - c.P->ext]; }
DFD_FN double empty_bound_k(const Ctx& c, int j) { return j < c.P->ext ? cw_bound(c.seen_k[j]) : c.iblk[c.P->p + (j - c.P->ext)]; }

// Router phase: executes the router's events in key order while the earliest one lies below every bound of its
// empty lanes.  Candidates are the heads of the chunk lanes, the heads of the credit lanes and the R_SEND timers;
// an empty lane contributes its bound as a blocker {bound, 0, 0}, which beats an event of the same timestamp.
// Every lane scans its share of the 3*radix entries, a warp argmin picks, lane 0 executes.
// Timestamps are non-negative doubles, so a key orders like the integer pair (timestamp bits, src << 32 | seq).
DFD_FN void key_min_at(unsigned long long& bh, unsigned long long& bl, int& bidx, const Key* p, int idx) {
    const int4 v = *(const int4*)p;
    const unsigned long long h = ((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x;
    const unsigned long long l = ((unsigned long long)(unsigned)v.z << 32) | (unsigned)v.w;
    if (h < bh || (h == bh && l < bl)) {
        bh = h;
        bl = l;
        bidx = idx;
    }
}

// What an activation needs to know about the router once nothing more of it can be executed: the earliest pending event
// (ev_min), the router's clock (a lower bound on its next event) and whether anything is pending at all.
// `hard`: the minimum without the bounds of empty TERMINAL lanes.  Those bounds are derived from the router's own clock
// and would let it creep upwards round by round; the next router event, however, can only come from a pending router
// event, an external lane, or an event pending at one of the terminals.
DFD_FN void router_horizon(Ctx& c, double best_ts) {
    const DfdParams& P = *c.P;
    double hard = DFD_TS_BIG;  // bounds of the empty EXTERNAL lanes
    double evm = DFD_TS_BIG;   // pending events: heads of the non-empty lanes, R_SEND timers
    LANE_FOR(j, P.radix) {
        const LaneCtl L = c.lane[j];
        const bool ext = j < P.ext;
        double t = c.ckey[j].ts;
        if (L.chead != L.ctail)
            evm = mind(evm, t);
        else if (ext)
            hard = mind(hard, t);
        t = c.kkey[j].ts;
        if (L.khead != L.ktail)
            evm = mind(evm, t);
        else if (ext)
            hard = mind(hard, t);
        evm = mind(evm, c.skey[j].ts);
    }
    if (P.nsnap) evm = mind(evm, snap_next_ts(c));
    c.ev_min = mind(warp_min_d(mind(evm, c.node_push_min)), c.node_floor);
    hard = mind(warp_min_d(hard), c.ev_min);
    c.rclock = maxd(best_ts, mind(hard, DFD_TS_BIG));
    c.nothing_pending = !(hard < DFD_TS_BIG);
}

// SPLIT: the horizon (router_horizon) is computed by a pass of its own when the phase ends -- fewer instructions per
// event, the right trade where instruction supply bounds the run (several groups per warp); otherwise every pick scan
// folds it in -- one pass less on the critical path of a latency-bound run (one group per warp).
template <bool SPLIT> DFD_FN unsigned router_phase(Ctx& c, RouterHdr* hd, unsigned budget) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const int radix = P.radix;
    unsigned nproc = 0;
    for (;;) {
        // pick: the smallest key among the 3*radix entries (every lane scans its share, a warp argmin settles it)
        unsigned long long bh = d2bits(DFD_TS_BIG), bl = ~0ULL;
        int bidx = 3 * radix;
        double hard = DFD_TS_BIG;
        if (SPLIT) {
            LANE_FOR(j, radix) {
                key_min_at(bh, bl, bidx, &c.ckey[j], j);
                key_min_at(bh, bl, bidx, &c.kkey[j], radix + j);
                key_min_at(bh, bl, bidx, &c.skey[j], 2 * radix + j);
            }
        } else {
            double evm = DFD_TS_BIG;  // as in router_horizon
            LANE_FOR(j, radix) {
                const LaneCtl L = c.lane[j];
                const bool ext = j < P.ext;
                key_min_at(bh, bl, bidx, &c.ckey[j], j);
                key_min_at(bh, bl, bidx, &c.kkey[j], radix + j);
                key_min_at(bh, bl, bidx, &c.skey[j], 2 * radix + j);
                double t = c.ckey[j].ts;
                if (L.chead != L.ctail)
                    evm = mind(evm, t);
                else if (ext)
                    hard = mind(hard, t);
                t = c.kkey[j].ts;
                if (L.khead != L.ktail)
                    evm = mind(evm, t);
                else if (ext)
                    hard = mind(hard, t);
                evm = mind(evm, c.skey[j].ts);
            }
            if (P.nsnap) evm = mind(evm, snap_next_ts(c));
            c.ev_min = mind(warp_min_d(mind(evm, c.node_push_min)), c.node_floor);
            hard = mind(warp_min_d(hard), c.ev_min);
        }
        Key best = mk_key(bits2d(bh), (unsigned)(bl >> 32), (unsigned)bl);
        warp_argmin_key(best, bidx);
        bool snap = false;
        if (P.nsnap) {  // the pending R_SNAPSHOT competes with the winner of the scan
            Key sk;
            if (snap_pending(c, &sk) && key_lt(sk, best)) {
                best = sk;
                snap = true;
            }
        }
        if (!SPLIT) {
            c.rclock = maxd(best.ts, mind(hard, DFD_TS_BIG));
            c.nothing_pending = !(hard < DFD_TS_BIG);
        }
        int kind = 0, j = 0;
        LaneCtl L = LaneCtl();
        bool is_event = false;
        if (snap) {
            kind = 3;
            is_event = true;
        } else if (bidx < 3 * radix) {
            kind = bidx >= 2 * radix ? 2 : bidx >= radix ? 1 : 0;
            j = bidx - kind * radix;
            L = c.lane[j];
            is_event = kind == 2 || (kind == 0 ? L.chead != L.ctail : L.khead != L.ktail);
        }
        if (!is_event || nproc >= budget) {
            if (SPLIT) router_horizon(c, best.ts);
            break;
        }
        if (IS_LANE0()) {
            const double now = best.ts;
            c.h = hd;
            check_order(c, &hd->last_ts, &hd->last_src, &hd->last_seq, best.ts, best.src, best.seq, kind < 2, c.r);
            if (kind == 3) {
                count_ev(c.ev, EV_R_SNAPSHOT);
                router_snapshot(c, hd);
            } else if (kind == 0) {
                ChunkRec rec;
                const ChunkRec* src = c.ring_c + ring_c_idx(A, j, L.chead);
#pragma unroll
                for (int i = 0; i < 6; i++) ((int4*)&rec)[i] = ld_cg16((const int4*)src + i);
                L.chead++;
                c.lane[j].chead = L.chead;
                if (L.chead != L.ctail) {
                    const ChunkRec* nx = c.ring_c + ring_c_idx(A, j, L.chead);
                    c.ckey[j] = key_from_int4(ld_cg16(nx));
                } else
                    c.ckey[j] = mk_key(empty_bound_c(c, j), 0, 0);
                count_ev(c.ev, EV_R_ARRIVE);
                router_packet_receive(c, now, rec, j);
            } else if (kind == 1) {
                const Key* src = c.ring_k + ring_k_idx(A, j, L.khead);
                Key rec = key_from_int4(ld_cg16(src));
                L.khead++;
                c.lane[j].khead = L.khead;
                if (L.khead != L.ktail) {
                    Key nx = key_from_int4(ld_cg16(c.ring_k + ring_k_idx(A, j, L.khead)));
                    nx.src &= DFD_GID_MASK;
                    c.kkey[j] = nx;
                } else
                    c.kkey[j] = mk_key(empty_bound_k(c, j), 0, 0);
                count_ev(c.ev, EV_R_BUFFER);
                router_buf_update(c, now, j, (int)(rec.src >> 30));
            } else {
                c.skey[j].ts = DINF;
                count_ev(c.ev, EV_R_SEND);
                router_packet_send(c, now, j);
            }
        }
        nproc++;
        WARP_SYNC();
    }
    return nproc;
}

// ------------------------------------------------------------------------------------------------
// group activation
// ------------------------------------------------------------------------------------------------
DFD_FN void ctx_bind_group(Ctx& c, unsigned rl) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    c.rl = rl;
    c.r = (unsigned)c.X->r0 + rl;
    const size_t ro = (size_t)rl * P.radix;
    c.h = &A.rh[rl];
    c.port = A.port + ro;
    c.pq = A.portq + ro;
    c.skey = A.skey + ro;
    c.ckey = A.ckey + ro;
    c.kkey = A.kkey + ro;
    c.lane = A.lane + ro;
    c.seen_c = A.seen_c + ro;
    c.seen_k = A.seen_k + ro;
    c.nq = A.nq + ro * P.radix;
    c.qbits = A.qbits + 2 * ro;
    c.pool = A.pool + (size_t)rl * P.rpool_cap;
    c.pfree = A.pfree + (size_t)rl * P.rpool_cap;
    c.ring_c = A.ring_c + (size_t)rl * P.cslots;
    c.ring_k = A.ring_k + (size_t)rl * P.kslots;
    c.iblk = A.iblk + (size_t)rl * 2 * P.p;
}

// Channel word to a neighbour.  Sharded runs with publisher queues (DfdParams::npubq, see PubCtl): a word whose target lives on
// another rank is handed to this group's publisher queue instead of being stored behind a system-scope fence of this warp.
DFD_FN void publish_word(const Ctx& c, unsigned long long* target, unsigned long long w, bool sys) {
#ifdef __CUDACC__
    const DfdParams& P = *c.P;
    if (P.npubq && (unsigned long long)((const char*)target - c.X->local_base) >= c.X->local_bytes) {
        const DfdArrays& A = *c.A;
        const unsigned qi = c.rl % (unsigned)P.npubq;
        PubCtl* ctl = &A.pubctl[qi];
        const unsigned long long t = atomicAdd(&ctl->reserve, 1ULL);
        while (t - ld_strong64(&ctl->head) >= (unsigned long long)DFD_PUBQ_CAP) {  // ring full: wait for the publisher
            if (ld_strong64(&A.ctrl->error) | ld_strong64(&A.ctrl->done)) return;
            __nanosleep(200);
        }
        const unsigned long long tag = (t / DFD_PUBQ_CAP) % 3ULL + 1ULL;
        st_relaxed_b128(&A.pubq[(size_t)qi * DFD_PUBQ_CAP + (t & (DFD_PUBQ_CAP - 1))], (unsigned long long)target | tag, w);
        return;
    }
#endif
    st_strong64(target, w, sys);
}

struct ActResult {
    unsigned nexec;  // events executed (warp total)
    bool polled_change;
};


// earliest pending event of a terminal: self events and the heads of its two rings
DFD_FN double node_next_pending(const Ctx& c, unsigned nl, const NodeState* ns) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    double next = ns->self_next;
    if (ns->cin_head != A.ncin_tail[nl]) next = mind(next, A.ncin[(size_t)nl * P.ncin_cap + (ns->cin_head & (P.ncin_cap - 1))].ts);
    if (ns->kin_head != A.nkin_tail[nl]) next = mind(next, A.nkin[(size_t)nl * P.nkin_cap + (ns->kin_head & (P.nkin_cap - 1))].ts);
    return next;
}

// true when terminal t has an event it may execute now: its earliest pending timestamp lies below its horizon
DFD_FN bool node_has_work(const Ctx& c, unsigned nl, const NodeState* ns, const NodeBounds& b) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    double next = ns->self_next, horizon = DFD_TS_BIG;
    if (ns->cin_head != A.ncin_tail[nl])
        next = mind(next, A.ncin[(size_t)nl * P.ncin_cap + (ns->cin_head & (P.ncin_cap - 1))].ts);
    else
        horizon = b.rb;
    if (ns->kin_head != A.nkin_tail[nl])
        next = mind(next, A.nkin[(size_t)nl * P.nkin_cap + (ns->kin_head & (P.nkin_cap - 1))].ts);
    else
        horizon = mind(horizon, b.kb);
    return next < horizon;
}

// One activation of group rl.  Returns the number of events executed (0 when nothing was safe).
template <bool SPLIT = false> DFD_FN ActResult group_activate(Ctx& c, unsigned rl) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const DfdPeers& X = *c.X;
    ActResult res;
    res.nexec = 0;
    // ---- (1) poll the channel words of the external lanes.  The idle poll touches one 16-byte poll header and the two
    // word arrays; the group's pointers are bound only when something has to be done.
    const PollHdr ph = A.pollhdr[rl];
    // ---- (1) poll the channel words of the external lanes
    // A changed word is RELEVANT only if the lane was empty and its old bound did not lie beyond the group's earliest
    // pending event (hd.ev_min): only such a bound can have been what the group was waiting for, and only such a change
    // can move the group's clock.  Everything else (a neighbour that is ahead announcing that it is further ahead, records
    // queueing up behind the head of a lane) is absorbed into the lane state and acted upon by the next full activation.
    // Outputs that stay stale because of this all lie beyond ev_min + a credit delay, so they can never be what the
    // LP holding the globally earliest event waits for ("relevance filter").
    const bool first = (ph.flags & 2u) != 0;
    bool changed = ph.flags != 0;  // continuation of an activation that stopped on its budget, or the very first one
    const double ev_min = ph.ev_min;
    {
        const size_t ro = (size_t)rl * P.radix;
        const unsigned long long* cwc = A.cw_c + ro;
        const unsigned long long* cwk = A.cw_k + ro;
        unsigned long long* sc = A.seen_c + ro;
        unsigned long long* sk = A.seen_k + ro;
        LANE_FOR(j, P.ext) {
            unsigned long long w = ld_strong64(&cwc[j]);
            unsigned long long v = ld_strong64(&cwk[j]);
            const unsigned long long w0 = sc[j];
            const unsigned long long v0 = sk[j];
            if (w != w0) {
                sc[j] = w;
                LaneCtl& L = A.lane[ro + j];
                bool was_empty = L.chead == L.ctail;
                if (was_empty && cw_bound(w0) <= ev_min) changed = true;
                L.ctail = cw_tail(w, L.chead);
                if (was_empty) {
                    if (L.chead != L.ctail)
                        A.ckey[ro + j] = key_from_int4(ld_cg16(A.ring_c + (size_t)rl * P.cslots + ring_c_idx(A, j, L.chead)));
                    else
                        A.ckey[ro + j] = mk_key(cw_bound(w), 0, 0);
                }
            }
            if (v != v0) {
                sk[j] = v;
                LaneCtl& L = A.lane[ro + j];
                bool was_empty = L.khead == L.ktail;
                if (was_empty && cw_bound(v0) <= ev_min) changed = true;
                L.ktail = cw_tail(v, L.khead);
                if (was_empty) {
                    if (L.khead != L.ktail) {
                        Key nx = key_from_int4(ld_cg16(A.ring_k + (size_t)rl * P.kslots + ring_k_idx(A, j, L.khead)));
                        nx.src &= DFD_GID_MASK;
                        A.kkey[ro + j] = nx;
                    } else
                        A.kkey[ro + j] = mk_key(cw_bound(v), 0, 0);
                }
            }
        }
    }
    res.polled_change = warp_any(changed);
    if (!res.polled_change) return res;  // the previous activation ran to its fixed point on exactly this input
    if (c.rl != rl || c.port == nullptr) ctx_bind_group(c, rl);
    if (P.prefetch) {
        // Partitions far larger than L2: the handlers below walk port state, FIFO heads and terminal state through chains
        // of dependent loads, each a DRAM miss.  Asking for all of it now, in parallel, turns the chain into one miss.
        for (int j = DFD_LANE() * 2; j < P.radix; j += 64) {  // two 64-byte entries per line
            prefetch_l1(&c.port[j]);
            prefetch_l1(&c.pq[j]);
        }
        LANE_FOR(j, P.radix / 8 + 1) prefetch_l1(&c.lane[j * 8]);
    }
    c.ncreated = 0;
    c.nexec = 0;
    c.wrote_remote = false;
    RouterHdr hd = A.rh[rl];  // lane 0's working copy; written back once
    WARP_SYNC();
    // phase cycle accounting: only in the one-warp-per-block kernels (the pooled kernel is short of registers)
    long long t0 = SPLIT ? 0LL : CYC_NOW(), t1;
    // ---- (2) rounds of router phase + node phase until nothing more is safe (or the event budget is spent)
    c.gid = hd.gid;
    c.lid = hd.lid;
    c.grp = hd.grp;
    c.rclock = hd.clock < 0.0 ? 0.0 : hd.clock;
    const unsigned nbase = c.r * (unsigned)P.p;        // first node of the group (global id)
    const unsigned nlbase = nbase - (unsigned)X.n0;     // ... local index
    unsigned total = 0, nrouter = 0;
    bool more = false;
    for (;;) {
        const double clock_in = c.rclock;
        if (!SPLIT) c.nrounds++;
        if (!SPLIT) {
            t1 = CYC_NOW();
            c.cyc[0] += t1 - t0;
            t0 = t1;
        }
        // bounds of the terminal lanes as seen by the router
        bool nwork = false;
        double nfloor = DFD_TS_BIG;
        LANE_FOR(t, P.p) {
            const NodeState* ns = &A.node[nlbase + t];
            NodeBounds b = node_bounds(c, t, nlbase + t, ns);
            nfloor = mind(nfloor, node_next_pending(c, nlbase + t, ns));
            const int s = P.ext + t;
            c.iblk[t] = b.x;
            c.iblk[P.p + t] = b.y;
            LaneCtl L = c.lane[s];
            if (L.chead == L.ctail) c.ckey[s] = mk_key(b.x, 0, 0);
            if (L.khead == L.ktail) c.kkey[s] = mk_key(b.y, 0, 0);
            nwork = nwork || node_has_work(c, nlbase + t, ns, b);
        }
        c.node_floor = warp_min_d(nfloor);
        c.node_push_min = DFD_TS_BIG;
        WARP_SYNC();
        if (!SPLIT) {
            t1 = CYC_NOW();
            c.cyc[1] += t1 - t0;
            t0 = t1;
        }
        unsigned nr = router_phase<SPLIT>(c, &hd, (unsigned)P.act_budget - nrouter);
        nrouter += nr;
        WARP_SYNC();
        if (!SPLIT) {
            t1 = CYC_NOW();
            c.cyc[2] += t1 - t0;
            t0 = t1;
        }
        unsigned nn = 0;
        // the router phase can only raise a terminal's horizon or hand it new records: without router events the
        // terminals that had nothing to do before it still have nothing to do
        if (nr != 0 || warp_any(nwork)) {
            LANE_FOR(t, P.p) {
                const NodeState* ns = &A.node[nlbase + t];
                NodeBounds b = node_bounds(c, t, nlbase + t, ns);
                if (node_has_work(c, nlbase + t, ns, b)) nn += node_process(c, t, nbase + t, nlbase + t, b.rb, b.kb);
            }
            WARP_SYNC();
        }
        c.nexec += nn;
        if (IS_LANE0()) c.nexec += nr;
        nn = warp_sum_u(nn);
        total += nr + nn;
        if (!SPLIT) {
            t1 = CYC_NOW();
            c.cyc[3] += t1 - t0;
            t0 = t1;
        }
#ifdef DFD_DEBUG_HOST
        {
            static unsigned long long dbg_rounds = 0;
            if (++dbg_rounds % 1000000 == 0)
                fprintf(stderr, "rounds %llu r=%u nr=%u nn=%u rclock=%.6f clock_in=%.6f total=%u nrouter=%u\n", dbg_rounds, c.r, nr, nn, c.rclock, clock_in, total, nrouter);
        }
#endif
        if ((nr + nn) == 0 && (!(c.rclock > clock_in) || c.nothing_pending)) break;  // fixed point
        if (nrouter >= (unsigned)P.act_budget) {  // publish what is known so far; the next pass continues
            more = true;
            break;
        }
    }
    res.nexec = total;
    if (IS_LANE0()) {
        hd.clock = c.rclock;
        hd.more = more ? 1u : 0u;
        hd.ev_min = c.ev_min;  // exact after a round that executed nothing (the fixed-point exit); unused after `more`
        PollHdr nph;
        nph.ev_min = c.ev_min;
        nph.flags = more ? 1u : 0u;
        nph.pad = 0;
        A.pollhdr[rl] = nph;
        if (total) hd.activations++;
        A.rh[rl] = hd;
    }
    c.h = &A.rh[rl];
    WARP_SYNC();
    // ---- (3) publish: created count, one release fence, consumed count and the channel words.  Bounds alone need no
    // fence: they only promise something about records that do not exist yet.
    const unsigned created = warp_sum_u(c.ncreated);
    const bool sys = X.world > 1;
    if (total | created) {
        if (IS_LANE0() && created) red_add_i64((long long*)&A.ctrl->created, (long long)created, sys);
        fence_release(sys && P.npubq == 0);  // with publisher queues the system-scope fence is the publisher's
        if (IS_LANE0() && total) red_add_i64((long long*)&A.ctrl->consumed, (long long)total, sys);
    }
    {
        unsigned long long* pubc = A.pub_c + (size_t)rl * P.radix;
        unsigned long long* pubk = A.pub_k + (size_t)rl * P.radix;
        unsigned long long** tgtc = A.tgt_c + (size_t)rl * P.radix;
        unsigned long long** tgtk = A.tgt_k + (size_t)rl * P.radix;
        if (first) {  // resolve once where this router's channel words go
            LANE_FOR(q, P.ext) {
                int dst, dlane, src, sport;
                if (q < P.intra_radix) {
                    int dlid = A.loc_dest[c.lid * P.intra_radix + q];
                    dst = dlid < 0 ? -1 : c.grp * P.a + dlid;
                    dlane = dlid < 0 ? -1 : A.lout_lane[c.lid * P.intra_radix + q];
                    int v = A.lin_src[c.lid * P.intra_radix + q];
                    src = v < 0 ? -1 : c.grp * P.a + (v >> 16);
                    sport = v & 0xFFFF;
                } else {
                    dst = A.gdest[(size_t)c.r * P.h + (q - P.intra_radix)];
                    dlane = dst < 0 ? -1 : A.gout_lane[(size_t)c.r * P.h + (q - P.intra_radix)];
                    src = A.gin_src[(size_t)c.r * P.h + (q - P.intra_radix)];
                    sport = src < 0 ? -1 : A.gin_port[(size_t)c.r * P.h + (q - P.intra_radix)];
                }
                unsigned long long* tc = nullptr;
                unsigned long long* tk = nullptr;
                if (dst >= 0) {
                    unsigned dl;
                    int owner = router_home(c, (unsigned)dst, &dl);
                    tc = A.cw_c + (size_t)dl * P.radix + dlane;
                    if (owner != X.rank) tc = on_rank(X, tc, owner);
                }
                if (src >= 0) {
                    unsigned ul;
                    int owner = router_home(c, (unsigned)src, &ul);
                    tk = A.cw_k + (size_t)ul * P.radix + sport;
                    if (owner != X.rank) tk = on_rank(X, tk, owner);
                }
                tgtc[q] = tc;
                tgtk[q] = tk;
            }
        }
        LANE_FOR(q, P.ext) {
            // chunk bound of my port q -> channel word of the in-lane it feeds.  Far ahead of anything the neighbour can
            // be waiting for (it carries at least router_delay + link delay), so it is re-published only when it announces
            // new records or has moved by a good part of that lookahead: every store wakes the neighbour up.
            unsigned long long* tc = tgtc[q];
            if (tc) {
                const PortState& ps = c.port[q];
                const bool loc = q < P.intra_radix;
                double st = ps.in_send_loop ? c.skey[q].ts : c.rclock;
                double bound = lower((maxd(ps.noat, st) + (loc ? P.lb_injrd_local : P.lb_injrd_global)) + (loc ? P.local_delay : P.global_delay));
                unsigned long long w = cw_pack(bound, c.lane[q].sent_c);
                unsigned long long old = pubc[q];
                if (w != old && (((w ^ old) & DFD_STAMP_MASK) != 0 || old == ~0ULL || !(bound < cw_bound(old) + (loc ? P.hyst_local : P.hyst_global)))) {
                    pubc[q] = w;
                    publish_word(c, tc, w, sys);
                }
            }
            // credit bound of my in-lane q -> channel word of the port that feeds it
            unsigned long long* tk = tgtk[q];
            if (tk) {
                const LaneCtl L = c.lane[q];
                double a_in = L.chead != L.ctail ? c.ckey[q].ts : cw_bound(c.seen_c[q]);
                double cd = q < P.intra_radix ? P.local_credit : P.global_credit;
                unsigned long long w = cw_pack(cd + mind(a_in, late_credit_bound(c, q)), L.sent_k);
                if (w != pubk[q]) {
                    pubk[q] = w;
                    publish_word(c, tk, w, sys);
                }
            }
        }
    }
    if (!SPLIT) {
        t1 = CYC_NOW();
        c.cyc[4] += t1 - t0;
    }
    return res;
}
