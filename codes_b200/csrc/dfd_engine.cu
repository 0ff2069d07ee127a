// dfd_engine.cu -- B200 (sm_100a) window engine for the CODES dragonfly-dally hot path.
//
// One persistent cooperative kernel executes the whole simulation.  Simulated time advances in
// conservative windows [T, T+L): L is the smallest cross-LP delay (a credit delay by default), T the
// global minimum pending timestamp.  Inside a window every LP that has something due processes its own
// events strictly in (timestamp, sender gid, sender sequence) order -- one thread per LP, 32 LPs per
// warp -- and appends the events it emits to the destination LP's inbox.  One grid barrier closes the
// window; the next T is an atomicMin reduction folded into the same barrier.
//
// LP classes:
//   node   = svr ("nw-lp", model-net-synthetic-dragonfly-all.c) + model-net base LP (model-net-lp.c)
//            + dragonfly-dally terminal (dragonfly-dally.cxx) of one compute node, fused because they
//            only exchange zero / sub-ns delay events with each other;
//   router = dragonfly-dally router LP.
// The handlers restate the reference's forward handlers; each cites the lines it follows
// (paths relative to the reference tree, "dally" = src/networks/model-net/dragonfly-dally.cxx,
// "synthetic" = src/network-workloads/model-net-synthetic-dragonfly-all.c).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (timestamps must be plain IEEE double
// operations in the reference's order, so FMA contraction is disabled for the whole file).

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dfd_engine.h"


#define INF_BITS 0x7FF0000000000000ULL
#define NIL 0xFFFFFFFFu
#define DFD_BLOCK 512

__device__ __forceinline__ double bits2d(unsigned long long b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ unsigned long long d2bits(double d) { return (unsigned long long)__double_as_longlong(d); }
#define DINF bits2d(INF_BITS)

struct Ctx {
    const DfdParams* P;
    const DfdArrays* A;
    double t_end;
    long long b_beg;  // timing-wheel bucket of the window start
    double cand;  // min timestamp of everything this thread left pending or pushed this window
    const DfdPeers* X;
    int cur, nxt;
    unsigned widx;
    bool remote_writes;  // this thread wrote into another rank's memory during the window
#ifdef DFD_PROFILE
    long long pst;
#endif
    unsigned* ev;  // per-block event counters by type (shared memory)
};
#define COUNT_EV(c, t) atomicAdd(&(c).ev[t], 1u)
#ifdef DFD_PROFILE
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(c, t) atomicAdd(&(c).A->counters[32 + (t)], (unsigned long long)(clock64() - prof_t0))
// fine-grained stamps inside the R_ARRIVE path: counters[58 + k] accumulates the cycles since the previous stamp
#define PSTAMP0(c) (c).pst = clock64()
#define PSTAMP(c, k)                                                                     \
    do {                                                                                 \
        long long _t = clock64();                                                        \
        atomicAdd(&(c).A->counters[58 + (k)], (unsigned long long)(_t - (c).pst));       \
        (c).pst = clock64();                                                             \
    } while (0)
#else
#define PROF_T0()
#define PROF_ADD(c, t)
#define PSTAMP0(c)
#define PSTAMP(c, k)
#endif

__device__ __forceinline__ void set_error(const DfdArrays& A, unsigned code, unsigned long long info) {
    if (atomicCAS(&A.counters[2], 0ULL, (unsigned long long)code) == 0ULL) A.counters[3] = info;
}

// ------------------------------------------------------------------------------------------------
// CLCG4 (L'Ecuyer-Andres), int32 Schrage form; identical values to ROSS' rng_gen_val
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double clcg4_unif(int32_t* s) {
    int k, v;
    double u = 0.0;
    v = s[0];
    k = v / 46693;
    v = 45991 * (v - k * 46693) - k * 25884;
    if (v < 0) v = v + 2147483647;
    s[0] = v;
    u = u + 4.65661287524579692e-10 * v;
    v = s[1];
    k = v / 10339;
    v = 207707 * (v - k * 10339) - k * 870;
    if (v < 0) v = v + 2147483543;
    s[1] = v;
    u = u - 4.65661310075985993e-10 * v;
    if (u < 0) u = u + 1.0;
    v = s[2];
    k = v / 15499;
    v = 138556 * (v - k * 15499) - k * 3979;
    if (v < 0) v = v + 2147483423;
    s[2] = v;
    u = u + 4.65661336096842131e-10 * v;
    if (u >= 1.0) u = u - 1.0;
    v = s[3];
    k = v / 43218;
    v = 49689 * (v - k * 43218) - k * 24121;
    if (v < 0) v = v + 2147483323;
    s[3] = v;
    u = u - 4.65661357780891134e-10 * v;
    if (u < 0) u = u + 1.0;
    return u;
}
// tw_rand_integer: low + (long)(unif * (high + 1 - low)); no draw when high < low
__device__ __forceinline__ long long clcg4_integer(int32_t* s, long long low, long long high, unsigned long long* draws) {
    if (high < low) return 0;
    (*draws)++;
    return low + (long long)(clcg4_unif(s) * (double)(high + 1 - low));
}

__device__ __forceinline__ double bytes_to_ns(unsigned long long bytes, double GB_p_s) {  // dally:1588-1599
    double time = ((double)bytes) / (1024.0 * 1024.0 * 1024.0);
    time = time / GB_p_s;
    time = time * 1000.0 * 1000.0 * 1000.0;
    return time;
}
__device__ __forceinline__ double maxd(double a, double b) { return a < b ? b : a; }
// exact n / d for n < 2^32 via the precomputed reciprocal (magic == 0 encodes d == 1)
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned long long magic) { return magic ? (unsigned)__umul64hi((unsigned long long)n, magic) : n; }
// bytes_to_ns(bytes, bw) for bw in {cn, local, global}: the two sizes every chunk of this workload uses are precomputed
enum { BW_CN = 0, BW_LOCAL = 1, BW_GLOBAL = 2 };
__device__ __forceinline__ double inj_delay(const DfdParams& P, unsigned bytes, int which) {
    if (bytes == (unsigned)P.chunk_size) return which == BW_CN ? P.inj_chunk_cn : which == BW_LOCAL ? P.inj_chunk_local : P.inj_chunk_global;
    if (bytes == (unsigned)P.payload_sz % (unsigned)P.chunk_size) return which == BW_CN ? P.inj_pay_cn : which == BW_LOCAL ? P.inj_pay_local : P.inj_pay_global;
    return bytes_to_ns(bytes, which == BW_CN ? P.cn_bw : which == BW_LOCAL ? P.local_bw : P.global_bw);
}
__device__ __forceinline__ unsigned mod_chunk(unsigned x, unsigned chunk) { return x < chunk ? x : x % chunk; }

// gid arithmetic (codes_mapping.c:149-378 collapsed to closed forms)
__device__ __forceinline__ unsigned svr_gid(const DfdParams& P, unsigned k) { return (k / P.p) * P.lps_per_rep + P.off_svr + k % P.p; }
__device__ __forceinline__ unsigned term_gid(const DfdParams& P, unsigned k) { return (k / P.p) * P.lps_per_rep + P.off_term + k % P.p; }
__device__ __forceinline__ unsigned rtr_gid(const DfdParams& P, unsigned r) { return r * P.lps_per_rep + P.off_rtr; }

__device__ __forceinline__ bool key_less(double ts, unsigned src, unsigned seq, double ts2, unsigned src2, unsigned seq2) {
    if (ts != ts2) return ts < ts2;
    if (src != src2) return src < src2;
    return seq < seq2;
}

// ------------------------------------------------------------------------------------------------
// group sharding: which rank owns a group, and the address of one of its exchange arrays in my address space
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int owner_of_group(const DfdPeers& X, int g) {
    int k = 0;
    while (g >= X.grp_lo[k + 1]) k++;
    return k;
}
template <typename T> __device__ __forceinline__ T* on_rank(const DfdPeers& X, T* local, int owner) {
    return (T*)(X.base[owner] + ((char*)local - X.local_base));
}

// ------------------------------------------------------------------------------------------------
// event delivery
// ------------------------------------------------------------------------------------------------
// chunk towards a router on this GPU: append to the router's inbox of the NEXT window parity (drained by the
// router at the start of the next window, so a record is never read in the window it is written in)
__device__ __forceinline__ void push_chunk_to_router(Ctx& c, unsigned r, const ChunkRec& rec) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (rec.ts >= P.ts_end) return;
    unsigned* tail = &A.rcin_tail[(size_t)c.nxt * P.NR + r];
    ChunkRec* box = &A.rcin[((size_t)c.nxt * P.NR + r) * P.rcin_cap];
    unsigned idx = atomicAdd(tail, 1u);
    if (idx >= (unsigned)P.rcin_cap) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, r);
        return;
    }
    int4* dst = (int4*)&box[idx];
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 0; i < 6; i++) __stcg(dst + i, src[i]);
    c.cand = fmin(c.cand, rec.ts);
}
__device__ __forceinline__ void push_credit_to_router(Ctx& c, unsigned r, double ts, unsigned src, unsigned seq, unsigned port, unsigned vc) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (ts >= P.ts_end) return;
    unsigned* tail = &A.rkin_tail[(size_t)c.nxt * P.NR + r];
    CreditRec* box = &A.rkin[((size_t)c.nxt * P.NR + r) * P.rkin_cap];
    unsigned idx = atomicAdd(tail, 1u);
    if (idx >= (unsigned)P.rkin_cap) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, r);
        return;
    }
    CreditRec rec;
    rec.ts = ts;
    rec.src = src;
    rec.seq = seq;
    rec.port = (uint16_t)port;
    rec.vc = (uint8_t)vc;
    rec.pad0 = 0;
    rec.pad1 = 0;
    rec.pad2 = 0;
    int4* dst = (int4*)&box[idx];
    __stcg(dst, ((const int4*)&rec)[0]);
    __stcg(dst + 1, ((const int4*)&rec)[1]);
    c.cand = fmin(c.cand, ts);
}
// Global link whose far end lives on another GPU.  Every such link direction has exactly one sending router, so
// the receiver keeps one lane per (router, global port): the sender counts its own pushes, writes record, count
// and lane bit as posted stores / reductions -- nothing has to return over NVLink inside a handler.
__device__ __forceinline__ unsigned xlane_reserve(Ctx& c, unsigned my_port_idx, bool credit, unsigned* packed) {
    uint2* sp = &c.A->xsend[my_port_idx];
    uint2 v = *sp;
    if (v.x != c.widx + 1u) {
        v.x = c.widx + 1u;
        v.y = 0;
    }
    unsigned idx = credit ? (v.y >> 16) : (v.y & 0xFFFFu);
    v.y += credit ? 0x10000u : 1u;
    *sp = v;
    *packed = v.y;
    return idx;
}
__device__ __forceinline__ void xlane_publish(Ctx& c, size_t lane, unsigned r, unsigned kp, unsigned packed, int owner) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const DfdPeers& X = *c.X;
    __stcg(on_rank(X, &A.xcnt[lane], owner), packed);
    unsigned long long* mk = on_rank(X, &A.xmask[(size_t)c.nxt * P.NR + r], owner);
    asm volatile("red.relaxed.sys.global.or.b64 [%0], %1;" ::"l"(mk), "l"(1ULL << kp) : "memory");
    c.remote_writes = true;
}
__device__ __forceinline__ void push_chunk_remote(Ctx& c, unsigned me, unsigned my_gport, unsigned r, const ChunkRec& rec, int owner) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const DfdPeers& X = *c.X;
    if (rec.ts >= P.ts_end) return;
    unsigned kp = (unsigned)A.grev[(size_t)me * P.h + my_gport];
    size_t lane = ((size_t)c.nxt * P.NR + r) * P.h + kp;
    unsigned packed;
    unsigned idx = xlane_reserve(c, me * P.h + my_gport, false, &packed);
    if (idx >= (unsigned)P.xlane_cap) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, r);
        return;
    }
    int4* dst = (int4*)on_rank(X, &A.xc[lane * P.xlane_cap + idx], owner);
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 0; i < 6; i++) __stcg(dst + i, src[i]);
    xlane_publish(c, lane, r, kp, packed, owner);
    c.cand = fmin(c.cand, rec.ts);
}
__device__ __forceinline__ void push_credit_remote(Ctx& c, unsigned me, unsigned r, double ts, unsigned src, unsigned seq, unsigned port, unsigned vc, int owner) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    const DfdPeers& X = *c.X;
    if (ts >= P.ts_end) return;
    unsigned kp = port - (unsigned)P.intra_radix;                   // the receiver's own port: its lane
    unsigned my_gport = (unsigned)A.grev[(size_t)r * P.h + kp];     // my end of that link
    size_t lane = ((size_t)c.nxt * P.NR + r) * P.h + kp;
    unsigned packed;
    unsigned idx = xlane_reserve(c, me * P.h + my_gport, true, &packed);
    if (idx >= (unsigned)P.xlane_cap) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, r);
        return;
    }
    CreditRec rec;
    rec.ts = ts;
    rec.src = src;
    rec.seq = seq;
    rec.port = (uint16_t)port;
    rec.vc = (uint8_t)vc;
    rec.pad0 = 0;
    rec.pad1 = 0;
    rec.pad2 = 0;
    int4* dst = (int4*)on_rank(X, &A.xk[lane * P.xlane_cap + idx], owner);
    __stcg(dst, ((const int4*)&rec)[0]);
    __stcg(dst + 1, ((const int4*)&rec)[1]);
    xlane_publish(c, lane, r, kp, packed, owner);
    c.cand = fmin(c.cand, ts);
}
// chunk / credit towards a terminal: single-producer rings (the only sender is the router the terminal
// hangs off); an empty slot holds ts = +inf, the consumer restores that when it pops.
__device__ __forceinline__ void push_chunk_to_node(Ctx& c, unsigned n, const ChunkRec& rec) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (rec.ts >= P.ts_end) return;
    unsigned idx = A.ncin_tail[n]++;
    ChunkRec* slot = &A.ncin[(size_t)n * P.ncin_cap + (idx & (P.ncin_cap - 1))];
#ifdef DFD_CHECKED
    if (d2bits(__ldcg(&slot->ts)) != INF_BITS) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, 0x80000000u | n);
        return;
    }
#endif
    int4* dst = (int4*)slot;
    const int4* src = (const int4*)&rec;
#pragma unroll
    for (int i = 5; i >= 0; i--) __stcg(dst + i, src[i]);
    atomicMin(&A.wake_in[(size_t)c.nxt * P.NT + n], d2bits(rec.ts));
    c.cand = fmin(c.cand, rec.ts);
}
__device__ __forceinline__ void push_credit_to_node(Ctx& c, unsigned n, double ts, unsigned src, unsigned seq) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (ts >= P.ts_end) return;
    unsigned idx = A.nkin_tail[n]++;
    CreditRec* slot = &A.nkin[(size_t)n * P.nkin_cap + (idx & (P.nkin_cap - 1))];
#ifdef DFD_CHECKED
    if (d2bits(__ldcg(&slot->ts)) != INF_BITS) {
        set_error(A, DFD_ERR_INBOX_OVERFLOW, 0xC0000000u | n);
        return;
    }
#endif
    CreditRec rec;
    rec.ts = ts;
    rec.src = src;
    rec.seq = seq;
    rec.port = 0;
    rec.vc = 0;
    rec.pad0 = 0;
    rec.pad1 = 0;
    rec.pad2 = 0;
    __stcg((int4*)slot + 1, ((const int4*)&rec)[1]);
    __stcg((int4*)slot, ((const int4*)&rec)[0]);
    atomicMin(&A.wake_in[(size_t)c.nxt * P.NT + n], d2bits(ts));
    c.cand = fmin(c.cand, ts);
}

// ================================================================================================
// node LP
// ================================================================================================
__device__ __forceinline__ void node_add_event(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double ts, unsigned src, unsigned seq,
                                               unsigned type, unsigned a = 0, unsigned b = 0, unsigned cc = 0, double d = 0.0, unsigned e = 0) {
    if (ts >= c.P->ts_end) return;
    if (s->nev >= (unsigned)c.P->nev_cap) {
        set_error(*c.A, DFD_ERR_NEV_OVERFLOW, n);
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
}

__device__ __forceinline__ double node_cll(const DfdParams& P, NodeState* s, int32_t* rng) {  // codes/codes.h:61-66
    s->rng_draws++;
    return P.lookahead + 0.5 + clcg4_unif(rng) * 0.5;
}

// issue_event synthetic:231-271: next KICKOFF at now + mean_interval
__device__ __forceinline__ void svr_issue_event(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    node_add_event(c, n, s, list, now + P.mean_interval, s->svr_gid, s->svr_seq++, EV_KICKOFF);
}

// handle_kickoff_event synthetic:329-413 + model_net_event_impl_base model-net.c:288-380
__device__ void svr_kickoff(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
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
__device__ void svr_remote(Ctx& c, unsigned n, NodeState* s, double now, unsigned src_svr, double msg_start_time) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (now >= P.warm_up_time) {
        s->msg_recvd_count++;
        double packet_latency = now - msg_start_time;
        s->sum_server_latency += packet_latency;
        if (packet_latency > s->max_server_latency) s->max_server_latency = packet_latency;
        s->svr_seq++;  // tw_event_new for the ACK
        if (now < P.ts_end) {
            const DfdPeers& X = *c.X;
            int owner = X.world > 1 ? owner_of_group(X, (int)fastdiv(fastdiv(src_svr, P.magic_p), P.magic_a)) : X.rank;
            if (owner != X.rank) {
                atomicAdd_system(on_rank(X, &A.ack_count[src_svr], owner), 1u);
                atomicMin_system(on_rank(X, &A.ack_first[src_svr], owner), d2bits(now));
                atomicMax_system(on_rank(X, &A.ack_last[src_svr], owner), d2bits(now));
                c.remote_writes = true;
            } else {
                atomicAdd(&A.ack_count[src_svr], 1u);
                atomicMin(&A.ack_first[src_svr], d2bits(now));
                atomicMax(&A.ack_last[src_svr], d2bits(now));
            }
            COUNT_EV(c, EV_ACK);
        }
    }
}

// dragonfly_dally_packet_event dally:5122-5183 via fcfs_next model-net-sched-impl.c:187-261
__device__ int mn_fcfs_next(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (s->sq_count == 0) return -1;
    SchedReq* q = &A.sq[(size_t)n * P.nsq_cap + (s->sq_head & (P.nsq_cap - 1))];
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
__device__ __forceinline__ void mn_sched_next(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {  // model-net-lp.c:838-873
    if (mn_fcfs_next(c, n, s, list, now) == -1) s->in_sched_send_loop = 0;
}

// handle_new_msg model-net-lp.c:643-802
__device__ void mn_new_msg(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now, const NodeEv& ev) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (ev.a == n) {
        // lp->gid == dest_mn_lp: a message to the sender's own terminal that did not take the no-op shortcut (size >=
        // node_eager_limit, or codes_noop_bypass): the base LP copies it inside the node, model-net-lp.c:672-704 (one
        // node-copy queue); remote event to the destination svr, then the self event to the sender, both from this LP
        double exp_time = (s->node_copy_next_available_time > now) ? s->node_copy_next_available_time : now;
        exp_time += (double)(unsigned long long)ev.b * P.codes_cn_delay;
        exp_time -= now;
        double delay = node_cll(P, s, s->rng_term2);
        s->node_copy_next_available_time = now + exp_time;
        exp_time += delay;
        node_add_event(c, n, s, list, now + exp_time, s->term_gid, s->term_seq++, EV_REMOTE, n, 0, 0, ev.d);
        exp_time += delay;
        node_add_event(c, n, s, list, now + exp_time, s->term_gid, s->term_seq++, EV_LOCAL);
        return;
    }
    if (ev.type & NEV_FLAG_QREQ) {
        double exp_time = (s->next_available_time > now) ? s->next_available_time : now;
        exp_time += P.nic_seq_delay + node_cll(P, s, s->rng_term2);
        s->next_available_time = exp_time;
        node_add_event(c, n, s, list, now + (exp_time - now), s->term_gid, s->term_seq++, EV_NEW_MSG, ev.a, ev.b, 0, ev.d);
        return;
    }
    if (s->sq_count >= (unsigned)P.nsq_cap) {
        set_error(A, DFD_ERR_SQ_OVERFLOW, n);
        return;
    }
    SchedReq* q = &A.sq[(size_t)n * P.nsq_cap + ((s->sq_head + s->sq_count) & (P.nsq_cap - 1))];
    q->dst_svr = ev.a;
    q->msg_size = ev.b;
    q->rem = ev.b;
    q->msg_id = s->msg_id++;
    q->msg_start_time = ev.d;
    s->sq_count++;
    if (s->sq_count > s->sq_hwm) s->sq_hwm = s->sq_count;
    if (s->in_sched_send_loop == 0) {
        s->in_sched_send_loop = 1;
        mn_sched_next(c, n, s, list, now);
    }
}

__device__ __forceinline__ unsigned num_chunks_for(unsigned sz, unsigned chunk) { return sz <= chunk ? 1u : (sz + chunk - 1) / chunk; }  // dally:103-104

// packet_generate dally:5417-5774
__device__ void term_packet_generate(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now, const NodeEv& ev) {
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
            set_error(A, DFD_ERR_TQ_OVERFLOW, n);
            return;
        }
        TermChunk* t = &A.tq[(size_t)n * P.ntq_cap + ((s->tq_head + s->tq_count) & (P.ntq_cap - 1))];
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
__device__ void term_packet_send(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (s->tq_count == 0 || s->vc_occupancy + P.chunk_size > P.cn_vc) {
        s->in_send_loop = 0;
        s->stalled_chunks++;
        if (!s->last_buf_full) s->last_buf_full = now;
        return;
    }
    TermChunk cur = A.tq[(size_t)n * P.ntq_cap + (s->tq_head & (P.ntq_cap - 1))];
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
        push_chunk_to_router(c, s->router, rec);
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
__device__ __forceinline__ void term_buf_update(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now) {
    const DfdParams& P = *c.P;
    s->vc_occupancy -= P.chunk_size;
    if (s->in_send_loop == 0 && s->tq_count != 0) {
        node_add_event(c, n, s, list, now + 0.0, s->term_gid, s->term_seq++, EV_T_SEND);
        s->in_send_loop = 1;
    }
}

// packet_arrive dally:6453-6744.  Packets of one chunk that are a whole message (the BASELINE workloads) never touch
// the reassembly tables: their rank_tbl entry would be created and retired inside this one call.
__device__ void term_packet_arrive(Ctx& c, unsigned n, NodeState* s, NodeEv* list, double now, const ChunkRec& m) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    if (m.dst_term != n) {
        set_error(A, DFD_ERR_WRONG_TERMINAL, n);
        return;
    }
    unsigned tg = s->term_gid;
    push_credit_to_router(c, m.prev_lp, now + P.cn_credit, tg, s->term_seq++, m.prev_port, m.prev_vc);
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
        uint4* pt = &A.ptab[(size_t)n * P.nreasm_cap];
        unsigned pi = 0;
        for (; pi < s->pt_n; pi++)
            if (pt[pi].x == m.packet_id && pt[pi].y == m.src_term) break;
        if (pi < s->pt_n) {
            unsigned actual = min((unsigned)P.chunk_size, pt[pi].z);
            pt[pi].z -= actual;
            if (pt[pi].z == 0) {
                packet_done = true;
                pt[pi] = pt[--s->pt_n];
            }
        } else if ((unsigned)P.chunk_size < m.packet_size) {
            if (s->pt_n >= (unsigned)P.nreasm_cap) {
                set_error(A, DFD_ERR_REASM_OVERFLOW, n);
                return;
            }
            pt[s->pt_n++] = make_uint4(m.packet_id, m.src_term, m.packet_size - P.chunk_size, 0);
        } else
            packet_done = true;
        // rank_tbl (dally:6620-6720): packets still missing of the message, keyed by (message_id, sender)
        uint4* mt = &A.mtab[(size_t)n * P.nreasm_cap];
        unsigned mi = 0;
        for (; mi < s->mt_n; mi++)
            if (mt[mi].x == m.message_id && mt[mi].y == m.src_svr) break;
        if (mi == s->mt_n) {
            if (s->mt_n >= (unsigned)P.nreasm_cap) {
                set_error(A, DFD_ERR_REASM_OVERFLOW, n);
                return;
            }
            unsigned total_packets = m.total_size / (unsigned)P.packet_size + (m.total_size % (unsigned)P.packet_size ? 1 : 0);
            if (total_packets == 0) total_packets = 1;
            mt[s->mt_n++] = make_uint4(m.message_id, m.src_svr, total_packets, 0);
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

__device__ void node_process(Ctx& c, unsigned n) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
#ifdef DFD_PROFILE
    long long prof_lp0 = clock64();
#endif
    NodeState* s = &A.node[n];
#ifdef DFD_PROFILE
    unsigned long long ev0 = s->events;
#endif
    NodeEv* list = &A.nev[(size_t)n * P.nev_cap];
    const unsigned cmask = P.ncin_cap - 1, kmask = P.nkin_cap - 1;
    double next_ts;
    // heads of the two single-producer rings: read once, re-read only after a pop (a record written during
    // this window is not due before the next one)
    ChunkRec* ch = &A.ncin[(size_t)n * P.ncin_cap + (s->cin_head & cmask)];
    CreditRec* kr = &A.nkin[(size_t)n * P.nkin_cap + (s->kin_head & kmask)];
    int4 chh = __ldcg((const int4*)ch);
    int4 krh = __ldcg((const int4*)kr);
    for (;;) {
        // earliest self event
        int bi = -1;
        double bts = DINF;
        unsigned bsrc = 0, bseq = 0;
        int kind = 0;  // 0 self, 1 chunk ring, 2 credit ring
        for (unsigned i = 0; i < s->nev; i++) {
            const NodeEv& e = list[i];
            if (bi < 0 || key_less(e.ts, e.src, e.seq, bts, bsrc, bseq)) {
                bi = (int)i;
                bts = e.ts;
                bsrc = e.src;
                bseq = e.seq;
            }
        }
        double cts = __longlong_as_double(((long long)(unsigned)chh.y << 32) | (unsigned)chh.x);
        if (d2bits(cts) != INF_BITS && (bts == DINF || key_less(cts, (unsigned)chh.z, (unsigned)chh.w, bts, bsrc, bseq))) {
            kind = 1;
            bts = cts;
            bsrc = (unsigned)chh.z;
            bseq = (unsigned)chh.w;
        }
        double kts = __longlong_as_double(((long long)(unsigned)krh.y << 32) | (unsigned)krh.x);
        if (d2bits(kts) != INF_BITS && (bts == DINF || key_less(kts, (unsigned)krh.z, (unsigned)krh.w, bts, bsrc, bseq))) {
            kind = 2;
            bts = kts;
            bsrc = (unsigned)krh.z;
            bseq = (unsigned)krh.w;
        }
        if (!(bts < c.t_end)) {
            next_ts = bts;
            break;
        }
        double now = bts;
        s->events++;
        PROF_T0();
        if (kind == 1) {
            ChunkRec m;
            int4* dst = (int4*)&m;
            dst[0] = chh;
#pragma unroll
            for (int i = 1; i < 6; i++) dst[i] = __ldcg((const int4*)ch + i);
            __stcg((unsigned long long*)&ch->ts, INF_BITS);
            s->cin_head++;
            ch = &A.ncin[(size_t)n * P.ncin_cap + (s->cin_head & cmask)];
            chh = __ldcg((const int4*)ch);
            COUNT_EV(c, EV_T_ARRIVE);
            term_packet_arrive(c, n, s, list, now, m);
            PROF_ADD(c, EV_T_ARRIVE);
        } else if (kind == 2) {
            __stcg((unsigned long long*)&kr->ts, INF_BITS);
            s->kin_head++;
            kr = &A.nkin[(size_t)n * P.nkin_cap + (s->kin_head & kmask)];
            krh = __ldcg((const int4*)kr);
            COUNT_EV(c, EV_T_BUFFER);
            term_buf_update(c, n, s, list, now);
            PROF_ADD(c, EV_T_BUFFER);
        } else {
            NodeEv ev = list[bi];
            list[bi] = list[s->nev - 1];
            s->nev--;
            unsigned type = ev.type & 0xFF;
            COUNT_EV(c, type);
            switch (type) {
            case EV_KICKOFF: svr_kickoff(c, n, s, list, now); break;
            case EV_REMOTE: svr_remote(c, n, s, now, ev.a, ev.d); break;
            case EV_LOCAL:  // handle_local_event synthetic:508-516
                if (now >= P.warm_up_time) s->local_recvd_count++;
                break;
            case EV_NEW_MSG: mn_new_msg(c, n, s, list, now, ev); break;
            case EV_SCHED_NEXT: mn_sched_next(c, n, s, list, now); break;
            case EV_T_GENERATE: term_packet_generate(c, n, s, list, now, ev); break;
            case EV_T_SEND: term_packet_send(c, n, s, list, now); break;
            default: break;
            }
            PROF_ADD(c, type);
        }
    }
    A.wake_self[n] = d2bits(next_ts);
    c.cand = fmin(c.cand, next_ts);
#ifdef DFD_PROFILE
    {
        unsigned long long dt = (unsigned long long)(clock64() - prof_lp0);
        atomicAdd(&A.counters[48], dt);
        atomicAdd(&A.counters[49], 1ULL);
        atomicMax(&A.winmax[c.widx & 0xFFFF], (dt << 16) | ((unsigned long long)min(255u, (unsigned)(s->events - ev0)) << 8) | 0ULL);
    }
#endif
}

// ================================================================================================
// router LP
// ================================================================================================
struct RCtx {
    unsigned r;       // router id
    unsigned gid;
    int lid, grp;
    RouterHdr* h;
    PortState* port;
    PortQ* pq;
    PoolSlot* pool;
    // one pending event that is due in this window stays in registers instead of going through the wheel (the
    // R_SEND that follows an arrival at the same timestamp, a credit drained in the window it is due in)
    double hot_ts;
    unsigned hot_src, hot_seq, hot_ref;  // hot_ref == 0: empty
    unsigned* pfree;            // stack of free pool-slot indices
    unsigned* efree;            // stack of free wheel-entry indices
    PendEnt* pent;              // wheel entries
    unsigned* wheel;            // bucket heads
    unsigned long long* wbits;  // non-empty-bucket bitmap
    long long b_beg;            // absolute bucket index of the window start
};

// Pending events of a router: timing wheel, bucket width = one lookahead window (see PendEnt).
__device__ __forceinline__ long long wheel_bucket(const DfdParams& P, double ts) { return (long long)(ts * P.inv_window); }

// earliest timestamp in the first non-empty bucket at or after bucket b0 (bucket index is monotone in the
// timestamp, so no later bucket can hold an earlier event); +inf when there is none
__device__ __forceinline__ double rwheel_min_from(Ctx& c, RCtx& R, long long b0) {
    const DfdParams& P = *c.P;
    if (R.h->pend_n == 0) return DINF;
    const unsigned nb = (unsigned)P.wheel_nb, mask = nb - 1, nwords = nb >> 6;
    unsigned start = (unsigned)b0 & mask;
    unsigned w = start >> 6;
    unsigned long long bits = R.wbits[w] & (~0ULL << (start & 63));
    for (unsigned k = 0; k <= nwords; k++) {
        if (bits) {
            unsigned bi = (w << 6) + (unsigned)__ffsll((long long)bits) - 1;
            double m = DINF;
            for (unsigned e = R.wheel[bi]; e != NIL; e = R.pent[e].next) m = fmin(m, R.pent[e].ts);
            return m;
        }
        w = (w + 1) & (nwords - 1);
        bits = R.wbits[w];
        if (k + 1 == nwords) bits &= ~(~0ULL << (start & 63));  // wrapped around to the first word: the part before start
    }
    return DINF;
}

__device__ __forceinline__ void rheap_push(Ctx& c, RCtx& R, double ts, unsigned src, unsigned seq, unsigned ref) {
    const DfdParams& P = *c.P;
    if (ts >= P.ts_end) return;
    if (R.hot_ref == 0 && ts < c.t_end) {
        R.hot_ts = ts;
        R.hot_src = src;
        R.hot_seq = seq;
        R.hot_ref = ref;
        return;
    }
    long long b = wheel_bucket(P, ts);
    if (R.h->ent_nfree == 0 || b - R.b_beg >= (long long)P.wheel_nb || b < R.b_beg) {
        // info: router | (bucket distance from the window start, clamped to +-2^23) << 32 | event kind << 60
        long long dist = b - R.b_beg;
        dist = dist > 8388607 ? 8388607 : dist < -8388607 ? -8388607 : dist;
        set_error(*c.A, R.h->ent_nfree == 0 ? DFD_ERR_HEAP_OVERFLOW : DFD_ERR_WHEEL_RANGE,
                  (unsigned long long)R.r | ((unsigned long long)(dist & 0xFFFFFF) << 32) | ((unsigned long long)(ref >> 28) << 60));
        return;
    }
    unsigned e = R.efree[--R.h->ent_nfree];
    PendEnt* pe = &R.pent[e];
    unsigned bi = (unsigned)b & (unsigned)(P.wheel_nb - 1);
    unsigned head = R.wheel[bi];
    pe->ts = ts;
    pe->src = src;
    pe->seq = seq;
    pe->ref = ref;
    pe->next = head;
    R.wheel[bi] = e;
    if (head == NIL) R.wbits[bi >> 6] |= 1ULL << (bi & 63);
    R.h->pend_n++;
    R.h->wmin = fmin(R.h->wmin, ts);
}

// earliest pending event with ts < t_end, by (ts, src, seq): the register-held entry or, when the wheel's
// earliest timestamp says something in it is due, the best entry of the buckets the window overlaps.  Removes
// it and returns its payload reference (0 when nothing is due).
__device__ __forceinline__ unsigned rwheel_pop_due(Ctx& c, RCtx& R, double* now_out) {
    const DfdParams& P = *c.P;
    if (!(R.h->wmin < c.t_end)) {  // nothing in the wheel is due
        unsigned ref = R.hot_ref;
        *now_out = R.hot_ts;
        R.hot_ref = 0;
        return ref;
    }
    const unsigned mask = (unsigned)(P.wheel_nb - 1);
    const long long b_end = wheel_bucket(P, c.t_end);
    unsigned best = NIL, best_prev = NIL, best_b = 0;
    double bts = DINF, second = DINF;  // second: earliest timestamp among the scanned entries other than best
    unsigned bsrc = 0, bseq = 0;
    for (long long b = wheel_bucket(P, R.h->wmin); b <= b_end; b++) {
        unsigned bi = (unsigned)b & mask;
        unsigned prev = NIL;
        for (unsigned e = R.wheel[bi]; e != NIL; prev = e, e = R.pent[e].next) {
            const PendEnt& pe = R.pent[e];
            if (best == NIL || key_less(pe.ts, pe.src, pe.seq, bts, bsrc, bseq)) {
                second = bts;
                best = e;
                best_prev = prev;
                best_b = bi;
                bts = pe.ts;
                bsrc = pe.src;
                bseq = pe.seq;
            } else
                second = fmin(second, pe.ts);
        }
    }
    if (R.hot_ref && key_less(R.hot_ts, R.hot_src, R.hot_seq, bts, bsrc, bseq)) {
        unsigned ref = R.hot_ref;
        *now_out = R.hot_ts;
        R.hot_ref = 0;
        return ref;
    }
    PendEnt* pe = &R.pent[best];
    unsigned nxt = pe->next, ref = pe->ref;
    if (best_prev == NIL) {
        R.wheel[best_b] = nxt;
        if (nxt == NIL) R.wbits[best_b >> 6] &= ~(1ULL << (best_b & 63));
    } else
        R.pent[best_prev].next = nxt;
    R.efree[R.h->ent_nfree++] = best;
    R.h->pend_n--;
    R.h->wmin = second < DINF ? second : rwheel_min_from(c, R, b_end + 1);
    *now_out = bts;
    return ref;
}

// --- connection queries (DragonflyConnectionManager, dragonfly-network-manager.cxx) ----------------
// local ports from my local id to router-local id `lid` (get_connections_to_gid(.., CONN_LOCAL)): file order
__device__ __forceinline__ void loc_range(const DfdParams& P, const DfdArrays& A, int mylid, int lid, int* lo, int* hi) {
    *lo = A.loc_off[mylid * P.a + lid];
    *hi = A.loc_off[mylid * P.a + lid + 1];
}
// direct global connections to group g (get_connections_to_group): contiguous range of the by-dest-gid list
__device__ __forceinline__ void grp_range(const DfdParams& P, const DfdArrays& A, unsigned r, int g, int* lo, int* hi) {
    const int* gg = A.gs_group + (size_t)r * P.h;
    int l = 0;
    while (l < P.h && gg[l] < g) l++;
    int u = l;
    while (u < P.h && gg[u] == g) u++;
    *lo = l;
    *hi = u;
}

// a candidate list of next-hop connections, never materialised:
//  kind 0: empty; 1: direct global links [lo,hi) of the by-gid list; 2: "routed" list towards group g
//  (for each global link my group -> g in inter-file order: all my local links to its owner,
//  get_routed_connections_to_group(g, true) network-manager.cxx:772-814); 3: local links to lid lo
struct Cands {
    int kind, lo, hi, n;
};
__device__ __forceinline__ int routed_count(const DfdParams& P, const DfdArrays& A, const RCtx& R, int g) {
    int o0 = A.routed_off[R.grp * P.G + g], o1 = A.routed_off[R.grp * P.G + g + 1];
    int n = 0;
    for (int i = o0; i < o1; i++) {
        int lid = A.routed_lid[i];
        n += A.loc_off[R.lid * P.a + lid + 1] - A.loc_off[R.lid * P.a + lid];
    }
    return n;
}
__device__ __forceinline__ int cand_port(const DfdParams& P, const DfdArrays& A, const RCtx& R, const Cands& cd, int idx) {
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
__device__ __forceinline__ Cands legal_minimal_stops(const DfdParams& P, const DfdArrays& A, const RCtx& R, int fdest_router, int fg) {
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

__device__ __forceinline__ int score_port(const DfdParams& P, const RCtx& R, int port, bool nonmin) {  // dally:1649-1684
    if (port < 0) return 2147483647;
    const PortState& ps = R.port[port];
    int score = ps.vc_occ[0] + ps.vc_occ[1] + ps.vc_occ[2] + ps.vc_occ[3] + ps.queued_count;
    if (P.scoring == DFD_SCORE_DELTA && nonmin) score = score * 2;
    return score;
}
__device__ __forceinline__ long long r_integer(RCtx& R, long long lo, long long hi) { return clcg4_integer(R.h->rng, lo, hi, &R.h->rng_draws); }

// get_absolute_best_connection_from_conns dally:1687-1723 over a candidate list
__device__ __forceinline__ int best_from_cands(const DfdParams& P, const DfdArrays& A, RCtx& R, const Cands& cd) {
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
__device__ __forceinline__ int best_from_k2(const DfdParams& P, const DfdArrays& A, RCtx& R, const Cands& cd) {
    if (cd.n == 0) return -1;
    if (cd.n == 1) return cand_port(P, A, R, cd, 0);
    long long n = cd.n;
    long long s1 = r_integer(R, 0, n - 1);
    long long off = r_integer(R, 1, n - 1);
    long long s2 = (s1 + off) % n;
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
__device__ int best_from_k(Ctx& c, const DfdParams& P, const DfdArrays& A, RCtx& R, const Cands& cd, int k) {
    if (cd.n == 0) return -1;
    if (cd.n == 1) return cand_port(P, A, R, cd, 0);
    if (k > cd.n || cd.n > 64 || k < 1) {
        set_error(A, DFD_ERR_DEAD_END, 0x4000000000000000ULL | R.r);
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
        int sc = score_port(P, R, cand_port(P, A, R, cd, __ffsll((long long)m) - 1), false);
        if (sc < best_score) {
            best_score = sc;
            nbest = 1;
        } else if (sc == best_score)
            nbest++;
    }
    int pick = (int)r_integer(R, 0, nbest - 1);
    for (unsigned long long m = sel; m; m &= m - 1) {
        int port = cand_port(P, A, R, cd, __ffsll((long long)m) - 1);
        if (score_port(P, R, port, false) == best_score) {
            if (pick == 0) return port;
            pick--;
        }
    }
    return -1;
}

// dfdally_select_intermediate_group dally:9738-9790
__device__ __forceinline__ void select_intermediate_group(const DfdParams& P, const DfdArrays& A, RCtx& R, PoolSlot& m) {
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
__device__ __forceinline__ Cands legal_nonminimal_stops(const DfdParams& P, const DfdArrays& A, RCtx& R, PoolSlot& m) {
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
__device__ __forceinline__ int do_routing(Ctx& c, RCtx& R, PoolSlot& m, int fdest_router) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    int fg = m.dst_grp;
    if ((int)R.r == fdest_router) {  // destination router: the terminal port (one rail -> no draw)
        return P.intra_radix + P.global_radix + (int)(m.dst_term - (unsigned)fdest_router * P.p);
    } else if (R.grp == fg) {  // destination group: local link(s) to the destination router
        Cands cd;
        loc_range(P, A, R.lid, fdest_router - fg * P.a, &cd.lo, &cd.hi);
        cd.kind = 3;
        cd.n = cd.hi - cd.lo;
        if (cd.n < 1) {
            set_error(A, DFD_ERR_DEAD_END, R.r);
            return -1;
        }
        if (P.routing == DFD_PROG_ADAPTIVE) return best_from_cands(P, A, R, cd);
        return cand_port(P, A, R, cd, (int)r_integer(R, 0, cd.n - 1));
    }
    if (m.last_hop == DFD_HOP_TERMINAL && m.intm_grp_id == -1) select_intermediate_group(P, A, R, m);
    if (P.routing == DFD_MINIMAL) {  // dfdally_minimal_routing dally:9913-9932
        Cands cd = legal_minimal_stops(P, A, R, fdest_router, fg);
        if (cd.n < 1) {
            set_error(A, DFD_ERR_DEAD_END, R.r);
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
            set_error(A, DFD_ERR_DEAD_END, R.r);
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
        best_min = P.global_k_picks == 2 ? best_from_k2(P, A, R, mins) : best_from_k(c, P, A, R, mins, P.global_k_picks);
    else
        best_min = best_from_cands(P, A, R, mins);
    if (nonmins.n > 0 && nonmins.kind == 1)
        best_non = P.global_k_picks == 2 ? best_from_k2(P, A, R, nonmins) : best_from_k(c, P, A, R, nonmins, P.global_k_picks);
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
template <typename T> __device__ __forceinline__ T get4(const T* a, int i) { return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : a[3]; }
template <typename T> __device__ __forceinline__ void set4(T* a, int i, T v) {
    if (i == 0) a[0] = v;
    else if (i == 1) a[1] = v;
    else if (i == 2) a[2] = v;
    else a[3] = v;
}
// whole-struct copies as 128-bit accesses (all of these structs are 16-byte aligned multiples of 16 bytes)
template <typename T> __device__ __forceinline__ void load_vec(T& dst, const T* src) {
    static_assert(sizeof(T) % 16 == 0, "vector copy");
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) ((int4*)&dst)[i] = ((const int4*)src)[i];
}
template <typename T> __device__ __forceinline__ void store_vec(T* dst, const T& src) {
#pragma unroll
    for (unsigned i = 0; i < sizeof(T) / 16; i++) ((int4*)dst)[i] = ((const int4*)&src)[i];
}

// router_credit_send dally:7269-7326
__device__ __forceinline__ void router_credit_send(Ctx& c, RCtx& R, double now, unsigned last_hop, unsigned prev_lp, unsigned port, unsigned vc) {
    const DfdParams& P = *c.P;
    unsigned seq = R.h->seq++;
    if (last_hop == DFD_HOP_TERMINAL)
        push_credit_to_node(c, prev_lp, now + P.cn_credit, R.gid, seq);
    else if (last_hop == DFD_HOP_GLOBAL) {
        const DfdPeers& X = *c.X;
        int owner = X.world > 1 ? owner_of_group(X, (int)fastdiv(prev_lp, P.magic_a)) : X.rank;
        if (owner != X.rank)
            push_credit_remote(c, R.r, prev_lp, now + P.global_credit, R.gid, seq, port, vc, owner);
        else
            push_credit_to_router(c, prev_lp, now + P.global_credit, R.gid, seq, port, vc);
    } else
        push_credit_to_router(c, prev_lp, now + P.local_credit, R.gid, seq, port, vc);
}

__device__ __forceinline__ double router_send_core(Ctx& c, RCtx& R, double now, int output_port, PortState& ps, const PoolSlot& m);

// router_packet_receive dally:7377-7595; the chunk already sits in pool slot `slot`
__device__ __forceinline__ void router_packet_receive(Ctx& c, RCtx& R, double now, unsigned slot) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
    PoolSlot m;  // the chunk: read once, routed in registers, written back once
    load_vec(m, &R.pool[slot]);
    int dest_router_id = (int)m.dst_router;
    int in_path_type = m.path_type;  // msg->path_type (as received)
    unsigned in_port = m.prev_port, in_vc = m.prev_vc, in_last_hop = m.last_hop, in_prev = m.prev_lp;
    if (m.last_hop == DFD_HOP_TERMINAL) m.path_type = DFD_MINIMAL;
    int output_port = do_routing(c, R, m, dest_router_id);
    if (output_port < 0 || output_port >= P.radix) {
        set_error(A, DFD_ERR_DEAD_END, R.r);
        return;
    }
    PSTAMP(c, 1);  // chunk loaded + routed
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
        set_error(A, DFD_ERR_BAD_VC, R.r);
        return;
    }
    m.out_vc = (uint8_t)output_chan;
    m.n_hop++;
    m.next = NIL;
    int occ = get4(ps.vc_occ, output_chan);
    if (occ + P.chunk_size <= max_vc_size) {
        PSTAMP(c, 2);  // port state loaded, next stop known
        router_credit_send(c, R, now, in_last_hop, in_prev, in_port, in_vc);
        PSTAMP(c, 3);  // credit sent
        set4(ps.vc_occ, output_chan, occ + P.chunk_size);
        if (ps.in_send_loop == 0) {
            double ts = maxd(ps.noat, now) - now;
            unsigned send_seq = R.h->seq++;  // the R_SEND event
            if (ts == 0.0 && R.hot_ref == 0 && R.h->wmin > now) {
                // The port is idle (so all of its pending lists are empty) and nothing else of this router is due at
                // or before `now`: the R_SEND scheduled here is the very next event this router executes, and it
                // sends exactly this chunk.  Run it on the state already in registers: no list append / pop, no
                // slot write-back, no trip through the pending-event structure.
                COUNT_EV(c, EV_R_SEND);
                ps.last_qos_lvl = (uint8_t)output_chan;
                router_send_core(c, R, now, output_port, ps, m);
                R.pfree[R.h->pool_nfree++] = slot;
                store_vec(&R.port[output_port], ps);
                PSTAMP(c, 4);
                return;
            }
            rheap_push(c, R, now + ts, R.gid, send_seq, (RK_SEND << 28) | (unsigned)output_port);
            ps.in_send_loop = 1;
        }
        unsigned tl = get4(pq.pend_tail, output_chan);  // append to pending_msgs[port][vc]
        if (tl == NIL)
            set4(pq.pend_head, output_chan, slot);
        else
            R.pool[tl].next = slot;
        set4(pq.pend_tail, output_chan, slot);
        PSTAMP(c, 4);  // queued + R_SEND scheduled
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
    }
    store_vec(&R.pool[slot], m);
    store_vec(&R.port[output_port], ps);
    store_vec(&R.pq[output_port], pq);
}

// router_packet_send dally:7705-8032 (+ get_next_router_vcg :3499-3557)
// The part of router_packet_send that moves one chunk out of `output_port`: serialisation delay, the record for the
// next hop, link statistics.  `m` is the chunk (pool-slot layout); returns the injection time without the router
// delay (when the port can start its next chunk, relative to now).
__device__ __forceinline__ double router_send_core(Ctx& c, RCtx& R, double now, int output_port, PortState& ps, const PoolSlot& m) {
    const DfdParams& P = *c.P;
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
        // outgoing record = the pool slot with a new key and the "previous hop" fields replaced; the two
        // layouts agree from byte 16 on
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
            push_chunk_to_node(c, m.next_stop, u.rec);
        else {
            const DfdPeers& X = *c.X;
            int owner = (global && X.world > 1) ? owner_of_group(X, (int)fastdiv(m.next_stop, P.magic_a)) : X.rank;
            if (owner != X.rank)
                push_chunk_remote(c, R.r, (unsigned)(output_port - P.intra_radix), m.next_stop, u.rec, owner);
            else
                push_chunk_to_router(c, m.next_stop, u.rec);
        }
    }
    if ((tail || m.packet_size == 0) && (m.chunk_id == nchunks - 1))
        ps.link_traffic += tail;
    else
        ps.link_traffic += P.chunk_size;
    ps.total_chunks++;
    ps.noat -= P.router_delay;
    return injection_ts - P.router_delay;
}

__device__ __forceinline__ void router_packet_send(Ctx& c, RCtx& R, double now, int output_port) {
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
    double injection_ts = router_send_core(c, R, now, output_port, ps, m);
    set4(pq.pend_head, output_chan, m.next);  // pop pending_msgs[port][vc]
    if (m.next == NIL) set4(pq.pend_tail, output_chan, NIL);
    R.pfree[R.h->pool_nfree++] = slot;  // free the slot
    bool more = (pq.pend_head[0] != NIL) | (pq.pend_head[1] != NIL) | (pq.pend_head[2] != NIL) | (pq.pend_head[3] != NIL);
    if (!more)
        ps.in_send_loop = 0;
    else
        rheap_push(c, R, now + injection_ts, R.gid, R.h->seq++, (RK_SEND << 28) | (unsigned)output_port);
    store_vec(&R.port[output_port], ps);
    store_vec(&R.pq[output_port], pq);
}

// router_buf_update dally:8069-8145
__device__ __forceinline__ void router_buf_update(Ctx& c, RCtx& R, double now, int indx, int output_chan) {
    const DfdParams& P = *c.P;
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
        const PoolSlot* hd = &R.pool[qh];
        int4 v0 = ((const int4*)hd)[0];
        int4 v3 = ((const int4*)hd)[3];
        unsigned nxt = (unsigned)v0.x;                        // next
        unsigned prev_lp = (unsigned)v3.x;                    // prev_lp
        unsigned prev_port = (unsigned)v3.y & 0xFFFFu;        // prev_port
        unsigned prev_vc = ((unsigned)v3.y >> 16) & 0xFFu;    // prev_vc
        unsigned last_hop = ((unsigned)v3.y >> 24) & 0xFFu;   // last_hop
        set4(pq.queue_head, output_chan, nxt);
        if (nxt == NIL) set4(pq.queue_tail, output_chan, NIL);
        router_credit_send(c, R, now, last_hop, prev_lp, prev_port, prev_vc);
        R.pool[qh].next = NIL;
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
        rheap_push(c, R, now + ts, R.gid, R.h->seq++, (RK_SEND << 28) | (unsigned)indx);
        ps.in_send_loop = 1;
    }
    store_vec(&R.port[indx], ps);
    store_vec(&R.pq[indx], pq);
}

// inbox records -> pool slots / wheel entries
__device__ __forceinline__ void drain_chunks(Ctx& c, RCtx& R, const ChunkRec* in, unsigned n) {
    for (unsigned i = 0; i < n; i++) {
        if (R.h->pool_nfree == 0) {
            set_error(*c.A, DFD_ERR_POOL_OVERFLOW, R.r);
            break;
        }
        unsigned slot = R.pfree[--R.h->pool_nfree];
        int4 v0 = __ldcg((const int4*)&in[i]);
        int4* dst = (int4*)&R.pool[slot];
#pragma unroll
        for (int k = 1; k < 6; k++) dst[k] = __ldcg((const int4*)&in[i] + k);
        double ts = __longlong_as_double(((long long)(unsigned)v0.y << 32) | (unsigned)v0.x);
        rheap_push(c, R, ts, (unsigned)v0.z, (unsigned)v0.w, (RK_CHUNK << 28) | slot);
    }
}
__device__ __forceinline__ void drain_credits(Ctx& c, RCtx& R, const CreditRec* in, unsigned n) {
    for (unsigned i = 0; i < n; i++) {
        int4 v0 = __ldcg((const int4*)&in[i]);
        int4 v1 = __ldcg((const int4*)&in[i] + 1);
        double ts = __longlong_as_double(((long long)(unsigned)v0.y << 32) | (unsigned)v0.x);
        unsigned port = (unsigned)v1.x & 0xFFFFu, vc = ((unsigned)v1.x >> 16) & 0xFFu;
        rheap_push(c, R, ts, (unsigned)v0.z, (unsigned)v0.w, (RK_CREDIT << 28) | (vc << 16) | port);
    }
}

__device__ void router_process(Ctx& c, unsigned r, unsigned ntail_c, unsigned ntail_k) {
    const DfdParams& P = *c.P;
    const DfdArrays& A = *c.A;
#ifdef DFD_PROFILE
    long long prof_lp0 = clock64();
#endif
    RCtx R;
    R.r = r;
    RouterHdr hd = A.rh[r];  // header lives in registers for the whole activation, written back once at the end
    R.h = &hd;
    R.gid = R.h->gid;
    R.lid = R.h->lid;
    R.grp = R.h->grp;
    R.port = &A.port[(size_t)r * P.radix];
    R.pq = &A.portq[(size_t)r * P.radix];
    R.pool = &A.pool[(size_t)r * P.rpool_cap];
    R.pent = &A.pent[(size_t)r * P.rheap_cap];
    R.pfree = &A.pfree[(size_t)r * P.rpool_cap];
    R.efree = &A.efree[(size_t)r * P.rheap_cap];
    R.wheel = &A.wheel[(size_t)r * P.wheel_nb];
    R.wbits = &A.wbits[(size_t)r * (P.wheel_nb >> 6)];
    R.b_beg = c.b_beg;
    R.hot_ref = 0;
    // drain the inbox filled during the previous window: chunks go to pool slots, keys to the heap
#ifdef DFD_PROFILE
    long long pd0 = clock64();
#endif
    if (ntail_c) {
        drain_chunks(c, R, &A.rcin[((size_t)c.cur * P.NR + r) * P.rcin_cap], ntail_c);
        A.rcin_tail[(size_t)c.cur * P.NR + r] = 0;
    }
    if (ntail_k) {
        drain_credits(c, R, &A.rkin[((size_t)c.cur * P.NR + r) * P.rkin_cap], ntail_k);
        A.rkin_tail[(size_t)c.cur * P.NR + r] = 0;
    }
    if (c.X->world > 1) {  // lanes of the global links that come from other GPUs
        unsigned long long xm = __ldcg(&A.xmask[(size_t)c.cur * P.NR + r]);
        if (xm) {
            A.xmask[(size_t)c.cur * P.NR + r] = 0;
            do {
                unsigned kp = (unsigned)__ffsll((long long)xm) - 1u;
                xm &= xm - 1;
                size_t lane = ((size_t)c.cur * P.NR + r) * P.h + kp;
                unsigned cnt = __ldcg(&A.xcnt[lane]);
                A.xcnt[lane] = 0;
                if (cnt & 0xFFFFu) drain_chunks(c, R, &A.xc[lane * P.xlane_cap], cnt & 0xFFFFu);
                if (cnt >> 16) drain_credits(c, R, &A.xk[lane * P.xlane_cap], cnt >> 16);
            } while (xm);
        }
    }
#ifdef DFD_PROFILE
    if (ntail_c | ntail_k) {
        atomicAdd(&A.counters[56], (unsigned long long)(clock64() - pd0));
        atomicAdd(&A.counters[57], (unsigned long long)(ntail_c + ntail_k));
    }
#endif
#ifdef DFD_PROFILE
    unsigned prof_nev = 0;
#endif
#ifdef DFD_PROFILE2
    long long p2_t0 = clock64();
    unsigned p2_seq = 0, p2_n = 0;
#endif
    for (;;) {
        double now;
        PSTAMP0(c);
        unsigned ref = rwheel_pop_due(c, R, &now);
        if (ref == 0) break;
#ifdef DFD_PROFILE
        prof_nev++;
#endif
        unsigned kind = ref >> 28;
#ifdef DFD_PROFILE2
        if (p2_n < 10) p2_seq |= kind << (2 * p2_n);
        p2_n++;
#endif
        PROF_T0();
        if (kind == RK_CHUNK) {
            COUNT_EV(c, EV_R_ARRIVE);
            PSTAMP(c, 0);  // pop
            router_packet_receive(c, R, now, ref & 0x0FFFFFFFu);
            PSTAMP(c, 5);  // state written back
            PROF_ADD(c, EV_R_ARRIVE);
        } else if (kind == RK_CREDIT) {
            COUNT_EV(c, EV_R_BUFFER);
            router_buf_update(c, R, now, ref & 0xFFFFu, (ref >> 16) & 0xFFu);
            PROF_ADD(c, EV_R_BUFFER);
        } else {
            COUNT_EV(c, EV_R_SEND);
            router_packet_send(c, R, now, ref & 0xFFFFu);
            PROF_ADD(c, EV_R_SEND);
        }
    }
    double next_ts = hd.wmin;
    A.rh[r] = hd;
    A.wake_self[P.NT + r] = d2bits(next_ts);
    c.cand = fmin(c.cand, next_ts);
#ifdef DFD_PROFILE2
    {   // slowest activation of the window: cycles (from after the drain) | drained chunks | credits | events | event kinds
        unsigned long long dt = (unsigned long long)(clock64() - p2_t0);
        atomicMax(&A.winmax[c.widx & 0xFFFF], (dt << 32) | ((unsigned long long)min(15u, ntail_c) << 28) | ((unsigned long long)min(15u, ntail_k) << 24) |
                                                  ((unsigned long long)min(15u, p2_n) << 20) | (p2_seq & 0xFFFFFu));
    }
#endif
#ifdef DFD_PROFILE
    {
        unsigned long long dt = (unsigned long long)(clock64() - prof_lp0);
        atomicAdd(&A.counters[50], dt);
        atomicAdd(&A.counters[51], 1ULL);
        if (ntail_c | ntail_k) atomicAdd(&A.counters[52], 1ULL);
        atomicMax(&A.winmax[c.widx & 0xFFFF], (dt << 16) | ((unsigned long long)min(255u, prof_nev) << 8) | ((ntail_c | ntail_k) ? 3ULL : 1ULL));
    }
#endif
}

// ================================================================================================
// the window loop
// ================================================================================================
// Grid barrier that keeps L1 valid: the arrival is a release (MEMBAR.ALL.GPU + RED, no CCTL.IVALL), the wait
// is a relaxed poll.  Everything another LP may have written is read with ld.cg (L2), so no acquire -- and
// therefore no L1 invalidation -- is needed; LP-private state stays L1-resident across windows because an LP
// is always processed by the same block.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// Each block owns a contiguous slice of the nodes and of the routers.  Per window: (1) a coalesced scan of the
// slice's wake words compacts the LPs that have work into a shared-memory list, (2) the list is dealt out
// warp-first (lane 0 of every warp, then lane 1, ...) so that concurrently active LPs do not serialise inside
// one warp, (3) block-min of the next timestamps -> one atomicMin, (4) grid barrier.
// Cross-GPU part of the window close (thread 0 of block 0, after the local barrier): publish this rank's
// minimum and window count into every peer's sync block, wait for theirs, combine.  Bounded spin: a peer that
// never arrives raises DFD_ERR_PEER_TIMEOUT instead of hanging the GPU.
__device__ void peer_window_close(const DfdArrays& A, const DfdPeers& X, unsigned long long w) {
    const unsigned slot = (unsigned)((w + 1) % 3);
    unsigned long long m = __ldcg(&A.gmin[slot]);
    const unsigned long long myerr = __ldcg(&A.counters[2]);
    DfdSync* mine = (DfdSync*)X.local_base;
    // one 8-byte store per peer is both the message and the flag: a slot holds DFD_SYNC_EMPTY until the peer's
    // minimum (or DFD_SYNC_FAILED) lands in it.  Every block fenced its own remote records before it arrived at
    // the local barrier, so nothing else has to be ordered here.
    const unsigned long long msg = myerr ? DFD_SYNC_FAILED : m;
    for (int k = 0; k < X.world; k++) {
        if (k == X.rank) continue;
        DfdSync* peer = (DfdSync*)X.base[k];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&peer->mins[slot][X.rank]), "l"(msg) : "memory");
    }
    unsigned long long err = myerr;
    for (int k = 0; k < X.world; k++) {
        if (k == X.rank) continue;
        unsigned long long pm;
        unsigned long long spins = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pm) : "l"(&mine->mins[slot][k]) : "memory");
            if (++spins > (1ULL << 27)) {  // ~ a minute of polling
                err = DFD_ERR_PEER_TIMEOUT;
                break;
            }
        } while (pm == DFD_SYNC_EMPTY);
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&mine->mins[slot][k]), "l"(DFD_SYNC_EMPTY) : "memory");  // re-arm for window w + 3
        if (pm == DFD_SYNC_FAILED) {
            if (!err) err = DFD_ERR_PEER_ERROR;
        } else if (pm < m)
            m = pm;
    }
    __stcg(&A.gmin[slot], m);
    if (err && !myerr) set_error(A, (unsigned)err, 0);
}
// Window close of a sharded run: one local barrier whose last arriving block exchanges minima and completion
// flags with the other GPUs before it releases the rest through the go word.
__device__ __forceinline__ void grid_barrier_peers(const DfdArrays& A, const DfdPeers& X, unsigned long long w, unsigned long long* ctr,
                                                   unsigned long long target, unsigned long long* go) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long old;
        asm volatile("atom.release.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(ctr), "l"(1ULL) : "memory");
        if (old + 1 == target) {
            peer_window_close(A, X, w);
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(go), "l"(w + 1) : "memory");
        } else {
            unsigned long long v;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(go) : "memory");
            } while (v < w + 1);
        }
    }
    __syncthreads();
}

// Two instantiations: <128, 2> lets the compiler use up to 255 registers per thread (186 used, nothing spilled) for
// partitions whose grid fits 2 blocks per SM; <512, 1> caps at 128 registers so that 16 warps per SM stay resident
// for large partitions.
template <int MAX_THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS) dfd_run_kernel(DfdParams P, DfdArrays A, DfdPeers X, unsigned long long max_windows) {
    extern __shared__ unsigned s_list[];  // [lps_per_block]: bit31 = router, low bits = index; then tc/tk halves
    __shared__ double s_min[MAX_THREADS / 32];
    __shared__ unsigned s_ev[EV_NTYPES];
    __shared__ unsigned s_n, s_nr;  // active nodes (list front) / active routers (list back)
    const unsigned tid = threadIdx.x, nth = blockDim.x, nwarps = nth >> 5;
    // this rank owns nodes [X.n0, X.n1) and routers [X.r0, X.r1); the block takes a contiguous slice of each
    const unsigned ntb = ((unsigned)(X.n1 - X.n0) + gridDim.x - 1) / gridDim.x, nrb = ((unsigned)(X.r1 - X.r0) + gridDim.x - 1) / gridDim.x;
    const unsigned n0 = min((unsigned)X.n1, X.n0 + blockIdx.x * ntb), n1 = min((unsigned)X.n1, n0 + ntb);
    const unsigned r0 = min((unsigned)X.r1, X.r0 + blockIdx.x * nrb), r1 = min((unsigned)X.r1, r0 + nrb);
    unsigned* s_tails = s_list + (ntb + nrb);
    Ctx c;
    c.P = &P;
    c.A = &A;
    c.X = &X;
    c.ev = s_ev;
    if (tid < EV_NTYPES) s_ev[tid] = 0;
    if (tid == 0) {
        s_n = 0;
        s_nr = 0;
    }
    __syncthreads();
    const unsigned lcap = ntb + nrb;
    unsigned long long w = __ldcg(&A.counters[1]);  // windows executed so far (the kernel may be re-entered)
    const unsigned long long w_stop = w + max_windows;
    const unsigned long long bar0 = __ldcg(&A.counters[28]);  // barrier count at entry
    unsigned long long nbar = 0;
    long long t_scan = 0, t_proc = 0, t_bar = 0, t_procmax = 0;  // cycle accounting (thread 0 of each block)
    for (;;) {
        long long c0 = clock64();
        // one L2 round trip: next window start, error flag and this thread's first wake words are all requested
        // before any of them is consumed
        const int cur = (int)(w & 1);
        const unsigned nA = n0 + tid, rA = r0 + tid;
        unsigned long long tb = __ldcg(&A.gmin[w % 3]);
        unsigned long long errw = __ldcg(&A.counters[2]);
        unsigned long long wiA = INF_BITS, wsA = INF_BITS, wsR = INF_BITS;
        unsigned tcA = 0, tkA = 0;
        if (nA < n1) {
            wiA = __ldcg(&A.wake_in[(size_t)cur * P.NT + nA]);
            wsA = A.wake_self[nA];
        }
        if (rA < r1) {
            tcA = __ldcg(&A.rcin_tail[(size_t)cur * P.NR + rA]);
            tkA = __ldcg(&A.rkin_tail[(size_t)cur * P.NR + rA]);
            wsR = A.wake_self[P.NT + rA];
        }
        if (tb >= INF_BITS || errw != 0 || w >= w_stop) break;
        double T = bits2d(tb);
        if (T >= P.ts_end) break;
        c.t_end = T + P.window;
        c.b_beg = (long long)(T * P.inv_window);
        c.cur = cur;
        c.nxt = cur ^ 1;
        c.widx = (unsigned)w;
        c.cand = DINF;
        c.remote_writes = false;
        if (blockIdx.x == 0 && tid == 0) __stcg(&A.gmin[(w + 2) % 3], INF_BITS);
        const unsigned long long te_bits = d2bits(c.t_end);
        // (1) scan + compact
        for (unsigned n = nA; n < n1; n += nth) {
            unsigned long long ws, wi;
            if (n == nA) {
                ws = wsA;
                wi = wiA;
            } else {
                ws = A.wake_self[n];
                wi = __ldcg(&A.wake_in[(size_t)cur * P.NT + n]);
            }
            if (wi != INF_BITS) {
                __stcg(&A.wake_in[(size_t)cur * P.NT + n], INF_BITS);
                if (wi < ws) {
                    ws = wi;
                    A.wake_self[n] = ws;
                }
            }
            if (ws < te_bits)
                s_list[atomicAdd(&s_n, 1u)] = n;
            else
                c.cand = fmin(c.cand, bits2d(ws));
        }
        for (unsigned r = rA; r < r1; r += nth) {
            unsigned long long ws;
            unsigned tc, tk;
            if (r == rA) {
                ws = wsR;
                tc = tcA;
                tk = tkA;
            } else {
                ws = A.wake_self[P.NT + r];
                tc = __ldcg(&A.rcin_tail[(size_t)cur * P.NR + r]);
                tk = __ldcg(&A.rkin_tail[(size_t)cur * P.NR + r]);
            }
            unsigned xany = 0;
            if (X.world > 1) xany = __ldcg(&A.xmask[(size_t)cur * P.NR + r]) != 0ULL;
            if (tc | tk | xany | (unsigned)(ws < te_bits)) {
                unsigned pos = lcap - 1 - atomicAdd(&s_nr, 1u);
                s_list[pos] = r;
                s_tails[pos] = (tc << 16) | (tk & 0xFFFFu);
            } else
                c.cand = fmin(c.cand, bits2d(ws));
        }
        __syncthreads();
        long long c1 = clock64();
        // (2) process.  Nodes and routers are dealt out separately so that the lanes of a warp run the same kind of
        // LP: nodes warp-first from warp 0 upwards, routers warp-first from the last warp downwards (with few active
        // LPs the two kinds do not share a warp, and no LP shares one with another).
        const unsigned nact = s_n, ract = s_nr;
        for (unsigned j = (tid & 31) * nwarps + (tid >> 5); j < nact; j += nth) node_process(c, s_list[j]);
        for (unsigned j = (tid & 31) * nwarps + (nwarps - 1 - (tid >> 5)); j < ract; j += nth) {
            unsigned pos = lcap - 1 - j;
            unsigned t = s_tails[pos];
            router_process(c, s_list[pos], t >> 16, t & 0xFFFFu);
        }
        if (c.remote_writes) __threadfence_system();  // records written over NVLink must be visible before the flags
        // (3) block-min of the candidates, one atomicMin per block
        double v = c.cand;
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
        if ((tid & 31) == 0) s_min[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double m = s_min[0];
            for (unsigned i = 1; i < nwarps; i++) m = fmin(m, s_min[i]);
            if (d2bits(m) != INF_BITS) atomicMin(&A.gmin[(w + 1) % 3], d2bits(m));
            s_n = 0;
            s_nr = 0;
        }
        // (4) close the window
        long long c2 = clock64();
        nbar++;
        if (X.world > 1)
            grid_barrier_peers(A, X, w, &A.counters[29], (bar0 + nbar) * gridDim.x, &A.counters[30]);
        else
            grid_barrier(&A.counters[29], (bar0 + nbar) * gridDim.x);
        long long c3 = clock64();
        t_scan += c1 - c0;
        t_proc += c2 - c1;
        t_bar += c3 - c2;
        if (c2 - c1 > t_procmax) t_procmax = c2 - c1;
        w++;
    }
    __syncthreads();
    if (tid < EV_NTYPES && s_ev[tid]) {
        atomicAdd(&A.counters[4 + tid], (unsigned long long)s_ev[tid]);
        atomicAdd(&A.counters[0], (unsigned long long)s_ev[tid]);
    }
    if (tid == 0) {  // per-block cycle sums: scan / process / barrier wait
        atomicAdd(&A.counters[20], (unsigned long long)t_scan);
        atomicAdd(&A.counters[21], (unsigned long long)t_proc);
        atomicAdd(&A.counters[22], (unsigned long long)t_bar);
        atomicMax(&A.counters[23], (unsigned long long)t_proc);
        atomicMax(&A.counters[24], (unsigned long long)t_procmax);
    }
    if (blockIdx.x == 0 && tid == 0) {
        A.counters[1] = w;
        A.counters[28] = bar0 + nbar;
    }
}

// ================================================================================================
// state initialisation (device side: nothing but scalars crosses PCIe)
// ================================================================================================
__device__ unsigned long long mulmod_u64(unsigned long long a, unsigned long long b, unsigned long long m) { return (a * b) % m; }  // a,b < 2^31
__device__ void seed_stream(int32_t* s, unsigned long long stream, const DfdSeedConsts& K) {
    for (int j = 0; j < 4; j++) {
        unsigned long long base = K.avw[j], acc = 1, e = stream, m = K.m[j];
        while (e) {
            if (e & 1) acc = mulmod_u64(acc, base, m);
            base = mulmod_u64(base, base, m);
            e >>= 1;
        }
        s[j] = (int32_t)mulmod_u64(acc, K.seed0[j], m);
    }
}

__global__ void dfd_init_kernel(DfdParams P, DfdArrays A, DfdSeedConsts K) {
    const unsigned nthreads = gridDim.x * blockDim.x;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid == 0) {
        A.gmin[0] = d2bits(P.mean_interval);
        A.gmin[1] = INF_BITS;
        A.gmin[2] = INF_BITS;
        for (int i = 0; i < 64; i++) A.counters[i] = 0;
    }
    for (unsigned n = gtid; n < (unsigned)P.NT; n += nthreads) {
        NodeState* s = &A.node[n];
        memset(s, 0, sizeof(NodeState));
        unsigned sg = svr_gid(P, n), tg = term_gid(P, n);
        seed_stream(s->rng_svr0, (unsigned long long)sg * 3 + 0, K);
        seed_stream(s->rng_svr2, (unsigned long long)sg * 3 + 2, K);
        seed_stream(s->rng_term2, (unsigned long long)tg * 3 + 2, K);
        s->dest_id = -1;
        s->svr_gid = sg;
        s->term_gid = tg;
        s->router = n / P.p;
        s->grp = s->router / P.a;
        s->min_latency = 2147483647.0;  // INT_MAX dally:4766
        // svr_init synthetic:287-298: first KICKOFF at mean_interval
        NodeEv* ev = &A.nev[(size_t)n * P.nev_cap];
        ev->ts = 0.0 + P.mean_interval;
        ev->src = sg;
        ev->seq = s->svr_seq++;
        ev->type = EV_KICKOFF;
        s->nev = 1;
        s->nev_hwm = 1;
        A.wake_self[n] = d2bits(ev->ts);
        A.wake_in[n] = INF_BITS;
        A.wake_in[(size_t)P.NT + n] = INF_BITS;
        A.ncin_tail[n] = 0;
        A.nkin_tail[n] = 0;
        A.ack_count[n] = 0;
        A.ack_first[n] = INF_BITS;
        A.ack_last[n] = 0;
        for (int i = 0; i < P.ncin_cap; i++) *(unsigned long long*)&A.ncin[(size_t)n * P.ncin_cap + i].ts = INF_BITS;
        for (int i = 0; i < P.nkin_cap; i++) *(unsigned long long*)&A.nkin[(size_t)n * P.nkin_cap + i].ts = INF_BITS;
    }
    for (unsigned r = gtid; r < (unsigned)P.NR; r += nthreads) {
        RouterHdr* h = &A.rh[r];
        memset(h, 0, sizeof(RouterHdr));
        seed_stream(h->rng, (unsigned long long)rtr_gid(P, r) * 3 + 0, K);
        h->pool_nfree = (uint32_t)P.rpool_cap;
        h->ent_nfree = (uint32_t)P.rheap_cap;
        h->wmin = DINF;
        h->gid = rtr_gid(P, r);
        h->lid = (uint16_t)(r % P.a);
        h->grp = (uint16_t)(r / P.a);
        A.wake_self[P.NT + r] = INF_BITS;
        A.rcin_tail[r] = 0;
        A.rcin_tail[(size_t)P.NR + r] = 0;
        A.rkin_tail[r] = 0;
        A.rkin_tail[(size_t)P.NR + r] = 0;
    }
    // per-port state and pool free lists, flat over all routers
    size_t nports = (size_t)P.NR * P.radix;
    for (size_t i = gtid; i < nports; i += nthreads) {
        PortState* ps = &A.port[i];
        memset(ps, 0, sizeof(PortState));
        PortQ* pq = &A.portq[i];
        for (int v = 0; v < DFD_NUM_VCS; v++) pq->pend_head[v] = pq->pend_tail[v] = pq->queue_head[v] = pq->queue_tail[v] = NIL;
    }
    size_t nent = (size_t)P.NR * P.rheap_cap;
    for (size_t i = gtid; i < nent; i += nthreads) {
        unsigned k = (unsigned)(i % P.rheap_cap);
        A.efree[i] = (unsigned)P.rheap_cap - 1 - k;  // popped from the top: entry 0 first
    }
    size_t nbk = (size_t)P.NR * P.wheel_nb;
    for (size_t i = gtid; i < nbk; i += nthreads) A.wheel[i] = NIL;
    for (size_t i = gtid; i < (nbk >> 6); i += nthreads) A.wbits[i] = 0ULL;
    size_t nslots = (size_t)P.NR * P.rpool_cap;
    for (size_t i = gtid; i < nslots; i += nthreads) {
        unsigned k = (unsigned)(i % P.rpool_cap);
        A.pfree[i] = (unsigned)P.rpool_cap - 1 - k;
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
    size_t NT = P.NT, NR = P.NR, NLP = NT + NR;
    if (dalloc(dev, &A.wake_self, NLP)) return 1;
    if (dalloc(dev, &A.wake_in, 2 * NT)) return 1;
    if (dalloc(dev, &A.gmin, 4)) return 1;
    if (dalloc(dev, &A.node, NT)) return 1;
    if (dalloc(dev, &A.nev, NT * P.nev_cap)) return 1;
    if (dalloc(dev, &A.sq, NT * P.nsq_cap)) return 1;
    if (dalloc(dev, &A.tq, NT * P.ntq_cap)) return 1;
    if (dalloc(dev, &A.ncin, NT * P.ncin_cap)) return 1;
    if (dalloc(dev, &A.nkin, NT * P.nkin_cap)) return 1;
    if (dalloc(dev, &A.mtab, NT * (size_t)P.nreasm_cap)) return 1;
    if (dalloc(dev, &A.ptab, NT * (size_t)P.nreasm_cap)) return 1;
    if (dalloc(dev, &A.ncin_tail, NT)) return 1;
    if (dalloc(dev, &A.nkin_tail, NT)) return 1;
    if (dalloc(dev, &A.rh, NR)) return 1;
    if (dalloc(dev, &A.port, NR * P.radix)) return 1;
    if (dalloc(dev, &A.portq, NR * P.radix)) return 1;
    if (dalloc(dev, &A.pool, NR * P.rpool_cap)) return 1;
    if (dalloc(dev, &A.pent, NR * P.rheap_cap)) return 1;
    if (dalloc(dev, &A.pfree, NR * P.rpool_cap)) return 1;
    if (dalloc(dev, &A.efree, NR * P.rheap_cap)) return 1;
    if (dalloc(dev, &A.wheel, NR * (size_t)P.wheel_nb)) return 1;
    if (dalloc(dev, &A.wbits, NR * (size_t)(P.wheel_nb >> 6))) return 1;
    // exchange region: everything another rank may write (router inboxes and their tails, ACK statistics), behind
    // the sync block -- one allocation, one CUDA-IPC handle
    size_t off = align256(sizeof(DfdSync));
    size_t o_rcin = off;
    off = align256(off + 2 * NR * P.rcin_cap * sizeof(ChunkRec));
    size_t o_rkin = off;
    off = align256(off + 2 * NR * P.rkin_cap * sizeof(CreditRec));
    size_t o_rct = off;
    off = align256(off + 2 * NR * sizeof(unsigned));
    size_t o_rkt = off;
    off = align256(off + 2 * NR * sizeof(unsigned));
    size_t o_xc = off, o_xk = off, o_xn = off, o_xm = off;
    if (world > 1) {  // lanes of the global links (only sharded runs use them)
        if (P.h > 64) {
            snprintf(dev->error, sizeof dev->error, "sharded runs support at most 64 global ports per router (have %d)", P.h);
            return 1;
        }
        const size_t lanes = 2 * NR * (size_t)P.h;
        off = align256(off + lanes * (size_t)P.xlane_cap * sizeof(ChunkRec));
        o_xk = off;
        off = align256(off + lanes * (size_t)P.xlane_cap * sizeof(CreditRec));
        o_xn = off;
        off = align256(off + lanes * sizeof(unsigned));
        o_xm = off;
        off = align256(off + 2 * NR * sizeof(unsigned long long));
    }
    size_t o_ac = off;
    off = align256(off + NT * sizeof(unsigned));
    size_t o_af = off;
    off = align256(off + NT * sizeof(unsigned long long));
    size_t o_al = off;
    off = align256(off + NT * sizeof(unsigned long long));
    char* xr = nullptr;
    if (dalloc(dev, &xr, off)) return 1;
    dev->xregion = xr;
    dev->xregion_bytes = off;
    A.rcin = (ChunkRec*)(xr + o_rcin);
    A.rkin = (CreditRec*)(xr + o_rkin);
    A.rcin_tail = (unsigned*)(xr + o_rct);
    A.rkin_tail = (unsigned*)(xr + o_rkt);
    if (world > 1) {
        A.xc = (ChunkRec*)(xr + o_xc);
        A.xk = (CreditRec*)(xr + o_xk);
        A.xcnt = (unsigned*)(xr + o_xn);
        A.xmask = (unsigned long long*)(xr + o_xm);
        dev->x_zero_off = o_xn;  // counts and masks are cleared at reset
        dev->x_zero_bytes = o_ac - o_xn;
        if (dalloc(dev, &A.xsend, NR * (size_t)P.h)) return 1;
        // far-end port of every global link: the j-th link from s to r pairs with the j-th link from r to s
        std::vector<int> grev((size_t)NR * P.h, 0);
        for (size_t sr = 0; sr < NR; sr++)
            for (int k = 0; k < P.h; k++) {
                int r = T.gdest[sr * P.h + k];
                if (r < 0) continue;
                int j = 0;
                for (int k2 = 0; k2 < k; k2++) j += T.gdest[sr * P.h + k2] == r;
                int kp = -1;
                for (int k2 = 0; k2 < P.h; k2++)
                    if (T.gdest[(size_t)r * P.h + k2] == (int)sr && j-- == 0) {
                        kp = k2;
                        break;
                    }
                if (kp < 0) {
                    snprintf(dev->error, sizeof dev->error, "global link %zu -> %d has no reverse link: sharded runs need symmetric global connections", sr, r);
                    return 1;
                }
                grev[sr * P.h + k] = kp;
            }
        int* d_grev = nullptr;
        if (dalloc(dev, &d_grev, grev.size())) return 1;
        CK(cudaMemcpy(d_grev, grev.data(), grev.size() * sizeof(int), cudaMemcpyHostToDevice));
        A.grev = d_grev;
    }
    A.ack_count = (unsigned*)(xr + o_ac);
    A.ack_first = (unsigned long long*)(xr + o_af);
    A.ack_last = (unsigned long long*)(xr + o_al);
    for (int k = 0; k < DFD_MAX_RANKS; k++) dev->X.base[k] = nullptr;
    dev->X.base[rank] = xr;
    dev->X.local_base = xr;
    if (dalloc(dev, &A.counters, 64)) return 1;
    if (dalloc(dev, &A.winmax, 65536)) return 1;
    CK(cudaMemset(A.winmax, 0, 65536 * sizeof(unsigned long long)));
    int *loc_off, *loc_port, *loc_dest, *gs_port, *gs_group, *gdest, *routed_off, *routed_lid;
    if (dalloc(dev, &loc_off, T.loc_off.size())) return 1;
    if (dalloc(dev, &loc_port, T.loc_port.size())) return 1;
    if (dalloc(dev, &loc_dest, T.loc_dest.size())) return 1;
    if (dalloc(dev, &gs_port, T.gs_port.size())) return 1;
    if (dalloc(dev, &gs_group, T.gs_group.size())) return 1;
    if (dalloc(dev, &gdest, T.gdest.size())) return 1;
    if (dalloc(dev, &routed_off, T.routed_off.size())) return 1;
    if (dalloc(dev, &routed_lid, T.routed_lid.size())) return 1;
    A.loc_off = loc_off;
    A.loc_port = loc_port;
    A.loc_dest = loc_dest;
    A.gs_port = gs_port;
    A.gs_group = gs_group;
    A.gdest = gdest;
    A.routed_off = routed_off;
    A.routed_lid = routed_lid;
    dev->topo_bytes = 0;
    // grid: persistent and co-resident (cooperative launch).  The engine is latency-bound (one grid barrier per
    // lookahead window), so LPs are spread thinly: ~DFD_LPS_PER_BLOCK LPs per block until every SM holds
    // blocks_per_sm blocks, after which the per-block slice grows.
    const size_t ownT = dev->X.n1 - dev->X.n0, ownR = dev->X.r1 - dev->X.r0;
    dev->block = 128;
    if (const char* ev = getenv("DFD_BLOCK_THREADS")) dev->block = std::max(32, std::min(DFD_BLOCK, atoi(ev) / 32 * 32));
    // small partitions are latency-bound with at most 2 blocks per SM: give the compiler the registers
    bool wide = ownT + ownR <= 12288 && dev->block <= 128;
    // ensemble: this engine shares the GPU with ensemble-1 others -> 1/ensemble of the block slots.  The 186-register
    // instantiation offers 2 slots per SM, the 128-register one 4: more than 3 simulations take the latter.
    if (ensemble > 3) wide = false;
    if (const char* ev = getenv("DFD_WIDE_REGS")) wide = atoi(ev) != 0 && dev->block <= 128;
    dev->kernel = wide ? (const void*)dfd_run_kernel<128, 2> : (const void*)dfd_run_kernel<DFD_BLOCK, 1>;
    int lps_per_block = 16;
    if (const char* ev = getenv("DFD_LPS_PER_BLOCK")) lps_per_block = std::max(1, atoi(ev));
    size_t want = (ownT + ownR + lps_per_block - 1) / lps_per_block;
    for (int iter = 0; iter < 2; iter++) {
        // dynamic shared memory depends on the slice size, which depends on the grid: settle in two passes
        size_t g = iter == 0 ? std::max<size_t>(1, want) : (size_t)dev->grid;
        size_t ntb = (ownT + g - 1) / g, nrb = (ownR + g - 1) / g;
        dev->smem = (ntb + nrb) * 2 * sizeof(unsigned);
        if (dev->smem > 48 * 1024) CK(cudaFuncSetAttribute(dev->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dev->smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (void (*)(DfdParams, DfdArrays, DfdPeers, unsigned long long))dev->kernel, dev->block, dev->smem));
        if (per_sm < 1) {
            snprintf(dev->error, sizeof dev->error, "run kernel does not fit on an SM (%zu bytes of shared memory)", dev->smem);
            return 1;
        }
        size_t maxb = (size_t)per_sm * dev->num_sms;
        dev->grid = (int)std::max<size_t>(1, std::min(want, maxb));
        dev->blocks_per_sm = per_sm;
    }
    if (ensemble > 1) {
        int slots = dev->blocks_per_sm * dev->num_sms;
        dev->grid = std::max(1, std::min(dev->grid, slots / ensemble));
    }
    if (const char* ev = getenv("DFD_GRID")) dev->grid = std::max(1, std::min(dev->grid, atoi(ev)));
    {
        size_t g = dev->grid, ntb = (ownT + g - 1) / g, nrb = (ownR + g - 1) / g;
        dev->smem = (ntb + nrb) * 2 * sizeof(unsigned);
        if (dev->smem > 48 * 1024) CK(cudaFuncSetAttribute(dev->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dev->smem));
    }
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
    size_t total = 0;
#define UP(dst, vec)                                                                                         \
    do {                                                                                                     \
        if (!(vec).empty()) CK(cudaMemcpyAsync((void*)(dst), (vec).data(), (vec).size() * sizeof(int), cudaMemcpyHostToDevice, st)); \
        total += (vec).size() * sizeof(int);                                                                 \
    } while (0)
    UP(A.loc_off, T.loc_off);
    UP(A.loc_port, T.loc_port);
    UP(A.loc_dest, T.loc_dest);
    UP(A.gs_port, T.gs_port);
    UP(A.gs_group, T.gs_group);
    UP(A.gdest, T.gdest);
    UP(A.routed_off, T.routed_off);
    UP(A.routed_lid, T.routed_lid);
#undef UP
    dev->topo_bytes = total;
    return 0;
}

int dfd_device_reset(DfdDevice* dev, cudaStream_t st) {
    int blocks = dev->num_sms * 8;
    CK(cudaMemsetAsync(dev->xregion, 0xFF, sizeof(DfdSync), st));  // every slot DFD_SYNC_EMPTY
    if (dev->X.world > 1) {
        CK(cudaMemsetAsync(dev->xregion + dev->x_zero_off, 0, dev->x_zero_bytes, st));
        CK(cudaMemsetAsync(dev->A.xsend, 0, sizeof(uint2) * (size_t)dev->P.NR * dev->P.h, st));
    }
    dfd_init_kernel<<<blocks, 256, 0, st>>>(dev->P, dev->A, dev->seeds);
    CK(cudaGetLastError());
    return 0;
}

int dfd_device_run(DfdDevice* dev, cudaStream_t st, unsigned long long max_windows) {
    for (int k = 0; k < dev->X.world; k++)
        if (!dev->X.base[k]) {
            snprintf(dev->error, sizeof dev->error, "rank %d: the exchange region of rank %d has not been imported", dev->X.rank, k);
            return 1;
        }
    void* args[] = {(void*)&dev->P, (void*)&dev->A, (void*)&dev->X, (void*)&max_windows};
    CK(cudaLaunchCooperativeKernel(dev->kernel, dim3(dev->grid), dim3(dev->block), args, dev->smem, st));
    return 0;
}

int dfd_device_download(DfdDevice* dev, DfdHostResults* out, cudaStream_t st) {
    const DfdParams& P = dev->P;
    const DfdArrays& A = dev->A;
    out->node.resize(P.NT);
    out->rh.resize(P.NR);
    out->port.resize((size_t)P.NR * P.radix);
    out->ack_count.resize(P.NT);
    out->ack_first.resize(P.NT);
    out->ack_last.resize(P.NT);
    CK(cudaMemcpyAsync(out->node.data(), A.node, sizeof(NodeState) * P.NT, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->rh.data(), A.rh, sizeof(RouterHdr) * P.NR, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->port.data(), A.port, sizeof(PortState) * P.NR * P.radix, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->ack_count.data(), A.ack_count, sizeof(unsigned) * P.NT, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->ack_first.data(), A.ack_first, sizeof(unsigned long long) * P.NT, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->ack_last.data(), A.ack_last, sizeof(unsigned long long) * P.NT, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->counters, A.counters, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost, st));
    out->winmax.clear();
    if (getenv("DFD_PROFILE_PRINT")) {  // per-window slowest activation, filled by -DDFD_PROFILE builds only
        out->winmax.resize(65536);
        CK(cudaMemcpyAsync(out->winmax.data(), A.winmax, sizeof(unsigned long long) * 65536, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    out->d2h_bytes = sizeof(NodeState) * (size_t)P.NT + sizeof(RouterHdr) * (size_t)P.NR + sizeof(PortState) * (size_t)P.NR * P.radix +
                     (sizeof(unsigned) + 2 * sizeof(unsigned long long)) * (size_t)P.NT + sizeof(unsigned long long) * 32;
    return 0;
}

int dfd_device_read_counters(DfdDevice* dev, unsigned long long* counters, cudaStream_t st) {
    CK(cudaMemcpyAsync(counters, dev->A.counters, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

void dfd_device_destroy(DfdDevice* dev) {
    for (void* p : dev->ipc_opened) cudaIpcCloseMemHandle(p);
    dev->ipc_opened.clear();
    for (void* p : dev->allocs) cudaFree(p);
    dev->allocs.clear();
}
