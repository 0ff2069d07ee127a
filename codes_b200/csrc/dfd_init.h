// dfd_init.h -- initial LP state (svr_init synthetic:287-298, terminal_dally_init dally:4713-4900,
// router_dally_init dally:4902-5111, model_net_base_lp_init model-net-lp.c:450-533) written in place on the device:
// nothing but scalars crosses PCIe.  Dual compile like dfd_core.h.
#pragma once
#include "dfd_core.h"

DFD_FN unsigned long long mulmod_u64(unsigned long long a, unsigned long long b, unsigned long long m) { return (a * b) % m; }  // a,b < 2^31
// seed of stream k = default seed jumped ahead k * 2^72 steps: s_j * (a_j^(2^72))^k mod m_j
DFD_FN void seed_stream(int32_t* s, unsigned long long stream, const DfdSeedConsts& K) {
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

DFD_FN void init_ctrl(const DfdParams& P, const DfdArrays& A, const DfdPeers& X) {
    (void)P;
    memset(A.ctrl, 0, sizeof(DfdCtrl));
    A.ctrl->created = 0.0 + P.mean_interval < P.ts_end ? (unsigned long long)(X.n1 - X.n0) : 0ULL;  // one KICKOFF per server
    A.ctrl->created += (unsigned long long)(X.r1 - X.r0) * (unsigned long long)P.nsnap;          // R_SNAPSHOT events per router
    for (int i = 0; i < 64; i++) A.counters[i] = 0;
}

DFD_FN void init_node(const DfdParams& P, const DfdArrays& A, const DfdPeers& X, const DfdSeedConsts& K, unsigned nl) {
    const unsigned n = (unsigned)X.n0 + nl;
    NodeState* s = &A.node[nl];
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
    NodeEv* ev = &A.nev[(size_t)nl * P.nev_cap];
    ev->ts = 0.0 + P.mean_interval;
    ev->src = sg;
    ev->seq = s->svr_seq++;
    ev->type = EV_KICKOFF;
    ev->a = ev->b = ev->c = ev->e = ev->pad = 0;
    ev->d = 0.0;
    s->nev = 1;
    s->nev_hwm = 1;
    s->self_next = ev->ts;
    if (!(ev->ts < P.ts_end)) {  // the run ends before the first KICKOFF (--end below the inter-arrival time): never sent
        s->nev = 0;
        s->self_next = DINF;
    }
    A.ncin_tail[nl] = 0;
    A.nkin_tail[nl] = 0;
    A.ack_count[nl] = 0;
    A.ack_first[nl] = INF_BITS;
    A.ack_last[nl] = 0;
}

DFD_FN void init_router(const DfdParams& P, const DfdArrays& A, const DfdPeers& X, const DfdSeedConsts& K, unsigned rl) {
    const unsigned r = (unsigned)X.r0 + rl;
    RouterHdr* h = &A.rh[rl];
    memset(h, 0, sizeof(RouterHdr));
    seed_stream(h->rng, (unsigned long long)rtr_gid(P, r) * 3 + 0, K);
    h->pool_nfree = (uint32_t)P.rpool_cap;
    h->gid = rtr_gid(P, r);
    h->lid = (uint16_t)(r % P.a);
    h->grp = (uint16_t)(r / P.a);
    h->clock = -1.0;
    h->ev_min = 0.0;
    A.pollhdr[rl].ev_min = 0.0;
    A.pollhdr[rl].flags = 2u;
    A.pollhdr[rl].pad = 0;
    h->pad1 = 0.0;
    // router_send_snapshot_events dally:4407-4414: the router's first nsnap_all events are its R_SNAPSHOTs (sequence numbers
    // 0..nsnap_all-1, also those past the end of the run, which are never sent)
    h->seq = (uint32_t)P.nsnap_all;
    h->snap_next = 0;
    for (int i = 0; i < 2 * P.p; i++) A.iblk[(size_t)rl * 2 * P.p + i] = 0.0;
}

// per (router, index j) state, flat over i = rl * radix + j
DFD_FN void init_lane(const DfdParams& P, const DfdArrays& A, const DfdPeers& X, size_t i) {
    const unsigned rl = (unsigned)(i / P.radix);
    const int j = (int)(i % P.radix);
    const unsigned r = (unsigned)X.r0 + rl;
    const int lid = (int)(r % P.a);
    memset(&A.port[i], 0, sizeof(PortState));
    PortQ* pq = &A.portq[i];
    for (int v = 0; v < DFD_NUM_VCS; v++) pq->pend_head[v] = pq->pend_tail[v] = pq->queue_head[v] = pq->queue_tail[v] = NIL;
    memset(&A.lane[i], 0, sizeof(LaneCtl));
    A.skey[i] = mk_key(DINF, 0, 0);
    A.qbits[2 * i] = A.qbits[2 * i + 1] = 0ULL;
    // an in-lane / port without a link never blocks: its bound is "never"
    bool in_used = true, out_used = true;
    if (j < P.intra_radix) {
        in_used = A.lin_src[lid * P.intra_radix + j] >= 0;
        out_used = A.loc_dest[lid * P.intra_radix + j] >= 0;
    } else if (j < P.ext) {
        in_used = A.gin_src[(size_t)r * P.h + (j - P.intra_radix)] >= 0;
        out_used = A.gdest[(size_t)r * P.h + (j - P.intra_radix)] >= 0;
    }
    const unsigned long long never = cw_pack(DFD_TS_BIG, 0);
    // seen != channel word for the lanes in use: the first poll takes every bound (0) as news
    const size_t xi = (size_t)rl * P.radix + j;
    A.cw_c[xi] = in_used ? 0ULL : never;
    A.seen_c[i] = in_used ? ~0ULL : never;
    A.ckey[i] = mk_key(in_used ? 0.0 : DFD_TS_BIG, 0, 0);
    A.cw_k[xi] = out_used ? 0ULL : never;
    A.seen_k[i] = out_used ? ~0ULL : never;
    A.kkey[i] = mk_key(out_used ? 0.0 : DFD_TS_BIG, 0, 0);
    A.pub_c[i] = ~0ULL;
    A.pub_k[i] = ~0ULL;
}
