// Shared host/device data layout of the dragonfly-dally window engine.
// Terminology follows the reference (terminal, router, chunk, credit, VC, port).
#pragma once
#include <stdint.h>
#include <vector_types.h>

#define DFD_NUM_VCS 4  // dragonfly-dally.cxx:2387 (num_qos_levels == 1)

// event types; numbering is internal to the engine (only used for dispatch and per-type counters)
enum DfdEv {
    EV_KICKOFF = 0,
    EV_REMOTE,
    EV_LOCAL,
    EV_ACK,
    EV_NEW_MSG,
    EV_SCHED_NEXT,
    EV_T_GENERATE,
    EV_T_ARRIVE,
    EV_T_SEND,
    EV_T_BUFFER,
    EV_R_SEND,
    EV_R_ARRIVE,
    EV_R_BUFFER,
    EV_NTYPES
};

// enum ROUTING_ALGO / last_hop values of the reference (dragonfly-dally.cxx:576-598)
enum { DFD_MINIMAL = 1, DFD_NON_MINIMAL = 2, DFD_PROG_ADAPTIVE = 4 };
enum { DFD_HOP_GLOBAL = 1, DFD_HOP_LOCAL = 2, DFD_HOP_TERMINAL = 3 };
enum { DFD_SCORE_ALPHA = 1, DFD_SCORE_DELTA = 4 };
enum { DFD_TR_UNIFORM = 1, DFD_TR_RAND_PERM = 2, DFD_TR_NEAREST_GROUP = 3, DFD_TR_NEAREST_NEIGHBOR = 4, DFD_TR_RANDOM_OTHER_GROUP = 5 };

// 96-byte packed chunk event (R_ARRIVE / T_ARRIVE): the fields of terminal_dally_message
// (codes/net/dragonfly-dally.h:19-139) that are consumed downstream.  The first 16 bytes are the
// event key (timestamp + deterministic tie-break: sender gid, sender's event-creation sequence).
struct __attribute__((aligned(16))) ChunkRec {
    double ts;
    uint32_t src;  // gid of the sending LP
    uint32_t seq;
    uint32_t packet_id;
    uint32_t src_term;  // dfdally_src_terminal_id
    uint32_t dst_term;  // dfdally_dest_terminal_id
    uint32_t message_id;
    uint32_t packet_size;
    uint32_t total_size;
    uint16_t chunk_id;
    uint16_t dst_grp;    // group of the destination router (carried so that no hop has to divide)
    int16_t intm_grp_id;
    uint16_t org_grp;    // group of origin_router
    uint32_t prev_lp;   // intm_lp_id as router id (last_hop LOCAL/GLOBAL) or source terminal id (TERMINAL)
    uint16_t prev_port; // vc_index of the upstream hop
    uint8_t prev_vc;    // output_chan of the upstream hop
    uint8_t last_hop;
    int8_t path_type;
    uint8_t is_intm_visited;
    uint8_t n_hop, l_hop, g_hop, hops_cur_group;
    uint8_t flags;  // bit0: carries the remote event, bit1: carries the local (self) event
    uint8_t pad0;
    double travel_start_time;
    double msg_start_time;
    uint32_t origin_router;
    uint32_t src_svr;  // sender_lp as svr id
    uint32_t dst_svr;  // final_dest_gid as svr id
    uint32_t dst_router;  // router of dst_term
};

// 32-byte credit event (T_BUFFER / R_BUFFER)
struct __attribute__((aligned(16))) CreditRec {
    double ts;
    uint32_t src;
    uint32_t seq;
    uint16_t port;  // vc_index at the receiver (its output port)
    uint8_t vc;     // output_chan
    uint8_t pad0;
    uint32_t pad1;
    uint64_t pad2;
};

// router chunk pool slot: a chunk waiting in pending_msgs / queued_msgs (or in flight towards the
// router: copied here when the inbox is drained).  Same 96 bytes; the key bytes are reused.
struct __attribute__((aligned(16))) PoolSlot {
    uint32_t next;        // FIFO link / free-list link
    uint16_t out_port;    // vc_index chosen by this router
    uint8_t out_vc;       // output_chan chosen by this router
    uint8_t next_is_term; // next stop is the destination terminal
    uint32_t next_stop;   // router id or terminal id
    uint32_t pad;
    uint32_t packet_id, src_term, dst_term, message_id;
    uint32_t packet_size, total_size;
    uint16_t chunk_id, dst_grp;
    int16_t intm_grp_id;
    uint16_t org_grp;
    uint32_t prev_lp;
    uint16_t prev_port;
    uint8_t prev_vc;
    uint8_t last_hop;
    int8_t path_type;
    uint8_t is_intm_visited;
    uint8_t n_hop, l_hop, g_hop, hops_cur_group;
    uint8_t flags;
    uint8_t pad0;
    double travel_start_time;
    double msg_start_time;
    uint32_t origin_router, src_svr, dst_svr, dst_router;
};

// A router's pending events (chunk arrivals, credits, its own R_SEND timers) sit in a timing wheel: bucket
// b = floor(ts / (4 * window)) mod wheel_nb holds a singly linked list of entries; a bitmap marks non-empty buckets.
// Push is O(1); a window [T, T+L) only has to look at the (at most three) buckets it overlaps.
struct __attribute__((aligned(8))) PendEnt {
    double ts;
    uint32_t src;
    uint32_t seq;
    uint32_t ref;   // kind<<28 | payload (pool slot / port | vc<<16)
    uint32_t next;  // next entry of the bucket, or of the free list
};
enum { RK_CHUNK = 1, RK_CREDIT = 2, RK_SEND = 3 };

// 64-byte hot state of one router output port (router_state dally:822-888)
struct __attribute__((aligned(16))) PortState {
    double noat;  // next_output_available_time
    double last_buf_full;
    double busy_time;
    long long link_traffic;
    int32_t vc_occ[DFD_NUM_VCS];
    int32_t queued_count;
    uint32_t stalled_chunks;
    uint32_t total_chunks;
    uint8_t in_send_loop;
    uint8_t last_qos_lvl;
    uint16_t pad;
};
// FIFO heads of one port: pending_msgs[port][vc] and queued_msgs[port][vc] as pool-slot lists
struct __attribute__((aligned(16))) PortQ {
    uint32_t pend_head[DFD_NUM_VCS], pend_tail[DFD_NUM_VCS];
    uint32_t queue_head[DFD_NUM_VCS], queue_tail[DFD_NUM_VCS];
};

struct __attribute__((aligned(16))) RouterHdr {
    int32_t rng[4];  // CLCG4 stream gid*3+0
    uint32_t seq;
    uint32_t pend_n;     // pending events (entries in the wheel)
    uint32_t pool_nfree; // entries on the free-slot stack
    uint32_t gid;        // codes_mapping gid of this router
    uint16_t lid, grp;   // id inside the group, group id
    uint32_t ent_nfree;  // entries on the free wheel-entry stack
    uint32_t pad0, pad1;
    double wmin;         // exact earliest timestamp among the wheel entries (+inf when the wheel is empty)
    unsigned long long rng_draws;
};

// self-scheduled event of a node (svr + base LP + terminal fused), 48 bytes
struct __attribute__((aligned(16))) NodeEv {
    double ts;
    uint32_t src, seq;
    uint32_t type;  // DfdEv | flags<<8
    uint32_t a, b, c;
    double d;
    uint32_t e;
    uint32_t pad;
};
#define NEV_FLAG_QREQ 0x100
#define NEV_FLAG_LAST 0x200

// one queued model-net request (mn_sched_qitem model-net-sched-impl.c:54-63), 24 bytes
struct SchedReq {
    uint32_t dst_svr;
    uint32_t msg_size;
    uint32_t rem;
    uint32_t msg_id;
    double msg_start_time;
};
// one chunk waiting in terminal_msgs (dally:5653-5680), 40 bytes
struct TermChunk {
    uint32_t dst_svr, packet_id, packet_size, total_size, message_id, chunk_id;
    uint32_t flags;  // bit0/1: remote/local event; bits 16..31: destination group
    uint32_t dst_router;
    double msg_start_time;
};

// node = svr LP + model-net base LP + dragonfly-dally terminal LP of one compute node
struct __attribute__((aligned(16))) NodeState {
    // --- svr_state synthetic:100-117
    int32_t rng_svr0[4];  // stream svr_gid*3+0 (destination draws)
    int32_t rng_svr2[4];  // stream svr_gid*3+2 (codes_local_latency)
    int32_t rng_term2[4]; // stream term_gid*3+2
    uint32_t svr_seq, term_seq;
    int32_t msg_sent_count, warm_msg_sent_count, msg_recvd_count, local_recvd_count;
    int32_t dest_id;
    uint32_t router;     // router this terminal hangs off
    long long rperm_data;
    double last_send_ts;
    double max_server_latency, sum_server_latency;
    // --- model_net_base_state model-net-lp.c:50-74
    double next_available_time;
    uint32_t msg_id;
    uint32_t in_sched_send_loop;
    uint32_t sq_head, sq_count;  // fcfs queue ring
    // --- terminal_state dally:703-820
    uint32_t packet_counter;
    int32_t packet_gen, packet_fin, total_gen_size;
    int32_t vc_occupancy, terminal_length;
    uint32_t in_send_loop, issueIdle;
    uint32_t tq_head, tq_count;  // terminal_msgs ring
    uint32_t nev;                // number of pending self events
    uint32_t cin_head, kin_head;
    uint32_t grp;        // group of `router`
    double terminal_available_time, last_buf_full, busy_time;
    unsigned long long link_traffic, total_msg_size;
    uint32_t total_chunks, stalled_chunks, injected_chunks, ejected_chunks;
    double total_time, total_hops, max_latency, min_latency;
    long long finished_msgs, finished_chunks, finished_packets;
    // mn_stats category "test"
    long long send_count, send_bytes, recv_count, recv_bytes, max_event_size;
    double send_time, recv_time;
    uint32_t stats_used;
    uint32_t nev_hwm, sq_hwm, tq_hwm;
    unsigned long long events;
    unsigned long long rng_draws;
    // global counters contributed by this node (packet_generate dally:5601-5615, packet_arrive :6512-6522)
    uint32_t local_sr, local_sg, remote_pk, minimal_count, nonmin_count, pad2;
    uint32_t svr_gid, term_gid;
    uint32_t mt_n, pt_n;  // live entries of the message / packet reassembly tables
    double node_copy_next_available_time;  // model_net_base_state.node_copy_next_available_time[0] (model-net-lp.c:672-704)
    double pad_d;
};

// scalar parameters (kernel argument, lives in the constant bank)
struct DfdParams {
    int NT, NR, a, p, G, h;
    int intra_radix, global_radix, radix, num_global_channels;
    int chunk_size, packet_size, local_vc, global_vc, cn_vc;
    unsigned mn_packet_size;
    int routing, scoring, adaptive_threshold, global_k_picks;
    double local_bw, global_bw, cn_bw, router_delay;
    double cn_delay, local_delay, global_delay;
    double cn_credit, local_credit, global_credit;
    int credit_size;
    double nic_seq_delay, lookahead, codes_cn_delay;
    double window, ts_end;
    // host-precomputed values of bytes_to_ns()/reciprocals (same IEEE operations as the device would do)
    double inj_chunk_cn, inj_chunk_local, inj_chunk_global;  // bytes_to_ns(chunk_size, bw)
    double inj_pay_cn, inj_pay_local, inj_pay_global;        // bytes_to_ns(payload_sz % chunk_size, bw)
    double inj_packet_cn;                                    // bytes_to_ns(payload_sz, cn_bw) (nic_ts of packet_generate)
    double inv_cn_bw;                                        // 1 / cn_bandwidth
    unsigned long long magic_p, magic_a;                     // floor(2^64/d)+1: n/d == __umul64hi(n, magic) for n < 2^32, d > 1
    int traffic, num_msgs, payload_sz;
    long long rperm_threshold;
    double warm_up_time, mean_interval;
    int num_nodes, nodes_per_grp, node_eager_limit, noop_bypass;
    int lps_per_rep, off_svr, off_term, off_rtr;
    // ring capacities (powers of two) and heap/pool capacities
    int ncin_cap, nkin_cap, nsq_cap, ntq_cap, nev_cap;
    int nreasm_cap;  // per-terminal reassembly table entries (0: every message is one chunk)
    int rcin_cap, rkin_cap, rpool_cap, rheap_cap;
    int xlane_cap;   // records one cross-GPU global link can carry per window and direction (chunks and credits each)
    int wheel_nb;        // buckets of a router's timing wheel (power of two, >= max event delay / window + slack)
    double inv_window;   // 1 / (wheel bucket width)
};

// device arrays
struct DfdArrays {
    // wake times (IEEE doubles compared as uint64: all timestamps are >= 0)
    unsigned long long* wake_self;  // [NLP]
    unsigned long long* wake_in;    // [2][NLP]
    unsigned long long* gmin;       // [3]
    // nodes
    NodeState* node;        // [NT]
    NodeEv* nev;            // [NT][nev_cap]
    SchedReq* sq;           // [NT][nsq_cap]
    TermChunk* tq;          // [NT][ntq_cap]
    ChunkRec* ncin;         // [NT][ncin_cap]
    CreditRec* nkin;        // [NT][nkin_cap]
    uint4* mtab;            // [NT][nreasm_cap] rank_tbl: {message_id, sender svr, remaining_packets, has_remote}
    uint4* ptab;            // [NT][nreasm_cap] remaining_sz_packets: {packet_id, source terminal, remaining bytes, -}
    unsigned* ncin_tail;    // [NT]
    unsigned* nkin_tail;    // [NT]
    unsigned* ack_count;             // [NT]
    unsigned long long* ack_first;   // [NT]
    unsigned long long* ack_last;    // [NT]
    // routers
    RouterHdr* rh;        // [NR]
    PortState* port;      // [NR][radix]
    PortQ* portq;         // [NR][radix]
    PoolSlot* pool;       // [NR][rpool_cap]
    PendEnt* pent;               // [NR][rheap_cap]   wheel entries
    unsigned* pfree;             // [NR][rpool_cap]   stack of free pool slots
    unsigned* efree;             // [NR][rheap_cap]   stack of free wheel entries
    unsigned* wheel;             // [NR][wheel_nb]    bucket heads
    unsigned long long* wbits;   // [NR][wheel_nb/64] non-empty-bucket bitmap
    // cross-GPU global links (world > 1): one single-writer lane per (router, global port) and window parity
    ChunkRec* xc;                // [2][NR][h][xlane_cap]
    CreditRec* xk;               // [2][NR][h][xlane_cap]
    unsigned* xcnt;              // [2][NR][h]  chunks | credits << 16 written into the lane this window
    unsigned long long* xmask;   // [2][NR]     bit k: lane k holds records
    uint2* xsend;                // [NR][h]     sender side: {window stamp, packed counts} of the lane behind my port
    const int* grev;             // [NR][h]     far-end global-port index of each global link
    ChunkRec* rcin;       // [NR][rcin_cap]
    CreditRec* rkin;      // [NR][rkin_cap]
    unsigned* rcin_tail;  // [NR]
    unsigned* rkin_tail;  // [NR]
    // topology (read-only)
    const int* loc_off;     // [a*a+1]   CSR over (my lid, dest lid) -> local ports
    const int* loc_port;    // [..]
    const int* loc_dest;    // [a][intra_radix] destination local id of local port j of local router i
    const int* gs_port;     // [NR][h]   global ports in by-dest-gid order (get_connections_by_type)
    const int* gs_group;    // [NR][h]   destination group of gs_port
    const int* gdest;       // [NR][h]   destination router of global port intra_radix+i
    const int* routed_off;  // [G*G+1]   CSR over (src group, dest group)
    const int* routed_lid;  // [..]      local id of the router owning each global link, inter-file order
    // run control / results
    unsigned long long* counters;  // [64]: 0 events, 1 windows, 2 error code, 3 error info, 4.. per-type, 20.. cycles
    unsigned long long* winmax;    // [65536] profile builds: slowest LP activation of each window
};

// Group-sharded multi-GPU run: rank k owns the dragonfly groups [grp_lo[k], grp_lo[k+1]) with all
// their routers, terminals and servers.  Only three things cross ranks: chunks and credits over global links
// (written straight into the owner's router inbox through a CUDA-IPC mapping of its exchange region) and the
// svr->svr ACK statistics.  One window = local grid barrier, flag + minimum exchange over NVLink, local barrier.
#define DFD_MAX_RANKS 8
// lives at the start of every rank's exchange region.  mins[w%3][k]: rank k's earliest pending timestamp for
// window w, written by rank k when it closes window w-1; DFD_SYNC_EMPTY until then, DFD_SYNC_FAILED if rank k
// raised a device error.
struct DfdSync {
    unsigned long long mins[3][DFD_MAX_RANKS];
};
#define DFD_SYNC_EMPTY 0xFFFFFFFFFFFFFFFFULL
#define DFD_SYNC_FAILED 0xFFFFFFFFFFFFFFFEULL
struct DfdPeers {
    int rank, world;
    int grp_lo[DFD_MAX_RANKS + 1];
    int n0, n1, r0, r1;                 // owned node / router ranges
    char* base[DFD_MAX_RANKS];          // exchange-region base of every rank as mapped into this process
    char* local_base;                   // == base[rank]
};

enum { DFD_ERR_NONE = 0, DFD_ERR_NEV_OVERFLOW = 1, DFD_ERR_SQ_OVERFLOW, DFD_ERR_TQ_OVERFLOW, DFD_ERR_POOL_OVERFLOW,
       DFD_ERR_HEAP_OVERFLOW, DFD_ERR_INBOX_OVERFLOW, DFD_ERR_DEAD_END, DFD_ERR_WRONG_TERMINAL, DFD_ERR_BAD_VC, DFD_ERR_SELF_SEND,
       DFD_ERR_LATE_EVENT, DFD_ERR_PEER_TIMEOUT, DFD_ERR_PEER_ERROR, DFD_ERR_REASM_OVERFLOW, DFD_ERR_WHEEL_RANGE };
