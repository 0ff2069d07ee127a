// Shared host/device data layout of the asynchronous dragonfly-dally engine.
// Terminology follows the reference (terminal, router, chunk, credit, VC, port).
//
// Scheduling unit ("group"): one router LP plus the p compute nodes (svr + base LP + terminal, fused) that hang off
// it, owned by one warp.  LPs talk through single-writer FIFO lanes, one per directed link and record kind:
//   chunk lane  (u, port q) -> in-lane s of router d : 96-byte ChunkRec ring at d, written by u only
//   credit lane  d -> port q of router u              : 16-byte Key ring at u, written by d only
// Records of one lane are pushed in increasing (ts, src, seq) order (a port serialises its chunks, credits leave at
// the sender's event times plus a constant), so a router's pending events are the heads of its lanes plus one R_SEND
// timer per port.  Next to every lane the producer publishes a CHANNEL WORD: a lower bound on the timestamp of any
// record it has not pushed yet, with the lane's record count stamped into the low bits (one 8-byte store is both).
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
    EV_R_SNAPSHOT,  // router VC-occupancy snapshot (PARAMS router_buffer_snapshots, dally:4363-4437)
    EV_NTYPES
};

// enum ROUTING_ALGO / last_hop values of the reference (dragonfly-dally.cxx:576-598)
enum { DFD_MINIMAL = 1, DFD_NON_MINIMAL = 2, DFD_PROG_ADAPTIVE = 4 };
enum { DFD_HOP_GLOBAL = 1, DFD_HOP_LOCAL = 2, DFD_HOP_TERMINAL = 3 };
enum { DFD_SCORE_ALPHA = 1, DFD_SCORE_DELTA = 4 };
enum { DFD_TR_UNIFORM = 1, DFD_TR_RAND_PERM = 2, DFD_TR_NEAREST_GROUP = 3, DFD_TR_NEAREST_NEIGHBOR = 4, DFD_TR_RANDOM_OTHER_GROUP = 5 };

// event key: timestamp + deterministic tie-break (sender gid, sender's event-creation sequence number).
// Also the 16-byte credit record (T_BUFFER / R_BUFFER): the receiving port is the lane, the VC rides in the two top
// bits of src (gids are < 2^30).
struct __attribute__((aligned(16))) Key {
    double ts;
    uint32_t src;
    uint32_t seq;
};
#define DFD_GID_MASK 0x3FFFFFFFu

// 96-byte packed chunk event (R_ARRIVE / T_ARRIVE): the fields of terminal_dally_message
// (codes/net/dragonfly-dally.h:19-139) that are consumed downstream.  The first 16 bytes are the event key.
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

// router chunk pool slot: a chunk waiting in pending_msgs / queued_msgs.  Same 96 bytes as ChunkRec from offset 16
// on; the key bytes are reused for the list link and the routing decision.
struct __attribute__((aligned(16))) PoolSlot {
    uint32_t next;        // FIFO link / free-list link
    uint16_t out_port;    // vc_index chosen by this router
    uint8_t out_vc;       // output_chan chosen by this router
    uint8_t next_is_term; // next stop is the destination terminal
    uint32_t next_stop;   // router id or terminal id
    uint16_t in_lane;     // in-lane of this router the chunk arrived on (its credit goes back there)
    uint16_t pad;
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

// Lane cursors of a router, one entry per index j, read as in-lane j (chunks arriving) and as port j (credits
// arriving, chunks leaving).  All counters count records modulo 2^16; a ring index is counter & (capacity - 1).
struct __attribute__((aligned(16))) LaneCtl {
    uint16_t chead, ctail;  // chunk in-lane j: records consumed / records known to be there
    uint16_t khead, ktail;  // credit lane of port j
    uint16_t sent_c;        // chunks this router pushed out of port j
    uint16_t sent_k;        // credits this router pushed back on in-lane j
    uint32_t pad;
};

struct __attribute__((aligned(16))) RouterHdr {
    int32_t rng[4];  // CLCG4 stream gid*3+0
    uint32_t seq;
    uint32_t pool_nfree; // entries on the free-slot stack
    uint32_t gid;        // codes_mapping gid of this router
    uint16_t lid, grp;   // id inside the group, group id
    double clock;        // lower bound on the timestamp of this router's next event (< 0: never activated)
    double last_ts;      // key of the last executed event (causality check)
    uint32_t last_src, last_seq;
    unsigned long long rng_draws;
    uint32_t more;       // the last activation stopped on its event budget: safe events may remain
    uint32_t snap_next;  // R_SNAPSHOT events executed so far = index of the next one in the time-sorted list (DfdArrays::snap_ts)
    unsigned long long activations;  // activations that executed at least one event
    double ev_min;       // earliest pending event of the group (router and terminals) when the last activation ended
    double pad1;
};

// what the idle poll of a group reads besides the channel words
struct __attribute__((aligned(16))) PollHdr {
    double ev_min;   // RouterHdr::ev_min
    uint32_t flags;  // bit0: the last activation stopped on its event budget, bit1: never activated
    uint32_t pad;
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
// A message between the two stages of handle_new_msg (model-net-lp.c:706-716): the second MN_BASE_NEW_MSG is
// scheduled at next_available_time, which only grows, so these events form a FIFO of their own.
struct MsgReq {
    double ts;
    uint32_t seq;
    uint32_t dst_svr;
    uint32_t msg_size;
    uint32_t pad;
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
struct __attribute__((aligned(128))) NodeState {
    // --- what the router's bounds pass reads of every terminal in every round: one 32-byte sector at the start of a line
    double self_next;            // earliest timestamp among the self-scheduled events (nev list, mq); +inf: none
    double terminal_available_time;
    uint32_t cin_head, kin_head; // records consumed from the chunk / credit ring
    uint32_t tq_head, tq_count;  // terminal_msgs ring
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
    uint32_t nev;                // number of pending self events
    uint32_t grp;        // group of `router`
    double last_buf_full, busy_time;
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
    uint32_t mq_head, mq_count;  // FIFO of second-stage MN_BASE_NEW_MSG events
    double last_ts;              // key of the last executed event (causality check)
    uint32_t last_src, last_seq;
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
    double min_delay, ts_end;  // smallest cross-LP delay (a credit delay by default); end of the simulation
    double hyst_local, hyst_global;  // a chunk bound is re-published only after it has moved by this much (or carries new records)
    // host-precomputed values of bytes_to_ns()/reciprocals (same IEEE operations as the device would do)
    double inj_chunk_cn, inj_chunk_local, inj_chunk_global;  // bytes_to_ns(chunk_size, bw)
    double inj_pay_cn, inj_pay_local, inj_pay_global;        // bytes_to_ns(payload_sz % chunk_size, bw)
    double inj_packet_cn;                                    // bytes_to_ns(payload_sz, cn_bw) (nic_ts of packet_generate)
    double inv_cn_bw;                                        // 1 / cn_bandwidth
    // lower bounds used by the channel words: the smallest serialisation time any chunk of this workload can have on
    // a link of each kind (plus router_delay for router ports, added first as router_packet_send does)
    double lb_injrd_cn, lb_injrd_local, lb_injrd_global;     // router port: min injection delay + router_delay
    double lb_inj_term;                                      // terminal NIC: min injection delay
    unsigned long long magic_p, magic_a;                     // floor(2^64/d)+1: n/d == __umul64hi(n, magic) for n < 2^32, d > 1
    int traffic, num_msgs, payload_sz;
    long long rperm_threshold;
    double warm_up_time, mean_interval;
    int num_nodes, nodes_per_grp, node_eager_limit, noop_bypass;
    int lps_per_rep, off_svr, off_term, off_rtr;
    // ring capacities (powers of two) and pool capacities
    int ncin_cap, nkin_cap, nsq_cap, ntq_cap, nev_cap;
    int nreasm_cap;  // per-terminal reassembly table entries (0: every message is one chunk)
    int cap_l, cap_g, cap_tc, cap_tk;  // lane capacities: local link, global link, terminal -> router chunks, terminal -> router credits
    int cslots, kslots;                // records per router in ring_c / ring_k
    int rpool_cap;
    int ext;                           // intra_radix + global_radix: lanes / ports [0, ext) connect routers, [ext, radix) terminals
    int idle_sleep_ns;                 // pause of a warp after a pass over its groups that executed nothing (0: poll again at once)
    int act_budget;                    // router events after which an activation publishes its bounds even if more would be safe
    int npubq, npub_blocks;            // sharded runs: publisher queues / publisher blocks at the end of the grid (0: every activation ends in its own system-scope fence)
    int nsnap, nsnap_all;              // R_SNAPSHOT events per router: those before ts_end (the ones that exist) / configured
    int prefetch;                      // an activation that gets past the poll prefetches the router's port state into L1 (partitions larger than L2)
};

struct DfdSeedConsts {  // CLCG4 constants: moduli, default seed, a^(2^(31+41)) mod m (stream jump)
    unsigned long long m[4], seed0[4], avw[4];
};

// run control block at the start of every rank's exchange region; rank 0's copy carries the global counters
// Termination: `created` counts the events this rank brought into being (self events, timers, records pushed to any
// rank; added BEFORE the release fence that publishes them), `consumed` the events it executed (added after that
// fence).  Both only grow.  A monitor that reads every rank's `consumed` first and every rank's `created` afterwards
// and finds the sums equal has seen a moment with nothing pending anywhere: the simulation is over.
struct DfdCtrl {
    unsigned long long created;
    unsigned long long consumed;
    unsigned long long error;     // first device error code (other ranks receive DFD_ERR_PEER_ERROR)
    unsigned long long error_info;
    unsigned long long done;      // set by this rank's monitor warp
    unsigned long long pad[27];
};

// Publisher queues (sharded runs).  fence.release.sys costs 2,800 cycles on B200 with nothing outstanding and 10,000 when a
// thousand warps issue it concurrently (tools/ubench/fence_sys.cu) -- a fifth of an 8-GPU run when every activation ends in one.
// So an activation ends in a gpu-scope release fence (700 cycles) and hands the channel words that go to ANOTHER rank to a
// publisher warp: {target | tag, word} entries in a multi-producer ring, one 16-byte relaxed store each.  The publisher
// takes a batch of entries, executes ONE fence.acq_rel.sys -- fences are cumulative: the records the producers wrote before their
// own release fences are ordered before it -- and stores the words with st.relaxed.sys.  Words of one source group always
// travel through the same queue, so they reach their target in order.
#define DFD_PUBQ_CAP 8192  // entries per queue (power of two)
struct PubCtl {
    unsigned long long reserve;  // next ticket (producers, atomicAdd)
    unsigned long long pad0[15];
    unsigned long long head;     // entries the publisher is done with (back-pressure)
    unsigned long long pad1[15];
};

// device arrays.  Everything indexed by router / node uses the LOCAL index inside this rank's partition.
struct DfdArrays {
    // nodes
    NodeState* node;        // [NTl]
    NodeEv* nev;            // [NTl][nev_cap]
    SchedReq* sq;           // [NTl][nsq_cap]
    MsgReq* mq;             // [NTl][nsq_cap]
    TermChunk* tq;          // [NTl][ntq_cap]
    ChunkRec* ncin;         // [NTl][ncin_cap]   router -> terminal chunks
    Key* nkin;              // [NTl][nkin_cap]   router -> terminal credits
    uint4* mtab;            // [NTl][nreasm_cap] rank_tbl: {message_id, sender svr, remaining_packets, has_remote}
    uint4* ptab;            // [NTl][nreasm_cap] remaining_sz_packets: {packet_id, source terminal, remaining bytes, -}
    unsigned* ncin_tail;    // [NTl]
    unsigned* nkin_tail;    // [NTl]
    unsigned* ack_count;             // [NTmax] exchange region
    unsigned long long* ack_first;   // [NTmax]
    unsigned long long* ack_last;    // [NTmax]
    // routers: private state
    RouterHdr* rh;        // [NRl]
    PollHdr* pollhdr;     // [NRl]
    PortState* port;      // [NRl][radix]
    PortQ* portq;         // [NRl][radix]
    Key* skey;            // [NRl][radix]  pending R_SEND of the port (ts = +inf: none)
    Key* ckey;            // [NRl][radix]  head of chunk in-lane j, or {bound, 0, 0} while the lane is empty
    Key* kkey;            // [NRl][radix]  head of the credit lane of port j, or {bound, 0, 0}
    LaneCtl* lane;        // [NRl][radix]
    unsigned long long* seen_c;  // [NRl][radix] channel words as last polled
    unsigned long long* seen_k;
    unsigned long long* pub_c;   // [NRl][radix] channel words as last published
    unsigned long long* pub_k;
    unsigned long long** tgt_c;  // [NRl][radix] where the chunk word of port j goes (neighbour's cw_c slot, possibly on a peer GPU; 0: no link)
    unsigned long long** tgt_k;  // [NRl][radix] where the credit word of in-lane j goes (neighbour's cw_k slot)
    unsigned char* nq;           // [NRl][radix][radix] overflowed (queued_msgs) chunks by (in-lane, out port)
    unsigned long long* qbits;   // [NRl][radix][2]     bit q of in-lane s: nq[s][q] > 0
    double* iblk;                // [NRl][2][p]         bounds of the terminal lanes (chunks, credits) for the router phase
    PoolSlot* pool;       // [NRl][rpool_cap]
    unsigned* pfree;      // [NRl][rpool_cap]   stack of free pool slots
    // routers: exchange region (written by neighbours, possibly from another GPU)
    unsigned long long* cw_c;    // [NRmax][radix] chunk channel words (by in-lane)
    unsigned long long* cw_k;    // [NRmax][radix] credit channel words (by port)
    ChunkRec* ring_c;            // [NRmax][cslots]
    Key* ring_k;                 // [NRmax][kslots]
    DfdCtrl* ctrl;
    ulonglong2* pubq;       // [npubq][DFD_PUBQ_CAP] {target pointer | tag in the low 3 bits, channel word}
    PubCtl* pubctl;         // [npubq]
    int4* snap;             // [NRmax][nsnap_all][radix] vc_occupancy[port][0..3] captured by R_SNAPSHOT number i (statistics: exchange region)
    const double* snap_ts;  // [nsnap]  snapshot times before ts_end, sorted by (time, configured index)
    const unsigned* snap_id; // [nsnap] configured index of each (= packet_ID = sequence number of the event, dally:4407-4414)
    const uint2* lgeo;      // [radix][2]  {ring offset, capacity - 1} of chunk in-lane j and of credit lane j (same for every router)
    // topology (read-only, global router ids)
    const int* loc_off;     // [a*a+1]   CSR over (my lid, dest lid) -> local ports
    const int* loc_port;    // [..]
    const int* loc_dest;    // [a][intra_radix] destination local id of local port j of local router i
    const int* lout_lane;   // [a][intra_radix] in-lane of that destination my local port feeds
    const int* lin_src;     // [a][intra_radix] in-lane s of local router i: source lid << 16 | its port, or -1
    const int* gs_port;     // [NR][h]   global ports in by-dest-gid order (get_connections_by_type)
    const int* gs_group;    // [NR][h]   destination group of gs_port
    const int* gdest;       // [NR][h]   destination router of global port intra_radix+i
    const int* gout_lane;   // [NR][h]   in-lane (absolute index) of that destination my global port feeds
    const int* gin_src;     // [NR][h]   source router of global in-lane intra_radix+i, or -1
    const int* gin_port;    // [NR][h]   its (absolute) port
    const int* routed_off;  // [G*G+1]   CSR over (src group, dest group)
    const int* routed_lid;  // [..]      local id of the router owning each global link, inter-file order
    // run control / results
    unsigned long long* counters;  // [64]: 0 events, 1 productive activations, 2 error code, 3 error info, 4.. per-type, 20.. cycles
};

// Group-sharded multi-GPU run: rank k owns the dragonfly groups [grp_lo[k], grp_lo[k+1]) with all
// their routers, terminals and servers.  Only three things cross ranks: chunks and credits over global links
// (records and channel words written straight into the owner's exchange region through a CUDA-IPC mapping) and the
// svr->svr ACK statistics.  There is no barrier and no collective: a router that owns a global link simply sees its
// peer's channel word like any other neighbour's.
#define DFD_MAX_RANKS 8
struct DfdPeers {
    int rank, world;
    int grp_lo[DFD_MAX_RANKS + 1];
    int n0, n1, r0, r1;                 // owned node / router ranges (global ids)
    char* base[DFD_MAX_RANKS];          // exchange-region base of every rank as mapped into this process
    char* local_base;                   // == base[rank]
    unsigned long long local_bytes;     // size of this rank's exchange region: a channel-word target outside it lives on another rank
};

enum { DFD_ERR_NONE = 0, DFD_ERR_NEV_OVERFLOW = 1, DFD_ERR_SQ_OVERFLOW, DFD_ERR_TQ_OVERFLOW, DFD_ERR_POOL_OVERFLOW,
       DFD_ERR_HEAP_OVERFLOW, DFD_ERR_INBOX_OVERFLOW, DFD_ERR_DEAD_END, DFD_ERR_WRONG_TERMINAL, DFD_ERR_BAD_VC, DFD_ERR_SELF_SEND,
       DFD_ERR_LATE_EVENT, DFD_ERR_PEER_TIMEOUT, DFD_ERR_PEER_ERROR, DFD_ERR_REASM_OVERFLOW, DFD_ERR_WHEEL_RANGE, DFD_ERR_STALL };

// channel words: the low DFD_STAMP_BITS of the bound's bit pattern are replaced by the lane's record count
#define DFD_STAMP_BITS 12
#define DFD_STAMP_MASK 0xFFFULL
#define DFD_TS_BIG 1e300  // "never": a finite stand-in for +inf that survives stamping
