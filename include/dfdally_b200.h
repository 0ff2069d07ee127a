/* dfdally_b200.h -- C-ABI of the B200 dragonfly-dally engine.
 *
 * Drop-in boundary for ONE hot path of CODES: the dragonfly-dally network method driven by the
 * synthetic traffic generator.  The call sequence mirrors main() of
 * src/network-workloads/model-net-synthetic-dragonfly-all.c:660-817; each entry point names the
 * reference interface it replaces (paths relative to the reference tree).  The engine is selected by the
 * reference's own .conf: LPGROUPS names "nw-lp", "modelnet_dragonfly_dally",
 * "modelnet_dragonfly_dally_router" (codes/model-net.h:93-95) and the dragonfly-dally PARAMS keys
 * (dragonfly-dally.cxx:2151-2757).  No torch types cross this boundary: plain pointers and sizes only.
 * The library fails loudly (non-zero return + dfd_last_error) when no CUDA device is usable; there is no
 * CPU fallback.
 */
#ifndef DFDALLY_B200_H
#define DFDALLY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfd_engine dfd_engine;

/* Command line of the synthetic driver: tw_optdef app_opt[] (synthetic:173-194) plus ROSS' --end. */
typedef struct dfd_synthetic_opts {
    int traffic;               /* --traffic: 1 UNIFORM, 2 RAND_PERM, 3 NEAREST_GROUP, 4 NEAREST_NEIGHBOR, 5 RANDOM_OTHER_GROUP */
    int num_messages;          /* --num_messages (default 20) */
    int payload_sz;            /* --payload_sz (default 2048) */
    long long rperm_threshold; /* --rperm_threshold (default 131072) */
    double arrival_time;       /* --arrival_time: constant inter-arrival in ns (0 = unset) */
    double load_per_svr;       /* --load_per_svr (0 = unset) */
    double warm_up_time;       /* --warm_up_time */
    double end_time;           /* --end (0 = the driver's 5-day default, synthetic:720-721) */
} dfd_synthetic_opts;

/* Global totals: what dragonfly_dally_report_stats (dally:3125-3192), svr_report_stats and
 * aggregate_svr_stats (synthetic:621-658) reduce with MPI_Reduce, plus engine counters. */
typedef struct dfd_summary {
    unsigned long long events;          /* "Net Events Processed": reference-equivalent committed events */
    unsigned long long events_by_type[16];
    unsigned long long windows;         /* group activations that executed at least one event (the engine has no global windows) */
    long long total_hops, finished_packets, finished_msgs, finished_chunks, total_msg_sz;
    double total_time, max_latency;
    long long packet_gen, packet_fin, local_same_router, local_same_group, remote;
    long long minimal_count, nonmin_count;
    long long svr_msgs_recvd;
    double svr_sum_latency, svr_max_latency, max_end_ts;
    double total_offered_load, total_observed_load;
    unsigned long long rng_draws;
    double window_ns;                   /* smallest cross-LP delay of the configuration (a credit delay by default) */
    unsigned long long h2d_bytes, d2h_bytes, device_bytes;
    int grid_blocks, block_threads, num_sms;
    int num_terminals, num_routers, radix;
    /* cycle accounting of the event-loop kernel per warp: cycles spent in passes over its groups that executed
     * events / that found nothing safe (averages over the warps), and the busiest warp's busy cycles */
    double cyc_busy, cyc_idle, cyc_busy_max_warp;
    unsigned long long activations;     /* activations that got past the channel-word poll */
    unsigned long long passes;          /* passes of the warps over their groups */
} dfd_summary;

void dfd_synthetic_opts_default(dfd_synthetic_opts* o);

/* handle management (the reference keeps this state in file-scope globals) */
int dfd_engine_create(dfd_engine** out);
void dfd_engine_destroy(dfd_engine* e);
const char* dfd_last_error(const dfd_engine* e);

/* configuration_load(argv[2], MPI_COMM_CODES, &config)            codes/configuration.h:89 */
int dfd_configuration_load(dfd_engine* e, const char* conf_path);
/* configuration_get_value(&config, "PARAMS", key, NULL, buf, len)  codes/configuration.h:142; returns the
 * length written (>0) or 0 when the key is absent */
int dfd_configuration_get_value(const dfd_engine* e, const char* section, const char* key, char* buf, int len);
/* model_net_register()                                             codes/model-net.h:175 -- scans LPGROUPS for the
 * dragonfly-dally LP names; fails if the config does not select this method */
int dfd_model_net_register(dfd_engine* e);
/* lp_type_lookup(name) != NULL                                     codes/lp-type-lookup.h:24 */
int dfd_lp_type_lookup(const dfd_engine* e, const char* lp_type_name);
/* codes_mapping_setup()                                            codes/codes_mapping.h -- LP counts, gid layout */
int dfd_codes_mapping_setup(dfd_engine* e);
/* codes_mapping_get_lp_count(group, 0, lp_type_name, NULL, 1)      codes/codes_mapping.h */
long long dfd_codes_mapping_get_lp_count(const dfd_engine* e, const char* lp_type_name);
/* codes_mapping_get_lpid_from_relative(rel, group, lp_type_name, NULL, 0) */
long long dfd_codes_mapping_get_lpid_from_relative(const dfd_engine* e, long long relative_id, const char* lp_type_name);
/* codes_mapping_get_lp_relative_id(gid, 0, 0) */
long long dfd_codes_mapping_get_lp_relative_id(const dfd_engine* e, long long gid);
/* model_net_configure(&num_nets)                                   codes/model-net.h:184 -- reads the PARAMS keys
 * (dragonfly_read_config dally:2151-3015), the intra/inter link files, builds the connection tables.
 * net_ids receives the ids in modelnet_order order (DRAGONFLY_DALLY=15, DRAGONFLY_DALLY_ROUTER=16 in the
 * reference's NETWORK_DEF list, codes/model-net.h:74-97). */
int dfd_model_net_configure(dfd_engine* e, int* net_ids, int max_ids, int* id_count);
/* the tw_optdef options of the synthetic driver */
int dfd_synthetic_configure(dfd_engine* e, const dfd_synthetic_opts* opts);

/* tw_run() (synthetic:803) split in three so that a caller can keep state resident:
 *   upload  : host -> device copy of the connection tables, device-side LP init (svr_init, terminal_dally_init,
 *             router_dally_init)
 *   run     : the event loop on the GPU; `cuda_stream` is a cudaStream_t (NULL = default stream); asynchronous
 *   finish  : waits, copies the per-LP statistics device -> host, runs the *_final handlers on the host */
int dfd_engine_upload(dfd_engine* e, void* cuda_stream);
int dfd_engine_reset(dfd_engine* e, void* cuda_stream); /* re-initialise LP state on the device (new run, same tables) */
int dfd_engine_run(dfd_engine* e, void* cuda_stream);
int dfd_engine_finish(dfd_engine* e, void* cuda_stream);
/* all of tw_run() in one call with host buffers on both sides */
int dfd_tw_run(dfd_engine* e, void* cuda_stream);
/* Ensembles (parameter sweeps: the usual way model-net-synthetic-* is used, one process per parameter point in the
 * reference).  A network of a few thousand LPs fills a small part of a B200, so several engines can run their event
 * loops at the same time on different CUDA streams of one GPU.  Call before dfd_engine_upload: `simulations` is the
 * number of engines that will share the GPU; this engine then sizes its persistent grid to 1/simulations of the
 * device's co-resident block slots (results do not depend on the grid).  simulations <= 1 restores the default. */
int dfd_engine_set_ensemble(dfd_engine* e, int simulations);

/* Group-sharded run over several GPUs of one box, one process per GPU (the reference's analogue is
 * codes_mapping()'s block mapping of LPs to MPI ranks, src/util/codes_mapping.c:78-86, with ROSS moving remote
 * events over MPI).  Rank k owns the dragonfly groups [k*G/world, (k+1)*G/world).  Sequence on every rank:
 *   set_partition -> engine_upload -> ipc_export -> (all-gather the 64-byte handles) -> ipc_import -> (barrier)
 *   -> engine_run -> (wait for the stream, barrier: every rank's event loop is over) -> engine_finish
 *   -> lp_io_flush / report_stats (rank 0).
 * engine_finish on rank 0 reads the per-LP statistics of EVERY partition over the CUDA-IPC mappings and runs the final
 * handlers; on the other ranks it only checks the rank's own control block.  export_partition / import_partition /
 * engine_finalize remain for callers that move the partitions themselves (e.g. over MPI): they are no-ops after a
 * gathering engine_finish. */
int dfd_engine_set_partition(dfd_engine* e, int rank, int world);
int dfd_engine_ipc_export(dfd_engine* e, void* handle64);
int dfd_engine_ipc_import(dfd_engine* e, const void* handles, int world);
/* 1 when every peer's exchange region is mapped: a re-upload of the same shape keeps the device arrays and the mappings,
 * so the handle exchange is needed only after the first upload (or a change of shape). */
int dfd_engine_ipc_ready(const dfd_engine* e);
/* All ranks of a sharded run as engines of ONE process on one GPU (each sized with dfd_engine_set_ensemble(world)): replaces
 * ipc_export / ipc_import; engines[k] is rank k.  Exercises the cross-rank path where only one GPU is available. */
int dfd_engine_attach_local_peers(dfd_engine* e, dfd_engine* const* engines, int world);
long long dfd_engine_partition_bytes(const dfd_engine* e);
/* ownership arithmetic: out4 = {first node, end node, first router, end router} of `rank` */
int dfd_partition_bounds(int num_groups, int routers_per_group, int cns_per_router, int rank, int world, int* out4);
int dfd_engine_export_partition(dfd_engine* e, void* buf, long long len);
int dfd_engine_import_partition(dfd_engine* e, const void* buf, long long len);
int dfd_engine_finalize(dfd_engine* e);

/* lp_io_prepare + lp_io_flush (codes/lp-io.h:23-32): writes dragonfly-link-stats, dragonfly-cn-stats,
 * model-net-category-*, synthetic-stats (+ raw-lp-stats with exact hex floats) into `directory` */
int dfd_lp_io_flush(dfd_engine* e, const char* directory);
/* model_net_report_stats(net_id) + svr_report_stats + aggregate_svr_stats (synthetic:808-810): the stdout
 * block, written into buf */
int dfd_model_net_report_stats(dfd_engine* e, char* buf, int len);
int dfd_get_summary(const dfd_engine* e, dfd_summary* out);

/* convenience: the whole main() of the synthetic driver */
int dfd_run_synthetic(const char* conf_path, const dfd_synthetic_opts* opts, const char* lp_io_dir, dfd_summary* out, char* err, int errlen);

/* small pure helpers exported for known-answer tests */
double dfd_bytes_to_ns(unsigned long long bytes, double gib_per_s); /* dally:1588-1599 */
int dfd_cuda_device_count(void);
/* sizeof(dfd_summary) / sizeof(dfd_synthetic_opts) of this build: lets a foreign-language binding (ctypes, cgo ...)
 * check its mirror of the two structs before it passes pointers to them */
void dfd_abi_sizes(int* summary_bytes, int* synthetic_opts_bytes);

#ifdef __cplusplus
}
#endif
#endif
