/* codes_compat.h -- the B200 dragonfly-dally engine under the reference's OWN entry-point names.
 *
 * libcodes_compat_b200.so exports, with the reference's names, argument lists and return conventions, every libcodes /
 * ROSS function and object that main() of src/network-workloads/model-net-synthetic-dragonfly-all.c (:660-817) and its
 * LP handlers reference, over one process-wide engine (dfd_engine, include/dfdally_b200.h).  The UNMODIFIED reference
 * driver compiles against the reference's own codes/ headers plus include/codes_compat/{ross.h,mpi.h} and links against
 * this library instead of libcodes + ROSS + MPI (tests/test_compat_boundary.py builds and runs exactly
 * that).  This header declares the same interface for builds that do not have the reference headers; when they HAVE been
 * included first (their include guards are visible) the type definitions below are skipped and only the prototypes
 * remain -- the compiler then checks them against the reference's declarations.
 *
 * What runs where: configuration, LP-type lookup, mapping arithmetic, option parsing, lp-io and the statistics report
 * are host code that forwards into the engine; tw_run() executes the whole simulation on the GPU.  The per-event
 * members (model_net_event, the method structs' packet_event, tw_event_*) exist so that a driver's host LP handlers link;
 * the engine runs the svr / base / terminal / router LPs as device code and never calls host handlers, so these members
 * abort with a message when called.
 */
#ifndef CODES_COMPAT_H
#define CODES_COMPAT_H

#include <ross.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codes/configuration.h:78-198 ------------------------------------------------------------------------- */
#ifndef __CONFIGURATION_H__
typedef struct ConfigVTable* ConfigHandle;
extern ConfigHandle config; /* the process-wide configuration (configuration.c) */
#endif
int configuration_load(const char* filepath, MPI_Comm comm, ConfigHandle* handle);
int configuration_get_value(ConfigHandle* handle, const char* section_name, const char* key_name, const char* annotation, char* value,
                            size_t length);
int configuration_get_value_int(ConfigHandle* handle, const char* section_name, const char* key_name, const char* annotation, int* value);
int configuration_get_value_double(ConfigHandle* handle, const char* section_name, const char* key_name, const char* annotation,
                                   double* value);

/* ---- codes/lp-type-lookup.h:21-24 -------------------------------------------------------------------------- */
const tw_lptype* lp_type_lookup(const char* name);
void lp_type_register(const char* name, const tw_lptype* type);

/* ---- codes/codes_mapping.h:30-132 -------------------------------------------------------------------------- */
tw_peid codes_mapping(tw_lpid gid);
void codes_mapping_setup(void);
int codes_mapping_get_lp_count(const char* group_name, int ignore_repetitions, const char* lp_type_name, const char* annotation,
                               int ignore_annos);
int codes_mapping_get_lp_relative_id(tw_lpid gid, int group_wise, int annotation_wise);
tw_lpid codes_mapping_get_lpid_from_relative(int relative_id, const char* group_name, const char* lp_type_name, const char* annotation,
                                             int annotation_wise);
void codes_mapping_get_lp_info(tw_lpid gid, char* group_name, int* group_index, char* lp_type_name, int* lp_type_index, char* annotation,
                               int* rep_id, int* offset);

/* ---- codes/lp-io.h:16-32 ----------------------------------------------------------------------------------- */
#ifndef LP_IO_H
typedef char* lp_io_handle;
#define LP_IO_UNIQ_SUFFIX 1
#endif
int lp_io_prepare(char* directory, int flags, lp_io_handle* handle, MPI_Comm comm);
int lp_io_write(tw_lpid gid, char* identifier, int size, void* buffer);
int lp_io_flush(lp_io_handle handle, MPI_Comm comm);

/* ---- codes/model-net.h:120-376 ----------------------------------------------------------------------------- */
#ifndef MODELNET_H
typedef int model_net_event_return;
/* enum NETWORKS (codes/model-net.h:74-100): the two ids this engine implements */
enum { DRAGONFLY_DALLY = 15, DRAGONFLY_DALLY_ROUTER = 16 };
typedef struct model_net_request model_net_request;
typedef struct mn_sched_params_s mn_sched_params;
#endif
void model_net_register(void);
int* model_net_configure(int* id_count); /* malloc'ed array, caller frees */
void model_net_enable_sampling(tw_stime interval, tw_stime end);
model_net_event_return model_net_event(int net_id, char const* category, tw_lpid final_dest_lp, uint64_t message_size, tw_stime offset,
                                       int remote_event_size, void const* remote_event, int self_event_size, void const* self_event,
                                       tw_lp* sender);
void model_net_event_rc2(tw_lp* sender, model_net_event_return const* ret);
void model_net_report_stats(int net_id);

/* ---- codes/model-net-method.h:24-66, dragonfly-dally.cxx:10972-11024 --------------------------------------- */
#ifndef MODELNET_METHOD_H
struct model_net_method {
    uint64_t packet_size;
    void (*mn_configure)(void);
    void (*mn_register)(tw_lptype* base_type);
    tw_stime (*model_net_method_packet_event)(model_net_request const* req, uint64_t message_offset, uint64_t packet_size, tw_stime offset,
                                              mn_sched_params const* sched_params, void const* remote_event, void const* self_event,
                                              tw_lp* sender, int is_last_pckt, bool is_there_another_pckt_in_queue);
    void (*model_net_method_packet_event_rc)(tw_lp* sender);
    tw_stime (*model_net_method_recv_msg_event)(const char* category, tw_lpid final_dest_lp, tw_lpid src_mn_lp, uint64_t msg_size,
                                                int is_pull, uint64_t pull_size, tw_stime offset, int remote_event_size,
                                                const void* remote_event, tw_lpid src_lp, tw_lp* sender);
    void (*model_net_method_recv_msg_event_rc)(tw_lp* lp);
    const tw_lptype* (*mn_get_lp_type)(void);
    int (*mn_get_msg_sz)(void);
    void (*mn_report_stats)(void);
    void (*mn_collective_call)(char const* category, int message_size, int remote_event_size, const void* remote_event, tw_lp* sender);
    void (*mn_collective_call_rc)(int message_size, tw_lp* sender);
    event_f mn_sample_fn;
    revent_f mn_sample_rc_fn;
    init_f mn_sample_init_fn;
    final_f mn_sample_fini_fn;
    void (*mn_model_stat_register)(st_model_types* base_type);
    const st_model_types* (*mn_get_model_stat_types)(void);
    event_f mn_end_notif_fn;
    revent_f mn_end_notif_rc_fn;
    event_f cc_congestion_event_fn;
    revent_f cc_congestion_event_rc_fn;
    commit_f cc_congestion_event_commit_fn;
    crv_checkpointer* checkpointer;
};
#endif
extern struct model_net_method dragonfly_dally_method;        /* terminal LP "modelnet_dragonfly_dally" */
extern struct model_net_method dragonfly_dally_router_method; /* router LP "modelnet_dragonfly_dally_router" */

/* ---- engine access for callers that want more than the reference API offers -------------------------------- */
struct dfd_engine;
struct dfd_engine* codes_compat_engine(void);     /* the process-wide engine (created on first use) */
const char* codes_compat_last_error(void);
void codes_compat_set_stream(void* cuda_stream);  /* stream tw_run() launches on (default: the NULL stream) */

#ifdef __cplusplus
}
#endif
#endif
