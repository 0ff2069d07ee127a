/* ross.h -- the slice of the ROSS API (ROSS-org/ROSS, an external dependency of CODES: CMakeLists.txt:78) that a CODES
 * model driver for the dragonfly-dally path is written against, declared for builds that link the B200 engine
 * (libcodes_compat_b200.so) instead of ROSS + libcodes.
 *
 * What is real here: command-line handling (tw_opt_add / tw_init), g_tw_ts_end, tw_run() -- which runs the whole
 * simulation on the GPU through dfd_tw_run (include/dfdally_b200.h) -- and tw_end().  What is declared only so that a
 * driver's LP handlers compile and link: tw_event_new / tw_event_send / tw_rand_* / tw_now.  The engine executes the
 * svr, model-net base, terminal and router LPs as device code; host handlers registered through
 * lp_type_register() are recorded by name and never called, and calling one of these functions aborts with a
 * message saying so.
 */
#ifndef CODES_COMPAT_ROSS_H
#define CODES_COMPAT_ROSS_H

#include <assert.h>
#include <errno.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

#include <mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double tw_stime;
typedef uint64_t tw_lpid;
typedef unsigned long tw_peid;
typedef unsigned long tw_kpid;
typedef int64_t tw_seed;
#define TW_STIME_DBL(x) (x)
#define TW_STIME_CRT(x) (x)
#define TW_STIME_ADD(x, y) ((x) + (y))
#define TW_STIME_CMP(x, y) (((x) < (y)) ? -1 : ((x) > (y)))
#define TW_STIME_MAX DBL_MAX

/* reverse-computation bit field of an event (unused by a conservative engine; handlers take a pointer to it) */
typedef struct tw_bf {
    unsigned int c0 : 1, c1 : 1, c2 : 1, c3 : 1, c4 : 1, c5 : 1, c6 : 1, c7 : 1, c8 : 1, c9 : 1, c10 : 1, c11 : 1, c12 : 1,
        c13 : 1, c14 : 1, c15 : 1, c16 : 1, c17 : 1, c18 : 1, c19 : 1, c20 : 1, c21 : 1, c22 : 1, c23 : 1, c24 : 1, c25 : 1,
        c26 : 1, c27 : 1, c28 : 1, c29 : 1, c30 : 1, c31 : 1;
} tw_bf;

typedef struct tw_rng_stream {
    long Ig[4], Lg[4], Cg[4]; /* CLCG4: initial, last and current seeds of the four component generators */
    unsigned long count;
} tw_rng_stream;
typedef tw_rng_stream tw_generator;

typedef struct tw_lp tw_lp;
typedef struct tw_pe tw_pe;
typedef struct tw_kp tw_kp;
typedef struct tw_event tw_event;

typedef void (*init_f)(void* sv, tw_lp* me);
typedef void (*pre_run_f)(void* sv, tw_lp* me);
typedef void (*event_f)(void* sv, tw_bf* cv, void* msg, tw_lp* me);
typedef void (*revent_f)(void* sv, tw_bf* cv, void* msg, tw_lp* me);
typedef void (*commit_f)(void* sv, tw_bf* cv, void* msg, tw_lp* me);
typedef void (*final_f)(void* sv, tw_lp* me);
typedef tw_peid (*map_f)(tw_lpid);
typedef void (*map_local_f)(void);
typedef tw_lp* (*map_custom_f)(tw_lpid);

typedef struct tw_lptype {
    init_f init;
    pre_run_f pre_run;
    event_f event;
    revent_f revent;
    commit_f commit;
    final_f final;
    map_f map;
    size_t state_sz;
} tw_lptype;

/* instrumentation hooks (types only: the engine has no ROSS instrumentation layer) */
typedef void (*ev_trace_f)(void* msg, tw_lp* lp, char* buffer, int* collect_flag);
typedef void (*model_stat_f)(void* sv, tw_lp* lp, char* buffer);
typedef void (*sample_event_f)(void* sv, tw_bf* cv, tw_lp* lp, void* sample);
typedef void (*sample_revent_f)(void* sv, tw_bf* cv, tw_lp* lp, void* sample);
typedef struct st_model_types {
    ev_trace_f ev_trace;
    size_t ev_sz;
    model_stat_f model_stat_fn;
    size_t mstat_sz;
    sample_event_f sample_event_fn;
    sample_revent_f sample_revent_fn;
    size_t sample_struct_sz;
} st_model_types;

/* state check-pointing of the rollback checker (types only) */
typedef void (*save_checkpoint_state_f)(void* into, void const* from);
typedef void (*clean_checkpoint_state_f)(void* state);
typedef bool (*check_states_f)(void* before, void* after);
typedef void (*print_lpstate_f)(FILE* out, char const* prefix, void* state);
typedef void (*print_checkpoint_state_f)(FILE* out, char const* prefix, void* state);
typedef void (*print_event_f)(FILE* out, char const* prefix, void* state, void* msg);
typedef struct crv_checkpointer {
    tw_lptype* lptype;
    size_t sz_storage;
    save_checkpoint_state_f save_lp;
    clean_checkpoint_state_f clean_lp;
    check_states_f check_lps;
    print_lpstate_f print_lp;
    print_checkpoint_state_f print_checkpoint;
    print_event_f print_event;
} crv_checkpointer;
void crv_add_custom_state_checkpoint(crv_checkpointer* chkptr);

typedef struct tw_event_sig {
    tw_stime recv_ts;
    tw_stime priority;
    tw_stime event_tiebreak[1];
    unsigned int tie_lineage_length;
} tw_event_sig;

struct tw_event {
    tw_lpid dest_lpid;
    tw_lp* src_lp;
    tw_stime recv_ts;
};
struct tw_pe {
    tw_peid id;
    void* pq;
    tw_event_sig GVT_sig;
    tw_stime GVT;
};
struct tw_kp {
    tw_kpid id;
    tw_pe* pe;
};
struct tw_lp {
    tw_lpid id;
    tw_lpid gid;
    tw_pe* pe;
    tw_kp* kp;
    void* cur_state;
    tw_lptype* type;
    tw_rng_stream* rng;
    st_model_types* model_types;
    unsigned int critical_path;
    tw_stime now;
};

/* command-line options (tw_opt_add + tw_init) */
typedef enum {
    TWOPTTYPE_GROUP = 1,
    TWOPTTYPE_ULONG,
    TWOPTTYPE_ULONGLONG,
    TWOPTTYPE_UINT,
    TWOPTTYPE_STIME,
    TWOPTTYPE_DOUBLE,
    TWOPTTYPE_CHAR,
    TWOPTTYPE_FLAG,
    TWOPTTYPE_SHOWHELP
} tw_opttype;
typedef struct tw_optdef {
    tw_opttype type;
    const char* name;
    const char* help;
    void* value;
} tw_optdef;
#define TWOPT_GROUP(h) {TWOPTTYPE_GROUP, NULL, (h), NULL}
#define TWOPT_ULONG(n, v, h) {TWOPTTYPE_ULONG, (n), (h), &(v)}
#define TWOPT_ULONGLONG(n, v, h) {TWOPTTYPE_ULONGLONG, (n), (h), &(v)}
#define TWOPT_UINT(n, v, h) {TWOPTTYPE_UINT, (n), (h), &(v)}
#define TWOPT_STIME(n, v, h) {TWOPTTYPE_STIME, (n), (h), &(v)}
#define TWOPT_DOUBLE(n, v, h) {TWOPTTYPE_DOUBLE, (n), (h), &(v)}
#define TWOPT_CHAR(n, v, h) {TWOPTTYPE_CHAR, (n), (h), &(v)}
#define TWOPT_FLAG(n, v, h) {TWOPTTYPE_FLAG, (n), (h), &(v)}
#define TWOPT_END() {(tw_opttype)0, NULL, NULL, NULL}

typedef enum {
    NO_SYNCH = 0,
    SEQUENTIAL = 1,
    CONSERVATIVE = 2,
    OPTIMISTIC = 3,
    OPTIMISTIC_DEBUG = 4,
    OPTIMISTIC_REALTIME = 5,
    SEQUENTIAL_ROLLBACK_CHECK = 6
} tw_synch;
typedef enum { LINEAR = 0, ROUND_ROBIN, CUSTOM } tw_lp_map;

extern tw_synch g_tw_synchronization_protocol; /* --sync / --synch is accepted and recorded; the engine is conservative */
extern tw_peid g_tw_mynode;
extern tw_stime g_tw_ts_end; /* --end */
extern tw_stime g_tw_lookahead;
extern unsigned int g_tw_nRNG_per_lp;
extern tw_lp** g_tw_lp;
extern tw_lpid g_tw_nlp;
extern tw_pe* g_tw_pe;
extern tw_kp** g_tw_kp;
extern tw_kpid g_tw_nkp;
extern size_t g_tw_msg_sz;
extern tw_lp_map g_tw_mapping;
extern unsigned int g_tw_events_per_pe;
extern map_local_f g_tw_custom_initial_mapping;
extern map_custom_f g_tw_custom_lp_global_to_local_map;
extern int g_st_ev_trace, g_st_model_stats, g_st_use_analysis_lps, g_st_ross_rank;
extern MPI_Comm MPI_COMM_ROSS;

#define TW_LOC __FILE__, __LINE__
void tw_error(const char* file, int line, const char* fmt, ...) __attribute__((noreturn));
void tw_warning(const char* file, int line, const char* fmt, ...);
int tw_output(tw_lp* lp, const char* fmt, ...);

void tw_opt_add(const tw_optdef* options);
void tw_init(int* argc, char*** argv);
void tw_define_lps(tw_lpid nlp, size_t msg_sz);
void tw_run(void); /* the whole simulation on the GPU: dfd_tw_run */
void tw_end(void);
tw_peid tw_nnodes(void);
void tw_lp_settype(tw_lpid lp, tw_lptype* type);
void st_model_settype(tw_lpid lp, st_model_types* model_types);

static inline tw_stime tw_now(tw_lp const* lp) { return lp->now; }
/* event and RNG calls of host LP handlers: link, never run (see the file comment) */
tw_event* tw_event_new(tw_lpid dest_gid, tw_stime offset_ts, tw_lp* sender);
void* tw_event_data(tw_event* event);
void tw_event_send(tw_event* event);
double tw_rand_unif(tw_rng_stream* g);
long tw_rand_integer(tw_rng_stream* g, long low, long high);
double tw_rand_exponential(tw_rng_stream* g, double lambda);
double tw_rand_reverse_unif(tw_rng_stream* g);

#ifdef __cplusplus
}
#endif
#endif
