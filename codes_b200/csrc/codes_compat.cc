// codes_compat.cc -- libcodes_compat_b200.so: the reference's entry-point names (include/codes_compat.h) as thin forwards
// into the C-ABI of the B200 engine (include/dfdally_b200.h) over one process-wide engine handle.
//
// The reference keeps this state in file-scope globals of libcodes and ROSS (`config`, the lp-type table, the mapping
// tables, g_tw_*); here one dfd_engine plays that role.  Host-callable entry points do real work; the per-event ones
// (model_net_event, tw_event_*, tw_rand_*, the method structs' packet_event) exist so that a driver's LP handlers link
// and abort when called, because the engine runs every LP of this path as device code.
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/codes_compat.h"
#include "../../include/dfdally_b200.h"

namespace {
struct LpType {
    std::string name;
    const tw_lptype* type;
};
struct Compat {
    dfd_engine* eng = nullptr;
    std::string last_error;
    std::vector<const tw_optdef*> opt_tables;
    std::vector<LpType> lp_types;
    void* stream = nullptr;
    bool registered = false, mapped = false, configured = false, ran = false;
    double wall_run_s = 0.0;
};
Compat& G() {
    static Compat g;
    return g;
}
dfd_engine* engine() {
    Compat& g = G();
    if (!g.eng && dfd_engine_create(&g.eng) != 0) {
        fprintf(stderr, "codes_compat: cannot create the engine\n");
        exit(1);
    }
    return g.eng;
}
// libcodes reports fatal conditions with tw_error(); so does this layer
[[noreturn]] void die(const char* what) {
    Compat& g = G();
    const char* msg = g.eng ? dfd_last_error(g.eng) : "";
    g.last_error = std::string(what) + ": " + msg;
    fprintf(stderr, "codes_compat: %s\n", g.last_error.c_str());
    exit(1);
}
[[noreturn]] void device_side(const char* fn) {
    fprintf(stderr,
            "codes_compat: %s() was called on the host.  The B200 engine executes the svr, model-net base, terminal and router LPs of the "
            "dragonfly-dally path as device code; host LP handlers registered through lp_type_register() are never run.\n",
            fn);
    abort();
}
const tw_optdef* find_opt(const char* name) {
    for (const tw_optdef* t : G().opt_tables)
        for (const tw_optdef* o = t; o->type; o++)
            if (o->name && !strcmp(o->name, name)) return o;
    return nullptr;
}
double opt_number(const char* name, double dflt) {
    const tw_optdef* o = find_opt(name);
    if (!o || !o->value) return dflt;
    switch (o->type) {
    case TWOPTTYPE_UINT:
    case TWOPTTYPE_FLAG: return (double)*(unsigned int*)o->value;
    case TWOPTTYPE_ULONG: return (double)*(unsigned long*)o->value;
    case TWOPTTYPE_ULONGLONG: return (double)*(unsigned long long*)o->value;
    case TWOPTTYPE_STIME:
    case TWOPTTYPE_DOUBLE: return *(double*)o->value;
    default: return dflt;
    }
}
void set_opt(const tw_optdef* o, const char* text) {
    switch (o->type) {
    case TWOPTTYPE_UINT: *(unsigned int*)o->value = (unsigned int)strtoul(text, nullptr, 10); break;
    case TWOPTTYPE_FLAG: *(unsigned int*)o->value = 1; break;
    case TWOPTTYPE_ULONG: *(unsigned long*)o->value = strtoul(text, nullptr, 10); break;
    case TWOPTTYPE_ULONGLONG: *(unsigned long long*)o->value = strtoull(text, nullptr, 10); break;
    case TWOPTTYPE_STIME:
    case TWOPTTYPE_DOUBLE: *(double*)o->value = atof(text); break;
    case TWOPTTYPE_CHAR: strcpy((char*)o->value, text); break;  // ROSS copies without a bound as well; drivers size these 256+
    default: break;
    }
}
const char* LPGROUP = "MODELNET_GRP";
// LPs of one type inside ONE repetition of the LP group: their gids are consecutive inside a repetition and jump at its end
long long count_per_rep(dfd_engine* e, const char* lp_type_name) {
    const long long total = dfd_codes_mapping_get_lp_count(e, lp_type_name);
    if (total <= 0) return 0;
    const long long g0 = dfd_codes_mapping_get_lpid_from_relative(e, 0, lp_type_name);
    long long k = 1;
    while (k < total && dfd_codes_mapping_get_lpid_from_relative(e, k, lp_type_name) == g0 + k) k++;
    return k;
}
}  // namespace

extern "C" {

// ---- ROSS globals the driver reads or writes ---------------------------------------------------------------
tw_synch g_tw_synchronization_protocol = SEQUENTIAL;
tw_peid g_tw_mynode = 0;
tw_stime g_tw_ts_end = 100000.0;  // ROSS' compiled-in default; the synthetic driver replaces it by 5 days (synthetic:720-721)
tw_stime g_tw_lookahead = 0.005;
unsigned int g_tw_nRNG_per_lp = 1;
tw_lp** g_tw_lp = nullptr;
tw_lpid g_tw_nlp = 0;
tw_pe* g_tw_pe = nullptr;
tw_kp** g_tw_kp = nullptr;
tw_kpid g_tw_nkp = 1;
size_t g_tw_msg_sz = 0;
tw_lp_map g_tw_mapping = CUSTOM;
unsigned int g_tw_events_per_pe = 0;
map_local_f g_tw_custom_initial_mapping = nullptr;
map_custom_f g_tw_custom_lp_global_to_local_map = nullptr;
int g_st_ev_trace = 0, g_st_model_stats = 0, g_st_use_analysis_lps = 0, g_st_ross_rank = 1;
MPI_Comm MPI_COMM_ROSS = MPI_COMM_WORLD;
MPI_Comm MPI_COMM_CODES = MPI_COMM_WORLD;        // codes/codes.h:21
ConfigHandle config = nullptr;                   // codes/configuration.h:221
char g_nm_link_failure_filepath[1024] = {0};    // codes/model-net.h:45 (link failures are out of scope: DESIGN.md 7)

struct dfd_engine* codes_compat_engine(void) { return engine(); }
const char* codes_compat_last_error(void) { return G().last_error.c_str(); }
void codes_compat_set_stream(void* cuda_stream) { G().stream = cuda_stream; }

// ---- ROSS: messages, options, run ----------------------------------------------------------------------------
void tw_error(const char* file, int line, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "error: %s:%d: ", file, line);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(1);
}
void tw_warning(const char* file, int line, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "warning: %s:%d: ", file, line);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}
int tw_output(tw_lp*, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vprintf(fmt, ap);
    va_end(ap);
    return n;
}
void tw_opt_add(const tw_optdef* options) { G().opt_tables.push_back(options); }

// tw_init: consumes the "--name=value" arguments that belong to ROSS or to a registered option table.  What is left is
// argv[0], "--" and the positional arguments, which is why the reference driver finds its .conf in argv[2].
void tw_init(int* argc, char*** argv) {
    char** av = *argv;
    int out = 1;
    bool positional = false;
    for (int i = 1; i < *argc; i++) {
        const char* a = av[i];
        if (positional || strncmp(a, "--", 2) != 0) {
            av[out++] = av[i];
            continue;
        }
        if (!strcmp(a, "--")) {
            positional = true;
            av[out++] = av[i];
            continue;
        }
        std::string name(a + 2), val;
        size_t eq = name.find('=');
        if (eq != std::string::npos) {
            val = name.substr(eq + 1);
            name.resize(eq);
        }
        if (name == "sync" || name == "synch") {
            g_tw_synchronization_protocol = (tw_synch)atoi(val.c_str());  // recorded; the engine is conservative in every mode
        } else if (name == "end") {
            g_tw_ts_end = atof(val.c_str());
        } else if (const tw_optdef* o = find_opt(name.c_str())) {
            set_opt(o, val.c_str());
        } else if (name == "help") {
            for (const tw_optdef* t : G().opt_tables)
                for (const tw_optdef* o = t; o->type; o++)
                    if (o->type == TWOPTTYPE_GROUP)
                        printf("\n%s\n", o->help);
                    else
                        printf("  --%-22s %s\n", o->name, o->help ? o->help : "");
            exit(0);
        } else {
            // the remaining ROSS kernel options (--extramem, --nkp, --gvt-interval ...) tune Time Warp and have no meaning here
            fprintf(stderr, "codes_compat: ROSS option --%s ignored by the B200 engine\n", name.c_str());
        }
    }
    *argc = out;
}
void tw_define_lps(tw_lpid nlp, size_t msg_sz) {
    g_tw_nlp = nlp;
    g_tw_msg_sz = msg_sz;
}
tw_peid tw_nnodes(void) { return 1; }
void tw_lp_settype(tw_lpid, tw_lptype*) {}
void st_model_settype(tw_lpid, st_model_types*) {}
void crv_add_custom_state_checkpoint(crv_checkpointer*) {}

// tw_run(): the synthetic driver's options are read from the option table it registered (synthetic:173-194), the run
// length from g_tw_ts_end, then the whole simulation runs on the GPU.
void tw_run(void) {
    Compat& g = G();
    if (!g.configured) die("tw_run() before model_net_configure()");
    dfd_synthetic_opts o;
    dfd_synthetic_opts_default(&o);
    o.traffic = (int)opt_number("traffic", o.traffic);
    o.num_messages = (int)opt_number("num_messages", o.num_messages);
    o.payload_sz = (int)opt_number("payload_sz", o.payload_sz);
    o.rperm_threshold = (long long)opt_number("rperm_threshold", (double)o.rperm_threshold);
    o.arrival_time = opt_number("arrival_time", o.arrival_time);
    o.load_per_svr = opt_number("load_per_svr", o.load_per_svr);
    o.warm_up_time = opt_number("warm_up_time", o.warm_up_time);
    // 5 days is the driver's own default (it overwrites ROSS' 100000.0); 0 selects the same default inside the engine
    o.end_time = g_tw_ts_end;
    if (dfd_synthetic_configure(g.eng, &o)) die("synthetic options");
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (dfd_tw_run(g.eng, g.stream)) die("tw_run");
    clock_gettime(CLOCK_MONOTONIC, &t1);
    g.wall_run_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    g.ran = true;
}
// tw_end(): ROSS prints its run statistics here; the line the CODES tests read is "Net Events Processed"
void tw_end(void) {
    Compat& g = G();
    if (g.ran) {
        dfd_summary s;
        dfd_get_summary(g.eng, &s);
        printf("\n\t: Running Time = %.4f seconds\n", g.wall_run_s);
        printf("\nTW Library Statistics:\n");
        printf("\t%-50s %11llu\n", "Net Events Processed", (unsigned long long)s.events);
        printf("\t%-50s %11.2f\n", "Event Rate (events/sec)", g.wall_run_s > 0 ? (double)s.events / g.wall_run_s : 0.0);
    }
    if (g.eng) dfd_engine_destroy(g.eng);
    g.eng = nullptr;
}
tw_event* tw_event_new(tw_lpid, tw_stime, tw_lp*) { device_side("tw_event_new"); }
void* tw_event_data(tw_event*) { device_side("tw_event_data"); }
void tw_event_send(tw_event*) { device_side("tw_event_send"); }
double tw_rand_unif(tw_rng_stream*) { device_side("tw_rand_unif"); }
long tw_rand_integer(tw_rng_stream*, long, long) { device_side("tw_rand_integer"); }
double tw_rand_exponential(tw_rng_stream*, double) { device_side("tw_rand_exponential"); }
double tw_rand_reverse_unif(tw_rng_stream*) { device_side("tw_rand_reverse_unif"); }

// ---- single-rank MPI ----------------------------------------------------------------------------------------
int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Initialized(int* flag) {
    *flag = 1;
    return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) {
    *rank = 0;
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int* size) {
    *size = 1;
    return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    if (s != r) memcpy(r, s, (size_t)n * (size_t)t);
    return MPI_SUCCESS;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
    if (s != r) memcpy(r, s, (size_t)n * (size_t)t);
    return MPI_SUCCESS;
}
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) { exit(code ? code : 1); }
double MPI_Wtime(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
void codes_comm_update(void) {}  // codes/codes.h:89: MPI_COMM_CODES = MPI_COMM_ROSS

// ---- configuration (codes/configuration.h) -------------------------------------------------------------------
int configuration_load(const char* filepath, MPI_Comm, ConfigHandle* handle) {
    Compat& g = G();
    if (dfd_configuration_load(engine(), filepath)) {
        g.last_error = dfd_last_error(g.eng);
        fprintf(stderr, "configuration_load: %s\n", g.last_error.c_str());
        return 1;  // the reference returns non-zero and leaves the decision to the caller
    }
    if (handle) *handle = (ConfigHandle)g.eng;  // opaque to callers; the getters below go through the engine
    return 0;
}
// returns the length of the value (as the reference does) or 0 when section / key do not exist
int configuration_get_value(ConfigHandle*, const char* section_name, const char* key_name, const char*, char* value, size_t length) {
    return dfd_configuration_get_value(engine(), section_name, key_name, value, (int)length);
}
int configuration_get_value_int(ConfigHandle*, const char* section_name, const char* key_name, const char*, int* value) {
    char buf[256];
    if (dfd_configuration_get_value(engine(), section_name, key_name, buf, sizeof buf) <= 0) return 1;
    char* end = nullptr;
    errno = 0;
    long v = strtol(buf, &end, 10);
    if (end == buf || errno) return 1;
    *value = (int)v;
    return 0;
}
int configuration_get_value_double(ConfigHandle*, const char* section_name, const char* key_name, const char*, double* value) {
    char buf[256];
    if (dfd_configuration_get_value(engine(), section_name, key_name, buf, sizeof buf) <= 0) return 1;
    char* end = nullptr;
    double v = strtod(buf, &end);
    if (end == buf) return 1;
    *value = v;
    return 0;
}

// ---- lp-type-lookup (codes/lp-type-lookup.h) -----------------------------------------------------------------
void lp_type_register(const char* name, const tw_lptype* type) {
    for (LpType& t : G().lp_types)
        if (t.name == name) {
            t.type = type;
            return;
        }
    G().lp_types.push_back({name, type});
}
const tw_lptype* lp_type_lookup(const char* name) {
    for (LpType& t : G().lp_types)
        if (t.name == name) return t.type;
    return nullptr;
}
void st_model_type_register(const char*, const st_model_types*) {}

// ---- model-net (codes/model-net.h) ---------------------------------------------------------------------------
static const tw_lptype device_lp_type = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (map_f)codes_mapping, 0};
static const tw_lptype* dally_cn_lp_type(void) { return &device_lp_type; }
static const tw_lptype* dally_router_lp_type(void) { return &device_lp_type; }
static int dally_msg_sz(void) { return 0; }  // events never exist as host messages
static void dally_configure(void) {
    int ids[4], n = 0;
    if (dfd_model_net_configure(engine(), ids, 4, &n)) die("dragonfly_dally_configure");
    G().configured = true;
}
static void dally_register(tw_lptype* base_type) { lp_type_register("modelnet_dragonfly_dally", base_type ? base_type : &device_lp_type); }
static void dally_router_register(tw_lptype* base_type) {
    lp_type_register("modelnet_dragonfly_dally_router", base_type ? base_type : &device_lp_type);
}
static void dally_report_stats(void) { model_net_report_stats(DRAGONFLY_DALLY); }
static tw_stime dally_packet_event(model_net_request const*, uint64_t, uint64_t, tw_stime, mn_sched_params const*, void const*, void const*,
                                   tw_lp*, int, bool) {
    device_side("dragonfly_dally_packet_event");
}
static void dally_packet_event_rc(tw_lp*) { device_side("dragonfly_dally_packet_event_rc"); }

// field order of codes/model-net-method.h:24-66; members this engine has no host-side counterpart for stay NULL, as
// they do for the reference's router method (dragonfly-dally.cxx:10999-11024)
struct model_net_method dragonfly_dally_method = {
    0,                        // packet_size: filled by model_net_configure from PARAMS/packet_size
    dally_configure,          // mn_configure
    dally_register,           // mn_register
    dally_packet_event,       // model_net_method_packet_event (device side)
    dally_packet_event_rc,    // model_net_method_packet_event_rc (no reverse computation: conservative engine)
    nullptr, nullptr,         // recv_msg_event(_rc)
    dally_cn_lp_type,         // mn_get_lp_type
    dally_msg_sz,             // mn_get_msg_sz
    dally_report_stats,       // mn_report_stats
    nullptr, nullptr,         // collective calls
    nullptr, nullptr, nullptr, nullptr,  // sampling
    nullptr, nullptr,         // model-stat types
    nullptr, nullptr,         // end-of-simulation notification
    nullptr, nullptr, nullptr,  // congestion control
    nullptr,                  // checkpointer
};
struct model_net_method dragonfly_dally_router_method = {
    0,       nullptr, dally_router_register, nullptr, nullptr, nullptr, nullptr, dally_router_lp_type, dally_msg_sz, nullptr, nullptr, nullptr,
    nullptr, nullptr, nullptr,               nullptr, nullptr, nullptr, nullptr, nullptr,              nullptr,      nullptr, nullptr, nullptr,
};

void model_net_register(void) {
    Compat& g = G();
    if (dfd_model_net_register(engine())) die("model_net_register");  // fails unless LPGROUPS selects the dragonfly-dally method
    dragonfly_dally_method.mn_register(nullptr);
    dragonfly_dally_router_method.mn_register(nullptr);
    g.registered = true;
}
int* model_net_configure(int* id_count) {
    Compat& g = G();
    if (!g.mapped) die("model_net_configure() before codes_mapping_setup()");
    int ids[4], n = 0;
    if (dfd_model_net_configure(g.eng, ids, 4, &n)) die("model_net_configure");
    g.configured = true;
    char buf[64];
    if (dfd_configuration_get_value(g.eng, "PARAMS", "packet_size", buf, sizeof buf) > 0)
        dragonfly_dally_method.packet_size = dragonfly_dally_router_method.packet_size = strtoull(buf, nullptr, 10);
    int* out = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) out[i] = ids[i];
    if (id_count) *id_count = n;
    return out;
}
void model_net_enable_sampling(tw_stime, tw_stime) {}  // sampling outputs are SURVEY 8f-3 (not built)
model_net_event_return model_net_event(int, char const*, tw_lpid, uint64_t, tw_stime, int, void const*, int, void const*, tw_lp*) {
    device_side("model_net_event");
}
void model_net_event_rc2(tw_lp*, model_net_event_return const*) { device_side("model_net_event_rc2"); }
void model_net_report_stats(int net_id) {
    Compat& g = G();
    if (net_id != DRAGONFLY_DALLY) return;
    static char report[16384];
    if (!g.ran || dfd_model_net_report_stats(g.eng, report, sizeof report)) die("model_net_report_stats");
    fputs(report, stdout);
}

// ---- codes_mapping (codes/codes_mapping.h) -------------------------------------------------------------------
tw_peid codes_mapping(tw_lpid) { return 0; }
void codes_mapping_setup(void) {
    if (dfd_codes_mapping_setup(engine())) die("codes_mapping_setup");
    G().mapped = true;
}
int codes_mapping_get_lp_count(const char* group_name, int ignore_repetitions, const char* lp_type_name, const char*, int) {
    (void)group_name;  // one LP group on this path (MODELNET_GRP)
    if (ignore_repetitions) return (int)count_per_rep(engine(), lp_type_name);
    long long n = dfd_codes_mapping_get_lp_count(engine(), lp_type_name);
    return n < 0 ? 0 : (int)n;
}
int codes_mapping_get_lp_relative_id(tw_lpid gid, int, int) { return (int)dfd_codes_mapping_get_lp_relative_id(engine(), (long long)gid); }
tw_lpid codes_mapping_get_lpid_from_relative(int relative_id, const char*, const char* lp_type_name, const char*, int) {
    return (tw_lpid)dfd_codes_mapping_get_lpid_from_relative(engine(), relative_id, lp_type_name);
}
void codes_mapping_get_lp_info(tw_lpid gid, char* group_name, int* group_index, char* lp_type_name, int* lp_type_index, char* annotation,
                               int* rep_id, int* offset) {
    // one LP group with three LP types per repetition: invert get_lpid_from_relative by search over the types
    static const char* names[3] = {"nw-lp", "modelnet_dragonfly_dally", "modelnet_dragonfly_dally_router"};
    dfd_engine* e = engine();
    long long per_rep[3], lps_per_rep = 0;
    for (int i = 0; i < 3; i++) {
        per_rep[i] = count_per_rep(e, names[i]);
        lps_per_rep += per_rep[i];
    }
    long long rep = lps_per_rep ? (long long)gid / lps_per_rep : 0, off = lps_per_rep ? (long long)gid % lps_per_rep : 0;
    int ti = 0;
    while (ti < 2 && off >= per_rep[ti]) off -= per_rep[ti++];
    if (group_name) strcpy(group_name, LPGROUP);
    if (group_index) *group_index = 0;
    if (lp_type_name) strcpy(lp_type_name, names[ti]);
    if (lp_type_index) *lp_type_index = ti;
    if (annotation) annotation[0] = '\0';
    if (rep_id) *rep_id = (int)rep;
    if (offset) *offset = (int)off;
}

// ---- lp-io (codes/lp-io.h) -----------------------------------------------------------------------------------
int lp_io_prepare(char* directory, int flags, lp_io_handle* handle, MPI_Comm) {
    // lp-io.c:150-190: optional "-<pid>-<time>" suffix; the directory must not exist yet
    std::string dir(directory);
    if (flags & LP_IO_UNIQ_SUFFIX) dir += "-" + std::to_string((long)getpid()) + "-" + std::to_string((long)time(nullptr));
    if (mkdir(dir.c_str(), 0775) != 0) {
        fprintf(stderr, "mkdir(\"%s\") failed: %s\n", dir.c_str(), strerror(errno));
        return -1;
    }
    *handle = strdup(dir.c_str());
    return 0;
}
int lp_io_write(tw_lpid, char*, int, void*) { device_side("lp_io_write"); }
int lp_io_flush(lp_io_handle handle, MPI_Comm) {
    Compat& g = G();
    if (!g.ran) die("lp_io_flush() before tw_run()");
    if (dfd_lp_io_flush(g.eng, handle)) die("lp_io_flush");
    free(handle);
    return 0;
}

// ---- pieces of libcodes the driver touches that have nothing to do on this path ---------------------------------
struct codes_jobmap_ctx {
    int num_ranks;
};
struct codes_jobmap_ctx* codes_jobmap_configure(int /*enum codes_jobmap_type*/, void const* params) {
    codes_jobmap_ctx* c = (codes_jobmap_ctx*)malloc(sizeof *c);
    c->num_ranks = params ? *(const int*)params : 0;  // identity map: struct codes_jobmap_params_identity { int num_ranks; }
    return c;
}
// YAML workload sections (codes/codes-workload-config.h): a .conf carries none, so these leave the command-line values alone
struct codes_workload_traffic_name;
void codes_workload_config_check_unsupported_jobs(const char*) {}
void codes_workload_config_apply_int(const char*, int*, int) {}
void codes_workload_config_apply_double(const char*, double*, double) {}
void codes_workload_config_apply_traffic(int*, int, const struct codes_workload_traffic_name*) {}

}  // extern "C"
