/* ross_shim.c -- sequential stand-in for the ROSS engine (TEST INFRASTRUCTURE, see ross.h).
 * Scheduler: binary heap ordered by (recv_ts, sender gid, sender sequence); tw_run = init all LPs in gid
 * order, pop/execute until empty or recv_ts >= g_tw_ts_end, final handlers in gid order.
 * RNG: L'Ecuyer-Andres CLCG4, stream seeds by jump-ahead (same construction as oracle/dfdally_oracle.cpp). */
#include <ross.h>
#include <time.h>

tw_synch g_tw_synchronization_protocol = SEQUENTIAL;
tw_peid g_tw_mynode = 0;
tw_stime g_tw_ts_end = 100000.0;
tw_stime g_tw_lookahead = 0.005;
unsigned int g_tw_nRNG_per_lp = 1;
tw_lp** g_tw_lp = NULL;
tw_lpid g_tw_nlp = 0;
tw_pe* g_tw_pe = NULL;
tw_kp** g_tw_kp = NULL;
tw_kpid g_tw_nkp = 16;
size_t g_tw_msg_sz = 0;
tw_lp_map g_tw_mapping = LINEAR;
unsigned int g_tw_events_per_pe = 0;
map_local_f g_tw_custom_initial_mapping = NULL;
map_custom_f g_tw_custom_lp_global_to_local_map = NULL;
int g_st_ev_trace = 0, g_st_model_stats = 0, g_st_use_analysis_lps = 0, g_st_ross_rank = 1;
MPI_Comm MPI_COMM_ROSS = 0;
unsigned long long g_shim_events_processed = 0;
double g_shim_run_seconds = 0;

void tw_error(const char* file, int line, const char* fmt, ...) {
    va_list ap;
    fprintf(stderr, "error: %s:%d: ", file, line);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
    exit(1);
}
void tw_warning(const char* file, int line, const char* fmt, ...) {
    va_list ap;
    fprintf(stderr, "warning: %s:%d: ", file, line);
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}
int tw_output(tw_lp* lp, const char* fmt, ...) {
    (void)lp;
    va_list ap;
    va_start(ap, fmt);
    int r = vprintf(fmt, ap);
    va_end(ap);
    return r;
}
void tw_fprint_binary_array(FILE* f, char const* prefix, void const* array, size_t size) {
    (void)array;
    fprintf(f, "%s<%zu bytes>\n", prefix, size);
}
void crv_add_custom_state_checkpoint(crv_checkpointer* c) { (void)c; }

/* ---- options ---------------------------------------------------------------------------------- */
static const tw_optdef* opt_groups[16];
static int n_opt_groups = 0;
void tw_opt_add(const tw_optdef* options) {
    if (n_opt_groups < 16) opt_groups[n_opt_groups++] = options;
}
static int apply_opt(const char* name, const char* val) {
    for (int g = 0; g < n_opt_groups; g++)
        for (const tw_optdef* o = opt_groups[g]; o->type; o++) {
            if (o->type == TWOPTTYPE_GROUP || strcmp(o->name, name)) continue;
            switch (o->type) {
            case TWOPTTYPE_ULONG: *(unsigned long*)o->value = strtoul(val, NULL, 10); break;
            case TWOPTTYPE_ULONGLONG: *(unsigned long long*)o->value = strtoull(val, NULL, 10); break;
            case TWOPTTYPE_UINT: *(unsigned int*)o->value = (unsigned int)strtoul(val, NULL, 10); break;
            case TWOPTTYPE_STIME:
            case TWOPTTYPE_DOUBLE: *(double*)o->value = strtod(val, NULL); break;
            case TWOPTTYPE_CHAR: strcpy((char*)o->value, val); break;
            case TWOPTTYPE_FLAG: *(unsigned int*)o->value = 1; break;
            default: break;
            }
            return 1;
        }
    return 0;
}
void tw_init(int* argc, char*** argv) {
    /* consume --name=value options; leave argv = {prog, "--", <remaining>} as ROSS does */
    char** av = *argv;
    int n = *argc, out = 1, i = 1;
    for (; i < n; i++) {
        const char* a = av[i];
        if (!strcmp(a, "--")) break;
        if (a[0] == '-' && a[1] == '-') {
            char name[128];
            const char* eq = strchr(a, '=');
            size_t len = eq ? (size_t)(eq - a - 2) : strlen(a) - 2;
            if (len >= sizeof name) len = sizeof name - 1;
            memcpy(name, a + 2, len);
            name[len] = 0;
            const char* val = eq ? eq + 1 : "1";
            if (!strcmp(name, "sync") || !strcmp(name, "synch")) {
                if (atoi(val) != 1) fprintf(stderr, "shim: only --sync=1 (sequential) exists; ignoring --sync=%s\n", val);
            } else if (!strcmp(name, "end")) {
                g_tw_ts_end = strtod(val, NULL);
            } else if (!apply_opt(name, val)) {
                fprintf(stderr, "shim: unknown option %s\n", a);
                exit(2);
            }
        } else
            break;
    }
    for (; i < n; i++) av[out++] = av[i];
    *argc = out;
    static tw_pe pe;
    memset(&pe, 0, sizeof pe);
    g_tw_pe = &pe;
}
tw_peid tw_nnodes(void) { return 1; }

/* ---- LPs ---------------------------------------------------------------------------------------- */
void tw_define_lps(tw_lpid nlp, size_t msg_sz) {
    g_tw_nlp = nlp;
    g_tw_msg_sz = msg_sz;
    g_tw_lp = (tw_lp**)calloc(nlp, sizeof(tw_lp*));
    g_tw_kp = (tw_kp**)calloc(g_tw_nkp, sizeof(tw_kp*));
    if (g_tw_mapping == CUSTOM && g_tw_custom_initial_mapping)
        g_tw_custom_initial_mapping();
    else
        tw_error(TW_LOC, "shim: only the CUSTOM mapping of codes_mapping is supported");
    for (tw_lpid i = 0; i < nlp; i++) {
        tw_lp* lp = g_tw_lp[i];
        lp->rng = (tw_rng_stream*)calloc(g_tw_nRNG_per_lp, sizeof(tw_rng_stream));
        for (unsigned int s = 0; s < g_tw_nRNG_per_lp; s++) tw_rand_initial_seed(&lp->rng[s], lp->gid * g_tw_nRNG_per_lp + s, NULL);
    }
}
void tw_kp_onpe(tw_kpid id, tw_pe* pe) {
    g_tw_kp[id] = (tw_kp*)calloc(1, sizeof(tw_kp));
    g_tw_kp[id]->id = id;
    g_tw_kp[id]->pe = pe;
}
void tw_lp_onpe(tw_lpid index, tw_pe* pe, tw_lpid gid) {
    tw_lp* lp = (tw_lp*)calloc(1, sizeof(tw_lp));
    lp->id = index;
    lp->gid = gid;
    lp->pe = pe;
    g_tw_lp[index] = lp;
}
void tw_lp_onkp(tw_lp* lp, tw_kp* kp) { lp->kp = kp; }
void tw_lp_settype(tw_lpid lp, tw_lptype* type) { g_tw_lp[lp]->type = type; }
void tw_lp_suspend(tw_lp* lp, int a, int b) {
    (void)a;
    (void)b;
    tw_error(TW_LOC, "tw_lp_suspend called for LP %llu", (unsigned long long)lp->gid);
}
void st_model_settype(tw_lpid lp, st_model_types* t) { g_tw_lp[lp]->model_types = t; }
const st_model_types* st_model_type_lookup(const char* name);

/* ---- events ------------------------------------------------------------------------------------- */
static tw_event** heap = NULL;
static size_t heap_n = 0, heap_cap = 0;
static tw_event* free_events = NULL;
static tw_event abort_event_storage;

static int ev_less(const tw_event* a, const tw_event* b) {
    if (a->recv_ts != b->recv_ts) return a->recv_ts < b->recv_ts;
    if (a->shim_src != b->shim_src) return a->shim_src < b->shim_src;
    return a->shim_seq < b->shim_seq;
}
static void heap_push(tw_event* e) {
    if (heap_n == heap_cap) {
        heap_cap = heap_cap ? heap_cap * 2 : 1024;
        heap = (tw_event**)realloc(heap, heap_cap * sizeof *heap);
    }
    size_t i = heap_n++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!ev_less(e, heap[p])) break;
        heap[i] = heap[p];
        i = p;
    }
    heap[i] = e;
}
static tw_event* heap_pop(void) {
    tw_event* top = heap[0];
    tw_event* last = heap[--heap_n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m;
        if (l >= heap_n) break;
        m = (r < heap_n && ev_less(heap[r], heap[l])) ? r : l;
        if (!ev_less(heap[m], last)) break;
        heap[i] = heap[m];
        i = m;
    }
    if (heap_n) heap[i] = last;
    return top;
}
tw_event* tw_event_new(tw_lpid dest_gid, tw_stime offset_ts, tw_lp* sender) {
    tw_event* e = free_events;
    if (e)
        free_events = e->shim_next;
    else
        e = (tw_event*)malloc(sizeof(tw_event) + g_tw_msg_sz);
    memset(e, 0, sizeof(tw_event));
    e->dest_lpid = dest_gid;
    e->src_lp = sender;
    e->recv_ts = tw_now(sender) + offset_ts;
    e->shim_src = (uint32_t)sender->gid;
    e->shim_seq = sender->shim_seq++;
    return e;
}
tw_event* tw_event_new_user_prio(tw_lpid dest_gid, tw_stime offset_ts, tw_lp* sender, tw_stime prio) {
    (void)prio;
    return tw_event_new(dest_gid, offset_ts, sender);
}
void* tw_event_data(tw_event* event) { return (void*)(event + 1); }
void tw_event_send(tw_event* e) {
    if (e->recv_ts >= g_tw_ts_end) { /* events at/after the end barrier are never executed */
        e->shim_next = free_events;
        free_events = e;
        return;
    }
    heap_push(e);
}
tw_event_sig tw_now_sig(tw_lp const* lp) {
    tw_event_sig s;
    memset(&s, 0, sizeof s);
    s.recv_ts = lp->shim_now;
    return s;
}
int tw_event_sig_compare_ptr(tw_event_sig const* a, tw_event_sig const* b) { return a->recv_ts < b->recv_ts ? -1 : a->recv_ts > b->recv_ts; }

void tw_run(void) {
    (void)abort_event_storage;
    for (tw_lpid i = 0; i < g_tw_nlp; i++) {
        tw_lp* lp = g_tw_lp[i];
        lp->cur_state = calloc(1, lp->type->state_sz);
        lp->shim_now = 0;
    }
    for (tw_lpid i = 0; i < g_tw_nlp; i++) {
        tw_lp* lp = g_tw_lp[i];
        if (lp->type->init) lp->type->init(lp->cur_state, lp);
    }
    for (tw_lpid i = 0; i < g_tw_nlp; i++) {
        tw_lp* lp = g_tw_lp[i];
        if (lp->type->pre_run) lp->type->pre_run(lp->cur_state, lp);
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    tw_bf bf;
    while (heap_n) {
        tw_event* e = heap_pop();
        tw_lp* lp = g_tw_custom_lp_global_to_local_map ? g_tw_custom_lp_global_to_local_map(e->dest_lpid) : g_tw_lp[e->dest_lpid];
        lp->shim_now = e->recv_ts;
        memset(&bf, 0, sizeof bf);
        g_shim_events_processed++;
        lp->type->event(lp->cur_state, &bf, tw_event_data(e), lp);
        if (lp->type->commit) lp->type->commit(lp->cur_state, &bf, tw_event_data(e), lp);
        e->shim_next = free_events;
        free_events = e;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    g_shim_run_seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    for (tw_lpid i = 0; i < g_tw_nlp; i++) {
        tw_lp* lp = g_tw_lp[i];
        if (lp->type->final) lp->type->final(lp->cur_state, lp);
    }
    printf("\tNet Events Processed                               %llu\n", g_shim_events_processed);
    fprintf(stderr, "shim: %llu events in %.3f s (%.3f Mev/s)\n", g_shim_events_processed, g_shim_run_seconds,
            g_shim_events_processed / g_shim_run_seconds / 1e6);
}
void tw_end(void) {}

/* ---- CLCG4 -------------------------------------------------------------------------------------- */
static const long long RM[4] = {2147483647LL, 2147483543LL, 2147483423LL, 2147483323LL};
static const long long RA[4] = {45991LL, 207707LL, 138556LL, 49689LL};
static const long long RS[4] = {11111111LL, 22222222LL, 33333333LL, 44444444LL};
static long long mulmod(long long a, long long b, long long m) { return (long long)(((__int128)a * b) % m); }
void tw_rand_initial_seed(tw_rng_stream* g, tw_lpid id, void* unused) {
    (void)unused;
    for (int j = 0; j < 4; j++) {
        long long avw = RA[j], acc = 1;
        for (int i = 0; i < 31 + 41; i++) avw = mulmod(avw, avw, RM[j]);
        for (tw_lpid e = id; e; e >>= 1) {
            if (e & 1) acc = mulmod(acc, avw, RM[j]);
            avw = mulmod(avw, avw, RM[j]);
        }
        g->Ig[j] = g->Lg[j] = g->Cg[j] = mulmod(acc, RS[j], RM[j]);
    }
    g->count = 0;
}
double tw_rand_unif(tw_rng_stream* g) {
    long k, s;
    double u = 0.0;
    g->count++;
    s = g->Cg[0];
    k = s / 46693;
    s = 45991 * (s - k * 46693) - k * 25884;
    if (s < 0) s = s + 2147483647;
    g->Cg[0] = s;
    u = u + 4.65661287524579692e-10 * s;
    s = g->Cg[1];
    k = s / 10339;
    s = 207707 * (s - k * 10339) - k * 870;
    if (s < 0) s = s + 2147483543;
    g->Cg[1] = s;
    u = u - 4.65661310075985993e-10 * s;
    if (u < 0) u = u + 1.0;
    s = g->Cg[2];
    k = s / 15499;
    s = 138556 * (s - k * 15499) - k * 3979;
    if (s < 0) s = s + 2147483423;
    g->Cg[2] = s;
    u = u + 4.65661336096842131e-10 * s;
    if (u >= 1.0) u = u - 1.0;
    s = g->Cg[3];
    k = s / 43218;
    s = 49689 * (s - k * 43218) - k * 24121;
    if (s < 0) s = s + 2147483323;
    g->Cg[3] = s;
    u = u - 4.65661357780891134e-10 * s;
    if (u < 0) u = u + 1.0;
    return u;
}
long tw_rand_integer(tw_rng_stream* g, long low, long high) {
    if (high < low) return 0;
    return low + (long)(tw_rand_unif(g) * (high + 1 - low));
}
double tw_rand_exponential(tw_rng_stream* g, double lambda) { return -lambda * log(tw_rand_unif(g)); }
double tw_rand_reverse_unif(tw_rng_stream* g) {
    (void)g;
    tw_error(TW_LOC, "shim: reverse computation is not available (sequential only)");
}
