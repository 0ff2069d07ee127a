// oracle/dfdally_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or execute this file.  The product (codes_b200/) never calls it.
//
// What it is: a single-threaded CPU restatement of the CODES dragonfly-dally hot path (svr "nw-lp"
// traffic generator + model-net base LP + dragonfly-dally terminal and router LPs) driven by a
// sequential discrete-event loop, producing the same lp-io text records and summary lines as the
// reference.  Every function cites the reference file:line it follows (paths relative to
// /root/reference).
//
// PARITY STATUS vs ROSS: *unpinned*.  ROSS (the PDES engine: event queue, CLCG4 RNG seeding, tie
// order) is an external dependency of the reference pinned at 9b6ccb18f9b9 and is NOT present under
// /root/reference nor installed here.  The model arithmetic follows the reference sources line by
// line; the engine semantics this file supplies are stated here and in DESIGN.md:
//   * RNG: L'Ecuyer-Andres CLCG4 (published algorithm), 3 streams per LP, stream id = gid*3+i,
//     seed(stream) = default seed jumped ahead stream * 2^(31+41) steps; tw_rand_integer(lo,hi) =
//     lo + (long)(unif * (hi+1-lo)).
//   * tw_now / tw_event_new: recv_ts = now(sender) + offset (IEEE double, no FMA).
//   * Event order: events are executed by repeatedly popping the minimum of
//     (recv_ts, sender gid, sender's event-creation sequence number) over all pending events.
//   * g_tw_lookahead = 0.005 (ROSS default) enters codes_local_latency.
// The restatement IS pinned against the reference's own handler code whenever oracle/_ref (the
// reference sources compiled unchanged against oracle/ross_shim) is built: tests diff both.

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <queue>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <time.h>

namespace {

[[noreturn]] void die(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw std::runtime_error(buf);
}

// ------------------------------------------------------------------------------------------------
// .conf reader (stands in for src/modelconfig/configlex.l + configparser.y; same grammar subset:
// SECTION { key="v"; key=("a","b"); SUB { ... } } with '#' comments)
// ------------------------------------------------------------------------------------------------
struct ConfNode {
    std::map<std::string, std::vector<std::string>> kv;
    std::vector<std::string> key_order;
    std::map<std::string, ConfNode> sub;
    std::vector<std::string> sub_order;
};

struct ConfLexer {
    std::string s;
    size_t i = 0;
    explicit ConfLexer(std::string t) : s(std::move(t)) {}
    void skip() {
        for (;;) {
            while (i < s.size() && isspace((unsigned char)s[i])) i++;
            if (i < s.size() && s[i] == '#') {
                while (i < s.size() && s[i] != '\n') i++;
            } else
                break;
        }
    }
    // returns token; kind: 0 eof, 1 punct, 2 ident, 3 string
    int next(std::string& tok) {
        skip();
        tok.clear();
        if (i >= s.size()) return 0;
        char c = s[i];
        if (strchr("{}=;(),", c)) {
            tok = c;
            i++;
            return 1;
        }
        if (c == '"') {
            i++;
            while (i < s.size() && s[i] != '"') tok += s[i++];
            i++;
            return 3;
        }
        while (i < s.size() && !isspace((unsigned char)s[i]) && !strchr("{}=;(),\"#", s[i]))
            tok += s[i++];
        return 2;
    }
};

void parse_block(ConfLexer& lx, ConfNode& node, bool top) {
    std::string tok;
    for (;;) {
        int k = lx.next(tok);
        if (k == 0) {
            if (!top) die("config: unexpected end of file");
            return;
        }
        if (k == 1 && tok == "}") {
            if (top) die("config: unexpected '}'");
            return;
        }
        if (k != 2 && k != 3) die("config: unexpected token '%s'", tok.c_str());
        std::string name = tok;
        k = lx.next(tok);
        if (k == 1 && tok == "{") {
            node.sub_order.push_back(name);
            parse_block(lx, node.sub[name], false);
        } else if (k == 1 && tok == "=") {
            std::vector<std::string> vals;
            k = lx.next(tok);
            if (k == 1 && tok == "(") {
                for (;;) {
                    k = lx.next(tok);
                    if (k == 1 && tok == ")") break;
                    if (k == 1 && tok == ",") continue;
                    if (k == 0) die("config: unterminated list for %s", name.c_str());
                    vals.push_back(tok);
                }
            } else {
                vals.push_back(tok);
            }
            k = lx.next(tok);
            if (!(k == 1 && tok == ";")) die("config: expected ';' after %s", name.c_str());
            if (!node.kv.count(name)) node.key_order.push_back(name);
            node.kv[name] = vals;
        } else
            die("config: expected '{' or '=' after %s", name.c_str());
    }
}

ConfNode load_conf(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open config file %s", path.c_str());
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    ConfLexer lx(text);
    ConfNode root;
    parse_block(lx, root, true);
    return root;
}

// configuration_get_value_* (codes/configuration.h:142-198): non-zero return when missing.
struct ParamReader {
    const ConfNode* p;
    bool has(const char* k) const { return p->kv.count(k) && !p->kv.at(k).empty(); }
    int get_int(const char* k, int* out) const {
        if (!has(k)) return 1;
        *out = (int)strtol(p->kv.at(k)[0].c_str(), nullptr, 10);
        return 0;
    }
    int get_double(const char* k, double* out) const {
        if (!has(k)) return 1;
        *out = strtod(p->kv.at(k)[0].c_str(), nullptr);
        return 0;
    }
    std::string get_str(const char* k) const { return has(k) ? p->kv.at(k)[0] : std::string(); }
};

// ------------------------------------------------------------------------------------------------
// time: dragonfly-dally.cxx:1588-1599 (identical copy synthetic:205-216) -- three separate ops.
// ------------------------------------------------------------------------------------------------
double bytes_to_ns(uint64_t bytes, double GB_p_s) {
    double time;
    time = ((double)bytes) / (1024.0 * 1024.0 * 1024.0);
    time = time / GB_p_s;
    time = time * 1000.0 * 1000.0 * 1000.0;
    return time;
}
// synthetic:224-228
double ns_to_GBps(double ns, uint64_t bytes) {
    return (bytes / (1024.0 * 1024.0 * 1024.0)) / (ns / (1000.0 * 1000.0 * 1000.0));
}
double maxd(double a, double b) { return a < b ? b : a; }  // dally:180-182

// ------------------------------------------------------------------------------------------------
// CLCG4 (L'Ecuyer & Andres 1997; ROSS rand-clcg4.c -- external, restated from the published
// algorithm; see header for the seeding assumption)
// ------------------------------------------------------------------------------------------------
const int64_t RNG_M[4] = {2147483647LL, 2147483543LL, 2147483423LL, 2147483323LL};
const int64_t RNG_A[4] = {45991LL, 207707LL, 138556LL, 49689LL};
const int64_t RNG_SEED0[4] = {11111111LL, 22222222LL, 33333333LL, 44444444LL};
int64_t mulmod(int64_t a, int64_t b, int64_t m) { return (int64_t)(((__int128)a * b) % m); }
int64_t powmod(int64_t a, uint64_t e, int64_t m) {
    int64_t r = 1;
    while (e) {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}
struct Rng {
    int64_t s[4];
    uint64_t draws = 0;
    void seed_stream(uint64_t stream) {
        for (int j = 0; j < 4; j++) {
            int64_t avw = RNG_A[j];
            for (int i = 0; i < 31 + 41; i++) avw = mulmod(avw, avw, RNG_M[j]);  // a^(2^(v+w))
            s[j] = mulmod(powmod(avw, stream, RNG_M[j]), RNG_SEED0[j], RNG_M[j]);
        }
        draws = 0;
    }
    double unif() {
        long k, v;
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
        draws++;
        return u;
    }
    long integer(long low, long high) {
        if (high < low) return 0;
        return low + (long)(unif() * (high + 1 - low));
    }
};

// ------------------------------------------------------------------------------------------------
// parameters: dragonfly_read_config dally:2151-2757, base_read_config model-net-lp.c:261-375,
// model_net_configure model-net.c:90-174, codes_mapping_setup codes_mapping.c:500-574
// ------------------------------------------------------------------------------------------------
enum Routing { MINIMAL = 1, NON_MINIMAL = 2, PROG_ADAPTIVE = 4 };  // values only matter internally;
// path_type uses MINIMAL/NON_MINIMAL = enum values of the reference's ROUTING_ALGO (dally:446-455)
enum Scoring { ALPHA = 0, DELTA = 3 };
enum LastHop { GLOBAL = 1, LOCAL, TERMINAL };  // dally:516-520 (values opaque)
enum ConnType { CONN_LOCAL = 1, CONN_GLOBAL, CONN_TERMINAL };

struct Params {
    int num_routers = 0, num_groups = 0, num_cn = 0, num_global_channels = 0;
    int local_vc_size = 0, global_vc_size = 0, cn_vc_size = 0, chunk_size = 0, packet_size = 0;
    double local_bandwidth = 0, global_bandwidth = 0, cn_bandwidth = 0, router_delay = 0;
    int num_vcs = 4, adaptive_threshold = 0, global_k_picks = 2, num_parallel_switch_conns = 1;
    int routing = MINIMAL, scoring = DELTA;
    int intra_grp_radix = 0, global_radix = 0, cn_radix = 0, radix = 0;
    int total_routers = 0, total_terminals = 0;
    double cn_delay = 0, local_delay = 0, global_delay = 0;
    int credit_size = 8;
    double local_credit_delay = 0, global_credit_delay = 0, cn_credit_delay = 0;
    // base LP
    uint64_t mn_packet_size = 0;
    double nic_seq_delay = 10;
    // model-net.c
    double mn_cn_bandwidth = 20.0, codes_cn_delay = 0;  // model-net.c:60,144-157
    int node_eager_limit = 16000, codes_noop_bypass = 0;
    int message_size = 256;
    double end_time = 0;
    double svr_link_bandwidth = 0;  // synthetic:236-240 (re-reads PARAMS cn_bandwidth; 4.7 when unset/0)
    std::string intra_file, inter_file;
    // LPGROUPS
    int repetitions = 0;
    int n_svr_per_rep = 0, n_term_per_rep = 0, n_rtr_per_rep = 0;
    int off_svr = 0, off_term = 0, off_rtr = 0, lps_per_rep = 0;
    std::vector<double> snapshot_times;  // PARAMS router_buffer_snapshots dally:2783-2797
};

std::string resolve_path(const std::string& p, const std::string& conf_dir) {
    struct stat st;
    if (stat(p.c_str(), &st) == 0) return p;
    if (!p.empty() && p[0] != '/') {
        std::string q = conf_dir + "/" + p;
        if (stat(q.c_str(), &st) == 0) return q;
    }
    return p;
}

Params read_params(const ConfNode& root, const std::string& conf_dir, FILE* log) {
    Params p;
    if (!root.sub.count("PARAMS")) die("config has no PARAMS section");
    if (!root.sub.count("LPGROUPS")) die("config has no LPGROUPS section");
    ParamReader r{&root.sub.at("PARAMS")};

    // ---- LPGROUPS (codes_mapping.c: one group, repetitions x [types in file order])
    const ConfNode& lg = root.sub.at("LPGROUPS");
    if (lg.sub_order.size() != 1) die("exactly one LPGROUPS group is supported");
    const ConfNode& grp = lg.sub.at(lg.sub_order[0]);
    int off = 0;
    bool seen_s = false, seen_t = false, seen_r = false;
    for (const std::string& k : grp.key_order) {
        int v = (int)strtol(grp.kv.at(k)[0].c_str(), nullptr, 10);
        if (k == "repetitions") {
            p.repetitions = v;
            continue;
        }
        if (k == "nw-lp") {
            p.n_svr_per_rep = v;
            p.off_svr = off;
            seen_s = true;
        } else if (k == "modelnet_dragonfly_dally") {
            p.n_term_per_rep = v;
            p.off_term = off;
            seen_t = true;
        } else if (k == "modelnet_dragonfly_dally_router") {
            p.n_rtr_per_rep = v;
            p.off_rtr = off;
            seen_r = true;
        } else
            die("unsupported LP type '%s' in LPGROUPS", k.c_str());
        off += v;
    }
    p.lps_per_rep = off;
    if (!seen_s || !seen_t || !seen_r)
        die("LPGROUPS must hold nw-lp, modelnet_dragonfly_dally and modelnet_dragonfly_dally_router");

    // ---- dragonfly_read_config dally:2165-2234
    if (r.get_int("local_vc_size", &p.local_vc_size)) {
        p.local_vc_size = 1024;
        fprintf(log, "Buffer size of local channels not specified, setting to %d\n", p.local_vc_size);
    }
    if (r.get_int("global_vc_size", &p.global_vc_size)) {
        p.global_vc_size = 2048;
        fprintf(log, "Buffer size of global channels not specified, setting to %d\n", p.global_vc_size);
    }
    if (r.get_int("cn_vc_size", &p.cn_vc_size)) {
        p.cn_vc_size = 1024;
        fprintf(log, "Buffer size of compute node channels not specified, setting to %d\n", p.cn_vc_size);
    }
    if (r.get_int("chunk_size", &p.chunk_size)) {
        p.chunk_size = 512;
        fprintf(log, "Chunk size for packets is specified, setting to %d\n", p.chunk_size);
    }
    if (r.get_int("packet_size", &p.packet_size)) {
        p.chunk_size = 512;  // sic, dally:2197-2203
        fprintf(log, "Packet size not specificied, it is assumed to be %d\n", p.packet_size);
    }
    if (r.get_double("local_bandwidth", &p.local_bandwidth)) p.local_bandwidth = 5.25;
    if (r.get_double("global_bandwidth", &p.global_bandwidth)) p.global_bandwidth = 4.7;
    if (r.get_double("cn_bandwidth", &p.cn_bandwidth)) p.cn_bandwidth = 5.25;
    if (r.get_double("router_delay", &p.router_delay)) p.router_delay = 100;
    int nq = 1;
    if (r.get_int("num_qos_levels", &nq)) nq = 1;
    if (nq != 1) die("num_qos_levels > 1 is out of scope (SURVEY 8f-1)");
    if (r.get_int("adaptive_threshold", &p.adaptive_threshold)) p.adaptive_threshold = 0;
    std::string rt = r.get_str("routing");  // dally:2283-2305
    if (rt == "minimal")
        p.routing = MINIMAL;
    else if (rt == "nonminimal" || rt == "non-minimal")
        p.routing = NON_MINIMAL;
    else if (rt == "prog-adaptive")
        p.routing = PROG_ADAPTIVE;
    else if (rt == "adaptive")
        die("Standard Adaptive Routing not implemented - try Progressive Adaptive instead");  // dally:7190
    else if (rt == "prog-adaptive-legacy" || rt == "smart-prog-adaptive" || rt == "smart-minimal" ||
             rt == "smart-nonminimal")
        die("routing '%s' is out of scope (SURVEY 2 row 1)", rt.c_str());
    else {
        fprintf(log, "No routing protocol specified, setting to minimal routing\n");
        p.routing = MINIMAL;
    }
    int t = 1;
    if (!r.get_int("num_injection_queues", &t) && t != 1) die("num_injection_queues > 1 out of scope");
    t = 1;
    if (!r.get_int("num_rails", &t) && t != 1) die("num_rails > 1 out of scope");
    t = 1;
    if (!r.get_int("num_planes", &t) && t != 1) die("num_planes > 1 out of scope");
    if (r.get_int("global_k_picks", &p.global_k_picks)) p.global_k_picks = 2;
    std::string sc = r.get_str("route_scoring_metric");  // dally:2359-2377
    if (sc == "alpha")
        p.scoring = ALPHA;
    else if (sc == "beta")
        die("Beta scoring not implemented");
    else if (sc == "gamma")
        die("Gamma scoring protocol currently non-functional");
    else
        p.scoring = DELTA;
    p.num_vcs = 4;
    if (r.get_int("num_groups", &p.num_groups)) die("\nnum_groups not specified, Aborting\n");
    if (r.get_int("num_routers", &p.num_routers)) die("\nnum_routers not specified, Aborting\n");
    if (r.get_int("num_cns_per_router", &p.num_cn)) p.num_cn = p.num_routers / 2;
    if (r.get_int("num_global_channels", &p.num_global_channels)) p.num_global_channels = 10;
    if (r.get_int("num_parallel_switch_conns", &p.num_parallel_switch_conns))
        p.num_parallel_switch_conns = 1;
    p.intra_grp_radix = (p.num_routers - 1) * p.num_parallel_switch_conns;  // dally:2447-2456
    p.cn_radix = p.num_cn;
    p.global_radix = p.num_global_channels * p.num_parallel_switch_conns;
    p.radix = p.intra_grp_radix + p.global_radix + p.cn_radix;
    p.total_routers = p.num_groups * p.num_routers;
    p.total_terminals = p.total_routers * p.num_cn;
    p.intra_file = resolve_path(r.get_str("intra-group-connections"), conf_dir);
    p.inter_file = resolve_path(r.get_str("inter-group-connections"), conf_dir);
    if (r.get_str("intra-group-connections").empty())
        die("Intra group connections file not specified. Aborting");
    if (r.get_str("inter-group-connections").empty())
        die("Inter group connections file not specified. Aborting");
    if (r.has("link-failure-file")) die("link failures are out of scope");
    int cc = 0;
    if (!r.get_int("congestion_control_enabled", &cc) && cc) die("congestion control out of scope");

    // delays dally:2639-2660
    if (r.get_double("cn_delay", &p.cn_delay)) p.cn_delay = bytes_to_ns(p.chunk_size, p.cn_bandwidth * 1);
    if (r.get_double("local_delay", &p.local_delay)) p.local_delay = bytes_to_ns(p.chunk_size, p.local_bandwidth);
    if (r.get_double("global_delay", &p.global_delay)) p.global_delay = bytes_to_ns(p.chunk_size, p.global_bandwidth);
    // credit delays dally:2663-2756
    if (r.get_int("credit_size", &p.credit_size)) p.credit_size = 8;
    double general = 0;
    int cu = r.get_double("credit_delay", &general);
    int lu = r.get_double("local_credit_delay", &p.local_credit_delay);
    int gu = r.get_double("global_credit_delay", &p.global_credit_delay);
    int nu = r.get_double("cn_credit_delay", &p.cn_credit_delay);
    int autoflag = 0;
    if (r.get_int("auto_credit_delay", &autoflag)) autoflag = 0;
    if (!cu && !(lu || gu || nu))
        die("\nCannot set both a general credit delay and specific (local/global/cn) credit delays. Check configuration file.");
    if ((!cu || !lu || !gu || !nu) && autoflag)
        die("\nCannot set both a credit delay (general or specific) and also enable auto credit delay calculation. Check Configuration file.");
    if (cu && lu && gu && nu && !autoflag) {
        p.local_credit_delay = bytes_to_ns(p.credit_size, p.local_bandwidth);
        p.global_credit_delay = p.local_credit_delay;
        p.cn_credit_delay = p.local_credit_delay;
    } else if (cu) {
        if (lu) p.local_credit_delay = bytes_to_ns(p.credit_size, p.local_bandwidth);
        if (gu) p.global_credit_delay = bytes_to_ns(p.credit_size, p.global_bandwidth);
        if (nu) p.cn_credit_delay = bytes_to_ns(p.credit_size, p.cn_bandwidth);
    } else {
        p.local_credit_delay = p.global_credit_delay = p.cn_credit_delay = general;
    }

    // ---- base LP model-net-lp.c:261-375
    std::string sched = r.get_str("modelnet_scheduler");
    if (!sched.empty() && sched != "fcfs") die("modelnet_scheduler '%s' out of scope (fcfs only)", sched.c_str());
    long ps = 0;
    if (r.has("packet_size")) ps = strtol(r.get_str("packet_size").c_str(), nullptr, 10);
    p.mn_packet_size = ps ? (uint64_t)ps : 512;
    if (r.get_double("nic_seq_delay", &p.nic_seq_delay)) p.nic_seq_delay = 10;
    t = 1;
    if (!r.get_int("node_copy_queues", &t) && t != 1) die("node_copy_queues > 1 out of scope");
    // ---- model-net.c:143-171
    if (r.get_double("cn_bandwidth", &p.mn_cn_bandwidth)) p.mn_cn_bandwidth = 20.0;
    p.codes_cn_delay = 1 / p.mn_cn_bandwidth;
    if (r.get_int("node_eager_limit", &p.node_eager_limit)) p.node_eager_limit = 16000;
    int nb = 0;
    p.codes_noop_bypass = r.get_int("codes_noop_bypass", &nb) ? 0 : nb;
    // ---- codes_mapping.c:519-550
    if (r.get_int("message_size", &p.message_size) || !p.message_size) p.message_size = 256;
    if (r.get_double("end_time", &p.end_time)) p.end_time = 0;
    if (r.get_double("cn_bandwidth", &p.svr_link_bandwidth)) p.svr_link_bandwidth = 0;
    if (!p.svr_link_bandwidth) p.svr_link_bandwidth = 4.7;

    // ---- router VC-occupancy snapshots dally:2783-2797: a list of timestamps, each read with strtod
    if (r.p->kv.count("router_buffer_snapshots"))
        for (const std::string& v : r.p->kv.at("router_buffer_snapshots")) p.snapshot_times.push_back(strtod(v.c_str(), nullptr));

    // consistency dally:4925-4931
    if (p.n_rtr_per_rep != 1 || p.n_term_per_rep != p.num_cn)
        die("LPGROUPS layout must be num_cns_per_router terminals and 1 router per repetition");
    if (p.total_routers != p.repetitions * p.n_rtr_per_rep)
        die("\n Config error: num_routers specified %d total routers computed in the network %d does not match with repetitions * dragonfly_router %d  ",
            p.num_routers, p.total_routers, p.repetitions * p.n_rtr_per_rep);
    if (p.n_svr_per_rep % p.n_term_per_rep != 0 && p.n_svr_per_rep > p.n_term_per_rep)
        die("nw-lp count must be a multiple of the terminal count");
    return p;
}

// ------------------------------------------------------------------------------------------------
// connection tables: dragonfly-network-manager.cxx:91-175 (add_link), :457-500, :672-741
// (add_connection / port numbering), :772-814, :1038-1105, :1134-1178 (ordering of returned vectors)
// ------------------------------------------------------------------------------------------------
struct Conn {
    int port = -1;
    int dest_gid = -1;
    int dest_group = -1;
    int type = 0;
};

struct RouterTopo {
    std::map<int, std::vector<Conn>> local_gid;  // intraGroupConnectionsGID: dest gid -> file order
    std::map<int, std::vector<Conn>> global;     // globalConnections: dest gid -> file order
    std::map<int, std::vector<Conn>> term;       // terminalConnections
    std::map<int, std::vector<Conn>> to_group;   // _connections_to_groups_map
    int used_intra = 0, used_inter = 0, used_term = 0;
};

struct Topo {
    std::vector<RouterTopo> r;
    // _global_group_connection_map[(src group, dest group)] -> src router gid per link, add order
    std::map<std::pair<int, int>, std::vector<int>> group_links;
    static const std::vector<Conn>& empty() {
        static const std::vector<Conn> e;
        return e;
    }
    const std::vector<Conn>& conns_to_gid_local(int rtr, int gid) const {
        auto it = r[rtr].local_gid.find(gid);
        return it == r[rtr].local_gid.end() ? empty() : it->second;
    }
    const std::vector<Conn>& conns_to_group(int rtr, int g) const {
        auto it = r[rtr].to_group.find(g);
        return it == r[rtr].to_group.end() ? empty() : it->second;
    }
    // get_routed_connections_to_group(g, true) network-manager.cxx:772-814
    std::vector<Conn> routed_to_group(int rtr, int my_group, int g) const {
        std::vector<Conn> out;
        auto it = group_links.find({my_group, g});
        if (it == group_links.end()) return out;
        for (int src : it->second) {
            const std::vector<Conn>& lc = conns_to_gid_local(rtr, src);
            out.insert(out.end(), lc.begin(), lc.end());
        }
        return out;
    }
};

Topo build_topo(const Params& p) {
    Topo T;
    T.r.resize(p.total_routers);
    // order of add_connection per router = order of add_link for that src: all intra links (file
    // order) first, then inter links (file order); terminals separately (dally:2478-2580,
    // network-manager.cxx:457-474)
    struct IntraRec {
        int src, dest, type;
    };
    struct InterRec {
        int src, dest;
    };
    FILE* f = fopen(p.intra_file.c_str(), "rb");
    if (!f) die("intra-group file not found ");
    IntraRec ir;
    while (fread(&ir, sizeof ir, 1, f) != 0) {
        for (int g = 0; g < p.num_groups; g++) {
            int src = g * p.num_routers + ir.src;
            int dst = g * p.num_routers + ir.dest;
            RouterTopo& R = T.r[src];
            if (R.used_intra >= p.intra_grp_radix)
                die("Attempting to add too many local connections per router - exceeding configuration value: %d", p.intra_grp_radix);
            Conn c;
            c.port = R.used_intra++;
            c.dest_gid = dst;
            c.dest_group = g;
            c.type = CONN_LOCAL;
            R.local_gid[dst].push_back(c);
        }
    }
    fclose(f);
    f = fopen(p.inter_file.c_str(), "rb");
    if (!f) die("inter-group file not found ");
    InterRec er;
    while (fread(&er, sizeof er, 1, f) != 0) {
        if (er.src < 0 || er.src >= p.total_routers || er.dest < 0 || er.dest >= p.total_routers)
            die("inter-group file: router id out of range");
        RouterTopo& R = T.r[er.src];
        if (R.used_inter >= p.global_radix)
            die("Attempting to add too many global connections per router - exceeding configuration value: %d", p.global_radix);
        Conn c;
        c.port = p.intra_grp_radix + R.used_inter++;
        c.dest_gid = er.dest;
        c.dest_group = er.dest / p.num_routers;
        c.type = CONN_GLOBAL;
        R.global[er.dest].push_back(c);
        T.group_links[{er.src / p.num_routers, c.dest_group}].push_back(er.src);
    }
    fclose(f);
    for (int t = 0; t < p.total_terminals; t++) {  // dally:2503-2515
        int rtr = t / p.num_cn;
        RouterTopo& R = T.r[rtr];
        Conn c;
        c.port = p.intra_grp_radix + p.global_radix + R.used_term++;
        c.dest_gid = t;
        c.dest_group = rtr / p.num_routers;
        c.type = CONN_TERMINAL;
        R.term[t].push_back(c);
    }
    for (RouterTopo& R : T.r)  // solidify_connections network-manager.cxx:1136-1143
        for (auto& kv : R.global)
            for (const Conn& c : kv.second) R.to_group[c.dest_group].push_back(c);
    return T;
}

// ------------------------------------------------------------------------------------------------
// events and LP state
// ------------------------------------------------------------------------------------------------
enum EvType : uint8_t {
    SVR_KICKOFF,
    SVR_REMOTE,
    SVR_LOCAL,
    SVR_ACK,
    MN_NEW_MSG,
    MN_SCHED_NEXT,
    T_GENERATE,
    T_ARRIVE,
    T_SEND,
    T_BUFFER,
    R_SEND,
    R_ARRIVE,
    R_BUFFER,
    R_SNAPSHOT,
    N_EV_TYPES
};

// the subset of terminal_dally_message (codes/net/dragonfly-dally.h:19-139) that is consumed
// downstream, plus the model_net_request fields carried by MN_BASE_NEW_MSG
struct Msg {
    uint64_t packet_ID = 0;
    uint64_t final_dest_gid = 0;  // destination svr gid
    uint64_t sender_lp = 0;       // source svr gid
    uint64_t dest_terminal_lpid = 0;
    uint64_t src_terminal_id = 0;  // gid of source terminal (set in packet_send)
    unsigned dfdally_src_terminal_id = 0, dfdally_dest_terminal_id = 0;
    unsigned origin_router_id = 0;
    unsigned intm_lp_id = 0;  // gid of previous router
    short my_N_hop = 0, my_l_hop = 0, my_g_hop = 0, my_hops_cur_group = 0;
    short saved_channel = 0, saved_vc = 0;
    short last_hop = 0, is_intm_visited = 0;
    int intm_grp_id = -1;
    uint32_t chunk_id = 0, packet_size = 0, message_id = 0, total_size = 0;
    int remote_event_size_bytes = 0, local_event_size_bytes = 0;
    short vc_index = 0;
    int output_chan = 0;
    int path_type = 0;
    int next_stop = 0;
    double travel_start_time = 0, msg_start_time = 0;
    // remote/self event payloads of the synthetic driver (svr_msg synthetic:119-130)
    uint64_t ev_src = 0;
    double ev_start_time = 0;
};

struct Req {  // model_net_request (codes/model-net-lp.h) subset
    uint64_t final_dest_lp = 0, dest_mn_lp = 0, src_lp = 0;
    uint64_t msg_size = 0, packet_size = 0, msg_id = 0;
    int remote_event_size = 0, self_event_size = 0;
    double msg_start_time = 0;
    uint64_t rem = 0;  // mn_sched_qitem.rem
    uint64_t ev_src = 0;
    double ev_start_time = 0;
};

struct Event {
    double ts;
    uint32_t src, seq;
    uint32_t dst;
    uint8_t type;
    uint8_t qreq = 0;
    Msg m;
    Req rq;
};

struct HeapKey {
    double ts;
    uint32_t src, seq;
    uint32_t idx;
};
struct HeapCmp {
    bool operator()(const HeapKey& a, const HeapKey& b) const {  // min-heap
        if (a.ts != b.ts) return a.ts > b.ts;
        if (a.src != b.src) return a.src > b.src;
        return a.seq > b.seq;
    }
};

struct SvrState {  // svr_state synthetic:100-117
    int msg_sent_count = 0, warm_msg_sent_count = 0, msg_recvd_count = 0, local_recvd_count = 0;
    double start_ts = 0, last_send_ts = 0, first_recv_ts = -1, last_recv_ts = 0;
    int svr_id = 0, dest_id = -1;
    long long rperm_data = 0;
    int msg_complete_count = 0;
    double max_server_latency = 0, sum_server_latency = 0;
};

struct MnStats {  // struct mn_stats codes/model-net.h
    bool used = false;
    long send_count = 0, send_bytes = 0;
    double send_time = 0;
    long recv_count = 0, recv_bytes = 0;
    double recv_time = 0;
    long max_event_size = 0;
};

struct RankEntry {  // dfly_qhash_entry dally:466-474
    int num_chunks = 0;
    uint64_t remaining_packets = 0;
    bool has_remote = false;
    int remote_event_size = 0;
    uint64_t ev_src = 0;
    double ev_start_time = 0;
};

struct TermState {
    // model_net_base_state model-net-lp.c:50-74
    uint64_t msg_id = 0;
    double next_available_time = 0;
    double node_copy_next_available_time = 0;  // model_net_base_state.node_copy_next_available_time[0]
    int in_sched_send_loop = 0;
    std::deque<Req> sched_q;  // fcfs queue model-net-sched-impl.c:128-261
    // terminal_state dally:703-820
    int64_t packet_counter = 0;
    int packet_gen = 0, packet_fin = 0, total_gen_size = 0;
    unsigned router_id = 0, terminal_id = 0;
    uint64_t router_lp = 0;
    int vc_occupancy = 0;
    double terminal_available_time = 0;
    std::deque<Msg> terminal_msgs;
    int in_send_loop = 0, issueIdle = 0, terminal_length = 0;
    MnStats stats;  // category "test"
    std::map<std::pair<uint64_t, uint64_t>, RankEntry> rank_tbl;
    uint64_t rank_tbl_pop = 0;
    double total_time = 0;
    uint64_t total_msg_size = 0;
    double total_hops = 0;
    long finished_msgs = 0, finished_chunks = 0, finished_packets = 0;
    double last_buf_full = 0, busy_time = 0;
    uint64_t link_traffic = 0;
    unsigned long total_chunks = 0, stalled_chunks = 0, injected_chunks = 0, ejected_chunks = 0;
    double max_latency = 0, min_latency = INT_MAX;
    std::map<std::pair<uint64_t, unsigned>, uint32_t> remaining_sz_packets;
    double last_in_queue_time = 0;
};

struct RouterState {  // router_state dally:822-888
    unsigned router_id = 0;
    int group_id = 0;
    std::vector<double> next_output_available_time, last_buf_full, busy_time;
    std::vector<unsigned long> stalled_chunks, total_chunks;
    std::vector<std::deque<Msg>> pending_msgs, queued_msgs;  // [port*num_vcs+vc]
    std::vector<int> in_send_loop, queued_count, last_qos_lvl, vc_max_sizes;
    std::vector<int> vc_occupancy;  // [port*num_vcs+vc]
    std::vector<int64_t> link_traffic;
    std::vector<std::vector<int>> snapshot_data;  // [snapshot][port*num_vcs+vc] dally:867, 5084-5091
};

struct RunOpts {  // CLI of model-net-synthetic-dragonfly-all synthetic:173-194
    int traffic = 1;
    int num_msgs = 20;
    int payload_sz = 2048;
    long long rperm_threshold = 131072;
    double arrival_time = 0.0, warm_up_time = 0.0, load = 0.0;
    double end = 0.0;  // --end (ROSS)
    std::string lp_io_dir;
    int quiet = 0;
};

struct Summary {
    uint64_t events = 0;
    uint64_t ev_by_type[N_EV_TYPES] = {0};
    long long total_hops = 0, finished_packets = 0, finished_msgs = 0, finished_chunks = 0, total_msg_sz = 0;
    double total_time = 0, max_latency = 0;
    long packet_gen = 0, packet_fin = 0, local_sr = 0, local_sg = 0, remote = 0;
    int minimal_count = 0, nonmin_count = 0;
    long long svr_msgs_recvd = 0;
    double svr_sum_latency = 0, svr_max_latency = 0, max_end_ts = 0;
    double total_offered_load = 0, total_observed_load = 0;
    double last_event_ts = 0;
    uint64_t rng_draws = 0;
    double run_seconds = 0;
};

enum { TR_UNIFORM = 1, TR_RAND_PERM = 2, TR_NEAREST_GROUP = 3, TR_NEAREST_NEIGHBOR = 4, TR_RANDOM_OTHER_GROUP = 5 };

class Sim {
public:
    Params p;
    Topo topo;
    RunOpts o;
    Summary sum;
    double lookahead = 0.005;  // g_tw_lookahead (ROSS default)
    double ts_end = 0;
    int num_nodes = 0, num_nodes_per_grp = 0;
    double mean_interval = 0, link_bandwidth = 0, load = 0;
    double now = 0;
    uint32_t cur_lp = 0;

    std::vector<SvrState> svr;
    std::vector<TermState> term;
    std::vector<RouterState> rtr;
    std::vector<Rng> rng0, rng2;  // per gid: stream 0 and stream 2 (codes_local_latency)
    std::vector<uint32_t> seq;    // per gid: events created so far
    std::vector<Event> pool;
    std::vector<uint32_t> free_list;
    std::priority_queue<HeapKey, std::vector<HeapKey>, HeapCmp> heap;
    std::map<std::string, std::string> lpio;  // identifier -> concatenated records
    std::vector<std::string> lpio_order;
    FILE* out = stdout;

    // ---- codes_mapping closed forms (codes_mapping.c:149-378)
    int n_lps() const { return p.repetitions * p.lps_per_rep; }
    uint64_t svr_gid(int k) const { return (uint64_t)(k / p.n_svr_per_rep) * p.lps_per_rep + p.off_svr + k % p.n_svr_per_rep; }
    uint64_t term_gid(int k) const { return (uint64_t)(k / p.n_term_per_rep) * p.lps_per_rep + p.off_term + k % p.n_term_per_rep; }
    uint64_t rtr_gid(int r) const { return (uint64_t)r * p.lps_per_rep + p.off_rtr; }
    int gid_rep(uint64_t g) const { return (int)(g / p.lps_per_rep); }
    int gid_off(uint64_t g) const { return (int)(g % p.lps_per_rep); }
    bool is_svr(uint64_t g) const { int o = gid_off(g); return o >= p.off_svr && o < p.off_svr + p.n_svr_per_rep; }
    bool is_term(uint64_t g) const { int o = gid_off(g); return o >= p.off_term && o < p.off_term + p.n_term_per_rep; }
    int svr_rel(uint64_t g) const { return gid_rep(g) * p.n_svr_per_rep + gid_off(g) - p.off_svr; }
    int term_rel(uint64_t g) const { return gid_rep(g) * p.n_term_per_rep + gid_off(g) - p.off_term; }
    int rtr_rel(uint64_t g) const { return gid_rep(g); }
    // model_net_find_local_device_mctx with CODES_MCTX_GROUP_MODULO (codes-mapping-context.c:131-132)
    uint64_t svr_to_term_gid(uint64_t sg) const {
        int rep = gid_rep(sg), off = gid_off(sg) - p.off_svr;
        return (uint64_t)rep * p.lps_per_rep + p.off_term + off % p.n_term_per_rep;
    }

    // ---- engine services (ROSS stand-ins)
    Event& ev_new(uint64_t dst, double offset, uint8_t type) {
        uint32_t idx;
        if (!free_list.empty()) {
            idx = free_list.back();
            free_list.pop_back();
        } else {
            idx = (uint32_t)pool.size();
            pool.emplace_back();
        }
        Event& e = pool[idx];
        e = Event();
        e.ts = now + offset;
        if (debug_maxoff && type < 32 && offset > maxoff[type]) maxoff[type] = offset;
        e.src = cur_lp;
        e.seq = seq[cur_lp]++;
        e.dst = (uint32_t)dst;
        e.type = type;
        pending_idx = idx;
        return e;
    }
    uint32_t pending_idx = 0;
    bool debug_maxoff = getenv("DFD_ORACLE_MAXOFF") != nullptr;  // largest (timestamp - now) per event type, printed at the end
    double maxoff[32] = {0};
    void ev_send() {
        Event& e = pool[pending_idx];
        if (e.ts >= ts_end) {  // events at/after the end barrier are never executed
            free_list.push_back(pending_idx);
            return;
        }
        heap.push(HeapKey{e.ts, e.src, e.seq, pending_idx});
    }
    double codes_local_latency(uint64_t gid) {  // codes/codes.h:61-66
        return lookahead + 0.5 + rng2[gid].unif() * 0.5;
    }

    // ============================== svr LP (synthetic) ==========================================
    double issue_event(uint64_t gid) {  // synthetic:231-271
        if (mean_interval == 0.0) {
            if (o.arrival_time != 0.0) {
                mean_interval = o.arrival_time;
                load = ns_to_GBps(mean_interval, o.payload_sz);
            } else if (load != 0.0) {
                mean_interval = bytes_to_ns(o.payload_sz, load * link_bandwidth);
            } else {
                mean_interval = 1000.0;
                load = ns_to_GBps(mean_interval, o.payload_sz);
            }
        }
        Event& e = ev_new(gid, mean_interval, SVR_KICKOFF);
        (void)e;
        ev_send();
        return mean_interval;
    }

    // model_net_event -> model_net_event_impl_base model-net.c:288-380 (+ noop path :252-286)
    void model_net_event(uint64_t sender, uint64_t final_dest_lp, uint64_t message_size, double offset) {
        if ((size_t)(72 + 72 + 416) > (size_t)p.message_size)
            die("Error: model_net trying to transmit an event of size %d but ROSS is configured for events of size %d\n", 560, p.message_size);
        uint64_t src_mn_lp = svr_to_term_gid(sender);
        uint64_t dest_mn_lp = svr_to_term_gid(final_dest_lp);
        if (src_mn_lp == dest_mn_lp && message_size < (uint64_t)p.node_eager_limit && !p.codes_noop_bypass) {
            double poffset = 0.0;  // model_net_noop_event
            double delay = codes_local_latency(sender);
            double sendTime = message_size * p.codes_cn_delay;
            poffset += delay;
            {
                Event& e = ev_new(sender, poffset + offset + sendTime, SVR_LOCAL);
                e.m.ev_src = sender;
                e.m.ev_start_time = now;
                ev_send();
            }
            poffset += delay;
            {
                Event& e = ev_new(final_dest_lp, poffset + offset + sendTime, SVR_REMOTE);
                e.m.ev_src = sender;
                e.m.ev_start_time = now;
                ev_send();
            }
            return;
        }
        double poffset = codes_local_latency(sender);
        Event& e = ev_new(src_mn_lp, poffset + offset, MN_NEW_MSG);
        e.qreq = 1;
        Req& r = e.rq;
        r.final_dest_lp = final_dest_lp;
        r.dest_mn_lp = dest_mn_lp;
        r.src_lp = sender;
        r.msg_size = message_size;
        r.remote_event_size = 72;
        r.self_event_size = 72;
        r.msg_start_time = now;
        r.ev_src = sender;
        r.ev_start_time = now;
        ev_send();
    }

    void handle_kickoff_event(uint64_t gid) {  // synthetic:329-413
        SvrState& ns = svr[svr_rel(gid)];
        if (ns.msg_sent_count >= o.num_msgs) return;
        int local_id = svr_rel(gid);
        long local_dest = -1;
        if (o.traffic == TR_UNIFORM) {
            local_dest = rng0[gid].integer(1, num_nodes - 2);
            local_dest = (ns.svr_id + local_dest) % num_nodes;
        } else if (o.traffic == TR_NEAREST_GROUP) {
            local_dest = (local_id + num_nodes_per_grp) % num_nodes;
        } else if (o.traffic == TR_NEAREST_NEIGHBOR) {
            local_dest = (local_id + 1) % num_nodes;
        } else if (o.traffic == TR_RAND_PERM) {
            if (ns.dest_id == -1 || ns.rperm_data > o.rperm_threshold) {
                ns.dest_id = (int)rng0[gid].integer(0, num_nodes - 1);
                local_dest = ns.dest_id;
                ns.rperm_data = o.payload_sz;
            } else {
                local_dest = ns.dest_id;
                ns.rperm_data += o.payload_sz;
            }
        } else if (o.traffic == TR_RANDOM_OTHER_GROUP) {
            int my_group_id = local_id / num_nodes_per_grp;
            std::vector<int> other;
            for (int i = 0; i < p.num_groups; i++)
                if (i != my_group_id) other.push_back(i);
            int rand_group = other[rng0[gid].integer(0, (long)other.size() - 1)];
            int rand_node = (int)rng0[gid].integer(0, num_nodes_per_grp - 1);
            local_dest = (long)rand_group * num_nodes_per_grp + rand_node;
        } else
            die("unknown traffic pattern %d", o.traffic);
        if (now >= o.warm_up_time) ns.warm_msg_sent_count++;
        assert(local_dest < num_nodes);
        uint64_t global_dest = svr_gid((int)local_dest);
        ns.msg_sent_count++;
        ns.last_send_ts = now;
        model_net_event(gid, global_dest, o.payload_sz, 0.0);
        issue_event(gid);
    }

    void handle_remote_event(uint64_t gid, const Event& ev) {  // synthetic:468-498
        SvrState& ns = svr[svr_rel(gid)];
        if (now >= o.warm_up_time) {
            ns.msg_recvd_count++;
            double packet_latency = now - ev.m.ev_start_time;
            ns.sum_server_latency += packet_latency;
            if (packet_latency > ns.max_server_latency) ns.max_server_latency = packet_latency;
            Event& e = ev_new(ev.m.ev_src, 0, SVR_ACK);
            (void)e;
            ev_send();
        }
    }
    void handle_local_event(uint64_t gid) {  // synthetic:508-516
        if (now >= o.warm_up_time) svr[svr_rel(gid)].local_recvd_count++;
    }
    void handle_ack_event(uint64_t gid) {  // synthetic:415-432
        SvrState& ns = svr[svr_rel(gid)];
        ns.msg_complete_count++;
        if (ns.first_recv_ts == -1) ns.first_recv_ts = now;
        ns.last_recv_ts = now;
        if (ns.last_recv_ts > sum.max_end_ts) sum.max_end_ts = ns.last_recv_ts;
    }

    // ============================== model-net base LP ===========================================
    // fcfs_next model-net-sched-impl.c:187-261; returns -1 when the queue is empty
    int fcfs_next(uint64_t gid, TermState& s) {
        if (s.sched_q.empty()) return -1;
        Req& q = s.sched_q.front();
        int is_last_packet;
        uint64_t psize;
        if (q.packet_size >= q.rem) {
            psize = q.rem;
            is_last_packet = 1;
        } else {
            psize = q.packet_size;
            is_last_packet = 0;
        }
        dragonfly_dally_packet_event(gid, q, psize, is_last_packet);
        if (is_last_packet) {
            s.sched_q.pop_front();
            return 1;
        }
        q.rem -= psize;
        return 0;
    }
    // dragonfly_dally_packet_event dally:5122-5183
    void dragonfly_dally_packet_event(uint64_t gid, const Req& req, uint64_t packet_size, int is_last_pckt) {
        Event& e = ev_new(gid, 0.0, T_GENERATE);
        Msg& m = e.m;
        m.final_dest_gid = req.final_dest_lp;
        m.total_size = (uint32_t)req.msg_size;
        m.sender_lp = req.src_lp;
        m.packet_size = (uint32_t)packet_size;
        m.remote_event_size_bytes = 0;
        m.local_event_size_bytes = 0;
        m.dest_terminal_lpid = req.dest_mn_lp;
        m.dfdally_src_terminal_id = term_rel(gid);
        m.dfdally_dest_terminal_id = term_rel(req.dest_mn_lp);
        m.message_id = (uint32_t)req.msg_id;
        m.msg_start_time = req.msg_start_time;
        if (is_last_pckt) {
            if (req.remote_event_size > 0) m.remote_event_size_bytes = req.remote_event_size;
            if (req.self_event_size > 0) m.local_event_size_bytes = req.self_event_size;
            m.ev_src = req.ev_src;
            m.ev_start_time = req.ev_start_time;
        }
        ev_send();
    }
    void handle_sched_next(uint64_t gid, TermState& s) {  // model-net-lp.c:838-873
        int ret = fcfs_next(gid, s);
        if (ret == -1) s.in_sched_send_loop = 0;
    }
    void handle_new_msg(uint64_t gid, const Event& ev) {  // model-net-lp.c:643-802
        TermState& s = term[term_rel(gid)];
        if (gid == ev.rq.dest_mn_lp) {
            // message to the sender's own terminal that did not take the no-op shortcut (size >= node_eager_limit, or
            // codes_noop_bypass): the base LP copies it inside the node, model-net-lp.c:672-704 (one node-copy queue)
            const Req& r = ev.rq;
            double exp_time = (s.node_copy_next_available_time > now) ? s.node_copy_next_available_time : now;
            exp_time += r.msg_size * p.codes_cn_delay;
            exp_time -= now;
            double delay = codes_local_latency(gid);
            s.node_copy_next_available_time = now + exp_time;
            if (r.remote_event_size > 0) {
                exp_time += delay;
                Event& e = ev_new(r.final_dest_lp, exp_time, SVR_REMOTE);
                e.m.ev_src = r.ev_src;
                e.m.ev_start_time = r.ev_start_time;
                ev_send();
            }
            if (r.self_event_size > 0) {
                exp_time += delay;
                Event& e = ev_new(r.src_lp, exp_time, SVR_LOCAL);
                e.m.ev_src = r.ev_src;
                e.m.ev_start_time = r.ev_start_time;
                ev_send();
            }
            return;
        }
        if (ev.qreq) {
            double exp_time = (s.next_available_time > now) ? s.next_available_time : now;
            exp_time += p.nic_seq_delay + codes_local_latency(gid);
            s.next_available_time = exp_time;
            Event& e = ev_new(gid, exp_time - now, MN_NEW_MSG);
            e.rq = ev.rq;
            e.qreq = 0;
            ev_send();
            return;
        }
        Req r = ev.rq;
        r.packet_size = p.mn_packet_size;
        r.msg_id = s.msg_id++;
        r.rem = r.msg_size;  // fcfs_add :142-170
        s.sched_q.push_back(r);
        if (s.in_sched_send_loop == 0) {
            s.in_sched_send_loop = 1;
            handle_sched_next(gid, s);
        }
    }
    void idle_event2(uint64_t gid, double offset) {  // model_net_method_idle_event2 model-net-lp.c:969-982
        Event& e = ev_new(gid, offset, MN_SCHED_NEXT);
        (void)e;
        ev_send();
    }

    // ============================== terminal LP =================================================
    uint64_t num_chunks_for(uint64_t sz) const { return sz ? (sz + p.chunk_size - 1) / p.chunk_size : 1; }  // dally:103-104

    void packet_generate(uint64_t gid, const Event& ev) {  // dally:5417-5774
        TermState& s = term[term_rel(gid)];
        Msg msg = ev.m;
        sum.packet_gen++;
        s.packet_gen++;
        s.total_gen_size += msg.packet_size;
        assert(gid != msg.dest_terminal_lpid);
        uint64_t num_chunks = num_chunks_for(msg.packet_size);
        int dest_router_id = msg.dfdally_dest_terminal_id / p.num_cn;
        int dest_grp_id = dest_router_id / p.num_routers;
        int src_grp_id = s.router_id / p.num_routers;
        if (src_grp_id == dest_grp_id) {
            if (dest_router_id == (int)s.router_id)
                sum.local_sr++;
            else
                sum.local_sg++;
        } else
            sum.remote++;
        msg.packet_ID = s.packet_counter;
        s.packet_counter++;
        msg.my_N_hop = 0;
        msg.my_l_hop = 0;
        msg.my_g_hop = 0;
        msg.my_hops_cur_group = 0;
        s.last_in_queue_time = now;
        for (uint64_t i = 0; i < num_chunks; i++) {
            Msg c = msg;
            c.output_chan = 0;
            c.chunk_id = (uint32_t)i;
            c.origin_router_id = s.router_id;
            s.terminal_msgs.push_back(c);
            s.terminal_length += p.chunk_size;
        }
        double injection_ts = bytes_to_ns(msg.packet_size, p.cn_bandwidth);
        double nic_ts = injection_ts;
        if (s.terminal_length < p.cn_vc_size) {
            idle_event2(gid, nic_ts);
        } else {
            s.issueIdle = 1;
            if (s.last_buf_full == 0.0) s.last_buf_full = now;
        }
        if (s.in_send_loop == 0) {
            Event& e = ev_new(gid, 0.0, T_SEND);
            (void)e;
            s.in_send_loop = 1;
            ev_send();
        }
        long total_event_size = 416 + msg.remote_event_size_bytes + msg.local_event_size_bytes;  // model_net_get_msg_sz() = sizeof(model_net_wrap_msg) = 416 (model-net.c:500-504)
        MnStats& st = s.stats;
        st.used = true;
        st.send_count++;
        st.send_bytes += msg.packet_size;
        st.send_time += (1 / p.cn_bandwidth) * msg.packet_size;
        if (st.max_event_size < total_event_size) st.max_event_size = total_event_size;
    }

    void packet_send(uint64_t gid) {  // dally:5836-6026, get_next_vcg :3440-3449
        TermState& s = term[term_rel(gid)];
        if (s.terminal_msgs.empty() || s.vc_occupancy + p.chunk_size > p.cn_vc_size) {
            s.in_send_loop = 0;
            s.stalled_chunks++;
            if (!s.last_buf_full) s.last_buf_full = now;
            return;
        }
        Msg cur = s.terminal_msgs.front();
        int data_size = p.chunk_size;
        uint64_t num_chunks = num_chunks_for(cur.packet_size);
        cur.travel_start_time = now;
        double injection_delay = bytes_to_ns(p.chunk_size, p.cn_bandwidth);
        if ((cur.packet_size < (uint32_t)p.chunk_size) && (cur.chunk_id == num_chunks - 1)) {
            data_size = cur.packet_size % p.chunk_size;
            injection_delay = bytes_to_ns(data_size, p.cn_bandwidth);
        }
        double propagation_delay = p.cn_delay;
        s.terminal_available_time = maxd(s.terminal_available_time, now);
        s.terminal_available_time += injection_delay;
        double injection_ts = s.terminal_available_time - now;
        double propagation_ts = injection_ts + propagation_delay;
        {
            Event& e = ev_new(s.router_lp, propagation_ts, R_ARRIVE);
            e.m = cur;
            Msg& m = e.m;
            m.src_terminal_id = gid;
            m.dfdally_src_terminal_id = s.terminal_id;
            m.vc_index = 0;
            m.last_hop = TERMINAL;
            m.path_type = -1;
            m.local_event_size_bytes = 0;
            m.is_intm_visited = 0;
            m.intm_grp_id = -1;
            ev_send();
        }
        if (cur.chunk_id == num_chunks - 1 && cur.local_event_size_bytes > 0) {
            Event& e = ev_new(cur.sender_lp, 0, SVR_LOCAL);
            e.m.ev_src = cur.ev_src;
            e.m.ev_start_time = cur.ev_start_time;
            ev_send();
        }
        s.vc_occupancy += p.chunk_size;
        s.terminal_msgs.pop_front();
        s.terminal_length -= p.chunk_size;
        s.link_traffic += p.chunk_size;
        s.total_chunks++;
        s.injected_chunks++;
        if (!s.terminal_msgs.empty() && s.vc_occupancy + p.chunk_size <= p.cn_vc_size) {
            Event& e = ev_new(gid, injection_ts, T_SEND);
            (void)e;
            ev_send();
        } else {
            s.in_send_loop = 0;
        }
        if (s.issueIdle && s.terminal_length < p.cn_vc_size) {
            s.issueIdle = 0;
            idle_event2(gid, injection_ts);
            if (s.last_buf_full > 0.0) {
                s.busy_time += (now - s.last_buf_full);
                s.last_buf_full = 0.0;
            }
        }
    }

    void terminal_buf_update(uint64_t gid) {  // dally:6762-6789
        TermState& s = term[term_rel(gid)];
        s.vc_occupancy -= p.chunk_size;
        if (s.in_send_loop == 0 && !s.terminal_msgs.empty()) {
            Event& e = ev_new(gid, 0.0, T_SEND);
            (void)e;
            s.in_send_loop = 1;
            ev_send();
        }
    }

    void packet_arrive(uint64_t gid, const Event& ev) {  // dally:6453-6744
        TermState& s = term[term_rel(gid)];
        const Msg& msg = ev.m;
        if (msg.dfdally_dest_terminal_id != s.terminal_id) die("Packet arrived at wrong terminal\n");
        assert(gid == msg.dest_terminal_lpid);
        {
            Event& e = ev_new(msg.intm_lp_id, p.cn_credit_delay, R_BUFFER);
            e.m.vc_index = msg.vc_index;
            e.m.output_chan = msg.output_chan;
            ev_send();
        }
        sum.finished_chunks++;
        s.finished_chunks++;
        s.ejected_chunks++;
        assert(gid != msg.src_terminal_id);
        uint64_t num_chunks = num_chunks_for(msg.packet_size);
        if (msg.path_type == MINIMAL) sum.minimal_count++;
        if (msg.path_type == NON_MINIMAL) sum.nonmin_count++;
        if (msg.path_type != MINIMAL && msg.path_type != NON_MINIMAL)
            fprintf(out, "\n Wrong message path type %d ", msg.path_type);
        double ete_latency = now - msg.travel_start_time;
        s.total_time += ete_latency;
        sum.total_hops += msg.my_N_hop;
        s.total_hops += msg.my_N_hop;
        MnStats& st = s.stats;
        st.used = true;
        st.recv_time += ete_latency;
        if (msg.chunk_id == num_chunks - 1) {
            s.packet_fin++;
            sum.packet_fin++;
            st.recv_count++;
            st.recv_bytes += msg.packet_size;
            sum.finished_packets++;
            s.finished_packets++;
        }
        if (s.min_latency > ete_latency) s.min_latency = ete_latency;
        if (s.max_latency < ete_latency) s.max_latency = ete_latency;
        std::pair<uint64_t, unsigned> packet_key{msg.packet_ID, msg.dfdally_src_terminal_id};
        bool has_remaining_sz = s.remaining_sz_packets.count(packet_key) == 1;
        bool is_packet_completed = false;
        int chunk_size = p.chunk_size;
        if (has_remaining_sz) {
            int actual = std::min(chunk_size, (int)s.remaining_sz_packets[packet_key]);
            s.remaining_sz_packets[packet_key] -= actual;
            if (s.remaining_sz_packets[packet_key] == 0) {
                is_packet_completed = true;
                s.remaining_sz_packets.erase(packet_key);
            }
        } else {
            if ((uint32_t)chunk_size < msg.packet_size)
                s.remaining_sz_packets[packet_key] = msg.packet_size - chunk_size;
            else
                is_packet_completed = true;
        }
        std::pair<uint64_t, uint64_t> key{msg.message_id, msg.sender_lp};
        auto it = s.rank_tbl.find(key);
        if (it == s.rank_tbl.end()) {
            uint64_t packet_size = p.packet_size;
            uint64_t total_packets = msg.total_size / packet_size + (msg.total_size % packet_size ? 1 : 0);
            if (total_packets == 0) total_packets = 1;
            RankEntry d;
            d.remaining_packets = total_packets;
            it = s.rank_tbl.emplace(key, d).first;
            s.rank_tbl_pop++;
            if (s.rank_tbl_pop >= 4999) die("\n Exceeded allocated qhash size, increase hash size in dragonfly model");
        }
        RankEntry& tmp = it->second;
        tmp.num_chunks++;
        if (msg.remote_event_size_bytes > 0 && !tmp.has_remote) {
            tmp.has_remote = true;
            tmp.remote_event_size = msg.remote_event_size_bytes;
            tmp.ev_src = msg.ev_src;
            tmp.ev_start_time = msg.ev_start_time;
        }
        if (is_packet_completed) tmp.remaining_packets--;
        if (tmp.remaining_packets == 0) {
            sum.finished_msgs++;
            sum.total_msg_sz += msg.total_size;
            s.total_msg_size += msg.total_size;
            s.finished_msgs++;
            if (tmp.has_remote && tmp.remote_event_size > 0) {  // send_remote_event dally:6124-6151
                Event& e = ev_new(msg.final_dest_gid, 0, SVR_REMOTE);
                e.m.ev_src = tmp.ev_src;
                e.m.ev_start_time = tmp.ev_start_time;
                ev_send();
            }
            s.rank_tbl.erase(it);
            s.rank_tbl_pop--;
        }
    }

    // ============================== router LP ===================================================
    int score_connection(const RouterState& s, const Conn& c, bool nonmin) const {  // dally:1649-1684
        if (c.port == -1) return INT_MAX;
        int score = 0;
        for (int k = 0; k < p.num_vcs; k++) score += s.vc_occupancy[c.port * p.num_vcs + k];
        score += s.queued_count[c.port];
        if (p.scoring == DELTA && nonmin) score = score * 2;
        return score;
    }
    Conn best_from_conns(const RouterState& s, uint64_t gid, const std::vector<Conn>& conns) {  // dally:1687-1723
        if (conns.empty()) return Conn();
        if (conns.size() == 1) return conns[0];
        std::vector<Conn> best;
        int best_score = INT_MAX;
        for (const Conn& c : conns) {
            int sc = score_connection(s, c, false);
            if (sc <= best_score) {
                if (sc < best_score) {
                    best_score = sc;
                    best.clear();
                }
                best.push_back(c);
            }
        }
        assert(!best.empty());
        return best[rng0[gid].integer(0, (long)best.size() - 1)];
    }
    std::vector<Conn> poll_k(uint64_t gid, const std::vector<Conn>& conns, int k) {  // dally:1796-1849
        std::vector<Conn> out;
        if (conns.empty()) return out;
        if (conns.size() == 1) {
            out.push_back(conns[0]);
            return out;
        }
        if (k == 2) {
            long n = (long)conns.size();
            long s1 = rng0[gid].integer(0, n - 1);
            long off = rng0[gid].integer(1, n - 1);
            long s2 = (s1 + off) % n;
            out.push_back(conns[s1]);
            out.push_back(conns[s2]);
            return out;
        }
        long last_sel = 0;
        std::set<long> sels;
        for (int i = 0; i < k; i++) {
            long ri = rng0[gid].integer(0, ((long)conns.size() - 1) - (long)sels.size());
            long att = (last_sel + ri) % (long)conns.size();
            while (sels.count(att)) att = (att + 1) % (long)conns.size();
            sels.insert(att);
            last_sel = att;
        }
        for (long i : sels) out.push_back(conns[i]);
        return out;
    }
    void select_intermediate_group(RouterState& s, uint64_t gid, Msg& msg, int fdest_router_id) {  // dally:9738-9790
        int fdest_group_id = fdest_router_id / p.num_routers;
        int origin_group_id = msg.origin_router_id / p.num_routers;
        if (msg.intm_grp_id == -1) {
            assert(s.router_id == msg.origin_router_id);
            std::vector<int> group_list;
            for (int i = 0; i < p.num_groups; i++)
                if (i != origin_group_id && i != fdest_group_id) group_list.push_back(i);
            long sel = rng0[gid].integer(0, (long)group_list.size() - 1);
            msg.intm_grp_id = group_list[sel];
        } else {
            assert(s.router_id != msg.origin_router_id);
            std::set<int> valid;
            for (auto& kv : topo.r[s.router_id].global)
                for (const Conn& c : kv.second)
                    if (c.dest_group != fdest_group_id && c.dest_group != origin_group_id) valid.insert(c.dest_group);
            long sel = rng0[gid].integer(0, (long)valid.size() - 1);
            auto it = valid.begin();
            std::advance(it, sel);
            msg.intm_grp_id = *it;
        }
    }
    std::vector<Conn> legal_minimal_stops(const RouterState& s, int fdest_router_id) {  // dally:9793-9849
        int fdest_group_id = fdest_router_id / p.num_routers;
        if (s.group_id != fdest_group_id) {
            const std::vector<Conn>& d = topo.conns_to_group(s.router_id, fdest_group_id);
            if (!d.empty()) return d;
            return topo.routed_to_group(s.router_id, s.group_id, fdest_group_id);
        }
        return topo.conns_to_gid_local(s.router_id, fdest_router_id);
    }
    std::vector<Conn> legal_nonminimal_stops(RouterState& s, uint64_t gid, Msg& msg, int fdest_router_id) {  // dally:9853-9910
        int origin_group_id = msg.origin_router_id / p.num_routers;
        int fdest_group_id = fdest_router_id / p.num_routers;
        bool in_intm = (s.group_id != origin_group_id) && (s.group_id != fdest_group_id);
        if (s.group_id == origin_group_id) {
            std::vector<Conn> c = topo.conns_to_group(s.router_id, msg.intm_grp_id);
            if (s.router_id == msg.origin_router_id) {
                if (!c.empty()) return c;
                return topo.routed_to_group(s.router_id, s.group_id, msg.intm_grp_id);
            }
            if (!c.empty()) return c;
            select_intermediate_group(s, gid, msg, fdest_router_id);
            return topo.conns_to_group(s.router_id, msg.intm_grp_id);
        }
        (void)in_intm;
        return std::vector<Conn>();
    }
    Conn minimal_routing(RouterState& s, uint64_t gid, Msg& msg, int fdest) {  // dally:9913-9932
        std::vector<Conn> c = legal_minimal_stops(s, fdest);
        if (c.empty())
            die("%d group %d: MINIMAL DEAD END to %d in group %d - (s%d -> d%d)\n", s.router_id, s.group_id, fdest,
                fdest / p.num_routers, msg.origin_router_id, fdest);
        return c[rng0[gid].integer(0, (long)c.size() - 1)];
    }
    Conn nonminimal_routing(RouterState& s, uint64_t gid, Msg& msg, int fdest) {  // dally:9939-9992
        int fdest_group_id = fdest / p.num_routers;
        assert(msg.intm_grp_id != -1);
        if (s.group_id == msg.intm_grp_id) msg.is_intm_visited = 1;
        int next_group = msg.is_intm_visited == 1 ? fdest_group_id : msg.intm_grp_id;
        const std::vector<Conn>& d = topo.conns_to_group(s.router_id, next_group);
        if (!d.empty()) return d[rng0[gid].integer(0, (long)d.size() - 1)];
        std::vector<Conn> t = topo.routed_to_group(s.router_id, s.group_id, next_group);
        if (t.empty()) die("router %d: NONMINIMAL DEAD END toward group %d", s.router_id, next_group);
        return t[rng0[gid].integer(0, (long)t.size() - 1)];
    }
    Conn prog_adaptive_routing(RouterState& s, uint64_t gid, Msg& msg, int fdest) {  // dally:9995-10061
        if (s.group_id == msg.intm_grp_id) msg.is_intm_visited = 1;
        std::vector<Conn> mins = legal_minimal_stops(s, fdest);
        std::vector<Conn> nonmins = legal_nonminimal_stops(s, gid, msg, fdest);
        int type_min = CONN_LOCAL, type_non = CONN_LOCAL;
        if (!mins.empty()) type_min = mins[0].type;
        if (!nonmins.empty()) type_non = nonmins[0].type;
        Conn best_min, best_non;
        if (type_min == CONN_GLOBAL)
            best_min = best_from_conns(s, gid, poll_k(gid, mins, p.global_k_picks));
        else
            best_min = best_from_conns(s, gid, mins);
        if (type_non == CONN_GLOBAL)
            best_non = best_from_conns(s, gid, poll_k(gid, nonmins, p.global_k_picks));
        else
            best_non = best_from_conns(s, gid, nonmins);
        int min_score = score_connection(s, best_min, false);
        int non_score = score_connection(s, best_non, true);
        if (msg.path_type == NON_MINIMAL && msg.is_intm_visited != 1) return best_non;
        if (min_score <= p.adaptive_threshold) return best_min;
        if (min_score <= non_score) return best_min;
        msg.path_type = NON_MINIMAL;
        return best_non;
    }
    Conn do_routing(RouterState& s, uint64_t gid, Msg& msg, int fdest) {  // dally:7110-7215
        int fdest_group_id = fdest / p.num_routers;
        if ((int)s.router_id == fdest) {
            auto it = topo.r[s.router_id].term.find(msg.dfdally_dest_terminal_id);
            if (it == topo.r[s.router_id].term.end() || it->second.empty())
                die("Destination Router %d: No connection to destination terminal %d\n", s.router_id, msg.dfdally_dest_terminal_id);
            return best_from_conns(s, gid, it->second);
        } else if (s.group_id == fdest_group_id) {
            std::vector<Conn> c = topo.conns_to_gid_local(s.router_id, fdest);
            if (c.empty()) die("Destination Group %d: No connection to destination router %d\n", s.router_id, fdest);
            if (p.routing == PROG_ADAPTIVE) return best_from_conns(s, gid, c);
            return c[rng0[gid].integer(0, (long)c.size() - 1)];
        }
        if (msg.last_hop == TERMINAL && msg.intm_grp_id == -1) select_intermediate_group(s, gid, msg, fdest);
        switch (p.routing) {
        case MINIMAL: return minimal_routing(s, gid, msg, fdest);
        case NON_MINIMAL: msg.path_type = NON_MINIMAL; return nonminimal_routing(s, gid, msg, fdest);
        case PROG_ADAPTIVE: return prog_adaptive_routing(s, gid, msg, fdest);
        }
        die("Error: No valid routing algorithm specified");
    }

    void router_credit_send(const Msg& msg, bool from_queue) {  // dally:7269-7326
        uint64_t dest;
        uint8_t type = R_BUFFER;
        double credit_delay;
        if (msg.last_hop == TERMINAL) {
            dest = msg.src_terminal_id;
            type = T_BUFFER;
            credit_delay = p.cn_credit_delay;
        } else if (msg.last_hop == GLOBAL) {
            dest = msg.intm_lp_id;
            credit_delay = p.global_credit_delay;
        } else {
            dest = msg.intm_lp_id;
            credit_delay = p.local_credit_delay;
        }
        Event& e = ev_new(dest, credit_delay, type);
        if (!from_queue) {
            e.m.vc_index = msg.vc_index;
            e.m.output_chan = msg.output_chan;
        } else {
            e.m.vc_index = msg.saved_vc;
            e.m.output_chan = msg.saved_channel;
        }
        ev_send();
    }

    void router_packet_receive(uint64_t gid, const Event& ev) {  // dally:7377-7595
        RouterState& s = rtr[rtr_rel(gid)];
        const Msg& msg = ev.m;
        int num_routers = p.num_routers;
        int dest_router_id = msg.dfdally_dest_terminal_id / p.num_cn;
        Msg cur = msg;
        if (cur.last_hop == TERMINAL) cur.path_type = MINIMAL;
        Conn next = do_routing(s, gid, cur, dest_router_id);
        int output_port = next.port;
        int next_stop;
        if (next.type != CONN_TERMINAL)
            next_stop = (int)rtr_gid(next.dest_gid);
        else
            next_stop = (int)cur.dest_terminal_lpid;
        assert(output_port >= 0);
        cur.vc_index = (short)output_port;
        cur.next_stop = next_stop;
        int max_vc_size = p.cn_vc_size;
        int my_group_id = s.group_id;
        int src_group_id = msg.origin_router_id / num_routers;
        int dest_group_id = dest_router_id / num_routers;
        int output_chan = 0;
        if (my_group_id == src_group_id) {
            output_chan = cur.my_l_hop;
            if (msg.path_type == NON_MINIMAL) output_chan = 1;
        } else if (my_group_id != src_group_id && my_group_id != dest_group_id) {
            assert(msg.path_type == NON_MINIMAL);
            output_chan = 2;
        } else if (my_group_id == dest_group_id) {
            output_chan = 3;
        }
        if (next.type == CONN_LOCAL) {
            max_vc_size = p.local_vc_size;
            cur.my_l_hop++;
            cur.my_hops_cur_group++;
        }
        if (next.type == CONN_GLOBAL) {
            max_vc_size = p.global_vc_size;
            cur.my_hops_cur_group = 0;
            cur.my_g_hop++;
        }
        if (!(output_chan < p.num_vcs && output_chan >= 0))
            die("\n Output channel %d great than available VCs %d", output_chan, p.num_vcs - 1);
        cur.output_chan = output_chan;
        cur.my_N_hop++;
        if (output_port >= p.radix) die("\n Output port greater than router radix %d ", output_port);
        int q = output_port * p.num_vcs + output_chan;
        if (s.vc_occupancy[q] + p.chunk_size <= max_vc_size) {
            router_credit_send(msg, false);
            s.pending_msgs[q].push_back(cur);
            s.vc_occupancy[q] += p.chunk_size;
            if (s.in_send_loop[output_port] == 0) {
                double ts = maxd(s.next_output_available_time[output_port], now) - now;
                Event& e = ev_new(gid, ts, R_SEND);
                e.m.vc_index = (short)output_port;
                ev_send();
                s.in_send_loop[output_port] = 1;
            }
        } else {
            s.stalled_chunks[output_port]++;
            cur.saved_vc = msg.vc_index;
            cur.saved_channel = (short)msg.output_chan;
            s.queued_msgs[q].push_back(cur);
            s.queued_count[output_port] += p.chunk_size;
            if (s.last_buf_full[output_port] == 0.0) s.last_buf_full[output_port] = now;
        }
    }

    void router_packet_send(uint64_t gid, const Event& ev) {  // dally:7705-8032
        RouterState& s = rtr[rtr_rel(gid)];
        int output_port = ev.m.vc_index;
        // get_next_router_vcg dally:3499-3557 (num_qos_levels == 1)
        int output_chan = -1;
        {
            int next_rr_vc = (s.last_qos_lvl[output_port] + 1) % p.num_vcs;
            for (int i = 0; i < p.num_vcs; i++) {
                if (!s.pending_msgs[output_port * p.num_vcs + next_rr_vc].empty()) {
                    s.last_qos_lvl[output_port] = next_rr_vc;
                    output_chan = next_rr_vc;
                    break;
                }
                next_rr_vc = (next_rr_vc + 1) % p.num_vcs;
            }
        }
        if (output_chan < 0) {
            s.in_send_loop[output_port] = 0;
            if (s.queued_count[output_port] && !s.last_buf_full[output_port]) s.last_buf_full[output_port] = now;
            return;
        }
        int q = output_port * p.num_vcs + output_chan;
        Msg cur = s.pending_msgs[q].front();
        if (s.last_buf_full[output_port]) {
            s.busy_time[output_port] += (now - s.last_buf_full[output_port]);
            s.last_buf_full[output_port] = 0.0;
        }
        int to_terminal = 1, global = 0;
        double delay = p.cn_delay;
        double bandwidth = p.cn_bandwidth;
        if (output_port < p.intra_grp_radix) {
            to_terminal = 0;
            delay = p.local_delay;
            bandwidth = p.local_bandwidth;
        } else if (output_port < p.intra_grp_radix + p.num_global_channels) {
            to_terminal = 0;
            global = 1;
            delay = p.global_delay;
            bandwidth = p.global_bandwidth;
        }
        uint64_t num_chunks = num_chunks_for(cur.packet_size);
        double propagation_delay = delay;
        double injection_delay = bytes_to_ns(p.chunk_size, bandwidth);
        if (cur.packet_size == 0) injection_delay = bytes_to_ns(p.credit_size, bandwidth);
        if ((cur.packet_size < (uint32_t)p.chunk_size) && (cur.chunk_id == num_chunks - 1))
            injection_delay = bytes_to_ns(cur.packet_size % p.chunk_size, bandwidth);
        injection_delay += p.router_delay;
        s.next_output_available_time[output_port] = maxd(s.next_output_available_time[output_port], now);
        s.next_output_available_time[output_port] += injection_delay;
        double injection_ts = s.next_output_available_time[output_port] - now;
        double propagation_ts = injection_ts + propagation_delay;
        {
            Event& e = ev_new((uint64_t)cur.next_stop, propagation_ts, to_terminal ? T_ARRIVE : R_ARRIVE);
            e.m = cur;
            e.m.last_hop = global ? GLOBAL : LOCAL;
            e.m.intm_lp_id = (unsigned)gid;
            ev_send();
        }
        if (((cur.packet_size % p.chunk_size) || cur.packet_size == 0) && (cur.chunk_id == num_chunks - 1))
            s.link_traffic[output_port] += (cur.packet_size % p.chunk_size);
        else
            s.link_traffic[output_port] += p.chunk_size;
        s.total_chunks[output_port]++;
        s.pending_msgs[q].pop_front();
        s.next_output_available_time[output_port] -= p.router_delay;
        injection_ts -= p.router_delay;
        int next_output_chan = -1;
        for (int k = 0; k < p.num_vcs; k++)
            if (!s.pending_msgs[output_port * p.num_vcs + k].empty()) {
                next_output_chan = k;
                break;
            }
        if (next_output_chan < 0) {
            s.in_send_loop[output_port] = 0;
            return;
        }
        Event& e = ev_new(gid, injection_ts, R_SEND);
        e.m.vc_index = (short)output_port;
        ev_send();
    }

    void router_buf_update(uint64_t gid, const Event& ev) {  // dally:8069-8145
        RouterState& s = rtr[rtr_rel(gid)];
        int indx = ev.m.vc_index;
        int output_chan = ev.m.output_chan;
        int q = indx * p.num_vcs + output_chan;
        s.vc_occupancy[q] -= p.chunk_size;
        if (s.last_buf_full[indx] > 0.0) {
            s.busy_time[indx] += (now - s.last_buf_full[indx]);
            s.last_buf_full[indx] = 0.0;
        }
        if (!s.queued_msgs[q].empty()) {
            Msg head = s.queued_msgs[q].front();
            s.queued_msgs[q].pop_front();
            router_credit_send(head, true);
            s.pending_msgs[q].push_back(head);
            s.vc_occupancy[q] += p.chunk_size;
            s.queued_count[indx] -= p.chunk_size;
        }
        if (s.in_send_loop[indx] == 0 && !s.pending_msgs[q].empty()) {
            double ts = maxd(s.next_output_available_time[indx], now) - now;
            Event& e = ev_new(gid, ts, R_SEND);
            e.m.vc_index = (short)indx;
            s.in_send_loop[indx] = 1;
            ev_send();
        }
    }

    // ============================== router VC-occupancy snapshots ================================
    // router_send_snapshot_events dally:4363-4417: router 0 writes the header record at init; every router sends itself
    // one R_SNAPSHOT per configured timestamp (packet_ID = index of the snapshot)
    void router_send_snapshot_events(uint64_t gid, RouterState& s) {
        if (s.router_id == 0) {
            std::string line = "#Time of snapshot,Router ID,";
            for (int i = 0; i < p.radix; i++)
                for (int j = 0; j < p.num_vcs; j++) line += fmt("Port %d VC %d,", i, j);
            line[line.size() - 1] = '\n';
            lp_io_write("dragonfly-snapshots.csv", line);
        }
        for (size_t i = 0; i < p.snapshot_times.size(); i++) {
            Event& e = ev_new(gid, p.snapshot_times[i], R_SNAPSHOT);
            e.m.packet_ID = i;
            ev_send();
        }
    }
    // router_handle_snapshot_event dally:4419-4437 stores the occupancies; the commit handler dally:4687-4709 writes the
    // record.  A sequential run commits every event right after executing it, so both happen here.
    void router_handle_snapshot_event(uint64_t gid, const Event& ev) {
        RouterState& s = rtr[rtr_rel(gid)];
        size_t k = (size_t)ev.m.packet_ID;
        if (k >= p.snapshot_times.size()) return;
        for (int i = 0; i < p.radix; i++)
            for (int j = 0; j < p.num_vcs; j++) s.snapshot_data[k][i * p.num_vcs + j] = s.vc_occupancy[i * p.num_vcs + j];
        std::string line = fmt("%.4f, %d, ", p.snapshot_times[k], (int)s.router_id);
        for (int i = 0; i < p.radix; i++)
            for (int j = 0; j < p.num_vcs; j++) line += fmt("%d, ", s.snapshot_data[k][i * p.num_vcs + j]);
        line.resize(line.size() - 2);
        line += '\n';
        lp_io_write("dragonfly-snapshots.csv", line);
    }

    // ============================== set-up / run / finalize =====================================
    void init() {
        int n = n_lps();
        num_nodes = p.repetitions * p.n_svr_per_rep;
        num_nodes_per_grp = num_nodes / p.num_groups;
        link_bandwidth = p.svr_link_bandwidth;
        load = o.load;
        seq.assign(n, 0);
        rng0.resize(n);
        rng2.resize(n);
        for (int g = 0; g < n; g++) {
            rng0[g].seed_stream((uint64_t)g * 3 + 0);
            rng2[g].seed_stream((uint64_t)g * 3 + 2);
        }
        svr.assign(num_nodes, SvrState());
        term.assign(p.total_terminals, TermState());
        rtr.assign(p.total_routers, RouterState());
        ts_end = o.end > 0 ? o.end : (p.end_time > 0 ? p.end_time : 5.0 * 24 * 60 * 60 * 1e9);  // synthetic:720-721
        // LP init in gid order (tw_run): svr_init synthetic:287-298, terminal_dally_init dally:4713-4904,
        // router_dally_init dally:4906-5111
        for (int g = 0; g < n; g++) {
            cur_lp = g;
            now = 0;
            if (is_svr(g)) {
                SvrState& ns = svr[svr_rel(g)];
                ns.svr_id = svr_rel(g);
                ns.start_ts = issue_event(g);
            } else if (is_term(g)) {
                TermState& s = term[term_rel(g)];
                s.terminal_id = term_rel(g);
                s.router_id = s.terminal_id / p.num_cn;
                s.router_lp = rtr_gid(s.router_id);
            } else {
                RouterState& r = rtr[rtr_rel(g)];
                r.router_id = rtr_rel(g);
                r.group_id = r.router_id / p.num_routers;
                int R = p.radix;
                r.next_output_available_time.assign(R, 0);
                r.last_buf_full.assign(R, 0);
                r.busy_time.assign(R, 0);
                r.stalled_chunks.assign(R, 0);
                r.total_chunks.assign(R, 0);
                r.in_send_loop.assign(R, 0);
                r.queued_count.assign(R, 0);
                r.last_qos_lvl.assign(R, 0);
                r.link_traffic.assign(R, 0);
                r.vc_occupancy.assign(R * p.num_vcs, 0);
                r.pending_msgs.assign(R * p.num_vcs, std::deque<Msg>());
                r.queued_msgs.assign(R * p.num_vcs, std::deque<Msg>());
                if (!p.snapshot_times.empty()) {  // dally:5084-5091
                    r.snapshot_data.assign(p.snapshot_times.size(), std::vector<int>((size_t)R * p.num_vcs, 0));
                    router_send_snapshot_events(g, r);
                }
            }
        }
    }

    void dispatch(const Event& ev) {
        uint64_t gid = ev.dst;
        switch (ev.type) {
        case SVR_KICKOFF: handle_kickoff_event(gid); break;
        case SVR_REMOTE: handle_remote_event(gid, ev); break;
        case SVR_LOCAL: handle_local_event(gid); break;
        case SVR_ACK: handle_ack_event(gid); break;
        case MN_NEW_MSG: handle_new_msg(gid, ev); break;
        case MN_SCHED_NEXT: handle_sched_next(gid, term[term_rel(gid)]); break;
        case T_GENERATE: packet_generate(gid, ev); break;
        case T_ARRIVE: packet_arrive(gid, ev); break;
        case T_SEND: packet_send(gid); break;
        case T_BUFFER: terminal_buf_update(gid); break;
        case R_SEND: router_packet_send(gid, ev); break;
        case R_ARRIVE: router_packet_receive(gid, ev); break;
        case R_BUFFER: router_buf_update(gid, ev); break;
        case R_SNAPSHOT: router_handle_snapshot_event(gid, ev); break;
        default: die("bad event type");
        }
    }

    void run() {
        while (!heap.empty()) {
            HeapKey k = heap.top();
            heap.pop();
            Event ev = pool[k.idx];  // copy: handlers allocate from the pool
            free_list.push_back(k.idx);
            now = ev.ts;
            cur_lp = ev.dst;
            sum.events++;
            sum.ev_by_type[ev.type]++;
            sum.last_event_ts = now;
            dispatch(ev);
        }
        if (debug_maxoff)
            for (int t = 0; t < 32; t++)
                if (maxoff[t] > 0) fprintf(stderr, "oracle: event type %d: largest (timestamp - now) = %.3f ns\n", t, maxoff[t]);
    }

    void lp_io_write(const std::string& id, const std::string& data) {  // lp-io.c:39-100
        if (!lpio.count(id)) lpio_order.push_back(id);
        lpio[id] += data;
    }
    static std::string fmt(const char* f, ...) {
        char buf[2048];
        va_list ap;
        va_start(ap, f);
        vsnprintf(buf, sizeof buf, f, ap);
        va_end(ap);
        return buf;
    }
    void write_mn_stats(uint64_t gid, const char* cat, const MnStats& st) {  // model-net.c:186-203
        lp_io_write(std::string("model-net-category-") + cat,
                    fmt("lp:%ld\tsend_count:%ld\tsend_bytes:%ld\tsend_time:%f\trecv_count:%ld\trecv_bytes:%ld\trecv_time:%f\tmax_event_size:%ld\n",
                        (long)gid, st.send_count, st.send_bytes, st.send_time, st.recv_count, st.recv_bytes, st.recv_time, st.max_event_size));
    }

    std::string raw;  // exact (hex-float) per-LP statistics for bit-exact comparison

    void finalize() {
        int n = n_lps();
        for (int g = 0; g < n; g++) {
            if (is_svr(g)) {  // svr_finalize synthetic:523-577
                SvrState& ns = svr[svr_rel(g)];
                sum.svr_sum_latency += ns.sum_server_latency;
                sum.svr_msgs_recvd += ns.msg_recvd_count;
                if (ns.max_server_latency > sum.svr_max_latency) sum.svr_max_latency = ns.max_server_latency;
                double observed_load_time = ((double)ns.last_recv_ts - o.warm_up_time) - ns.first_recv_ts;
                double observed_load = ((double)o.payload_sz * (double)ns.msg_recvd_count) / observed_load_time;
                observed_load = observed_load * (double)(1000 * 1000 * 1000);
                observed_load = observed_load / (double)(1024 * 1024 * 1024);
                double offered_load_time = ((double)ns.last_send_ts - o.warm_up_time) - ns.start_ts;
                double offered_load = ((double)o.payload_sz * (double)ns.warm_msg_sent_count) / offered_load_time;
                offered_load = offered_load * (double)(1000 * 1000 * 1000);
                offered_load = offered_load / (double)(1024 * 1024 * 1024);
                sum.total_offered_load += offered_load;
                sum.total_observed_load += observed_load;
                std::string rec;
                if (g == 0)
                    rec += "# Format <LP id> <Msgs Sent> <Msgs Recvd> <Bytes Sent> <Bytes Recvd> <Offered Load [GBps]> <Observed Load [GBps]> <End Time [ns]>\n";
                rec += fmt("%llu %d %d %d %d %f %f %f %f\n", (unsigned long long)g, ns.msg_sent_count, ns.msg_recvd_count,
                           o.payload_sz * ns.msg_sent_count, o.payload_sz * ns.msg_recvd_count, load * link_bandwidth, observed_load,
                           ns.last_recv_ts, observed_load_time);
                lp_io_write("synthetic-stats", rec);
                raw += fmt("S %d %d %d %d %d %a %a %a %a %a\n", g, ns.msg_sent_count, ns.msg_recvd_count, ns.local_recvd_count,
                           ns.msg_complete_count, ns.first_recv_ts, ns.last_recv_ts, ns.last_send_ts, ns.sum_server_latency,
                           ns.max_server_latency);
            } else if (is_term(g)) {  // dragonfly_dally_terminal_final dally:6791-6862
                TermState& s = term[term_rel(g)];
                sum.total_time += s.total_time;
                if (s.max_latency > sum.max_latency) sum.max_latency = s.max_latency;
                if (s.stats.used) write_mn_stats(g, "test", s.stats);  // model_net_print_stats model-net.c:205-226
                MnStats all = s.stats;
                if (!s.stats.used) all = MnStats();
                write_mn_stats(g, "all", all);
                std::string rec;
                if (s.terminal_id == 0)
                    rec += "# Format <source_id> <source_type> <dest_id> < dest_type>  <link_type> <link_traffic> <link_saturation> <stalled_chunks>\n";
                rec += fmt("\n%u %s %u %s %s %llu %lf %lu", s.terminal_id, "T", s.router_id, "R", "CN", (unsigned long long)s.link_traffic,
                           s.busy_time, s.stalled_chunks);
                lp_io_write("dragonfly-link-stats", rec);
                rec.clear();
                if (s.terminal_id == 0)
                    rec += "# Format <LP id> <Terminal ID> <Total Data Sent> <Total Data Received> <Avg packet latency> <Max packet Latency> <Min packet Latency> <# Packets finished> <Avg Hops> <Avg Busy Time (over rails)>\n";
                double avg_busy_time = 0;
                avg_busy_time += s.busy_time;
                avg_busy_time = avg_busy_time / 1;
                rec += fmt("%llu %u %d %llu %lf %lf %lf %ld %lf %lf\n", (unsigned long long)g, s.terminal_id, s.total_gen_size,
                           (unsigned long long)s.total_msg_size, s.total_time / s.finished_chunks, s.max_latency, s.min_latency,
                           s.finished_packets, (double)s.total_hops / s.finished_chunks, avg_busy_time);
                if (!s.terminal_msgs.empty()) fprintf(out, "[%llu] leftover terminal messages \n", (unsigned long long)g);
                lp_io_write("dragonfly-cn-stats", rec);
                raw += fmt("T %d %d %d %d %llu %ld %ld %ld %a %a %a %a %a %llu %lu %lu %lu %lu %a %a %ld %ld\n", g, s.packet_gen, s.packet_fin,
                           s.total_gen_size, (unsigned long long)s.total_msg_size, s.finished_msgs, s.finished_chunks, s.finished_packets,
                           s.total_time, s.total_hops, s.max_latency, s.min_latency, s.busy_time, (unsigned long long)s.link_traffic,
                           s.total_chunks, s.stalled_chunks, s.injected_chunks, s.ejected_chunks, s.stats.send_time, s.stats.recv_time,
                           s.stats.send_count, s.stats.recv_count);
            } else {  // dragonfly_dally_router_final dally:6974-7041
                RouterState& s = rtr[rtr_rel(g)];
                for (int i = 0; i < p.radix; i++)
                    for (int j = 0; j < p.num_vcs; j++) {
                        if (!s.queued_msgs[i * p.num_vcs + j].empty())
                            fprintf(out, "[%llu] leftover queued messages %d %d %d\n", (unsigned long long)g, i, j, s.vc_occupancy[i * p.num_vcs + j]);
                        if (!s.pending_msgs[i * p.num_vcs + j].empty())
                            fprintf(out, "[%llu] lefover pending messages %d %d\n", (unsigned long long)g, i, j);
                    }
                std::string rec;
                int src_rel_id = s.router_id % p.num_routers;
                int local_grp_id = s.router_id / p.num_routers;
                for (int d = 0; d <= p.intra_grp_radix; d++) {
                    if (d != src_rel_id) {
                        int dest_ab_id = local_grp_id * p.num_routers + d;
                        rec += fmt("\n%d %s %d %s %s %llu %lf %lu", s.router_id, "R", dest_ab_id, "R", "L", (unsigned long long)s.link_traffic[d],
                                   s.busy_time[d], s.stalled_chunks[d]);
                    }
                }
                for (auto& kv : topo.r[s.router_id].global)  // get_connections_by_type(CONN_GLOBAL)
                    for (const Conn& c : kv.second)
                        rec += fmt("\n%d %s %d %s %s %llu %lf %lu", s.router_id, "R", c.dest_gid, "R", "G", (unsigned long long)s.link_traffic[c.port],
                                   s.busy_time[c.port], s.stalled_chunks[c.port]);
                for (auto& kv : topo.r[s.router_id].term)
                    for (const Conn& c : kv.second)
                        rec += fmt("\n%d %s %d %s %s %llu %lf %lu", s.router_id, "R", c.dest_gid, "T", "CN", (unsigned long long)s.link_traffic[c.port],
                                   s.busy_time[c.port], s.stalled_chunks[c.port]);
                lp_io_write("dragonfly-link-stats", rec);
                for (int i = 0; i < p.radix; i++)
                    raw += fmt("R %d %d %lld %a %lu %lu %a\n", g, i, (long long)s.link_traffic[i], s.busy_time[i], s.stalled_chunks[i],
                               s.total_chunks[i], s.next_output_available_time[i]);
            }
        }
        for (int g = 0; g < n; g++) sum.rng_draws += rng0[g].draws + rng2[g].draws;
    }

    void report(FILE* f) {
        // ROSS-style marker line used by tests/equivalence-run.sh
        fprintf(f, "\tNet Events Processed                               %llu\n", (unsigned long long)sum.events);
        // dragonfly_dally_report_stats dally:3169-3190
        fprintf(f,
                "\nAverage number of hops traversed %f average chunk latency %lf us maximum chunk latency %lf us avg message size %lf bytes finished messages %lld finished chunks %lld\n",
                (float)sum.total_hops / sum.finished_chunks, (float)sum.total_time / sum.finished_chunks / 1000, sum.max_latency / 1000,
                (float)sum.total_msg_sz / sum.finished_msgs, sum.finished_msgs, sum.finished_chunks);
        fprintf(f, "\nADAPTIVE ROUTING STATS: %d chunks routed minimally %d chunks routed non-minimally completed packets %lld \n", sum.minimal_count,
                sum.nonmin_count, sum.finished_chunks);
        fprintf(f, "\nTotal packets generated %ld finished %ld Locally routed- same router %ld different-router %ld Remote (inter-group) %ld \n",
                sum.packet_gen, sum.packet_fin, sum.local_sr, sum.local_sg, sum.remote);
        // svr_report_stats synthetic:621-643
        double mean_latency = sum.svr_sum_latency / sum.svr_msgs_recvd;
        fprintf(f, "\nSynthetic Workload LP Stats: Mean Message Latency: %lf us,  Maximum Message Latency: %lf us, Total Messages Received: %lld\n",
                (float)mean_latency / 1000, (float)sum.svr_max_latency / 1000, sum.svr_msgs_recvd);
        fprintf(f, "\tMaximum Workload End Time %.2f\n", sum.max_end_ts);
        // aggregate_svr_stats synthetic:645-658
        int num_servers = num_nodes, spt = num_nodes / p.total_terminals;
        fprintf(f, "\nSynthetic-All Stats ---\n");
        fprintf(f, "AVG OFFERED LOAD = %.3f     |     AVG OBSERVED LOAD = %.3f\n", sum.total_offered_load / num_servers * spt,
                sum.total_observed_load / num_servers * spt);
    }

    void write_outputs() {
        if (o.lp_io_dir.empty()) return;
        mkdir(o.lp_io_dir.c_str(), 0775);
        for (const std::string& id : lpio_order) {
            std::string path = o.lp_io_dir + "/" + id;
            FILE* f = fopen(path.c_str(), "wb");
            if (!f) die("cannot write %s", path.c_str());
            fwrite(lpio[id].data(), 1, lpio[id].size(), f);
            fclose(f);
        }
        {
            std::string path = o.lp_io_dir + "/raw-lp-stats";
            FILE* f = fopen(path.c_str(), "wb");
            if (!f) die("cannot write %s", path.c_str());
            fwrite(raw.data(), 1, raw.size(), f);
            fclose(f);
        }
        std::string path = o.lp_io_dir + "/summary.txt";
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) die("cannot write %s", path.c_str());
        report(f);
        fclose(f);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// C entry points for the tests / bench (ctypes) and the CLI
// ------------------------------------------------------------------------------------------------
extern "C" {

struct dfdally_oracle_result {
    unsigned long long events;
    unsigned long long ev_by_type[16];
    long long total_hops, finished_packets, finished_msgs, finished_chunks, total_msg_sz;
    double total_time, max_latency;
    long long packet_gen, packet_fin, local_sr, local_sg, remote;
    long long minimal_count, nonmin_count;
    long long svr_msgs_recvd;
    double svr_sum_latency, svr_max_latency, max_end_ts;
    double last_event_ts;
    unsigned long long rng_draws;
    double run_seconds;  // event loop only (excludes config/table build)
    char error[512];
};

int dfdally_oracle_run(const char* conf_path, int traffic, int num_msgs, int payload_sz, double arrival_time, double load,
                       double warm_up_time, double end_time, const char* lp_io_dir, int quiet, dfdally_oracle_result* res) {
    memset(res, 0, sizeof *res);
    try {
        std::string cp(conf_path);
        std::string dir = cp.find('/') == std::string::npos ? "." : cp.substr(0, cp.rfind('/'));
        ConfNode root = load_conf(cp);
        FILE* log = quiet ? fopen("/dev/null", "w") : stderr;
        Sim sim;
        sim.p = read_params(root, dir, log);
        if (quiet && log) fclose(log);
        sim.topo = build_topo(sim.p);
        sim.o.traffic = traffic;
        sim.o.num_msgs = num_msgs;
        sim.o.payload_sz = payload_sz;
        sim.o.arrival_time = arrival_time;
        sim.o.load = load;
        sim.o.warm_up_time = warm_up_time;
        sim.o.end = end_time;
        sim.o.lp_io_dir = lp_io_dir ? lp_io_dir : "";
        sim.o.quiet = quiet;
        sim.init();
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        sim.run();
        clock_gettime(CLOCK_MONOTONIC, &t1);
        sim.sum.run_seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        sim.finalize();
        sim.write_outputs();
        if (!quiet) sim.report(stdout);
        const Summary& s = sim.sum;
        res->events = s.events;
        for (int i = 0; i < N_EV_TYPES; i++) res->ev_by_type[i] = s.ev_by_type[i];
        res->total_hops = s.total_hops;
        res->finished_packets = s.finished_packets;
        res->finished_msgs = s.finished_msgs;
        res->finished_chunks = s.finished_chunks;
        res->total_msg_sz = s.total_msg_sz;
        res->total_time = s.total_time;
        res->max_latency = s.max_latency;
        res->packet_gen = s.packet_gen;
        res->packet_fin = s.packet_fin;
        res->local_sr = s.local_sr;
        res->local_sg = s.local_sg;
        res->remote = s.remote;
        res->minimal_count = s.minimal_count;
        res->nonmin_count = s.nonmin_count;
        res->svr_msgs_recvd = s.svr_msgs_recvd;
        res->svr_sum_latency = s.svr_sum_latency;
        res->svr_max_latency = s.svr_max_latency;
        res->max_end_ts = s.max_end_ts;
        res->last_event_ts = s.last_event_ts;
        res->rng_draws = s.rng_draws;
        res->run_seconds = s.run_seconds;
        return 0;
    } catch (const std::exception& e) {
        snprintf(res->error, sizeof res->error, "%s", e.what());
        return 1;
    }
}

// CLCG4 known-answer hooks for tests
void dfdally_oracle_rng_stream(unsigned long long stream, int n, double* out_unif, long long* out_state) {
    Rng r;
    r.seed_stream(stream);
    if (out_state)
        for (int j = 0; j < 4; j++) out_state[j] = r.s[j];
    for (int i = 0; i < n; i++) out_unif[i] = r.unif();
}
double dfdally_oracle_bytes_to_ns(unsigned long long bytes, double gibps) { return bytes_to_ns(bytes, gibps); }
}

#ifdef DFDALLY_ORACLE_MAIN
// twin of: model-net-synthetic-dragonfly-all --sync=1 [--traffic=..] [--num_messages=..] ... -- <conf>
int main(int argc, char** argv) {
    int traffic = 1, num_msgs = 20, payload = 2048, quiet = 0;
    double arrival = 0, load = 0, warm = 0, end = 0;
    std::string lpio, conf;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* k) -> const char* {
            size_t n = strlen(k);
            return a.compare(0, n, k) == 0 ? a.c_str() + n : nullptr;
        };
        const char* v;
        if ((v = val("--traffic="))) traffic = atoi(v);
        else if ((v = val("--num_messages="))) num_msgs = atoi(v);
        else if ((v = val("--payload_sz="))) payload = atoi(v);
        else if ((v = val("--arrival_time="))) arrival = atof(v);
        else if ((v = val("--load_per_svr="))) load = atof(v);
        else if ((v = val("--warm_up_time="))) warm = atof(v);
        else if ((v = val("--end="))) end = atof(v);
        else if ((v = val("--lp-io-dir="))) lpio = v;
        else if (a == "--quiet") quiet = 1;
        else if ((v = val("--sync=")) || (v = val("--synch="))) { /* sequential only */ }
        else if (a == "--") continue;
        else if (a[0] != '-') conf = a;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (conf.empty()) {
        printf("\n Usage: dfdally_oracle <args> --sync=1 -- <config_file.conf> ");
        return 0;
    }
    dfdally_oracle_result res;
    int rc = dfdally_oracle_run(conf.c_str(), traffic, num_msgs, payload, arrival, load, warm, end, lpio.empty() ? nullptr : lpio.c_str(), quiet, &res);
    if (rc) {
        fprintf(stderr, "error: %s\n", res.error);
        return 1;
    }
    fprintf(stderr, "oracle: %llu events in %.3f s (%.3f Mev/s)\n", res.events, res.run_seconds, res.events / res.run_seconds / 1e6);
    return 0;
}
#endif
