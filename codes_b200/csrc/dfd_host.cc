// dfd_host.cc -- host side of the B200 dragonfly-dally engine: .conf reader, PARAMS handling, connection
// table construction (O(links), flattened for the device), statistics output, and the C-ABI declared in
// include/dfdally_b200.h.  Reference citations: "dally" = src/networks/model-net/dragonfly-dally.cxx,
// "synthetic" = src/network-workloads/model-net-synthetic-dragonfly-all.c, "netman" =
// src/networks/model-net/network-managers/dragonfly-network-manager.cxx.

#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dfdally_b200.h"
#include "dfd_engine.h"

namespace {

// ---------------------------------------------------------------------------------------------------
// configuration store: sections -> (sub)sections -> key -> values; grammar of src/modelconfig
// (configlex.l / configparser.y): NAME { key="v"; key=("a","b"); NAME { ... } }  with '#' comments
// ---------------------------------------------------------------------------------------------------
struct Section {
    std::vector<std::pair<std::string, std::vector<std::string>>> entries;  // file order
    std::vector<std::pair<std::string, Section>> children;
    const std::vector<std::string>* find(const std::string& k) const {
        for (const auto& e : entries)
            if (e.first == k) return &e.second;
        return nullptr;
    }
    const Section* child(const std::string& k) const {
        for (const auto& c : children)
            if (c.first == k) return &c.second;
        return nullptr;
    }
};

struct Parser {
    const char* p;
    const char* end;
    std::string err;
    void ws() {
        while (p < end) {
            if (*p == '#') {
                while (p < end && *p != '\n') ++p;
            } else if (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') {
                ++p;
            } else
                break;
        }
    }
    bool word(std::string* out) {
        ws();
        out->clear();
        if (p >= end) return false;
        if (*p == '"') {
            ++p;
            const char* q = p;
            while (q < end && *q != '"') ++q;
            if (q >= end) {
                err = "unterminated string";
                return false;
            }
            out->assign(p, q);
            p = q + 1;
            return true;
        }
        const char* q = p;
        while (q < end && !strchr(" \t\r\n{}=;(),\"#", *q)) ++q;
        if (q == p) return false;
        out->assign(p, q);
        p = q;
        return true;
    }
    bool take(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    bool body(Section* s, bool top) {
        for (;;) {
            ws();
            if (p >= end) {
                if (!top) err = "missing '}'";
                return top;
            }
            if (*p == '}') {
                if (top) {
                    err = "stray '}'";
                    return false;
                }
                ++p;
                return true;
            }
            std::string name;
            if (!word(&name)) {
                if (err.empty()) err = std::string("unexpected character '") + *p + "'";
                return false;
            }
            if (take('{')) {
                s->children.emplace_back(name, Section());
                if (!body(&s->children.back().second, false)) return false;
            } else if (take('=')) {
                std::vector<std::string> vals;
                if (take('(')) {
                    for (;;) {
                        if (take(')')) break;
                        if (take(',')) continue;
                        std::string v;
                        if (!word(&v)) {
                            err = "bad list for key " + name;
                            return false;
                        }
                        vals.push_back(v);
                    }
                } else {
                    std::string v;
                    if (!word(&v)) {
                        err = "missing value for key " + name;
                        return false;
                    }
                    vals.push_back(v);
                }
                if (!take(';')) {
                    err = "missing ';' after key " + name;
                    return false;
                }
                bool replaced = false;
                for (auto& e : s->entries)
                    if (e.first == name) {
                        e.second = vals;
                        replaced = true;
                    }
                if (!replaced) s->entries.emplace_back(name, vals);
            } else {
                err = "expected '{' or '=' after " + name;
                return false;
            }
        }
    }
};

// dragonfly_param (dally:391-442) + base LP + model-net + mapping parameters
struct HostParams {
    int num_routers = 0, num_groups = 0, num_cn = 0, num_global_channels = 0, num_parallel_switch_conns = 1;
    int local_vc_size = 0, global_vc_size = 0, cn_vc_size = 0, chunk_size = 0, packet_size = 0;
    double local_bandwidth = 0, global_bandwidth = 0, cn_bandwidth = 0, router_delay = 0;
    int adaptive_threshold = 0, global_k_picks = 2, routing = DFD_MINIMAL, scoring = DFD_SCORE_DELTA;
    int intra_grp_radix = 0, global_radix = 0, cn_radix = 0, radix = 0, total_routers = 0, total_terminals = 0;
    double cn_delay = 0, local_delay = 0, global_delay = 0;
    int credit_size = 8;
    double local_credit_delay = 0, global_credit_delay = 0, cn_credit_delay = 0;
    unsigned long long mn_packet_size = 0;
    double nic_seq_delay = 10, mn_cn_bandwidth = 20.0, svr_link_bandwidth = 4.7, end_time = 0;
    int node_eager_limit = 16000, codes_noop_bypass = 0, message_size = 256;
    std::string intra_file, inter_file;
    int repetitions = 0, n_svr = 0, n_term = 0, n_rtr = 0, off_svr = 0, off_term = 0, off_rtr = 0, lps_per_rep = 0;
    std::vector<std::string> modelnet_order;
    std::vector<double> snapshot_times;  // PARAMS router_buffer_snapshots (dally:2783-2797)
};

struct LpioFile {
    std::string name, data;
};

}  // namespace

struct dfd_engine {
    std::string error;
    std::string conf_path, conf_dir;
    Section root;
    bool loaded = false, registered = false, mapped = false, configured = false, opts_set = false;
    bool device_ready = false, uploaded = false, finished = false, downloaded = false;
    int rank = 0, world = 1;
    int ensemble = 1, dev_ensemble = 1;  // engines sharing the GPU (dfd_engine_set_ensemble) / what the device was sized for
    HostParams hp;
    dfd_synthetic_opts opts;
    double mean_interval = 0, load = 0;
    DfdHostTopo topo;
    // per-router lists kept for the *_final output order
    std::vector<std::vector<std::pair<int, int>>> glinks_by_gid;  // router -> (dest gid, port), by-type order
    DfdDevice dev;
    DfdHostResults res;
    dfd_summary sum;
    std::vector<LpioFile> lpio;
    bool slice_downloaded = false;  // e->res holds this rank's per-LP statistics
    bool gathered = false;          // ... and every other rank's (rank 0 of a sharded run after engine_finish)
    bool lpio_rendered = false;  // the lp-io text records are rendered on demand (render_lpio)
    std::string raw;
    unsigned long long h2d_bytes = 0;
};

namespace {

int fail(dfd_engine* e, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    e->error = buf;
    return 1;
}

// configuration_get_value_int/double: return non-zero when the key is missing
struct Getter {
    const Section* s;
    int geti(const char* k, int* out) const {
        const auto* v = s ? s->find(k) : nullptr;
        if (!v || v->empty()) return 1;
        *out = (int)strtol((*v)[0].c_str(), nullptr, 10);
        return 0;
    }
    int getd(const char* k, double* out) const {
        const auto* v = s ? s->find(k) : nullptr;
        if (!v || v->empty()) return 1;
        *out = strtod((*v)[0].c_str(), nullptr);
        return 0;
    }
    std::string gets(const char* k) const {
        const auto* v = s ? s->find(k) : nullptr;
        return (!v || v->empty()) ? std::string() : (*v)[0];
    }
    bool has(const char* k) const {
        const auto* v = s ? s->find(k) : nullptr;
        return v && !v->empty();
    }
};

double h_bytes_to_ns(unsigned long long bytes, double GB_p_s) {  // dally:1588-1599
    double time = ((double)bytes) / (1024.0 * 1024.0 * 1024.0);
    time = time / GB_p_s;
    time = time * 1000.0 * 1000.0 * 1000.0;
    return time;
}
double h_ns_to_GBps(double ns, unsigned long long bytes) {  // synthetic:224-228
    return (bytes / (1024.0 * 1024.0 * 1024.0)) / (ns / (1000.0 * 1000.0 * 1000.0));
}

std::string locate(const std::string& path, const std::string& dir) {
    struct stat st;
    if (stat(path.c_str(), &st) == 0) return path;
    if (!path.empty() && path[0] != '/') {
        std::string q = dir + "/" + path;
        if (stat(q.c_str(), &st) == 0) return q;
    }
    return path;
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------------
// dragonfly_read_config dally:2151-2757 (+ base_read_config model-net-lp.c:261-375, model_net_configure
// model-net.c:143-171, codes_mapping_setup codes_mapping.c:519-550)
// ---------------------------------------------------------------------------------------------------
int read_params(dfd_engine* e) {
    HostParams& p = e->hp;
    Getter g{e->root.child("PARAMS")};
    if (g.geti("local_vc_size", &p.local_vc_size)) p.local_vc_size = 1024;
    if (g.geti("global_vc_size", &p.global_vc_size)) p.global_vc_size = 2048;
    if (g.geti("cn_vc_size", &p.cn_vc_size)) p.cn_vc_size = 1024;
    if (g.geti("chunk_size", &p.chunk_size)) p.chunk_size = 512;
    if (g.geti("packet_size", &p.packet_size)) p.chunk_size = 512;  // sic dally:2197-2203
    if (g.getd("local_bandwidth", &p.local_bandwidth)) p.local_bandwidth = 5.25;
    if (g.getd("global_bandwidth", &p.global_bandwidth)) p.global_bandwidth = 4.7;
    if (g.getd("cn_bandwidth", &p.cn_bandwidth)) p.cn_bandwidth = 5.25;
    if (g.getd("router_delay", &p.router_delay)) p.router_delay = 100;
    int v = 1;
    if (!g.geti("num_qos_levels", &v) && v != 1) return fail(e, "\n priority needs to be specified with qos_levels>1 (the synthetic driver sends category \"test\": the reference aborts here too, dally:3248-3254)");
    if (g.geti("adaptive_threshold", &p.adaptive_threshold)) p.adaptive_threshold = 0;
    std::string rt = g.gets("routing");
    if (rt == "minimal")
        p.routing = DFD_MINIMAL;
    else if (rt == "nonminimal" || rt == "non-minimal")
        p.routing = DFD_NON_MINIMAL;
    else if (rt == "prog-adaptive")
        p.routing = DFD_PROG_ADAPTIVE;
    else if (rt == "adaptive")
        return fail(e, "Standard Adaptive Routing not implemented - try Progressive Adaptive instead");
    else if (rt == "prog-adaptive-legacy" || rt == "smart-prog-adaptive" || rt == "smart-minimal" || rt == "smart-nonminimal")
        return fail(e, "routing \"%s\" is not supported by this engine", rt.c_str());
    else
        p.routing = DFD_MINIMAL;
    v = 1;
    if (!g.geti("num_injection_queues", &v) && v != 1) return fail(e, "num_injection_queues > 1 is not supported by this engine");
    v = 1;
    if (!g.geti("num_rails", &v) && v != 1) return fail(e, "num_rails > 1 is not supported by this engine");
    v = 1;
    if (!g.geti("num_planes", &v) && v != 1) return fail(e, "num_planes > 1 is not supported by this engine");
    if (g.geti("global_k_picks", &p.global_k_picks)) p.global_k_picks = 2;
    std::string sc = g.gets("route_scoring_metric");
    if (sc == "alpha")
        p.scoring = DFD_SCORE_ALPHA;
    else if (sc == "beta")
        return fail(e, "Beta scoring not implemented");
    else if (sc == "gamma")
        return fail(e, "Gamma scoring protocol currently non-functional");
    else
        p.scoring = DFD_SCORE_DELTA;
    if (p.routing == DFD_PROG_ADAPTIVE && p.global_k_picks < 1) return fail(e, "global_k_picks must be at least 1");
    if (g.geti("num_groups", &p.num_groups)) return fail(e, "\nnum_groups not specified, Aborting\n");
    if (g.geti("num_routers", &p.num_routers)) return fail(e, "\nnum_routers not specified, Aborting\n");
    if (g.geti("num_cns_per_router", &p.num_cn)) p.num_cn = p.num_routers / 2;
    if (g.geti("num_global_channels", &p.num_global_channels)) p.num_global_channels = 10;
    if (g.geti("num_parallel_switch_conns", &p.num_parallel_switch_conns)) p.num_parallel_switch_conns = 1;
    p.intra_grp_radix = (p.num_routers - 1) * p.num_parallel_switch_conns;
    p.cn_radix = p.num_cn;
    p.global_radix = p.num_global_channels * p.num_parallel_switch_conns;
    p.radix = p.intra_grp_radix + p.global_radix + p.cn_radix;
    p.total_routers = p.num_groups * p.num_routers;
    p.total_terminals = p.total_routers * p.num_cn;
    if (!g.has("intra-group-connections")) return fail(e, "Intra group connections file not specified. Aborting");
    if (!g.has("inter-group-connections")) return fail(e, "Inter group connections file not specified. Aborting");
    p.intra_file = locate(g.gets("intra-group-connections"), e->conf_dir);
    p.inter_file = locate(g.gets("inter-group-connections"), e->conf_dir);
    if (g.has("link-failure-file")) return fail(e, "link failures are not supported by this engine");
    v = 0;
    if (!g.geti("congestion_control_enabled", &v) && v) return fail(e, "congestion control is not supported by this engine");
    if (g.getd("cn_delay", &p.cn_delay)) p.cn_delay = h_bytes_to_ns(p.chunk_size, p.cn_bandwidth * 1);
    if (g.getd("local_delay", &p.local_delay)) p.local_delay = h_bytes_to_ns(p.chunk_size, p.local_bandwidth);
    if (g.getd("global_delay", &p.global_delay)) p.global_delay = h_bytes_to_ns(p.chunk_size, p.global_bandwidth);
    if (g.geti("credit_size", &p.credit_size)) p.credit_size = 8;
    double general = 0;
    int cu = g.getd("credit_delay", &general), lu = g.getd("local_credit_delay", &p.local_credit_delay);
    int gu = g.getd("global_credit_delay", &p.global_credit_delay), nu = g.getd("cn_credit_delay", &p.cn_credit_delay);
    int autoflag = 0;
    if (g.geti("auto_credit_delay", &autoflag)) autoflag = 0;
    if (!cu && !(lu || gu || nu))
        return fail(e, "\nCannot set both a general credit delay and specific (local/global/cn) credit delays. Check configuration file.");
    if ((!cu || !lu || !gu || !nu) && autoflag)
        return fail(e, "\nCannot set both a credit delay (general or specific) and also enable auto credit delay calculation. Check Configuration file.");
    if (cu && lu && gu && nu && !autoflag) {
        p.local_credit_delay = h_bytes_to_ns(p.credit_size, p.local_bandwidth);
        p.global_credit_delay = p.local_credit_delay;
        p.cn_credit_delay = p.local_credit_delay;
    } else if (cu) {
        if (lu) p.local_credit_delay = h_bytes_to_ns(p.credit_size, p.local_bandwidth);
        if (gu) p.global_credit_delay = h_bytes_to_ns(p.credit_size, p.global_bandwidth);
        if (nu) p.cn_credit_delay = h_bytes_to_ns(p.credit_size, p.cn_bandwidth);
    } else {
        p.local_credit_delay = p.global_credit_delay = p.cn_credit_delay = general;
    }
    std::string sched = g.gets("modelnet_scheduler");
    if (!sched.empty() && sched != "fcfs") return fail(e, "modelnet_scheduler \"%s\" is not supported by this engine (fcfs only)", sched.c_str());
    long ps = g.has("packet_size") ? strtol(g.gets("packet_size").c_str(), nullptr, 10) : 0;
    p.mn_packet_size = ps ? (unsigned long long)ps : 512;
    if (g.getd("nic_seq_delay", &p.nic_seq_delay)) p.nic_seq_delay = 10;
    v = 1;
    if (!g.geti("node_copy_queues", &v) && v != 1) return fail(e, "node_copy_queues > 1 is not supported by this engine");
    if (g.getd("cn_bandwidth", &p.mn_cn_bandwidth)) p.mn_cn_bandwidth = 20.0;
    if (g.getd("cn_bandwidth", &p.svr_link_bandwidth) || !p.svr_link_bandwidth) p.svr_link_bandwidth = 4.7;  // synthetic:236-240
    if (g.geti("node_eager_limit", &p.node_eager_limit)) p.node_eager_limit = 16000;
    v = 0;
    p.codes_noop_bypass = g.geti("codes_noop_bypass", &v) ? 0 : v;
    if (g.geti("message_size", &p.message_size) || !p.message_size) p.message_size = 256;
    if (g.getd("end_time", &p.end_time)) p.end_time = 0;
    const auto* order = g.s ? g.s->find("modelnet_order") : nullptr;
    if (!order) return fail(e, "unable to read PARAMS:modelnet_order variable\n");
    p.modelnet_order = *order;
    // router VC-occupancy snapshots: a list of timestamps, each read with strtod (dally:2783-2797)
    p.snapshot_times.clear();
    if (const auto* snaps = g.s ? g.s->find("router_buffer_snapshots") : nullptr)
        for (const std::string& t : *snaps) p.snapshot_times.push_back(strtod(t.c_str(), nullptr));
    if (p.snapshot_times.size() > 4096) return fail(e, "router_buffer_snapshots: more than 4096 snapshots are not supported by this engine");
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// connection tables (netman:91-175 add_link, :457-500, :672-741 port numbering, :772-814, :1134-1178
// ordering of the returned vectors), built in O(links) and flattened
// ---------------------------------------------------------------------------------------------------
int build_topology(dfd_engine* e) {
    const HostParams& p = e->hp;
    DfdHostTopo& T = e->topo;
    const int a = p.num_routers, G = p.num_groups, h = p.global_radix, NR = p.total_routers;
    // --- intra file: one group's worth of {src, dest, type}; port = order of appearance per source
    FILE* f = fopen(p.intra_file.c_str(), "rb");
    if (!f) return fail(e, "intra-group file not found: %s", p.intra_file.c_str());
    std::vector<std::vector<std::vector<int>>> loc(a, std::vector<std::vector<int>>(a));  // [src lid][dest lid] -> ports
    std::vector<int> used(a, 0);
    T.loc_dest.assign((size_t)a * std::max(1, p.intra_grp_radix), -1);
    int rec3[3];
    while (fread(rec3, sizeof rec3, 1, f) == 1) {
        int s = rec3[0], d = rec3[1];
        if (s < 0 || s >= a || d < 0 || d >= a) {
            fclose(f);
            return fail(e, "intra-group file: router id out of range");
        }
        if (used[s] >= p.intra_grp_radix) {
            fclose(f);
            return fail(e, "Attempting to add too many local connections per router - exceeding configuration value: %d", p.intra_grp_radix);
        }
        int port = used[s]++;
        loc[s][d].push_back(port);
        T.loc_dest[(size_t)s * p.intra_grp_radix + port] = d;
    }
    fclose(f);
    T.loc_off.assign((size_t)a * a + 1, 0);
    T.loc_port.clear();
    for (int s = 0; s < a; s++)
        for (int d = 0; d < a; d++) {
            T.loc_off[(size_t)s * a + d] = (int)T.loc_port.size();
            for (int port : loc[s][d]) T.loc_port.push_back(port);
        }
    T.loc_off[(size_t)a * a] = (int)T.loc_port.size();
    // --- inter file: {src gid, dest gid}; global port = intra_radix + order of appearance per source
    f = fopen(p.inter_file.c_str(), "rb");
    if (!f) return fail(e, "inter-group file not found: %s", p.inter_file.c_str());
    std::vector<int> gused(NR, 0);
    T.gdest.assign((size_t)NR * std::max(1, h), -1);
    std::vector<std::vector<int>> routed((size_t)G * G);  // (src group, dest group) -> owner lids, file order
    int rec2[2];
    while (fread(rec2, sizeof rec2, 1, f) == 1) {
        int s = rec2[0], d = rec2[1];
        if (s < 0 || s >= NR || d < 0 || d >= NR) {
            fclose(f);
            return fail(e, "inter-group file: router id out of range");
        }
        if (gused[s] >= h) {
            fclose(f);
            return fail(e, "Attempting to add too many global connections per router - exceeding configuration value: %d", h);
        }
        T.gdest[(size_t)s * h + gused[s]++] = d;
        routed[(size_t)(s / a) * G + d / a].push_back(s % a);
    }
    fclose(f);
    T.routed_off.assign((size_t)G * G + 1, 0);
    T.routed_lid.clear();
    for (size_t i = 0; i < (size_t)G * G; i++) {
        T.routed_off[i] = (int)T.routed_lid.size();
        T.routed_lid.insert(T.routed_lid.end(), routed[i].begin(), routed[i].end());
    }
    T.routed_off[(size_t)G * G] = (int)T.routed_lid.size();
    // --- by-type order of the global connections: std::map keyed by dest gid, file order inside a key
    T.gs_port.assign((size_t)NR * std::max(1, h), -1);
    T.gs_group.assign((size_t)NR * std::max(1, h), INT_MAX);
    e->glinks_by_gid.assign(NR, {});
    for (int r = 0; r < NR; r++) {
        std::vector<std::pair<int, int>> v;  // (dest gid, port)
        for (int i = 0; i < gused[r]; i++) v.emplace_back(T.gdest[(size_t)r * h + i], p.intra_grp_radix + i);
        std::stable_sort(v.begin(), v.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; });
        for (size_t i = 0; i < v.size(); i++) {
            T.gs_port[(size_t)r * h + i] = v[i].second;
            T.gs_group[(size_t)r * h + i] = v[i].first / a;
        }
        e->glinks_by_gid[r] = v;
    }
    // --- lanes: every directed link feeds one in-lane of its destination.  Local in-lanes of router-local id d are
    // numbered in (source lid, source port) order, global in-lanes of router d in (source router, source port) order;
    // a lane has the same index range as the ports of its kind, so capacities follow from the index alone.
    T.lout_lane.assign((size_t)a * std::max(1, p.intra_grp_radix), -1);
    T.lin_src.assign((size_t)a * std::max(1, p.intra_grp_radix), -1);
    {
        std::vector<int> nin(a, 0);
        for (int sl = 0; sl < a; sl++)
            for (int port = 0; port < p.intra_grp_radix; port++) {
                int d = T.loc_dest[(size_t)sl * p.intra_grp_radix + port];
                if (d < 0) continue;
                if (nin[d] >= p.intra_grp_radix)
                    return fail(e, "local router %d is the destination of more than %d local links (> intra-group radix)", d, p.intra_grp_radix);
                int lane = nin[d]++;
                T.lout_lane[(size_t)sl * p.intra_grp_radix + port] = lane;
                T.lin_src[(size_t)d * p.intra_grp_radix + lane] = (sl << 16) | port;
            }
    }
    T.gout_lane.assign((size_t)NR * std::max(1, h), -1);
    T.gin_src.assign((size_t)NR * std::max(1, h), -1);
    T.gin_port.assign((size_t)NR * std::max(1, h), -1);
    {
        std::vector<int> nin(NR, 0);
        for (int r = 0; r < NR; r++)
            for (int i = 0; i < h; i++) {
                int d = T.gdest[(size_t)r * h + i];
                if (d < 0) continue;
                if (nin[d] >= h)
                    return fail(e, "router %d is the destination of more than %d global links (> num_global_channels): asymmetric link files are not supported", d, h);
                int k = nin[d]++;
                T.gout_lane[(size_t)r * h + i] = p.intra_grp_radix + k;
                T.gin_src[(size_t)d * h + k] = r;
                T.gin_port[(size_t)d * h + k] = p.intra_grp_radix + i;
            }
    }
    return 0;
}

void lpio_write(dfd_engine* e, const std::string& id, const std::string& data) {  // lp-io.c:39-100
    for (auto& f : e->lpio)
        if (f.name == id) {
            f.data += data;
            return;
        }
    e->lpio.push_back({id, data});
}
std::string sfmt(const char* fmtstr, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmtstr);
    vsnprintf(buf, sizeof buf, fmtstr, ap);
    va_end(ap);
    return buf;
}

unsigned h_svr_gid(const HostParams& p, unsigned k) { return (k / p.n_svr) * p.lps_per_rep + p.off_svr + k % p.n_svr; }
unsigned h_term_gid(const HostParams& p, unsigned k) { return (k / p.n_term) * p.lps_per_rep + p.off_term + k % p.n_term; }
unsigned h_rtr_gid(const HostParams& p, unsigned r) { return r * p.lps_per_rep + p.off_rtr; }

double bits_to_double(unsigned long long b) {
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
}

// The *_final handlers: svr_finalize synthetic:523-577, dragonfly_dally_terminal_final dally:6791-6862 (+ model_net_print_stats
// model-net.c:186-226), dragonfly_dally_router_final dally:6974-7041.  They do two things in the reference: fold the LP's
// counters into the run totals and lp_io_write() a text record.  Here the totals are folded when the run finishes
// (finalize, part of tw_run) and the text records are rendered when lp_io_flush asks for them (render_lpio): a caller
// that only reads the summary never pays for 1.2 M snprintf calls on a 131,584-terminal network.
struct SvrLoad {
    double first_recv_ts, last_recv_ts, observed_load, observed_load_time, offered_load;
};
static SvrLoad svr_load(const dfd_engine* e, unsigned k) {  // synthetic:533-560
    const dfd_synthetic_opts& o = e->opts;
    const unsigned long long INF = 0x7FF0000000000000ULL;
    const NodeState& ns = e->res.node[k];
    SvrLoad L;
    L.first_recv_ts = e->res.ack_first[k] == INF ? -1.0 : bits_to_double(e->res.ack_first[k]);
    L.last_recv_ts = bits_to_double(e->res.ack_last[k]);
    double start_ts = e->mean_interval;
    L.observed_load_time = ((double)L.last_recv_ts - o.warm_up_time) - L.first_recv_ts;
    double observed_load = ((double)o.payload_sz * (double)ns.msg_recvd_count) / L.observed_load_time;
    observed_load = observed_load * (double)(1000 * 1000 * 1000);
    L.observed_load = observed_load / (double)(1024 * 1024 * 1024);
    double offered_load_time = ((double)ns.last_send_ts - o.warm_up_time) - start_ts;
    double offered_load = ((double)o.payload_sz * (double)ns.warm_msg_sent_count) / offered_load_time;
    offered_load = offered_load * (double)(1000 * 1000 * 1000);
    L.offered_load = offered_load / (double)(1024 * 1024 * 1024);
    return L;
}

void finalize(dfd_engine* e) {
    const HostParams& p = e->hp;
    dfd_summary& S = e->sum;
    e->lpio.clear();
    e->lpio_rendered = false;
    e->raw.clear();  // built on demand (build_raw)
    const unsigned long long* C = e->res.counters;
    S.events = C[0];
    S.windows = C[1];
    for (int i = 0; i < EV_NTYPES && i < 16; i++) S.events_by_type[i] = C[4 + i];
    // the same accumulation order as the reference's final handlers (gid order: floating-point sums depend on it)
    for (int rep = 0; rep < p.repetitions; rep++) {
        for (int i = 0; i < p.n_svr; i++) {
            unsigned k = rep * p.n_svr + i;
            const NodeState& ns = e->res.node[k];
            SvrLoad L = svr_load(e, k);
            if (L.last_recv_ts > S.max_end_ts) S.max_end_ts = L.last_recv_ts;  // pe_max_end_ts synthetic:427-430
            S.svr_sum_latency += ns.sum_server_latency;
            S.svr_msgs_recvd += ns.msg_recvd_count;
            if (ns.max_server_latency > S.svr_max_latency) S.svr_max_latency = ns.max_server_latency;
            S.total_offered_load += L.offered_load;
            S.total_observed_load += L.observed_load;
        }
        for (int i = 0; i < p.n_term; i++) {
            unsigned k = rep * p.n_term + i;
            const NodeState& s = e->res.node[k];
            S.total_time += s.total_time;
            if (s.max_latency > S.max_latency) S.max_latency = s.max_latency;
            S.total_hops += (long long)s.total_hops;
            S.finished_packets += s.finished_packets;
            S.finished_msgs += s.finished_msgs;
            S.finished_chunks += s.finished_chunks;
            S.total_msg_sz += (long long)s.total_msg_size;
            S.packet_gen += s.packet_gen;
            S.packet_fin += s.packet_fin;
            S.local_same_router += s.local_sr;
            S.local_same_group += s.local_sg;
            S.remote += s.remote_pk;
            S.minimal_count += s.minimal_count;
            S.nonmin_count += s.nonmin_count;
            S.rng_draws += s.rng_draws;
        }
        S.rng_draws += e->res.rh[rep].rng_draws;
    }
}

// the lp-io text records of the final handlers, in gid order
void render_lpio(dfd_engine* e) {
    if (e->lpio_rendered) return;
    const HostParams& p = e->hp;
    const dfd_synthetic_opts& o = e->opts;
    e->lpio.clear();
    if (!p.snapshot_times.empty()) {
        // dragonfly-snapshots.csv: router 0 writes the header at init (dally:4364-4403); every R_SNAPSHOT writes its record
        // when it commits (dally:4687-4709), i.e. in the order of the sequential run: by (time, router gid)
        std::string text = "#Time of snapshot,Router ID,";
        for (int i = 0; i < p.radix; i++)
            for (int j = 0; j < DFD_NUM_VCS; j++) text += sfmt("Port %d VC %d,", i, j);
        text[text.size() - 1] = '\n';
        const size_t per_router = p.snapshot_times.size() * (size_t)p.radix;
        const std::vector<double>& sts = e->topo.snap_ts;
        for (size_t k0 = 0; k0 < sts.size();) {  // runs of equal timestamps: (time, router gid, configured index) order
            size_t k1 = k0;
            while (k1 < sts.size() && sts[k1] == sts[k0]) k1++;
            for (int r = 0; r < p.total_routers; r++)
                for (size_t k = k0; k < k1; k++) {
                    const unsigned id = e->topo.snap_id[k];
                    const int4* occ = e->res.snap.data() + (size_t)r * per_router + (size_t)id * p.radix;
                    text += sfmt("%.4f, %d, ", p.snapshot_times[id], r);
                    for (int i = 0; i < p.radix; i++) text += sfmt("%d, %d, %d, %d, ", occ[i].x, occ[i].y, occ[i].z, occ[i].w);
                    text.resize(text.size() - 2);
                    text += '\n';
                }
            k0 = k1;
        }
        lpio_write(e, "dragonfly-snapshots.csv", text);
    }
    for (int rep = 0; rep < p.repetitions; rep++) {
        for (int off = 0; off < p.lps_per_rep; off++) {
            unsigned g = (unsigned)rep * p.lps_per_rep + off;
            if (off >= p.off_svr && off < p.off_svr + p.n_svr) {
                unsigned k = rep * p.n_svr + (off - p.off_svr);
                const NodeState& ns = e->res.node[k];
                SvrLoad L = svr_load(e, k);
                std::string rec;
                if (g == 0)
                    rec += "# Format <LP id> <Msgs Sent> <Msgs Recvd> <Bytes Sent> <Bytes Recvd> <Offered Load [GBps]> <Observed Load [GBps]> <End Time [ns]>\n";
                rec += sfmt("%llu %d %d %d %d %f %f %f %f\n", (unsigned long long)g, ns.msg_sent_count, ns.msg_recvd_count,
                            o.payload_sz * ns.msg_sent_count, o.payload_sz * ns.msg_recvd_count, e->load * p.svr_link_bandwidth, L.observed_load,
                            L.last_recv_ts, L.observed_load_time);
                lpio_write(e, "synthetic-stats", rec);
            } else if (off >= p.off_term && off < p.off_term + p.n_term) {
                unsigned k = rep * p.n_term + (off - p.off_term);
                const NodeState& s = e->res.node[k];
                unsigned router_id = k / p.num_cn;
                const char* line = "lp:%ld\tsend_count:%ld\tsend_bytes:%ld\tsend_time:%f\trecv_count:%ld\trecv_bytes:%ld\trecv_time:%f\tmax_event_size:%ld\n";
                if (s.stats_used)
                    lpio_write(e, "model-net-category-test", sfmt(line, (long)g, (long)s.send_count, (long)s.send_bytes, s.send_time,
                                                                  (long)s.recv_count, (long)s.recv_bytes, s.recv_time, (long)s.max_event_size));
                if (s.stats_used)
                    lpio_write(e, "model-net-category-all", sfmt(line, (long)g, (long)s.send_count, (long)s.send_bytes, s.send_time,
                                                                 (long)s.recv_count, (long)s.recv_bytes, s.recv_time, (long)s.max_event_size));
                else
                    lpio_write(e, "model-net-category-all", sfmt(line, (long)g, 0L, 0L, 0.0, 0L, 0L, 0.0, 0L));
                std::string rec;
                if (k == 0)
                    rec += "# Format <source_id> <source_type> <dest_id> < dest_type>  <link_type> <link_traffic> <link_saturation> <stalled_chunks>\n";
                rec += sfmt("\n%u %s %u %s %s %llu %lf %lu", k, "T", router_id, "R", "CN", (unsigned long long)s.link_traffic, s.busy_time,
                            (unsigned long)s.stalled_chunks);
                lpio_write(e, "dragonfly-link-stats", rec);
                rec.clear();
                if (k == 0)
                    rec += "# Format <LP id> <Terminal ID> <Total Data Sent> <Total Data Received> <Avg packet latency> <Max packet Latency> <Min packet Latency> <# Packets finished> <Avg Hops> <Avg Busy Time (over rails)>\n";
                double avg_busy_time = 0;
                avg_busy_time += s.busy_time;
                avg_busy_time = avg_busy_time / 1;
                rec += sfmt("%llu %u %d %llu %lf %lf %lf %ld %lf %lf\n", (unsigned long long)g, k, s.total_gen_size, (unsigned long long)s.total_msg_size,
                            s.total_time / s.finished_chunks, s.max_latency, s.min_latency, (long)s.finished_packets,
                            (double)s.total_hops / s.finished_chunks, avg_busy_time);
                lpio_write(e, "dragonfly-cn-stats", rec);
            } else {
                unsigned r = rep;
                const PortState* ps = &e->res.port[(size_t)r * p.radix];
                std::string rec;
                int src_rel_id = r % p.num_routers, local_grp_id = r / p.num_routers;
                for (int d = 0; d <= p.intra_grp_radix; d++) {  // indexes ports by router-local id (sic, dally:7010-7017)
                    if (d != src_rel_id) {
                        int dest_ab_id = local_grp_id * p.num_routers + d;
                        rec += sfmt("\n%d %s %d %s %s %llu %lf %lu", (int)r, "R", dest_ab_id, "R", "L", (unsigned long long)ps[d].link_traffic,
                                    ps[d].busy_time, (unsigned long)ps[d].stalled_chunks);
                    }
                }
                for (const auto& gl : e->glinks_by_gid[r])
                    rec += sfmt("\n%d %s %d %s %s %llu %lf %lu", (int)r, "R", gl.first, "R", "G", (unsigned long long)ps[gl.second].link_traffic,
                                ps[gl.second].busy_time, (unsigned long)ps[gl.second].stalled_chunks);
                for (int t = 0; t < p.num_cn; t++) {
                    int port = p.intra_grp_radix + p.global_radix + t;
                    rec += sfmt("\n%d %s %d %s %s %llu %lf %lu", (int)r, "R", (int)(r * p.num_cn + t), "T", "CN",
                                (unsigned long long)ps[port].link_traffic, ps[port].busy_time, (unsigned long)ps[port].stalled_chunks);
                }
                lpio_write(e, "dragonfly-link-stats", rec);
            }
        }
    }
    e->lpio_rendered = true;
}

// raw-lp-stats: the same per-LP statistics as C99 hex floats, for bit-exact comparison in the tests.  Built only
// when lp_io_flush asks for it: it is not part of the reference's output and costs a third of the final handlers.
void build_raw(dfd_engine* e) {
    if (!e->raw.empty()) return;
    const HostParams& p = e->hp;
    const unsigned long long INF = 0x7FF0000000000000ULL;
    for (int rep = 0; rep < p.repetitions; rep++) {
        for (int off = 0; off < p.lps_per_rep; off++) {
            unsigned g = (unsigned)rep * p.lps_per_rep + off;
            if (off >= p.off_svr && off < p.off_svr + p.n_svr) {
                unsigned k = rep * p.n_svr + (off - p.off_svr);
                const NodeState& ns = e->res.node[k];
                double first_recv_ts = e->res.ack_first[k] == INF ? -1.0 : bits_to_double(e->res.ack_first[k]);
                double last_recv_ts = bits_to_double(e->res.ack_last[k]);
                e->raw += sfmt("S %d %d %d %d %d %a %a %a %a %a\n", (int)g, ns.msg_sent_count, ns.msg_recvd_count, ns.local_recvd_count,
                               (int)e->res.ack_count[k], first_recv_ts, last_recv_ts, ns.last_send_ts, ns.sum_server_latency, ns.max_server_latency);
            } else if (off >= p.off_term && off < p.off_term + p.n_term) {
                unsigned k = rep * p.n_term + (off - p.off_term);
                const NodeState& s = e->res.node[k];
                e->raw += sfmt("T %d %d %d %d %llu %ld %ld %ld %a %a %a %a %a %llu %lu %lu %lu %lu %a %a %ld %ld\n", (int)g, s.packet_gen, s.packet_fin,
                               s.total_gen_size, (unsigned long long)s.total_msg_size, (long)s.finished_msgs, (long)s.finished_chunks,
                               (long)s.finished_packets, s.total_time, s.total_hops, s.max_latency, s.min_latency, s.busy_time,
                               (unsigned long long)s.link_traffic, (unsigned long)s.total_chunks, (unsigned long)s.stalled_chunks,
                               (unsigned long)s.injected_chunks, (unsigned long)s.ejected_chunks, s.send_time, s.recv_time, (long)s.send_count,
                               (long)s.recv_count);
            } else {
                unsigned r = rep;
                const PortState* ps = &e->res.port[(size_t)r * p.radix];
                for (int i = 0; i < p.radix; i++)
                    e->raw += sfmt("R %d %d %lld %a %lu %lu %a\n", (int)g, i, (long long)ps[i].link_traffic, ps[i].busy_time,
                                   (unsigned long)ps[i].stalled_chunks, (unsigned long)ps[i].total_chunks, ps[i].noat);
            }
        }
    }
}

std::string report_text(const dfd_engine* e) {
    const dfd_summary& s = e->sum;
    const HostParams& p = e->hp;
    std::string out;
    out += sfmt("\tNet Events Processed                               %llu\n", s.events);
    out += sfmt("\nAverage number of hops traversed %f average chunk latency %lf us maximum chunk latency %lf us avg message size %lf bytes finished messages %lld finished chunks %lld\n",
                (float)s.total_hops / s.finished_chunks, (float)s.total_time / s.finished_chunks / 1000, s.max_latency / 1000,
                (float)s.total_msg_sz / s.finished_msgs, s.finished_msgs, s.finished_chunks);  // dally:3169-3177
    out += sfmt("\nADAPTIVE ROUTING STATS: %d chunks routed minimally %d chunks routed non-minimally completed packets %lld \n",
                (int)s.minimal_count, (int)s.nonmin_count, s.finished_chunks);
    out += sfmt("\nTotal packets generated %ld finished %ld Locally routed- same router %ld different-router %ld Remote (inter-group) %ld \n",
                (long)s.packet_gen, (long)s.packet_fin, (long)s.local_same_router, (long)s.local_same_group, (long)s.remote);
    double mean_latency = s.svr_sum_latency / s.svr_msgs_recvd;  // synthetic:633-641
    out += sfmt("\nSynthetic Workload LP Stats: Mean Message Latency: %lf us,  Maximum Message Latency: %lf us, Total Messages Received: %lld\n",
                (float)mean_latency / 1000, (float)s.svr_max_latency / 1000, s.svr_msgs_recvd);
    out += sfmt("\tMaximum Workload End Time %.2f\n", s.max_end_ts);
    int num_servers = p.repetitions * p.n_svr, spt = num_servers / p.total_terminals;  // synthetic:645-657
    out += "\nSynthetic-All Stats ---\n";
    out += sfmt("AVG OFFERED LOAD = %.3f     |     AVG OBSERVED LOAD = %.3f\n", s.total_offered_load / num_servers * spt,
                s.total_observed_load / num_servers * spt);
    return out;
}

const char* device_error_name(unsigned long long code) {
    switch (code) {
    case DFD_ERR_NEV_OVERFLOW: return "node self-event list overflow";
    case DFD_ERR_SQ_OVERFLOW: return "model-net scheduler queue overflow";
    case DFD_ERR_TQ_OVERFLOW: return "terminal_msgs queue overflow";
    case DFD_ERR_POOL_OVERFLOW: return "router chunk pool overflow";
    case DFD_ERR_HEAP_OVERFLOW: return "router event heap overflow";
    case DFD_ERR_INBOX_OVERFLOW: return "inbox overflow";
    case DFD_ERR_DEAD_END: return "routing dead end (no legal next stop)";
    case DFD_ERR_WRONG_TERMINAL: return "Packet arrived at wrong terminal";
    case DFD_ERR_BAD_VC: return "Output channel greater than available VCs";
    case DFD_ERR_SELF_SEND: return "self-addressed network message above the eager limit";
    case DFD_ERR_LATE_EVENT: return "causality violation: an LP received an event older than one it had already executed";
    case DFD_ERR_PEER_TIMEOUT: return "a peer GPU stopped responding";
    case DFD_ERR_STALL: return "no event became safe for the watchdog period although events are pending (info = pending events)";
    case DFD_ERR_PEER_ERROR: return "a peer GPU reported an error";
    case DFD_ERR_REASM_OVERFLOW: return "too many messages being reassembled at one terminal (raise DFD_REASM_CAP)";
    }
    return "unknown";
}

}  // namespace

// =====================================================================================================
// C-ABI
// =====================================================================================================
extern "C" {

void dfd_synthetic_opts_default(dfd_synthetic_opts* o) {
    memset(o, 0, sizeof *o);
    o->traffic = 1;
    o->num_messages = 20;
    o->payload_sz = 2048;
    o->rperm_threshold = 131072;
}

int dfd_engine_create(dfd_engine** out) {
    *out = new dfd_engine();
    dfd_synthetic_opts_default(&(*out)->opts);
    memset(&(*out)->sum, 0, sizeof(dfd_summary));
    return 0;
}
void dfd_engine_destroy(dfd_engine* e) {
    if (!e) return;
    if (e->device_ready) dfd_device_destroy(&e->dev);
    delete e;
}
const char* dfd_last_error(const dfd_engine* e) { return e ? e->error.c_str() : "null engine"; }

int dfd_configuration_load(dfd_engine* e, const char* conf_path) {
    FILE* f = fopen(conf_path, "rb");
    if (!f) return fail(e, "cannot open config file %s: %s", conf_path, strerror(errno));
    std::string text;
    char buf[8192];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    Parser ps{text.data(), text.data() + text.size(), ""};
    e->root = Section();
    if (!ps.body(&e->root, true)) return fail(e, "config %s: %s", conf_path, ps.err.c_str());
    e->conf_path = conf_path;
    size_t slash = e->conf_path.rfind('/');
    e->conf_dir = slash == std::string::npos ? "." : e->conf_path.substr(0, slash);
    e->loaded = true;
    return 0;
}

int dfd_configuration_get_value(const dfd_engine* e, const char* section, const char* key, char* buf, int len) {
    const Section* s = e->root.child(section);
    if (!s) return 0;
    const auto* v = s->find(key);
    if (!v || v->empty()) return 0;
    snprintf(buf, len, "%s", (*v)[0].c_str());
    return (int)strlen(buf);
}

int dfd_model_net_register(dfd_engine* e) {  // model-net.c:71-88
    if (!e->loaded) return fail(e, "model_net_register: no configuration loaded");
    const Section* lg = e->root.child("LPGROUPS");
    if (!lg) return fail(e, "configuration has no LPGROUPS section");
    bool term = false, rtr = false;
    for (const auto& grp : lg->children)
        for (const auto& kv : grp.second.entries) {
            if (kv.first == "modelnet_dragonfly_dally") term = true;
            if (kv.first == "modelnet_dragonfly_dally_router") rtr = true;
        }
    if (!term || !rtr) return fail(e, "LPGROUPS does not select the dragonfly-dally method (modelnet_dragonfly_dally + modelnet_dragonfly_dally_router)");
    e->registered = true;
    return 0;
}

int dfd_lp_type_lookup(const dfd_engine* e, const char* name) {
    (void)e;
    return !strcmp(name, "nw-lp") || !strcmp(name, "modelnet_dragonfly_dally") || !strcmp(name, "modelnet_dragonfly_dally_router");
}

int dfd_codes_mapping_setup(dfd_engine* e) {  // codes_mapping.c:500-574 (layout: repetitions x [types in file order])
    if (!e->registered) return fail(e, "codes_mapping_setup: model_net_register has not been called");
    const Section* lg = e->root.child("LPGROUPS");
    if (lg->children.size() != 1) return fail(e, "exactly one LPGROUPS group is supported");
    HostParams& p = e->hp;
    int off = 0;
    bool ss = false, st = false, sr = false;
    for (const auto& kv : lg->children[0].second.entries) {
        int v = (int)strtol(kv.second.empty() ? "0" : kv.second[0].c_str(), nullptr, 10);
        if (kv.first == "repetitions") {
            p.repetitions = v;
            continue;
        }
        if (kv.first == "nw-lp") {
            p.n_svr = v;
            p.off_svr = off;
            ss = true;
        } else if (kv.first == "modelnet_dragonfly_dally") {
            p.n_term = v;
            p.off_term = off;
            st = true;
        } else if (kv.first == "modelnet_dragonfly_dally_router") {
            p.n_rtr = v;
            p.off_rtr = off;
            sr = true;
        } else
            return fail(e, "LP type \"%s\" is not supported by this engine", kv.first.c_str());
        off += v;
    }
    p.lps_per_rep = off;
    if (!ss || !st || !sr) return fail(e, "LPGROUPS must hold nw-lp, modelnet_dragonfly_dally and modelnet_dragonfly_dally_router");
    if (p.n_svr != p.n_term) return fail(e, "this engine needs one nw-lp per terminal (got %d nw-lp, %d terminals per repetition)", p.n_svr, p.n_term);
    if (p.n_rtr != 1) return fail(e, "this engine needs one router per repetition");
    if ((long long)p.repetitions * p.lps_per_rep >= (1LL << 31)) return fail(e, "too many LPs");
    e->mapped = true;
    return 0;
}

long long dfd_codes_mapping_get_lp_count(const dfd_engine* e, const char* name) {
    const HostParams& p = e->hp;
    if (!strcmp(name, "nw-lp")) return (long long)p.repetitions * p.n_svr;
    if (!strcmp(name, "modelnet_dragonfly_dally")) return (long long)p.repetitions * p.n_term;
    if (!strcmp(name, "modelnet_dragonfly_dally_router")) return (long long)p.repetitions * p.n_rtr;
    return 0;
}
long long dfd_codes_mapping_get_lpid_from_relative(const dfd_engine* e, long long rel, const char* name) {
    const HostParams& p = e->hp;
    if (!strcmp(name, "nw-lp")) return h_svr_gid(p, (unsigned)rel);
    if (!strcmp(name, "modelnet_dragonfly_dally")) return h_term_gid(p, (unsigned)rel);
    if (!strcmp(name, "modelnet_dragonfly_dally_router")) return h_rtr_gid(p, (unsigned)rel);
    return -1;
}
long long dfd_codes_mapping_get_lp_relative_id(const dfd_engine* e, long long gid) {
    const HostParams& p = e->hp;
    long long rep = gid / p.lps_per_rep, off = gid % p.lps_per_rep;
    if (off >= p.off_svr && off < p.off_svr + p.n_svr) return rep * p.n_svr + off - p.off_svr;
    if (off >= p.off_term && off < p.off_term + p.n_term) return rep * p.n_term + off - p.off_term;
    return rep;
}

int dfd_model_net_configure(dfd_engine* e, int* net_ids, int max_ids, int* id_count) {  // model-net.c:90-174
    if (!e->mapped) return fail(e, "model_net_configure: codes_mapping_setup has not been called");
    if (read_params(e)) return 1;
    HostParams& p = e->hp;
    if (p.n_term != p.num_cn) return fail(e, "LPGROUPS holds %d terminals per repetition but num_cns_per_router is %d", p.n_term, p.num_cn);
    if (p.total_routers != p.repetitions * p.n_rtr)  // dally:4925-4931
        return fail(e, "\n Config error: num_routers specified %d total routers computed in the network %d does not match with repetitions * dragonfly_router %d  ",
                    p.num_routers, p.total_routers, p.repetitions * p.n_rtr);
    if (p.modelnet_order.size() != 2) return fail(e, "number of networks in PARAMS:modelnet_order do not match number in LPGROUPS\n");
    int n = 0;
    for (const std::string& name : p.modelnet_order) {
        int id;
        if (name == "dragonfly_dally")
            id = 15;  // DRAGONFLY_DALLY in NETWORK_DEF, codes/model-net.h:74-97
        else if (name == "dragonfly_dally_router")
            id = 16;
        else
            return fail(e, "unknown network in PARAMS:modelnet_order: %s\n", name.c_str());
        if (net_ids && n < max_ids) net_ids[n] = id;
        n++;
    }
    if (id_count) *id_count = n;
    if (p.num_groups > 1 && p.num_groups < 3) return fail(e, "num_groups must be 1 or >= 3 (an intermediate group is drawn for every inter-group packet)");
    if (p.chunk_size <= 0 || p.cn_vc_size < p.chunk_size || p.local_vc_size < p.chunk_size || p.global_vc_size < p.chunk_size)
        return fail(e, "VC sizes must hold at least one chunk");
    if (build_topology(e)) return 1;
    e->configured = true;
    return 0;
}

int dfd_synthetic_configure(dfd_engine* e, const dfd_synthetic_opts* o) {
    if (!e->configured) return fail(e, "synthetic_configure: model_net_configure has not been called");
    e->opts = *o;
    const HostParams& p = e->hp;
    if (o->traffic < 1 || o->traffic > 5) return fail(e, "unknown traffic pattern %d", o->traffic);
    if (o->payload_sz <= 0) return fail(e, "payload_sz must be positive");
    if (p.packet_size <= 0) return fail(e, "packet_size must be set (the reference's fallback for a missing packet_size is not supported)");
    if ((unsigned long long)o->payload_sz > 0xFFFFFFFFull || (p.mn_packet_size + p.chunk_size - 1) / p.chunk_size > 65535)
        return fail(e, "message / packet too large for the packed chunk record");
    if (72 + 72 + 416 > p.message_size)  // model-net.c:294-300
        return fail(e, "Error: model_net trying to transmit an event of size %d but ROSS is configured for events of size %d\n", 560, p.message_size);
    // issue_event synthetic:243-254
    e->load = o->load_per_svr;
    if (o->arrival_time != 0.0) {
        e->mean_interval = o->arrival_time;
        e->load = h_ns_to_GBps(e->mean_interval, o->payload_sz);
    } else if (e->load != 0.0) {
        e->mean_interval = h_bytes_to_ns(o->payload_sz, e->load * p.svr_link_bandwidth);
    } else {
        e->mean_interval = 1000.0;
        e->load = h_ns_to_GBps(e->mean_interval, o->payload_sz);
    }
    if (!(e->mean_interval > 0)) return fail(e, "inter-arrival time must be positive");
    e->opts_set = true;
    return 0;
}

static int make_device_params(dfd_engine* e, DfdParams* P) {
    const HostParams& p = e->hp;
    const dfd_synthetic_opts& o = e->opts;
    memset(P, 0, sizeof *P);
    P->NT = p.total_terminals;
    P->NR = p.total_routers;
    P->a = p.num_routers;
    P->p = p.num_cn;
    P->G = p.num_groups;
    P->h = p.global_radix;
    P->intra_radix = p.intra_grp_radix;
    P->global_radix = p.global_radix;
    P->radix = p.radix;
    P->num_global_channels = p.num_global_channels;
    P->chunk_size = p.chunk_size;
    P->packet_size = p.packet_size;
    P->local_vc = p.local_vc_size;
    P->global_vc = p.global_vc_size;
    P->cn_vc = p.cn_vc_size;
    P->mn_packet_size = (unsigned)std::min<unsigned long long>(p.mn_packet_size, 0xFFFFFFFFull);
    P->routing = p.routing;
    P->scoring = p.scoring;
    P->adaptive_threshold = p.adaptive_threshold;
    P->global_k_picks = p.global_k_picks;
    P->local_bw = p.local_bandwidth;
    P->global_bw = p.global_bandwidth;
    P->cn_bw = p.cn_bandwidth;
    P->router_delay = p.router_delay;
    P->cn_delay = p.cn_delay;
    P->local_delay = p.local_delay;
    P->global_delay = p.global_delay;
    P->cn_credit = p.cn_credit_delay;
    P->local_credit = p.local_credit_delay;
    P->global_credit = p.global_credit_delay;
    P->credit_size = p.credit_size;
    P->nic_seq_delay = p.nic_seq_delay;
    P->lookahead = 0.005;  // g_tw_lookahead, ROSS default
    P->codes_cn_delay = 1 / p.mn_cn_bandwidth;
    P->inj_chunk_cn = h_bytes_to_ns(p.chunk_size, p.cn_bandwidth);
    P->inj_chunk_local = h_bytes_to_ns(p.chunk_size, p.local_bandwidth);
    P->inj_chunk_global = h_bytes_to_ns(p.chunk_size, p.global_bandwidth);
    {
        unsigned tail = (unsigned)o.payload_sz % (unsigned)p.chunk_size;
        P->inj_pay_cn = h_bytes_to_ns(tail, p.cn_bandwidth);
        P->inj_pay_local = h_bytes_to_ns(tail, p.local_bandwidth);
        P->inj_pay_global = h_bytes_to_ns(tail, p.global_bandwidth);
    }
    P->inj_packet_cn = h_bytes_to_ns(o.payload_sz, p.cn_bandwidth);
    P->inv_cn_bw = 1 / p.cn_bandwidth;
    P->magic_p = p.num_cn > 1 ? (~0ULL) / (unsigned long long)p.num_cn + 1 : 0;
    P->magic_a = p.num_routers > 1 ? (~0ULL) / (unsigned long long)p.num_routers + 1 : 0;
    if (p.num_groups > 32767 || p.total_routers > (1 << 30)) return fail(e, "topology too large for the packed chunk record");
    // smallest delay of any event that crosses LPs (every chunk hop is at least its link delay; every credit exactly
    // its credit delay): the asynchronous scheduler needs it strictly positive to make progress
    double L = p.cn_delay;
    L = std::min(L, p.local_delay);
    L = std::min(L, p.global_delay);
    L = std::min(L, p.cn_credit_delay);
    L = std::min(L, p.local_credit_delay);
    L = std::min(L, p.global_credit_delay);
    if (!(L > 0)) return fail(e, "this engine needs strictly positive link and credit delays (smallest is %g ns)", L);
    P->min_delay = L;
    // hysteresis of the chunk bounds: a quarter of the smallest chunk lookahead of the link kind
    double hf = 0.25;
    if (const char* ev = getenv("DFD_HYST_FRAC")) hf = std::min(0.9, std::max(0.0, atof(ev)));
    P->hyst_local = hf * (p.router_delay + p.local_delay);
    P->hyst_global = hf * (p.router_delay + p.global_delay);
    P->act_budget = 4;
    P->idle_sleep_ns = 0;
    P->prefetch = 0;
    if (const char* ev = getenv("DFD_ACT_BUDGET")) P->act_budget = std::max(1, atoi(ev));
    P->ts_end = o.end_time > 0 ? o.end_time : (p.end_time > 0 ? p.end_time : 5.0 * 24 * 60 * 60 * 1000.0 * 1000.0 * 1000.0);
    {   // R_SNAPSHOT schedule: an event at or past the end of the run is never sent (tw_event_new); the others in key order
        std::vector<std::pair<double, unsigned>> sched;
        for (size_t i = 0; i < p.snapshot_times.size(); i++)
            if (0.0 + p.snapshot_times[i] < P->ts_end) sched.push_back({0.0 + p.snapshot_times[i], (unsigned)i});
        std::sort(sched.begin(), sched.end());
        e->topo.snap_ts.clear();
        e->topo.snap_id.clear();
        for (const auto& s : sched) {
            e->topo.snap_ts.push_back(s.first);
            e->topo.snap_id.push_back(s.second);
        }
        P->nsnap = (int)sched.size();
        P->nsnap_all = (int)p.snapshot_times.size();
    }
    P->traffic = o.traffic;
    P->num_msgs = o.num_messages;
    P->payload_sz = o.payload_sz;
    P->rperm_threshold = o.rperm_threshold;
    P->warm_up_time = o.warm_up_time;
    P->mean_interval = e->mean_interval;
    P->num_nodes = p.repetitions * p.n_svr;
    P->nodes_per_grp = P->num_nodes / p.num_groups;
    P->node_eager_limit = p.node_eager_limit;
    P->noop_bypass = p.codes_noop_bypass;
    P->lps_per_rep = p.lps_per_rep;
    P->off_svr = p.off_svr;
    P->off_term = p.off_term;
    P->off_rtr = p.off_rtr;
    // capacities: every bound below follows from the credit protocol (a sender never has more than
    // vc_size/chunk_size uncredited chunks per VC outstanding)
    int cn_chunks = p.cn_vc_size / p.chunk_size, loc_chunks = p.local_vc_size / p.chunk_size, glb_chunks = p.global_vc_size / p.chunk_size;
    P->ncin_cap = next_pow2(DFD_NUM_VCS * cn_chunks);
    P->nkin_cap = next_pow2(cn_chunks);
    P->nsq_cap = next_pow2(std::max(2, o.num_messages + 1));
    int chunks_per_packet = (int)((std::min<unsigned long long>(p.mn_packet_size, (unsigned long long)o.payload_sz) + p.chunk_size - 1) / p.chunk_size);
    if (chunks_per_packet < 1) chunks_per_packet = 1;
    P->ntq_cap = next_pow2(cn_chunks + chunks_per_packet + 2);
    // reassembly tables are only needed when a message is more than one chunk; 4999 is the reference's own limit on
    // concurrently reassembled messages per terminal (DFLY_HASH_TABLE_SIZE, dally:68, :6676-6678)
    bool multi = (unsigned long long)o.payload_sz > p.mn_packet_size || o.payload_sz > p.chunk_size;
    P->nreasm_cap = multi ? 256 : 0;
    if (const char* ev = getenv("DFD_REASM_CAP")) if (multi) P->nreasm_cap = std::max(4, std::min(4999, atoi(ev)));
    // self events of a node: KICKOFF, the first-stage NEW_MSG, LOCAL / REMOTE completions, SCHED_NEXT, T_GENERATE, T_SEND.
    // The second-stage NEW_MSG events, whose number grows with the burst length, live in their own FIFO (mq, nsq_cap).
    P->nev_cap = 24;
    if (const char* ev = getenv("DFD_NEV_CAP")) P->nev_cap = std::max(8, std::min(4096, atoi(ev)));
    // lanes: one link carries at most VCs * vc_size / chunk_size uncredited chunks, and returns at most that many credits
    P->ext = p.intra_grp_radix + p.global_radix;
    P->cap_l = next_pow2(DFD_NUM_VCS * loc_chunks);
    P->cap_g = next_pow2(DFD_NUM_VCS * glb_chunks);
    P->cap_tc = next_pow2(cn_chunks);                // terminal -> router chunks: the terminal's single VC
    P->cap_tk = next_pow2(DFD_NUM_VCS * cn_chunks);  // terminal -> router credits: the router's terminal port has all VCs
    if (std::max(std::max(P->cap_l, P->cap_g), P->cap_tk) > 2048)
        return fail(e, "VC sizes / chunk_size allow more than 2048 chunks in flight on one link: not supported");
    if (p.radix > 128) return fail(e, "router radix %d is larger than 128: not supported", p.radix);
    if (p.radix != P->ext + p.num_cn) return fail(e, "radix %d != local + global + terminal ports", p.radix);
    if ((unsigned long long)p.total_routers * p.lps_per_rep >= (1ull << 30)) return fail(e, "more than 2^30 LPs: not supported");
    P->cslots = p.intra_grp_radix * P->cap_l + p.global_radix * P->cap_g + p.num_cn * P->cap_tc;
    P->kslots = p.intra_grp_radix * P->cap_l + p.global_radix * P->cap_g + p.num_cn * P->cap_tk;
    // pool: pending chunks are bounded by the output VCs, overflowed (uncredited) ones by what the upstream links may have in flight
    long long in_chunks = (long long)p.intra_grp_radix * DFD_NUM_VCS * loc_chunks + (long long)p.global_radix * DFD_NUM_VCS * glb_chunks + (long long)p.num_cn * cn_chunks;
    long long out_chunks = (long long)p.intra_grp_radix * DFD_NUM_VCS * loc_chunks + (long long)p.global_radix * DFD_NUM_VCS * glb_chunks + (long long)p.num_cn * DFD_NUM_VCS * cn_chunks;
    if (in_chunks + out_chunks > (1 << 24)) return fail(e, "VC sizes / chunk_size give more than 2^24 chunks per router: not supported");
    P->rpool_cap = (int)(in_chunks + out_chunks);
    {   // lower bounds of the serialisation time of any chunk this workload can emit (channel words, dfd_core.h): a chunk is
        // chunk_size bytes, or a whole packet smaller than a chunk (packet sizes: the model-net packet size and the tail)
        unsigned long long mb = (unsigned long long)p.chunk_size;
        unsigned long long pay = (unsigned long long)o.payload_sz, mnp = p.mn_packet_size;
        unsigned long long last = pay <= mnp ? pay : (pay % mnp ? pay % mnp : mnp);
        if (pay > mnp) mb = std::min(mb, mnp);
        mb = std::min(mb, last);
        if (mb < 1) mb = 1;
        P->lb_inj_term = h_bytes_to_ns(mb, p.cn_bandwidth);
        P->lb_injrd_cn = h_bytes_to_ns(mb, p.cn_bandwidth) + p.router_delay;
        P->lb_injrd_local = h_bytes_to_ns(mb, p.local_bandwidth) + p.router_delay;
        P->lb_injrd_global = h_bytes_to_ns(mb, p.global_bandwidth) + p.router_delay;
    }
    return 0;
}

static void seed_consts(DfdSeedConsts* K) {
    const unsigned long long M[4] = {2147483647ULL, 2147483543ULL, 2147483423ULL, 2147483323ULL};
    const unsigned long long A0[4] = {45991ULL, 207707ULL, 138556ULL, 49689ULL};
    const unsigned long long S0[4] = {11111111ULL, 22222222ULL, 33333333ULL, 44444444ULL};
    for (int j = 0; j < 4; j++) {
        unsigned long long v = A0[j];
        for (int i = 0; i < 31 + 41; i++) v = (v * v) % M[j];
        K->m[j] = M[j];
        K->seed0[j] = S0[j];
        K->avw[j] = v;
    }
}

static int check_in_degree(dfd_engine* e) {
    // the capacity bounds assume every router is the destination of at most intra_radix local and
    // global_radix global links (true for every symmetric link file)
    const HostParams& p = e->hp;
    std::vector<int> gin(p.total_routers, 0);
    for (int r = 0; r < p.total_routers; r++)
        for (int i = 0; i < p.global_radix; i++) {
            int d = e->topo.gdest[(size_t)r * p.global_radix + i];
            if (d >= 0) gin[d]++;
        }
    for (int r = 0; r < p.total_routers; r++)
        if (gin[r] > p.global_radix) return fail(e, "router %d is the destination of %d global links (> num_global_channels): asymmetric link files are not supported", r, gin[r]);
    std::vector<int> lin(p.num_routers, 0);
    for (int s = 0; s < p.num_routers; s++)
        for (int j = 0; j < p.intra_grp_radix; j++) {
            int d = e->topo.loc_dest[(size_t)s * p.intra_grp_radix + j];
            if (d >= 0) lin[d]++;
        }
    for (int s = 0; s < p.num_routers; s++)
        if (lin[s] > p.intra_grp_radix) return fail(e, "local router %d is the destination of %d local links (> intra-group radix)", s, lin[s]);
    return 0;
}

int dfd_engine_upload(dfd_engine* e, void* stream) {
    if (!e->opts_set) return fail(e, "engine_upload: synthetic_configure has not been called");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(e, "no CUDA device available: the dragonfly-dally B200 engine has no CPU fallback");
    DfdParams P;
    if (make_device_params(e, &P)) return 1;
    if (check_in_degree(e)) return 1;
    // the device arrays are kept across runs of the same shape (cudaMalloc/cudaFree dominate a short run otherwise)
    // (a sharded engine keeps its exchange region too: the peers' CUDA-IPC mappings of it stay valid, dfd_engine_ipc_ready)
    bool reuse = e->device_ready && e->dev.X.rank == e->rank && e->dev.X.world == e->world && e->dev_ensemble == e->ensemble;
    if (reuse) {
        const DfdHostTopo& T = e->topo;
        const std::vector<int>* tabs[] = {&T.loc_off, &T.loc_port, &T.loc_dest, &T.lout_lane, &T.lin_src, &T.gs_port, &T.gs_group, &T.gdest, &T.gout_lane, &T.gin_src, &T.gin_port, &T.routed_off, &T.routed_lid};
        reuse = e->dev.topo_sizes.size() == sizeof tabs / sizeof tabs[0];
        for (size_t i = 0; reuse && i < e->dev.topo_sizes.size(); i++) reuse = tabs[i]->size() == e->dev.topo_sizes[i];  // every table, not the total
    }
    if (reuse) {
        const DfdParams& Q = e->dev.P;
        reuse = Q.NT == P.NT && Q.NR == P.NR && Q.radix == P.radix && Q.a == P.a && Q.p == P.p && Q.G == P.G && Q.h == P.h &&
                Q.ncin_cap == P.ncin_cap && Q.nkin_cap == P.nkin_cap && Q.nsq_cap == P.nsq_cap && Q.ntq_cap == P.ntq_cap &&
                Q.nev_cap == P.nev_cap && Q.nreasm_cap == P.nreasm_cap && Q.cap_l == P.cap_l && Q.cap_g == P.cap_g &&
                Q.cap_tc == P.cap_tc && Q.cap_tk == P.cap_tk && Q.rpool_cap == P.rpool_cap && Q.nsnap_all == P.nsnap_all;
    }
    if (reuse) {
        const int budget = e->dev.P.act_budget, idle_sleep = e->dev.P.idle_sleep_ns, prefetch = e->dev.P.prefetch;  // chosen by dfd_device_create from the grid it sized
        const int npubq = e->dev.P.npubq, npub_blocks = e->dev.P.npub_blocks;
        e->dev.P = P;
        if (!getenv("DFD_ACT_BUDGET")) e->dev.P.act_budget = budget;
        e->dev.P.idle_sleep_ns = idle_sleep;
        e->dev.P.prefetch = prefetch;
        e->dev.P.npubq = npubq;
        e->dev.P.npub_blocks = npub_blocks;
    } else {
        if (e->device_ready) {
            dfd_device_destroy(&e->dev);
            e->device_ready = false;
        }
        e->dev = DfdDevice();
        seed_consts(&e->dev.seeds);
        if (dfd_device_create(&e->dev, P, e->topo, e->rank, e->world, e->ensemble)) {
            std::string msg = e->dev.error;
            dfd_device_destroy(&e->dev);  // frees what a create that failed midway had already allocated
            return fail(e, "%s", msg.c_str());
        }
        e->dev_ensemble = e->ensemble;
        e->device_ready = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dfd_device_upload_topo(&e->dev, e->topo, st)) return fail(e, "%s", e->dev.error);
    e->h2d_bytes = e->dev.topo_bytes + sizeof(DfdParams) + sizeof(DfdArrays) + sizeof(DfdSeedConsts);
    if (dfd_device_reset(&e->dev, st)) return fail(e, "%s", e->dev.error);
    e->uploaded = true;
    e->finished = false;
    return 0;
}

int dfd_engine_reset(dfd_engine* e, void* stream) {
    if (!e->uploaded) return fail(e, "engine_reset: engine_upload has not been called");
    if (dfd_device_reset(&e->dev, (cudaStream_t)stream)) return fail(e, "%s", e->dev.error);
    e->finished = false;
    return 0;
}

int dfd_engine_run(dfd_engine* e, void* stream) {
    if (!e->uploaded) return fail(e, "engine_run: engine_upload has not been called");
    if (dfd_device_run(&e->dev, (cudaStream_t)stream)) return fail(e, "%s", e->dev.error);
    return 0;
}

static void fill_summary_meta(dfd_engine* e) {
    dfd_summary& S = e->sum;
    S.window_ns = e->dev.P.min_delay;
    S.h2d_bytes = e->h2d_bytes;
    S.d2h_bytes = e->res.d2h_bytes;
    S.device_bytes = e->dev.bytes_allocated;
    S.grid_blocks = e->dev.grid;
    S.block_threads = e->dev.block;
    S.num_sms = e->dev.num_sms;
    S.num_terminals = e->hp.total_terminals;
    S.num_routers = e->hp.total_routers;
    S.radix = e->hp.radix;
    const unsigned long long* C = e->res.counters;
    S.cyc_busy = (double)C[20] / e->dev.warps;
    S.cyc_idle = (double)C[21] / e->dev.warps;
    S.cyc_busy_max_warp = (double)C[24];
    S.activations = C[22];
    S.passes = C[23];
}

static void profile_print(dfd_engine* e) {
    if (!getenv("DFD_PROFILE_PRINT")) return;
    const unsigned long long* C = e->res.counters;
    const char* names[] = {"KICKOFF", "REMOTE", "LOCAL", "ACK", "NEW_MSG", "SCHED_NEXT", "T_GENERATE", "T_ARRIVE", "T_SEND", "T_BUFFER", "R_SEND", "R_ARRIVE", "R_BUFFER", "R_SNAPSHOT"};
    for (int i = 0; i < EV_NTYPES; i++)
        if (C[4 + i]) fprintf(stderr, "  %-11s n=%9llu\n", names[i], C[4 + i]);
    fprintf(stderr, "  warps %d, passes %llu, activations past the poll %llu, productive %llu, events/productive activation %.2f\n", e->dev.warps, C[23], C[22], C[1],
            C[1] ? (double)C[0] / C[1] : 0.0);
    fprintf(stderr, "  rounds %llu; cycles per warp inside activations: poll-tail %.0f, terminal bounds %.0f, router phase %.0f, node phase %.0f, publish %.0f\n", C[35],
            (double)C[30] / e->dev.warps, (double)C[31] / e->dev.warps, (double)C[32] / e->dev.warps, (double)C[33] / e->dev.warps, (double)C[34] / e->dev.warps);
    fprintf(stderr, "  cycles per warp: busy passes %.0f, idle passes %.0f, busiest warp %.0f\n", (double)C[20] / e->dev.warps, (double)C[21] / e->dev.warps, (double)C[24]);
}

int dfd_engine_finish(dfd_engine* e, void* stream) {
    if (!e->uploaded) return fail(e, "engine_finish: engine_upload has not been called");
    memset(&e->sum, 0, sizeof e->sum);
    // Sharded run: the per-LP statistics of every partition live in the exchange regions, so rank 0 reads all of them
    // over its CUDA-IPC mappings (the caller has waited for every rank's event loop: stream sync + barrier) and runs the
    // final handlers; the other ranks only check their own control block.
    const int mode = e->world == 1 ? DFD_DL_OWN : (e->rank == 0 ? DFD_DL_ALL : DFD_DL_CTRL);
    if (dfd_device_download(&e->dev, &e->res, (cudaStream_t)stream, mode)) return fail(e, "%s", e->dev.error);
    if (e->res.counters[2] != 0)
        return fail(e, "device error %llu (%s), info 0x%llx", e->res.counters[2], device_error_name(e->res.counters[2]), e->res.counters[3]);
    e->downloaded = true;
    e->slice_downloaded = mode != DFD_DL_CTRL;
    e->gathered = mode == DFD_DL_ALL;
    if (mode == DFD_DL_CTRL) return 0;
    finalize(e);
    fill_summary_meta(e);
    profile_print(e);
    e->finished = true;
    return 0;
}

// ---- group-sharded runs (one process per GPU) ---------------------------------------------------------
int dfd_engine_set_partition(dfd_engine* e, int rank, int world) {
    if (world < 1 || world > DFD_MAX_RANKS || rank < 0 || rank >= world) return fail(e, "bad partition: rank %d of %d (at most %d ranks)", rank, world, DFD_MAX_RANKS);
    if (e->uploaded) return fail(e, "set_partition must be called before engine_upload");
    e->rank = rank;
    e->world = world;
    return 0;
}
int dfd_engine_ipc_export(dfd_engine* e, void* handle64) {
    if (!e->uploaded) return fail(e, "ipc_export: engine_upload has not been called");
    if (dfd_device_ipc_export(&e->dev, handle64)) return fail(e, "%s", e->dev.error);
    return 0;
}
int dfd_engine_ipc_import(dfd_engine* e, const void* handles, int world) {
    if (!e->uploaded) return fail(e, "ipc_import: engine_upload has not been called");
    if (dfd_device_ipc_import(&e->dev, handles, world)) return fail(e, "%s", e->dev.error);
    return 0;
}
// Several ranks of a sharded run inside ONE process on one GPU (each engine sized with dfd_engine_set_ensemble(world)):
// the peers' exchange regions are plain device pointers, no CUDA-IPC needed.  Runs the very same cross-rank code path.
int dfd_engine_attach_local_peers(dfd_engine* e, dfd_engine* const* engines, int world) {
    if (!e->uploaded) return fail(e, "attach_local_peers: engine_upload has not been called");
    if (world != e->world) return fail(e, "attach_local_peers: %d engines for a world of %d", world, e->world);
    for (int k = 0; k < world; k++) {
        const dfd_engine* q = engines[k];
        if (!q || !q->uploaded || q->world != world || q->rank != k) return fail(e, "attach_local_peers: engine %d is not rank %d of this run", k, k);
        if (q->dev.xregion_bytes != e->dev.xregion_bytes) return fail(e, "attach_local_peers: engine %d has a different layout", k);
        e->dev.X.base[k] = q->dev.xregion;
    }
    return 0;
}
int dfd_engine_ipc_ready(const dfd_engine* e) {
    if (!e->uploaded) return 0;
    for (int k = 0; k < e->dev.X.world; k++)
        if (!e->dev.X.base[k]) return 0;
    return 1;
}
namespace {
struct PartHeader {
    int rank, n0, n1, r0, r1, radix;
    unsigned long long counters[64];
};
}
int dfd_partition_bounds(int num_groups, int routers_per_group, int cns_per_router, int rank, int world, int* out4) {
    if (world < 1 || world > DFD_MAX_RANKS || rank < 0 || rank >= world || world > num_groups) return 1;
    DfdPeers X;
    dfd_partition(num_groups, routers_per_group, cns_per_router, rank, world, &X);
    out4[0] = X.n0;
    out4[1] = X.n1;
    out4[2] = X.r0;
    out4[3] = X.r1;
    return 0;
}
long long dfd_engine_partition_bytes(const dfd_engine* e) {
    const DfdPeers& X = e->dev.X;
    size_t nn = X.n1 - X.n0, nr = X.r1 - X.r0;
    return (long long)(sizeof(PartHeader) + nn * (sizeof(NodeState) + sizeof(unsigned) + 2 * sizeof(unsigned long long)) +
                       nr * (sizeof(RouterHdr) + (size_t)e->hp.radix * sizeof(PortState)) +
                       nr * e->hp.snapshot_times.size() * (size_t)e->hp.radix * sizeof(int4));
}
int dfd_engine_export_partition(dfd_engine* e, void* buf, long long len) {
    if (!e->downloaded) return fail(e, "export_partition: engine_finish has not been called");
    if (len < dfd_engine_partition_bytes(e)) return fail(e, "export_partition: buffer too small");
    if (!e->slice_downloaded) {  // engine_finish of a non-root rank fetches only the control block
        if (dfd_device_download(&e->dev, &e->res, nullptr, DFD_DL_OWN)) return fail(e, "%s", e->dev.error);
        e->slice_downloaded = true;
    }
    const DfdPeers& X = e->dev.X;
    size_t nn = X.n1 - X.n0, nr = X.r1 - X.r0, radix = e->hp.radix;
    char* p = (char*)buf;
    PartHeader h;
    memset(&h, 0, sizeof h);
    h.rank = X.rank;
    h.n0 = X.n0;
    h.n1 = X.n1;
    h.r0 = X.r0;
    h.r1 = X.r1;
    h.radix = (int)radix;
    memcpy(h.counters, e->res.counters, sizeof h.counters);
    memcpy(p, &h, sizeof h);
    p += sizeof h;
#define PUT(vec, off, cnt)                                   \
    memcpy(p, (vec).data() + (off), (cnt) * sizeof((vec)[0])); \
    p += (cnt) * sizeof((vec)[0]);
    PUT(e->res.node, X.n0, nn)
    PUT(e->res.ack_count, X.n0, nn)
    PUT(e->res.ack_first, X.n0, nn)
    PUT(e->res.ack_last, X.n0, nn)
    PUT(e->res.rh, X.r0, nr)
    PUT(e->res.port, (size_t)X.r0 * radix, nr * radix)
    PUT(e->res.snap, (size_t)X.r0 * radix * e->hp.snapshot_times.size(), nr * radix * e->hp.snapshot_times.size())
#undef PUT
    return 0;
}
int dfd_engine_import_partition(dfd_engine* e, const void* buf, long long len) {
    if (!e->downloaded) return fail(e, "import_partition: engine_finish has not been called");
    if (len < (long long)sizeof(PartHeader)) return fail(e, "import_partition: truncated buffer");
    const char* p = (const char*)buf;
    PartHeader h;
    memcpy(&h, p, sizeof h);
    p += sizeof h;
    const size_t radix = e->hp.radix;
    if (h.radix != (int)radix || h.rank < 0 || h.rank >= e->world || h.rank == e->rank) return fail(e, "import_partition: partition does not belong to this run");
    {   // the blob must be exactly the partition of h.rank
        DfdPeers Y;
        dfd_partition(e->hp.num_groups, e->hp.num_routers, e->hp.num_cn, h.rank, e->world, &Y);
        if (h.n0 != Y.n0 || h.n1 != Y.n1 || h.r0 != Y.r0 || h.r1 != Y.r1) return fail(e, "import_partition: bounds do not match the partition of rank %d", h.rank);
    }
    const size_t nn = (size_t)(h.n1 - h.n0), nr = (size_t)(h.r1 - h.r0);
    const size_t nsnap = e->hp.snapshot_times.size();
    const size_t need = sizeof(PartHeader) + nn * (sizeof(NodeState) + sizeof(unsigned) + 2 * sizeof(unsigned long long)) +
                        nr * (sizeof(RouterHdr) + radix * sizeof(PortState)) + nr * nsnap * radix * sizeof(int4);
    if ((size_t)len < need) return fail(e, "import_partition: truncated buffer (%lld bytes, need %zu)", len, need);
    if (h.counters[2]) return fail(e, "rank %d reported device error %llu (%s)", h.rank, h.counters[2], device_error_name(h.counters[2]));
    if (e->gathered) return 0;  // engine_finish has already read this partition over the peer mapping
#define GET(vec, off, cnt)                                   \
    memcpy((vec).data() + (off), p, (cnt) * sizeof((vec)[0])); \
    p += (cnt) * sizeof((vec)[0]);
    GET(e->res.node, h.n0, nn)
    GET(e->res.ack_count, h.n0, nn)
    GET(e->res.ack_first, h.n0, nn)
    GET(e->res.ack_last, h.n0, nn)
    GET(e->res.rh, h.r0, nr)
    GET(e->res.port, (size_t)h.r0 * radix, nr * radix)
    GET(e->res.snap, (size_t)h.r0 * radix * nsnap, nr * radix * nsnap)
#undef GET
    e->res.counters[0] += h.counters[0];  // committed events
    e->res.counters[1] += h.counters[1];
    for (int i = 0; i < EV_NTYPES; i++) e->res.counters[4 + i] += h.counters[4 + i];
    return 0;
}
int dfd_engine_finalize(dfd_engine* e) {
    if (!e->downloaded) return fail(e, "engine_finalize: engine_finish has not been called");
    memset(&e->sum, 0, sizeof e->sum);
    finalize(e);
    fill_summary_meta(e);
    e->finished = true;
    return 0;
}

void dfd_abi_sizes(int* summary_bytes, int* synthetic_opts_bytes) {
    if (summary_bytes) *summary_bytes = (int)sizeof(dfd_summary);
    if (synthetic_opts_bytes) *synthetic_opts_bytes = (int)sizeof(dfd_synthetic_opts);
}

int dfd_engine_set_ensemble(dfd_engine* e, int simulations) {
    e->ensemble = simulations > 1 ? simulations : 1;
    return 0;
}

int dfd_tw_run(dfd_engine* e, void* stream) {
    const bool timing = getenv("DFD_TIMING") != nullptr;  // host wall clock of the three phases on stderr
    auto t0 = std::chrono::steady_clock::now();
    if (dfd_engine_upload(e, stream)) return 1;
    if (timing) cudaStreamSynchronize((cudaStream_t)stream);
    auto t1 = std::chrono::steady_clock::now();
    if (dfd_engine_run(e, stream)) return 1;
    if (timing) cudaStreamSynchronize((cudaStream_t)stream);
    auto t2 = std::chrono::steady_clock::now();
    int rc = dfd_engine_finish(e, stream);
    auto t3 = std::chrono::steady_clock::now();
    if (timing) {
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "dfd_tw_run: upload+init %.2f ms, event loop %.2f ms, download+final %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
    }
    return rc;
}

int dfd_lp_io_flush(dfd_engine* e, const char* directory) {
    if (!e->finished) return fail(e, "lp_io_flush: the engine has not finished a run");
    if (mkdir(directory, 0775) != 0 && errno != EEXIST) return fail(e, "cannot create %s: %s", directory, strerror(errno));
    render_lpio(e);
    build_raw(e);
    std::vector<LpioFile> files = e->lpio;
    files.push_back({"raw-lp-stats", e->raw});
    files.push_back({"summary.txt", report_text(e)});
    for (const auto& f : files) {
        std::string path = std::string(directory) + "/" + f.name;
        FILE* fp = fopen(path.c_str(), "wb");
        if (!fp) return fail(e, "cannot write %s: %s", path.c_str(), strerror(errno));
        fwrite(f.data.data(), 1, f.data.size(), fp);
        fclose(fp);
    }
    return 0;
}

int dfd_model_net_report_stats(dfd_engine* e, char* buf, int len) {
    if (!e->finished) return fail(e, "model_net_report_stats: the engine has not finished a run");
    std::string t = report_text(e);
    snprintf(buf, len, "%s", t.c_str());
    return 0;
}

int dfd_get_summary(const dfd_engine* e, dfd_summary* out) {
    *out = e->sum;
    return e->finished ? 0 : 1;
}

int dfd_run_synthetic(const char* conf_path, const dfd_synthetic_opts* opts, const char* lp_io_dir, dfd_summary* out, char* err, int errlen) {
    dfd_engine* e = nullptr;
    dfd_engine_create(&e);
    int ids[4], n = 0;
    int rc = dfd_configuration_load(e, conf_path) || dfd_model_net_register(e) || dfd_codes_mapping_setup(e) ||
             dfd_model_net_configure(e, ids, 4, &n) || dfd_synthetic_configure(e, opts) || dfd_tw_run(e, nullptr) ||
             (lp_io_dir && lp_io_dir[0] ? dfd_lp_io_flush(e, lp_io_dir) : 0);
    if (rc && err) snprintf(err, errlen, "%s", dfd_last_error(e));
    if (!rc && out) dfd_get_summary(e, out);
    dfd_engine_destroy(e);
    return rc;
}

double dfd_bytes_to_ns(unsigned long long bytes, double gib_per_s) { return h_bytes_to_ns(bytes, gib_per_s); }
int dfd_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
}
