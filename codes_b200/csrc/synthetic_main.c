/* Twin of src/network-workloads/model-net-synthetic-dragonfly-all.c main() (synthetic:660-817) on top of
 * the B200 engine: same command line, same call sequence, tw_run() replaced by dfd_tw_run().
 *   model-net-synthetic-dragonfly-all --sync=1 --traffic=1 --num_messages=20 --lp-io-dir=out -- cfg.conf */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/dfdally_b200.h"

static const char* optval(const char* arg, const char* key) {
    size_t n = strlen(key);
    return strncmp(arg, key, n) == 0 ? arg + n : NULL;
}

int main(int argc, char** argv) {
    dfd_synthetic_opts o;
    const char *conf = NULL, *lpio = NULL, *v;
    int i, net_ids[4], num_nets = 0;
    static char report[8192];
    dfd_engine* e;
    dfd_synthetic_opts_default(&o);
    for (i = 1; i < argc; i++) {
        const char* a = argv[i];
        if ((v = optval(a, "--traffic="))) o.traffic = atoi(v);
        else if ((v = optval(a, "--num_messages="))) o.num_messages = atoi(v);
        else if ((v = optval(a, "--payload_sz="))) o.payload_sz = atoi(v);
        else if ((v = optval(a, "--rperm_threshold="))) o.rperm_threshold = atoll(v);
        else if ((v = optval(a, "--arrival_time="))) o.arrival_time = atof(v);
        else if ((v = optval(a, "--load_per_svr="))) o.load_per_svr = atof(v);
        else if ((v = optval(a, "--warm_up_time="))) o.warm_up_time = atof(v);
        else if ((v = optval(a, "--end="))) o.end_time = atof(v);
        else if ((v = optval(a, "--lp-io-dir="))) lpio = v;
        else if ((v = optval(a, "--sync=")) || (v = optval(a, "--synch="))) { /* the engine has one (conservative) mode */ }
        else if (!strcmp(a, "--")) continue;
        else if (a[0] != '-') conf = a;
        else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (!conf) {
        printf("\n Usage: mpirun <args> --sync=1/2/3 -- <config_file.conf> ");
        return 0;
    }
    dfd_engine_create(&e);
    if (dfd_configuration_load(e, conf) || dfd_model_net_register(e) || dfd_codes_mapping_setup(e) ||
        dfd_model_net_configure(e, net_ids, 4, &num_nets) || dfd_synthetic_configure(e, &o) || dfd_tw_run(e, NULL) ||
        (lpio && dfd_lp_io_flush(e, lpio)) || dfd_model_net_report_stats(e, report, sizeof report)) {
        fprintf(stderr, "error: %s\n", dfd_last_error(e));
        dfd_engine_destroy(e);
        return 1;
    }
    fputs(report, stdout);
    dfd_engine_destroy(e);
    return 0;
}
