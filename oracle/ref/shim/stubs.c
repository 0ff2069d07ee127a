/* stubs.c -- link-time stand-ins for subsystems that are off in every dragonfly-dally hot-path configuration
 * (TEST INFRASTRUCTURE, oracle/).  model-net.c's method table references every network model
 * (codes/model-net.h:74-97, src/networks/model-net/core/model-net.c:21-51); only the two dragonfly-dally methods are
 * compiled from the reference, the others are zero-filled placeholders that are never selected by a dally config. */
#include <ross.h>
#include <stdbool.h>
#include "codes/model-net-method.h"
#include "codes/codes-workload-config.h"

struct model_net_method simplenet_method, simplep2p_method, torus_method, slimfly_method, slimfly_router_method, fattree_method,
    dragonfly_method, dragonfly_router_method, dragonfly_custom_method, dragonfly_custom_router_method, loggp_method,
    express_mesh_method, express_mesh_router_method, dragonfly_plus_method, dragonfly_plus_router_method;

/* the ML surrogate is only configured when the config has a NETWORK_SURROGATE section (dally:2960-3007) */
struct network_surrogate_config;
struct packet_latency_predictor;
bool network_surrogate_configure(char const* const annotation, struct network_surrogate_config* const config,
                                 struct packet_latency_predictor** pl_pred) {
    (void)annotation;
    (void)config;
    (void)pl_pred;
    tw_error(TW_LOC, "shim: the network surrogate is not part of the oracle build");
}

/* YAML workload: section helpers (src/util/codes-workload-config.cxx): inert for a legacy .conf */
void codes_workload_config_check_unsupported_jobs(const char* model_name) { (void)model_name; }
void codes_workload_config_apply_int(const char* key, int* val, int cli_default) { (void)key; (void)val; (void)cli_default; }
void codes_workload_config_apply_double(const char* key, double* val, double cli_default) { (void)key; (void)val; (void)cli_default; }
void codes_workload_config_apply_traffic(int* val, int cli_default, const struct codes_workload_traffic_name* names) { (void)val; (void)cli_default; (void)names; }
