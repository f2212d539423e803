/* examples/mm1_ci.c -- the C ABI used from plain C99, no Python, no torch.
 *
 * The replication experiment of sim/tests/simulations.rs:20-128 and sim/tests/web.rs:417-616 as ONE batched call:
 * post the M/M/1/14 network for N replications, step every replication until t = 10000, and read the IID
 * confidence interval of the mean waiting time (IndependentSample::confidence_interval_mean) over the
 * replication means.
 *
 *   gcc -std=c99 -O2 -I include examples/mm1_ci.c -o mm1_ci -L sim_b200 -lsimb200 -Wl,-rpath,$PWD/sim_b200
 *   ./mm1_ci [replications] [until]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "simb.h"

static const char* MODELS =
    "[{\"id\":\"generator-01\",\"type\":\"Generator\",\"portsIn\":{},\"portsOut\":{\"job\":\"job\"},"
    "\"messageInterdepartureTime\":{\"exp\":{\"lambda\":0.5}}},"
    "{\"id\":\"processor-01\",\"type\":\"Processor\",\"portsIn\":{\"job\":\"job\"},\"portsOut\":{\"job\":\"processed\"},"
    "\"serviceTime\":{\"exp\":{\"lambda\":0.333333}},\"queueCapacity\":14},"
    "{\"id\":\"storage-01\",\"type\":\"Storage\",\"portsIn\":{\"put\":\"store\",\"get\":\"read\"},"
    "\"portsOut\":{\"stored\":\"stored\"}}]";
static const char* CONNECTORS =
    "[{\"id\":\"connector-01\",\"sourceID\":\"generator-01\",\"targetID\":\"processor-01\",\"sourcePort\":\"job\",\"targetPort\":\"job\"},"
    "{\"id\":\"connector-02\",\"sourceID\":\"processor-01\",\"targetID\":\"storage-01\",\"sourcePort\":\"processed\",\"targetPort\":\"store\"}]";

static int check(int status, const char* what) {
    if (status != SIMB_OK) {
        fprintf(stderr, "%s: %s (%s)\n", what, simb_status_name(status), simb_last_error_message());
        exit(1);
    }
    return status;
}

int main(int argc, char** argv) {
    simb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.replications = argc > 1 ? strtoull(argv[1], NULL, 10) : 65536;
    cfg.seed = 42;
    cfg.rng_kind = SIMB_RNG_PHILOX;
    cfg.stat_model_id = "processor-01";
    const double until = argc > 2 ? atof(argv[2]) : 10000.0;

    simb_sim* sim = NULL;
    check(simb_post_json(MODELS, CONNECTORS, &cfg, &sim), "simb_post_json");
    check(simb_step_until(sim, until), "simb_step_until");

    simb_counters c;
    check(simb_get_counters(sim, &c), "simb_get_counters");
    simb_ci ci;
    check(simb_independent_sample_ci(sim, 0.05, &ci), "simb_independent_sample_ci");
    printf("replications %llu steps %llu events %llu kernel_ms %.3f events_per_s %.4e\n",
           (unsigned long long)cfg.replications, (unsigned long long)c.steps, (unsigned long long)c.events, c.kernel_ms,
           (double)c.events / (c.kernel_ms * 1e-3));
    printf("mean_wait %.17g lower %.17g upper %.17g n %llu\n", ci.mean, ci.lower, ci.upper, (unsigned long long)ci.n);
    check(simb_destroy(sim), "simb_destroy");
    return 0;
}
