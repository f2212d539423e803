// sim_b200/csrc/lower.hpp -- descriptor lowering: the native equivalent of the reference's
// serde + model_factory path (sim/src/models/model.rs:28-50, model_factory.rs:65-77,
// sim_derive/src/lib.rs:9-32): JSON models/connectors -> validated SimbNetDesc + host metadata.
#pragma once
#include <string>
#include <vector>

#include "net.h"

namespace simb {

struct LowerOptions {
    uint32_t queue_capacity = 0;  // 0 = defaults (64 for unbounded Processor queues, 16 for transient queues)
    uint32_t max_pending = 0;
    bool instrument = false;
    int rng_kind = 0;
    std::string stat_model_id;
    uint32_t batch_warmup = 0, batch_size = 0;  // streaming batch means of the statistic (batch_size 0 = off)
};

struct HostModel {
    std::string id, type;
    std::vector<std::string> ports_in, ports_out;
    bool store_records = false;
    int top = -1;      // index of the top-level entry (HostNet::tops) it is or belongs to
    int coupled = -1;  // index into HostNet::coupled when it is a component of a Coupled model
};
// a Coupled model of the posted document (sim/src/models/coupled.rs:14-67)
struct HostCoupled {
    std::string id;
    std::vector<std::string> ports_in, ports_out;
    int first = 0, count = 0;  // its components in HostNet::models
    struct Coupling {
        std::string source_id, target_id, source_port, target_port;
    };
    struct Parked {
        std::string component_id, port, content;
    };
    std::vector<Parked> parked;          // `state.parkedMessages` given on input (coupled.rs:69-84)
    std::vector<Coupling> eic, eoc, ic;  // external input (target_id, source_port, target_port), external output
                                         // (source_id, source_port, target_port), internal
};
struct HostTop {  // one entry of the posted model list
    std::string id;
    int flat = -1;     // its model when atomic
    int coupled = -1;  // its HostCoupled otherwise
};
struct HostConnector {
    std::string id, source_id, target_id, source_port, target_port;
};
struct RingPreload {
    uint32_t slot;   // absolute slot in ring_w; with RING_PRELOAD_F64: slot in ring_d, one 32-bit half of the f64
    uint32_t value;
};
#define RING_PRELOAD_F64 0x80000000u
#define RING_PRELOAD_HI 0x40000000u

struct HostNet {
    SimbNetDesc desc;
    std::vector<HostModel> models;  // the atomic models, flattened (components in place of their Coupled model)
    std::vector<HostTop> tops;
    std::vector<HostCoupled> coupled;
    std::vector<HostConnector> connectors;
    std::vector<std::string> names;       // content intern table (prefixes and literals)
    std::vector<double> d_init;           // initial value of every f64 row
    std::vector<uint32_t> w_init;         // initial value of every u32 row (ctl / rng rows are special-cased)
    std::vector<RingPreload> ring_preload;
    uint32_t ring_w_slots = 0, ring_d_slots = 0;
    int stat_model = -1;

    int find_model(const std::string& id) const;  // a TOP-LEVEL atomic model (what Simulation::get_status etc. can name)
    int find_top(const std::string& id) const;
    int find_component(int coupled_index, const std::string& id) const;  // flat index, or -1
    // Every events_ext call one Message addressed to (id, port) leads to, in the order the reference makes them
    // (mod.rs:203-221: EVERY posted model whose id matches, in model order; coupled.rs:102-131, 166-184: for a Coupled
    // model the components its external input couplings name for that port, in coupling order).
    struct Delivery {
        int model;      // flat index
        uint32_t port;  // input-port index as the device sees it
    };
    std::vector<Delivery> deliveries(const std::string& id, const std::string& port, bool* bad_coupling = nullptr) const;
    uint32_t intern(const std::string& s, bool* overflow);
    // canonical content code of an arbitrary string (DESIGN.md "content codes")
    uint32_t content_code(const std::string& content, bool* overflow);
    std::string render(uint32_t code) const;
    // index of `port` among model m's input ports as the device sees it
    uint32_t resolve_in_port(int m, const std::string& port) const;
};

// returns a simb_status; *err receives a human-readable reason
int lower_network(const char* models_json, const char* connectors_json, const LowerOptions& opt, HostNet* out,
                  std::string* err);

// one random variable (JSON form of Continuous / Boolean / Discrete / Index, random_variable.rs:17-63) lowered into
// net->models[0] (+ net->wcum); family = SimbFamily
int lower_random_variable(int family, const char* variable_json, SimbNetDesc* net, std::string* err);

const char* action_name(int action);
const char* status_name(int status);

}  // namespace simb
