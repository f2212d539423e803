//! Raw FFI declarations for include/simb.h (UNVERIFIED: no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct simb_sim {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct simb_config {
    pub struct_size: u32,
    pub flags: u32,
    pub replications: u64,
    pub replication_offset: u64,
    pub seed: u64,
    pub rng_kind: i32,
    pub device: i32,
    pub steps_per_launch: u32,
    pub queue_capacity: u32,
    pub max_pending: u32,
    pub trace_replications: u32,
    pub trace_capacity: u32,
    pub reserved0: u32,
    pub stat_model_id: *const c_char,
    pub batch_warmup: u32,
    pub batch_size: u32,
    pub series_capacity: u32,
    pub reserved1: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct simb_message {
    pub source_id: [c_char; 48],
    pub source_port: [c_char; 48],
    pub target_id: [c_char; 48],
    pub target_port: [c_char; 48],
    pub content: [c_char; 48],
    pub time: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct simb_counters {
    pub steps: u64,
    pub events: u64,
    pub launches: u64,
    pub kernel_ms: f64,
    pub state_bytes_per_replication: u64,
    pub failed: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct simb_ci {
    pub mean: f64,
    pub variance: f64,
    pub lower: f64,
    pub upper: f64,
    pub n: u64,
}

pub const SIMB_RNG_PHILOX: i32 = 0;
pub const SIMB_RNG_PCG64MCG: i32 = 1;
pub const SIMB_RNG_EXTERNAL: i32 = 2;
pub const SIMB_FLAG_INSTRUMENT: u32 = 1;

extern "C" {
    pub fn simb_last_error_message() -> *const c_char;
    pub fn simb_status_name(status: c_int) -> *const c_char;
    pub fn simb_post_json(models: *const c_char, connectors: *const c_char, config: *const simb_config,
                          out: *mut *mut simb_sim) -> c_int;
    pub fn simb_put_json(sim: *mut simb_sim, models: *const c_char, connectors: *const c_char) -> c_int;
    pub fn simb_reset(sim: *mut simb_sim) -> c_int;
    pub fn simb_reset_messages(sim: *mut simb_sim) -> c_int;
    pub fn simb_reset_global_time(sim: *mut simb_sim) -> c_int;
    pub fn simb_set_rng(sim: *mut simb_sim, rng_kind: c_int, seed: u64) -> c_int;
    pub fn simb_set_rng_stream(sim: *mut simb_sim, draws: *const u64, draws_per_replication: u64) -> c_int;
    pub fn simb_destroy(sim: *mut simb_sim) -> c_int;
    pub fn simb_inject_input(sim: *mut simb_sim, target_id: *const c_char, target_port: *const c_char,
                             content: *const c_char, replica_mask: *const u8) -> c_int;
    pub fn simb_step(sim: *mut simb_sim) -> c_int;
    pub fn simb_step_n(sim: *mut simb_sim, n: u64) -> c_int;
    pub fn simb_step_until(sim: *mut simb_sim, until: f64) -> c_int;
    pub fn simb_get_global_time(sim: *mut simb_sim, out: *mut f64, capacity: u64) -> c_int;
    pub fn simb_get_messages(sim: *mut simb_sim, replication: u64, out: *mut simb_message, capacity: u64,
                             count: *mut u64) -> c_int;
    pub fn simb_get_status(sim: *mut simb_sim, model_id: *const c_char, replication: u64, out: *mut c_char,
                           capacity: u64) -> c_int;
    pub fn simb_get_counters(sim: *mut simb_sim, out: *mut simb_counters) -> c_int;
    pub fn simb_get_wait_stats(sim: *mut simb_sim, sum: *mut f64, count: *mut u32, capacity: u64) -> c_int;
    pub fn simb_independent_sample_ci(sim: *mut simb_sim, alpha: f64, out: *mut simb_ci) -> c_int;
    // WebSimulation string API (sim/src/simulator/web.rs:87-215); *json is owned by the handle
    pub fn simb_get_messages_json(sim: *mut simb_sim, replication: u64, json: *mut *const c_char) -> c_int;
    pub fn simb_get_records_json(sim: *mut simb_sim, model_id: *const c_char, replication: u64,
                                 json: *mut *const c_char) -> c_int;
    pub fn simb_inject_input_json(sim: *mut simb_sim, message_json: *const c_char, replica_mask: *const u8) -> c_int;
    pub fn simb_step_json(sim: *mut simb_sim, replication: u64, json: *mut *const c_char) -> c_int;
    pub fn simb_step_n_json(sim: *mut simb_sim, n: u64, replication: u64, json: *mut *const c_char) -> c_int;
    pub fn simb_step_until_json(sim: *mut simb_sim, until: f64, replication: u64, json: *mut *const c_char) -> c_int;
}
