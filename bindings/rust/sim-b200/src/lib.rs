//! `BatchedSimulation`: N lockstep replications of a `sim` network on one B200, behind the method
//! names of `sim::simulator::Simulation` (sim/src/simulator/mod.rs:49-303).
//! UNVERIFIED SOURCE: the build image has no Rust toolchain; the tested boundary is the C ABI.
use std::ffi::{CStr, CString};
use std::ptr;

use sim::models::Model;
use sim::simulator::{Connector, Message};
use sim_b200_sys as sys;

#[derive(Debug)]
pub struct SimbError {
    pub status: i32,
    pub name: String,
    pub message: String,
}

fn check(status: i32) -> Result<(), SimbError> {
    if status == 0 {
        return Ok(());
    }
    unsafe {
        Err(SimbError {
            status,
            name: CStr::from_ptr(sys::simb_status_name(status)).to_string_lossy().into_owned(),
            message: CStr::from_ptr(sys::simb_last_error_message()).to_string_lossy().into_owned(),
        })
    }
}

pub struct BatchedSimulation {
    raw: *mut sys::simb_sim,
    replications: u64,
}

impl BatchedSimulation {
    /// `Simulation::post(models, connectors)` for `replications` replications (mod.rs:49-56).
    pub fn post(models: Vec<Model>, connectors: Vec<Connector>, replications: u64, seed: u64)
                -> Result<Self, SimbError> {
        // Model/Connector implement Serialize with exactly the wire form simb_post_json expects
        // (model.rs:28-41: id, type, flattened fields; coupling.rs:7-17).
        let mj = CString::new(serde_json::to_string(&models).unwrap()).unwrap();
        let cj = CString::new(serde_json::to_string(&connectors).unwrap()).unwrap();
        let cfg = sys::simb_config {
            struct_size: std::mem::size_of::<sys::simb_config>() as u32,
            flags: 0,
            replications,
            replication_offset: 0,
            seed,
            rng_kind: sys::SIMB_RNG_PHILOX,
            device: 0,
            steps_per_launch: 0,
            queue_capacity: 0,
            max_pending: 0,
            trace_replications: 0,
            trace_capacity: 0,
            reserved0: 0,
            stat_model_id: ptr::null(),
            batch_warmup: 0,
            batch_size: 0,
            series_capacity: 0,
            reserved1: 0,
        };
        let mut raw = ptr::null_mut();
        check(unsafe { sys::simb_post_json(mj.as_ptr(), cj.as_ptr(), &cfg, &mut raw) })?;
        Ok(Self { raw, replications })
    }

    /// `Simulation::put` (mod.rs:82-85).
    pub fn put(&mut self, models: Vec<Model>, connectors: Vec<Connector>) -> Result<(), SimbError> {
        let mj = CString::new(serde_json::to_string(&models).unwrap()).unwrap();
        let cj = CString::new(serde_json::to_string(&connectors).unwrap()).unwrap();
        check(unsafe { sys::simb_put_json(self.raw, mj.as_ptr(), cj.as_ptr()) })
    }

    /// `Simulation::reset` (mod.rs:131-134).
    pub fn reset(&mut self) -> Result<(), SimbError> {
        check(unsafe { sys::simb_reset(self.raw) })
    }

    /// `Simulation::inject_input` (mod.rs:189-191) on every replication.
    pub fn inject_input(&mut self, message: Message) -> Result<(), SimbError> {
        let t = CString::new(message.target_id()).unwrap();
        let p = CString::new(message.target_port()).unwrap();
        let c = CString::new(message.content()).unwrap();
        check(unsafe { sys::simb_inject_input(self.raw, t.as_ptr(), p.as_ptr(), c.as_ptr(), ptr::null()) })
    }

    /// `Simulation::step` (mod.rs:198-272).
    pub fn step(&mut self) -> Result<(), SimbError> {
        check(unsafe { sys::simb_step(self.raw) })
    }

    /// `Simulation::step_n` (mod.rs:293-303).
    pub fn step_n(&mut self, n: usize) -> Result<(), SimbError> {
        check(unsafe { sys::simb_step_n(self.raw, n as u64) })
    }

    /// `Simulation::step_until` (mod.rs:277-288), per replication.
    pub fn step_until(&mut self, until: f64) -> Result<(), SimbError> {
        check(unsafe { sys::simb_step_until(self.raw, until) })
    }

    /// `Simulation::get_global_time` (mod.rs:99) of every replication.
    pub fn get_global_time(&self) -> Result<Vec<f64>, SimbError> {
        let mut out = vec![0.0f64; self.replications as usize];
        check(unsafe { sys::simb_get_global_time(self.raw, out.as_mut_ptr(), self.replications) })?;
        Ok(out)
    }

    /// `WebSimulation::step_json` (web.rs:161-163) for one observed replication of an instrumented batch:
    /// the text deserialises into the reference's own `Vec<Message>`.
    pub fn step_json(&mut self, replication: u64) -> Result<String, SimbError> {
        let mut out: *const std::os::raw::c_char = ptr::null();
        check(unsafe { sys::simb_step_json(self.raw, replication, &mut out) })?;
        Ok(unsafe { CStr::from_ptr(out) }.to_string_lossy().into_owned())
    }

    /// `WebSimulation::get_records_json` (web.rs:109-111): deserialises into `Vec<ModelRecord>`.
    pub fn get_records_json(&self, model_id: &str, replication: u64) -> Result<String, SimbError> {
        let id = CString::new(model_id).unwrap();
        let mut out: *const std::os::raw::c_char = ptr::null();
        check(unsafe { sys::simb_get_records_json(self.raw, id.as_ptr(), replication, &mut out) })?;
        Ok(unsafe { CStr::from_ptr(out) }.to_string_lossy().into_owned())
    }

    pub fn counters(&self) -> Result<sys::simb_counters, SimbError> {
        let mut c = sys::simb_counters::default();
        check(unsafe { sys::simb_get_counters(self.raw, &mut c) })?;
        Ok(c)
    }
}

impl Drop for BatchedSimulation {
    fn drop(&mut self) {
        unsafe {
            sys::simb_destroy(self.raw);
        }
    }
}
