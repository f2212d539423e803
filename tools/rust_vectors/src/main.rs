//! Golden vectors from the reference itself, for tests/test_rust_vectors.py.
//!
//! For the reference's default stream (`Pcg64Mcg::new(42)`, sim/src/input_modeling/dynamic_rng.rs:7-9):
//!  * the first 10 000 variates of every random variable the engine implements (one fresh stream per variable),
//!    f64 values as their IEEE bit patterns (hex) so that the comparison is exact;
//!  * the complete message list of `poisson_generator_processor_with_capacity` (sim/tests/simulations.rs:20-72):
//!    `Simulation::post(models, connectors).step_n(3000)`.
use std::env;
use std::fs;

use rand_pcg::Pcg64Mcg;
use serde_json::{json, Value};

use sim::input_modeling::dynamic_rng::dyn_rng;
use sim::input_modeling::{
    BooleanRandomVariable, ContinuousRandomVariable, DiscreteRandomVariable, IndexRandomVariable,
};
use sim::models::{Generator, Model, Processor, Storage};
use sim::simulator::{Connector, Simulation};

const N: usize = 10_000;

fn continuous(name: &str, json_form: Value, mut v: ContinuousRandomVariable) -> Value {
    let rng = dyn_rng(Pcg64Mcg::new(42));
    let bits: Vec<String> = (0..N)
        .map(|_| format!("{:016x}", v.random_variate(rng.clone()).unwrap().to_bits()))
        .collect();
    json!({"family": "continuous", "name": name, "variable": json_form, "f64_bits": bits})
}

fn integers(family: &str, name: &str, json_form: Value, values: Vec<u64>) -> Value {
    json!({"family": family, "name": name, "variable": json_form, "u64": values})
}

fn main() {
    let out = env::args().nth(1).unwrap_or_else(|| String::from("rust_vectors.json"));
    use ContinuousRandomVariable as C;
    let mut vars: Vec<Value> = vec![
        continuous("exp", json!({"exp": {"lambda": 0.5}}), C::Exp { lambda: 0.5 }),
        continuous("normal", json!({"normal": {"mean": 3.0, "std_dev": 0.5}}), C::Normal { mean: 3.0, std_dev: 0.5 }),
        continuous("triangular", json!({"triangular": {"min": 0.2, "max": 1.4, "mode": 0.6}}),
                   C::Triangular { min: 0.2, max: 1.4, mode: 0.6 }),
        continuous("uniform", json!({"uniform": {"min": 0.3, "max": 0.9}}), C::Uniform { min: 0.3, max: 0.9 }),
        continuous("gamma", json!({"gamma": {"shape": 2.5, "scale": 0.8}}), C::Gamma { shape: 2.5, scale: 0.8 }),
        continuous("gamma_small", json!({"gamma": {"shape": 0.6, "scale": 1.5}}), C::Gamma { shape: 0.6, scale: 1.5 }),
        continuous("lognormal", json!({"logNormal": {"mu": -0.2, "sigma": 0.5}}), C::LogNormal { mu: -0.2, sigma: 0.5 }),
        continuous("weibull", json!({"weibull": {"shape": 1.2, "scale": 2.0}}), C::Weibull { shape: 1.2, scale: 2.0 }),
        continuous("beta_bb", json!({"beta": {"alpha": 2.0, "beta": 5.0}}), C::Beta { alpha: 2.0, beta: 5.0 }),
        continuous("beta_bc", json!({"beta": {"alpha": 0.5, "beta": 0.7}}), C::Beta { alpha: 0.5, beta: 0.7 }),
    ];
    {
        let rng = dyn_rng(Pcg64Mcg::new(42));
        let mut v = BooleanRandomVariable::Bernoulli { p: 0.3 };
        let vals = (0..N).map(|_| v.random_variate(rng.clone()).unwrap() as u64).collect();
        vars.push(integers("boolean", "bernoulli", json!({"bernoulli": {"p": 0.3}}), vals));
    }
    for (name, form, mut v) in vec![
        ("geometric", json!({"geometric": {"p": 0.2}}), DiscreteRandomVariable::Geometric { p: 0.2 }),
        ("geometric_trivial", json!({"geometric": {"p": 0.8}}), DiscreteRandomVariable::Geometric { p: 0.8 }),
        ("poisson", json!({"poisson": {"lambda": 7.0}}), DiscreteRandomVariable::Poisson { lambda: 7.0 }),
        ("poisson_large", json!({"poisson": {"lambda": 40.0}}), DiscreteRandomVariable::Poisson { lambda: 40.0 }),
        ("uniform", json!({"uniform": {"min": 7, "max": 11}}), DiscreteRandomVariable::Uniform { min: 7, max: 11 }),
    ] {
        let rng = dyn_rng(Pcg64Mcg::new(42));
        let vals = (0..N).map(|_| v.random_variate(rng.clone()).unwrap()).collect();
        vars.push(integers("discrete", name, form, vals));
    }
    for (name, form, mut v) in vec![
        ("uniform", json!({"uniform": {"min": 7, "max": 11}}), IndexRandomVariable::Uniform { min: 7, max: 11 }),
        ("weighted", json!({"weightedIndex": {"weights": [1, 2, 3, 4]}}),
         IndexRandomVariable::WeightedIndex { weights: vec![1, 2, 3, 4] }),
    ] {
        let rng = dyn_rng(Pcg64Mcg::new(42));
        let vals = (0..N).map(|_| v.random_variate(rng.clone()).unwrap() as u64).collect();
        vars.push(integers("index", name, form, vals));
    }

    // sim/tests/simulations.rs:20-72
    let models = vec![
        Model::new(String::from("generator-01"), Box::new(Generator::new(
            C::Exp { lambda: 0.5 }, None, String::from("job"), false, None))),
        Model::new(String::from("processor-01"), Box::new(Processor::new(
            C::Exp { lambda: 0.333333 }, Some(14), String::from("job"), String::from("processed"), false, None))),
        Model::new(String::from("storage-01"), Box::new(Storage::new(
            String::from("store"), String::from("read"), String::from("stored"), false))),
    ];
    let connectors = vec![
        Connector::new(String::from("connector-01"), String::from("generator-01"), String::from("processor-01"),
                       String::from("job"), String::from("job")),
        Connector::new(String::from("connector-02"), String::from("processor-01"), String::from("storage-01"),
                       String::from("processed"), String::from("store")),
    ];
    let mut simulation = Simulation::post(models, connectors);
    let messages = simulation.step_n(3000).unwrap();
    let msgs: Vec<Value> = messages.iter().map(|m| json!({
        "sourceId": m.source_id(), "sourcePort": m.source_port(), "targetId": m.target_id(),
        "targetPort": m.target_port(), "time_bits": format!("{:016x}", m.time().to_bits()), "content": m.content(),
    })).collect();
    let doc = json!({
        "generator": "tools/rust_vectors (ndebuhr/sim, Pcg64Mcg::new(42))",
        "count": N,
        "variables": vars,
        "poisson_generator_processor_with_capacity": {
            "steps": 3000,
            "global_time_bits": format!("{:016x}", simulation.get_global_time().to_bits()),
            "messages": msgs,
        },
    });
    fs::write(&out, serde_json::to_string(&doc).unwrap()).unwrap();
    println!("wrote {}", out);
}
