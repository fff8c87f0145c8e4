//! Dumps the outputs of the reference crate's own `step` / `reset` for the inputs listed in
//! tests/golden/reference_inputs.txt (written by tests/golden/make_reference_inputs.py).
//!
//! Every float crosses the file boundary as the 16-hex-digit bit pattern of the f64, so nothing is
//! lost to decimal printing.  Line formats (space separated):
//!
//!   in : step cartpole <integrator 0|1> <action> <x> <x_dot> <theta> <theta_dot>
//!   out: step cartpole <integrator> <action> <x> <x_dot> <theta> <theta_dot> -> <x'> <x_dot'> <theta'> <theta_dot'> <reward> <done 0|1> <truncated 0|1>
//!   in : step mountain_car <action> <position> <velocity>
//!   out: step mountain_car <action> <position> <velocity> -> <position'> <velocity'> <reward> <done> <truncated>
//!   in : seq cartpole <action> <n_steps> <x> <x_dot> <theta> <theta_dot>      (no reset in between: cartpole.rs:455-464)
//!   out: seq cartpole <action> <n_steps> ... -> then n_steps lines `  <x'> <x_dot'> <theta'> <theta_dot'> <reward> <done>`
//!   in : reset cartpole|mountain_car <seed>
//!   out: reset <env> <seed> -> <state values...>          (rand_pcg stream: range / distribution evidence only)
//!
//! The envs are built with RenderMode::None, under which `step` never touches SDL
//! (src/utils/renderer.rs:40-62).
use std::env;
use std::fmt::Write as _;
use std::fs;

use gym_rs::core::Env;
use gym_rs::envs::classical_control::cartpole::{CartPoleEnv, CartPoleObservation, KinematicsIntegrator};
use gym_rs::envs::classical_control::mountain_car::{MountainCarEnv, MountainCarObservation};
use gym_rs::utils::renderer::RenderMode;
use ordered_float::OrderedFloat;

fn f(hex: &str) -> f64 {
    f64::from_bits(u64::from_str_radix(hex, 16).expect("16 hex digits"))
}

fn h(v: f64) -> String {
    format!("{:016x}", v.to_bits())
}

fn cartpole_at(s: &[f64], integrator: u32) -> CartPoleEnv {
    let mut env = CartPoleEnv::new(RenderMode::None);
    env.state = CartPoleObservation::new(OrderedFloat(s[0]), OrderedFloat(s[1]), OrderedFloat(s[2]), OrderedFloat(s[3]));
    env.steps_beyond_terminated = None;
    if integrator == 1 {
        env.kinematics_integrator = KinematicsIntegrator::Other;
    }
    env
}

fn main() {
    let args: Vec<String> = env::args().collect();
    if args.len() != 3 {
        eprintln!("usage: gymrs_ref_fixtures <reference_inputs.txt> <reference_fixtures.txt>");
        std::process::exit(2);
    }
    let input = fs::read_to_string(&args[1]).expect("read inputs");
    let mut out = String::new();
    writeln!(out, "# outputs of gym-rs {} (the reference crate itself); see oracle/ref_fixtures/src/main.rs", "0.3.1").unwrap();
    for line in input.lines() {
        let t: Vec<&str> = line.split_whitespace().collect();
        if t.is_empty() || t[0].starts_with('#') {
            continue;
        }
        match (t[0], t[1]) {
            ("step", "cartpole") => {
                let integrator: u32 = t[2].parse().unwrap();
                let action: usize = t[3].parse().unwrap();
                let s: Vec<f64> = t[4..8].iter().map(|x| f(x)).collect();
                let mut env = cartpole_at(&s, integrator);
                let r = env.step(action);
                let o: Vec<f64> = r.observation.into();
                writeln!(out, "{} -> {} {} {} {} {} {} {}", line.trim(), h(o[0]), h(o[1]), h(o[2]), h(o[3]),
                         h(r.reward.into_inner()), r.done as u8, r.truncated as u8).unwrap();
            }
            ("step", "mountain_car") => {
                let action: usize = t[2].parse().unwrap();
                let mut env = MountainCarEnv::new(RenderMode::None);
                env.state = MountainCarObservation { position: OrderedFloat(f(t[3])), velocity: OrderedFloat(f(t[4])) };
                let r = env.step(action);
                let o: Vec<f64> = r.observation.into();
                writeln!(out, "{} -> {} {} {} {} {}", line.trim(), h(o[0]), h(o[1]), h(r.reward.into_inner()),
                         r.done as u8, r.truncated as u8).unwrap();
            }
            ("seq", "cartpole") => {
                let action: usize = t[2].parse().unwrap();
                let n: usize = t[3].parse().unwrap();
                let s: Vec<f64> = t[4..8].iter().map(|x| f(x)).collect();
                let mut env = cartpole_at(&s, 0);
                writeln!(out, "{} ->", line.trim()).unwrap();
                for _ in 0..n {
                    let r = env.step(action);
                    let o: Vec<f64> = r.observation.into();
                    writeln!(out, "  {} {} {} {} {} {}", h(o[0]), h(o[1]), h(o[2]), h(o[3]), h(r.reward.into_inner()),
                             r.done as u8).unwrap();
                }
            }
            ("reset", "cartpole") => {
                let seed: u64 = t[2].parse().unwrap();
                let mut env = CartPoleEnv::new(RenderMode::None);
                let (obs, _) = env.reset(Some(seed), false, None);
                let o: Vec<f64> = obs.into();
                writeln!(out, "{} -> {} {} {} {}", line.trim(), h(o[0]), h(o[1]), h(o[2]), h(o[3])).unwrap();
            }
            ("reset", "mountain_car") => {
                let seed: u64 = t[2].parse().unwrap();
                let mut env = MountainCarEnv::new(RenderMode::None);
                let (obs, _) = env.reset(Some(seed), false, None);
                let o: Vec<f64> = obs.into();
                writeln!(out, "{} -> {} {}", line.trim(), h(o[0]), h(o[1])).unwrap();
            }
            _ => panic!("unknown input line: {line}"),
        }
    }
    fs::write(&args[2], out).expect("write fixtures");
}
