//! Drop-in for `gym_rs::envs::classical_control::cartpole::CartPoleEnv` (cartpole.rs:51-87).
use std::os::raw::c_void;

use gym_rs::core::{ActionReward, Env, EnvProperties};
use gym_rs::envs::classical_control::cartpole::{CartPoleObservation, KinematicsIntegrator};
use gym_rs::spaces::{BoxR, Discrete};
use gym_rs::utils::custom::structs::Metadata;
use gym_rs::utils::custom::types::O64;
use gym_rs::utils::renderer::{RenderMode, Renders};
use gym_rs::utils::seeding::rand_random;
use ordered_float::OrderedFloat;
use rand_pcg::Pcg64;
use serde::Serialize;

use crate::ffi;

const RENDER_MODES: &[RenderMode] = &[RenderMode::None];

/// One CartPole instance living in GPU memory.  Field names follow the reference struct.  The
/// reference's `pub` fields are plain data a caller may assign between steps (`env.gravity = ...`,
/// `env.state = ...`, cartpole.rs:60-81) and the next `step` uses them; here `step` / `reset` compare
/// the fields with the copy last exchanged with the device and push whatever changed, so assignments
/// take effect exactly as in the reference ([`CartPoleEnv::sync_params`] remains for eager pushes).
#[derive(Debug, Serialize)]
pub struct CartPoleEnv {
    pub action_space: Discrete,
    pub observation_space: BoxR<CartPoleObservation>,
    pub render_mode: RenderMode,
    pub state: CartPoleObservation,
    pub metadata: Metadata<Self>,
    pub gravity: O64,
    pub masscart: O64,
    pub masspole: O64,
    pub length: O64,
    pub force_mag: O64,
    pub tau: O64,
    pub kinematics_integrator: KinematicsIntegrator,
    pub theta_threshold_radians: O64,
    pub x_threshold: O64,
    pub steps_beyond_terminated: Option<usize>,
    #[serde(skip_serializing)]
    rand_random: Pcg64,
    #[serde(skip_serializing)]
    handle: *mut ffi::gymrs_env,
    /// what the device holds: the parameter block last pushed, the state / counter last pulled
    #[serde(skip_serializing)]
    device_params: ffi::gymrs_cartpole_params,
    #[serde(skip_serializing)]
    device_state: CartPoleObservation,
    #[serde(skip_serializing)]
    device_sbt: Option<usize>,
}

fn obs_from(v: &[f32; 4]) -> CartPoleObservation {
    CartPoleObservation::new(
        OrderedFloat(v[0] as f64),
        OrderedFloat(v[1] as f64),
        OrderedFloat(v[2] as f64),
        OrderedFloat(v[3] as f64),
    )
}

impl CartPoleEnv {
    /// `CartPoleEnv::new(render_mode)` (cartpole.rs:91-144).  Only `RenderMode::None` is supported:
    /// SDL2 rendering is out of scope for the GPU path.
    pub fn new(render_mode: RenderMode) -> Self {
        assert!(render_mode == RenderMode::None, "the B200 path renders nothing");
        let mut p = ffi::gymrs_cartpole_params::default();
        let mut handle = std::ptr::null_mut();
        unsafe {
            ffi::check(ffi::gymrs_default_params(ffi::GYMRS_CARTPOLE, &mut p as *mut _ as *mut c_void));
            ffi::check(ffi::gymrs_create(ffi::GYMRS_CARTPOLE, 1, 0, 0, std::ptr::null(), 0, &mut handle));
        }
        let (rng, _) = rand_random(None);
        let high = CartPoleObservation::new(
            OrderedFloat(p.x_threshold * 2.),
            OrderedFloat(f64::INFINITY),
            OrderedFloat(p.theta_threshold_radians * 2.),
            OrderedFloat(f64::INFINITY),
        );
        let mut env = Self {
            action_space: Discrete(2),
            observation_space: BoxR::new(-high, high),
            render_mode,
            state: obs_from(&[0.0; 4]),
            metadata: Metadata::new(RENDER_MODES, 50),
            gravity: OrderedFloat(p.gravity),
            masscart: OrderedFloat(p.masscart),
            masspole: OrderedFloat(p.masspole),
            length: OrderedFloat(p.length),
            force_mag: OrderedFloat(p.force_mag),
            tau: OrderedFloat(p.tau),
            kinematics_integrator: KinematicsIntegrator::Euler, // cartpole.rs:100
            theta_threshold_radians: OrderedFloat(p.theta_threshold_radians),
            x_threshold: OrderedFloat(p.x_threshold),
            steps_beyond_terminated: None,
            rand_random: rng,
            handle,
            device_params: p,
            device_state: obs_from(&[0.0; 4]),
            device_sbt: None,
        };
        env.pull_state();
        env
    }

    /// The parameter block the `pub` fields describe right now.
    fn params_from_fields(&self) -> ffi::gymrs_cartpole_params {
        let mut p = self.device_params; // keeps max_episode_steps
        p.gravity = self.gravity.into_inner();
        p.masscart = self.masscart.into_inner();
        p.masspole = self.masspole.into_inner();
        p.length = self.length.into_inner();
        p.force_mag = self.force_mag.into_inner();
        p.tau = self.tau.into_inner();
        p.kinematics_integrator = match self.kinematics_integrator {
            KinematicsIntegrator::Euler => 0,
            KinematicsIntegrator::Other => 1, // cartpole.rs:437-441
        };
        p.theta_threshold_radians = self.theta_threshold_radians.into_inner();
        p.x_threshold = self.x_threshold.into_inner();
        p
    }

    /// Push the `pub` physics fields to the device now (`step` and `reset` do it on their own when a
    /// field changed).
    pub fn sync_params(&mut self) {
        let p = self.params_from_fields();
        unsafe { ffi::check(ffi::gymrs_set_params(self.handle, &p as *const _ as *const c_void)) };
        self.device_params = p;
    }

    /// Bring the device in line with fields the caller assigned since the last exchange:
    /// physics constants (cartpole.rs:63-80), `state` (:60) and `steps_beyond_terminated` (:81).
    fn push_if_changed(&mut self) {
        if self.params_from_fields() != self.device_params {
            self.sync_params();
        }
        if self.state != self.device_state || self.steps_beyond_terminated != self.device_sbt {
            let v: Vec<f64> = self.state.into();
            let s = [v[0] as f32, v[1] as f32, v[2] as f32, v[3] as f32];
            let sbt = [self.steps_beyond_terminated.map_or(-1i32, |k| k as i32)];
            unsafe { ffi::check(ffi::gymrs_set_state(self.handle, s.as_ptr(), sbt.as_ptr())) };
            self.device_state = self.state;
            self.device_sbt = self.steps_beyond_terminated;
        }
    }

    fn pull_state(&mut self) {
        let mut s = [0f32; 4];
        let mut sbt = [-1i32; 1];
        unsafe { ffi::check(ffi::gymrs_get_state(self.handle, s.as_mut_ptr(), sbt.as_mut_ptr())) };
        self.state = obs_from(&s);
        self.steps_beyond_terminated = if sbt[0] < 0 { None } else { Some(sbt[0] as usize) };
        self.device_state = self.state;
        self.device_sbt = self.steps_beyond_terminated;
    }
}

impl Env for CartPoleEnv {
    type Action = usize;
    type Observation = CartPoleObservation;
    type Info = ();
    type ResetInfo = ();

    fn step(&mut self, action: Self::Action) -> ActionReward<Self::Observation, Self::Info> {
        // same check and message as the reference (cartpole.rs:402-406); the device validates again
        assert!(unsafe { ffi::gymrs_discrete_contains(2, action as u64) } != 0, "{} usize invalid", action);
        self.push_if_changed();
        let was_terminated = self.steps_beyond_terminated.is_some();
        let act = [action as i32];
        let (mut obs, mut reward, mut done, mut truncated) = ([0f32; 4], [0f32; 1], [0u8; 1], [0u8; 1]);
        unsafe {
            ffi::check(ffi::gymrs_step_host(self.handle, act.as_ptr() as *const c_void, 0, obs.as_mut_ptr(),
                                            reward.as_mut_ptr(), done.as_mut_ptr(), truncated.as_mut_ptr()));
            ffi::check(ffi::gymrs_sync(self.handle, std::ptr::null_mut()));
        }
        self.pull_state();
        if done[0] != 0 && was_terminated {
            // the reference's message, word for word (cartpole.rs:461)
            log::warn!("Calling step after termination may result in undefined behaviour. Consider reseting.");
        }
        ActionReward {
            observation: obs_from(&obs),
            reward: OrderedFloat(reward[0] as f64),
            done: done[0] != 0,
            truncated: truncated[0] != 0,
            info: Some(()),
        }
    }

    fn reset(&mut self, seed: Option<u64>, return_info: bool, options: Option<BoxR<Self::Observation>>)
             -> (Self::Observation, Option<Self::ResetInfo>) {
        let (rng, seed_no) = rand_random(seed);
        self.rand_random = rng;
        if self.params_from_fields() != self.device_params {
            self.sync_params();
        }
        let bounds = options.map(|b| {
            let lo: Vec<f64> = b.low.into();
            let hi: Vec<f64> = b.high.into();
            (lo.iter().map(|v| *v as f32).collect::<Vec<f32>>(), hi.iter().map(|v| *v as f32).collect::<Vec<f32>>())
        });
        let (lo, hi) = match &bounds {
            Some((l, h)) => (l.as_ptr(), h.as_ptr()),
            None => (std::ptr::null(), std::ptr::null()),
        };
        unsafe { ffi::check(ffi::gymrs_reset(self.handle, &seed_no, lo, hi, std::ptr::null(), std::ptr::null_mut())) };
        self.pull_state();
        if return_info { (self.state, Some(())) } else { (self.state, None) }
    }

    fn render(&mut self, _mode: RenderMode) -> Renders {
        Renders::None
    }

    fn close(&mut self) {
        if !self.handle.is_null() {
            unsafe { ffi::gymrs_destroy(self.handle) };
            self.handle = std::ptr::null_mut();
        }
    }
}

impl Clone for CartPoleEnv {
    fn clone(&self) -> Self {
        let mut handle = std::ptr::null_mut();
        unsafe { ffi::check(ffi::gymrs_clone(self.handle, &mut handle)) };
        Self {
            action_space: self.action_space.clone(),
            observation_space: self.observation_space.clone(),
            render_mode: self.render_mode,
            state: self.state,
            metadata: self.metadata.clone(),
            gravity: self.gravity,
            masscart: self.masscart,
            masspole: self.masspole,
            length: self.length,
            force_mag: self.force_mag,
            tau: self.tau,
            kinematics_integrator: self.kinematics_integrator.clone(),
            theta_threshold_radians: self.theta_threshold_radians,
            x_threshold: self.x_threshold,
            steps_beyond_terminated: self.steps_beyond_terminated,
            rand_random: self.rand_random.clone(),
            handle,
            device_params: self.device_params,
            device_state: self.device_state,
            device_sbt: self.device_sbt,
        }
    }
}

impl Drop for CartPoleEnv {
    fn drop(&mut self) {
        self.close();
    }
}

impl EnvProperties for CartPoleEnv {
    type ActionSpace = Discrete;
    type ObservationSpace = BoxR<CartPoleObservation>;

    fn metadata(&self) -> &Metadata<Self> {
        &self.metadata
    }
    fn rand_random(&self) -> &Pcg64 {
        &self.rand_random
    }
    fn action_space(&self) -> &Self::ActionSpace {
        &self.action_space
    }
    fn observation_space(&self) -> &Self::ObservationSpace {
        &self.observation_space
    }
}
