//! Drop-in for `gym_rs::envs::classical_control::mountain_car::MountainCarEnv` (mountain_car.rs:46-84).
use std::os::raw::c_void;

use gym_rs::core::{ActionReward, Env, EnvProperties};
use gym_rs::envs::classical_control::mountain_car::MountainCarObservation;
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

/// One MountainCar instance living in GPU memory.  As in the reference, the `pub` fields are plain
/// data a caller may assign between steps (mountain_car.rs:49-74); `step` / `reset` push whatever
/// changed since the last exchange with the device.
#[derive(Debug, Serialize)]
pub struct MountainCarEnv {
    pub min_position: O64,
    pub max_position: O64,
    pub max_speed: O64,
    pub goal_position: O64,
    pub goal_velocity: O64,
    pub force: O64,
    pub gravity: O64,
    pub render_mode: RenderMode,
    pub action_space: Discrete,
    pub observation_space: BoxR<MountainCarObservation>,
    pub state: MountainCarObservation,
    pub metadata: Metadata<Self>,
    #[serde(skip_serializing)]
    rand_random: Pcg64,
    #[serde(skip_serializing)]
    handle: *mut ffi::gymrs_env,
    /// what the device holds: the parameter block last pushed, the state last pulled
    #[serde(skip_serializing)]
    device_params: ffi::gymrs_mountain_car_params,
    #[serde(skip_serializing)]
    device_state: MountainCarObservation,
}

fn obs_from(v: &[f32; 2]) -> MountainCarObservation {
    MountainCarObservation::new(OrderedFloat(v[0] as f64), OrderedFloat(v[1] as f64))
}

impl MountainCarEnv {
    /// `MountainCarEnv::new(render_mode)` (mountain_car.rs:341-389).
    pub fn new(render_mode: RenderMode) -> Self {
        assert!(render_mode == RenderMode::None, "the B200 path renders nothing");
        let mut p = ffi::gymrs_mountain_car_params::default();
        let mut handle = std::ptr::null_mut();
        unsafe {
            ffi::check(ffi::gymrs_default_params(ffi::GYMRS_MOUNTAIN_CAR, &mut p as *mut _ as *mut c_void));
            ffi::check(ffi::gymrs_create(ffi::GYMRS_MOUNTAIN_CAR, 1, 0, 0, std::ptr::null(), 0, &mut handle));
        }
        let (rng, _) = rand_random(None);
        let low = MountainCarObservation::new(OrderedFloat(p.min_position), OrderedFloat(-p.max_speed));
        let high = MountainCarObservation::new(OrderedFloat(p.max_position), OrderedFloat(p.max_speed));
        let mut env = Self {
            min_position: OrderedFloat(p.min_position),
            max_position: OrderedFloat(p.max_position),
            max_speed: OrderedFloat(p.max_speed),
            goal_position: OrderedFloat(p.goal_position),
            goal_velocity: OrderedFloat(p.goal_velocity),
            force: OrderedFloat(p.force),
            gravity: OrderedFloat(p.gravity),
            render_mode,
            action_space: Discrete(3),
            observation_space: BoxR::new(low, high),
            state: obs_from(&[0.0; 2]),
            metadata: Metadata::new(RENDER_MODES, 30),
            rand_random: rng,
            handle,
            device_params: p,
            device_state: obs_from(&[0.0; 2]),
        };
        env.pull_state();
        env
    }

    /// The parameter block the `pub` fields describe right now.
    fn params_from_fields(&self) -> ffi::gymrs_mountain_car_params {
        let mut p = self.device_params; // keeps max_episode_steps
        p.min_position = self.min_position.into_inner();
        p.max_position = self.max_position.into_inner();
        p.max_speed = self.max_speed.into_inner();
        p.goal_position = self.goal_position.into_inner();
        p.goal_velocity = self.goal_velocity.into_inner();
        p.force = self.force.into_inner();
        p.gravity = self.gravity.into_inner();
        p
    }

    /// Push the `pub` physics fields to the device now (`step` and `reset` do it on their own when a
    /// field changed).
    pub fn sync_params(&mut self) {
        let p = self.params_from_fields();
        unsafe { ffi::check(ffi::gymrs_set_params(self.handle, &p as *const _ as *const c_void)) };
        self.device_params = p;
    }

    /// Bring the device in line with fields the caller assigned since the last exchange.
    fn push_if_changed(&mut self) {
        if self.params_from_fields() != self.device_params {
            self.sync_params();
        }
        if self.state != self.device_state {
            let s = [self.state.position.into_inner() as f32, self.state.velocity.into_inner() as f32];
            unsafe { ffi::check(ffi::gymrs_set_state(self.handle, s.as_ptr(), std::ptr::null())) };
            self.device_state = self.state;
        }
    }

    fn pull_state(&mut self) {
        let mut s = [0f32; 2];
        unsafe { ffi::check(ffi::gymrs_get_state(self.handle, s.as_mut_ptr(), std::ptr::null_mut())) };
        self.state = obs_from(&s);
        self.device_state = self.state;
    }
}

impl Env for MountainCarEnv {
    type Action = usize;
    type Observation = MountainCarObservation;
    type Info = ();
    type ResetInfo = ();

    fn step(&mut self, action: Self::Action) -> ActionReward<Self::Observation, Self::Info> {
        // mountain_car.rs:402-406
        assert!(unsafe { ffi::gymrs_discrete_contains(3, action as u64) } != 0, "{} (usize) invalid", action);
        self.push_if_changed();
        let act = [action as i32];
        let (mut obs, mut reward, mut done, mut truncated) = ([0f32; 2], [0f32; 1], [0u8; 1], [0u8; 1]);
        unsafe {
            ffi::check(ffi::gymrs_step_host(self.handle, act.as_ptr() as *const c_void, 0, obs.as_mut_ptr(),
                                            reward.as_mut_ptr(), done.as_mut_ptr(), truncated.as_mut_ptr()));
            ffi::check(ffi::gymrs_sync(self.handle, std::ptr::null_mut()));
        }
        self.state = obs_from(&obs);
        self.device_state = self.state;
        ActionReward {
            observation: self.state,
            reward: OrderedFloat(reward[0] as f64),
            done: done[0] != 0,
            truncated: truncated[0] != 0,
            info: None, // mountain_car.rs:433
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
            ([b.low.position.into_inner() as f32, b.low.velocity.into_inner() as f32],
             [b.high.position.into_inner() as f32, b.high.velocity.into_inner() as f32])
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

impl Clone for MountainCarEnv {
    fn clone(&self) -> Self {
        let mut handle = std::ptr::null_mut();
        unsafe { ffi::check(ffi::gymrs_clone(self.handle, &mut handle)) };
        Self {
            min_position: self.min_position,
            max_position: self.max_position,
            max_speed: self.max_speed,
            goal_position: self.goal_position,
            goal_velocity: self.goal_velocity,
            force: self.force,
            gravity: self.gravity,
            render_mode: self.render_mode,
            action_space: self.action_space.clone(),
            observation_space: self.observation_space.clone(),
            state: self.state,
            metadata: self.metadata.clone(),
            rand_random: self.rand_random.clone(),
            handle,
            device_params: self.device_params,
            device_state: self.device_state,
        }
    }
}

impl Drop for MountainCarEnv {
    fn drop(&mut self) {
        self.close();
    }
}

impl EnvProperties for MountainCarEnv {
    type ActionSpace = Discrete;
    type ObservationSpace = BoxR<MountainCarObservation>;

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
