//! Raw `extern "C"` declarations of `include/gymrs_b200.h` (ABI version 1).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct gymrs_env {
    _private: [u8; 0],
}

pub const GYMRS_CARTPOLE: c_int = 0;
pub const GYMRS_MOUNTAIN_CAR: c_int = 1;
pub const GYMRS_PENDULUM: c_int = 2;

pub const GYMRS_OK: c_int = 0;
pub const GYMRS_ERR_INVALID_ACTION: c_int = 1;
pub const GYMRS_ERR_BAD_ARG: c_int = 2;
pub const GYMRS_ERR_CUDA: c_int = 3;
pub const GYMRS_ERR_NO_DEVICE: c_int = 4;
pub const GYMRS_ERR_ALLOC: c_int = 5;
pub const GYMRS_ERR_UNSUPPORTED: c_int = 6;

pub const GYMRS_FLAG_TIME_LIMIT: u32 = 0x1;
pub const GYMRS_STEP_AUTORESET: u32 = 0x1;
pub const GYMRS_HOST_U8_ACTIONS: u32 = 0x1;
pub const GYMRS_HOST_PACKED_DONE: u32 = 0x2;

pub type gymrs_host_step_fn = Option<unsafe extern "C" fn(user: *mut c_void, step: u32, slot: u32)>;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct gymrs_host_rollout_desc {
    pub actions: *const c_void,
    pub obs: *mut f32,
    pub reward: *mut f32,
    pub done: *mut u8,
    pub truncated: *mut u8,
    pub action_slots: u32,
    pub result_slots: u32,
    pub transport: u32,
    pub _pad: u32,
    pub on_step: gymrs_host_step_fn,
    pub user: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct gymrs_cartpole_params {
    pub gravity: f64,
    pub masscart: f64,
    pub masspole: f64,
    pub length: f64,
    pub force_mag: f64,
    pub tau: f64,
    pub theta_threshold_radians: f64,
    pub x_threshold: f64,
    pub kinematics_integrator: i32,
    pub max_episode_steps: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct gymrs_mountain_car_params {
    pub min_position: f64,
    pub max_position: f64,
    pub max_speed: f64,
    pub goal_position: f64,
    pub goal_velocity: f64,
    pub force: f64,
    pub gravity: f64,
    pub max_episode_steps: i32,
    pub _pad: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct gymrs_pendulum_params {
    pub max_speed: f64,
    pub max_torque: f64,
    pub dt: f64,
    pub g: f64,
    pub m: f64,
    pub l: f64,
    pub max_episode_steps: i32,
    pub _pad: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct gymrs_buffers {
    pub num_envs: u64,
    pub ld: u64,
    pub state_dim: u32,
    pub obs_dim: u32,
    pub state: *mut f32,
    pub obs: *mut f32,
    pub reward: *mut f32,
    pub done: *mut u8,
    pub truncated: *mut u8,
    pub steps_beyond_terminated: *mut i32,
    pub elapsed_steps: *mut u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct gymrs_checkpoint_info {
    pub kind: i32,
    pub flags: u32,
    pub num_envs: u64,
    pub global_env_offset: u64,
    pub seed: u64,
    pub step_count: u64,
    pub bytes: u64,
}

extern "C" {
    pub fn gymrs_abi_version() -> c_int;
    pub fn gymrs_last_error() -> *const c_char;
    pub fn gymrs_device_count() -> c_int;
    pub fn gymrs_default_params(kind: c_int, params: *mut c_void) -> c_int;
    pub fn gymrs_create(kind: c_int, num_envs: u64, device: c_int, global_env_offset: u64,
                        params: *const c_void, flags: u32, out: *mut *mut gymrs_env) -> c_int;
    pub fn gymrs_destroy(env: *mut gymrs_env) -> c_int;
    pub fn gymrs_clone(env: *const gymrs_env, out: *mut *mut gymrs_env) -> c_int;
    pub fn gymrs_set_params(env: *mut gymrs_env, params: *const c_void) -> c_int;
    pub fn gymrs_get_params(env: *const gymrs_env, params: *mut c_void) -> c_int;
    pub fn gymrs_set_stream(env: *mut gymrs_env, cuda_stream: *mut c_void) -> c_int;
    pub fn gymrs_get_stream(env: *const gymrs_env, cuda_stream: *mut *mut c_void) -> c_int;
    pub fn gymrs_reset(env: *mut gymrs_env, seed: *const u64, low: *const f32, high: *const f32,
                       mask: *const u8, seed_used: *mut u64) -> c_int;
    pub fn gymrs_step(env: *mut gymrs_env, actions: *const c_void, step_flags: u32) -> c_int;
    pub fn gymrs_step_many(envs: *const *mut gymrs_env, actions: *const *const c_void, count: u32, step_flags: u32,
                           done: *mut u32) -> c_int;
    pub fn gymrs_step_pass(envs: *const *mut gymrs_env, actions: *const *const c_void, count: u32, step_flags: u32,
                           begin_event: *mut c_void, end_event: *mut c_void, done: *mut u32) -> c_int;
    pub fn gymrs_step_host(env: *mut gymrs_env, actions: *const c_void, step_flags: u32, obs: *mut f32,
                           reward: *mut f32, done: *mut u8, truncated: *mut u8) -> c_int;
    pub fn gymrs_step_host_async(env: *mut gymrs_env, actions: *const c_void, step_flags: u32, obs: *mut f32,
                                 reward: *mut f32, done: *mut u8, truncated: *mut u8, ticket: *mut u64) -> c_int;
    pub fn gymrs_host_wait(env: *mut gymrs_env, ticket: u64) -> c_int;
    pub fn gymrs_rollout_host(env: *mut gymrs_env, n_steps: u32, step_flags: u32,
                              desc: *const gymrs_host_rollout_desc) -> c_int;
    pub fn gymrs_rollout(env: *mut gymrs_env, actions: *const c_void, n_steps: u32, step_flags: u32,
                         obs_out: *mut f32, reward_out: *mut f32, done_out: *mut u8) -> c_int;
    pub fn gymrs_get_state(env: *mut gymrs_env, state: *mut f32, sbt: *mut i32) -> c_int;
    pub fn gymrs_set_state(env: *mut gymrs_env, state: *const f32, sbt: *const i32) -> c_int;
    pub fn gymrs_get_buffers(env: *mut gymrs_env, out: *mut gymrs_buffers) -> c_int;
    pub fn gymrs_checkpoint_size(env: *const gymrs_env, bytes: *mut usize) -> c_int;
    pub fn gymrs_checkpoint_save(env: *mut gymrs_env, buf: *mut c_void, bytes: usize) -> c_int;
    pub fn gymrs_checkpoint_load(env: *mut gymrs_env, buf: *const c_void, bytes: usize) -> c_int;
    pub fn gymrs_checkpoint_create(buf: *const c_void, bytes: usize, device: c_int, out: *mut *mut gymrs_env) -> c_int;
    pub fn gymrs_checkpoint_info_of(buf: *const c_void, bytes: usize, info: *mut gymrs_checkpoint_info) -> c_int;
    pub fn gymrs_action_space(env: *const gymrs_env, n: *mut u64, low: *mut f32, high: *mut f32) -> c_int;
    pub fn gymrs_observation_space(env: *const gymrs_env, low: *mut f64, high: *mut f64) -> c_int;
    pub fn gymrs_reward_range(env: *const gymrs_env, low: *mut f64, high: *mut f64) -> c_int;
    pub fn gymrs_num_envs(env: *const gymrs_env, n: *mut u64) -> c_int;
    pub fn gymrs_kind_of(env: *const gymrs_env, kind: *mut c_int) -> c_int;
    pub fn gymrs_sync(env: *mut gymrs_env, bad_env: *mut u64) -> c_int;
    pub fn gymrs_set_launch_config(env: *mut gymrs_env, vec: c_int, block: c_int, pdl: c_int) -> c_int;
    pub fn gymrs_set_launch_occupancy(env: *mut gymrs_env, wide: c_int) -> c_int;
    pub fn gymrs_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn gymrs_host_free(p: *mut c_void) -> c_int;
    pub fn gymrs_clip(value: f64, left_bound: f64, right_bound: f64) -> f64;
    pub fn gymrs_discrete_contains(n: u64, value: u64) -> c_int;
    pub fn gymrs_rand_random(seed: *const u64) -> u64;
}

/// Turns a non-zero status into a panic carrying the library's message: the reference's error
/// model is `panic!`/`assert!`, never `Result` (cartpole.rs:402-406), and nothing unwinds across
/// the FFI boundary itself.
pub fn check(rc: c_int) {
    if rc != GYMRS_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(gymrs_last_error()) };
        panic!("gymrs error {}: {}", rc, msg.to_string_lossy());
    }
}
