//! N env instances per handle: the shape the GPU path is built for.  `ActionReward` carries a scalar
//! reward/done (core.rs:94-106), so the batched surface returns SoA slices instead.
use std::os::raw::c_void;

use crate::ffi;

/// Which environment a [`BatchedEnv`] holds.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Kind {
    CartPole,
    MountainCar,
    Pendulum,
}

/// Host-side view of one batched step: `observation[k * num_envs + i]` is field k of env i.
pub struct BatchStep<'a> {
    pub observation: &'a [f32],
    pub reward: &'a [f32],
    pub done: &'a [u8],
    pub truncated: &'a [u8],
}

pub struct BatchedEnv {
    handle: *mut ffi::gymrs_env,
    kind: Kind,
    num_envs: usize,
    global_env_offset: u64,
    obs_dim: usize,
    actions: Vec<i32>,
    obs: Vec<f32>,
    reward: Vec<f32>,
    done: Vec<u8>,
    truncated: Vec<u8>,
}

impl BatchedEnv {
    /// `global_env_offset` keys the reset RNG, so shards of one logical batch on several GPUs give
    /// the same per-env results as a single handle.
    pub fn new(kind: Kind, num_envs: usize, device: i32, global_env_offset: u64) -> Self {
        let k = match kind {
            Kind::CartPole => ffi::GYMRS_CARTPOLE,
            Kind::MountainCar => ffi::GYMRS_MOUNTAIN_CAR,
            Kind::Pendulum => ffi::GYMRS_PENDULUM,
        };
        let mut handle = std::ptr::null_mut();
        unsafe {
            ffi::check(ffi::gymrs_create(k, num_envs as u64, device, global_env_offset, std::ptr::null(), 0, &mut handle));
        }
        let obs_dim = match kind { Kind::CartPole => 4, Kind::MountainCar => 2, Kind::Pendulum => 3 };
        Self {
            handle,
            kind,
            num_envs,
            global_env_offset,
            obs_dim,
            actions: vec![0; num_envs],
            obs: vec![0.0; obs_dim * num_envs],
            reward: vec![0.0; num_envs],
            done: vec![0; num_envs],
            truncated: vec![0; num_envs],
        }
    }

    pub fn reset(&mut self, seed: Option<u64>) -> u64 {
        let mut used = 0u64;
        let p = seed.as_ref().map_or(std::ptr::null(), |s| s as *const u64);
        unsafe { ffi::check(ffi::gymrs_reset(self.handle, p, std::ptr::null(), std::ptr::null(), std::ptr::null(), &mut used)) };
        used
    }

    /// Discrete envs (`Action = usize`, cartpole.rs:390, mountain_car.rs:393).  Panics like the
    /// reference on an action outside the action space.
    pub fn step(&mut self, actions: &[usize], autoreset: bool) -> BatchStep<'_> {
        assert!(self.kind != Kind::Pendulum, "Pendulum takes f32 torques: use step_f32");
        assert_eq!(actions.len(), self.num_envs);
        for (d, a) in self.actions.iter_mut().zip(actions) {
            *d = i32::try_from(*a).unwrap_or(i32::MAX); // out-of-range values are rejected on the device
        }
        let flags = if autoreset { ffi::GYMRS_STEP_AUTORESET } else { 0 };
        unsafe {
            ffi::check(ffi::gymrs_step_host(self.handle, self.actions.as_ptr() as *const c_void, flags,
                                            self.obs.as_mut_ptr(), self.reward.as_mut_ptr(),
                                            self.done.as_mut_ptr(), self.truncated.as_mut_ptr()));
            let mut bad = 0u64;
            let rc = ffi::gymrs_sync(self.handle, &mut bad);
            if rc == ffi::GYMRS_ERR_INVALID_ACTION {
                // gymrs_sync reports the GLOBAL env id
                panic!("{} usize invalid", actions[((bad - self.global_env_offset) as usize) % self.num_envs]);
            }
            ffi::check(rc);
        }
        BatchStep { observation: &self.obs, reward: &self.reward, done: &self.done, truncated: &self.truncated }
    }

    /// Continuous envs (Pendulum): one f32 torque per env, clipped to the action box on the device.
    pub fn step_f32(&mut self, actions: &[f32], autoreset: bool) -> BatchStep<'_> {
        assert!(self.kind == Kind::Pendulum, "discrete envs take usize actions: use step");
        assert_eq!(actions.len(), self.num_envs);
        let flags = if autoreset { ffi::GYMRS_STEP_AUTORESET } else { 0 };
        unsafe {
            ffi::check(ffi::gymrs_step_host(self.handle, actions.as_ptr() as *const c_void, flags,
                                            self.obs.as_mut_ptr(), self.reward.as_mut_ptr(),
                                            self.done.as_mut_ptr(), self.truncated.as_mut_ptr()));
            ffi::check(ffi::gymrs_sync(self.handle, std::ptr::null_mut()));
        }
        BatchStep { observation: &self.obs, reward: &self.reward, done: &self.done, truncated: &self.truncated }
    }

    /// Step with actions that already live on the device (`int32[num_envs]`, or `float[num_envs]` for
    /// Pendulum): no host copies at all.  Asynchronous; results land in [`BatchedEnv::buffers`].
    ///
    /// # Safety
    /// `actions_dev` must be a device pointer readable from the handle's GPU, valid until the step has
    /// run ([`BatchedEnv::sync`]).
    pub unsafe fn step_device(&mut self, actions_dev: *const c_void, autoreset: bool) {
        let flags = if autoreset { ffi::GYMRS_STEP_AUTORESET } else { 0 };
        ffi::check(ffi::gymrs_step(self.handle, actions_dev, flags));
    }

    /// Launch tuning (no counterpart in the reference): `pdl` 0 / 1 / 2 and the step kernel's occupancy, see
    /// `gymrs_set_launch_config` / `gymrs_set_launch_occupancy` in `include/gymrs_b200.h`.  `wide` suits several
    /// handles stepped round-robin on one stream; the defaults suit everything else.
    pub fn set_launch_tuning(&mut self, pdl: i32, wide: bool) {
        unsafe {
            ffi::check(ffi::gymrs_set_launch_config(self.handle, 0, 0, pdl));
            ffi::check(ffi::gymrs_set_launch_occupancy(self.handle, if wide { 1 } else { 0 }));
        }
    }

    /// Device pointers to the handle's SoA arrays (observation, reward, done, truncated, ...).
    pub fn buffers(&self) -> ffi::gymrs_buffers {
        let mut b = std::mem::MaybeUninit::<ffi::gymrs_buffers>::zeroed();
        unsafe {
            ffi::check(ffi::gymrs_get_buffers(self.handle, b.as_mut_ptr()));
            b.assume_init()
        }
    }

    /// Wait for queued device steps; panics like the reference if one of them met an invalid action.
    pub fn sync(&mut self) {
        let mut bad = 0u64;
        let rc = unsafe { ffi::gymrs_sync(self.handle, &mut bad) };
        if rc == ffi::GYMRS_ERR_INVALID_ACTION {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::gymrs_last_error()) };
            panic!("{}", msg.to_string_lossy()); // "<action> usize invalid (env <global id>)"
        }
        ffi::check(rc);
    }

    /// The whole handle as one blob (`Env: Serialize`, core.rs:25).  Unlike the reference's serde
    /// derive, which skips the RNG (cartpole.rs:85-86), the blob carries the counter-based reset
    /// stream, so [`BatchedEnv::restore`] resumes bit-identically.
    pub fn checkpoint(&mut self) -> Vec<u8> {
        let mut bytes = 0usize;
        unsafe { ffi::check(ffi::gymrs_checkpoint_size(self.handle, &mut bytes)) };
        let mut blob = vec![0u8; bytes];
        unsafe { ffi::check(ffi::gymrs_checkpoint_save(self.handle, blob.as_mut_ptr() as *mut c_void, bytes)) };
        blob
    }

    /// Loads a blob written by a handle of the same kind, size and time-limit flag.
    pub fn restore(&mut self, blob: &[u8]) {
        unsafe { ffi::check(ffi::gymrs_checkpoint_load(self.handle, blob.as_ptr() as *const c_void, blob.len())) };
    }

    pub fn num_envs(&self) -> usize {
        self.num_envs
    }
    pub fn obs_dim(&self) -> usize {
        self.obs_dim
    }
    pub fn kind(&self) -> Kind {
        self.kind
    }
    /// Raw handle for callers that keep actions / results on the device (`gymrs_step`, `gymrs_rollout`).
    pub fn raw(&self) -> *mut ffi::gymrs_env {
        self.handle
    }
}

impl Drop for BatchedEnv {
    fn drop(&mut self) {
        unsafe { ffi::gymrs_destroy(self.handle) };
    }
}
