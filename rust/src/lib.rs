//! `gym_rs` `Env` implementations whose `step`/`reset` run on a B200 through libgymrs_b200.so.
//!
//! Two shapes (`ActionReward.reward: O64` and `.done: bool` are scalars, so the trait
//! cannot express a batch):
//!
//! * [`cartpole::CartPoleEnv`] / [`mountain_car::MountainCarEnv`] — batch-of-1 handles implementing
//!   `gym_rs::core::Env` verbatim, so generic code and `examples/cartpole.rs` run unchanged.
//! * [`batched::BatchedEnv`] — N instances per handle; `step(&[usize]) -> BatchStep<'_>` returns SoA
//!   slices, plus `EnvProperties`.
pub mod batched;
pub mod cartpole;
pub mod ffi;
pub mod mountain_car;
