//! BASELINE config 1 (plumbing, one env): random-action CartPole episodes through the GPU-backed
//! `Env` implementation.  Same protocol as the reference example (15 episodes, at most 475 steps
//! each, reset after every episode) but with `RenderMode::None`, since the B200 path draws nothing.
use gym_rs::core::Env;
use gym_rs::utils::renderer::RenderMode;
use gym_rs_b200::cartpole::CartPoleEnv;
use rand::Rng;

const EPISODES: usize = 15;
const MAX_STEPS: usize = 475;

fn run_episode(env: &mut CartPoleEnv, rng: &mut impl Rng) -> f64 {
    let mut total = 0.0;
    for _ in 0..MAX_STEPS {
        let outcome = env.step(rng.gen_range(0..2usize));
        total += outcome.reward.into_inner();
        if outcome.done {
            break;
        }
    }
    total
}

fn main() {
    let mut rng = rand::thread_rng();
    let mut env = CartPoleEnv::new(RenderMode::None);
    let returns: Vec<f64> = (0..EPISODES)
        .map(|_| {
            env.reset(None, false, None);
            run_episode(&mut env, &mut rng)
        })
        .collect();
    println!("episode returns: {returns:?}");
    env.close();
}
