//! Scalar adapter: ONE env behind the reference's `Gym` shape (`/root/reference/src/classic_control/cartpole.rs:234-357`),
//! for callers that still step a single env.  It owns a 1-env handle in MANUAL mode (`auto_reset = 0`), so
//! `steps_since_reset` / `steps_beyond_terminated` behave exactly as in `cartpole.rs:296-347`: the caller resets after
//! `done || truncated`, as every reference caller does (`cartpole.rs:468-470`).  NOT COMPILED HERE (no cargo/rustc).
//!
//! The `impl modurl::Gym for ScalarGym` block is behind the `modurl` feature: the trait's source (modurl rev 4ddf128)
//! is not on disk in this image, so its shape is the one the three reference impls imply.

use crate::{GpuVecEnv, HostActions, Kind, MgymError};

/// `StepInfo { state, reward, done, truncated }` (cartpole.rs:300-305) with the state on the host.
#[derive(Debug, Clone, PartialEq)]
pub struct StepInfo {
    pub state: Vec<f32>,
    pub reward: f32,
    pub done: bool,
    pub truncated: bool,
}

/// A scalar action: `Discrete(n)` index (the reference passes a rank-0 u32 tensor, cartpole.rs:257) or a Box value.
#[derive(Debug, Clone, Copy)]
pub enum Action {
    Discrete(u32),
    Box(f32),
}

pub struct ScalarGym {
    env: GpuVecEnv,
    obs_dim: usize,
}

impl ScalarGym {
    pub fn new(kind: Kind, device: i32, seed: u64) -> Result<Self, MgymError> {
        // validate_actions: the reference asserts action_space.contains(&action) (cartpole.rs:252); here the
        // offending step returns Err(InvalidAction) and nothing is stepped
        let env = GpuVecEnv::builder(kind, 1)?.device(device).seed(seed).auto_reset(false).validate_actions(true).build()?;
        let obs_dim = env.obs_dim();
        Ok(Self { env, obs_dim })
    }

    /// Gym::reset (cartpole.rs:238-249): a fresh state from the per-env Philox stream.
    pub fn reset(&mut self) -> Result<Vec<f32>, MgymError> {
        self.env.reset(std::ptr::null_mut())?;
        self.env.get_obs_host()
    }

    /// Gym::step (cartpole.rs:251-348).
    pub fn step(&mut self, action: Action) -> Result<StepInfo, MgymError> {
        let mut state = vec![0f32; self.obs_dim];
        let (mut reward, mut flags) = ([0f32; 1], [0u8; 1]);
        match action {
            Action::Discrete(a) => {
                // a value that does not fit a byte is out of range for every Discrete space of this crate: 255 fails
                // the validation exactly like the original would
                let byte = [u8::try_from(a).unwrap_or(u8::MAX)];
                self.env.step_host(HostActions::Discrete(&byte), &mut state, &mut reward, &mut flags)?
            }
            Action::Box(a) => self.env.step_host(HostActions::Box(&[a]), &mut state, &mut reward, &mut flags)?,
        }
        Ok(StepInfo { state, reward: reward[0], done: flags[0] & 1 != 0, truncated: flags[0] & 2 != 0 })
    }

    /// Testable::set_state (cartpole.rs:444-446).
    pub fn set_state(&mut self, state: &[f32]) -> Result<(), MgymError> {
        self.env.set_state(state)
    }
}

#[cfg(feature = "modurl")]
mod gym_trait {
    use super::*;
    use candle_core::{Device, Tensor};
    use modurl::gym::{Gym, StepInfo as ModurlStepInfo};
    use modurl::spaces::{BoxSpace, Discrete, Space};

    impl Gym for ScalarGym {
        type Error = candle_core::Error;
        type SpaceError = candle_core::Error;

        fn reset(&mut self) -> Result<Tensor, Self::Error> {
            let obs = ScalarGym::reset(self).map_err(|e| candle_core::Error::Msg(e.to_string()))?;
            Tensor::from_vec(obs, vec![self.obs_dim], &Device::Cpu)
        }

        fn step(&mut self, action: Tensor) -> Result<ModurlStepInfo, Self::Error> {
            let a = if self.env.is_continuous() {
                Action::Box(action.to_vec0::<f32>()?)
            } else {
                Action::Discrete(action.to_vec0::<u32>()?) // cartpole.rs:257
            };
            let info = ScalarGym::step(self, a).map_err(|e| candle_core::Error::Msg(e.to_string()))?;
            Ok(ModurlStepInfo {
                state: Tensor::from_vec(info.state, vec![self.obs_dim], &Device::Cpu)?,
                reward: info.reward,
                done: info.done,
                truncated: info.truncated,
            })
        }

        fn observation_space(&self) -> Box<dyn Space<Error = Self::SpaceError>> {
            let (low, high) = self.env.observation_bounds();
            Box::new(BoxSpace::new(
                Tensor::from_vec(low, vec![self.obs_dim], &Device::Cpu).unwrap(),
                Tensor::from_vec(high, vec![self.obs_dim], &Device::Cpu).unwrap(),
            ))
        }

        fn action_space(&self) -> Box<dyn Space<Error = Self::SpaceError>> {
            match self.env.num_actions() {
                0 => {
                    let (low, high) = self.env.action_bounds();
                    Box::new(BoxSpace::new(
                        Tensor::from_vec(vec![low], vec![1], &Device::Cpu).unwrap(),
                        Tensor::from_vec(vec![high], vec![1], &Device::Cpu).unwrap(),
                    ))
                }
                n => Box::new(Discrete::new(n as usize)), // cartpole.rs:68
            }
        }
    }
}
