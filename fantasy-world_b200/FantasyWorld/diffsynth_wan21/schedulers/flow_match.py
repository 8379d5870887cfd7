"""Mirror of FantasyWorld/diffsynth_wan21/schedulers/flow_match.py (inference subset): rectified-flow sigma schedule
with the Wan shift and the Euler update.  Host scalars only; the latent update of the fused sampler runs in
fwb_cfg_euler_step.  `step` keeps the reference's generic torch form for external callers.
"""
import torch


class FlowMatchScheduler():
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False, reverse_sigmas=False):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.sigma_max, self.sigma_min = sigma_max, sigma_min
        self.inverse_timesteps, self.extra_one_step, self.reverse_sigmas = inverse_timesteps, extra_one_step, reverse_sigmas
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False, shift=None):
        """sigma_i = linspace(start, min, n[+1])[:n]; sigma <- s*sigma / (1 + (s-1)*sigma); t = 1000*sigma.
        ref: flow_match.py:18-40."""
        if training:
            raise NotImplementedError("inference-only mirror")
        if shift is not None:
            self.shift = shift
        start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        n = num_inference_steps
        sig = torch.linspace(start, self.sigma_min, n + 1)[:-1] if self.extra_one_step else torch.linspace(start, self.sigma_min, n)
        if self.inverse_timesteps:
            sig = torch.flip(sig, dims=[0])
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        if self.reverse_sigmas:
            sig = 1 - sig
        self.sigmas = sig
        self.timesteps = sig * self.num_train_timesteps
        self.training = False

    def _index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.cpu()
        return int(torch.argmin((self.timesteps - timestep).abs()))

    def sigma_pair(self, timestep, to_final=False):
        """(sigma_t, sigma_next) for the step that starts at `timestep`.  ref: flow_match.py:43-51."""
        i = self._index(timestep)
        if to_final or i + 1 >= len(self.timesteps):
            nxt = 1.0 if (self.inverse_timesteps or self.reverse_sigmas) else 0.0
        else:
            nxt = float(self.sigmas[i + 1])
        return float(self.sigmas[i]), nxt

    def dsigma(self, timestep, to_final=False) -> float:
        """sigma_next - sigma evaluated in fp32 exactly as the reference's tensor subtraction does."""
        i = self._index(timestep)
        if to_final or i + 1 >= len(self.timesteps):
            nxt = torch.tensor(1.0 if (self.inverse_timesteps or self.reverse_sigmas) else 0.0, dtype=self.sigmas.dtype)
        else:
            nxt = self.sigmas[i + 1]
        return float(nxt - self.sigmas[i])

    def step(self, model_output, timestep, sample, to_final=False, **kwargs):
        """x <- x + v * (sigma_next - sigma).  ref: flow_match.py:43-53."""
        s, nxt = self.sigma_pair(timestep, to_final)
        return sample + model_output * (nxt - s)

    def add_noise(self, original_samples, noise, timestep):
        s, _ = self.sigma_pair(timestep)
        return (1 - s) * original_samples + s * noise
