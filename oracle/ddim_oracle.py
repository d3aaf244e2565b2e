"""TEST INFRASTRUCTURE — CPU restatement of the DDIM scheduler OFT's DiffusionActionHead uses.

PARITY UNPINNED: the reference imports `diffusers.schedulers.scheduling_ddim.DDIMScheduler`
(dexbotic/model/oft/action_model/model.py:9,220); diffusers is a pip dependency (pyproject.toml `diffusers`, no version
pin), not vendored under /root/reference and not installed in this image, so this file restates the published algorithm
(Song, Meng, Ermon: "Denoising Diffusion Implicit Models", eq. 12, eta = 0) with the class's documented defaults:
  beta_schedule "squaredcos_cap_v2": beta_i = min(1 - abar((i+1)/T) / abar(i/T), 0.999), abar(s) = cos^2((s+0.008)/1.008 * pi/2)
  prediction_type "epsilon", clip_sample True (range 1), set_alpha_to_one True, timestep_spacing "leading", steps_offset 0.
It is anchored on the reference's call sites: add_noise (model.py:243), set_timesteps / timesteps / step().prev_sample
(oft_arch.py:225-249).  `oracle/ref_loader.py` installs this class as the `diffusers` stub so that the reference's own
DiffusionActionHead / OFTForCausalLM code runs here around it.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch


class DDIMSchedulerOracle:
    def __init__(self, num_train_timesteps: int = 1000, beta_schedule: str = "linear", **_):
        assert beta_schedule == "squaredcos_cap_v2", "only the schedule the reference configures is restated"
        T = num_train_timesteps
        self.config = SimpleNamespace(num_train_timesteps=T)
        s = np.arange(T + 1, dtype=np.float64) / T
        abar = np.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
        betas = np.minimum(1.0 - abar[1:] / abar[:-1], 0.999).astype(np.float32)
        self.alphas_cumprod = torch.from_numpy(np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32))
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = acp[timesteps] ** 0.5
        sb = (1 - acp[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def step(self, model_output, timestep, sample):
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        x0 = x0.clamp(-1.0, 1.0)
        direction = (1 - a_p) ** 0.5 * model_output
        return SimpleNamespace(prev_sample=a_p ** 0.5 * x0 + direction, pred_original_sample=x0)
