"""DDIM scheduler of OFT's DiffusionActionHead.

The reference builds `diffusers.schedulers.scheduling_ddim.DDIMScheduler(num_train_timesteps=100,
beta_schedule="squaredcos_cap_v2")` (oft/action_model/model.py:220) and uses three things of it: `add_noise`
(model.py:243), `set_timesteps` + `timesteps` (oft_arch.py:225,232) and `step(...).prev_sample` (oft_arch.py:249).
diffusers is a third-party dependency that is not vendored in the reference tree (and absent here), so this is a
restatement of the published DDIM update (Song et al. 2021, eq. 12 with eta = 0) under the scheduler's documented
defaults: epsilon prediction, `clip_sample=True` with range 1, `set_alpha_to_one=True`, `timestep_spacing="leading"`,
`steps_offset=0`.  Parity for this class is unpinned.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch


class DDIMScheduler:
    def __init__(self, num_train_timesteps: int = 100, beta_schedule: str = "squaredcos_cap_v2",
                 clip_sample: bool = True, clip_sample_range: float = 1.0):
        if beta_schedule != "squaredcos_cap_v2":
            raise NotImplementedError("DiffusionActionHead only configures the squaredcos_cap_v2 schedule")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_schedule=beta_schedule,
                                      clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                      prediction_type="epsilon", timestep_spacing="leading", steps_offset=0)
        T = num_train_timesteps
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        betas = [min(1.0 - bar((i + 1) / T) / bar(i / T), 0.999) for i in range(T)]
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(T - 1, -1, -1, dtype=torch.int64)
        self._dev_tables: dict = {}

    def _acp(self, device, dtype) -> torch.Tensor:
        key = (str(device), dtype)
        if key not in self._dev_tables:
            self._dev_tables[key] = self.alphas_cumprod.to(device=device, dtype=dtype)
        return self._dev_tables[key]

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """sqrt(acp[t]) * x0 + sqrt(1 - acp[t]) * noise, `t` one integer per sample."""
        acp = self._acp(original_samples.device, original_samples.dtype)[timesteps.to(original_samples.device).long()]
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return (acp ** 0.5).view(shape) * original_samples + ((1 - acp) ** 0.5).view(shape) * noise

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(f"num_inference_steps {num_inference_steps} > num_train_timesteps {T}")
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps
        ts = [int(round(i * ratio)) for i in range(num_inference_steps)][::-1]
        self.timesteps = torch.tensor(ts, dtype=torch.int64, device=device)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps() before step()")
        if eta != 0.0:
            raise NotImplementedError("the reference steps with eta = 0")
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            r = self.config.clip_sample_range
            x0 = x0.clamp(-r, r)
        prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * model_output
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)
