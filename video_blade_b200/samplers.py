"""The 8-step samplers of record around the hot path, restated without diffusers.

The reference never touches the sampler: its inference scripts install ASA into the transformer and hand the pipeline
a stock diffusers scheduler --
  * Wan:  `UniPCMultistepScheduler(prediction_type='flow_prediction', use_flow_sigmas=True,
           num_train_timesteps=1000, flow_shift=3.0)`                     (wanx/train/inference.py:48-52, 8 steps :88-97)
  * Cog:  `CogVideoXDPMScheduler.from_config(..., timestep_spacing="trailing")`  (cogvideox/train/inference.py:64-66,
           8 steps :85-90)
and the trainer's own K-step rollout is `generate_new` (train_wanx_tdm.py:1402-1443 / train_cogvideo_tdm.py:1415-1449,
restated in `dit.generate_new`).  diffusers (pinned 0.34.0, requirements.txt) is not installable in this image, so the two
schedulers below restate the published algorithms of that release (UniPC: Zhao et al. 2023, the `bh2` predictor /
corrector pair on flow sigmas; CogVideoX-DPM: SDE DPM-Solver++(2M) on the v-prediction cosine-free "scaled_linear"
schedule with zero terminal SNR).  **Parity unpinned** for both: there is no diffusers build to compare with; what the
tests pin is (a) the timestep / sigma schedules against the closed forms above, (b) that the first-order UniPC update
IS `generate_new`'s update (eta = 1) for the same sigmas, (c) exactness properties of both solvers on problems with a
known solution (tests/test_bench_and_sampler.py).  They are scaffolding for the clip benchmark, not part of the
hot path: the same velocity / v-prediction callable is sampled with either, ASA untouched.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np
import torch


# ================================================================================================
# Wan: UniPC multistep on flow sigmas
# ================================================================================================
def flow_sigmas(num_inference_steps: int, flow_shift: float = 3.0, num_train_timesteps: int = 1000) -> np.ndarray:
    """`set_timesteps` of UniPCMultistepScheduler with use_flow_sigmas: N shifted sigmas, descending, then the final 0."""
    alphas = np.linspace(1.0, 1.0 / num_train_timesteps, num_inference_steps + 1)
    sig = 1.0 - alphas
    sig = np.flip(flow_shift * sig / (1.0 + (flow_shift - 1.0) * sig))[:-1].copy()
    return np.concatenate([sig, [0.0]])


class UniPCFlowScheduler:
    """UniPCMultistepScheduler(prediction_type='flow_prediction', use_flow_sigmas=True): solver_order 2, `bh2`,
    predict_x0, lower_order_final, corrector on every step but the first."""

    def __init__(self, num_train_timesteps: int = 1000, flow_shift: float = 3.0, solver_order: int = 2,
                 solver_type: str = "bh2", lower_order_final: bool = True, use_corrector: bool = True):
        assert solver_type in ("bh1", "bh2")
        self.T, self.shift, self.order_max, self.solver_type = num_train_timesteps, flow_shift, solver_order, solver_type
        self.lower_order_final, self.use_corrector = lower_order_final, use_corrector
        self.set_timesteps(8)

    def set_timesteps(self, n: int):
        self.sigmas = torch.from_numpy(flow_sigmas(n, self.shift, self.T)).double()
        self.timesteps = (self.sigmas[:-1] * self.T).to(torch.int64)
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.order_max
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    @staticmethod
    def _lambda(sigma):
        return torch.log(1.0 - sigma) - torch.log(sigma)          # alpha_t = 1 - sigma, sigma_t = sigma (flow)

    def _coeffs(self, h, order, rks):
        hh = -h                                                    # predict_x0
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1.0
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        return h_phi_1, B_h, torch.stack(R), torch.stack(b)

    def _history(self, first_index, order, lam_s0, h, m0):
        rks, D1s = [], []
        for i in range(1, order):
            lam_si = self._lambda(self.sigmas[first_index - i])
            rk = (lam_si - lam_s0) / h
            rks.append(rk)
            D1s.append((self.model_outputs[-(i + 1)] - m0) / rk)
        rks.append(torch.tensor(1.0, dtype=torch.float64))
        return torch.stack(rks), D1s

    def _predict(self, sample, order):
        s_t, s_0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        m0 = self.model_outputs[-1]
        lam_t, lam_0 = self._lambda(s_t), self._lambda(s_0)
        h = lam_t - lam_0
        rks, D1s = self._history(self.step_index, order, lam_0, h, m0)
        h_phi_1, B_h, R, b = self._coeffs(h, order, rks)
        x_t = (s_t / s_0) * sample - (1.0 - s_t) * h_phi_1 * m0
        if D1s:
            rhos = torch.tensor([0.5], dtype=torch.float64) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            pred = sum(r * d for r, d in zip(rhos, D1s))
            x_t = x_t - (1.0 - s_t) * B_h * pred
        return x_t

    def _correct(self, model_t, last_sample, order):
        s_t, s_0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        m0 = self.model_outputs[-1]
        lam_t, lam_0 = self._lambda(s_t), self._lambda(s_0)
        h = lam_t - lam_0
        rks, D1s = self._history(self.step_index - 1, order, lam_0, h, m0)
        h_phi_1, B_h, R, b = self._coeffs(h, order, rks)
        rhos = torch.tensor([0.5], dtype=torch.float64) if order == 1 else torch.linalg.solve(R, b)
        x_t = (s_t / s_0) * last_sample - (1.0 - s_t) * h_phi_1 * m0
        corr = sum(r * d for r, d in zip(rhos[:-1], D1s)) if D1s else 0.0
        return x_t - (1.0 - s_t) * B_h * (corr + rhos[-1] * (model_t - m0))

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """One scheduler step: `model_output` is the transformer's flow / velocity prediction at sigmas[step_index]."""
        dt = sample.dtype
        sample64 = sample.double()
        x0 = sample64 - self.sigmas[self.step_index] * model_output.double()      # convert_model_output, flow_prediction
        if self.use_corrector and self.step_index > 0 and self.last_sample is not None:
            sample64 = self._correct(x0, self.last_sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [x0]
        n = len(self.timesteps)
        order = min(self.order_max, n - self.step_index) if self.lower_order_final else self.order_max
        self.this_order = min(order, self.lower_order_nums + 1)
        self.last_sample = sample64
        prev = self._predict(sample64, self.this_order)
        if self.lower_order_nums < self.order_max:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev.to(dt)


@torch.no_grad()
def sample_unipc_flow(velocity_fn: Callable, noise: torch.Tensor, steps: int = 8, flow_shift: float = 3.0,
                      **kw) -> torch.Tensor:
    """WanPipeline's denoising loop (inference.py:88-97) on a guided-velocity callable `velocity_fn(x_t, T)`."""
    sch = UniPCFlowScheduler(flow_shift=flow_shift, **kw)
    sch.set_timesteps(steps)
    x = noise
    for t in sch.timesteps:
        T = torch.full((noise.shape[0],), int(t), device=noise.device, dtype=torch.long)
        x = sch.step(velocity_fn(x, T), x)
    return x


# ================================================================================================
# CogVideoX: SDE DPM-Solver++(2M), "trailing" spacing, v-prediction, zero terminal SNR
# ================================================================================================
def cogvideox_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                             snr_shift_scale: float = 1.0, rescale_zero_snr: bool = True) -> torch.Tensor:
    """scaled_linear betas -> alphas_cumprod, SNR shift, terminal SNR rescaled to zero (CogVideoX-5B's scheduler
    config: snr_shift_scale 1.0, rescale_betas_zero_snr true)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1.0 - snr_shift_scale) * ac)
    if rescale_zero_snr:
        s = ac.sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s = (s - sT) * (s0 / (s0 - sT))
        ac = s ** 2
    return ac


def trailing_timesteps(n: int, num_train_timesteps: int = 1000) -> np.ndarray:
    """timestep_spacing="trailing": round(arange(T, 0, -T/n)) - 1  ->  999, 874, 749, ... for n = 8."""
    return (np.round(np.arange(num_train_timesteps, 0, -num_train_timesteps / n)) - 1).astype(np.int64)


class CogVideoXDPMScheduler:
    def __init__(self, num_train_timesteps: int = 1000, **kw):
        self.T = num_train_timesteps
        self.alphas_cumprod = cogvideox_alphas_cumprod(num_train_timesteps, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0, dtype=torch.float64)      # set_alpha_to_one
        self.set_timesteps(8)

    def set_timesteps(self, n: int):
        self.n = n
        self.timesteps = torch.from_numpy(trailing_timesteps(n, self.T))

    @staticmethod
    def _lam(a):
        return 0.5 * (torch.log(a) - torch.log1p(-a))                          # log sqrt(a / (1 - a))

    def step(self, model_output, old_x0, timestep: int, timestep_back: Optional[int], sample, noise_fn=None):
        """Returns (prev_sample, pred_original_sample).  `noise_fn()` draws N(0,1) of sample's shape (two draws on
        second-order steps, as diffusers does); None = zeros (the deterministic skeleton, used by the tests)."""
        dt = sample.dtype
        x, v = sample.double(), model_output.double()
        prev_t = timestep - self.T // self.n
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_b = self.alphas_cumprod[timestep_back] if timestep_back is not None else None
        x0 = a_t.sqrt() * x - (1.0 - a_t).sqrt() * v                           # v_prediction
        h = self._lam(a_p) - self._lam(a_t)
        m1 = ((1.0 - a_p) / (1.0 - a_t)).sqrt() * torch.exp(-h)
        m2 = torch.expm1(-2.0 * h) * a_p.sqrt()
        m_noise = (1.0 - a_p).sqrt() * (1.0 - torch.exp(-2.0 * h)).sqrt()
        draw = (lambda: noise_fn().double()) if noise_fn is not None else (lambda: torch.zeros_like(x))
        prev = m1 * x - m2 * x0 + m_noise * draw()
        if old_x0 is not None and prev_t >= 0 and a_b is not None:
            r = (self._lam(a_t) - self._lam(a_b)) / h
            d = (1.0 + 1.0 / (2.0 * r)) * x0 - (1.0 / (2.0 * r)) * old_x0.double()
            prev = m1 * x - m2 * d + m_noise * draw()
        return prev.to(dt), x0.to(dt)


@torch.no_grad()
def sample_cogvideox_dpm(v_fn: Callable, noise: torch.Tensor, steps: int = 8, generator=None, **kw) -> torch.Tensor:
    """CogVideoXPipeline's denoising loop with the DPM scheduler (cogvideox/train/inference.py:64-66,85-90) on a guided
    v-prediction callable `v_fn(x_t, T)`."""
    sch = CogVideoXDPMScheduler(**kw)
    sch.set_timesteps(steps)
    x, old_x0 = noise, None
    ts = [int(t) for t in sch.timesteps]

    def nf():
        return torch.randn(noise.shape, device=noise.device, dtype=torch.float32, generator=generator)
    for i, t in enumerate(ts):
        T = torch.full((noise.shape[0],), t, device=noise.device, dtype=torch.long)
        x, old_x0 = sch.step(v_fn(x, T), old_x0, t, ts[i - 1] if i > 0 else None, x, noise_fn=nf)
    return x
