"""Host side of the diffusion sampler: DPM-Solver++ (2M, midpoint) scheduler collapsed to scalar tables.

Mirrors the surface of the reference's `DPMSolverMultistepScheduler`
(`vibevoice/schedule/dpm_solver.py:122-295`: `.config`, `.from_config`, `.set_timesteps`, `.timesteps`,
`.sigmas`) for the configuration VibeVoice instantiates (`modeling_vibevoice.py:138-142`: cosine betas,
v-prediction, solver_order 2, dpmsolver++, midpoint, lower_order_final, final sigma zero, linspace
spacing).  Where the reference calls `scheduler.step()` once per inner iteration with CPU 0-dim tensors
(`dpm_solver.py:935-1022`), this class evaluates the same fp32 expressions once per `set_timesteps` and
hands per-step coefficients {a0, s0, ks, kx, rinv, order} to the CUDA sampler:

    x0 = a0*z - s0*v                                   (:581-584)
    z' = ks*z - kx*x0                                  order 1, steps 0 and N-1   (:669-677)
    z' = ks*z - kx*x0 - 0.5*kx*rinv*(x0 - x0_prev)      order 2 midpoint           (:738-764)

`algorithm_type="sde-dpmsolver++"` (what `demo/gradio_demo.py:141-146` switches to; :680-686, :785-793) keeps both forms with
ks = sigma_t/sigma_s0 * e^-h, kx = -(alpha_t (1 - e^-2h)) and adds `+ kn * noise_i`, kn = sigma_t sqrt(1 - e^-2h): a seventh
coefficient column and one [2n,64] normal draw per step (`randn_tensor`, :993-997).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

_SUPPORTED = dict(solver_order=2, algorithm_type="dpmsolver++", solver_type="midpoint", lower_order_final=True,
                  final_sigmas_type="zero", timestep_spacing="linspace", prediction_type="v_prediction",
                  use_karras_sigmas=False, use_lu_lambdas=False, thresholding=False, euler_at_final=False)


class DPMSolverMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_schedule: str = "cosine", prediction_type: str = "v_prediction",
                 **kwargs):
        cfg = dict(_SUPPORTED, num_train_timesteps=num_train_timesteps, beta_schedule=beta_schedule,
                   prediction_type=prediction_type)
        for k, v in kwargs.items():
            if k == "algorithm_type" and v == "sde-dpmsolver++":
                cfg[k] = v
                continue
            if k in _SUPPORTED and v != _SUPPORTED[k]:
                raise NotImplementedError(
                    "DPMSolverMultistepScheduler(%s=%r) is not on the accelerated path (only %r; algorithm_type may also be "
                    "'sde-dpmsolver++')" % (k, v, _SUPPORTED[k]))
            cfg[k] = v
        if beta_schedule not in ("cosine", "squaredcos_cap_v2"):
            raise NotImplementedError("beta_schedule %r" % beta_schedule)
        self.config = SimpleNamespace(**cfg)
        self._cfg_dict = cfg
        betas = []
        for i in range(num_train_timesteps):            # betas_for_alpha_bar, dpm_solver.py:52-83
            t1, t2 = i / num_train_timesteps, (i + 1) / num_train_timesteps
            ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
            betas.append(min(1 - ab(t2) / ab(t1), 0.999))
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.coef: Optional[np.ndarray] = None

    @classmethod
    def from_config(cls, config, **kwargs):
        d = dict(config.__dict__ if isinstance(config, SimpleNamespace) else config)
        d.update(kwargs)
        return cls(**d)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        ac = self.alphas_cumprod
        sig_all = (((1 - ac) / ac) ** 0.5).numpy()
        ts = np.linspace(0, n_train - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
        sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
        self.sigmas = sigmas
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = len(ts)

        def a_s(sigma):                                   # _sigma_to_alpha_sigma_t, :483-487
            alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
            return alpha_t, sigma * alpha_t

        n = len(ts)
        sde = self.config.algorithm_type == "sde-dpmsolver++"
        coef = np.zeros((n, 7 if sde else 6), np.float32)
        for i in range(n):
            alpha_s0, sigma_s0 = a_s(sigmas[i])
            alpha_t, sigma_t = a_s(sigmas[i + 1])
            lam_t = torch.log(alpha_t) - torch.log(sigma_t)
            lam_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
            h = lam_t - lam_s0
            first = i == 0 or i == n - 1
            rinv = 0.0
            if not first:
                alpha_s1, sigma_s1 = a_s(sigmas[i - 1])
                lam_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
                rinv = (1.0 / ((lam_s0 - lam_s1) / h)).item()
            if sde:
                coef[i] = [alpha_s0.item(), sigma_s0.item(), (sigma_t / sigma_s0 * torch.exp(-h)).item(),
                           -(alpha_t * (1 - torch.exp(-2.0 * h))).item(), rinv, 1.0 if first else 2.0,
                           (sigma_t * torch.sqrt(1.0 - torch.exp(-2 * h))).item()]
            else:
                coef[i] = [alpha_s0.item(), sigma_s0.item(), (sigma_t / sigma_s0).item(),
                           (alpha_t * (torch.exp(-h) - 1.0)).item(), rinv, 1.0 if first else 2.0]
        self.coef = coef
        return self
