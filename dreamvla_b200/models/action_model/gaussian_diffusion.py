"""IDDPM Gaussian diffusion + DDIM respacing with DEVICE-RESIDENT fp32 schedule tables.

Mirror of reference models/action_model/gaussian_diffusion.py:98-353,522-689,870-882 and respace.py:12-116,193-205,
restricted to what DreamVLA executes: epsilon-prediction, fixed-small variance, MSE loss, `q_sample`,
`ddim_sample_loop` with eta = 0 and clip_denoised = False.  The reference keeps float64 numpy tables and uploads
them on every call (`_extract_into_tensor`, gaussian_diffusion.py:870-882); here they are computed once in float64,
rounded to fp32 exactly as `.float()` does, and cached per device.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "squaredcos_cap_v2":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def space_timesteps(num_timesteps, section_counts):
    """respace.py:12-65."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            if desired_count == 1:
                return set([50])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx = 0.0
        for _ in range(section_count):
            all_steps.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        start_idx += size
    return set(all_steps)


class GaussianDiffusion:
    def __init__(self, *, betas):
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self._dev_tables = {}

    def _table(self, name, device):
        key = (name, str(device))
        t = self._dev_tables.get(key)
        if t is None:
            t = torch.from_numpy(getattr(self, name)).to(device=device).float()
            self._dev_tables[key] = t
        return t

    def _extract(self, name, timesteps, ndim):
        res = self._table(name, timesteps.device)[timesteps]
        while res.dim() < ndim:
            res = res[..., None]
        return res

    def q_sample(self, x_start, t, noise=None):
        """gaussian_diffusion.py:215-230."""
        if noise is None:
            noise = torch.randn_like(x_start)
        return (self._extract("sqrt_alphas_cumprod", t, x_start.dim()) * x_start
                + self._extract("sqrt_one_minus_alphas_cumprod", t, x_start.dim()) * noise)

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (self._extract("sqrt_recip_alphas_cumprod", t, x_t.dim()) * x_t
                - self._extract("sqrt_recipm1_alphas_cumprod", t, x_t.dim()) * eps)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return ((self._extract("sqrt_recip_alphas_cumprod", t, x_t.dim()) * x_t - pred_xstart)
                / self._extract("sqrt_recipm1_alphas_cumprod", t, x_t.dim()))


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:67-116: keep `use_timesteps` of a base process; the wrapped model sees original timestep ids."""

    def __init__(self, use_timesteps, betas):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(betas)
        base = GaussianDiffusion(betas=betas)
        last_alpha_cumprod = 1.0
        new_betas = []
        for i, alpha_cumprod in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - alpha_cumprod / last_alpha_cumprod)
                last_alpha_cumprod = alpha_cumprod
                self.timestep_map.append(i)
        super().__init__(betas=np.array(new_betas))
        self._map_dev = {}

    def _map(self, device):
        m = self._map_dev.get(str(device))
        if m is None:
            m = torch.tensor(self.timestep_map, device=device, dtype=torch.long)
            self._map_dev[str(device)] = m
        return m

    @torch.no_grad()
    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=False, model_kwargs=None, device=None,
                         progress=False, eta=0.0):
        """gaussian_diffusion.py:609-689 + ddim_sample :522-569 + p_mean_variance :255-341 (epsilon, eta = 0)."""
        assert eta == 0.0 and not clip_denoised, "DreamVLA samples with eta=0, clip_denoised=False (dreamvla_model.py:966-974)"
        model_kwargs = model_kwargs or {}
        img = noise if noise is not None else torch.randn(*shape, device=device)
        device = img.device
        n = shape[0]
        tmap = self._map(device)
        for i in range(self.num_timesteps - 1, -1, -1):
            t = torch.full((n,), i, device=device, dtype=torch.long)
            model_output = model(img, tmap[t], **model_kwargs)
            pred_xstart = self._predict_xstart_from_eps(img, t, model_output)
            eps = self._predict_eps_from_xstart(img, t, pred_xstart)
            alpha_bar_prev = self._extract("alphas_cumprod_prev", t, img.dim())
            img = pred_xstart * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev) * eps
        return img


class FMDiffusion(GaussianDiffusion):
    """Flow-matching sampler of `--use_fm` (reference respace.py:118-191): NOT a diffusion process -- the betas only fix
    `num_timesteps` -- but an explicit Euler integration of the predicted velocity over t = 0, 1/T, ..., (T-1)/T."""

    @torch.no_grad()
    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=False, model_kwargs=None, device=None,
                         progress=False, eta=0.0, start=None):
        """respace.py:122-156.  As in the reference the guidance scale is forced to 1.0 and the `noise` argument is IGNORED:
        the start state is drawn here, [shape] fp32 (`start=` injects it for parity tests); `final += (1/T) * u_t`."""
        model_kwargs = dict(model_kwargs or {})
        if "cfg_scale" in model_kwargs:
            model_kwargs["cfg_scale"] = 1.0                                              # :136-138
        final = start.to(device=device, dtype=torch.float32) if start is not None else torch.randn(*shape, device=device)
        delta = 1.0 / self.num_timesteps
        for i in range(self.num_timesteps):
            t = torch.full((shape[0],), float(i), device=final.device) / self.num_timesteps       # :144-145
            ut = model(final, t, **model_kwargs)                                         # p_sample -> p_mean_variance :158-172
            final = final + delta * ut
        return final


def create_diffusion(timestep_respacing, noise_schedule="linear", diffusion_steps=1000, **_unused):
    """action_model/__init__.py:10-46 (epsilon / fixed-small / MSE configuration)."""
    betas = get_named_beta_schedule(noise_schedule, diffusion_steps)
    if timestep_respacing is None or timestep_respacing == "":
        timestep_respacing = [diffusion_steps]
    return SpacedDiffusion(use_timesteps=space_timesteps(diffusion_steps, timestep_respacing), betas=betas)
