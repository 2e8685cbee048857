"""Shared test helpers (test infrastructure: may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import unet_oracle as O  # noqa: E402


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def cfg_kwargs(cfg):
    return dict(cfg)


class FakeWrapper(torch.nn.Module):
    """DiffusionWrapper('hybrid') stand-in (ddpm.py:1327-1351)."""

    def __init__(self, unet):
        super().__init__()
        self.diffusion_model = unet
        self.conditioning_key = "hybrid"

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        xc = torch.cat([x] + c_concat, dim=1)
        cc = torch.cat(c_crossattn, 1)
        return self.diffusion_model(xc, t, context=cc)


class FakeLDM:
    """What DDIMSampler needs from LatentDiffusion: schedule buffers (ddpm.py:149-170) + apply_model (:865-880)."""

    def __init__(self, unet, device):
        ac = O.make_alphas_cumprod()
        betas = np.linspace(0.00085 ** 0.5, 0.0120 ** 0.5, 1000, dtype=np.float64) ** 2
        self.num_timesteps = 1000
        self.betas = torch.tensor(betas, dtype=torch.float32, device=device)
        self.alphas_cumprod = torch.tensor(ac, dtype=torch.float32, device=device)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32, device=device)
        self.parameterization = "eps"
        self.device = device
        self.model = FakeWrapper(unet)

    def apply_model(self, x_noisy, t, cond):
        return self.model(x_noisy, t, **cond)


def err_stats(got, ref):
    got = torch.as_tensor(got).float().cpu()
    ref = torch.as_tensor(ref).float().cpu()
    d = (got - ref).abs()
    return dict(max_abs=d.max().item(), ref_max=ref.abs().max().item(), ref_rms=ref.pow(2).mean().sqrt().item(),
                rms=d.pow(2).mean().sqrt().item(),
                rel_rms=(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(),
                finite=bool(torch.isfinite(got).all()))


def synthetic_inputs(batch, h=64, w=128, ctx_dim=1024, L=77, seed=1234, device="cpu"):
    """SURVEY §8d synthetic workload: x_T, c_concat (mask: left half 0, right half a random rectangle; masked-image
    latent), context and a repeated unconditional context."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, 4, h, w, generator=g)
    mask = torch.zeros(batch, 1, h, w)
    for b in range(batch):
        rh = int(torch.randint(h // 2, 3 * h // 4, (1,), generator=g))
        rw = int(torch.randint(w // 4, 3 * w // 8, (1,), generator=g))
        y0 = int(torch.randint(0, h - rh + 1, (1,), generator=g))
        x0 = w // 2 + int(torch.randint(0, w // 2 - rw + 1, (1,), generator=g))
        mask[b, :, y0:y0 + rh, x0:x0 + rw] = 1.0
    lat = 0.18215 * torch.randn(batch, 4, h, w, generator=g) * (1.0 - mask)
    c_concat = torch.cat([mask, lat], dim=1)
    ctx = torch.randn(batch, L, ctx_dim, generator=g)
    uc = torch.randn(1, L, ctx_dim, generator=g).repeat(batch, 1, 1)
    return [t.to(device) for t in (x, c_concat, ctx, uc)]
