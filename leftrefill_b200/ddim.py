"""Drop-in `DDIMSampler` (reference ldm/models/diffusion/ddim.py: make_schedule :23-52, sample :54-144,
ddim_sampling :224-302, p_sample_ddim :304-386).

Same constructor / `sample` signature and return value `(samples, {'x_inter': [...], 'pred_x0': [...]})`, same RNG
consumption (one `torch.randn` for x_T if not given, one per step even when sigma == 0), so callers such as
inpainting_ldm/ref_inpainting_ldm.py:74-81 and ref_inpainting_gradio.py:119-129 run unchanged.

What changes is where the work happens:
  * the CFG combine + x0 prediction + x_{t-1} update is ONE fused native kernel (lr_ddim_update), fp32 state;
  * when the wrapped model is a `leftrefill_b200.UNetModel` under 'hybrid' conditioning, the step-invariant work is
    hoisted out of the loop: cat(uncond, cond) context is built once and its cross-attention K/V cached in the
    engine, c_concat is staged once, and the UNet runs the CFG pair as one batched native forward per step.
  * on that fast path the per-step work (stage x_t into the 9-channel input, UNet CFG forward, fused update) is
    captured ONCE as a CUDA graph whose step-dependent scalars (timestep, a_t, a_prev, sigma_t, ...) live in device
    memory, and replayed for every step: ~410 kernel launches per step cost the host one graph launch. The graph is
    cached on the UNet module (keyed by shapes and the engine's plan generation), so later `sample` calls reuse it.
    `LR_NO_CUDA_GRAPH=1` keeps the eager loop (bit-identical results; the kernels and their order are the same).
Anything else (other models / conditioning layouts) goes through `model.apply_model` exactly like the reference.
"""
import os

import numpy as np
import torch

from . import ops
from .unet import UNetModel


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:63-74. `alphacums` is the float32 alphas_cumprod of the model."""
    ac = np.asarray(alphacums, dtype=np.float32)
    alphas = ac[ddim_timesteps].astype(np.float64)
    alphas_prev = np.asarray([ac[0]] + ac[ddim_timesteps[:-1]].tolist(), dtype=np.float64)
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule for ddim "
              f"sampler {sigmas}")
    return sigmas, alphas, alphas_prev


class _StepGraph:
    """One captured CUDA graph of a whole DDIM step on the native fast path, replayed for every step.

    Static device buffers (graph inputs / outputs): x (the fp32 latent state, updated in place), xc (the UNet input
    with c_concat pre-staged), t (int64 timesteps), coef (cfg, a_t, a_prev, sigma_t, sqrt(1 - a_t)), noise, pred_x0.
    """

    def __init__(self, unet, nb, b, cx, c_total, H, W, pair, use_cfg, temperature, device):
        from . import _native as N
        self.unet, self.nb, self.b, self.pair, self.use_cfg = unet, nb, b, pair, use_cfg
        self.temperature = float(temperature)
        self.x = torch.zeros(b, cx, H, W, dtype=torch.float32, device=device)
        self.xc = torch.zeros(nb, c_total, H, W, dtype=torch.float32, device=device)
        self.t = torch.full((nb,), 1, dtype=torch.long, device=device)
        self.coef = torch.tensor([1.0, 0.5, 0.6, 0.1, 0.5 ** 0.5], dtype=torch.float32, device=device)
        self.noise = torch.zeros_like(self.x)
        self.pred_x0 = torch.zeros_like(self.x)
        self.graph = None
        self.device = device
        # eager warm-up on a side stream: builds the engine's plan (device allocations are illegal during capture)
        with torch.cuda.device(device):
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                self._body()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self.generation = N.lib().lr_unet_plan_generation(unet.engine())
            graph = torch.cuda.CUDAGraph()
            before = N.lib().lr_launch_count()
            with torch.cuda.graph(graph):
                self._body()
        self.kernels = N.lib().lr_launch_count() - before  # native kernel nodes of one replay
        N.lib().lr_launch_count_add(-self.kernels)          # the capture itself executed nothing
        self.graph = graph

    def replay(self):
        from . import _native as N
        self.graph.replay()
        N.lib().lr_launch_count_add(self.kernels)

    def _body(self):
        cx = self.x.shape[1]
        if self.pair and self.use_cfg:
            self.xc[:, :cx] = self.x
            e = self.unet.forward_native_cfg_pair(self.xc, self.t)
        else:
            self.xc[:self.b, :cx] = self.x
            if self.use_cfg:
                self.xc[self.b:, :cx] = self.x
            e = self.unet.forward_native(self.xc, self.t, None)
        e_u, e_c = (e[:self.b], e[self.b:]) if self.use_cfg else (e, None)
        ops.ddim_update_dev(self.x, e_u, e_c, self.noise, self.coef, self.temperature, self.x, self.pred_x0)

    def valid(self):
        from . import _native as N
        return self.graph is not None and N.lib().lr_unet_plan_generation(self.unet.engine()) == self.generation


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.noise_source = None  # optional callable(shape, device, step) -> noise (tests inject recorded noise)

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        ac = alphas_cumprod.detach().float().cpu()
        dev = self._device()
        to_dev = lambda t: t.clone().detach().to(torch.float32).to(dev)
        self.register_buffer("betas", to_dev(self.model.betas))
        self.register_buffer("alphas_cumprod", to_dev(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_dev(self.model.alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_dev(ac.sqrt()))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_dev((1. - ac).sqrt()))
        self.register_buffer("log_one_minus_alphas_cumprod", to_dev((1. - ac).log()))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_dev((1. / ac).sqrt()))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_dev((1. / ac - 1).sqrt()))
        sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(ac.numpy(), self.ddim_timesteps, ddim_eta, verbose)
        self.register_buffer("ddim_sigmas", sigmas)
        self.register_buffer("ddim_alphas", alphas)
        self.register_buffer("ddim_alphas_prev", alphas_prev)
        self.register_buffer("ddim_sqrt_one_minus_alphas", np.sqrt(1. - alphas))
        acp = self.alphas_cumprod_prev
        self.register_buffer("ddim_sigmas_for_original_num_steps", ddim_eta * torch.sqrt(
            (1 - acp) / (1 - self.alphas_cumprod) * (1 - self.alphas_cumprod / acp)))

    def _device(self):
        return self.model.betas.device

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None, **kwargs):
        if conditioning is not None and isinstance(conditioning, dict):
            ctmp = conditioning[list(conditioning.keys())[0]]
            while isinstance(ctmp, list):
                ctmp = ctmp[0]
            if ctmp.shape[0] != batch_size:
                print(f"Warning: Got {ctmp.shape[0]} conditionings but batch-size is {batch_size}")
        if kwargs.get("return_attn", False):
            raise NotImplementedError("return_attn is a debugging feature of the reference and is not supported")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        if type(conditioning) == list:  # noqa: E721 - same dispatch as the reference (ddim.py:104)
            return self.ddim_multi_sampling(conditioning, (batch_size, C, H, W), callback=callback,
                                            img_callback=img_callback, quantize_denoised=quantize_x0, mask=mask, x0=x0,
                                            ddim_use_original_steps=False, noise_dropout=noise_dropout,
                                            temperature=temperature, score_corrector=score_corrector,
                                            corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                            unconditional_guidance_scale=unconditional_guidance_scale,
                                            unconditional_conditioning=unconditional_conditioning,
                                            dynamic_threshold=dynamic_threshold, ucg_schedule=ucg_schedule)
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning,
                                  dynamic_threshold=dynamic_threshold, ucg_schedule=ucg_schedule)

    # ------------------------------------------------------------------------------------------------------------
    def _native_unet(self, cond, ucond, cfg_scale):
        """Returns the leftrefill_b200.UNetModel behind model.apply_model when the hoisted fast path applies."""
        wrapper = getattr(self.model, "model", None)
        unet = getattr(wrapper, "diffusion_model", None)
        if not isinstance(unet, UNetModel) or getattr(wrapper, "conditioning_key", None) != "hybrid":
            return None
        if type(unet).forward is not UNetModel.forward:
            return None  # a subclass overrides forward (e.g. the reference NVSUnetModel after install()): honour it
        if getattr(self.model, "parameterization", "eps") != "eps":
            return None

        def ok(c):
            return (isinstance(c, dict) and set(c.keys()) == {"c_concat", "c_crossattn"} and
                    all(isinstance(c[k], list) and len(c[k]) == 1 for k in c))

        if not ok(cond):
            return None
        if ucond is not None and cfg_scale != 1. and not ok(ucond):
            return None
        return unet

    @torch.no_grad()
    def ddim_multi_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                            quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                            temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                            unconditional_guidance_scale=1., unconditional_conditioning=None, dynamic_threshold=None,
                            ucg_schedule=None, **kwargs):
        """Conditionally-consistent multi-reference sampling (reference ddim.py:146-222): `cond` /
        `unconditional_conditioning` / `x_T` are lists with one entry per reference view; every step denoises each
        stitched [reference | target] canvas on its own, then ONE of the predicted right (target) halves - chosen with
        `random.shuffle`, exactly like the reference - is written into all canvases. Returns (first canvas, {}).
        Each per-view step is `p_sample_ddim`: the native UNet through `model.apply_model` + the fused update kernel."""
        import random
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is never used by the LeftRefill drivers")
        device = self._device()
        b = shape[0]
        img = [torch.randn(shape, device=device)] * len(cond) if x_T is None else list(x_T)
        if unconditional_conditioning is None:
            unconditional_conditioning = [None] * len(cond)
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if ucg_schedule is not None:
                assert len(ucg_schedule) == len(time_range)
                unconditional_guidance_scale = ucg_schedule[i]
            pred_img_right, new_img = [], []
            for img_, cond_, uc_ in zip(img, cond, unconditional_conditioning):
                img_, _ = self.p_sample_ddim(img_, cond_, ts, index=index, quantize_denoised=quantize_denoised,
                                             temperature=temperature, noise_dropout=noise_dropout,
                                             score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                             unconditional_guidance_scale=unconditional_guidance_scale,
                                             unconditional_conditioning=uc_, dynamic_threshold=dynamic_threshold)
                pred_img_right.append(img_[:, :, :, img_.shape[-1] // 2:])
                new_img.append(img_)
            random.shuffle(pred_img_right)  # the reference picks a random view's target half (ddim.py:205-207)
            random_right = pred_img_right[0].clone()
            for j in range(len(new_img)):
                new_img[j][:, :, :, random_right.shape[-1]:] = random_right
            img = new_img
            if callback:
                callback(i)
        return img[0], intermediates

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, dynamic_threshold=None, ucg_schedule=None, **kwargs):
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is never used by the LeftRefill drivers")
        if quantize_denoised or score_corrector is not None or dynamic_threshold is not None:
            raise NotImplementedError("quantize_denoised / score_corrector / dynamic_threshold are not supported")
        device = self._device()
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]

        use_cfg = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        unet = None
        if ucg_schedule is None:
            unet = self._native_unet(cond, unconditional_conditioning, unconditional_guidance_scale)
        staged = None
        if unet is not None:
            # ---- step-invariant hoisting (the reference rebuilds all of this every step, ddim.py:317-333) ----
            cc = cond["c_crossattn"][0]
            c_cat = cond["c_concat"][0].float()
            if use_cfg:
                cc = torch.cat([unconditional_conditioning["c_crossattn"][0], cc])
                c_cat = torch.cat([unconditional_conditioning["c_concat"][0].float(), c_cat])
            # both CFG halves see the same x, t and (in every LeftRefill driver) the same c_concat: then everything
            # before the first cross-attention is computed once (lr_unet_forward_cfg_pair)
            pair = (use_cfg and unet.view_num == 1 and not unet.use_sep and
                    torch.equal(unconditional_conditioning["c_concat"][0], cond["c_concat"][0]))
            if pair:
                c_cat = cond["c_concat"][0].float()
            nb = c_cat.shape[0]
            if getattr(unet, "_cin_active", False):  # a c_input staged by an earlier forward(c_input=...) call
                unet.set_c_input(None, shape[3])
                unet._cin_active = False
            unet.set_context(cc)
            xc = torch.empty(nb, img.shape[1] + c_cat.shape[1], shape[2], shape[3], dtype=torch.float32, device=device)
            xc[:, img.shape[1]:] = c_cat
            staged = (xc, nb, pair)
            if noise_dropout == 0. and os.environ.get("LR_NO_CUDA_GRAPH") is None and torch.cuda.is_available():
                sg = self._step_graph(unet, nb, b, img.shape[1], xc.shape[1], shape[2], shape[3], pair, use_cfg,
                                      temperature, device, cc.shape[1])
                if sg is not None:
                    return self._graphed_loop(sg, img, c_cat, time_range, total_steps, mask, x0, callback,
                                              img_callback, log_every_t, unconditional_guidance_scale, intermediates)

        x = img.float()
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            if mask is not None:
                assert x0 is not None
                ts = torch.full((b,), int(step), device=device, dtype=torch.long)
                img_orig = self.model.q_sample(x0, ts)
                x = img_orig * mask + (1. - mask) * x
            if ucg_schedule is not None:
                assert len(ucg_schedule) == len(time_range)
                unconditional_guidance_scale = ucg_schedule[i]
                use_cfg = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
            if staged is not None:
                xc, nb, pair = staged
                t_in = torch.full((nb,), int(step), device=device, dtype=torch.long)
                if pair and use_cfg:
                    xc[:, :x.shape[1]] = x
                    e = unet.forward_native_cfg_pair(xc, t_in)
                else:
                    xc[:, :x.shape[1]] = torch.cat([x, x]) if use_cfg else x
                    e = unet.forward_native(xc, t_in, None)
                e_u, e_c = (e[:b], e[b:]) if use_cfg else (e, None)
            else:
                e_u, e_c = self._apply_model(x, cond, int(step), unconditional_conditioning, use_cfg)
            x, pred_x0 = self._update(x, e_u, e_c, index, unconditional_guidance_scale, temperature, noise_dropout, i)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(x)
                intermediates["pred_x0"].append(pred_x0)
        return x, intermediates

    # ------------------------------------------------------------------------------------------------------------
    def _step_graph(self, unet, nb, b, cx, c_total, H, W, pair, use_cfg, temperature, device, ctx_len=0):
        """Returns the cached (or freshly captured) step graph for this configuration, None if capture is unavailable.
        The context length is part of the key: the cached cross-attention K/V buffers (and the attention launch
        geometry baked into the graph) depend on it; the engine also bumps its plan generation when they are
        reallocated, which `_StepGraph.valid` checks."""
        key = (nb, b, cx, c_total, H, W, bool(pair), bool(use_cfg), float(temperature), str(device), int(ctx_len))
        cache = unet.__dict__.setdefault("_step_graphs", {})
        sg = cache.get(key)
        if sg is not None and sg is not False and sg.valid():
            return sg
        if sg is False:
            return None
        try:
            sg = _StepGraph(unet, nb, b, cx, c_total, H, W, pair, use_cfg, temperature, device)
        except Exception as e:  # noqa: BLE001 - capture is an optimisation: fall back to the eager loop, say so once
            print(f"leftrefill_b200: CUDA graph capture of the DDIM step failed ({str(e)[:200]}); using the eager loop")
            torch.cuda.synchronize(device)
            cache[key] = False
            return None
        cache[key] = sg
        return sg

    def _graphed_loop(self, sg, img, c_cat, time_range, total_steps, mask, x0, callback, img_callback, log_every_t,
                      cfg_scale, intermediates):
        """ddim_sampling's loop (ddim.py:253-296) with the whole step as one graph replay. RNG consumption is the
        reference's: one normal draw of x's shape per step, after the model call of that step."""
        device = sg.x.device
        cx = sg.x.shape[1]
        sg.xc[:, cx:] = c_cat
        sg.x.copy_(img.float())
        idx = [total_steps - i - 1 for i in range(total_steps)]
        table = np.stack([np.full(total_steps, cfg_scale, dtype=np.float64), self.ddim_alphas[idx],
                          self.ddim_alphas_prev[idx], self.ddim_sigmas[idx], self.ddim_sqrt_one_minus_alphas[idx]],
                         axis=1).astype(np.float32)
        coef_table = torch.from_numpy(table).to(device)
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            if mask is not None:
                assert x0 is not None
                ts = torch.full((sg.b,), int(step), device=device, dtype=torch.long)
                img_orig = self.model.q_sample(x0, ts)
                sg.x.copy_(img_orig * mask + (1. - mask) * sg.x)
            sg.t.fill_(int(step))
            sg.coef.copy_(coef_table[i])
            if self.noise_source is not None:
                sg.noise.copy_(self.noise_source(sg.x.shape, device, i))
            else:
                sg.noise.normal_()
            sg.replay()
            if callback:
                callback(i)
            if img_callback:
                img_callback(sg.pred_x0.clone(), i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(sg.x.clone())
                intermediates["pred_x0"].append(sg.pred_x0.clone())
        return sg.x.clone(), intermediates

    def _apply_model(self, x, c, step, uc, use_cfg):
        """Generic path: the reference's CFG batching around model.apply_model (ddim.py:311-343)."""
        b, device = x.shape[0], x.device
        t = torch.full((b,), step, device=device, dtype=torch.long)
        if not use_cfg:
            return self.model.apply_model(x, t, c).float().contiguous(), None
        x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
        if isinstance(c, dict):
            assert isinstance(uc, dict)
            c_in = {}
            for k in c:
                if isinstance(c[k], list):
                    c_in[k] = [torch.cat([uc[k][i], c[k][i]]) for i in range(len(c[k]))]
                else:
                    c_in[k] = torch.cat([uc[k], c[k]])
        else:
            c_in = torch.cat([uc, c])
        e = self.model.apply_model(x_in, t_in, c_in).float().contiguous()
        return e[:b], e[b:]

    def _update(self, x, e_u, e_c, index, cfg_scale, temperature, noise_dropout, step_i):
        if getattr(self.model, "parameterization", "eps") != "eps":
            raise NotImplementedError("only the eps parameterisation (SD2-inpainting) is supported")
        a_t, a_prev = float(self.ddim_alphas[index]), float(self.ddim_alphas_prev[index])
        sigma_t, s1m = float(self.ddim_sigmas[index]), float(self.ddim_sqrt_one_minus_alphas[index])
        # the reference draws noise every step, even when sigma_t == 0 (ddim.py:378): keep the RNG stream identical
        if self.noise_source is not None:
            noise = self.noise_source(x.shape, x.device, step_i)
        else:
            noise = torch.randn(x.shape, device=x.device)
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        return ops.ddim_update(x.contiguous(), e_u.contiguous(), None if e_c is None else e_c.contiguous(),
                               noise.float().contiguous(), cfg_scale, a_t, a_prev, sigma_t, s1m, temperature)

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, dynamic_threshold=None,
                      **kwargs):
        """One reference-compatible step: returns (x_prev, pred_x0)."""
        if use_original_steps or quantize_denoised or score_corrector is not None or dynamic_threshold is not None \
                or repeat_noise:
            raise NotImplementedError("unsupported p_sample_ddim option")
        use_cfg = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
        e_u, e_c = self._apply_model(x.float(), c, int(t[0]), unconditional_conditioning, use_cfg)
        return self._update(x.float(), e_u, e_c, index, unconditional_guidance_scale, temperature, noise_dropout, 0)
