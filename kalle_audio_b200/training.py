"""Training step of the sigmaVAE / Oobleck autoencoder (BASELINE config 5).

Reference: ``AutoencoderTrainingWrapper.training_step`` (stable_audio_tools/training/autoencoders.py:221-352):
``encode(return_info=True)`` -> bottleneck -> ``decode`` -> losses -> ``manual_backward`` -> ``opt_gen.step()``, with
the KL term wired as ``ValueLoss(key='kl')`` (:446-456) and the sample/KL arithmetic of ``vae_sample``
(models/bottleneck.py:51-62).  The reconstruction term here is the sigma-VAE Gaussian negative log-likelihood
(BASELINE.json config 5; no in-tree definition in the reference):

    nll = sum_{c,t} [ 0.5 ((x - x_hat)/sigma)^2 + log sigma + 0.5 log 2 pi ]   averaged over the batch
    loss = nll + kl_weight * kl

Everything differentiable runs in libkvae: the encoder/decoder stacks are ONE autograd node each
(``_plan.PlanFunction``), ``vae_sample`` and the NLL are nodes of their own, so ``loss.backward()`` is the
reference's ``manual_backward`` and any torch optimizer works on the module's parameters.  ``AutoencoderTrainer``
adds what a data-parallel job needs on top: parameters flattened into one fp32 master buffer per direction,
gradients produced directly as flat buffers, an NCCL all-reduce per direction launched the moment that
direction's backward finishes (the decoder's overlaps the encoder's backward) and a fused AdamW step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib


# ----------------------------------------------------------------------------------------- autograd nodes
class _VaeSampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, scale, noise):
        dt = mean.dtype if mean.dtype in (torch.float32, torch.bfloat16) else torch.float32
        m, s, n = (t.detach().to(dt).contiguous() for t in (mean, scale, noise))
        out = torch.empty_like(m)
        kl = torch.empty((), dtype=torch.float32, device=m.device)
        scratch = torch.empty(8 * 1024, dtype=torch.uint8, device=m.device)
        B, D, T = m.shape
        _lib.check(_lib.lib().kvae_vae_sample(m.data_ptr(), s.data_ptr(), n.data_ptr(), out.data_ptr(), kl.data_ptr(),
                                              B, D, T, _lib.dtype_code(dt), scratch.data_ptr(),
                                              _lib.stream_ptr(m.device)))
        ctx.save_for_backward(m, s, n)
        ctx.in_dtype = mean.dtype
        return out.to(mean.dtype), kl

    @staticmethod
    def backward(ctx, gz, gkl):
        m, s, n = ctx.saved_tensors
        B, D, T = m.shape
        gz_c = None if gz is None else gz.detach().to(m.dtype).contiguous()
        gkl_c = None if gkl is None else gkl.detach().float().contiguous()
        gm, gs = torch.empty_like(m), torch.empty_like(m)
        _lib.check(_lib.lib().kvae_vae_sample_bwd(m.data_ptr(), s.data_ptr(), n.data_ptr(), _lib.ptr(gz_c),
                                                  _lib.ptr(gkl_c), gm.data_ptr(), gs.data_ptr(), B, D, T,
                                                  _lib.dtype_code(m.dtype), _lib.stream_ptr(m.device)))
        return gm.to(ctx.in_dtype), gs.to(ctx.in_dtype), None


def vae_sample_with_grad(mean: torch.Tensor, scale: torch.Tensor, noise: Optional[torch.Tensor] = None):
    """Differentiable ``vae_sample`` (bottleneck.py:51-62): (latents, kl)."""
    _lib.require_cuda(mean, "vae_sample")
    if mean.shape != scale.shape or mean.dim() != 3:
        raise ValueError("mean and scale must both be [B, D, T]")
    if noise is None:
        noise = torch.randn_like(mean)
    return _VaeSampleFn.apply(mean, scale, noise)


class _GaussianNllFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, xhat, log_sigma):
        dt = xhat.dtype if xhat.dtype in (torch.float32, torch.bfloat16) else torch.float32
        xc, hc = x.detach().to(dt).contiguous(), xhat.detach().to(dt).contiguous()
        B = xc.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=xc.device)
        g = torch.empty_like(hc)
        scratch = torch.empty(8 * 1024, dtype=torch.uint8, device=xc.device)
        _lib.check(_lib.lib().kvae_gaussian_nll(xc.data_ptr(), hc.data_ptr(), g.data_ptr(), loss.data_ptr(), B,
                                                xc[0].numel(), float(log_sigma), _lib.dtype_code(dt),
                                                scratch.data_ptr(), _lib.stream_ptr(xc.device)))
        ctx.save_for_backward(g)
        ctx.in_dtype = xhat.dtype
        return loss

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return None, (g * gl.to(g.dtype)).to(ctx.in_dtype), None


def gaussian_nll(x: torch.Tensor, xhat: torch.Tensor, log_sigma: float = 0.0) -> torch.Tensor:
    """sigma-VAE reconstruction term with a fixed scalar sigma = exp(log_sigma): per-clip sum, batch mean."""
    _lib.require_cuda(xhat, "gaussian_nll")
    if x.shape != xhat.shape:
        raise ValueError("x and xhat must have the same shape")
    return _GaussianNllFn.apply(x, xhat, log_sigma)


# ----------------------------------------------------------------------------------------- flat buffers
def flatten_parameters(module: nn.Module) -> torch.Tensor:
    """Re-homes every parameter of ``module`` as a view of ONE contiguous fp32 buffer (in ``parameters()`` order)
    and returns that buffer.  State-dict keys, shapes and values are unchanged."""
    params = list(module.parameters())
    if not params:
        raise ValueError("module has no parameters")
    flat = torch.cat([p.detach().reshape(-1).float() for p in params])
    off = 0
    for p in params:
        n = p.numel()
        p.data = flat[off:off + n].view(p.shape)
        off += n
    return flat


class GradSync:
    """Sum-all-reduce of flat gradient buffers over a process group, launched asynchronously so that a later
    backward overlaps it.  Without an initialised process group (single GPU) it is a no-op."""

    def __init__(self, process_group=None):
        self.group = process_group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.world = dist.get_world_size(process_group) if self.enabled else 1
        self._pending: List = []

    def launch(self, flat_grads: torch.Tensor) -> None:
        if self.enabled:
            self._pending.append(dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending.clear()

    def broadcast(self, tensors: List[torch.Tensor], src: int = 0) -> None:
        """Every rank's ``tensors`` become rank ``src``'s (in place)."""
        if self.enabled:
            for t in tensors:
                dist.broadcast(t, src, group=self.group)

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world


class FlatAdamW:
    """torch.optim.AdamW semantics over flat fp32 buffers, one kernel per buffer (kvae_adamw_step)."""

    def __init__(self, flats: List[torch.Tensor], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 modules: Optional[List[nn.Module]] = None):
        self.flats = flats
        # modules whose parameters are views of ``flats``: every plan of theirs (any precision / device key) must
        # re-pack after a step.  The kernel writes through the flat buffer, which does NOT bump the version counters
        # of the parameter views, so the plans' (data_ptr, _version) fingerprint cannot see it: bump an explicit
        # weights epoch that the fingerprint includes (PlanRunner.sync_weights).
        self.modules = list(modules or [])
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = [torch.zeros_like(f) for f in flats]
        self.exp_avg_sq = [torch.zeros_like(f) for f in flats]
        self.step_count = 0

    def step(self, grads: List[torch.Tensor], grad_scale: float = 1.0) -> None:
        self.step_count += 1
        L = _lib.lib()
        for f, g, m, v in zip(self.flats, grads, self.exp_avg, self.exp_avg_sq):
            if g.numel() != f.numel():
                raise ValueError("gradient buffer does not match the parameter buffer")
            _lib.check(L.kvae_adamw_step(f.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), f.numel(), self.lr,
                                         self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                                         grad_scale, _lib.stream_ptr(f.device)))
        for m in self.modules:
            m._weights_epoch = getattr(m, "_weights_epoch", 0) + 1

    def state(self) -> List[torch.Tensor]:
        return self.exp_avg + self.exp_avg_sq


class AutoencoderTrainer:
    """Data-parallel training step for an ``AudioAutoencoder`` with Oobleck encoder/decoder.

    ``training_step(reals)`` = encode -> split mean/scale -> vae_sample -> decode -> Gaussian NLL + kl_weight * KL
    (+ spectral_weight * multi-resolution STFT loss when ``spectral_loss`` is given) -> backward -> gradient
    all-reduce -> AdamW, the generator branch of the reference's ``training_step``.  With ``discriminator`` (an
    ``OobleckDiscriminator``) the step alternates the way training/autoencoders.py:288-337 does once warmed up: odd steps
    train the discriminator on ``loss_dis``, even steps train the autoencoder with ``adversarial_weight * loss_adv +
    feature_matching_weight * feature_matching_distance`` added to its loss.
    One process per GPU; pass the process group (NCCL over NVLink) or leave ``None`` for the default group /
    single-GPU operation."""

    def __init__(self, autoencoder: nn.Module, lr: float = 1e-4, betas=(0.8, 0.99), eps: float = 1e-8,
                 weight_decay: float = 1e-3, kl_weight: float = 1e-6, log_sigma: float = 0.0,
                 precision: Optional[str] = "bf16", process_group=None, data_parallel: bool = True,
                 spectral_loss: Optional[nn.Module] = None, spectral_weight: float = 1.0, nll_weight: float = 1.0,
                 discriminator: Optional[nn.Module] = None, adversarial_weight: float = 0.1,
                 feature_matching_weight: float = 5.0, disc_lr: Optional[float] = None, warmup_steps: int = 0):
        self.autoencoder = autoencoder
        self.discriminator, self.adversarial_weight = discriminator, adversarial_weight
        self.feature_matching_weight, self.warmup_steps, self.global_step = feature_matching_weight, warmup_steps, 0
        # the reference's generator loss is MR-STFT + adversarial + feature matching + KL (training/autoencoders.py:
        # 150-200); ``spectral_loss`` (kalle_audio_b200.SumAndDifferenceSTFTLoss / MultiResolutionSTFTLoss, called as
        # module(reals, decoded) like the reference's AuralossLoss) adds its spectral term, ``nll_weight=0`` drops the
        # Gaussian NLL of BASELINE config 5
        self.spectral_loss, self.spectral_weight, self.nll_weight = spectral_loss, spectral_weight, nll_weight
        self.encoder, self.decoder = autoencoder.encoder, autoencoder.decoder
        dev = next(self.encoder.parameters()).device
        if dev.type != "cuda":
            raise _lib.KvaeError("AutoencoderTrainer needs the autoencoder on a CUDA device (no CPU path)")
        for m in (self.encoder, self.decoder):
            m.set_precision(precision)
        self.flat_enc = flatten_parameters(self.encoder)
        self.flat_dec = flatten_parameters(self.decoder)
        self.kl_weight, self.log_sigma = kl_weight, log_sigma
        self.sync = GradSync(process_group)
        if not data_parallel:       # a single-process trainer inside an initialised process group
            self.sync.enabled, self.sync.world = False, 1
        self.opt = FlatAdamW([self.flat_enc, self.flat_dec], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                             modules=[self.encoder, self.decoder])
        if discriminator is not None:
            # the discriminator has its own optimizer in the reference (optimizer_configs['discriminator'])
            self.flat_disc = flatten_parameters(discriminator)
            self.opt_disc = FlatAdamW([self.flat_disc], lr=lr if disc_lr is None else disc_lr, betas=betas, eps=eps,
                                      weight_decay=weight_decay)
        self._grads: Dict[str, torch.Tensor] = {}
        for name, mod, flat in (("enc", self.encoder, self.flat_enc), ("dec", self.decoder, self.flat_dec)):
            # applied to every runner the module creates from now on (another precision, a rebuilt cache after .to()):
            # the flat master buffer, the all-reduce hook, and no per-parameter gradient views -- the trainer consumes
            # the flat gradient buffer, and views handed to autograd would be cloned by AccumulateGrad while NCCL
            # reduces the buffer in place
            mod._runner_init = self._make_runner_init(name, flat)
            mod._plans.clear()
            mod.runner(dev)
        self.broadcast_parameters()

    def _make_runner_init(self, name, flat):
        def init(runner):
            runner.flat_master = flat
            runner.grads_ready_hook = self._make_hook(name)
            runner.return_param_grads = False
        return init

    def broadcast_parameters(self, src: int = 0) -> None:
        """Every rank continues from rank ``src``'s parameters and optimizer moments (what DDP / Lightning do at
        construction and on resume in the reference's trainer): replicas that were seeded or loaded differently would
        otherwise diverge silently, since only gradients are exchanged afterwards."""
        if not self.sync.enabled:
            return
        step = torch.tensor([self.opt.step_count], device=self.flat_enc.device, dtype=torch.int64)
        self.sync.broadcast([self.flat_enc, self.flat_dec] + self.opt.state() + [step], src)
        self.opt.step_count = int(step)
        if self.discriminator is not None:
            dstep = torch.tensor([self.opt_disc.step_count, self.global_step], device=self.flat_enc.device, dtype=torch.int64)
            self.sync.broadcast([self.flat_disc] + self.opt_disc.state() + [dstep], src)
            self.opt_disc.step_count, self.global_step = int(dstep[0]), int(dstep[1])
        for m in (self.encoder, self.decoder):
            m._weights_epoch = getattr(m, "_weights_epoch", 0) + 1

    def replica_checksum_spread(self) -> float:
        """max - min over ranks of the fp64 sum of all parameters: 0.0 when the replicas are in sync."""
        cs = (self.flat_enc.double().sum() + self.flat_dec.double().sum()).reshape(1)
        if not self.sync.enabled:
            return 0.0
        hi, lo = cs.clone(), cs.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.sync.group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.sync.group)
        return float(hi - lo)

    def _make_hook(self, name):
        def hook(runner, grads):
            self._grads[name] = grads
            self.sync.launch(grads)       # overlaps whatever backward work is still queued behind it
        return hook

    def forward_loss(self, reals: torch.Tensor, noise: Optional[torch.Tensor] = None):
        enc = self.encoder(reals)
        mean, scale = enc.chunk(2, dim=1)
        latents, kl = vae_sample_with_grad(mean, scale, noise)
        decoded = self.decoder(latents)
        nll = gaussian_nll(reals, decoded, self.log_sigma)
        loss = self.nll_weight * nll + self.kl_weight * kl
        info = {"nll": nll.detach(), "kl": kl.detach(), "latents": latents.detach(), "decoded": decoded.detach()}
        if self.spectral_loss is not None:
            st = self.spectral_loss(reals, decoded)
            loss = loss + self.spectral_weight * st
            info["mrstft"] = st.detach()
        if self._warmed_up():       # training/autoencoders.py:287-296, 144-145: the GAN terms of the generator loss
            for p in self.discriminator.parameters():
                p.requires_grad_(False)       # only the gradient w.r.t. the decoded signal is wanted here
            loss_dis, loss_adv, fm = self.discriminator.loss(reals, decoded)
            loss = loss + self.adversarial_weight * loss_adv + self.feature_matching_weight * fm
            info.update(loss_dis=loss_dis.detach(), loss_adv=loss_adv.detach(), feature_matching_distance=fm.detach())
        return loss, info

    def _warmed_up(self) -> bool:
        return self.discriminator is not None and self.global_step >= self.warmup_steps

    def discriminator_step(self, reals: torch.Tensor, noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """training/autoencoders.py:308-321: loss_dis -> backward -> discriminator optimizer.  The reference leaves
        ``decoded`` attached, so its backward also runs through the autoencoder and throws those gradients away at the
        next ``opt_gen.zero_grad()``; here the autoencoder runs without autograd."""
        disc = self.discriminator
        with torch.no_grad():
            mean, scale = self.encoder(reals).chunk(2, dim=1)
            latents, _ = vae_sample_with_grad(mean, scale, noise)
            decoded = self.decoder(latents)
        params = list(disc.parameters())
        for p in params:
            p.requires_grad_(True)
            p.grad = None
        with torch.enable_grad():
            loss_dis, loss_adv, fm = disc.loss(reals, decoded)
            loss_dis.backward()
        grads = torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1).float() for p in params])
        for p in params:
            p.grad = None
        self.sync.launch(grads)
        self.sync.wait()
        self.opt_disc.step([grads], grad_scale=self.sync.grad_scale)
        return {"loss": loss_dis.detach(), "loss_dis": loss_dis.detach(), "loss_adv": loss_adv.detach(),
                "feature_matching_distance": fm.detach(), "decoded": decoded}

    def training_step(self, reals: torch.Tensor, noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        step = self.global_step
        if self._warmed_up() and step % 2:      # training/autoencoders.py:309: odd steps train the discriminator
            info = self.discriminator_step(reals, noise)
            self.global_step = step + 1
            return info
        info = self.generator_step(reals, noise)
        self.global_step = step + 1
        return info

    def generator_step(self, reals: torch.Tensor, noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        for p in self.autoencoder.parameters():
            p.grad = None
        self._grads.clear()
        with torch.enable_grad():
            loss, info = self.forward_loss(reals, noise)
            loss.backward()
        self.sync.wait()
        self.opt.step([self._grads["enc"], self._grads["dec"]], grad_scale=self.sync.grad_scale)
        info["loss"] = loss.detach()
        return info
