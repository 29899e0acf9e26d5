"""Generates the golden fixtures in this directory by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports un-vendored third-party packages; they are stubbed exactly as SURVEY.md
section 8c describes: ``dac.nn.layers.WNConv1d/WNConvTranspose1d`` are the published two-liners
(old-style ``torch.nn.utils.weight_norm`` over ``nn.Conv1d`` / ``nn.ConvTranspose1d``), everything
else on the import path that the hot path never executes is a MagicMock.

Fixtures (all float32, little-endian .npz):
  tiny_ae.npz       self-contained: state_dicts + inputs + outputs of a 3-stage Oobleck at C=8
  mid_ae.npz        C=64 three-stage model (tensor-core path shapes); outputs + param checksums
  sao_full.npz      Stable-Audio-Open-shape model, [1,64,216] <-> [1,2,442368]; output slices
  o12_d512.npz      12.5 Hz shape (strides 2,4,4,5,8), latent 512, [1,512,16] -> [1,1,20480]
  chunked.npz       decode_audio / encode_audio with chunked=True on the tiny model
  sampling.npz      vae_sample (bottleneck.py:51) and sample() (model_sigmaVAE.py:187) outputs
  train_tiny.npz    one training step's loss and EVERY parameter gradient of the tiny model, from autograd through
                    the reference modules + the reference's vae_sample (generator branch of
                    training/autoencoders.py:221-352 with the Gaussian-NLL + KL objective of BASELINE config 5)
  train_mid.npz     same on the C=64 model: loss terms, per-parameter gradient norms and a strided sample
  glue.npz          the reference's own ``Llasa.infer`` loop (model_sigmaVAE.py:105-148) run with a small deterministic
                    stand-in for the Llama backbone: generated latents with and without the KL stop, the parameters of
                    audio_linear / distribution_linear, and the stand-in's matrix (so the test can re-create it)
  dataset.npz       twj_dataset.py:231-256 on three clips of different, non-multiple-of-the-ratio lengths: the
                    reference's pretransform.encode + vae_sample per clip (tiny model, self-contained)
  nearest.npz       decoders built with use_nearest_upsample=True (autoencoders.py:87-96): a tiny one (self-contained) and
                    a C=64 one (tensor-core path; checksums), outputs of the reference's OobleckDecoder
  bigvgan.npz       the reference's BigVGANFlowVAE (backup/flows.py, its 12.5 Hz VAE) at a small synthetic config, causal
                    and non-causal, AMPBlock1: state_dict, extract_latents and inference_from_latents outputs (with and
                    without sampling).  flows.py does ``from alias_free_torch import *`` (un-vendored): the published
                    algorithm restated in oracle/alias_free_restated.py is installed under that name
  o12_d256.npz      12.5 Hz shape, latent 256 ("dim512"), [1,256,16] <-> [1,1,20480], full outputs
  o12_d1024.npz     12.5 Hz shape, latent 1024 ("dim2048"), [1,1024,16] <-> [1,1,20480], full outputs
  o12_full.npz      BASELINE configs 3 and 4 at their own length: latent 512, [1,512,375] -> [1,1,480000] (config 3's
                    clip), and latent 1024, decode_audio(chunked=True, chunk_size=128, overlap=32) AND unchunked over
                    T=375 (config 4); sampled points (clip edges, every window seam, every 97th sample) + sums
  disc.npz          the reference's OobleckDiscriminator (models/discriminators.py:240-297; multi-scale Conv1d nets +
                    multi-period 15 x 15 Conv2d nets) at random init, stereo and mono (a length that is no multiple of
                    any period): loss(reals, fakes) = hinge discriminator / generator losses and the feature-matching
                    distance, the summed scores, a checksum of each of the 40 feature tensors, autograd gradients of
                    each loss w.r.t. the fakes and of the discriminator loss w.r.t. every parameter (norm + a strided
                    sample); weights re-created from the seed, per-parameter checksums recorded
Weights of the larger models are NOT stored: they are re-created from the recorded seed by the
same construction order (nn.Conv1d / nn.ConvTranspose1d default init), and the fixture carries a
float64 checksum of every parameter so a mismatch is detected rather than silently compared.
"""
import io
import os
import sys
import types
import warnings
from contextlib import redirect_stdout
from unittest.mock import MagicMock

import numpy as np
import torch
from torch import nn
from torch.nn.utils import weight_norm

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF)
    dac = types.ModuleType("dac")
    dac_nn = types.ModuleType("dac.nn")
    layers = types.ModuleType("dac.nn.layers")
    layers.WNConv1d = lambda *a, **k: weight_norm(nn.Conv1d(*a, **k))
    layers.WNConvTranspose1d = lambda *a, **k: weight_norm(nn.ConvTranspose1d(*a, **k))
    layers.Snake1d = MagicMock()
    sys.modules.update({"dac": dac, "dac.nn": dac_nn, "dac.nn.layers": layers})
    for m in ["dac.nn.quantize", "dac.model", "dac.model.dac", "dac.model.discriminator", "alias_free_torch",
              "vector_quantize_pytorch", "k_diffusion", "x_transformers", "einops_exts", "audiotools",
              "encodec", "pywt"]:
        sys.modules[m] = MagicMock()
    from stable_audio_tools.models import autoencoders, bottleneck
    return autoencoders, bottleneck


def ae_config(channels, c_mults, strides, enc_latent, dec_latent, io_channels, sample_rate):
    ratio = int(np.prod(strides))
    return {
        "model_type": "autoencoder",
        "sample_rate": sample_rate,
        "model": {
            "encoder": {"type": "oobleck", "config": {"in_channels": io_channels, "channels": channels,
                                                      "c_mults": list(c_mults), "strides": list(strides),
                                                      "latent_dim": enc_latent, "use_snake": True}},
            "decoder": {"type": "oobleck", "config": {"out_channels": io_channels, "channels": channels,
                                                      "c_mults": list(c_mults), "strides": list(strides),
                                                      "latent_dim": dec_latent, "use_snake": True,
                                                      "final_tanh": False}},
            "bottleneck": {"type": "vae"},
            "latent_dim": dec_latent,
            "downsampling_ratio": ratio,
            "io_channels": io_channels,
        },
    }


CONFIGS = {
    "tiny": ae_config(8, [1, 2, 4], [2, 4, 5], 8, 4, 2, 16000),
    # encode_audio(chunked=True) pastes encoder output into a [B, latent_dim, T] buffer, so it only
    # runs when the encoder emits latent_dim channels (autoencoders.py:471,496)
    "tiny_sym": ae_config(8, [1, 2, 4], [2, 4, 5], 4, 4, 2, 16000),
    "mid": ae_config(64, [1, 2, 4], [2, 4, 5], 128, 64, 2, 16000),
    "sao": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 8, 8], 128, 64, 2, 44100),
    "o12_d512": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 1024, 512, 1, 16000),
    "o12_d256": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 512, 256, 1, 16000),
    "o12_d1024": ae_config(128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 2048, 1024, 1, 16000),
}


def checksums(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items()}


def randomize_snake(model, seed):
    """alpha/beta are zero at init (exp -> 1); perturb them so the fixture exercises the per-channel
    parameters.  Deterministic given the seed; the product tests apply the same perturbation."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".alpha") or name.endswith(".beta"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))


def build(ae_mod, name, seed=0, snake_seed=None):
    torch.manual_seed(seed)
    m = ae_mod.create_autoencoder_from_config(CONFIGS[name]).eval()
    if snake_seed is not None:
        randomize_snake(m, snake_seed)
    return m


KL_WEIGHT, LOG_SIGMA = 1e-2, -1.0   # large enough that both loss terms shape the gradients


def ref_training_grads(ae_mod, bn_mod, name, x, snake_seed):
    """Generator branch of the reference's training_step on reference modules, autograd gradients."""
    import math
    m = build(ae_mod, name, 0, snake_seed=snake_seed).train()
    with torch.enable_grad():
        enc = m.encoder(x)
        mean, scale = enc.chunk(2, dim=1)
        torch.manual_seed(3)
        noise = torch.randn_like(mean)
        torch.manual_seed(3)
        with redirect_stdout(io.StringIO()):
            lat, kl = bn_mod.vae_sample(mean, scale)       # draws the same noise
        dec = m.decoder(lat)
        nll = (0.5 * ((x - dec) / math.exp(LOG_SIGMA)) ** 2 + LOG_SIGMA + 0.5 * math.log(2 * math.pi)).flatten(1).sum(1).mean()
        loss = nll + KL_WEIGHT * kl
        loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    return m, noise, loss.detach(), nll.detach(), kl.detach(), dec.detach(), grads


def train_fixtures(ae_mod, bn_mod):
    x = 0.1 * torch.randn(2, 2, 40 * 9, generator=torch.Generator().manual_seed(2))
    m, noise, loss, nll, kl, dec, grads = ref_training_grads(ae_mod, bn_mod, "tiny", x, 7)
    out = {"x": x, "noise": noise, "loss": loss, "nll": nll, "kl": kl, "decoded": dec,
           "kl_weight": torch.tensor(KL_WEIGHT), "log_sigma": torch.tensor(LOG_SIGMA)}
    for k, v in m.state_dict().items():
        out["sd." + k] = v
    for k, v in grads.items():
        out["g." + k] = v
    np.savez_compressed(os.path.join(HERE, "train_tiny.npz"), **{k: v.numpy() for k, v in out.items()})

    x = 0.1 * torch.randn(2, 2, 40 * 24, generator=torch.Generator().manual_seed(2))
    m, noise, loss, nll, kl, dec, grads = ref_training_grads(ae_mod, bn_mod, "mid", x, 7)
    cs = checksums(m.state_dict())
    out = {"x": x.numpy(), "noise": noise.numpy(), "loss": float(loss), "nll": float(nll), "kl": float(kl),
           "kl_weight": KL_WEIGHT, "log_sigma": LOG_SIGMA,
           "cs_keys": np.array(list(cs.keys())), "cs_vals": np.array(list(cs.values()), dtype=np.float64),
           "g_keys": np.array(list(grads.keys())),
           "g_norm": np.array([float(g.double().norm()) for g in grads.values()], dtype=np.float64),
           "g_sum": np.array([float(g.double().sum()) for g in grads.values()], dtype=np.float64)}
    for i, g in enumerate(grads.values()):
        out[f"g_sample{i}"] = g.reshape(-1)[::max(1, g.numel() // 256)][:256].numpy()
    np.savez_compressed(os.path.join(HERE, "train_mid.npz"), **out)


def o12_fixtures(ae_mod):
    """BASELINE configs 3 / 4 architectures (12.5 Hz, strides 2,4,4,5,8) at latent 256 / 1024 (short clip, full
    outputs) and at the configs' own length T = 375 (sampled points)."""
    torch.set_grad_enabled(False)
    for name, D in (("o12_d256", 256), ("o12_d1024", 1024)):
        m = build(ae_mod, name, 0)
        z = torch.randn(1, D, 16, generator=torch.Generator().manual_seed(1))
        x = 0.1 * torch.randn(1, 1, 1280 * 16, generator=torch.Generator().manual_seed(2))
        cs = checksums(m.state_dict())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), dec_out=m.decode(z).numpy(), enc_out=m.encode(x).numpy(),
                            cs_keys=np.array(list(cs.keys())), cs_vals=np.array(list(cs.values()), dtype=np.float64))
    L = 480000
    seams = [s * 1280 for s in (16, 96, 112, 192, 208, 247, 263, 288, 359)]    # window starts / paste boundaries of 128/32 over 375
    idx = torch.cat([torch.arange(0, 4096), torch.arange(L - 4096, L), torch.arange(0, L, 97)] +
                    [torch.arange(s - 640, s + 640) for s in seams]).unique()
    out = {"idx": idx.numpy()}
    m = build(ae_mod, "o12_d512", 0)
    z = torch.randn(1, 512, 375, generator=torch.Generator().manual_seed(1))
    y = m.decode(z)
    out.update(d512_at_idx=y[:, :, idx].numpy(), d512_abs_max=float(y.abs().max()), d512_sum=float(y.double().sum()),
               d512_sq_sum=float((y.double() ** 2).sum()))
    m = build(ae_mod, "o12_d1024", 0)
    z = torch.randn(1, 1024, 375, generator=torch.Generator().manual_seed(1))
    y = m.decode_audio(z, chunked=False)
    yc = m.decode_audio(z, chunked=True, overlap=32, chunk_size=128)
    out.update(d1024_at_idx=y[:, :, idx].numpy(), d1024_chunked_at_idx=yc[:, :, idx].numpy(),
               d1024_abs_max=float(y.abs().max()), d1024_sq_sum=float((y.double() ** 2).sum()),
               d1024_chunked_sq_sum=float((yc.double() ** 2).sum()),
               d1024_chunked_vs_full_max=float((y - yc).abs().max()))
    np.savez_compressed(os.path.join(HERE, "o12_full.npz"), **out)
    torch.set_grad_enabled(True)


class _FakeBackbone(nn.Module):
    """Deterministic stand-in for ``AutoModelForCausalLM``: hidden_t = tanh(mean_{s<=t}(embed_s) @ M).  Only what
    ``Llasa.__init__`` / ``Llasa.infer`` touch."""

    class _Inner(nn.Module):
        def __init__(self, hidden, vocab, M):
            super().__init__()
            self.embed_tokens = nn.Embedding(vocab, hidden)
            self.register_buffer("M", M)

        def forward(self, inputs_embeds=None, attention_mask=None):
            c = inputs_embeds.cumsum(dim=1) / torch.arange(1, inputs_embeds.shape[1] + 1).view(1, -1, 1)
            return (torch.tanh(c @ self.M),)

    def __init__(self, hidden, vocab, M):
        super().__init__()
        self.model = self._Inner(hidden, vocab, M)
        self.config = types.SimpleNamespace(vocab_size=vocab, hidden_size=hidden)

    def resize_token_embeddings(self, n):
        return None


def glue_fixtures():
    """Runs the reference's Llasa.infer (the per-frame distribution_linear -> sample -> KL stop -> audio_linear loop)
    unmodified; only AutoModelForCausalLM.from_pretrained is replaced by the stand-in backbone above."""
    import transformers
    sys.path.insert(0, REF)
    H, D, V = 256, 64, 50
    torch.manual_seed(41)
    M = torch.randn(H, H) / H ** 0.5
    fake = _FakeBackbone(H, V, M)
    orig = transformers.AutoModelForCausalLM.from_pretrained
    transformers.AutoModelForCausalLM.from_pretrained = staticmethod(lambda *a, **k: fake)
    try:
        import importlib
        msv = importlib.import_module("model_sigmaVAE")
        torch.manual_seed(42)
        llasa = msv.Llasa({"llm_model_name_or_path": "stand-in", "latent_dim": D, "audio_proj_dim": H},
                          tokenizer=list(range(V)), use_flash_attention=False).eval()
    finally:
        transformers.AutoModelForCausalLM.from_pretrained = orig
    ids = torch.arange(7) % V
    prompt = torch.randn(1, 3, D, generator=torch.Generator().manual_seed(43))
    out = {"M": M, "ids": ids, "prompt": prompt, "embed_tokens": fake.model.embed_tokens.weight.detach()}
    for k, v in llasa.state_dict().items():
        if k.startswith("audio_linear") or k.startswith("distribution_linear"):
            out["sd." + k] = v
    with torch.no_grad():
        torch.manual_seed(44)
        out["latents_no_stop"] = llasa.infer(ids, prompt, end_disp_kl_thres=0.0, max_length=7)      # [1, D, 6]
        torch.manual_seed(44)
        out["latents_kl_stop"] = llasa.infer(ids, prompt, end_disp_kl_thres=1e9, max_length=20)     # stops at i = 4
        torch.manual_seed(44)
        out["noise"] = torch.stack([torch.randn(1, 1, D) for _ in range(7)])
    np.savez_compressed(os.path.join(HERE, "glue.npz"), **{k: v.numpy() for k, v in out.items()})


def dataset_fixtures(ae_mod, bn_mod):
    """twj_dataset.py:231-256 per clip on the reference's modules.  librosa is not installed here; its
    ``util.normalize`` (norm=inf, the default) is the published ``x / max|x|`` (left alone below float tiny)."""
    torch.set_grad_enabled(False)
    from stable_audio_tools.models.factory import create_pretransform_from_config
    torch.manual_seed(0)
    pt = create_pretransform_from_config({"type": "autoencoder", "config": CONFIGS["tiny"]["model"], "scale": 1.0,
                                          "iterate_batch": True}, 16000)
    randomize_snake(pt.model, 7)
    out = {}
    for k, v in pt.model.state_dict().items():
        out["sd." + k] = v
    for i, L in enumerate((40 * 23 + 17, 40 * 9, 40 * 31 + 39)):
        wav = (0.3 * torch.randn(L, generator=torch.Generator().manual_seed(60 + i))).numpy()
        peak = np.abs(wav).max()
        norm_wav = (wav / peak if peak > np.finfo(np.float32).tiny else wav) * 0.95
        norm_wav = torch.from_numpy(norm_wav.astype(np.float32)).reshape(1, -1)
        dual_norm_wav = norm_wav.repeat(2, 1).unsqueeze(0)
        mean_scale_latent = pt.encode(dual_norm_wav).detach()
        mean, scale = mean_scale_latent.chunk(2, dim=1)
        torch.manual_seed(70 + i)
        noise = torch.randn_like(mean)
        torch.manual_seed(70 + i)
        with redirect_stdout(io.StringIO()):
            latents, kl = bn_mod.vae_sample(mean, scale)
        latents = latents.squeeze(0).transpose(0, 1)
        out.update({f"wav{i}": torch.from_numpy(wav), f"dual{i}": dual_norm_wav, f"mean_scale{i}": mean_scale_latent,
                    f"noise{i}": noise, f"latents{i}": latents})
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **{k: v.numpy() for k, v in out.items()})
    torch.set_grad_enabled(True)


def nearest_fixtures(ae_mod):
    torch.set_grad_enabled(False)
    out = {}
    torch.manual_seed(0)
    d = ae_mod.OobleckDecoder(out_channels=2, channels=8, latent_dim=4, c_mults=[1, 2, 4], strides=[2, 4, 5], use_snake=True,
                              use_nearest_upsample=True, final_tanh=False).eval()
    randomize_snake(d, 7)
    z = torch.randn(2, 4, 13, generator=torch.Generator().manual_seed(1))
    out.update(tiny_z=z, tiny_out=d(z))
    for k, v in d.state_dict().items():
        out["tiny_sd." + k] = v
    torch.manual_seed(0)
    d = ae_mod.OobleckDecoder(out_channels=2, channels=64, latent_dim=64, c_mults=[1, 2, 4], strides=[2, 4, 5], use_snake=True,
                              use_nearest_upsample=True, final_tanh=True).eval()
    randomize_snake(d, 7)
    z = torch.randn(2, 64, 24, generator=torch.Generator().manual_seed(1))
    cs = checksums(d.state_dict())
    out.update(mid_z=z, mid_out=d(z))
    out = {k: v.numpy() for k, v in out.items()}
    out.update(cs_keys=np.array(list(cs.keys())), cs_vals=np.array(list(cs.values()), dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, "nearest.npz"), **out)
    torch.set_grad_enabled(True)


BIGVGAN_H = dict(latent_dim=16, use_vae=True, downsample_channels=[12, 24, 48], downsample_rates=[2, 4],
                 flow_hidden_channels=8, resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]],
                 upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=32, resblock="1",
                 activation="snakebeta", snake_logscale=True)


def bigvgan_fixtures():
    import importlib
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import alias_free_restated
    sys.modules["alias_free_torch"] = alias_free_restated
    sys.path.insert(0, os.path.join(REF, "backup"))
    flows = importlib.import_module("flows")

    class AttrDict(dict):
        __getattr__ = dict.__getitem__

    out = {}
    torch.set_grad_enabled(False)
    for tag, causal in (("causal", True), ("noncausal", False)):
        h = AttrDict(BIGVGAN_H, causal=causal)
        torch.manual_seed(0)
        m = flows.BigVGANFlowVAE(h).eval()
        g = torch.Generator().manual_seed(7)
        for name, p in m.named_parameters():                 # alpha / beta start at zero: perturb them
            if name.endswith(".alpha") or name.endswith(".beta"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
        x = 0.3 * torch.randn(2, 1, 8 * 37, generator=torch.Generator().manual_seed(1))
        lat = m.extract_latents(x)
        torch.manual_seed(5)
        noise = torch.randn(2, 16, 37)
        torch.manual_seed(5)
        y_s = m.inference_from_latents(lat)                  # do_sample: draws randn_like(m_q) -> the same noise
        z = torch.randn(2, 16, 21, generator=torch.Generator().manual_seed(2))
        y_z = m.inference_from_latents(z, do_sample=False)
        out.update({f"{tag}.x": x, f"{tag}.latents": lat, f"{tag}.noise": noise, f"{tag}.wav_sampled": y_s, f"{tag}.z": z,
                    f"{tag}.wav": y_z})
        if causal:                                           # same seed, same shapes: both variants share the parameters
            for k, v in m.state_dict().items():
                out[f"sd.{k}"] = v
        else:
            assert all(torch.equal(v, out[f"sd.{k}"]) for k, v in m.state_dict().items())
    np.savez_compressed(os.path.join(HERE, "bigvgan.npz"), **{k: v.numpy() for k, v in out.items()})
    torch.set_grad_enabled(True)



MRSTFT_ARGS = dict(fft_sizes=[2048, 1024, 512, 256, 128, 64, 32], hop_sizes=[512, 256, 128, 64, 32, 16, 8],
                   win_lengths=[2048, 1024, 512, 256, 128, 64, 32], perceptual_weighting=True)


def mrstft_fixtures():
    """SURVEY section 8(f) item 4 (loss half): the reference's own auraloss copy (training/losses/auraloss.py:443-606)
    with the stft_loss_args of the Stable Audio autoencoder configs, called the way the training wrapper calls it
    (training/autoencoders.py:163: AuralossLoss(self.sdstft, 'reals', 'decoded') -> module(input = reals, target =
    decoded), so the gradient that matters is the one w.r.t. the TARGET argument).  Values + autograd gradients w.r.t.
    both arguments, for the stereo sum-and-difference form, the plain multi-resolution form on stereo and on mono
    input, and a short-window variant (win_length < fft_size, no pre-filter)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_auraloss", os.path.join(REF, "stable_audio_tools", "training", "losses",
                                                                               "auraloss.py"))
    al = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(al)
    torch.manual_seed(11)
    T = 4500
    out = {}

    def signals(B, C):
        t = torch.arange(T, dtype=torch.float32) / 44100.0
        base = 0.3 * torch.sin(2 * np.pi * 440.0 * t) + 0.1 * torch.sin(2 * np.pi * 3100.0 * t + 0.5)
        x = (base + 0.05 * torch.randn(B, C, T)).contiguous()
        y = (0.9 * base + 0.07 * torch.randn(B, C, T)).contiguous()
        return x, y

    def run(tag, mod, x, y):
        x = x.clone().requires_grad_(True)
        y = y.clone().requires_grad_(True)
        loss = mod(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y))
        out[f"{tag}.loss"] = np.float64(loss.item())
        out[f"{tag}.gx"] = gx.numpy()
        out[f"{tag}.gy"] = gy.numpy()
        print(tag, "loss", loss.item(), "|gx|", float(gx.abs().max()), "|gy|", float(gy.abs().max()))

    x2, y2 = signals(2, 2)
    out["x2"], out["y2"] = x2.numpy(), y2.numpy()
    run("sd", al.SumAndDifferenceSTFTLoss(sample_rate=44100, **MRSTFT_ARGS), x2, y2)
    run("mr_stereo", al.MultiResolutionSTFTLoss(sample_rate=44100, **MRSTFT_ARGS), x2, y2)
    x1, y1 = signals(2, 1)
    out["x1"], out["y1"] = x1.numpy(), y1.numpy()
    run("mr_mono", al.MultiResolutionSTFTLoss(sample_rate=44100, **MRSTFT_ARGS), x1, y1)
    run("short_win", al.MultiResolutionSTFTLoss(fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50],
                                                win_lengths=[600, 1200, 240]), x1, y1)
    # the A-weighting FIR taps the reference designs with scipy (101 taps at 44.1 kHz)
    out["aw_taps_44100"] = al.FIRFilter(filter_type="aw", fs=44100).fir.weight.data.view(-1).numpy()
    np.savez_compressed(os.path.join(HERE, "mrstft.npz"), **out)
    print("wrote mrstft.npz", os.path.getsize(os.path.join(HERE, "mrstft.npz")) // 1024, "KiB")


def disc_fixtures():
    """SURVEY section 8(f) item 4 (discriminator half): the reference's OobleckDiscriminator, imported unmodified
    (audiotools / dac.model.discriminator / encodec, which this class never touches, are stubbed)."""
    import importlib.util
    for m in ["audiotools", "dac", "dac.model", "dac.model.discriminator", "encodec", "encodec.msstftd"]:
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location("ref_discriminators",
                                                  os.path.join(REF, "stable_audio_tools", "models", "discriminators.py"))
    dm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dm)
    out = {}
    for tag, C, B, T, seed in (("stereo", 2, 2, 3000, 21), ("mono", 1, 3, 2999, 22)):
        torch.manual_seed(seed)
        m = dm.OobleckDiscriminator(in_channels=C)
        g = torch.Generator().manual_seed(seed + 100)
        t = torch.arange(T, dtype=torch.float32) / 44100.0
        base = 0.3 * torch.sin(2 * np.pi * 440.0 * t) + 0.1 * torch.sin(2 * np.pi * 3100.0 * t + 0.5)
        reals = (base + 0.05 * torch.randn(B, C, T, generator=g)).contiguous()
        fakes = (0.9 * base + 0.07 * torch.randn(B, C, T, generator=g)).contiguous().requires_grad_(True)
        out[f"{tag}.seed"] = np.int64(seed)
        out[f"{tag}.reals"], out[f"{tag}.fakes"] = reals.numpy(), fakes.detach().numpy()
        for k, v in checksums(m.state_dict()).items():
            out[f"{tag}.cs.{k}"] = v
        dis, gen, fm = m.loss(reals, fakes)
        out[f"{tag}.dis"], out[f"{tag}.gen"], out[f"{tag}.fm"] = (np.float64(dis.item()), np.float64(gen.item()),
                                                                  np.float64(fm.item()))
        inputs = m.multi_discriminator({"reals": reals, "fakes": fakes})
        out[f"{tag}.score_reals"] = inputs["score_reals"].detach().numpy()
        out[f"{tag}.score_fakes"] = inputs["score_fakes"].detach().numpy()
        fr, ff = inputs["features_reals"], inputs["features_fakes"]
        out[f"{tag}.n_features"] = np.int64(len(fr))
        for i, (a, b) in enumerate(zip(fr, ff)):
            out[f"{tag}.feat{i}.shape"] = np.array(a.shape, dtype=np.int64)
            out[f"{tag}.feat{i}.sums"] = np.array([a.double().sum().item(), a.double().abs().sum().item(),
                                                   b.double().sum().item(), b.double().abs().sum().item()])
        params = [p for p in m.parameters()]
        names = [k for k, _ in m.named_parameters()]
        for lname, loss in (("dis", dis), ("gen", gen), ("fm", fm)):
            gr = torch.autograd.grad(loss, [fakes] + params, retain_graph=True, allow_unused=True)
            out[f"{tag}.g_fakes.{lname}"] = gr[0].numpy()
            if lname == "gen":
                continue
            for k, gp in zip(names, gr[1:]):
                gp = torch.zeros(1) if gp is None else gp
                out[f"{tag}.gp.{lname}.{k}.norm"] = np.float64(gp.double().norm().item())
                flat = gp.reshape(-1)
                out[f"{tag}.gp.{lname}.{k}.sample"] = flat[:: max(1, flat.numel() // 61)][:64].numpy()
        print(tag, "dis", dis.item(), "gen", gen.item(), "fm", fm.item(), "features", len(fr))
    np.savez_compressed(os.path.join(HERE, "disc.npz"), **out)
    print("wrote disc.npz", os.path.getsize(os.path.join(HERE, "disc.npz")) // 1024, "KiB")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "disc":
        disc_fixtures()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "mrstft":
        mrstft_fixtures()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "bigvgan":
        bigvgan_fixtures()
        return
    ae_mod, bn_mod = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "nearest":
        nearest_fixtures(ae_mod)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "glue":       # only the LM-glue and dataset-side fixtures
        glue_fixtures()
        dataset_fixtures(ae_mod, bn_mod)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "train":      # only the training fixtures
        train_fixtures(ae_mod, bn_mod)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "o12":        # only the 12.5 Hz fixtures at configs 3 / 4 sizes
        o12_fixtures(ae_mod)
        return
    bigvgan_fixtures()
    disc_fixtures()
    mrstft_fixtures()
    o12_fixtures(ae_mod)
    nearest_fixtures(ae_mod)
    glue_fixtures()
    dataset_fixtures(ae_mod, bn_mod)
    train_fixtures(ae_mod, bn_mod)
    torch.set_grad_enabled(False)

    # ---- tiny: self-contained
    m = build(ae_mod, "tiny", 0, snake_seed=7)
    z = torch.randn(2, 4, 13, generator=torch.Generator().manual_seed(1))
    x = 0.1 * torch.randn(2, 2, 40 * 9, generator=torch.Generator().manual_seed(2))
    out = {"z": z, "x": x, "dec_out": m.decode(z), "enc_out": m.encode(x)}
    # per-layer activations of the decoder (layer-level parity)
    h = z
    for i, layer in enumerate(m.decoder.layers):
        h = layer(h)
        out[f"dec_layer{i}"] = h
    for k, v in m.state_dict().items():
        out["sd." + k] = v
    np.savez_compressed(os.path.join(HERE, "tiny_ae.npz"), **{k: v.numpy() for k, v in out.items()})

    # ---- chunked on the tiny model
    zc = torch.randn(1, 4, 300, generator=torch.Generator().manual_seed(11))
    xc = 0.1 * torch.randn(1, 2, 40 * 300, generator=torch.Generator().manual_seed(12))
    ms = build(ae_mod, "tiny_sym", 5, snake_seed=9)
    ch = {"z": zc, "x": xc,
          "dec_chunked": m.decode_audio(zc, chunked=True, overlap=32, chunk_size=128),
          "dec_full": m.decode_audio(zc, chunked=False),
          "dec_chunked_64_16": m.decode_audio(zc, chunked=True, overlap=16, chunk_size=64),
          "enc_chunked": ms.encode_audio(xc, chunked=True, overlap=32, chunk_size=128),
          "enc_full": ms.encode_audio(xc, chunked=False)}
    for k, v in ms.state_dict().items():
        ch["sym_sd." + k] = v
    np.savez_compressed(os.path.join(HERE, "chunked.npz"), **{k: v.numpy() for k, v in ch.items()})

    # ---- mid
    m = build(ae_mod, "mid", 0, snake_seed=7)
    z = torch.randn(2, 64, 24, generator=torch.Generator().manual_seed(1))
    x = 0.1 * torch.randn(2, 2, 40 * 24, generator=torch.Generator().manual_seed(2))
    out = {"z": z, "x": x, "dec_out": m.decode(z), "enc_out": m.encode(x)}
    cs = checksums(m.state_dict())
    out["cs_keys"] = np.array(list(cs.keys()))
    out["cs_vals"] = np.array(list(cs.values()), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "mid_ae.npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})

    # ---- SAO full size (BASELINE config 1 input shape)
    m = build(ae_mod, "sao", 0)
    z = torch.randn(1, 64, 216, generator=torch.Generator().manual_seed(1))
    y = m.decode(z)
    x = 0.1 * torch.randn(1, 2, 442368, generator=torch.Generator().manual_seed(2))
    e = m.encode(x)
    idx = torch.cat([torch.arange(0, 4096), torch.arange(221184 - 2048, 221184 + 2048),
                     torch.arange(442368 - 4096, 442368), torch.arange(0, 442368, 97)]).unique()
    cs = checksums(m.state_dict())
    np.savez_compressed(os.path.join(HERE, "sao_full.npz"), dec_idx=idx.numpy(), dec_out_at_idx=y[:, :, idx].numpy(),
                        dec_abs_max=float(y.abs().max()), dec_sum=float(y.double().sum()),
                        dec_sq_sum=float((y.double() ** 2).sum()), enc_out=e.numpy(),
                        cs_keys=np.array(list(cs.keys())), cs_vals=np.array(list(cs.values()), dtype=np.float64))

    # ---- O12 latent 512
    m = build(ae_mod, "o12_d512", 0)
    z = torch.randn(1, 512, 16, generator=torch.Generator().manual_seed(1))
    y = m.decode(z)
    x = 0.1 * torch.randn(1, 1, 1280 * 16, generator=torch.Generator().manual_seed(2))
    e = m.encode(x)
    cs = checksums(m.state_dict())
    np.savez_compressed(os.path.join(HERE, "o12_d512.npz"), dec_out=y.numpy(), enc_out=e.numpy(),
                        cs_keys=np.array(list(cs.keys())), cs_vals=np.array(list(cs.values()), dtype=np.float64))

    # ---- sampling
    sys.path.insert(0, REF)
    mean = torch.randn(3, 64, 50, generator=torch.Generator().manual_seed(21))
    scale = torch.randn(3, 64, 50, generator=torch.Generator().manual_seed(22))
    torch.manual_seed(3)
    noise = torch.randn_like(mean)
    torch.manual_seed(3)
    with redirect_stdout(io.StringIO()):
        lat, kl = bn_mod.vae_sample(mean, scale)
    # model_sigmaVAE.py imports transformers' Llama at module import; the free function sample() at
    # :187-213 is self-contained, so execute just that function's source.
    src = open(os.path.join(REF, "model_sigmaVAE.py")).read()
    start = src.index("\ndef sample(mean, dist_type='fix'):")
    ns = {"torch": torch}
    exec(src[start:], ns)
    sample = ns["sample"]
    torch.manual_seed(3)
    fix = sample(mean, "fix")
    torch.manual_seed(3)
    std_noise = torch.randn(3)
    noise_g = torch.randn_like(mean)
    torch.manual_seed(3)
    gau = sample(mean, "gaussian")
    mean_bf = mean.bfloat16()
    torch.manual_seed(3)
    noise_bf = torch.randn_like(mean_bf)
    torch.manual_seed(3)
    fix_bf = sample(mean_bf, "fix")
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), mean=mean.numpy(), scale=scale.numpy(),
                        noise=noise.numpy(), vae_latents=lat.numpy(), vae_kl=float(kl), fix=fix.numpy(),
                        std_noise=std_noise.numpy(), noise_g=noise_g.numpy(), gaussian=gau.numpy(),
                        noise_bf=noise_bf.float().numpy(), fix_bf=fix_bf.float().numpy(),
                        none=sample(mean, "other").numpy())
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
