"""Pins the CPU oracle (oracle/oobleck_oracle.py) against outputs of the reference's own modules
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import oobleck_oracle as O

torch.set_grad_enabled(False)


def _sd_from_golden(g, prefix="sd."):
    return {k[len(prefix):]: H.t(g[k]) for k in g.files if k.startswith(prefix)}


def test_tiny_decoder_encoder_match_reference():
    g = H.golden("tiny_ae")
    sd = _sd_from_golden(g)
    st = H.strides_of("tiny")
    y = O.oobleck_decoder(H.split_sd(sd, "decoder."), H.t(g["z"]), st)
    e = O.oobleck_encoder(H.split_sd(sd, "encoder."), H.t(g["x"]), st)
    assert y.shape == g["dec_out"].shape and e.shape == g["enc_out"].shape
    assert np.abs(y.numpy() - g["dec_out"]).max() <= 2e-6
    assert np.abs(e.numpy() - g["enc_out"]).max() <= 2e-6


def test_tiny_decoder_layerwise():
    g = H.golden("tiny_ae")
    sd = H.split_sd(_sd_from_golden(g), "decoder.")
    st = H.strides_of("tiny")
    n = len(st)
    h = O._wn_conv1d(sd, "layers.0", H.t(g["z"]), padding=3)
    assert np.abs(h.numpy() - g["dec_layer0"]).max() <= 1e-6
    for i in range(n):
        h = O.decoder_block(sd, f"layers.{1 + i}", h, st[n - 1 - i])
        ref = g[f"dec_layer{1 + i}"]
        assert np.abs(h.numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    h = O._snake(sd, f"layers.{1 + n}", h)
    assert np.abs(h.numpy() - g[f"dec_layer{1 + n}"]).max() <= 2e-6


def test_weight_norm_fold_matches_torch():
    v = torch.randn(6, 5, 7)
    g = torch.rand(6, 1, 1) + 0.5
    assert torch.allclose(O.weight_norm_fold(v, g), torch._weight_norm(v, g, 0), atol=1e-7)
    vt = torch.randn(5, 6, 4)       # transposed conv: norm over dim 0 = in-channels
    gt = torch.rand(5, 1, 1) + 0.5
    assert torch.allclose(O.weight_norm_fold(vt, gt), torch._weight_norm(vt, gt, 0), atol=1e-7)


def test_chunked_paths_match_reference():
    g = H.golden("chunked")
    tiny = H.golden("tiny_ae")
    sd = H.split_sd(_sd_from_golden(tiny), "decoder.")
    st = H.strides_of("tiny")
    dec = lambda z: O.oobleck_decoder(sd, z, st)
    y = O.decode_audio_chunked(dec, H.t(g["z"]), 40, 2, overlap=32, chunk_size=128)
    assert np.abs(y.numpy() - g["dec_chunked"]).max() <= 2e-6
    y2 = O.decode_audio_chunked(dec, H.t(g["z"]), 40, 2, overlap=16, chunk_size=64)
    assert np.abs(y2.numpy() - g["dec_chunked_64_16"]).max() <= 2e-6
    # receptive field +-10 frames < overlap/2 = 16: chunked == unchunked to fp32 noise
    assert np.abs(g["dec_chunked"] - g["dec_full"]).max() <= 1e-5
    sym = H.split_sd({k[len("sym_sd."):]: H.t(g[k]) for k in g.files if k.startswith("sym_sd.")}, "encoder.")
    enc = lambda x: O.oobleck_encoder(sym, x, st)
    e = O.encode_audio_chunked(enc, H.t(g["x"]), 40, 4, overlap=32, chunk_size=128)
    assert np.abs(e.numpy() - g["enc_chunked"]).max() <= 2e-6
    with pytest.raises(UnboundLocalError):
        O.decode_audio_chunked(dec, H.t(g["z"])[:, :, :100], 40, 2)


def test_sampling_bit_exact():
    g = H.golden("sampling")
    mean, scale, noise = H.t(g["mean"]), H.t(g["scale"]), H.t(g["noise"])
    lat, kl = O.vae_sample(mean, scale, noise)
    assert torch.equal(lat, H.t(g["vae_latents"]))
    assert abs(float(kl) - float(g["vae_kl"])) <= 1e-4 * abs(float(g["vae_kl"]))
    assert torch.equal(O.sigma_sample(mean, noise, "fix"), H.t(g["fix"]))
    assert torch.equal(O.sigma_sample(mean, H.t(g["noise_g"]), "gaussian", H.t(g["std_noise"])), H.t(g["gaussian"]))
    assert torch.equal(O.sigma_sample(mean, noise, "other"), H.t(g["none"]))
    fb = O.sigma_sample(mean.bfloat16(), H.t(g["noise_bf"]).bfloat16(), "fix")
    assert fb.dtype == torch.bfloat16 and torch.equal(fb.float(), H.t(g["fix_bf"]))


def test_mid_model_same_seed_same_weights_and_outputs():
    g = H.golden("mid_ae")
    m = H.build("mid", 0, snake_seed=7)
    sd = m.state_dict()
    H.check_checksums(sd, g)
    st = H.strides_of("mid")
    y = O.oobleck_decoder(H.split_sd(sd, "decoder."), H.t(g["z"]), st)
    e = O.oobleck_encoder(H.split_sd(sd, "encoder."), H.t(g["x"]), st)
    assert np.abs(y.numpy() - g["dec_out"]).max() <= 5e-6
    assert np.abs(e.numpy() - g["enc_out"]).max() <= 5e-6


def test_o12_latent512_matches_reference():
    g = H.golden("o12_d512")
    m = H.build("o12_d512", 0)
    sd = m.state_dict()
    H.check_checksums(sd, g)
    st = H.strides_of("o12_d512")
    z = torch.randn(1, 512, 16, generator=torch.Generator().manual_seed(1))
    y = O.oobleck_decoder(H.split_sd(sd, "decoder."), z, st)
    assert y.shape == (1, 1, 1280 * 16)
    assert np.abs(y.numpy() - g["dec_out"]).max() <= 5e-6


@pytest.mark.parametrize("name,D", [("o12_d256", 256), ("o12_d1024", 1024)])
def test_o12_latent256_and_1024_match_reference(name, D):
    """The other two 12.5 Hz widths the reference ships ("dim512" / "dim2048"): decoder and encoder, full outputs."""
    g = H.golden(name)
    m = H.build(name, 0)
    sd = m.state_dict()
    H.check_checksums(sd, g)
    st = H.strides_of(name)
    z = torch.randn(1, D, 16, generator=torch.Generator().manual_seed(1))
    x = 0.1 * torch.randn(1, 1, 1280 * 16, generator=torch.Generator().manual_seed(2))
    y = O.oobleck_decoder(H.split_sd(sd, "decoder."), z, st)
    e = O.oobleck_encoder(H.split_sd(sd, "encoder."), x, st)
    assert y.shape == (1, 1, 1280 * 16) and e.shape == (1, 2 * D, 16)
    assert np.abs(y.numpy() - g["dec_out"]).max() <= 5e-6
    assert np.abs(e.numpy() - g["enc_out"]).max() <= 5e-6 * max(1.0, np.abs(g["enc_out"]).max())


def test_o12_configs_3_and_4_at_full_length_match_reference():
    """BASELINE config 3's clip (latent 512, T = 375 -> 480 000 samples) and config 4 (latent 1024, T = 375, unchunked
    and decode_audio(chunked=True, chunk_size=128, overlap=32)) against the reference's outputs at the sampled points
    (clip edges, every window seam, every 97th sample) and its energy."""
    g = H.golden("o12_full")
    idx = H.t(g["idx"]).long()
    sd = H.split_sd(H.build("o12_d512", 0).state_dict(), "decoder.")
    st = H.strides_of("o12_d512")
    z = torch.randn(1, 512, 375, generator=torch.Generator().manual_seed(1))
    y = O.oobleck_decoder(sd, z, st)
    assert y.shape == (1, 1, 480000)
    assert np.abs(y[:, :, idx].numpy() - g["d512_at_idx"]).max() <= 5e-6
    assert abs(float((y.double() ** 2).sum()) - float(g["d512_sq_sum"])) <= 1e-4 * float(g["d512_sq_sum"])
    sd = H.split_sd(H.build("o12_d1024", 0).state_dict(), "decoder.")
    z = torch.randn(1, 1024, 375, generator=torch.Generator().manual_seed(1))
    dec = lambda zz: O.oobleck_decoder(sd, zz, st)
    yc = O.decode_audio_chunked(dec, z, 1280, 1, overlap=32, chunk_size=128)
    assert np.abs(yc[:, :, idx].numpy() - g["d1024_chunked_at_idx"]).max() <= 5e-6
    assert abs(float((yc.double() ** 2).sum()) - float(g["d1024_chunked_sq_sum"])) <= 1e-4 * float(g["d1024_chunked_sq_sum"])
    # the reference's own chunked and unchunked outputs agree to fp32 noise (overlap / 2 = 16 > receptive field 10)
    assert float(g["d1024_chunked_vs_full_max"]) <= 1e-6
    assert np.abs(g["d1024_chunked_at_idx"] - g["d1024_at_idx"]).max() <= 1e-6


def test_sao_full_size_decode_matches_reference():
    """BASELINE config 1: SAO-shape decoder, z [1,64,216] -> [1,2,442368], fp32 CPU."""
    g = H.golden("sao_full")
    m = H.build("sao", 0)
    sd = m.state_dict()
    H.check_checksums(sd, g)
    z = torch.randn(1, 64, 216, generator=torch.Generator().manual_seed(1))
    y = O.oobleck_decoder(H.split_sd(sd, "decoder."), z, H.strides_of("sao"))
    assert y.shape == (1, 2, 442368)
    idx = H.t(g["dec_idx"]).long()
    assert np.abs(y[:, :, idx].numpy() - g["dec_out_at_idx"]).max() <= 5e-6
    assert abs(float(y.abs().max()) - float(g["dec_abs_max"])) <= 1e-5
    assert abs(float((y.double() ** 2).sum()) - float(g["dec_sq_sum"])) <= 1e-4 * float(g["dec_sq_sum"])


def test_flop_model_matches_survey():
    # SAO decode 1089.145 GF, encode 1089.089 GF; O12 D=512 decode 1255.219 GF
    f = O.conv_flops_decoder(64, 128, [1, 2, 4, 8, 16], [2, 4, 4, 8, 8], 2, 1, 216)
    assert abs(f / 1e9 - 1089.145) < 0.01
    f = O.conv_flops_encoder(128, 128, [1, 2, 4, 8, 16], [2, 4, 4, 8, 8], 2, 1, 442368)
    assert abs(f / 1e9 - 1089.089) < 0.01
    f = O.conv_flops_decoder(512, 128, [1, 2, 4, 8, 16], [2, 4, 4, 5, 8], 1, 1, 375)
    assert abs(f / 1e9 - 1255.219) < 0.01


def _fake_backbone(M):
    def run(inputs_embeds):
        c = inputs_embeds.cumsum(dim=1) / torch.arange(1, inputs_embeds.shape[1] + 1, device=inputs_embeds.device).view(1, -1, 1)
        return torch.tanh(c @ M)
    return run


def test_lm_glue_loop_matches_reference_infer():
    """Llasa.infer of the reference (run unmodified around a stand-in backbone by make_golden.py) against the
    oracle's restatement of its per-frame glue, with and without the KL stop."""
    g = H.golden("glue")
    sd = {k[3:]: H.t(g[k]) for k in g.files if k.startswith("sd.")}
    text = H.t(g["embed_tokens"])[H.t(g["ids"]).long()].unsqueeze(0)
    noises = list(H.t(g["noise"]))
    a = O.llasa_infer(sd, _fake_backbone(H.t(g["M"])), text, H.t(g["prompt"]), noises, end_disp_kl_thres=0.0, max_length=7)
    assert a.shape == g["latents_no_stop"].shape and np.abs(a.numpy() - g["latents_no_stop"]).max() <= 1e-6
    b = O.llasa_infer(sd, _fake_backbone(H.t(g["M"])), text, H.t(g["prompt"]), noises, end_disp_kl_thres=1e9, max_length=20)
    assert b.shape == g["latents_kl_stop"].shape == (1, 64, 4) and np.abs(b.numpy() - g["latents_kl_stop"]).max() <= 1e-6


def test_dataset_side_latents_match_reference():
    """twj_dataset.py:231-256 per clip (lengths that are not multiples of the ratio; the stride-5 stage floors
    (T + 1) / 5, so 1279 samples give 32 frames, not 31)."""
    g = H.golden("dataset")
    enc = H.split_sd({k[3:]: H.t(g[k]) for k in g.files if k.startswith("sd.")}, "encoder.")
    for i in range(3):
        lat, ms = O.dataset_latents(enc, H.strides_of("tiny"), H.t(g[f"wav{i}"]), H.t(g[f"noise{i}"]))
        assert ms.shape == g[f"mean_scale{i}"].shape and np.abs(ms.numpy() - g[f"mean_scale{i}"]).max() <= 2e-6
        assert lat.shape == g[f"latents{i}"].shape and np.abs(lat.numpy() - g[f"latents{i}"]).max() <= 2e-6
    assert g["latents2"].shape[0] == 32


def test_nearest_upsample_decoder_matches_reference():
    """DecoderBlock's use_nearest_upsample branch (autoencoders.py:87-96), and the fold the CUDA plan uses for it:
    Upsample(nearest, x s) + Conv1d(k = 2s, 'same') == ConvTranspose1d(k = 3s - 1, stride s, padding s, output_padding 1)
    with summed taps."""
    g = H.golden("nearest")
    sd = {k[len("tiny_sd."):]: H.t(g[k]) for k in g.files if k.startswith("tiny_sd.")}
    y = O.oobleck_decoder(sd, H.t(g["tiny_z"]), [2, 4, 5], use_nearest_upsample=True)
    assert y.shape == g["tiny_out"].shape and np.abs(y.numpy() - g["tiny_out"]).max() <= 2e-6
    from kalle_audio_b200.layers import nearest_upsample_conv_taps
    import torch.nn.functional as F
    for s in (2, 4, 5, 8):
        w = torch.randn(5, 6, 2 * s, dtype=torch.float64)
        x = torch.randn(2, 6, 11, dtype=torch.float64)
        ref = F.conv1d(F.interpolate(x, scale_factor=s, mode="nearest"), w, padding="same")
        got = F.conv_transpose1d(x, nearest_upsample_conv_taps(w, s), stride=s, padding=s, output_padding=1)
        assert got.shape == ref.shape and float((got - ref).abs().max()) <= 1e-12


def test_bigvgan_oracle_matches_reference():
    """BigVGANFlowVAE.extract_latents / inference_from_latents of the reference (backup/flows.py:494-529), causal and
    non-causal, against the functional restatement in oracle/bigvgan_oracle.py."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import BIGVGAN_H
    from oracle import bigvgan_oracle as BO
    g = H.golden("bigvgan")
    for tag, causal in (("causal", True), ("noncausal", False)):
        h = dict(BIGVGAN_H, causal=causal)
        sd = {k[3:]: H.t(g[k]) for k in g.files if k.startswith("sd.")}
        lat = BO.extract_latents(sd, h, H.t(g[f"{tag}.x"]))
        assert lat.shape == g[f"{tag}.latents"].shape and np.abs(lat.numpy() - g[f"{tag}.latents"]).max() <= 2e-6
        y = BO.inference_from_latents(sd, h, H.t(g[f"{tag}.z"]))
        assert y.shape == g[f"{tag}.wav"].shape and np.abs(y.numpy() - g[f"{tag}.wav"]).max() <= 2e-6
        ys = BO.inference_from_latents(sd, h, H.t(g[f"{tag}.latents"]), noise=H.t(g[f"{tag}.noise"]))
        assert np.abs(ys.numpy() - g[f"{tag}.wav_sampled"]).max() <= 2e-6
