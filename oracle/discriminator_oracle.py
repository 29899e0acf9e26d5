"""CPU restatement of the reference's Oobleck discriminator.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU and
torch-eager legs): the product package never imports this file.

Follows /root/reference/stable_audio_tools/models/discriminators.py:
  * get_hinge_losses                      :11-14
  * SharedDiscriminatorConvNet            :62-116   n_layers=4, capacity=32, kernel_size=15, stride=4, SiLU,
                                                    weight_norm on the strided convs, plain 1x1 conv last; the features
                                                    are the outputs of every conv (before the activation)
  * MultiScaleDiscriminator               :119-138  3 nets on x, avg_pool1d(x, 2), avg_pool1d(avg_pool1d(x, 2), 2)
  * MultiPeriodDiscriminator              :140-168  Conv2d nets (an int kernel_size makes the kernel 15 x 15, stride
                                                    (4, 4), padding (7, 7)) on x folded to [B, C, ceil(T / n), n]
  * MultiDiscriminator                    :171-238  reals and fakes concatenated on the batch axis, scores summed over
                                                    the discriminators, features concatenated into one flat list
  * OobleckDiscriminator.loss             :240-297  hinge losses on the summed scores; the feature-matching distance
                                                    iterates over the flat list, so each term is the mean over the batch
                                                    of the per-item mean |real - fake| of ONE feature tensor
Plain functional code over a reference-keyed ``state_dict`` (torch CPU fp32 ops, i.e. the library kernels the
reference's modules call).  Pinned against the reference itself by tests/golden/make_golden.py (``disc.npz``)."""
import torch
import torch.nn.functional as F

PERIODS = (2, 3, 5, 7, 11)
N_SCALES = 3
N_LAYERS = 4
KERNEL = 15
STRIDE = 4


def _fold_wn(sd, prefix):
    """old-style torch.nn.utils.weight_norm, dim 0"""
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    dims = tuple(range(1, v.dim()))
    return v * (g / v.norm(2, dim=dims, keepdim=True))


def shared_convnet(sd, prefix, x, two_d):
    """SharedDiscriminatorConvNet.forward (:108-116): (score [B], features)"""
    conv = F.conv2d if two_d else F.conv1d
    feats = []
    for i in range(N_LAYERS):
        p = f"{prefix}.net.{2 * i}"
        x = conv(x, _fold_wn(sd, p), sd[p + ".bias"], stride=STRIDE, padding=KERNEL // 2)
        feats.append(x)
        x = F.silu(x)
    p = f"{prefix}.net.{2 * N_LAYERS}"
    x = conv(x, sd[p + ".weight"], sd[p + ".bias"])
    feats.append(x)
    return x.reshape(x.shape[0], -1).mean(-1), feats


def period_fold(x, n):
    """MultiPeriodDiscriminator.fold (:164-168)"""
    pad = (n - (x.shape[-1] % n)) % n
    x = F.pad(x, (0, pad))
    return x.reshape(*x.shape[:2], -1, n)


def multi_scale(sd, prefix, x):
    score, feats = 0, []
    for i in range(N_SCALES):
        s, f = shared_convnet(sd, f"{prefix}.layers.{i}", x, False)
        score = score + s
        feats.extend(f)
        x = F.avg_pool1d(x, 2)
    return score, feats


def multi_period(sd, prefix, x):
    score, feats = 0, []
    for i, n in enumerate(PERIODS):
        s, f = shared_convnet(sd, f"{prefix}.layers.{i}", period_fold(x, n), True)
        score = score + s
        feats.extend(f)
    return score, feats


def oobleck_discriminator(sd, x, prefix="multi_discriminator"):
    """scores [N] and the flat feature list of MultiDiscriminator.forward for a batch x [N, C, T]"""
    s0, f0 = multi_scale(sd, f"{prefix}.discriminators.0", x)
    s1, f1 = multi_period(sd, f"{prefix}.discriminators.1", x)
    return s0 + s1, f0 + f1


def hinge_losses(score_real, score_fake):
    gen_loss = -score_fake.mean()
    dis_loss = torch.relu(1 - score_real).mean() + torch.relu(1 + score_fake).mean()
    return dis_loss, gen_loss


def oobleck_discriminator_loss(sd, reals, fakes):
    """OobleckDiscriminator.loss (:269-297): (dis_loss, gen_loss, feature_matching_distance)"""
    B = reals.shape[0]
    scores, feats = oobleck_discriminator(sd, torch.cat([reals, fakes], 0))
    dis_loss, gen_loss = hinge_losses(scores[:B], scores[B:])
    fm = torch.tensor(0.)
    for f in feats:
        real, fake = f[:B], f[B:]
        fm = fm + sum((real[b] - fake[b]).abs().mean() for b in range(B)) / B
    return dis_loss, gen_loss, fm
