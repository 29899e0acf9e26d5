"""Imports the REFERENCE's own modules (unmodified) for the benchmark's reference arm.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python with no build step (no setup.py / pyproject, so ``pip install --target baseline/_ref
/root/reference`` has nothing to install).  ``__graft_entry__.build()`` therefore places an unmodified copy of the
package the hot path lives in (``stable_audio_tools/`` plus ``model_sigmaVAE.py``) under ``baseline/_ref/`` --
git-ignored, not gpurun-ignored, so it travels to the GPU box exactly like a built ``.so`` -- and this loader imports
it from there with the un-vendored third-party imports stubbed as SURVEY.md section 8c describes:
``dac.nn.layers.WNConv1d / WNConvTranspose1d`` are the published two-liners (old-style ``torch.nn.utils.weight_norm``
over ``nn.Conv1d`` / ``nn.ConvTranspose1d``); everything else the hot path never executes is a MagicMock.

Only ``bench.py`` (``--impl reference``, ``cpu_baseline``, ``gpu_eager_baseline``) and tests use this; the product
package never imports it.  Returns None when no copy of the reference is reachable (callers then fall back to the
oracle port and say so)."""
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]
_cached = None


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "stable_audio_tools", "models", "autoencoders.py")):
            return c
    return None


def load_reference():
    """(autoencoders module, bottleneck module, root path) of the reference, or None."""
    global _cached
    if _cached is not None:
        return _cached
    root = reference_root()
    if root is None:
        return None
    import torch  # noqa: F401
    from torch import nn
    from torch.nn.utils import weight_norm
    warnings.filterwarnings("ignore")
    if root not in sys.path:
        sys.path.insert(0, root)
    dac = types.ModuleType("dac")
    dac_nn = types.ModuleType("dac.nn")
    layers = types.ModuleType("dac.nn.layers")
    layers.WNConv1d = lambda *a, **k: weight_norm(nn.Conv1d(*a, **k))
    layers.WNConvTranspose1d = lambda *a, **k: weight_norm(nn.ConvTranspose1d(*a, **k))
    layers.Snake1d = MagicMock()
    sys.modules.update({"dac": dac, "dac.nn": dac_nn, "dac.nn.layers": layers})
    for m in ["dac.nn.quantize", "dac.model", "dac.model.dac", "dac.model.discriminator", "alias_free_torch",
              "vector_quantize_pytorch", "k_diffusion", "x_transformers", "einops_exts", "audiotools", "encodec", "pywt"]:
        sys.modules.setdefault(m, MagicMock())
    from stable_audio_tools.models import autoencoders, bottleneck
    _cached = (autoencoders, bottleneck, root)
    return _cached


def load_reference_discriminators():
    """The reference's stable_audio_tools/models/discriminators.py, imported unmodified from the same copy (audiotools /
    dac.model.discriminator / encodec -- which OobleckDiscriminator never touches -- stubbed), or None."""
    root = reference_root()
    if root is None:
        return None
    path = os.path.join(root, "stable_audio_tools", "models", "discriminators.py")
    if not os.path.isfile(path):
        return None
    import importlib.util
    warnings.filterwarnings("ignore")
    for m in ["audiotools", "dac", "dac.model", "dac.model.discriminator", "encodec", "encodec.msstftd"]:
        sys.modules.setdefault(m, MagicMock())
    spec = importlib.util.spec_from_file_location("ref_discriminators", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
