"""Make the reference's own scripts use this implementation without editing them.

    import affganwriting_b200.install as inst; inst.install()          # before `import network_tro` / main_run

replaces, inside the reference's already-importable modules `blocks` and `modules_tro`, the classes that sit on the
accelerated path by their drop-in counterparts, so `network_tro.ConTranModel`, `main_run.py` and the
`tt.test_single_writer.*` scripts pick them up through their normal `from modules_tro import ...` statements.
The reference tree must be on sys.path (it is not shipped here); see INTEGRATION.md.
"""
import importlib
import sys

from . import blocks as _blocks
from . import modules_tro as _modules
from . import resnet_encoder as _resnet

BLOCK_NAMES = ("Conv2dBlock", "ResBlock", "ResBlocks", "ActFirstResBlock", "LinearBlock", "AdaptiveInstanceNorm2d", "iAFF",
               "get_key", "mean_variance_norm")
MODULE_NAMES = ("GenModel_FC", "DisModel", "WriterClaModel", "TextEncoder_FC", "ImageEncoder", "Decoder", "MLP",
                "get_num_adain_params")


_saved = []      # (module object, attribute, original) of everything install() replaced


def _swap(mod, mod_name, attr, new, done):
    _saved.append((mod, attr, getattr(mod, attr)))
    setattr(mod, attr, new)
    done.append((mod_name, attr))


def install(ref_blocks="blocks", ref_modules="modules_tro", ref_network="network_tro", recogniser=True):
    """Patch the reference modules in place. Returns the list of (module, attribute) pairs that were replaced.
    recogniser=False keeps the reference's own RecModel (plain PyTorch) behind the native generator / discriminator / classifier.
    `network_tro` binds the model classes by name at import (network_tro.py:4): when it is already imported its bindings are
    replaced too, so install() works before or after `import network_tro`."""
    done = []
    rb = sys.modules.get(ref_blocks) or importlib.import_module(ref_blocks)
    for n in BLOCK_NAMES:
        if hasattr(rb, n):
            _swap(rb, ref_blocks, n, getattr(_blocks, n), done)
    rm = sys.modules.get(ref_modules) or importlib.import_module(ref_modules)
    for n in MODULE_NAMES + BLOCK_NAMES:
        src = _modules if n in MODULE_NAMES else _blocks
        if hasattr(rm, n):
            _swap(rm, ref_modules, n, getattr(src, n), done)
    if recogniser and hasattr(rm, "RecModel"):
        _swap(rm, ref_modules, "RecModel", _modules.RecModel, done)
    # the encoder the reference's GenModel_FC constructs by default (modules_tro.py:219)
    if hasattr(rm, "ImageEncoderResNet50"):
        _swap(rm, ref_modules, "ImageEncoderResNet50", _resnet.ImageEncoderResNet50, done)
    # the DINOv2 wrapper (dinomodel.py; modules_tro.py:29 imports the class by name) - generation only
    from . import dinomodel as _dino
    if hasattr(rm, "ImageEncoderDINOv2"):
        _swap(rm, ref_modules, "ImageEncoderDINOv2", _dino.ImageEncoderDINOv2, done)
    dm = sys.modules.get("dinomodel")
    if dm is not None and hasattr(dm, "ImageEncoderDINOv2"):
        _swap(dm, "dinomodel", "ImageEncoderDINOv2", _dino.ImageEncoderDINOv2, done)
    nt = sys.modules.get(ref_network)
    if nt is not None:
        for n in ("GenModel_FC", "DisModel", "WriterClaModel") + (("RecModel",) if recogniser else ()):
            if hasattr(nt, n):
                _swap(nt, ref_network, n, getattr(_modules, n), done)
    # the stand-alone ResNet-18 (Resnet18.py), when the caller has imported it
    r18 = sys.modules.get("Resnet18")
    if r18 is not None:
        from . import Resnet18 as _r18
        for n in ("conv3x3", "BasicBlock", "ResNet18"):
            _swap(r18, "Resnet18", n, getattr(_r18, n), done)
    return done


def install_line_generation(ref_module="model.pure_gen"):
    """Same for the line-level generator of `line_generation/` (generation only): replaces SpacedGenerator and its building
    blocks inside the reference's `model.pure_gen`, so `HWWithStyle` (hw_with_style.py:204) builds the libaffgw generator."""
    from . import linegen as _lg
    done = []
    mod = sys.modules.get(ref_module) or importlib.import_module(ref_module)
    for n in ("SpacedGenerator", "StyledConvBlock", "AdaptiveInstanceNorm", "NoiseInjection", "Blur", "FusedUpsample",
              "EqualConv2d", "PixelNorm"):
        if hasattr(mod, n):
            _swap(mod, ref_module, n, getattr(_lg, n), done)
    hw = sys.modules.get("model.hw_with_style")
    if hw is not None and hasattr(hw, "SpacedGenerator"):
        _swap(hw, "model.hw_with_style", "SpacedGenerator", _lg.SpacedGenerator, done)
    return done


def uninstall():
    """Undo every install() of this process (tests share one interpreter with the golden-vector generators)."""
    while _saved:
        mod, attr, orig = _saved.pop()
        setattr(mod, attr, orig)
