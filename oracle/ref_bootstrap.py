"""Container-only bootstrap that imports the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE, NOT PRODUCT.  Nothing under affganwriting_b200/ may import this.
It exists to (a) validate oracle/affgw_oracle.py against the real reference and
(b) generate the golden vectors under tests/golden/ (see oracle/make_golden.py).
/root/reference does not exist on the GPU box; there the staged copy under oracle/_ref/ (oracle/stage_reference.py) is
used by `bench.py --impl reference` and by the install() test that drives the reference's ConTranModel.

Shims follow SURVEY.md Appendix D; no reference file is modified or copied:
  1. builtins.open redirect of the hard-coded corpus path      (GAN_word/load_data.py:22-29)
  2. stub `Levenshtein` module                                  (GAN_word/loss_tro.py:2)
  3. modules_tro.gpu -> cpu, Tensor.cuda() no-op                (GAN_word/modules_tro.py:33,308)
  4. GenModel_FC built without its ctor (ResNet-50 weight load) (GAN_word/modules_tro.py:208-224)
  5. NUM_CHANNEL override for the 15-style-image config         (GAN_word/load_data.py:14-15)
"""
import builtins
import os
import sys
import types

# the reference tree where it lies (build container), else the byte-for-byte staged copy of the files this path needs
# (oracle/stage_reference.py -> oracle/_ref/, git-ignored; it travels to the GPU box with the snapshot)
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference/GAN_word") else _STAGED
REF_WORD = os.path.join(REF_ROOT, "GAN_word")
_HOME_PREFIX = "/home/woody/iwi5/iwi5333h/AFFGanWriting/"

_state = {}


def available():
    return os.path.isdir(REF_WORD)


def _lev_distance(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def load(num_channel=50):
    """Import the reference modules on CPU. Returns a namespace with blocks, modules_tro,
    load_data, vgg and a `Gen` class (GenModel_FC with the VGG ImageEncoder wiring)."""
    import torch
    from torch import nn

    if "ns" in _state:
        ns = _state["ns"]
        if ns.load_data.NUM_CHANNEL != num_channel:
            ns.load_data.NUM_CHANNEL = num_channel
            ns.vgg.NUM_CHANNEL = num_channel
        return ns
    if not available():
        raise RuntimeError("reference tree not present: " + REF_WORD)

    real_open = builtins.open

    def patched_open(file, *a, **k):
        if isinstance(file, str) and file.startswith(_HOME_PREFIX):
            file = os.path.join(REF_ROOT, file[len(_HOME_PREFIX):])
        return real_open(file, *a, **k)

    lev = types.ModuleType("Levenshtein")
    lev.distance = _lev_distance
    sys.modules.setdefault("Levenshtein", lev)

    sys.path.insert(0, REF_WORD)
    builtins.open = patched_open
    try:
        import load_data  # noqa
        load_data.NUM_CHANNEL = num_channel
        import vgg_tro_channel3_modi as vgg  # noqa
        vgg.NUM_CHANNEL = num_channel
        import blocks  # noqa
        import modules_tro  # noqa
    finally:
        builtins.open = real_open

    if not torch.cuda.is_available() or os.environ.get("AFFGW_REF_CPU", "0") == "1":
        modules_tro.gpu = torch.device("cpu")
        torch.Tensor.cuda = lambda self, *a, **k: self  # modules_tro.py:308
        import recognizer.models.encoder_vgg as _ev
        _ev.cuda = torch.device("cpu")                   # encoder_vgg.py:25

    class Gen(modules_tro.GenModel_FC):
        """GenModel_FC with the VGG ImageEncoder (modules_tro.py:211 commented line)."""

        def __init__(self, text_max_len=12, encoder=None):
            nn.Module.__init__(self)
            self.enc_image = encoder if encoder is not None else modules_tro.ImageEncoder()
            self.enc_text = modules_tro.TextEncoder_FC(text_max_len)
            self.dec = modules_tro.Decoder()
            self.linear_mix = nn.Linear(1024, 512)
            self.max_conv = nn.MaxPool2d(kernel_size=2, stride=2)

    ns = types.SimpleNamespace(load_data=load_data, vgg=vgg, blocks=blocks,
                               modules_tro=modules_tro, Gen=Gen)
    _state["ns"] = ns
    return ns


def load_network(num_channel=50):
    """The reference's network_tro module (ConTranModel) on top of load(); RecModel is built without its pre-trained
    VGG download (encoder_vgg.py:30, SURVEY.md §8(c) shim 6)."""
    import importlib

    import torch
    ns = load(num_channel)
    import recognizer.models.encoder_vgg as ev
    ev.PRE_TRAIN_VGG = False
    nt = sys.modules.get("network_tro") or importlib.import_module("network_tro")
    if not torch.cuda.is_available() or os.environ.get("AFFGW_REF_CPU", "0") == "1":
        nt.gpu = torch.device("cpu")
    ns.network_tro = nt
    return ns


def gen_forward(g, tr_img, label):
    """network_tro.py:60-66 composition."""
    f_xss = g.enc_image(tr_img)
    f_xt, f_embed = g.enc_text(label, f_xss[-1].shape)
    f_mix = g.mix(f_xss, f_embed)
    return g.decode(f_mix, f_xss, f_embed, f_xt)
