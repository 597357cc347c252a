"""CPU: the C-ABI library loads, exports every symbol include/affgw.h declares, the Python binding covers them all,
the drop-in classes reproduce the reference's checkpoint key layout, and the product never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "affgw.h")).read()
    return sorted(set(re.findall(r"\b(affgw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from affganwriting_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(h, n), n
    assert h.affgw_version() == 107


def test_python_binding_covers_the_header():
    from affganwriting_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.lib()


def test_conv_descriptor_layout_matches_header():
    from affganwriting_b200 import _lib
    src = open(os.path.join(ROOT, "include", "affgw.h")).read()
    body = src[src.index("typedef struct affgw_conv_desc {"):src.index("} affgw_conv_desc;")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0]
        if "int32_t" in line:
            fields += [f.strip() for f in line.replace("int32_t", "").replace(";", "").split(",") if f.strip()]
    assert fields == [f[0] for f in _lib.ConvDesc._fields_]
    assert ctypes.sizeof(_lib.ConvDesc) == 4 * len(fields)


def test_host_side_argument_validation_needs_no_gpu():
    from affganwriting_b200 import _lib
    h = _lib.lib()
    d = _lib.ConvDesc()
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0
    assert b"non-positive" in h.affgw_last_error()
    d.N, d.H, d.W, d.Cin, d.Cout, d.KH, d.KW = 2, 8, 27, 512, 512, 3, 3
    d.stride, d.pad, d.pad_mode, d.upsample, d.Ho, d.Wo = 1, 1, 1, 1, 8, 27
    d.in_pitch, d.out_pitch, d.x_dtype, d.w_dtype, d.y_dtype = 512, 512, 1, 1, 1
    d.algo, d.passes, d.grad_dtype = 2, 3, 0
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 1
    assert h.affgw_conv2d_dgrad_ws_bytes(ctypes.byref(d)) == 2 * 10 * 29 * 512 * 4     # reflect: folded path, fp32 dx
    d.pad_mode = 0
    assert h.affgw_conv2d_dgrad_ws_bytes(ctypes.byref(d)) == 0                          # zero pad: direct
    d.passes = 2
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0 and b"passes" in h.affgw_last_error()
    d.passes = 1
    d.Ho = 9
    assert h.affgw_conv_tc_supported(ctypes.byref(d)) == 0 and b"output extent" in h.affgw_last_error()
    for layout in (1, 2):                                                               # im2col tiles / shifted-kernel planes
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 1, layout) == 512 * 512 * 9 * 2
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 3, layout) == 512 * 512 * 9 * 2 * 2
        assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 510, 0, 1, layout) < 0     # c_store must be a multiple of 8
    assert h.affgw_pack_weight_tc_bytes(512, 512, 3, 3, 512, 0, 1, 7) < 0
    d.Ho, d.passes = 8, 3
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 2 and h.affgw_conv_tc_layout(ctypes.byref(d), 1) == 2
    assert h.affgw_conv_tc_prefer_shift(0) == 1
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 1
    assert h.affgw_conv_tc_prefer_shift(1) == 0
    d.stride, d.Ho, d.Wo = 2, 4, 14                                                      # strided: im2col kernel only
    assert h.affgw_conv_tc_layout(ctypes.byref(d), 0) == 1
    assert h.affgw_operand_planes_bytes(100, 56, 3) == 100 * 56 * 2 * 2


@pytest.mark.parametrize("key,build", [("gen_c50", "gen"), ("dis", "dis"), ("cla", "cla")])
def test_state_dict_keys_match_reference(key, build, specs):
    from affganwriting_b200 import modules_tro as M
    m = {"gen": lambda: M.GenModel_FC(12), "dis": M.DisModel, "cla": lambda: M.WriterClaModel(500)}[build]()
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == specs[key]
    assert list(mine) == list(specs[key])          # same registration order, as torch.save would write it


def test_block_constructor_signatures(specs):
    from affganwriting_b200 import blocks as B
    for name, sp in specs.items():
        if not name.startswith("blocks.conv_"):
            continue
        m = B.Conv2dBlock(**sp["ctor"])
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == sp["spec"], name
    m = B.AdaptiveInstanceNorm2d(512)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == specs["blocks.adain_plain"]["spec"]
    with pytest.raises(AssertionError):
        B.Conv2dBlock(4, 4, 3, 1, norm="groupnorm")
    with pytest.raises(AssertionError):
        m(__import__("torch").zeros(1, 512, 2, 2))      # "Please assign AdaIN weight first"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "affganwriting_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", "").replace("CPU oracle", "") or fn == "__never__", fn


def test_missing_library_fails_loudly(monkeypatch):
    from affganwriting_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libaffgw.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
