"""Golden vectors of the DINOv2 style-encoder WRAPPER produced by the UNMODIFIED reference (GAN_word/dinomodel.py, imported in
place) around a small stand-in backbone (oracle/dino_oracle.StandInViT answers the wrapper's torch.hub.load: the real DINOv2
checkout is an absent, un-vendored dependency).  Container-only:  python -m oracle.make_golden_dino     (TEST INFRASTRUCTURE)
Writes tests/golden/dino.npz + dino_spec.json and checks oracle.dino_oracle.dino_encoder against the reference wrapper."""
import importlib.util
import json
import os

import numpy as np
import torch

from oracle import affgw_oracle as O
from oracle import dino_oracle as DO
from oracle import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ARCH = dict(embed_dim=192, depth=6, num_heads=3)
TAPS = [1, 2, 4, 5]


def main():
    sp = importlib.util.spec_from_file_location("ref_dinomodel", "/root/reference/GAN_word/dinomodel.py")
    ref = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(ref)
    real = torch.hub.load
    torch.hub.load = lambda *a, **k: DO.StandInViT(**ARCH)
    try:
        enc = ref.ImageEncoderDINOv2("/nonexistent/dinov2", arch="vitl14", ckpt_path=None, in_channels=50, final_size=(8, 27),
                                     tap_blocks=TAPS)
    finally:
        torch.hub.load = real
    spec = W.spec_of(enc)
    sd = W.make_state(spec)
    for k in spec:                                  # LayerNorm / LayerScale scales around 1 (make_state centres 1-D tensors on 0)
        if k.endswith(("norm1.weight", "norm2.weight", "norm.weight", ".gamma")):
            sd[k] = sd[k] + 1.0
    enc.load_state_dict(sd)
    enc.eval()
    x = O.synthetic_batch(2, 50)["tr_img"]
    with torch.no_grad():
        ref_out = enc(x)
        mine = DO.dino_encoder(x, sd, ARCH["num_heads"], TAPS)
    report = []
    for i, (a, b) in enumerate(zip(mine, ref_out)):
        assert a.shape == b.shape, (i, a.shape, b.shape)
        e = float((a - b).abs().max() / max(1.0, float(b.abs().max())))
        report.append({"name": f"dino.result{i}", "max_abs": e, "tol": 1e-5})
        print(f"result{i} {tuple(b.shape)}: oracle vs reference wrapper {e:.2e}, |map| max {float(b.abs().max()):.2f}")
        assert e <= 1e-5
    out = {"result1": ref_out[1].numpy(), "result4": ref_out[4].numpy()}
    for i, r in enumerate(ref_out):
        out[f"result{i}.abs_mean"] = np.float32(r.abs().mean())
        out[f"result{i}.shape"] = np.array(r.shape)
    np.savez_compressed(os.path.join(OUT, "dino.npz"), **out)
    json.dump({"spec": spec, "arch": ARCH, "taps": TAPS, "report": report}, open(os.path.join(OUT, "dino_spec.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
