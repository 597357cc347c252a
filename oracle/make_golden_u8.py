"""Golden vectors for the uint8 wire format, produced by the UNMODIFIED reference loader:
`IAM_words.read_image_single` (GAN_word/load_data.py:141-167) is run on PNG files written to a temporary directory while a
spy on cv2.resize records the uint8 image it hands to the normalisation; the fixture holds those uint8 images and the float32
canvases the reference returned.  Container-only:  python -m oracle.make_golden_u8      (TEST INFRASTRUCTURE)"""
import os
import tempfile

import numpy as np

from oracle import affgw_oracle as O
from oracle import ref_bootstrap as rb

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    import cv2
    ns = rb.load(50)
    ld = ns.load_data
    rng = np.random.RandomState(7)
    ds = ld.IAM_words({}, True)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        old_base, real_resize = ld.img_base, cv2.resize
        ld.img_base = tmp
        seen = []

        def spy(*a, **k):
            r = real_resize(*a, **k)
            seen.append(r.copy())
            return r

        cv2.resize = spy
        try:
            # (height, width) of the source scans: narrower than, equal to and wider than the 216-column canvas after resize
            for i, (h, w) in enumerate([(64, 100), (37, 90), (128, 431), (50, 400), (64, 215), (91, 13)]):
                src = rng.randint(0, 256, size=(h, w)).astype(np.uint8)
                src[:, : w // 3] = 255                                  # blank paper
                src[:, w // 3: w // 2] = 0                              # solid ink
                cv2.imwrite(os.path.join(tmp, f"w{i}.png"), src)
                ref, ref_width = ds.read_image_single(f"w{i}")
                u8 = seen[-1]
                assert u8.dtype == np.uint8 and u8.shape[0] == 64
                mine, my_width = O.normalize_resized_u8(u8)
                assert ref.dtype == np.float32 and mine.dtype == np.float32
                assert my_width == ref_width and np.array_equal(mine, ref), i
                wire = O.pad_resized_u8(u8)
                assert np.array_equal(O.decode_u8(wire), ref), i
                out[f"u8.{i}"] = u8
                out[f"ref.{i}"] = ref
                out[f"width.{i}"] = np.int64(ref_width)
                print(f"w{i}: source {h}x{w} -> resized {u8.shape}, width {ref_width}: oracle == reference bit for bit")
            out["count"] = np.int64(i + 1)
        finally:
            cv2.resize = real_resize
            ld.img_base = old_base
    np.savez_compressed(os.path.join(OUT, "u8_wire.npz"), **out)
    print("wrote u8_wire.npz,", os.path.getsize(os.path.join(OUT, "u8_wire.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
