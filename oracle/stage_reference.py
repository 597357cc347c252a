"""Stage the UNMODIFIED reference sources the hot path needs into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it
travels to the GPU box like a built .so, and never enters history).

TEST / BENCH INFRASTRUCTURE.  Used by
  * bench.py --impl reference   : times the reference's own nn.Modules (blocks.py, modules_tro.py, ...) on the host cores;
  * tests that drive the reference's network_tro.ConTranModel through affganwriting_b200.install on the GPU box.

Run by __graft_entry__.build() whenever /root/reference is present (in the build container); the GPU box only uses what
was staged.  Nothing is edited: files are copied byte for byte and a manifest with their sha256 is written next to them.

    python oracle/stage_reference.py            # -> oracle/_ref/GAN_word/...
"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/GAN_word"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "GAN_word")

# the python modules reachable from network_tro / modules_tro (modules_tro.py:1-30 imports them at module level), the
# recogniser package, and the one data file load_data.py:22-29 opens at import
FILES = ["blocks.py", "modules_tro.py", "network_tro.py", "load_data.py", "loss_tro.py", "vgg_tro_channel3_modi.py",
         "Resnet18.py", "dinomodel.py", "inception.py", "inceptionrecognizer.py", "cnn.py", "trocr_recognizer.py",
         "pairs_idx_wid_iam.py", "helpers.py", "cer.py", "corpora_english/brown-azAZ.tr"]
DIRS = ["recognizer/models"]


def stage():
    if not os.path.isdir(SRC):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}

    def put(rel):
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()

    for rel in FILES:
        if os.path.isfile(os.path.join(SRC, rel)):
            put(rel)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                put(os.path.join(d, name))
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not present: nothing staged")
