"""TEST INFRASTRUCTURE - never imported by the product (tests/test_cabi.py::test_product_never_imports_the_oracle).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs (`--impl reference`, `cpu_baseline`) may import this package,
and only as the checker / the reported CPU baseline.  Parity is PINNED: every restatement below is checked against outputs of the
UNMODIFIED reference, imported in place from /root/reference by the committed generator scripts (the reference ships no tests
or golden vectors of its own).

    affgw_oracle.py     blocks / generator / discriminator / classifier / step losses (GAN_word/blocks.py, modules_tro.py,
                        network_tro.py)                          <- make_golden.py, make_golden_resnet*.py, make_golden_u8.py
    rec_oracle.py       RecModel incl. the per-sample beam search (recognizer/models/*)          <- make_golden_rec.py
    linegen_oracle.py   line-level generator (line_generation/model/pure_gen.py)                  <- make_golden_linegen.py
    dino_oracle.py      DINOv2 wrapper (dinomodel.py) + the public ViT definition of its absent hub backbone
                                                                  <- make_golden_dino.py (reference wrapper),
                                                                     make_golden_dino_hf.py (transformers.Dinov2Model)
    weights.py          seeded, platform-independent parameter tensors from {key: shape} specs
    ref_bootstrap.py    imports the reference on the CPU without touching its files
    stage_reference.py  byte-for-byte copy (sha256 manifest) of the reference files the path imports into oracle/_ref/
                        (git-ignored; travels to the GPU box for the CPU baseline arm and the drop-in tests)
"""
