"""CPU checks of bench.py's host side: the JSON contract of the reference arm (the line the driver divides by), the clock /
throttle-reason summary, the synthetic batch in the wire format.  No GPU work, no timing of the real reference (the timed
function is replaced by a stub - what is pinned here is the arithmetic around it and the keys of the line)."""
import argparse
import json

import numpy as np
import torch

import bench


def _args(**kw):
    return argparse.Namespace(gpus=kw.get("gpus", 1), steps=kw.get("steps", 4), warmup=kw.get("warmup", 1))


def test_reference_arm_line(monkeypatch, capsys):
    calls = {}

    def fake(batch, steps, warmup, threads):
        calls["args"] = (batch, steps, warmup, threads)
        return [2.0] * steps                                    # seconds per 8-sample iteration

    monkeypatch.setattr(bench, "reference_modules_steps", fake)
    monkeypatch.setattr("oracle.ref_bootstrap.available", lambda: True)
    bench.run_reference_arm(_args(), rank=0)
    line = json.loads(capsys.readouterr().out.strip())
    assert calls["args"][:3] == (bench.REFERENCE_SAMPLE_BATCH, 4, 1)              # every W + K step is executed
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "steps/s"
    # 8 of the 64 samples in 2 s -> one 64-sample step in 16 s
    assert abs(line["value"] - (bench.REFERENCE_SAMPLE_BATCH / bench.BATCH_PER_GPU) / 2.0) < 1e-12
    assert line["ms_per_step"] == 2000.0 and line["sample_fraction_of_step"] == bench.REFERENCE_SAMPLE_BATCH / bench.BATCH_PER_GPU
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["config"]["workload"] == bench.WORKLOAD
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == line["value"] and cb["cores"] >= 1 and "8 of the 64" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 4 and line["warmup"] == 1 and line["n_gpus"] == 1


def test_reference_arm_other_ranks_print_nothing(monkeypatch, capsys):
    monkeypatch.setattr(bench, "reference_modules_steps", lambda *a: (_ for _ in ()).throw(AssertionError("rank 1 must not run")))
    bench.run_reference_arm(_args(gpus=2), rank=1)
    assert capsys.readouterr().out == ""


def test_reference_arm_says_port_without_the_staged_tree(monkeypatch, capsys):
    monkeypatch.setattr("oracle.ref_bootstrap.available", lambda: False)
    monkeypatch.setattr(bench, "cpu_reference_steps", lambda b, s, w, t: 4.0)
    bench.run_reference_arm(_args(steps=2), rank=0)
    line = json.loads(capsys.readouterr().out.strip())
    assert line["cpu_baseline"]["kind"] == "port" and "oracle port" in line["cpu_baseline"]["sample"]
    assert abs(line["value"] - 0.125 / 4.0) < 1e-12


def test_clock_summary_reads_throttle_reasons():
    s = bench.ClockSampler(0)
    s.rows = [["1837", "1965", "950.1", "Not Active", "Not Active", "Not Active", "Active"],
              ["1845", "1965", "948.0", "Not Active", "Not Active", "Not Active", "Not Active"],
              ["1830", "1965", "951.3", "Not Active", "Active", "Not Active", "Active"],
              ["[N/A]", "x", "", "", "", "", ""],              # a malformed sample is dropped
              ["short"]]
    out = s.summary()
    assert out["sm_mhz"] == 1837.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    assert bench.ClockSampler(0).summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_synthetic_batch_is_the_loader_tuple_in_wire_format():
    from affganwriting_b200 import load_data as ld
    b = bench.synthetic_batch(3, 5, seed=7)
    assert len(b) == 9                                                             # main_run.py:108-118
    tr_img, widths, tr_label, img_xt, label_xt, label_xt_swap = b[3], b[4], b[5], b[6], b[7], b[8]
    assert tr_img.dtype == torch.uint8 and tuple(tr_img.shape) == (3, 5, ld.IMG_HEIGHT, ld.IMG_WIDTH)
    assert img_xt.dtype == torch.uint8 and tuple(img_xt.shape) == (3, 1, ld.IMG_HEIGHT, ld.IMG_WIDTH)
    assert tuple(tr_label.shape) == (3, 5, ld.OUTPUT_MAX_LEN) and tuple(label_xt.shape) == (3, ld.OUTPUT_MAX_LEN)
    assert label_xt_swap.dtype == torch.int64
    # right of each image's width: background (255), like the loader's padding (load_data.py:153-166)
    for n in range(3):
        for c in range(5):
            w = int(widths[n, c])
            assert 40 <= w <= ld.IMG_WIDTH and bool((tr_img[n, c, :, w:] == 255).all())
    # labels: <GO> first, <END> once, then padding only (load_data.py:169-179)
    lab = label_xt.numpy()
    assert (lab[:, 0] == 0).all()
    for row in lab:
        end = int(np.where(row == 1)[0][0])
        assert (row[end + 1:] == 2).all() and (row[1:end] >= 3).all()
    # seeded: the same call gives the same batch, and its byte count is what e2e reports as h2d_bytes_per_step
    b2 = bench.synthetic_batch(3, 5, seed=7)
    assert all(torch.equal(x, y) for x, y in zip(b, b2) if torch.is_tensor(x))
    assert bench.batch_bytes(b) == sum(t.numel() * t.element_size() for t in b if torch.is_tensor(t))
