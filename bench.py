"""Benchmark of the hot path: one full GAN training iteration (cla_update -> dis_update -> gen_update, forward + backward
+ Adam, no recogniser) at batch 64 per GPU, 50 style planes, 64x216 synthetic IAM-shaped words (BASELINE.json configs[1]).
Activations, gradients and parameters are fp32 in HBM; the convolutions run on tcgen05 tensor cores over 16-bit operand
planes (fp16 by default, `--precision bf16` for bf16 planes).

    python bench.py --gpus N --steps K --warmup W            # this implementation (torchrun for N > 1)
    python bench.py --impl reference ...                      # the reference's own nn.Modules (staged in oracle/_ref) on the host cores

Prints ONE JSON line on rank 0 (see DESIGN.md "measurement" for every field).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:] and "--impl" in sys.argv[1:]:
    # the reference arm is a CPU measurement: hide the GPUs before torch initialises CUDA, so that nothing in the reference
    # (nn.DataParallel around the VGG slices, modules_tro.py:341-346; `.cuda()` at :308) reaches for a device
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH_PER_GPU = 64
NUM_CHANNEL = 50
STEP_GFLOP_PER_SAMPLE = 389.7        # cla 6.14 + dis_update 108.4 + gen_update 275.2 (no recogniser)
METRIC = "train_steps_per_sec"
WORKLOAD = ("configs[1]: full GAN training step (GenModel_FC + DisModel + WriterClaModel fwd/bwd + Adam, no recogniser), "
            "bf16, batch 64 per GPU, 50 style planes 64x216, synthetic IAM-shaped words")


# --------------------------------------------------------------------------------------------------- synthetic data
def synthetic_batch(batch, num_channel, seed):
    """The 9-tuple main_run.sort_batch builds (main_run.py:108-118), filled with seeded synthetic data.
    Images are in the uint8 wire format (grey levels as the loader's cv2.resize leaves them, right padding 255 =
    background): `load_data.batch_to_device` normalises them on the GPU exactly like load_data.py:152-166,
    which gives ink uniform over the 256 levels of (-1, 1] and -1 to the right of a random width."""
    from affganwriting_b200 import load_data as ld
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    letters = list(ld.letter2index)

    def imgs(n, c):
        x = torch.randint(0, 256, (n, c, ld.IMG_HEIGHT, ld.IMG_WIDTH), generator=g, dtype=torch.uint8)
        widths = torch.randint(40, ld.IMG_WIDTH + 1, (n, c), generator=g)
        cols = torch.arange(ld.IMG_WIDTH).view(1, 1, 1, -1)
        return torch.where(cols >= widths.view(n, c, 1, 1), torch.full_like(x, 255), x), widths

    def labels(shape):
        flat = []
        for _ in range(int(np.prod(shape))):
            ln = rng.randint(1, ld.MAX_CHARS + 1)
            flat.append(ld.label_padding("".join(letters[i] for i in rng.randint(0, len(letters), ln))))
        return torch.tensor(flat, dtype=torch.int64).view(*shape, ld.OUTPUT_MAX_LEN)

    tr_img, widths = imgs(batch, num_channel)
    img_xt, _ = imgs(batch, 1)
    return (np.zeros(batch, dtype=np.int64),
            torch.randint(0, ld.NUM_WRITERS, (batch,), generator=g),
            np.arange(batch),
            tr_img, widths, labels((batch, num_channel)), img_xt, labels((batch,)), labels((batch,)))


def batch_bytes(batch):
    return int(sum(t.numel() * t.element_size() for t in batch if torch.is_tensor(t)))


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.thread.join(6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_steps(batch, steps, warmup, threads):
    """The reference's algorithm for the path (CPU oracle port, oracle/affgw_oracle.py: fp32 PyTorch ops on the host
    cores), one cla+dis+gen update per step at `batch` samples.  Returns seconds per step."""
    from oracle import affgw_oracle as O
    from oracle import weights as W
    torch.set_num_threads(threads)
    spec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_spec.json")))
    full = {}
    for pre, key in (("gen.", "gen_c50"), ("dis.", "dis"), ("cla.", "cla")):
        for k, v in W.make_state(spec[key]).items():
            full[pre + k] = v.clone().requires_grad_(v.is_floating_point())
    data = O.synthetic_batch(batch, NUM_CHANNEL)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in full.values():
            p.grad = None
        O.cla_update(data, full).backward()
        l_real, l_fake = O.dis_update(data, full)
        (l_real + l_fake).backward()
        O.gen_update(data, full)[0].backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def reference_modules_steps(batch, steps, warmup, threads):
    """The REFERENCE's own modules (GAN_word/blocks.py, modules_tro.py, vgg_tro_channel3_modi.py - unmodified, imported from
    /root/reference or from the staged copy oracle/_ref/) on the host cores: one iteration of main_run.py:146-167 without the
    recogniser = cla_update, dis_update, gen_update composed exactly like network_tro.py:50-138 (minus the l_rec lines),
    each followed by its torch.optim.Adam step (main_run.py:275-278).  Returns the list of per-step wall times (seconds)."""
    os.environ["AFFGW_REF_CPU"] = "1"                    # keep the reference on the CPU even though the box has a GPU
    from oracle import affgw_oracle as O
    from oracle import ref_bootstrap as rb
    torch.set_num_threads(threads)
    ns = rb.load(NUM_CHANNEL)
    m = ns.modules_tro
    torch.manual_seed(0)
    gen, dis, cla = ns.Gen(12).train(), m.DisModel().train(), m.WriterClaModel(500).train()
    opt = {"cla": torch.optim.Adam(cla.parameters(), lr=1e-5), "dis": torch.optim.Adam(dis.parameters(), lr=1e-4),
           "gen": torch.optim.Adam(gen.parameters(), lr=1e-4)}
    d = O.synthetic_batch(batch, NUM_CHANNEL)
    tr_img, tr_wid, label_xt, label_xt_swap = d["tr_img"], d["tr_wid"], d["label_xt"], d["label_xt_swap"]

    def generate():
        f_xss = gen.enc_image(tr_img)
        f_xs = f_xss[-1]
        f_xt, f_embed = gen.enc_text(label_xt, f_xs.shape)
        xg = gen.decode(gen.mix(f_xss, f_embed), f_xss, f_embed, f_xt)
        f_xt_s, f_embed_s = gen.enc_text(label_xt_swap, f_xs.shape)
        xg_swap = gen.decode(gen.mix(f_xss, f_embed_s), f_xss, f_embed_s, f_xt_s)
        return xg, xg_swap

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt["cla"].zero_grad()                                                      # network_tro.py:50-55
        cla(tr_img[:, 0:1].requires_grad_(), tr_wid).backward()
        opt["cla"].step()
        opt["dis"].zero_grad()                                                      # network_tro.py:105-130
        s1, s2 = tr_img[:, 0:1].requires_grad_(), tr_img[:, 1:2].requires_grad_()
        l_real = (dis.calc_dis_real_loss(s1) + dis.calc_dis_real_loss(s2)) / 2.
        l_real.backward(retain_graph=True)
        with torch.no_grad():
            xg, xg_swap = generate()
        ((dis.calc_dis_fake_loss(xg) + dis.calc_dis_fake_loss(xg_swap)) / 2.).backward()
        opt["dis"].step()
        opt["gen"].zero_grad()                                                      # network_tro.py:57-103 without l_rec
        xg, xg_swap = generate()
        l_dis = (dis.calc_gen_loss(xg) + dis.calc_gen_loss(xg_swap)) / 2.
        l_cla = (cla(xg, tr_wid) + cla(xg_swap, tr_wid)) / 2.
        (l_dis + l_cla).backward()
        opt["gen"].step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


REFERENCE_SAMPLE_BATCH = 8      # the reference driver's own BATCH_SIZE (main_run.py:58); 1/8 of the benchmarked batch


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's own implementation of the path on the host cores.  Every one of the W + K steps
    is executed and timed for real; each is a BOUNDED SAMPLE of the benchmarked step - the same iteration on the first 8 of
    the 64 samples (a full 64-sample iteration takes ~90 s on 8 cores, 25 of them would not end within minutes).
    `ms_per_step` is the measured time of such a sample step; `value` converts it to steps/s of the 64-sample workload
    (per-sample cost is what scales: every layer but the 3 BatchNorm-ed MLP / iAFF branches is per-sample work)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import ref_bootstrap as rb
    b = REFERENCE_SAMPLE_BATCH
    if rb.available():
        kind = "reference"
        times = reference_modules_steps(b, args.steps, args.warmup, threads)
        what = ("the reference's own nn.Modules (GAN_word blocks.py / modules_tro.py / vgg_tro_channel3_modi.py, unmodified, "
                "from %s) composed as network_tro.py:50-138 without the recogniser + torch.optim.Adam" % rb.REF_ROOT)
    else:
        kind = "port"
        sec = cpu_reference_steps(b, args.steps, args.warmup, threads)
        times = [sec] * args.steps
        what = "CPU oracle port (oracle/affgw_oracle.py): the reference tree is neither at /root/reference nor staged in oracle/_ref"
    sec = sum(times) / len(times)
    frac = b / BATCH_PER_GPU
    value = frac / sec
    sample = (f"{len(times)} timed iterations (after {args.warmup} warm-up), each the full cla/dis/gen iteration on {b} of the "
              f"{BATCH_PER_GPU} samples ({sec:.2f} s each, {NUM_CHANNEL} style planes, fp32, {threads} threads); steps/s = "
              f"({b}/{BATCH_PER_GPU}) / seconds per sample step; {what}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec, "ms_per_step_is": "measured wall time of one sample step",
            "sample_fraction_of_step": frac, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH_PER_GPU,
                                                             "sample_batch": b},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="affgw", choices=["affgw", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encoder", default="vgg", choices=["vgg", "resnet18", "resnet50"],
                    help="style encoder: vgg = configs[1] (default, the headline), resnet18 = configs[2]")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="gradient exchange + Adam on the main stream (no side-stream overlap)")
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "bf16x3"],
                    help="16-bit tensor-core mode: f16 (default; fp16 operand planes) or the same kernels on bf16 planes")
    ap.add_argument("--no-rec-extra", action="store_true", help="skip the extra measurement of the full iteration with the recogniser")
    ap.add_argument("--quick", action="store_true", help="timed steps only (no e2e / generation / CPU legs): the command ncu profiles")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "affgw" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist

    import affganwriting_b200 as A
    from affganwriting_b200 import _lib, network_tro, ops
    from affganwriting_b200 import load_data as LD
    from affganwriting_b200.trainer import Trainer

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().affgw_device_ok(), _lib.last_error()

    A.set_precision(args.precision)
    torch.manual_seed(0)
    trainer = Trainer(num_writers=500, device=dev, encoder=None if args.encoder == "vgg" else args.encoder,
                      cuda_graph=not args.no_graph, overlap_exchange=not args.no_overlap)
    B = args.batch
    host = synthetic_batch(B, NUM_CHANNEL, seed=1234 + rank)
    host = tuple(t.pin_memory() if torch.is_tensor(t) else t for t in host)
    resident = LD.batch_to_device(host, dev)              # uint8 images normalised on the GPU (affgw_u8_to_image)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms.item())

    def step_resident():
        trainer.train_step(resident)

    prefetch = LD.DevicePrefetcher(dev)                   # copy stream + two device slots

    def step_e2e():
        # every step: take the batch staged during the previous step, queue the H2D copy (uint8 canvases + labels from
        # pinned memory) and the GPU normalisation of the next one on the copy stream, run the step, read the losses back
        dev_batch = prefetch.get()
        prefetch.stage(host)
        losses = trainer.train_step(dev_batch)
        prefetch.release()
        return float(losses["gen"].item()) + float(losses["dis"].item()) + float(losses["cla"].item())

    if not args.no_graph:                                # eager iterations + CUDA-graph capture, before the warm-up steps
        for _ in range(Trainer.GRAPH_WARMUP + 1):
            step_resident()
    for _ in range(args.warmup):
        step_resident()
    A.check_device_errors()

    # ---- timed region (inputs resident in HBM)
    with ClockSampler(local_rank) as clocks:
        n0 = A.launch_count()
        ms_total = timed(step_resident, args.steps)
        launches = A.launch_count() - n0
    if trainer.graph_launches:                           # replayed launches are not seen by the library's launch counter
        launches += trainer.graph_launches * args.steps
    ms_step = ms_total / args.steps
    value = world / (ms_step / 1e3)                       # whole-job steps/s: every rank completes one step per step time

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> losses read back
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms_step, "gpu_launches": launches}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- instrumented pass: per-launch CUDA events (on the launching stream) around every convolution kernel
    # (single stream here: with the weight-gradient GEMMs on their side stream a kernel's events would also cover whatever ran
    # beside it - each kernel is timed ALONE for its roofline figure; the step numbers above come from the two-stream replay)
    ws_flag, trainer.wgrad_stream = trainer.wgrad_stream, False
    trainer.train_step_eager(resident)                   # eager: events cannot sit inside a graph replay
    ops.start_kernel_timing()
    timed(lambda: trainer.train_step_eager(resident), 2)
    streams = ops.stop_stream_timing()
    kern = ops.stop_kernel_timing(by_kernel=True)
    trainer.wgrad_stream = ws_flag
    instr_steps = 2

    prefetch.stage(host)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = world / (ms_e2e / 1e3)

    # ---- generation throughput (secondary number of BASELINE.json's metric): style encode + decode per image
    trainer.join()                                       # the last generator step may still be on the side stream
    gen = trainer.model.gen
    from affganwriting_b200.inference import GraphedGenerator
    gen.eval()                                           # the reference generates under model.eval() (tt.test_single_writer.4_scenarios.py:146)
    gen_fn = gen if args.no_graph else GraphedGenerator(gen)
    with torch.no_grad():
        for _ in range(4):
            gen_fn(resident[3], resident[7])
        ms_gen = timed(lambda: gen_fn(resident[3], resident[7]), 10) / 10
    # opt-in: one tensor-core pass in VGG convolutions 9-16 and the decoder ResBlocks (image 6e-3 instead of 1.5e-3 of the fp32
    # reference's, bar 2e-2: tests/test_gpu_parity_c50.py::test_c50_b8_relaxed_generation_image)
    gen_relaxed = None
    if not args.no_graph and args.encoder == "vgg" and args.precision == "f16":
        fast_fn = GraphedGenerator(gen, relaxed=True)
        with torch.no_grad():
            for _ in range(4):
                fast_fn(resident[3], resident[7])
            ms_fast = timed(lambda: fast_fn(resident[3], resident[7]), 10) / 10
        gen_relaxed = {"images_per_sec": world * B / (ms_fast / 1e3), "ms_per_batch": ms_fast,
                       "note": "inference.GraphedGenerator(gen, relaxed=True): generated image within 6e-3 of the fp32 reference "
                               "(three passes: 1.5e-3; bar 2e-2)"}
        del fast_fn
    gen.train()
    gen_img_s = world * B / (ms_gen / 1e3)

    # ---- the reference driver's COMPLETE iteration (main_run.py:146-167: rec_update -> cla_update -> dis_update -> gen_update
    # with the w_rec * l_rec term) with the native recogniser attached; reported beside the headline, whose workload
    # (BASELINE.json configs[1]) names the three convolutional models only.  The recogniser's beam search selects hypotheses on
    # the host every decoding step, so rec_update / gen_update are issued eagerly here (cla / dis still replay).
    full_iter = None
    if not args.no_rec_extra:
        try:
            del gen_fn
            t2 = Trainer(num_writers=500, device=dev, encoder=None if args.encoder == "vgg" else args.encoder,
                         cuda_graph=not args.no_graph, rec=True)
            for _ in range(Trainer.GRAPH_WARMUP + 2):
                t2.train_step(resident)
            n0 = A.launch_count()
            # host-bound (thousands of eager launches and Python objects per iteration): single iterations jitter by +-50 %
            # with the interpreter's garbage collector, so five are timed one by one and the median is reported
            each = sorted(timed(lambda: t2.train_step(resident), 1) for _ in range(5))
            ms_full = each[2]
            full_iter = {"ms_per_step": ms_full, "steps_per_sec": world / (ms_full / 1e3), "ms_min_max": [each[0], each[-1]],
                         "eager_launches_per_step": (A.launch_count() - n0) / 5, "recogniser": "affganwriting_b200.recognizer.RecModel",
                         "note": "rec_update + cla_update + dis_update + gen_update(l_dis + l_cla + l_rec), batch %d per GPU; "
                                 "median of 5 iterations timed one by one" % B}
            del t2
        except Exception as e:           # never fatal for the headline
            full_iter = {"error": repr(e)[:300]}

    # ---- the same iteration with ONE generator forward: dis_update and gen_update of the reference each run the generator on the
    # same batch with the same weights; Trainer(share_generator_forward=True) keeps the first one's autograd graph.  NOT the
    # headline: `value` times the reference's composition.
    shared_fwd = None
    if not args.no_rec_extra and args.encoder == "vgg":
        try:
            trainer.join()
            torch.cuda.synchronize()
            t3 = Trainer(num_writers=500, device=dev, cuda_graph=not args.no_graph, overlap_exchange=not args.no_overlap,
                         share_generator_forward=True)
            for _ in range(Trainer.GRAPH_WARMUP + 2):
                t3.train_step(resident)
            ms_sh = timed(lambda: t3.train_step(resident), 5) / 5
            t3.join()
            shared_fwd = {"ms_per_step": ms_sh, "steps_per_sec": world / (ms_sh / 1e3),
                          "note": "Trainer(share_generator_forward=True): same losses / gradients / BatchNorm buffers "
                                  "(tests/test_gpu_models.py::test_shared_generator_forward_matches_the_reference_composition), "
                                  "one generator forward per iteration instead of the reference's two"}
            del t3
            torch.cuda.empty_cache()
        except Exception as e:
            shared_fwd = {"error": repr(e)[:300]}

    # ---- BASELINE.json configs[4]: line-level generator, 64 x 1024 lines (T = 256 spaced characters), batch 32 per GPU
    line_gen = None
    try:
        from affganwriting_b200.linegen import SpacedGenerator
        lg = SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).to(dev).eval()
        gl = torch.Generator(device=dev).manual_seed(5)
        idx = torch.randint(0, 80, (256, 32), device=dev, generator=gl)
        content = torch.zeros(256, 32, 80, device=dev).scatter_(2, idx.unsqueeze(2), 1.0)
        style = torch.randn(32, 128, device=dev, generator=gl)
        from affganwriting_b200.inference import GraphedForward
        lg_fn = lg if args.no_graph else GraphedForward(lg)
        with torch.no_grad():
            for _ in range(4):
                lg_fn(content, style)
            ms_line = timed(lambda: lg_fn(content, style), 10) / 10
        line_gen = {"images_per_sec": world * 32 / (ms_line / 1e3), "ms_per_batch": ms_line, "batch_per_gpu": 32,
                    "image": "64x1024", "note": "SpacedGenerator forward (line_generation/model/pure_gen.py:42-50), " +
                    ("eager launches" if args.no_graph else "CUDA-graph replay (inference.GraphedForward)") + ", noise drawn on the device"}
        del lg
    except Exception as e:
        line_gen = {"error": repr(e)[:300]}

    # ---- BASELINE.json configs[3]: DINOv2 ViT-L/14 style encoder + AdaIN decoder, generation at batch 256 per GPU
    dino_gen = None
    try:
        from affganwriting_b200 import modules_tro as M
        torch.manual_seed(1)
        dg = M.GenModel_FC(12, encoder="dino").to(dev).eval()
        big = LD.batch_to_device(synthetic_batch(256, NUM_CHANNEL, seed=99 + rank), dev)
        dg_fn = dg if args.no_graph else GraphedGenerator(dg)
        with torch.no_grad():
            for _ in range(4):
                dg_fn(big[3], big[7])
            ms_dino = timed(lambda: dg_fn(big[3], big[7]), 3) / 3
        dino_gen = {"images_per_sec": world * 256 / (ms_dino / 1e3), "ms_per_batch": ms_dino, "batch_per_gpu": 256,
                    "note": "GenModel_FC(encoder=ImageEncoderDINOv2 vitl14, taps [4, 8, 16, 23]) forward under eval(), random weights, " +
                    ("eager launches" if args.no_graph else "CUDA-graph replay")}
        del dg, big
    except Exception as e:
        dino_gen = {"error": repr(e)[:300]}

    # ---- BASELINE.json configs[2]: the same training step with the ResNet-18 style encoder (torchvision wrapper wiring,
    # modules_tro2.py:447-516), so that the driver's N = 1, 2, 4, 8 runs of this file also carry that configuration.  Last of
    # the extras and never fatal.  dis_update's generator forward keeps three passes everywhere here: the relaxed forward
    # (ops.relaxed_forward) was validated against the oracle with the VGG encoder only.
    r18_step = None
    if not args.no_rec_extra and args.encoder == "vgg":
        relaxed_was = network_tro._RELAXED_DIS_FWD
        try:
            trainer.join()
            torch.cuda.synchronize()
            network_tro._RELAXED_DIS_FWD = False
            t4 = Trainer(num_writers=500, device=dev, encoder="resnet18", cuda_graph=not args.no_graph,
                         overlap_exchange=not args.no_overlap)
            for _ in range(Trainer.GRAPH_WARMUP + 2):
                t4.train_step(resident)
            ms_r18 = timed(lambda: t4.train_step(resident), 5) / 5
            t4.join()
            r18_step = {"ms_per_step": ms_r18, "steps_per_sec": world / (ms_r18 / 1e3), "n_gpus": world, "batch_per_gpu": B,
                        "note": "configs[2]: full GAN training step (cla_update -> dis_update -> gen_update + Adam, no recogniser) with "
                                "GenModel_FC(encoder='resnet18'), same batch / precision mode / stream options as the headline, "
                                "5 timed CUDA-graph iterations, max over ranks; relaxed dis_update forward off"}
            del t4
            torch.cuda.empty_cache()
        except Exception as e:
            r18_step = {"error": repr(e)[:300]}
        finally:
            network_tro._RELAXED_DIS_FWD = relaxed_was

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    launched_tflop = sum(d["flops"] for d in kern.values()) / instr_steps / 1e12
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PF sustained"
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    kernels = {}

    def mma_passes(name):                                # conv_*_kernel<BN, NPASS, ...>: tensor-core MMAs issued per product
        try:
            return int(name.split("<")[1].split(",")[1].strip(" >"))
        except Exception:
            return 1
    for name, d in kern.items():
        tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        kernels[name] = {"launches_per_step": d["launches"] / instr_steps, "ms_per_step": d["ms"] / instr_steps,
                         "share_of_step": d["ms"] / instr_steps / ms_step, "bound": "tensor", "tflops": tf,
                         "frac_of_peak": tf / tf_peak, "mma_passes": mma_passes(name)}
    # streaming kernels (normalisation, operand split, pools, optimiser): algorithmic bytes (each tensor read once + written
    # once) over their CUDA-event time, against the measured HBM copy bandwidth
    for name, d in streams.items():
        gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        kernels[name] = {"launches_per_step": d["launches"] / instr_steps, "ms_per_step": d["ms"] / instr_steps,
                         "share_of_step": d["ms"] / instr_steps / ms_step, "bound": "hbm", "gbs": gbs,
                         "frac_of_peak": gbs / hbm_peak, "algorithmic_bytes_per_step": d["bytes"] / instr_steps}
    tensor_kernels = {k: v for k, v in kernels.items() if v["bound"] == "tensor"}
    dominant = max(tensor_kernels, key=lambda k: tensor_kernels[k]["ms_per_step"]) if tensor_kernels else None
    roofline = None
    if dominant:
        k = kernels[dominant]
        roofline = {"kernel": dominant, "bound": "tensor", "achieved": k["tflops"], "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": k["frac_of_peak"], "traffic": None, "peak_source": peak_src,
                    "avg_launch_ms": k["ms_per_step"] / max(k["launches_per_step"], 1e-9),
                    "share_of_step": k["share_of_step"], "mma_passes": k["mma_passes"],
                    "executed_tflops": k["mma_passes"] * k["tflops"], "executed_frac": k["mma_passes"] * k["frac_of_peak"],
                    "note": "achieved = algorithmic FLOPs (direct-convolution MAC x 2) of every launch of this kernel function in "
                            "one eager step / their summed CUDA-event time on the launching stream.  Forward GEMMs issue 3 "
                            "tensor-core MMAs per product (split-bf16 operands, DESIGN.md section 3), backward GEMMs 1: "
                            "executed_* = mma_passes x the algorithmic figure = tensor-pipe work against the measured cuBLAS "
                            "peak (a 3-pass kernel's frac is capped at 1/3).  traffic: ncu DRAM bytes of this kernel on one layer"}
    hbm_ms = sum(v["ms_per_step"] for v in kernels.values() if v["bound"] == "hbm")
    hbm_bytes = sum(v["algorithmic_bytes_per_step"] for v in kernels.values() if v["bound"] == "hbm")
    hbm_summary = {"ms_per_step": hbm_ms, "share_of_step": hbm_ms / ms_step, "algorithmic_gb_per_step": hbm_bytes / 1e9,
                   "achieved_gbs": hbm_bytes / max(hbm_ms, 1e-9) / 1e6, "peak_gbs": hbm_peak,
                   "frac": hbm_bytes / max(hbm_ms, 1e-9) / 1e6 / hbm_peak}

    try:        # DRAM bytes of the dominant kernel from the committed ncu --set full capture (one layer, see profiles/README.md)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(dominant.replace(" f16", ""))
        if tr and roofline:
            roofline["traffic"] = tr["dram_bytes_per_launch"]
            roofline["traffic_note"] = ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch on " + tr["layer"] +
                                        "; algorithmic bytes of that launch: %d" % sum(tr["algorithmic_bytes_per_launch"].values()))
    except Exception:
        pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        # the reference arm of this same file in a child process with the GPUs hidden (3 timed sample steps after 1 warm-up)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                                 capture_output=True, text=True, timeout=900, cwd=ROOT).stdout
            cpu_baseline = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])["cpu_baseline"]
        except Exception as e:                      # reported, never fatal for the GPU measurement
            cpu_baseline = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)[:200]}

    h2d = batch_bytes(host)
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "f16" else "bf16",
        "dtype_detail": ("tcgen05 kind::f16 MMAs on %s operand planes with fp32 TMEM accumulation (forward: split operands, 3 MMAs per "
                         "product - 1 in the decoder's up-convolutions in mode f16 and, in the no_grad generator forward of dis_update whose image "
                         "only feeds the discriminator, in VGG convolutions 6-16 and the decoder ResBlock convolutions "
                         "(ops.relaxed_forward); backward: 1 MMA per product); activations, "
                         "statistics, gradients and parameters are stored in fp32.  fp16 planes carry 11 significant bits against "
                         "bf16's 8 at the same tensor throughput; their range is handled by exact power-of-two scales "
                         "(DESIGN.md section 3)" % ("fp16" if args.precision == "f16" else "bf16")),
        "value_note": "whole-job aggregate under weak scaling: every rank completes one batch-%d training step per step time, value = "
                      "n_gpus / step time (rank-steps/s; one optimiser step of global batch n_gpus x %d per step time); "
                      "config.samples_per_sec = value x batch_per_gpu" % (B, B),
        "data": "synthetic (seeded uint8 grey-level canvases normalised on the GPU like load_data.py:152-166; random-init weights)",
        "config": {"workload": WORKLOAD if args.encoder == "vgg" else WORKLOAD.replace(
                       "configs[1]", "configs[2] (%s style encoder)" % args.encoder),
                   "encoder": args.encoder, "precision_mode": args.precision, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "cuda_graph": bool(trainer.graph_launches), "overlap_exchange": bool(trainer.overlap_exchange),
                   "wgrad_side_stream": bool(trainer.wgrad_stream), "concurrent_cla_dis": bool(trainer.concurrent_cla_dis),
                   "early_generator_forward": bool(trainer.early_generator_forward),
                   "concurrent_gen_heads": bool(trainer.concurrent_gen_heads),
                   "side_text_encoder": bool(trainer.model.side_text_encoder),
                   "relaxed_dis_update_forward": bool(network_tro._RELAXED_DIS_FWD and args.precision == "f16"),
                   "l2": "inputs larger than L2: 177 MB of style images are re-read every step (L2 is 126 MB)",
                   "samples_per_sec": value * B},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "ms_per_step": ms_e2e,
                "wire_format": "uint8 images (a quarter of the float32 bytes) + int64 labels from pinned host memory; "
                               "affgw_u8_to_image normalises on the GPU, bit-exact with the reference loader; the copy of "
                               "batch k+1 runs on a copy stream during step k (one H2D per step inside the timed region)"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "hbm_kernels": hbm_summary,
        "kernels": kernels,
        "step_tflops": None if args.encoder != "vgg" else {
            "algorithmic_tflop_per_step": STEP_GFLOP_PER_SAMPLE * B / 1e3,
            "achieved_tflops_per_gpu": STEP_GFLOP_PER_SAMPLE * B / 1e3 / (ms_step / 1e3),
            "frac_of_peak": STEP_GFLOP_PER_SAMPLE * B / 1e3 / (ms_step / 1e3) / tf_peak,
            # the direct-convolution FLOPs of the launches that actually ran (the reference's dead weight gradients of gen_update
            # and the dead stem dgrads are skipped, trainer.skip_unused_wgrad): the figure the fraction should be read against
            "launched_tflop_per_step": launched_tflop,
            "launched_tflops_per_gpu": launched_tflop / (ms_step / 1e3),
            "launched_frac_of_peak": launched_tflop / (ms_step / 1e3) / tf_peak},
        "cpu_baseline": cpu_baseline,
        "extra": {"full_iteration_with_recogniser": full_iter, "iteration_with_shared_generator_forward": shared_fwd, "line_generator": line_gen, "dino_generation": dino_gen, "resnet18_encoder_step": r18_step, "gen_images_per_sec": gen_img_s, "gen_images_relaxed": gen_relaxed, "gen_batch_per_gpu": B, "gen_ms_per_batch": ms_gen,
                  "gen_frac_of_peak": None if args.encoder != "vgg" else 62.17e-3 * gen_img_s / world / tf_peak},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
