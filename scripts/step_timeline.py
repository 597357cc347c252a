"""GPU timeline of replayed training steps (CUPTI through torch.profiler): every kernel / memset / memcpy of `--steps` CUDA-graph
replays with stream, start and duration -> gpurun_out/timeline_step.csv, plus a summary (busy time per stream, time with no
kernel running on any stream, concurrency histogram).  Run under gpurun; analyse the CSV anywhere."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/timeline_step.csv")
    args = ap.parse_args()
    import affganwriting_b200 as A
    import bench
    from affganwriting_b200 import load_data as LD
    from affganwriting_b200.trainer import Trainer
    from torch.profiler import ProfilerActivity, profile

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    A.set_precision("f16")
    torch.manual_seed(0)
    t = Trainer(num_writers=500, device=dev, cuda_graph=True, overlap_exchange=True)
    host = bench.synthetic_batch(bench.BATCH_PER_GPU, bench.NUM_CHANNEL, seed=1234)
    batch = LD.batch_to_device(host, dev)
    for _ in range(Trainer.GRAPH_WARMUP + 4):
        t.train_step(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            t.train_step(batch)
        t.join()
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    # stream ids: kineto events carry them in the chrome trace only; export that too (compact) and parse it
    trace = args.out.replace(".csv", "_trace.json")
    prof.export_chrome_trace(trace)
    import json
    ev = [e for e in json.load(open(trace))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    with open(args.out, "w") as f:
        f.write("start_us,dur_us,stream,cat,name\n")
        for e in ev:
            f.write("%.3f,%.3f,%s,%s,%s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream", e.get("tid")), e["cat"],
                                            e["name"].replace(",", ";")[:90]))
    os.remove(trace)
    print("events", len(ev), "span ms", (ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3)


if __name__ == "__main__":
    main()
