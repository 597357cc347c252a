"""Turn the artifacts of scripts/gpu_profile_pass.sh (gpurun_out/) into the committed summaries under profiles/:
   launches_r01b.csv (ncu launch list)           -> profiles/r01_launch_list_step.csv  (per-kernel shares)
   prof_shift_r01b.ncu-rep (ncu --set full)      -> profiles/r01_ncu_full_position_kernels_vgg256.csv + r01_traffic.json
   bench.json                                    -> profiles/r01_bench_n1.json
Runs in the build container (needs `ncu` for the report import only)."""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# round tag and the artifact names of scripts/gpu_profile_pass_r02.sh (round 1 used launches_r01b.csv / prof_shift_r01b / bench.json)
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
LAUNCHES = {"r01": "launches_r01b.csv"}.get(R, f"launches_{R}.csv")
REPORT = {"r01": "prof_shift_r01b.ncu-rep"}.get(R, f"prof_shift_{R}.ncu-rep")
BENCH = {"r01": "bench.json"}.get(R, f"bench_{R}.json")
MODE = {"r01": "bf16 mode (3 MMAs per product)"}.get(R, "f16 mode (forward 3 MMAs per product, backward 1)")


# algorithmic bytes of one launch on the profiled layer (two fp16 planes of 122.6 MB for the split forward, one for the single-pass
# backward GEMMs; 226.5 MB of fp32 activations / 2.4 MB of fp32 weight gradients written)
ALGO_BYTES = {
    "conv_shift_tcgen05_kernel<256, 1, 2>": {"operand plane read (dY, one fp16 plane)": 122552320, "fp32 output written": 226492416},
    "conv_wgrad_shift_kernel<128, 1>": {"operand planes read (x and dY, one fp16 plane each)": 245104640,
                                        "fp32 weight-gradient partials written": 2359296},
}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return name.split("(")[0]


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(G, LAUNCHES))) if len(r) > 8]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        a = agg[short(r[ki])]
        a[0] += float(r[vi].replace(",", "")) / 1e6          # ns -> ms
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{R}_launch_list_step.csv"), "w") as f:
        f.write(f"# ncu launch list of one eager training step (bench.py --quick --no-graph --steps 1 --warmup 3), {R} kernels, {MODE}\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -s <3 warm-up steps> -c <1 step + margin> --csv python bench.py --quick --no-graph --steps 1 --warmup 3\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py \"kernels[*].share_of_step\", not absolutes\n")
        f.write(f"# launches in window {n}, summed kernel time {total:.1f} ms\n")
        f.write("share_pct,ms,launches,kernel\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{100 * a[0] / total:.2f},{a[0]:.3f},{a[1]},\"{k}\"\n")
    top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:3]
    print("launch list:", n, "launches,", f"{total:.1f} ms;", [(k, f"{100 * a[0] / total:.1f}%") for k, a in top])


def full_capture():
    rep = os.path.join(G, REPORT)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
    idx = [hdr.index(k) for k in keep if k in hdr]
    with open(os.path.join(P, f"{R}_ncu_full_position_kernels_vgg256.csv"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:shift -c 6 python scripts/conv_microbench.py --only vgg_256_256 --reps 1   ({R} kernels)\n")
        f.write(f"# layer: VGG conv3x3 256->256 @ 32x108, batch 64, {MODE}: 260.9 algorithmic GFLOP per launch\n")
        w = csv.writer(f)
        w.writerow([f"{hdr[i]} [{units[i]}]" for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])

    def col(r, k):
        return float(r[hdr.index(k)].replace(",", ""))

    def to_bytes(r, k):
        u = units[hdr.index(k)].lower()
        return col(r, k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    out = {}
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        if "pack_weight" in name or "split" in name:
            continue
        d = out.setdefault(name, {"layer": "VGG conv3x3 256->256 @32x108, batch 64 (scripts/conv_microbench.py --only vgg_256_256)",
                                  "launches_profiled": 0, "dram": 0.0, "us": 0.0, "tensor": 0.0, "l1": 0.0})
        d["launches_profiled"] += 1
        d["dram"] += to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
        d["us"] += col(r, "gpu__time_duration.sum")
        d["tensor"] += col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        d["l1"] += col(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
    res = {}
    for name, d in out.items():
        n = d["launches_profiled"]
        res[name] = {"layer": d["layer"], "launches_profiled": n, "dram_bytes_per_launch": d["dram"] / n,
                     "algorithmic_bytes_per_launch": ALGO_BYTES.get(name, {"operand planes read": 245104640, "fp32 output written": 226492416}),
                     "us_per_launch_under_ncu": d["us"] / n, "tensor_pipe_active_pct": d["tensor"] / n,
                     "l1tex_throughput_pct": d["l1"] / n}
        print(name, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res[name].items() if k != "layer"})
    json.dump(res, open(os.path.join(P, f"{R}_traffic.json"), "w"), indent=1)


def bench():
    line = open(os.path.join(G, BENCH)).read().strip().splitlines()[-1]
    d = json.loads(line)
    json.dump(d, open(os.path.join(P, f"{R}_bench_n1.json"), "w"), indent=1)
    print("bench:", d["value"], "steps/s,", d["ms_per_step"], "ms; e2e", d["e2e"]["value"])


if __name__ == "__main__":
    launch_list()
    full_capture()
    bench()
