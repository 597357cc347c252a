run() { tag=$1; shift; env "$@" python bench.py --no-rec-extra --no-cpu-baseline > gpurun_out/ab_$tag.log 2>/dev/null; python - <<P
import json
try:
    l=[x for x in open("gpurun_out/ab_$tag.log") if x.startswith("{")][-1]; d=json.loads(l); print("$tag", round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
except Exception as e: print("$tag", "ERR", e)
P
}
run early_p0 A=1
run late_p0 AFFGW_WGRAD_FORK=late
run early_mainhi AFFGW_MAIN_PRIO=-1
run late_mainhi AFFGW_WGRAD_FORK=late AFFGW_MAIN_PRIO=-1
