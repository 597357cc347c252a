"""Pretty-print a bench.py JSON line: python scripts/show_bench.py gpurun_out/bench.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], '| roofline', {k: d['roofline'][k] for k in ('kernel', 'frac', 'executed_frac', 'share_of_step')})
print('hbm', d.get('hbm_kernels')); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores']); print(d['extra'], d.get('step_tflops'))
tot = {'tensor': 0, 'hbm': 0}
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    tot[v.get('bound', 'tensor')] += v['ms_per_step']
    rate = '%7.1f TF/s' % v['tflops'] if v.get('bound', 'tensor') == 'tensor' else '%7.0f GB/s' % v['gbs']
    print('%-52s %6.1f x %7.3f ms %s  frac %.3f' % (k, v['launches_per_step'], v['ms_per_step'], rate, v['frac_of_peak']))
print('totals', tot)
