#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_l_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_l_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','warp_kcycles_max_median') if k in d}, d['e2e']['value'], d['roofline']['frac'], d.get('large_batch'), d['latency']['p50_ms'], d['latency']['min_latency_flag']['p50_ms'])
for k,v in (d.get('other_configs') or {}).items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity','error')}, (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'))
"
