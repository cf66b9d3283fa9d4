#!/bin/bash
# round 2, run c: full GPU test suite with the new parity tests, default bench line with other_configs, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_c_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -3 gpurun_out/r02_c_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','warp_kcycles_max_median','env_knobs') if k in d}, d['e2e']['value'], d['roofline']['frac'], d.get('large_batch'))
for k,v in (d.get('other_configs') or {}).items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity','error')}, (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'))
"
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_c_bench_reference.json 2>&1 ) 2>&1 | grep real
cut -c1-600 gpurun_out/r02_c_bench_reference.json
