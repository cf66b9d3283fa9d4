#!/bin/bash
# round 2 final-state measurements for the documentation: full test suite, default bench (N=1) with other_configs,
# reference arm, batch sweep, shard times, phase profile, latency
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_x7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_x7_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_x7_bench.json 2> gpurun_out/r02_x7_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_x7_bench_reference.json 2>/dev/null; echo "ref rc=$?"
python profiles/lane_sweep.py 1 8 64 4096 4736 8192 16384 65536 262144 > gpurun_out/r02_x7_batch_sweep.jsonl
python profiles/shard_times.py 8 --why | grep -v '"instance"' > gpurun_out/r02_x7_shard_times.jsonl
python profiles/phase_profile.py config2 config2-large config4 > gpurun_out/r02_x7_phase_profile.jsonl
python -c "
import json
d=json.loads(open('gpurun_out/r02_x7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','warp_kcycles_max_median') if k in d}, d['e2e']['value'], d['e2e']['synchronous']['value'], d['roofline']['frac'], d.get('large_batch'), d.get('large_batch_262144'), d['latency']['p50_ms'], d['latency']['min_latency_flag']['p50_ms'], d['cpu_baseline'])
for k,v in (d.get('other_configs') or {}).items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity','error')}, (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'))
"
cat gpurun_out/r02_x7_batch_sweep.jsonl; tail -1 gpurun_out/r02_x7_shard_times.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r02_x7_launches_65536.csv \
    python profiles/prof_target.py 65536 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_x7_launches_65536.csv | tail -3 | cut -d, -f5,15
