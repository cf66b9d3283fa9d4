#!/bin/bash
# launch list of the default bench command with the final build (numbers printed under ncu are never bench values)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_y_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r02_y_ncu_bench.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, io, collections
rows=[l for l in open('gpurun_out/r02_y_launches_bench.csv') if not l.startswith('==')]
agg=collections.Counter(); cnt=collections.Counter()
for r in csv.DictReader(io.StringIO(''.join(rows))):
    n=r['Kernel Name'].split('(')[0][:60]; agg[n]+=float(r['Metric Value']); cnt[n]+=1
for n,v in agg.most_common(8): print(f"{n:62s} launches {cnt[n]:4d}  total {v/1e3:10.1f} us  mean {v/cnt[n]/1e3:9.1f} us")
PY
