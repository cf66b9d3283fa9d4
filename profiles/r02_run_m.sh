#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pipelined or reset_and_masks or device_pointer" 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-other-configs > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_m_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_m_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
"
