#!/bin/bash
python bench.py --workload config5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config5', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
python bench.py --workload config3 --no-cpu-baseline --cl-steps 50 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config3', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
python bench.py --workload config4 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config4', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
python -m pytest tests -m gpu -x -q -k "double_pendulum" 2>&1 | tail -2
