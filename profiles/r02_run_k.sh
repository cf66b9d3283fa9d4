#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "first_qp_blocks or teacher_forced or long_horizon_512 or batches_match or swingup_batch or terminal_weight or variants_against or maximum_sizes or non_finite or latency_mapping or per_instance" 2>&1 | tail -8
python profiles/shard_times.py 8 --why | grep -v instance | cut -c1-600 | grep -E "shard\": [056]|max_over"
python profiles/phase_profile.py config2 config4
python bench.py --workload config4 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config4', d['value'], d['ms_per_step'], d['roofline']['frac'])"
