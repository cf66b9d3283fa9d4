#!/bin/bash
# final build of round 2 on one 8 x B200 box: N = 1, 2, 4, 8 of the default bench (with other_configs at N = 8)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r02_z3_scale_n1.json 2> gpurun_out/r02_z3_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 $( [ $n != 8 ] && echo --no-other-configs ) > gpurun_out/r02_z3_scale_n$n.json 2> gpurun_out/r02_z3_scale_n$n.err
  tail -c 300 gpurun_out/r02_z3_scale_n$n.err
done
for f in gpurun_out/r02_z3_scale_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], l["n_gpus"], round(l["value"]), l["ms_per_step"], l.get("e2e",{}).get("value"), {k: round(v["value"]) for k, v in (l.get("other_configs") or {}).items() if "value" in v})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
