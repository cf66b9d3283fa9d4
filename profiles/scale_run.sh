set -x
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r01_k_scale_n$n.json 2> gpurun_out/r01_k_scale_n$n.err
  tail -c 300 gpurun_out/r01_k_scale_n$n.err
done
python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r01_k_scale_n1.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload config4 --steps 3 --warmup 1 > gpurun_out/r01_k_config4_n8.json 2> gpurun_out/r01_k_config4_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload config5 --steps 3 --warmup 1 > gpurun_out/r01_k_config5_n8.json 2> gpurun_out/r01_k_config5_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --allgather --steps 50 --warmup 5 > gpurun_out/r01_k_scale_n8_allgather.json 2> gpurun_out/r01_k_n8ag.err
for f in gpurun_out/r01_k_scale_n*.json gpurun_out/r01_k_config*_n8.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], l["n_gpus"], round(l["value"]), l["ms_per_step"], l.get("e2e",{}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
