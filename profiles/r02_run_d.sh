#!/bin/bash
mkdir -p gpurun_out
python profiles/occupancy_probe.py > gpurun_out/r02_d_occupancy_probe.jsonl 2>&1
cat gpurun_out/r02_d_occupancy_probe.jsonl
python -m pytest tests -m gpu -x -q -k "long_horizon_512 or two_live or rejected or native_pypendulum" 2>&1 | tail -5
