#!/bin/bash
python -m pytest tests -m gpu -x -q -k "single_solve or batches_match or swingup_batch or ragged or closed_loop or full_size or latency_mapping or first_qp" 2>&1 | tail -5
python profiles/phase_profile.py config2 config2-large config4
python profiles/lane_sweep.py 4096 65536
