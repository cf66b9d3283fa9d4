#!/bin/bash
python profiles/lane_sweep.py 4096 65536
python profiles/phase_profile.py config2
