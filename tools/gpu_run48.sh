#!/bin/bash
SWEEP_COUNT_LOG2=20 timeout 600 python tools/k2_sweep.py m512k24 '' 2>/dev/null | cut -c1-120
timeout 600 python tools/k2_sweep.py cfg2 '' 2>/dev/null | cut -c1-120
