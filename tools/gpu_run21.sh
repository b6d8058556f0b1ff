#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/k2_sweep.py cfg2 'PLS_K3_PROW16=0' '' 'PLS_K3_PROW16=104' 'PLS_K3_PROW16=128' 'PLS_K3_PROW16=0,PLS_K2_PHASES=1' 'PLS_K2_PHASES=1' 'PLS_K3_PROW16=104,PLS_K2_PHASES=1' > gpurun_out/k2_sweep21_cfg2.jsonl 2> gpurun_out/sweep21.err; echo "sweep rc=$?"
cut -c1-130 gpurun_out/k2_sweep21_cfg2.jsonl
