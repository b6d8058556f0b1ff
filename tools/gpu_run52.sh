#!/bin/bash
python tools/debug_case.py tests/golden/stall_n144_m117_rho09.npz 'PLS_K2_IMPL=v3' 'PLS_K2_IMPL=v2' 'PLS_K2_IMPL=v1' 'PLS_K2_IMPL=v4,PLS_K4_GRID=4' 'PLS_K2_IMPL=v4,PLS_K4_GRID=1,PLS_K4_L=2' 2>&1 | tail -8
