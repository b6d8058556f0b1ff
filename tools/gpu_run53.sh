#!/bin/bash
python tools/debug_case.py tests/golden/stall_n144_m117_rho09.npz 'PLS_K2_IMPL=v3,PLS_K3_CHAIN=0' 2>&1 | tail -30
