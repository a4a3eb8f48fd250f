#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "laguerre or golden or reference_live or otm_newton_vs or clustered or fused_walk" 2>&1 | tail -5
timeout 500 python tools/config_probe.py otm 10000000 5 1e-3
timeout 300 python tools/config_probe.py laguerre 10000000 5 0.25
