#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "otm" 2>&1 | tail -4
timeout 500 python tools/config_probe.py otm 10000000 5 1e-3
