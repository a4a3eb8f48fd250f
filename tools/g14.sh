#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 60 python tools/config_probe.py clustered 200000 || echo "TIMEOUT 200k"
timeout 120 python tools/config_probe.py clustered 2000000 || echo "TIMEOUT 2m"
timeout 200 python tools/config_probe.py clustered 10000000 || echo "TIMEOUT 10m"
