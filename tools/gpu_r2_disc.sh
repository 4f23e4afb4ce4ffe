#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "discretize or golden" 2>&1 | tail -2
echo "== default"; timeout 120 python tools/disc_probe.py all
echo "== no MMA"; SCHEMANET_DISC_DEBUG=2 timeout 120 python tools/disc_probe.py
