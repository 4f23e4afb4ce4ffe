#!/bin/bash
SCHEMANET_GEMM_TRACE=1 timeout 400 python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-graph 2>&1 >/dev/null | grep "gemm trace" | tail -24
