#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_cpp.py -q -m gpu 2>&1 | tail -25
