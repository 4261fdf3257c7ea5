#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_face_elem.py -q 2>&1 | tail -6
timeout 900 python -m pytest tests/test_host_cpp.py -q 2>&1 | tail -3
