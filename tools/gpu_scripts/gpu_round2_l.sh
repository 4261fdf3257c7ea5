#!/bin/bash
# face/element BC path + config-4 tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_face_elem.py tests/test_gpu_parity.py -q -m gpu -k "face or config4" > gpurun_out/s3c_tests.txt 2>&1
tail -40 gpurun_out/s3c_tests.txt
