#!/bin/bash
# compute-sanitizer over the Hex27 kernels added last (one-dof, mdot, nodal gradient, fused mass, aliasing, empty parts)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_hex27.py -q -m gpu > gpurun_out/s4v_memcheck_hex27.txt 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/s4v_memcheck_hex27.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_hex27.py -q -m gpu -k "one_dof or fused or nodal or mdot" > gpurun_out/s4v_racecheck_hex27.txt 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/s4v_racecheck_hex27.txt
