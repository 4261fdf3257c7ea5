#!/bin/bash
# compute-sanitizer memcheck over one small case of every kernel family
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_face_elem.py -q -m gpu -x \
  -k "golden or test_box_kernel_tiles_and_chunks or test_cont_box_kernel_options_and_chunks or test_scalar_box_kernel_options_and_chunks or test_row_owner or test_row_thread_unstructured or test_fused_lumped_mass_list or test_config4_lists_tiled or test_tpetra or test_general_kernel_is_deterministic or test_node_mass or test_compute_mdot or test_assemble_nodal_grad or test_dirichlet or face_elem_algorithm_vs_oracle" \
  > gpurun_out/s4n_memcheck.txt 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/s4n_memcheck.txt
grep -c "Invalid\|out of bounds" gpurun_out/s4n_memcheck.txt
