#!/bin/bash
# compute-sanitizer evidence: racecheck + memcheck on the kernels with hand-rolled synchronisation --
# lock-free union-find (cc.cu), the ticket-counter BatchNorm finalize inside conv_tc, fp32 red.global.add weight gradients,
# the mbarrier / tcgen05 pipelines -- through small parity tests (the sanitizer serialises everything; keep the set small).
set -u
mkdir -p gpurun_out
T="tests/test_gpu_step.py::test_largest_cc_kernel_matches_host_oracle tests/test_gpu_ops.py::test_tensor_core_path_is_taken_and_accurate_to_tf32 tests/test_gpu_ops.py::test_perturbation_generator tests/test_gpu_ops.py::test_feature_dropout_kernel_matches_reference_fixture"
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -q -x -p no:cacheprovider > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
