#!/bin/bash
# The binning / engine tests against a -DGSR_DEBUG_CHECKS build (index checks that trap), then the product build again.
#   gpurun -- 'bash tools/debug_checks.sh'
O=gpurun_out
GSR_EXTRA_NVCC_FLAGS="-DGSR_DEBUG_CHECKS" python gs-slam-analytica_jacobian_b200/build.py --force > $O/debug_build.log 2>&1; echo "debug build rc=$?"
python -m pytest tests/test_binning_gpu.py tests/test_engine_gpu.py "tests/test_config_scale_gpu.py::test_config_scale_parity[C2]" "tests/test_config_scale_gpu.py::test_config_scale_parity[C3]" -m gpu -x -q > $O/pytest_debug_checks.log 2>&1; echo "pytest (debug checks) rc=$?"; tail -2 $O/pytest_debug_checks.log
python gs-slam-analytica_jacobian_b200/build.py --force > /dev/null 2>&1
