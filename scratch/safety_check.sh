#!/bin/bash
# the checking build (-DMR_CHECK: device-side bounds assertions, scratch/variants/check.so) over
# the sanitizer cases and the parity tests
export MINRMS_B200_LIB=$PWD/scratch/variants/check.so
python scratch/sanitize_case.py 2>&1 | tail -8
python -m pytest tests/test_gpu_parity.py tests/test_gpu_repeatability.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
