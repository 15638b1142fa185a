#!/bin/bash
# The frame parity tests against a build with -DVX_DEBUG_CHECKS (index checks inside the frame kernels; stands in for
# compute-sanitizer memcheck, which is closed on the GPU pool).  usage: bash tools/debug_checks.sh <tag>
TAG=${1:-x}
bash tools/build_variant.sh dbg "-DVX_DEBUG_CHECKS" > /dev/null 2>&1 || true
VX_B200_LIB=$PWD/variants/libvx_dbg.so python -m pytest tests/test_frame_gpu.py tests/test_golden.py tests/test_world_gpu.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/debug_checks_${TAG}.log 2>&1
echo "debug-checks build: rc=$? $(tail -1 gpurun_out/debug_checks_${TAG}.log)"
VX_B200_LIB=$PWD/variants/libvx_dbg.so python tools/sanitize_target.py >> gpurun_out/debug_checks_${TAG}.log 2>&1; echo "target rc=$?"
