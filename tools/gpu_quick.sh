# usage: bash tools/gpu_quick.sh  -- gpu tests + quick device timing + work-item trace, logs under gpurun_out/
python -m pytest tests -m gpu -x -q -s > gpurun_out/quick_pytest.log 2>&1; echo pytest rc=$?
python tools/variant_bench.py > gpurun_out/quick_bench.log 2>&1; tail -1 gpurun_out/quick_bench.log
python tools/frame_trace.py > gpurun_out/quick_trace.log 2>&1; echo trace rc=$?
