# usage: bash tools/gpu_session.sh <tag> [ncu]   -- gpu tests + bench (+ ncu launch list and full capture)
TAG=${1:-x}
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench rc=$?
tail -3 gpurun_out/bench_${TAG}.err
python tools/show_bench.py gpurun_out/bench_${TAG}.json 2>&1 | tail -12
timeout 300 python tools/ncu_target.py 3 > gpurun_out/ncu_plain_${TAG}.log 2>&1; tail -4 gpurun_out/ncu_plain_${TAG}.log | cut -c1-300
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/ncu_target.py 3 > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frame_raster_kernel|frame_setup_kernel|mesh_chunks_kernel|frame_cull" -c 4 -o gpurun_out/prof_${TAG} -f python tools/ncu_target.py 1 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo ncu rc=$?
fi
