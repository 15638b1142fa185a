# usage: bash tools/gpu_session.sh <tag> [ncu]   -- gpu tests + bench (+ ncu launch list and full capture, each only after
# the program itself has exited 0 without ncu)
TAG=${1:-x}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench rc=$?
tail -3 gpurun_out/bench_${TAG}.err
for l in 1 3 8; do
timeout 900 python bench.py --steps 200 --warmup 10 --lanes $l > gpurun_out/bench_${TAG}_lanes$l.json 2> gpurun_out/bench_${TAG}_lanes$l.err; echo bench lanes $l rc=$?
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo ref rc=$?
timeout 300 python tools/ncu_target.py 3 > gpurun_out/ncu_plain_${TAG}.log 2>&1; echo target rc=$?
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/ncu_target.py 3 > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frame_raster_kernel|frame_setup_kernel|mesh_chunks_kernel|frame_cull" -s 1 -c 7 -o gpurun_out/prof_${TAG} -f python tools/ncu_target.py 3 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo ncu rc=$?
python tools/frame_trace.py > gpurun_out/trace_${TAG}_1280x720_vd12.txt 2>&1
python tools/frame_trace.py 3840 2160 32 > gpurun_out/trace_${TAG}_3840x2160_vd32.txt 2>&1
fi
