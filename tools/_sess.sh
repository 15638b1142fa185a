timeout 600 python -m pytest tests/test_frame_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for x in 200 100 140 300; do
  echo "== target x$x"; VX_ITEM_TARGET_X100=$x timeout 300 python tools/overlap_probe.py 2>&1 | grep "group mode"
done
for x in 200 100 50; do
  echo "== 4K target x$x"; VX_ITEM_TARGET_X100=$x timeout 300 python tools/overlap_probe.py 3840 2160 32 2>&1 | grep "group mode" | head -3
done
