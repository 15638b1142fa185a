import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("extra", {})
print("fps", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e fps", d["e2e"]["value"] and round(d["e2e"]["value"], 1),
      "warm-L2 fps", round(e.get("frames_per_sec_warm_l2", 0), 1))
print("kernel_ms", {k: round(v, 4) for k, v in e.get("kernel_ms", {}).items()})
print("roofline", d["roofline"]["kernel"], "GB/s", round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 4), "share", d["roofline"]["kernel_share_of_step"])
print("cpu fps", round(d["cpu_baseline"]["value"], 1), "cores", d["cpu_baseline"]["cores"], "clocks", d["clocks"])
print("mesh: world chunks/s", round(e.get("chunks_meshed_per_sec", 0)), "ms", round(e.get("remesh_world_ms", 0), 4), "frac", round(e.get("remesh_hbm_frac", 0), 4),
      "| large batch chunks/s", round(e.get("chunks_meshed_per_sec_large_batch", 0)), "frac", round(e.get("large_batch_hbm_frac", 0), 4), "| cpu", round(e.get("cpu_chunks_meshed_per_sec_1_thread", 0)))
print("frame", e.get("frame_stats"), "launches", d["gpu_launches"], "lanes", e.get("lanes"), "frame alone ms", e.get("frame_alone_ms"),
      "host submit us", e.get("host_submit_us_per_frame"))
if d.get("n_gpus", 1) > 1:
    print("stripes", e.get("stripes"), "guards", {k: e.get(k) for k in ("composite_bit_identical", "sharded_batch_bit_identical", "e2e_frames_identical")},
          "stripe ms without hand-off (max over ranks)", e.get("stripe_ms_without_handoff_max_over_ranks"))
c5 = e.get("cfg5_3840x2160_vd32") or {}
print("cfg5 3840x2160 vd32 frame ms", c5.get("frame_ms"), "one GPU", c5.get("frame_ms_one_gpu_whole_frame"))
