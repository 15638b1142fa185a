import json, sys
d = json.load(open(sys.argv[1]))
e = d.get("extra", {})
print("fps", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e fps", d["e2e"]["value"] and round(d["e2e"]["value"], 1),
      "warm-L2 fps", round(e.get("frames_per_sec_warm_l2", 0), 1))
print("kernel_ms", {k: round(v, 4) for k, v in e.get("kernel_ms", {}).items()})
print("roofline", d["roofline"]["kernel"], "GB/s", round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 4), "share", d["roofline"]["kernel_share_of_step"])
print("cpu fps", round(d["cpu_baseline"]["value"], 1), "cores", d["cpu_baseline"]["cores"], "clocks", d["clocks"])
print("mesh: world chunks/s", round(e.get("chunks_meshed_per_sec", 0)), "ms", round(e.get("remesh_world_ms", 0), 4), "frac", round(e.get("remesh_hbm_frac", 0), 4),
      "| large batch chunks/s", round(e.get("chunks_meshed_per_sec_large_batch", 0)), "frac", round(e.get("large_batch_hbm_frac", 0), 4), "| cpu", round(e.get("cpu_chunks_meshed_per_sec_1_thread", 0)))
print("config", {k: d["config"][k] for k in ("visible_meshes", "visible_quads", "triangles")}, "launches", d["gpu_launches"])
