"""DRAM bytes per launch of every kernel in an `ncu --set full` report -> JSON (bench.py reads the raster kernel's entry for
roofline.traffic):  python tools/ncu_traffic.py rep out.json "source note" """
import csv, io, json, subprocess, sys
rep, out_path = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
out = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[hdr.index("Kernel Name")]
    short = next((k for k in ("mesh_chunks_kernel", "frame_cull_kernel", "frame_setup_kernel", "frame_raster_kernel") if k in name), name)
    def val(metric):
        i = hdr.index(metric)
        v = float(r[i].replace(",", ""))
        unit = rows[1][i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                  "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))}
out["_source"] = note
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
