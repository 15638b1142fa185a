"""Print the hottest SASS lines (warp-stall samples) of one kernel from an .ncu-rep: python tools/ncu_top.py rep kernel_regex [n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
blk = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    blk.append(r)
ci = {k: hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed") if k in hdr}
tot = sum(int(r[ci["# Samples"]] or 0) for r in blk)
inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in blk)
print(f"total samples {tot}, warp instructions executed {inst}, sass lines {len(blk)}")
order = sorted(range(len(blk)), key=lambda i: -int(blk[i][ci["# Samples"]] or 0))
for i in order[:n]:
    r = blk[i]
    print(f"{i:5d} {int(r[ci['# Samples']]):7d} {100*int(r[ci['# Samples']])/max(tot,1):5.1f}%  exec {int(r[ci['Instructions Executed']]):9d}  thr {r[ci['Avg. Threads Executed']]:>5}  {r[ci['Source']].strip()[:110]}")
