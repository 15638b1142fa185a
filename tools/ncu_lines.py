"""Per-source-line totals (samples, warp instructions) of one kernel from an .ncu-rep (needs -lineinfo):
   python tools/ncu_lines.py rep kernel_regex [n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, seen_kernel = [], "", 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        seen_kernel += 1
        if seen_kernel > 1 and "Kernel" in r[0]:
            pass
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0].isdigit():
        try:
            si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
            extra = len(r) - len(hdr)  # unquoted commas inside the source text shift the columns
            lines.append((fname, int(r[0]), ",".join(r[1:2 + extra]).strip(), int(r[si + extra] or 0), int(r[ei + extra] or 0)))
        except (ValueError, IndexError):
            pass
# several launches may repeat: collapse by (file,line)
agg = {}
for f, ln, src, s, e in lines:
    k = (f, ln)
    a = agg.setdefault(k, [src, 0, 0])
    a[1] += s; a[2] += e
tot_s = sum(a[1] for a in agg.values()); tot_e = sum(a[2] for a in agg.values())
print(f"samples {tot_s}  warp-instructions {tot_e}")
for (f, ln), (src, s, e) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:n]:
    print(f"{f}:{ln:<5d} inst {e:10d} {100*e/max(tot_e,1):5.1f}%  samp {s:6d} {100*s/max(tot_s,1):5.1f}%  {src[:90]}")
