"""Static resource table of every kernel in libvx_b200.so from the ptxas logs the csrc Makefile writes
(build/*.ptxas.log):   python tools/ptxas_table.py > profiles/rNN_static_kernels.md"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                 r".*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?")


def main():
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, "differential_projection_voxel_renderer_b200", "csrc", "build", "*.ptxas.log"))):
        for m in PAT.finditer(open(f).read()):
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*$", "", name)
            rows.append((os.path.basename(f).replace(".ptxas.log", ".cu"), name, int(m.group(5)), int(m.group(2)), int(m.group(3)),
                         int(m.group(4)), int(m.group(8) or 0)))
    print("# Static resources of every kernel in libvx_b200.so\n")
    print("`nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v` (the `build/*.ptxas.log` files the csrc Makefile writes);")
    print("regenerate with `python tools/ptxas_table.py`.\n")
    print("| file | kernel | registers | stack B | spill st B | spill ld B | static smem B |\n|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %s | `%s` | %d | %d | %d | %d | %d |" % r)
    print("\nOnly the traced (diagnostic) raster variant and the horizon kernel spill; the three hot kernels of a frame")
    print("(`frame_cull_kernel`, `frame_setup_kernel<false>`, `frame_raster_kernel<false, false>`) and `mesh_chunks_kernel` do not.")


if __name__ == "__main__":
    main()
