"""The C-ABI library loads and exports every symbol include/vx_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vx_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^VX_API[^;(]*?\b(vx_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("vx_context_create", "vx_mesh_chunks", "vx_mesh_chunks_device", "vx_greedy_mesh_slices", "vx_cull_chunks",
                 "vx_render_frame", "vx_render_frame_device", "vx_render_mesh", "vx_face_basis", "vx_project_packet",
                 "vx_transform_vertices", "vx_mesh_batch_download", "vx_mesh_batch_upload"):
        assert must in syms
    assert len(syms) >= 31


def test_library_exports_every_declared_symbol():
    from differential_projection_voxel_renderer_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libvx_b200.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"


def test_python_prototypes_cover_the_header():
    from differential_projection_voxel_renderer_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    _lib.load()


def test_struct_layouts_match_the_header():
    from differential_projection_voxel_renderer_b200 import _lib
    assert ctypes.sizeof(_lib.VxAtlas) == 4 * 16 * 4 + 4 * 32
    assert ctypes.sizeof(_lib.VxFrameConfig) == 20 * 4
    assert ctypes.sizeof(_lib.VxMeshBatchInfo) == 16
    assert ctypes.sizeof(_lib.VxFrameStats) == 32
    lib = _lib.load()
    cfg = _lib.VxFrameConfig()
    lib.vx_default_frame_config(ctypes.byref(cfg), 1280, 720)  # host-only helper
    assert (cfg.width, cfg.height, cfg.clear_color, cfg.backface_culling, cfg.enable_shading) == (1280, 720, 0xFF87CEEB, 1, 1)
    assert (cfg.macrotile, cfg.occlusion_culling, cfg.occlusion_grid_w, cfg.occlusion_grid_h) == (0, 0, 128, 72)  # main.rs:46-47, :112
    a = _lib.VxAtlas()
    lib.vx_default_atlas(ctypes.byref(a))
    assert a.palette[1][0] == 0xFF007D00


def test_no_device_means_error_not_fallback():
    """Without a GPU the context cannot be created; with one, creation succeeds.  Either way: no CPU path."""
    from differential_projection_voxel_renderer_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.vx_context_create(0, ctypes.byref(h))
    if rc == 0:
        lib.vx_context_destroy(h)
    else:
        assert rc == _lib.VX_ERR_NO_DEVICE
        assert b"no CPU fallback" in lib.vx_error_string(rc)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "differential_projection_voxel_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "vx_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/vx_b200.h is the drop-in boundary: it must compile as C99 (no C++ in the signatures) and a C program
    must link against libvx_b200.so and find the host-only helpers."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "vx_b200.h"\n'
        "int main(void) {\n"
        "    VxFrameConfig cfg;\n"
        "    vx_default_frame_config(&cfg, 1280, 720);\n"
        "    VxAtlas atlas;\n"
        "    vx_default_atlas(&atlas);\n"
        '    printf("%d %d %d %d %u %zu %zu\\n", cfg.width, cfg.height, cfg.occlusion_grid_w, cfg.occlusion_grid_h, atlas.palette[1][0],\n'
        "           sizeof(VxFrameConfig), sizeof(VxFacePacket32));\n"
        "    return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "differential_projection_voxel_renderer_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lvx_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["1280", "720", "128", "72", str(0xFF007D00), "80", "224"]


def test_ctypes_structs_match_the_c_layout_field_by_field(tmp_path):
    """Every struct that crosses the ABI by pointer: sizeof and the offset of each field as the C compiler sees them in
    include/vx_b200.h equal the ctypes declarations in _lib.py (names included) -- a reordered or resized field on either
    side would otherwise corrupt a call silently."""
    import subprocess
    from differential_projection_voxel_renderer_b200 import _lib
    structs = ["VxFrameConfig", "VxStripeSync", "VxShardLayout", "VxMeshBatchInfo", "VxFrameStats", "VxTerrainParams", "VxMeshBatchDevice"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vx_b200.h"', "int main(void) {"]
    for sname in structs:
        cls = getattr(_lib, sname)
        lines.append(f'    printf("{sname} %zu", sizeof({sname}));')
        for fname, _ in cls._fields_:
            lines.append(f'    printf(" {fname}:%zu", offsetof({sname}, {fname}));')
        lines.append('    printf("\\n");')
    lines += ["    return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    seen = {}
    for row in out.stdout.strip().splitlines():
        parts = row.split()
        seen[parts[0]] = (int(parts[1]), {kv.split(":")[0]: int(kv.split(":")[1]) for kv in parts[2:]})
    for sname in structs:
        cls = getattr(_lib, sname)
        size, offs = seen[sname]
        assert ctypes.sizeof(cls) == size, sname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == offs[fname], f"{sname}.{fname}"
