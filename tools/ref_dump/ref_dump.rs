//! Reference vector dump for the B200 port's parity tests (tests/test_reference_golden.py).
//!
//! Drop this file into the reference crate as `examples/ref_dump.rs` and run, from the crate root:
//!
//!     cargo run --release --example ref_dump > ref_vectors.json
//!
//! then copy `ref_vectors.json` to `tests/golden/ref_vectors.json` of the port.  It uses only the crate's public API
//! (`Chunk`, `BinaryGreedyMesher`, `Camera`, `Rasterizer`, `Framebuffer`) and no extra dependency; JSON is written by
//! hand.  What it pins (each item is checked against the CPU oracle and against the CUDA path):
//!   * `heights`   sample of `Chunk::generate_terrain` heights (top solid voxel per column) -> pins `noise 0.9` Perlin
//!   * `chunks`    per chunk: uniform kind, FNV-1a of the 32768 voxel bytes, quads per (face, slice) in mesher order
//!   * `camera`    the 16 f32 bit patterns of `Camera::new((0,10,20), 16/9).view_projection_matrix()` -> pins glam
//!   * `frames`    `Rasterizer::render_mesh` of the listed meshes, in list order, into a cleared framebuffer:
//!                 FNV-1a of colour and of depth bits, covered-pixel count, plus the raw rows of one scanline band
use glam::{IVec3, Vec3};
use voxel_engine::meshing::FaceDir;
use voxel_engine::{BinaryGreedyMesher, Camera, Chunk, ChunkMesh, Framebuffer, Rasterizer, CHUNK_SIZE};

fn fnv1a(bytes: impl Iterator<Item = u8>) -> u64 {
    let mut h: u64 = 0xcbf29ce484222325;
    for b in bytes {
        h ^= b as u64;
        h = h.wrapping_mul(0x100000001b3);
    }
    h
}

fn chunk_voxel_bytes(chunk: &Chunk) -> Vec<u8> {
    // index = z * 32 * 32 + y * 32 + x (voxel/chunk.rs:52)
    let mut out = Vec::with_capacity(CHUNK_SIZE * CHUNK_SIZE * CHUNK_SIZE);
    for z in 0..CHUNK_SIZE {
        for y in 0..CHUNK_SIZE {
            for x in 0..CHUNK_SIZE {
                out.push(chunk.get_block(x, y, z).block_type as u8);
            }
        }
    }
    out
}

const FACES: [FaceDir; 6] = [FaceDir::PosX, FaceDir::NegX, FaceDir::PosY, FaceDir::NegY, FaceDir::PosZ, FaceDir::NegZ];

fn mesh_json(mesh: &ChunkMesh) -> String {
    // [[face, slice, [[u, v, w, h, block_type], ...]], ...] for non-empty slices, in FaceDir / slice order
    let mut parts = Vec::new();
    for (fi, f) in FACES.iter().enumerate() {
        let list = mesh.face_list(*f);
        for s in 0..32 {
            if list.slice_quads[s].is_empty() {
                continue;
            }
            let quads: Vec<String> = list.slice_quads[s]
                .iter()
                .map(|q| format!("[{},{},{},{},{}]", q.u(), q.v(), q.width(), q.height(), q.block_type()))
                .collect();
            parts.push(format!("[{},{},[{}]]", fi, s, quads.join(",")));
        }
    }
    format!("[{}]", parts.join(","))
}

fn main() {
    let positions: Vec<IVec3> = vec![
        IVec3::new(0, 0, 0), IVec3::new(0, -1, 0), IVec3::new(1, 0, 0), IVec3::new(-1, 0, 0), IVec3::new(0, 0, 1),
        IVec3::new(0, 0, -1), IVec3::new(0, 1, 0), IVec3::new(3, 0, -5), IVec3::new(-12, 0, 11), IVec3::new(2, -2, 2),
        IVec3::new(0, 0, -2), IVec3::new(1, 0, -2), IVec3::new(-1, 0, -2), IVec3::new(0, -1, -2),
    ];
    let chunks: Vec<Chunk> = positions.iter().map(|p| Chunk::generate_terrain(*p)).collect();
    let refs: Vec<&Chunk> = chunks.iter().collect();

    // ---- heights: top solid voxel of each column of the y = 0 and y = -1 chunk layers (world y in -32..32)
    let mut heights = Vec::new();
    for (cx, cz) in [(0, 0), (3, -5), (-12, 11)] {
        let upper = Chunk::generate_terrain(IVec3::new(cx, 0, cz));
        let lower = Chunk::generate_terrain(IVec3::new(cx, -1, cz));
        let mut rows = Vec::new();
        for z in 0..CHUNK_SIZE {
            let mut row = Vec::new();
            for x in 0..CHUNK_SIZE {
                let mut h: i32 = -33;
                for wy in (-32..32).rev() {
                    let (c, ly) = if wy >= 0 { (&upper, wy as usize) } else { (&lower, (wy + 32) as usize) };
                    if c.get_block(x, ly, z).block_type as u8 != 0 {
                        h = wy;
                        break;
                    }
                }
                row.push(h.to_string());
            }
            rows.push(format!("[{}]", row.join(",")));
        }
        heights.push(format!("{{\"chunk_xz\":[{},{}],\"top\":[{}]}}", cx, cz, rows.join(",")));
    }

    // ---- chunks: voxels + meshes with neighbours (mesh_chunk_in_world over the whole list)
    let mut chunk_json = Vec::new();
    let mut meshes: Vec<Option<ChunkMesh>> = Vec::new();
    for (i, c) in chunks.iter().enumerate() {
        let p = positions[i];
        let kind = if c.is_uniform() { c.uniform_block_type().map(|b| b as i32 + 1).unwrap_or(-1) } else { 0 };
        let vox = chunk_voxel_bytes(c);
        let mesh = BinaryGreedyMesher::mesh_chunk_in_world(c, &refs);
        let alone = BinaryGreedyMesher::mesh_chunk(c);
        chunk_json.push(format!(
            "{{\"pos\":[{},{},{}],\"uniform\":{},\"voxels_fnv\":\"{:016x}\",\"quads_in_world\":{},\"quad_count_in_world\":{},\"quads_alone\":{}}}",
            p.x, p.y, p.z, kind, fnv1a(vox.iter().copied()),
            mesh.as_ref().map(mesh_json).unwrap_or_else(|| "null".to_string()),
            mesh.as_ref().map(|m| m.quad_count()).unwrap_or(0),
            alone.as_ref().map(mesh_json).unwrap_or_else(|| "null".to_string()),
        ));
        meshes.push(mesh);
    }

    // ---- camera (main.rs:51, benches/rendering.rs:17)
    let camera = Camera::new(Vec3::new(0.0, 10.0, 20.0), 1280.0 / 720.0);
    let vp = camera.view_projection_matrix();
    let vp_bits: Vec<String> = vp.to_cols_array().iter().map(|f| f.to_bits().to_string()).collect();

    // ---- frames: listed meshes drawn in list order with Rasterizer::render_mesh (rasterizer.rs:385)
    let mut frames = Vec::new();
    for (w, h, list) in [(1280usize, 720usize, vec![0usize]), (640, 360, vec![0, 2, 3, 4, 5, 10, 11, 12]), (320, 180, vec![12, 11, 10, 5, 0])] {
        let cam = Camera::new(Vec3::new(0.0, 10.0, 20.0), w as f32 / h as f32);
        let vpm = cam.view_projection_matrix();
        let mut fb = Framebuffer::new(w, h);
        fb.clear(0xFF87CEEB);
        let mut rast = Rasterizer::new();
        let mut drawn = Vec::new();
        for &i in &list {
            if let Some(m) = &meshes[i] {
                rast.render_mesh(m, &vpm, &mut fb);
                drawn.push(i.to_string());
            }
        }
        let covered = fb.color_buffer.iter().filter(|&&c| c != 0xFF87CEEB).count();
        let color_fnv = fnv1a(fb.color_buffer.iter().flat_map(|c| c.to_le_bytes()));
        let depth_fnv = fnv1a(fb.depth_buffer.iter().flat_map(|d| d.to_bits().to_le_bytes()));
        // four raw scanlines around the lower third of the frame (ground is there for this camera)
        let y0 = h * 2 / 3;
        let rows_c: Vec<String> = (y0..y0 + 4)
            .map(|y| format!("[{}]", fb.color_buffer[y * w..(y + 1) * w].iter().map(|c| c.to_string()).collect::<Vec<_>>().join(",")))
            .collect();
        let rows_d: Vec<String> = (y0..y0 + 4)
            .map(|y| format!("[{}]", fb.depth_buffer[y * w..(y + 1) * w].iter().map(|d| d.to_bits().to_string()).collect::<Vec<_>>().join(",")))
            .collect();
        let vpb: Vec<String> = vpm.to_cols_array().iter().map(|f| f.to_bits().to_string()).collect();
        frames.push(format!(
            "{{\"width\":{},\"height\":{},\"meshes\":[{}],\"vp_bits\":[{}],\"covered\":{},\"color_fnv\":\"{:016x}\",\"depth_fnv\":\"{:016x}\",\"row0\":{},\"color_rows\":[{}],\"depth_bits_rows\":[{}]}}",
            w, h, drawn.join(","), vpb.join(","), covered, color_fnv, depth_fnv, y0, rows_c.join(","), rows_d.join(",")
        ));
    }

    println!(
        "{{\"format\":1,\"crate\":\"voxel_engine 0.1.0\",\"heights\":[{}],\"chunks\":[{}],\"camera\":{{\"position\":[0.0,10.0,20.0],\"aspect\":\"1280/720\",\"vp_bits\":[{}]}},\"frames\":[{}]}}",
        heights.join(","), chunk_json.join(","), vp_bits.join(","), frames.join(",")
    );
}
