"""Camera / matrix helpers that generate the view-projection INPUT (16 f32,
column-major like glam `Mat4::to_cols_array`).

Restates glam 0.25 `Mat4::perspective_rh`, `Mat4::look_at_rh`, `Mat4 * Mat4` and
the reference `Camera` (/root/reference/src/camera/mod.rs:20-61).  glam is not
vendored with the reference, so the bit pattern of VP is parity-unpinned; VP is
therefore an input that crosses the C ABI as 16 floats, and the oracle and the
CUDA path always consume identical bits.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def perspective_rh(fov_y: float, aspect: float, z_near: float, z_far: float) -> np.ndarray:
    fov_y, aspect, z_near, z_far = f32(fov_y), f32(aspect), f32(z_near), f32(z_far)
    s, c = f32(np.sin(f32(0.5) * fov_y)), f32(np.cos(f32(0.5) * fov_y))
    h = f32(c / s)
    w = f32(h / aspect)
    r = f32(z_far / f32(z_near - z_far))
    m = np.zeros((4, 4), dtype=f32)  # m[col, row]
    m[0, 0] = w
    m[1, 1] = h
    m[2, 2] = r
    m[2, 3] = f32(-1.0)
    m[3, 2] = f32(r * z_near)
    return m


def _normalize(v):
    v = np.asarray(v, dtype=f32)
    return (v / f32(np.sqrt(f32(np.dot(v, v))))).astype(f32)


def look_at_rh(eye, center, up) -> np.ndarray:
    eye = np.asarray(eye, dtype=f32)
    f = _normalize(np.asarray(center, dtype=f32) - eye)
    s = _normalize(np.cross(f, np.asarray(up, dtype=f32)).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.zeros((4, 4), dtype=f32)
    m[0] = [s[0], u[0], -f[0], 0]
    m[1] = [s[1], u[1], -f[1], 0]
    m[2] = [s[2], u[2], -f[2], 0]
    m[3] = [-f32(np.dot(eye, s)), -f32(np.dot(eye, u)), f32(np.dot(eye, f)), 1]
    return m


def mat4_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """glam Mat4 * Mat4: each column of b through a.mul_vec4 (unfused, left to right)."""
    out = np.zeros((4, 4), dtype=f32)
    for c in range(4):
        acc = (a[0] * b[c, 0]).astype(f32)
        acc = (acc + a[1] * b[c, 1]).astype(f32)
        acc = (acc + a[2] * b[c, 2]).astype(f32)
        acc = (acc + a[3] * b[c, 3]).astype(f32)
        out[c] = acc
    return out


def _rot_y(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float64)


def _rot_x(a):
    return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]], dtype=np.float64)


class Camera:
    """camera/mod.rs:5-61: fov 70 deg, near 0.1, far 1000, yaw/pitch FPS camera."""

    def __init__(self, position, aspect_ratio: float, yaw: float = 0.0, pitch: float = 0.0,
                 fov_deg: float = 70.0, near: float = 0.1, far: float = 1000.0):
        self.position = np.asarray(position, dtype=f32)
        self.aspect_ratio = float(aspect_ratio)
        self.yaw = float(yaw)
        self.pitch = float(pitch)
        self.fov = float(np.radians(f32(fov_deg)))
        self.near = near
        self.far = far

    def view_matrix(self) -> np.ndarray:
        rot = _rot_y(self.yaw) @ _rot_x(self.pitch)
        fwd = (rot @ np.array([0.0, 0.0, -1.0])).astype(f32)
        up = (rot @ np.array([0.0, 1.0, 0.0])).astype(f32)
        return look_at_rh(self.position, self.position + fwd, up)

    def projection_matrix(self) -> np.ndarray:
        return perspective_rh(self.fov, self.aspect_ratio, self.near, self.far)

    def view_projection(self) -> np.ndarray:
        """16 f32, column-major (camera/mod.rs:59-61)."""
        return np.ascontiguousarray(mat4_mul(self.projection_matrix(), self.view_matrix()).reshape(16))
