"""Viewer pass: camera constants against an independent numpy restatement
of DirectXMath's LookAtLH / PerspectiveFovLH chain, the oracle's ray-march on the CPU, and (GPU) the CUDA
kernel against the oracle with a tolerance on the 8-bit image."""
import numpy as np
import pytest

import dxrvoxelizer_b200 as d


def _numpy_view(bound, w, h, pos_scale=(0, 0, 0, 1)):
    def T(x, y, z):
        m = np.eye(4); m[3, :3] = (x, y, z); return m
    def S(s):
        m = np.eye(4); m[0, 0] = m[1, 1] = m[2, 2] = s; return m
    eye, focus, up = np.array([8.0, 12.0, -14.0]), np.array([0.0, 4.0, 0.0]), np.array([0.0, 1.0, 0.0])
    z = focus - eye; z /= np.linalg.norm(z)
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    view = np.eye(4); view[:3, 0], view[:3, 1], view[:3, 2] = x, y, z
    view[3, :3] = (-x @ eye, -y @ eye, -z @ eye)
    hh = 1.0 / np.tan(0.785398163 / 2); zn, zf = 1.0, 1000.0
    proj = np.zeros((4, 4)); proj[0, 0] = hh / (w / h); proj[1, 1] = hh; proj[2, 2] = zf / (zf - zn); proj[2, 3] = 1; proj[3, 2] = -zn * zf / (zf - zn)
    world = S(bound[3]) @ T(*bound[:3]) @ S(pos_scale[3]) @ T(*pos_scale[:3])
    to_screen = np.array([[0.5 * w, 0, 0, 0], [0, -0.5 * h, 0, 0], [0, 0, 1, 0], [0.5 * w, 0.5 * h, 0, 1]])
    s2l = np.linalg.inv(world @ view @ proj @ to_screen)
    wi = np.linalg.inv(world)
    def coord(p):
        r = np.append(p, 1.0) @ wi
        return r[:3] / r[3]
    return s2l, coord(eye), coord(np.array([-10.0, 45.0, -75.0]))


@pytest.mark.parametrize("pos_scale", [None, (0.0, 2.8, 0.0, 0.03)])
def test_default_view_matches_directxmath_restatement(pos_scale):
    bound = np.array([0.0, 4.96995, 0.0, 7.0467], np.float32)
    m, eye, light = d.default_view(bound, 1280, 720, pos_scale)
    s2l, e, l = _numpy_view(bound.astype(np.float64), 1280, 720, pos_scale or (0, 0, 0, 1))
    np.testing.assert_allclose(m, s2l, rtol=1e-5, atol=1e-6 * np.abs(s2l).max())
    np.testing.assert_allclose(eye, e, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(light, l, rtol=1e-5, atol=1e-6)


def test_oracle_view_of_a_sphere_is_sane(meshes_mod, oracle_mod):
    m = meshes_mod.icosphere(3, seed=2)
    bits = oracle_mod.voxelize(m.vertices, m.indices, 32, 1)["bits"]
    s2l, eye, light = d.default_view([0, 4, 0, 6], 160, 90)
    img = oracle_mod.render_view(bits, 32, 160, 90, s2l, eye, light)
    clear = np.array([0, 51, 102, 0], np.uint8)                  # CLEAR_COLOR as UNORM8, alpha 0 on a miss
    assert (img[0, 0] == clear).all() and (img[-1, -1] == clear).all()
    centre = img[45, 80]
    assert centre[3] == 255 and not (centre[:3] == clear[:3]).all()   # the object covers the middle of the frame
    empty = oracle_mod.render_view(np.zeros_like(bits), 32, 160, 90, s2l, eye, light)
    # empty volume: rays that enter the cube leave with transmit 1 -> sqrt(clear^2) = clear colour, alpha 1
    inside = empty[..., 3] == 255
    assert inside.any() and (np.abs(empty[inside][:, :3].astype(int) - clear[:3].astype(int)) <= 1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,N", [("bunny.obj", 64), ("dragon.obj", 128)])
def test_gpu_view_matches_oracle(vox, assets, oracle_mod, name, N):
    m = assets(name)
    vox.build_bvh(m)
    vox.voxelize(N, d.MODE_SHADER)                                 # what the reference's viewer shows
    bits = vox.fetch_bits()
    s2l, eye, light = d.default_view(vox.bound(), 320, 180)
    got = vox.render_view(320, 180, s2l, eye, light).astype(int)
    ref = oracle_mod.render_view(bits, N, 320, 180, s2l, eye, light).astype(int)
    diff = np.abs(got - ref).max(axis=-1)
    assert (got[..., 3] == ref[..., 3]).mean() > 0.999             # hit / miss classification
    assert (diff <= 2).mean() > 0.995 and np.median(diff) == 0     # fp32 reassociation / FMA noise only
    assert (ref[..., 3] == 255).mean() > 0.05
    with pytest.raises(d.DxrvError):
        vox.voxelize(N, d.MODE_PARITY, 0, N // 2)
        vox.render_view(320, 180, s2l, eye, light)                 # needs the full grid
