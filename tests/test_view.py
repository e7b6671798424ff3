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


# ---- the one artefact of the voxel stage the reference tree holds: its README screenshot -------------------------
def _silhouette(rgb):
    clear = np.array([0, 51, 102], np.float32)                    # CLEAR_COLOR (SharedConst.h:8) as UNORM8
    return np.abs(rgb[..., :3].astype(np.float32) - clear).sum(-1) > 60


def _iou(a, b):
    return float((a & b).sum()) / float(max(1, (a | b).sum()))


def _screenshot_ious(render, bound):
    """IoU of the rendered silhouette with the reference screenshot (tests/golden/make_reference_screenshot.py),
    under the reference's DEFAULT camera (DXRVoxelizer.cpp:222-234), for the grid as the current sources define it
    and for its reflection in local x (same camera: reflecting the object = reflecting eye, light and ray origins)."""
    import os
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_screenshot_bunny64.npz"))["rgb"]
    want = _silhouette(ref)
    s2l, eye, light = d.default_view(bound, 480, 270)
    mirror = np.diag([-1.0, 1.0, 1.0, 1.0]).astype(np.float32)
    flip = np.array([-1.0, 1.0, 1.0], np.float32)
    as_is = _iou(_silhouette(render(s2l, eye, light)), want)
    mirrored = _iou(_silhouette(render(s2l @ mirror, eye * flip, light * flip)), want)
    return as_is, mirrored, float(want.mean())


def test_oracle_grid_against_the_reference_screenshot(assets, oracle_mod):
    """bunny, GRID_SIZE 64, MODE_SHADER (what the reference's window shows) through the oracle's viewer pass against
    Doc/Images/SolidVoxelization.jpg.  Finding: the screenshot matches the silhouette to
    IoU 0.99 -- volume, proportions, camera and viewer constants all agree -- but with the OBJECT REFLECTED IN LOCAL X
    under the unchanged default camera; as the current sources define the grid (z flip in the loader,
    XUSGObjLoader.cpp:198,227; x = launch index, hlsl:46-49; tex = (0.5,-0.5,0.5)*pos+0.5, PSRayCast.hlsl:137) the IoU
    is 0.55.  The screenshot therefore predates the shipped loader/shader conventions (or was taken with another
    asset orientation); it pins the shape, not the handedness."""
    m = assets("bunny.obj")
    bits = oracle_mod.voxelize(m.vertices, m.indices, 64, oracle_mod.MODE_SHADER)["bits"]
    as_is, mirrored, cover = _screenshot_ious(lambda s, e, l: oracle_mod.render_view(bits, 64, 480, 270, s, e, l), oracle_mod.bound(m.vertices))
    assert 0.15 < cover < 0.21                                     # the bunny covers ~18 % of the reference's frame
    assert mirrored > 0.97, (as_is, mirrored)
    assert as_is < mirrored


@pytest.mark.gpu
def test_gpu_grid_against_the_reference_screenshot(vox, assets):
    """Same pin for the product path: dxrv_voxelize (MODE_SHADER, 64^3) + dxrv_render_view on the GPU."""
    m = assets("bunny.obj")
    vox.build_bvh(m)
    vox.voxelize(64, d.MODE_SHADER)
    as_is, mirrored, _ = _screenshot_ious(lambda s, e, l: vox.render_view(480, 270, s, e, l), vox.bound())
    assert mirrored > 0.97, (as_is, mirrored)


# ---- the second screenshot: a hi-res bunny that shows the artefact ONLY the reference's shader produces --------------
def _stray_pixels(mask):
    """Pixels of `mask` that are not connected to its largest component (the body), a 4-pixel frame ignored."""
    from scipy import ndimage
    m = mask[4:-4, 4:-4]
    lab, n = ndimage.label(m, structure=np.ones((3, 3)))
    if n == 0:
        return np.zeros_like(m)
    sizes = ndimage.sum(m, lab, range(1, n + 1))
    return m & (lab != 1 + int(np.argmax(sizes)))


def _hires_pin(render_shader, render_parity, bound):
    """Doc/Images/VoxelizationHiRes.jpg (README.md:10, grid size not stated) against a 1920 x 1080 view of the bunny
    under the reference's default camera (reflected in local x like the first screenshot).
    Besides the silhouette, the screenshot shows a spray of stray voxels in the air behind the base of the left ear:
    voxels OUTSIDE the mesh whose radial ray's closest hit passes the normal threshold (hlsl:132-140).  A parity or
    winding-number voxelizer cannot produce them; a faithful restatement of the shader must, and in the same place."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_screenshot_bunny_hires.npz"))
    H, W = (int(v) for v in fx["shape"])
    want = np.unpackbits(fx["mask20"], axis=1)[:, :W].astype(bool)
    want_stray = _stray_pixels(want)
    wy, wx = np.nonzero(want_stray)
    assert 400 < wy.size < 800 and 900 < wx.mean() < 1100 and 250 < wy.mean() < 350      # the fixture itself (578 pixels)

    s2l, eye, light = d.default_view(bound, W, H)
    mirror = np.diag([-1.0, 1.0, 1.0, 1.0]).astype(np.float32)
    flip = np.array([-1.0, 1.0, 1.0], np.float32)
    clear = np.array([0, 51, 102], np.float32)
    out = {}
    for name, render in (("shader", render_shader), ("parity", render_parity)):
        img = render(s2l @ mirror, eye * flip, light * flip)
        got = np.abs(img[..., :3].astype(np.float32) - clear).sum(-1) > 20
        stray = _stray_pixels(got)
        out[name] = (_iou(got[4:-4, 4:-4], want[4:-4, 4:-4]), stray)               # (the capture has a window border)
    iou, stray = out["shader"]
    sy, sx = np.nonzero(stray)
    assert iou > 0.99, iou                                                     # silhouette (0.995 at 256^3)
    assert 0.4 * wy.size < sy.size < 2.5 * wy.size, (sy.size, wy.size)         # as many stray pixels as the screenshot
    assert abs(sx.mean() - wx.mean()) < 40 and abs(sy.mean() - wy.mean()) < 40, (sx.mean(), sy.mean(), wx.mean(), wy.mean())
    box = (sx >= wx.min() - 40) & (sx <= wx.max() + 40) & (sy >= wy.min() - 40) & (sy <= wy.max() + 40)
    assert box.mean() > 0.95, box.mean()                                       # ... in the same region of the frame
    assert out["parity"][0] > 0.99 and out["parity"][1].sum() == 0             # MODE_PARITY: same body, no artefact
    return iou, sy.size


def test_oracle_shader_artefact_against_the_hires_screenshot(assets, oracle_mod):
    """The oracle's MODE_SHADER at 256^3 reproduces the reference screenshot's stray voxels behind the left ear
    (578 pixels around (1003, 307) of the 1920 x 1080 frame; ours ~380 around (1010, 311), 455 / 474 at 384^3 / 512^3),
    its MODE_PARITY shows none: the one place where something the reference holds tells the two functions apart."""
    m = assets("bunny.obj")
    N = 256
    grids = {mode: oracle_mod.voxelize(m.vertices, m.indices, N, mode)["bits"] for mode in (oracle_mod.MODE_SHADER, oracle_mod.MODE_PARITY)}
    _hires_pin(lambda s, e, l: oracle_mod.render_view(grids[oracle_mod.MODE_SHADER], N, 1920, 1080, s, e, l),
               lambda s, e, l: oracle_mod.render_view(grids[oracle_mod.MODE_PARITY], N, 1920, 1080, s, e, l), oracle_mod.bound(m.vertices))


@pytest.mark.gpu
def test_gpu_shader_artefact_against_the_hires_screenshot(vox, assets):
    """Same pin for the product path (direction-bin kernel at 256^3 + dxrv_render_view)."""
    m = assets("bunny.obj")
    vox.build_bvh(m)

    def render(mode):
        def f(s, e, l):
            vox.voxelize(256, mode)
            return vox.render_view(1920, 1080, s, e, l)
        return f
    _hires_pin(render(d.MODE_SHADER), render(d.MODE_PARITY), vox.bound())


def test_save_image_writes_the_reference_screenshot_format(tmp_path):
    """dxrv_save_image = DXRVoxelizer::SaveImage (DXRVoxelizer.cpp:531-551): an RGB (default) or RGBA PNG from an
    R8G8B8A8 buffer with a row pitch; decoded here by an independent reader (PIL)."""
    Image = pytest.importorskip("PIL.Image")
    from dxrvoxelizer_b200 import _lib as L
    rng = np.random.default_rng(7)
    for h, w in [(720, 1280), (1, 1), (37, 513), (3, 30000)]:          # (the last two: rows that straddle stored-block limits)
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        for comp in (3, 4):
            p = tmp_path / ("img_%d_%d_%d.png" % (h, w, comp))
            d.save_image(p, img, comp)
            got = np.asarray(Image.open(p))
            assert got.shape == (h, w, comp) and np.array_equal(got, img[:, :, :comp])
    # a row pitch wider than the image (the reference's read-back buffer is pitched: rowPitch / 4 pixels per row)
    wide = rng.integers(0, 256, (20, 48, 4), dtype=np.uint8)
    p = tmp_path / "pitched.png"
    assert L.lib().dxrv_save_image(str(p).encode(), wide.ctypes.data, 40, 20, 48 * 4, 3) == L.OK
    assert np.array_equal(np.asarray(Image.open(p)), wide[:, :40, :3])
    # argument checks
    lib = L.lib()
    assert lib.dxrv_save_image(str(p).encode(), wide.ctypes.data, 40, 20, 40 * 4 - 1, 3) == L.ERR_INVALID_ARG
    assert lib.dxrv_save_image(str(p).encode(), wide.ctypes.data, 40, 20, 48 * 4, 2) == L.ERR_INVALID_ARG
    assert lib.dxrv_save_image(None, wide.ctypes.data, 40, 20, 48 * 4, 3) == L.ERR_INVALID_ARG
    assert lib.dxrv_save_image(str(tmp_path / "no_such_dir" / "x.png").encode(), wide.ctypes.data, 40, 20, 48 * 4, 3) == L.ERR_IO
