"""Known answers for the Book-2 features the reference does not contain (quads, rotate_y, sphere (u,v), image and
checker textures) - independent of both the oracle and the kernels: the expected values below are worked out by hand
from the book's definitions (and, for the checker, from the reference's own checker_texture::value,
main/src/rt_engine/shaders/cu_Textures.cuh:32-39).  The same cases run against the CPU oracle (tests/test_cpu_oracle.py)
and against the CUDA path (tests/test_gpu_parity.py); `trace(scene, rays)` returns rtb_hit records and
`render(scene, cam, w, h, spp, depth)` returns radiance sums with the sample count in alpha."""
import math

import numpy as np


def _rays(rtb, o, d, time=0.0):
    o = np.atleast_2d(np.asarray(o, dtype=np.float32)); d = np.atleast_2d(np.asarray(d, dtype=np.float32))
    r = np.zeros(len(o), dtype=rtb.RAY_DTYPE)
    r["o"] = o; r["d"] = d; r["time"] = time
    return r


def check_quad_alpha_beta(rtb, trace):
    """quad(Q, u, v): a ray through Q + a u + b v reports (u, v) = (a, b), t = the plane distance, the normal facing the
    ray, front_face by the side it came from - for a skewed (non-rectangular, non-axis-aligned) parallelogram too; rays
    just outside any of the four edges miss."""
    s = rtb.Scene(); m = s.lambertian(tex=s.image(np.full((4, 4, 3), 128, dtype=np.uint8)))
    Q = np.array([1.0, 2.0, -3.0]); u = np.array([4.0, 0.5, 1.0]); v = np.array([-1.0, 3.0, 0.25])
    s.set_root(s.list([s.quad(Q, u, v, m)]))
    n = np.cross(u, v); n /= np.linalg.norm(n)
    ab = np.array([[0.5, 0.5], [0.125, 0.875], [0.9375, 0.0625], [0.001, 0.001], [0.999, 0.999], [0.25, 0.75]])
    pts = Q + ab[:, :1] * u + ab[:, 1:] * v
    for side in (+1.0, -1.0):
        o = pts + side * 7.0 * n + np.array([0.3, -0.2, 0.1])
        h = trace(s, _rays(rtb, o, pts - o))
        assert (h["object"] == 0).all()
        np.testing.assert_allclose(h["t"], 1.0, rtol=2e-6)
        np.testing.assert_allclose(np.stack([h["u"], h["v"]], 1), ab, atol=3e-6)
        np.testing.assert_allclose(h["p"], pts, atol=2e-5)
        np.testing.assert_allclose(h["n"], np.broadcast_to(side * n, (len(pts), 3)), atol=2e-6)     # facing the ray (book set_face_normal)
        # the quad's own normal is n = unit(u x v): front face <=> the ray runs against it
        assert (h["front_face"] == (1 if side > 0 else 0)).all()
    out = np.array([[-0.001, 0.5], [1.001, 0.5], [0.5, -0.001], [0.5, 1.001]])
    po = Q + out[:, :1] * u + out[:, 1:] * v
    o = po + 5.0 * n
    assert (trace(s, _rays(rtb, o, po - o))["object"] == -1).all()
    # a ray parallel to the plane misses; one starting behind the quad and pointing away misses (t >= 0 only)
    assert trace(s, _rays(rtb, Q + 2.0 * n, u))["object"][0] == -1
    assert trace(s, _rays(rtb, pts[0] + 3.0 * n, n))["object"][0] == -1


def check_rotate_y_quarter_turn(rtb, trace):
    """rotate_y(theta) maps an object point (x, y, z) to (cos x + sin z, y, -sin x + cos z): at 90 degrees (x, y, z) ->
    (z, y, -x).  A unit quad at object x = 2 facing +x therefore stands at world z = -2 facing -z; translate then adds its
    offset.  Hit distance, point, normal and (u, v) are the un-rotated quad's."""
    s = rtb.Scene(); m = s.lambertian(tex=s.image(np.full((4, 4, 3), 128, dtype=np.uint8)))
    quad = s.quad((2.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), m)          # spans z, y in [0,1] at x = 2; normal u x v = (-1, 0, 0)
    s.set_root(s.list([s.translate(s.rotate_y(quad, 90.0), (10.0, 20.0, 30.0))]))
    # object (2, y, z) -> world (z, y, -2) + offset
    a, b = 0.25, 0.75                                                              # u runs along object z, v along y
    target = np.array([a + 10.0, b + 20.0, -2.0 + 30.0])
    for dz, front in ((-5.0, 0), (+5.0, 1)):      # object normal (-1,0,0) is world (0,0,+1): coming from z > 28 runs against it
        o = target + np.array([0.0, 0.0, dz])
        h = trace(s, _rays(rtb, o, target - o))
        assert h["object"][0] >= 0
        np.testing.assert_allclose(h["t"][0], 1.0, rtol=1e-5)
        np.testing.assert_allclose(h["p"][0], target, atol=1e-4)
        np.testing.assert_allclose([h["u"][0], h["v"][0]], [a, b], atol=1e-5)
        np.testing.assert_allclose(h["n"][0], [0.0, 0.0, -1.0 if dz < 0 else 1.0], atol=1e-6)
        assert h["front_face"][0] == front
    # where the un-rotated quad would have been there is nothing
    o = np.array([15.0, 20.5, 30.5])
    assert trace(s, _rays(rtb, o, np.array([12.0, 20.5, 30.5]) - o))["object"][0] == -1


def check_sphere_uv(rtb, trace):
    """book sphere::get_sphere_uv on the outward unit normal: theta = acos(-y), phi = atan2(-z, x) + pi, u = phi / 2 pi,
    v = theta / pi.  Known points (the book's own table): +x -> (0.5, 0.5); +y -> v = 1; -y -> v = 0; +z -> u = 0.25; -z -> u = 0.75;
    the seam at -x: just on the +z side u ~ 0, just on the -z side u ~ 1."""
    s = rtb.Scene(); m = s.lambertian(tex=s.image(np.full((4, 4, 3), 128, dtype=np.uint8)))
    c = np.array([3.0, -2.0, 5.0]); R = 2.0
    s.set_root(s.list([s.sphere(c, R, m)]))
    e = 1e-3
    cases = [((1, 0, 0), 0.5, 0.5), ((0, 0, 1), 0.25, 0.5), ((0, 0, -1), 0.75, 0.5), ((0, 1, 0), None, 1.0), ((0, -1, 0), None, 0.0),
             ((-1, 0, e), 0.0, 0.5), ((-1, 0, -e), 1.0, 0.5), ((1, 1, 0), 0.5, 0.75), ((0, -1, -1), 0.75, 0.25)]
    for n, u, v in cases:
        n = np.array(n, dtype=np.float64); n /= np.linalg.norm(n)
        o = c + 5.0 * R * n
        h = trace(s, _rays(rtb, o, -n))
        assert h["object"][0] == 0
        np.testing.assert_allclose(h["t"][0], 4.0 * R, rtol=1e-6)
        np.testing.assert_allclose(h["n"][0], n, atol=2e-6)
        np.testing.assert_allclose(h["v"][0], v, atol=3e-4)
        if u is not None:
            np.testing.assert_allclose(h["u"][0], u, atol=3e-4)
        assert 0.0 <= h["u"][0] <= 1.0 and 0.0 <= h["v"][0] <= 1.0
        assert h["front_face"][0] == 1
    # from inside: same point and outward normal (spheres never flip it, SphereHittable.cu:64), back face
    h = trace(s, _rays(rtb, c, np.array([1.0, 0.0, 0.0])))
    np.testing.assert_allclose(h["t"][0], R, rtol=1e-6); np.testing.assert_allclose(h["n"][0], [1, 0, 0], atol=1e-6)
    assert h["front_face"][0] == 0


def check_checker_at_negative_coordinates(rtb, trace, render):
    """checker_texture::value (cu_Textures.cuh:32-39): i = ivec3(pos * inv_scale) is a C cast - truncation toward zero, not
    floor - and parity is the C `%` of the component sum.  So the cells on either side of a zero coordinate merge into one
    double-width cell (both (-1,0) and (0,1) truncate to 0), unlike the book's floor-based checker.  A checkered floor quad
    under a white sky, one camera sample per pixel at depth 2: every pixel is exactly the texel colour at its hit point."""
    even, odd = (0.2, 0.3, 0.1), (0.9, 0.9, 0.9)
    s = rtb.Scene()
    tex = s.checker(1.0, s.solid(even), s.solid(odd))
    s.set_root(s.list([s.quad((-4.0, 0.0, -4.0), (8.0, 0.0, 0.0), (0.0, 0.0, 8.0), s.lambertian(tex=tex))]))
    s.set_background(rtb.BG_CONSTANT, (1.0, 1.0, 1.0))
    W = H = 64
    cam = rtb.make_camera("pinhole", (0.0, 10.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, -1.0), 40.0, 1.0)
    acc = render(s, cam, W, H, 1, 2)
    # the camera's own rays through the pixel centres would land on the same cells except where a cell border crosses the
    # pixel filter's half-pixel disc: compare on pixels whose hit point (from pixel-centre rays) is well inside a cell
    xs = (np.arange(W, dtype=np.float32) + np.float32(0.5)) / np.float32(W) * np.float32(2) - np.float32(1)
    U, V = np.meshgrid(xs, xs)
    o = np.array(cam.o[:], dtype=np.float32); cu = np.array(cam.u[:], dtype=np.float32); cv = np.array(cam.v[:], dtype=np.float32); cw = np.array(cam.w[:], dtype=np.float32)
    d = cw[None, None, :] + cu[None, None, :] * U[..., None] + cv[None, None, :] * V[..., None]
    h = trace(s, _rays(rtb, np.broadcast_to(o, (W * H, 3)), d.reshape(-1, 3)))
    assert (h["object"] >= 0).all()
    p = h["p"].astype(np.float64)
    cell = np.trunc(p)                                                    # C cast
    parity = np.fmod(cell.sum(axis=1), 2.0)                               # C %: sign of the dividend
    expect = np.where((parity == 0)[:, None], np.array(even), np.array(odd))
    frac = np.abs(p - np.round(p))
    inside = (np.minimum(frac[:, 0], frac[:, 2]) > 0.12)                  # 8 / 64 = 0.125 units per pixel: stay a pixel away from borders
    got = acc[..., :3].reshape(-1, 3)
    assert inside.sum() > 1500
    np.testing.assert_allclose(got[inside], expect[inside], atol=1e-6)
    # the merged cells around zero: x in (-1, 1), z in (0, 1) is ONE colour under truncation (floor would alternate)
    left = inside & (p[:, 0] > -0.85) & (p[:, 0] < -0.15) & (p[:, 2] > 0.15) & (p[:, 2] < 0.85)
    right = inside & (p[:, 0] > 0.15) & (p[:, 0] < 0.85) & (p[:, 2] > 0.15) & (p[:, 2] < 0.85)
    assert left.any() and right.any()
    assert np.allclose(got[left], even, atol=1e-6) and np.allclose(got[right], even, atol=1e-6)
    # a negative odd sum gives % == -1, i.e. "not even" (x in (-2,-1), z in (0,1): cell sum -1)
    neg = inside & (p[:, 0] > -1.85) & (p[:, 0] < -1.15) & (p[:, 2] > 0.15) & (p[:, 2] < 0.85)
    assert neg.any() and np.allclose(got[neg], odd, atol=1e-6)


def check_image_texture_lookup(rtb, render):
    """book image_texture::value: nearest texel, u across, v up (row 0 of the image is the top), bytes / 255.  A quad
    carrying a 4x2 image of distinct colours under a white sky: eight flat rectangles of exactly those colours."""
    img = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0]],
                    [[0, 255, 255], [255, 0, 255], [51, 102, 153], [204, 17, 34]]], dtype=np.uint8)      # row 0 = top
    s = rtb.Scene()
    s.set_root(s.list([s.quad((-2.0, -1.0, 0.0), (4.0, 0.0, 0.0), (0.0, 2.0, 0.0), s.lambertian(tex=s.image(img)))]))
    s.set_background(rtb.BG_CONSTANT, (1.0, 1.0, 1.0))
    W, H = 64, 32
    # orthographic-like view from far away so the quad fills the image exactly: vfov such that the viewport is 4 x 2 at z = 0
    dist = 1000.0
    cam = rtb.make_camera("pinhole", (0.0, 0.0, dist), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), math.degrees(2.0 * math.atan(1.0 / dist)), 2.0)
    acc = render(s, cam, W, H, 1, 2)[..., :3]
    # camera u = cross(up, w) with w = -z points to -x: image column x looks at world x = +2 - ...; find the mapping from the data
    # rows: image row 0 of the float buffer is the bottom (v = 0 .. 0.5 -> image row 1, the lower texel row)
    cells = acc.reshape(2, 16, 4, 16, 3)
    for j in range(2):
        for i in range(4):
            block = cells[j, 2:14, i, 2:14]                 # interior of the rectangle
            assert np.ptp(block.reshape(-1, 3), axis=0).max() < 1e-6, "a texel rectangle is not flat"
    bottom_left_to_right = [tuple(np.round(cells[0, 8, i, 8] * 255).astype(int)) for i in range(4)]
    top_left_to_right = [tuple(np.round(cells[1, 8, i, 8] * 255).astype(int)) for i in range(4)]
    want_top = [tuple(int(c) for c in img[0, i]) for i in range(4)]; want_bottom = [tuple(int(c) for c in img[1, i]) for i in range(4)]
    # the view may be mirrored left-right (the reference's camera basis has u = cross(up, w)): accept the mirror, not a swap of rows
    assert (top_left_to_right == want_top and bottom_left_to_right == want_bottom) or \
           (top_left_to_right == want_top[::-1] and bottom_left_to_right == want_bottom[::-1]), (top_left_to_right, bottom_left_to_right)
    np.testing.assert_allclose(cells[1, 8, 0, 8] if top_left_to_right == want_top else cells[1, 8, 3, 8], np.array(img[0, 0], dtype=np.float32) / np.float32(255.0), atol=1e-6)
