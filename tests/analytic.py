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


def check_box_faces(rtb, trace):
    """book box(a, b) = six quads whose own normals point outward.  Axis-parallel rays from outside meet the face they are
    aimed at: t = the gap, the point on that face, the normal against the ray, front face; a ray from inside meets the far
    face from behind (normal still reported against the ray, back face); (u, v) run along the face's edge vectors as the book
    lays them out (front face: u along +x from min.x, v along +y from min.y); rays past an edge miss."""
    s = rtb.Scene(); m = s.lambertian(tex=s.image(np.full((4, 4, 3), 128, dtype=np.uint8)))
    a = np.array([-1.0, 2.0, 3.0]); b = np.array([2.0, 4.0, 8.0])
    s.set_root(s.list([s.box(a, b, m)]))
    c = a + (b - a) * np.array([0.25, 0.625, 0.375])
    for axis in range(3):
        e = np.zeros(3); e[axis] = 1.0
        for sign in (+1.0, -1.0):
            o = c.copy(); o[axis] = (b[axis] + 5.0) if sign > 0 else (a[axis] - 5.0)
            h = trace(s, _rays(rtb, o, -sign * e))
            want_p = c.copy(); want_p[axis] = b[axis] if sign > 0 else a[axis]
            assert h["object"][0] >= 0
            np.testing.assert_allclose(h["t"][0], 5.0, rtol=1e-6)
            np.testing.assert_allclose(h["p"][0], want_p, atol=1e-5)
            np.testing.assert_allclose(h["n"][0], sign * e, atol=1e-6)
            assert h["front_face"][0] == 1
            # from the inside the same face is met from behind
            h = trace(s, _rays(rtb, c, sign * e))
            np.testing.assert_allclose(h["t"][0], abs(want_p[axis] - c[axis]), rtol=1e-6)
            np.testing.assert_allclose(h["n"][0], -sign * e, atol=1e-6)
            assert h["front_face"][0] == 0
    # front face (z = max.z): Q = (min.x, min.y, max.z), u = dx, v = dy
    h = trace(s, _rays(rtb, [c[0], c[1], 20.0], [0.0, 0.0, -1.0]))
    np.testing.assert_allclose([h["u"][0], h["v"][0]], [0.25, 0.625], atol=1e-6)
    # just past every edge of the silhouette seen along -z: nothing
    for dx, dy in ((-1e-3, 0.5), (1.0 + 1e-3, 0.5), (0.5, -1e-3), (0.5, 1.0 + 1e-3)):
        o = np.array([a[0] + dx * (b[0] - a[0]), a[1] + dy * (b[1] - a[1]), 20.0])
        assert trace(s, _rays(rtb, o, [0.0, 0.0, -1.0]))["object"][0] == -1


def check_triangle_barycentric(rtb, trace):
    """triangle(Q, u, v): the half of the parallelogram with alpha >= 0, beta >= 0, alpha + beta <= 1 - same plane
    arithmetic as the quad, (u, v) = (alpha, beta)."""
    s = rtb.Scene(); m = s.lambertian(tex=s.image(np.full((4, 4, 3), 128, dtype=np.uint8)))
    Q = np.array([0.5, -1.0, 2.0]); u = np.array([3.0, 0.0, 1.0]); v = np.array([0.5, 2.0, -0.5])
    s.set_root(s.list([s.triangle(Q, u, v, m)]))
    n = np.cross(u, v); n /= np.linalg.norm(n)
    inside = np.array([[0.25, 0.25], [0.01, 0.01], [0.98, 0.01], [0.01, 0.98], [0.5, 0.499], [0.125, 0.75]])
    outside = np.array([[0.5, 0.501], [0.75, 0.75], [0.99, 0.02], [-0.01, 0.5], [0.5, -0.01], [1.01, -0.005]])
    for ab, hit in ((inside, True), (outside, False)):
        pts = Q + ab[:, :1] * u + ab[:, 1:] * v
        o = pts + 4.0 * n + np.array([0.2, 0.1, -0.3])
        h = trace(s, _rays(rtb, o, pts - o))
        if hit:
            assert (h["object"] == 0).all()
            np.testing.assert_allclose(h["t"], 1.0, rtol=2e-6)
            np.testing.assert_allclose(np.stack([h["u"], h["v"]], 1), ab, atol=3e-6)
            np.testing.assert_allclose(h["n"], np.broadcast_to(n, (len(pts), 3)), atol=2e-6)
        else:
            assert (h["object"] == -1).all()


def check_emitter_and_mirror(rtb, render):
    """book diffuse_light: a path that reaches an emitter ends there with throughput x emitted (no scattering, either side);
    metal with fuzz 0 (cu_materials.cuh:77-95) multiplies the throughput by its albedo and reflects.  Under a black
    background, a camera looking straight at a light of radiance E sees exactly E; looking down at a mirror of albedo A
    under a ceiling of that light, exactly A * E (float32 products); and nothing but black where neither is."""
    E = np.array([4.0, 2.0, 1.0], dtype=np.float32); A = np.array([0.8, 0.6, 0.4], dtype=np.float32)
    s = rtb.Scene()
    light = s.diffuse_light(s.solid(tuple(float(x) for x in E)))
    s.set_root(s.list([s.quad((-50.0, 5.0, -50.0), (100.0, 0.0, 0.0), (0.0, 0.0, 100.0), light),          # ceiling
                       s.quad((-1.0, 0.0, -1.0), (2.0, 0.0, 0.0), (0.0, 0.0, 2.0), s.metal(tuple(float(x) for x in A), 0.0))]))
    s.set_background(rtb.BG_CONSTANT, (0.0, 0.0, 0.0))
    W = H = 32
    up = render(s, rtb.make_camera("pinhole", (0.0, 1.0, 0.0), (0.0, 5.0, 0.0), (0.0, 0.0, -1.0), 40.0, 1.0), W, H, 4, 8)
    np.testing.assert_array_equal(up[..., :3], np.broadcast_to(np.float32(4.0) * E, (H, W, 3)))            # 4 samples of exactly E
    down = render(s, rtb.make_camera("pinhole", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, -1.0), 40.0, 1.0), W, H, 4, 8)
    # vfov 40 at distance 1: the view is 0.73 wide - entirely on the 2 x 2 mirror
    want = np.float32(4.0) * (A * E)
    np.testing.assert_allclose(down[..., :3], np.broadcast_to(want, (H, W, 3)), rtol=3e-7)
    away = render(s, rtb.make_camera("pinhole", (60.0, 1.0, 0.0), (100.0, 1.0, 0.0), (0.0, 1.0, 0.0), 20.0, 1.0), W, H, 4, 8)
    assert np.abs(away[..., :3]).max() == 0.0                                                               # past the edge of the ceiling, looking away: black


def check_white_furnace(rtb, render):
    """Energy conservation: under a uniform white sky every material with albedo 1 must return exactly the sky - Lambertian
    on spheres, quads and boxes (inter-reflections included), dielectric (reflect or refract, attenuation 1), a white
    isotropic medium.  Every sample is 1 unless its path runs out of depth (black, Renderer.cu:180), which at depth 50 is
    vanishingly rare for these objects: a convex Lambertian sphere alone gives exactly spp in every pixel."""
    W = H = 48; spp = 16
    one = (1.0, 1.0, 1.0)
    s = rtb.Scene()
    s.set_root(s.list([s.sphere((0.0, 0.0, 0.0), 1.0, s.lambertian(albedo=one))]))
    s.set_background(rtb.BG_CONSTANT, one)
    cam = rtb.make_camera("pinhole", (0.0, 0.5, 4.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, 1.0)
    acc = render(s, cam, W, H, spp, 50)
    np.testing.assert_array_equal(acc[..., :3], np.full((H, W, 3), np.float32(spp)))
    s = rtb.Scene()
    white = s.lambertian(albedo=one)
    ball = s.sphere((1.2, 0.5, 0.0), 0.5, white)
    s.set_root(s.list([s.quad((-3.0, 0.0, -3.0), (6.0, 0.0, 0.0), (0.0, 0.0, 6.0), white),
                       s.sphere((-1.2, 0.5, 0.0), 0.5, s.dielectric(1.5)),
                       s.translate(s.rotate_y(s.box((-0.4, 0.0, -0.4), (0.4, 0.8, 0.4), white), 30.0), (0.0, 0.0, -0.8)),
                       s.constant_medium(ball, 1.0, s.isotropic(s.solid(one)))]))
    s.set_background(rtb.BG_CONSTANT, one)
    cam = rtb.make_camera("pinhole", (0.0, 1.5, 5.0), (0.0, 0.4, 0.0), (0.0, 1.0, 0.0), 40.0, 1.0)
    acc = render(s, cam, W, H, spp, 50)[..., :3]
    assert acc.max() <= spp                                                      # no path ever gains energy
    assert acc.min() >= spp - 1                                                  # at most one depth-out per pixel
    assert acc.mean() >= spp * (1.0 - 2e-4)


def check_beer_lambert(rtb, render):
    """book constant_medium: the free path is -log(u) / density, so a purely absorbing medium (isotropic albedo 0) of
    density s transmits exp(-s L) over a chord of length L - through a slab (box boundary, L = thickness) and through the
    centre of a sphere boundary (L = 2 R)."""
    black_sky_free = (1.0, 1.0, 1.0)
    for kind, dens, L in (("slab", 0.25, 4.0), ("ball", 0.4, 3.0)):
        s = rtb.Scene()
        black = s.isotropic(s.solid((0.0, 0.0, 0.0)))
        shell = s.lambertian(albedo=(1.0, 1.0, 1.0))
        boundary = s.box((-50.0, -50.0, 0.0), (50.0, 50.0, L), shell) if kind == "slab" else s.sphere((0.0, 0.0, 0.0), 0.5 * L, shell)
        s.set_root(s.list([s.constant_medium(boundary, dens, black)]))
        s.set_background(rtb.BG_CONSTANT, black_sky_free)
        # a very narrow view along the axis: every ray crosses (almost exactly) the full chord
        cam = rtb.make_camera("pinhole", (0.0, 0.0, -10.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 0.2, 1.0)
        acc = render(s, cam, 32, 32, 256, 8)
        mean = float((acc[..., 0] / acc[..., 3]).mean())
        want = math.exp(-dens * L)
        sigma = math.sqrt(want * (1.0 - want) / (32 * 32 * 256))
        assert abs(mean - want) < 4.0 * sigma + 1e-4, (kind, mean, want)


def check_perlin_lattice(rtb, render):
    """book noise_texture: 0.5 (1 + sin(scale z + 10 turb(p, 7))) with gradient (Perlin) noise, which vanishes at every
    lattice point - and so does every octave of the turbulence there (2^k p is a lattice point too).  At integer (x, y, z)
    the texture is therefore 0.5 (1 + sin(scale z)) whatever the random tables hold; value noise, a phase taken from another
    axis, or a missing turbulence term all fail this.  White Lambertian quad carrying the texture under a white sky, a view
    four thousandths of a unit wide centred on the lattice point: every sample returns the texel under it."""
    scale = 4.0
    for (x, y, z) in ((3.0, -2.0, 0.0), (-5.0, 7.0, 1.0), (12.0, 4.0, 2.0), (0.0, 0.0, -3.0)):
        s = rtb.Scene()
        s.set_root(s.list([s.quad((x - 1.0, y - 1.0, z), (2.0, 0.0, 0.0), (0.0, 2.0, 0.0), s.lambertian(tex=s.noise(scale)))]))
        s.set_background(rtb.BG_CONSTANT, (1.0, 1.0, 1.0))
        cam = rtb.make_camera("pinhole", (x, y, z + 10.0), (x, y, z), (0.0, 1.0, 0.0), math.degrees(2.0 * math.atan(0.002 / 10.0)), 1.0)
        acc = render(s, cam, 4, 4, 4, 2)
        got = acc[..., :3] / acc[..., 3:4]
        assert np.ptp(got, axis=2).max() < 1e-6                              # grey
        want = 0.5 * (1.0 + math.sin(scale * z))
        # half a view (0.002 units) away from the lattice point each of the 7 octaves is at most ~0.002 |gradient| in: 10 turb < 0.2
        assert np.abs(got[..., 0] - want).max() < 0.08, (x, y, z, got[..., 0], want)
    # away from the lattice the turbulence term is there: half-way between lattice points the value differs from the bare sine
    s = rtb.Scene()
    s.set_root(s.list([s.quad((-8.0, -8.0, 1.0), (16.0, 0.0, 0.0), (0.0, 16.0, 0.0), s.lambertian(tex=s.noise(scale)))]))
    s.set_background(rtb.BG_CONSTANT, (1.0, 1.0, 1.0))
    cam = rtb.make_camera("pinhole", (0.0, 0.0, 11.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), math.degrees(2.0 * math.atan(0.7)), 1.0)
    acc = render(s, cam, 64, 64, 1, 2)
    got = acc[..., 0] / acc[..., 3]
    assert 0.0 <= got.min() and got.max() <= 1.0 and got.std() > 0.1
