"""Seeded, deterministic synthetic stand-ins for the reference's benchmark scenes.

The reference repository ships no geometry (*.bim, textures and env maps are git-ignored,
.gitignore:57-65), only the scene scripts (scenes/*.txt: camera, point lights, resolution) and the
material tables (scenes/*/*.json).  Each generator below rebuilds a scene of the named *shape* --
same materials / lights / camera as the script, procedural geometry of the stated triangle count --
and hands it to the host layer (Scene.from_mesh), which builds the bim hierarchy.  SURVEY.md 8d.

This module only produces INPUT data (numpy); no light transport happens here.
"""
import os

import numpy as np

from . import Scene

# --------------------------------------------------------------------------------------------------
# mesh assembly
# --------------------------------------------------------------------------------------------------


class MeshBuilder:
    def __init__(self):
        self.pos, self.nrm, self.uv, self.tri, self.mat = [], [], [], [], []
        self.nv = 0

    def add(self, pos, nrm, uv, tri, material):
        pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
        nrm = np.asarray(nrm, dtype=np.float32).reshape(-1, 3)
        uv = np.asarray(uv, dtype=np.float32).reshape(-1, 2)
        tri = np.asarray(tri, dtype=np.uint32).reshape(-1, 3)
        self.pos.append(pos); self.nrm.append(nrm); self.uv.append(uv)
        self.tri.append(tri + np.uint32(self.nv))
        self.mat.append(np.full(tri.shape[0], material, dtype=np.uint32))
        self.nv += pos.shape[0]

    def num_triangles(self):
        return int(sum(t.shape[0] for t in self.tri))

    def grid(self, origin, eu, ev, nu, nv, material, uv_scale=(1.0, 1.0), displace=None, flip=False):
        """(nu x nv) quads spanning origin + s*eu + t*ev, s,t in [0,1]; normal = normalize(eu x ev) (flipped if flip)."""
        origin = np.asarray(origin, dtype=np.float64); eu = np.asarray(eu, dtype=np.float64); ev = np.asarray(ev, dtype=np.float64)
        s, t = np.meshgrid(np.linspace(0.0, 1.0, nu + 1), np.linspace(0.0, 1.0, nv + 1), indexing="xy")
        p = origin[None, None, :] + s[..., None] * eu[None, None, :] + t[..., None] * ev[None, None, :]
        n = np.cross(eu, ev)
        n = n / np.linalg.norm(n)
        if flip:
            n = -n
        normals = np.broadcast_to(n, p.shape).copy()
        if displace is not None:
            d, normals = displace(s, t, p, n)
            p = p + d
        uv = np.stack([s * uv_scale[0], t * uv_scale[1]], axis=-1)
        idx = np.arange((nu + 1) * (nv + 1)).reshape(nv + 1, nu + 1)
        a, b, c, d_ = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
        if flip:
            tris = np.concatenate([np.stack([a, c, b], 1), np.stack([a, d_, c], 1)], 0)
        else:
            tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d_], 1)], 0)
        self.add(p.reshape(-1, 3), normals.reshape(-1, 3), uv.reshape(-1, 2), tris, material)

    def box(self, lo, hi, material, tess=1, uv_scale=(1.0, 1.0), inward=False, skip_faces=()):
        """Axis-aligned box; outward normals (inward=True: a room seen from inside).
        skip_faces: indices into (-x,+x,-y,+y,-z,+z) to leave out (e.g. a bottom face resting on a floor)."""
        lo = np.asarray(lo, dtype=np.float64); hi = np.asarray(hi, dtype=np.float64)
        dx, dy, dz = np.array([hi[0] - lo[0], 0, 0.0]), np.array([0, hi[1] - lo[1], 0.0]), np.array([0, 0, hi[2] - lo[2]])
        faces = [
            (lo, dz, dy),                     # -x
            (lo + dx, dy, dz),                # +x
            (lo, dx, dz),                     # -y
            (lo + dy, dz, dx),                # +y
            (lo, dy, dx),                     # -z
            (lo + dz, dx, dy),                # +z
        ]
        mats = material if isinstance(material, (list, tuple)) else [material] * 6
        for k, ((o, eu, ev), m) in enumerate(zip(faces, mats)):
            if k in skip_faces:
                continue
            self.grid(o, eu, ev, tess, tess, m, uv_scale, flip=inward)

    def cylinder(self, base, radius, height, material, segments=24, rings=8, uv_scale=(1.0, 1.0), axis=1):
        base = np.asarray(base, dtype=np.float64)
        th = np.linspace(0.0, 2.0 * np.pi, segments + 1)
        hh = np.linspace(0.0, 1.0, rings + 1)
        T, H = np.meshgrid(th, hh, indexing="xy")
        ring = np.stack([np.cos(T), np.zeros_like(T), np.sin(T)], -1)
        up = np.zeros(3); up[1] = 1.0
        p = ring * radius + H[..., None] * height * up
        n = ring.copy()
        if axis != 1:
            perm = [1, 0, 2] if axis == 0 else [0, 2, 1]
            p = p[..., perm]; n = n[..., perm]
        p = p + base
        uv = np.stack([T / (2 * np.pi) * uv_scale[0], H * uv_scale[1]], -1)
        idx = np.arange((segments + 1) * (rings + 1)).reshape(rings + 1, segments + 1)
        a, b, c, d_ = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
        tris = np.concatenate([np.stack([a, c, b], 1), np.stack([a, d_, c], 1)], 0)
        self.add(p.reshape(-1, 3), n.reshape(-1, 3), uv.reshape(-1, 2), tris, material)

    def sphere(self, center, radius, material, segments=24, rings=12, uv_scale=(1.0, 1.0), squash=(1.0, 1.0, 1.0)):
        center = np.asarray(center, dtype=np.float64)
        th = np.linspace(0.0, 2.0 * np.pi, segments + 1)
        ph = np.linspace(1e-3, np.pi - 1e-3, rings + 1)
        T, Ph = np.meshgrid(th, ph, indexing="xy")
        n = np.stack([np.sin(Ph) * np.cos(T), np.cos(Ph), np.sin(Ph) * np.sin(T)], -1)
        p = n * radius * np.asarray(squash) + center
        nn = n / np.asarray(squash)
        nn = nn / np.linalg.norm(nn, axis=-1, keepdims=True)
        uv = np.stack([T / (2 * np.pi) * uv_scale[0], Ph / np.pi * uv_scale[1]], -1)
        idx = np.arange((segments + 1) * (rings + 1)).reshape(rings + 1, segments + 1)
        a, b, c, d_ = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
        tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d_], 1)], 0)
        self.add(p.reshape(-1, 3), nn.reshape(-1, 3), uv.reshape(-1, 2), tris, material)

    def arrays(self):
        return (np.concatenate(self.pos), np.concatenate(self.nrm), np.concatenate(self.uv), np.concatenate(self.tri), np.concatenate(self.mat))


def legacy_material(albedo=(0.5, 0.5, 0.5), specular=(0.0, 0.0, 0.0), roughness=1.0, opacity=1.0, emissivity=(0.0, 0.0, 0.0), refraction_index=None,
                    albedo_texture=None):
    """One entry of the JSON 'materials' object with the keys Scene::LoadMaterial reads (scene.cpp:138-146)."""
    m = {"type": "legacy", "specularColor": list(map(float, specular)), "roughness": [float(roughness)], "opacity": [float(opacity)] * 3,
         "emissivity": list(map(float, emissivity))}
    if albedo_texture is not None:
        m["albedo"] = albedo_texture
    else:
        m["albedo"] = list(map(float, albedo))
    if refraction_index is not None:
        m["refractionIndex"] = [float(refraction_index)]
    return m


# --------------------------------------------------------------------------------------------------
# procedural textures (binary PPM, read by the host Scene as SRGB8_ALPHA8)
# --------------------------------------------------------------------------------------------------


def _write_ppm(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb, dtype=np.uint8).tobytes())


def procedural_texture(kind, size, seed):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / np.float32(size)
    base = rng.uniform(0.25, 0.85, size=3).astype(np.float32)
    alt = np.clip(base * rng.uniform(0.3, 0.8) + rng.uniform(-0.1, 0.1, size=3), 0.02, 0.95).astype(np.float32)
    if kind == 0:    # bricks
        rows = 16
        yy = y * rows
        xx = x * rows / 2 + 0.5 * (np.floor(yy) % 2)
        mortar = ((yy % 1.0) < 0.08) | ((xx % 1.0) < 0.04)
        t = mortar.astype(np.float32)
    elif kind == 1:  # checker
        n = 8 + (seed % 5) * 4
        t = ((np.floor(x * n) + np.floor(y * n)) % 2).astype(np.float32)
    elif kind == 2:  # stripes
        n = 6 + (seed % 7) * 3
        t = (0.5 + 0.5 * np.sin(2 * np.pi * n * (x + 0.3 * y))).astype(np.float32)
    elif kind == 3:  # value noise (wraps)
        t = np.zeros_like(x)
        amp = 0.5
        for octave in range(5):
            n = 4 << octave
            g = rng.rand(n, n).astype(np.float32)
            fx, fy = x * n, y * n
            ix, iy = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
            ax, ay = fx - ix, fy - iy
            g00 = g[iy % n, ix % n]; g10 = g[iy % n, (ix + 1) % n]; g01 = g[(iy + 1) % n, ix % n]; g11 = g[(iy + 1) % n, (ix + 1) % n]
            t += amp * ((g00 * (1 - ax) + g10 * ax) * (1 - ay) + (g01 * (1 - ax) + g11 * ax) * ay)
            amp *= 0.5
        t = np.clip(t, 0, 1)
    else:            # tiles with rosettes
        n = 4 + (seed % 4) * 2
        u, v = (x * n) % 1.0 - 0.5, (y * n) % 1.0 - 0.5
        r = np.sqrt(u * u + v * v)
        t = (0.5 + 0.5 * np.cos(18.0 * r + 3.0 * np.arctan2(v, u))).astype(np.float32)
    grain = rng.rand(size, size).astype(np.float32) * 0.08
    rgb = base[None, None, :] * (1 - t[..., None]) + alt[None, None, :] * t[..., None] + grain[..., None] - 0.04
    return np.clip(rgb * 255.0 + 0.5, 0, 255).astype(np.uint8)


def write_textures(directory, count, size, seed=1):
    os.makedirs(directory, exist_ok=True)
    names = []
    for i in range(count):
        name = "tex%02d.ppm" % i
        _write_ppm(os.path.join(directory, name), procedural_texture(i % 5, size, seed * 1000 + i))
        names.append(name)
    return names


# --------------------------------------------------------------------------------------------------
# C1  box1 (scenes/box1.txt, scenes/box1/box1.json)
# --------------------------------------------------------------------------------------------------

BOX1_MATERIAL_NAMES = ["DefaultMaterial", "Glass", "Floor", "RightWall", "Wall", "LeftWall", "White", "LightR", "LightL"]

BOX1_CAMERA = {"pos": (0.0, 1.5, 2.27), "lookat": (0.0, 1.5, 2.0), "fov": 67.0}   # scenes/box1.txt:2-4


def box1_materials():
    """scenes/box1/box1.json:4-86, verbatim values."""
    return {
        "DefaultMaterial": legacy_material((0.6, 0.6, 0.6), (0, 0, 0), 1.0, 1.0),
        "Glass": legacy_material((0.5882, 0.5882, 0.5882), (0.035, 0.035, 0.035), 0.01, 0.0),
        "Floor": legacy_material((0.5882, 0.5882, 0.5882), (0, 0, 0), 0.9, 1.0),
        "RightWall": legacy_material((0.02, 0.2863, 0.02), (0, 0, 0), 0.9, 1.0),
        "Wall": legacy_material((0.5882, 0.5882, 0.5882), (0, 0, 0), 0.9, 1.0),
        "LeftWall": legacy_material((0.02, 0.02, 0.3373), (0, 0, 0), 0.9, 1.0),
        "White": legacy_material((1, 1, 1), (0, 0, 0), 0.9, 1.0),
        "LightR": legacy_material((0.5882, 0.5882, 0.5882), (0, 0, 0), 1.0, 1.0, (803.9, 192.2, 19.6)),
        "LightL": legacy_material((0.5882, 0.5882, 0.5882), (0, 0, 0), 1.0, 1.0, (98.0, 392.2, 788.2)),
    }


def box1_mesh(tess=4, coincident_faces=False):
    """The box1 mesh as arrays (positions, normals, texcoords, triangles, triangle materials)."""
    M = {n: i for i, n in enumerate(BOX1_MATERIAL_NAMES)}
    mb = MeshBuilder()
    lo, hi = np.array([-1.5, 0.0, -3.0]), np.array([1.5, 3.0, 2.3])
    # room, seen from inside: faces -x,+x,-y,+y,-z,+z
    mb.box(lo, hi, [M["LeftWall"], M["RightWall"], M["Floor"], M["Wall"], M["Wall"], M["Wall"]], tess=tess, inward=True)
    skip = () if coincident_faces else (2,)
    mb.box([-1.0, 0.0, -2.2], [-0.1, 1.7, -1.3], M["White"], tess=max(1, tess // 2), skip_faces=skip)
    mb.box([0.25, 0.0, -1.3], [1.05, 0.8, -0.5], M["Glass"], tess=max(1, tess // 2), skip_faces=skip)
    # emissive quads just below the ceiling, facing down (normal -y): eu x ev must point to -y
    mb.grid([0.35, 2.98, -1.9], [0.6, 0, 0], [0, 0, 0.6], 2, 2, M["LightR"])
    mb.grid([-0.95, 2.98, -1.9], [0.6, 0, 0], [0, 0, 0.6], 2, 2, M["LightL"])
    return mb.arrays()


def make_box1(tess=4, build_method="sah", coincident_faces=False, bvh_type=0):
    """Closed room x in [-1.5,1.5], y in [0,3], z in [-3,2.3] + a white box, a glass box and two emissive quads.
    coincident_faces=True keeps the boxes' bottom faces, which lie exactly in the floor plane: a
    deliberate exact-tie case for the traversal parity tests."""
    pos, nrm, uv, tri, mat = box1_mesh(tess, coincident_faces)
    return Scene.from_mesh(pos, nrm, uv, tri, mat, box1_materials(), "", build_method, bvh_type=bvh_type)


# --------------------------------------------------------------------------------------------------
# C2  Sponza-shaped atrium (scenes/cryteksponza.txt)
# --------------------------------------------------------------------------------------------------

SPONZA_CAMERA = {"pos": (784.069214, 525.171387, -123.871895), "lookat": (-784.925598, -525.603210, 124.154999), "fov": 70.0}  # cryteksponza.txt:2-4
SPONZA_POINT_LIGHTS = [((-382.0, 160.0, -180.0), (50000.0, 50000.0, 50000.0)), ((-382.0, 160.0, 150.0), (50000.0, 50000.0, 50000.0)),
                       ((487.4, 160.0, -180.0), (50000.0, 50000.0, 50000.0)), ((487.4, 160.0, 150.0), (50000.0, 50000.0, 50000.0))]   # :8-11


def make_sponza_like(texture_dir, target_triangles=262144, texture_size=1024, num_materials=25, seed=7, build_method="sah", detail=1.0):
    """Two-storey colonnaded atrium, extent ~3700 x 1500 x 2300, `num_materials` textured materials.
    detail < 1 coarsens every tessellation (small test scenes); target_triangles=None skips the fill-up."""

    def T(n):
        return max(1, int(round(n * detail)))
    rng = np.random.RandomState(seed)
    tex = write_textures(texture_dir, num_materials, texture_size, seed)
    materials = {}
    for i in range(num_materials):
        rough = [1.0, 0.9, 0.7, 0.5, 0.3][i % 5]
        spec = 0.0 if i % 3 else 0.04
        materials["mat%02d" % i] = legacy_material(specular=(spec, spec, spec), roughness=rough, albedo_texture=tex[i])
    mb = MeshBuilder()
    X, Y, Z = 1850.0, 1500.0, 1150.0

    def m(i):
        return i % num_materials

    # shell: floor, roof, four walls (finely tessellated so the BVH has work everywhere)
    mb.grid([-X, 0, -Z], [0, 0, 2 * Z], [2 * X, 0, 0], T(48), T(64), m(0), (12, 20))                 # floor, normal +y
    mb.grid([-X, Y, -Z], [2 * X, 0, 0], [0, 0, 2 * Z], T(32), T(24), m(1), (10, 6))                  # roof, normal -y
    mb.grid([-X, 0, -Z], [2 * X, 0, 0], [0, Y, 0], T(64), T(24), m(2), (16, 6))                      # z = -Z, normal +z
    mb.grid([-X, 0, Z], [0, Y, 0], [2 * X, 0, 0], T(24), T(64), m(3), (6, 16))                       # z = +Z, normal -z
    mb.grid([-X, 0, -Z], [0, Y, 0], [0, 0, 2 * Z], T(24), T(40), m(4), (6, 10))                      # x = -X, normal +x
    mb.grid([X, 0, -Z], [0, 0, 2 * Z], [0, Y, 0], T(40), T(24), m(5), (10, 6))                       # x = +X, normal -x
    # galleries: slabs along both long sides on the upper floor, with a parapet
    for side in (-1.0, 1.0):
        z0, z1 = (side * 420.0, side * Z) if side > 0 else (side * Z, side * 420.0)
        mb.box([-X + 1, 480.0, z0], [X - 1, 520.0, z1], m(6), tess=T(12), uv_scale=(12, 3))
        zp = side * 420.0
        mb.box([-X + 1, 520.0, zp - 8], [X - 1, 600.0, zp + 8], m(7), tess=T(10), uv_scale=(20, 1))
    # colonnades: two rows x two storeys of columns with capitals and bases
    col_x = np.linspace(-1600.0, 1600.0, 14)
    k = 0
    for side in (-1.0, 1.0):
        for x in col_x:
            for (y0, hgt, rad) in ((0.0, 480.0, 42.0), (520.0, 520.0, 34.0)):
                mb.cylinder([x, y0 + 30, side * 420.0], rad, hgt - 60, m(8 + k % 3), segments=max(6, T(28)), rings=T(10), uv_scale=(2, 4))
                mb.box([x - rad * 1.4, y0, side * 420.0 - rad * 1.4], [x + rad * 1.4, y0 + 30, side * 420.0 + rad * 1.4], m(11), tess=T(2))
                mb.box([x - rad * 1.4, y0 + hgt - 30, side * 420.0 - rad * 1.4], [x + rad * 1.4, y0 + hgt, side * 420.0 + rad * 1.4], m(11), tess=T(2))
                k += 1
    # arches between columns approximated by lintels
    for side in (-1.0, 1.0):
        for i in range(len(col_x) - 1):
            xa, xb = col_x[i] + 50, col_x[i + 1] - 50
            mb.box([xa, 400.0, side * 420.0 - 25], [xb, 480.0, side * 420.0 + 25], m(12), tess=T(3), uv_scale=(3, 1))
            mb.box([xa, 960.0, side * 420.0 - 20], [xb, 1040.0, side * 420.0 + 20], m(12), tess=T(3), uv_scale=(3, 1))
    # hanging curtains: wavy, finely tessellated sheets between upper columns
    def wave(amp, freq, phase):
        def f(s, t, p, n):
            d = amp * np.sin(2 * np.pi * freq * s + phase) * (0.3 + 0.7 * t)
            dd = amp * 2 * np.pi * freq * np.cos(2 * np.pi * freq * s + phase) * (0.3 + 0.7 * t)
            disp = d[..., None] * n[None, None, :]
            # approximate normal: n - dd * tangent_u (unit) / |eu|
            nn = np.broadcast_to(n, p.shape).copy()
            nn[..., 0] -= dd / 300.0
            nn /= np.linalg.norm(nn, axis=-1, keepdims=True)
            return disp, nn
        return f
    for i in range(0, len(col_x) - 1, 2):
        for side in (-1.0, 1.0):
            xa, xb = col_x[i] + 40, col_x[i + 1] - 40
            mb.grid([xa, 560.0, side * 400.0], [xb - xa, 0, 0], [0, 380.0, 0], T(40), T(28), m(13 + (i // 2) % 4), (2, 2), displace=wave(14.0, 3.0, 0.7 * i), flip=side < 0)
    # props on the floor: vases (squashed spheres) and plinths
    for i in range(24):
        x = rng.uniform(-1500, 1500); z = rng.uniform(-330, 330)
        mb.box([x - 40, 0, z - 40], [x + 40, 60, z + 40], m(17), tess=T(2), skip_faces=(2,))
        mb.sphere([x, 60 + 55, z], 55.0, m(18 + i % 4), segments=max(6, T(28)), rings=max(3, T(16)), squash=(1.0, 1.3, 1.0))
    # a central fountain basin
    mb.cylinder([0, 0, 0], 260.0, 70.0, m(22), segments=max(8, T(64)), rings=T(4), uv_scale=(8, 1))
    mb.sphere([0, 150, 0], 90.0, m(23), segments=max(8, T(40)), rings=max(4, T(24)))
    # fill up to the exact triangle target with a relief frieze along the end walls
    remaining = 0 if target_triangles is None else target_triangles - mb.num_triangles()
    if remaining < 0:
        raise ValueError("sponza-like base geometry already has %d triangles" % mb.num_triangles())
    def relief(s, t, p, n):
        d = 6.0 * np.sin(40 * np.pi * s) * np.sin(6 * np.pi * t)
        disp = d[..., None] * n[None, None, :]
        return disp, np.broadcast_to(n, p.shape).copy()
    quads = remaining // 2
    if quads > 0:
        rows = 64
        cols = quads // rows
        if cols > 0:
            mb.grid([-X + 2, 600.0, -Z + 4], [0, 500.0, 0], [0, 0, 2 * Z - 8], rows, cols, m(24), (2, 12), displace=relief)
        rest = quads - rows * cols
        if rest > 0:
            mb.grid([X - 2, 600.0, -Z + 4], [0, 0, 2 * Z - 8], [0, 500.0, 0], rest, 1, m(24), (12, 2))
    if target_triangles is not None and (target_triangles - mb.num_triangles()) == 1:  # odd target: one more small triangle
        mb.add([[0, 1400, 0], [10, 1400, 0], [0, 1400, 10]], [[0, -1, 0]] * 3, [[0, 0], [1, 0], [0, 1]], [[0, 1, 2]], m(24))
    pos, nrm, uv, tri, mat = mb.arrays()
    scene = Scene.from_mesh(pos, nrm, uv, tri, mat, materials, texture_dir, build_method)
    for p, i in SPONZA_POINT_LIGHTS:
        scene.add_point_light(p, i)
    return scene


# --------------------------------------------------------------------------------------------------
# small analytic scenes for tests
# --------------------------------------------------------------------------------------------------


def make_random_triangles(n, seed=0, extent=10.0, size=1.0, build_method="sah", bvh_type=0, offset=(0.0, 0.0, 0.0)):
    """n random triangles in a cube around `offset`: the adversarial case for traversal parity (overlaps, slivers)."""
    rng = np.random.RandomState(seed)
    c = rng.uniform(-extent, extent, size=(n, 1, 3)) + np.asarray(offset, dtype=np.float64)
    p = (c + rng.uniform(-size, size, size=(n, 3, 3))).astype(np.float32).reshape(-1, 3)
    nrm = np.tile(np.array([0, 1, 0], dtype=np.float32), (3 * n, 1))
    uv = rng.rand(3 * n, 2).astype(np.float32)
    tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mat = (np.arange(n) % 2).astype(np.uint32)
    mats = {"a": legacy_material((0.7, 0.7, 0.7)), "b": legacy_material((0.3, 0.5, 0.7), (0.04, 0.04, 0.04), 0.3)}
    return Scene.from_mesh(p, nrm, uv, tri, mat, mats, "", build_method, bvh_type=bvh_type)


def make_furnace(albedo=0.5, emission=1.0, tess=2):
    """Closed box whose walls all emit `emission` with albedo `albedo`: radiance = emission/(1-albedo) everywhere."""
    mb = MeshBuilder()
    mb.box([-1, -1, -1], [1, 1, 1], 0, tess=tess, inward=True)
    pos, nrm, uv, tri, mat = mb.arrays()
    mats = {"wall": legacy_material((albedo,) * 3, (0, 0, 0), 1.0, 1.0, (emission,) * 3)}
    return Scene.from_mesh(pos, nrm, uv, tri, mat, mats, "")


# --------------------------------------------------------------------------------------------------
# C3 / C4 / C5 stand-ins: scenes of the named SHAPE for the other BASELINE configs.  Cameras, point lights and
# the quoted material values are the reference's script / material-table data (input data, not code).
# --------------------------------------------------------------------------------------------------

BATHROOM_CAMERA = {"pos": (1.315, 1.183, -0.5367), "lookat": (0.389973, 1.086791, -0.904588), "fov": 50.0}           # scenes/bathroom.txt:2-4
VILLA_CAMERA = {"pos": (11.901048, 3.936197, 6.616631), "lookat": (12.886539, 3.829603, 6.748702), "fov": 47.0}       # scenes/villa-night.txt:4-6
VILLA_POINT_LIGHTS = [((23.5658, 4.35, 2.46295), (3.0, 3.0, 3.0)), ((21.3179, 4.35, 2.46144), (3.0, 3.0, 3.0)), ((16.8016, 4.18727, 8.15809), (8.0, 8.0, 8.0)),
                      ((21.2769, 3.96, 7.20042), (1.0, 1.0, 1.0)), ((22.1166, 3.96, 7.18543), (1.0, 1.0, 1.0)), ((23.0586, 3.96, 7.1867), (1.0, 1.0, 1.0))]  # villa-night.txt:10-17
VILLAGE_CAMERA = {"pos": (99.868530, 41.301537, -84.863884), "lookat": (99.381393, 40.570442, -84.386185), "fov": 67.0}  # scenes/medival.txt:2-4
# scenes/medival.txt:9-56 (x, y, z, r, g, b)
VILLAGE_POINT_LIGHTS = [
    (50.961807, 52.331261, -48.856907, 56.0, 60.0, 74.0),
    (93.699875, 9.947262, -68.010811, 9.6, 9.0, 8.0),
    (93.893402, 7.766943, -68.173920, 9.6, 9.0, 8.0),
    (90.212570, 8.170808, -64.181938, 9.6, 9.0, 8.0),
    (87.710732, 7.788787, -60.950520, 9.6, 9.0, 8.0),
    (96.882660, 10.159486, -49.179569, 4.8, 4.5, 4.0),
    (93.794876, 10.159486, -49.886395, 4.8, 4.5, 4.0),
    (90.595482, 10.159486, -50.397888, 4.8, 4.5, 4.0),
    (87.744881, 10.159486, -50.774151, 4.8, 4.5, 4.0),
    (84.929131, 10.159486, -51.592556, 4.8, 4.5, 4.0),
    (76.179848, 7.677871, -72.380875, 9.6, 9.0, 8.0),
    (72.044197, 7.767259, -65.459724, 9.6, 9.0, 8.0),
    (70.393097, 7.530910, -61.896576, 9.6, 9.0, 8.0),
    (67.061455, 7.728245, -63.888161, 9.6, 9.0, 8.0),
    (67.829430, 8.177100, -79.097054, 9.6, 9.0, 8.0),
    (58.257168, 8.014050, -82.186745, 9.6, 9.0, 8.0),
    (57.641644, 8.066957, -71.122726, 9.6, 9.0, 8.0),
    (59.883175, 8.035384, -69.444321, 9.6, 9.0, 8.0),
    (35.047054, 8.077431, -79.554718, 9.6, 9.0, 8.0),
    (41.835251, 10.939795, -77.081375, 9.6, 9.0, 8.0),
    (41.836338, 8.048097, -77.093620, 9.6, 9.0, 8.0),
    (46.435116, 8.234673, -75.694122, 9.6, 9.0, 8.0),
    (46.435116, 10.939795, -75.694122, 9.6, 9.0, 8.0),
    (51.466606, 8.112941, -74.541481, 9.6, 9.0, 8.0),
    (51.466606, 10.939795, -74.541481, 9.6, 9.0, 8.0),
    (40.675018, 7.732394, -65.276169, 9.6, 9.0, 8.0),
    (44.817814, 7.609001, -60.265152, 9.6, 9.0, 8.0),
    (46.405258, 7.907426, -56.999210, 9.6, 9.0, 8.0),
    (51.176723, 10.372055, -51.016087, 9.6, 9.0, 8.0),
    (51.176723, 8.112055, -51.016087, 9.6, 9.0, 8.0),
    (56.440342, 7.825526, -46.059772, 9.6, 9.0, 8.0),
    (60.088978, 8.039399, -40.409771, 9.6, 9.0, 8.0),
    (63.068680, 8.044291, -36.820271, 9.6, 9.0, 8.0),
    (65.726524, 10.416846, -30.324221, 9.6, 9.0, 8.0),
    (65.726524, 7.5116846, -30.324221, 9.6, 9.0, 8.0),
    (83.404610, 10.873559, -42.846710, 9.6, 9.0, 8.0),
    (83.404610, 7.873559, -42.846710, 9.6, 9.0, 8.0),
    (83.043175, 10.922251, -38.061066, 9.6, 9.0, 8.0),
    (83.043175, 7.922251, -38.061066, 9.6, 9.0, 8.0),
    (82.306511, 8.227152, -31.170267, 9.6, 9.0, 8.0),
    (80.517807, 8.11138, -25.950634, 9.6, 9.0, 8.0),
    (79.047409, 8.272710, -20.515154, 9.6, 9.0, 8.0),
    (88.518288, 10.999722, -19.305355, 9.6, 9.0, 8.0),
    (88.518288, 7.999722, -19.305355, 9.6, 9.0, 8.0),
    (96.576538, 8.258466, -24.095053, 9.6, 9.0, 8.0),
    (98.233070, 8.421268, -33.393890, 9.6, 9.0, 8.0),
    (99.365128, 8.412631, -37.839222, 9.6, 9.0, 8.0),
]


def _relief(amp, fu, fv):
    def f(s, t, p, n):
        d = amp * np.sin(2 * np.pi * fu * s) * np.sin(2 * np.pi * fv * t)
        return d[..., None] * n[None, None, :], np.broadcast_to(n, p.shape).copy()
    return f


def _fill_to_target(mb, target_triangles, origin, eu, ev, material, displace):
    """Tops the mesh up to exactly `target_triangles` with a finely tessellated relief panel."""
    if target_triangles is None:
        return
    remaining = target_triangles - mb.num_triangles()
    if remaining < 0:
        raise ValueError("base geometry already has %d triangles" % mb.num_triangles())
    quads = remaining // 2
    rows = 256 if quads >= 65536 else 16
    cols = quads // rows
    if cols > 0:
        mb.grid(origin, eu, ev, cols, rows, material, (8, 8), displace=displace)
    rest = quads - rows * cols
    if rest > 0:
        eu64, ev64 = np.asarray(eu, dtype=np.float64), np.asarray(ev, dtype=np.float64)
        nn = np.cross(eu64, ev64)
        o2 = np.asarray(origin, dtype=np.float64) + 1.0e-3 * nn / np.linalg.norm(nn)
        mb.grid(o2, eu, ev, rest, 1, material, (8, 1))
    if target_triangles - mb.num_triangles() == 1:
        o = np.asarray(origin, dtype=np.float64)
        mb.add([o, o + 0.01 * np.asarray(eu), o + 0.01 * np.asarray(ev)], [[0, 1, 0]] * 3, [[0, 0], [1, 0], [0, 1]], [[0, 1, 2]], material)


def make_bathroom_like(texture_dir, target_triangles=500000, texture_size=512, seed=11, build_method="sah", detail=1.0):
    """C3: closed interior around the camera of scenes/bathroom.txt, ~0.5 M triangles, two emissive materials
    (bathroom.json 'light' 13/13/13 and 'low_light' 64/18.1951/1.8709), glossy tub, glass globe, water, a mirror-like
    drawing (roughness 6.77e-6), 4 textured materials.  Lit by area lights only (BPT's home turf)."""

    def T(n):
        return max(2, int(round(n * detail)))
    tex = write_textures(texture_dir, 4, texture_size, seed)
    materials = {
        "DefaultMaterial": legacy_material((0.6, 0.6, 0.6)),
        "bathtube": legacy_material((0.64, 0.64, 0.64), (0.3, 0.3, 0.3), 0.001),
        "carpet": legacy_material(specular=(0.1, 0.1, 0.1), roughness=0.9, albedo_texture=tex[0]),
        "ceilling_white": legacy_material((0.8, 0.8, 0.8)),
        "drawing": legacy_material(specular=(0.0, 0.0, 0.0), roughness=6.77062e-06, albedo_texture=tex[1]),
        "floor": legacy_material(specular=(0.05, 0.05, 0.05), roughness=0.3, albedo_texture=tex[2]),
        "wood": legacy_material(specular=(0.04, 0.04, 0.04), roughness=0.5, albedo_texture=tex[3]),
        "glass_globe": legacy_material((1.0, 1.0, 1.0), (0.04, 0.04, 0.04), 0.0001, 0.0, refraction_index=1.5),
        "water": legacy_material((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), 1e-6, 0.0, refraction_index=1.33),
        "light": legacy_material((0.6, 0.6, 0.6), emissivity=(13.0, 13.0, 13.0)),
        "low_light": legacy_material((0.6, 0.6, 0.6), emissivity=(64.0, 18.1951, 1.8709)),
        "wall": legacy_material((0.75, 0.72, 0.68), roughness=0.9),
    }
    M = {n: i for i, n in enumerate(materials)}
    mb = MeshBuilder()
    lo, hi = np.array([-2.2, 0.0, -3.0]), np.array([2.2, 2.6, 1.2])
    mb.box(lo, hi, [M["wall"], M["wall"], M["floor"], M["ceilling_white"], M["wall"], M["wall"]], tess=T(48), inward=True, uv_scale=(4, 4))
    mb.grid([-1.0, 0.002, -2.4], [0, 0, 1.6], [1.8, 0, 0], T(64), T(64), M["carpet"], (3, 3))
    # tub: glossy squashed sphere half sunk into a plinth, water surface inside
    mb.box([-2.0, 0.0, -2.9], [-0.4, 0.45, -1.9], M["bathtube"], tess=T(16), skip_faces=(2,))
    mb.sphere([-1.2, 0.55, -2.4], 0.45, M["bathtube"], segments=T(96), rings=T(48), squash=(1.6, 0.5, 1.0))
    mb.grid([-1.8, 0.62, -2.75], [0, 0, 0.7], [1.2, 0, 0], T(48), T(48), M["water"], displace=_relief(0.004, 6, 5))
    # glass globe lamp + its low light, mirror-like drawing on the wall, wooden shelf + props
    mb.sphere([0.6, 1.9, -1.6], 0.22, M["glass_globe"], segments=T(96), rings=T(48))
    mb.sphere([0.6, 1.9, -1.6], 0.05, M["low_light"], segments=max(6, T(24)), rings=max(3, T(12)))
    mb.grid([-0.8, 0.9, -2.995], [1.4, 0, 0], [0, 1.0, 0], T(32), T(32), M["drawing"])
    mb.box([1.2, 0.0, -2.9], [2.1, 0.9, -2.3], M["wood"], tess=T(24), skip_faces=(2,), uv_scale=(2, 2))
    rng = np.random.RandomState(seed)
    for i in range(12):
        x, z = rng.uniform(1.25, 2.05), rng.uniform(-2.85, -2.35)
        mb.sphere([x, 0.9 + 0.06, z], 0.06, M["bathtube" if i % 2 else "wood"], segments=max(6, T(32)), rings=max(3, T(16)))
    # ceiling panels (emissive quads facing down)
    mb.grid([-0.9, 2.59, -2.0], [0.8, 0, 0], [0, 0, 0.8], 4, 4, M["light"])
    mb.grid([0.6, 2.59, -0.6], [0.6, 0, 0], [0, 0, 0.6], 4, 4, M["light"])
    _fill_to_target(mb, target_triangles, [2.195, 0.3, 1.1], [0, 0, -3.8], [0, 2.0, 0], M["wall"], _relief(0.004, 60, 40))
    pos, nrm, uv, tri, mat = mb.arrays()
    return Scene.from_mesh(pos, nrm, uv, tri, mat, materials, texture_dir, build_method)


def make_villa_like(texture_dir, target_triangles=1000000, texture_size=512, seed=13, build_method="sah", detail=1.0):
    """C4: night-time villa around the camera of scenes/villa-night.txt: terrace + pool (water: opacity 0, roughness 1e-6, villa.json:6-11),
    window panes (glass: specular 0.14, roughness 1e-4, :46-51), textured boards / walls, ~1 M triangles, the 6 point lights of villa-night.txt:10-17."""

    def T(n):
        return max(2, int(round(n * detail)))
    tex = write_textures(texture_dir, 8, texture_size, seed)
    materials = {
        "pool water": legacy_material((0.0, 0.0, 0.0), (0.05, 0.05, 0.05), 1e-6, 0.0, refraction_index=1.33),
        "steel": legacy_material((0.3, 0.3, 0.3), (0.6, 0.6, 0.6), 0.05),
        "windows/glass": legacy_material((1.0, 1.0, 1.0), (0.14, 0.14, 0.14), 0.0001, 0.0, refraction_index=1.5),
        "plaster": legacy_material((0.85, 0.85, 0.82), roughness=0.9),
    }
    for i in range(8):
        materials["boards%02d" % i] = legacy_material(specular=(0.04, 0.04, 0.04), roughness=[0.9, 0.6, 0.4, 0.8][i % 4], albedo_texture=tex[i])
    M = {n: i for i, n in enumerate(materials)}
    mb = MeshBuilder()
    # ground + terrace deck, pool basin with a rippled water surface
    mb.grid([0, 0, -6], [0, 0, 24], [36, 0, 0], T(160), T(160), M["boards00"], (30, 30), displace=_relief(0.03, 9, 7))
    mb.box([14.0, 0.0, 9.5], [24.0, 3.2, 10.0], M["boards01"], tess=T(48), uv_scale=(8, 3))
    mb.box([15.0, 0.02, 3.5], [25.0, 0.5, 8.5], M["plaster"], tess=T(40), skip_faces=(2, 3))
    mb.grid([15.2, 0.45, 3.7], [0, 0, 4.6], [9.6, 0, 0], T(200), T(200), M["pool water"], displace=_relief(0.01, 14, 11))
    # house: two storeys, glass front, lamella screens
    mb.box([12.0, 0.0, -4.0], [30.0, 3.3, 2.0], [M["plaster"], M["plaster"], M["boards02"], M["boards03"], M["plaster"], M["plaster"]], tess=T(64), uv_scale=(6, 3))
    mb.box([14.0, 3.3, -4.0], [28.0, 6.4, 1.0], M["plaster"], tess=T(48), uv_scale=(6, 3))
    for i in range(10):
        x = 13.0 + 1.6 * i
        mb.grid([x, 0.2, 2.03], [1.4, 0, 0], [0, 2.8, 0], T(24), T(24), M["windows/glass"])
    for i in range(60):
        x = 12.2 + 0.3 * i
        mb.box([x, 3.4, 1.05], [x + 0.05, 6.3, 1.25], M["boards04"], tess=T(6))
    # garden furniture
    rng = np.random.RandomState(seed)
    for i in range(20):
        x, z = rng.uniform(13.0, 26.0), rng.uniform(2.6, 9.0)
        if 15.0 < x < 25.0 and 3.5 < z < 8.5:
            continue
        mb.cylinder([x, 0.0, z], 0.03, 0.45, M["steel"], segments=12, rings=4)
        mb.box([x - 0.3, 0.45, z - 0.3], [x + 0.3, 0.5, z + 0.3], M["boards05"], tess=T(6))
    _fill_to_target(mb, target_triangles, [30.5, 0.0, -6.0], [0, 0, 24], [0, 5.0, 0], M["boards06"], _relief(0.02, 80, 30))
    pos, nrm, uv, tri, mat = mb.arrays()
    scene = Scene.from_mesh(pos, nrm, uv, tri, mat, materials, texture_dir, build_method)
    for p, i in VILLA_POINT_LIGHTS:
        scene.add_point_light(p, i)
    return scene


def make_village_like(target_triangles=4000000, seed=17, build_method="sah", num_lights=48):
    """C5: large outdoor scene around the camera of scenes/medival.txt: rolling terrain + ~400 houses (boxes with gable roofs) and towers,
    ~4 M triangles, deep BVH, the two materials of MedivalVillage.json ('DefaultMaterial' opacity 0 -- unused by geometry here -- and
    'FrontColor' albedo 0.9 roughness 0.6) and the 48 point lights of medival.txt:9-56."""
    materials = {
        "DefaultMaterial": legacy_material((0.6, 0.6, 0.6), roughness=1.0, opacity=0.0),
        "FrontColor": legacy_material((0.9, 0.9, 0.9), roughness=0.6),
    }
    F = 1
    rng = np.random.RandomState(seed)
    mb = MeshBuilder()

    def hills(s, t, p, n):
        d = 6.0 * np.sin(2 * np.pi * 1.5 * s) * np.cos(2 * np.pi * 1.2 * t) + 1.5 * np.sin(2 * np.pi * 9 * s + 1.0) * np.sin(2 * np.pi * 7 * t)
        return d[..., None] * n[None, None, :], np.broadcast_to(n, p.shape).copy()

    scale = max(0.05, min(1.0, (target_triangles or 4000000) / 4000000.0)) ** 0.5
    g = max(8, int(700 * scale))
    mb.grid([0, 0, -160], [0, 0, 200], [200, 0, 0], g, g, F, (40, 40), displace=hills)
    houses = max(4, int(400 * scale * scale))
    for i in range(houses):
        x, z = rng.uniform(20, 180), rng.uniform(-140, 20)
        w, d, h = rng.uniform(4, 9), rng.uniform(4, 9), rng.uniform(3, 7)
        y0 = -8.0
        t = max(2, int(12 * scale))
        mb.box([x, y0, z], [x + w, y0 + 8.0 + h, z + d], F, tess=t, skip_faces=(2,))
        # gable roof: two sloped grids
        r = rng.uniform(1.5, 3.5)
        mb.grid([x - 0.3, y0 + 8.0 + h, z - 0.3], [0, 0, d + 0.6], [w / 2 + 0.3, r, 0], t, t, F)
        mb.grid([x + w + 0.3, y0 + 8.0 + h, z - 0.3], [-(w / 2 + 0.3), r, 0], [0, 0, d + 0.6], t, t, F)
    for i in range(max(2, int(24 * scale))):
        x, z = rng.uniform(30, 170), rng.uniform(-130, 10)
        mb.cylinder([x, -8.0, z], rng.uniform(2.0, 3.5), rng.uniform(20, 40), F, segments=max(8, int(48 * scale)), rings=max(4, int(40 * scale)))
    _fill_to_target(mb, target_triangles, [0.0, -20.0, -160.0], [200.0, 0, 0], [0, 0, 200.0], F, _relief(0.5, 90, 70))
    pos, nrm, uv, tri, mat = mb.arrays()
    scene = Scene.from_mesh(pos, nrm, uv, tri, mat, materials, "", build_method)
    for (x, y, z, r, g_, b) in VILLAGE_POINT_LIGHTS[:num_lights]:
        scene.add_point_light((x, y, z), (r, g_, b))
    return scene
