"""CPU: the float32 decision path of the fused kernel (csrc/smap_fuse.cuh) restated in numpy with the constants the
library itself computes (smap_debug_fast32, host only), checked against the oracle.

What must hold, by construction of the error bounds (header of smap_fuse.cuh):
  * cull32 never drops a point the reference keeps;
  * a point the float32 path CERTIFIES gets the reference's decision: kept-and-on-grid or not, and if so the
    reference's pixel and cell.
The emulation differs from the hardware only in the reciprocal (correctly rounded here, <= 2 ulp assumed by the
bound) and in rare double roundings of the emulated FMA, both far inside the guard bands being tested.
"""
import ctypes

import numpy as np
import pytest

from oracle import c_oracle
from vision_semantic_segmentation_b200 import _native, synthetic as syn
from vision_semantic_segmentation_b200.camera import camera_setup_1, camera_setup_6
from vision_semantic_segmentation_b200.utils import transforms as tr

F = np.float32
MAGIC = F(12582912.0)
MB = 0x4B400000
OFF = (1369.0496826171875, 562.84814453125)


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def fast32(cfg_kw, T, P, hw):
    lib = _native.load()
    cfg = _native.SmapConfig()
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    fr = _native.SmapFrame()
    fr.image_height, fr.image_width = hw
    fr.has_transform = 0 if T is None else 1
    Tm = np.eye(4) if T is None else np.ascontiguousarray(T, dtype=np.float64)
    ctypes.memmove(fr.world_to_velodyne, Tm.ctypes.data, 128)
    Pm = np.ascontiguousarray(P, dtype=np.float64)
    out = (ctypes.c_double * 69)()
    _native.check(lib.smap_debug_fast32(ctypes.byref(cfg), ctypes.byref(fr), Pm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), out))
    v = np.array(out[:], dtype=np.float64)
    k, i = {}, [0]

    def take(n):
        r = v[i[0]:i[0] + n]
        i[0] += n
        return r
    k["c_dc"] = take(8).reshape(4, 2).astype(F); k["c_ab"] = take(8).reshape(4, 2).astype(F); k["c_wh"] = take(2).astype(F)
    for name in ("c_bw", "c_rh", "c_rthr", "c_depth", "c_lo_u", "c_hi_u", "c_lo_v", "c_hi_v"):
        k[name] = F(take(1)[0])
    k["n_ctr"] = take(3).astype(F); k["coord_l"] = F(take(1)[0])
    k["d_uv"] = take(8).reshape(4, 2).astype(F); k["d_cd"] = take(8).reshape(4, 2).astype(F)
    for name in ("g_k1", "g_k0", "g_hc", "r_h", "r_kd", "r_thr0"):
        k[name] = F(take(1)[0])
    k["mid_uv"] = take(2).astype(F); k["half_uv"] = take(2).astype(F); k["cell_f0"] = take(2).astype(F)
    for name in ("cell_rf", "cell_kc", "cell_hg0"):
        k[name] = F(take(1)[0])
    k["mid_c"] = take(2).astype(F); k["half_c"] = take(2).astype(F); k["clamp_c"] = take(2).astype(F)
    k["pix_k"], k["cell_k"] = int(take(1)[0]), int(take(1)[0])
    return k


def cull32(k, p):
    """smap_fuse.cuh cull32, rows evaluated as the kernel's fused chains"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]

    def chain(m, col):
        return fma(m[0, col], x, fma(m[1, col], y, fma(m[2, col], z, np.full_like(x, m[3, col]))))
    d, c = chain(k["c_dc"], 0), chain(k["c_dc"], 1)
    a, b = chain(k["c_ab"], 0), chain(k["c_ab"], 1)
    lo_u, lo_v = (a + c).astype(F), (b + c).astype(F)
    hi_u, hi_v = fma(k["c_wh"][0], c, -a), fma(k["c_wh"][1], c, -b)
    with np.errstate(invalid="ignore"):
        ok = (lo_u > k["c_lo_u"]) & (hi_u > k["c_hi_u"]) & (lo_v > k["c_lo_v"]) & (hi_v > k["c_hi_v"])
        ok |= ~(c > k["c_depth"])
        ok &= np.abs((d - k["c_rh"]).astype(F)) < k["c_rthr"]
        ok |= np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z)) > k["c_bw"]
    return ok


def decide32(k, p, img_w, map_w):
    """smap_fuse.cuh decide32: returns (certified, inside, pixel index, cell index)"""
    xl, yl, zl = (p[:, 0] + k["n_ctr"][0]).astype(F), (p[:, 1] + k["n_ctr"][1]).astype(F), (p[:, 2] + k["n_ctr"][2]).astype(F)
    rho = ((np.abs(xl) + np.abs(yl)).astype(F) + np.abs(zl)).astype(F)

    def chain(m, col):
        return fma(m[0, col], xl, fma(m[1, col], yl, fma(m[2, col], zl, np.full_like(xl, m[3, col]))))
    qu, qv = chain(k["d_uv"], 0), chain(k["d_uv"], 1)
    q2, d = chain(k["d_cd"], 0), chain(k["d_cd"], 1)
    with np.errstate(all="ignore"):
        rc = (F(1.0) / q2).astype(F)
        hg = fma(-fma(k["g_k1"], rho, np.full_like(rho, k["g_k0"])), np.abs(rc), np.full_like(rho, k["g_hc"]))
        tpx, tpy = fma(qu, rc, np.full_like(rho, MAGIC)), fma(qv, rc, np.full_like(rho, MAGIC))
        rpx, rpy = (tpx - MAGIC).astype(F), (tpy - MAGIC).astype(F)
        dpx, dpy = fma(qu, rc, -rpx), fma(qv, rc, -rpy)
        opx, opy = (rpx - k["mid_uv"][0]).astype(F), (rpy - k["mid_uv"][1]).astype(F)
        tsx, tsy = fma(xl, k["cell_rf"], np.full_like(rho, k["cell_f0"][0])), fma(yl, k["cell_rf"], np.full_like(rho, k["cell_f0"][1]))
        hgc = fma(-k["cell_kc"], rho, np.full_like(rho, k["cell_hg0"]))
        tcx, tcy = (tsx + MAGIC).astype(F), (tsy + MAGIC).astype(F)
        rcx, rcy = (tcx - MAGIC).astype(F), (tcy - MAGIC).astype(F)
        dcx, dcy = (tsx - rcx).astype(F), (tsy - rcy).astype(F)
        ocx, ocy = (rcx - k["mid_c"][0]).astype(F), (rcy - k["mid_c"][1]).astype(F)
        cert = rho < k["coord_l"]
        cert &= np.abs((d - k["r_h"]).astype(F)) < fma(-k["r_kd"], rho, np.full_like(rho, k["r_thr0"]))
        cert &= (np.abs(dpx) < hg) & (np.abs(dpy) < hg) & (np.abs(dcx) < hgc) & (np.abs(dcy) < hgc)
        inside = (np.abs(opx) <= k["half_uv"][0]) & (np.abs(opy) <= k["half_uv"][1]) & \
                 (np.abs(ocx) <= k["half_c"][0]) & (np.abs(ocy) <= k["half_c"][1])
    tpx, tpy = np.maximum(tpx, MAGIC), np.maximum(tpy, MAGIC)
    tcx, tcy = np.maximum(tcx, k["clamp_c"][0]), np.maximum(tcy, k["clamp_c"][1])
    u32 = lambda a: a.view(np.uint32).astype(np.uint64)
    pix = (u32(tpy) * img_w + u32(tpx) + k["pix_k"]) & 0xffffffff
    cell = (u32(tcx) * map_w + u32(tcy) + k["cell_k"]) & 0xffffffff
    return cert, inside, pix, cell


CASES = [
    ("default grid, cam1", camera_setup_1, [[100, 300], [800, 1000]], 0.1, 2000, 2000, (1440, 1920), 0),
    ("2 km grid 0.2 m, cam1, turned vehicle", camera_setup_1, [[0, 2000], [0, 2000]], 0.2, 10000, 10000, (1440, 1920), 9),
    ("cam6, 0.2 m", camera_setup_6, [[0, 600], [0, 1400]], 0.2, 3000, 7000, (1440, 1920), 3),
    ("odd image size, 0.5 m", camera_setup_1, [[100, 300], [800, 1000]], 0.5, 400, 400, (1439, 1917), 1),
]


@pytest.mark.parametrize("name,cam_fn,boundary,res,mh,mw,hw,fidx", CASES)
def test_float32_decisions_agree_with_the_oracle_whenever_certified(name, cam_fn, boundary, res, mh, mw, hw, fidx):
    cam = cam_fn()
    fr = syn.synthetic_frame(123, fidx, 400000, height=hw[0], width=hw[1])
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    k = fast32(dict(map_height=mh, map_width=mw, num_classes=5, use_intensity=1, lane_index=2, device=0,
                    boundary_x_min=float(boundary[0][0]), boundary_y_min=float(boundary[1][0]), resolution=res,
                    origin_offset_x=OFF[0], origin_offset_y=OFF[1], range_max=100.0), T, cam.P, hw)
    assert k["coord_l"] > 0, "float32 path unexpectedly switched off"
    pts = fr["points"]
    # the oracle's decisions for every point
    _, _, uv, keep = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
    x64, y64 = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    gx = ((x64 + OFF[0]) - boundary[0][0]) / res
    gy = ((y64 + OFF[1]) - boundary[1][0]) / res
    on_grid = (gx > -1) & (gx < mh) & (gy > -1) & (gy < mw)
    ref_cell = np.trunc(gx).astype(np.int64) * mw + np.trunc(gy).astype(np.int64)
    ref_pix = np.zeros(len(pts), np.int64)
    ref_pix[keep] = uv[1].astype(np.int64) * hw[1] + uv[0].astype(np.int64)

    passed = cull32(k, pts)
    assert not np.any(keep & ~passed), "the cull dropped a point the reference keeps"
    assert passed.mean() < keep.mean() + 0.05, "the cull is much looser than the reference rule"

    cert, inside, pix, cell = decide32(k, pts[passed], hw[1], mw)
    want = (keep & on_grid)[passed]
    have = cert & inside
    assert np.array_equal(have[cert], want[cert]), "a certified decision differs from the reference"
    sel = have
    assert np.array_equal(pix[sel].astype(np.int64), ref_pix[passed][sel]), "certified pixel differs"
    assert np.array_equal(cell[sel].astype(np.int64), ref_cell[passed][sel]), "certified cell differs"
    assert cert.mean() > 0.9, "too few points certified in float32 (%.3f)" % cert.mean()


def test_float32_path_switches_off_when_its_preconditions_fail():
    cam = camera_setup_1()
    base = dict(map_height=2000, map_width=2000, num_classes=5, use_intensity=1, lane_index=2, device=0,
                boundary_x_min=100.0, boundary_y_min=800.0, resolution=0.1, origin_offset_x=OFF[0],
                origin_offset_y=OFF[1], range_max=100.0)
    T = np.linalg.inv(tr.get_transform_from_pose(syn.synthetic_pose(0)) @ syn.velodyne_to_baselink())
    assert fast32(base, T, cam.P, (1440, 1920))["coord_l"] > 0
    far = np.array(T); far[0, 3] += 1e9          # vehicle millions of cells away from the grid
    assert fast32(base, far, cam.P, (1440, 1920))["coord_l"] < 0
    assert fast32(dict(base, resolution=1e-7), T, cam.P, (1440, 1920))["coord_l"] < 0
    assert fast32(dict(base, range_max=float("inf")), T, cam.P, (1440, 1920))["coord_l"] < 0
    bad = np.array(cam.P); bad[0, 0] = np.nan
    assert fast32(base, T, bad, (1440, 1920))["coord_l"] < 0


def test_certification_holds_on_points_snapped_to_cell_and_range_borders():
    """Adversarial inputs: kept points moved to within one float32 ulp of BEV cell edges (grid edges and the (-1, 0)
    strip included) and along their viewing ray to the range limits.  Whatever the float32 path still certifies
    must agree with the reference; the rest is what the kernel hands to its float64 path."""
    cam = camera_setup_1()
    boundary, res, mh, mw, hw = [[100, 300], [800, 1000]], 0.1, 2000, 2000, (1440, 1920)
    fr = syn.synthetic_frame(321, 2, 300000)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    k = fast32(dict(map_height=mh, map_width=mw, num_classes=5, use_intensity=1, lane_index=2, device=0,
                    boundary_x_min=100.0, boundary_y_min=800.0, resolution=res, origin_offset_x=OFF[0],
                    origin_offset_y=OFF[1], range_max=100.0), T, cam.P, hw)
    rng = np.random.default_rng(8)
    _, _, _, keep = c_oracle.project_pcd(fr["pcd"], T, cam.P, fr["semantic_image"], 100.0)
    base = fr["pcd"][:, keep][:, :40000]
    clouds = []
    for axis in (0, 1):
        v = base.copy()
        g = ((v[axis] + OFF[axis]) - boundary[axis][0]) / res
        kk = np.where(rng.random(g.shape) < 0.3, rng.choice([-1.0, 0.0, float(mh)], size=g.shape), np.rint(g))
        x32 = (kk * res + boundary[axis][0] - OFF[axis]).astype(F)
        step = rng.integers(-1, 2, size=x32.shape)
        x32 = np.where(step < 0, np.nextafter(x32, F(-np.inf)), np.where(step > 0, np.nextafter(x32, F(np.inf)), x32))
        v[axis] = x32.astype(np.float64)
        clouds.append(v)
    velo = T @ np.vstack((base[0:3], np.ones((1, base.shape[1]))))
    for target in (100.0, 1e-4, 99.9999, 0.0):
        vv = velo.copy()
        vv[0:3] *= (target + rng.normal(0, 2e-5, vv.shape[1])) / vv[0]
        w = (np.linalg.inv(T) @ vv)[0:3].astype(F).astype(np.float64)
        clouds.append(np.vstack((w, base[3:4])))
    pcd = np.ascontiguousarray(np.hstack(clouds))
    pts = np.ascontiguousarray(pcd.T.astype(F))
    pcd = np.ascontiguousarray(pts.T.astype(np.float64))
    _, _, uv, keep = c_oracle.project_pcd(pcd, T, cam.P, fr["semantic_image"], 100.0)
    gx = ((pcd[0] + OFF[0]) - boundary[0][0]) / res
    gy = ((pcd[1] + OFF[1]) - boundary[1][0]) / res
    on_grid = (gx > -1) & (gx < mh) & (gy > -1) & (gy < mw)
    ref_cell = np.trunc(gx).astype(np.int64) * mw + np.trunc(gy).astype(np.int64)
    ref_pix = np.zeros(pcd.shape[1], np.int64)
    ref_pix[keep] = uv[1].astype(np.int64) * hw[1] + uv[0].astype(np.int64)
    passed = cull32(k, pts)
    assert not np.any(keep & ~passed)
    cert, inside, pix, cell = decide32(k, pts[passed], hw[1], mw)
    want = (keep & on_grid)[passed]
    have = cert & inside
    assert np.array_equal(have[cert], want[cert])
    assert np.array_equal(pix[have].astype(np.int64), ref_pix[passed][have])
    assert np.array_equal(cell[have].astype(np.int64), ref_cell[passed][have])
    assert 0.05 < (~cert).mean() < 0.9   # the borders really are exercised: many points are NOT certified


def test_velodyne_frame_cloud_without_transform():
    """pcd_frame_id == "velodyne": no world -> velodyne transform, re-centring point at the origin."""
    cam = camera_setup_1()
    boundary, res, mh, mw, hw = [[1360, 1500], [500, 630]], 0.5, 280, 260, (1440, 1920)
    fr = syn.synthetic_frame(77, 1, 200000)
    T = np.linalg.inv(tr.get_transform_from_pose(fr["pose"]) @ syn.velodyne_to_baselink())
    pcd = fr["pcd"].copy()
    pcd[0:3] = (T @ np.vstack((pcd[0:3], np.ones((1, pcd.shape[1])))))[0:3].astype(F)
    pts = np.ascontiguousarray(pcd.T.astype(F))
    k = fast32(dict(map_height=mh, map_width=mw, num_classes=5, use_intensity=1, lane_index=2, device=0,
                    boundary_x_min=float(boundary[0][0]), boundary_y_min=float(boundary[1][0]), resolution=res,
                    origin_offset_x=OFF[0], origin_offset_y=OFF[1], range_max=100.0), None, cam.P, hw)
    assert k["coord_l"] > 0 and np.all(k["n_ctr"] == 0)
    _, _, uv, keep = c_oracle.project_pcd(pcd, None, cam.P, fr["semantic_image"], 100.0)
    gx = ((pcd[0] + OFF[0]) - boundary[0][0]) / res
    gy = ((pcd[1] + OFF[1]) - boundary[1][0]) / res
    on_grid = (gx > -1) & (gx < mh) & (gy > -1) & (gy < mw)
    ref_cell = np.trunc(gx).astype(np.int64) * mw + np.trunc(gy).astype(np.int64)
    ref_pix = np.zeros(pcd.shape[1], np.int64)
    ref_pix[keep] = uv[1].astype(np.int64) * hw[1] + uv[0].astype(np.int64)
    passed = cull32(k, pts)
    assert not np.any(keep & ~passed)
    cert, inside, pix, cell = decide32(k, pts[passed], hw[1], mw)
    have = cert & inside
    assert np.array_equal(have[cert], (keep & on_grid)[passed][cert])
    assert np.array_equal(pix[have].astype(np.int64), ref_pix[passed][have])
    assert np.array_equal(cell[have].astype(np.int64), ref_cell[passed][have])
    assert cert.mean() > 0.9


@pytest.mark.parametrize("seed", range(12))
def test_certification_holds_for_random_poses_grids_and_cameras(seed):
    """Random vehicle positions (up to tens of kilometres from the origin, where float32 world coordinates are
    centimetre-coarse), headings, small roll / pitch, grid resolutions and origins, both cameras: the cull stays
    conservative and every certified float32 decision equals the reference's."""
    from vision_semantic_segmentation_b200.synthetic import Pose
    import math
    rng = np.random.default_rng(1000 + seed)
    cam = (camera_setup_1, camera_setup_6)[seed % 2]()
    scale = (50.0, 2000.0, 60000.0)[seed % 3]
    px, py = rng.uniform(-scale, scale, 2)
    yaw, pitch, roll = rng.uniform(-math.pi, math.pi), rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)
    cy, sy, cp, sp, cr, sr = math.cos(yaw / 2), math.sin(yaw / 2), math.cos(pitch / 2), math.sin(pitch / 2), math.cos(roll / 2), math.sin(roll / 2)
    quat = (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)
    pose = Pose((px, py, rng.uniform(-2, 2)), quat)
    res = float(rng.choice([0.05, 0.1, 0.2, 0.5, 1.0]))
    mh, mw = int(rng.integers(200, 3000)), int(rng.integers(200, 3000))
    # grid placed so that the vehicle is somewhere inside it (map coordinate = world + offset - boundary_min)
    bx = px + OFF[0] - rng.uniform(0.1, 0.9) * mh * res
    by = py + OFF[1] - rng.uniform(0.1, 0.9) * mw * res
    hw = (1440, 1920)
    fr = syn.synthetic_frame(500 + seed, 0, 120000, pose=pose)
    T = np.linalg.inv(tr.get_transform_from_pose(pose) @ syn.velodyne_to_baselink())
    k = fast32(dict(map_height=mh, map_width=mw, num_classes=5, use_intensity=1, lane_index=2, device=0,
                    boundary_x_min=float(bx), boundary_y_min=float(by), resolution=res, origin_offset_x=OFF[0],
                    origin_offset_y=OFF[1], range_max=100.0), T, cam.P, hw)
    assert k["coord_l"] > 0
    pts, pcd = fr["points"], fr["pcd"]
    _, _, uv, keep = c_oracle.project_pcd(pcd, T, cam.P, fr["semantic_image"], 100.0)
    gx = ((pcd[0] + OFF[0]) - bx) / res
    gy = ((pcd[1] + OFF[1]) - by) / res
    on_grid = (gx > -1) & (gx < mh) & (gy > -1) & (gy < mw)
    ref_cell = np.trunc(gx).astype(np.int64) * mw + np.trunc(gy).astype(np.int64)
    ref_pix = np.zeros(pcd.shape[1], np.int64)
    ref_pix[keep] = uv[1].astype(np.int64) * hw[1] + uv[0].astype(np.int64)
    passed = cull32(k, pts)
    assert not np.any(keep & ~passed), "the cull dropped a point the reference keeps"
    cert, inside, pix, cell = decide32(k, pts[passed], hw[1], mw)
    have = cert & inside
    assert np.array_equal(have[cert], (keep & on_grid)[passed][cert]), "a certified decision differs from the reference"
    assert np.array_equal(pix[have].astype(np.int64), ref_pix[passed][have])
    assert np.array_equal(cell[have].astype(np.int64), ref_cell[passed][have])
    if keep.sum() > 1000:
        assert cert[keep[passed]].mean() > 0.5, "the float32 path certifies too little to be useful (%.3f)" % cert[keep[passed]].mean()
