"""ctypes loader for the CPU oracle (oracle/gm_oracle.cpp).

TEST INFRASTRUCTURE ONLY — parity unpinned (see the header of gm_oracle.cpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
nothing under geometric_mapping_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libgm_oracle.so")
_L = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "gm_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        env = dict(os.environ)
        env.pop("CXX", None)
        res = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


def lib():
    global _L
    if _L is None:
        build()
        _L = C.CDLL(LIB)
        _L.gmo_crop.restype = C.c_int64
        _L.gmo_compact.restype = C.c_int64
        _L.gmo_voxel.restype = C.c_int64
        _L.gmo_refit_plane.restype = C.c_int64
        _L.gmo_refit_cylinder.restype = C.c_int64
        _L.gmo_argmax.restype = C.c_int32
        _L.gmo_polyline.restype = C.c_int32
        _L.gmo_num_threads.restype = C.c_int32
        _L.gmo_map_insert.restype = C.c_int64
    return _L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def num_threads() -> int:
    return int(lib().gmo_num_threads())


def crop(pts, bound, is_dense=True):
    pts = _f32(pts)
    n = pts.shape[0]
    out = np.empty((n, 4), np.float32)
    idx = np.empty(n, np.int32)
    m = lib().gmo_crop(_p(pts), C.c_int64(n), C.c_double(bound), C.c_int(1 if is_dense else 0), _p(out), _p(idx))
    return out[:m].copy(), idx[:m].copy()


def normals(pts, radius, mode=0, order=0, truth=False, nthreads=0):
    pts = _f32(pts)
    n = pts.shape[0]
    out = np.empty((n, 8), np.float32)
    cnt = np.empty(n, np.int32)
    tr = np.empty((n, 4), np.float64) if truth else None
    lib().gmo_normals(_p(pts), C.c_int64(n), C.c_double(radius), _p(out), _p(cnt), C.c_int(mode), C.c_int(order), _p(tr),
                      C.c_int(nthreads))
    return out, cnt, tr


def normals_knn(pts, k, cell=0.05, mode=0, nthreads=0, with_indices=False, max_radius=0.0):
    """k-nearest-neighbour normals (pcl::NormalEstimation::setKSearch).  -> normals8, count[, n x k neighbour indices]"""
    pts = _f32(pts)
    n = pts.shape[0]
    out = np.empty((n, 8), np.float32)
    cnt = np.empty(n, np.int32)
    idx = np.empty((n, k), np.int32) if with_indices else None
    lib().gmo_normals_knn(_p(pts), C.c_int64(n), C.c_int32(k), C.c_double(cell), _p(out), _p(cnt), _p(idx), C.c_int(mode), C.c_int(nthreads),
                          C.c_double(max_radius))
    return (out, cnt, idx) if with_indices else (out, cnt)


def compact(pts, normals8):
    pts, normals8 = _f32(pts), _f32(normals8)
    n = pts.shape[0]
    po, no, mp = np.empty((n, 4), np.float32), np.empty((n, 8), np.float32), np.empty(n, np.int32)
    m = lib().gmo_compact(_p(pts), _p(normals8), C.c_int64(n), _p(po), _p(no), _p(mp))
    return po[:m].copy(), no[:m].copy(), mp


def voxel(pts, leaf):
    pts = _f32(pts)
    n = pts.shape[0]
    keys, assign = np.empty(n, np.int32), np.empty(n, np.int32)
    cen = np.empty((max(n, 1), 4), np.float32)
    vkeys, vcnt = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
    grid6 = np.zeros(6, np.int32)
    status = C.c_int32(0)
    V = lib().gmo_voxel(_p(pts), C.c_int64(n), C.c_double(leaf), _p(keys), _p(assign), _p(cen), _p(vkeys), _p(vcnt),
                        _p(grid6), C.byref(status))
    cen_fx = cen[:V].copy()
    if status.value == 0 and V > 0:
        lib().gmo_voxel_fx_centroids(_p(pts), C.c_int64(n), _p(assign), C.c_int64(V), _p(cen_fx))
    # "centroids": pcl::VoxelGrid's float accumulation (stable order); "centroids_fx": the order-independent
    # fixed-point form the CUDA path computes (equal within ~1e-6 m)
    return {"V": int(V), "keys": keys, "assign": assign, "centroids": cen[:V].copy(), "centroids_fx": cen_fx,
            "voxel_keys": vkeys[:V].copy(), "voxel_counts": vcnt[:V].copy(), "grid6": grid6, "status": status.value}


def nn1(query, pts, nthreads=0):
    query, pts = _f32(query), _f32(pts)
    idx = np.empty(query.shape[0], np.int32)
    d2 = np.empty(query.shape[0], np.float32)
    lib().gmo_nn1(_p(query), C.c_int64(query.shape[0]), _p(pts), C.c_int64(pts.shape[0]), _p(idx), _p(d2), C.c_int(nthreads))
    return idx, d2


def nn1_grid(query, pts, cell, nthreads=0):
    query, pts = _f32(query), _f32(pts)
    idx = np.empty(query.shape[0], np.int32)
    d2 = np.empty(query.shape[0], np.float32)
    lib().gmo_nn1_grid(_p(query), C.c_int64(query.shape[0]), _p(pts), C.c_int64(pts.shape[0]), C.c_double(cell), _p(idx),
                       _p(d2), C.c_int(nthreads))
    return idx, d2


def local_frame(normals8, wf, weight_mode=0):
    normals8 = _f32(normals8)
    S, vals, vecs = np.empty(9, np.float32), np.empty(3, np.float32), np.empty(9, np.float32)
    St = np.empty(9, np.float64)
    lib().gmo_local_frame(_p(normals8), C.c_int64(normals8.shape[0]), C.c_double(wf), _p(S), _p(vals), _p(vecs), _p(St),
                          C.c_int32(weight_mode))
    return {"scatter": S.reshape(3, 3), "vals": vals, "vecs": vecs.reshape(3, 3), "scatter_truth": St.reshape(3, 3)}


def local_frame_dense(normals8, wf):
    normals8 = _f32(normals8)
    S = np.empty(9, np.float32)
    lib().gmo_local_frame_dense(_p(normals8), C.c_int64(normals8.shape[0]), C.c_double(wf), _p(S))
    return S.reshape(3, 3)


def eigen_markers(vals, vecs):
    out = np.empty(30, np.float32)
    lib().gmo_eigen_markers(_p(_f32(vals)), _p(_f32(vecs).reshape(-1)), _p(out))
    return out.reshape(3, 10)


def plane_hypotheses(pts, samples):
    pts = _f32(pts)
    s = np.ascontiguousarray(samples, np.int32)
    H = s.shape[0]
    coef, valid = np.empty((H, 4), np.float32), np.empty(H, np.int32)
    lib().gmo_plane_hypotheses(_p(pts), C.c_int64(pts.shape[0]), _p(s), C.c_int32(H), _p(coef), _p(valid))
    return coef, valid


def count_plane(pts, coef, valid, tau, nthreads=0):
    pts, coef = _f32(pts), _f32(coef)
    valid = np.ascontiguousarray(valid, np.int32)
    H = coef.shape[0]
    counts = np.empty(H, np.int32)
    lib().gmo_count_plane(_p(pts), C.c_int64(pts.shape[0]), _p(coef), _p(valid), C.c_int32(H), C.c_double(tau), _p(counts),
                          C.c_int(nthreads))
    return counts


def cyl_hypotheses(pts, normals8, samples, rmin, rmax, tau):
    pts, normals8 = _f32(pts), _f32(normals8)
    s = np.ascontiguousarray(samples, np.int32)
    H = s.shape[0]
    m7, t12, valid = np.empty((H, 7), np.float32), np.empty((H, 12), np.float32), np.empty(H, np.int32)
    lib().gmo_cyl_hypotheses(_p(pts), _p(normals8), C.c_int64(pts.shape[0]), _p(s), C.c_int32(H), C.c_double(rmin),
                             C.c_double(rmax), C.c_double(tau), _p(m7), _p(t12), _p(valid))
    return m7, t12, valid


def cyl_test_params(model7, tau):
    m = _f32(model7).reshape(-1, 7)
    t = np.empty((m.shape[0], 12), np.float32)
    lib().gmo_cyl_test_params(_p(m), C.c_int32(m.shape[0]), C.c_double(tau), _p(t))
    return t


def count_cyl(pts, test12, valid, nthreads=0):
    pts, test12 = _f32(pts), _f32(test12)
    valid = np.ascontiguousarray(valid, np.int32)
    H = test12.shape[0]
    counts = np.empty(H, np.int32)
    lib().gmo_count_cyl(_p(pts), C.c_int64(pts.shape[0]), _p(test12), _p(valid), C.c_int32(H), _p(counts), C.c_int(nthreads))
    return counts


def argmax(counts):
    c = np.ascontiguousarray(counts, np.int32)
    return int(lib().gmo_argmax(_p(c), C.c_int32(c.shape[0])))


def refit_plane(pts, coef_in, tau):
    pts = _f32(pts)
    ci = _f32(coef_in)
    co = np.empty(4, np.float32)
    cnt = lib().gmo_refit_plane(_p(pts), C.c_int64(pts.shape[0]), _p(ci), C.c_double(tau), _p(co))
    return co, int(cnt)


def refit_cylinder(pts, model7_in, test12_in, iters):
    pts = _f32(pts)
    mi, ti = _f32(model7_in), _f32(test12_in)
    mo = np.empty(7, np.float32)
    rms = C.c_double(0.0)
    cnt = lib().gmo_refit_cylinder(_p(pts), C.c_int64(pts.shape[0]), _p(mi), _p(ti), C.c_int32(iters), _p(mo), C.byref(rms))
    return mo, int(cnt), rms.value


def labels(pts, plane4, tau, cyl_test12):
    pts = _f32(pts)
    out = np.empty(pts.shape[0], np.uint8)
    p4 = None if plane4 is None else _f32(plane4)
    t12 = None if cyl_test12 is None else _f32(cyl_test12)
    lib().gmo_labels(_p(pts), C.c_int64(pts.shape[0]), _p(p4), C.c_double(tau), _p(t12), _p(out))
    return out


def polyline(pts, normals8, labels_u8, want, axis3, wf, L, max_slices, weight_mode=0):
    pts, normals8 = _f32(pts), _f32(normals8)
    lab = None if labels_u8 is None else np.ascontiguousarray(labels_u8, np.uint8)
    ax = _f32(axis3)
    out = np.zeros((max_slices, 12), np.float64)
    t0 = C.c_double(0.0)
    S = lib().gmo_polyline(_p(pts), _p(normals8), _p(lab), C.c_int64(pts.shape[0]), C.c_int(want), _p(ax), C.c_double(wf),
                           C.c_double(L), C.c_int32(max_slices), _p(out), C.byref(t0), C.c_int32(weight_mode))
    return out[:S].copy(), t0.value


def compress(pts, labels_u8, plane4, cyl7, tau, leaf):
    pts = _f32(pts)
    lab = np.ascontiguousarray(labels_u8, np.uint8)
    p4 = None if plane4 is None else _f32(plane4)
    c7 = None if cyl7 is None else _f32(cyl7)
    ints = np.zeros(6, np.int32)
    fl = np.zeros(32, np.float32)
    cen = np.zeros((max(pts.shape[0], 1), 4), np.float32)
    lib().gmo_compress(_p(pts), _p(lab), C.c_int64(pts.shape[0]), _p(p4), _p(c7), C.c_double(tau), C.c_double(leaf), _p(ints), _p(fl),
                       _p(cen))
    return {"n_points": int(ints[0]), "n_plane": int(ints[1]), "n_cylinder": int(ints[2]), "n_residual": int(ints[3]),
            "n_residual_voxels": int(ints[4]), "plane_u": fl[0:3].copy(), "plane_v": fl[3:6].copy(), "plane_bounds": fl[6:10].copy(),
            "plane_rms": float(fl[10]), "cyl_t_range": fl[11:13].copy(), "cyl_rms": float(fl[13]), "residual_rms": float(fl[14]),
            "total_rms": float(fl[15]), "residual_centroids": cen[: int(ints[4])].copy()}


def map_insert(state, pts, leaf, labels=None, label_filter=-1, pose34=None):
    """Aggregated voxel map (gmo_map_insert).  state = None or the dict returned by a previous call;
    -> {"keys" u64, "ijk" int32 Vx3, "counts", "sums" int64 Vx3, "centroids" Vx4, "out_of_range"} sorted by key."""
    pts = _f32(pts)
    n = pts.shape[0]
    if state is None:
        k0, c0, s0 = np.empty(0, np.uint64), np.empty(0, np.int32), np.empty((0, 3), np.int64)
        oor0 = 0
    else:
        k0, c0, s0, oor0 = state["keys"], state["counts"], state["sums"], state["out_of_range"]
    V = len(k0)
    cap = V + n + 1
    ko, co, so = np.empty(cap, np.uint64), np.empty(cap, np.int32), np.empty((cap, 3), np.int64)
    ceo = np.empty((cap, 4), np.float32)
    oor = C.c_int64(0)
    lab = None if labels is None else np.ascontiguousarray(labels, np.uint8)
    pose = None if pose34 is None else _f32(np.asarray(pose34).reshape(12))
    Vn = lib().gmo_map_insert(_p(np.ascontiguousarray(k0)), _p(np.ascontiguousarray(c0)), _p(np.ascontiguousarray(s0)), C.c_int64(V),
                              _p(pts), _p(lab), C.c_int64(n), C.c_int32(label_filter if lab is not None else -1), _p(pose),
                              C.c_double(leaf), _p(ko), _p(co), _p(so), _p(ceo), C.byref(oor))
    keys = ko[:Vn].copy()
    m = np.uint64((1 << 21) - 1)
    ijk = np.stack([(keys & m).astype(np.int64) - (1 << 20), ((keys >> np.uint64(21)) & m).astype(np.int64) - (1 << 20),
                    ((keys >> np.uint64(42)) & m).astype(np.int64) - (1 << 20)], axis=1).astype(np.int32)
    return {"keys": keys, "ijk": ijk, "counts": co[:Vn].copy(), "sums": so[:Vn].copy(), "centroids": ceo[:Vn].copy(),
            "out_of_range": oor0 + int(oor.value)}



def trig(y, x):
    """The oracle's libm-independent atan2(y>=0, x), sin(x), cos(x) (x in [0, pi/3] for sin/cos), elementwise."""
    y, x = _f32(y), _f32(x)
    n = x.shape[0]
    a, s, c = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    lib().gmo_trig(_p(y), _p(x), C.c_int64(n), _p(a), _p(s), _p(c))
    return a, s, c
