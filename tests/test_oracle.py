"""CPU tests of the oracle (oracle/gm_oracle.cpp): analytic known answers, self-consistency of the
accelerated forms against the brute-force definitions, and the committed golden vectors.

The reference has no tests or fixtures; the only expected values it pins are
analytic: for a straight cylinder the min-eigenvector of the normal scatter is the cylinder axis
(src/tunnel_processing.cpp:91, :139-143).  Everything else here pins the builder's restatement.
"""
import os

import numpy as np
import pytest

from geometric_mapping_b200 import synth
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz")


def _angle(a, b):
    a = a / np.linalg.norm(a, axis=-1, keepdims=True)
    b = b / np.linalg.norm(b, axis=-1, keepdims=True)
    return np.arccos(np.clip(np.abs((a * b).sum(-1)), -1.0, 1.0))


# ---- a1 ----------------------------------------------------------------------------------------
def test_crop_inclusive_bounds_and_nan_quirk():
    b = np.float32(5.0)
    up = np.nextafter(b, np.float32(6))
    pts = np.array([[5, 5, 5, 1], [-5, -5, -5, 1], [up, 0, 0, 1], [0, -up, 0, 1], [np.nan, 0, 0, 1], [0, np.inf, 0, 1],
                    [1, 2, 3, 7]], np.float32)
    out, idx = O.crop(pts, 5.0, True)
    assert list(idx) == [0, 1, 4, 6]            # bounds inclusive; NaN kept in a dense cloud (A.1); inf dropped
    assert out[3, 3] == 7                       # padding word is carried through
    out, idx = O.crop(pts, 5.0, False)
    assert list(idx) == [0, 1, 6]               # !is_dense: non-finite skipped
    assert O.crop(np.zeros((0, 4), np.float32), 5.0)[0].shape == (0, 4)


# ---- a2/a3 -------------------------------------------------------------------------------------
def test_normals_pure_plane_known_answer():
    nrm = np.array([1.0, 2.0, 2.0]) / 3.0
    pts = synth.plane_patch(4000, seed=5, normal=nrm, offset=-1.5, half=1.0)
    out, cnt, truth = O.normals(pts, 0.25, truth=True)
    ok = cnt >= 6
    assert ok.mean() > 0.9
    assert _angle(out[ok, :3].astype(np.float64), nrm[None]).max() < 2e-2      # float, unshifted covariance (A.3)
    assert _angle(truth[ok, :3], nrm[None]).max() < 1e-6                       # double truth
    assert np.median(out[ok, 4]) < 1e-3                                        # curvature ~ 0 (float noise level)
    assert (-(pts[ok, :3] * out[ok, :3]).sum(1) >= 0).all()                    # flipped towards the origin
    assert (out[ok][:, [3, 5, 6, 7]] == 0).all()


def test_normals_brute_force_equals_grid_and_orders_agree():
    pts = synth.curved_tunnel(5000, seed=9)
    a, ca, _ = O.normals(pts, 0.3, mode=1, order=0)
    b, cb, _ = O.normals(pts, 0.3, mode=0, order=0)
    assert np.array_equal(ca, cb)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))                # same neighbours, same order
    c, cc, _ = O.normals(pts, 0.3, mode=0, order=1)                            # index order: same set, other order
    assert np.array_equal(cc, ca)
    ok = ca >= 8
    assert np.quantile(_angle(a[ok, :3].astype(np.float64), c[ok, :3].astype(np.float64)), 0.99) < 2e-2
    # threads do not change anything (parallel over points only)
    d, _, _ = O.normals(pts, 0.3, mode=0, order=0, nthreads=1)
    assert np.array_equal(b.view(np.uint32), d.view(np.uint32))


def test_normals_fewer_than_three_neighbours_is_nan_and_compaction_is_stable():
    pts = np.array([[0, 0, 0, 1], [0.01, 0, 0, 1], [3, 3, 3, 1], [0, 0.01, 0, 1], [0.01, 0.01, 0.002, 1]], np.float32)
    nr, cnt, _ = O.normals(pts, 0.05, mode=1)
    assert list(cnt) == [4, 4, 1, 4, 4] and np.isnan(nr[2, :3]).all() and np.isnan(nr[2, 4])
    pc, nc, mp = O.compact(pts, nr)
    assert list(mp) == [0, 1, -1, 2, 3] and np.array_equal(pc, pts[[0, 1, 3, 4]])
    # strict d2 < r2: a point exactly at distance r is not a neighbour
    two = np.array([[0, 0, 0, 1], [0.5, 0, 0, 1]], np.float32)
    assert list(O.normals(two, 0.5, mode=1)[1]) == [1, 1]
    assert list(O.normals(two, float(np.nextafter(np.float32(0.5), np.float32(1))), mode=1)[1]) == [2, 2]


# ---- a4 ----------------------------------------------------------------------------------------
def test_voxel_lattice_known_keys():
    g = np.arange(-2, 3, dtype=np.float32) * 0.25 + 0.01
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.ones(xx.size)], 1).astype(np.float32)
    v = O.voxel(pts, 0.25)
    assert v["V"] == 125 and list(v["grid6"]) == [-2, -2, -2, 5, 5, 5]
    i, j, k = np.meshgrid(np.arange(5), np.arange(5), np.arange(5), indexing="ij")
    assert np.array_equal(v["keys"], (i + 5 * j + 25 * k).ravel())             # key = ijk . (1, dx, dx*dy)
    assert np.array_equal(v["voxel_keys"], np.arange(125))                     # ascending key order
    assert np.array_equal(v["centroids"][v["assign"]][:, :3], pts[:, :3])      # one point per voxel
    assert (v["centroids"][:, 3] == 1.0).all()


def test_voxel_centroids_counts_and_overflow_rule():
    pts = synth.curved_tunnel(20000, seed=4)
    v = O.voxel(pts, 0.2)
    assert v["voxel_counts"].sum() == len(pts) and (np.diff(v["voxel_keys"]) > 0).all()
    for j in (0, v["V"] // 2, v["V"] - 1):
        m = pts[v["assign"] == j]
        s = np.zeros(3, np.float32)
        for q in m:                                                             # float sum in index order
            s = s + q[:3]
        assert np.array_equal(v["centroids"][j, :3], s / np.float32(len(m)))
    big = O.voxel(pts, 0.001)                                                   # dx*dy*dz > INT32_MAX
    assert big["status"] == 1 and big["V"] == len(pts) and np.array_equal(big["centroids"], pts)
    assert O.voxel(np.zeros((0, 4), np.float32), 0.1)["V"] == 0


def test_nn1_grid_equals_brute_force_with_lowest_index_ties():
    pts = synth.curved_tunnel(20000, seed=7)
    pts[100] = pts[50]                                                          # exact duplicate -> tie
    q = np.concatenate([synth.curved_tunnel(300, seed=8), pts[[100]], np.array([[4.9, 4.9, 4.9, 1]], np.float32)])
    a, da = O.nn1(q, pts)
    assert a[300] == 50
    for cell in (0.05, 0.3, 3.0):
        b, db = O.nn1_grid(q, pts, cell)
        assert np.array_equal(a, b) and np.array_equal(da, db)


# ---- a5 / a6 -----------------------------------------------------------------------------------
def test_local_frame_dense_form_is_bit_identical_to_diagonal_form():
    g = np.random.Generator(np.random.Philox(1))
    nr = np.zeros((300, 8), np.float32)
    v = g.normal(size=(300, 3))
    nr[:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    nr[:, 4] = g.uniform(0, 0.33, 300)
    a = O.local_frame(nr, 0.2)["scatter"]
    b = O.local_frame_dense(nr, 0.2)                                            # literal n x n weights (:100-124)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_local_frame_weight_formula_and_straight_cylinder_axis():
    nr = np.zeros((1, 8), np.float32)
    nr[0, :3] = [0, 1, 0]
    nr[0, 4] = 0.1
    fr = O.local_frame(nr, 0.2)
    w = np.float32(np.exp((np.float64(np.float32(0.1)) + 0.001 / 0.2) ** 2))    # :106, (curv + .001/wf)^2
    assert fr["scatter"][1, 1] == np.float32(w * np.float32(1)) * np.float32(w * np.float32(1))
    pts = synth.straight_cylinder(30000, seed=1, noise=0.0)
    nrm, cnt, _ = O.normals(pts, 0.2)
    cloud, nc, _ = O.compact(pts, nrm)
    fr = O.local_frame(nc, 0.2)
    assert abs(abs(fr["vecs"][0, 0]) - 1.0) < 1e-4                              # axis = generator axis (x)
    assert fr["vals"][0] < 1e-3 * fr["vals"][1] and fr["vals"][0] <= fr["vals"][1] <= fr["vals"][2]
    assert np.abs(fr["scatter"] - fr["scatter_truth"]).max() <= 1e-4 * fr["vals"][2]
    assert all(fr["vecs"][np.abs(fr["vecs"][:, k]).argmax(), k] > 0 for k in range(3))   # sign convention


def test_eigen_markers_formula():
    vals = np.array([1.0, 2.0, 2.0], np.float32)
    vecs = np.eye(3, dtype=np.float32)
    m = O.eigen_markers(vals, vecs)
    e = np.float32(1.0) / np.float32(3.0) * np.abs(vals)                       # (1/norm) * |vals|  (:265)
    for i in range(3):
        assert np.array_equal(m[i, 3:6], vecs[:, i]) and (m[i, 0:3] == 0).all()
        assert m[i, 6] == np.float32(0.1 - 0.05 * np.float64(e[i]))
        assert m[i, 7] == np.float32(0.3 - 0.15 * np.float64(e[i]))
        assert m[i, 8] == np.float32(0.25 - 0.125 * np.float64(e[i]))


# ---- a8 (builder-defined) --------------------------------------------------------------------------
def test_plane_hypothesis_known_answer_and_degenerate_samples():
    pts = np.array([[0, 0, 1, 1], [1, 0, 1, 1], [0, 1, 1, 1], [2, 2, 1, 1], [3, 3, 1.04, 1], [0, 0, 1.06, 1], [2, 0, 1, 1]], np.float32)
    coef, valid = O.plane_hypotheses(pts, np.array([[0, 1, 2], [0, 0, 1], [0, 1, 9], [0, 1, 6]], np.int32))
    assert list(valid) == [1, 0, 0, 0]                                          # duplicate, out of range, collinear
    assert np.allclose(coef[0], [0, 0, 1, -1])
    counts = O.count_plane(pts, coef, valid, 0.05)
    assert list(counts) == [6, -1, -1, -1]                                      # |d| < tau strict: 0.04 in, 0.06 out
    assert O.argmax(counts) == 0 and O.argmax(np.array([3, 7, 7, -1], np.int32)) == 1 and O.argmax(np.array([-1, -1], np.int32)) == -1


def test_cylinder_hypothesis_and_refit_recover_exact_cylinder():
    pts = synth.straight_cylinder(20000, seed=6, noise=0.0, radius=2.0)
    nr = np.zeros((len(pts), 8), np.float32)
    nr[:, 1:3] = -pts[:, 1:3] / 2.0                                            # exact inward normals
    samples = synth.sample_indices(len(pts), 64, 2, seed=1)
    m7, t12, valid = O.cyl_hypotheses(pts, nr, samples, 0.5, 10.0, 0.05)
    assert valid.sum() > 48
    good = valid == 1
    assert np.abs(m7[good, 6] - 2.0).max() < 1e-3                              # radius
    assert np.abs(np.abs(m7[good, 3]) - 1.0).max() < 1e-4                      # axis = x
    assert np.abs(m7[good, 1:3]).max() < 2e-3                                  # axis passes through y=z=0
    counts = O.count_cyl(pts, t12, valid)
    assert counts[good].min() > 0.999 * len(pts)
    # radius limits reject
    assert O.cyl_hypotheses(pts, nr, samples, 2.5, 10.0, 0.05)[2].sum() == 0
    # refit from a perturbed model converges to the exact one
    start = m7[np.argmax(counts)].copy()
    start[6] += 0.02
    ref, n_in, rms = O.refit_cylinder(pts, start, O.cyl_test_params(start, 0.05)[0], 5)
    assert n_in == len(pts) and abs(ref[6] - 2.0) < 1e-5 and rms < 1e-5
    # test parameters: |dist - r| < tau  <=>  |A^2 + B^2 - mid| < half
    t = O.cyl_test_params(np.array([0, 0, 0, 1, 0, 0, 2.0], np.float32), 0.05)[0]
    assert np.isclose(t[8], 0.5 * (2.05 ** 2 + 1.95 ** 2)) and np.isclose(t[9], 0.5 * (2.05 ** 2 - 1.95 ** 2))
    lab = O.labels(np.array([[0, 2.04, 0, 1], [0, 2.06, 0, 1], [0, 0, -1.96, 1], [0, 0, 1.94, 1]], np.float32), None, 0.05, t)
    assert list(lab) == [2, 0, 2, 0]


def test_refit_plane_and_polyline_known_answers():
    pts = synth.plane_patch(5000, seed=2, normal=(0, 0, 1), offset=-1.5, half=2.0, noise=0.01)
    coef, cnt = O.refit_plane(pts, np.array([0, 0, 1, 1.48], np.float32), 0.05)
    assert cnt > 4500 and abs(coef[2] - 1.0) < 1e-4 and abs(coef[3] - 1.5) < 1e-3
    cyl = synth.straight_cylinder(40000, seed=3, noise=0.0, radius=2.5)
    nr = np.zeros((len(cyl), 8), np.float32)
    nr[:, 1:3] = -cyl[:, 1:3] / 2.5
    lab = np.full(len(cyl), 2, np.uint8)
    poly, t0 = O.polyline(cyl, nr, lab, 2, np.array([1, 0, 0], np.float32), 0.2, 1.0, 64)
    assert len(poly) == 10 and t0 == -5.0
    assert np.abs(poly[:, 6] - 2.5).max() < 1e-5 and np.abs(poly[:, 1:3]).max() < 1e-5   # radius, centre on the axis
    assert np.allclose(poly[:, 0], np.arange(10) - 4.5) and np.abs(np.abs(poly[:, 3]) - 1).max() < 1e-6
    assert poly[:, 7].sum() == len(cyl)


# ---- golden vectors ---------------------------------------------------------------------------------
def test_golden_vectors():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.compute()
    ref = np.load(GOLDEN)
    exact = ["crop_src", "nbr_count", "valid_map", "voxel_keys", "voxel_assign", "voxel_grid", "nn_index", "plane_samples",
             "cyl_samples", "plane_valid", "plane_counts", "cyl_valid", "cyl_counts", "labels", "plane_refit_count",
             "cyl_refit_count"]
    for k in exact:
        assert np.array_equal(got[k], ref[k]), k
    # float outputs: the fixture was produced by the same code on x86-64 with the same flags, so they
    # are expected bit-equal; libm (sin/cos/atan2/exp) differences across glibc builds are tolerated.
    for k in ["normals", "voxel_centroids", "frame_scatter", "frame_vals", "frame_vecs", "plane_coef", "cyl_model", "cyl_test",
              "plane_refit", "cyl_refit", "cyl_refit_rms", "polyline", "polyline_t0"]:
        a, b = np.asarray(got[k], np.float64), np.asarray(ref[k], np.float64)
        assert a.shape == b.shape, k
        both_nan = np.isnan(a) & np.isnan(b)
        assert np.allclose(np.where(both_nan, 0, a), np.where(both_nan, 0, b), rtol=1e-5, atol=1e-6), k


def test_map_insert_known_answers_and_order_independence():
    """Aggregated voxel map oracle: lattice points land in the expected global cells; inserting the same
    clouds in another order, or split differently, gives the identical state (integer sums)."""
    leaf = 0.1
    pts = np.array([[0.05, 0.05, 0.05, 1], [0.06, 0.01, 0.09, 1], [-0.01, 0.0, 0.0, 1], [1.234, -5.678, 0.25, 1], [np.nan, 0, 0, 1]],
                   np.float32)
    st = O.map_insert(None, pts, leaf)
    assert st["out_of_range"] == 1
    assert st["ijk"].tolist() == [[-1, 0, 0], [0, 0, 0], [12, -57, 2]]           # sorted by (z, y, x)
    by_cell = {tuple(c): k for c, k in zip(st["ijk"].tolist(), st["counts"].tolist())}
    assert by_cell == {(0, 0, 0): 2, (-1, 0, 0): 1, (12, -57, 2): 1}
    c000 = st["centroids"][[tuple(c) == (0, 0, 0) for c in st["ijk"].tolist()].index(True)]
    assert np.abs(c000[:3] - np.array([0.055, 0.03, 0.07])).max() < 2e-6
    g = np.random.Generator(np.random.Philox(3))
    a = g.uniform(-3, 3, (5000, 4)).astype(np.float32)
    b = g.uniform(-3, 3, (7000, 4)).astype(np.float32)
    pose = np.array([[0.6, -0.8, 0, 1.5], [0.8, 0.6, 0, -2.0], [0, 0, 1, 0.25]], np.float32)
    s1 = O.map_insert(O.map_insert(None, a, leaf, pose34=pose), b, leaf)
    s2 = O.map_insert(O.map_insert(None, b, leaf), a, leaf, pose34=pose)
    s3 = O.map_insert(O.map_insert(O.map_insert(None, b[:100], leaf), a[::-1], leaf, pose34=pose), b[100:], leaf)
    for s in (s2, s3):
        assert np.array_equal(s1["keys"], s["keys"]) and np.array_equal(s1["counts"], s["counts"]) and np.array_equal(s1["sums"], s["sums"])
    assert s1["counts"].sum() == 12000
    lab = (np.arange(5000) % 3).astype(np.uint8)
    only = O.map_insert(None, a, leaf, labels=lab, label_filter=2)
    assert only["counts"].sum() == (lab == 2).sum()



# ---- second source for the oracle (scipy / numpy, double precision): the restatement of the UPSTREAM semantics is a
# single-author reading of PCL / FLANN / Eigen; these checks compare it with independent library implementations -----
def test_oracle_against_scipy_and_numpy_second_source():
    from scipy.spatial import cKDTree

    pts = synth.curved_tunnel(8_000, seed=71, outlier_frac=0.02)
    cropped, _ = O.crop(pts, 5.0, True)
    xyz = cropped[:, :3].astype(np.float64)
    r = 0.25
    nrm, cnt, truth = O.normals(cropped, r, mode=0, order=0, truth=True)
    tree = cKDTree(xyz)
    # radius neighbour sets: FLANN's strict float test vs scipy's double test differ only for pairs within rounding of r
    lists = tree.query_ball_point(xyz, r * (1 - 1e-6))
    lists_hi = tree.query_ball_point(xyz, r * (1 + 1e-6))
    lo = np.array([len(l) for l in lists]); hi = np.array([len(l) for l in lists_hi])
    assert ((cnt >= lo) & (cnt <= hi)).all()
    assert (cnt == lo).mean() > 0.999
    # normals: the double "truth" of the oracle equals numpy's eigh of the two-pass covariance
    for i in range(0, len(xyz), 97):
        nb = xyz[lists[i]]
        if len(nb) != cnt[i] or len(nb) < 3:
            continue
        w, v = np.linalg.eigh(np.cov(nb.T, bias=True))
        n0 = v[:, 0] * (1 if np.dot(v[:, 0], -xyz[i]) >= 0 else -1)
        assert np.arccos(min(1.0, abs(float(np.dot(n0, truth[i, :3]))))) < 1e-6
        assert abs(w[0] / w.sum() - truth[i, 3]) < 1e-9
        # ... and the float single-pass result (PCL's form) is near it
        assert np.arccos(min(1.0, abs(float(np.dot(n0, nrm[i, :3].astype(np.float64)))))) < 5e-2
    # exact 1-NN and k-NN sets against cKDTree
    q = cropped[::7].copy()
    idx, _ = O.nn1(q, cropped)
    d, j = tree.query(q[:, :3].astype(np.float64), k=1)
    assert (idx == j).mean() > 0.999  # (ties / rounding at equal distances aside)
    _, kc, ki = O.normals_knn(cropped, 12, cell=0.1, mode=0, with_indices=True)
    dk, jk = tree.query(xyz, k=12)
    same = sum(set(jk[t]) == set(ki[t]) for t in range(len(xyz)))
    assert same >= 0.999 * len(xyz)
    # VoxelGrid keys against a numpy restatement in float32
    leaf = 0.2
    vox = O.voxel(cropped, leaf)
    inv = np.float32(1.0) / np.float32(leaf)
    mn = cropped[:, :3].min(0); mx = cropped[:, :3].max(0)
    minb = np.floor(mn * inv).astype(np.int64); maxb = np.floor(mx * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(cropped[:, :3] * inv) - minb.astype(np.float32)).astype(np.int64)
    keys = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    assert np.array_equal(keys.astype(np.int32), vox["keys"])
    assert np.array_equal(np.unique(keys).astype(np.int32), vox["voxel_keys"])
    # getLocalFrame: eigen-decomposition against numpy.linalg.eigh on the oracle's own scatter matrix
    cloud, nrm_c, _ = O.compact(cropped, nrm)
    fr = O.local_frame(nrm_c, 0.2)
    w, v = np.linalg.eigh(fr["scatter"].astype(np.float64))
    assert np.abs(w - fr["vals"]).max() <= 1e-6 * np.abs(w).max()
    assert abs(abs(float(np.dot(v[:, 0], fr["vecs"][:, 0]))) - 1.0) < 1e-6


def test_oracle_trig_is_correctly_rounded_and_libm_independent():
    g = np.random.Generator(np.random.Philox(5))
    n = 200_000
    y = (np.abs(g.normal(size=n)) * 10.0 ** g.uniform(-6, 2, n)).astype(np.float32)
    x = (g.normal(size=n) * 10.0 ** g.uniform(-6, 2, n)).astype(np.float32)
    th = g.uniform(0, np.pi / 3, n).astype(np.float32)
    a, _, _ = O.trig(y, x)
    _, s, c = O.trig(y, th)
    assert np.array_equal(a, np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32))
    assert np.array_equal(s, np.sin(th.astype(np.float64)).astype(np.float32))
    assert np.array_equal(c, np.cos(th.astype(np.float64)).astype(np.float32))
    z = np.zeros(1, np.float32)
    assert O.trig(z, z)[0][0] == 0.0 and abs(O.trig(z, -z)[0][0] - np.float32(np.pi)) == 0.0   # atan2(+0, -0) = pi


def test_knn_oracle_grid_equals_brute_force():
    pts = synth.curved_tunnel(3_000, seed=9, outlier_frac=0.1)
    pts[4] = [np.nan, 0, 0, 1]
    a = O.normals_knn(pts, 20, cell=0.15, mode=0, with_indices=True)
    b = O.normals_knn(pts, 20, cell=0.15, mode=1, with_indices=True)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
