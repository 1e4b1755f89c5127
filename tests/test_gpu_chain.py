"""End-to-end parity from RAW POINTS: the CUDA path (through the C-ABI) against the CPU oracle running ALONE
(oracle/chain.py: every oracle stage is fed the oracle's own previous output, never the GPU's), at the sizes
BASELINE.json names:

  C0  100k-point straight cylinder, the reference's CPU-runnable case
  C1  1M-point curved tunnel + floor + noise + outliers, 1024 hypotheses (512 plane + 512 cylinder)
  C2  the same scan with 4096 hypotheses (2048 + 2048)

Two modes of gm_normals:
  canonical (gm_set_normals_mode(1)): neighbourhood sums in FLANN's (d2, index) order -> normals are BIT-IDENTICAL to
     the oracle's, and so is every integer output downstream (valid map, voxel keys / assignment / order, hypothesis
     coefficients, inlier counts, argmax); refits / frame / polyline within 1e-4 relative (north_star).
  fast (default): the neighbour sets and everything that does not depend on the float sums stay exact; the measured
     divergence of everything that does is asserted against explicit bars below.
Also: the CUDA path against the committed golden fixture (tests/golden/oracle_v1.npz).
"""
import json
import os

import numpy as np
import pytest

from geometric_mapping_b200 import capi, synth
from oracle import chain
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TAU = 0.05


gpu_chain = chain.gpu_chain


def _assert_frame(m):
    """getLocalFrame: the scatter matrix and its eigenvalues within 1e-4 relative (north_star), the axis within 1e-3 rad.
    The yardstick is the oracle's DOUBLE accumulation of the same float products: its float accumulation (the literal
    restatement, Eigen's order being unspecified anyway) carries a rounding error of its own that grows with n -- it is
    reported as float_oracle_scatter_rel_err_vs_double and bounds how far the GPU may sit from the float oracle."""
    # bit-identical normals: only the reduction differs (1e-5); fast normals: the scatter entries move with them (1e-4)
    assert m["scatter_rel_diff_vs_double"] <= (1e-5 if m["normals_bit_identical"] else 1e-4) and m["eigenvalue_rel_diff_vs_double"] <= 1e-4, m
    assert m["axis_angle_rad_vs_double"] <= 1e-3, m
    assert m["scatter_rel_diff"] <= 1e-4 + 2.0 * m["float_oracle_scatter_rel_err_vs_double"], m


def _assert_canonical(m):
    """Bars of the canonical mode (integer outputs were already asserted exact inside chain.compare)."""
    assert m["normals_bit_identical"]
    _assert_frame(m)
    assert m["plane_refit_count_equal"] and m["plane_refit_normal_angle_rad"] <= 1e-4 and m["plane_refit_d_abs_diff"] <= 1e-4 * 5.0, m
    assert m["cyl_refit_axis_angle_rad"] <= 1e-4 and m["cyl_refit_radius_rel_diff"] <= 1e-4 and m["cyl_refit_axis_offset_rel"] <= 1e-4, m
    assert m["cyl_refit_count_rel_diff"] == 0, m
    # labels come from the refined models (floats within 1e-4): only points within that distance of a band edge may flip
    assert m["label_mismatch_fraction"] <= 1e-4, m
    assert m["polyline_slices"][0] == m["polyline_slices"][1]
    assert m["polyline_center_abs_diff_max_m"] <= 1e-4 * 5.0 and m["polyline_radius_rel_diff_max"] <= 1e-4, m


def _assert_fast(m, well_conditioned=True):
    """Measured divergence of the default (fast) summation order, with the bars it has to stay under.  The neighbour
    SETS are identical (asserted exactly in chain.compare); only the ORDER of the float additions inside a neighbourhood
    differs from FLANN's, and PCL's single-pass covariance E[xx]-E[x]^2 amplifies that rounding:

    * well-conditioned neighbourhoods (C0: r = 0.15 m over 1 cm noise): normals within ~1e-3 rad, the refined cylinder
      within 1e-4 relative of the oracle's (north_star's bar), labels identical;
    * C1/C2 as SURVEY 8(d) fixes them (r = 0.05 m over 2 cm noise, ~39 neighbours): the 5 cm blob is nearly isotropic, so
      the smallest eigenvector is decided by the last bits of the float sums -- the oracle's own float normals sit as far
      from the double-precision truth as the GPU's do (tests/test_gpu_parity.py:test_normals_parity) -- and every
      cylinder hypothesis is built from two such normals: counts move by percents, RANSAC may elect a different winner and
      the refit then converges to a different (equally weak: ~15 % inliers) cylinder.  Everything that does NOT pass
      through the float sums stays exact, the frame stays within 1e-4, and the canonical mode above reproduces the
      oracle bit for bit; the divergence of the rest is recorded, not bounded tightly."""
    _assert_frame(m)
    assert m["plane_refit_count_equal"] and m["plane_refit_normal_angle_rad"] <= 1e-4 and m["plane_refit_d_abs_diff"] <= 1e-4 * 5.0, m
    assert m["cyl_valid_pattern_equal"] or not well_conditioned, m
    if well_conditioned:
        assert m["normal_angle_rad"]["p50"] <= 1e-3 and m["normal_angle_rad"]["p99"] <= 1e-2, m
        assert m["cyl_best_id_equal"] and m["cyl_best_count_rel_diff"] <= 1e-2, m
        assert m["cyl_refit_axis_angle_rad"] <= 1e-3 and m["cyl_refit_radius_rel_diff"] <= 1e-3 and m["cyl_refit_axis_offset_rel"] <= 1e-3, m
        assert m["label_mismatch_fraction"] <= 5e-3, m
        assert m["polyline_slices"][0] == m["polyline_slices"][1]
    else:
        assert m["normal_angle_rad"]["p50"] <= 1e-2 and m["normal_angle_rad"]["p99"] <= 0.1, m
        assert m["cyl_best_count_rel_diff"] <= 0.25, m


def _record(name, m):
    """Keep the measured numbers with the run (gpurun_out/ travels back from the GPU box)."""
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "chain_parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **m}) + "\n")
    except OSError:
        pass


def _run_case(name, pts, Hp, Hc, radius, leaf, bound=5.0, well_conditioned=True):
    f = chain.front(pts, bound=bound, radius=radius, leaf=leaf)
    nv = f["n_valid"]
    assert nv > 0.5 * len(pts)
    ps = synth.sample_indices(nv, Hp, 3, seed=3)
    cs = synth.sample_indices(nv, Hc, 2, seed=4)
    b = chain.back(f, ps, cs, tau=TAU)
    assert b["plane_best"] >= 0 and b["cyl_best"] >= 0
    g = gpu_chain(pts, ps, cs, True, radius, leaf, bound)
    m = chain.compare(g, f, b, exact_normals=True)
    _record(name + "/canonical", m)
    _assert_canonical(m)
    g = gpu_chain(pts, ps, cs, False, radius, leaf, bound)
    m = chain.compare(g, f, b, exact_normals=False)
    _record(name + "/fast", m)
    _assert_fast(m, well_conditioned)
    return f, b


def test_c0_straight_cylinder_100k_from_raw_points():
    """BASELINE configs[0]: 100k-point straight cylinder (R = 2.5 m, axis = x, 1 cm noise).  r = 0.15 m gives ~45 neighbours."""
    pts = synth.straight_cylinder(100_000, seed=1)
    f, b = _run_case("C0_100k", pts, 512, 512, radius=0.15, leaf=0.1)
    # analytic known answer (src/tunnel_processing.cpp:91): the centre axis of a straight cylinder is its generator
    axis = f["frame"]["vecs"][:, 0]
    assert abs(abs(axis[0]) - 1.0) < 1e-3
    assert abs(b["cyl_refit"][6] - 2.5) < 5e-3


@pytest.mark.parametrize("H", [1024, 4096])
def test_c1_c2_curved_tunnel_1m_from_raw_points(H):
    """BASELINE configs[1] (H = 1024) and configs[2] (H = 4096): 1M points, r = 0.05, leaf = 0.1, tau = 0.05."""
    pts = synth.curved_tunnel(1_000_000, seed=2)
    _run_case(f"C1_1M_H{H}", pts, H // 2, H - H // 2, radius=0.05, leaf=0.1, well_conditioned=False)


def test_canonical_mode_small_radius_with_junk_points():
    """NaN / inf / out-of-box points, isolated points (NaN normals) and duplicates through the canonical mode."""
    pts = synth.curved_tunnel(40_000, seed=5, outlier_frac=0.05)
    pts[100] = [np.nan, 0, 0, 1]
    pts[101] = [0, np.inf, 0, 1]
    pts[102] = [9.0, 0, 0, 1]
    pts[200:210] = pts[300:310]          # exact duplicates: equal d2, tie broken by index
    _run_case("junk_40k", pts, 128, 128, radius=0.12, leaf=0.2)


def test_cuda_path_against_the_golden_fixture():
    """tests/golden/oracle_v1.npz was written by the oracle (make_golden.py); the CUDA path in canonical mode reproduces
    its integer vectors exactly and its floats within 1e-4, from the same raw points."""
    import importlib.util

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    p = mg.PARAMS
    gold = np.load(os.path.join(here, "golden", "oracle_v1.npz"))
    pts = mg.scan(p)
    g = gpu_chain(pts, gold["plane_samples"], gold["cyl_samples"], True, p["radius"], p["leaf"], p["bound"])
    assert g["n_cropped"] == len(gold["crop_src"])
    assert np.array_equal(g["nbr_count"], gold["nbr_count"])
    assert np.array_equal(chain._bits(g["normals"]), chain._bits(gold["normals"]))   # NaN rows compare as NaN
    assert np.array_equal(g["valid_map"], gold["valid_map"])
    assert np.array_equal(g["vox_key_pt"], gold["voxel_keys"]) and np.array_equal(g["vox_assign"], gold["voxel_assign"])
    assert np.array_equal(g["grid6"], gold["voxel_grid"])
    assert np.abs(g["centroids"][:, :3] - gold["voxel_centroids"][:, :3]).max() <= 2e-6
    assert np.array_equal(g["nn_index"], gold["nn_index"])
    assert np.array_equal(g["plane_coef"].view(np.uint32), gold["plane_coef"].view(np.uint32))
    assert np.array_equal(g["plane_counts"], gold["plane_counts"])
    assert np.array_equal(g["cyl_model"].view(np.uint32), gold["cyl_model"].view(np.uint32))
    assert np.array_equal(g["cyl_test"].view(np.uint32), gold["cyl_test"].view(np.uint32))
    assert np.array_equal(g["cyl_counts"], gold["cyl_counts"])
    assert np.abs(g["frame"]["vals"] - gold["frame_vals"]).max() <= 1e-4 * np.abs(gold["frame_vals"]).max()
    assert np.abs(g["plane"]["coef"] - gold["plane_refit"]).max() <= 1e-4 * 5.0
    assert g["plane"]["refit_count"] == int(gold["plane_refit_count"])
    assert abs(g["cyl"]["coef"][6] - gold["cyl_refit"][6]) <= 1e-4 * gold["cyl_refit"][6]
    assert g["cyl"]["refit_count"] == int(gold["cyl_refit_count"])
    assert (g["labels"] != gold["labels"]).mean() <= 1e-3


# ---- k-nearest-neighbour normals (the north-star's "grid-hashed k-NN": pcl::NormalEstimation::setKSearch) ---------
@pytest.mark.parametrize("n,k,radius,outliers,cap", [(30_000, 16, 0.12, 0.05, 0.0), (30_000, 32, 0.3, 0.0, 0.0), (3_000, 8, 0.05, 0.3, 0.0),
                                                     (200, 64, 0.5, 0.0, 0.0), (40, 64, 1.0, 0.0, 0.0), (30_000, 16, 0.1, 0.05, 0.35),
                                                     (30_000, 32, 0.2, 0.1, 0.15)])
def test_knn_neighbour_sets_and_normals_bit_exact(n, k, radius, outliers, cap):
    """Neighbour index lists (in FLANN's result order) against the BRUTE-FORCE oracle, normals bit-identical.  The grid
    cell (neighborRadius) only sizes the search: dense stencils (pass A), sparse ones and outliers (pass B, block shells),
    clouds smaller than k; with and without a radius cap (the k nearest within `cap`, inside and beyond one cell)."""
    pts = synth.curved_tunnel(n, seed=61, outlier_frac=outliers)
    pts[3] = [np.nan, 0, 0, 1]
    pts[5:8] = pts[9]                      # duplicates: equal distances, tie broken by index
    cropped, _ = O.crop(pts, 5.0, True)
    ref_n, ref_c, ref_i = O.normals_knn(cropped, k, cell=radius, mode=1, with_indices=True, max_radius=cap)
    with capi.Context(capi.default_params(neighborRadius=radius), max_points=n, max_hypotheses=4) as ctx:
        ctx.set_knn(k, keep_indices=True, max_radius=cap)
        ctx.upload_scan(pts)
        ctx.crop()
        ctx.normals()
        assert ctx.counts().device_error == 0
        idx, cnt, nrm = ctx.download_knn_indices(), ctx.download_neighbor_counts(), ctx.download_normals(0)
    fin = np.isfinite(cropped[:, :3]).all(1)
    assert np.array_equal(cnt[fin], ref_c[fin])
    assert np.array_equal(idx[fin], ref_i[fin])
    assert np.array_equal(chain._bits(nrm), chain._bits(ref_n))


def test_knn_chain_c1_1m_k32_from_raw_points():
    """BASELINE configs[1] in k mode (k = 32): the whole chain from raw points against the oracle alone; in this
    mode the default path already sums in FLANN's order, so everything integer is exact without a special mode."""
    pts = synth.curved_tunnel(1_000_000, seed=2)
    f = chain.front(pts, radius=0.07, leaf=0.1, knn=32, knn_max_radius=0.25)
    nv = f["n_valid"]
    assert nv < f["n_cropped"]   # the isolated outliers end with < 3 neighbours inside the cap and leave with the compaction
    ps, cs = synth.sample_indices(nv, 512, 3, seed=3), synth.sample_indices(nv, 512, 2, seed=4)
    b = chain.back(f, ps, cs, tau=TAU)
    g = gpu_chain(pts, ps, cs, False, 0.07, 0.1, knn=32, knn_max_radius=0.25)
    m = chain.compare(g, f, b, exact_normals=True)
    _record("C1_1M_knn32", m)
    _assert_canonical(m)
