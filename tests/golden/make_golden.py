"""Generates tests/golden/oracle_v1.npz from the CPU oracle on small seeded scans.

The reference ships no fixtures and cannot be built here, so these vectors
pin the ORACLE (oracle/gm_oracle.cpp) against regressions; its semantics are pinned separately by
the analytic known answers in tests/test_oracle.py.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from geometric_mapping_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")
PARAMS = dict(n=6000, seed=17, bound=5.0, radius=0.3, leaf=0.25, wf=0.2, tau=0.05, H=96)


def scan(p=PARAMS):
    pts = synth.curved_tunnel(p["n"], seed=p["seed"], outlier_frac=0.03)
    pts[5, 0] = 7.5          # outside the crop box
    pts[6] = [np.nan, 0, 0, 1]  # kept by the dense-cloud quirk, dropped with its NaN normal
    return pts


def compute(p=PARAMS):
    pts = scan(p)
    cropped, src = O.crop(pts, p["bound"], True)
    nrm, cnt, _ = O.normals(cropped, p["radius"], mode=1, order=0)
    cloud, nrm_c, vmap = O.compact(cropped, nrm)
    vox = O.voxel(cloud, p["leaf"])
    # 1-NN query = the order-independent fixed-point centroid (what the CUDA path defines; within 1e-6 m of "centroids")
    nn, _ = O.nn1(vox["centroids_fx"], np.where(np.isfinite(cropped), cropped, 1e30).astype(np.float32))
    fr = O.local_frame(nrm_c, p["wf"])
    ps = synth.sample_indices(len(cloud), p["H"], 3, seed=3)
    cs = synth.sample_indices(len(cloud), p["H"], 2, seed=4)
    pcoef, pvalid = O.plane_hypotheses(cloud, ps)
    pcnt = O.count_plane(cloud, pcoef, pvalid, p["tau"])
    m7, t12, cvalid = O.cyl_hypotheses(cloud, nrm_c, cs, 0.5, 10.0, p["tau"])
    ccnt = O.count_cyl(cloud, t12, cvalid)
    pb, cb = O.argmax(pcnt), O.argmax(ccnt)
    prefit, prc = O.refit_plane(cloud, pcoef[pb], p["tau"])
    crefit, crc, crms = O.refit_cylinder(cloud, m7[cb], t12[cb], 5)
    lab = O.labels(cloud, prefit, p["tau"], O.cyl_test_params(crefit, p["tau"])[0])
    poly, t0 = O.polyline(cloud, nrm_c, lab, 2, fr["vecs"][:, 0], p["wf"], 1.0, 64)
    return dict(
        crop_src=src, nbr_count=cnt, normals=nrm, valid_map=vmap, voxel_keys=vox["keys"], voxel_assign=vox["assign"],
        voxel_centroids=vox["centroids"], voxel_grid=vox["grid6"], nn_index=nn, frame_scatter=fr["scatter"],
        frame_vals=fr["vals"], frame_vecs=fr["vecs"], plane_samples=ps, cyl_samples=cs, plane_coef=pcoef,
        plane_valid=pvalid, plane_counts=pcnt, cyl_model=m7, cyl_test=t12, cyl_valid=cvalid, cyl_counts=ccnt,
        plane_refit=prefit, plane_refit_count=np.int64(prc), cyl_refit=crefit, cyl_refit_count=np.int64(crc),
        cyl_refit_rms=np.float64(crms), labels=lab, polyline=poly, polyline_t0=np.float64(t0))


if __name__ == "__main__":
    np.savez_compressed(OUT, **compute())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
