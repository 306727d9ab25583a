"""Generates the committed golden vectors.  Run HERE (the container with /root/reference); the outputs travel.

1. bspline_notebook.npz — executes the numeric half of the reference's own
   /root/reference/scripts/CubicBSpline3D.ipynb (the only reference artefact that runs in this image) and
   stores its sample points p, control points Q and the printed BSpline rows.
2. c1_oracle.npz — oracle outputs on the seeded C1 window (surfels, per-point assignment, correspondences,
   per-iteration LM costs, final data_cor): a regression pin of the restatement itself (NOT a reference
   fixture: BuildSurfels / Match / the solve are untested upstream, parity for them is unpinned).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)


def notebook():
    nb = json.load(open("/root/reference/scripts/CubicBSpline3D.ipynb"))
    src = "".join(nb["cells"][0]["source"])
    src = src.split("# Spline display")[0].replace("import matplotlib.pyplot as plt", "")
    src = src.replace("print(BSpline)", "")
    env = {}
    exec(compile(src, "CubicBSpline3D.ipynb", "exec"), env)
    stored = "".join(nb["cells"][0]["outputs"][0]["text"])
    first = np.array([float(v) for v in stored.split("]")[0].replace("[", "").split()])
    assert np.allclose(env["BSpline"][0], first, atol=5e-9), "notebook no longer reproduces its stored output"
    np.savez(os.path.join(HERE, "bspline_notebook.npz"), p=env["p"].astype(np.float64), Q=env["Q"], BSpline=env["BSpline"],
             Nbs=env["Nbs"])
    print("bspline_notebook.npz", env["BSpline"].shape)


def c1():
    from oracle import wc_oracle as O
    from wildcat_slam_b200 import synthetic as S

    w = S.make_window("C1")
    r = O.build_surfels(w.points, want_assign=True, want_info=True)
    st, sld = O.update_surfel_poses(w.imu, r["surfels"])
    rf = O.build_surfels(w.fix_points)
    st, fix = O.update_surfel_poses(w.fix_imu, rf["surfels"])
    cs, _ = O.match(sld, sld, True)
    cf, _ = O.match(sld, fix, False)
    st, smp, summ = O.window_solve(sld, fix, cs, cf, w.imu, w.samples)
    n = summ.num_iterations
    np.savez_compressed(
        os.path.join(HERE, "c1_oracle.npz"),
        points_xyz=np.stack([w.points["x"], w.points["y"], w.points["z"]], 1), points_t=w.points["time"],
        surfels=r["surfels"], assign=r["assign"], n_points=r["info"]["n_points"], layer=r["info"]["layer"],
        sld_body=sld, fix_body=fix, sld_corr=cs, fix_corr=cf, data_cor=smp["data_cor"],
        iter_cost=np.array(summ.iter_cost[: n + 1]), iter_accepted=np.array(summ.iter_accepted[: n + 1]),
        initial_cost=summ.initial_cost, final_cost=summ.final_cost, termination=summ.termination)
    print("c1_oracle.npz", len(r["surfels"]), len(cs), len(cf), n)


def sweep():
    """3. c1_sweep_oracle.npz — the sweep-preparation row (AddLidarScan's extrinsic + range / blind-box filter and
    UndistortSweep) of the oracle on a seeded C1-sized cloud that exercises every filter outcome: a regression pin of the
    restatement (untested upstream: parity unpinned)."""
    from oracle import wc_oracle as O
    from wildcat_slam_b200 import synthetic as S

    w = S.make_window("C1")
    n = 3000  # a prefix keeps the fixture small
    pts = w.points[:n].copy()
    rng = np.random.default_rng(11)
    k = len(pts) // 7
    pts["x"][:k] = rng.uniform(-1.0, 1.0, k).astype(np.float32)
    pts["y"][:k] = rng.uniform(-1.0, 1.0, k).astype(np.float32)
    pts["z"][:k] = rng.uniform(-0.6, 0.6, k).astype(np.float32)
    pts["x"][k:2 * k] = rng.uniform(-130.0, 130.0, k).astype(np.float32)
    pts["y"][k:2 * k] = rng.uniform(-130.0, 130.0, k).astype(np.float32)
    st, kept = O.filter_points(pts)
    assert st == 0
    st, und = O.undistort_sweep(w.imu, w.points[:n])
    assert st == 0
    xyz = lambda a: np.stack([a["x"], a["y"], a["z"]], 1)
    np.savez_compressed(os.path.join(HERE, "c1_sweep_oracle.npz"), raw_xyz=xyz(pts), raw_t=pts["time"], kept_xyz=xyz(kept),
                        kept_t=kept["time"], undistorted_xyz=xyz(und))
    print("c1_sweep_oracle.npz", len(pts), len(kept))


if __name__ == "__main__":
    notebook()
    c1()
    sweep()
