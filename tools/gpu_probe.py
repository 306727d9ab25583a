"""First-contact probe for a GPU box: runs every stage once on C1/C2 against the oracle and prints diagnostics
(not a test; the tests are tests/test_gpu_parity.py)."""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
from oracle import wc_oracle as O  # noqa: E402
from wildcat_slam_b200 import odometry as od  # noqa: E402
from wildcat_slam_b200 import synthetic as S  # noqa: E402
from wildcat_slam_b200 import types as T  # noqa: E402


def stage(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[ok ] {name}  ({time.time() - t:.3f}s)  {r if r is not None else ''}", flush=True)
        return True
    except Exception:
        print(f"[ERR] {name}", flush=True)
        traceback.print_exc()
        return False


def main():
    ctx = od.Context(0)
    for name in sys.argv[1:] or ["C1", "C2"]:
        print("=====", name, flush=True)
        w = S.make_window(name)
        ref = O.build_surfels(w.points, want_assign=True)
        res = {}

        def extract():
            tm = {}
            g, a = od.BuildSurfels(w.points, ctx=ctx, want_assign=True, timing=tm)
            res["g"] = g
            same = a.tobytes() == ref["assign"].tobytes()
            o = ref["surfels"]
            msg = f"S gpu={len(g)} oracle={len(o)} assign_exact={same} gpu_ms={tm['gpu_ms']:.3f}"
            if len(g) == len(o) and len(o):
                msg += (f" dt={np.abs(g['timestamp'] - o['timestamp']).max():.2e} dc={np.abs(g['center'] - o['center']).max():.2e}"
                        f" dcov={np.abs(g['covariance'] - o['covariance']).max():.2e}"
                        f" dn={1 - np.sum(g['norm'] * o['norm'], 1).min():.2e} res_eq={(g['resolution'] == o['resolution']).all()}")
            return msg

        stage("extract", extract)
        st, sld = O.update_surfel_poses(w.imu, ref["surfels"])
        st, fix = O.update_surfel_poses(w.fix_imu, O.build_surfels(w.fix_points)["surfels"])

        def poses():
            g = od.UpdateSurfelPoses(w.imu, ref["surfels"], ctx=ctx)
            return "max diff " + " ".join(f"{f}={np.abs(g[f] - sld[f]).max():.2e}" for f in ("pos", "rot", "center", "norm", "covariance"))

        stage("update_surfel_poses", poses)
        cs, _ = O.match(sld, sld, True, use_kdtree=False)
        cf, _ = O.match(sld, fix, False, use_kdtree=False)

        def match():
            m = od.KnnSurfelMatcher(ctx)
            m.BuildIndex(sld)
            tm = {}
            g, _ = m.Match(sld, timing=tm)
            m2 = od.KnnSurfelMatcher(ctx)
            m2.BuildIndex(fix)
            g2, _ = m2.Match(sld)
            return (f"sld gpu={len(g)} oracle={len(cs)} equal={g.tobytes() == cs.tobytes()} ms={tm['gpu_ms']:.3f}; "
                    f"fix gpu={len(g2)} oracle={len(cf)} equal={g2.tobytes() == cf.tobytes()}")

        stage("match", match)

        def evaluate():
            rng = np.random.default_rng(7)
            smp = w.samples.copy()
            smp["data_cor"] = rng.normal(size=(len(smp), 12)) * 1e-3
            out = []
            for mode in (0, 1):
                o = T.default_solve_opts()
                o.jacobian_mode = mode
                st, c_o, g_o, H_o = O.window_evaluate(sld, fix, cs, cf, w.imu, smp, opts=o)
                c_g, g_g, H_g = od.EvaluateWindow(sld, fix, cs, cf, w.imu, smp, opts=o, ctx=ctx)
                out.append(f"mode{mode}: cost rel={abs(c_g - c_o) / c_o:.2e} g={np.abs(g_g - g_o).max() / np.abs(g_o).max():.2e} "
                           f"H={np.abs(H_g - H_o).max() / np.abs(H_o).max():.2e}")
            o = T.default_solve_opts()
            o.use_imu_factors = 0
            st, c_o, g_o, H_o = O.window_evaluate(sld, fix, cs, cf, None, smp, opts=o)
            c_g, g_g, H_g = od.EvaluateWindow(sld, fix, cs, cf, None, smp, opts=o, ctx=ctx)
            out.append(f"lidar-only: cost rel={abs(c_g - c_o) / c_o:.2e} g={np.abs(g_g - g_o).max() / np.abs(g_o).max():.2e} "
                       f"H={np.abs(H_g - H_o).max() / np.abs(H_o).max():.2e}")
            return " | ".join(out)

        stage("evaluate", evaluate)

        def solve():
            st, smp_o, so = O.window_solve(sld, fix, cs, cf, w.imu, w.samples)
            smp_g, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
            n = min(so.num_iterations, sg.num_iterations)
            dc = np.abs(np.array(sg.iter_cost[1:n + 1]) / np.array(so.iter_cost[1:n + 1]) - 1).max() if n else 0
            return (f"iters gpu={sg.num_iterations} oracle={so.num_iterations} term gpu={sg.termination} oracle={so.termination} "
                    f"cost gpu={sg.final_cost:.9g} oracle={so.final_cost:.9g} iter_cost_rel={dc:.2e} "
                    f"dx={np.abs(smp_g['data_cor'] - smp_o['data_cor']).max():.2e} gpu_ms={sg.gpu_ms_total:.3f}")

        stage("solve", solve)

        def spline():
            st, smp_o, so = O.window_solve(sld, fix, cs, cf, w.imu, w.samples)
            st, s_o, i_o = O.apply_corrections(smp_o, w.imu)
            s_g, i_g = od.ApplyCorrections(smp_o, w.imu, ctx=ctx)
            return (f"oracle st={st} samples d={max(np.abs(s_g[f] - s_o[f]).max() for f in ('rot', 'pos', 'data_cor')):.2e} "
                    f"imu d={max(np.abs(i_g[f] - i_o[f]).max() for f in ('rot', 'pos')):.2e}")

        stage("apply_corrections", spline)
    ctx.close()


if __name__ == "__main__":
    main()
