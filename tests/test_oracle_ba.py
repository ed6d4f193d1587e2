"""CPU: pin the BA oracle (oracle/ba_oracle.c).  The reference pins nothing (no tests, no g2o binary here), so the
restatement is checked for internal consistency and against scipy on the same Huber-ised objective."""
import numpy as np
import pytest

from oracle import ba_oracle as B


@pytest.fixture(scope="module")
def prob(pkg):
    return pkg.synth.synth_ba_problem(3, 6, 150, outlier_frac=0.05)


def _project(T, p, K):
    T = T.reshape(3, 4)
    q = K @ (T[:, :3] @ p + T[:, 3])
    return q[:2] / q[2]


def test_se3_exp_matches_closed_form(pkg):
    rng = np.random.default_rng(0)
    for _ in range(20):
        xi = rng.normal(0, 0.3, 6)
        R, t = B.se3_exp(xi)
        R2, t2 = pkg.synth.se3_exp(xi)
        assert np.allclose(R, R2, atol=1e-13) and np.allclose(t, t2, atol=1e-13)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-13)
    R, t = B.se3_exp(np.array([1.0, 2.0, 3.0, 0, 0, 0]))
    assert np.allclose(R, np.eye(3)) and np.allclose(t, [1, 2, 3])


def test_jacobians_vs_finite_differences(pkg, prob):
    """optimization.cpp:52-73: A = de/d(xi) for T <- exp(xi) T, B = de/dp"""
    K = prob["K"]
    rng = np.random.default_rng(1)
    for i in rng.integers(0, len(prob["obs_pose"]), 10):
        T = prob["poses"][prob["obs_pose"][i]]
        p = prob["points"][prob["obs_point"][i]]
        z = prob["obs_uv"][i]
        e, A, Bm = B.residual_and_jacobians(T, p, K, z)
        assert np.allclose(e, z - _project(T, p, K), atol=1e-12)
        h = 1e-6
        for a in range(6):
            xi = np.zeros(6); xi[a] = h
            dR, dt = B.se3_exp(xi)
            T4 = T.reshape(3, 4)
            Tn = np.hstack([dR @ T4[:, :3], (dR @ T4[:, 3] + dt)[:, None]])
            num = ((z - _project(Tn, p, K)) - e) / h
            assert np.allclose(num, A[:, a], rtol=2e-4, atol=2e-4)
        for a in range(3):
            dp = np.zeros(3); dp[a] = h
            num = ((z - _project(T, p + dp, K)) - e) / h
            assert np.allclose(num, Bm[:, a], rtol=2e-4, atol=2e-4)
        _, A2, _ = B.residual_and_jacobians(T, p, K, z, pose_only=True)   # optimization.cpp:84-101
        assert np.allclose(A, A2, rtol=1e-12)


def test_first_lm_step_equals_dense_normal_equations(pkg, prob):
    """Schur-complement solve == solving the full (6K+3L) damped system"""
    H, b = B.dense_system(prob["poses"], prob["points"], prob["obs_pose"], prob["obs_point"], prob["obs_uv"], prob["K"])
    assert np.allclose(H, H.T)
    lam = 1e-5 * np.abs(np.diag(H)).max()
    x = np.linalg.solve(H + lam * np.eye(len(b)), b)
    r = B.optimize(prob["poses"], prob["points"], prob["obs_pose"], prob["obs_point"], prob["obs_uv"], prob["K"],
                   num_iterations=1)
    assert r["trials"] == 1 and r["accepted"] == 1
    assert np.isclose(r["trace"][0, 0], lam, rtol=1e-12)
    nK = len(prob["poses"])
    for k in range(nK):
        dR, dt = B.se3_exp(x[6 * k:6 * k + 6])
        T4 = prob["poses"][k].reshape(3, 4)
        Tn = np.hstack([dR @ T4[:, :3], (dR @ T4[:, 3] + dt)[:, None]]).reshape(-1)
        assert np.allclose(r["poses"][k], Tn, rtol=1e-8, atol=1e-9)
    assert np.allclose(r["points"], prob["points"] + x[6 * nK:].reshape(-1, 3), rtol=1e-8, atol=1e-8)


def test_lm_monotone_and_schedule(pkg, prob):
    r = B.optimize(prob["poses"], prob["points"], prob["obs_pose"], prob["obs_point"], prob["obs_uv"], prob["K"],
                   num_iterations=10)
    tr = r["trace"]
    acc = tr[tr[:, 3] == 1]
    chi = np.concatenate([[r["chi2_initial"]], acc[:, 1]])
    assert (np.diff(chi) < 0).all()                 # accepted steps strictly decrease the robust chi2
    assert np.isclose(r["chi2_final"], chi[-1])
    assert np.isclose(B.chi2(r["poses"], r["points"], prob["obs_pose"], prob["obs_point"], prob["obs_uv"], prob["K"]),
                      r["chi2_final"], rtol=1e-12)
    assert r["iterations"] == 10 and r["trials"] >= 10
    # lambda schedule: after an accepted step lambda shrinks by a factor in [1/3, 2/3]
    for a, b in zip(tr[:-1], tr[1:]):
        if a[3] == 1:
            assert 1 / 3 - 1e-12 <= b[0] / a[0] <= 2 / 3 + 1e-12


def test_optimum_agrees_with_scipy(pkg):
    """the converged robust cost equals scipy.optimize.least_squares(loss='huber') on the same residuals"""
    from scipy.optimize import least_squares
    p = pkg.synth.synth_ba_problem(5, 4, 40, outlier_frac=0.1)
    K, op, ol, uv = p["K"], p["obs_pose"], p["obs_point"], p["obs_uv"]
    nK, nL = len(p["poses"]), len(p["points"])
    r = B.optimize(p["poses"], p["points"], op, ol, uv, K, num_iterations=300)

    def fun_from(base_poses, base_pts):
        def unpack(x):
            poses = []
            for k in range(nK):
                dR, dt = B.se3_exp(x[6 * k:6 * k + 6])
                T4 = base_poses[k].reshape(3, 4)
                poses.append(np.hstack([dR @ T4[:, :3], (dR @ T4[:, 3] + dt)[:, None]]).reshape(-1))
            return np.array(poses), base_pts + x[6 * nK:].reshape(-1, 3)

        def fun(x):
            poses, pts = unpack(x)
            T = poses[op].reshape(-1, 3, 4)
            pc = np.einsum("nij,nj->ni", T[:, :, :3], pts[ol]) + T[:, :, 3]
            q = pc @ K.T
            e = uv - q[:, :2] / q[:, 2:3]
            return np.sqrt((e * e).sum(1) + 1e-30)   # one residual per edge: Huber acts on the 2-D edge norm like g2o
        return fun

    x0 = np.zeros(6 * nK + 3 * nL)
    # (a) same objective: scipy's cost at the oracle's inputs / outputs equals the oracle's robust chi2
    f0 = fun_from(p["poses"], p["points"])(x0)
    huber = lambda f: np.where(f <= 5.991, f * f, 2 * 5.991 * f - 5.991 ** 2).sum()
    assert np.isclose(huber(f0), r["chi2_initial"], rtol=1e-9)
    fun = fun_from(r["poses"], r["points"])
    assert np.isclose(huber(fun(x0)), r["chi2_final"], rtol=1e-6)
    # (b) the oracle's converged estimate is a local minimum of that objective: scipy cannot improve it
    s = least_squares(fun, x0, loss="huber", f_scale=5.991, method="trf", max_nfev=50)
    chi_scipy = 2 * s.cost
    assert r["chi2_final"] < r["chi2_initial"] * 0.5
    assert chi_scipy <= r["chi2_final"] * (1 + 1e-6)
    assert (r["chi2_final"] - chi_scipy) / chi_scipy < 2e-3, (r["chi2_final"], chi_scipy)


def test_relabel_semantics(pkg):
    p = pkg.synth.synth_ba_problem(7, 5, 80, outlier_frac=0.15)
    r = B.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=5)
    chi = r["chi2_per_obs"]
    th = r["chi2_threshold"]
    assert th == 5.991 or (chi <= th / 2).mean() <= 0.5          # doubled only while the inlier ratio was <= 0.5
    assert r["n_inlier_obs"] == (chi <= th).sum() and r["n_outlier_obs"] == (chi > th).sum()
    last = {}
    for i, l in enumerate(p["obs_point"]):
        last[l] = chi[i] <= th                                    # last observation of a landmark wins
    assert all(r["point_inlier"][l] == v for l, v in last.items())
    # pose-only variant: landmarks untouched
    r2 = B.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=10,
                    pose_only=True)
    assert np.array_equal(r2["points"], p["points"]) and r2["chi2_final"] < r2["chi2_initial"]


def test_golden_ba(pkg):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ba_seed11_K5_L120_it10.npz"))
    p = pkg.synth.synth_ba_problem(11, 5, 120, outlier_frac=0.05)
    r = B.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=10)
    assert r["trials"] == int(g["trials"])
    assert np.allclose(r["poses"], g["poses"], rtol=1e-9, atol=1e-12) and np.allclose(r["points"], g["points"], rtol=1e-9)
    assert np.allclose([r["chi2_initial"], r["chi2_final"]], g["chi2"], rtol=1e-10)
    assert np.array_equal(r["point_inlier"], g["point_inlier"])
