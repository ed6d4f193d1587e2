"""CPU stand-in for ffi.GpuBaSession (same phase interface, numpy + the C oracle's edge functions), used by the
gloo tests to exercise the landmark-sharded LM driver (stereo-visual-slam_b200/sharding.py) without a GPU."""
import numpy as np

from oracle import ba_oracle as B


class CpuBaSession:
    BUILD, IMPORT_BUILD, SCHUR, SOLVE_UPDATE, RELABEL_COUNT, RELABEL_APPLY = 1, 2, 3, 4, 5, 6

    def __init__(self, problem, shard, r1, r2, r3, huber_delta=5.991, chi2_th=5.991):
        self.K = np.asarray(problem["K"], float)
        self.poses = [np.array(problem["poses"], float).reshape(-1, 12).copy() for _ in range(2)]
        self.points = [np.array(problem["points"], float).reshape(-1, 3).copy() for _ in range(2)]
        self.op, self.ol = np.asarray(problem["obs_pose"]), np.asarray(problem["obs_point"])
        self.uv = np.asarray(problem["obs_uv"], float).reshape(-1, 2)
        self.l0, self.l1 = shard
        self.mine = np.nonzero((self.ol >= self.l0) & (self.ol < self.l1))[0]
        self.r1, self.r2, self.r3 = r1, r2, r3     # torch CPU tensors (float64)
        self.nk, self.delta, self.th0 = len(self.poses[0]), huber_delta, chi2_th
        self.cur = 0
        self.err = np.zeros((len(self.op), 2))
        self.chi2_out = np.zeros(len(self.op)); self.inl = np.zeros(len(self.points[0]), np.uint8)

    def _huber(self, e2):
        d2 = self.delta ** 2
        return (e2, 1.0) if e2 <= d2 else (2 * np.sqrt(e2) * self.delta - d2, self.delta / np.sqrt(e2))

    def phase(self, ph, value=0.0):
        nk, n = self.nk, 6 * self.nk
        P, X = self.poses[self.cur], self.points[self.cur]
        if ph == self.BUILD:
            Hpp = np.zeros((nk, 6, 6)); bp = np.zeros(n); chi = 0.0
            self.Hll, self.bl, self.Hpl = {}, {}, {}
            for i in self.mine:
                k, l = self.op[i], self.ol[i]
                e, A, Bm = B.residual_and_jacobians(P[k], X[l], self.K, self.uv[i])
                self.err[i] = e
                r0, w = self._huber(e @ e)
                chi += r0
                Hpp[k] += w * A.T @ A; bp[6 * k:6 * k + 6] += -w * A.T @ e
                self.Hll[l] = self.Hll.get(l, 0) + w * Bm.T @ Bm
                self.bl[l] = self.bl.get(l, 0) + -w * Bm.T @ e
                self.Hpl[i] = w * A.T @ Bm
            md = max([np.abs(np.diag(h)).max() for h in self.Hll.values()], default=0.0)
            self.r1[:36 * nk] = self._t(Hpp.ravel()); self.r1[36 * nk:42 * nk] = self._t(bp)
            self.r1[42 * nk] = chi; self.r1[42 * nk + 1] = md
        elif ph == self.SCHUR:
            lam = value
            S = np.zeros((n, n)); bs = np.zeros(n)
            self.Dinv = {}
            by_l = {}
            for i in self.mine:
                by_l.setdefault(self.ol[i], []).append(i)
            self.by_l = by_l
            for l, obs in by_l.items():
                Di = np.linalg.inv(self.Hll[l] + lam * np.eye(3)); self.Dinv[l] = Di
                for i in obs:
                    ki = self.op[i]
                    bs[6 * ki:6 * ki + 6] -= self.Hpl[i] @ Di @ self.bl[l]
                    for j in obs:
                        kj = self.op[j]
                        S[6 * ki:6 * ki + 6, 6 * kj:6 * kj + 6] -= self.Hpl[i] @ Di @ self.Hpl[j].T
            self.r2[:n * n] = self._t(S.ravel()); self.r2[n * n:] = self._t(bs)
        elif ph == self.SOLVE_UPDATE:
            lam = value
            r1 = self.r1.numpy(); r2 = self.r2.numpy()
            Hpp = r1[:36 * nk].reshape(nk, 6, 6); bp = r1[36 * nk:42 * nk]
            S = r2[:n * n].reshape(n, n).copy(); bs = r2[n * n:].copy() + bp
            for k in range(nk):
                S[6 * k:6 * k + 6, 6 * k:6 * k + 6] += Hpp[k]
            S += lam * np.eye(n)
            xp = np.linalg.solve(S, bs)
            T, Y = self.poses[1 - self.cur], self.points[1 - self.cur]
            scale = 0.0
            for k in range(nk):
                dR, dt = B.se3_exp(xp[6 * k:6 * k + 6])
                T4 = P[k].reshape(3, 4)
                T[k] = np.hstack([dR @ T4[:, :3], (dR @ T4[:, 3] + dt)[:, None]]).reshape(-1)
            if self.l0 == 0:
                scale += float(xp @ (lam * xp + bp))
            Y[:] = X
            for l, obs in self.by_l.items():
                c = self.bl[l].copy()
                for i in obs:
                    c -= self.Hpl[i].T @ xp[6 * self.op[i]:6 * self.op[i] + 6]
                xl = self.Dinv[l] @ c
                Y[l] = X[l] + xl
                scale += float(xl @ (lam * xl + self.bl[l]))
            chi = 0.0
            for i in self.mine:
                e, _, _ = B.residual_and_jacobians(T[self.op[i]], Y[self.ol[i]], self.K, self.uv[i])
                self.err[i] = e
                chi += self._huber(e @ e)[0]
            self.r3[0] = chi; self.r3[1] = scale; self.r3[2] = 1.0
        elif ph == self.RELABEL_COUNT:
            c2 = (self.err[self.mine] ** 2).sum(1)
            for r in range(6):
                self.r3[4 + r] = float((c2 <= self.th0 * 2 ** r).sum())
        elif ph == self.RELABEL_APPLY:
            th = value
            for i in self.mine:
                c = float(self.err[i] @ self.err[i])
                self.chi2_out[i] = c
                self.inl[self.ol[i]] = 0 if c > th else 1   # insertion order: the last edge of a landmark wins

    @staticmethod
    def _t(a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a))

    def trial_done(self, accept):
        if accept:
            self.cur = 1 - self.cur

    def end(self):
        pts = np.zeros_like(self.points[0]); pts[self.l0:self.l1] = self.points[self.cur][self.l0:self.l1]
        return self.poses[self.cur].copy(), pts, self.chi2_out, self.inl
