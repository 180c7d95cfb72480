"""fp64 numpy LM engine for the CPU-side distributed tests (TEST INFRASTRUCTURE).

Plugs oracle/lm.py into the product's LMSequencer (skiing_analysis_pytorch_b200/ba.py) so the
multi-rank plumbing - frame sharding, the packed all-reduce payloads `red` / `red2`, the global
sum of confidences, identical controller decisions on every rank - runs over gloo on the CPU-only
box with exactly the payload layout the CUDA engine uses (include/ska.h SkaBaProblem.d_red).
"""
import numpy as np
import torch

from oracle import geometry as G
from oracle import lm
from skiing_analysis_pytorch_b200 import _cabi
from skiing_analysis_pytorch_b200.ba import LMSequencer
from skiing_analysis_pytorch_b200.ba_reg import RegLMSequencer


class OracleBundleAdjuster(LMSequencer):
    def __init__(self, x2d, conf, K, R0, t0, X0, mode="full", lam0=1e-3, max_iters=32, group=None):
        T, C, J, _ = x2d.shape
        self.C, self.N = C, T * J
        self.group, self.max_iters = group, max_iters
        self.X = np.asarray(X0, float).reshape(-1, 3).copy()
        self.R, self.t, self.K = np.asarray(R0, float).copy(), np.asarray(t0, float).copy(), np.asarray(K, float)
        self.x = np.asarray(x2d, float).transpose(0, 2, 1, 3).reshape(self.N, C, 2)
        self.cw = np.asarray(conf, float).transpose(0, 2, 1).reshape(self.N, C)  # RAW conf: the payload is unscaled
        self.free = lm.free_mask(C, mode)
        self.L = _cabi.red_layout(C)
        self.red = torch.zeros(self.L["size"], dtype=torch.float64)
        self.red2 = torch.zeros(_cabi.BA_RED2_DOUBLES, dtype=torch.float64)
        sc = torch.tensor([float(self.cw.sum())], dtype=torch.float64)
        self._allreduce(sc)  # global sum of confidences (loss.py:94 denominator), like the CUDA engine
        self.sumconf = float(sc.item())
        self.lam, self.nu = float(lam0), 2.0
        self.history = []

    # ---- the four steps, same payload layout as ska_ba_*.cu
    def linearize(self):
        lin = lm.linearise(self.X, self.R, self.t, self.K, self.x, self.cw, self.lam)
        L, n = self.L, self.L["n"]
        iu = np.triu_indices(n)
        r = np.zeros(L["size"])
        r[L["sw"]: L["sw"] + len(iu[0])] = lin.Sw[6:, 6:][iu]        # free cameras only (camera 0 is the gauge)
        r[L["bw"]: L["bw"] + n] = lin.bw[6:]
        r[L["gc"]: L["gc"] + n] = lin.gc[1:].reshape(-1)
        i6 = np.triu_indices(6)
        for c in range(1, self.C):
            r[L["hcc"] + 21 * (c - 1): L["hcc"] + 21 * c] = lin.Hcc[c][i6]
        r[L["cost"]] = lin.cost
        r[L["clamp"]] = lin.n_clamped
        self.red.copy_(torch.from_numpy(r))

    def solve(self):
        L, n, C = self.L, self.L["n"], self.C
        r = self.red.numpy()
        s = 1.0 / (self.sumconf + 1e-6)
        Sw = np.zeros((6 * C, 6 * C))
        iu = np.triu_indices(n)
        blk = np.zeros((n, n))
        blk[iu] = r[L["sw"]: L["sw"] + len(iu[0])]
        blk = blk + np.triu(blk, 1).T
        Sw[6:, 6:] = blk
        Hcc = np.zeros((C, 6, 6))
        i6 = np.triu_indices(6)
        for c in range(1, C):
            h = np.zeros((6, 6))
            h[i6] = r[L["hcc"] + 21 * (c - 1): L["hcc"] + 21 * c]
            Hcc[c] = h + np.triu(h, 1).T
        gc = np.zeros((C, 6))
        gc[1:] = r[L["gc"]: L["gc"] + n].reshape(C - 1, 6)
        bw = np.zeros(6 * C)
        bw[6:] = r[L["bw"]: L["bw"] + n]
        lin = lm.Linearisation(Hcc * s, gc * s, Sw * s, bw * s, r[L["cost"]] * s, int(r[L["clamp"]]))
        self.delta, self.pred_cam, self.ok = lm.solve_reduced(lin, self.lam, self.free)
        self.F, self.ncl = lin.cost, lin.n_clamped
        self.Rn, self.tn = lm.apply_camera_step(self.R, self.t, self.delta)

    def backsub(self):
        dp, pred_pts = lm.back_substitute(self.X, self.R, self.t, self.K, self.x, self.cw, self.lam, self.delta)
        self.Xn = self.X + dp
        c, k = lm.cost_only(self.Xn, self.Rn, self.tn, self.K, self.x, self.cw)
        self.red2.copy_(torch.tensor([c, pred_pts, float(k), 0.0], dtype=torch.float64))

    def control(self):
        s = 1.0 / (self.sumconf + 1e-6)
        Ft = float(self.red2[0]) * s
        pred = self.pred_cam + float(self.red2[1]) * s
        rho = (self.F - Ft) / pred if pred > 0 else 0.0
        accepted = bool(self.ok and np.isfinite(Ft) and Ft < self.F)
        self.history.append(dict(iter=len(self.history), cost=self.F, trial_cost=Ft, lam=self.lam, rho=rho, accepted=accepted,
                                 n_clamped=self.ncl, pred=pred))
        self.lam, self.nu = lm.nielsen_update(self.lam, self.nu, rho, accepted)
        if accepted:
            self.X, self.R, self.t = self.Xn, self.Rn, self.tn


__all__ = ["OracleBundleAdjuster", "G"]


class OracleCalibratingBundleAdjuster(LMSequencer):
    """The calibrating BA (free intrinsics + distortion, oracle/lm_calib.py) behind the product's sequencer with the
    CUDA engine's payload layout (include/ska.h: Sw triangle over n = 15 C - 6 parameters, then one 160-double block per
    camera = upper triangle of the 17 x 17 row products, slot 153 = clamp count)."""

    def __init__(self, x2d, conf, theta, R0, t0, X0, free=None, lam0=1e-3, max_iters=32, group=None, prior_theta=None, prior_rho=None):
        from oracle import lm_calib as lc

        self.lc = lc
        T, C, J, _ = x2d.shape
        self.C, self.N = C, T * J
        self.group, self.max_iters = group, max_iters
        self.X = np.asarray(X0, float).reshape(-1, 3).copy()
        self.R, self.t, self.th = np.asarray(R0, float).copy(), np.asarray(t0, float).copy(), np.asarray(theta, float).copy()
        self.x = np.asarray(x2d, float).transpose(0, 2, 1, 3).reshape(self.N, C, 2)
        self.cw = np.asarray(conf, float).transpose(0, 2, 1).reshape(self.N, C)
        self.free = lc.free_mask(C) if free is None else np.asarray(free, bool)
        self.pth = self.th.copy() if prior_theta is None else np.asarray(prior_theta, float)
        self.rho = np.zeros((C, lc.NI)) if prior_rho is None else np.broadcast_to(np.asarray(prior_rho, float), (C, lc.NI)).copy()
        self.L = _cabi.calib_red_layout(C)
        self.red = torch.zeros(self.L["size"], dtype=torch.float64)
        self.red2 = torch.zeros(_cabi.BA_RED2_DOUBLES, dtype=torch.float64)
        sc = torch.tensor([float(self.cw.sum())], dtype=torch.float64)
        self._allreduce(sc)
        self.sumconf = float(sc.item())
        self.lam, self.nu = float(lam0), 2.0
        self.history = []
        P = lc.P
        self.cols = np.array([P * c + r for c in range(C) for r in range(P) if not (c == 0 and r < 6)])

    def linearize(self):
        lc, L, P = self.lc, self.L, self.lc.P
        e, A, B, clamped = lc.residual_blocks(self.X, self.R, self.t, self.th, self.x)
        lin = lc.linearise(self.X, self.R, self.t, self.th, self.x, self.cw, self.lam)
        r = np.zeros(L["size"])
        n = L["n"]
        r[: n * (n + 1) // 2] = lin.Sw[np.ix_(self.cols, self.cols)][np.triu_indices(n)]
        for c in range(self.C):
            blk = np.zeros(_cabi.BA_CALIB_CAM_BLOCK)
            for a in range(P):
                for b in range(a, P):
                    blk[_cabi.calib_tri(a, b)] = lin.Hcc[c, a, b]
                blk[_cabi.calib_tri(a, 15)] = lin.gc[c, a]
                blk[_cabi.calib_tri(a, 16)] = lin.bw[P * c + a]
            blk[_cabi.calib_tri(15, 15)] = float((self.cw[:, c, None] * e[:, c] ** 2).sum())
            blk[153] = float(clamped[:, c].sum())
            r[L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * c: L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * (c + 1)] = blk
        self.red.copy_(torch.from_numpy(r))

    def solve(self):
        lc, L, P, C = self.lc, self.L, self.lc.P, self.C
        r = self.red.numpy()
        s = 1.0 / (self.sumconf + 1e-6)
        n = L["n"]
        S = np.zeros((n, n))
        S[np.triu_indices(n)] = r[: n * (n + 1) // 2]
        S = S + np.triu(S, 1).T
        Sw = np.zeros((P * C, P * C))
        Sw[np.ix_(self.cols, self.cols)] = S
        Hcc, gc, bw = np.zeros((C, P, P)), np.zeros((C, P)), np.zeros(P * C)
        cost = ncl = 0.0
        for c in range(C):
            blk = r[L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * c: L["cam"] + _cabi.BA_CALIB_CAM_BLOCK * (c + 1)]
            for a in range(P):
                for b in range(a, P):
                    Hcc[c, a, b] = Hcc[c, b, a] = blk[_cabi.calib_tri(a, b)]
                gc[c, a] = blk[_cabi.calib_tri(a, 15)]
                bw[P * c + a] = blk[_cabi.calib_tri(a, 16)]
            cost += blk[_cabi.calib_tri(15, 15)]
            ncl += blk[153]
        lin = lc.Linearisation(Hcc * s, gc * s, Sw * s, bw * s, cost * s, int(ncl))
        self.delta, self.pred_cam, self.ok = lc.solve_reduced(lin, self.lam, self.free, self.th, self.pth, self.rho)
        self.F, self.ncl = lin.cost + lc.prior_cost(self.th, self.pth, self.rho), lin.n_clamped
        self.Rn, self.tn, self.thn = lc.apply_camera_step(self.R, self.t, self.th, self.delta)

    def backsub(self):
        lc = self.lc
        dp, pred_pts = lc.back_substitute(self.X, self.R, self.t, self.th, self.x, self.cw, self.lam, self.delta)
        self.Xn = self.X + dp
        c, k = lc.cost_only(self.Xn, self.Rn, self.tn, self.thn, self.x, self.cw)
        self.red2.copy_(torch.tensor([c, pred_pts, float(k), 0.0], dtype=torch.float64))

    def control(self):
        lc = self.lc
        s = 1.0 / (self.sumconf + 1e-6)
        Ft = float(self.red2[0]) * s + lc.prior_cost(self.thn, self.pth, self.rho)
        pred = self.pred_cam + float(self.red2[1]) * s
        rho = (self.F - Ft) / pred if pred > 0 else 0.0
        accepted = bool(self.ok and np.isfinite(Ft) and Ft < self.F)
        self.history.append(dict(iter=len(self.history), cost=self.F, trial_cost=Ft, lam=self.lam, rho=rho, accepted=accepted,
                                 n_clamped=self.ncl, pred=pred))
        self.lam, self.nu = lm.nielsen_update(self.lam, self.nu, rho, accepted)
        if accepted:
            self.X, self.R, self.t, self.th = self.Xn, self.Rn, self.tn, self.thn


class OracleRegularisedBundleAdjuster(RegLMSequencer):
    """fp64 numpy engine of the regularised LM (oracle/lm_reg.py) behind the product's RegLMSequencer
    (skiing_analysis_pytorch_b200/ba_reg.py), with the CUDA engine's payloads: the 40-double sums rows (include/ska.h
    SKA_BA_REG_SUMS; same slots as csrc/ska_ba_reg.cu), the CG dot scalar, and one-frame halo rows of the CG direction and
    of the state.  A rank multiplies only ITS rows of the damped normal matrix with [halo | own | halo] entries of p - a
    wrong or missing halo exchange changes the trajectory.  (To obtain its rows the test engine assembles the global
    system from an all-gather of the current state; that gather is test scaffolding, not part of the sequencer.)"""

    SUM_REPROJ, SUM_CLAMP, SUM_TEMP, SUM_SMOOTH, SUM_B, SUM_B2, SUM_PRED, SUM_L, SUM_L2 = 0, 1, 2, 3, 4, 5, 6, 8, 24

    def __init__(self, x2d, conf, K, R, t, X0, frame_range, global_obs, mode="pose_only", lam0=1e-3, max_iters=16, cg_iters=60,
                 cg_tol=1e-10, group=None, weights=None):
        import torch.distributed as dist

        from oracle import lm_reg

        self.lr = lm_reg
        self.group, self.max_iters, self.cg_iters, self.mode = group, max_iters, cg_iters, mode
        self.a, self.b = frame_range
        self.gx2d, self.gconf = (np.asarray(v, float) for v in global_obs)
        self.Tg = self.gx2d.shape[0]
        Tl, C, J, _ = x2d.shape
        self.Tl, self.C, self.J = Tl, C, J
        self.x2d, self.conf, self.K = np.asarray(x2d, float), np.asarray(conf, float), np.asarray(K, float)
        self.kf = {"pose_only": 0, "pose_cam_t": 3, "full": 6}[mode]
        self.nfree = 3 * J + C * self.kf
        self.state = torch.zeros((2, Tl + 2, 3 * J + 12 * C), dtype=torch.float64)
        self.state[0, 1:-1, : 3 * J] = torch.from_numpy(np.asarray(X0, float).reshape(Tl, -1))
        self.state[0, 1:-1, 3 * J:] = torch.from_numpy(np.concatenate([np.asarray(R, float).reshape(Tl, C, 9), np.asarray(t, float)], -1).reshape(Tl, -1))
        self.state[1] = self.state[0]
        self.cur = 0
        self.p = torch.zeros((Tl + 2, self.nfree), dtype=torch.float64)
        self.sums = torch.zeros((2, 40), dtype=torch.float64)
        self.dot = torch.zeros(1, dtype=torch.float64)
        self.coef = lm_reg.coefficients(self.Tg, J, C, float(self.gconf.sum()), weights)
        self.bones = lm_reg.bones_for(J)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lam, self.nu, self.tol2 = float(lam0), 2.0, cg_tol ** 2
        self.history, self.iters_done, self._graph = [], 0, None
        self.done = False
        self.setup_cost()

    # ---- state access
    def _unpack(self, buf, halo=False):
        s = self.state[buf].numpy()
        s = s if halo else s[1:-1]
        n = s.shape[0]
        X = s[:, : 3 * self.J].reshape(n, self.J, 3)
        cam = s[:, 3 * self.J:].reshape(n, self.C, 12)
        return X, cam[..., :9].reshape(n, self.C, 3, 3), cam[..., 9:]

    def edges(self, kind):
        if kind == "p":
            return self.p[1], self.p[self.Tl], self.p[0], self.p[self.Tl + 1]
        s = self.state[self.cur if kind == "current" else 1 - self.cur]
        return s[1], s[self.Tl], s[0], s[self.Tl + 1]

    # ---- steps
    def cost(self, which):
        buf = self.cur if which == 0 else 1 - self.cur
        X, R, t = self._unpack(buf, halo=True)
        has_next = self.rank < self.world - 1
        own = slice(1, self.Tl + 1)
        _, uv, _, clamped = self.lr.project(X[own], R[own], t[own], self.K)
        out = np.zeros(40)
        out[self.SUM_REPROJ] = (self.conf * ((uv - self.x2d) ** 2).sum(-1)).sum()
        out[self.SUM_CLAMP] = clamped.sum()
        hi = self.Tl + 2 if has_next else self.Tl + 1  # pairs (t, t+1) owned by t
        out[self.SUM_TEMP] = ((X[2:hi] - X[1:hi - 1]) ** 2).sum()
        Cc = self.lr.centres(R, t)
        out[self.SUM_SMOOTH] = ((Cc[2:hi] - Cc[1:hi - 1]) ** 2).sum()
        for k, (i, j) in enumerate(self.bones):
            L = np.linalg.norm(X[own][:, i] - X[own][:, j], axis=-1)
            out[self.SUM_L + k], out[self.SUM_L2 + k] = L.sum(), (L ** 2).sum()
        if self.C >= 2:
            bl = np.linalg.norm(Cc[own][:, 0] - Cc[own][:, 1], axis=-1)
            out[self.SUM_B], out[self.SUM_B2] = bl.sum(), (bl ** 2).sum()
        if which == 1:
            out[self.SUM_PRED] = self.pred_local
        self.sums[which] = torch.from_numpy(out)

    def _F(self, s):
        c = self.coef
        bone = sum(s[self.SUM_L2 + k] - s[self.SUM_L + k] ** 2 / self.Tg for k in range(len(self.bones)))
        base = s[self.SUM_B2] - s[self.SUM_B] ** 2 / self.Tg if self.C >= 2 else 0.0
        return (c["reproj"] * s[self.SUM_REPROJ] + c["smooth"] * s[self.SUM_SMOOTH] + c["baseline"] * base + c["bone_length"] * bone
                + c["pose_temporal"] * s[self.SUM_TEMP])

    def finish_cost(self, which):
        s = self.sums[which].numpy()
        if which == 0:
            self.F, self.ncl = float(self._F(s)), int(s[self.SUM_CLAMP])
        else:
            self.Ft, self.pred = float(self._F(s)), float(s[self.SUM_PRED])

    def _gather_global(self):
        import torch.distributed as dist

        mine = self.state[self.cur, 1:-1].numpy().copy()
        if self.world == 1:
            return mine
        parts = [None] * self.world
        dist.all_gather_object(parts, mine, group=self.group)
        return np.concatenate(parts, 0)

    def linearize(self):
        lr, J, C, kf = self.lr, self.J, self.C, self.kf
        g_state = self._gather_global()
        X = g_state[:, : 3 * J].reshape(self.Tg, J, 3)
        cam = g_state[:, 3 * J:].reshape(self.Tg, C, 12)
        H, g, free, _ = lr.normal_system(X, cam[..., :9].reshape(self.Tg, C, 3, 3), cam[..., 9:], self.K, self.gx2d, self.gconf, self.coef, self.mode)
        import scipy.sparse as sp

        D = H.diagonal()
        A = (H + self.lam * sp.diags(D)).tocsr()
        nX = self.Tg * J * 3

        def cols(f0, f1):  # free-column indices of frames [f0, f1), frame-row order [3J | C kf]
            out = []
            for f in range(f0, f1):
                out.append(np.concatenate([np.arange(f * J * 3, (f + 1) * J * 3), nX + np.arange(f * C * kf, (f + 1) * C * kf)]))
            return out

        own = cols(self.a, self.b)
        self.own_idx = np.concatenate(own)
        lo, hi = max(self.a - 1, 0), min(self.b + 1, self.Tg)
        self.ext_idx = np.concatenate(cols(lo, hi))
        self.ext_rows = slice(1 - (self.a - lo), self.Tl + 1 + (hi - self.b))  # rows of [halo | own | halo] that exist
        self.A_rows = A[self.own_idx][:, self.ext_idx]
        self.Minv = [np.linalg.inv(A[ix][:, ix].toarray()) for ix in own]
        self.g_own, self.D_own = g[self.own_idx], D[self.own_idx]

    def _prec(self, r):
        return np.concatenate([Mi @ r[k * self.nfree:(k + 1) * self.nfree] for k, Mi in enumerate(self.Minv)])

    def cg(self, op):
        from skiing_analysis_pytorch_b200 import _cabi as k

        if op == k.BA_REG_CG_BEGIN:
            self.x = np.zeros_like(self.g_own)
            self.r = -self.g_own
            self.p.zero_()
            self.z = self._prec(self.r)
            self.dot[0] = float(self.r @ self.z)
            self.done, self.cg_used = False, 0
        elif op == k.BA_REG_CG_INIT:
            self.rz0 = self.rz = float(self.dot)
            self.beta, self.alpha = 0.0, 0.0
            self.done = not (self.rz0 > 0)
        elif op == k.BA_REG_CG_DIR:
            if self.done and self.cg_used > 0:
                return
            pn = self.z + self.beta * self.p[1:-1].numpy().reshape(-1)
            self.p[1:-1] = torch.from_numpy(pn.reshape(self.Tl, self.nfree))
        elif self.done:
            return
        elif op == k.BA_REG_CG_MATVEC:
            p_ext = self.p.numpy()[self.ext_rows].reshape(-1)
            self.y = self.A_rows @ p_ext
            self.dot[0] = float(self.p[1:-1].numpy().reshape(-1) @ self.y)
        elif op == k.BA_REG_CG_ALPHA:
            pap = float(self.dot)
            self.alpha = self.rz / pap
        elif op == k.BA_REG_CG_UPDATE:
            self.x = self.x + self.alpha * self.p[1:-1].numpy().reshape(-1)
            self.r = self.r - self.alpha * self.y
            self.z = self._prec(self.r)
            self.dot[0] = float(self.r @ self.z)
        elif op == k.BA_REG_CG_BETA:
            v = float(self.dot)
            self.beta, self.rz = v / self.rz, v
            self.cg_used += 1
            if not (v > self.tol2 * self.rz0):
                self.done = True

    def apply(self):
        x = self.x.reshape(self.Tl, self.nfree)
        self.pred_local = float(self.x @ (self.lam * self.D_own * self.x - self.g_own))
        X, R, t = self._unpack(self.cur)
        delta = np.zeros(self.Tl * self.J * 3 + self.Tl * self.C * 6)
        delta[: self.Tl * self.J * 3] = x[:, : 3 * self.J].reshape(-1)
        dc = np.zeros((self.Tl, self.C, 6))
        if self.kf == 3:
            dc[..., 3:] = x[:, 3 * self.J:].reshape(self.Tl, self.C, 3)
        elif self.kf == 6:
            dc[:] = x[:, 3 * self.J:].reshape(self.Tl, self.C, 6)
        delta[self.Tl * self.J * 3:] = dc.reshape(-1)
        Xn, Rn, tn = self.lr.apply_step(X.copy(), R.copy(), t.copy(), delta, self.mode)
        s = self.state[1 - self.cur]
        s[1:-1, : 3 * self.J] = torch.from_numpy(Xn.reshape(self.Tl, -1))
        s[1:-1, 3 * self.J:] = torch.from_numpy(np.concatenate([Rn.reshape(self.Tl, self.C, 9), tn], -1).reshape(self.Tl, -1))

    def control(self):
        rho = (self.F - self.Ft) / self.pred if self.pred > 0 else 0.0
        accepted = bool(np.isfinite(self.Ft) and self.Ft < self.F)
        self.history.append(dict(iter=len(self.history), cost=self.F, trial_cost=self.Ft, lam=self.lam, rho=rho, accepted=accepted,
                                 pred=self.pred, cg_iters=self.cg_used, n_clamped=self.ncl))
        if accepted:
            self.lam, self.nu = self.lam * max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 2.0
            self.cur = 1 - self.cur
            self.sums[0] = self.sums[1]
            self.F, self.ncl = self.Ft, int(self.sums[1][self.SUM_CLAMP])
        else:
            self.lam, self.nu = self.lam * self.nu, 2.0 * self.nu

    @property
    def X(self):
        return self._unpack(self.cur)[0]
