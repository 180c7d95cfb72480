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
