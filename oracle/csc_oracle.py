"""Oracle for the unrolled convolutional-sparse-coding network (SURVEY §8a part 1: A0-A16), PyTorch on CPU.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference has no tests or fixtures and
cannot run here; this file restates src/model.jl, src/train.jl and src/inference/_1_code_retrieval.jl.

Two restatements that are tested against each other (tests/test_csc_oracle.py):
  * `forward_literal`  — conv-by-conv transcription of model.jl on Julia-ordered arrays, through the NNlib 0.9.8
    conventions (source not in /root/reference, published behaviour): conv(x,w) is a true convolution (kernel
    reversed), flipped=true is cross-correlation, weights are (kW[,kH], Cin/groups, Cout), pad pads both ends,
    groups split channels, batched_mul multiplies per trailing index, upsample_nearest repeats entries.
  * `forward_pos`      — the index-level ("position space") form of SURVEY Appendix B that the CUDA kernels
    implement: only rows 4p of Z,Y are ever non-zero (z_mask_n), so the network lives on c = Lb-7 positions.
Gradients come from torch.autograd (stand-in for Zygote 0.6.67; masks/top-q/duals are constants exactly where the
reference wraps them in @ignore; relu'(0) = 0 in both).

Flat parameter vector = Flux.params(cdl) order (model.jl:67-137 field order, arrays only):
  lambda_sparsity[6], kappa_sparsity[3], lambda_stepsize[6], omega_stepsize[6], kappa_stepsize[3],
  D[32*M] (Julia (32,1,M) column-major: k + 32 m), F[h*2M*K] (Julia (h,2M,1,K): a + h (j + 2M k)),
  penalty_xyz[6], mu[3]   (= 30433 floats by default)  followed by the three NON-trainable warm-up scalars
  lambda_sparsity_warmup, lambda_stepsize_warmup, omega_stepsize_warmup.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as Fn


@dataclass
class Hyperparam:                      # model.jl:1-14
    filter_len: int = 8
    M: int = 50
    h: int = 12
    K: int = 24
    q: int = 32
    batch_size: int = 6
    num_pass_xyz: int = 6
    num_pass_df: int = 3
    magnifying_factor: float = 10.0
    gamma: float = 0.1                 # unused by the reference too (model.jl:13)

    @property
    def f_len(self):
        return 4 * self.filter_len

    @property
    def twoM(self):
        return 2 * self.M


@dataclass
class LengthInfo:                      # model.jl:16-37
    L: int
    C: int
    c: int
    l: int
    CS_vlen: int

    @staticmethod
    def of(hp: Hyperparam, Lb: int):
        L = 4 * Lb
        C = L - hp.f_len + 1
        c = Lb - hp.filter_len + 1
        return LengthInfo(L, C, c, c - hp.h + 1, C + L - 1)


PARAM_FIELDS = ["lambda_sparsity", "kappa_sparsity", "lambda_stepsize", "omega_stepsize", "kappa_stepsize", "D", "F",
                "penalty_xyz", "mu"]
WARM_FIELDS = ["lambda_sparsity_warmup", "lambda_stepsize_warmup", "omega_stepsize_warmup"]


def param_sizes(hp: Hyperparam):
    return {"lambda_sparsity": hp.num_pass_xyz, "kappa_sparsity": hp.num_pass_df, "lambda_stepsize": hp.num_pass_xyz,
            "omega_stepsize": hp.num_pass_xyz, "kappa_stepsize": hp.num_pass_df, "D": hp.f_len * hp.M,
            "F": hp.h * hp.twoM * hp.K, "penalty_xyz": hp.num_pass_xyz, "mu": hp.num_pass_df}


def n_params(hp):
    return sum(param_sizes(hp).values())


def init_params(hp: Hyperparam, seed: int) -> np.ndarray:
    """ucdl(hp) (model.jl:84-106) + randomly_initialize_filters (MOTIFs.jl:17-33) with numpy's PCG64 (the
    reference uses Julia's global RNG; the stream is ours, the distributions are the reference's).
    Returns the flat vector (n_params + 3 warm-up scalars), float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    eta1 = np.float32(0.05)
    M, fl = hp.M, hp.filter_len
    D = np.zeros((4 * fl, M), np.float32)                       # D[4 i + a, m]
    for i in range(fl):
        for j in range(M):
            u = np.sort(rng.random(3))
            arr = np.concatenate([[0.0], u, [1.0]])
            D[4 * i:4 * i + 4, j] = np.diff(arr).astype(np.float32)
    D = np.sqrt(D)
    Fw = np.abs(np.float32(0.1) * rng.standard_normal((hp.K, hp.twoM, hp.h)).astype(np.float32))   # [k][j][a] = Julia order reversed
    p = {}
    lsw = eta1 * np.float32(rng.random())
    p["lambda_sparsity"] = eta1 * rng.random(hp.num_pass_xyz).astype(np.float32)
    p["kappa_sparsity"] = eta1 * rng.random(hp.num_pass_df).astype(np.float32)
    lstw = eta1 * np.float32(rng.random())
    osw = eta1 * np.float32(rng.random())
    p["lambda_stepsize"] = eta1 * rng.random(hp.num_pass_xyz).astype(np.float32)
    p["omega_stepsize"] = eta1 * rng.random(hp.num_pass_xyz).astype(np.float32)
    p["kappa_stepsize"] = eta1 * rng.random(hp.num_pass_df).astype(np.float32)
    p["penalty_xyz"] = eta1 * rng.random(hp.num_pass_xyz).astype(np.float32)
    p["mu"] = eta1 * rng.random(hp.num_pass_df).astype(np.float32)
    p["D"] = D.T.reshape(-1)                                     # k + 32 m  -> m-major blocks of 32
    p["F"] = Fw.reshape(-1)                                      # a + h (j + 2M k)
    flat = np.concatenate([np.asarray(p[f], np.float32).reshape(-1) for f in PARAM_FIELDS] + [np.array([lsw, lstw, osw], np.float32)])
    return flat.astype(np.float32)


def unpack(flat: torch.Tensor, hp: Hyperparam):
    """flat vector -> dict of tensors (views, so autograd flows to `flat`).  D -> (32, M) [k, m]; F -> (h, 2M, K) [a, j, k]."""
    out, o = {}, 0
    for f, n in param_sizes(hp).items():
        out[f] = flat[o:o + n]
        o += n
    for f in WARM_FIELDS:
        out[f] = flat[o]
        o += 1
    out["D"] = out["D"].reshape(hp.M, hp.f_len).t()
    out["F"] = out["F"].reshape(hp.K, hp.twoM, hp.h).permute(2, 1, 0)
    return out


def onehot_from_codes(codes: np.ndarray, dtype=torch.float32) -> torch.Tensor:
    """(B, Lb) codes -> S (4Lb, 1, B) in Julia index order (row 4p + a)."""
    B, Lb = codes.shape
    S = torch.zeros(B, Lb, 4, dtype=dtype)
    S.scatter_(2, torch.as_tensor(codes.astype(np.int64))[:, :, None], 1.0)
    return S.reshape(B, 4 * Lb).t().reshape(4 * Lb, 1, B).contiguous()


# ==================================================================================================
# shared selections (constants for AD): batch median mask (model.jl:194-204) and top-q (model.jl:181-192)
# ==================================================================================================
def median_of_positives(v: torch.Tensor):
    """Statistics.median(ZY[ZY .> 0]); None when empty.  Even count: middle(a,b) = a/2 + b/2."""
    nz = v[v > 0]
    n = nz.numel()
    if n == 0:
        return None
    s, _ = torch.sort(nz.reshape(-1))
    if n % 2 == 1:
        return s[n // 2]
    return s[n // 2 - 1] / 2 + s[n // 2] / 2


def qth_largest(v: torch.Tensor, q: int):
    """partialsort(col, q, rev=true)."""
    return torch.topk(v.reshape(-1), q, largest=True, sorted=True).values[q - 1]


# ==================================================================================================
# 1. literal transcription (Julia index order, NNlib conventions)
# ==================================================================================================
def _jconv1d(x, w, pad=0, flipped=False, groups=1):
    xt, wt = x.permute(2, 1, 0), w.permute(2, 1, 0)
    if not flipped:
        wt = wt.flip(-1)
    return Fn.conv1d(xt, wt, padding=pad, groups=groups).permute(2, 1, 0)


def _jconv2d(x, w, pad=(0, 0), flipped=False, groups=1):
    xt, wt = x.permute(3, 2, 1, 0), w.permute(3, 2, 1, 0)
    if not flipped:
        wt = wt.flip(-1, -2)
    return Fn.conv2d(xt, wt, padding=(pad[1], pad[0]), groups=groups).permute(3, 2, 1, 0)


def _jreshape(x, shape):
    """Julia (column-major) reshape of a tensor held in Julia index order."""
    nd = x.dim()
    y = x.permute(*reversed(range(nd))).reshape(*reversed(shape))
    return y.permute(*reversed(range(len(shape))))


class Literal:
    """model.jl, function by function, same names."""

    def __init__(self, hp: Hyperparam, Lb: int, dtype=torch.float32):
        self.hp, self.len, self.dt = hp, LengthInfo.of(hp, Lb), dtype
        ln = self.len
        B = hp.batch_size
        self.mapdrange = torch.zeros(hp.f_len, ln.C + ln.L - 1, dtype=dtype)                 # model.jl:46-47
        self.mapdrange[:, ln.C - 1:ln.C - 1 + hp.f_len] = torch.eye(hp.f_len, dtype=dtype)
        self.mapclarge = torch.zeros(ln.C, ln.c, dtype=dtype)                               # :50-51
        self.mapclarge[0::4, :] = torch.eye(ln.c, dtype=dtype)
        zcol = torch.zeros(ln.C, dtype=dtype)
        zcol[0::4] = 1
        self.z_mask_n = zcol[:, None, None].expand(ln.C, hp.M, B)                            # :54-55
        self.pseudocount = torch.full((4, hp.filter_len, hp.M), 0.001, dtype=dtype)          # :56

    # prep (:139-169) — every scalar squared
    def prep_filters(self, D):
        hp = self.hp
        Dr = _jreshape(D ** 2, (4, hp.filter_len, hp.M)) + self.pseudocount
        Dr = Dr / Dr.sum(dim=0, keepdim=True)
        return _jreshape(Dr, (hp.f_len, 1, hp.M))

    @staticmethod
    def prep_syntax_filters(F):
        F = F ** 2
        return F / torch.sqrt((F ** 2).sum(dim=(0, 1), keepdim=True))

    def warmup_ZY(self, S, D, eta_w, lam_w):
        DtS = _jconv1d(S, D, 0, flipped=True)
        DS = _jconv1d(S, D, 0)
        Z = torch.relu(self.z_mask_n * (eta_w * DtS - lam_w * eta_w))
        Y = torch.relu(self.z_mask_n * (eta_w * DS - lam_w * eta_w))
        return Z, Y

    def project_X(self, X):
        hp = self.hp
        with torch.no_grad():
            Xr = _jreshape(X, (X.shape[0] * hp.K, hp.batch_size))
            vals = torch.stack([qth_largest(Xr[:, b], hp.q) for b in range(hp.batch_size)]).reshape(1, 1, 1, -1)
            bit = (X >= vals).to(X.dtype)
        return X * bit

    def cat_ZY(self, Z, Y):
        hp, ln = self.hp, self.len
        ZY = _jreshape(torch.cat([Z[0::4], Y[0::4]], dim=1), (ln.c, hp.twoM, 1, hp.batch_size))
        with torch.no_grad():
            med = median_of_positives(ZY)
            mask = None if med is None else (ZY >= med).to(ZY.dtype)
        return hp.magnifying_factor * ZY if mask is None else hp.magnifying_factor * (mask * ZY)

    def FX_of(self, X, F):
        hp = self.hp
        return _jconv2d(X, F, pad=(hp.h - 1, hp.twoM - 1), groups=hp.K).sum(dim=2, keepdim=True)

    def left_right(self, FX):
        hp, ln = self.hp, self.len
        return (_jreshape(FX[:, :hp.M], (ln.c, hp.M, hp.batch_size)), _jreshape(FX[:, hp.M:], (ln.c, hp.M, hp.batch_size)))

    def update_ZY(self, S, Z, Y, D, lFX, rFX, alpha, beta, lam, eta, rho):
        hp = self.hp
        ZD = _jconv1d(Z, D, hp.f_len - 1, groups=hp.M)
        YD = _jconv1d(Y, D, hp.f_len - 1, groups=hp.M, flipped=True)
        diff = (ZD + YD).sum(dim=1, keepdim=True) - S
        bm = lambda v: torch.einsum("Cc,cmb->Cmb", self.mapclarge, v)                        # batched_mul
        zg = _jconv1d(diff, D, 0, flipped=True) + rho * (Z - bm(lFX + alpha))
        yg = _jconv1d(diff, D, 0) + rho * (Y - bm(rFX + beta))
        Zu = Z - eta * zg - lam * eta
        Yu = Y - eta * yg - lam * eta
        return torch.relu(self.z_mask_n * Zu), torch.relu(self.z_mask_n * Yu)

    def update_X(self, FX, Z, Y, X, F, alpha, beta, omega):
        hp, ln = self.hp, self.len
        ab = _jreshape(torch.cat([alpha, beta], dim=1), (ln.c, hp.twoM, 1, hp.batch_size))
        ZY = self.cat_ZY(Z, Y)
        diff = FX.sum(dim=2, keepdim=True) - (ZY - ab)
        xg = _jconv2d(diff, F, flipped=True)
        return self.project_X(X - omega * xg)

    def conv_code_diff(self, code, diff):
        hp, ln = self.hp, self.len
        MB = hp.M * hp.batch_size
        up = diff.expand(ln.L, hp.M, hp.batch_size)                                         # upsample_nearest (1,M,1)
        out = _jconv1d(_jreshape(up, (ln.L, MB, 1)), _jreshape(code, (ln.C, 1, MB)), ln.C - 1, flipped=True, groups=MB)
        return _jreshape(out, (ln.CS_vlen, hp.M, hp.batch_size))

    def update_D(self, S, Z, Y, D, mu):
        hp, ln = self.hp, self.len
        sZD = _jconv1d(Z, D, hp.f_len - 1, groups=hp.M).sum(dim=1, keepdim=True)
        sYRD = _jconv1d(Y, D, hp.f_len - 1, groups=hp.M, flipped=True).sum(dim=1, keepdim=True)
        ccd = self.conv_code_diff
        tot = ccd(Z, sZD) + ccd(Z, sYRD) + ccd(Z, S) + (ccd(Y, sZD) + ccd(Y, sYRD) + ccd(Y, S)).flip(0)   # '+S' as in :285
        Dg = _jreshape(self.mapdrange @ _jreshape(tot.sum(dim=2, keepdim=True), (ln.CS_vlen, hp.M)), (hp.f_len, 1, hp.M))
        Br = _jreshape(D * torch.exp(-mu * Dg), (4, hp.filter_len, 1, hp.M))
        return _jreshape(Br / Br.sum(dim=0, keepdim=True), (hp.f_len, 1, hp.M))

    def F_gradient(self, ZY, X, F, theta):
        hp, ln = self.hp, self.len
        KB = hp.K * hp.batch_size
        d = self.FX_of(X, F) - (ZY + theta)
        d_up = d.expand(ln.c, hp.twoM, hp.K, hp.batch_size)                                  # upsample_nearest (1,1,K,1)
        diff_r = _jreshape(d_up, (ln.c, hp.twoM, KB, 1))
        X_r = _jreshape(X, (ln.l, 1, 1, KB))
        cv = _jconv2d(diff_r, X_r, flipped=True, groups=KB)
        Fc = _jreshape(cv, (hp.h, hp.twoM, hp.K, hp.batch_size))
        return _jreshape(Fc.sum(dim=3, keepdim=True), (hp.h, hp.twoM, 1, hp.K))

    def update_F(self, ZY, X, F, theta, kappa, kappa_s):
        Fu = torch.relu(F - kappa * self.F_gradient(ZY, X, F, theta) - kappa * kappa_s)
        return Fu / torch.sqrt((Fu ** 2).sum(dim=(0, 1), keepdim=True))

    def ADMM_XYZ(self, S, D, F, P):
        hp, ln = self.hp, self.len
        alpha = torch.zeros(ln.c, hp.M, hp.batch_size, dtype=self.dt)
        beta = torch.zeros_like(alpha)
        Z, Y = self.warmup_ZY(S, D, P["lambda_stepsize_warmup"], P["lambda_sparsity_warmup"])
        X = self.project_X(P["omega_stepsize_warmup"] * _jconv2d(self.cat_ZY(Z, Y), F, flipped=True))
        FX = self.FX_of(X, F)
        lFX, rFX = self.left_right(FX)
        for n in range(hp.num_pass_xyz):
            Z, Y = self.update_ZY(S, Z, Y, D, lFX, rFX, alpha, beta, P["lambda_sparsity"][n], P["lambda_stepsize"][n], P["penalty_xyz"][n])
            X = self.update_X(FX, Z, Y, X, F, alpha, beta, P["omega_stepsize"][n])
            FX = self.FX_of(X, F)
            lFX, rFX = self.left_right(FX)
            alpha = alpha + lFX - Z[0::4]
            beta = beta + rFX - Y[0::4]
        return Z, Y, X

    def forward(self, S, flat):
        """forward_pass_return_loss (model.jl:375-395).  Returns (loss, dict of intermediates)."""
        hp = self.hp
        P0 = unpack(flat, hp)
        P = {k: (v ** 2) for k, v in P0.items() if k not in ("D", "F")}
        D = self.prep_filters(P0["D"].reshape(hp.f_len, 1, hp.M))
        F = self.prep_syntax_filters(P0["F"].reshape(hp.h, hp.twoM, 1, hp.K))
        Z, Y, X = self.ADMM_XYZ(S, D, F, P)
        theta = torch.zeros(self.len.c, hp.twoM, 1, hp.batch_size, dtype=self.dt)
        ZY = self.cat_ZY(Z, Y)
        Dc, Fc = D, F
        for n in range(hp.num_pass_df):
            Dc = self.update_D(S, Z, Y, Dc, P["mu"][n])
            Fc = self.update_F(ZY, X, Fc, theta, P["kappa_stepsize"][n], P["kappa_sparsity"][n])
            theta = theta + self.FX_of(X, Fc) - ZY
        nf = 1.0 / hp.batch_size
        DZ = _jconv1d(Z, Dc, hp.f_len - 1, groups=hp.M).sum(dim=1, keepdim=True)
        DY = _jconv1d(Y, Dc, hp.f_len - 1, groups=hp.M, flipped=True).sum(dim=1, keepdim=True)
        rec = nf * ((DZ + DY - S) ** 2).sum()
        syn = nf * ((self.FX_of(X, Fc) - ZY) ** 2).sum()
        return rec + syn, {"Z": Z, "Y": Y, "X": X, "D": Dc, "F": Fc, "ZY": ZY, "rec": rec, "syn": syn}


# ==================================================================================================
# 2. position-space form — the specification of the CUDA kernels
#    layouts: z,y,alpha,beta (B,c,M); zy,fx,theta (B,c,2M); x (B,l,K); signals (B,4Lb); D (32,M); F (h,2M,K)
# ==================================================================================================
def prep_D(Draw, hp):
    D4 = (Draw ** 2 + 0.001).reshape(hp.filter_len, 4, hp.M)
    return (D4 / D4.sum(dim=1, keepdim=True)).reshape(hp.f_len, hp.M)


def prep_F(Fraw):
    F2 = Fraw ** 2
    return F2 / torch.sqrt((F2 ** 2).sum(dim=(0, 1), keepdim=True))


def gather_uf_ur(D, codes_t, hp, c):
    """A3: u_f[b,p,m] = sum_j D[4j + s[p+j], m];  u_r[b,p,m] = sum_j D[4(7-j) + 3 - s[p+j], m]."""
    fl = hp.filter_len
    win = codes_t.unfold(1, fl, 1)[:, :c]                                   # (B, c, fl)
    j = torch.arange(fl)
    uf = D[(4 * j)[None, None, :] + win].sum(dim=2)                         # (B, c, M)
    ur = D[(4 * (fl - 1 - j))[None, None, :] + (3 - win)].sum(dim=2)
    return uf, ur


def recon_pos(z, y, D, Lb):
    """A8/A10/A12: recon[b,t] = sum_m sum_p z[b,p,m] D[t-4p,m] + y[b,p,m] D[31-(t-4p),m]."""
    B, c, M = z.shape
    fl4 = D.shape[0]
    a = z @ D.t() + y @ D.flip(0).t()                                       # (B, c, 32): contribution of position p at lag k
    out = torch.zeros(B, 4 * Lb, dtype=z.dtype)
    idx = (4 * torch.arange(c))[:, None] + torch.arange(fl4)[None, :]        # (c, 32)
    return out.index_add(1, idx.reshape(-1), a.reshape(B, -1))


def corr_sig(r, D):
    """A8: gz[b,p,m] = sum_k r[b,4p+k] D[k,m];  gy[b,p,m] = sum_k r[b,4p+k] D[31-k,m]."""
    win = r.unfold(1, D.shape[0], 4)                                         # (B, c, 32)
    return win @ D, win @ D.flip(0)


def dgrad_pos(z, y, R, fl4):
    """A10: G[t,m] = sum_b sum_p z[b,p,m] R[b,4p+t] + y[b,p,m] R[b,4p+31-t]."""
    win = R.unfold(1, fl4, 4)                                                # (B, c, 32)
    return torch.einsum("bpt,bpm->tm", win, z) + torch.einsum("bpt,bpm->tm", win.flip(2), y)


def corr2d_pos(A, F):
    """A6/A9: out[b,i,k] = sum_{a,j} A[b,i+a,j] F[a,j,k]."""
    h = F.shape[0]
    return torch.einsum("bija,ajk->bik", A.unfold(1, h, 1), F)


def tconv_pos(x, F):
    """A7: fx[b,i,j] = sum_k sum_a x[b,i-a,k] F[a,j,k]   (0 <= i-a < l)."""
    h = F.shape[0]
    xp = Fn.pad(x, (0, 0, h - 1, h - 1))                                     # pad positions
    w = xp.unfold(1, h, 1)                                                   # (B, c, K, h): w[i][t] = xp[i+t] = x[i+t-(h-1)]
    return torch.einsum("bikt,tjk->bij", w, F.flip(0))


def fgrad_pos(e, x):
    """A11: Fg[a,j,k] = sum_b sum_i e[b,a+i,j] x[b,i,k]."""
    l = x.shape[1]
    return torch.einsum("bajl,blk->ajk", e.unfold(1, l, 1), x)


def mask_scale(z, y, mf):
    """A4: zy' = mf * [zy >= median(zy > 0 over the whole batch)] * zy ; mask is a constant."""
    zy = torch.cat([z, y], dim=2)
    with torch.no_grad():
        med = median_of_positives(zy)
        mask = None if med is None else (zy >= med).to(zy.dtype)
    return mf * zy if mask is None else mf * (mask * zy)


def topq(x, q):
    """A5: keep entries >= the q-th largest of each sequence's l*K values; bitmask is a constant."""
    with torch.no_grad():
        v = torch.stack([qth_largest(x[b], q) for b in range(x.shape[0])]).reshape(-1, 1, 1)
        bit = (x >= v).to(x.dtype)
    return x * bit


def admm_xyz_pos(codes_t, D, F, P, hp, Lb, keep=None):
    c = Lb - hp.filter_len + 1
    M = hp.M
    S = torch.zeros(codes_t.shape[0], Lb, 4, dtype=D.dtype).scatter_(2, codes_t[:, :, None], 1.0).reshape(codes_t.shape[0], 4 * Lb)
    eta_w, lam_w, om_w = P["lambda_stepsize_warmup"], P["lambda_sparsity_warmup"], P["omega_stepsize_warmup"]
    uf, ur = gather_uf_ur(D, codes_t, hp, c)
    z = torch.relu(eta_w * uf - lam_w * eta_w)
    y = torch.relu(eta_w * ur - lam_w * eta_w)
    x = topq(om_w * corr2d_pos(mask_scale(z, y, hp.magnifying_factor), F), hp.q)
    fx = tconv_pos(x, F)
    alpha = torch.zeros_like(z)
    beta = torch.zeros_like(z)
    for n in range(hp.num_pass_xyz):
        eta, lam, rho, om = P["lambda_stepsize"][n], P["lambda_sparsity"][n], P["penalty_xyz"][n], P["omega_stepsize"][n]
        r = recon_pos(z, y, D, Lb) - S
        gz, gy = corr_sig(r, D)
        left, right = fx[:, :, :M], fx[:, :, M:]
        z_new = torch.relu(z - eta * (gz + rho * (z - left - alpha)) - lam * eta)
        y_new = torch.relu(y - eta * (gy + rho * (y - right - beta)) - lam * eta)
        z, y = z_new, y_new
        zy = mask_scale(z, y, hp.magnifying_factor)
        d = fx - (zy - torch.cat([alpha, beta], dim=2))
        x = topq(x - om * corr2d_pos(d, F), hp.q)
        fx = tconv_pos(x, F)
        alpha = alpha + fx[:, :, :M] - z
        beta = beta + fx[:, :, M:] - y
        if keep is not None:
            keep.append({"z": z, "y": y, "x": x, "fx": fx, "alpha": alpha, "beta": beta})
    return z, y, x, S


def forward_pos(codes, flat, hp: Hyperparam, dtype=torch.float32):
    """forward_pass_return_loss in position space.  codes: (B, Lb) uint8/int.  Returns (loss, intermediates)."""
    codes_t = torch.as_tensor(np.asarray(codes).astype(np.int64))
    Lb = codes_t.shape[1]
    P0 = unpack(flat, hp)
    P = {k: (v ** 2) for k, v in P0.items() if k not in ("D", "F")}
    D = prep_D(P0["D"], hp)
    F = prep_F(P0["F"])
    z, y, x, S = admm_xyz_pos(codes_t, D, F, P, hp, Lb)
    zy = mask_scale(z, y, hp.magnifying_factor)
    theta = torch.zeros_like(zy)
    Dc, Fc = D, F
    for n in range(hp.num_pass_df):
        R = recon_pos(z, y, Dc, Lb) + S                                                       # '+S' (model.jl:282-285)
        G = dgrad_pos(z, y, R, hp.f_len)
        Dn = (Dc * torch.exp(-P["mu"][n] * G)).reshape(hp.filter_len, 4, hp.M)
        Dc = (Dn / Dn.sum(dim=1, keepdim=True)).reshape(hp.f_len, hp.M)
        e = tconv_pos(x, Fc) - (zy + theta)
        Fu = torch.relu(Fc - P["kappa_stepsize"][n] * fgrad_pos(e, x) - P["kappa_stepsize"][n] * P["kappa_sparsity"][n])
        Fc = Fu / torch.sqrt((Fu ** 2).sum(dim=(0, 1), keepdim=True))
        theta = theta + tconv_pos(x, Fc) - zy
    nf = 1.0 / hp.batch_size
    rec = nf * ((recon_pos(z, y, Dc, Lb) - S) ** 2).sum()
    syn = nf * ((tconv_pos(x, Fc) - zy) ** 2).sum()
    return rec + syn, {"z": z, "y": y, "x": x, "D": Dc, "F": Fc, "zy": zy, "rec": rec, "syn": syn}


def loss_and_grad(codes, flat_np, hp, form="pos", dtype=torch.float32):
    """gradient(ps) do forward_pass_return_loss(...) end (train.jl:42-44) -> (loss, grads over the n_params trainable
    entries; the three warm-up scalars are not parameters: model.jl:68,72-73,137)."""
    flat = torch.tensor(np.asarray(flat_np), dtype=dtype, requires_grad=True)
    if form == "pos":
        loss, aux = forward_pos(codes, flat, hp, dtype)
    else:
        Lb = np.asarray(codes).shape[1]
        loss, aux = Literal(hp, Lb, dtype).forward(onehot_from_codes(np.asarray(codes), dtype), flat)
    loss.backward()
    g = flat.grad.detach().clone()
    g[n_params(hp):] = 0
    return float(loss.detach()), g.numpy(), aux


def code_retrieval(codes_all, flat_np, hp, dtype=torch.float32):
    """code_retrieval (_1_code_retrieval.jl:33-56): forward-only ADMM_XYZ over consecutive groups of batch_size
    (partial=false drops the tail); non-zeros of X as (position, fil, seq, Float16 mag), 0-based, ordered like
    findall on (l,1,K,B): seq, then fil, then position."""
    flat = torch.tensor(np.asarray(flat_np), dtype=dtype)
    P0 = unpack(flat, hp)
    P = {k: (v ** 2) for k, v in P0.items() if k not in ("D", "F")}
    D, F = prep_D(P0["D"], hp), prep_F(P0["F"])
    N, Lb = np.asarray(codes_all).shape
    recs = []
    with torch.no_grad():
        for i in range(0, N - N % hp.batch_size, hp.batch_size):
            ct = torch.as_tensor(np.asarray(codes_all[i:i + hp.batch_size]).astype(np.int64))
            _, _, x, _ = admm_xyz_pos(ct, D, F, P, hp, Lb)
            b, pos, fil = torch.nonzero(x.permute(0, 2, 1) > 0, as_tuple=True)[0], None, None
            nz = torch.nonzero(x.permute(0, 2, 1) > 0)                       # rows (b, k, i) sorted: seq, fil, position
            for bb, kk, ii in nz.tolist():
                recs.append((ii, kk, i + bb, np.float16(float(x[bb, ii, kk]))))
    out = np.zeros(len(recs), np.dtype([("position", "<u2"), ("fil", "<u2"), ("seq", "<u4"), ("mag_f16", "<u2"), ("_pad", "<u2")]))
    for n, (ii, kk, ss, mg) in enumerate(recs):
        out[n] = (ii, kk, ss, np.float16(mg).view(np.uint16), 0)
    return out


# ==================================================================================================
# AdaBelief (Flux 0.14.6 Optimise.AdaBelief; source not in /root/reference — the published rule of that release,
# restated [inferred]: bias-corrected moments and epsilon^2):
#   mt = b1 mt + (1-b1) g ;  st = b2 st + (1-b2) (g - mt)^2 + eps^2
#   p -= eta * mt / (1 - b1^t) / (sqrt(st / (1 - b2^t)) + eps^2)
#   defaults eta=1e-3, beta=(0.9, 0.999), eps=1e-8.  Per-step loss/gradient parity does not depend on this rule;
#   trajectory parity does.
# ==================================================================================================
def adabelief_step(p, g, mt, st, t, eta=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    mt = b1 * mt + (1 - b1) * g
    st = b2 * st + (1 - b2) * (g - mt) ** 2 + eps ** 2
    p = p - eta * mt / (1 - b1 ** t) / (np.sqrt(st / (1 - b2 ** t)) + eps ** 2)
    return p, mt, st
