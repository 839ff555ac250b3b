"""numpy stand-in for the fw_shard_* kernels (TEST INFRASTRUCTURE): same data flow --
local row shard, local column snapshots Cp/NCp, broadcast row-snapshot panel Rw -- so the
multi-rank schedule in floydwarshall_b200.sharded can be exercised on CPU under gloo."""
import numpy as np

B = 128


class NumpyShardBackend:
    def __init__(self, n, row0, rate, nxt, block=B):
        self.n, self.row0, self.rows, self.B = n, row0, rate.shape[0], block
        self.rate, self.next = rate, nxt
        self.Rw2 = [np.zeros((block, n)) for _ in range(16)]
        self.Rw = self.Rw2[0]
        self.Cp = np.zeros((self.rows, block))
        self.NCp = np.zeros((self.rows, block), dtype=np.int32)
        # the diagonal is held as NaN while solving (see csrc/fw_common.cuh) and restored by finish()
        li = np.arange(self.rows)
        self.diag = rate[li, li + row0].copy()
        rate[li, li + row0] = np.nan

    def finish(self):
        li = np.arange(self.rows)
        self.rate[li, li + self.row0] = self.diag

    def _relax(self, R, X, a, an, b):
        with np.errstate(invalid="ignore", over="ignore"):
            nv = np.outer(a, b)
            upd = R < nv
        R[upd] = nv[upd]
        X[upd] = np.broadcast_to(an[:, None], R.shape)[upd]

    def pivot(self, b0, buf=0):
        B_ = self.B
        Rw = self.Rw2[buf]
        lr = b0 - self.row0
        ks = slice(b0, b0 + B_)
        D = self.rate[lr:lr + B_, ks]
        DX = self.next[lr:lr + B_, ks]
        for kk in range(B_):
            self.Cp[lr:lr + B_, kk] = D[:, kk]
            self.NCp[lr:lr + B_, kk] = DX[:, kk]
            Rw[kk, ks] = D[kk, :]
            self._relax(D, DX, D[:, kk].copy(), DX[:, kk].copy(), D[kk, :].copy())
        out = np.ones(self.n, bool)
        out[ks] = False
        Xr = self.rate[lr:lr + B_][:, out]
        XX = self.next[lr:lr + B_][:, out]
        Cd = self.Cp[lr:lr + B_]
        NCd = self.NCp[lr:lr + B_]
        Rwo = np.empty((B_, Xr.shape[1]))
        for kk in range(B_):
            Rwo[kk] = Xr[kk]
            self._relax(Xr, XX, Cd[:, kk], NCd[:, kk], Xr[kk].copy())
        self.rate[lr:lr + B_, out] = Xr
        self.next[lr:lr + B_, out] = XX
        Rw[:, out] = Rwo

    def update(self, b0, buf=0, mode=0, lr0=0, lrn=None, extra_skip=None):
        """mode 0: all rows outside the k-block; 1: only rows [lr0, lr0+lrn); 2: mode 0 minus those rows.
        extra_skip: (first local row, count) additionally left out (the other block of a pair)."""
        B_ = self.B
        lrn = B_ if lrn is None else lrn
        Rw = self.Rw2[buf]
        ks = slice(b0, b0 + B_)
        rout = np.ones(self.rows, bool)
        if self.row0 <= b0 < self.row0 + self.rows:
            rout[b0 - self.row0:b0 - self.row0 + B_] = False
        if extra_skip is not None and 0 <= extra_skip[0] < self.rows:
            rout[extra_skip[0]:extra_skip[0] + extra_skip[1]] = False
        if mode == 1:
            only = np.zeros(self.rows, bool)
            only[lr0:lr0 + lrn] = True
            rout &= only
        elif mode == 2:
            rout[lr0:lr0 + lrn] = False
        if not rout.any():
            return
        Y = self.rate[rout][:, ks]
        YX = self.next[rout][:, ks]
        Rd = Rw[:, ks]
        Cc = np.empty((Y.shape[0], B_))
        NCc = np.empty((Y.shape[0], B_), dtype=np.int32)
        for kk in range(B_):
            Cc[:, kk] = Y[:, kk]
            NCc[:, kk] = YX[:, kk]
            self._relax(Y, YX, Y[:, kk].copy(), YX[:, kk].copy(), Rd[kk])
        ridx = np.where(rout)[0]
        self.rate[np.ix_(ridx, np.arange(b0, b0 + B_))] = Y
        self.next[np.ix_(ridx, np.arange(b0, b0 + B_))] = YX
        self.Cp[ridx] = Cc
        self.NCp[ridx] = NCc
        cout = np.ones(self.n, bool)
        cout[ks] = False
        cidx = np.where(cout)[0]
        Rb = self.rate[np.ix_(ridx, cidx)]
        Xb = self.next[np.ix_(ridx, cidx)]
        for kk in range(B_):
            self._relax(Rb, Xb, self.Cp[ridx, kk], self.NCp[ridx, kk], Rw[kk, cidx])
        self.rate[np.ix_(ridx, cidx)] = Rb
        self.next[np.ix_(ridx, cidx)] = Xb

    def update_group(self, b0, nb, buf=0, mode=0, lr0=0, lrn=0):
        """k-blocks b0, b0+B, ... (panels Rw2[buf..buf+nb-1]) on every local row outside the blocks' own rows
        (modes as in update): by definition the single-block updates one after the other."""
        B_ = self.B
        for i in range(nb):
            self.update(b0 + i * B_, buf + i, mode, lr0, lrn if lrn else None, extra_skip=(b0 - self.row0, nb * B_))

    def update_pair(self, b0, buf=0, mode=0, lr0=0, lrn=0):
        self.update_group(b0, 2, buf, mode, lr0, lrn)
