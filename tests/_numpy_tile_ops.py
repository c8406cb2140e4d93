"""NumPy provider of the tile operations of lsqfitgp_b200._dist.DistChol -- TEST INFRASTRUCTURE ONLY.

It lets the CPU test tier run the real host orchestration (ownership maps, panel/trailing ordering, collectives)
over gloo without a GPU.  The product provider is lsqfitgp_b200._dist.CudaTileOps (C ABI); nothing under
lsqfitgp_b200/ imports this file.  Each method restates what the corresponding lgp_dist_*/lgp_tile_* entry point of
include/lgp_b200.h computes.
"""
import contextlib

import numpy as np
import scipy.linalg
import torch


class _Ev:
    def record(self):
        pass

    def wait(self):
        pass


class NumpyTileOps:
    def __init__(self, K):
        self.K = np.asarray(K, dtype=float)

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    def empty(self, *shape, dtype=torch.float64):
        return torch.full(shape if not (len(shape) == 1 and isinstance(shape[0], tuple)) else shape[0], float('nan'),
                          dtype=dtype) if dtype.is_floating_point else torch.zeros(*shape, dtype=dtype)

    def main_stream(self):
        return contextlib.nullcontext()

    def panel_stream(self):
        return contextlib.nullcontext()

    def event(self):
        return _Ev()

    def synchronize(self):
        pass

    def gram_local(self, descs, x, rows, cols, lay, out=None):
        n = lay.n
        r = np.minimum(rows, n - 1)
        c = np.minimum(cols, n - 1)
        blk = torch.from_numpy(np.ascontiguousarray(self.K[np.ix_(r, c)]))
        if out is not None:
            out.copy_(blk)
            return out
        return blk

    # ---- lower-packed storage (lgp_dist_panel_* of include/lgp_b200.h)
    @staticmethod
    def _panel_geom(lay, lj):
        from lsqfitgp_b200._dist import tiles_before
        J = lay.pc + lay.Pc * lj
        first = tiles_before(J, lay.pr, lay.Pr)
        gi = lay.global_rows()[first * lay.T:]
        gj = np.arange(J * lay.T, (J + 1) * lay.T)
        return J, first, gi, gj

    def panel_diag(self, lay, lj, panel, d):
        J, first, gi, gj = self._panel_geom(lay, lj)
        if J % lay.Pr != lay.pr or panel.numel() == 0:
            return
        t = np.arange(lay.T)
        ok = gj < lay.n
        d.numpy()[gj[ok]] = panel.numpy()[t[ok], t[ok]]

    def panel_prepare(self, lay, lj, panel, sinv, rowsum):
        J, first, gi, gj = self._panel_geom(lay, lj)
        a = panel.numpy()
        si = sinv.numpy()
        a *= si[gi][:, None]
        a *= si[gj][None, :]
        pad = (gi[:, None] >= lay.n) | (gj[None, :] >= lay.n)
        a[pad] = 0
        a[(gi[:, None] == gj[None, :]) & pad] = 1
        absa = np.abs(np.where(pad, 0, a))
        rs = rowsum.numpy()
        ok = gi < lay.n
        np.add.at(rs, gi[ok], absa.sum(axis=1)[ok])
        # tiles strictly below the diagonal also stand for their mirror images
        offdiag = (gi // lay.T) != J
        okc = gj < lay.n
        np.add.at(rs, gj[okc], absa[offdiag].sum(axis=0)[okc])

    def panel_add_diag(self, lay, lj, panel, eps):
        J, first, gi, gj = self._panel_geom(lay, lj)
        if J % lay.Pr != lay.pr or panel.numel() == 0:
            return
        t = np.arange(lay.T)[gj < lay.n]
        panel.numpy()[t, t] += float(eps[0])

    def trailing_update_packed(self, lay, colpanels, k, panel, lj_begin, lj_end):
        from lsqfitgp_b200._dist import tiles_before
        T = lay.T
        li0 = tiles_before(k + 1, lay.pr, lay.Pr)
        for lj in range(max(lj_begin, 0), min(lj_end, lay.LC)):
            J = lay.pc + lay.Pc * lj
            if J <= k:
                continue
            li_s = tiles_before(J, lay.pr, lay.Pr)
            if li_s >= lay.LR:
                continue
            rJ = J % lay.Pr
            Aop = panel[lay.pr].numpy().reshape(-1, T)[(li_s - li0) * T:(lay.LR - li0) * T]
            off = J // lay.Pr - tiles_before(k + 1, rJ, lay.Pr)
            Bop = panel[rJ].numpy().reshape(-1, T)[off * T:(off + 1) * T]
            colpanels[lj].numpy()[...] -= Aop @ Bop.T

    @staticmethod
    def _owned_diag(lay):
        i = np.arange(lay.n)
        I = i // lay.T
        m = (I % lay.Pr == lay.pr) & (I % lay.Pc == lay.pc)
        i = i[m]
        I = I[m]
        return i, (I // lay.Pr) * lay.T + i % lay.T, (I // lay.Pc) * lay.T + i % lay.T

    def diag(self, lay, A, d):
        i, lr, lc = self._owned_diag(lay)
        d.numpy()[i] = A.numpy()[lr, lc]

    def scale_from_diag(self, lay, d):
        dn = d.numpy()[:lay.n]
        s = np.ones(lay.npad)
        nz = dn != 0
        s[:lay.n][nz] = np.exp2(np.rint(0.5 * np.log2(dn[nz])))
        return torch.from_numpy(s), torch.from_numpy(1 / s)

    def prepare(self, lay, A, sinv, rowsum):
        a = A.numpy()
        gi, gj = lay.global_rows(), lay.global_cols()
        si = sinv.numpy()
        a *= si[gi][:, None]
        a *= si[gj][None, :]
        pad = (gi[:, None] >= lay.n) | (gj[None, :] >= lay.n)
        a[pad] = 0
        a[(gi[:, None] == gj[None, :]) & pad] = 1
        rs = np.abs(np.where(pad, 0, a)).sum(axis=1)
        ok = gi < lay.n
        rowsum.numpy()[gi[ok]] = rs[ok]

    def eps(self, lay, rowsum, epsrel, epsabs):
        if epsrel < 0:
            epsrel = lay.n * np.finfo(float).eps
        m = rowsum.numpy().max()
        return torch.tensor([m, epsrel * m + epsabs])

    def add_diag(self, lay, A, eps):
        i, lr, lc = self._owned_diag(lay)
        A.numpy()[lr, lc] += float(eps[0])

    def potrf_tile(self, tile, invd, dvec, info, j0):
        t = tile.numpy()
        T = t.shape[0]
        try:
            L = np.linalg.cholesky(np.tril(t) + np.tril(t, -1).T)
        except np.linalg.LinAlgError:
            info[0] = min(int(info[0]), j0 + 1)
            L = np.full_like(t, np.nan)
        t[np.tril_indices(T)] = L[np.tril_indices(T)]
        dvec.numpy()[j0:j0 + T] = np.diag(L)
        iv = invd.numpy().reshape(T // 128, 128, 128)
        for b in range(T // 128):
            blk = L[128 * b:128 * (b + 1), 128 * b:128 * (b + 1)]
            iv[b] = scipy.linalg.solve_triangular(blk, np.eye(128), lower=True) if np.isfinite(blk).all() else np.nan

    def trsm_right(self, L, invd, B):
        if B.shape[0] == 0:
            return
        b = B.numpy()
        b[...] = scipy.linalg.solve_triangular(np.tril(L.numpy()), b.T, lower=True, check_finite=False).T

    def copy2d(self, src, dst):
        dst.copy_(src)

    def trailing_update(self, lay, A, k, panel, lj_begin, lj_end):
        from lsqfitgp_b200._dist import tiles_before
        a = A.numpy()
        T = lay.T
        li0 = tiles_before(k + 1, lay.pr, lay.Pr)
        for lj in range(max(lj_begin, 0), min(lj_end, lay.LC)):
            J = lay.pc + lay.Pc * lj
            if J <= k:
                continue
            li_s = tiles_before(J, lay.pr, lay.Pr)
            if li_s >= lay.LR:
                continue
            rJ = J % lay.Pr
            Aop = panel[lay.pr].numpy().reshape(-1, T)[(li_s - li0) * T:(lay.LR - li0) * T]
            off = J // lay.Pr - tiles_before(k + 1, rJ, lay.Pr)
            Bop = panel[rJ].numpy().reshape(-1, T)[off * T:(off + 1) * T]
            a[li_s * T:, lj * T:(lj + 1) * T] -= Aop @ Bop.T

    def trsv_tile(self, L, invd, b, trans):
        bn = b.numpy()
        bn[...] = scipy.linalg.solve_triangular(np.tril(L.numpy()), bn, lower=True, trans=1 if trans else 0, check_finite=False)

    def trmv_tile(self, L, x, y, trans):
        l = np.tril(L.numpy())
        y.numpy()[...] += (l.T if trans else l) @ x.numpy()

    def gemv(self, P, x, y, alpha, trans):
        if P.shape[0] == 0 or P.shape[1] == 0:
            return
        p = P.numpy()
        y.numpy()[...] += alpha * ((p.T @ x.numpy()) if trans else (p @ x.numpy()))


class NumpyPeerTileOps(NumpyTileOps):
    """ Adds a simulation of peer-mapped memory to the NumPy provider, so that the fused "panel solve -> broadcast" branch of
    DistChol._factor (slab / diagonal buffers living in a symmetric allocation, release/acquire counters) runs on CPU:
    the symmetric allocation of rank q is a tensor in POSIX shared memory that every process of the test maps
    (`bufs[q]`), "addresses" are (rank + 1) << 44 | byte offset (rank index 15 = the multicast mapping: a store goes to every
    rank), stores are NumPy copies, counters are int64 words of the same buffers polled with a sleep.  The processes run
    concurrently, so the READY / DONE protocol is exercised for real. """

    NFLAGS = 64
    SHIFT = 44
    MC = 15

    def __init__(self, K, bufs, rank, multicast):
        super().__init__(K)
        self.bufs, self.rank, self.multicast = bufs, rank, multicast
        self.device = torch.device('cpu')

    def peer_setup(self, count, group):
        import types
        from lsqfitgp_b200 import _dist
        t = self.bufs[self.rank]
        assert t.numel() >= self.NFLAGS + count
        t[:self.NFLAGS].zero_()
        h = types.SimpleNamespace(buffer_ptrs=[(q + 1) << self.SHIFT for q in range(len(self.bufs))],
                                  multicast_ptr=(self.MC + 1) << self.SHIFT if self.multicast else 0)
        return _dist._PeerBuffer(t, h, self.NFLAGS)

    def _targets(self, addr):
        q = (addr >> self.SHIFT) - 1
        off = (addr & ((1 << self.SHIFT) - 1)) // 8
        return (list(range(len(self.bufs))) if q == self.MC else [q]), off

    def _store(self, src, dst_ptrs, ld_dst):
        rows, cols = src.shape
        for addr in dst_ptrs:
            ranks, off = self._targets(addr)
            for q in ranks:
                self.bufs[q][off:off + rows * ld_dst].view(rows, ld_dst)[:, :cols] = src

    def trsm_right_bcast(self, L, invd, B, dst_ptrs, ld_dst, multimem):
        if B.shape[0] == 0:
            return
        assert bool(multimem) == self.multicast and (len(dst_ptrs) == 1 or not multimem)
        self.trsm_right(L, invd, B)
        self._store(B, dst_ptrs, ld_dst)

    def copy2d_bcast(self, src, dst_ptrs, ld_dst, multimem):
        self._store(src, dst_ptrs, ld_dst)

    def flag_signal(self, ptrs, value):
        for addr in ptrs:
            ranks, off = self._targets(addr)
            assert off < self.NFLAGS and len(ranks) == 1
            self.bufs[ranks[0]][:self.NFLAGS].view(torch.int64)[off] = int(value)

    def flag_wait(self, addr, n, value, timeout_ms, err):
        import time
        ranks, off = self._targets(addr)
        assert ranks == [self.rank] and off + n <= self.NFLAGS     # waits are on local memory only
        flags = self.bufs[self.rank][:self.NFLAGS].view(torch.int64)
        t0 = time.monotonic()
        while not bool((flags[off:off + n] >= int(value)).all()):
            if (time.monotonic() - t0) * 1e3 > timeout_ms:
                err[0] = 1
                return
            time.sleep(0.0005)
