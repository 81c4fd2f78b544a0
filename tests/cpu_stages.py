"""TEST DOUBLE for ``abc_dls_b200.stages.CudaStages`` — numpy implementations of the stage methods so that
``engine.py``'s sequencing of stages and collectives (row sharding, histogram all-reduce per radix pass, the
cross-rank cap rule, normal-equation / residual / summary reductions) can run under ``gloo`` with world_size 2
on a machine without a GPU.  It is test infrastructure only: the product never imports it and has no CPU path.
"""
import numpy as np
import torch

BINS = 2048


def _key(v):
    b = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def _unkey(k):
    k = np.uint64(k)
    b = (k & np.uint64((1 << 63) - 1)) if (k >> np.uint64(63)) else ~k
    return np.array([b], dtype=np.uint64).view(np.float64)[0]


class _SelWs:
    pass


class CpuStages:
    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0
        self.timer = None

    def new(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype)

    # ---- radix select ------------------------------------------------------------------------------
    def select_workspace(self, njobs):
        w = _SelWs()
        w.hist = torch.zeros((njobs, 2, BINS), dtype=torch.int64)
        return w

    def select_begin(self, ws, njobs, base, n, ld, counts, center, k0, k1):
        ws.keys = []
        for j in range(njobs):
            x = base[j, :n].numpy()
            if center is not None:
                x = np.abs(x - center[j].item())
            ws.keys.append(_key(x))
        ws.k = np.array([[int(k0), int(k1)]] * njobs, dtype=np.int64)
        ws.krem = ws.k.copy()
        ws.prefix = np.zeros((njobs, 2), dtype=np.uint64)
        ws.hist.zero_()

    def select_hist_tensor(self, ws, njobs):
        return ws.hist

    def select_hist(self, ws, njobs, n_max, p):
        sd = 64 - 11 * (p + 1) if p < 5 else 0
        mask = 2047 if p < 5 else 511
        for j in range(njobs):
            for q in range(2):
                key = ws.keys[j]
                if p > 0:
                    key = key[(key >> np.uint64(64 - 11 * p)) == ws.prefix[j, q]]
                dig = ((key >> np.uint64(sd)) & np.uint64(mask)).astype(np.int64)
                ws.hist[j, q] += torch.from_numpy(np.bincount(dig, minlength=BINS)[:BINS])

    def select_pick(self, ws, njobs, p):
        bits = 11 if p < 5 else 9
        for j in range(njobs):
            for q in range(2):
                h = ws.hist[j, q].numpy()
                cum = np.cumsum(h)
                b = int(np.searchsorted(cum, ws.krem[j, q], side="right"))
                ws.krem[j, q] -= (cum[b - 1] if b > 0 else 0)
                ws.prefix[j, q] = (ws.prefix[j, q] << np.uint64(bits)) | np.uint64(b)
        ws.hist.zero_()

    def select_end(self, ws, njobs, out):
        for j in range(njobs):
            for q in range(2):
                out[j, q] = float(_unkey(ws.prefix[j, q]))

    def median_from_ranks(self, ranks2, ncols, scale, out):
        r = ranks2.numpy()
        out[:ncols] = torch.from_numpy(scale * (0.5 * r[:ncols, 0] + 0.5 * r[:ncols, 1]))

    # ---- distance / cut -------------------------------------------------------------------------------
    def scale_dist(self, ss, n, d, ld, mad, target, dist, status):
        s = np.zeros(n)
        for j in range(d):
            m = mad[j].item()
            col = ss[j, :n].numpy()
            t = target[j].item()
            a = (col - t) if m == 0 else (col / m - t / m)
            s = s + a * a
        dist[:n] = torch.from_numpy(np.sqrt(s))
        if not np.isfinite(s).all():
            status |= 1

    def accept_workspace(self, n):
        return _SelWs()

    def count_le(self, dist, n, n_dev, ds, le_count, ws):
        m = n if n_dev is None else min(n, int(n_dev))
        ws.flags = dist[:m].numpy() <= ds.item()
        le_count[0] = int(ws.flags.sum())

    def accept(self, dist, n, n_dev, src_idx, ds, quota, row_offset, cap, idx, dist_out, n_out, ws, status,
               slot_out=None):
        rows = np.flatnonzero(ws.flags)[: max(int(quota), 0)][:cap]
        if slot_out is not None:
            slot_out[: rows.size] = torch.from_numpy(rows)
        idx[: rows.size] = torch.from_numpy((src_idx.numpy()[rows] if src_idx is not None else rows + row_offset))
        if dist_out is not None:
            dist_out[: rows.size] = dist[torch.from_numpy(rows)]
        n_out[0] = rows.size

    # ---- loclinear -----------------------------------------------------------------------------------
    def gather(self, ss, ld_ss, param, ld_param, dist_acc, idx, n_acc, row_offset, cap, d, P, mad, target, ds, xs, y,
               ss_out, w):
        n = int(n_acc)
        loc = (idx[:n] - row_offset)
        for j in range(d):
            m = mad[j].item()
            col = ss[j][loc].numpy()
            ss_out[j, :n] = torch.from_numpy(col)
            xs[j, :n] = torch.from_numpy(col - target[j].item() if m == 0 else col / m - target[j].item() / m)
        if param is not None:
            for p in range(P):
                y[p, :n] = param[p][loc]
        q = dist_acc[:n].numpy() / ds.item()
        w[:n] = torch.from_numpy(1.0 - q * q)

    def gather_param_rows(self, par, row_pitch, P, idx, n_acc, row_offset, cap, y):
        n = int(n_acc)
        assert par.stride(0) == 1 and par.stride(1) == row_pitch
        y[:, :n] = par[:, idx[:n] - row_offset]

    def normal(self, xs, rhs, w, n_acc, cap, d, P, with_gram, partial):
        n = int(n_acc)
        X = np.column_stack([np.ones(n)] + [xs[j, :n].numpy() for j in range(d)]) if n else np.zeros((0, d + 1))
        W = w[:n].numpy()[:, None]
        Y = np.column_stack([rhs[p, :n].numpy() for p in range(P)]) if n else np.zeros((0, P))
        parts = ([((X * W).T @ X).ravel()] if with_gram else []) + [((X * W).T @ Y).ravel()]
        partial[:] = torch.from_numpy(np.concatenate(parts))

    def solve(self, gram, rhs, d, P, beta, status):
        M = d + 1
        beta[:] = torch.from_numpy(np.linalg.solve(gram.numpy().reshape(M, M), rhs.numpy().reshape(M, P)))

    def residual(self, xs, y, beta, n_acc, cap, d, P, res):
        n = int(n_acc)
        X = np.column_stack([np.ones(n)] + [xs[j, :n].numpy() for j in range(d)])
        res[:, :n] = torch.from_numpy((y[:, :n].numpy().T - X @ beta.numpy()).T.copy())

    def residual_stats(self, xs, y, beta, w, n_acc, cap, d, P, res, stat):
        self.residual(xs, y, beta, n_acc, cap, d, P, res)
        n = int(n_acc)
        r, ww = res[:, :n], w[:n]
        stat[0], stat[1], stat[2] = r.sum(dim=1), (r * ww).sum(dim=1), (r * r * ww).sum(dim=1)

    def normal_logres(self, xs, res, the_m, w, n_acc, cap, d, P, partial):
        n = int(n_acc)
        res[:, :n] -= the_m[:, None]
        self.normal(xs, torch.log(res ** 2), w, n_acc, cap, d, P, False, partial)

    def recentre_log(self, res, the_m, n_acc, cap, P, logres2):
        n = int(n_acc)
        res[:, :n] -= the_m[:, None]
        if logres2 is not None:
            logres2[:, :n] = torch.log(res[:, :n] ** 2)

    def adjust(self, xs, res, beta, the_m, gamma, n_acc, cap, d, P, hcorr, adj):
        n = int(n_acc)
        r = res[:, :n].numpy()
        if hcorr:
            X = np.column_stack([np.ones(n)] + [xs[j, :n].numpy() for j in range(d)])
            g = gamma.numpy()
            r = (np.sqrt(np.exp(g[0]))[:, None] * r) / np.sqrt(np.exp(X @ g)).T
            res[:, :n] = torch.from_numpy(r.copy())
        adj[:, :n] = torch.from_numpy(((beta[0] + the_m).numpy()[:, None] + r).copy())

    def col_stats(self, values, w, n_acc, cap, ncols, out):
        n = int(n_acc)
        ww = w[:n].numpy() if w is not None else np.ones(n)
        for c in range(ncols):
            x = values[c, :n].numpy()
            out[c] = torch.tensor([x.min() if n else np.inf, x.max() if n else -np.inf, x.sum(), (ww * x).sum(),
                                   (ww * x * x).sum(), ww.sum()])

    def resid_moments(self, rs, n_acc, P, sums):
        sums[:P], sums[P:2 * P], sums[2 * P:3 * P] = rs[:, 2], rs[:, 3], rs[:, 4]
        sums[3 * P] = float(n_acc)

    def resid_finish(self, sums, gram, P, the_m, sig, logsig):
        sw = gram[0]
        m = sums[:P] / sums[3 * P]
        the_m.copy_(m)
        sig.copy_((sums[2 * P:3 * P] - 2.0 * m * sums[P:2 * P] + m * m * sw) / sw)
        logsig[0] = torch.log(sig).sum()

    def final_small_len(self, d, P, loclinear) -> int:
        return 8 + d + 6 * P + (2 * P + 2 * (d + 1) * P if loclinear else 0)

    def final_pack(self, status, reg, stats, d, P, pack, sm):
        bits = ((int(status) >> np.arange(8)) & 1).astype(np.float64)
        pack.copy_(torch.cat([reg[:, 0], -reg[:, 1], stats[:, 0], -stats[:, 1], -torch.from_numpy(bits)]))
        sm.copy_(stats[:, 2:])

    def final_small(self, status, n_acc, ds, reg, stats, mad, target, logsig, beta, gamma, the_m, sig, pack, sm, d, P,
                    small):
        M = d + 1
        st = int(status)
        rmin, rmax, stt = reg[:, 0], reg[:, 1], stats.clone()
        if pack is not None:
            rmin, rmax = pack[:d], -pack[d:2 * d]
            stt = torch.cat([pack[2 * d:2 * d + P, None], -pack[2 * d + P:2 * d + 2 * P, None], sm], dim=1)
            st = int(sum(1 << b for b in range(8) if float(pack[2 * d + 2 * P + b]) != 0.0))
        small.zero_()
        small[:6] = torch.tensor([st, float(n_acc), float(ds), float((rmin == rmax).all()), float((mad == 0).all()),
                                  float(logsig) if logsig is not None else 0.0], dtype=torch.float64)
        o = 8
        small[o:o + d] = mad
        o += d
        small[o:o + 6 * P] = stt.reshape(-1)
        o += 6 * P
        if beta is not None:
            tsc = torch.where(mad == 0, target, target / torch.where(mad == 0, torch.ones_like(mad), mad))
            small[o:o + P] = sig
            small[o + P:o + 2 * P] = beta[0] + the_m
            o += 2 * P
            for b in (beta, gamma):
                if b is not None:
                    c = b.clone()
                    c[0] = b[0] - (tsc[:, None] * b[1:]).sum(dim=0)
                    small[o:o + M * P] = c.reshape(-1)
                o += M * P

    def postpr_counts(self, model_id, idx, n_acc, row_offset, cap, K, counts):
        n = int(n_acc)
        counts[:] = torch.bincount(model_id[idx[:n] - row_offset].long(), minlength=K)[:K]

    # ---- SMC --------------------------------------------------------------------------------------------
    def smc_oldrange(self, table, n, P, ld, minmax):
        minmax[:P] = table[:, :n].min(dim=1).values
        minmax[P:] = table[:, :n].max(dim=1).values

    def smc_box_filter(self, table, n, P, ld, lo, hi, row_offset, cap, idx, n_out, status):
        keep = ((table[:, :n] >= lo[:, None]) & (table[:, :n] <= hi[:, None])).all(dim=0)
        rows = torch.nonzero(keep).view(-1)
        idx[: rows.numel()] = rows + row_offset
        n_out[0] = rows.numel()
