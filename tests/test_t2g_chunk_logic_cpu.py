"""Lane-by-lane NumPy emulation of k_t2g_chunk's aggregation logic (pylamp_b200/csrc/markers.cu):
4 consecutive markers per lane, first/last run selection (s, e), masked sums, warp-level segmented
reduction of the last runs (head / run_end / span), one-marker path for the markers in between and
for chunks with a marker outside, generic path for the tail.  Checked against the oracle's
trac2grid on clouds that exercise every branch: whatever the marker order, every marker must be
added exactly once.  (The kernel itself is tested on the GPU in test_markers_gpu.py; this pins the
algorithm on CPU.)"""
import numpy as np
import pytest

from oracle import pylamp_oracle as O
from pylamp_b200 import setups


def emulate_chunk_kernel(x, f, axz, axx, merge_first=False):
    """Raw sums (wsum, fsum) over the extended axes axz/axx, computed the way the kernel does."""
    nze, nxe = len(axz), len(axx)
    z0, x0 = axz[0], axx[0]
    sz, sx = (nze - 1) / (axz[-1] - axz[0]), (nxe - 1) / (axx[-1] - axx[0])
    wsum, fsum = np.zeros(nze * nxe), np.zeros(nze * nxe)
    stats = {"single": 0, "first": 0, "merged_lanes": 0}

    def cell_weights(p):
        ie, je = int(np.floor((p[0] - z0) * sz)), int(np.floor((p[1] - x0) * sx))
        if ie == nze - 1 and p[0] <= axz[-1]:
            ie = nze - 2
        if je == nxe - 1 and p[1] <= axx[-1]:
            je = nxe - 2
        if not (0 <= ie <= nze - 2 and 0 <= je <= nxe - 2):
            return None, None
        az = (p[0] - axz[ie]) * (1.0 / (axz[ie + 1] - axz[ie]))
        ax = (p[1] - axx[je]) * (1.0 / (axx[je + 1] - axx[je]))
        bz, bx = 1 - az, 1 - ax
        return ie * nxe + je, np.array([(1 - ax) * (1 - az), (1 - ax) * (1 - bz), (1 - bx) * (1 - az), (1 - bx) * (1 - bz)])

    off = np.array([0, nxe, 1, nxe + 1])

    def single(m):
        c, w = cell_weights(x[m])
        if c is None:
            return
        stats["single"] += 1
        wsum[c + off] += w
        fsum[c + off] += f[m] * w

    M = x.shape[0]
    nchunk = M // 4
    for base in range(0, nchunk, 32):                 # one warp pass
        lanes = []
        for lane in range(32):
            ch = base + lane
            live = ch < nchunk
            cell, wu, ok = [0] * 4, np.zeros((4, 4)), live
            if live:
                for u in range(4):
                    c, w = cell_weights(x[4 * ch + u])
                    if c is None:
                        ok = False
                    else:
                        cell[u], wu[u] = c, w
            s, e = 4, (0 if live else 4)
            if ok:
                s = 3
                if cell[2] == cell[3]:
                    s = 2
                if s == 2 and cell[1] == cell[2]:
                    s = 1
                if s == 1 and cell[0] == cell[1]:
                    s = 0
                if s > 0:
                    e = 1
                    if s > 1 and cell[1] == cell[0]:
                        e = 2
                    if e == 2 and s > 2 and cell[2] == cell[0]:
                        e = 3
            ef = e if ok else 0
            lanes.append(dict(ch=ch, live=live, ok=ok, cell=cell, wu=wu, s=s, e=e, ef=ef,
                              cell_last=cell[3] if ok else -1 - lane, cell_first=cell[0] if ef > 0 else -1 - lane))

        def runs(cells):
            head = [lane == 0 or cells[lane - 1] != cells[lane] for lane in range(32)]
            run_end = []
            for lane in range(32):
                nxt = [j for j in range(lane + 1, 32) if head[j]]
                run_end.append(nxt[0] - 1 if nxt else 31)
            return head, run_end, max(run_end[lane] - lane for lane in range(32))

        def seg_reduce(S, run_end, span):
            o = 1
            while o <= span:                       # __shfl_down_sync reads the pre-step values of all lanes
                t = np.vstack([S[o:], np.zeros((o, 4))])
                for lane in range(32):
                    if lane + o <= run_end[lane]:
                        S[lane] += t[lane]
                o <<= 1

        head, run_end, span = runs([d["cell_last"] for d in lanes])
        head_f, run_end_f, span_f = runs([d["cell_first"] for d in lanes])
        for plane, vals in ((wsum, None), (fsum, f)):
            L, F = np.zeros((32, 4)), np.zeros((32, 4))
            for lane, d in enumerate(lanes):
                v = np.ones(4) if vals is None else (vals[4 * d["ch"]:4 * d["ch"] + 4] if d["ok"] else np.zeros(4))
                for u in range(4):
                    if u >= d["s"]:
                        L[lane] += v[u] * d["wu"][u]
                for u in range(3):
                    if u < d["ef"]:
                        F[lane] += v[u] * d["wu"][u]
            seg_reduce(L, run_end, span)
            for lane, d in enumerate(lanes):
                if head[lane] and d["ok"]:
                    plane[d["cell_last"] + off] += L[lane]
                    stats["merged_lanes"] += (plane is wsum) * (run_end[lane] - lane)
            if merge_first:
                seg_reduce(F, run_end_f, span_f)
            for lane, d in enumerate(lanes):
                if d["ef"] > 0 and (head_f[lane] or not merge_first):
                    plane[d["cell_first"] + off] += F[lane]
                    stats["first"] += plane is wsum
        for d in lanes:
            for u in range(4):
                if d["live"] and d["e"] <= u < d["s"]:
                    single(4 * d["ch"] + u)
    for m in range(4 * nchunk, M):                    # tail: generic kernel
        single(m)
    return wsum.reshape(nze, nxe), fsum.reshape(nze, nxe), stats


@pytest.mark.parametrize("merge_first", [False, True])
@pytest.mark.parametrize("cloud", ["random", "sorted", "drifted", "outside", "ragged"])
def test_chunk_aggregation_adds_every_marker_once(cloud, merge_first):
    rng = np.random.default_rng(4)
    ncz, ncx, L = 10, 9, [1.0, 0.75]
    nx = [ncz + 1, ncx + 1]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    x = setups.lattice_markers(ncz, ncx, L, 4, seed=3)[0]
    if cloud == "random":
        x = rng.random((1500, 2)) * L
    elif cloud == "drifted":
        x = np.clip(x + np.array([0.37 * L[0] / ncz, 0.61 * L[1] / ncx]), 1e-9, np.array(L) - 1e-9)
    elif cloud == "ragged":
        x = x[:-3]
    f = rng.uniform(1, 2, x.shape[0])
    for target in ([grid[0], grid[1]], [gridmp[0], grid[1]]):
        axz, axx = np.array(target[0], dtype=float), np.array(target[1], dtype=float)
        if cloud == "outside":           # ghost extension on the low z side and the high x side, a few markers beyond it
            axz = np.concatenate([[axz[0] - (axz[1] - axz[0])], axz])
            axx = np.concatenate([axx, [axx[-1] + (axx[-1] - axx[-2])]])
        xs = x.copy()
        if cloud == "outside":
            xs = xs + np.array([-0.4 * L[0] / ncz, 0.3 * L[1] / ncx])
            xs[7] = [-5.0, 0.1]          # outside even the extended axes: skipped, its chunk takes the one-marker path
        wsum, fsum, stats = emulate_chunk_kernel(xs, f, axz, axx, merge_first)
        # brute force: every marker added once
        nze, nxe = len(axz), len(axx)
        wref, fref = np.zeros((nze, nxe)), np.zeros((nze, nxe))
        for m in range(xs.shape[0]):
            ie = np.searchsorted(axz, xs[m, 0], side="right") - 1
            je = np.searchsorted(axx, xs[m, 1], side="right") - 1
            if not (0 <= ie <= nze - 2 and 0 <= je <= nxe - 2):
                continue
            az = (xs[m, 0] - axz[ie]) / (axz[ie + 1] - axz[ie])
            ax = (xs[m, 1] - axx[je]) / (axx[je + 1] - axx[je])
            for (di, dj, w) in ((0, 0, (1 - ax) * (1 - az)), (1, 0, (1 - ax) * az), (0, 1, ax * (1 - az)), (1, 1, ax * az)):
                wref[ie + di, je + dj] += w
                fref[ie + di, je + dj] += w * f[m]
        assert np.allclose(wsum, wref, rtol=1e-12, atol=1e-13)
        assert np.allclose(fsum, fref, rtol=1e-12, atol=1e-13)
        if cloud == "sorted":
            assert stats["single"] == 0 and stats["first"] == 0 and stats["merged_lanes"] > 0
        if cloud == "drifted":
            assert stats["first"] > 0
        if cloud in ("random", "outside"):
            assert stats["single"] > 0
