"""Bounding-box statistics of the sampling windows per (8x8 query patch, head, level) for distribution E."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from richsem_b200 import synthetic as syn
from richsem_b200._capi import build_patch_order
PH, PW = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 8
shapes = syn.level_shapes(800, 1333)
i = syn.make_inputs("E", 1, shapes, "cpu", seed=1)
loc = i["loc"][0].numpy()  # (S, M, L, P, 2)
S = loc.shape[0]
starts = [0]
for h, w in shapes[:-1]: starts.append(starts[-1] + h * w)
# patch tiles
tiles = []
for (h, w), st in zip(shapes, starts):
    for ty in range(0, h, PH):
        for tx in range(0, w, PW):
            ys, xs = np.meshgrid(np.arange(ty, min(ty + PH, h)), np.arange(tx, min(tx + PW, w)), indexing="ij")
            tiles.append((len(tiles), st + (ys * w + xs).ravel()))
print("tiles", len(tiles))
ext = {l: [] for l in range(4)}
qlevel = []
for tid, toks in tiles:
    ql = sum(1 for s_ in starts if toks[0] >= s_) - 1
    for m in range(8):
        for l, (H, W) in enumerate(shapes):
            x = loc[toks, m, l, :, 0].astype(np.float32) * np.float32(W) - np.float32(0.5)
            y = loc[toks, m, l, :, 1].astype(np.float32) * np.float32(H) - np.float32(0.5)
            ok = (x > -1) & (y > -1) & (x < W) & (y < H)
            if not ok.any(): continue
            w0 = np.floor(x[ok]); h0 = np.floor(y[ok])
            ext[l].append((ql, h0.max() - h0.min() + 2, w0.max() - w0.min() + 2))
for l in range(4):
    a = np.array(ext[l])
    for ql in range(4):
        b = a[a[:, 0] == ql]
        if len(b) == 0: continue
        print(f"sample level {l} query level {ql}: n={len(b)} H ext mean {b[:,1].mean():.1f} p50 {np.percentile(b[:,1],50):.0f} p99 {np.percentile(b[:,1],99):.0f} max {b[:,1].max():.0f} | W ext mean {b[:,2].mean():.1f} p99 {np.percentile(b[:,2],99):.0f} max {b[:,2].max():.0f} | area mean {(b[:,1]*b[:,2]).mean():.0f}")

# greedy coarse-to-fine allocation: fraction of samples served from the window pool
import collections
per_tile = collections.defaultdict(dict)
k = {l: 0 for l in range(4)}
rows_all = []
for tid, toks in tiles:
    for m in range(8):
        rws = []
        for l, (H, W) in enumerate(shapes):
            x = loc[toks, m, l, :, 0].astype(np.float32) * np.float32(W) - np.float32(0.5)
            y = loc[toks, m, l, :, 1].astype(np.float32) * np.float32(H) - np.float32(0.5)
            ok = (x > -1) & (y > -1) & (x < W) & (y < H)
            if not ok.any(): rws.append((0, 0)); continue
            w0 = np.floor(x[ok]); h0 = np.floor(y[ok])
            rws.append((int((h0.max() - h0.min() + 2) * (w0.max() - w0.min() + 2)), int(ok.sum())))
        rows_all.append((len(toks), rws))
for P in (320, 384, 448, 512, 640, 768):
    tot = 0; win = 0; staged = 0
    for nq, rws in rows_all:
        used = 0
        for l in (3, 2, 1, 0):
            r, n = rws[l]
            tot += n
            if n and used + r <= P:
                used += r; win += n; staged += r
    print(f"pool {P} rows: windowed samples {100*win/tot:.1f}%  staged rows per windowed sample {staged/win:.3f}")
