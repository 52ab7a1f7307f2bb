"""What would sorting the bounce rays buy?  (GPU box; not a pytest file.)

Builds the first-bounce rays of a C2 frame on the host (first hits from the pipeline's own k_primary, cosine-weighted
directions), and traces the SAME rays through k_trace_generic (the traversal loop of k_trace) in different queue orders:
the renderer's (tile-major), scanline, shuffled, and binned by direction octant (with and without an origin cell).
The ratio between the orders is an upper bound for what a binning k_shade could gain in k_trace.
"""
import sys
import time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dsrt_b200 as d

pose = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 12
W, H = 1920, 1080
poses = d.read_pose_file(ROOT / "tests" / "golden" / "rendezvous_1s_dt0_01s.txt")
tris, mats = d.synth_mesh(1_000_000)
V0 = np.stack([tris["v0"][c] for c in "xyz"], -1).astype(np.float32)
E1 = np.stack([tris["v1"][c] for c in "xyz"], -1).astype(np.float32) - V0
E2 = np.stack([tris["v2"][c] for c in "xyz"], -1).astype(np.float32) - V0
rng = np.random.default_rng(7)
with d.Context(0) as c:
    c.upload_scene(tris, mats)
    cam_pos, sun, sep = d.pose_to_model_frame(poses[pose])
    cam = d.camera_look_at_origin(cam_pos, 40.0, W, H, 64, 10)
    p = d.make_params(W, H, 64, 10)
    O, D, key_tile, key_scan = [], [], [], []
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    tile_key = ((ys // 4) * (W // 8) + xs // 8).ravel() * 32 + ((ys % 4) * 8 + xs % 8).ravel()
    for s in range(nsamp):
        tri, t, u, v, _, _ = c.primary_hits(cam, p, sample=s)
        m = tri >= 0
        ti = tri[m]
        P = V0[ti] + u[m, None] * E1[ti] + v[m, None] * E2[ti]
        N = np.cross(E1[ti], E2[ti])
        N /= np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-30)
        flip = ((np.asarray(cam_pos, np.float32)[None, :] - P) * N).sum(1) < 0
        N[flip] *= -1
        # cosine-weighted direction about N
        r1, r2 = rng.random(len(ti), dtype=np.float32), rng.random(len(ti), dtype=np.float32)
        phi = 2 * np.pi * r1
        loc = np.stack([np.cos(phi) * np.sqrt(r2), np.sin(phi) * np.sqrt(r2), np.sqrt(1 - r2)], -1)
        A = np.where(np.abs(N[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
        T = np.cross(N, A); T /= np.linalg.norm(T, axis=1, keepdims=True)
        B = np.cross(N, T)
        Dd = loc[:, :1] * T + loc[:, 1:2] * B + loc[:, 2:] * N
        O.append(P); D.append(Dd.astype(np.float32))
        key_tile.append(s * (1 << 32) + tile_key[m].astype(np.int64))
        key_scan.append(s * (1 << 32) + np.nonzero(m)[0].astype(np.int64))
    O, D = np.concatenate(O), np.concatenate(D)
    key_tile, key_scan = np.concatenate(key_tile), np.concatenate(key_scan)
    n = len(O)
    rays = np.empty((n, 8), np.float32)
    rays[:, 0:3] = O; rays[:, 3] = 1e-3; rays[:, 4:7] = D; rays[:, 7] = 1e30
    octant = (D[:, 0] < 0).astype(np.int64) | ((D[:, 1] < 0).astype(np.int64) << 1) | ((D[:, 2] < 0).astype(np.int64) << 2)
    lo, hi = O.min(0), O.max(0)
    cell = np.minimum(((O - lo) / np.maximum(hi - lo, 1e-9) * 16).astype(np.int64), 15)
    cell_key = (cell[:, 0] * 16 + cell[:, 1]) * 16 + cell[:, 2]
    base = np.argsort(key_tile, kind="stable")
    orders = {
        "tile-major (the renderer's)": base,
        "scanline": np.argsort(key_scan, kind="stable"),
        "shuffled": rng.permutation(n),
        "octant bins, tile-major inside": base[np.argsort(octant[base], kind="stable")],
        "octant x 16^3 origin cells": base[np.argsort((octant * 4096 + cell_key)[base], kind="stable")],
        "16^3 origin cells only": base[np.argsort(cell_key[base], kind="stable")],
    }
    dev = torch.device("cuda:0")
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    print(f"pose {pose}: {n} first-bounce rays from {nsamp} samples", flush=True)
    ref = None
    for name, order in orders.items():
        r = torch.from_numpy(rays[order]).to(dev)
        best = 1e9
        for _ in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c.trace_closest_device(r.data_ptr(), n, hits.data_ptr())
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        h = hits.cpu().numpy()
        back = np.empty_like(h); back[order] = h
        if ref is None:
            ref = back
        same = np.array_equal(ref.view(np.uint32), back.view(np.uint32))
        print(f"{name:34s} {best * 1e3:7.3f} ms  {n / best / 1e6:7.0f} Mrays/s  hits identical to the first order: {same}", flush=True)
