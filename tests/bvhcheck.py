"""Host-side invariants of the GPU-built wide BVH (layout in csrc/bvh8.cuh).

Decodes the downloaded 96-byte nodes and 48-byte triangle records and checks:
  * every uploaded triangle index appears in exactly one leaf record,
  * v0 / e1 / e2 of each record equal the uploaded vertices (e = one fp32 subtraction),
  * every leaf child's decoded (de-quantised) box contains its triangles,
  * every inner child's decoded box contains the whole subtree below it,
  * inner children are contiguous and ordered by slot; meta bytes are well formed,
  * the bfloat16 triangle weights the traversal sums (word 5) equal unary(count) << offset.
"""
from __future__ import annotations

import numpy as np


def _bytes(word):
    return [(int(word) >> (8 * k)) & 0xFF for k in range(4)]


def decode_node(n):
    """n: 24 uint32 words -> dict."""
    p = n[0:3].view(np.float32).astype(np.float64)
    ew = int(n[3])
    e = [(ew >> (8 * k)) & 0xFF for k in range(3)]
    imask = ew >> 24
    scale = np.array([2.0 ** (x - 127) for x in e])
    meta = _bytes(n[6]) + _bytes(n[7])
    q = {}
    q["lox"] = _bytes(n[8]) + _bytes(n[9])
    q["loy"] = _bytes(n[10]) + _bytes(n[11])
    q["loz"] = _bytes(n[12]) + _bytes(n[13])
    q["hix"] = _bytes(n[14]) + _bytes(n[15])
    q["hiy"] = _bytes(n[16]) + _bytes(n[17])
    q["hiz"] = _bytes(n[18]) + _bytes(n[19])
    lo = np.stack([np.array(q["lox"]), np.array(q["loy"]), np.array(q["loz"])], 1) * scale + p
    hi = np.stack([np.array(q["hix"]), np.array(q["hiy"]), np.array(q["hiz"])], 1) * scale + p
    # word 5: eight bfloat16 numbers = the top halves of float32 bit patterns
    wbits = np.array([(int(n[20 + s // 2]) >> (16 * (s % 2))) & 0xFFFF for s in range(8)], np.uint32) << 16
    weight = wbits.view(np.float32).astype(np.float64)
    return dict(p=p, imask=imask, child_base=int(n[4]), tri_base=int(n[5]), meta=meta, lo=lo, hi=hi, weight=weight)


def check_bvh(nodes: np.ndarray, recs: np.ndarray, tris: np.ndarray) -> dict:
    """Returns statistics; raises AssertionError on a violated invariant."""
    ntris = len(tris)
    nn = len(nodes)
    v = np.stack([np.stack([tris[name][k] for k in "xyz"], 1) for name in ("v0", "v1", "v2")], 1)  # (n,3,3) f32
    rec_idx = recs[:, 3].view(np.uint32).astype(np.int64)
    assert len(recs) == ntris
    assert np.array_equal(np.sort(rec_idx), np.arange(ntris)), "every triangle must appear exactly once"
    # record contents
    v0 = v[rec_idx, 0]
    assert np.array_equal(recs[:, 0:3], v0), "record v0"
    assert np.array_equal(recs[:, 4:7], (v[rec_idx, 1] - v0).astype(np.float32)), "record e1 = v1 - v0 in fp32"
    assert np.array_equal(recs[:, 8:11], (v[rec_idx, 2] - v0).astype(np.float32)), "record e2 = v2 - v0 in fp32"

    tri_lo = v.min(1).astype(np.float64)[rec_idx]
    tri_hi = v.max(1).astype(np.float64)[rec_idx]

    visited = np.zeros(nn, bool)
    tri_seen = np.zeros(ntris, np.int32)
    stats = dict(nodes=nn, leaves=0, max_depth=0, children_hist=[0] * 9, leaf_hist=[0] * 4)

    # iterative post-order: compute the true bounds of each subtree, compare with parent's boxes
    sub_lo = np.full((nn, 3), np.inf)
    sub_hi = np.full((nn, 3), -np.inf)
    order = []
    stack = [(0, 1)]
    dec = {}
    while stack:
        ni, depth = stack.pop()
        assert 0 <= ni < nn, "node index in range"
        assert not visited[ni], "node reachable twice"
        visited[ni] = True
        stats["max_depth"] = max(stats["max_depth"], depth)
        d = decode_node(nodes[ni])
        dec[ni] = d
        order.append(ni)
        rank = 0
        nchild = 0
        for s in range(8):
            m = d["meta"][s]
            inner = (d["imask"] >> s) & 1
            if m == 0:
                assert not inner
                assert d["weight"][s] == 0.0, "empty slot carries no triangle weight"
                continue
            nchild += 1
            if inner:
                assert m == (0x20 | (24 + s)), f"inner meta of slot {s}"
                assert d["weight"][s] == 0.0, "inner slot carries no triangle weight"
                stack.append((d["child_base"] + rank, depth + 1))
                rank += 1
            else:
                cnt_bits = m >> 5
                assert cnt_bits in (1, 3, 7), "unary triangle count"
                assert d["weight"][s] == float(cnt_bits << (m & 31)), f"triangle weight of slot {s}"
        stats["children_hist"][nchild] += 1
    assert visited.all(), "all nodes reachable"

    for ni in reversed(order):
        d = dec[ni]
        rank = 0
        for s in range(8):
            m = d["meta"][s]
            if m == 0:
                continue
            if (d["imask"] >> s) & 1:
                ci = d["child_base"] + rank
                rank += 1
                lo, hi = sub_lo[ci], sub_hi[ci]
            else:
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]
                off = m & 31
                a = d["tri_base"] + off
                assert off + cnt <= 24
                tri_seen[a:a + cnt] += 1
                lo, hi = tri_lo[a:a + cnt].min(0), tri_hi[a:a + cnt].max(0)
                stats["leaves"] += 1
                stats["leaf_hist"][cnt] += 1
            assert np.all(d["lo"][s] <= lo) and np.all(d["hi"][s] >= hi), \
                f"node {ni} slot {s}: decoded box {d['lo'][s]}..{d['hi'][s]} must contain {lo}..{hi}"
            sub_lo[ni] = np.minimum(sub_lo[ni], lo)
            sub_hi[ni] = np.maximum(sub_hi[ni], hi)
    assert np.all(tri_seen == 1), "every record referenced by exactly one leaf"
    return stats
