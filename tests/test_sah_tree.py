"""Sweep-SAH tree (csrc/sah.cu, host code of the library) against the oracle's independent restatement
(oracle/oracle_geom.h SahBuilder) -- runs without a GPU: pt_sah_build_host is plain host code.

Bars: node records bit-identical (topology, preorder numbering, child boxes); every leaf referenced exactly once;
the tree answers closest-hit / any-hit queries exactly like brute force over all triangles (oracle side)."""
import numpy as np
import pytest

import oracle_api as O
from acgpathtracing_b200 import scenes as S


def _boxes(arr):
    v = arr.verts[arr.indices][:, :, :3]
    return v.min(1), v.max(1)


def _soup(rng, n):
    verts = rng.uniform(0, 500, (3 * n, 3)).astype(np.float32)
    verts[1::3] = verts[0::3] + rng.normal(0, 40, (n, 3)).astype(np.float32)
    verts[2::3] = verts[0::3] + rng.normal(0, 40, (n, 3)).astype(np.float32)
    if n >= 8:                       # duplicates and degenerate triangles: equal sort keys, zero-area boxes
        verts[3:6] = verts[0:3]
        verts[6:9] = verts[6]
    v4 = np.ones((3 * n, 4), np.float32)
    v4[:, :3] = verts
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return S.SceneArrays(verts=v4, indices=idx, mat_ids=np.zeros(n, np.uint32),
                         materials=[dict(name="m", Kd=(0.5, 0.5, 0.5), Ke=(0, 0, 0))], name="soup")


def _cases():
    rng = np.random.default_rng(5)
    out = [("cornell", S.cornell32()), ("material", S.material_scene(2)), ("tess40", S.tessellated_scene(40))]
    out += [(f"soup{n}", _soup(rng, n)) for n in (2, 3, 9, 64, 257, 1500)]
    return out


@pytest.mark.parametrize("name,arr", _cases(), ids=lambda x: x if isinstance(x, str) else "")
def test_host_builder_matches_oracle(port, name, arr):
    from acgpathtracing_b200 import api
    osc = O.OracleScene(port, arr)
    b, so = osc.bvh(), osc.bvh_sah()
    lo, hi = _boxes(arr)
    nodes, depth = api.sah_build_host(lo, hi, b["sorted_prim"])
    ref = so["nodes"].copy()
    got = nodes.copy()
    ref[:, [11, 15]] = 0             # padding words
    got[:, [11, 15]] = 0
    assert depth == so["depth"]
    assert np.array_equal(got, ref)
    # structure: n - 1 nodes in preorder, every leaf slot and every node but the root referenced exactly once
    ch = nodes[:, [3, 7]].view(np.int32)
    assert np.array_equal(np.sort(~ch[ch < 0]), np.arange(arr.n_tri))
    assert np.array_equal(np.sort(ch[ch >= 0]), np.arange(1, arr.n_tri - 1))
    assert np.all(ch[:, 0][ch[:, 0] >= 0] == np.nonzero(ch[:, 0] >= 0)[0] + 1)       # left child = next node (preorder)
    # child boxes are the exact unions of their leaves' boxes (checked on the root's two children)
    f = nodes.view(np.float32)
    slot_lo, slot_hi = lo[b["sorted_prim"]], hi[b["sorted_prim"]]

    def leaves(ref_):
        if ref_ < 0:
            return [~ref_]
        return leaves(int(ch[ref_, 0])) + leaves(int(ch[ref_, 1]))
    import sys
    sys.setrecursionlimit(10000)
    for side, (lo_c, hi_c) in enumerate(((f[0, 0:3], f[0, 4:7]), (f[0, 8:11], f[0, 12:15]))):
        ls = leaves(int(ch[0, side]))
        assert np.array_equal(lo_c, slot_lo[ls].min(0)) and np.array_equal(hi_c, slot_hi[ls].max(0))


@pytest.mark.parametrize("name", ["cornell", "material"])
def test_sah_tree_answers_like_brute_force(port, name):
    arr = {"cornell": S.cornell32, "material": lambda: S.material_scene(2)}[name]()
    osc = O.OracleScene(port, arr)
    rng = np.random.default_rng(9)
    n = 20000
    lo = arr.verts[:, :3].min(0) - 50
    hi = arr.verts[:, :3].max(0) + 50
    org = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    assert osc.set_tree(2) == "sah"
    sp, st, sa, sc = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    assert osc.set_tree(1) == "lbvh"
    lp, lt, la, lc = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    bp, bt, ba, _ = osc.trace(org, d, 0.01, 1e16, use_bvh=False, want_any=True)
    osc.set_tree(0)
    for p, t, a in ((sp, st, sa), (lp, lt, la)):
        assert np.array_equal(p, bp) and np.array_equal(t.view(np.uint32), bt.view(np.uint32)) and np.array_equal(a, ba)
    assert sc[0] < lc[0]             # the point of the exercise: fewer node visits than the Morton-order tree


def test_auto_rule(port):
    """auto = SAH for 2..8192 triangles, PLOC above (each at most 64 levels deep), LBVH for a single triangle."""
    assert O.OracleScene(port, S.cornell32()).set_tree(0) == "sah"
    one = _soup(np.random.default_rng(1), 1)
    assert O.OracleScene(port, one).set_tree(0) == "lbvh"
    big = S.tessellated_scene(70)                         # 9812 triangles
    assert big.n_tri > 8192
    osc = O.OracleScene(port, big)
    assert osc.set_tree(0) == "ploc" and osc.set_tree(2) == "lbvh" and osc.set_tree(3) == "ploc"


@pytest.mark.parametrize("name,arr", _cases() + [("tess70", S.tessellated_scene(70))], ids=lambda x: x if isinstance(x, str) else "")
def test_ploc_tree_structure_and_answers(port, name, arr):
    """The oracle's PLOC tree (the GPU build is compared with it word for word in tests/test_gpu_parity.py): n - 1 nodes,
    root 0, every leaf and every other node referenced once, child boxes = exact unions, and the same answers as the
    LBVH and as brute force."""
    osc = O.OracleScene(port, arr)
    p = osc.bvh_ploc()
    nodes = p["nodes"]
    ch = nodes[:, [3, 7]].view(np.int32)
    assert np.array_equal(np.sort(~ch[ch < 0]), np.arange(arr.n_tri))
    assert np.array_equal(np.sort(ch[ch >= 0]), np.arange(1, arr.n_tri - 1))
    assert np.all(ch[ch >= 0] > np.repeat(np.arange(arr.n_tri - 1), 2).reshape(-1, 2)[ch >= 0])   # children were created earlier
    b = osc.bvh()
    lo, hi = _boxes(arr)
    slot_lo, slot_hi = lo[b["sorted_prim"]], hi[b["sorted_prim"]]
    f = nodes.view(np.float32)
    box_lo, box_hi = np.zeros((arr.n_tri - 1, 3), np.float32), np.zeros((arr.n_tri - 1, 3), np.float32)
    for i in range(arr.n_tri - 2, -1, -1):               # children have larger ids: bottom-up in one sweep
        for side, (l, h) in enumerate(((f[i, 0:3], f[i, 4:7]), (f[i, 8:11], f[i, 12:15]))):
            c = int(ch[i, side])
            el, eh = (slot_lo[~c], slot_hi[~c]) if c < 0 else (box_lo[c], box_hi[c])
            assert np.array_equal(l, el) and np.array_equal(h, eh)
        box_lo[i] = np.minimum(f[i, 0:3], f[i, 8:11])
        box_hi[i] = np.maximum(f[i, 4:7], f[i, 12:15])
    rng = np.random.default_rng(13)
    n = 4000
    org = rng.uniform(arr.verts[:, :3].min(0) - 30, arr.verts[:, :3].max(0) + 30, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    assert osc.set_tree(3) == "ploc"
    pp, pt, pa, _ = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    bp, bt, ba, _ = osc.trace(org, d, 0.01, 1e16, use_bvh=False, want_any=True)
    assert np.array_equal(pp, bp) and np.array_equal(pt.view(np.uint32), bt.view(np.uint32)) and np.array_equal(pa, ba)
