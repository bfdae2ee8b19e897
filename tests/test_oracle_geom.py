"""The oracle's own LBVH / traversal statement (oracle_geom.h): structural invariants and
BVH traversal == brute force over all triangles (test 5)."""
import numpy as np
import pytest

import oracle_api as O
from acgpathtracing_b200 import scenes as S


def _rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo = scene.verts[:, :3].min(0) - 50
    hi = scene.verts[:, :3].max(0) + 50
    org = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # a share of exactly axis-aligned and surface-origin rays (the path tracer's bounce rays start on surfaces)
    d[: n // 20] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 20)] * rng.choice([-1, 1], (n // 20, 1))
    tri = scene.indices[rng.integers(0, scene.n_tri, n // 4)]
    bary = rng.dirichlet((1, 1, 1), n // 4)
    p = (scene.verts[tri[:, 0], :3] * bary[:, :1] + scene.verts[tri[:, 1], :3] * bary[:, 1:2] +
         scene.verts[tri[:, 2], :3] * bary[:, 2:3])
    org[n // 2: n // 2 + n // 4] = p.astype(np.float32)
    return org, d


@pytest.fixture(scope="module", params=["cornell", "material", "tess40"])
def any_scene(request):
    return {"cornell": S.cornell32, "material": lambda: S.material_scene(2),
            "tess40": lambda: S.tessellated_scene(40)}[request.param]()


def test_lbvh_structure(port, any_scene):
    sc = O.OracleScene(port, any_scene)
    b = sc.bvh()
    nT = any_scene.n_tri
    keys = b["sorted_keys"]
    assert np.all(keys[1:] > keys[:-1])                                     # strictly ascending, unique
    assert np.array_equal(np.sort(b["sorted_prim"]), np.arange(nT, dtype=np.uint32))
    assert np.array_equal((keys >> np.uint64(32)).astype(np.uint32), b["morton"][b["sorted_prim"]])
    assert np.all(b["morton"] < (1 << 30))
    nodes = b["nodes"]
    assert nodes.shape[0] == nT - 1
    c0 = nodes[:, 3].view(np.int32)
    c1 = nodes[:, 7].view(np.int32)
    # every internal node except the root and every leaf is referenced exactly once
    refs = np.concatenate([c0, c1])
    internal = np.sort(refs[refs >= 0])
    leaves = np.sort(~refs[refs < 0])
    assert np.array_equal(internal, np.arange(1, nT - 1))
    assert np.array_equal(leaves, np.arange(nT))
    # child boxes of a node contain the boxes stored in the child node
    f = nodes.view(np.float32)
    lo0, hi0, lo1, hi1 = f[:, 0:3], f[:, 4:7], f[:, 8:11], f[:, 12:15]
    for child, lo, hi in ((c0, lo0, hi0), (c1, lo1, hi1)):
        m = child >= 0
        cl = np.minimum(lo0[child[m]], lo1[child[m]])
        ch = np.maximum(hi0[child[m]], hi1[child[m]])
        assert np.array_equal(cl, lo[m]) and np.array_equal(ch, hi[m])
    # leaf boxes == triangle AABBs; triangle records are (v0, e1, e2) in leaf order
    tris = b["tris"]
    assert np.array_equal(tris[:, 3], b["sorted_prim"])
    v = any_scene.verts[:, :3]
    idx = any_scene.indices[b["sorted_prim"]]
    tf = tris.view(np.float32)
    assert np.array_equal(tf[:, 0:3], v[idx[:, 0]])
    assert np.array_equal(tf[:, 4:7], v[idx[:, 1]] - v[idx[:, 0]])
    assert np.array_equal(tf[:, 8:11], v[idx[:, 2]] - v[idx[:, 0]])
    assert np.array_equal(tris[:, 7], any_scene.mat_ids[b["sorted_prim"]])


def test_bvh_equals_brute_force(port, any_scene):
    sc = O.OracleScene(port, any_scene)
    n = 200_000 if any_scene.n_tri < 2000 else 60_000
    org, d = _rays(any_scene, n, 7)
    pb, tb, ab, _ = sc.trace(org, d, 0.01, 1e16, use_bvh=False, want_any=True)
    pv, tv, av, ctr = sc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    assert np.array_equal(pb, pv)
    assert np.array_equal(tb.view(np.uint32), tv.view(np.uint32))
    assert np.array_equal(ab, av)
    assert (pb >= 0).mean() > 0.3
    # bounded shadow-style queries
    tmax = np.random.default_rng(3).uniform(1, 600, n).astype(np.float32)
    _, _, ab, _ = sc.trace(org, d, 0.01, tmax, use_bvh=False, want_any=True)
    _, _, av, _ = sc.trace(org, d, 0.01, tmax, use_bvh=True, want_any=True)
    assert np.array_equal(ab, av)


def test_single_triangle_and_empty(port):
    one = S.SceneArrays(np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]], np.float32),
                        np.array([[0, 1, 2]], np.uint32), np.zeros(1, np.uint32),
                        [dict(name="m", Kd=(1, 1, 1), Ke=(0, 0, 0))])
    sc = O.OracleScene(port, one)
    org = np.array([[0.2, 0.2, -1], [2, 2, -1]], np.float32)
    d = np.array([[0, 0, 1], [0, 0, 1]], np.float32)
    p, t, a, _ = sc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    assert p.tolist() == [0, -1] and t[0] == 1.0 and a.tolist() == [1, 0]
    empty = S.SceneArrays(np.ones((3, 4), np.float32), np.zeros((0, 3), np.uint32), np.zeros(0, np.uint32),
                          [dict(name="m", Kd=(1, 1, 1), Ke=(0, 0, 0))])
    sc = O.OracleScene(port, empty)
    p, t, a, _ = sc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    assert p.tolist() == [-1, -1] and a.tolist() == [0, 0]
