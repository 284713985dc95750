import numpy as np

INT_FIELDS = ("count1", "count2", "verts1", "verts2", "tri1", "tri2", "edge1", "n_edge1", "edge2")
REAL_FIELDS = ("dist", "nor1", "nor2", "pos1", "pos2", "pos1_", "weights1", "weights2", "edgeDir")


def block_row_scale(outer, vals, n_nodes):
    """s = max |entry| over the 3 rows (== columns, symmetric) of each node, broadcast to every entry.  With Eulerian dofs behind
    the 3 n_nodes Lagrangian ones (EOL meshes, dof = 3N + 2 EoL_Count) the two columns of an EoL node form a group of their own."""
    outer = np.asarray(outer, dtype=np.int64)
    counts = np.diff(outer)
    col_of = np.repeat(np.arange(outer.size - 1), counts)
    node_of = np.where(col_of < 3 * n_nodes, col_of // 3, n_nodes + (col_of - 3 * n_nodes) // 2)
    s = np.zeros(n_nodes + max(outer.size - 1 - 3 * n_nodes, 0) // 2)
    np.maximum.at(s, node_of, np.abs(vals))
    return s[node_of]


def assert_close_tol(a, b, scale, tol=1e-10, what=""):
    """SURVEY §8c rule 5: |a-b| <= tol * max(|a|, |b|, s)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bound = tol * np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    err = np.abs(a - b)
    bad = err > bound
    assert not bad.any(), f"{what}: {bad.sum()} entries outside tol; worst ratio {np.max(err / np.maximum(bound, 1e-300)):.3e}"


def assert_contacts_equal(got, ref, real_tol=1e-12, what=""):
    """Count, order and every integer field bit-exact; real fields to 1e-12 relative (SURVEY §8c rule 5)."""
    assert len(got) == len(ref), f"{what}: {len(got)} contacts vs {len(ref)}"
    for f in INT_FIELDS:
        assert np.array_equal(got[f], ref[f]), f"{what}: integer field {f} differs at {np.nonzero(np.any(np.atleast_2d((got[f] != ref[f]).T).T.reshape(len(got), -1), axis=1))[0][:5]}"
    for f in REAL_FIELDS:
        a, b = got[f], ref[f]
        scale = np.maximum(np.abs(b).max(initial=0.0), 1e-300)
        err = np.abs(a - b).max(initial=0.0)
        assert err <= real_tol * max(scale, 1.0), f"{what}: real field {f} differs by {err:.3e}"


def bit_identical_fraction(got, ref):
    n = 0; tot = 0
    for f in REAL_FIELDS:
        n += np.count_nonzero(got[f].view(np.uint64) == ref[f].view(np.uint64)); tot += got[f].size
    return n / max(tot, 1)


def fan_mesh(k, seed=0):
    """One centre node joined to a ring of k nodes: valence-k node (rows with k + 1 blocks, k faces and k bending stencils on one node)."""
    import eol_cloth_b200 as E
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    X = np.r_[[[0.0, 0.0]], np.c_[np.cos(ang), np.sin(ang)]]
    fn = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], np.int32)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = np.c_[X, 0.05 * np.sin(3 * X[:, 0])] + 1e-3 * np.random.default_rng(seed + k).standard_normal((X.shape[0], 3))
    return dict(x=x, X=X, face_nodes=fn, edge_stencil=es)


def strip_mesh(m):
    """A 2 x m strip: every node on the boundary, bending stencils only across the rungs and diagonals."""
    import eol_cloth_b200 as E
    X = np.array([[i / (m - 1), j * 0.02] for i in range(m) for j in range(2)])
    fn = []
    for i in range(m - 1):
        a = 2 * i
        fn += [[a, a + 2, a + 1], [a + 1, a + 2, a + 3]]
    fn = np.array(fn, np.int32)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    return dict(x=np.c_[X, 0.05 * np.sin(3 * X[:, 0])], X=X, face_nodes=fn, edge_stencil=es)
