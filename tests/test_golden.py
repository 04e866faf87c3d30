"""Golden vectors recorded from the unmodified reference ikd-Tree (tests/golden/make_golden.py): the oracle's independent
restatement (PortMap) and the device voxel-hash map must reproduce the neighbour sets and the tree contents the reference
returned -- bit for bit -- without the reference build being present."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as gold  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "ikd_tree_reference.npz"))


def srt(f):
    f = np.asarray(f, np.float32).reshape(-1, 4)
    return f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))]


def check_knn(pts, d2, cnt, tag):
    np.testing.assert_array_equal(cnt, G[f"{tag}_cnt"])
    np.testing.assert_array_equal(d2, G[f"{tag}_d2"])
    np.testing.assert_array_equal(pts[:, :, :3], G[f"{tag}_pts"])


def test_port_map_reproduces_reference_golden(oracle):
    import oracle_binding as ob

    m = oracle.new_map(ob.MAP_PORT, 0.5)
    m.build(gold.cloud(6000, 1))
    q = gold.cloud(400, 2, -21, 21)[:, :3]
    check_knn(*m.knn(q), "knn0")
    for i, (kind, arg) in enumerate(gold.steps()):
        if kind == "add_ds":
            m.add(arg, True)
        elif kind == "add_raw":
            m.add(arg, False)
        else:
            m.delete_boxes(arg)
        assert m.validnum() == int(G["counts"][i]), i
        if i in gold.CHECKPOINTS:
            np.testing.assert_array_equal(srt(m.flatten()), G[f"contents_{i}"])
    check_knn(*m.knn(q), "knn1")


def test_device_map_reproduces_reference_golden(dev):
    from daliti_b200.binding import ScanToMap

    lib, _ = dev
    dm = ScanToMap(lib, max_scan_points=8192, max_map_points=1 << 16)
    dm.map_build(gold.cloud(6000, 1))
    q = gold.cloud(400, 2, -21, 21)[:, :3]
    pts, d2, cnt = dm.map_knn(q)
    check_knn(pts, d2, cnt, "knn0")
    for i, (kind, arg) in enumerate(gold.steps()):
        if kind == "add_ds":
            dm.map_add(arg, True)
        elif kind == "add_raw":
            dm.map_add(arg, False)
        else:
            dm.map_delete_boxes(arg)
        assert dm.map_valid_count() == int(G["counts"][i]), i
        if i in gold.CHECKPOINTS:
            np.testing.assert_array_equal(srt(dm.map_export()), G[f"contents_{i}"])
    pts, d2, cnt = dm.map_knn(q)
    check_knn(pts, d2, cnt, "knn1")
    dm.close()
