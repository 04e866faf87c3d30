#!/usr/bin/env python
"""Golden vectors from the UNMODIFIED reference ikd-Tree (oracle/_ref/libikd_ref.so, compiled from
/root/reference/eskf_lio/include/ikd-Tree/ikd_Tree.cpp by oracle/Makefile): k = 5 neighbours of seeded queries, and the
tree contents after a seeded sequence of downsample-inserts, raw inserts and box deletes.  The reference has no test
fixtures of its own; these pin the oracle's independent restatement (PortMap) and the device map to what
the reference code returned when this script was run, wherever the reference itself cannot be rebuilt.

    python tests/golden/make_golden.py        # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def cloud(n, seed, lo=-20.0, hi=20.0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    p[:, 2] = (rng.uniform(-1.0, 3.0, n)).astype(np.float32)
    return np.column_stack([p, rng.uniform(1, 100, n).astype(np.float32)])


def steps():
    """the seeded mutation sequence shared by this script and tests/test_golden.py"""
    out = []
    for s in range(4):
        out.append(("add_ds", cloud(1500, 100 + s, -22, 22)))
        out.append(("add_raw", cloud(120, 200 + s, -22, 22)))
        if s % 2 == 1:
            out.append(("delete", np.array([[-22, -22, -2, -22 + 4.0 * s, 22, 5]], np.float32)))
    return out


CHECKPOINTS = (2, 5, 9)  # mutation steps after which the whole tree contents are stored (the live count is stored for all)


def main():
    import oracle_binding as ob

    orc = ob.load()
    assert orc.ref_ok, "the reference build (oracle/_ref) is required to make golden vectors"
    m = orc.new_map(ob.MAP_REF, 0.5)
    base = cloud(6000, 1)
    m.build(base)
    q = cloud(400, 2, -21, 21)[:, :3]
    pts0, d0, c0 = m.knn(q)
    contents = []
    for kind, arg in steps():
        if kind == "add_ds":
            m.add(arg, True)
        elif kind == "add_raw":
            m.add(arg, False)
        else:
            m.delete_boxes(arg)
        f = m.flatten()
        contents.append(f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))])
    pts1, d1, c1 = m.knn(q)
    np.savez_compressed(os.path.join(HERE, "ikd_tree_reference.npz"), knn0_pts=pts0[:, :, :3], knn0_d2=d0, knn0_cnt=c0, knn1_pts=pts1[:, :, :3],
                        knn1_d2=d1, knn1_cnt=c1, **{f"contents_{i}": contents[i] for i in CHECKPOINTS}, counts=np.array([len(c) for c in contents]))
    print("wrote", os.path.join(HERE, "ikd_tree_reference.npz"), [len(c) for c in contents])


if __name__ == "__main__":
    main()
