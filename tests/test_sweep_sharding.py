"""CPU: host-side sweep logic -- range plumbing, grid arithmetic, phase.txt format, and the N > 1 sharding
path over torch.distributed (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def sweep_mod():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from bose_hubbard_phase_transition_b200 import sweep
    return sweep


def test_grid_matches_reference_plumbing():
    sw = sweep_mod()
    for name, args in [("phase_m5_fJ.txt", ("J", 1, 0, 0, 2, 1)), ("phase_m5_fU.txt", ("U", 0.5, 2, 1, 1, 0.5)),
                       ("phase_m5_fu.txt", ("u", 0.5, 2, 1, 1, 0.5)), ("phase_m6_fJ.txt", ("J", 1, 0, 0, 3, 1))]:
        lines = open(os.path.join(GOLD, name)).read().split("\n")
        g = sw.make_grid(*args)
        assert lines[0] == f"{g['fixed']} {sw.fmt(g['fixed_value'])}"
        rows = [l.split() for l in lines[1:] if l]
        assert len(rows) == g["num1"] * g["num2"]
        table = {}
        for (index, p1, p2, cJ, cU, cmu) in g["points"]:
            table[index] = (p1, p2)
        for k, row in enumerate(rows):
            assert row[0] == sw.fmt(table[k][0]) and row[1] == sw.fmt(table[k][1])
    # non-integer step: the count is int((max - min) / s) + 1 evaluated in double
    g = sw.make_grid("J", 0.1, 0, 0, 1.0, 0.1)
    assert g["num1"] == int((0.1 + 1.0 - 0.1) / 0.1) + 1


def test_variance_rule_and_format():
    sw = sweep_mod()
    assert sw.needs_more_eigenvalues([0.2, 0.2, 0.2])
    assert not sw.needs_more_eigenvalues([0.2, 0.25, 0.1])
    assert sw.fmt(0.0750343123) == "0.0750343" and sw.fmt(1.0) == "1" and sw.fmt(1.5) == "1.5"


def fake_point(cJ, cU, cmu, nb):
    return (np.sin(cJ + 2 * cU) ** 2, 0.5 + 0.01 * cU, 0.25 * cJ + 0.125 * cmu)


def fake_points(cJ, cU, cmu, nb):
    return np.array([fake_point(a, b, c, nb) for a, b, c in zip(cJ, cU, cmu)])


def test_chunked_evaluation_equals_point_by_point():
    sw = sweep_mod()
    g = sw.make_grid("J", 1, 0, 0, 4, 1)
    serial = sw.run_sweep(fake_point, g)
    for chunk in (1, 3, 16, 100):
        assert np.array_equal(sw.run_sweep(fake_point, g, points_fn=fake_points, chunk=chunk), serial)
    for world in (2, 3):   # the shards of every rank, evaluated in chunks, cover the grid exactly once
        seen = sorted(t for r in range(world) for t in sw.shard(len(g["points"]), world, r))
        assert seen == list(range(len(g["points"])))


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sw = sweep_mod()
    g = sw.make_grid("J", 1, 0, 0, 4, 1)
    rows = sw.run_sweep(fake_point, g, world, rank, dist, points_fn=fake_points if rank == 0 else None, chunk=4)
    q.put((rank, rows))
    dist.destroy_process_group()


def test_two_rank_sweep_equals_serial():
    sw = sweep_mod()
    g = sw.make_grid("J", 1, 0, 0, 4, 1)
    serial = sw.run_sweep(fake_point, g)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert np.array_equal(got[0], serial) and np.array_equal(got[1], serial)
    assert sorted(sw.shard(25, 2, 0) + sw.shard(25, 2, 1)) == list(range(25))


def test_read_phase_round_trip(tmp_path):
    """Reader side of the phase.txt contract: every committed reference file parses, write -> read is the identity at the
    file's 6-digit precision, and a half-written file (a sweep interrupted under --resume) gives NaN for the missing points."""
    sw = sweep_mod()
    for name, fixed, shape in [("phase_m5_fJ.txt", "J", (3, 3)), ("phase_m5_fU.txt", "U", (3, 3)), ("phase_m5_fu.txt", "u", None),
                               ("phase_m8_C1.txt", "J", (11, 11)), ("phase_m10_fJ.txt", "J", None)]:
        ph = sw.read_phase(os.path.join(GOLD, name))
        assert ph["fixed"] == fixed and ph["axes"] == sw.AXES[fixed]
        if shape:
            assert ph["gap_ratio"].shape == shape and not np.isnan(ph["gap_ratio"]).any()
        # row k = i1 * n2 + i2 in writing order (full rectangular grids only)
        n1, n2 = len(ph["p1"]), len(ph["p2"])
        if len(ph["rows"]) == n1 * n2:
            assert np.array_equal(ph["coherence"].reshape(-1), ph["rows"][:, 4])
            assert np.array_equal(ph["condensate_fraction"].reshape(-1), ph["rows"][:, 3])
        out = tmp_path / name
        sw.write_phase(str(out), {"fixed": ph["fixed"], "fixed_value": ph["fixed_value"]}, ph["rows"])
        assert open(out).read() == open(os.path.join(GOLD, name)).read()
    text = open(os.path.join(GOLD, "phase_m8_C1.txt")).read().split("\n")
    part = tmp_path / "partial.txt"
    part.write_text("\n".join(text[:1 + 60]) + "\n")
    ph = sw.read_phase(str(part))
    assert len(ph["rows"]) == 60 and np.isnan(ph["gap_ratio"]).sum() == ph["gap_ratio"].size - 60
    bad = tmp_path / "bad.txt"
    bad.write_text("X 1\n1 2 3 4 5\n")
    with pytest.raises(ValueError):
        sw.read_phase(str(bad))
