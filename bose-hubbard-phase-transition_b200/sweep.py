"""Host-side sweep logic shared by bench.py, the multi-process driver (tools/sweep_mgpu.py) and the tests:
the reference's range plumbing / grid arithmetic / phase.txt format, and the sharding of the grid over ranks.

Follows Analysis::exact_parameters and calculate_and_save (reference src/analysis.cpp:242-256, :281-282,
:303-308, :341, :364-387).  The C++ host layer (host/analysis.cpp) implements the same logic for the CLI;
tests/test_shim.py checks both against the golden phase.txt files.
"""
import numpy as np


def make_grid(fixed, J, U, mu, r, s):
    """-> dict(fixed, fixed_value, num1, num2, points=[(index, p1, p2, cJ, cU, cmu)] in loop order)."""
    J_min, J_max, mu_min, mu_max, U_min, U_max = J, J + r, mu, mu + r, U, U + r
    if fixed == "J":
        fv, p1min, p1max, p2min, p2max = J, J_min, J_max, mu_min, mu_max
    elif fixed == "U":
        fv, p1min, p1max, p2min, p2max = U, J_min, J_max, U_min, U_max   # sic
    elif fixed == "u":
        fv, p1min, p1max, p2min, p2max = mu, J_min, J_max, mu_min, mu_max
    else:
        raise ValueError("fixed parameter must be J, U or u.")
    num1 = int((p1max - p1min) / s) + 1
    num2 = int((p2max - p2min) / s) + 1
    pts = []
    for i in range(num1):
        for j in range(num2):
            p1 = p1min + i * s
            p2 = p2min + j * s
            if fixed == "J":
                c = (fv, p1, p2)
            elif fixed == "U":
                c = (p1, fv, p2)
            else:
                c = (p1, p2, fv)
            pts.append((i * num1 + j, p1, p2) + c)   # index as in src/analysis.cpp:341
    return dict(fixed=fixed, fixed_value=fv, num1=num1, num2=num2, points=pts)


def shard(npoints, world, rank):
    """Static interleave of the loop-order point list over ranks (no data-path collective is needed)."""
    return list(range(rank, npoints, world))


def needs_more_eigenvalues(gap_ratios):
    """The variance-restart rule of src/analysis.cpp:364-380: repeat with nb_eigen += 5 unless
    var(gap_ratio) > 1e-8 * mean(gap_ratio)."""
    g = np.asarray(gap_ratios, dtype=np.float64)
    mean = g.sum() / len(g)
    var = ((g - mean) ** 2).sum() / len(g)
    return not (var > 1e-8 * mean)


def run_sweep(point_fn, grid, world=1, rank=0, dist=None, nb_eigen=20, points_fn=None, chunk=16):
    """point_fn(cJ, cU, cmu, nb_eigen) -> (gap_ratio, condensate_fraction, coherence) for the local shard;
    results are exchanged once at the end (all_gather of 3 doubles per point).  Returns the rows of phase.txt.
    points_fn(cJ[], cU[], cmu[], nb_eigen) -> array[n, 3], when given, evaluates the shard in chunks of `chunk` points
    (bh_points: the points of a chunk are solved in lockstep and share their H.v launches; same results)."""
    pts = grid["points"]
    total = grid["num1"] * grid["num2"]
    while True:
        mine = shard(len(pts), world, rank)
        local = np.zeros((len(mine), 4))
        if points_fn is not None:
            for k0 in range(0, len(mine), max(1, chunk)):
                sel = mine[k0:k0 + max(1, chunk)]
                c = np.array([pts[t][3:6] for t in sel], dtype=np.float64).reshape(len(sel), 3)
                local[k0:k0 + len(sel), 0] = sel
                local[k0:k0 + len(sel), 1:] = np.asarray(points_fn(c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy(), nb_eigen)).reshape(len(sel), 3)
        for k, t in enumerate(mine if points_fn is None else []):
            _, _, _, cJ, cU, cmu = pts[t]
            local[k, 0] = t
            local[k, 1:] = point_fn(cJ, cU, cmu, nb_eigen)
        if world > 1:
            import torch
            cnt = (len(pts) + world - 1) // world
            buf = torch.full((cnt, 4), -1.0, dtype=torch.float64)
            buf[: len(mine)] = torch.from_numpy(local)
            if dist.get_backend() == "nccl":
                buf = buf.cuda()
            out = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(out, buf)
            allr = torch.cat([o.cpu() for o in out]).numpy()
            allr = allr[allr[:, 0] >= 0]
        else:
            allr = local
        rows = np.zeros((total, 5))
        for t, g, c, k in allr:
            index, p1, p2 = pts[int(t)][:3]
            if 0 <= index < total:
                rows[index] = (p1, p2, g, c, k)
        if not needs_more_eigenvalues(rows[:, 2]):
            return rows
        nb_eigen += 5


def fmt(x):
    """C++ default ostream formatting of a double (precision 6)."""
    return "%g" % x


def write_phase(path, grid, rows):
    with open(path, "w") as f:
        f.write(f"{grid['fixed']} {fmt(grid['fixed_value'])}\n")
        for r in rows:
            f.write(" ".join(fmt(v) for v in r) + "\n")


AXES = {"J": ("U", "mu"), "U": ("J", "mu"), "u": ("J", "U")}


def read_phase(path):
    """Reader side of the phase.txt contract (the reference's only consumer is plot.py:15-64).

    Line 1: "<fixed parameter: J|U|u> <its value>"; then one row per grid point, "p1 p2 gap_ratio condensate_fraction
    coherence", p1 outer / p2 inner (src/analysis.cpp:384-387).  Returns the header, the two axes and the three observable
    grids indexed [i1, i2] in WRITING order (plot.py reshapes to (len(y), len(x)), which only agrees with this for square
    grids).  An incomplete file (a sweep still running under --resume) yields NaN for the missing points.
    """
    import numpy as np
    with open(path) as f:
        head = f.readline().split()
        if len(head) != 2 or head[0] not in AXES:
            raise ValueError("Invalid fixed parameter in phase.txt")
        rows = [[float(v) for v in line.split()] for line in f if line.strip()]
    data = np.array(rows, dtype=float).reshape(-1, 5)
    x = np.unique(data[:, 0])
    y = np.unique(data[:, 1])
    grids = np.full((3, len(x), len(y)), np.nan)
    ix = np.searchsorted(x, data[:, 0])
    iy = np.searchsorted(y, data[:, 1])
    for c in range(3):
        grids[c, ix, iy] = data[:, 2 + c]
    return {"fixed": head[0], "fixed_value": float(head[1]), "axes": AXES[head[0]], "p1": x, "p2": y, "rows": data,
            "gap_ratio": grids[0], "condensate_fraction": grids[1], "coherence": grids[2]}
