#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the real (patched P1..P6) reference compiled into
oracle/_ref (run `make -C oracle ref` first; needs /root/reference).  The reference's own test-suite holds no
vectors for this path, so these outputs of the reference itself are the pins.

    python tests/golden/make_golden.py            # everything up to m = n = 10   (~1 min)
    python tests/golden/make_golden.py --m12      # also one m = n = 12 point     (~5 min)
    python tests/golden/make_golden.py --points "12,12,1,16,3,chain;12,5,1,4,1,rect:4:3" --part /tmp/p1.npz
    python tests/golden/make_golden.py --merge /tmp/p1.npz /tmp/p2.npz   # fold part files into the fixtures
(--points runs only the listed grid points (m,n,cJ,cU,cu,lattice) and writes them to a part file, so that the
slow m = n = 12 points can run as concurrent processes; --merge folds part files into reference_golden.npz.)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_lib as R  # noqa: E402
from checksums import checksum_basis, checksum_csc  # noqa: E402

assert R.available(), "oracle/_ref is missing: make -C oracle ref"
gold = {}
OUT = os.path.join(HERE, "reference_golden.npz")
MFILE = os.path.join(HERE, "reference_golden_meta.json")


def point_key(m, n, cJ, cU, cu, lat="chain"):
    key = f"point_{m}_{n}_{cJ:g}_{cU:g}_{cu:g}"
    return key if lat == "chain" else key + "_" + lat.replace(":", "-")


if "--merge" in sys.argv:
    old = dict(np.load(OUT))
    meta = json.load(open(MFILE))
    for f in sys.argv[sys.argv.index("--merge") + 1:]:
        old.update(dict(np.load(f)))
        meta.update(json.load(open(f + ".json")))
    np.savez_compressed(OUT, **old)
    json.dump(meta, open(MFILE, "w"), indent=1, sort_keys=True)
    sys.exit(0)

if "--dense" in sys.argv:
    # --dense: for every golden grid point with D <= 5000, the 24 lowest eigenvalues by DENSE diagonalisation (numpy) of the
    # same H (built by the oracle, whose H is bit-identical to the reference's: tests/test_oracle_vs_ref.py).  They show
    # where the reference's single-vector Krylov solver misses copies of (>= 3-fold) degenerate levels.
    import oracle_lib as O
    old = dict(np.load(OUT))
    for k in [k for k in old if k.startswith("point_") and k.endswith("_evals")]:
        f = k[len("point_"):-len("_evals")].split("_")
        m, n, cJ, cU, cu = int(f[0]), int(f[1]), float(f[2]), float(f[3]), float(f[4])
        lat = f[5] if len(f) > 5 else "chain"
        if O.dimension(m, n) > 5000:
            continue
        nbr = O.chain(m) if lat == "chain" else O.rect(*[int(v) for v in lat.split("-")[1:]])
        t, b = O.basis(m, n)
        o, i, v = O.hsum_csc(O.hopping_csc(m, nbr, t, b), *O.diagonals(m, b), cJ, cU, cu)
        D = len(t)
        H = np.zeros((D, D))
        for c in range(D):
            H[i[o[c]:o[c + 1]], c] = v[o[c]:o[c + 1]]
        old[k[:-len("_evals")] + "_dense"] = np.linalg.eigvalsh(H)[:24]
        print(k, "reference == dense:", bool(np.abs(np.sort(old[k]) - old[k[:-len("_evals")] + "_dense"][:20]).max() < 1e-9))
    np.savez_compressed(OUT, **old)
    sys.exit(0)

if "--grid" in sys.argv:
    # --grid "m,n,fixed,cfix,p1min,p2min,step,n1,n2,lattice" --part FILE: a small sweep grid through the reference's
    # per-point body (ref_harness points), for lattices the reference CLI cannot build itself
    part = sys.argv[sys.argv.index("--part") + 1]
    f = sys.argv[sys.argv.index("--grid") + 1].split(",")
    m, n, fixed, lat = int(f[0]), int(f[1]), f[2], f[9]
    r, info = R.points(m, n, fixed, float(f[3]), float(f[4]), float(f[5]), float(f[6]), int(f[7]), int(f[8]), threads=8, lattice=lat)
    key = f"grid_{m}_{n}_{lat.replace(':', '-')}"
    np.savez_compressed(part, **{key + "_out5": r["out5"], key + "_evals": r["evals"]})
    json.dump({key: info}, open(part + ".json", "w"), indent=1, sort_keys=True)
    print(key, info, r["out5"][:3])
    sys.exit(0)

if "--points" in sys.argv:
    part = sys.argv[sys.argv.index("--part") + 1]
    meta = {}
    for spec in sys.argv[sys.argv.index("--points") + 1].split(";"):
        f = spec.split(",")
        m, n, cJ, cU, cu, lat = int(f[0]), int(f[1]), float(f[2]), float(f[3]), float(f[4]), f[5]
        r, info = R.eigs(m, n, cJ, cU, cu, lattice=lat)
        key = point_key(m, n, cJ, cU, cu, lat)
        gold[key + "_evals"], gold[key + "_rho"], gold[key + "_out5"] = r["evals"], r["rho"], r["out5"]
        meta[key] = info
        print(key, info, flush=True)
    np.savez_compressed(part, **gold)
    json.dump(meta, open(part + ".json", "w"), indent=1, sort_keys=True)
    sys.exit(0)

# --- basis / tags, both orders (small shapes; the large ones are pinned by checksums below) ---
for (m, n) in [(3, 2), (4, 4), (5, 3)]:
    for unp in (False, True):
        t, b, _ = R.basis(m, n, unpatched=unp)
        key = f"basis_{m}_{n}_{'scatter' if unp else 'sorted'}"
        gold[key + "_tags"] = t
        gold[key + "_states"] = b.astype(np.int8)
for (m, n) in [(6, 6), (8, 8), (10, 10)]:
    for unp in (False, True):
        t, b, _ = R.basis(m, n, unpatched=unp)
        key = f"basissum_{m}_{n}_{'scatter' if unp else 'sorted'}"
        gold[key] = checksum_basis(t, b)

# --- CSC of the three terms and of H ---
for (m, n, lat) in [(3, 2, "chain"), (4, 4, "chain"), (5, 5, "chain"), (2, 5, "chain"), (6, 3, "rect:3:2"),
                    (4, 3, "rect:2:2")]:
    tag = f"{m}_{n}_{lat.replace(':', '-')}"
    for term in "JUu":
        (o, i, v), _ = R.csc(m, n, term, lat)
        gold[f"csc_{term}_{tag}_outer"], gold[f"csc_{term}_{tag}_inner"], gold[f"csc_{term}_{tag}_val"] = o, i, v
    (o, i, v), _ = R.hsum(m, n, 1.0, 4.0, 1.0, lat)
    gold[f"hsum_{tag}_outer"], gold[f"hsum_{tag}_inner"], gold[f"hsum_{tag}_val"] = o, i, v
for (m, n) in [(8, 8), (10, 10)]:
    (o, i, v), info = R.csc(m, n, "J")
    gold[f"cscsum_J_{m}_{n}"] = checksum_csc(o, i, v)

# --- variable-N API: max_set_basis / max_bosons_hamiltonian ---
t, b = R.max_basis(4, 3)
gold["maxbasis_4_3_tags"], gold["maxbasis_4_3_states"] = t, b.astype(np.int8)
for term, (J, U, mu) in {"J": (1, 0, 0), "U": (0, 1, 0), "u": (0, 0, 1)}.items():
    o, i, v = R.max_hamiltonian(4, 1, 3, J, U, mu)
    gold[f"maxham_{term}_4_1_3_outer"], gold[f"maxham_{term}_4_1_3_inner"], gold[f"maxham_{term}_4_1_3_val"] = o, i, v

# ... including the empty sector (n_min = 0: a 1 x 1 block with the explicit entry U * 0 / -mu * 0, src/hamiltonian.cpp:268)
for term, (J, U, mu) in {"J": (1, 0, 0), "U": (0, 1.5, 0), "u": (0, 0, 0.75)}.items():
    o, i, v = R.max_hamiltonian(4, 0, 2, J, U, mu)
    gold[f"maxham_{term}_4_0_2_outer"], gold[f"maxham_{term}_4_0_2_inner"], gold[f"maxham_{term}_4_0_2_val"] = o, i, v

# --- finite-temperature branch (src/analysis.cpp:474-494): dense density matrix of the 20 Ritz pairs ---
for (m, n, T) in [(5, 5, 0.5), (5, 5, 3.0), (6, 4, 1.0)]:
    r = R.thermal(m, n, 1.0, 4.0, 1.0, T)
    key = f"thermal_{m}_{n}_{T:g}"
    gold[key + "_dm"], gold[key + "_evals"], gold[key + "_out2"] = r["dm"], r["evals"], r["out2"]

# --- H.v through Spectra's MatOp ---
for (m, n) in [(6, 6), (8, 8)]:
    r, _ = R.hv(m, n, 1.0, 4.0, 1.0)
    gold[f"hv_{m}_{n}_x"], gold[f"hv_{m}_{n}_y"] = r["x"], r["y"]

# --- grid points: eigenvalues, rho, out5 ---
points = [(5, 5, 1, 4, 1), (5, 5, 0.5, 1, 0), (6, 6, 1, 4, 1), (6, 6, 1, 0.5, 0), (7, 6, 1, 2, 3), (8, 8, 1, 4, 1),
          (8, 8, 1, 1, 0), (8, 8, 1, 10, 0), (10, 10, 1, 4, 1), (10, 10, 1, 1, 0)]
if "--m12" in sys.argv:
    points = [(12, 12, 1, 4, 1)]
meta = {}
for (m, n, cJ, cU, cu) in points:
    r, info = R.eigs(m, n, cJ, cU, cu)
    key = f"point_{m}_{n}_{cJ}_{cU}_{cu}"
    gold[key + "_evals"], gold[key + "_rho"], gold[key + "_out5"] = r["evals"], r["rho"], r["out5"]
    meta[key] = info
    print(key, info, flush=True)
out = os.path.join(HERE, "reference_golden.npz")
mfile = os.path.join(HERE, "reference_golden_meta.json")
if "--m12" in sys.argv:
    old = dict(np.load(out))
    old.update({k: v for k, v in gold.items() if k.startswith("point_12")})
    gold = old
    meta = {**json.load(open(mfile)), **meta}
np.savez_compressed(out, **gold)
json.dump(meta, open(mfile, "w"), indent=1, sort_keys=True)
if "--m12" in sys.argv:
    sys.exit(0)

# --- phase.txt of the patched CLI (text) ---
runs = {
    "phase_m5_fJ.txt": ["-m", 5, "-n", 5, "-J", 1, "-U", 0, "-u", 0, "-r", 2, "-s", 1, "-f", "J", "-t", "exact"],
    "phase_m5_fU.txt": ["-m", 5, "-n", 5, "-J", 0.5, "-U", 2, "-u", 1, "-r", 1, "-s", 0.5, "-f", "U", "-t", "exact"],
    "phase_m5_fu.txt": ["-m", 5, "-n", 5, "-J", 0.5, "-U", 2, "-u", 1, "-r", 1, "-s", 0.5, "-f", "u", "-t", "exact"],
    "phase_m6_fJ.txt": ["-m", 6, "-n", 6, "-J", 1, "-U", 0, "-u", 0, "-r", 3, "-s", 1, "-f", "J", "-t", "exact"],
    "phase_m8_fJ.txt": ["-m", 8, "-n", 8, "-J", 1, "-U", 0, "-u", 0, "-r", 2, "-s", 1, "-f", "J", "-t", "exact"],
    # config 1 of BASELINE.json in full (121 points) and a 3 x 3 corner of config 2 (85 s: the reference's thread
    # heuristic leaves it single-threaded at m = n = 10)
    "phase_m8_C1.txt": ["-m", 8, "-n", 8, "-J", 1, "-U", 0, "-u", 0, "-r", 10, "-s", 1, "-f", "J", "-t", "exact"],
    "phase_m10_fJ.txt": ["-m", 10, "-n", 10, "-J", 1, "-U", 0, "-u", 0, "-r", 2, "-s", 1, "-f", "J", "-t", "exact"],
}
for name, args in runs.items():
    open(os.path.join(HERE, name), "w").write(R.cli_phase(args, threads=8))
    print("wrote", name, flush=True)
