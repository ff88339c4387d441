#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the real (patched P1..P6) reference compiled into
oracle/_ref (run `make -C oracle ref` first; needs /root/reference).  The reference's own test-suite holds no
vectors for this path, so these outputs of the reference itself are the pins.

    python tests/golden/make_golden.py            # everything up to m = n = 10   (~1 min)
    python tests/golden/make_golden.py --m12      # also one m = n = 12 point     (~5 min)
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
