"""Input structures for MADDY runs: 13-protofilament microtubule lattices as PDB pairs.

The reference ships its inputs as data files (initial/xyz_N.pdb + ang_N.pdb) produced by its
offline Python-2 tools (scripts/make_mt.py, scripts/make_mt_cncntr.py).  The lattice is a closed
formula, so this module regenerates the same files (tests/test_structures.py checks them
byte-for-byte against /root/reference/initial when that tree is present) instead of copying
data, and adds the synthetic systems the benchmark configs need.

Geometry (reference scripts/make_mt.py:4-9): monomer radius 2.0 nm, tube radius 8.12 nm,
13 protofilaments 2*pi/13 apart, helical rise 6/13*r_mon per protofilament, 4 nm per monomer
along a protofilament; the `ang` file stores (fi, psi, theta) as (x, y, z).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path
from typing import List, Tuple

import numpy as np

R_MON = 2.0
R_MT = 8.12
N_PF = 13
ALPHA_STEP = 2 * math.pi / 13.0
Z_STEP = 6.0 / 13.0 * R_MON
THETA_TAIL = 0.21


@dataclass
class Atom:
    number: int
    name: str      # 'CA' (alpha) / 'CB' (beta): second letter is the monomer type (preparator.cpp:259-265)
    chain: str
    resid: int
    x: float
    y: float
    z: float


def _fmt(a: Atom, resname: str) -> str:
    # fixed columns of scripts/pdb.py:36-52 with the template atom of make_mt.py:30-33
    return ("ATOM  " + str(a.number).rjust(5) + " " + a.name.rjust(4) + " " + resname + " " + a.chain
            + str(a.resid).rjust(4) + " " + "   " + ("%.3f" % a.x).rjust(8) + ("%.3f" % a.y).rjust(8)
            + ("%.3f" % a.z).rjust(8) + "  0.00" + "  8.12" + "      " + "   A" + "  ")


def _write(path: Path, atoms: List[Atom], resname: str) -> None:
    with open(path, "w") as f:
        for a in atoms:
            f.write(_fmt(a, resname) + "\n")
        f.write("END")


def lattice(mt_len: int, tail_len: int = 0) -> Tuple[List[Atom], List[Atom]]:
    """scripts/make_mt.py: `mt_len` monomers per protofilament, the last tail_len-1 of each curled outward."""
    xyz, ang = [], []
    for c in range(N_PF):
        chain = chr(ord("A") + c)
        for mon in range(mt_len):
            if (mt_len - mon) < tail_len:
                R = R_MT + (tail_len - (mt_len - mon)) * 2 * R_MON * math.sin(THETA_TAIL)
                theta = THETA_TAIL
            else:
                R = R_MT
                theta = 0.0
            z = 2 * R_MON * mon + Z_STEP * c
            alpha = ALPHA_STEP * c
            name = "CA" if mon % 2 == 0 else "CB"
            num = mon + mt_len * c + 1
            xyz.append(Atom(num, name, chain, mon // 2 + 1, R * math.cos(-alpha), R * math.sin(-alpha), z))
            ang.append(Atom(num, name, chain, mon // 2 + 1, 0.0, -alpha, theta))
    return xyz, ang


def lattice_with_reserve(mt_len: int, mt_extra: int) -> Tuple[List[Atom], List[Atom]]:
    """scripts/make_mt_cncntr.py: lattice + `mt_extra` reserve monomers per protofilament parked on
    chain 'X' at z = 600/604 (the pool that constant-concentration insertion draws from)."""
    if mt_len % 2:
        mt_len += 1
    xyz, ang = lattice(mt_len, 0)
    for c in range(N_PF):
        for mon in range(mt_len, mt_len + mt_extra):
            name = "CA" if mon % 2 == 0 else "CB"
            num = mt_len * (N_PF - 1) + mon + mt_extra * c + 1
            xyz.append(Atom(num, name, "X", mon // 2 + 1, 0.0, 0.0, 600.0 if mon % 2 == 0 else 604.0))
            ang.append(Atom(num, name, "X", mon // 2 + 1, 0.0, 0.0, 0.0))
    return xyz, ang


def free_dimers(n_dimers: int, radius: float, height: float, seed: int = 1, seed_rings: int = 1) -> Tuple[List[Atom], List[Atom]]:
    """SYNTHETIC: `seed_rings` fixed lattice dimers per protofilament plus free dimers placed at random
    (non-overlapping, random orientation) in a cylinder — the kind of system initial/cylinder_xyz.pdb
    holds (520 monomers: one seed dimer ring + free tubulin).  Not a copy of that file."""
    rng = np.random.default_rng(seed)
    xyz, ang = lattice(2 * seed_rings, 0)
    # keep the reference's chain/resid convention: residues continue per chain
    centers = [(a.x, a.y, a.z) for a in xyz]
    n_placed = 0
    per_chain = [0] * N_PF
    while n_placed < n_dimers:
        r = radius * math.sqrt(rng.random())
        phi = 2 * math.pi * rng.random()
        z0 = 4.0 * seed_rings * 2 + 8.0 + (height - 4.0 * seed_rings * 2 - 16.0) * rng.random()
        fi, psi, theta = (rng.random(3) * 2 - 1) * np.array([3.0, 3.0, 1.5])
        # body z axis (third column of Rz(psi) Ry(theta) Rx(fi))
        n = (math.sin(fi) * math.sin(psi) + math.cos(fi) * math.cos(psi) * math.sin(theta),
             -math.cos(psi) * math.sin(fi) + math.cos(fi) * math.sin(psi) * math.sin(theta),
             math.cos(fi) * math.cos(theta))
        c1 = (r * math.cos(phi), r * math.sin(phi), z0)
        c2 = (c1[0] + 4 * n[0], c1[1] + 4 * n[1], c1[2] + 4 * n[2])
        ok = all((c[0] - q[0]) ** 2 + (c[1] - q[1]) ** 2 + (c[2] - q[2]) ** 2 > 4.6 ** 2 for c in (c1, c2) for q in centers)
        if not ok:
            continue
        ch = n_placed % N_PF
        chain = chr(ord("A") + ch)
        per_chain[ch] += 1
        resid = seed_rings + per_chain[ch]
        base = len(xyz)
        for k, (c, name) in enumerate(((c1, "CA"), (c2, "CB"))):
            xyz.append(Atom(base + k + 1, name, chain, resid, round(c[0], 3), round(c[1], 3), round(c[2], 3)))
            ang.append(Atom(base + k + 1, name, chain, resid, round(float(fi), 3), round(float(psi), 3), round(float(theta), 3)))
            centers.append(c)
        n_placed += 1
    return xyz, ang


def write_pair(xyz: List[Atom], ang: List[Atom], xyz_path, ang_path) -> None:
    _write(Path(xyz_path), xyz, "ALA")
    _write(Path(ang_path), ang, "GLY")
