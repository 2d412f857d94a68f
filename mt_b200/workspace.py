"""Run directories for the BASELINE configs: config.conf + forcefield + conditions + PDB inputs.

The key names are the reference's parameter surface (src/parameters.h:18-243); the values are
those of template/config.conf, template/morse.conf and template/cond.conf unless a config of
BASELINE.json overrides them.  Files are written in the `name value` format both
hosts parse, so the SAME directory drives this repo's host and the reference binary.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Optional

from . import structures

CONFIG_DEFAULT: Dict[str, object] = {
    "device": 0, "mpi_device_auto": "no", "mpi_dpn": 2, "mpi_firstrun_auto": "yes",
    "rseed": 1234567, "dt": 200, "LJPairsCutoff": 15, "LJPairsUpdateFreq": 20, "stride": 1000,
    "forcefield": "morse.conf", "conditions": "cond.conf",
    "coordinates_xyz": "dcd/xyz.pdb", "coordinates_ang": "dcd/ang.pdb",
    "restart_xyz": "restart/xyz_<run>.xyz", "restart_ang": "restart/ang_<run>.xyz", "restartkey": "restart/key.txt",
    "dcd_xyz": "dcd/run_<run>.dcd", "dcd_ang": "dcd/run_<run>.dcd_ang",
    "steps": 100000, "fix": 1, "runnum": 1,
}
FORCEFIELD_DEFAULT: Dict[str, object] = {
    "C": 300.0, "D_long": 15.0, "A_long": 3.0, "D_lat": 7.0, "A_lat": 5.0,
    "LJ_on": "yes", "LJSigma": 3.8, "LJScale": 0.1,
    "B_psi": 9125, "B_fi": 9125, "B_theta": 90, "psi0": 0, "fi0": 0, "theta0_gdp": 0.2, "theta0_gtp": 0.0,
    "a_barr_long": 0, "r_barr_long": 0.4, "w_barr_long": 0.10, "a_barr_lat": 0, "r_barr_lat": 0.4, "w_barr_lat": 0.15,
    "rep_h": 160, "rep_r": 80.0, "rep_eps": 0.005, "repulsive_walls": "no", "rep_leftborder": 0.0,
    "seam_coeff": 1, "barrier": "yes",
}
CONDITIONS_DEFAULT: Dict[str, object] = {
    "Temp": 300, "viscosity": 2.85e4, "is_const_conc": "no", "conc": 30, "hydrolysis": "yes", "khydro": 1000000,
    "freeze_temp": 1.0, "alpha": 5.0,
}


def _write_conf(path: Path, values: Dict[str, object]) -> None:
    with open(path, "w") as f:
        for k, v in values.items():
            f.write(f"{k} {v}\n")


def make_rundir(root, structure=("lattice", 40, 0), config: Optional[Dict[str, object]] = None,
                forcefield: Optional[Dict[str, object]] = None, conditions: Optional[Dict[str, object]] = None) -> Path:
    """Create <root>/{config.conf,morse.conf,cond.conf,dcd/xyz.pdb,dcd/ang.pdb,restart/} and return <root>.

    structure: ("lattice", mt_len, tail_len) | ("reserve", mt_len, mt_extra) | ("free", n_dimers, radius, height, seed)
               | ("files", xyz.pdb, ang.pdb)
    """
    root = Path(root)
    (root / "dcd").mkdir(parents=True, exist_ok=True)
    (root / "restart").mkdir(exist_ok=True)
    kind = structure[0]
    if kind == "lattice":
        xyz, ang = structures.lattice(structure[1], structure[2])
    elif kind == "reserve":
        xyz, ang = structures.lattice_with_reserve(structure[1], structure[2])
    elif kind == "free":
        xyz, ang = structures.free_dimers(*structure[1:])
    elif kind == "files":  # an existing PDB pair (e.g. the reference's initial/*.pdb), copied byte for byte
        import shutil
        shutil.copyfile(structure[1], root / "dcd" / "xyz.pdb")
        shutil.copyfile(structure[2], root / "dcd" / "ang.pdb")
        xyz = None
    else:
        raise ValueError(kind)
    if xyz is not None:
        structures.write_pair(xyz, ang, root / "dcd" / "xyz.pdb", root / "dcd" / "ang.pdb")
    cfg = dict(CONFIG_DEFAULT)
    cfg.update(config or {})
    ff = dict(FORCEFIELD_DEFAULT)
    ff.update(forcefield or {})
    cond = dict(CONDITIONS_DEFAULT)
    cond.update(conditions or {})
    _write_conf(root / "config.conf", cfg)
    _write_conf(root / str(cfg["forcefield"]), ff)
    _write_conf(root / str(cfg["conditions"]), cond)
    return root


# The five configurations of BASELINE.json.  `runnum` is the headline ensemble size;
# callers override it (and steps/stride) for small parity cases.
BASELINE_CONFIGS = {
    # 1: 13-PF seed, 1 trajectory, Morse lateral bonds, 1e5 steps
    "mt40_single": dict(structure=("lattice", 40, 0), config={"runnum": 1, "steps": 100000}),
    # 2: same seed, 256 trajectories on one B200 (the config the metric is quoted on)
    "mt40_ensemble": dict(structure=("lattice", 40, 0), config={"runnum": 256, "steps": 100000}),
    # 3: long MT + free tubulin at constant concentration, LJ list, walls (make_mt_cncntr.py recipe)
    "mt120_constconc": dict(structure=("reserve", 120, 40), config={"runnum": 128, "steps": 100000},
                            forcefield={"repulsive_walls": "yes", "rep_r": 80.0, "rep_h": 160},
                            conditions={"is_const_conc": "yes", "conc": 30}),
    # 4: disassembly: curled tails + Morse barrier
    "mt120_disassembly": dict(structure=("lattice", 120, 3), config={"runnum": 256, "steps": 100000},
                              forcefield={"barrier": "yes", "a_barr_long": 3.4, "a_barr_lat": 1.9, "theta0_gdp": 0.2},
                              conditions={"hydrolysis": "yes"}),
    # 5: TEA hydrodynamics on free dimers in a cylinder (SYNTHETIC structure, see structures.free_dimers)
    "cylinder_tea": dict(structure=("free", 247, 30.0, 160.0, 1), config={"runnum": 64, "steps": 10000, "tea_on": "yes", "tea_a": 1.5,
                                                                    "tea_epsilon_freq": 100, "tea_capricious": "yes"},
                         conditions={"hydrolysis": "no"}),
    # 5b: the "large-N single system" of config 5 (SYNTHETIC: 2600 free dimers + the seed ring = 5226 beads, cylinder scaled
    #     to the same density): longer than one CTA's shared memory, runs on the wide path (maddy_wide.cuh)
    "cylinder_tea_large": dict(structure=("free", 2600, 65.7, 350.5, 1), config={"runnum": 1, "steps": 10000, "tea_on": "yes", "tea_a": 1.5,
                                                                               "tea_epsilon_freq": 100, "tea_capricious": "yes"},
                               conditions={"hydrolysis": "no"}),
    # large-N lattice (SYNTHETIC: make_mt.py recipe, 13 x 400 = 5200 monomers), wide path without TEA
    "mt400_single": dict(structure=("lattice", 400, 3), config={"runnum": 1, "steps": 100000}),
}


def make_baseline_rundir(root, name: str, **config_overrides) -> Path:
    spec = BASELINE_CONFIGS[name]
    cfg = dict(spec.get("config", {}))
    cfg.update(config_overrides)
    return make_rundir(root, spec["structure"], cfg, spec.get("forcefield"), spec.get("conditions"))


class chdir:
    """The hosts resolve forcefield/conditions/coordinates/dcd paths relative to the cwd, like the reference."""

    def __init__(self, path):
        self.path = str(path)

    def __enter__(self):
        self.old = os.getcwd()
        os.chdir(self.path)
        return self

    def __exit__(self, *a):
        os.chdir(self.old)
