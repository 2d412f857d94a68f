"""Drop-in host: config surface, derived constants and topology against the reference's own host code
(the `params_raw` / topology records that oracle/ref_probe.cu dumped from the reference's initParameters)."""
import numpy as np
import pytest

from conftest import golden_npz
from helpers import ref_parameters, system_from_golden
from mt_b200 import HostSystem, MaddyError, workspace

FLOAT_FIELDS = ["Temp", "varR", "gammaR", "varTheta", "gammaTheta", "freeze_temp", "alpha", "dt", "C", "B_psi", "B_fi", "B_theta",
                "psi_0", "fi_0", "theta0_gtp", "theta0_gdp", "A_lat", "A_long", "D_lat", "D_long", "seam_coeff", "rep_h", "rep_r",
                "rep_eps", "rep_leftborder", "a_barr_long", "r_barr_long", "w_barr_long", "a_barr_lat", "r_barr_lat", "w_barr_lat",
                "ljpairscutoff", "ljscale", "ljsigma6"]


@pytest.mark.parametrize("name", ["mt40", "mt120_gdp_barrier", "reserve_walls", "mt40_static", "tea"])
def test_derived_parameters_bit_equal_to_reference_host(name, rundir, load_system):
    g = golden_npz(name)
    ref = ref_parameters(g["params_raw"])
    s = system_from_golden(g, rundir, load_system)
    for f in FLOAT_FIELDS:
        mine = getattr(s.par, f)
        assert np.float32(mine).tobytes() == np.float32(getattr(ref, f)).tobytes(), f"{f}: {mine!r} vs reference {getattr(ref, f)!r}"
    assert (s.par.n_tot, s.par.n_tr, s.par.rseed) == (ref.Ntot, ref.Ntr, ref.rseed)
    assert bool(s.par.barrier) == ref.barrier and bool(s.par.lj_on) == ref.lj_on and bool(s.par.is_wall) == ref.is_wall
    assert bool(s.par.is_assembly) == ref.is_assembly and s.par.ljpairsupdatefreq == ref.ljpairsupdatefreq
    assert bool(s.par.tea_on) == ref.hdi_on
    assert (s.host.steps, s.host.stride) == (ref.steps, ref.stride)
    assert bool(s.host.hydrolysis) == ref.hydrolysis and bool(s.host.is_const_conc) == ref.is_const_conc


@pytest.mark.parametrize("name", ["mt40", "mt120_gdp_barrier", "reserve_walls", "mt40_static", "tea"])
def test_topology_equal_to_reference_host(name, rundir, load_system):
    g = golden_npz(name)
    s = system_from_golden(g, rundir, load_system)
    assert np.array_equal(s.harmonic, g["harm"]) and np.array_equal(s.harmonic_count, g["harmcnt"])
    assert np.array_equal(s.mon_type, g["montype"]) and np.array_equal(s.fixed, g["fixed"]) and np.array_equal(s.extra, g["extra"])
    assert (s.par.max_longitudinal, s.par.max_lateral) == (int(g["cap_long"]), int(g["cap_lat"]))
    # host coordinates after initIntegration's angle wrap == first device frame of the reference
    c = s.coords.copy()
    for k in (3, 4, 5):
        a = c[..., k].astype(np.float64)
        c[..., k] = (a - 2 * np.pi * np.trunc(a / (2 * np.pi))).astype(np.float32)
    assert np.array_equal(c[..., :6], g["coords0"][..., :6])


def test_static_topology_lists_match_reference(rundir, load_system):
    """is_assembly = no: lists built on the host from the initial structure (preparator.cpp:357-561)."""
    g = golden_npz("mt40_static")
    s = system_from_golden(g, rundir, load_system)
    from helpers import lists_equal
    assert lists_equal(s.longitudinal_count, s.longitudinal, g["longcnt0"], g["long0"])
    assert lists_equal(s.lateral_count, s.lateral, g["latcnt0"], g["lat0"])
    # structural fixtures of SURVEY.md 8c
    assert np.bincount(s.lateral_count[0]).tolist() == [0, 6, 514]
    assert s.lateral[0, 0, :1].tolist() == [40] and s.lateral[0, 45, :2].tolist() == [-5, 85] and s.lateral[0, 500, :2].tolist() == [23, -460]
    assert np.bincount(s.longitudinal_count[0]).tolist() == [26, 494]


def test_template_defaults_and_hydrostep(rundir, load_system):
    s = load_system(rundir(runnum=3))
    assert (s.Ntot, s.Ntr) == (520, 3)
    assert s.host.hydrostep == 100 and s.host.stride == 1000 and s.host.fix == 1
    assert s.par.is_assembly == 1 and (s.par.max_longitudinal, s.par.max_lateral) == (8, 16)  # AssemblyInit capacities
    assert (s.gtp == 1).all() and (s.on_tubule_cur == 0).all()
    assert s.fixed.sum() == 26  # resid <= fix: the bottom dimer of every protofilament
    assert abs(s.par.gammaR - 6 * np.pi * 2.85e4 * 2.0) / s.par.gammaR < 1e-6


def test_config_parser_semantics(tmp_path):
    """configreader.cpp behaviours: comments, table replaced per file, argv overrides re-applied, masks, defaults, DIE."""
    d = workspace.make_baseline_rundir(tmp_path / "r", "mt40_single", runnum=2)
    conf = (d / "config.conf").read_text()
    (d / "config.conf").write_text("# a comment line\n \tindented lines are skipped too\n" + conf.replace("stride 1000", "stride\t 250   # trailing comment")
                                   .replace("dcd_xyz dcd/run_<run>.dcd", "dcd_xyz dcd/<name>_<run>.dcd") + "is_const_conc yes\nname mt\n")
    with workspace.chdir(d):
        s = HostSystem("config.conf", ["dt=100", "Temp=310"], write_files=True)
    assert s.host.stride == 250
    assert s.par.dt == 100.0  # override applied to config.conf
    assert s.par.Temp == 310.0  # ... and re-applied to conditions (a different file)
    assert s.host.is_const_conc == 0  # the key in config.conf is invisible once cond.conf replaced the table
    assert (d / "dcd" / "mt_0.dcd").exists() and (d / "dcd" / "mt_1.dcd").exists()  # <name> mask and <run> replacement
    s.close()
    # a missing mandatory key is fatal (the reference DIEs)
    (d / "morse.conf").write_text((d / "morse.conf").read_text().replace("seam_coeff 1\n", ""))
    with workspace.chdir(d):
        with pytest.raises(MaddyError) as e:
            HostSystem("config.conf")
    assert "seam_coeff" in str(e.value)
    # wrong type is fatal too
    d2 = workspace.make_baseline_rundir(tmp_path / "r2", "mt40_single", rseed="abc")
    with workspace.chdir(d2):
        with pytest.raises(MaddyError) as e:
            HostSystem("config.conf")
    assert "Should be integer" in str(e.value)


def test_const_conc_needs_reserve_particles(rundir, load_system):
    with pytest.raises(MaddyError) as e:
        load_system(rundir(runnum=1), ["is_const_conc=yes"])
    assert "chain X" in str(e.value)


def test_dcd_and_pdb_writers_roundtrip(tmp_path):
    import mt_b200
    d = workspace.make_baseline_rundir(tmp_path / "r", "mt40_single", runnum=1, steps=10, stride=5)
    with workspace.chdir(d):
        s = HostSystem("config.conf", write_files=True)
    raw = (d / "dcd" / "run_0.dcd").read_bytes()
    # header layout of dcdio.cpp:98-150: 84,'CORD',NFILE,NPRIV,NSAVC,NPRIV-NSAVC ... N at the end
    hdr = np.frombuffer(raw[:24], dtype=np.int32)
    assert hdr[0] == 84 and raw[4:8] == b"CORD" and hdr[2] == 2 and hdr[3] == 1 and hdr[4] == 5 and hdr[5] == 1 - 5
    assert np.frombuffer(raw[44:48], dtype=np.float32)[0] == 200.0
    assert len(raw) == 276 and np.frombuffer(raw[268:272], dtype=np.int32)[0] == 520
    assert raw[100:100 + 26] == b"REMARKS CREATED BY dcdio.c"
    assert mt_b200.read_dcd(d / "dcd" / "run_0.dcd").shape == (0, 520, 3)
    s.close()


def test_pdb_writer_matches_printf_formatting(tmp_path):
    """The fast ATOM-line formatter (std::to_chars) must be byte-identical to the reference's fprintf format."""
    from mt_b200 import capi
    d = workspace.make_baseline_rundir(tmp_path / "r", "mt40_single", runnum=1)
    with workspace.chdir(d):
        s = HostSystem("config.conf")
    rng = np.random.default_rng(0)
    c = s.coords
    c[0, :, :3] = rng.normal(0, 90, (520, 3)).astype(np.float32)
    c[0, :10, 0] = [0.0005, -0.0005, 1.0005, 2.5, -0.0, 999.9995, -99.9995, 0.0015, 0.0025, 1e-9]
    c[0, :, 3:6] = rng.normal(0, 3, (520, 3)).astype(np.float32)
    assert capi.hostlib.mt_system_save_pdb(s._h, str(d / "x.pdb").encode(), str(d / "a.pdb").encode()) == 0
    xyz = (d / "x.pdb").read_text().split("\n")
    ang = (d / "a.pdb").read_text().split("\n")
    assert xyz[520] == "END" and len(xyz) == 521
    for i in range(520):
        x, y, z, fi, theta, psi = (float(v) for v in c[0, i, :6])
        name = "CA" if i % 2 == 0 else "CB"
        chain = chr(ord("A") + i // 40)
        head = "ATOM  %5d %-4s%c%3s %c%4d    " % (i + 1, name, " ", "ALA", chain, (i % 40) // 2 + 1)
        assert xyz[i] == head + "%8.3f%8.3f%8.3f%6.2f%6.2f" % (x, y, z, 0.0, 8.12)
        assert ang[i] == head + "%8.3f%8.3f%8.3f%6.2f%6.2f" % (fi, psi, theta, 0.0, 8.12)
    s.close()


def test_checkpoint_keys_are_validated(rundir, load_system):
    from mt_b200 import MaddyError
    d = rundir("mt40_single", runnum=1, steps=100, stride=50)
    with pytest.raises(MaddyError, match="multiple of LJPairsUpdateFreq"):
        load_system(d, ["checkpoint=ck.bin", "steps=110"])
    with pytest.raises(MaddyError, match="checkpoint_freq"):
        load_system(d, ["checkpoint=ck.bin", "checkpoint_freq=30"])
    s = load_system(d, ["checkpoint=ck.bin", "checkpoint_freq=100", "is_restart=no"])
    assert s.host.steps == 100
    # is_restart with a checkpoint key but no checkpoint file falls back to the reference's XYZ restart (which needs its key file)
    with pytest.raises(MaddyError, match="restart/key.txt"):
        load_system(d, ["checkpoint=missing.bin", "is_restart=yes"])
