"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the headers
declare, the host C object-database front end reproduces what the reference's own init chain
produced for the same decks (tests/golden/*/ref.npz, written by oracle/_ref/ref_dump), and
compute entry points fail loudly without a device (there is no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ddcmd_b200 as dd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = ["waterbox", "popc_small", "ras_small"]


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddcb200_[A-Za-z0-9]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol():
    L = dd.lib()
    names = _declared("ddcmd_b200.h") + _declared("ddcmd_b200_host.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libddcmd_b200.so does not export %s" % n
    # and the Python table binds exactly the declared set
    assert sorted(dd.EXPORTS) == sorted(names)


def test_abi_struct_sizes_match_header():
    # compile-time layout probe: a tiny C program prints sizeof() of the public structs
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include "ddcmd_b200_host.h"\nint main(){printf("%zu %zu %zu\\n",'
                             'sizeof(ddcb200_params),sizeof(ddcb200_etype),sizeof(ddcb200_deck));return 0;}\n')
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        a, b, c = map(int, subprocess.check_output([exe]).split())
    assert (a, b, c) == (C.sizeof(dd.Params), C.sizeof(dd.EType), C.sizeof(dd.DeckStruct))


@pytest.mark.parametrize("name", DECKS)
def test_deck_matches_reference_init(golden_dir, name):
    ref = np.load(os.path.join(golden_dir, name, "ref.npz"))
    deck = dd.Deck(os.path.join(golden_dir, name, "object.data"))
    n = int(ref["nion"][0])
    assert deck.n == n
    # collection_read: labels, species, positions, velocities bit for bit (input order = reference order at np=1)
    assert np.array_equal(deck.array("gid"), ref["s0_label"])
    assert np.array_equal(deck.array("species"), ref["s0_species"])
    for k in ("rx", "ry", "rz", "vx", "vy", "vz"):
        assert np.array_equal(deck.array(k), ref["s0_" + k]), k
    # species table: names in creation order, masses, charges
    names = ref["species_names"].tobytes().split(b"\0")[0].decode().split()
    assert deck.species_names == names
    assert np.array_equal(deck.array("specMass"), ref["species_mass"])
    assert np.array_equal(deck.array("specCharge")[deck.array("species")], ref["s0_q"])
    # box, units, MARTINI scalars
    assert np.array_equal(np.array(deck.s.params.h[:]), ref["h"])
    un = ref["units"]   # 1 internal in Angstrom, kJ/mol, amu, bar, K ; ke ; kB ; fs
    assert deck.s.lengthPerAngstrom == pytest.approx(1.0 / un[0], rel=1e-15)
    assert deck.s.energyPerKJmol == pytest.approx(1.0 / un[1], rel=1e-15)
    assert deck.s.massPerAmu == pytest.approx(1.0 / un[2], rel=1e-15)
    assert deck.s.pressurePerBar == pytest.approx(1.0 / un[3], rel=1e-15)
    assert deck.s.ke == un[5] and deck.s.kB == un[6]
    mp = ref["martini_parms"]   # rmax rcoulomb epsilon_r epsilon_rf krf crf
    assert deck.s.params.rmax == mp[0] and deck.s.rcoulomb == mp[1]
    assert deck.s.epsilon_r == mp[2] and deck.s.epsilon_rf == mp[3]
    assert deck.s.params.krf == mp[4] and deck.s.params.crf == mp[5]
    assert deck.s.dt == ref["dt"][0]


def test_units_convert_known_values():
    # SURVEY.md Appendix B (src/units.c:450-486 with CODATA 2014)
    assert dd.units_convert(1.0, "Angstrom", None) == pytest.approx(1.8897261254578, rel=1e-12)
    assert dd.units_convert(1.0, "kJ/mol", None) == pytest.approx(7.61759769773e-4, rel=1e-10)
    assert dd.units_convert(1.0, "amu", None) == pytest.approx(2.13314461, rel=1e-8)
    assert dd.units_convert(1.0, "bar", None) == pytest.approx(6.79786195e-9, rel=1e-8)
    assert dd.units_convert(11.0, "Angstrom", "nm") == pytest.approx(1.1, rel=1e-14)


def test_bonded_term_tables(golden_dir):
    deck = dd.Deck(os.path.join(golden_dir, "ras_small", "object.data"))
    kind = deck.array("termKind")
    idx = deck.array("termIdx").reshape(-1, 4)
    assert set(np.unique(kind)) <= {0, 1, 2, 3, 4, 5}
    assert (kind == 0).sum() > 0 and (kind == 2).sum() > 0 and (kind >= 4).sum() > 0
    need = np.where(kind == 0, 2, np.where(kind <= 3, 3, 4))
    for a in range(4):
        used = need > a
        assert np.all(idx[used, a] >= 0) and np.all(idx[used, a] < deck.n)
    # every term lives inside one molecule (ddcRule MARTINI keeps molecules whole)
    gid = deck.array("gid")
    mol = (gid >> np.uint64(32)).astype(np.int64)
    for a in range(1, 4):
        used = need > a
        assert np.array_equal(mol[idx[used, 0]], mol[idx[used, a]])
    # exclusion keys exist for multi-species molecule types
    off = deck.array("bpairOffset")
    assert off[-1] > 0 and np.all(np.diff(off) >= 0)


def test_waterbox_has_no_terms_and_single_species_molecules(golden_dir):
    deck = dd.Deck(os.path.join(golden_dir, "waterbox", "object.data"))
    assert deck.s.nTerms == 0 and deck.n == 6173
    assert np.all(deck.array("molTypeNSpecies") == 1)
    assert deck.s.params.updateRate == 20


def test_deck_errors_are_reported(tmp_path):
    with pytest.raises(dd.DdcError):
        dd.Deck(str(tmp_path / "missing_object.data"))
    bad = tmp_path / "object.data"
    bad.write_text("simulate SIMULATE { type = MD; system=nosuch; }\n")
    with pytest.raises(dd.DdcError):
        dd.Deck(str(bad))


def test_no_device_is_a_loud_error(golden_dir):
    L = dd.lib()
    if L.ddcb200_deviceCount() > 0:
        pytest.skip("a CUDA device is present")
    deck = dd.Deck(os.path.join(golden_dir, "waterbox", "object.data"))
    with pytest.raises(dd.DdcError, match="no CUDA device|no CPU path"):
        dd.Simulate(deck)
    ctx = C.c_void_p()
    p = dd.Params()
    rc = L.ddcb200_create(C.byref(p), C.byref(ctx))
    assert rc == -2 and b"no CPU path" in L.ddcb200_lastError()


def test_product_does_not_reference_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ddcmd_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                # prose may mention the oracle; code may not name it in a string literal, import or include
                bad = re.findall(r"""["'][^"'\n]*(?:oracle|ref_dump|_ref)[^"'\n]*["']|import\s+oracle|from\s+oracle|#include\s*[<"][^>"]*oracle""", txt)
                assert not bad, (os.path.join(root, f), bad)


def test_nglfconstraint_deck_parsing(golden_dir, tmp_path):
    """INTEGRATOR NGLFCONSTRAINT keys, LANGEVIN groups, the default per-bead LCG64 streams (lcg64_default + nextPrime) and the
    constraint clusters of genConstraint, against what the reference's own init produced (tests/golden/nglfc.npz)."""
    import nglfc_decks
    g = np.load(os.path.join(golden_dir, "nglfc.npz"))
    d = nglfc_decks.make_variant(golden_dir, "waterbox", "full", tmp_path)
    deck = dd.Deck(os.path.join(d, "object.data"))
    s = deck.s
    assert int(s.integratorType) == 1 and int(s.nGroups) == 2 and [s.groupType[0], s.groupType[1]] == [1, 1]
    K, bar, ps = dd.units_convert(1.0, "K", None), dd.units_convert(1.0, "bar", None), dd.units_convert(1.0, "ps", None)
    assert abs(s.ncT - 310 * K) < 1e-15 * 310 * K and abs(s.ncP0 - bar) < 1e-12 * bar
    assert abs(s.ncBeta - 3.0e-4 / bar) < 1e-12 * 3.0e-4 / bar and abs(s.ncTauBarostat - ps) < 1e-12 * ps
    assert abs(s.groupTeq[0] - 310 * K) < 1e-15 * 310 * K and abs(s.groupTau[1] - ps) < 1e-12 * ps
    assert np.array_equal(deck.array("rngState"), g["waterbox_rng0_state"])
    assert np.array_equal(deck.array("rngMult"), g["waterbox_rng0_mp"][:, 0].astype(np.uint32))
    assert np.array_equal(deck.array("rngPrime"), g["waterbox_rng0_mp"][:, 1].astype(np.uint32))
    assert int(s.nCons) == 0 and set(deck.array("groupOfBead").tolist()) == {0}      # every waterbox record names "group"
    # ras_small: one CONSLISTPARMS on the protein-like residue -> one cluster
    d2 = nglfc_decks.make_variant(golden_dir, "ras_small", "full", tmp_path)
    deck2 = dd.Deck(os.path.join(d2, "object.data"))
    assert int(deck2.s.nCons) == 1
    ao, po = deck2.array("consAtomOffset"), deck2.array("consPairOffset")
    atoms, pa, pb, dist = deck2.array("consAtomBead"), deck2.array("consPairA"), deck2.array("consPairB"), deck2.array("consPairDist")
    assert ao[1] == 2 * po[1] and len(set(atoms.tolist())) == len(atoms)             # disjoint (i, i+4) pairs
    assert np.all(np.diff(atoms) > 0) and np.all(dist > 0)
    assert np.all(pa < ao[1]) and np.all(pb < ao[1]) and np.all(pa != pb)
    # the cluster's distances are the r0 of the CONSPARMS objects (nm -> internal length)
    txt = open(os.path.join(d2, "martini.data")).read()
    r0 = [float(x) for x in re.findall(r"CONSPARMS\{[^}]*r0=([0-9.eE+-]+) nm", txt)]
    assert np.allclose(dist, np.array(r0) * dd.units_convert(1.0, "nm", None), rtol=1e-14)


def test_nglfconstraint_deck_errors(golden_dir, tmp_path):
    import nglfc_decks
    d = nglfc_decks.make_variant(golden_dir, "popc_small", "lang", tmp_path)
    p = os.path.join(d, "object.data")
    s = open(p).read()
    open(p, "w").write(s.replace("group GROUP { type = LANGEVIN; Teq=310K;", "group GROUP { type = LANGEVIN; Teq=310K+t*0.1;"))
    with pytest.raises(dd.DdcError, match="constant temperature"):
        dd.Deck(p)
    open(p, "w").write(s.replace("random = lcg64;", ""))
    with pytest.raises(dd.DdcError, match="RANDOM"):
        dd.Deck(p)
    open(p, "w").write(s.replace("type = NGLFCONSTRAINT;", "type = NEXTGEN;"))
    with pytest.raises(dd.DdcError, match="NEXTGEN"):
        dd.Deck(p)


def test_standalone_driver_fails_loudly_without_a_device(golden_dir, tmp_path):
    """ddcmd_b200/ddcMD_b200 (the reference's `ddcMD -o object.data` for Martini decks): no device, no run, a clear message."""
    import shutil
    import subprocess
    import ddcmd_b200.build as b
    exe = b.EXE
    if not os.path.exists(exe):
        b.build(force=True)
    if dd.lib().ddcb200_deviceCount() > 0:
        pytest.skip("a CUDA device is present")
    d = os.path.join(str(tmp_path), "popc_small")
    shutil.copytree(os.path.join(golden_dir, "popc_small"), d, symlinks=True)
    r = subprocess.run([exe, "-o", "object.data"], cwd=d, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    assert not os.path.exists(os.path.join(d, "data"))
    r = subprocess.run([exe, "analysis"], cwd=d, capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "not supported" in r.stderr
