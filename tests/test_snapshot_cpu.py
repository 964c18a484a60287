"""ddcMD-format writers and the control file, host side - no GPU needed.

writeRestart (src/io.c:58-113) / collection_writeBLOCK (src/collection_write.c:57-186) must write the record block the
UNMODIFIED reference writes for the same state, byte for byte (tests/golden/snapshot.json, made by
tests/golden/make_snapshot_golden.py with `ddcMD_ref readWrite`); the reader must take those files back; and, when the
oracle binary is present, the reference itself must read a snapshot written here and re-write the same bytes."""
import hashlib
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import ddcmd_b200 as dd
import nglfc_decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ddcMD_ref")
CASES = [("waterbox", None), ("waterbox", "full"), ("popc_small", None), ("popc_small", "full"), ("ras_small", "full")]
WRITER_LINES = ("create_time", "code_version")     # header lines that name the writer


def stage(golden_dir, deck, variant, tmp_path):
    if variant:
        return nglfc_decks.make_variant(golden_dir, deck, variant, tmp_path)
    dst = os.path.join(str(tmp_path), deck)
    shutil.copytree(os.path.join(golden_dir, deck), dst, symlinks=True)
    return dst


def split_atoms(path):
    raw = open(path, "rb").read()
    k = raw.index(b"}")
    return raw[:k].decode(), raw[k:]


def strip_ids(text):
    return re.sub(r"run_id=0x[0-9a-f]{8}", "run_id=X", text)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return json.load(open(os.path.join(golden_dir, "snapshot.json")))


@pytest.mark.parametrize("deck,variant", CASES)
def test_writeRestart_matches_reference_bytes(golden_dir, gold, tmp_path, deck, variant):
    g = gold[deck + ("_" + variant if variant else "")]["loop0"]
    d = stage(golden_dir, deck, variant, tmp_path)
    dk = dd.Deck(os.path.join(d, "object.data"))
    snap = dk.writeRestart()
    assert os.path.basename(snap) == g["snapshot"]
    header, body = split_atoms(os.path.join(snap, "atoms#000000"))
    assert len(body) == g["body_bytes"]
    first = body.index(b"\n\n") + 2
    assert body[first:first + 2 * g["lrec"]].decode() == g["first_records"]
    assert hashlib.sha256(body).hexdigest() == g["body_sha256"]          # every record, CRC32 and LCG64 field included
    ours = [ln for ln in strip_ids(header).splitlines() if not ln.startswith(WRITER_LINES[1])]
    ref = [ln for ln in strip_ids(g["header"]).splitlines() if not ln.startswith(WRITER_LINES[1])]
    assert [re.sub(r"create_time=[^;]*;", "", x) for x in ours] == [re.sub(r"create_time=[^;]*;", "", x) for x in ref]
    assert strip_ids(open(os.path.join(snap, "restart")).read()) == strip_ids(g["restart"])


@pytest.mark.parametrize("deck,variant", CASES)
@pytest.mark.parametrize("mode", ["binary", "brief"])
def test_binary_restart_matches_reference_bytes(golden_dir, gold, tmp_path, deck, variant, mode):
    """SIMULATE checkpointmode=BINARY (and checkpointprecision=BRIEF): FIXRECORDBINARY records of collection_writeBLOCK_binary."""
    g = gold[deck + ("_" + variant if variant else "")]["loop0_" + mode]
    d = stage(golden_dir, deck, variant, tmp_path)
    p = os.path.join(d, "object.data")
    extra = "checkpointmode=BINARY;" + (" checkpointprecision=BRIEF;" if mode == "brief" else "")
    text = open(p).read()
    open(p, "w").write(re.sub(r"checkpointrate=\d+;", "checkpointrate=10; " + extra, text, count=1))
    dk = dd.Deck(p)
    assert int(dk.s.checkpointBinary) == 1 and int(dk.s.checkpointBrief) == (mode == "brief")
    snap = dk.writeRestart(restart_link=True)
    raw = open(os.path.join(snap, "atoms#000000"), "rb").read()
    k = raw.index(b"}")
    assert len(raw) - k == g["body_bytes"] and hashlib.sha256(raw[k:]).hexdigest() == g["body_sha256"]
    strip = lambda t: [re.sub(r"create_time=[^;]*;", "", x) for x in strip_ids(t).splitlines() if not x.startswith("code_version")]   # noqa: E731
    assert strip(raw[:k].decode()) == strip(g["header"])
    # and back through the reader: ids, species, groups and LCG64 states exactly; positions and velocities to the format's precision
    d2 = dd.Deck(p)
    assert np.array_equal(d2.array("gid"), dk.array("gid")) and np.array_equal(d2.array("species"), dk.array("species"))
    assert np.array_equal(d2.array("groupOfBead"), dk.array("groupOfBead"))
    if int(dk.s.haveRandom):
        assert np.array_equal(d2.array("rngState"), dk.array("rngState")) and np.array_equal(d2.array("rngPrime"), dk.array("rngPrime"))
    h = np.array(dk.s.params.h[:])[[0, 4, 8]]
    for kx, a in zip(("rx", "ry", "rz"), h):
        dx = d2.array(kx) - dk.array(kx)
        dx -= a * np.rint(dx / a)
        assert np.abs(dx).max() <= 1e-15 * a
    vtol = 1e-7 if mode == "brief" else 1e-15
    for kv in ("vx", "vy", "vz"):
        assert np.abs(d2.array(kv) - dk.array(kv)).max() <= vtol * max(np.abs(dk.array(kv)).max(), 1e-30)


def test_corrupt_binary_record_is_rejected(golden_dir, tmp_path):
    d = stage(golden_dir, "popc_small", None, tmp_path)
    p = os.path.join(d, "object.data")
    text = open(p).read()
    open(p, "w").write(re.sub(r"checkpointrate=\d+;", "checkpointrate=10; checkpointmode=BINARY;", text, count=1))
    snap = dd.Deck(p).writeRestart(restart_link=True)
    f = os.path.join(snap, "atoms#000000")
    raw = bytearray(open(f, "rb").read())
    k = raw.index(b"\n\n", raw.index(b"}")) + 2 + 7 * 75 + 30
    raw[k] ^= 0x10
    open(f, "wb").write(raw)
    with pytest.raises(dd.DdcError, match="CRC32 mismatch in record 7"):
        dd.Deck(p)


@pytest.mark.parametrize("deck,variant", CASES)
def test_writeBXYZ_matches_reference_bytes(golden_dir, gold, tmp_path, deck, variant):
    """bxyz#000000 (collection_writeBXYZ): single-precision snapshot records with CRC32, as readWriteMaster writes them at loop 0."""
    g = gold[deck + ("_" + variant if variant else "")]["bxyz0"]
    d = stage(golden_dir, deck, variant, tmp_path)
    dk = dd.Deck(os.path.join(d, "object.data"))
    dk.writeBXYZ()
    raw = open(os.path.join(d, "snapshot.%0*d" % (int(dk.s.nLoopDigits), 0), "bxyz#000000"), "rb").read()
    k = raw.index(b"}")
    assert len(raw) - k == g["body_bytes"] and hashlib.sha256(raw[k:]).hexdigest() == g["body_sha256"]
    strip = lambda t: [re.sub(r"create_time=[^;]*;", "", x) for x in strip_ids(t).splitlines() if not x.startswith("code_version")]   # noqa: E731
    assert strip(raw[:k].decode()) == strip(g["header"])


def test_restart_round_trip_through_the_reader(golden_dir, tmp_path):
    """write -> read back through ddcb200_deckLoad (CRC32 records, hexadecimal ids, LCG64 fields)."""
    d = stage(golden_dir, "popc_small", "full", tmp_path)
    dk = dd.Deck(os.path.join(d, "object.data"))
    rng = np.random.default_rng(5)
    st = {k: dk.array(k) + 1e-3 * rng.standard_normal(dk.n) for k in ("rx", "ry", "rz", "vx", "vy", "vz")}
    states = rng.integers(0, 2 ** 63, dk.n, dtype=np.uint64)
    h = np.array(dk.s.params.h[:]) * 1.01
    snap = dk.writeRestart(loop=120, time=dk.s.dt * 120, h=h, rng=states, restart_link=True, **st)
    assert os.path.basename(snap) == "snapshot.000000000120"
    assert os.readlink(os.path.join(d, "restart")) == "./snapshot.000000000120/restart"
    d2 = dd.Deck(os.path.join(d, "object.data"))
    assert int(d2.s.loop) == 120 and abs(d2.s.time - dk.s.dt * 120) < 1e-9 * dk.s.dt * 120
    assert np.array_equal(d2.array("gid"), dk.array("gid")) and np.array_equal(d2.array("species"), dk.array("species"))
    assert np.allclose(np.array(d2.s.params.h[:]), h, rtol=1e-14)
    hh = h[[0, 4, 8]]
    for k, a in zip(("rx", "ry", "rz"), hh):
        dx = d2.array(k) - st[k]
        dx -= a * np.rint(dx / a)                       # records are written back in the box
        assert np.abs(dx).max() < 1e-11
    for k in ("vx", "vy", "vz"):
        assert np.abs(d2.array(k) - st[k]).max() <= 1e-13 * np.abs(st[k]).max()
    assert np.array_equal(d2.array("rngState"), states)
    assert np.array_equal(d2.array("rngMult"), dk.array("rngMult")) and np.array_equal(d2.array("rngPrime"), dk.array("rngPrime"))


def test_corrupt_record_is_rejected(golden_dir, tmp_path):
    d = stage(golden_dir, "popc_small", None, tmp_path)
    dk = dd.Deck(os.path.join(d, "object.data"))
    snap = dk.writeRestart(restart_link=True)
    p = os.path.join(snap, "atoms#000000")
    raw = bytearray(open(p, "rb").read())
    k = raw.index(b"\n\n") + 2 + 5 * 208 + 100
    raw[k] = ord("7") if raw[k] != ord("7") else ord("3")
    open(p, "wb").write(raw)
    with pytest.raises(dd.DdcError, match="CRC32 mismatch in record 5"):
        dd.Deck(os.path.join(d, "object.data"))


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ddcMD_ref not built")
@pytest.mark.parametrize("deck,variant,binary", [("popc_small", "full", False), ("waterbox", None, False), ("popc_small", "full", True)])
def test_reference_reads_our_snapshot(golden_dir, tmp_path, deck, variant, binary):
    """Drop-in check in the other direction: the unmodified reference starts from a restart written here (perturbed state,
    loop 40) and its own readWrite pass re-writes exactly the records it was given."""
    d = stage(golden_dir, deck, variant, tmp_path)
    if binary:
        p = os.path.join(d, "object.data")
        text = open(p).read()
        open(p, "w").write(re.sub(r"checkpointrate=\d+;", "checkpointrate=10; checkpointmode=BINARY;", text, count=1))
    dk = dd.Deck(os.path.join(d, "object.data"))
    rng = np.random.default_rng(7)
    st = {k: dk.array(k) * (1.0 + 1e-4 * rng.standard_normal(dk.n)) for k in ("rx", "ry", "rz", "vx", "vy", "vz")}
    snap = dk.writeRestart(loop=40, time=dk.s.dt * 40, restart_link=True, **st)
    _, ours = split_atoms(os.path.join(snap, "atoms#000000"))
    shutil.move(snap, snap + ".ours")
    os.unlink(os.path.join(d, "restart"))
    os.symlink("./snapshot.000000000040.ours/restart", os.path.join(d, "restart"))
    s = open(os.path.join(snap + ".ours", "restart")).read().replace("snapshot.000000000040/", "snapshot.000000000040.ours/")
    open(os.path.join(snap + ".ours", "restart"), "w").write(s)
    s = open(os.path.join(d, "object.data")).read()
    open(os.path.join(d, "object.data"), "w").write(re.sub(r"checkpointrate=\d+;", "checkpointrate=10;", s))
    r = subprocess.run([REF, "readWrite"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _, theirs = split_atoms(os.path.join(snap, "atoms#000000"))
    assert "loop=40;" in open(os.path.join(snap, "restart")).read()
    if not binary:
        assert theirs == ours
        return
    # binary fields carry all 53 bits, so the reference's internal-units round trip (x * l_in * l_out) may move a last bit:
    # same record count and layout, ids / species / LCG64 states identical, coordinates to 4 ulp
    assert len(theirs) == len(ours)
    os.unlink(os.path.join(d, "restart"))
    os.symlink("./snapshot.000000000040/restart", os.path.join(d, "restart"))
    back = dd.Deck(os.path.join(d, "object.data"))
    assert int(back.s.loop) == 40
    assert np.array_equal(back.array("gid"), dk.array("gid")) and np.array_equal(back.array("species"), dk.array("species"))
    assert np.array_equal(back.array("rngState"), dk.array("rngState"))
    h = np.array(dk.s.params.h[:])[[0, 4, 8]]
    for k, a in zip(("rx", "ry", "rz"), h):
        dx = back.array(k) - st[k]
        dx -= a * np.rint(dx / a)
        assert np.abs(dx).max() <= 1e-15 * a
    for k in ("vx", "vy", "vz"):
        assert np.abs(back.array(k) - st[k]).max() <= 1e-15 * np.abs(st[k]).max()


def test_readCMDS(tmp_path):
    p = os.path.join(str(tmp_path), "ddcMD_CMDS")
    assert dd.read_cmds(p) == 0                                  # no file
    open(p, "w").write("checkpoint\n")
    assert dd.read_cmds(p) == 1 and os.path.getsize(p) == 0      # CHECKPOINT, file truncated
    open(p, "w").write("profile\nexit\n")
    assert dd.read_cmds(p) == (4 | 2 | 1)
    open(p, "w").write("kill\n")
    assert dd.read_cmds(p) == 2
    assert dd.read_cmds(p) == 0


def test_printinfo_header_matches_reference(golden_dir, gold):
    dk = dd.Deck(os.path.join(golden_dir, "popc_small", "object.data"))
    assert dk.printinfoHeader() == gold["popc_small"]["run"]["data"].splitlines()[0]


def test_writer_argument_errors(golden_dir, tmp_path):
    """bad arguments come back as errors with text, never as a crash or a half-written file"""
    import ctypes as C
    d = stage(golden_dir, "popc_small", None, tmp_path)
    dk = dd.Deck(os.path.join(d, "object.data"))
    L = dd.lib()
    pd = C.POINTER(C.c_double)
    h = np.array(dk.s.params.h[:])
    a = [np.ascontiguousarray(dk.array(k)) for k in ("rx", "ry", "rz", "vx", "vy", "vz")]
    ptr = [x.ctypes.data_as(pd) for x in a]
    assert L.ddcb200_subsetWrite(dk._p, 0, None, 0, 0.0, h.ctypes.data_as(pd), *ptr) < 0          # the deck has no subsetWrite analysis
    assert b"no such ANALYSIS" in L.ddcb200_lastHostError()
    g = np.zeros(10)
    assert L.ddcb200_pairCorrelationWrite(dk._p, 3, None, 0, 1.0, g.ctypes.data_as(pd), 1) < 0
    hb = h.copy()
    hb[1] = 0.5                                                                               # not orthorhombic
    assert L.ddcb200_writeRestart(dk._p, None, 0, 0.0, hb.ctypes.data_as(pd), *ptr, None, 0, None, 0) < 0
    assert b"orthorhombic" in L.ddcb200_lastHostError()
    assert L.ddcb200_writeBXYZ(dk._p, None, 0, 0.0, None, *ptr) < 0
    assert not [x for x in os.listdir(d) if x.startswith("snapshot.0")]
    with pytest.raises(dd.DdcError, match="cannot create"):
        dk.writeRestart(dirname="/nonexistent_dir_xyz/snap")
