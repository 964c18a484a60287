"""Snapshot / restart / data-file goldens from the UNMODIFIED reference (oracle/_ref/ddcMD_ref).

  loop0:  `ddcMD_ref readWrite` (readWriteMaster, src/masters.c:100-124) on a golden deck: the reference reads the deck and
          writes snapshot.000000000000/{atoms#000000,restart} without stepping -> header text, sha256 of the record block,
          the first records and the restart text.  The writer must reproduce the record block byte for byte.
  run:    `ddcMD_ref` (simulateMaster) for 10 steps with printrate=5 and checkpointrate=10 -> the `data` file, the restart
          text and the parsed records of snapshot.000000000010 (popc_small only) for the simulateMaster test.

Run in the build container after make_golden.py:  python tests/golden/make_snapshot_golden.py
Writes tests/golden/snapshot.json and tests/golden/snapshot_run.npz.
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nglfc_decks  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ddcMD_ref")
CASES = [("waterbox", None), ("waterbox", "full"), ("popc_small", None), ("popc_small", "full"), ("ras_small", "full")]


def stage(deck, variant, tmp):
    if variant:
        return nglfc_decks.make_variant(HERE, deck, variant, tmp)
    dst = os.path.join(tmp, deck)
    shutil.copytree(os.path.join(HERE, deck), dst, symlinks=True)
    return dst


def split_atoms(path):
    raw = open(path, "rb").read()
    k = raw.index(b"}")
    return raw[:k].decode(), raw[k:]


def run_deck_edit(d):
    p = os.path.join(d, "object.data")
    s = open(p).read()
    s = re.sub(r"deltaloop=\d+;", "deltaloop=10;", s)
    s = re.sub(r"printrate=\d+;", "printrate=5;", s)
    s = re.sub(r"checkpointrate=\d+;", "checkpointrate=10;", s)
    open(p, "w").write(s)


SUBSET_DTYPE = np.dtype([("gid", "<u8"), ("pin", "<u4"), ("r", "<f4", 3)])


def subset_deck(text):
    """popc_small with a subsetWrite analysis: the phosphate and sodium beads above z = -30 A, every 5 loops"""
    text = text.replace("printinfo=printinfo;", "printinfo=printinfo; analysis=subset;", 1)
    text = re.sub(r"printrate=\d+;", "printrate=5;", text)
    return text + ("\nsubset ANALYSIS { type = subsetWrite; outputrate=5; format=binaryCharmm; filename=pos; "
                   "species = POPCxPO4 POPExPO4 NAxNA; zmin=-30 Ang; }\n")


def paircorr_deck(text):
    """popc_small with a pair-correlation analysis: 22 bins of 0.5 A, sampled every 5 loops, written every 10"""
    text = text.replace("printinfo=printinfo;", "printinfo=printinfo; analysis=gr;", 1)
    text = re.sub(r"printrate=\d+;", "printrate=5;", text)
    return text + "\ngr ANALYSIS { type = PAIRCORRELATION; eval_rate=5; outputrate=10; length=22; delta_r=0.5 Ang; rmin=0 Ang; filename=gr.dat; }\n"


def parse_records(body, lrec):
    recs = body[body.index(b"\n\n") + 2:] if not body.startswith(b"}") else body[body.index(b"\n\n", 1) + 2:]
    n = len(recs) // lrec
    out = np.zeros((n, 6))
    gid = np.zeros(n, np.uint64)
    for i in range(n):
        f = recs[i * lrec:(i + 1) * lrec].split()
        gid[i] = int(f[1], 16)
        out[i] = [float(x) for x in f[5:11]]
    return gid, out


if __name__ == "__main__":
    gold = {}
    runz = {}
    for deck, variant in CASES:
        key = deck + ("_" + variant if variant else "")
        tmp = tempfile.mkdtemp(prefix="snap_")
        d = stage(deck, variant, tmp)
        subprocess.check_call([REF, "readWrite"], cwd=d, stdout=open(os.path.join(d, "_rw.log"), "w"), stderr=subprocess.STDOUT)
        snap = [x for x in os.listdir(d) if x.startswith("snapshot.0")][0]
        header, body = split_atoms(os.path.join(d, snap, "atoms#000000"))
        lrec = int(re.search(r"lrec=(\d+)", header).group(1))
        first = body.index(b"\n\n") + 2
        braw = open(os.path.join(d, snap, "bxyz#000000"), "rb").read()     # readWriteMaster also writes bxyz at loop 0
        bk = braw.index(b"}")
        bxyz = {"header": braw[:bk].decode(), "body_bytes": len(braw) - bk, "body_sha256": hashlib.sha256(braw[bk:]).hexdigest()}
        gold[key] = {"bxyz0": bxyz, "loop0": {"snapshot": snap, "header": header, "lrec": lrec, "body_bytes": len(body), "body_sha256": hashlib.sha256(body).hexdigest(),
                               "first_records": body[first:first + 2 * lrec].decode(), "restart": open(os.path.join(d, snap, "restart")).read()}}
        shutil.rmtree(tmp)

        # the same with SIMULATE checkpointmode=BINARY (collection_writeBLOCK_binary), full and brief precision
        for mode in ("BINARY", "BRIEF"):
            tmp = tempfile.mkdtemp(prefix="snap_")
            d = stage(deck, variant, tmp)
            p = os.path.join(d, "object.data")
            text = open(p).read()
            extra = "checkpointmode=BINARY;" + (" checkpointprecision=BRIEF;" if mode == "BRIEF" else "")
            open(p, "w").write(re.sub(r"checkpointrate=\d+;", "checkpointrate=10; " + extra, text, count=1))
            subprocess.check_call([REF, "readWrite"], cwd=d, stdout=open(os.path.join(d, "_rw.log"), "w"), stderr=subprocess.STDOUT)
            snap = [x for x in os.listdir(d) if x.startswith("snapshot.0")][0]
            raw = open(os.path.join(d, snap, "atoms#000000"), "rb").read()
            k = raw.index(b"}")
            gold[key]["loop0_" + mode.lower()] = {"header": raw[:k].decode(), "body_bytes": len(raw) - k, "body_sha256": hashlib.sha256(raw[k:]).hexdigest()}
            shutil.rmtree(tmp)

        tmp = tempfile.mkdtemp(prefix="snap_")
        d = stage(deck, variant, tmp)
        run_deck_edit(d)
        subprocess.check_call([REF], cwd=d, stdout=open(os.path.join(d, "_run.log"), "w"), stderr=subprocess.STDOUT)
        snap = "snapshot.000000000010"
        header, body = split_atoms(os.path.join(d, snap, "atoms#000000"))
        lrec = int(re.search(r"lrec=(\d+)", header).group(1))
        gold[key]["run"] = {"data": open(os.path.join(d, "data")).read(), "restart": open(os.path.join(d, snap, "restart")).read(), "lrec": lrec,
                            "restart_link": os.readlink(os.path.join(d, "restart"))}
        if deck == "popc_small":
            gid, rv = parse_records(body, lrec)
            runz[key + "_gid"] = gid
            runz[key + "_rv"] = rv
        shutil.rmtree(tmp)
        print(key, "lrec", lrec, gold[key]["loop0"]["body_bytes"], "bytes")
    # ANALYSIS type = subsetWrite, format = binaryCharmm (src/subsetWrite.c): 10 steps of popc_small, output every 5 loops
    tmp = tempfile.mkdtemp(prefix="snap_")
    d = stage("popc_small", None, tmp)
    p = os.path.join(d, "object.data")
    text = open(p).read()
    open(p, "w").write(subset_deck(text))
    subprocess.check_call([REF], cwd=d, stdout=open(os.path.join(d, "_run.log"), "w"), stderr=subprocess.STDOUT)
    gold["popc_small_subset"] = {}
    for loop in (5, 10):
        raw = open(os.path.join(d, "snapshot.%012d" % loop, "pos#000000"), "rb").read()
        k = raw.index(b"}")
        rec = np.frombuffer(raw[raw.index(b"\n\n", k) + 2:], dtype=SUBSET_DTYPE)
        gold["popc_small_subset"]["header_%d" % loop] = raw[:k].decode()
        for f in ("gid", "pin", "r"):
            runz["subset_%d_%s" % (loop, f)] = rec[f].copy()
    shutil.rmtree(tmp)
    # ANALYSIS type = PAIRCORRELATION (src/paircorrelation.c): 10 steps of popc_small, sampled every 5 loops, written at loop 10
    tmp = tempfile.mkdtemp(prefix="snap_")
    d = stage("popc_small", None, tmp)
    p = os.path.join(d, "object.data")
    text = open(p).read()
    open(p, "w").write(paircorr_deck(text))
    subprocess.check_call([REF], cwd=d, stdout=open(os.path.join(d, "_run.log"), "w"), stderr=subprocess.STDOUT)
    lines = open(os.path.join(d, "snapshot.%012d" % 10, "gr.dat")).read().splitlines()
    gold["popc_small_gr"] = {"header": lines[:3]}
    runz["gr_table"] = np.array([[float(x) for x in ln.split()] for ln in lines[3:]])
    shutil.rmtree(tmp)
    json.dump(gold, open(os.path.join(HERE, "snapshot.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "snapshot_run.npz"), **runz)
