"""ddcmd_b200 - B200-native Martini MD step behind ddcMD's object-database API.

Python here is plumbing only (ctypes over the C-ABI in include/ddcmd_b200.h and the plain-C
host layer in include/ddcmd_b200_host.h); the product is libddcmd_b200.so.  There is no CPU
path: compute calls raise when no sm_100 device is present.

Reference-facing names are kept: ``simulate_init`` (src/simulate.c:104), ``ddcenergy``
(src/ddcenergy.c:160), ``nglf`` (src/nglf.c:67), ``eval_energyInfo`` (src/energyInfo.c:75),
``constructList`` (src/nlistGPU.h:184).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Params(C.Structure):
    _fields_ = [("h", C.c_double * 9), ("pbc", C.c_int), ("updateRate", C.c_int), ("rmax", C.c_double),
                ("deltaR", C.c_double), ("minBoxSide", C.c_double), ("keR", C.c_double), ("krf", C.c_double),
                ("crf", C.c_double), ("center", C.c_double * 3), ("nConstraints", C.c_int), ("device", C.c_int)]


class EType(C.Structure):
    _fields_ = [("eion", C.c_double), ("rk", C.c_double), ("virial", C.c_double * 6), ("tion", C.c_double * 6),
                ("sion", C.c_double * 6), ("pion", C.c_double), ("temperature", C.c_double), ("number", C.c_double),
                ("volume", C.c_double), ("eLJ", C.c_double), ("eEle", C.c_double), ("eBond", C.c_double),
                ("eAngle", C.c_double), ("eTorsion", C.c_double), ("eImproper", C.c_double), ("eRestraint", C.c_double),
                ("molVirial", C.c_double * 3), ("molPressure", C.c_double * 3), ("pMolecular", C.c_double),
                ("loop", C.c_int64), ("time", C.c_double), ("nMolecules", C.c_int64), ("nPairsListed", C.c_int64)]


_P = C.POINTER


class DeckStruct(C.Structure):
    _fields_ = [("dt", C.c_double), ("time", C.c_double), ("loop", C.c_int64), ("maxloop", C.c_int64),
                ("printrate", C.c_int), ("deltaloop", C.c_int), ("snapshotrate", C.c_int), ("checkpointrate", C.c_int),
                ("params", Params), ("ddc_lx", C.c_int), ("ddc_ly", C.c_int), ("ddc_lz", C.c_int),
                ("rcoulomb", C.c_double), ("epsilon_r", C.c_double), ("epsilon_rf", C.c_double), ("rmax4all", C.c_double),
                ("excludePotentialTerm", C.c_int), ("potentialShift", C.c_int),
                ("kB", C.c_double), ("ke", C.c_double), ("lengthPerAngstrom", C.c_double), ("energyPerKJmol", C.c_double),
                ("massPerAmu", C.c_double), ("pressurePerBar", C.c_double), ("timePerFs", C.c_double),
                ("printMolecularPressure", C.c_int),
                ("nspecies", C.c_int), ("speciesName", _P(C.c_char_p)), ("specLJ", _P(C.c_int)),
                ("specCharge", _P(C.c_double)), ("specMass", _P(C.c_double)), ("specMolType", _P(C.c_int)),
                ("specResidue", _P(C.c_int)), ("specAtom", _P(C.c_int)),
                ("ntypes", C.c_int), ("ljEps", _P(C.c_double)), ("ljSigma", _P(C.c_double)), ("ljShift", _P(C.c_double)),
                ("nMolTypes", C.c_int), ("molTypeNSpecies", _P(C.c_int)), ("molTypeResidue", _P(C.c_int)),
                ("molTypeOwnerOffset", _P(C.c_int)), ("bpairOffset", _P(C.c_int)), ("bpairI", _P(C.c_int)), ("bpairJ", _P(C.c_int)),
                ("n", C.c_int64), ("gid", _P(C.c_uint64)), ("species", _P(C.c_int)),
                ("rx", _P(C.c_double)), ("ry", _P(C.c_double)), ("rz", _P(C.c_double)),
                ("vx", _P(C.c_double)), ("vy", _P(C.c_double)), ("vz", _P(C.c_double)),
                ("nTerms", C.c_int64), ("termKind", _P(C.c_int)), ("termIdx", _P(C.c_int)), ("termParm", _P(C.c_double)),
                ("nRestraints", C.c_int64), ("restrBead", _P(C.c_int)), ("restrFrac0", _P(C.c_double)),
                ("restrKb", _P(C.c_double)), ("restrFc", _P(C.c_double)), ("restrOrigin", C.c_int),
                ("nMol", C.c_int64), ("nMolTotal", C.c_int64), ("molOffset", _P(C.c_int64)), ("molBeads", _P(C.c_int)),
                ("integratorType", C.c_int), ("ncT", C.c_double), ("ncP0", C.c_double), ("ncBeta", C.c_double),
                ("ncTauBarostat", C.c_double), ("ncIsotropic", C.c_int),
                ("nGroups", C.c_int), ("groupName", _P(C.c_char_p)), ("groupType", _P(C.c_int)), ("groupTeq", _P(C.c_double)),
                ("groupTau", _P(C.c_double)), ("groupVcm", _P(C.c_double)), ("groupOfBead", _P(C.c_ubyte)),
                ("haveRandom", C.c_int), ("randomSeed", C.c_uint64), ("rngState", _P(C.c_uint64)), ("rngMult", _P(C.c_uint32)),
                ("rngPrime", _P(C.c_uint32)),
                ("nCons", C.c_int64), ("consAtomOffset", _P(C.c_int64)), ("consPairOffset", _P(C.c_int64)),
                ("consAtomBead", _P(C.c_int)), ("consPairA", _P(C.c_int)), ("consPairB", _P(C.c_int)), ("consPairDist", _P(C.c_double)),
                ("runDir", C.c_char_p), ("simulateName", C.c_char_p), ("boxName", C.c_char_p), ("collectionName", C.c_char_p),
                ("atomsdir", C.c_char_p), ("nLoopDigits", C.c_int), ("gidFormatHex", C.c_int), ("runId", C.c_uint),
                ("speciesType", _P(C.c_char_p)), ("printUnit", C.c_char_p * 6), ("printConvert", C.c_double * 6),
                ("reducedCorner", C.c_double * 3), ("checkpointBinary", C.c_int), ("checkpointBrief", C.c_int),
                ("nSubsets", C.c_int), ("subsets", C.c_void_p), ("nPairCorr", C.c_int), ("pairCorr", C.c_void_p)]


class DdcError(RuntimeError):
    pass


def _preload_nccl():
    """libddcmd_b200.so needs libnccl.so.2.  When PyTorch is installed its bundled NCCL (newer than the system
    one) must be the copy the process loads first, or a later `import torch` fails on missing symbols."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
    except Exception:
        pass   # the system libnccl.so.2 is resolved by the loader


def lib():
    """Load libddcmd_b200.so (building it in-tree with nvcc when sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "libddcmd_b200.so")
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # a prebuilt .so that travelled with the snapshot is still usable
            if not os.path.exists(path):
                raise DdcError("libddcmd_b200.so is missing and could not be built: %s" % e)
    _preload_nccl()
    _lib = _declare(C.CDLL(path))
    return _lib


def _declare(L):
    """ctypes signatures of every entry point of include/ddcmd_b200.h and include/ddcmd_b200_host.h."""
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    pd, pi = _P(C.c_double), _P(C.c_int)
    sig = {
        "ddcb200_lastError": (C.c_char_p, []),
        "ddcb200_deviceCount": (i32, []),
        "ddcb200_create": (i32, [_P(Params), _P(vp)]),
        "ddcb200_destroy": (None, [vp]),
        "ddcb200_sync": (i32, [vp]),
        "ddcb200_martiniNonBondParms": (i32, [vp, i32, pd, pd, pd]),
        "ddcb200_setSpecies": (i32, [vp, i32, pi, pd, pd]),
        "ddcb200_setBeads": (i32, [vp, i64, _P(C.c_uint64), pi]),
        "ddcb200_setExclusions": (i32, [vp, i32, pi, pi, pi, pi, pi]),
        "ddcb200_martiniBondParms": (i32, [vp, i64, pi, pi, pd]),
        "ddcb200_setRestraints": (i32, [vp, i64, pi, pd, pd, pd, i32]),
        "ddcb200_setMolecules": (i32, [vp, i64, _P(C.c_int64), pi, i64]),
        "ddcb200_sendState": (i32, [vp, i64, pi, pd, pd, pd, pd, pd, pd, i64, dbl]),
        "ddcb200_updateState": (i32, [vp, i64, pi, pd, pd, pd, pd, pd, pd, i64, dbl]),
        "ddcb200_numLocal": (i64, [vp]),
        "ddcb200_getLocalBeads": (i32, [vp, pi]),
        "ddcb200_getState": (i32, [vp] + [pd] * 9),
        "ddcb200_constructList": (i32, [vp]),
        "ddcb200_ddcenergy": (i32, [vp, i32]),
        "ddcb200_nglf": (i32, [vp, i32, dbl]),
        "ddcb200_energyInfo": (i32, [vp, dbl, _P(EType)]),
        "ddcb200_setGroups": (i32, [vp, i32, pi, pd, pd, pd, i64, _P(C.c_ubyte)]),
        "ddcb200_setRandom": (i32, [vp, i64, _P(C.c_uint64), _P(C.c_uint32), _P(C.c_uint32)]),
        "ddcb200_getRandom": (i32, [vp, i64, _P(C.c_uint64)]),
        "ddcb200_setConstraints": (i32, [vp, i64, _P(C.c_int64), pi, _P(C.c_int64), pi, pi, pd]),
        "ddcb200_nglfconstraintParms": (i32, [vp, dbl, dbl, dbl, dbl]),
        "ddcb200_nglfconstraint": (i32, [vp, i32, dbl]),
        "ddcb200_getBox": (i32, [vp, pd]),
        "ddcb200_setBox": (i32, [vp, pd]),
        "ddcb200_constraintFailures": (i64, [vp]),
        "ddcb200_getCells": (i32, [vp, pi, pi, pd]),
        "ddcb200_getPairs": (i64, [vp, i64, pi, pi, pi]),
        "ddcb200_pairSetHash": (i32, [vp, _P(C.c_uint64)]),
        "ddcb200_profile": (i32, [vp, i32]),
        "ddcb200_profileRead": (i32, [vp, pd, _P(C.c_int64), i32]),
        "ddcb200_timerRecord": (i32, [vp, i32]),
        "ddcb200_timerElapsed": (i32, [vp, i32, i32, pd]),
        "ddcb200_kernelLaunches": (i64, [vp]),
        "ddcb200_lastListBuild": (i64, [vp]),
        "ddcb200_listBuildInfo": (i32, [vp, pi, pd]),
        "ddcb200_pruneInfo": (i32, [vp, _P(C.c_int64)]),
        "ddcb200_kineticByClass": (i32, [vp, i32, i32, pd]),
        "ddcb200_pairCorrelation": (i32, [vp, i32, dbl, dbl, i32, dbl, _P(C.c_uint64), _P(C.c_uint64)]),
        "ddcb200_pairCorrelationWrite": (i32, [_P(DeckStruct), i32, C.c_char_p, i64, dbl, pd, i32]),
        "ddcb200_ncclUniqueId": (i32, [C.c_char_p]),
        "ddcb200_ddcInit": (i32, [vp, i32, i32, i32, i32, i32, C.c_char_p]),
        "ddcb200_ddcPlan": (i32, [pd, i32, i32, i32, dbl, i64, pd, pd, pd, pi, i32, pi, _P(C.c_uint32)]),
        "ddcb200_deckLoad": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, _P(_P(DeckStruct))]),
        "ddcb200_deckFree": (None, [_P(DeckStruct)]),
        "ddcb200_lastHostError": (C.c_char_p, []),
        "ddcb200_simulateBind": (i32, [_P(DeckStruct), i32, _P(vp)]),
        "ddcb200_simulateBindRank": (i32, [_P(DeckStruct), i32, i32, i32, i32, i32, i32, C.c_char_p, _P(vp)]),
        "ddcb200_printinfoLine": (i32, [_P(DeckStruct), _P(EType), C.c_char_p, C.c_size_t]),
        "ddcb200_unitsConvert": (dbl, [dbl, C.c_char_p, C.c_char_p]),
        "ddcb200_printinfoHeader": (i32, [_P(DeckStruct), C.c_char_p, C.c_size_t]),
        "ddcb200_writeRestart": (i32, [_P(DeckStruct), C.c_char_p, i64, dbl, pd, pd, pd, pd, pd, pd, pd, _P(C.c_uint64), i32, C.c_char_p, C.c_size_t]),
        "ddcb200_readCMDS": (i32, [C.c_char_p]),
        "ddcb200_writeBXYZ": (i32, [_P(DeckStruct), C.c_char_p, i64, dbl, pd, pd, pd, pd, pd, pd, pd]),
        "ddcb200_subsetWrite": (i64, [_P(DeckStruct), i32, C.c_char_p, i64, dbl, pd, pd, pd, pd, pd, pd, pd]),
        "ddcb200_simulateMaster": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L


EXPORTS = ["ddcb200_lastError", "ddcb200_deviceCount", "ddcb200_create", "ddcb200_destroy", "ddcb200_sync",
           "ddcb200_martiniNonBondParms", "ddcb200_setSpecies", "ddcb200_setBeads", "ddcb200_setExclusions",
           "ddcb200_martiniBondParms", "ddcb200_setRestraints", "ddcb200_setMolecules", "ddcb200_sendState",
           "ddcb200_updateState", "ddcb200_numLocal", "ddcb200_getLocalBeads", "ddcb200_getState", "ddcb200_constructList", "ddcb200_ddcenergy",
           "ddcb200_nglf", "ddcb200_energyInfo", "ddcb200_setGroups", "ddcb200_setRandom", "ddcb200_getRandom", "ddcb200_setConstraints",
           "ddcb200_nglfconstraintParms", "ddcb200_nglfconstraint", "ddcb200_getBox", "ddcb200_setBox", "ddcb200_constraintFailures", "ddcb200_getCells", "ddcb200_getPairs", "ddcb200_profile",
           "ddcb200_profileRead", "ddcb200_timerRecord", "ddcb200_timerElapsed", "ddcb200_kernelLaunches", "ddcb200_lastListBuild", "ddcb200_ncclUniqueId", "ddcb200_ddcInit", "ddcb200_ddcPlan", "ddcb200_deckLoad", "ddcb200_deckFree",
           "ddcb200_lastHostError", "ddcb200_simulateBind", "ddcb200_simulateBindRank", "ddcb200_printinfoLine", "ddcb200_unitsConvert",
           "ddcb200_printinfoHeader", "ddcb200_writeRestart", "ddcb200_readCMDS", "ddcb200_simulateMaster", "ddcb200_listBuildInfo", "ddcb200_pruneInfo", "ddcb200_subsetWrite", "ddcb200_writeBXYZ", "ddcb200_pairCorrelation",
           "ddcb200_pairCorrelationWrite", "ddcb200_kineticByClass", "ddcb200_pairSetHash"]


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


def units_convert(value, frm, to):
    """units_convert (src/units.c:515-551); None = internal units."""
    return lib().ddcb200_unitsConvert(value, frm.encode() if frm else None, to.encode() if to else None)


class Deck:
    """Parsed object database of one Martini deck (object.data + restart + parm files + atoms#)."""

    def __init__(self, object_file, restart_file=None, simulate_name=None):
        L = lib()
        p = _P(DeckStruct)()
        rc = L.ddcb200_deckLoad(os.fsencode(object_file), os.fsencode(restart_file) if restart_file else None,
                                simulate_name.encode() if simulate_name else None, C.byref(p))
        if rc != 0:
            raise DdcError(L.ddcb200_lastHostError().decode())
        self._p = p
        self.s = p.contents

    def __del__(self):
        if getattr(self, "_p", None) is not None and _lib is not None:
            _lib.ddcb200_deckFree(self._p)
            self._p = None

    # numpy views (valid while the Deck lives)
    @property
    def n(self):
        return int(self.s.n)

    def array(self, name):
        s = self.s
        n, ns, nt, nm = int(s.n), int(s.nspecies), int(s.ntypes), int(s.nMolTypes)
        table = {
            "gid": (s.gid, n, np.uint64), "species": (s.species, n, np.int32),
            "rx": (s.rx, n, np.float64), "ry": (s.ry, n, np.float64), "rz": (s.rz, n, np.float64),
            "vx": (s.vx, n, np.float64), "vy": (s.vy, n, np.float64), "vz": (s.vz, n, np.float64),
            "specLJ": (s.specLJ, ns, np.int32), "specCharge": (s.specCharge, ns, np.float64),
            "specMass": (s.specMass, ns, np.float64), "specMolType": (s.specMolType, ns, np.int32),
            "specResidue": (s.specResidue, ns, np.int32), "specAtom": (s.specAtom, ns, np.int32),
            "ljEps": (s.ljEps, nt * nt, np.float64), "ljSigma": (s.ljSigma, nt * nt, np.float64),
            "ljShift": (s.ljShift, nt * nt, np.float64),
            "molTypeNSpecies": (s.molTypeNSpecies, nm, np.int32), "bpairOffset": (s.bpairOffset, nm + 1, np.int32),
            "termKind": (s.termKind, int(s.nTerms), np.int32), "termIdx": (s.termIdx, 4 * int(s.nTerms), np.int32),
            "termParm": (s.termParm, 3 * int(s.nTerms), np.float64),
            "molOffset": (s.molOffset, int(s.nMol) + 1, np.int64),
            "restrBead": (s.restrBead, int(s.nRestraints), np.int32),
            "groupOfBead": (s.groupOfBead, n if s.groupOfBead else 0, np.uint8),
            "rngState": (s.rngState, n if s.rngState else 0, np.uint64), "rngMult": (s.rngMult, n if s.rngState else 0, np.uint32),
            "rngPrime": (s.rngPrime, n if s.rngState else 0, np.uint32),
            "consAtomOffset": (s.consAtomOffset, int(s.nCons) + 1, np.int64), "consPairOffset": (s.consPairOffset, int(s.nCons) + 1, np.int64),
        }
        if name in ("consAtomBead",):
            return _arr(s.consAtomBead, int(self.array("consAtomOffset")[-1]), np.int32)
        if name in ("consPairA", "consPairB", "consPairDist"):
            return _arr(getattr(s, name), int(self.array("consPairOffset")[-1]), np.float64 if name == "consPairDist" else np.int32)
        if name == "bpairI" or name == "bpairJ":
            nb = int(self.array("bpairOffset")[-1])
            return _arr(getattr(s, name), nb, np.int32)
        if name == "molBeads":
            return _arr(s.molBeads, int(self.array("molOffset")[-1]), np.int32)
        ptr, cnt, dt = table[name]
        return _arr(ptr, cnt, dt)

    def writeRestart(self, rx=None, ry=None, rz=None, vx=None, vy=None, vz=None, loop=None, time=None, h=None, rng=None, dirname=None,
                     restart_link=False):
        """writeRestart (src/io.c:58-113) of a state in this deck's bead order (default: the state the deck was read with).
        Returns the snapshot directory."""
        L = lib()
        a = [np.ascontiguousarray(x if x is not None else self.array(k), np.float64)
             for k, x in zip(("rx", "ry", "rz", "vx", "vy", "vz"), (rx, ry, rz, vx, vy, vz))]
        hh = np.ascontiguousarray(h if h is not None else np.array(self.s.params.h[:]), np.float64)
        r = np.ascontiguousarray(rng, np.uint64) if rng is not None else None
        pd = _P(C.c_double)
        out = C.create_string_buffer(1024)
        rc = L.ddcb200_writeRestart(self._p, os.fsencode(dirname) if dirname else None, int(self.s.loop if loop is None else loop),
                                    float(self.s.time if time is None else time), hh.ctypes.data_as(pd), *[x.ctypes.data_as(pd) for x in a],
                                    r.ctypes.data_as(_P(C.c_uint64)) if r is not None else None, int(restart_link), out, 1024)
        if rc != 0:
            raise DdcError(L.ddcb200_lastHostError().decode())
        return os.fsdecode(out.value)

    def writeBXYZ(self, dirname=None):
        """writeBXYZ (src/io.c:144-155) of the state the deck was read with; returns nothing (the file is <snapshotdir>/bxyz#000000)."""
        L = lib()
        a = [np.ascontiguousarray(self.array(k), np.float64) for k in ("rx", "ry", "rz", "vx", "vy", "vz")]
        hh = np.ascontiguousarray(np.array(self.s.params.h[:]), np.float64)
        pd = _P(C.c_double)
        rc = L.ddcb200_writeBXYZ(self._p, os.fsencode(dirname) if dirname else None, int(self.s.loop), float(self.s.time), hh.ctypes.data_as(pd),
                                 *[x.ctypes.data_as(pd) for x in a])
        if rc != 0:
            raise DdcError(L.ddcb200_lastHostError().decode())

    def printinfoHeader(self):
        buf = C.create_string_buffer(1024)
        lib().ddcb200_printinfoHeader(self._p, buf, 1024)
        return buf.value.decode()

    @property
    def species_names(self):
        return [self.s.speciesName[i].decode() for i in range(int(self.s.nspecies))]


class Simulate:
    """simulate_init + the simulateMaster loop pieces, on one GPU.

    ``nglf(nsteps)`` is the reference's ``eval_integrator`` called nsteps times; ``energyInfo()``
    is ``kinetic_terms`` + ``eval_energyInfo`` (+ molecular pressure).
    """

    def __init__(self, deck, device=0, rank=0, nranks=1, lattice=None, nccl_id=None):
        L = lib()
        if L.ddcb200_deviceCount() <= 0:
            raise DdcError("no CUDA device: ddcmd_b200 has no CPU path")
        self.deck = deck
        self.ctx = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        if nranks == 1:
            rc = L.ddcb200_simulateBind(deck._p, device, C.byref(self.ctx))
        else:
            lx, ly, lz = lattice or default_lattice(nranks, [deck.s.params.h[0], deck.s.params.h[4], deck.s.params.h[8]])
            rc = L.ddcb200_simulateBindRank(deck._p, device, rank, nranks, lx, ly, lz, bytes(nccl_id), C.byref(self.ctx))
        if rc != 0:
            raise DdcError(L.ddcb200_lastHostError().decode())
        self.dt = deck.s.dt

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value and _lib is not None:
            _lib.ddcb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise DdcError(lib().ddcb200_lastError().decode())

    def constructList(self):
        self._ck(lib().ddcb200_constructList(self.ctx))

    def ddcenergy(self, e_eval_flag=1):
        self._ck(lib().ddcb200_ddcenergy(self.ctx, int(e_eval_flag)))

    def nglf(self, nsteps=1, dt=None):
        self._ck(lib().ddcb200_nglf(self.ctx, int(nsteps), float(self.dt if dt is None else dt)))

    def nglfconstraint(self, nsteps=1, dt=None):
        """nglfconstraint (src/nglfconstraint.c:510-574) called nsteps times."""
        self._ck(lib().ddcb200_nglfconstraint(self.ctx, int(nsteps), float(self.dt if dt is None else dt)))

    def eval_integrator(self, nsteps=1, dt=None):
        """simulate->integrator->eval_integrator (src/masters.c:445): the INTEGRATOR object of the deck decides."""
        if int(self.deck.s.integratorType) == 1:
            self.nglfconstraint(nsteps, dt)
        else:
            self.nglf(nsteps, dt)

    def getBox(self):
        h = np.empty(9, np.float64)
        self._ck(lib().ddcb200_getBox(self.ctx, h.ctypes.data_as(_P(C.c_double))))
        return h

    def getRandom(self):
        """current per-bead LCG64 states, input order.  On several ranks (collective call over the caller's torch.distributed
        group) every rank's entries for its own local beads are merged: a bead's state is current only where the bead lives."""
        st = np.empty(self.deck.n, np.uint64)
        self._ck(lib().ddcb200_getRandom(self.ctx, st.size, st.ctypes.data_as(_P(C.c_uint64))))
        if self.nranks > 1:
            import torch.distributed as dist
            beads = self.getLocalBeads()
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, (beads, st[beads]))
            for b, v in parts:
                st[b] = v
        return st

    def constraintFailures(self):
        return int(lib().ddcb200_constraintFailures(self.ctx))

    def sync(self):
        self._ck(lib().ddcb200_sync(self.ctx))

    def energyInfo(self):
        e = EType()
        self._ck(lib().ddcb200_energyInfo(self.ctx, float(self.deck.s.kB), C.byref(e)))
        return e

    def sendState(self, rx, ry, rz, vx, vy, vz, loop=0, time=0.0, bead=None):
        """bead = input-order index of each row (None: 0..n-1); on several ranks any partition of the beads works."""
        a = [np.ascontiguousarray(x, np.float64) for x in (rx, ry, rz, vx, vy, vz)]
        pd = _P(C.c_double)
        b = np.ascontiguousarray(bead, np.int32) if bead is not None else None
        self._ck(lib().ddcb200_sendState(self.ctx, a[0].size, b.ctypes.data_as(_P(C.c_int)) if b is not None else None,
                                         *[x.ctypes.data_as(pd) for x in a], int(loop), float(time)))

    def updateState(self, rx, ry, rz, vx, vy, vz, loop=0, time=0.0, bead=None):
        """per-step upload of a host-side integrator (keeps cells and the neighbor list, unlike sendState)"""
        a = [np.ascontiguousarray(x, np.float64) for x in (rx, ry, rz, vx, vy, vz)]
        pd = _P(C.c_double)
        b = np.ascontiguousarray(bead, np.int32) if bead is not None else None
        self._ck(lib().ddcb200_updateState(self.ctx, a[0].size, b.ctypes.data_as(_P(C.c_int)) if b is not None else None,
                                           *[x.ctypes.data_as(pd) for x in a], int(loop), float(time)))

    def setBox(self, h):
        hh = np.ascontiguousarray(h, np.float64)
        self._ck(lib().ddcb200_setBox(self.ctx, hh.ctypes.data_as(_P(C.c_double))))

    def getForces(self, out=None):
        """fx fy fz of the local beads as a (3, numLocal) array (`out`: caller-owned, e.g. pinned)"""
        n = int(lib().ddcb200_numLocal(self.ctx))
        if out is None:
            out = np.empty((3, n), np.float64)
        elif out.shape != (3, n) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise DdcError("getForces: out must be a C-contiguous float64 array of shape (3, %d)" % n)
        pd = _P(C.c_double)
        self._ck(lib().ddcb200_getState(self.ctx, None, None, None, None, None, None, *[out[k].ctypes.data_as(pd) for k in range(3)]))
        return out

    def numLocal(self):
        return int(lib().ddcb200_numLocal(self.ctx))

    def getLocalBeads(self):
        n = int(lib().ddcb200_numLocal(self.ctx))
        bead = np.empty(n, np.int32)
        self._ck(lib().ddcb200_getLocalBeads(self.ctx, bead.ctypes.data_as(_P(C.c_int))))
        return bead

    def getState(self, out=None):
        """rx ry rz vx vy vz fx fy fz of the local beads; `out` = a caller-owned C-contiguous float64 array of shape (9, numLocal),
        e.g. a view of pinned memory, to copy into instead of a fresh (pageable) array."""
        n = int(lib().ddcb200_numLocal(self.ctx))
        if out is None:
            out = np.empty((9, n), np.float64)
        elif out.shape != (9, n) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise DdcError("getState: out must be a C-contiguous float64 array of shape (9, %d)" % n)
        pd = _P(C.c_double)
        self._ck(lib().ddcb200_getState(self.ctx, *[out[k].ctypes.data_as(pd) for k in range(9)]))
        return dict(zip(("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"), out))

    def getCells(self):
        n = int(lib().ddcb200_numLocal(self.ctx))
        cell = np.empty(n, np.int32)
        dims = (C.c_int * 3)()
        geom = (C.c_double * 9)()
        self._ck(lib().ddcb200_getCells(self.ctx, cell.ctypes.data_as(_P(C.c_int)), dims, geom))
        return cell, np.array(dims[:]), np.array(geom[:])

    def getPairs(self):
        L = lib()
        n = int(L.ddcb200_getPairs(self.ctx, 0, None, None, None))
        if n < 0:
            self._ck(n)
        bi, bj, pr = (np.empty(n, np.int32) for _ in range(3))
        pi = _P(C.c_int)
        L.ddcb200_getPairs(self.ctx, n, bi.ctypes.data_as(pi), bj.ctypes.data_as(pi), pr.ctypes.data_as(pi))
        return bi, bj, pr

    def pairSetHash(self):
        """(count, sum, xor) of the interacting list and of the pruned list: order-independent 64-bit hashes of the (gid, gid) pairs
        this rank owns, comparable with oracle/ref_dump's "pairhash" record."""
        h = np.zeros(6, np.uint64)
        self._ck(lib().ddcb200_pairSetHash(self.ctx, h.ctypes.data_as(_P(C.c_uint64))))
        return h

    def profile(self, enable=True):
        self._ck(lib().ddcb200_profile(self.ctx, int(enable)))

    def profileRead(self, reset=True):
        ms = (C.c_double * 8)()
        ln = (C.c_int64 * 8)()
        self._ck(lib().ddcb200_profileRead(self.ctx, ms, ln, int(reset)))
        names = ("integrate", "pair", "bonded", "list", "reduce", "halo", "pair_prune")
        return {k: (ms[i], int(ln[i])) for i, k in enumerate(names)}

    def timerRecord(self, which):
        self._ck(lib().ddcb200_timerRecord(self.ctx, int(which)))

    def timerElapsed(self, a, b):
        ms = C.c_double()
        self._ck(lib().ddcb200_timerElapsed(self.ctx, int(a), int(b), C.byref(ms)))
        return ms.value

    def kernelLaunches(self):
        return int(lib().ddcb200_kernelLaunches(self.ctx))

    def lastListBuild(self):
        """sys->neighbor->lastUpdate: loop of the last list build."""
        return int(lib().ddcb200_lastListBuild(self.ctx))

    def writeRestart(self, dirname=None, restart_link=True):
        """checkpointSimulate (src/masters.c:51-56): CreateSnapshotdir + writeRestart of the current device state."""
        e = self.energyInfo()
        st = self.getState()
        if self.nranks == 1:
            rng = self.getRandom() if int(self.deck.s.haveRandom) else None
            return self.deck.writeRestart(st["rx"], st["ry"], st["rz"], st["vx"], st["vy"], st["vz"], loop=int(e.loop), time=float(e.time),
                                          h=self.getBox(), rng=rng, dirname=dirname, restart_link=restart_link)
        # several ranks (collective call): the reference funnels every task's records to the writer tasks of pio
        # (src/pio.c); here the ranks' beads are gathered over the caller's torch.distributed group and rank 0 writes the
        # single atoms#000000 in the deck's bead order.  Returns the snapshot directory on rank 0, None elsewhere.
        import torch.distributed as dist
        keys = ("rx", "ry", "rz", "vx", "vy", "vz")
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, (self.getLocalBeads(), {k: st[k] for k in keys}))
        if dist.get_rank() != 0:
            return None
        n = self.deck.n
        full = {k: np.full(n, np.nan) for k in keys}
        for beads, d in parts:
            for k in keys:
                full[k][beads] = d[k]
        if any(np.isnan(full[k]).any() for k in keys):
            raise DdcError("writeRestart: some beads are local on no rank")
        return self.deck.writeRestart(*[full[k] for k in keys], loop=int(e.loop), time=float(e.time), h=self.getBox(), dirname=dirname,
                                      restart_link=restart_link)

    def kineticByClass(self, by_species=False):
        """per-GROUP or per-SPECIES kinetic terms of kinetic_terms (src/energy.c:116-143): array [class, 12] =
        rk, mass, number, sum m v_a v_b (xx yy zz xy xz yz), sum K v (x y z)"""
        n = int(self.deck.s.nspecies) if by_species else max(1, int(self.deck.s.nGroups))
        out = np.zeros((n, 12), np.float64)
        self._ck(lib().ddcb200_kineticByClass(self.ctx, int(bool(by_species)), n, out.ctypes.data_as(_P(C.c_double))))
        return out

    def pairCorrelation(self, nbins, rmin, delta, rmax, log_scale=False):
        """paircorrelation_eval's counts (src/paircorrelation.c:158-420): (counts[np, nbins] uint64, atoms per species uint64), internal units."""
        ns = int(self.deck.s.nspecies)
        npair = ns * (ns + 1) // 2
        counts = np.zeros(npair * int(nbins), np.uint64)
        natoms = np.zeros(ns, np.uint64)
        self._ck(lib().ddcb200_pairCorrelation(self.ctx, int(nbins), float(rmin), float(delta), int(bool(log_scale)), float(rmax),
                                               counts.ctypes.data_as(_P(C.c_uint64)), natoms.ctypes.data_as(_P(C.c_uint64))))
        return counts.reshape(npair, int(nbins)), natoms

    def listBuildInfo(self):
        """(1, [device ms of the last list build (candidate pass + exact pass), 0.0])"""
        v = C.c_int()
        ms = (C.c_double * 2)()
        self._ck(lib().ddcb200_listBuildInfo(self.ctx, C.byref(v), ms))
        return int(v.value), [ms[0], ms[1]]

    def pruneInfo(self):
        """Pruned rows of the pair walk: dict(every, since, walk_next, beads_pruned, pruned_entries, full_entries) - ddcb200_pruneInfo"""
        a = (C.c_int64 * 6)()
        self._ck(lib().ddcb200_pruneInfo(self.ctx, a))
        return dict(zip(("every", "since", "walk_next", "beads_pruned", "pruned_entries", "full_entries"), [int(x) for x in a]))

    def printinfo(self, e=None):
        e = e or self.energyInfo()
        buf = C.create_string_buffer(512)
        lib().ddcb200_printinfoLine(self.deck._p, C.byref(e), buf, 512)
        return buf.value.decode()


def nccl_unique_id():
    """128 bytes made on rank 0 (ncclGetUniqueId); distribute them to every rank before Simulate(..., nccl_id=)."""
    buf = C.create_string_buffer(128)
    L = lib()
    if L.ddcb200_ncclUniqueId(buf) != 0:
        raise DdcError(L.ddcb200_lastError().decode())
    return buf.raw


def default_lattice(nranks, box):
    """DDC lx ly lz when the deck gives none: split the longest box edges first (membranes: in the plane)."""
    lat = [1, 1, 1]
    n = nranks
    f = 2
    factors = []
    while n > 1:
        while n % f == 0:
            factors.append(f)
            n //= f
        f += 1
    for f in sorted(factors, reverse=True):
        a = max(range(3), key=lambda k: box[k] / lat[k])
        lat[a] *= f
    return tuple(lat)


def ddc_plan(h, lattice, rlist, rx, ry, rz, rank, owner_bead=None):
    """Host restatement of the domain classification (ddcb200_ddcPlan): returns (owner, mask)."""
    n = len(rx)
    a = [np.ascontiguousarray(x, np.float64) for x in (rx, ry, rz)]
    hh = np.ascontiguousarray(h, np.float64)
    owner = np.empty(n, np.int32)
    mask = np.empty(n, np.uint32)
    ob = np.ascontiguousarray(owner_bead, np.int32) if owner_bead is not None else None
    pd, pi = _P(C.c_double), _P(C.c_int)
    L = lib()
    rc = L.ddcb200_ddcPlan(hh.ctypes.data_as(pd), int(lattice[0]), int(lattice[1]), int(lattice[2]), float(rlist), n,
                           a[0].ctypes.data_as(pd), a[1].ctypes.data_as(pd), a[2].ctypes.data_as(pd),
                           ob.ctypes.data_as(pi) if ob is not None else None, int(rank), owner.ctypes.data_as(pi),
                           mask.ctypes.data_as(_P(C.c_uint32)))
    if rc != 0:
        raise DdcError(L.ddcb200_lastError().decode())
    return owner, mask


def read_cmds(filename):
    """readCMDS (src/readCmds.c:20-57)."""
    return int(lib().ddcb200_readCMDS(os.fsencode(filename)))


def simulateMaster(object_file, restart_file=None, simulate_name=None, device=0):
    """simulateMaster (src/masters.c:383-559): the whole run of a deck - data lines, ddcMD_CMDS, restarts - in the deck's directory."""
    L = lib()
    rc = L.ddcb200_simulateMaster(os.fsencode(object_file), os.fsencode(restart_file) if restart_file else None,
                                  simulate_name.encode() if simulate_name else None, int(device))
    if rc != 0:
        raise DdcError(L.ddcb200_lastHostError().decode())


def simulate_init(object_file, restart_file=None, device=0, simulate_name=None, rank=0, nranks=1, lattice=None, nccl_id=None):
    """Mirror of simulate_init(NULL, name, comm) (src/simulate.c:104) for a Martini deck."""
    return Simulate(Deck(object_file, restart_file, simulate_name), device, rank, nranks, lattice, nccl_id)
