"""Synthetic Martini systems in ddcMD's own deck formats (object.data, martini.data,
restraint.data, restart, snapshot.mem/atoms#000000), with fixed seeds.

The reference ships one Martini input (examples/waterbox: water only).  The configs named in
BASELINE.json (POPC bilayer ~100k beads, RAS-like bonded-heavy patch ~140k, 1M and 10M
membranes) are synthesised here so the SAME files feed the reference CPU path
(oracle/_ref) and this package.  Force-field numbers follow the public Martini 2.x
conventions (sigma 0.47 nm, 0.62 nm for charged-apolar pairs, epsilon levels 2.0-5.6
kJ/mol, 1250 kJ/mol/nm^2 bonds, G96 cosine angles); parity only needs both codes to read
the same file, so exact force-field fidelity is not a goal.

Deck conventions the reference requires (and this writer keeps):
  * species are named <RES>x<ATOM> (src/bioMartini.c:1280-1284), <RES> has at most 4
    characters and no lower-case c/n/x (src/bioCharmmParms.c:2488-2507);
  * gid = molecule<<32 | atom offset (src/bioGid.h:13-23): one residue per molecule, one
    group per residue, so (gid & 0xffff) is the atom offset reOrgPairs matches bpairs on;
  * theta0 of func-2 / func-10 angles is a cosine; func-1 angles and dihedral phases are radians.
"""
import math
import os

import numpy as np

KB_KJMOL = 8.3144598e-3  # kJ/mol/K

# ---------------------------------------------------------------------------------------------
# force field
# ---------------------------------------------------------------------------------------------
TYPES = ["BP4", "P4", "Q0", "Qa", "Qd", "Na", "C1", "C3", "P5", "P1", "C5", "Nda"]
_POLARITY = {"Q0": 0, "Qa": 0, "Qd": 0, "P5": 1, "P4": 2, "BP4": 2, "P1": 3, "Nda": 4, "Na": 4, "C5": 5, "C3": 6, "C1": 7}
_CHARGED = {"Q0", "Qa", "Qd"}


def lj_pair(a, b):
    """(sigma nm, eps kJ/mol) for a Martini-2-like interaction matrix."""
    pa, pb = _POLARITY[a], _POLARITY[b]
    eps = max(2.0, 5.0 - 0.43 * abs(pa - pb))
    if pa <= 2 and pb <= 2:
        eps = min(5.6, eps + 0.6)
    if pa >= 5 and pb >= 5:
        eps = 3.5
    sigma = 0.47
    if (a in _CHARGED and b == "C1") or (b in _CHARGED and a == "C1"):
        sigma, eps = 0.62, 2.0
    if {a, b} == {"P4", "BP4"}:
        sigma, eps = 0.57, 5.6
    if a == b == "P4" or a == b == "BP4":
        eps = 5.0
    return sigma, round(eps, 3)


class Residue:
    """One RESIPARMS template = one molecule type."""

    def __init__(self, name, atoms):
        self.name = name
        self.atoms = atoms            # list of (atomName, atomType, charge, mass_amu-ish in M_p)
        self.bonds = []               # (i, j, kb kJ/mol/nm^2 [ddcMD convention E=kb(b-b0)^2], b0 nm)
        self.angles = []              # (i, j, k, func, ktheta kJ/mol, theta0)
        self.dihedrals = []           # (i, j, k, l, func, n, kchi, delta)
        self.exclusions = []          # (i, j)
        self.constraints = []         # (i, j, r0 nm)

    @property
    def natoms(self):
        return len(self.atoms)


def popc_like(name="POPC", head=("NC3", "Q0", 1.0), second=("D2A", "C3"), kink=True):
    atoms = [(head[0], head[1], head[2], 72.0), ("PO4", "Qa", -1.0, 72.0), ("GL1", "Na", 0.0, 72.0), ("GL2", "Na", 0.0, 72.0),
             ("C1A", "C1", 0.0, 72.0), (second[0], second[1], 0.0, 72.0), ("C3A", "C1", 0.0, 72.0), ("C4A", "C1", 0.0, 72.0),
             ("C1B", "C1", 0.0, 72.0), ("C2B", "C1", 0.0, 72.0), ("C3B", "C1", 0.0, 72.0), ("C4B", "C1", 0.0, 72.0)]
    r = Residue(name, atoms)
    kb = 625.0  # = 1250/2
    for i, j, b0 in [(0, 1, 0.47), (1, 2, 0.47), (2, 3, 0.37), (2, 4, 0.47), (4, 5, 0.47), (5, 6, 0.47), (6, 7, 0.47),
                     (3, 8, 0.47), (8, 9, 0.47), (9, 10, 0.47), (10, 11, 0.47)]:
        r.bonds.append((i, j, kb, b0))
    c180, c120 = -1.0, -0.5
    for i, j, k, kt, c in [(1, 2, 3, 12.5, c120), (1, 2, 4, 12.5, c180), (2, 4, 5, 12.5, c180), (4, 5, 6, 22.5 if kink else 12.5, c120 if kink else c180),
                           (5, 6, 7, 12.5, c180), (3, 8, 9, 12.5, c180), (8, 9, 10, 12.5, c180), (9, 10, 11, 12.5, c180)]:
        r.angles.append((i, j, k, 2, kt, c))
    return r


def single_bead(name, atom, typ, charge=0.0, mass=72.0):
    return Residue(name, [(atom, typ, charge, mass)])


# ---------------------------------------------------------------------------------------------
# geometry helpers
# ---------------------------------------------------------------------------------------------
_LIPID_Z = [30.0, 25.5, 21.0, 21.0, 16.5, 12.0, 7.5, 3.0, 16.5, 12.0, 7.5, 3.0]


def _lipid_coords(x, y, s, up):
    """12 bead positions of one lipid whose chain A site is (x,y) and chain B site is (x+s,y)."""
    dx = [0.6, 0.3, 0.0, s, 0.0, 0.0, 0.0, 0.0, s, s, s, s]
    out = np.empty((12, 3))
    for k in range(12):
        out[k] = (x + dx[k], y, _LIPID_Z[k] * (1 if up else -1))
    return out


def _protein(rng, nbb, center, axis_len):
    """A helical-tube 'protein-like' chain: nbb backbone beads + a side-chain bead on every
    second residue, lying along x.  Topology is designed around the generated coordinates
    (b0 / theta0 / delta taken from the geometry) so the structure is at its own minimum."""
    R, rise = 9.0, 3.6          # helix radius, arc step (Angstrom)
    pitch = 6.2
    pts = []
    t = 0.0
    for _ in range(nbb):
        pts.append((center[0] - axis_len / 2 + pitch * t / (2 * math.pi), center[1] + R * math.cos(t), center[2] + R * math.sin(t)))
        t += rise / math.sqrt(R * R + (pitch / (2 * math.pi)) ** 2)
    bb = np.array(pts)
    atoms, coords = [], []
    bb_idx, sc_idx = [], {}
    bbtypes = ["P5", "Nda", "P5", "P1"]
    for i in range(nbb):
        bb_idx.append(len(atoms))
        atoms.append(("B%03d" % i, bbtypes[i % 4], 0.0, 72.0))
        coords.append(bb[i])
        if i % 2 == 1:
            out = bb[i] - np.array([bb[i][0], center[1], center[2]])
            out /= np.linalg.norm(out)
            q = 0.0
            typ = "C5"
            if i % 10 == 1:
                q, typ = 1.0, "Qd"
            elif i % 10 == 5:
                q, typ = -1.0, "Qa"
            elif i % 6 == 3:
                typ = "C3"
            sc_idx[i] = len(atoms)
            atoms.append(("S%03d" % i, typ, q, 72.0))
            coords.append(bb[i] + 3.6 * out + rng.normal(0, 0.05, 3))
    coords = np.array(coords)
    res = Residue("PROT", atoms)

    def dist(a, b):
        return float(np.linalg.norm(coords[a] - coords[b])) / 10.0  # nm

    def cosang(a, b, c):
        u, v = coords[a] - coords[b], coords[c] - coords[b]
        return float(np.dot(u, v) / np.linalg.norm(u) / np.linalg.norm(v))

    def dihed(a, b, c, d):
        # same convention as bioDihedralFast (src/bioCharmmCovalentEnergies.c:266-351)
        vij, vjk, vkl = coords[a] - coords[b], coords[b] - coords[c], coords[c] - coords[d]
        m, n = np.cross(vij, vjk), np.cross(vjk, vkl)
        x = np.dot(m, n) / np.linalg.norm(m) / np.linalg.norm(n)
        sign = -1.0 if np.dot(vjk, np.cross(m, n)) < 0 else 1.0
        return sign * math.acos(max(-1.0, min(1.0, x)))

    for i in range(nbb - 1):
        res.bonds.append((bb_idx[i], bb_idx[i + 1], 625.0, dist(bb_idx[i], bb_idx[i + 1])))
    for i, s in sc_idx.items():
        res.bonds.append((bb_idx[i], s, 2500.0, dist(bb_idx[i], s)))
    # elastic network: BB pairs |i-j|>=3 within 0.9 nm, at most 6 per bead
    cnt = np.zeros(nbb, int)
    for i in range(nbb):
        for j in range(i + 3, min(nbb, i + 40)):
            d = dist(bb_idx[i], bb_idx[j])
            if d < 0.9 and cnt[i] < 6 and cnt[j] < 6:
                res.bonds.append((bb_idx[i], bb_idx[j], 250.0, d))
                cnt[i] += 1
                cnt[j] += 1
    for i in range(nbb - 2):
        a, b, c = bb_idx[i], bb_idx[i + 1], bb_idx[i + 2]
        ca = cosang(a, b, c)
        if i % 3 == 0:
            res.angles.append((a, b, c, 1, 20.0, math.acos(ca)))     # harmonic in the angle
        elif i % 3 == 1:
            res.angles.append((a, b, c, 2, 20.0, ca))                # cosine harmonic
        else:
            res.angles.append((a, b, c, 10, 10.0, ca))               # restricted bending
        res.exclusions.append((a, c))
    for i, s in sc_idx.items():
        if i + 1 < nbb:
            res.angles.append((s, bb_idx[i], bb_idx[i + 1], 2, 12.5, cosang(s, bb_idx[i], bb_idx[i + 1])))
    for i in range(0, nbb - 3):
        a, b, c, d = (bb_idx[i + k] for k in range(4))
        phi = dihed(a, b, c, d)
        # E = kchi (1 + cos(n phi - delta)) has its minimum at n*phi - delta = pi
        res.dihedrals.append((a, b, c, d, 1, 1, 5.0, phi - math.pi if phi > 0 else phi + math.pi))
    for i in sorted(sc_idx)[: max(1, len(sc_idx) // 4)]:
        if 1 <= i < nbb - 1:
            a, b, c, d = bb_idx[i], bb_idx[i - 1], bb_idx[i + 1], sc_idx[i]
            res.dihedrals.append((a, b, c, d, 2, 1, 25.0, dihed(a, b, c, d)))   # improper at its own minimum
    # a few "constraints" (func 1) between BB i and i+3 that are not bonds/exclusions -> extra bpairs
    for i in range(0, nbb - 5, 17):
        res.constraints.append((bb_idx[i], bb_idx[i + 4], dist(bb_idx[i], bb_idx[i + 4])))
    # move every bead slightly off the minimum so all bonded terms carry force at step 0
    coords = coords + rng.normal(0, 0.12, coords.shape)
    return res, coords


# ---------------------------------------------------------------------------------------------
# system
# ---------------------------------------------------------------------------------------------
class SynthSystem:
    def __init__(self, box, residues, mol_res, mol_start, coords, seed, temperature):
        self.box = np.asarray(box, float)          # Angstrom, orthorhombic, centred on 0
        self.residues = residues                   # list of Residue (molecule types)
        self.mol_res = np.asarray(mol_res, np.int32)      # residue index of every molecule
        self.mol_start = np.asarray(mol_start, np.int64)  # first bead of every molecule (+ end)
        self.coords = coords                       # (n,3) Angstrom
        self.seed = seed
        n = len(coords)
        nat = np.array([r.natoms for r in residues])[self.mol_res]
        self.bead_mol = np.repeat(np.arange(len(self.mol_res), dtype=np.int64), nat)
        self.bead_atom = (np.arange(n, dtype=np.int64) - self.mol_start[self.bead_mol]).astype(np.int64)
        self.gid = (self.bead_mol.astype(np.uint64) << np.uint64(32)) | self.bead_atom.astype(np.uint64)
        rng = np.random.default_rng(seed + 1000)
        masses = np.empty(n)
        off = 0
        per_res_mass = [np.array([a[3] for a in r.atoms]) for r in residues]
        for m, r in enumerate(self.mol_res):
            k = residues[r].natoms
            masses[off:off + k] = per_res_mass[r]
            off += k
        self.mass = masses
        # Maxwell-Boltzmann in Angstrom/fs: sigma = sqrt(kB T / m); 1 (kJ/mol)/(g/mol) = 1e6 m^2/s^2 = 1e-4 (A/fs)^2
        sig = np.sqrt(KB_KJMOL * temperature / masses * 1e-4)
        v = rng.normal(size=(n, 3)) * sig[:, None]
        v -= (v * masses[:, None]).sum(0) / masses.sum()
        self.vel = v
        # shuffle file order a little?  No: the reference keeps file order; beads of a molecule are contiguous here.

    @property
    def n(self):
        return len(self.coords)

    def species_names(self):
        out = []
        for r in self.residues:
            out += ["%sx%s" % (r.name, a[0]) for a in r.atoms]
        return out

    # -- deck writer -----------------------------------------------------------------------
    def write_deck(self, path, dt=20.0, update_rate=20, deltaloop=100, printrate=100, cutoff=11.0, deltaR=4.0,
                   restraints=None):
        os.makedirs(os.path.join(path, "snapshot.mem"), exist_ok=True)
        res = self.residues
        heap = max(1000, int(self.n * 0.004))
        mol_names = " ".join(r.name + "x" for r in res)
        obj = []
        obj.append("""simulate SIMULATE
{
   type = MD; system=system; integrator=nglf;
   deltaloop=%d; maxloop=1000000; dt=%g; printrate=%d; snapshotrate=100000000; checkpointrate=100000000;
   nLoopDigits=12; gidFormat=hex; printinfo=printinfo; heap=heap; ddc=ddc;
}
energyInfo ENERGYINFO{}
heap HEAP { size = %d ;}
ddc DDC { updateRate=%d; }
printinfo PRINTINFO { PRESSURE=bar; VOLUME=Ang^3; TEMPERATURE=K; ENERGY=kJ/mol; TIME=ns; printStress=0; printMolecularPressure=1; }
martini POTENTIAL
{
   type = MARTINI; excludePotentialTerm=0; use_transform=0;
   cutoff=%g Angstrom; rcoulomb=%g Angstrom; epsilon_r=15; epsilon_rf=-1;
   function=lennardjones; parmfile=martini.data;
}
restraint POTENTIAL { type=RESTRAINT; parmfile=restraint.data; }
nglf INTEGRATOR { type = NGLF; }
system SYSTEM
{
   type = NORMAL; potential = martini%s; neighbor=nbr; groups= group free; random = lcg64;
   box = box; collection=collection; moleculeClass = moleculeClass; nConstraints=0;
}
box BOX { type=ORTHORHOMBIC; pbc=7; }
nbr NEIGHBOR { type = NORMAL; deltaR=%g; minBoxSide=6; }
group GROUP { type = FREE; }
free GROUP { type = FREE; }
lcg64 RANDOM {type = LCG64; randomizeSeed=0;}
moleculeClass MOLECULECLASS { molecules = %s; }
""" % (deltaloop, dt, printrate, heap, update_rate, cutoff, cutoff, " restraint" if restraints else "", deltaR, mol_names))
        sid = 0
        for r in res:
            sp = ["%sx%s" % (r.name, a[0]) for a in r.atoms]
            obj.append("%sx MOLECULE {ownershipSpecies = %s; species = %s;}\n" % (r.name, sp[0], " ".join(sp)))
            for a, name in zip(r.atoms, sp):
                obj.append("%s SPECIES { type = ATOM ; charge =%r; id=%d; mass =%r M_p ; }\n" % (name, float(a[2]), sid, float(a[3])))
                sid += 1
        with open(os.path.join(path, "object.data"), "w") as f:
            f.write("".join(obj))

        # martini.data
        used = sorted({a[1] for r in res for a in r.atoms}, key=TYPES.index)
        tid = {t: i for i, t in enumerate(used)}
        mm = ["martini MMFF\n{\nresiParms=%s ;\natomTypeList=%s ;\nljParms=%s ;\n}\n" % (
            " ".join(r.name for r in res), " ".join(used),
            " ".join("%s_%s" % (a, b) for i, a in enumerate(used) for b in used[i:]))]
        for t in used:
            mm.append("%s MASSPARMS { atomType=%s; atomTypeID=%d; mass=72.0M_p ; }\n" % (t, t, tid[t]))
        for ri, r in enumerate(res):
            def lst(prefix, n):
                return " ".join("%s_%s%d" % (r.name, prefix, k) for k in range(n))
            lines = ["%s RESIPARMS\n{\n  resID=%d; resType=0; resName=%s; charge=%r; groupList=%s_g0; centerAtom=0;\n" % (
                r.name, ri + 1, r.name, float(sum(a[2] for a in r.atoms)), r.name)]
            if r.bonds:
                lines.append("  bondList=%s;\n" % lst("b", len(r.bonds)))
            if r.angles:
                lines.append("  angleList=%s;\n" % lst("a", len(r.angles)))
            if r.dihedrals:
                lines.append("  dihedralList=%s;\n" % lst("d", len(r.dihedrals)))
            if r.exclusions:
                lines.append("  exclusionList=%s;\n" % lst("e", len(r.exclusions)))
            if r.constraints:
                lines.append("  constraintList=%s_cl0;\n" % r.name)
            lines.append("}\n")
            lines.append("%s_g0 GROUPPARMS{ groupID=0; atomList=%s ; }\n" % (r.name, " ".join("%s_%s" % (r.name, a[0]) for a in r.atoms)))
            for k, a in enumerate(r.atoms):
                lines.append("%s_%s ATOMPARMS{atomID=%d; atomName=%s; atomType=%s; atomTypeID=%d; charge=%r; mass=%r M_p ; }\n" % (
                    r.name, a[0], k, a[0], a[1], tid[a[1]], float(a[2]), float(a[3])))
            for k, (i, j, kb, b0) in enumerate(r.bonds):
                lines.append("%s_b%d BONDPARMS{atomI=%d; atomJ=%d; func=1; atomTypeI=%s; atomTypeJ=%s; kb=%r kJ*mol^-1*nm^-2; b0=%r nm; }\n" % (
                    r.name, k, i, j, r.atoms[i][1], r.atoms[j][1], float(kb), float(b0)))
            for k, (i, j, kk, func, kt, t0) in enumerate(r.angles):
                lines.append("%s_a%d ANGLEPARMS{atomI=%d; atomJ=%d; atomK=%d; func=%d; ktheta=%r kJ*mol^-1; theta0=%r; }\n" % (
                    r.name, k, i, j, kk, func, float(kt), float(t0)))
            for k, (i, j, kk, l, func, n, kchi, delta) in enumerate(r.dihedrals):
                lines.append("%s_d%d TORSPARMS{atomI=%d; atomJ=%d; atomK=%d; atomL=%d; func=%d; n=%d; kchi=%r kJ*mol^-1; delta=%r; }\n" % (
                    r.name, k, i, j, kk, l, func, n, float(kchi), float(delta)))
            for k, (i, j) in enumerate(r.exclusions):
                lines.append("%s_e%d EXCLUDEPARMS{atomI=%d; atomJ=%d; atomTypeI=%s; atomTypeJ=%s; }\n" % (
                    r.name, k, i, j, r.atoms[i][1], r.atoms[j][1]))
            if r.constraints:
                lines.append("%s_cl0 CONSLISTPARMS{ constraintSubList=%s; }\n" % (r.name, lst("c", len(r.constraints))))
                for k, (i, j, r0) in enumerate(r.constraints):
                    lines.append("%s_c%d CONSPARMS{atomI=%d; atomJ=%d; atomTypeI=%s; atomTypeJ=%s; func=1; r0=%r nm; }\n" % (
                        r.name, k, i, j, r.atoms[i][1], r.atoms[j][1], float(r0)))
            mm.append("".join(lines))
        for i, a in enumerate(used):
            for b in used[i:]:
                s, e = lj_pair(a, b)
                mm.append("%s_%s LJPARMS{atomtypeI=%s; indexI=%d; atomtypeJ=%s; indexJ=%d; sigma=%r nm; eps=%r kJ*mol^-1;}\n" % (a, b, a, tid[a], b, tid[b], s, e))
        with open(os.path.join(path, "martini.data"), "w") as f:
            f.write("".join(mm))

        # restraint.data: (bead index, kb, fc) tuples -> RESTRAINTPARMS with the bead's current fractional position.
        # (The reference's object syntax check rejects an empty "restraintList=;", so the RESTRAINT potential is
        # only listed in SYSTEM when there is at least one restraint.)
        rl = ["restraint RESTRAINTLIST{\n  xbox = %r;\n  ybox = %r;\n  zbox = %r;\n  restraintList=%s;\n}\n" % (
            float(self.box[0]), float(self.box[1]), float(self.box[2]), " ".join("rst%d" % k for k in range(len(restraints or []))))]
        for k, (bead, kb, fc) in enumerate(restraints or []):
            fr = self.coords[bead] / self.box + 0.5
            rl.append("rst%d RESTRAINTPARMS{ gid=%d; atomI=0; func=1; fcx=%d; fcy=%d; fcz=%d; x0=%r; y0=%r; z0=%r; kb=%r kJ*mol^-1*nm^-2; }\n" % (
                k, int(self.gid[bead]), fc[0], fc[1], fc[2], float(fr[0]), float(fr[1]), float(fr[2]), float(kb)))
        with open(os.path.join(path, "restraint.data"), "w") as f:
            f.write("".join(rl))

        # restart
        with open(os.path.join(path, "snapshot.mem", "restart"), "w") as f:
            f.write("simulate SIMULATE { loop=0; time=0.000000 ;}\nbox BOX {\nh= %r 0.0 0.0\n   0.0 %r 0.0\n   0.0 0.0 %r ;\n}\n"
                    "collection COLLECTION { mode=VARRECORDASCII; size=%d; files=snapshot.mem/atoms#;}\n" % (
                        float(self.box[0]), float(self.box[1]), float(self.box[2]), self.n))
        link = os.path.join(path, "restart")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("snapshot.mem/restart", link)

        # atoms
        names = self.species_names()
        res_off = np.cumsum([0] + [r.natoms for r in res])
        spec = res_off[self.mol_res][self.bead_mol] + self.bead_atom
        with open(os.path.join(path, "snapshot.mem", "atoms#000000"), "w") as f:
            f.write("particle FILEHEADER {type=MULTILINE; datatype=VARRECORDASCII; checksum=NONE;\nloop=0; time=0.000000;\n"
                    "nfiles=1; nrecord=%d; nfields=10;\nfield_names=id class type group rx ry rz vx vy vz;\nfield_types=u s s s f f f f f f;\n"
                    "h= %r 0.0 0.0\n   0.0 %r 0.0\n   0.0 0.0 %r ;\ngroups = group ;\ntypes = ATOM ;\n} \n\n" % (
                        self.n, float(self.box[0]), float(self.box[1]), float(self.box[2])))
            c, v = self.coords, self.vel
            chunk = 200000
            for a in range(0, self.n, chunk):
                b = min(self.n, a + chunk)
                f.write("".join("%16d ATOM %11s group %22.15e %22.15e %22.15e %22.15e %22.15e %22.15e\n" % (
                    int(self.gid[i]), names[spec[i]], c[i, 0], c[i, 1], c[i, 2], v[i, 0], v[i, 1], v[i, 2]) for i in range(a, b)))
        return path


def make_membrane(lx=150.0, ly=150.0, lz=110.0, seed=1, protein_beads=0, temperature=240.0, lipid_mix=True, ions=True):
    """Bilayer in the xy plane (centre z=0) + water (5% antifreeze) + ions (+ optional protein-like chain).
    ~0.0083 beads/A^3.  lx, ly are rounded so the lipid lattice tiles the box exactly."""
    rng = np.random.default_rng(seed)
    s = 5.657                               # chain-site spacing: 64 A^2 per lipid
    nx = max(1, int(round(lx / (2 * s))))
    ny = max(1, int(round(ly / s)))
    lx, ly = nx * 2 * s, ny * s
    residues = [popc_like("POPC"), popc_like("POPE", head=("NH3", "Qd", 1.0), second=("C2A", "C1"), kink=False),
                single_bead("W", "W", "P4"), single_bead("WF", "WF", "BP4"),
                single_bead("NA", "NA", "Qd", 1.0), single_bead("CL", "CL", "Qa", -1.0)]
    R_POPC, R_POPE, R_W, R_WF, R_NA, R_CL = range(6)
    mol_res, chunks = [], []
    prot_coords = None
    zwat = 34.5
    if protein_beads > 0:
        nbb = int(round(protein_beads / 1.5))
        axis_len = 6.2 * nbb * 3.6 / math.sqrt(81 + (6.2 / (2 * math.pi)) ** 2) / (2 * math.pi)
        if axis_len + 20 > lx or lz / 2 - zwat < 36:
            raise ValueError("box too small for the protein-like chain")
        pres, prot_coords = _protein(rng, nbb, (0.0, 0.0, zwat + (lz / 2 - zwat) / 2 + 1.0), axis_len)
        residues.append(pres)
        mol_res.append(len(residues) - 1)
        chunks.append(prot_coords)
    # lipids
    for up in (True, False):
        for ix in range(nx):
            for iy in range(ny):
                x = -lx / 2 + ix * 2 * s + 1.2 + (0.0 if up else s * 0.5)
                y = -ly / 2 + iy * s + 1.2
                c = _lipid_coords(x, y, s, up)
                c += rng.normal(0, 0.15, c.shape)
                kind = R_POPE if (lipid_mix and (ix * 7 + iy * 3) % 4 == 0) else R_POPC
                mol_res.append(kind)
                chunks.append(c)
    # water on an fcc lattice (nearest-neighbour distance ~5.5 A => ~0.0085 beads/A^3, every neighbour just
    # outside the LJ minimum of sigma = 4.7 A) in |z| > zwat
    nn = 5.5
    a = nn * math.sqrt(2.0)
    wx, wy = max(1, int(round(lx / a))), max(1, int(round(ly / a)))
    ax, ay = lx / wx, ly / wy
    nzw = max(1, int(round((lz / 2 - zwat - 1.0) / a)))
    az = (lz / 2 - zwat - 1.0) / nzw
    basis = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])
    cells = np.stack(np.meshgrid(np.arange(wx), np.arange(wy), np.arange(nzw), indexing="ij"), -1).reshape(-1, 1, 3)
    frac = (cells + basis[None, :, :]).reshape(-1, 3)
    Wup = np.stack([-lx / 2 + (frac[:, 0] + 0.25) * ax, -ly / 2 + (frac[:, 1] + 0.25) * ay, zwat + 1.5 + frac[:, 2] * az], 1)
    W = np.concatenate([Wup, Wup * np.array([1.0, 1.0, -1.0])])
    W += rng.uniform(-0.15, 0.15, W.shape)
    if prot_coords is not None:
        # drop water sites within 4.6 A of any protein bead (coarse grid search)
        from scipy.spatial import cKDTree
        near = cKDTree(prot_coords).query_ball_point(W, 4.6, return_length=True)
        W = W[near == 0]
    kinds = np.full(len(W), R_W)
    u = rng.random(len(W))
    kinds[u < 0.05] = R_WF
    if ions:
        nion = max(1, len(W) // 184)
        idx = rng.choice(len(W), 2 * nion, replace=False)
        kinds[idx[:nion]] = R_NA
        kinds[idx[nion:]] = R_CL
    mol_res += list(kinds)
    chunks.append(W)
    coords = np.concatenate([np.atleast_2d(c) for c in chunks])
    nat = np.array([residues[r].natoms for r in mol_res])
    mol_start = np.concatenate([[0], np.cumsum(nat)])
    box = np.array([lx, ly, lz])
    # wrap into [-L/2, L/2)
    coords = (coords + box / 2) % box - box / 2
    if protein_beads > 0:
        # net charge of the chain is balanced by flipping surplus counter-ions back to water
        q = sum(a[2] for a in residues[-1].atoms)
        flip = R_NA if q > 0 else R_CL
        k = int(abs(round(q)))
        mr = np.array(mol_res)
        cand = np.nonzero(mr == flip)[0][:k]
        mr[cand] = R_W
        mol_res = list(mr)
    return SynthSystem(box, residues, mol_res, mol_start, coords, seed, temperature)


def make_waterbox(l=60.0, seed=3, temperature=300.0):
    rng = np.random.default_rng(seed)
    a = 4.95
    n = int(l / a)
    g = -l / 2 + (np.arange(n) + 0.5) * (l / n)
    W = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + rng.uniform(-0.35, 0.35, (n ** 3, 3))
    residues = [single_bead("W", "W", "P4"), single_bead("WF", "WF", "BP4")]
    kinds = (rng.random(len(W)) < 0.1).astype(int)
    return SynthSystem([l, l, l], residues, list(kinds), np.arange(len(W) + 1), W, seed, temperature)


CONFIGS = {
    # name: (builder kwargs, description) - sizes follow BASELINE.json "configs"
    "popc_small": (dict(lx=68.0, ly=68.0, lz=112.0, seed=1), "small POPC/POPE bilayer + water + ions (~4k beads), test size"),
    "ras_small": (dict(lx=136.0, ly=68.0, lz=160.0, seed=2, protein_beads=120), "small bilayer + protein-like chain (~12k beads), test size"),
    "popc_100k": (dict(lx=300.0, ly=300.0, lz=130.0, seed=1), "POPC/POPE bilayer + water, ~100k beads"),
    "ras_140k": (dict(lx=340.0, ly=340.0, lz=150.0, seed=2, protein_beads=350), "RAS-like protein chain over a mixed membrane, ~140k beads"),
    "membrane_1m": (dict(lx=1060.0, ly=1060.0, lz=130.0, seed=3), "multi-lipid membrane, ~1M beads"),
    "membrane_10m": (dict(lx=3352.0, ly=3352.0, lz=130.0, seed=3), "multi-lipid membrane, ~10M beads"),
}


def make(name):
    kw, _ = CONFIGS[name]
    return make_membrane(**kw)
