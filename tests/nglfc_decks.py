"""Deck variants for the NGLFCONSTRAINT tests: a golden deck with its INTEGRATOR switched to
NGLFCONSTRAINT (Berendsen molecular-pressure barostat, src/nglfconstraint.c:64-84,510-574) and/or its GROUPs switched to
LANGEVIN (src/langevin.c:92-128) with a fixed per-bead LCG64 seed.  Used by the golden generator and by the tests, so
both sides run the same files."""
import os
import re
import shutil

VARIANTS = {
    # name: (langevin groups, barostat)
    "lang": (True, False),
    "baro": (False, True),
    "full": (True, True),
}


def make_variant(golden_dir, deck, variant, dst_root):
    lang, baro = VARIANTS[variant]
    dst = os.path.join(str(dst_root), "%s_%s" % (deck, variant))
    shutil.copytree(os.path.join(golden_dir, deck), dst, symlinks=True)
    p = os.path.join(dst, "object.data")
    s = open(p).read()
    integ = "nglf INTEGRATOR {type = NGLFCONSTRAINT; T=310K; P0 = 1.0 bar; beta = %s; tauBarostat = 1.0 ps;}" % ("3.0e-4/bar" if baro else "0.0/bar")
    s, n = re.subn(r"^nglf INTEGRATOR \{ *type = NGLF; *\}", integ, s, flags=re.M)
    assert n == 1, "INTEGRATOR line not found in %s" % p
    if lang:
        s, n = re.subn(r"^(group|free) GROUP \{ type = FREE; \}", r"\1 GROUP { type = LANGEVIN; Teq=310K; tau=1ps; useDefault=0;}", s, flags=re.M)
        assert n == 2
    s = s.replace("randomizeSeed=1;", "randomizeSeed=0;")
    open(p, "w").write(s)
    return dst
