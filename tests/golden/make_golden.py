"""Generates tests/golden/*.json.

multirand_kat.json holds (a) the known-answer sequences copied as DATA from the reference's RNG self test
(/root/reference/src/multirand.F90:396-425) -- the only golden vectors the reference has -- and (b) the first 16
markers of the restated particle_load for the default input (SuperKISS64, constant seeds, rank 0, warm-up 5),
hex-encoded, so a later change of the oracle is caught.  The reference itself (Fortran + PETSc + MPI) cannot be
built or imported in this image, so no reference-generated hot-path vectors exist ("parity unpinned").

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

KAT = {
    "kiss64": [8932985056925012148, 5710300428094272059, -104233206776033023, -4143107803135683366,
               542381058189297533, -4244931820854714191, 6853720724624422285, -767542866500872268,
               -257204313086867125, 8128797625455304420],
    "mt19937_64_head": [-3932459287431434586, 4620546740167642908, -5337173792191653896, -983805426561117294,
                        355488278567739596, 7469126240319926998, 4635995468481642529, 418970542659199878,
                        -8842573084457035060, 6358044926049913402],
    "mt19937_64_tail": [-7948593974297132281, 1921007855220546564, 7643484074408755248, -7128315020423208677,
                        1370093900783164344, 6776537281339823025, 3450492372588984223, -9045729527952115285,
                        7896519943553875907, -4143300141377237606],
    "superkiss64_head": [6140839658375754198, -95225469143006167, -9148462456964506707, 3912874252778582253,
                         6801212277726928591, -809575511391043410, -397286769868273005, 4963780769400405858,
                         2406624640673457322, 1246843699883922102],
    "superkiss64_tail": [-1387224431860786161, -8846516422183390713, 8111357788999165247, 444070776306226770,
                         -7730678117654887867, -296399128303442035, -1658509282659454084, -8190332265239255687,
                         -1492517620356299342, -5016179395587873849],
}

if __name__ == "__main__":
    o = O.Oracle(O.default_params())
    x, v, p, w = o.particle_load(0, 3, 0, 5, 16, 16)
    KAT["particle_load_default_rank0_n16"] = {
        k: [float.hex(float(t)) for t in a] for k, a in (("x", x), ("v", v), ("p", p), ("w", w))}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multirand_kat.json")
    json.dump(KAT, open(out, "w"), indent=1)
    print("wrote", out)
