#!/usr/bin/env python
"""Convert the CRUST2.0 text tables that ship with the reference (aux/crust2x2/CNtype2_key.txt,
CNtype2.txt, CNelevatio2.txt; public model of Laske, Masters & Reif) into the compact binary table
kiwi_b200/data/crust2x2.kcr that the engine (and the oracle) load at run time.

    python tools/make_crust2x2_table.py /root/reference/aux/crust2x2 kiwi_b200/data/crust2x2.kcr

Layout (little endian): magic "KCR1", int32 ntypes, nlo, nla; per type 31 float32 exactly as printed
in the key file (vp[8], vs[8], rho[8] in km/s and g/cm3, thickness[7] in km: unit conversion and the
ice/water swap of crust2x2.f90:281-297 are done by the loader in fp32, as the reference does);
int16 type index per (lat row, lon column); float32 elevation per (lat row, lon column).
"""
import struct
import sys

import numpy as np

NL, NTYPES, NLA, NLO = 7, 360, 90, 180


def main(src, dst):
    lines = open(src + "/CNtype2_key.txt").read().splitlines()[5:]
    ids, table = [], np.zeros((NTYPES, 31), np.float32)
    for i in range(NTYPES):
        blk = lines[5 * i:5 * i + 5]
        ids.append(blk[0].split()[0][:2])
        vp = [float(x) for x in blk[1].split()[:NL + 1]]
        vs = [float(x) for x in blk[2].split()[:NL + 1]]
        rho = [float(x) for x in blk[3].split()[:NL + 1]]
        th = [float(x) for x in blk[4].split()[:NL]]
        table[i] = np.array(vp + vs + rho + th, np.float32)
    index = {}
    for i, s in enumerate(ids):
        index.setdefault(s, i)          # first match wins, as in the reference's type_loop
    rows = open(src + "/CNtype2.txt").read().splitlines()[1:1 + NLA]
    tmap = np.zeros((NLA, NLO), np.int16)
    for j, row in enumerate(rows):
        toks = row.split()[1:1 + NLO]
        tmap[j] = [index[t[:2]] for t in toks]
    rows = open(src + "/CNelevatio2.txt").read().splitlines()[1:1 + NLA]
    elev = np.zeros((NLA, NLO), np.float32)
    for j, row in enumerate(rows):
        elev[j] = [float(x) for x in row.split()[1:1 + NLO]]
    with open(dst, "wb") as f:
        f.write(b"KCR1" + struct.pack("<iii", NTYPES, NLO, NLA))
        f.write(table.tobytes()); f.write(tmap.tobytes()); f.write(elev.tobytes())
    print("wrote", dst, table.shape, tmap.shape, elev.shape)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
