#!/usr/bin/env python
"""Generates the committed fixtures in tests/golden/ from the reference tree (/root/reference, read-only).
Run in the build container only; the GPU box has no /root/reference and uses the committed JSON files.

  sod_tube_ref.json   <- tests/hydro/sod-tube/sod-tube-ref.dat   (golden sums of the 1-D AMR Sod run, tol 3e-13)
  sod_tube_ana.json   <- tests/hydro/sod-tube/sod-tube-ana.dat   (exact Sod solution at t=0.245, 1024 points)
  indices3cube.json   <- amr/nbors_utils.f90:305-358             (lll/mmm neighbour lookup tables)
  implosion_ref.json  <- tests/hydro/implosion/implosion-ref.dat (golden sums, 2-D AMR; kept for later rounds)
  imhd_tube_ref.json  <- tests/mhd/imhd-tube/imhd-tube-ref.dat   (golden sums of the 1-D AMR MHD tube, hlld)
  imhd_tube_ana.json  <- tests/mhd/imhd-tube/imhd-tube-ana.dat   (exact solution at t=0.4, 2014 points)
  orszag_tang_ref.json <- tests/mhd/orszag-tang/orszag-tang-ref.dat (golden sums, 2-D AMR; kept for later rounds)
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def ref_dat(path):
    d = {}
    for line in open(path):
        k, v = line.split(":")
        d[k.strip()] = float(v)
    return d


def main():
    json.dump(ref_dat(f"{REF}/tests/hydro/sod-tube/sod-tube-ref.dat"), open(f"{OUT}/sod_tube_ref.json", "w"), indent=1)
    json.dump(ref_dat(f"{REF}/tests/hydro/implosion/implosion-ref.dat"), open(f"{OUT}/implosion_ref.json", "w"), indent=1)
    rows = [[float(x) for x in l.split()] for l in open(f"{REF}/tests/hydro/sod-tube/sod-tube-ana.dat") if l.strip()]
    json.dump({"columns": ["idx", "x", "u", "rho", "P", "e"], "rows": rows}, open(f"{OUT}/sod_tube_ana.json", "w"))
    json.dump(ref_dat(f"{REF}/tests/mhd/imhd-tube/imhd-tube-ref.dat"), open(f"{OUT}/imhd_tube_ref.json", "w"), indent=1)
    json.dump(ref_dat(f"{REF}/tests/mhd/orszag-tang/orszag-tang-ref.dat"), open(f"{OUT}/orszag_tang_ref.json", "w"), indent=1)
    rows = [[float(x) for x in l.split()] for l in open(f"{REF}/tests/mhd/imhd-tube/imhd-tube-ana.dat") if l.strip()]
    json.dump({"columns": ["x", "rho", "u", "v", "w", "Bx", "By", "Bz", "P"], "rows": rows},
              open(f"{OUT}/imhd_tube_ana.json", "w"))
    src = open(f"{REF}/amr/nbors_utils.f90").read()
    tab = {"lll": {}, "mmm": {}}
    for name in ("lll", "mmm"):
        for m in re.finditer(name + r"\(1:(\d+),(\d),(\d)\)=\(/([0-9,\s]+)/\)", src):
            n, ind, ndim, vals = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)
            v = [int(x) for x in vals.replace(" ", "").split(",")]
            assert len(v) == n
            tab[name][f"{ndim},{ind}"] = v
    json.dump(tab, open(f"{OUT}/indices3cube.json", "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
