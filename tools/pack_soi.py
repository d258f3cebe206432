"""Pack the SOI rotation sets into the factorised form the package ships.

Every `oim<inc>.eul` table of the reference (`/root/reference/data/oim*.eul`, one
`phi\t theta\t psi` line per rotation, radians, `%f`) is M sphere points x (360/inc)
in-plane angles: runs of 360/inc lines share (theta, psi) while phi steps by `inc`
degrees from -pi.  We therefore keep only the M (theta, psi) pairs per set, as the
decimal strings found in the table; `rotations.soi_euler_angles` rebuilds the
table and `rotations.write_eul` re-emits the text byte for byte.

Run in the authoring container only (needs /root/reference):
    python tools/pack_soi.py
"""
import json
import math
import os
import sys

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, "deeplocalproteindocking_b200", "data", "soi_axes.json")


def main():
    sets = {}
    for inc in (20, 15, 12, 10, 8, 6):
        path = os.path.join(REF, "oim%02d.eul" % inc)
        rows = [l.split() for l in open(path) if l.strip()]
        nphi = 360 // inc
        assert len(rows) % nphi == 0
        axes = []
        for m in range(len(rows) // nphi):
            run = rows[m * nphi:(m + 1) * nphi]
            assert all(r[1:] == run[0][1:] for r in run), "run does not share (theta, psi)"
            for k, r in enumerate(run):
                assert r[0] == "%f" % (-math.pi + k * inc * math.pi / 180.0), (inc, m, k)
            axes.append([run[0][1], run[0][2]])
        sets[str(inc)] = axes
        print(inc, len(rows), "rotations =", len(axes), "axes x", nphi, file=sys.stderr)
    with open(OUT, "w") as f:
        f.write("{\n")
        keys = list(sets)
        for i, k in enumerate(keys):
            body = ",".join('["%s","%s"]' % (t, p) for t, p in sets[k])
            f.write('"%s":[%s]%s\n' % (k, body, "," if i + 1 < len(keys) else ""))
        f.write("}\n")


if __name__ == "__main__":
    main()
