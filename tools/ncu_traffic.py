"""ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) -> profiles/ncu_traffic.json: measured DRAM bytes per launch of every
search kernel, at the launch shape the capture was taken with.  bench.py copies the dominant kernel's figure into
`roofline.traffic` when its own launches have the same shape.
usage: python tools/ncu_traffic.py raw.csv BOX ROTATIONS_PER_LAUNCH "source note" > profiles/ncu_traffic.json"""
import csv
import json
import sys

KIND = [("FwdZC<80", "fwdz_level0"), ("FwdZC<40", "fwdz_level1"), ("PlaneX<80", "plane_level0"), ("Plane<80", "plane_level0"),
        ("Plane<40", "plane_level1"), ("HiddenL1", "hidden_level1"), ("Epilogue", "epilogue"),
        ("FwdZC<128", "fwdz_level0"), ("FwdZC<64", "fwdz_level1"), ("PlaneX<128", "plane_level0"), ("Plane<64", "plane_level1")]


def main(path, box, rot, note):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    unit = {h: u for h, u in zip(hdr, rows[1])}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in data:
        name = r[idx["Kernel Name"]]
        kind = next((k for pat, k in KIND if pat in name), None)
        if kind is None or kind in out:
            continue
        rd = float(r[idx["dram__bytes_read.sum"]]) * scale[unit["dram__bytes_read.sum"]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * scale[unit["dram__bytes_write.sum"]]
        out[kind] = {"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                     "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) *
                     {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit["gpu__time_duration.sum"]]}
    json.dump({"box": int(box), "rotations_per_launch": float(rot), "source": note, "kernels": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:5])
