"""Turn `ncu -i <rep> --page raw --csv` output of ONE kernel launch into the record bench.py reads for `roofline.traffic`
(profiles/traffic.json): DRAM bytes per point of the dominant kernel, tagged with the library version it was captured on.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/r2_<mode>_ncu_full.csv
    python tools/ncu_traffic.py <mode> profiles/r2_<mode>_ncu_full.csv <points in the profiled launch> "<lib version>"
"""
import csv
import json
import os
import sys

mode, path, points, ver = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hdr = rows[0]
# first data row whose kernel name is the jet kernel
ki = hdr.index("Kernel Name")
data = [r for r in rows[2:] if len(r) == len(hdr) and "jet_" in r[ki]]
r = data[0]
col = lambda n: float(r[hdr.index(n)].replace(",", ""))
unit = lambda n: rows[1][hdr.index(n)]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rd = col("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
wr = col("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
db = json.load(open(out)) if os.path.exists(out) else {}
db[mode] = {"lib_version": ver, "kernel": r[ki], "points_in_capture": points, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_point": (rd + wr) / points, "source": os.path.relpath(path, os.path.dirname(out) + "/..")}
json.dump(db, open(out, "w"), indent=1)
print(json.dumps(db[mode], indent=1))
