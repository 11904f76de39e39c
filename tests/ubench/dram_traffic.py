#!/usr/bin/env python
"""profiles/dram_traffic.json from an ncu capture of one bench step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/<tag>_dram.csv python bench.py --steps 1 --warmup 3 ...
    python tests/ubench/dram_traffic.py gpurun_out/<tag>_dram.csv 32 > profiles/dram_traffic.json

Every launch is attributed to the libgbops entry-point family that issued it (sort / grid-build helper launches go to
the family of the main kernel that follows them); dram_bytes_per_launch = (read + write bytes of the family) / (number
of its main launches), i.e. per entry-point call, like bench.py's `achieved`."""
import csv
import json
import re
import sys

_I, _B0, _B1 = r"(\(int\))?", r"(\(bool\))?(0|false)", r"(\(bool\))?(1|true)"
MAIN = [  # (regex on the kernel name, family); seg_dense_kernel<CT, TPT, DIV, WEIGHTED, MUL>, seg_accum_kernel<CC, DIV, WEIGHTED>
    (r"group_fwd_kernel|group_fwd_generic", "gb_group_fwd"),
    (rf"scatter_private_kernel|seg_dense_kernel<[^>]*, ?{_I}1, ?{_B0}, ?{_I}\d+>|seg_accum_kernel<[^>]*, ?{_I}1, ?{_B0}>|group_bwd_kernel|group_bwd_generic", "gb_group_bwd"),
    (rf"seg_dense_kernel<[^>]*, ?{_I}3, ?{_B1}, ?{_I}\d+>|seg_accum_kernel<[^>]*, ?{_I}3, ?{_B1}>|interp_bwd_kernel", "gb_three_interp_bwd"),
    (r"interp_fwd", "gb_three_interp_fwd"),
    (rf"grid_query_kernel<{_B1}[,>]|query_kernel<{_B1}[,>]", "gb_cylinder_query"),
    (rf"grid_query_kernel<{_B0}[,>]|query_kernel<{_B0}[,>]", "gb_ball_query"),
    (r"fps_", "gb_fps"), (r"three_nn", "gb_three_nn"), (r"collision", "gb_collision_counts"), (r"group_xyz", "gb_group_xyz"),
    (r"gather_", "gb_gather"), (r"knn", "gb_knn"),
]
HELPER = r"seg_sort|seg_perm|grid_build"


def main():
    path, batch = sys.argv[1], int(sys.argv[2])
    launches = {}
    for r in csv.DictReader(l for l in open(path) if l.startswith('"')):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    fams, pending = {}, 0.0
    for i in sorted(launches):
        d = launches[i]
        nbytes = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        if not d["name"].startswith(("void gb::", "gb::")):
            continue
        if re.search(HELPER, d["name"]):
            pending += nbytes
            if "grid_build" in d["name"]:  # a new query call starts: the full-scan launches behind its grid query belong to it
                for f in fams.values():
                    f["_grid_pending"] = False
            continue
        for rx, fam in MAIN:
            if re.search(rx, d["name"]):
                f = fams.setdefault(fam, {"launches": 0, "dram_bytes": 0.0})
                # the full-scan query kernel launched behind a grid query is the same entry-point call
                if not (fam in ("gb_ball_query", "gb_cylinder_query") and "grid_query" not in d["name"] and f.get("_grid_pending")):
                    f["launches"] += 1
                f["_grid_pending"] = f.get("_grid_pending", False) or "grid_query" in d["name"]
                f["dram_bytes"] += nbytes + pending
                pending = 0.0
                break
    out = {"batch": batch, "backward": True, "source": path, "families": {
        k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / max(v["launches"], 1)} for k, v in sorted(fams.items())}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
