"""profiles/traffic.json from `ncu --set full` captures of one bench step per config:
dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels that make up
each timer of bench.py's per-kernel table (bench.py puts the dominant kernel's figure into
`roofline.traffic`).

    python profiles/make_traffic.py 1080p_full_n5_open3=gpurun_out/prof_r01d_1080p.ncu-rep \\
                                    4k_full_n9_oc5=gpurun_out/prof_r01d_4k.ncu-rep
"""
import collections
import csv
import json
import os
import subprocess
import sys

GROUPS = [("k_fg_bits", "fg_bits"), ("k_fg_n9", "fg_bits"), ("k_morph_mask", "morph_mask"), ("k_ccl_local", "ccl_merge"),
          ("k_ccl_boundary", "ccl_merge"), ("k_ccl_init", "ccl_merge"), ("k_ccl_merge", "ccl_merge"),
          ("k_root_count", "ccl_rank"), ("k_ccl_scan", "ccl_rank"), ("k_ccl_offsets", "ccl_rank"),
          ("k_seg_init", "ccl_rank"), ("k_root_place", "ccl_rank"), ("k_root_rank", "ccl_rank"),
          ("k_ccl_roots", "ccl_rank"), ("k_props_final", "ccl_label"), ("k_ccl_props_wide", "ccl_label"),
          ("k_write_labels", "write_labels")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def per_launch(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for row in rows[2:]:
        name = row[ik]
        for key, grp in GROUPS:
            if key in name:
                tot[(grp, key)] += float(row[ir]) * UNIT[units[ir]] + float(row[iw]) * UNIT[units[iw]]
                cnt[(grp, key)] += 1
                break
    out = collections.Counter()
    for (grp, key), b in tot.items():
        out[grp] += b / cnt[(grp, key)]          # average per launch of each kernel, summed per timer
    return {k: int(v) for k, v in out.items()}


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "traffic.json")
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    for arg in sys.argv[1:]:
        cfg, path = arg.split("=")
        data[cfg] = per_launch(path)
        data[cfg]["_source"] = os.path.basename(path) + ": ncu --set full --clock-control none, one step of bench.py"
    json.dump(data, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))
