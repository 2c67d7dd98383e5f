#!/usr/bin/env python3
"""Per-kernel-family constants for bench.py's roofline from an ncu raw-page CSV of tools/ncu_capture.sh
(first two wavefront iterations of scene 6 FHD: 4 147 200 camera rays, then the survivors).
usage: ncu_constants.py gpurun_out/TAG_raw.csv > profiles/r2_kernels_ncu.json"""
import csv
import json
import sys

FAMILY = [("k_top<1", "top_new"), ("k_top<(bool)1", "top_new"), ("k_top<0", "top_old"), ("k_top<(bool)0", "top_old"), ("k_mesh_multi", "mesh_multi"), ("k_mesh_enter", "mesh_enter"),
          ("k_mesh_walk", "mesh_walk"), ("k_shade<0", "miss"), ("k_shade<(int)0", "miss"), ("k_shade<1", "light"), ("k_shade<(int)1", "light"),
          ("k_shade<2", "diffuse"), ("k_shade<(int)2", "diffuse"), ("k_shade<3", "metal"), ("k_shade<(int)3", "metal"), ("k_shade<4", "glass"),
          ("k_shade<(int)4", "glass"), ("k_shade<5", "principled"), ("k_shade<(int)5", "principled")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)

    # iteration boundaries: a k_top launch starts an iteration; rays of the iteration = its grid x 128 threads
    its, cur = [], None
    for r in data:
        name = r[col["Kernel Name"]].replace("void ", "").replace("ptd::", "")
        if name.startswith("k_top"):
            cur = {"rays": int(float(r[col["launch__grid_size"]].replace(",", ""))) * 128, "launches": []}
            its.append(cur)
        if cur is not None:
            cur["launches"].append((name, r))
    fam = {}
    for it in its:
        seen = set()
        for name, r in it["launches"]:
            f = next((fam_name for prefix, fam_name in FAMILY if name.startswith(prefix)), None)
            if f is None:
                continue
            e = fam.setdefault(f, {"kernel": name.split("(")[0], "us": 0.0, "dram": 0.0, "issue_w": 0.0, "lanes_w": 0.0, "rays": 0, "winst": 0.0, "regs": 0, "occ_w": 0.0})
            us = val(r, "gpu__time_duration.sum")
            e["us"] += us
            e["dram"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            e["issue_w"] += us * val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
            e["lanes_w"] += us * val(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
            e["occ_w"] += us * val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
            e["winst"] += val(r, "smsp__inst_executed.sum")
            e["regs"] = int(val(r, "launch__registers_per_thread"))
            if f not in seen:
                e["rays"] += it["rays"]; seen.add(f)
    out = {}
    for f, e in fam.items():
        out[f] = {"kernel": e["kernel"], "dram_B_per_seg": e["dram"] / max(e["rays"], 1), "issue_pct": e["issue_w"] / max(e["us"], 1e-9),
                  "lanes": e["lanes_w"] / max(e["us"], 1e-9), "occupancy_pct": e["occ_w"] / max(e["us"], 1e-9), "registers": e["regs"],
                  "warp_inst_per_seg": e["winst"] / max(e["rays"], 1), "ns_per_seg_under_ncu": 1e3 * e["us"] / max(e["rays"], 1), "rays_in_capture": e["rays"]}
    out["_source"] = {"csv": sys.argv[1], "iterations": [it["rays"] for it in its],
                      "note": "per segment = per ray of the wavefront iteration the launch belongs to (mesh kernels: all rounds of the iteration summed)"}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
