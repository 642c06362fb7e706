#!/usr/bin/env python
"""Turns one `ncu --set full` capture of a bench step (tools/gpu_round.sh <tag> full -> gpurun_out/prof_<tag>_raw.csv) into
the two files that are committed under profiles/:

  profiles/r01_<tag>_ncu_summary.txt   one block per launch (tools/ncu_summary.py) + a per-stage table
  profiles/traffic.json                measured DRAM bytes per frame and launch time of every stage's kernels;
                                       bench.py reads it for `roofline.traffic` (per launch, like `achieved`)

Usage: python tools/profile_digest.py <tag> [frames_per_launch=64] [round=r01] [geometry, e.g. 640x480] [how it was captured]
"""
import csv
import io
import json
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

STAGE = {"k_resize": "pyramid", "k_resize_tma": "pyramid", "k_fast": "fast", "k_quadtree": "quadtree", "k_blur7": "blur", "k_blur_tc": "blur",
         "k_describe": "describe", "k_stereo_match": "stereo", "k_stereo_median": "stereo", "k_frustum": "track",
         "k_track_grid": "track", "k_track_enum": "track", "k_track_resolve": "track"}
PAIR_STAGES = ("stereo", "track")  # launched once per pair batch, not per eye


def num(v):
    return float(v.replace(",", ""))


def main(tag, frames, rnd="r01", geometry=None, how=None):
    raw = os.path.join(ROOT, "gpurun_out", "prof_%s_raw.csv" % tag)
    rows = list(csv.reader(open(raw)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[h], rows[h + 1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                     "dram__bytes_write.sum", "Grid Size")}
    # integer-ALU pipe occupancy (the roof that binds k_fast): a warp instruction holds its sub-partition's 16-lane ALU
    # pipe for 2 cycles, so  ALU warp instructions = pipe-active fraction x active cycles x 4 sub-partitions x SMs / 2
    c_alu = hdr.index("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") \
        if "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None
    c_act = hdr.index("sm__cycles_active.avg") if "sm__cycles_active.avg" in hdr else None
    n_sm = 148
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[col["gpu__time_duration.sum"]]]
    agg = {}
    seen_calls = 0
    for r in rows[h + 2:]:
        if len(r) != len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].replace("orbx::", "")
        st = STAGE.get(name)
        if st is None:
            continue
        if seen_calls >= 1 and st not in PAIR_STAGES:
            continue  # one extract call (left eye) is enough; the second repeats it
        if name == "k_describe":
            seen_calls += 1  # k_describe is the last kernel of an extract call
        a = agg.setdefault(st, {"kernel": name, "launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "alu_warp_insts": 0.0})
        a["launches"] += 1
        if c_alu is not None and c_act is not None:
            a["alu_warp_insts"] += num(r[c_alu]) / 100.0 * num(r[c_act]) * 4 * n_sm / 2.0
        a["time_us"] += num(r[col["gpu__time_duration.sum"]]) * tscale
        a["dram_bytes"] += num(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]] + \
            num(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    out = {"source": "profiles/%s_%s_ncu_summary.txt (ncu --set full --clock-control none, %d frames per launch, cold cache, "
                     "serialised)" % (rnd, tag, frames),
           "frames_per_launch": frames, "stages": {}}
    for st, a in agg.items():
        per = frames if st != "stereo" else frames  # stereo: pairs per launch == frames per eye launch
        out["stages"][st] = {"kernel": a["kernel"], "launches_per_call": a["launches"], "time_us_per_call": round(a["time_us"], 2),
                             "dram_bytes_per_frame": round(a["dram_bytes"] / per, 1)}
        if a["alu_warp_insts"] > 0:
            out["stages"][st]["alu_pipe_warp_insts_per_frame"] = round(a["alu_warp_insts"] / per, 1)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if geometry:  # keyed by geometry: bench.py looks up "<w>x<h>" first and falls back to the flat round-1 layout
        try:
            allj = json.load(open(tpath))
        except Exception:
            allj = {}
        allj[geometry] = out
        json.dump(allj, open(tpath, "w"), indent=1)
    else:
        json.dump(out, open(tpath, "w"), indent=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        print("# ncu --set full --clock-control none, one step at %d pairs per launch (%s)" % (frames, how or "tools/gpu_round.sh %s full" % tag))
        print("# per-stage totals of one extract call (+ the stereo kernels of the pair batch):")
        for st, a in out["stages"].items():
            print("#   %-9s %-16s launches %d  %8.1f us  DRAM %10.0f B/frame" % (st, a["kernel"], a["launches_per_call"],
                                                                               a["time_us_per_call"], a["dram_bytes_per_frame"]))
        ncu_summary.main(raw)
    open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.txt" % (rnd, tag)), "w").write(buf.getvalue())
    print(json.dumps(out["stages"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 64, sys.argv[3] if len(sys.argv) > 3 else "r01",
         sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else None)
