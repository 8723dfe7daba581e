#!/usr/bin/env python3
"""profiles/<step summary>.txt (tools/ncu_summary.py) -> profiles/dram_traffic.json: DRAM bytes per image of every
stage of one resident step (dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture), which
bench.py scales to its batch and reports as roofline.traffic.

    python tools/dram_traffic.py profiles/r01_v7_step_ncu_full.txt 64 32
(images per launch and stereo frames per launch of the captured run)"""
import json
import re
import sys

src, images, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
stage_of = {"pyr_resize_kernel": "pyramid", "fast_segments_kernel": "fast_cells", "octree_kernel": "quadtree",
            "blur_kernel": "blur", "orient_describe_kernel": "orient_describe", "stereo_match_kernel": "stereo_match",
            "track_grids_kernel": "track", "track_match_kernel": "track", "projection_decode_kernel": "track"}
per_frame = ("stereo_match", "track")   # these stages run once per stereo frame, the others once per image
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc, inst, launches, cur = {}, {}, {}, None
for line in open(src):
    m = re.match(r"== launch \d+: (?:void )?(\w+)", line)
    if m:
        cur = stage_of.get(m.group(1))
        if cur:
            launches[cur] = launches.get(cur, 0) + 1
        continue
    m = re.match(r"dram__bytes_(read|write)\.sum\s+([\d.]+)\s+(\w+)", line)
    if m and cur:
        acc[cur] = acc.get(cur, 0.0) + float(m.group(2)) * unit[m.group(3)]
    m = re.match(r"smsp__inst_executed\.sum\s+([\d.]+)", line)
    if m and cur:
        inst[cur] = inst.get(cur, 0.0) + float(m.group(1))
out = {"source": f"{src} (ncu --set full, {images} images per launch, dram__bytes_read.sum + dram__bytes_write.sum; pyramid = sum over "
                 f"its launches; stereo_match and track are per frame of a {frames}-frame batch; warp_instructions = smsp__inst_executed.sum)",
       "kernels": {k: {"dram_bytes_per_image": int(round(v / (frames if k in per_frame else images))),
                       "warp_instructions_per_image": int(round(inst.get(k, 0.0) / (frames if k in per_frame else images))),
                       "launches": launches[k]}
                   for k, v in acc.items()}}
json.dump(out, open("profiles/dram_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
