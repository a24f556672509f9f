#!/usr/bin/env python
"""profiles/r2_*_kernel.json (ncu --set full captures, summarize_kernel.py) -> profiles/traffic.json, which bench.py
reads for `roofline.traffic`: DRAM bytes per unit of each mapping kernel, from one capture, scaled by the units of a launch."""
import json, os
HERE = os.path.dirname(os.path.abspath(__file__))
out = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per unit from one `ncu --set full --clock-control none` capture per kernel "
               "(round 2, final kernels); bench.py scales it by the units of one launch -- a capture, not a measurement of the timed run",
       "kernels": {}}
for tag in ("r2_se_kernel", "r2_pe_kernel", "r2_rrbs_kernel", "r2_wide_kernel"):
    d = json.load(open(os.path.join(HERE, tag + ".json")))
    name = d["kernel"].split("::")[-1].split("(")[0]
    out["kernels"][name] = {"dram_bytes_per_unit": d["derived"]["dram_bytes_per_unit"], "units_in_capture": d["units_in_capture"],
                            "source": f"profiles/{tag}.json"}
json.dump(out, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))
