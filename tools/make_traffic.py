"""profiles/traffic.json from an ncu --set full capture of tools/ncu_all.py: DRAM bytes (read + write) per launch of each
shipped kernel, last captured launch of each kind.  python tools/make_traffic.py gpurun_out/r2_full.ncu-rep profiles/r2_ncu_full_summary.txt"""
import csv
import json
import os
import subprocess
import sys

rep, summary = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
name_i, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
units = rows[1]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


traffic = {"_source": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, last captured launch of each kernel in the ncu --set full capture of tools/ncu_all.py "
                      f"(default shipped instantiations, round 2), summarised in {summary}"}
seen_grouped = 0
for r in rows[2:]:
    n = r[name_i]
    b = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
    if "sgemm_tc3x_kernel" in n:
        traffic["sgemm_tc3x_kernel@4096"] = int(b)
    elif "split_lo_kernel" in n:
        traffic["split_lo_kernel@4096"] = int(b)
    elif "sgemm_simt_kernel" in n:
        traffic["sgemm_simt_kernel@4096"] = int(b)
    elif "gemv_stream_kernel" in n and "GemvF32" in n:
        traffic["gemv_stream_kernel<GemvF32>@4096x16384"] = int(b)
    elif "gemv_stream_kernel" in n and "GemvS8" in n:
        targs = [a.strip() for a in n.split("<", 1)[1].split(">", 1)[0].split(",")]  # <T, WARPS, UNROLL, LPR, MROWS, GROUPED, MINB, EAGER>
        grouped = len(targs) > 5 and targs[5] in ("1", "true", "(bool)1")
        traffic["gemv_stream_kernel<GemvS8,GROUPED>@4096x14336" if grouped else "gemv_stream_kernel<GemvS8>@4096x14336"] = int(b)
json.dump(traffic, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
