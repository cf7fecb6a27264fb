"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), duration and tensor-pipe activity of the
captured kernels from `ncu --set full` reports and write profiles/traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/r02_pair.ncu-rep [more.ncu-rep | raw-page .csv ...] [--note "how it was captured"]

A report too large to bring back is exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv`) and the CSV passed here.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(paths):
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for rep in paths:
        txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        txt = "".join(l for l in txt.splitlines(True) if not l.startswith("=="))
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        per_kernel = {}
        for r in data:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").strip()
            key = re.sub(r"<.*", "", name)
            key = {"pair_gemm_kernel": "tsp::pair_gemm_kernel", "dw_pair_kernel": "tsp::dw_pair_kernel", "split_gemm_kernel": "ts::split_gemm_kernel"}.get(key, key)
            ent = {"kernel": name, "grid": r[idx["Grid Size"]] if "Grid Size" in idx else None}
            for m in WANT:
                if m in idx:
                    ent[m] = r[idx[m]] + " " + units[idx[m]]
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            ent["dram_bytes"] = rd + wr
            per_kernel.setdefault(key, []).append(ent)
        for key, ents in per_kernel.items():
            res[key] = {"dram_bytes_per_launch": sum(e["dram_bytes"] for e in ents) / len(ents), "launches": ents,
                        "source": os.path.basename(rep) + " (ncu --set full --clock-control none" + (": " + NOTE if NOTE else "") + ")"}
            print(key, f"{res[key]['dram_bytes_per_launch'] / 1e6:.1f} MB per launch over {len(ents)} captured launches")
    json.dump(res, open(out_path, "w"), indent=1)


NOTE = ""
if __name__ == "__main__":
    args = sys.argv[1:]
    if "--note" in args:
        i = args.index("--note"); NOTE = args[i + 1]; del args[i:i + 2]
    main(args)
