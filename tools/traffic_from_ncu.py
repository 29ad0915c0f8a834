#!/usr/bin/env python
"""profiles/roofline_traffic.json from an `ncu --set full` capture of one step (bench.py's `roofline.traffic`).

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv      # here, after the GPU call
    python tools/traffic_from_ncu.py /tmp/raw.csv profiles/<extract>.csv

Writes (a) the selected-columns extract that is committed under profiles/ and (b) profiles/roofline_traffic.json:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the tcgen05 GEMM kernel, averaged over the
captured GEMM launches of one step, stamped with a digest of the GEMM kernel's sources.  bench.py reports the figure
only while that digest matches the sources it runs (a changed kernel silently invalidates an old capture).
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]


def gemm_source_digest() -> str:
    h = hashlib.sha256()
    for f in ("gemm_sm100.cuh", "gemm_host.cu", "ptx.cuh", "mathfn.cuh"):
        with open(os.path.join(ROOT, "prot2text-v2-esm3_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    raw, extract = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(extract, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in data:
            if len(r) == len(hdr):
                w.writerow([r[idx[c]] for c in cols])
    gemm = [r for r in data if len(r) == len(hdr) and "gemm_bf16_tcgen05_kernel" in r[idx["Kernel Name"]]]
    # one step = the LAST five-launch group fc1, fc2, dgrad, dW2, dW1 (other GEMM launches belong to other paths)
    step = gemm[-5:] if len(gemm) >= 5 else gemm
    total = sum(to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in step)
    out = {"kernel": "gemm_bf16_tcgen05_kernel", "dram_bytes_per_launch": total / max(1, len(step)),
           "launches_in_capture": len(step), "source": os.path.relpath(extract, ROOT) + " (ncu --set full, one step)",
           "gemm_source_sha16": gemm_source_digest()}
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
