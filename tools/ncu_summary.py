#!/usr/bin/env python
"""Compact, committable summary of an .ncu-rep (one `ncu --set full --import-source on` capture):

    python tools/ncu_summary.py <file.ncu-rep> [--out summary.json] [--instance -1]

Per captured kernel launch: duration, launch shape, DRAM traffic, pipe utilisation (fp64 / tensor-DMMA / issue slots),
warp-stall reasons aggregated over the kernel, the instructions with the most stall samples, and a SASS opcode histogram
(the proof of which hardware paths the kernel uses: DMMA.8x8x4, UBLKCP, SYNCS mbarriers, MATCH, ...).
Runs `ncu -i` itself, so it works wherever ncu is installed (no GPU needed)."""
import argparse
import csv
import io
import json
import subprocess
import sys

RAW_KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_of_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_of_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct_of_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__cycles_elapsed.avg": "sm_cycles",
}


def run(args):
    return subprocess.run(args, capture_output=True, text=True, check=True).stdout


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        return float(value.replace(",", "")) * scale.get(unit, 1)
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", args.rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    launches = []
    for row in raw[2:]:
        rec = {}
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                rec["kernel"] = row[i]
            elif h in RAW_KEYS:
                v = row[i]
                if RAW_KEYS[h] in ("dram_read", "dram_write", "l2_bytes"):
                    rec[RAW_KEYS[h] + "_bytes"] = to_bytes(v, units[i])
                else:
                    try:
                        x = float(v.replace(",", ""))
                        if RAW_KEYS[h] == "duration_us":   # ncu picks the unit per report
                            x *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
                                  "second": 1e6, "s": 1e6}.get(units[i], 1.0)
                        rec[RAW_KEYS[h]] = x
                    except ValueError:
                        rec[RAW_KEYS[h]] = v
        launches.append(rec)
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", args.rep, "--page", "source", "--csv"]))))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for k, start in enumerate(secs):
        end = secs[k + 1] if k + 1 < len(secs) else len(src)
        h = src[start + 1]
        idx = {name: i for i, name in enumerate(h)}
        stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        rows = [r for r in src[start + 2:end] if len(r) >= len(h)]
        total = sum(int(r[idx["# Samples"]] or 0) for r in rows)
        stalls = {c: sum(int(r[idx[c]] or 0) for r in rows) for c in stall_cols}
        ops = {}
        for r in rows:
            toks = r[idx["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(rows, key=lambda r: -int(r[idx["# Samples"]] or 0))[:12]
        rec = launches[k] if k < len(launches) else {}
        rec["stall_samples"] = total
        rec["stall_pct"] = {c[6:]: round(100.0 * v / max(total, 1), 1) for c, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        rec["hot_instructions"] = []
        for r in top:
            sr = {c: int(r[idx[c]] or 0) for c in stall_cols}
            m = max(sr, key=sr.get)
            rec["hot_instructions"].append({"samples": int(r[idx["# Samples"]] or 0), "sass": " ".join(r[idx["Source"]].split())[:90],
                                            "main_stall": m[6:]})
        rec["sass_opcode_histogram"] = dict(sorted(ops.items(), key=lambda kv: -kv[1])[:40])
    out = {"report": args.rep.split("/")[-1], "launches": launches}
    text = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
