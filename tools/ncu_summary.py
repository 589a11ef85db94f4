"""Condense an `ncu --page raw --csv` export into the per-kernel key metrics kept under profiles/.

    ncu --set full --clock-control none -k regex:"attn_tc|tail_tc" -s 14 -c 7 -o gpurun_out/r2_forward_full \
        python tools/prof_forward.py
    ncu -i gpurun_out/r2_forward_full.ncu-rep --page raw --csv > gpurun_out/r2_forward_full_raw.csv
    python tools/ncu_summary.py gpurun_out/r2_forward_full_raw.csv profiles/r2_forward_ncu

writes <out>_key_metrics.csv (one row per captured launch) and <out>_summary.json (the same, keyed by kernel, plus
the per-launch DRAM traffic bench.py's roofline.traffic reads for the attention kernel).
"""
import csv
import json
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "time_us", "time"),
    ("dram__bytes_read.sum", "dram_read_MB", None),
    ("dram__bytes_write.sum", "dram_write_MB", None),
    ("lts__t_sectors.sum", "l2_sectors", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct", 1.0),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct", 1.0),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe_pct", 1.0),
    ("smsp__inst_executed.sum", "warp_instructions", 1.0),
    ("sm__cycles_elapsed.max", "sm_cycles", 1.0),
    ("launch__registers_per_thread", "registers", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("launch__block_size", "block", 1.0),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_bytes", None),
    ("smsp__cycles_active.avg", "smsp_cycles_active", 1.0),
]

UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}     # -> MB
TIME_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}   # -> us


def short(name):
    m = re.search(r"(\w+(?:<[^()]*>)?)\s*\(", name)          # last identifier (+ template arguments) before the parameter list
    return m.group(1) if m else name


def main(src, out):
    rows = list(csv.reader(open(src)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {n: i for i, n in enumerate(hdr)}
    recs = []
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        rec = {"kernel": short(r[col["Kernel Name"]]), "id": int(r[col["ID"]])}
        for name, key, scale in KEYS:
            if name not in col or r[col[name]] in ("", "n/a"):
                continue
            v = float(r[col[name]].replace(",", ""))
            if scale == "time":
                v *= TIME_SCALE[units[col[name]]]
            elif scale is None:
                u = units[col[name]].split("/")[0]
                v *= UNIT_SCALE[u]
                if key == "dyn_smem_bytes":
                    v = v * 1e6              # back to bytes
            else:
                v *= scale
            rec[key] = round(v, 4)
        recs.append(rec)
    keys = ["kernel", "id"] + [k for _, k, _ in KEYS]
    with open(out + "_key_metrics.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in recs:
            w.writerow(rec)
    summary = {"source": src, "note": "ncu --set full --clock-control none, one launch per row; times are under the profiler "
               "(serialised, cold L2) -- shares, traffic and pipe utilisation are the evidence, not the absolute time",
               "launches": recs}
    att = [r for r in recs if r["kernel"].startswith("attn_tc_kernel")]
    if att:
        summary["final"] = {
            "kernel": "attn_tc_kernel",
            "launches_averaged": len(att),
            "traffic_bytes_per_launch": round(sum((r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) for r in att) / len(att) * 1e6),
            "tensor_pipe_pct": round(sum(r.get("tensor_pipe_pct", 0) for r in att) / len(att), 2),
        }
    json.dump(summary, open(out + "_summary.json", "w"), indent=1)
    for rec in recs:
        print(rec)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
