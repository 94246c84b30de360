"""Freeze the executed fp64 flop count per Green's entry of the assembly kernels from an ncu capture of
scripts/assembly_probe.py (ncu --set full): flop = dadd + dmul + 2 dfma (thread-level, predicated-on).
Writes profiles/r02_assembly_flops.json, which bench.py multiplies by entries / CUDA-event time for
extra.assembly.<kernel>.roofline.   usage: python scripts/ncu_assembly_flops.py <rep> [out.json]"""
import csv
import json
import re
import subprocess
import sys

# entries one launch of scripts/assembly_probe.py produces, per kernel
ENTRIES = {"gf_fault_fault_kernel<0>": 256 * 64 * 64, "gf_fault_fault_kernel<1>": 256 * 64 * 64,
           "gf_fault_mantle_kernel<0>": 12288 * 1024, "gf_mantle_fault_kernel<0>": 1024 * 12288,
           "gf_mantle_mantle_kernel": 12288 * 12288, "gf_mantle_mantle_tile_kernel": 12288 * 12288,
           "gf_mantle_fault_tile_kernel<0>": 1024 * 12288}


def num(v):
    return float(v.replace(",", ""))


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        m = re.search(r"(gf_\w+_kernel)(?:<\s*(\d)[^>]*>)?", r[col["Kernel Name"]])
        if not m:
            continue
        key = m.group(1) + (f"<{m.group(2)}>" if m.group(2) is not None else "")
        if key not in ENTRIES or key in res:
            continue
        cyc = num(r[col["sm__cycles_elapsed.avg"]])
        ops = {}
        for op in ("dadd", "dmul", "dfma"):
            name = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
            ops[op] = num(r[col[name]]) if name in col else num(r[col[name + ".per_cycle_elapsed"]]) * cyc
        flop = ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]
        t = num(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[col["gpu__time_duration.sum"]]]
        res[key] = {"flop_per_entry": flop / ENTRIES[key], "entries_per_launch": ENTRIES[key],
                    "dadd": ops["dadd"], "dmul": ops["dmul"], "dfma": ops["dfma"], "ncu_time_ms": t * 1e3,
                    "source": f"{rep} (ncu --set full, scripts/assembly_probe.py)"}
    json.dump({"kernels": res, "flop_definition": "dadd + dmul + 2*dfma, smsp__sass_thread_inst_executed_op_*_pred_on"},
              open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_assembly_flops.json")
