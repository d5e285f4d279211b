#!/usr/bin/env python
"""Regenerate profiles/hpsi_traffic.json -- the measured DRAM traffic per launch of the fused
H psi kernel that bench.py reports as roofline.traffic -- from kept ncu --set full reports:

    python tools/traffic_table.py gpurun_out/r02_hpsi256.ncu-rep gpurun_out/r02_hpsi128.ncu-rep

Each report REP.ncu-rep needs its sidecar REP.launches.jsonl written by tools/ncu_hpsi.py
(launch order = kernel order in the report).  An entry is keyed by box, dtype, operator and
the kernel signature (template arguments + tile configuration), so bench.py only quotes a
traffic figure for the very kernel configuration it launched."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def gb(row, key):
        v, u = float(row[idx[key]]), units[idx[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    res = []
    for d in data:
        if "k_hpsi" not in d[idx["Kernel Name"]]:
            continue
        dur, du = float(d[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]]
        res.append({"name": d[idx["Kernel Name"]],
                    "dram_read": gb(d, "dram__bytes_read.sum"),
                    "dram_write": gb(d, "dram__bytes_write.sum"),
                    "ms": dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[du]})
    return res


def main():
    caps = []
    for rep in sys.argv[1:]:
        side = rep.replace(".ncu-rep", ".launches.jsonl")
        launches = [json.loads(l) for l in open(side)]
        ks = kernels(rep)
        assert len(ks) == len(launches), (rep, len(ks), len(launches))
        for k, l in zip(ks, launches):
            S = 8 if l["dtype"] == "f64" else 4
            alg = 2.0 * S * l["grid"][0] * l["grid"][1] * l["grid"][2] * l["orbitals"]
            e = dict(l)
            e.update({"report": os.path.basename(rep), "ncu_kernel_name": k["name"],
                      "dram_bytes_per_launch": k["dram_read"] + k["dram_write"],
                      "dram_read": k["dram_read"], "dram_write": k["dram_write"],
                      "algorithmic_bytes": alg,
                      "traffic_over_algorithmic": (k["dram_read"] + k["dram_write"]) / alg,
                      "ncu_duration_ms": k["ms"]})
            caps.append(e)
    out = {"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of k_hpsi_tma, ncu "
                   "--set full --clock-control none; regenerate with tools/traffic_table.py",
           "captures": caps}
    with open(os.path.join(ROOT, "profiles", "hpsi_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in caps:
        print("%s %s lap%d %s: %.3f GB vs %.3f GB algorithmic (%.3fx)" % (
            c["grid"], c["dtype"], c["lap_type"], c["kernel"], c["dram_bytes_per_launch"] / 1e9,
            c["algorithmic_bytes"] / 1e9, c["traffic_over_algorithmic"]))


if __name__ == "__main__":
    main()
