"""profiles/kernel_traffic.json from an `ncu --page raw --csv` dump of the two neighbour kernels.

    ncu -i gpurun_out/<tag>_neighbour_kernels.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/kernel_traffic.py /tmp/raw.csv "profiles/<tag>_ncu_neighbour_kernels.txt" > profiles/kernel_traffic.json

The file is keyed by the sha256 of the kernel sources (gws.kernel_source_hash(); the .so is not byte-reproducible across
nvcc runs): bench.py only reports `roofline.traffic` and `roofline.binding_roof.frac` from it when the kernels it is
timing are the ones that were profiled.
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "time_us": "gpu__time_duration.sum",
    "issue_active_frac": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_frac": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1_data_pipe_frac": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "dram_frac": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_frac": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_frac": "lts__t_sector_hit_rate.pct",
    "lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers": "launch__registers_per_thread",
}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    sys.path.insert(0, ROOT)
    import gmu_water_simulation_b200 as gws

    out = {"kernel_source_sha256": gws.kernel_source_hash(), "source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1],
           "state": "dam break, 1,011,240 particles, step 201 (tools/profile_step.py), one launch each, cold caches under ncu",
           "kernels": {}}
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        k = {}
        for key, metric in WANT.items():
            v, unit = float(r[ix[metric]].replace(",", "")), units[ix[metric]]
            if key.endswith("_frac"):
                v /= 100.0
            if metric.startswith("dram__bytes"):
                v *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
            k[key] = v
        k["dram_bytes"] = int(k.pop("dram_bytes_read") + k.pop("dram_bytes_write"))
        out["kernels"][name] = k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
