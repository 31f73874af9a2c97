"""Turn .ncu-rep captures (ncu --set full) into the small text/JSON summaries committed under profiles/.
usage: python tools/summarize_ncu.py out_prefix rep1.ncu-rep [rep2 ...]   (needs the `ncu` CLI, no GPU)"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tma.sum", "dram__cycles_active.avg"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    rows_out, traffic = [], []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"report": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")]}
            for k in KEYS:
                if k in hdr:
                    d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            tot = float(r[rd]) * UNIT_SCALE.get(units[rd], 1.0) + float(r[wr]) * UNIT_SCALE.get(units[wr], 1.0)
            d["dram_bytes_total"] = tot
            rows_out.append(d)
            traffic.append({"kernel_sig": d["kernel"], "dram_bytes": tot, "report": d["report"]})
    with open(prefix + ".txt", "w") as f:
        for d in rows_out:
            f.write(f"== {d['report']}: {d['kernel']}\n")
            for k, v in d.items():
                if k not in ("report", "kernel"):
                    f.write(f"   {k:70s} {v}\n")
    json.dump(traffic, open(prefix + ".json", "w"), indent=1)
    print(open(prefix + ".txt").read())


if __name__ == "__main__":
    main()
