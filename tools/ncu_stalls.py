"""Warp-stall breakdown (pc sampling) per kernel of an ncu --set full report: python tools/ncu_stalls.py rep.ncu-rep [...]"""
import csv, io, subprocess, sys
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        vals = []
        for i in cols:
            try:
                vals.append((float(r[i].replace(",", "")), hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
        tot = sum(v for v, _ in vals) or 1.0
        vals.sort(reverse=True)
        print(f"== {rep.split('/')[-1]}: {r[ki][:70]}")
        print("   " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in vals[:8]))
