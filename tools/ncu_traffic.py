"""Map the kernel signatures of an ncu summary (tools/summarize_ncu.py) onto bench.py's launch names, so bench.py can
report roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) for its dominant kernel.
The capture is `tools/run_once.py 32` (one 32-page chunk = the shapes bench.py launches with device-resident pages)."""
import json
import re
import sys

PX = 32 * 1024 * 1024          # pixels of the captured chunk
rows = json.load(open(sys.argv[1]))
out, seen = [], {}
for r in rows:
    sig = r["kernel_sig"]
    name = None
    if "stem_planar_kernel" in sig:
        name = f"stem_planar[P={PX // 16}]"
    elif "head_planar_kernel" in sig:
        name = f"head_planar[P={PX // 16}]"
    elif "stem1_tc_kernel" in sig:
        name = f"stem1_tc[P={PX // 4},N=24]"
    elif "dwconv_tiled_h2_kernel" in sig:
        k = seen.get("dw7", 0); seen["dw7"] = k + 1
        name = f"dwconv7x7_h2[P={PX // 16 >> (2 * k)},C=96,s=1]" if k < 4 else None
    elif "mlp_tc_kernel" in sig:
        m = re.search(r"mlp_tc_kernel<(\d+), (\d+)", sig)
        c, n = int(m.group(1)), int(m.group(2))
        M = PX // 16 if (c, n) == (48, 48) else PX // 64
        name = f"mlp_tc[M={M},C={c},N={n},res={1 if c == n else 0}]"
    elif "mlp_big_kernel" in sig:
        m = re.search(r"mlp_big_kernel<(\d+), (\d+)", sig)
        c, n = int(m.group(1)), int(m.group(2))
        M = 122880 if "ncu_rec_top" in r["report"] else PX // 256
        name = f"mlp_tc[M={M},C={c},N={n},res=1]"
    if name and name not in [o["kernel"] for o in out]:
        out.append({"kernel": name, "dram_bytes": r["dram_bytes"], "source": f"profiles/r01_ncu_summary.txt ({r['report']}: {sig.split('(')[0]}, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"})
# carried over from the first capture of this round (kernel's data movement unchanged since: same tiles, same operands)
out.append({"kernel": "gemm_tc[M=122880,K=192,N=384,res=0]", "dram_bytes": 87916032,
            "source": "ncu --set full capture of the rec pw1 GEMM earlier in round 1 (git 66b01f3: profiles/r01_ncu_summary.txt, ncu_rec_gemm)"})
print(json.dumps(out, indent=1))
