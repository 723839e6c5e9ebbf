"""Per-source-line summary of an ncu report captured with --import-source on: shared-memory wavefronts, executed
instructions and stall samples per CUDA source line of the first kernel in the report.
usage: python tools/ncu_lines.py report.ncu-rep <iterations in that launch> [top N]"""
import csv, io, subprocess, sys

rep, iters = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
sect, hdr, out = 0, None, []
for r in rows:
    if r and r[0] == "Line No":
        sect += 1; hdr = r; continue
    if r and r[0] == "Function Name" and sect == 0:
        print(r[1][:150])
    if sect == 1 and hdr and len(r) >= len(hdr) - 2 and r[0].isdigit():
        out.append(r)
def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
iw, ii, isamp = hdr.index("L1 Wavefronts Shared"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tw, ti, ts = (sum(f(r[k]) for r in out) for k in (iw, ii, isamp))
print(f"shared wavefronts/iter {tw/iters:.1f}  instructions/iter {ti/iters:.1f}  samples {ts:.0f}")
for key, name in ((iw, "wavefronts"), (ii, "instructions"), (isamp, "samples")):
    print(f"--- by {name}")
    for r in sorted(out, key=lambda r: -f(r[key]))[:top]:
        print(f"{r[0]:>4} {f(r[iw])/iters:7.1f} wf/it {f(r[ii])/iters:7.1f} inst/it {100*f(r[isamp])/ts:5.1f}% samp  {r[1].strip()[:95]}")
