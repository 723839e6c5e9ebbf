"""Distils an ncu report (--set full) into profiles/r02_counters.json, the per-kernel constants bench.py's roofline records
read: FP64 warp-instructions per ellipsoid iteration, DRAM bytes per launch, pipe utilisations.
usage: python tools/ncu_counters.py <report.ncu-rep> <key> <kernel-regex> <units-per-launch> [<launch-index>]
  key               entry name in the JSON ("lowpass32", "sweep_n16384", ...)
  units-per-launch  ellipsoid iterations (batched kernels) or 1 (sweep) the profiled launch performed"""
import csv, io, json, os, re, subprocess, sys

rep, key, pat, units = sys.argv[1], sys.argv[2], re.compile(sys.argv[3]), float(sys.argv[4])
which = int(sys.argv[5]) if len(sys.argv) > 5 else -1
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, unit_row = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
cands = [r for r in rows[2:] if pat.search(r[col["Kernel Name"]])]
if not cands:
    sys.exit(f"no kernel matching {pat.pattern} in {rep}")
r = cands[which]

def val(name, scale_units=True):
    if name not in col:
        return None
    v = float(r[col[name]].replace(",", ""))
    u = unit_row[col[name]].lower()
    if scale_units:
        for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9),
                       ("second", 1.0)):
            if u.startswith(pre):
                return v * f
    return v

fp64 = val("smsp__inst_executed_pipe_fp64.sum", False) or val("sm__inst_executed_pipe_fp64.sum", False)
out = {
    "kernel": r[col["Kernel Name"]], "source": os.path.relpath(rep), "duration_s": val("gpu__time_duration.sum"),
    "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]], "registers": r[col["launch__registers_per_thread"]],
    "units_in_profiled_launch": units,
    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
    "dram_bytes_per_launch": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
    "fp64_warp_instr": fp64, "fp64_warp_instr_per_iteration": None if fp64 is None else fp64 / units,
    "warp_instr_per_iteration": (val("smsp__inst_executed.sum", False) or 0) / units,
    "shared_wavefronts_per_iteration": (val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False) or 0) / units,
    "pipes": {"fp64_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
              "shared_mem_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", False),
              "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
              "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False)},
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_counters.json")
try:
    allc = json.load(open(path))
except Exception:
    allc = {}
allc[key] = out
json.dump(allc, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
