"""SASS evidence for profiles/: per kernel of libellcuda.so the counts of the mnemonics that prove the Blackwell paths
(UTMALDG/UTMASTG/UBLKCP = TMA, LDTM/STTM = tcgen05.ld/st, DADD/DMUL/DFMA = the FP64 pipe, ACQBULK/PREEXIT... = PDL) and the
first lines of the hot loops.  usage: python tools/sass_excerpt.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ellalgo_rs_b200", "csrc", "libellcuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "DADD", "DMUL", "DFMA", "DSETP", "MUFU", "LDS", "STS",
        "LDG", "STG", "SHFL", "BAR", "NANOSLEEP", "ACQBULK", "PREEXIT", "CCTL", "FENCE", "ATOM", "RED", "HMMA", "UTCHMMA", "UTCQMMA"]
print(f"cuobjdump -sass {os.path.relpath(so, ROOT)} — per-kernel mnemonic counts (static instruction counts, not executed)\n")
for (name, lines), dn in zip(funcs.items(), demangle):
    cnt = collections.Counter()
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1).split(".")[0]
            cnt[op] += 1
    short = re.sub(r"ellcuda::\(anonymous namespace\)::", "", dn)
    short = re.sub(r"\((ellcuda::)?BatchArgs.*", "(BatchArgs…)", short)
    print(f"{short[:150]}\n    {len(lines)} instructions; " + ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]))
want = sys.argv[1:] or ["sweep_kernelILi64ELi64ELi3", "lowpass32_tmem_kernelILi0ELb0ELi16"]
for w in want:
    for name, lines in funcs.items():
        if w in name:
            print(f"\n==== {name}: lines with TMA / TMEM / PDL / mbarrier instructions")
            for l in lines:
                if re.search(r"UTMA|UBLKCP|LDTM|STTM|SYNCS|ACQBULK|PREEXIT|NANOSLEEP|DEPBAR|ERRBAR|MEMBAR|FENCE", l):
                    print(l.rstrip()[:140])
