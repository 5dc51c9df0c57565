"""SASS opcode summary per kernel of libmca_b200.so (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove
the Blackwell path (B200_PROFILING.md), the legacy tensor-core ones that must be absent, and the pipes the kernel leans on.
usage: python scripts/sass_summary.py [lib.so] > profiles/rN_sass_summary.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mca_paper_b200/csrc/libmca_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP",
         "SYNCS", "LDGMC", "PREEXIT", "ACQBULK", "HMMA", "IMMA", "HGMMA", "MUFU.EX2", "MUFU", "FFMA", "FFMA2", "FMUL", "F2FP", "LDS", "STS", "LDG", "STG", "RED", "ATOM",
         "BAR", "STL", "LDL", "USETMAXREG", "ELECT"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print(f"# {lib}: SASS opcode counts per kernel (sm_100a).  UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, "
      "UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce, SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce (NVSwitch; multimem.st is a plain STG.E.128.STRONG.SYS to the multicast address), "
      "PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait (programmatic dependent launch), STL/LDL = local-memory spills")
for k, c in counts.items():
    items = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{k[:70]:70s} instr={total[k]:6d}  {items}")
