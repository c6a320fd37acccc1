"""Opcode evidence per translation unit (no GPU needed): `cuobjdump -sass build/*.o` -> counts of the Blackwell-native
mnemonics (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA,
UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, SYNCS = mbarrier) and of what must NOT be there (HMMA = legacy mma.sync,
BRA.U.ANY = the ELECT / R2UR waterfall loop around an instruction whose operands the compiler could not prove uniform).
usage: python scripts/sass_histogram.py > profiles/rNN_sass_opcodes.txt"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "USETMAXREG", "SYNCS",
        "MUFU.EX2", "FFMA2", "FADD2", "F2FP", "HMMA", "BRA.U.ANY", "LDL", "STL", "RED", "ATOM"]
print(__doc__.split("usage")[0])
for obj in sorted(glob.glob(os.path.join(ROOT, "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn)
            per[fn] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            per[fn]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "RED" and op.startswith("RED.")) or (k == "ATOM" and op.startswith("ATOM")):
                    per[fn][k] += 1
            if op.startswith("UTCHMMA.2CTA"):
                per[fn]["UTCHMMA.2CTA"] += 0  # counted by the prefix rule above
    print(f"== {os.path.basename(obj)}")
    for fn, c in per.items():
        hits = ", ".join(f"{k} {c[k]}" for k in KEYS if c[k])
        print(f"   {fn[:70]:70s} {c['total']:6d} instr | {hits}")
