"""SASS digest of the built library: per kernel, instruction count, registers are in the ptxas log; here the opcode histogram
(top opcodes) and the Blackwell-specific markers (UBLKCP / UTMALDG = TMA, SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FADD2 /
FMUL2 = packed fp32, DFMA = fp64).  Usage: python scripts/dev/sass_digest.py > profiles/sass_digest_r2.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
objs = sorted(glob.glob(os.path.join(ROOT, "hipgp_b200", "csrc", "build", "*.o")))
MARK = ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL", "LDS", "STS", "BAR", "WARPSYNC", "HMMA", "UTCHMMA")
only = sys.argv[1:] or None
print("# opcode histogram per kernel (cuobjdump -sass of hipgp_b200/csrc/build/*.o, sm_100a)")
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    fn = None; hist = None
    def flush():
        if fn and hist:
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"hipgp::", "", name)
            if only and not any(t in name for t in only):
                return
            tot = sum(hist.values())
            marks = " ".join("%s=%d" % (m, hist[m]) for m in MARK if hist.get(m))
            top = " ".join("%s:%d" % kv for kv in hist.most_common(8))
            print("%s\n    %d instructions | %s\n    top: %s" % (name[:150], tot, marks, top))
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            flush(); fn = m.group(1); hist = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and hist is not None:
            hist[m.group(1)] += 1
    flush()
