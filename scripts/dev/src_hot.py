"""Summarise an ncu --page source CSV: stall samples and executed instructions per CUDA source line / per opcode."""
import csv, sys, collections, re, subprocess, io
rep, kern = sys.argv[1], sys.argv[2]
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass" if view == "cuda" else "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, ei, sa = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
stalls = collections.Counter()
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    try: n = int(r[ei]); sm = int(r[sa] or 0)
    except ValueError: continue
    s = r[si].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    op = s.split()[0].split(".")[0] if s else "?"
    ops[op] += n; samp[op] += sm; tot += n; tots += sm
    for c in stall_cols:
        try: stalls[h[c]] += int(r[c] or 0)
        except ValueError: pass
    lines.append((sm, n, r[si].strip()[:110], [ (h[c], int(r[c] or 0)) for c in stall_cols if (r[c] or "0") not in ("0", "")]))
print("total warp instr", tot, "samples", tots)
print("-- by opcode"); 
for op, n in ops.most_common(18): print(f"  {op:10s} {n:11d} {100*n/tot:5.1f}%  samples {samp[op]:6d} {100*samp[op]/max(tots,1):5.1f}%")
print("-- stall reasons"); 
for k, v in stalls.most_common(10): print(f"  {k:28s} {v:7d} {100*v/max(tots,1):5.1f}%")
print("-- top sampled instructions")
for sm, n, s, st in sorted(lines, key=lambda x: -x[0])[:40]:
    print(f"  {sm:6d} {n:9d}  {s:110s} {sorted(st, key=lambda x:-x[1])[:2]}")
