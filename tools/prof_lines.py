"""Dev tool: per-source-line counts of executed SASS instruction classes (loads, local-memory traffic, branches, float64)
and of the long-scoreboard stall samples, per element, from `ncu --page source --csv` + `nvdisasm -gi`.
  python tools/prof_lines.py <ncu_source.csv> <nvdisasm.txt> <mangled-substring> <elements>"""
import csv, re, sys, collections
src, dis, sym, nel = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l)
chains, cur, pend = [], [], []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: pend.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        if pend: cur = pend; pend = []
        chains.append(cur)
assert len(chains) == len(data), (len(chains), len(data))
N = nel / 32
tot = sum(int(r[ci['Instructions Executed']] or 0) for r in data)
print("warp-instructions per warp of elements", round(tot / N, 1), " samples", sum(int(r[ci['# Samples']] or 0) for r in data))
for pat, name in ((r"^(LD|LDG)$", "global loads"), (r"^(LDC|LDCU|ULDC)$", "constant loads"), (r"^(LDL|STL)$", "local"), (r"^(BRA|BSSY|BSYNC|BREAK)$", "branch"), (r"^D[A-Z]+$", "fp64")):
    agg = collections.Counter(); t = 0
    for r, ch in zip(data, chains):
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ci['Source']].strip())
        if m and re.match(pat, m.group(1)):
            e = int(r[ci['Instructions Executed']] or 0); t += e; agg[ch[0] if ch else None] += e
    print("==", name, round(t / N, 1), "per element")
    print("   ", "  ".join(f"{k[0].replace('nrt_','')}:{k[1]} {v / N:.1f}" for k, v in agg.most_common(18) if k))
agg = collections.Counter()
for r, ch in zip(data, chains): agg[ch[0] if ch else None] += int(r[ci['stall_long_sb']] or 0)
t = sum(agg.values())
print("== long scoreboard samples", t)
print("   ", "  ".join(f"{k[0].replace('nrt_','')}:{k[1]} {100 * v / t:.1f}%" for k, v in agg.most_common(18) if k))
