"""Dev tool: joins `ncu --page source --csv` (per-SASS-instruction executed counts and stall samples) with
`nvdisasm -gi` line info of the same kernel and aggregates per source line — both by the innermost line and
by the line of a chosen "frame" function body (call-site attribution through the inline chain).

  python tools/sass_by_line.py <ncu_source.csv> <nvdisasm_all.txt> <mangled-substring> [file:lo-hi of the frame body]
"""
import csv, re, sys, collections

src_csv, dis, sym = sys.argv[1:4]
frame = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
ex = [int(r[ci["Instructions Executed"]] or 0) for r in data]
smp = [int(r[ci["# Samples"]] or 0) for r in data]
sass = [r[ci["Source"]].strip() for r in data]

lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l)
chains, cur = [], []
pend = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        if pend:
            cur = pend
            pend = []
        chains.append(cur)
assert len(chains) == len(ex), (len(chains), len(ex))
tot = sum(ex)
tots = sum(smp)
inner = collections.Counter(); inner_s = collections.Counter()
fr = collections.Counter(); fr_s = collections.Counter()
ff, lo, hi = None, 0, 0
if frame:
    ff, rng = frame.split(":"); lo, hi = map(int, rng.split("-"))
for ch, e, s in zip(chains, ex, smp):
    if not ch:
        continue
    inner[ch[0]] += e; inner_s[ch[0]] += s
    if ff:
        k = next(((f, n) for f, n in ch if f == ff and lo <= n <= hi), ("?", 0))
        fr[k] += e; fr_s[k] += s
print(f"total warp-instructions {tot}  samples {tots}")
def show(c, cs, title, n=45):
    print(f"--- {title}")
    for k, v in c.most_common(n):
        print(f"{k[0]}:{k[1]:<5} {100.0 * v / tot:6.2f}% inst  {100.0 * cs[k] / max(tots,1):6.2f}% samples")
if ff:
    show(fr, fr_s, f"by line of the frame body {frame}")
show(inner, inner_s, "by innermost line")
# opcode histogram
op = collections.Counter()
for sline, e in zip(sass, ex):
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sline)
    if m:
        op[m.group(1).split(".")[0]] += e
print("--- opcodes")
print("  ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in op.most_common(40)))
