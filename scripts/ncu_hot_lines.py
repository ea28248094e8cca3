"""Join an `ncu --page source --csv` (SASS) dump with nvdisasm line info and print
the hottest source lines (dev helper).

  python scripts/ncu_hot_lines.py <sass.csv> <cubin> <kernel-substring> [N]
"""
import collections, csv, re, subprocess, sys
sass_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur_line, in_fn = [], None, False
for l in dis:
    if l.startswith(".text.") or l.strip().startswith(".section\t.text."):
        in_fn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if in_fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur_line)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Thread Instructions Executed" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0.0, 0.0])
k = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    ln = lines[k] if k < len(lines) else None
    k += 1
    agg[ln][0] += float(r[ci["Thread Instructions Executed"]] or 0)
    agg[ln][1] += float(r[ci["# Samples"]] or 0)
print(f"{k} SASS rows, {len(lines)} disassembled instructions")
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
src = {}
def text(ln):
    if not ln: return "?"
    f, n_ = ln
    if f not in src:
        try: src[f] = open(f"/root/repo/spaghettisearch_b200/csrc/{f}").read().splitlines()
        except Exception: src[f] = []
    return src[f][n_ - 1].strip()[:100] if n_ - 1 < len(src[f]) else ""
print("--- by thread instructions")
for ln, v in sorted(agg.items(), key=lambda x: -x[1][0])[:n]:
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/ts:5.1f}% samp  {ln}  {text(ln)}")
print("--- by stall samples")
for ln, v in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/ts:5.1f}% samp  {ln}  {text(ln)}")
