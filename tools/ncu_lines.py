"""Per-source-line and per-function breakdown of an ncu report captured with --import-source on (needs -lineinfo):

    ncu -i gpurun_out/X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/x.csv
    python tools/ncu_lines.py /tmp/x.csv [top_lines]

Prints the share of warp-stall samples and executed instructions of each function (a line belongs to the last
function definition above it in its file) and of the hottest lines, with the dominant stall reasons."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(path)))
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "":
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)  # "Source" appears twice: keep the CUDA one
    d["file"] = cur
    lines.append(d)


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


S, I = "Warp Stall Sampling (All Samples)", "Instructions Executed"
tot_s, tot_i = sum(num(d[S]) for d in lines), sum(num(d[I]) for d in lines)
thr = sum(num(d["Thread Instructions Executed"]) for d in lines)
print(f"{len(lines)} source lines, {tot_s:.0f} stall samples, {tot_i:.3e} warp instructions, {thr / max(tot_i, 1):.1f} active lanes / instruction")

# function of each line
func_re = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|static|inline|__host__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
starts = defaultdict(list)
for f in {d["file"] for d in lines}:
    try:
        src = open(f).read().split("\n")
    except OSError:
        continue
    for i, line in enumerate(src, 1):
        m = func_re.match(line)
        if m and not line.strip().endswith(";"):
            starts[f].append((i, m.group(1)))


def func_of(f, ln):
    name = "?"
    for i, n in starts.get(f, []):
        if i <= ln:
            name = n
        else:
            break
    return name


by_func = defaultdict(lambda: [0.0, 0.0, 0.0])
for d in lines:
    k = f"{d['file'].split('/')[-1]}: {func_of(d['file'], int(d['Line No']))}"
    by_func[k][0] += num(d[S])
    by_func[k][1] += num(d[I])
    by_func[k][2] += num(d["Thread Instructions Executed"])
print("\n== by function")
for k, (s, i, t) in sorted(by_func.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{100 * s / tot_s:5.1f}% samples {100 * i / tot_i:5.1f}% instructions  {t / max(i, 1):5.1f} lanes  {k}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("\n== stall reasons (all samples)")
tot = {h: sum(num(d.get(h, 0)) for d in lines) for h in stalls}
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print(f"{100 * v / max(sum(tot.values()), 1):5.1f}%  {h}")
print("\n== hottest lines")
for d in sorted(lines, key=lambda d: -num(d[S]))[:top]:
    reasons = sorted(((num(d.get(h, 0)), h[6:]) for h in stalls), reverse=True)[:2]
    rs = ", ".join(f"{n}:{int(v)}" for v, n in reasons if v)
    print(f"{100 * num(d[S]) / tot_s:5.1f}% s {100 * num(d[I]) / tot_i:5.1f}% i {num(d['Thread Instructions Executed']) / max(num(d[I]), 1):5.1f} ln  "
          f"{d['file'].split('/')[-1]}:{d['Line No']:>4}  {d['Source'].strip()[:90]}   [{rs}]")
