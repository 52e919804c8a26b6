"""Aggregate an `ncu --page source --csv` SASS export by source line, using `nvdisasm --print-line-info` of the same
cubin for the address -> file:line map.  Usage: ncu_by_line.py source.csv sass_lineinfo.txt [topN]"""
import collections
import csv
import re
import sys

src_csv, li_txt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
addr2line = {}
cur = None
infunc = False
for line in open(li_txt):
    if ".text." in line and "rg_step_kernel" in line:
        infunc = True
    elif line.startswith("//--------------------- .text.") and "rg_step_kernel" not in line:
        infunc = False
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
    if m and infunc:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
base = None
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    n = float(r[ix["# Samples"]] or 0)
    agg[key]["samples"] += n
    agg[key]["instr"] += float(r[ix["Instructions Executed"]] or 0)
    agg[key]["thr"] += float(r[ix["Thread Instructions Executed"]] or 0)
    tot["samples"] += n
    tot["instr"] += float(r[ix["Instructions Executed"]] or 0)
    for s in stalls:
        v = float(r[ix[s]] or 0)
        agg[key][s] += v
        tot[s] += v
print("total samples %d, warp instructions %.3g" % (tot["samples"], tot["instr"]))
print("stall mix: " + ", ".join("%s %.1f%%" % (s[6:], 100 * tot[s] / max(tot["samples"], 1)) for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
byfile = collections.Counter()
for k, c in agg.items():
    byfile[k[0]] += c["samples"]
print("by file:", {k: "%.1f%%" % (100 * v / tot["samples"]) for k, v in byfile.most_common()})
print("%-22s %7s %7s %5s  top stalls" % ("file:line", "samp%", "instr%", "thr"))
for k, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(stalls, key=lambda s: -c[s])[:3]
    print("%-22s %6.2f%% %6.2f%% %5.1f  %s" % ("%s:%d" % k, 100 * c["samples"] / tot["samples"], 100 * c["instr"] / max(tot["instr"], 1),
                                              c["thr"] / max(c["instr"], 1), ", ".join("%s %.0f%%" % (s[6:], 100 * c[s] / max(c["samples"], 1)) for s in st)))

# ---- per-function view: function (or struct) = nearest preceding definition line in the same file that the PRODUCT build
# compiles -- definitions inside preprocessor branches that are off in that build (RG_EMU, RG_SKEW > 0, RG_PROFILE, RG_STATS)
# are skipped, so a sample can no longer be credited to code that is not in the binary
import os
MACROS = {"RG_SKEW": 0, "RG_SYNC_LEVEL": 2}          # defined in the product build; RG_EMU / RG_PROFILE / RG_STATS / RG_DEBUG_NEWTON are not


def cond_true(expr):
    e = re.sub(r"defined\s*\(?\s*(\w+)\s*\)?", lambda m: "True" if m.group(1) in MACROS else "False", expr)
    e = e.replace("&&", " and ").replace("||", " or ")
    e = re.sub(r"!(?!=)", " not ", e)
    e = re.sub(r"\b([A-Z_][A-Z0-9_]*)\b", lambda m: str(MACROS.get(m.group(1), 0)) if m.group(1) not in ("True", "False") else m.group(1), e)
    try:
        return bool(eval(e, {"__builtins__": {}}, {}))
    except Exception:
        return True


defs = {}
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robogym_b200", "csrc")
for f in os.listdir(root):
    L = []
    stack = []          # (this branch active, some branch of this #if already taken)
    for n, line in enumerate(open(os.path.join(root, f)), 1):
        t = line.strip()
        if t.startswith("#"):
            d = t[1:].strip()
            if d.startswith("ifdef"):
                a = d.split()[1] in MACROS; stack.append([a, a])
            elif d.startswith("ifndef"):
                a = d.split()[1] not in MACROS; stack.append([a, a])
            elif d.startswith("if"):
                a = cond_true(d[2:]); stack.append([a, a])
            elif d.startswith("elif") and stack:
                a = (not stack[-1][1]) and cond_true(d[4:]); stack[-1] = [a, stack[-1][1] or a]
            elif d.startswith("else") and stack:
                stack[-1] = [not stack[-1][1], True]
            elif d.startswith("endif") and stack:
                stack.pop()
            continue
        if not all(a for a, _ in stack):
            continue
        m = re.match(r"^(?:RG_DEV_NOINLINE|RG_DEV|RG_HD|static inline|__global__|__device__)\b.*?\b(rg_\w+)\s*\(", line)
        if m:
            L.append((n, m.group(1)))
            continue
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?struct\s+(\w+)\s*\{", line)
        if m:
            L.append((n, "struct " + m.group(1)))
    defs[f] = L
fagg = collections.defaultdict(lambda: collections.Counter())
for (f, ln), c in agg.items():
    name = f
    for n, fn in defs.get(f, []):
        if n <= ln:
            name = fn
    for k, v in c.items():
        fagg[name][k] += v
print("\n%-26s %7s %7s %5s  top stalls" % ("function", "samp%", "instr%", "thr"))
for k, c in sorted(fagg.items(), key=lambda kv: -kv[1]["samples"])[:40]:
    st = sorted(stalls, key=lambda s: -c[s])[:4]
    print("%-26s %6.2f%% %6.2f%% %5.1f  %s" % (k, 100 * c["samples"] / tot["samples"], 100 * c["instr"] / max(tot["instr"], 1), c["thr"] / max(c["instr"], 1),
                                              ", ".join("%s %.0f%%" % (s[6:], 100 * c[s] / max(c["samples"], 1)) for s in st)))
