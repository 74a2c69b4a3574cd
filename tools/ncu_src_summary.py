"""dev tool: per-source-line warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_src_summary.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], newline='')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
src_text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr, r))
    try: line = int(r[0])
    except ValueError: continue
    key = (cur_file, line)
    src_text[key] = r[1]
    # second "Source" column is SASS; dict(zip) keeps the last -> fine
    n = int(d.get("# Samples", "0") or 0)
    agg[key]["samples"] += n
    agg[key]["inst"] += int(d.get("Instructions Executed", "0") or 0)
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try: agg[key][k] += int(v or 0)
            except ValueError: pass
tot = sum(a["samples"] for a in agg.values())
reasons = collections.Counter()
for a in agg.values():
    for k, v in a.items():
        if k.startswith("stall_"): reasons[k] += v
print("total samples", tot, " reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, tot)) for k, v in reasons.most_common(9)))
nonbar = sum(a["samples"] - a["stall_barrier"] for a in agg.values())
print("non-barrier samples", nonbar)
print("---- by all samples")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    rs = ", ".join("%s %d" % (k[6:], v) for k, v in a.most_common(6) if k.startswith("stall_") and v)
    print("%-14s %5d %7d %5.1f%% inst %8d | %s | %s" % (key[0], key[1], a["samples"], 100.0 * a["samples"] / tot, a["inst"], rs, src_text[key].strip()[:90]))
print("---- by non-barrier samples")
for key, a in sorted(agg.items(), key=lambda kv: -(kv[1]["samples"] - kv[1]["stall_barrier"]))[:top]:
    nb = a["samples"] - a["stall_barrier"]
    rs = ", ".join("%s %d" % (k[6:], v) for k, v in a.most_common(6) if k.startswith("stall_") and v and k != "stall_barrier")
    print("%-14s %5d %7d %5.1f%% inst %8d | %s | %s" % (key[0], key[1], nb, 100.0 * nb / max(1, nonbar), a["inst"], rs, src_text[key].strip()[:90]))
