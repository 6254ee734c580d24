"""Per-source-line summary of an ncu report: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv
usage: python profiles/ncu_lines.py X.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name") or len(r) < len(hdr) - 2: continue
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    def g(name):
        try: return float(r[ix[name]])
        except Exception: return 0.0
    out.append(dict(file=cur_file, line=r[0], src=r[1].strip(), samples=g("# Samples"), inst=g("Instructions Executed"),
                    shw=g("L1 Wavefronts Shared"), shx=g("L1 Wavefronts Shared Excessive"),
                    stalls={k[6:]: g(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
tot = sum(o["samples"] for o in out) or 1
tin = sum(o["inst"] for o in out) or 1
agg = {}
for o in out:
    for k, v in o["stalls"].items(): agg[k] = agg.get(k, 0) + v
print("total samples %d, warp-instructions %d" % (tot, tin))
print("stall mix: " + " ".join("%s=%.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v / tot > 0.01))
out.sort(key=lambda o: -o["samples"])
for o in out[:top]:
    st = " ".join("%s=%d" % (k, v) for k, v in sorted(o["stalls"].items(), key=lambda t: -t[1])[:3] if v > 0)
    print("%5.1f%% inst %4.1f%% shw %8d(+%d) %s:%s  %s | %s" % (100 * o["samples"] / tot, 100 * o["inst"] / tin, o["shw"], o["shx"],
                                                      o["file"], o["line"], o["src"][:70], st))
