"""Per-source-line stall samples from an ncu report.  usage: python tools/ncu_lines.py <rep> [top [launch-skip]]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30; skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file = cur = None
agg, inst, text = collections.Counter(), collections.Counter(), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0] != "":
        cur = (cur_file, r[0]); text[cur] = (r[1] if len(r) > 1 else '')[:100]; continue
    try:
        s, ie = int(r[4]), int(r[7])
    except Exception:
        continue
    agg[cur] += s; inst[cur] += ie
tot = sum(agg.values()) or 1
print("total samples", tot)
for k, v in agg.most_common(top):
    print(f"{v:6d} {v / tot:6.1%} inst={inst[k]:9d} {k[0]}:{k[1]}  {text.get(k, '')}")
