"""Top stall-sample SASS lines of one kernel from an ncu report:
   python profiles/hot_sass.py gpurun_out/prof_TAG.ncu-rep k_radius_query [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
rows = []
for i, r in enumerate(rd):
    if not r.get("Source") or r["Address"] == "Address":
        continue
    try:
        rows.append((int(r["# Samples"] or 0), i, r))
    except ValueError:
        pass
tot = sum(s for s, _, _ in rows) or 1
inst = sum(int(r["Instructions Executed"] or 0) for _, _, r in rows)
print(f"{kern}: {tot} samples, {inst} warp instructions, {len(rows)} SASS lines")
stalls = [k for k in rd[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k] or 0) for _, _, r in rows) for k in stalls}
print("stall reasons:", ", ".join(f"{k[6:]} {v * 100 // tot}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, i, r in sorted(rows, key=lambda t: -t[0])[:top]:
    why = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{s * 100 / tot:5.1f}%  line {i:4d}  exec {r['Instructions Executed']:>8}  {r['Source'].strip():60s} {why}")
