"""Executed warp instructions and stall samples per CUDA source line of one kernel from an ncu report:
   python profiles/hot_lines.py gpurun_out/prof_TAG.ncu-rep k_voxel_finalize [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows, cur_file, hdr = [], None, None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            rows.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)))
        except (ValueError, KeyError):
            pass
# the page repeats itself (two views): de-duplicate on (file, line)
seen = {}
for f, ln, src, inst, smp in rows:
    seen.setdefault((f, ln), (src, inst, smp))
tot_i = sum(v[1] for v in seen.values()) or 1
tot_s = sum(v[2] for v in seen.values()) or 1
print(f"{kern}: {tot_i} warp instructions, {tot_s} samples over {len(seen)} source lines")
for (f, ln), (src, inst, smp) in sorted(seen.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{inst * 100 / tot_i:5.1f}% inst {smp * 100 / tot_s:5.1f}% smp  {f}:{ln:<4d} {src[:110]}")
