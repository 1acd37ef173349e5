"""Turn the raw ncu outputs of profiles/gpu_round.sh into the small tracked summaries:

  python profiles/summarise_ncu.py TAG   (reads gpurun_out/launches_TAG.csv, gpurun_out/prof_TAG.ncu-rep)

  profiles/TAG_launch_shares.csv   per-kernel mean duration / launches / share of the bench command
  profiles/TAG_ncu_full.csv        per-kernel dram bytes, L2 / L1 hit rates, occupancy, registers
  profiles/TAG_traffic.json        {kernel: dram bytes per launch}  (bench.py's roofline.traffic)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")


def short(name):
    m = re.search(r"\b(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name.split("(")[0][:60]


def launch_shares():
    path = os.path.join(GP, f"launches_{tag}.csv")
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(k, [0.0, 0])
        a[0] += ns
        a[1] += 1
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    total = sum(v[0] for v in ours.values())
    other = sum(v[0] for k, v in agg.items() if not k.startswith("k_"))
    with open(os.path.join(OUT, f"{tag}_launch_shares.csv"), "w") as f:
        f.write(f"# ncu launch list of the small bench command of profiles/gpu_final*.sh ({tag})\n")
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised launches;\n")
        f.write("# compare SHARES with bench.py's CUDA-event 'kernels' table, not absolute times.\n")
        f.write(f"# apc kernels {total / 1e3:.1f} us over {sum(v[1] for v in ours.values())} launches; "
                f"other (torch copies / fills) {other / 1e3:.1f} us\n")
        f.write("kernel,us_per_launch,launches,share_of_apc_time\n")
        for k, (ns, n) in sorted(ours.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{k},{ns / n / 1e3:.2f},{n},{ns / total:.4f}\n")
    return ours


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def full_capture():
    rep = os.path.join(GP, f"prof_{tag}.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    head, units, body = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(head)}
    per = OrderedDict()
    for r in body:
        k = short(r[col["Kernel Name"]])
        d = {}
        for m in METRICS:
            if m not in col:
                continue
            v, u = float(r[col[m]].replace(",", "") or 0), units[col[m]]
            if u in UNIT_BYTES:
                v *= UNIT_BYTES[u]
            elif u in UNIT_US:
                v *= UNIT_US[u]
            d[m] = v
        per.setdefault(k, []).append(d)
    traffic = {}
    with open(os.path.join(OUT, f"{tag}_ncu_full.csv"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one eager C2 scan (profiles/run_pipeline.py), {tag}; mean over captured launches\n")
        f.write("kernel,launches,us,dram_read_MB,dram_write_MB,l2_bytes_MB,l2_hit_pct,l1_hit_pct,regs,grid,block,"
                "warps_active_pct,sm_throughput_pct,mem_throughput_pct,dram_throughput_pct,warp_insts\n")
        for k, ds in per.items():
            mean = lambda m: sum(d.get(m, 0.0) for d in ds) / len(ds)
            rd_b, wr_b = mean("dram__bytes_read.sum"), mean("dram__bytes_write.sum")
            traffic[k] = rd_b + wr_b
            f.write(f"{k},{len(ds)},{mean('gpu__time_duration.sum'):.2f},{rd_b / 1e6:.3f},{wr_b / 1e6:.3f},"
                    f"{mean('lts__t_bytes.sum') / 1e6:.2f},{mean('lts__t_sector_hit_rate.pct'):.1f},"
                    f"{mean('l1tex__t_sector_hit_rate.pct'):.1f},{int(mean('launch__registers_per_thread'))},"
                    f"{int(mean('launch__grid_size'))},{int(mean('launch__block_size'))},"
                    f"{mean('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f},"
                    f"{mean('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},"
                    f"{mean('gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed'):.1f},"
                    f"{mean('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},"
                    f"{int(mean('smsp__inst_executed.sum'))}\n")
    json.dump({"source": f"ncu --set full, {tag}, profiles/run_pipeline.py (one eager C2 scan, cold L2 per replayed pass)",
               "dram_bytes_per_launch": {k: round(v) for k, v in traffic.items()}},
              open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    return traffic


def steady_capture():
    """gpurun_out/steady_kernels_TAG.csv (profiles/steady_traffic.sh: the 8-lane bench under ncu kernel replay,
    --cache-control none, two DRAM counters = one pass per kernel) -> profiles/TAG_steady_traffic.json"""
    import collections
    path = os.path.join(GP, f"steady_kernels_{tag}.csv")
    rows = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in csv.DictReader(io.StringIO("".join(rows))):
        v = float(r["Metric Value"].replace(",", "") or 0) * UNIT_BYTES.get(r["Metric Unit"], 1.0)
        agg[short(r["Kernel Name"])][r["Metric Name"]].append(v)
    mean = lambda xs: sum(xs) / max(len(xs), 1)
    per = {k: {"launches": len(d["dram__bytes_read.sum"]), "dram_read": round(mean(d["dram__bytes_read.sum"])),
               "dram_write": round(mean(d["dram__bytes_write.sum"])),
               "l2_hit_pct": round(mean(d["lts__t_sector_hit_rate.pct"]), 1)} for k, d in agg.items() if k.startswith("k_")}
    json.dump({"source": f"ncu kernel replay of the 8-lane bench (64 resident frames), --cache-control none --clock-control none, {tag}: "
                         "launches serialised, L2 left as the other lanes' kernels left it (steady state); write-backs are "
                         "charged to the kernel that evicts them",
               "dram_bytes_per_launch": {k: v["dram_read"] + v["dram_write"] for k, v in per.items()},
               "per_kernel": per,
               "MB_per_scan": round(sum(v["dram_read"] + v["dram_write"] for v in per.values()) / 1e6, 2)},
              open(os.path.join(OUT, f"{tag}_steady_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    if os.path.exists(os.path.join(GP, f"steady_kernels_{tag}.csv")):
        steady_capture()
    if os.path.exists(os.path.join(GP, f"launches_{tag}.csv")):
        launch_shares()
    if os.path.exists(os.path.join(GP, f"prof_{tag}.ncu-rep")):
        full_capture()
