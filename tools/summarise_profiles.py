#!/usr/bin/env python
"""Turn the raw captures under gpurun_out/ into the small, tracked summaries under profiles/.
Usage: python tools/summarise_profiles.py <round-tag>   (e.g. r1)"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)
go = ROOT / "gpurun_out"

# 1. launch list -> shares
rows = [r for r in csv.reader(open(go / f"launches_{tag}.csv")) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
lines = [f"# ncu launch list, {tag}. Command: ncu --metrics gpu__time_duration.sum --clock-control none --csv "
         "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor",
         "# Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's, not absolutes.",
         f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total device time "
         "(6 resident steps = 3 warm-up + 3 timed; torch fill/randn kernels are input setup)",
         "# share_of_siss = share among this repo's kernels only (the step itself; at:: kernels are input setup)",
         "kernel,launches,total_us,avg_us,share,share_of_siss"]
tot_siss = sum(v[1] for k, v in agg.items() if "siss::" in k) or 1.0
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    own = f"{t / tot_siss:.4f}" if "siss::" in name else ""
    lines.append(f"\"{name[:110]}\",{c},{t / 1e3:.2f},{t / 1e3 / c:.2f},{t / tot:.4f},{own}")
(out / f"{tag}_launches_summary.csv").write_text("\n".join(lines) + "\n")
(out / f"{tag}_launches.csv").write_text((go / f"launches_{tag}.csv").read_text())

# 2. ncu --set full -> per-kernel metrics and DRAM traffic
raw = subprocess.run(["ncu", "-i", str(go / f"prof_{tag}_final.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, data = rr[0], rr[1], rr[2:]
idx = {k: i for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return v * scale


traffic, summary = {}, []
short = {"MixtureOp": "siss_add_noise_mixture", "WmseFwdBwdOp": "siss_wmse_fwd_bwd", "norm3_kernel": "siss_norm3",
         "combine_kernel": "siss_combine"}
for d in data:
    name = d[idx["Kernel Name"]]
    key = next((v for k, v in short.items() if k in name), name[:40])
    rec = {"kernel": key, "name": name[:120]}
    for w in want:
        col = next((c for c in h if c.endswith(w)), None)
        if col is not None:
            rec[w] = num(d[idx[col]])
            rec[w + ".unit"] = units[idx[col]]
    rd = to_bytes(rec["dram__bytes_read.sum"], rec["dram__bytes_read.sum.unit"])
    wr = to_bytes(rec["dram__bytes_write.sum"], rec["dram__bytes_write.sum.unit"])
    rec["dram_bytes_per_launch"] = rd + wr
    traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                    "duration_us_under_ncu": rec["gpu__time_duration.sum"],
                    "source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full --clock-control none, 1 launch each)"}
    summary.append(rec)
(out / f"{tag}_ncu_full_summary.json").write_text(json.dumps(summary, indent=1) + "\n")
(out / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

# 3. microbench sweep -> table
sw = [json.loads(l) for l in open(go / f"sweep_{tag}.jsonl")]
tl = ["# tools/microbench.py sweep (BASELINE.json configs[4]). Small shapes are bounded by the Python launch path "
      "(~15-35 us per call), not by the kernels; 'frac' is of the measured copy peak (MEASURED_PEAKS.json).",
      "kernel,config,B_or_P,dtype,us,alg_GB,GBps,frac,l2"]
for r in sw:
    tl.append(f"{r['kernel']},{r['config']},{r.get('B', r.get('P'))},{r.get('dtype', 'float32')},{r['us']:.1f},"
              f"{r['alg_bytes'] / 1e9:.4f},{r['gbs']:.0f},{r['frac_of_peak']:.3f},{r['l2']}")
(out / f"{tag}_microbench_sweep.csv").write_text("\n".join(tl) + "\n")
print("wrote", sorted(p.name for p in out.iterdir()))
for k, v in traffic.items():
    print(k, f"{v['dram_bytes_per_launch'] / 1e6:.1f} MB", f"{v['duration_us_under_ncu']:.1f} us")
