#!/usr/bin/env python
"""Round 2: turn the raw captures under gpurun_out/ (tools/ncu_r2.sh, tools/microbench.py) into the small, tracked
summaries under profiles/:
  r2_launches.csv / r2_launches_summary.csv   ncu launch list of the bench step (shares)
  r2_ncu_full_summary.json                    --set full metrics of every kernel, DRAM bytes next to algorithmic bytes
  ncu_traffic.json                            what bench.py's roofline.traffic reads
  r2_microbench_sweep.csv                     configs[4] sweep, CUDA-graph-timed"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = "r2"
out = ROOT / "profiles"
go = ROOT / "gpurun_out"
csv.field_size_limit(1 << 30)

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
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
own = lambda k: any(t in k for t in ("pipe_row_kernel", "norm3_kernel", "combine_kernel"))
tot_own = sum(v[1] for k, v in agg.items() if own(k)) or 1.0
lines = [f"# ncu launch list, {tag}. Command: ncu --metrics gpu__time_duration.sum --clock-control none --csv "
         "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor",
         "# Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's, not absolutes.",
         f"# {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total device time",
         "kernel,launches,total_us,avg_us,share,share_of_siss"]
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"\"{name[:110]}\",{c},{t / 1e3:.2f},{t / 1e3 / c:.2f},{t / tot:.4f},{(t / tot_own if own(name) else 0):.4f}")
(out / f"{tag}_launches_summary.csv").write_text("\n".join(lines) + "\n")
(out / f"{tag}_launches.csv").write_text((go / f"launches_{tag}.csv").read_text())

# 2. ncu --set full raw pages -> per-kernel metrics
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct"]
alg = json.loads((go / "r2_ncu_alg_bytes.json").read_text())
B, D, P, s_in = 64, 3 * 256 * 256, 113_673_220, 2
alg.update({"siss_add_noise_mixture": 4 * s_in * B * D, "siss_wmse_fwd_bwd": (12 + 3 * s_in) * B * D, "siss_norm3": 8 * P,
            "siss_combine": 12 * P})
names = [  # (substring of the demangled kernel name, C-ABI entry point) — first match wins
    ("MixtureOp<__nv_bfloat16, 1, 1>", "siss_add_noise_mixture_rng"), ("MixtureOp<__nv_bfloat16, 1, 0>", "siss_add_noise_mixture"),
    ("MixtureOp<__nv_bfloat16, 0", "siss_mixture_weights"), ("WmseFwdBwdOp", "siss_wmse_fwd_bwd"),
    ("AddNoiseOp<__nv_bfloat16, 2>", "siss_add_noise_pair"), ("AddNoiseOp<__nv_bfloat16, 1>", "siss_add_noise"),
    ("DualMseRngOp", "siss_dual_mse_rng_fwd_bwd"), ("DualMseOp", "siss_dual_mse_fwd_bwd"), ("dual_mse", "siss_dual_mse_fwd_bwd"),
    ("wmse_fwd_kernel", "siss_wmse_fwd(api-compat)"), ("batch_stats", "siss_batch_stats"), ("randn_kernel", "siss_randn"),
    ("draw_rows", "siss_draw_rows"), ("combine_adamw", "siss_combine_adamw"), ("mt_norm3", "siss_mt_norm3"),
    ("mt_combine", "siss_mt_combine"), ("membership_add_noise", "siss_membership_add_noise"),
    ("membership_sqerr", "siss_membership_sqerr"), ("MembershipSqerrOp", "siss_membership_sqerr"),
    ("norm3_kernel", "siss_norm3"), ("combine_kernel", "siss_combine"), ("counter", "siss_counter_add")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


summary, traffic = [], {}
for fname in ("prof_r2_final_raw.csv", "prof_r2_rest_raw.csv"):
    rr = list(csv.reader(open(go / fname)))
    h, units, data = rr[0], rr[1], rr[2:]
    idx = {k: i for i, k in enumerate(h)}
    seen = {}
    for d in data:
        name = d[idx["Kernel Name"]]
        key = next((v for k, v in names if k in name), name[:50])
        rec = {"kernel": key, "name": name[:140], "capture": fname}
        for w in want:
            col = next((c for c in h if c.endswith(w)), None)
            if col is not None:
                rec[w] = num(d[idx[col]])
                rec[w + ".unit"] = units[idx[col]]
        rd = to_bytes(rec["dram__bytes_read.sum"], rec["dram__bytes_read.sum.unit"])
        wr = to_bytes(rec["dram__bytes_write.sum"], rec["dram__bytes_write.sum.unit"])
        rec["dram_bytes_per_launch"] = rd + wr
        a = alg.get(key)
        if key in ("siss_mt_norm3", "siss_mt_combine"):
            a = (8 if key == "siss_mt_norm3" else 12) * P
        if a:
            rec["algorithmic_bytes_per_launch"] = a
            rec["dram_over_algorithmic"] = (rd + wr) / a
            dur = rec["gpu__time_duration.sum"] * ({"us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(rec["gpu__time_duration.sum.unit"], 1e-6))
            rec["algorithmic_GBps_under_ncu"] = a / dur / 1e9
        seen[key] = rec           # the LAST launch of each kernel (warm caches for code / constants; data is cold under ncu)
    for key, rec in seen.items():
        summary.append(rec)
        if fname == "prof_r2_final_raw.csv":
            traffic[key] = {"dram_bytes_per_launch": rec["dram_bytes_per_launch"],
                            "duration_us_under_ncu": rec["gpu__time_duration.sum"],
                            "source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full --clock-control none, 1 launch each)"}
(out / f"{tag}_ncu_full_summary.json").write_text(json.dumps(summary, indent=1) + "\n")
(out / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

# 3. microbench sweep -> table
sw = [json.loads(l) for l in open(go / f"sweep_{tag}.jsonl")]
tl = ["# tools/microbench.py sweep (BASELINE.json configs[4]), round 2: every launch is timed inside a CUDA-graph replay "
      "(the row is the kernel, not the host launch path); 'frac' is of the measured copy peak (MEASURED_PEAKS.json); "
      "siss_dual_mse_fwd_bwd now gets two DISTINCT prediction tensors (20 B/elem fp32, 18 bf16-in)",
      "kernel,config,B_or_P,dtype,us,alg_GB,GBps,frac,l2"]
for r in sw:
    tl.append(f"{r['kernel']},{r['config']},{r.get('B', r.get('P'))},{r.get('dtype', 'float32')},{r['us']:.1f},"
              f"{r['alg_bytes'] / 1e9:.4f},{r['gbs']:.0f},{r['frac_of_peak']:.3f},{r['l2']}")
(out / f"{tag}_microbench_sweep.csv").write_text("\n".join(tl) + "\n")
for rec in summary:
    print(f"{rec['kernel']:32s} {rec['gpu__time_duration.sum']:9.1f} {rec['gpu__time_duration.sum.unit']}  dram {rec['dram_bytes_per_launch'] / 1e6:9.1f} MB"
          f"  alg {rec.get('algorithmic_bytes_per_launch', 0) / 1e6:9.1f} MB  ratio {rec.get('dram_over_algorithmic', 0):5.2f}  "
          f"{rec.get('algorithmic_GBps_under_ncu', 0):7.0f} GB/s  active/elapsed {((rec.get('sm__cycles_active.avg') or 0) / (rec.get('sm__cycles_elapsed.avg') or 1)):.2f}")
