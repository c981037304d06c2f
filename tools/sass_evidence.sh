#!/bin/bash
# SASS evidence for the Blackwell-native paths, from the built library (no GPU needed).
LIB=${1:-siss_b200/libsiss_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a), instruction mnemonics of interest, counted over all kernels"
echo "# built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3"
cuobjdump -sass "$LIB" > /tmp/siss_sass.txt
echo "arch lines: $(grep -c 'arch = sm_100a' /tmp/siss_sass.txt) kernels for sm_100a, $(grep -c 'Function :' /tmp/siss_sass.txt) functions"
for m in "UBLKCP" "SYNCS.ARRIVE" "SYNCS.PHASECHK" "SYNCS.EXCH" "LDS.128" "LDG\.E(\.[A-Z]+)*\.128" "STG\.E(\.[A-Z]+)*\.128" "HMUL2.BF16_V2" "HADD2.BF16_V2\|HFMA2.BF16_V2" "DFMA" "F2FP.BF16" "MUFU.RCP" "LDGMC" "RED\.\|ATOMG" "MEMBAR" "BAR.SYNC" "HMMA\|UTC.MMA\|HGMMA\|LDTM"; do
  printf "%-34s %s\n" "$m" "$(grep -cE "$m" /tmp/siss_sass.txt)"
done
echo
echo "# per kernel family: TMA bulk copies / mbarrier ops / 128-bit global accesses"
python3 - <<'PY'
import re, collections
txt = open('/tmp/siss_sass.txt').read()
fam = collections.OrderedDict()
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function :|\Z)", txt, flags=re.S):
    name, body = m.group(1), m.group(2)
    key = ("pipe_row_kernel<" + re.search(r"(MixtureOp|AddNoiseOp|WmseFwdBwdOp|DualMseOp)", name).group(1) + ">") if "pipe_row_kernel" in name and re.search(r"(MixtureOp|AddNoiseOp|WmseFwdBwdOp|DualMseOp)", name) else \
          next((k for k in ("norm3_kernel", "combine_adamw_kernel", "combine_kernel", "p2p_reduce_norm3_kernel", "p2p_combine_allgather_kernel", "p2p_adamw_allgather_kernel",
                            "nvls_reduce_norm3_kernel", "nvls_xcombine_bcast_kernel", "scale_finalize_kernel", "ce_reduce_chunk_kernel",
                            "ce_combine_chunk_kernel", "publish_sums_kernel",
                            "membership_add_noise_kernel", "membership_sqerr_kernel", "randn_kernel", "draw_rows_kernel", "counter_add_kernel",
                            "mt_norm3_kernel", "mt_combine_kernel", "batch_stats_kernel", "mixture_kernel", "add_noise_kernel",
                            "wmse_fwd_bwd_kernel", "wmse_fwd_kernel", "wmse_bwd_kernel", "dual_mse_kernel", "sqerr") if k in name), "other")
    c = fam.setdefault(key, collections.Counter())
    c["kernels"] += 1
    for tag, pat in (("UBLKCP", r"UBLKCP"), ("SYNCS", r"SYNCS"), ("LDS.128", r"LDS\.128"), ("LDG.128", r"LDG\.E(\.[A-Z]+)*\.128"),
                     ("STG.128", r"STG\.E(\.[A-Z]+)*\.128"), ("BAR", r"BAR\.SYNC"), ("DFMA", r"DFMA"), ("MUFU", r"MUFU\.(LG2|SIN|COS)"), ("IMAD.WIDE", r"IMAD\.WIDE\.U32"),
                     ("LDGMC", r"LDGMC\.E\.ADD\.F32x4"), ("STG.SYS", r"STG\.E\.128\.STRONG\.SYS")):
        c[tag] += len(re.findall(pat, body))
print(f"{'family':44s} {'kernels':>7s} {'UBLKCP':>7s} {'SYNCS':>6s} {'LDS.128':>8s} {'LDG.128':>8s} {'STG.128':>8s} {'BAR':>5s} {'DFMA':>6s} {'MUFU.lg2/sin/cos':>17s} {'IMAD.WIDE.U32':>14s} {'LDGMC.ADD':>10s} {'STG.SYS':>8s}")
for k, c in fam.items():
    print(f"{k:44s} {c['kernels']:7d} {c['UBLKCP']:7d} {c['SYNCS']:6d} {c['LDS.128']:8d} {c['LDG.128']:8d} {c['STG.128']:8d} {c['BAR']:5d} {c['DFMA']:6d} {c['MUFU']:17d} {c['IMAD.WIDE']:14d} {c['LDGMC']:10d} {c['STG.SYS']:8d}")
PY
