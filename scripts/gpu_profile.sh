#!/bin/bash
# ncu evidence for profiles/: launch list (per-launch device time) + full captures of the dominant kernels.
# Every ncu run is preceded by the identical plain command (B200_PROFILING.md).  Usage: scripts/gpu_profile.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
REP=${CZ_NCU_REP_DIR:-/tmp/cz_ncu}
mkdir -p $REP
# ONLY="launches cdf_stats ..." restricts the run to the named steps (default: everything)
want() { [ -z "$ONLY" ] || [[ " $ONLY " == *" $1 "* ]]; }
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alice --no-gate --sharded-segments 0 --rwkv-bytes 0"
$CMD > gpurun_out/prof_plain_$TAG.json 2> gpurun_out/prof_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$TAG.err; exit 1; }
# launch list of the SECOND (timed) step: skip the warm-up step's launches
N=$(python -c "import json;print(json.load(open('gpurun_out/prof_plain_$TAG.json'))['gpu_launches'])")
echo "launches per step: $N"
if want launches; then
ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
fi
# full captures. -k filters by function base name; --launch-skip counts MATCHING launches.
full() { # name kernel-regex skip count [command]
  want $1 || return 0
  local cmd="${5:-$CMD}"
  # the .ncu-rep files stay on the box (gpurun merges at most 64 MiB back): only the raw / source CSV exports travel
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $REP/$1_$TAG -f $cmd > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "full capture $1 rc=$? $(ls -la $REP/$1_$TAG.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
J() { python -c "import json;d=json.load(open('gpurun_out/prof_plain_$TAG.json'))['kernel_launches_per_step'];print($1)"; }
G=$(J "d['gemm']"); H=$(J "d['gemm_head']"); A=$(J "d['attn']")
NB=$(J "d['gemm_head']")   # CDF batches per step = LM-head launches (each batch: offsets + stats, then two prefix launches of which one works)
# the bench runs: warm-up step, timed step, profiled step, e2e steps -> skip one whole step before capturing
full gemm_trunk "gemm_tc_kernel" $((G + 8)) 4          # layer 2 of the first wave: qkv(+rope), o, gate-up, down
full gemm_head "gemm_tc_kernel" $((2 * G - 1)) 1       # the last LM-head launch of the timed step
full attn "attn_tc_kernel" $((A + 2)) 1
full cdf_stats "cdf_stats_tma_kernel" $NB 1
full cdf_prefix "cdf_bounds_warp_kernel" $((2 * NB)) 1   # the cached variant comes first of each pair
full elem "embed_norm_kernel|rmsnorm_kernel|ac_encode_lanes_kernel" 2 3
# RWKV-7 0.1B kernels and the stepwise decoder, from their own commands
RCMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-alice --no-gate --sharded-segments 0 --rwkv-bytes 131072 --rwkv-segments 128"
full rwkv "wkv7_scan_kernel|rwkv_ln_mix_kernel" 14 3 "$RCMD"
DCMD="python scripts/decode_profile.py smollm 98304 384 corpus"
full decode_step "decode_step_kernel" 300 1 "$DCMD"
full decode_attn "attn_tc_kernel" 9000 1 "$DCMD"
for r in gemm_trunk gemm_head attn cdf_stats cdf_prefix elem rwkv decode_step decode_attn; do
  [ -f $REP/${r}_$TAG.ncu-rep ] && ncu -i $REP/${r}_$TAG.ncu-rep --page raw --csv > gpurun_out/${r}_${TAG}_raw.csv 2>/dev/null
done
for r in attn cdf_stats cdf_prefix; do
  [ -f $REP/${r}_$TAG.ncu-rep ] && ncu -i $REP/${r}_$TAG.ncu-rep --page source --csv > gpurun_out/${r}_${TAG}_source.csv 2>/dev/null
done
python scripts/ncu_summary.py gpurun_out/*_${TAG}_raw.csv > gpurun_out/ncu_table_$TAG.md 2>/dev/null
cat gpurun_out/ncu_table_$TAG.md
du -sh gpurun_out
