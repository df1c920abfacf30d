#!/bin/bash
# ncu evidence for profiles/: launch list (per-launch device time) + full captures of the dominant kernels.
# Every ncu run is preceded by the identical plain command (B200_PROFILING.md).  Usage: scripts/gpu_profile.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --decode-tokens 0 --no-cpu-baseline --rwkv-tokens 0"
$CMD > gpurun_out/prof_plain_$TAG.json 2> gpurun_out/prof_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$TAG.err; exit 1; }
# launch list of the SECOND (timed) step: skip the warm-up step's launches
N=$(python -c "import json;print(json.load(open('gpurun_out/prof_plain_$TAG.json'))['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
# full captures. -k filters by function base name; --launch-skip counts MATCHING launches.
# per step: 240 trunk gemm_tc launches (4 per layer x 60 waves-layers) then 9 LM-head launches; 60 attention; 9 cdf
full() { # name kernel-regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  echo "full capture $1 rc=$? $(ls -la gpurun_out/$1_$TAG.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
G=$(python -c "import json;d=json.load(open('gpurun_out/prof_plain_$TAG.json'))['kernel_launches_per_step'];print(d['gemm'])")
H=$(python -c "import json;d=json.load(open('gpurun_out/prof_plain_$TAG.json'))['kernel_launches_per_step'];print(d['gemm_head'])")
A=$(python -c "import json;d=json.load(open('gpurun_out/prof_plain_$TAG.json'))['kernel_launches_per_step'];print(d['attn'])")
C=$(python -c "import json;d=json.load(open('gpurun_out/prof_plain_$TAG.json'))['kernel_launches_per_step'];print(d['cdf'])")
# the bench runs: warm-up step, timed step, then two e2e steps -> skip one whole step (G gemm launches) before capturing
full gemm_trunk "gemm_tc_kernel" $((G + 8)) 4          # layer 2 of the first wave: qkv(+rope), o, gate-up, down
full gemm_head "gemm_tc_kernel" $((2 * G - H)) 1       # the first LM-head launch of the timed step
full attn "attn_tc_kernel|attn_mma_kernel" $((A + 2)) 1
full cdf "cdf_cols_kernel" $C 1
full elem "embed_norm_kernel|rmsnorm_kernel|ac_encode_lanes_kernel" 2 3
for r in gemm_trunk gemm_head attn cdf elem; do
  [ -f gpurun_out/${r}_$TAG.ncu-rep ] && ncu -i gpurun_out/${r}_$TAG.ncu-rep --page raw --csv > gpurun_out/${r}_${TAG}_raw.csv 2>/dev/null
done
[ -f gpurun_out/attn_$TAG.ncu-rep ] && ncu -i gpurun_out/attn_$TAG.ncu-rep --page source --csv > gpurun_out/attn_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/
