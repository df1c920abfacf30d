mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --decode-tokens 0 --no-cpu-baseline --rwkv-tokens 0"
$CMD > gpurun_out/prof_plain_v6.json 2> gpurun_out/prof_plain_v6.err || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 62 -c 1 -o gpurun_out/attn_v6 -f $CMD > gpurun_out/ncu_attn_v6.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/attn_v6.ncu-rep --page raw --csv > gpurun_out/attn_v6_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_v6.ncu-rep --page source --csv > gpurun_out/attn_v6_source.csv 2>/dev/null
ls -la gpurun_out/attn_v6*
