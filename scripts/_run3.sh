mkdir -p gpurun_out
timeout 60 scripts/ubench/tmem_bw 2>&1 | tee gpurun_out/tmem_bw.txt
bash scripts/_run2.sh
