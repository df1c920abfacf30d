mkdir -p gpurun_out
timeout 180 python -m pytest tests -m gpu -q -x -k "tiny and not simt" > gpurun_out/t_tiny.log 2>&1; echo "tiny rc=$?"; tail -n 15 gpurun_out/t_tiny.log
timeout 180 python -m pytest tests -m gpu -q -x -k "full_size" > gpurun_out/t_full.log 2>&1; echo "full rc=$?"; tail -n 5 gpurun_out/t_full.log
timeout 600 python bench.py --no-cpu-baseline --rwkv-tokens 0 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; cat gpurun_out/bench_v4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['decode'])"
