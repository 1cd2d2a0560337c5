run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'])"; }
run loop; BPK_VB_NO_LOOP=1 run noloop; run loop
