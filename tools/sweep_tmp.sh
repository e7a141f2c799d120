./tools/fp64_peak
for o in "cuda:chunk_mb=40" "cuda:chunk_mb=400" "cuda:chunk_mb=400 --opt cuda:seg=32" "cuda:chunk_mb=400 --opt cuda:strip=64" "cuda:chunk_mb=100"; do
  echo "== $o"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --opt $o 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:iwb_vstream -s 3 -c 1 -o gpurun_out/r1_vstream python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --opt cuda:chunk_mb=400 > gpurun_out/r1_vstream.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iwb_hstream -s 3 -c 1 -o gpurun_out/r1_hstream python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --opt cuda:chunk_mb=400 > gpurun_out/r1_hstream.log 2>&1
ls -la gpurun_out | tail -5
