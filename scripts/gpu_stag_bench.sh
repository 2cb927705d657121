mkdir -p gpurun_out
for cap in "64 192" "96 224" "128 448"; do
  timeout 300 python bench.py --variant stag --envs 8192 --cap $cap --no-cpu --no-e2e --steps 300 --warmup 1200 > gpurun_out/bench_stag_$(echo $cap | tr ' ' '_').json 2> gpurun_out/bench_stag.err
  cat gpurun_out/bench_stag_$(echo $cap | tr ' ' '_').json; tail -3 gpurun_out/bench_stag.err
done
