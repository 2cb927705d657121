#!/bin/bash
mkdir -p gpurun_out/r02u
python -m pytest tests/test_gpu_parity_traits.py tests/test_trait_known_answers.py "tests/test_gpu_dict_adapters.py::test_cadence_dict_adapter_replays_reference_episode" tests/test_abi.py -m gpu -q > gpurun_out/r02u/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02u/pytest.log
python bench.py --variant cadence --envs 16384 --steps 100 --warmup 20 --no-cpu --no-e2e --no-configs > gpurun_out/r02u/bench_cadence.json 2> gpurun_out/r02u/bench_cadence.err; head -c 900 gpurun_out/r02u/bench_cadence.json; tail -3 gpurun_out/r02u/bench_cadence.err
