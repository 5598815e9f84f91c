mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_z_bispec_pairs.py tests/test_gpu_z_mocks.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/late_tests.log 2>&1
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/late_smoke.log 2>&1
( timeout 90 python tools/bench_round_late.py 2>&1 | tail -5 ) > gpurun_out/late_bench.log 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/full_gpu_tests.log 2>&1
