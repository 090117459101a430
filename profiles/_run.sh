mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout=100 -k "dss_records or decode" > gpurun_out/t2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t2.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k1_decode_fixed -c 1 -o gpurun_out/r02_k1_dss14 -f python profiles/prof_k1.py > gpurun_out/ncu_k1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k1.log
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.log
