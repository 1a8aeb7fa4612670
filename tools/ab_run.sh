mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
python bench.py --workload c5 --no-cpu-baseline --no-evaluator --steps 50 > gpurun_out/r2o_bench_c5_n1.json 2> gpurun_out/r2o_bench_c5.err; cut -c1-300 gpurun_out/r2o_bench_c5_n1.json; tail -2 gpurun_out/r2o_bench_c5.err
python bench.py --no-cpu-baseline --no-evaluator > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench.err; cut -c1-300 gpurun_out/r2o_bench_n1.json
O=gpurun_out/r2o_step_ab.txt
: > $O
timeout 120 python tools/step_ab.py --reserve 8 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --tuning 0x40 --rounds 1 --only forward,overlapped >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --tuning 0x2 --rounds 1 --only forward,overlapped >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --tuning 0x42 --rounds 1 --only forward,overlapped >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 4 --decode-ctas 4 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --rounds 1 --only forward,overlapped >> $O 2>&1
cat $O
