mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
O=gpurun_out/r2l_step_ab.txt
: > $O
for r in 0 8 16; do
  for sc in 1 2; do
    timeout 120 python tools/step_ab.py --reserve $r --rounds 1 --only forward,overlapped,decode --decode-schedule $sc >> $O 2>&1
  done
done
timeout 120 python tools/step_ab.py --reserve 8 --rounds 1 --only forward,overlapped,decode --decode-schedule 1 >> $O 2>&1
cat $O
