mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_pipeline.py tests/test_gpu_next_rows.py -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
O=gpurun_out/r2m_step_ab.txt
: > $O
timeout 120 python tools/step_ab.py --reserve 8 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --decode-ctas 0 --rounds 1 --only overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --decode-ctas 4 --rounds 1 --only overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 0 --decode-ctas 8 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 4 --decode-ctas 4 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
timeout 120 python tools/step_ab.py --reserve 8 --rounds 1 --only forward,overlapped,decode >> $O 2>&1
cat $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"peaks_kernel|limbs_kernel|assemble_kernel" -c 30 --csv --log-file gpurun_out/r2m_decode_launches.csv python tools/step_ab.py --reserve 8 --rounds 1 --only decode --seconds 0.05 > /dev/null 2>&1
tail -12 gpurun_out/r2m_decode_launches.csv | cut -d, -f5,12-
