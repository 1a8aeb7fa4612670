mkdir -p gpurun_out
NG=${NG:-8}
CALIB=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/gather_check.py > gpurun_out/r2q_gather_check_n$NG.txt 2>&1
CALIB=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 tools/gather_check.py >> gpurun_out/r2q_gather_check_n$NG.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$\|W1017\|warn" gpurun_out/r2q_gather_check_n$NG.txt | tail -40
