EMRIFD_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/prof_params.py > gpurun_out/r2_prof_params_2gpu.txt 2>&1
grep -v Warning gpurun_out/r2_prof_params_2gpu.txt | tail -30
