python -m pytest tests/test_cuda_kernels.py -m gpu -x -q -k "peer or residual" > gpurun_out/pytest_peer.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_peer.log
for cfg in 0 6 2; do LDC_RING_CFG=$cfg python tools/dev_ring.py time 1024 1024 1 4; done > gpurun_out/ring_regs.jsonl 2>&1
python tools/dev_ring.py time 4096 4096 1 4 >> gpurun_out/ring_regs.jsonl 2>&1
python tools/dev_ring.py time 64 64 512 4 >> gpurun_out/ring_regs.jsonl 2>&1
for dbg in 0 81; do
    LDC_PEER_DEBUG=$dbg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu > gpurun_out/peer_exp_$dbg.out 2> gpurun_out/peer_exp_$dbg.err
    echo "dbg=$dbg rc=$?"
done
