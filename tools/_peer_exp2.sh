run() {
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $K --warmup 5 --no-cpu --no-strips --no-ensemble 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); h=d['halo']; print('K=$K $*', 'fused %.2f kernel %.2f nccl %.2f enq %.1f regions %s' % (h['fused_peer_ms_per_step']*1e3, h['kernel_only_ms']*1e3, h['nccl_sendrecv_then_kernel_ms_per_step']*1e3, h['host_enqueue_us_per_step'], [round(r*1e3,2) for r in d['regions_ms_per_step']]), h['bit_identical_to_nccl_path'], d['multi_gpu_parity'], 'e2e', d['e2e']['ms_per_call'])
"
}
python -m pytest tests/test_cuda_kernels.py tests/test_strips.py -m gpu -q -x -k "peer or strip" 2>&1 | tail -3
K=20 run LDC_PEER_DEBUG=0
K=200 run LDC_PEER_DEBUG=0
