# NOTE: experiment log of the 2-GPU strip-step runs quoted in DESIGN.md section 7 (LDC_PEER_DEBUG bits 512 / 1024 /
# 2048 / 4096 and LDC_PEER_NO_HELPER existed only in the builds those runs used).
run() {
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $K --warmup 5 --no-cpu --no-strips --no-ensemble 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); h=d['halo']; print('K=$K $*', 'fused %.2f kernel %.2f nccl %.2f enq %.1f regions %s' % (h['fused_peer_ms_per_step']*1e3, h['kernel_only_ms']*1e3, h['nccl_sendrecv_then_kernel_ms_per_step']*1e3, h['host_enqueue_us_per_step'], [round(r*1e3,2) for r in d['regions_ms_per_step']]), h['bit_identical_to_nccl_path'], d['multi_gpu_parity'])
"
}
python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "peer" 2>&1 | tail -2
K=20 run X=1
K=200 run X=1
K=20 run LDC_PEER_NO_HELPER=1
