run() {
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --no-strips --no-ensemble 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); h=d['halo']; print('$*', 'fused %.2f kernel %.2f nccl %.2f enq %.1f regions %s' % (h['fused_peer_ms_per_step']*1e3, h['kernel_only_ms']*1e3, h['nccl_sendrecv_then_kernel_ms_per_step']*1e3, h['host_enqueue_us_per_step'], [round(r*1e3,2) for r in d['regions_ms_per_step']]), h['bit_identical_to_nccl_path'], d['multi_gpu_parity'])
"
}
run LDC_PEER_DEBUG=512
run LDC_PEER_DEBUG=1024
run LDC_PEER_DEBUG=2048
run LDC_PEER_DEBUG=2560
run LDC_PEER_DEBUG=4096
run LDC_PEER_DEBUG=6144
