# NOTE: experiment log. Bits 2, 4, 8, 32 (fence / polling variants) existed only in the builds these runs used;
# the library keeps 1 (no flag waits), 16 (no push), 64 (no flow-control events), 256 (single-kernel fallback).
# breakdown of the N=2 fused step (LDC_PEER_DEBUG bits: 1 no flag waits, 16 no push, 32 single peer kernel, 64 no events)
python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "peer" 2>&1 | tail -30
for k in 200; do
for dbg in 4 8 12 32 40; do
    LDC_PEER_DEBUG=$dbg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps $k --warmup 10 --no-cpu --no-strips --no-ensemble 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); h=d['halo']; print('K=$k dbg=$dbg fused %.2f kernel %.2f nccl %.2f enq %.1f regions %s' % (h['fused_peer_ms_per_step']*1e3, h['kernel_only_ms']*1e3, h['nccl_sendrecv_then_kernel_ms_per_step']*1e3, h['host_enqueue_us_per_step'], [round(r*1e3,2) for r in d['regions_ms_per_step']]), d.get('multi_gpu_parity'), h['bit_identical_to_nccl_path'])
"
done; done
