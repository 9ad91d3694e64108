# replication threshold of the distributed multigrid (cells of the whole level at or below which a level is replicated)
for rep in 32768 131072 524288; do
  LDC_REP_CELLS=$rep python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/run_strips.py --grid 4096 2>/dev/null | tail -1 | sed "s/^/rep=$rep /"
done
