#!/bin/bash
for T in auto p2p; do
SISS_GAP_TRANSPORT=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29559 tools/gap_probe.py 2>&1 | grep '^{'
done
