#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs under ncu (NVLink byte counters + duration of the exchange kernels only; the
# symmetric-memory barrier kernels are NOT profiled — replaying a barrier would wait for a peer that does not replay)
METRICS=${SISS_NCU_METRICS:-nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes_data_protocol.sum,gpu__time_duration.sum}
if [ "$RANK" = "0" ]; then
  exec ncu --metrics $METRICS --clock-control none --cache-control none -k regex:"^(p2p_|nvls_|ce_|scale_finalize)" -c 60 --csv \
       --log-file gpurun_out/r2_ncu_nvlink_w${WORLD_SIZE}${SISS_NCU_TAG}.csv python "$@"
else
  exec python "$@"
fi
