#!/bin/bash
# development aid: CTA-0 durations of every UMMA kernel (needs a CGP_UMMA_TRACE build)
for k in kuf_fwd kdiag_fwd kdiag_bwd kuf_bwd_dz; do
  echo "== $k"; python tools/umma_trace.py $k 2>&1 | grep "CTA 0"
done
