#!/bin/bash
# Ablation of the activation quantizer (NB200_QUANT_DEBUG bits: 1 = no low-rank MMAs, 2 = no quantise phases)
for D in 0 1 2 3; do
  echo "#### quant debug=$D"
  NB200_QUANT_DEBUG=$D python tools/op_sweep.py --precision nvfp4 --shapes primary --bn 0 --out gpurun_out/tmp.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print(r['precision'], r['M'], r['K'], 'quant us', round(r['quant_us'],1))
"
done
