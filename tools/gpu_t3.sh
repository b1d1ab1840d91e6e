#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 0 2>/tmp/err.txt | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), 'trace avg ms', round(d['roofline']['avg_launch_ms'],3), 'share', round(d['roofline']['kernel_share_of_step'],3))"; grep "L2 window" /tmp/err.txt | head -1; }
run MSNE_L2_PERSIST=1 MSNE_LOG_BOUNCES=1
run MSNE_L2_PERSIST=0
run MSNE_L2_PERSIST=1
run MSNE_L2_PERSIST=0
