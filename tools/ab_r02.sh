#!/bin/bash
# round-2 measurement batch (run on the GPU box through gpurun); every line it prints is a record for profiles/r02_sweeps.md
mkdir -p gpurun_out
B="timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
pick='import json,sys; d=json.loads(sys.stdin.readline()); print(sys.argv[1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "interactive", round(d["e2e_interactive"]["value"],1), "ms/call", round(d["e2e_interactive"]["ms_per_call"],2), "ahead", round(d["e2e_interactive"].get("render_ahead",{}).get("value",0),1), round(d["e2e_interactive"].get("render_ahead",{}).get("ms_per_call",0),2), "trace share", round(d["roofline"]["kernel_share_of_step"],3), "reorder", round(d["roofline"]["reorder_share_of_step"],3))'
for v in "$@"; do
  case $v in
    ctx2) $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx1) MSNE_WAVE_CONTEXTS=1 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx2big) $B --paths-in-flight 268435456 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx2small) $B --paths-in-flight 67108864 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx1big) MSNE_WAVE_CONTEXTS=1 $B --paths-in-flight 268435456 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx4) MSNE_WAVE_CONTEXTS=4 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx4small) MSNE_WAVE_CONTEXTS=4 $B --paths-in-flight 67108864 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx2tail128k) MSNE_TAIL_THRESHOLD=131072 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ctx4tail128k) MSNE_WAVE_CONTEXTS=4 MSNE_TAIL_THRESHOLD=131072 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    nostage) MSNE_LIB=$PWD/moonshine_b200/libmoonshine_nostage.so $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    lib_*) MSNE_LIB=$PWD/moonshine_b200/libmoonshine_${v#lib_}.so $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    cfg5) python bench.py --workload cfg5 --steps 3 --warmup 3 --batch-spp 8 --no-cpu-baseline 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    cfg5_*) MSNE_LIB=$PWD/moonshine_b200/libmoonshine_${v#cfg5_}.so python bench.py --workload cfg5 --steps 3 --warmup 3 --batch-spp 8 --no-cpu-baseline 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    cfg2) python bench.py --workload cfg2 --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    leaf*) MSNE_MAX_LEAF=${v#leaf} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    tg*) MSNE_TRACE_GRID_PER_SM=${v#tg} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    sg*) MSNE_SHADE_GRID_PER_SM=${v#sg} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ahead*) $B --render-ahead ${v#ahead} 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    rsp*) MSNE_TRACE_SUSPEND=8 MSNE_TRACE_SUSPEND_PASSES=${v#rsp} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    rs*) MSNE_TRACE_SUSPEND=${v#rs} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    ttb*) MSNE_TAIL_BLOCK=32 $B --steps 3 --ahead-tail-paths ${v#ttb} 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    tt*) $B --steps 3 --ahead-tail-paths ${v#tt} 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    tb*) MSNE_TAIL_BLOCK=${v#tb} $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    nosort) MSNE_SORT_THRESHOLD=0 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    tail8k) MSNE_TAIL_THRESHOLD=8192 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    tail128k) MSNE_TAIL_THRESHOLD=131072 $B 2>gpurun_out/ab_$v.err | python -c "$pick" $v ;;
    *) echo "unknown variant $v" ;;
  esac
done
