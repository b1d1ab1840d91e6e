for rep in 1 2; do
echo "== default"; python bench.py --steps 3 --warmup 3 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],3))"
echo "== early uv"; MSNE_LIB=$PWD/build_variants/lib_earlyuv.so python bench.py --steps 3 --warmup 3 --batch-spp 8 --paths-in-flight 16777216 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['avg_launch_ms'],3))"
done
