tag=r03j; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
for kv in A=1 RS_PAIR_FROM=131072 RS_PAIR_FROM=524288 RS_PAIR=0; do
  echo "== $kv"
  env $kv timeout 300 python tools/quick.py cfg4 cfg2 heal:1024:512 2>&1 | cut -c1-105
done 2>&1 | tee $out/${tag}_ab.txt
