tag=r03g; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
bash tools/ab_variants.sh $tag - _tab
