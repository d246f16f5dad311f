tag=r03f; out=gpurun_out; mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
for rep in 1 2; do for v in "" _prev; do
  echo "== variant '$v' rep $rep"
  RS_LIB_VARIANT=$v timeout 300 python tools/quick.py cfg3 cfg4 cfg2 cfg1 cfg5 heal:1024:512 2>&1 | cut -c1-105
done; done 2>&1 | tee $out/${tag}_ab.txt
for t in 4 8 12; do echo "== RS_COPY_THREADS=$t"; RS_COPY_THREADS=$t timeout 200 python tools/phase_times.py cfg3 cfg5 2>&1 | grep -v "^PRNG" ; done | tee $out/${tag}_copy_threads.txt
timeout 300 python tools/batch_width_sweep.py --slots 2,3,4 --widths=-:- 2>&1 | tee $out/${tag}_batch.txt
