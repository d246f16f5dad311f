#!/bin/bash
# Round 2, call C: full GPU suite on the changed kernels, variant timings, new bench line, quality table, batch policy sweep.
# Every step under its own timeout; the first one is a 2-minute smoke that aborts the call if the engine hangs.
out=gpurun_out; tag=r02c
mkdir -p $out
timeout 180 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -20 $out/${tag}_smoke.txt; exit 1; }
tail -1 $out/${tag}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $out/${tag}_pytest_gpu.txt
tail -4 $out/${tag}_pytest_gpu.txt
timeout 240 python tools/quick.py cfg2 cfg4 cfg1 cfg5 cfg3 > $out/${tag}_quick_default.txt 2>&1
RS_NB_FULL=1 timeout 120 python tools/quick.py cfg2 cfg4 > $out/${tag}_quick_nbfull.txt 2>&1
RS_NO_CORPUS_BITS=1 timeout 200 python tools/quick.py cfg1 cfg5 cfg3 > $out/${tag}_quick_nobits.txt 2>&1
RS_LIB_VARIANT=_lut8 timeout 240 python tools/quick.py cfg2 cfg4 cfg1 cfg5 cfg3 > $out/${tag}_quick_lut8.txt 2>&1
for f in default nbfull nobits lut8; do echo "== $f"; cat $out/${tag}_quick_$f.txt; done
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 400 $out/${tag}_bench.err
timeout 300 python tools/quality_table.py --out $out/${tag}_quality > /dev/null 2> $out/${tag}_quality.err
tail -c 300 $out/${tag}_quality.err
timeout 300 python tools/batch_sweep.py > $out/${tag}_batch_sweep.txt 2>&1
cat $out/${tag}_batch_sweep.txt
