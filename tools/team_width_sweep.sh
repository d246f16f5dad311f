#!/bin/bash
# Sweep of team widths (RS_TEAM_PN) and pass-0 segment plans (RS_SEG_P0) over a set of workloads; run on a GPU box
# from the repo root: bash tools/team_width_sweep.sh   (see DESIGN.md section 8.4 for the results that set the defaults)
W="cfg2 cfg5 heal:2048:1024 heal:1024:384 heal:2048:512"
echo "== default"; python tools/quick.py $W cfg1
echo "== PN=8"; RS_TEAM_PN=8 python tools/quick.py cfg5 heal:1024:384 heal:2048:512
echo "== PN=4"; RS_TEAM_PN=4 python tools/quick.py cfg5 heal:1024:384 heal:2048:512 heal:2048:1024
echo "== PN=2"; RS_TEAM_PN=2 python tools/quick.py heal:1024:384 heal:2048:512 heal:2048:1024 cfg2
echo "== P0 plan D 32768:8,131072:4,524288:2"; RS_SEG_P0="32768:8,131072:4,524288:2,0:1" python tools/quick.py $W
echo "== P0 plan E 65536:8,262144:4,1048576:2"; RS_SEG_P0="65536:8,262144:4,1048576:2,0:1" python tools/quick.py $W
echo "== P0 plan G 32768:8,262144:4,1048576:2"; RS_SEG_P0="32768:8,262144:4,1048576:2,0:1" python tools/quick.py $W
