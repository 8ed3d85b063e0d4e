#!/bin/bash
# Round 2, GPU call 23 (one B200): ncu evidence for the final build — the launch list of two FLIP 256^3 sub-steps and full captures
# of the kernels added in the last session (natural-ordering row kernel, ICCG tile kernel, the SOR pass with dependent launches).
set -u
OUT=gpurun_out/r2c23
mkdir -p $OUT
run() { local name=$1; shift; echo "== $name"; local t0=$SECONDS; timeout "$@" > $OUT/$name.log 2>&1; echo "rc=$? $((SECONDS-t0))s $(tail -1 $OUT/$name.log | cut -c1-200)"; }
run ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches.csv python scripts/ncu_step.py 256 2
python scripts/ncu_summary.py launches $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -14 $OUT/launches_summary.txt
run ncu_lex 200 ncu --set full --clock-control none --import-source on -k regex:k_relax_lex_rows -s 150 -c 1 -f -o $OUT/lex_rows python scripts/ncu_step.py 256 1 flip natural
python scripts/ncu_summary.py rep $OUT/lex_rows.ncu-rep > $OUT/lex_rows_summary.txt 2>&1
run ncu_ic 200 ncu --set full --clock-control none --import-source on -k regex:k_ic_tiles -s 140 -c 1 -f -o $OUT/ic_tiles python scripts/ncu_step.py 128 1 flip iccg
python scripts/ncu_summary.py rep $OUT/ic_tiles.ncu-rep > $OUT/ic_tiles_summary.txt 2>&1
run ncu_relax 200 ncu --set full --clock-control none --import-source on -k regex:k_relax_colour_code -s 40 -c 1 -f -o $OUT/relax python scripts/ncu_step.py 256 1
python scripts/ncu_summary.py rep $OUT/relax.ncu-rep > $OUT/relax_summary.txt 2>&1
cat $OUT/lex_rows_summary.txt $OUT/ic_tiles_summary.txt $OUT/relax_summary.txt | head -70
