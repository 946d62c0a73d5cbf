#!/bin/bash
# Final single-GPU pass of round 2 (second session): tests, smoke, bench + reference arm, backward timing and launch
# lists, ncu captures of the two passes of the GAT backward, sanitizer passes.  Run through gpurun; outputs in gpurun_out/.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test_gpu.log 2>&1; echo "rc=$?" >> $OUT/test_gpu.log; tail -2 $OUT/test_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
python bench.py 2>$OUT/bench_n1.err > $OUT/bench_n1.json; tail -c 400 $OUT/bench_n1.json
python bench.py --impl reference 2>$OUT/bench_ref.err > $OUT/bench_ref_n1.json
for w in arxiv_gcn_layer_32 proteins_gcn_layer_64; do python bench.py --workload $w --cpu-seconds 1 --c5 0 --ref-kernels 0 2>>$OUT/bench_n1.err > $OUT/r2b_bench_n1_$w.json; done
rm -f $OUT/backward.jsonl
GNNAGG_HEAVY=1 timeout 600 python -m pytest tests/test_gpu_backward.py -q -k timing 2>&1 | tail -1
NCU="ncu --clock-control none"
for c in "arxiv 32" "reddit 128"; do set -- $c
  timeout 300 $NCU --metrics gpu__time_duration.sum -k "regex:agg_|fixup|gat_bwd|rowsum" --csv --log-file $OUT/r2b_bwd_launches_$1$2.csv python tools/bwd_loop.py $1 $2 2 > $OUT/bwd_loop_$1.log 2>&1
done
timeout 300 $NCU --metrics gpu__time_duration.sum -k 'regex:^(agg_|dense_|split_w|item_row|rowsum|edge_map|gat_)' -c 400 --csv --log-file $OUT/r2b_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --c5 0 --ref-kernels 0 > $OUT/r2b_launches.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:agg_kernel -s 1 -c 1 -f -o $OUT/r2b_prof_gatbwd_pass1_reddit128 python tools/bwd_loop.py reddit 128 1 > $OUT/r2b_prof_p1.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:agg_kernel -s 1 -c 2 -f -o $OUT/r2b_prof_gatbwd_arxiv32 python tools/bwd_loop.py arxiv 32 1 > $OUT/r2b_prof_p2.log 2>&1
for rep in $OUT/r2b_prof_*.ncu-rep; do ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2>/dev/null; rm -f $rep; done
bash tools/sanitize.sh memcheck | tail -4
bash tools/sanitize.sh racecheck | tail -4
ls $OUT | tail -5
