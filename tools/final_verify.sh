set -u
R=r1; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>$OUT/bench_n1.err > $OUT/r1_bench_n1.json
for w in arxiv_gcn_layer_32 proteins_gcn_layer_64 products_gcn_layer_256; do python bench.py --workload $w --cpu-seconds 3 2>>$OUT/bench_n1.err > $OUT/r1_bench_n1_$w.json; done
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --cpu-seconds 0"
timeout 300 $NCU --metrics gpu__time_duration.sum -k 'regex:^(agg_|dense_|split_w|item_row|rowsum|edge_map|gat_)' -c 400 --csv --log-file $OUT/${R}_launches.csv $BENCH > $OUT/${R}_launches.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:dense_tf32x3_ws -s 3 -c 1 -f -o $OUT/${R}_prof_dense $BENCH > $OUT/${R}_prof_dense.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:dense_tf32x3_ws -s 3 -c 1 -f -o $OUT/${R}_prof_dense_stream $BENCH --workload products_gcn_layer_256 > $OUT/${R}_prof_dense_stream.log 2>&1
for rep in $OUT/${R}_prof_dense.ncu-rep $OUT/${R}_prof_dense_stream.ncu-rep; do ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2>/dev/null; done
ls -la $OUT/${R}_prof_dense*
