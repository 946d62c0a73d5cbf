#!/bin/bash
# Profile pass for one round, run on the GPU box through gpurun:
#   tools/profile_round.sh r1
# Writes into gpurun_out/: the ncu launch list of the default bench command and one
# `--set full` report per hot kernel (read back with tools/ncu_summarize.py).
set -u
R=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 3 --cpu-seconds 0 --c5 0 --ref-kernels 0"

# launch list of the bench command (per-launch times are cold-cache and serialised)
# (only the library's kernels: the first thousands of launches of the process are torch element-wise kernels of the
# synthetic graph generator, outside the timed region)
timeout 600 $NCU --metrics gpu__time_duration.sum -k 'regex:^(agg_|dense_|split_w|item_row|rowsum|edge_map|gat_)' -c 400 --csv --log-file $OUT/${R}_launches.csv $BENCH > $OUT/${R}_launches.log 2>&1

full() {  # name, kernel regex, skip, extra bench args
    timeout 900 $NCU --set full --import-source on -k "regex:$2" -s $3 -c 1 -f -o $OUT/${R}_prof_$1 $BENCH $4 > $OUT/${R}_prof_$1.log 2>&1
}
full agg        'agg_kernel'                3 ""
timeout 900 $NCU --set full --import-source on -k regex:agg_fixup -s 6 -c 2 -f -o $OUT/${R}_prof_fixup $BENCH > $OUT/${R}_prof_fixup.log 2>&1
full dense      'dense_tf32x3_ws'           3 ""
full agg_uniform 'agg_kernel'               3 "--sources uniform"
full agg_products 'agg_kernel'              3 "--workload products_gcn_layer_256 --locality-slices 1"
full dense_stream 'dense_tf32x3_ws'         3 "--workload products_gcn_layer_256 --locality-slices 1"
full agg_proteins 'agg_kernel'              3 "--workload proteins_gcn_layer_64"
full agg_arxiv  'agg_kernel'                3 "--workload arxiv_gcn_layer_32"
full agg_rmat26 'agg_kernel'                3 "--workload rmat26_gcn_agg_64"
timeout 900 $NCU --set full --import-source on -k regex:agg_kernel -s 2 -c 1 -f -o $OUT/${R}_prof_gat python tools/gat_loop.py 4 > $OUT/${R}_prof_gat.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw pages of every capture, the full report (with
# source) only for the two kernels worth reading line by line
for rep in $OUT/${R}_prof_*.ncu-rep; do
    ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2>/dev/null
    case $rep in *_prof_agg.ncu-rep|*_prof_dense_stream.ncu-rep) ;; *) rm -f $rep ;; esac
done
ls -la $OUT/${R}_prof_*
