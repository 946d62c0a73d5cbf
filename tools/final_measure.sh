set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
GNNAGG_HEAVY=1 timeout 900 python -m pytest tests/test_gpu_reference_kernels.py tests/test_gpu_backward.py -m gpu -q -k "timing or three_layer" 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_n1.err > gpurun_out/r1_bench_n1.json
python bench.py --impl reference 2>gpurun_out/bench_ref.err > gpurun_out/r1_bench_reference_arm.json
for w in arxiv_gcn_layer_32 proteins_gcn_layer_64 products_gcn_layer_256; do python bench.py --workload $w --cpu-seconds 3 2>>gpurun_out/bench_n1.err > gpurun_out/r1_bench_n1_$w.json; done
python bench.py --sources uniform --cpu-seconds 3 2>>gpurun_out/bench_n1.err > gpurun_out/r1_bench_n1_uniform_sources.json
python bench.py --scheduled 1 --workload products_gcn_layer_256 --cpu-seconds 1 2>>gpurun_out/bench_n1.err > gpurun_out/r1_bench_n1_products_sched.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls gpurun_out/*.jsonl
