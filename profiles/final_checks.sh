set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 600 gpurun_out/r2_final_bench.json; grep -E "e2e\]|Update took" gpurun_out/r2_final_bench.err | tail -6
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; cat gpurun_out/r2_final_ref.json | head -c 700
