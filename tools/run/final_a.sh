cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -s -k "256_box" > gpurun_out/final_tests_256.log 2>&1; grep -E "margins|passed|failed|Error|assert" gpurun_out/final_tests_256.log | tail -8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 200 gpurun_out/final_ref.json
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 300 gpurun_out/final_n1.err
