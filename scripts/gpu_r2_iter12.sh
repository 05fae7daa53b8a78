# round 2, iteration 12: near-tie adjudication (literal phase 1 in the general kernel + batcher re-submission), full GPU suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "literal_phase1 or adjudicates" 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/r2_pytest_adj.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu12.log
