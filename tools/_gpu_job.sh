# the round's validation recipe on a GPU box: /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/_gpu_job.sh'
set +e
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r1_gpu_tests.log
