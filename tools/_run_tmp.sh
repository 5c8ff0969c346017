# scratch command file for `gpurun -- 'bash tools/_run_tmp.sh'`
python -m pytest tests/test_gpu_streams.py -x -q -k "random_configurations" 2>&1 | tail -15
