mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 2>&1 | tail -30) > gpurun_out/t7_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err
tail -5 gpurun_out/t7_all.log; tail -3 gpurun_out/bench7.err; head -c 1500 gpurun_out/bench7.json; echo; head -c 600 gpurun_out/bench7_ref.json
