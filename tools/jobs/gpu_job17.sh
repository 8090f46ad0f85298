mkdir -p gpurun_out
python tests/bench_reference_kernel.py --frames 4 > gpurun_out/refk_4.json 2>gpurun_out/refk.err
python tests/bench_reference_kernel.py --frames 16 > gpurun_out/refk_16.json 2>>gpurun_out/refk.err
cat gpurun_out/refk_4.json gpurun_out/refk_16.json; tail -2 gpurun_out/refk.err
