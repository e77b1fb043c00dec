RND=${1:-r02b}
export W2E_BENCH_EXTRAS=0 W2E_BENCH_NEXT_ROWS=0
python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/plain_b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${RND}.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/ncu1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:modconv_tc2_kernel --launch-skip 68 --launch-count 17 -f -o gpurun_out/${RND}_step_modconv_tc2 python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/ncu2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_act_nhwc --launch-skip 32 --launch-count 8 -f -o gpurun_out/${RND}_step_blur_act_nhwc python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/ncu3b.log 2>&1
mkdir -p gpurun_out/profiles
W2E_PROFILES_OUT=gpurun_out/profiles python tools/make_profiles.py ${RND} > gpurun_out/mkprof_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/*.ncu-rep
tail -2 gpurun_out/mkprof_b.log
