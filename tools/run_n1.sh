# N = 1: one ncu --set full capture of the shipped Viterbi kernel (one wave of utterances)
timeout 100 ncu --set full --clock-control none --import-source on -k regex:pitch_viterbi -s 1 -c 1 -f -o gpurun_out/r02_viterbi_final python tools/bench_configs.py --utts 4736 --only pitch_only > gpurun_out/r02l_ncu.log 2>&1
true
