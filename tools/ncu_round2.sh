# ncu evidence of round 2 (one GPU): launch list of an eager iteration, --set full of every kernel of the loop, racecheck.
# Only CSV exports are kept: the .ncu-rep files exceed what gpurun copies back.
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_iter.csv python tools/prof_iter.py 2 > gpurun_out/r2_ncu_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/r2_prof_iter python tools/prof_iter.py 1 > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/r2_prof_iter.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_iter_raw.csv 2>/dev/null
ncu --profile-from-start off --set full --clock-control none -k regex:'linear_tc|ffn_tc|attn_seq|attn_last|embed_lut|env_step|env_roll|replay_gather' -c 14 -f -o /tmp/r2_prof_mem python tools/prof_iter.py 1 Memory-5-v0 > gpurun_out/r2_ncu_full_mem.log 2>&1
ncu -i /tmp/r2_prof_mem.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_mem_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/prof_iter.py 1 DiscreteCarFlag-v0 512 > gpurun_out/r2_racecheck.log 2>&1
tail -3 gpurun_out/r2_racecheck.log
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
