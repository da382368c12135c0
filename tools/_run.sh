out=gpurun_out; tag=r02i
BENCH="python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_fp32_leg --no_config5"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $out/${tag}_launches.csv $BENCH > $out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_win_kernel -s 40 -c 8 -o $out/${tag}_conv_win -f $BENCH > $out/${tag}_ncu_cw.log 2>&1; echo "ncu conv_win rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:wgrad_(win|thin)_kernel" -c 6 -o $out/${tag}_wgrad -f $BENCH > $out/${tag}_ncu_wg.log 2>&1; echo "ncu wgrad rc=$?"
ls -la $out | grep r02i
